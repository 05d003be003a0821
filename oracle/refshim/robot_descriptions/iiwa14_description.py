"""Only so that the reference's test modules import; the iiwa14 collision URDF is not available
offline and collision geometry is outside the shim (tests that open these paths fail/skip)."""
PACKAGE_PATH = "/nonexistent/iiwa14_description"
REPOSITORY_PATH = "/nonexistent"
