"""Stand-in for the module NAME ``robot_descriptions`` (see ../README.md).  TEST INFRASTRUCTURE."""
