"""pytest plugin, TEST HARNESS ONLY (used by scripts/run_reference_tests.py --product): lets the
reference's own, unmodified test files (/root/reference/tests) run against THIS package.

* the module name ``pink`` (and ``pink.tasks``, ``pink.limits``, ...) is bound to ``pink_b200``;
* ``pinocchio`` / ``qpsolvers`` / ``robot_descriptions`` are the stand-ins of oracle/refshim (the
  tests call ``pin.integrate``, ``pin.centerOfMass``, ... to form their own expectations: the oracle
  is the checker there, as everywhere under tests/);
* without a GPU the engine is the host build of the kernels (tests/host_engine.py); with
  ``PK_REFALIAS_DEVICE=cuda`` the real library runs.
"""
import importlib
import os
import pkgutil
import sys

import numpy as np
import pytest

import pink_b200

sys.modules["pink"] = pink_b200
for _info in pkgutil.walk_packages(pink_b200.__path__, "pink_b200."):
    if _info.name.startswith("pink_b200.csrc") or _info.name.endswith("libpink_b200"):
        continue
    _mod = importlib.import_module(_info.name)
    sys.modules["pink" + _info.name[len("pink_b200"):]] = _mod

from tests.host_engine import HostEngine  # noqa: E402

# the stand-in pin.centerOfMass(model, data) reads the configuration the data was last updated
# with (Pinocchio keeps it inside Data; the product's Data has no such field): remember it
_set_q = pink_b200.Configuration._set_q


def _set_q_and_remember(self, q):
    _set_q(self, q)
    if self.data is not None and not self.batched:
        self.data._q = np.array(self.q, dtype=np.float64)


pink_b200.Configuration._set_q = _set_q_and_remember


@pytest.fixture(autouse=True)
def _route_engine(monkeypatch):
    if os.environ.get("PK_REFALIAS_DEVICE", "host") == "cuda":
        yield
        return
    import pink_b200.configuration as cfgmod

    cache = {}

    def get_engine(model, device=None):
        key = (id(model), len(model.frames))
        if key not in cache:
            cache[key] = HostEngine(model)
        return cache[key]

    monkeypatch.setattr(cfgmod, "get_engine", get_engine)
    yield
