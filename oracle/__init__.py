"""CPU oracle of the batched differential-IK hot path (fp64, numpy).

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
it.  ``pink_b200`` never imports it and has no CPU fallback.

It restates, in plain numpy, what the reference computes on the path
``pink.solve_ik`` (``/root/reference/pink/solve_ik.py:206-275``):

* ``lie.py``         SE(3)/SO(3) maps Pink calls through Pinocchio
                     (``pin.log``, ``pin.Jlog6``, ``SE3.actInv`` ...),
* ``kinematics.py``  forward kinematics, LOCAL frame Jacobians, centre of mass,
                     ``pin.difference`` / ``pin.integrate``
                     (``pink/configuration.py:131-164,203-293``),
* ``tasks.py``       FrameTask / PostureTask / ComTask / RelativeFrameTask
                     error + Jacobian and the generic ``(H, c)`` formula
                     (``pink/tasks/task.py:115-167``),
* ``limits.py``      ConfigurationLimit / VelocityLimit rows
                     (``pink/limits/*.py``),
* ``qp.py``          Goldfarb-Idnani dual active-set QP (what
                     ``solver="quadprog"`` runs),
* ``ik.py``          ``build_ik`` / ``solve_ik`` assembly in the reference's order.

``refshim/`` holds stand-ins for the module NAMES ``pinocchio`` / ``qpsolvers`` so that the
reference's unmodified Pink-layer Python can be executed in the build container to generate
``tests/golden/ref_pink_layer_*.npz`` (``scripts/make_reference_golden.py``); with those
fixtures the Pink layer of this oracle (task composition, weighting, stacking, limit /
barrier / equality rows) is pinned to the reference's code (1e-10).

PARITY UNPINNED at the third-party boundary: the arithmetic of this path lives
in Pinocchio (pin 3.8.0, ``uv.lock:495``) and quadprog (unpinned), neither of
which is vendored under ``/root/reference`` nor installable offline, and the
reference's golden matrices (``tests/test_configuration.py:30-346``,
``tests/test_relative_frame_task.py:146-264``) need URDFs fetched from the
network.  The oracle is therefore pinned only by the model-independent
invariants the reference's own tests assert (finite-difference Jacobians,
at-target identities, unit-cost ``H == J^T J``, ...; see
``tests/test_oracle_*.py``) plus independent numerical cross-checks
(``scipy.linalg.logm``, brute-force KKT enumeration, SLSQP).  The vectors under
``tests/golden/`` are frozen outputs of THIS oracle (``scripts/make_golden.py``), a
drift guard and a GPU-side fixture - not reference outputs.
"""
