import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pink_b200
from tests import extras
from tests.hostsim import HostSim
sc = extras.g1_extras(64)
cfg = pink_b200.Configuration(sc.model, None, torch.as_tensor(sc.q32, device="cuda"), collision_model=sc.collision_model)
for rep in range(3):
    v, st = pink_b200.solve_ik(cfg, sc.tasks, sc.dt, damping=sc.damping, limits=sc.limits, barriers=sc.barriers,
                               constraints=sc.constraints, safety_break=False, return_status=True)
    torch.cuda.synchronize()
    v, st = v.cpu().numpy(), st.cpu().numpy()
    hs = HostSim(sc.model)
    prob, targets, _ = sc.problem()
    v_h, st_h = hs.solve_ik(prob, sc.q32, targets)
    bad = np.nonzero(np.abs(v - v_h).max(axis=1) > 1e-3)[0]
    print("rep", rep, "bad instances", bad, "status gpu", st[bad], "host", st_h[bad])
    for i in bad:
        j = np.nonzero(np.abs(v[i] - v_h[i]) > 1e-3)[0]
        print(i, j, v[i][j], v_h[i][j])
