#!/usr/bin/env python3
"""Benchmark of the batched differential-IK hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one ``solve_ik`` pass over one batch of B = 65536 UR5 instances per
GPU (FrameTask(tool0) + PostureTask + default limits, examples/arm_ur5.py).
Prints ONE JSON line (rank 0).  See DESIGN.md section "Measurement".
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "IK solves/sec (batch=65536 per GPU, UR5 6-DOF, FrameTask+PostureTask, default limits)"
UNIT = "IK steps/s"
BYTES_PER_STEP = 96  # q 24 B + frame target 48 B read, v 24 B written
L2_BYTES = 126 * 1024 * 1024


def committed_traffic():
    """dram bytes per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/traffic.json names the capture); None when absent."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            return json.load(f)["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


# ---------------------------------------------------------------------------
# CPU arm: the oracle's C port of the reference path, all host cores
# ---------------------------------------------------------------------------


class CpuArm:
    """fp64 CPU implementation of the same step (oracle/c/pink_oracle.c: dense H,
    Goldfarb-Idnani QP with Givens updates, pthreads).  Pink + Pinocchio + quadprog
    cannot be installed offline, so this port stands in for them."""

    def __init__(self, batch, seed=20260922):
        from oracle import cport
        from oracle import kinematics as okin
        from pink_b200 import workloads as wl
        from pink_b200.robots import load_robot_description

        robot = load_robot_description("ur5_description")
        table = robot.model.table()
        f = table.frame_names.index("tool0")
        rng = np.random.default_rng(seed)
        q = wl.sample_configurations(table, batch, rng)
        qt = wl.perturb_configurations(table, q, rng)
        R, p = okin.frame_placement(table, okin.forward_kinematics(table, qt), f)
        self.T = np.concatenate([R, p[:, :, None]], axis=2).astype(np.float32).astype(np.float64)[:, None]
        self.q = q.astype(np.float32).astype(np.float64)
        tasks = [
            {"type": "frame", "frame": f, "cost": np.ones(6), "gain": 1.0, "lm_damping": 1.0},
            {"type": "posture", "cost": 1e-3, "gain": 1.0, "lm_damping": 0.0,
             "target": wl.ur5_posture_reference(robot.model)},
        ]
        self.port = cport.CPort(table, tasks, wl.UR5_DT, wl.UR5_DAMPING)
        self.batch = batch
        self.table, self.model, self.wl = table, robot.model, wl

    def step(self, threads):
        t0 = time.perf_counter()
        v, st = self.port.solve(self.q, self.T, threads=threads)
        return time.perf_counter() - t0, v, st

    def python_port_rate(self, n=200):
        """The numpy per-instance loop (structure of the reference's Python path), 1 core."""
        from oracle import ik as oik

        f = self.table.frame_names.index("tool0")
        tasks = [
            {"type": "frame", "frame": f, "cost": np.ones(6), "gain": 1.0, "lm_damping": 1.0,
             "target": (self.T[:n, 0, :, :3], self.T[:n, 0, :, 3])},
            {"type": "posture", "cost": 1e-3, "gain": 1.0, "lm_damping": 0.0,
             "target": self.wl.ur5_posture_reference(self.model)},
        ]
        t0 = time.perf_counter()
        oik.solve_ik_batch(self.table, self.q[:n], tasks, self.wl.UR5_DT, self.wl.UR5_DAMPING)
        return n / (time.perf_counter() - t0)


def cpu_baseline_block(batch, passes=5):
    cores = os.cpu_count() or 1
    arm = CpuArm(batch)
    arm.step(cores)  # warm-up (page faults, thread start)
    wall = sum(arm.step(cores)[0] for _ in range(passes))
    one = arm.step(1)[0]
    return {
        "value": batch * passes / wall, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{passes} passes over the same {batch}-instance UR5 workload, fp64 C port of the reference path "
                  f"(dense H, Goldfarb-Idnani QP), {cores} pthreads; Pink/Pinocchio/quadprog are not installable offline",
        "one_core": batch / one,
        "python_loop_one_core": arm.python_port_rate(),
    }


def run_reference_arm(args):
    """``--impl reference``: the reference's CPU path on all host cores (the
    oracle's C port; the real stack cannot be installed offline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    arm = CpuArm(args.batch)
    for _ in range(max(args.warmup, 1)):
        arm.step(cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step(cores)
    wall = time.perf_counter() - t0
    value = args.batch * args.steps / wall
    sample = (f"every step = the full {args.batch}-instance workload, fp64 C port of the reference path "
              f"(oracle/c/pink_oracle.c), {cores} pthreads")
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    cfg = {
        "workload": "UR5 6-DOF (hand-authored URDF), FrameTask(tool0, 1, 1, lm_damping=1) + PostureTask(1e-3), "
                    "ConfigurationLimit + VelocityLimit, dt=1/200, damping=1e-12 (examples/arm_ur5.py)",
        "batch_per_gpu": args.batch,
        "targets": "reachable: FK(q + N(0, 0.3^2)) clipped to limits; seed 20260922",
    }
    if not cpu:
        cfg["l2"] = f"rotating {args.nbuf} input/output sets ({args.nbuf * args.batch * BYTES_PER_STEP / 2**20:.0f} MiB > 126 MiB L2)"
    return cfg


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx.append(float(s[1]))
                for name, val in zip(names, s[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from pink_b200 import FrameTask, PostureTask, _cabi, workloads
    from pink_b200.engine import get_engine
    from pink_b200.limits import ConfigurationLimit, VelocityLimit
    from pink_b200.robots import load_robot_description
    from pink_b200.solve_ik import describe_problem

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; pink_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    real_stdout = None
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner on fd 1 at
        # communicator creation (NCCL_DEBUG >= VERSION); everything written to fd 1 from here
        # on goes to stderr, the JSON line is written to the saved descriptor at the end
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=device)

    # model constants: built on rank 0, broadcast over NCCL (north_star), then
    # turned into device tables by every rank
    from pink_b200 import parallel

    robot = load_robot_description("ur5_description") if rank == 0 else None
    model = parallel.broadcast_model(robot.model if rank == 0 else None, device)
    eng = get_engine(model, device)
    table = eng.table
    B, NBUF = args.batch, args.nbuf
    f = table.frame_names.index("tool0")

    # seeded synthetic inputs, one set per buffer; rank r owns shard r of the job
    rng = np.random.default_rng(workloads.SEED + 1000 * rank)
    qs, ts, vs, ss = [], [], [], []
    for b in range(NBUF):
        q = workloads.sample_configurations(table, B, rng)
        qt = workloads.perturb_configurations(table, q, rng)
        q_d = torch.as_tensor(q, dtype=torch.float32, device=device)
        oMf, _ = eng.forward_kinematics(torch.as_tensor(qt, dtype=torch.float32, device=device))
        qs.append(q_d)
        ts.append(oMf[:, f].reshape(B, 12).contiguous())
        vs.append(torch.empty((B, 6), dtype=torch.float32, device=device))
        ss.append(torch.empty((B,), dtype=torch.int32, device=device))
    frame_task = FrameTask("tool0", position_cost=1.0, orientation_cost=1.0, lm_damping=1.0)
    frame_task.set_target(ts[0])
    posture_task = PostureTask(cost=1e-3)
    posture_task.set_target(workloads.ur5_posture_reference(model))
    limits = [ConfigurationLimit(model), VelocityLimit(model)]
    from pink_b200 import BatchedIK

    ik = BatchedIK(model, [frame_task, posture_task], workloads.UR5_DT, damping=workloads.UR5_DAMPING,
                   limits=limits, safety_break=True, device=device, batch_size=B)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        i = k % NBUF
        ik.solve(qs[i], ts[i], vs[i], ss[i])

    # ---- device-resident throughput ("value") ---------------------------------
    # The K timed steps are K launches of the fused kernel on rotating buffer sets.  They
    # are submitted as replays of a CUDA graph holding NBUF consecutive steps (plus K mod
    # NBUF direct launches), so that the figure measures the device, not how fast this
    # host thread can issue 20 us kernels; `eager_ms_per_step` reports the direct-launch
    # loop next to it.
    for k in range(args.warmup):
        step(k)
    barrier()
    graph = None
    try:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            # thread-local capture mode: the NCCL watchdog thread of a multi-rank run may
            # issue CUDA calls (event queries) while this thread is capturing
            with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
                for k in range(NBUF):
                    step(k)
        torch.cuda.current_stream(device).wait_stream(side)
        graph.replay()
        barrier()
    except Exception as exc:  # pragma: no cover - graph capture is an optimisation only
        graph = None
        print(f"[bench] CUDA graph submission unavailable, direct launches: {exc}", file=sys.stderr)

    def run_steps(n):
        """exactly n steps; returns the number of launches issued outside graphs"""
        reps, tail = (n // NBUF, n % NBUF) if graph is not None else (0, n)
        for _ in range(reps):
            graph.replay()
        for k in range(tail):
            step(k)
        return tail

    launches0 = _cabi.load().pk_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        direct = run_steps(args.steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        # keep the GPU under the same load a little longer so the sampler sees it
        t_end = time.time() + 0.6
        while time.time() < t_end:
            run_steps(NBUF)
        torch.cuda.synchronize()
    # launches inside the timed region: every step is one launch of the fused kernel
    # (graph replays launch the captured kernels; pk_launch_count only sees direct calls)
    launches = args.steps
    assert _cabi.load().pk_launch_count() - launches0 >= direct
    bad = int(sum(int((s != 0).sum().item()) for s in ss))

    # direct-launch loop (host launch latency included), for reference
    barrier()
    d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_direct = min(args.steps, 2000)
    d0.record()
    for k in range(n_direct):
        step(k)
    d1.record()
    barrier()
    eager_ms = d0.elapsed_time(d1) / n_direct
    g_ms = ms / args.steps if graph is not None else None

    # ---- end to end through host buffers ("e2e") ---------------------------------
    # Every step copies its inputs from pinned host memory and its results back, inside
    # the timed region.  Steps are submitted round-robin on `--e2e-streams` CUDA streams
    # with one pinned buffer set per stream (double buffering): step k+1's H2D overlaps
    # step k's D2H on the full-duplex link; all streams are joined before the stop event.
    NS = max(1, args.e2e_streams)
    q_h = [torch.empty((B, 6), dtype=torch.float32).pin_memory() for _ in range(NS)]
    t_h = [torch.empty((B, 12), dtype=torch.float32).pin_memory() for _ in range(NS)]
    v_h = [torch.empty((B, 6), dtype=torch.float32).pin_memory() for _ in range(NS)]
    s_h = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(NS)]
    for i in range(NS):
        q_h[i].copy_(qs[i % NBUF].cpu())
        t_h[i].copy_(ts[i % NBUF].cpu())
    e2e_streams = [torch.cuda.Stream(device=device) for _ in range(NS)] if NS > 1 else [torch.cuda.current_stream(device)]

    def e2e_step(k):
        i = k % NS
        with torch.cuda.stream(e2e_streams[i]):
            ik.solve_host(q_h[i], t_h[i], v_h[i], s_h[i])

    def e2e_join():
        cur = torch.cuda.current_stream(device)
        for st_ in e2e_streams:
            if st_ is not cur:
                cur.wait_stream(st_)

    for k in range(max(3, args.warmup)):
        e2e_step(k)
    e2e_join()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = args.steps
    h0.record()
    for st_ in e2e_streams:
        st_.wait_event(h0)
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_join()
    h1.record()
    barrier()
    e2e_ms = h0.elapsed_time(h1)
    e2e_ok = bool(torch.allclose(v_h[0], vs[0].cpu(), atol=0, rtol=0)) if NBUF >= 1 else True

    # ---- max over ranks -----------------------------------------------------------
    times = torch.tensor([ms, e2e_ms, g_ms if g_ms is not None else 0.0], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, g_ms_max = (float(x) for x in times.cpu())
    if g_ms is not None:
        g_ms = g_ms_max

    # ---- solve + gather variant (multi-GPU only): v gathered on every rank --------
    gather_ms = None
    if world > 1:
        gathered = torch.empty((world * B, 6), dtype=torch.float32, device=device)
        for k in range(3):
            step(k)
            dist.all_gather_into_tensor(gathered, vs[k % NBUF])
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for k in range(args.steps):
            step(k)
            dist.all_gather_into_tensor(gathered, vs[k % NBUF])
        a1.record()
        barrier()
        t = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = float(t.item())

    if rank == 0:
        peak, peak_kind = load_peaks()
        total_steps = world * B * args.steps
        value = total_steps / (ms * 1e-3)
        kern_ms = g_ms if g_ms is not None else ms / args.steps
        achieved = B * BYTES_PER_STEP / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "gpu_launches": int(launches) * world,
            "clocks": clocks.summary(),
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": committed_traffic(), "peak_kind": peak_kind,
                "kernel": "pk::ik_chain_kernel<6>", "kernel_ms": kern_ms,
                "timing": "CUDA graph replay of %d launches" % NBUF if g_ms is not None else "eager launches",
                "eager_ms_per_step": eager_ms,
                "algorithmic_bytes_per_launch": B * BYTES_PER_STEP,
            },
            "e2e": {
                "value": world * B * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": B * (6 + 12) * 4, "d2h_bytes_per_step": B * (6 + 1) * 4,
                "ms_per_step": e2e_ms / e2e_steps, "bitwise_equal_to_device_path": e2e_ok,
                "api": "BatchedIK.solve_host -> pk_solve_ik_prepared_host (pinned host buffers; mode %s; %d submission stream(s))" % (os.environ.get("PK_HOST_MODE", "0"), NS),
            },
            "nonzero_status": bad,
        }
        if gather_ms is not None:
            line["solve_plus_allgather"] = {"value": total_steps / (gather_ms * 1e-3), "unit": UNIT}
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_block(B)
        if real_stdout is not None:
            os.write(real_stdout, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    # release the captured graph and the prepared problem before the process group goes
    graph = None
    ik = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--nbuf", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-streams", type=int, default=1, help="CUDA streams the e2e steps are submitted on (double buffering)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
