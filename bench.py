#!/usr/bin/env python
"""Benchmark of the SCP / fast-SLS robust-MPC hot path (BASELINE.json metric: rocket SCP-SLS solves/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

A "step" is one closed-loop MPC step of every scenario of the batch: warm-start shift, one SCP_SLS.solve (linearise,
nominal QP, eta, Riccati sweep, Phi propagation, back-offs, tightened QP, nominal update) and the disturbed plant
step -- i.e. `batch` solves.  Workload at N = 1: BASELINE.json configs[3] "rocket compare closed loop batched over 4,096
disturbance draws" (N = 15 horizon, rti = 1, fast_sls_rti_steps = 1, 29 MPC steps = default warmup 3 + steps 26).  With N > 1
(torchrun, one rank per GPU) every rank owns its own `batch` scenarios (weak scaling, no collective on the solve path,
one statistics all_gather at the end): 8 x 8,192 = configs[4]'s 65,536 scenarios when --batch 8192.

`value` times the device-resident loop; `e2e` times the same steps through the host-buffer entry point (pinned host
x0 / disturbance in, next state + first input + flags out, copies inside the timed region).  `--impl reference` times
the CPU oracle port of the same path (the reference itself needs CasADi + OSQP, absent from the image) on the box's
host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rocket SCP-SLS solves/sec (fp64)"
UNIT = "solves/s"


# ------------------------------------------------------------------------------------------------ flop accounting
def flops_model(nx=17, nu=4, N=15):
    """Algorithmic flops (canonical box-constraint structure)."""
    nz, P = nx + nu, N * (N + 1) // 2
    ni, ni_f = 2 * nz, 2 * nx
    f_ric = 6 * nu * nx**2 + 4 * nx**3 + 4 * nu**2 * nx + (2 / 3) * nu**3 + 2 * nx**2
    sls = P * (f_ric + nz) + (N + 1) * nx + P * (2 * nu * nx * nx + 2 * nx**2 * nu + nx**2 + 2 * nx**2 * nx) \
        + P * 2 * nz * nx + (N + 1) * 2 * nx * nx
    lin = N * 3 * 2 * nx**2 * nz
    qp_fact = N * f_ric                                                        # per factorisation
    qp_solve = N * (4 * nx**2 + 8 * nx * nu + 2 * nu**2) + 10 * (N * ni + ni_f)   # per KKT solve
    return dict(f_ric=f_ric, sls=sls, lin=lin, qp_fact=qp_fact, qp_solve=qp_solve)


# ------------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    QUERY = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------- CPU baseline
def cpu_loop(n_scen, warmup, steps, threads, seed=2025):
    """CPU port of the whole solve (oracle/cpu_port/rnmpc_cpu.cpp: C++, one scenario per std::thread worker, all host cores) on the
    GPU arm's workload: the same `n_scen` scenarios (same seed), `warmup` untimed MPC steps and `steps` timed ones, and the same
    Monte-Carlo rule (a scenario whose solve failed is frozen and is neither worked on nor counted afterwards).  The reference's own
    CPU path needs CasADi + OSQP (absent from the image: profiles/r2_probe_refdeps.log)."""
    from oracle import models as om, scp as oscp
    from oracle import cpu_port
    from rnmpc_b200.workloads import rocket_scenarios
    m, cfg = om.rocket_config()
    x0, W = rocket_scenarios(n_scen, seed=seed, steps=2 * (warmup + steps) + 4)
    X0, U0, ok = oscp.nominal_sqp(m, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], x0[0], iters=10)     # untimed (same x0 for all scenarios)
    loops = cpu_port.ResumableLoops(m, cfg, x0, np.repeat(X0.T[None], n_scen, 0), np.repeat(U0.T[None], n_scen, 0), True, threads)
    if warmup > 0:
        loops.advance(W[:, :warmup], warmup)
    t0 = time.perf_counter()
    live, failed = loops.advance(W[:, warmup:warmup + steps], steps)
    wall = time.perf_counter() - t0
    return live / wall, dict(solves=live, wall_s=wall, failed=failed, dead_after=int((loops.alive == 0).sum()), scenarios=n_scen)


def cpu_single_latency(steps=29):
    """Single-instance solve latency of the CPU port: one scenario, one thread, wall time per MPC step (median over 3 runs)."""
    from oracle import models as om, scp as oscp
    from oracle import cpu_port
    from tests.helpers import rocket_scenarios
    m, cfg = om.rocket_config()
    x0, W = rocket_scenarios(1, seed=777, steps=steps)
    X0, U0, ok = oscp.nominal_sqp(m, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], x0[0], iters=10)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        cpu_port.closed_loop(m, cfg, x0, W, steps, X0.T[None].copy(), U0.T[None].copy(), 1)
        ts.append((time.perf_counter() - t0) / steps)
    return 1e3 * float(np.median(ts))


def gpu_single_latency(m, cfg, device, steps=40):
    """Single-solve latency (BASELINE.json metric, second half): one rocket MPC instance, batch 1, driven from the host as a
    controller would (measured state in from host memory, solve, nominal trajectory + back-offs out to host memory); wall clock per
    call over a disturbed closed loop, for the C-ABI entry point (rnmpc_solve_host) and for the drop-in call itself
    (solver.SCP_SLS_jit.SCP_SLS.solve(x0), the call that replaces solver/SCP_SLS_jit.py:65-152)."""
    import torch
    from rnmpc_b200.engine import BatchSolver
    from rnmpc_b200.solver.SCP_SLS_jit import SCP_SLS
    from rnmpc_b200.workloads import rocket_scenarios
    x_ub, x_lb, u_ub, u_lb, _, _ = m.box()
    x0, W = rocket_scenarios(1, seed=777, steps=steps + 8)

    def pct(lat):
        lat = np.sort(np.array(lat))
        return {"p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)), "min_ms": float(lat[0]), "n": int(lat.size)}

    # (a) C ABI: rnmpc_solve_host
    s = BatchSolver("rocket", cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E, x_ub, x_lb,
                    u_ub, u_lb, rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=1, device=device)
    x = s.dev(x0).clone()
    s.init_nominal(x, 10)
    X0, U0 = s.buffer("X", 1), s.buffer("U", 1)
    lat, fails = [], 0
    for i in range(steps + 8):
        if i > 0:
            s.reset_warm_start(1)
        xh = np.array(torch.as_tensor(x).cpu().numpy())
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        out = s.solve_host(xh)
        dt = time.perf_counter() - t0
        fails += int(out["flags"][0] & 1)
        if i >= 8:
            lat.append(1e3 * dt)
        s.plant_step(x, s.dev(W[:, i]))
    res = pct(lat)
    res.update({"failed": fails, "what": "rocket SCP_SLS.solve, batch 1, host buffers in/out (rnmpc_solve_host: one CUDA graph per call), wall clock"})
    s.close()
    # (b) the drop-in class call
    sc = SCP_SLS(cfg["N"], cfg["Q"], cfg["R"], m, cfg["Qf"], Q_reg=cfg["Q_reg"], R_reg=cfg["R_reg"], Q_reg_f=cfg["Q_reg_f"], rti=cfg["rti"],
                 fast_sls_rti_steps=cfg["fast_sls_rti_steps"], verbose=False, save_it_data=False, timing=False, device=device)
    sc.set_nominal_trajectory(X0[0].T, U0[0].T)
    xk = x0[0].copy()
    lat2, fails2 = [], 0
    for i in range(steps + 8):
        if i > 0:
            sc.reset_warm_start()
        t0 = time.perf_counter()
        sol = sc.solve(xk)
        dt = time.perf_counter() - t0
        fails2 += 0 if sol["success"] else 1
        if i >= 8:
            lat2.append(1e3 * dt)
        xk = np.array(m.ddyn(xk, sol["primal_u"][:, 0])).reshape(-1) + m.E @ W[0, i]
    res["scp_sls_solve"] = dict(pct(lat2), failed=fails2,
                                what="solver.SCP_SLS_jit.SCP_SLS.solve(x0) of this package (NumPy in, result dict out; keys beyond the "
                                     "drivers' set are copied on first access), wall clock")
    return res


# ------------------------------------------------------------------------------------------------------------ main
def qp_traffic_record():
    """dram__bytes_read + dram__bytes_write of ONE launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r2_qp_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep of the same tools/profile_step.py command)."""
    path = os.path.join(ROOT, "profiles", "r2_qp_traffic.json")
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


def run_loops(solver, x, Wd, n_steps, first_step, stream):
    """n_steps closed-loop MPC steps of the solver's batch on `stream`, no host synchronisation."""
    import torch
    with torch.cuda.stream(stream):
        for j in range(n_steps):
            if first_step + j > 0:
                solver.reset_warm_start(x.shape[0])
            solver.solve(x)
            solver.plant_step(x, Wd[first_step + j])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=26)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=1, help="independent scenario groups per GPU, each on its own CUDA stream (the groups' "
                    "kernels overlap, so a slow interior-point straggler only delays its own group)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the 65,536-scenario Monte-Carlo sweep (BASELINE configs[4]) at N > 1")
    ap.add_argument("--config4-scenarios", type=int, default=65536)
    ap.add_argument("--option", action="append", default=[], help="solver option name=value (rnmpc_set_option), for experiments")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = (f"BASELINE configs[3]: rocket robust closed loop (SCP-SLS, N=15, rti=1, fast_sls_rti_steps=1), {a.batch} disturbance-draw "
                f"scenarios per GPU from the fixed x0, a failed scenario is frozen; one step = one MPC step of every scenario")

    if a.impl == "reference":
        # the reference's CPU path on the box's host cores (C++ port: CasADi + OSQP are absent); only rank 0 works.  Same scenarios,
        # same warm-up / timed MPC steps and the same freeze-on-failure rule as the GPU arm's rank 0.
        if rank != 0:
            return
        v, info = cpu_loop(a.batch, a.warmup, a.steps, cores)
        sample = (f"the GPU arm's rank-0 workload itself: {a.batch} rocket scenarios, {a.warmup} untimed + {a.steps} timed MPC steps, {cores} threads, "
                  f"{info['solves']} live solves in {info['wall_s']:.1f} s, {info['dead_after']} scenarios dead at the end (C++ CPU port oracle/cpu_port)")
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * info["wall_s"] / max(a.steps, 1), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "global_batch": a.batch,
                           "note": "reference needs CasADi+OSQP (absent: profiles/r2_probe_refdeps.log); timed: C++ CPU port of the same path on all host "
                                   "cores, one rank's batch (the CPU does not scale with --gpus); value counts live solves only"},
                "dead_fraction": 1.0 - info["solves"] / float(a.batch * max(a.steps, 1)),
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
                "single_solve_latency": {"p50_ms": cpu_single_latency(), "what": "CPU port, one scenario on one thread, wall per MPC step"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # The contract is ONE JSON line on stdout.  Native libraries write there too (with NCCL_DEBUG=VERSION, as on the GPU boxes,
    # NCCL printf()s its version banner to stdout; NCCL_DEBUG_FILE does not redirect that line -- tried), so file descriptor 1 is
    # pointed at stderr for the whole run and the result line goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    result_out = os.fdopen(result_fd, "w")

    import torch
    import torch.distributed as dist
    # this arm imports the product only: nothing under oracle/ (the checker) is loaded here
    from rnmpc_b200.engine import BatchSolver, fp64_peak_tflops
    from rnmpc_b200.sharding import gather_statistics, shard_bounds
    from rnmpc_b200.workloads import rocket_config, rocket_scenarios

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, Wm = a.batch, a.steps, a.warmup
    G = max(1, min(a.groups, B))
    total_steps = K + Wm
    m, cfg = rocket_config()
    x_ub, x_lb, u_ub, u_lb, _, _ = m.box()
    options = [o.split("=") for o in a.option]

    def make_solver(cap):
        s_ = BatchSolver("rocket", cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E, x_ub, x_lb,
                         u_ub, u_lb, rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=cap, device=local_rank)
        s_.set_option("sticky_fail", 1)     # Monte-Carlo semantics: a scenario whose solve failed is dead (the reference's driver would stop)
        for k_, v_ in options:
            s_.set_option(k_, float(v_))
        return s_

    x0, W = rocket_scenarios(B, seed=2025 + rank, steps=2 * total_steps + 4)
    bounds = [(g * B // G, (g + 1) * B // G) for g in range(G)]
    solvers, streams, xs, Wds = [], [], [], []
    for lo, hi in bounds:
        s = make_solver(hi - lo)
        s.enable_timing(2)                  # event pair around every hot-kernel launch, summed after the timed region (no sync inside)
        solvers.append(s)
        streams.append(torch.cuda.Stream(device=dev))
        Wds.append(s.dev(np.ascontiguousarray(np.swapaxes(W[lo:hi], 0, 1))))     # [steps][b][nx], resident
        xs.append(s.dev(x0[lo:hi]).clone())
    peak_dfma = fp64_peak_tflops(local_rank, "dfma")
    peak_dmma = fp64_peak_tflops(local_rank, "dmma")
    peak = max(peak_dfma, peak_dmma)
    for s, x in zip(solvers, xs):
        s.init_nominal(x, 10)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    step_idx = [0]

    def device_steps(n):
        """n closed-loop MPC steps of every group; each group's launches go to its own stream, no host synchronisation."""
        for s, st, x, Wd in zip(solvers, streams, xs, Wds):
            run_loops(s, x, Wd, n, step_idx[0], st)
        step_idx[0] += n

    barrier()
    device_steps(Wm)
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    barrier()
    launches0 = sum(s.launch_count() for s in solvers)
    for s in solvers:
        for kname in ("linearize", "qp", "sls"):
            s.kernel_time(kname)                    # forget the warm-up launches
    ipm0 = sum(s.ipm_iterations() for s in solvers)
    live0 = sum(s.live_solves() for s in solvers)
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream(dev))
    for st in streams:
        st.wait_event(e0)
    device_steps(K)
    ends = []
    for st in streams:
        e = torch.cuda.Event(enable_timing=True)
        e.record(st)
        ends.append(e)
    barrier()
    launches = sum(s.launch_count() for s in solvers) - launches0
    # hot kernels over the timed region: summed launch durations (CUDA events on the launching streams) and launch counts
    region = {kname: [sum(v) for v in zip(*[s.kernel_time(kname) for s in solvers])] for kname in ("linearize", "qp", "sls")}
    ipm_region = sum(s.ipm_iterations() for s in solvers) - ipm0
    live_region = sum(s.live_solves() for s in solvers) - live0      # solves of scenarios that were alive when their step began
    for s in solvers:
        s.enable_timing(1)
    ms = max(e0.elapsed_time(e) for e in ends)          # device time of the slowest group's stream
    flags = np.concatenate([s.buffer("flags") for s in solvers])
    n_dead = int((flags & 1).sum())                        # scenarios dead after W + K steps of the device-resident loop
    dev_x_checksum = float(sum(np.abs(x.cpu().numpy()).sum() for x in xs))
    # per-kernel times + iteration counters on two extra (untimed) steps of group 0 alone, so that the kernel times are not
    # inflated by the other groups' kernels sharing the SMs and get_timings' sync stays outside the timed region
    t_qp, t_sls, t_jac, qp_iters = [], [], [], []
    for _ in range(2):
        s, x, Wd = solvers[0], xs[0], Wds[0]
        s.reset_warm_start(x.shape[0])
        s.solve(x)
        s.plant_step(x, Wd[step_idx[0]])
        tm = s.timings()
        t_qp.append(tm["t_qp_ms"]); t_sls.append(tm["t_backward_ms"]); t_jac.append(tm["t_jac_ms"])
        qp_iters.append(s.buffer("qp_iters").astype(np.float64))
        for s2, st, x2, Wd2 in list(zip(solvers, streams, xs, Wds))[1:]:
            s2.reset_warm_start(x2.shape[0]); s2.solve(x2); s2.plant_step(x2, Wd2[step_idx[0]])
        step_idx[0] += 1
    B0 = bounds[0][1] - bounds[0][0]

    # ---- e2e: same steps through host buffers (pinned host memory in, results out, copies inside the timed region)
    xh = torch.empty((B, m.nx), dtype=torch.float64, pin_memory=True)
    uh = torch.empty((B, m.nu), dtype=torch.float64, pin_memory=True)
    fh = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    Wpin = torch.as_tensor(np.ascontiguousarray(np.swapaxes(W, 0, 1))).pin_memory()
    xds, wds = [torch.empty_like(x) for x in xs], [torch.empty_like(x) for x in xs]

    def e2e_step():
        i = step_idx[0]
        for (lo, hi), s, st, xd, wd in zip(bounds, solvers, streams, xds, wds):
            with torch.cuda.stream(st):
                xd.copy_(xh[lo:hi], non_blocking=True)              # measured state: pinned host -> device
                wd.copy_(Wpin[i, lo:hi], non_blocking=True)         # this step's disturbance draw arrives from the host (pinned)
                if i > 0:
                    s.reset_warm_start(hi - lo)
                s.solve(xd)
                s.plant_step(xd, wd)
                Uv = torch.as_tensor(s.buffer_view("U", hi - lo))
                xh[lo:hi].copy_(xd, non_blocking=True)              # next state, applied input and flags: device -> pinned host
                uh[lo:hi].copy_(Uv[:, 0], non_blocking=True)
                fh[lo:hi].copy_(torch.as_tensor(s.buffer_view("flags", hi - lo)), non_blocking=True)
        torch.cuda.synchronize(dev)
        step_idx[0] += 1

    # the SAME scenarios and the SAME MPC steps as the device-resident measurement: state and nominal trajectory are
    # re-initialised (untimed), W warm-up steps and then K timed steps run through the host-buffer path
    xh.copy_(torch.as_tensor(x0))
    for (lo, hi), s in zip(bounds, solvers):
        s.init_nominal(s.dev(x0[lo:hi]), 10)
    barrier()                                               # the initialiser ran on the default stream, the steps run on the groups' streams
    step_idx[0] = 0
    n_e2e = K
    for _ in range(Wm):
        e2e_step()
    barrier()
    live1 = sum(s.live_solves() for s in solvers)
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    e2e_live = sum(s.live_solves() for s in solvers) - live1
    e2e_dead = int((fh.numpy() & 1).sum())                 # scenarios dead after W + K steps of the host-buffer loop
    e2e_x_checksum = float(np.abs(xh.numpy()).sum())
    clk = clocks.stop() if rank == 0 else None
    single = gpu_single_latency(m, cfg, local_rank) if rank == 0 else None

    # ---- reduce over ranks: max time, summed solves
    # the host-buffer loop ran the same scenarios and steps as the device-resident loop: on EVERY rank it must end in the same
    # place.  A loop that diverged (dead scenarios cost nothing) would report a meaningless throughput, so it is not reported.
    e2e_invalid = None
    if e2e_dead != n_dead or e2e_live != live_region or not (abs(e2e_x_checksum - dev_x_checksum) <= 1e-6 * abs(dev_x_checksum)):
        e2e_invalid = "rank %d: host-buffer loop ended with %d dead scenarios / %d live solves / state checksum %r, device-resident loop with %d / %d / %r" % (
            rank, e2e_dead, e2e_live, e2e_x_checksum, n_dead, live_region, dev_x_checksum)
        print("bench.py: e2e measurement INVALID: " + e2e_invalid, file=sys.stderr)
    tvec = torch.tensor([ms, e2e_ms, 1.0 if e2e_invalid else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(tvec[0]), float(tvec[1])
    if float(tvec[2]) > 0.0 and e2e_invalid is None:
        e2e_invalid = "the host-buffer loop diverged from the device-resident loop on another rank (see stderr)"
    bo = solvers[0].buffer("backoff")
    stats, per_rank = gather_statistics(dict(solves=live_region, failed_solves=n_dead, violations=0,
                                             cost_sum=float(sum(s.buffer("cost").sum() for s in solvers)),
                                             max_backoff=float(bo.max()), qp_iters=float(np.mean(qp_iters)) * B, elapsed_ms=ms), device=dev)
    e2e_stats, _ = gather_statistics(dict(solves=e2e_live), device=dev)

    # ---- BASELINE configs[4] (N > 1): 65,536 perturbed-x0 scenarios sharded over the ranks, 29 MPC steps, device-resident and host-buffer loops
    config4 = None
    if world > 1 and not a.no_config4:
        for s in solvers:
            s.close()
        solvers.clear()
        torch.cuda.empty_cache()
        lo4, hi4 = shard_bounds(a.config4_scenarios, rank, world)
        B4, steps4 = hi4 - lo4, 29
        x04, W4 = rocket_scenarios(B4, seed=[2025, rank], steps=steps4, x0_scale=0.05)
        s4 = make_solver(B4)
        st4 = torch.cuda.Stream(device=dev)
        W4d = s4.dev(np.ascontiguousarray(np.swapaxes(W4, 0, 1)))
        x4 = s4.dev(x04).clone()
        s4.init_nominal(x4, 10)
        n_init_fail = int((s4.buffer("flags") & 1).sum())
        barrier()
        lv0 = s4.live_solves()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(torch.cuda.current_stream(dev)); st4.wait_event(ea)
        run_loops(s4, x4, W4d, steps4, 0, st4)
        eb.record(st4)
        barrier()
        live4 = s4.live_solves() - lv0
        ms4 = ea.elapsed_time(eb)
        dead4 = int((s4.buffer("flags") & 1).sum())
        chk4 = float(np.abs(x4.cpu().numpy()).sum())
        # host-buffer loop on the same scenarios
        xh4 = torch.empty((B4, m.nx), dtype=torch.float64, pin_memory=True); xh4.copy_(torch.as_tensor(x04))
        oh4 = torch.empty((B4, m.nx + m.nu), dtype=torch.float64, pin_memory=True)
        fh4 = torch.empty((B4,), dtype=torch.int32, pin_memory=True)
        W4pin = torch.as_tensor(np.ascontiguousarray(np.swapaxes(W4, 0, 1))).pin_memory()
        xd4, wd4 = torch.empty_like(x4), torch.empty_like(x4)
        s4.init_nominal(s4.dev(x04), 10)
        barrier()
        lv1 = s4.live_solves()
        t0 = time.perf_counter()
        for i in range(steps4):
            with torch.cuda.stream(st4):
                xd4.copy_(xh4, non_blocking=True); wd4.copy_(W4pin[i], non_blocking=True)
                if i > 0:
                    s4.reset_warm_start(B4)
                s4.solve(xd4); s4.plant_step(xd4, wd4)
                oh4[:, :m.nx].copy_(xd4, non_blocking=True)
                oh4[:, m.nx:].copy_(torch.as_tensor(s4.buffer_view("U", B4))[:, 0], non_blocking=True)
                fh4.copy_(torch.as_tensor(s4.buffer_view("flags", B4)), non_blocking=True)
            torch.cuda.synchronize(dev)
            xh4.copy_(oh4[:, :m.nx])
        barrier()
        e2e4_s = time.perf_counter() - t0
        live4e = s4.live_solves() - lv1
        ok4 = (int((fh4.numpy() & 1).sum()) == dead4) and live4e == live4 and abs(float(np.abs(xh4.numpy()).sum()) - chk4) <= 1e-6 * abs(chk4)
        v4 = torch.tensor([ms4, e2e4_s, 0.0 if ok4 else 1.0], dtype=torch.float64, device=dev)
        c4 = torch.tensor([float(live4), float(live4e), float(dead4), float(n_init_fail), float(B4)], dtype=torch.float64, device=dev)
        dist.all_reduce(v4, op=dist.ReduceOp.MAX)
        dist.all_reduce(c4, op=dist.ReduceOp.SUM)
        config4 = {"workload": "BASELINE configs[4]: rocket Monte-Carlo sweep, x0 = x0_fixed + 0.05 x half-range x U[-1,1] (quaternion renormalised), "
                               "disturbance draws U[-1,1]^17 per step, 29 MPC steps, scenarios sharded contiguously over the ranks (seed [2025, rank])",
                   "global_batch": int(c4[4]), "steps": steps4, "value": float(c4[0]) / (float(v4[0]) * 1e-3), "unit": UNIT,
                   "e2e": None if float(v4[2]) > 0 else float(c4[1]) / float(v4[1]), "live_solves": int(c4[0]),
                   "dead": int(c4[2]), "dead_fraction_of_scenarios": float(c4[2]) / float(c4[4]), "failed_in_initialiser": int(c4[3]),
                   "ms_total": float(v4[0]), "e2e_s_total": float(v4[1])}
        s4.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = stats["solves"] / (ms_all * 1e-3)
    e2e_value = None if e2e_invalid else e2e_stats["solves"] / (e2e_ms_all * 1e-3 * n_e2e)

    fm = flops_model()
    # roofline of the dominant kernel over the TIMED REGION (this rank): every QP launch of the region is bracketed by CUDA events
    # on its stream; the interior-point iterations of all instances are counted on the device
    per_it = fm["qp_fact"] + 2 * fm["qp_solve"]                # algorithmic flops of one interior-point iteration of one instance
    qp_ms_total, qp_launches = region["qp"]
    iters_per_solve = ipm_region / float(max(live_region, 1))  # over the 2 QPs of a live solve
    t_qp_ms = qp_ms_total / max(qp_launches, 1)                # average launch duration
    qp_flops_launch = ipm_region * per_it / max(qp_launches, 1)
    flops_per_solve = fm["lin"] + fm["sls"] + iters_per_solve * per_it
    achieved = ipm_region * per_it / (qp_ms_total * 1e-3) * 1e-12
    sls_ms_total, sls_launches = region["sls"]
    sls_achieved = fm["sls"] * B * sls_launches / max(G, 1) / (sls_ms_total * 1e-3) * 1e-12 if sls_ms_total > 0 else None
    h2d = 2 * B * m.nx * 8
    d2h = B * (m.nx + m.nu) * 8 + B * 4
    tr = qp_traffic_record()
    traffic = None
    if tr and tr.get("batch"):
        traffic = (tr["dram_bytes_read"] + tr["dram_bytes_write"]) * B0 / float(tr["batch"])
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_all / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "horizon": cfg["N"], "parallelism": f"scenario-sharded x{world} GPUs, {G} stream groups per GPU",
                   "l2": "per-step working set ~%.0f MB per GPU > 126 MB L2 (no flush needed)" % (B * 0.2),
                   "counting": "value and e2e count the solves of scenarios that were alive when their MPC step began (device counter); frozen scenarios are not counted",
                   "flops_per_solve": flops_per_solve, "qp_ipm_iters_per_solve": iters_per_solve,
                   "whole_solve_fp64_frac": value / world * flops_per_solve / (peak * 1e12),
                   "kernel_ms_avg_over_region": {k_: (v_[0] / max(v_[1], 1)) for k_, v_ in region.items()},
                   "kernel_launches_in_region": {k_: int(v_[1]) for k_, v_ in region.items()},
                   "kernel_share_of_step": {k_: v_[0] / ms for k_, v_ in region.items()},
                   "kernel_ms_after_region": {"linearize": float(np.mean(t_jac)), "qp_last": float(np.mean(t_qp)), "fast_sls": float(np.mean(t_sls))},
                   "fast_sls_fp64_tflops": sls_achieved},
        "clocks": clk,
        "dead_fraction": 1.0 - stats["solves"] / float(B * world * K),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_all,
                "steps": n_e2e, "live_solves": int(e2e_stats["solves"]), "dead_scenarios_after_loop": e2e_dead, "state_checksum": e2e_x_checksum,
                "invalid": e2e_invalid},
        "device_loop": {"live_solves": int(stats["solves"]), "dead_scenarios_after_loop": n_dead, "state_checksum": dev_x_checksum},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "kernel": "qp_tm_kernel<17,4,15>", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "peak_dfma": peak_dfma, "peak_dmma": peak_dmma,
                     "peak_source": "measured live on this GPU (rnmpc_fp64_peak / rnmpc_fp64_mma_peak: DFMA and mma.sync.m8n8k4.f64 probes; "
                                    "MEASURED_PEAKS.json has no fp64 entry); peak = the larger of the two",
                     "traffic": traffic,
                     "traffic_note": None if tr is None else ("dram__bytes_read + dram__bytes_write per launch from the committed ncu --set full capture (%s), scaled "
                                                              "to this batch; the launch's inputs + outputs are 0.26 GB at batch 4096" % tr.get("source", "?")),
                     "launch_ms_avg": t_qp_ms, "launches": int(qp_launches), "flops_per_launch": qp_flops_launch,
                     "note": "all QP launches of the timed region on rank 0: algorithmic flops (interior-point iterations counted on the device x "
                             "0.49 MFLOP) / summed CUDA-event durations.  The kernel is bound by the latency of its sequential stage recursion "
                             "(DESIGN.md section 5), not by the pipe or HBM"},
        "dead_scenarios": int(stats["failed_solves"]),
        "single_solve_latency": single,
    }
    if config4 is not None:
        line["config4"] = config4
        line["config"]["workload"] += "; config4 = BASELINE configs[4] (65,536 perturbed-x0 scenarios over the ranks) measured after it"
    if not a.no_cpu_baseline and world == 1:
        v, info = cpu_loop(B, Wm, K, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"this run's own workload: {B} rocket scenarios, {Wm} untimed + {K} timed MPC steps, {cores} threads, C++ CPU port "
                                          f"(oracle/cpu_port; the reference's CasADi+OSQP path is not installable here), {info['solves']} live solves in "
                                          f"{info['wall_s']:.1f} s, {info['dead_after']} scenarios dead at the end"}
        line["single_solve_latency"]["cpu_port_p50_ms"] = cpu_single_latency()
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
