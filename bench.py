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
def cpu_baseline(n_scen, steps, threads):
    """CPU port of the whole solve (oracle/cpu_port/rnmpc_cpu.cpp: C++, one scenario per std::thread worker, all host cores)
    on a bounded sample of the bench workload.  The reference's own CPU path needs CasADi + OSQP (absent from the image)."""
    from oracle import models as om, scp as oscp
    from oracle import cpu_port
    from tests.helpers import rocket_scenarios
    m, cfg = om.rocket_config()
    x0, W = rocket_scenarios(n_scen, seed=4242, steps=steps)
    X0, U0, ok = oscp.nominal_sqp(m, cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], x0[0], iters=10)     # untimed (same x0 for all scenarios)
    X0b, U0b = np.repeat(X0.T[None], n_scen, 0), np.repeat(U0.T[None], n_scen, 0)
    cpu_port.closed_loop(m, cfg, x0[:threads], W[:threads, :2], 2, X0b[:threads], U0b[:threads], threads)          # warm-up
    t0 = time.perf_counter()
    _, _, nfail = cpu_port.closed_loop(m, cfg, x0, W, steps, X0b, U0b, threads)
    wall = time.perf_counter() - t0
    return n_scen * steps / wall, dict(solves=n_scen * steps, wall_s=wall, failed=nfail)


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
    controller would (measured state in from host memory, solve, nominal trajectory + back-offs out to host memory); wall
    clock per rnmpc_solve_host call, percentiles over a disturbed closed loop."""
    import torch
    from rnmpc_b200.engine import BatchSolver
    from rnmpc_b200.workloads import rocket_scenarios
    x_ub, x_lb, u_ub, u_lb, _, _ = m.box()
    s = BatchSolver("rocket", cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E, x_ub, x_lb,
                    u_ub, u_lb, rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=1, device=device)
    x0, W = rocket_scenarios(1, seed=777, steps=steps + 8)
    x = s.dev(x0).clone()
    s.init_nominal(x, 10)
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
    lat = np.sort(np.array(lat))
    return {"p50_ms": float(np.percentile(lat, 50)), "p90_ms": float(np.percentile(lat, 90)), "min_ms": float(lat[0]), "n": int(lat.size),
            "failed": fails, "what": "rocket SCP_SLS.solve, batch 1, host buffers in/out (rnmpc_solve_host), wall clock"}


# ------------------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=26)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4096, help="scenarios per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--groups", type=int, default=1, help="independent scenario groups per GPU, each on its own CUDA stream (the groups' "
                    "kernels overlap, so a slow interior-point straggler only delays its own group)")
    ap.add_argument("--cpu-scenarios", type=int, default=0, help="CPU baseline sample: scenarios (default = 256 x host cores)")
    ap.add_argument("--cpu-steps", type=int, default=29, help="CPU baseline sample: MPC steps per scenario")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    workload = f"rocket robust closed loop (SCP-SLS, N=15, rti=1, fast_sls_rti_steps=1), {a.batch} disturbance-draw scenarios per GPU"

    if a.impl == "reference":
        # the reference's CPU path; only rank 0 works
        if rank != 0:
            return
        n_scen = a.cpu_scenarios or 256 * cores
        t_steps = []
        for _ in range(max(1, min(a.steps, 3))):          # bounded: each "step" = cpu_steps MPC steps of n_scen scenarios
            v, info = cpu_baseline(n_scen, a.cpu_steps, cores)
            t_steps.append((v, info))
        v = float(np.mean([t[0] for t in t_steps]))
        sample = f"{n_scen} rocket scenarios x {a.cpu_steps} MPC steps per timed step, {cores} threads (C++ CPU port oracle/cpu_port)"
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * float(np.mean([t[1]["wall_s"] for t in t_steps])), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload, "note": "reference needs CasADi+OSQP (absent); timed: C++ CPU port of the same path on all host cores"},
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
    from rnmpc_b200.sharding import gather_statistics
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
    x0, W = rocket_scenarios(B, seed=2025 + rank, steps=2 * total_steps + 4)
    bounds = [(g * B // G, (g + 1) * B // G) for g in range(G)]
    solvers, streams, xs, Wds = [], [], [], []
    for lo, hi in bounds:
        s = BatchSolver("rocket", cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], m.E, x_ub, x_lb,
                        u_ub, u_lb, rti=cfg["rti"], sls_steps=cfg["fast_sls_rti_steps"], capacity=hi - lo, device=local_rank)
        s.enable_timing(2)                  # event pair around every hot-kernel launch, summed after the timed region (no sync inside)
        s.set_option("sticky_fail", 1)      # Monte-Carlo semantics: a scenario whose solve failed is dead (the reference's driver would stop)
        solvers.append(s)
        streams.append(torch.cuda.Stream(device=dev))
        Wds.append(s.dev(np.ascontiguousarray(np.swapaxes(W[lo:hi], 0, 1))))     # [steps][b][nx], resident
        xs.append(s.dev(x0[lo:hi]).clone())
    peak = fp64_peak_tflops(local_rank)
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
        i0 = step_idx[0]
        for j in range(n):
            for s, st, x, Wd in zip(solvers, streams, xs, Wds):
                with torch.cuda.stream(st):
                    if i0 + j > 0:
                        s.reset_warm_start(x.shape[0])
                    s.solve(x)
                    s.plant_step(x, Wd[i0 + j])
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
    t_wall0 = time.perf_counter()
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
    for s in solvers:
        s.enable_timing(1)
    ms = max(e0.elapsed_time(e) for e in ends)          # device time of the slowest group's stream
    flags = np.concatenate([s.buffer("flags") for s in solvers])
    n_fail_last = int((flags & 1).sum())                   # scenarios dead after W + K steps of the device-resident loop
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
    wh = torch.empty((B, m.nx), dtype=torch.float64, pin_memory=True)
    oh = torch.empty((B, m.nx + m.nu), dtype=torch.float64, pin_memory=True)
    fh = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    for (lo, hi), x in zip(bounds, xs):
        xh[lo:hi].copy_(x)
    Wpin = torch.as_tensor(np.ascontiguousarray(np.swapaxes(W, 0, 1))).pin_memory()
    xds, wds = [torch.empty_like(x) for x in xs], [torch.empty_like(x) for x in xs]
    def e2e_step():
        i = step_idx[0]
        wh.copy_(Wpin[i])                                   # this step's disturbance draw arrives from the host
        for (lo, hi), s, st, xd, wd in zip(bounds, solvers, streams, xds, wds):
            with torch.cuda.stream(st):
                xd.copy_(xh[lo:hi], non_blocking=True)
                wd.copy_(wh[lo:hi], non_blocking=True)
                if i > 0:
                    s.reset_warm_start(hi - lo)
                s.solve(xd)
                s.plant_step(xd, wd)
                Uv = torch.as_tensor(s.buffer_view("U", hi - lo))
                oh[lo:hi, :m.nx].copy_(xd, non_blocking=True)
                oh[lo:hi, m.nx:].copy_(Uv[:, 0], non_blocking=True)
                fh[lo:hi].copy_(torch.as_tensor(s.buffer_view("flags", hi - lo)), non_blocking=True)
        torch.cuda.synchronize(dev)
        xh.copy_(oh[:, :m.nx])
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
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / n_e2e
    e2e_failed = int((fh.numpy() & 1).sum())               # scenarios dead after W + K steps of the host-buffer loop
    e2e_x_checksum = float(np.abs(xh.numpy()).sum())
    clk = clocks.stop() if rank == 0 else None
    single = gpu_single_latency(m, cfg, local_rank) if rank == 0 else None

    # ---- reduce over ranks: max time, summed solves
    # the host-buffer loop ran the same scenarios and steps as the device-resident loop: on EVERY rank it must end in the same
    # place.  A loop that diverged (dead scenarios cost nothing) would report a meaningless throughput, so it is not reported.
    e2e_invalid = None
    if e2e_failed != n_fail_last or not (abs(e2e_x_checksum - dev_x_checksum) <= 1e-6 * abs(dev_x_checksum)):
        e2e_invalid = "rank %d: host-buffer loop ended with %d dead scenarios / state checksum %r, device-resident loop with %d / %r" % (
            rank, e2e_failed, e2e_x_checksum, n_fail_last, dev_x_checksum)
        print("bench.py: e2e measurement INVALID: " + e2e_invalid, file=sys.stderr)
    tvec = torch.tensor([ms, e2e_ms, 1.0 if e2e_invalid else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tvec, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(tvec[0]), float(tvec[1])
    if float(tvec[2]) > 0.0 and e2e_invalid is None:
        e2e_invalid = "the host-buffer loop diverged from the device-resident loop on another rank (see stderr)"
    bo = solvers[0].buffer("backoff")
    stats, per_rank = gather_statistics(dict(solves=B * K, failed_solves=n_fail_last, violations=0,
                                             cost_sum=float(sum(s.buffer("cost").sum() for s in solvers)),
                                             max_backoff=float(bo.max()), qp_iters=float(np.mean(qp_iters)) * B, elapsed_ms=ms), device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = stats["solves"] / (ms_all * 1e-3)
    e2e_value = None if e2e_invalid else (B * world) / (e2e_ms_all * 1e-3)

    fm = flops_model()
    # roofline of the dominant kernel over the TIMED REGION (this rank): every QP launch of the region is bracketed by CUDA events
    # on its stream; the interior-point iterations of all instances are counted on the device
    per_it = fm["qp_fact"] + 2 * fm["qp_solve"]                # algorithmic flops of one interior-point iteration of one instance
    qp_ms_total, qp_launches = region["qp"]
    iters_per_solve = ipm_region / float(B * K)                # over the 2 QPs of a solve, frozen scenarios count as 0
    t_qp_ms = qp_ms_total / max(qp_launches, 1)                # average launch duration
    qp_flops_launch = ipm_region * per_it / max(qp_launches, 1)
    flops_per_solve = fm["lin"] + fm["sls"] + iters_per_solve * per_it
    achieved = ipm_region * per_it / (qp_ms_total * 1e-3) * 1e-12
    h2d = 2 * B * m.nx * 8
    d2h = B * (m.nx + m.nu) * 8 + B * 4
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms_all / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "global_batch": B * world, "horizon": cfg["N"], "parallelism": f"scenario-sharded x{world} GPUs, {G} stream groups per GPU",
                   "l2": "per-step working set ~%.0f MB per GPU > 126 MB L2 (no flush needed)" % (B * 0.2),
                   "flops_per_solve": flops_per_solve, "qp_ipm_iters_per_solve": iters_per_solve,
                   "whole_solve_fp64_frac": value / world * flops_per_solve / (peak * 1e12),
                   "kernel_ms_avg_over_region": {k_: (v_[0] / max(v_[1], 1)) for k_, v_ in region.items()},
                   "kernel_launches_in_region": {k_: int(v_[1]) for k_, v_ in region.items()},
                   "kernel_share_of_step": {k_: v_[0] / ms for k_, v_ in region.items()},
                   "kernel_ms_after_region": {"linearize": float(np.mean(t_jac)), "qp_last": float(np.mean(t_qp)), "fast_sls": float(np.mean(t_sls))}},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_all,
                "steps": n_e2e, "dead_scenarios_after_loop": e2e_failed, "state_checksum": e2e_x_checksum, "invalid": e2e_invalid},
        "device_loop": {"dead_scenarios_after_loop": n_fail_last, "state_checksum": dev_x_checksum},
        "gpu_launches": int(launches),
        "roofline": {"bound": "fp64", "kernel": "qp_ipm_kernel<17,4>", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak,
                     # dram__bytes_read + dram__bytes_write of ONE launch of this kernel at batch 4096 from the committed ncu capture
                     # (profiles/r1f_ncu_full_summary.csv: the tightened QP of a fresh step, 4.57 ms under ncu); scaled to this batch
                     "traffic": 11.80e9 * B0 / 4096.0,
                     "traffic_note": "static figure from the ncu capture, not measured in this run; 2.58 TB/s inside that capture = 40 % of the "
                                     "measured HBM copy bandwidth; the launch's inputs + outputs are only 0.26 GB (re-streamed working set)",
                     "launch_ms_avg": t_qp_ms, "launches": int(qp_launches), "flops_per_launch": qp_flops_launch,
                     "note": "all QP solves of the timed region on rank 0 (qp_ipm_kernel + its hand-off launch to qp_block_kernel): algorithmic flops (interior-point iterations counted on the "
                             "device x 0.49 MFLOP) / summed CUDA-event durations, vs the DFMA peak measured live on this GPU "
                             "(MEASURED_PEAKS.json has no fp64 entry).  Neither the FP64 nor the HBM roof is near: the kernel is latency "
                             "bound (DESIGN.md section 5)"},
        "failed_solves_last_step": int(stats["failed_solves"]),
        "single_solve_latency": single,
    }
    if not a.no_cpu_baseline:
        n_scen = a.cpu_scenarios or 256 * cores
        v, info = cpu_baseline(n_scen, a.cpu_steps, cores)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n_scen} rocket scenarios x {a.cpu_steps} MPC steps, {cores} threads, C++ CPU port (oracle/cpu_port; "
                                          f"the reference's CasADi+OSQP path is not installable here), {info['wall_s']:.1f} s wall, "
                                          f"{info['failed']} failed solves"}
        line["single_solve_latency"]["cpu_port_p50_ms"] = cpu_single_latency()
    result_out.write(json.dumps(line) + "\n")
    result_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
