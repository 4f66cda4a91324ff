"""CPU checks of the CUDA kernel SOURCES through the SIMT emulator (tools/simt_emu: one fiber per CUDA thread, warp collectives
and the m8n8k4 tensor-pipe instruction emulated).  The build container has no GPU; these tests run the unmodified kernel code of
robust-nonlinear-mpc_b200/csrc/ against the same golden fixtures as the `-m gpu` tests, so that logic errors are found before
GPU time is spent.  They say nothing about the compiled sm_100a code (that is what `-m gpu` is for) and they are not a product
path: nothing under robust-nonlinear-mpc_b200/ imports the emulator."""
import numpy as np
import pytest

from oracle import models as om
from tests.helpers import load, relf
from tools import simt_emu

MODELS = ["pendulum", "quadrotor", "rocket"]


def _linearize(m, X, U):
    B, N = U.shape[0], U.shape[1]
    A, Bm, c = np.zeros((B, N, m.nx, m.nx)), np.zeros((B, N, m.nx, m.nu)), np.zeros((B, N, m.nx))
    for b in range(B):
        for k in range(N):
            A[b, k], Bm[b, k] = m.jacobians(X[b, k], U[b, k])
            c[b, k] = m.ddyn(X[b, k], U[b, k]) - X[b, k + 1]
    return A, Bm, c


QP_KERNELS = {"qp_mma_kernel": 0, "qp_warp_kernel": 1, "qp_tm_kernel": 2}


@pytest.fixture(params=sorted(QP_KERNELS))
def qp_kernel(request):
    simt_emu.set_qp_kernel(QP_KERNELS[request.param])
    yield request.param
    simt_emu.set_qp_kernel(0)


@pytest.mark.parametrize("name", MODELS)
def test_qp_mma_kernel_source_vs_oracle_golden(name, qp_kernel):
    """qp_mma_kernel (rnmpc_qp_mma.cuh) and qp_warp_kernel (rnmpc_qp_warp.cuh) on the emulator vs the oracle's exact QP solutions
    (same bounds as the GPU test)."""
    g = load(f"qp_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    A, Bm, c = _linearize(m, g["X"], g["U"])
    simt_emu.set_order(0)
    for tight in (False, True):
        idx = np.where(g["tight"] == tight)[0][:2]
        if len(idx) == 0:
            continue
        out = simt_emu.qp_mma(name, m, cfg, A[idx], Bm[idx], c[idx], g["X"][idx], g["U"][idx], g["x0"][idx],
                              g["backoff"][idx] if tight else None, g["backoff_f"][idx] if tight else None)
        assert not out["status"].any() and out["iters"].max() <= 30
        for j, i in enumerate(idx):
            assert relf(out["dX"][j], g["dX"][i]) < 1e-6 and relf(out["dU"][j], g["dU"][i]) < 1e-6
            scale = max(np.abs(g["mu"][i]).max(), 1.0)
            assert np.abs(out["mu"][j] - g["mu"][i]).max() < 1e-6 * scale
            assert np.abs(out["mu_f"][j] - g["mu_f"][i]).max() < 1e-6 * scale
            assert out["mu"][j].min() > -1e-12


def test_qp_mma_kernel_source_is_schedule_independent(qp_kernel):
    """A missing barrier makes the result depend on the order in which the emulator runs the threads between barriers;
    ascending, descending and pseudo-random orders must agree bit for bit.  Also: the result must not depend on the block
    index (the kernel rotates its warps' roles with it)."""
    name = "rocket"
    g = load(f"qp_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    A, Bm, c = _linearize(m, g["X"], g["U"])
    i = int(np.where(g["tight"])[0][0])
    idx = [i, i, i, i]
    res = []
    for order in (0, 1, 2):
        simt_emu.set_order(order)
        res.append(simt_emu.qp_mma(name, m, cfg, A[idx], Bm[idx], c[idx], g["X"][idx], g["U"][idx], g["x0"][idx],
                                   g["backoff"][idx], g["backoff_f"][idx]))
    simt_emu.set_order(0)
    for r in res:
        for k in r:
            assert np.array_equal(r[k], res[0][k]), k
            assert all(np.array_equal(r[k][0], r[k][b]) for b in range(1, 4)), k


@pytest.mark.parametrize("name", MODELS)
def test_sls_mma_kernel_source_vs_reference_numba_golden(name):
    """sls_mma_kernel (rnmpc_sls_mma.cuh) on the emulator vs the outputs of the reference's own Numba kernels
    (tests/golden/sls_*.npz, solver/fast_SLS_jit.py:43-188), same assertions as the GPU test."""
    from oracle.gen_golden import sls_inputs
    g = load(f"sls_{name}.npz")
    m, cfg, A, B, eta, eta_f = sls_inputs(name)
    N = cfg["N"]
    simt_emu.set_order(2)
    out = {k: (None if v is None else v[0]) for k, v in simt_emu.sls_mma(name, m, cfg, A, B, eta=eta, eta_f=eta_f).items()}
    simt_emu.set_order(0)
    tol = 1e-9
    assert relf(out["K"], g["K"], 0.0) < tol
    assert relf(np.einsum("kjii->kji", out["S"]), g["S_diag"], 0.0) < tol
    assert relf(out["Phi_x"][N], g["Phi_x_last"], 0.0) < tol
    assert relf(out["Phi_u"].sum(axis=(2, 3)), g["Phi_u_sum"], 0.0) < tol
    for key in ("beta", "beta_f"):
        assert relf(out[key], g[key], 0.0) < tol, key
    assert np.array_equal(out["S"], np.swapaxes(out["S"], 2, 3))    # symmetrised exactly


def test_soft_qp_kernel_source_vs_dense_oracle():
    """The soft-constrained comparator's QP (qp_warp_kernel, SOFT: slack variables eliminated entry by entry inside the interior
    point) on the emulator vs the dense QP in (delta, gamma) of oracle/soft.py, for a state inside the box (all slacks zero: the
    hard-constrained solution) and one outside (active slacks)."""
    from oracle import scp as oscp, soft as osoft
    name = "pendulum"
    m, cfg = om.CONFIGS[name]()
    N, Q, R, Qf = cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"]
    X0, U0, ok = oscp.nominal_sqp(m, N, Q, R, Qf, om.PENDULUM_X0, iters=10)
    assert ok
    A, Bm, c = _linearize(m, X0.T[None], U0.T[None])
    sq = osoft.SoftQP(m, N, Q, R, Qf, 1e6, 1e6)
    x_out = om.PENDULUM_X0.copy()
    x_out[0] = 1.05 * m.x_ub[0]
    simt_emu.set_order(2)
    for xs, active in ((om.PENDULUM_X0, False), (x_out, True)):
        ref = sq.solve(X0, U0, xs)
        assert max(ref["kkt"].values()) < 1e-7
        out = simt_emu.qp_soft(name, m, cfg, A, Bm, c, X0.T[None], U0.T[None], np.asarray(xs)[None])
        assert out["status"][0] == 0 and out["iters"][0] <= 60
        assert np.abs(out["dX"][0].T - ref["dX"]).max() < 1e-9 and np.abs(out["dU"][0].T - ref["dU"]).max() < 1e-9
        assert np.abs(out["gamma"][0] - ref["gamma"]).max() < 1e-9 and np.abs(out["gamma_f"][0] - ref["gamma_f"]).max() < 1e-9
        assert (ref["gamma"].max() > 0.1) == active
    simt_emu.set_order(0)
