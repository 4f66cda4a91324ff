"""CPU tests: the oracle against the golden fixtures (reference Numba kernels, SymPy Jacobians), KKT checks of the
oracle QP, the host-side layout logic, and the exported C ABI.  No GPU compute."""
import ctypes
import os
import sys
import re

import numpy as np
import pytest

from oracle import models as om
from oracle import qp as oqp
from oracle import scp as oscp
from oracle import sls as osls
from oracle.gen_golden import jac_points, sls_inputs
from tests.helpers import load, rel, relf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_sls_restatement_matches_reference_kernels(name):
    """oracle/sls.py vs outputs of the reference's own Numba kernels (tests/golden/sls_*.npz)."""
    g = load(f"sls_{name}.npz")
    m, cfg, A, B, eta, eta_f = sls_inputs(name)
    N, nx, nu = cfg["N"], m.nx, m.nu
    S, K = osls.backward_solve(N, nx, nu, A, B, m.G, m.Gf, eta, eta_f, cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"])
    Px, Pu = osls.propagate(A, B, np.stack([m.E] * (N + 1)), K)
    beta, beta_f, bo, bof = osls.backoff_from_phi(Px, Pu, m.G[:, :nx], m.G[:, nx:], m.Gf, 1e-10)
    assert rel(K, g["K"]) < 1e-13
    assert rel(np.einsum("kjii->kji", S), g["S_diag"]) < 1e-13
    assert rel(S.sum(axis=(2, 3)), g["S_sum"]) < 1e-12
    assert rel(Px[N], g["Phi_x_last"]) < 1e-13
    assert rel(Pu.sum(axis=(2, 3)), g["Phi_u_sum"]) < 1e-12
    for a, key in ((beta, "beta"), (beta_f, "beta_f"), (bo, "backoff"), (bof, "backoff_f")):
        assert rel(a, g[key]) < 1e-13
    # structure the reference guarantees: K[k, j] = 0 for j > k, beta zero above the diagonal
    for k in range(N):
        assert not K[k, k + 1:].any() and not beta[k, k + 1:].any()


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_jacobians_match_sympy(name):
    """oracle/models.py complex-step A, B and RK4 step vs SymPy differentiation (tests/golden/jac_*.npz)."""
    g = load(f"jac_{name}.npz")
    m = om.MODELS[name]()
    xs, us = jac_points(name)
    assert np.array_equal(xs, g["x"]) and np.array_equal(us, g["u"])
    for x, u, A, B, xp in zip(xs, us, g["A"], g["B"], g["xp"]):
        a, b = m.jacobians(x, u)
        assert np.abs(a - A).max() < 1e-12 and np.abs(b - B).max() < 1e-12
        assert np.abs(m.ddyn(x, u) - xp).max() < 1e-13


def test_model_constants_follow_reference():
    """Sizes and constraint layout of SURVEY.md section 8: g = [x_ub, u_ub, -x_lb, -u_lb], G = [I; -I]."""
    for name, (nx, nu, ni, nif) in dict(pendulum=(4, 1, 10, 8), quadrotor=(13, 4, 34, 26), rocket=(17, 4, 42, 34)).items():
        m = om.MODELS[name]()
        assert (m.nx, m.nu, m.nw, m.ni, m.ni_f) == (nx, nu, nx, ni, nif)
        nz = nx + nu
        assert np.array_equal(m.G, np.vstack([np.eye(nz), -np.eye(nz)]))
        assert np.array_equal(m.g[:nx], m.x_ub) and np.array_equal(m.g[nz:nz + nx], -m.x_lb)
    r = om.Rocket()
    assert r.x_ub[3] == 1.0 and r.x_ub[13] == 50.0 and r.u_ub[0] == 50.0
    assert abs(r.ddyn(om.ROCKET_X0, np.zeros(4))[10] - (-4.92936284)) < 1e-7   # strong roll torque at the fixed x0


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_oracle_qp_kkt(name):
    """Every QP fixture: the stored solution satisfies the KKT conditions of the reference's QP to <= 1e-9 and the
    oracle solver reproduces it."""
    g = load(f"qp_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    N, nx = cfg["N"], m.nx
    solver = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=1)
    n_check = 0
    for i in range(len(g["X"])):
        X, U = g["X"][i].T, g["U"][i].T
        solver.set_nominal(X, U)
        solver.update_jacobian()
        f = solver.fast
        if g["tight"][i]:
            # rebuild the tightened bounds exactly like update_tightening (fast_SLS_jit.py:560-569)
            tight = np.stack(f.g_list[:-1]) - g["backoff"][i]
            cmat = np.stack(f.c_list, axis=1)
            stage = np.vstack([-cmat, tight.T]).reshape(-1, order="F")
            f.ubg = np.concatenate([stage, f.gf_model - g["backoff_f"][i]])
        l, u = f.qp.bounds(f.lbg, f.ubg, X[:, 0] - g["x0"][i])
        y = f.qp.pack(g["dX"][i].T, g["dU"][i].T)
        lam = np.zeros(f.qp.m)
        stage_dual = np.full((N, nx + m.ni), np.nan)
        stage_dual[:, nx:] = g["mu"][i]
        # dynamics multipliers are not part of the fixture: recover them from stationarity by least squares
        ineq = np.concatenate([stage_dual.reshape(-1), g["mu_f"][i], np.full(nx, np.nan)])
        known = ~np.isnan(ineq)
        lam[known] = ineq[known]
        rhs = -(2 * f.qp.H @ y + f.q + f.qp.Acon[known].T @ lam[known])
        sol, *_ = np.linalg.lstsq(f.qp.Acon[~known].T, rhs, rcond=None)
        lam[~known] = sol
        res = oqp.kkt_residuals(f.qp.H, f.q, f.qp.Acon, l, u, y, lam)
        assert res["stat"] < 1e-8 and res["prim"] < 1e-9 and res["comp"] < 1e-8, res
        if i < 2:
            x2, y2, info = oqp.solve_qp(f.qp.H, f.q, f.qp.Acon, l, u)
            assert np.abs(x2 - y).max() < 1e-8 * max(1.0, np.abs(y).max())
            n_check += 1
    assert n_check >= 1


def test_q3_and_q4_quirks():
    """Q3: the convergence test uses the previous first-in-step QP across calls and skips the tube update;
    Q4: the tightened terminal bound is gf - backoff_f (not shifted by x_N)."""
    m, cfg = om.pendulum_config()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=1)
    X = np.zeros((m.nx, N + 1)); U = np.zeros((m.nu, N))
    X[0, :] = 0.3                                               # x_N != 0 so that Q4 is visible
    s.set_nominal(X, U)
    s.update_jacobian()
    inner = s.fast.solve(X[:, 0] - np.zeros(m.nx))
    assert np.allclose(s.fast.ubg[-m.ni_f:], m.gf - inner["backoff_f"])            # Q4: no "- Gf x_N"
    assert not np.allclose(s.fast.ubg[-m.ni_f:], m.gf - m.Gf @ X[:, N] - inner["backoff_f"])
    s.fast._prev_primal_vec = None
    X[0, :] = 0.0
    s.set_nominal(X, U)
    sol1 = s.solve(np.zeros(m.nx))
    assert sol1["success"] and sol1["K"] is not None            # first call: prev is None -> not converged
    s.reset_warm_start()
    sol2 = s.solve(np.zeros(m.nx))                               # identical QP -> |delta - prev| = 0 <= 1e-3 -> early return
    assert sol2["success"] and sol2["K"] is None
    assert not sol2["backoff_x"].any() and np.allclose(sol2["backoff"], N * 1e-5)


def test_pack_unpack_layout():
    """Stage-wise layout [x0; u0; ...; xN] of solver/ocp.py:43-74 in the oracle and in the product's host mirror."""
    import rnmpc_b200  # noqa: F401
    from rnmpc_b200.dyn.rocket import Rocket
    from rnmpc_b200.solver.ocp import OCP
    m = Rocket()
    o = OCP(15, np.eye(17), np.eye(4), m, np.eye(17))
    rng = np.random.default_rng(0)
    X, U = rng.random((17, 16)), rng.random((4, 15))
    y = o.pack_primal_nominal(X, U)
    q = oqp.QPData(15, 17, 4, np.eye(17), np.eye(4), np.eye(17), m.G, m.Gf)
    assert np.array_equal(y, q.pack(X, U))
    assert np.array_equal(y[:17], X[:, 0]) and np.array_equal(y[17:21], U[:, 0]) and np.array_equal(y[-17:], X[:, 15])
    X2, U2 = o.unpack_primal_nominal(y)
    assert np.array_equal(X, X2) and np.array_equal(U, U2)
    assert np.array_equal(o.Q_reg, np.eye(17)) and np.array_equal(o.R_reg, np.eye(4))    # ocp.py:15-26 defaults
    # product models expose the reference attribute set
    om_r = om.Rocket()
    assert np.array_equal(m.g, om_r.g) and np.array_equal(m.gf, om_r.gf) and np.array_equal(m.E, om_r.E)
    assert np.array_equal(np.concatenate(m.box()[:2]), np.concatenate([om_r.x_ub, om_r.x_lb]))
    m.G = np.ones_like(m.G)
    with pytest.raises(NotImplementedError):
        m.box()


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports every function include/rnmpc_b200.h declares (no compute calls)."""
    from rnmpc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "rnmpc_b200.h")).read()
    declared = set(re.findall(r"\b(rnmpc_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"rnmpc_handle", "rnmpc_config"}
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    nx, nu = ctypes.c_int(), ctypes.c_int()
    assert lib.rnmpc_model_dims(2, ctypes.byref(nx), ctypes.byref(nu)) == 0 and (nx.value, nu.value) == (17, 4)
    assert lib.rnmpc_model_dims(9, ctypes.byref(nx), ctypes.byref(nu)) != 0
    assert b"unknown model" in lib.rnmpc_last_error()


def test_bench_gpu_arm_does_not_import_the_oracle():
    """bench.py may touch oracle/ only in its CPU-baseline leg (cpu_loop, cpu_single_latency); the measured arm imports the
    product alone, and the product-side workload is the same as the oracle-side one the tests and the CPU arm use."""
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    allowed = {"cpu_loop", "cpu_single_latency"}
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            mod = node.module if isinstance(node, ast.ImportFrom) else None
            names = [a.name for a in node.names] if isinstance(node, ast.Import) else []
            if (mod and (mod.split(".")[0] in ("oracle", "tests"))) or any(n.split(".")[0] in ("oracle", "tests") for n in names):
                assert fn.name in allowed, f"bench.py:{node.lineno} imports {mod or names} inside {fn.name}()"
    for node in tree.body:      # module level
        if isinstance(node, (ast.Import, ast.ImportFrom)):
            mod = getattr(node, "module", None) or ""
            assert mod.split(".")[0] not in ("oracle", "tests")
    from oracle import models as om
    from tests.helpers import rocket_scenarios as rs_o
    from rnmpc_b200.workloads import ROCKET_X0, rocket_config, rocket_scenarios as rs_p
    m_o, c_o = om.rocket_config()
    m_p, c_p = rocket_config()
    assert np.array_equal(ROCKET_X0, om.ROCKET_X0) and np.array_equal(m_o.E, m_p.E)
    assert all(np.array_equal(np.asarray(c_o[k]), np.asarray(c_p[k])) for k in c_o)
    xu, xl, uu, ul, _, _ = m_p.box()
    assert np.array_equal(xu, m_o.x_ub) and np.array_equal(xl, m_o.x_lb) and np.array_equal(uu, m_o.u_ub) and np.array_equal(ul, m_o.u_lb)
    for args in ((64, 2025, 29, 0.0), (16, 7, 5, 0.002)):
        a, b = rs_o(*args), rs_p(*args)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_product_has_no_cpu_fallback():
    """Without a CUDA device the product must refuse to run instead of falling back to the oracle / NumPy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from tests.helpers import make_solver
    with pytest.raises(RuntimeError):
        make_solver("pendulum", 1)
    # and nothing under the product imports the oracle
    pkg = os.path.join(ROOT, "robust-nonlinear-mpc_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.parametrize("tag,name", [("rocket_rng123", "rocket"), ("pendulum", "pendulum"), ("quadrotor", "quadrotor")])
def test_cpp_cpu_port_reproduces_oracle_loops(tag, name):
    """oracle/cpu_port (C++ CPU port timed by bench.py as the CPU baseline) against the NumPy oracle's closed loops."""
    from oracle import cpu_port
    g = load(f"loop_{tag}.npz")
    m, cfg = om.CONFIGS[name]()
    steps = len(g["inputs"])
    W = g["W"][None] if g["W"].size else None
    st, inp, nfail = cpu_port.closed_loop(m, cfg, g["x0"][None], W, steps, g["X0"][None], g["U0"][None], threads=1)
    assert nfail == 0
    assert np.abs(st[0] - g["states"]).max() < 1e-6 and np.abs(inp[0] - g["inputs"]).max() < 1e-6


def test_driver_modules_import_and_mirror_the_reference_command_line(capsys, monkeypatch):
    """The package's counterparts of the reference's expe/main_*.py import without a GPU, expose generate() and accept the
    reference's command line (--run / --N); without --run they do not compute."""
    import importlib
    import inspect
    import runpy
    for name in ("rocket_robust", "pendulum_robust", "quadrotor_robust", "rocket_compare"):
        mod = importlib.import_module(f"rnmpc_b200.expe.main_{name}_closed_loop")
        assert callable(mod.generate) and isinstance(mod.folder_name, str)
        if name != "rocket_compare":
            assert list(inspect.signature(mod.generate).parameters)[:2] == ["N", "out_dir"]
        monkeypatch.setattr("sys.argv", [f"main_{name}_closed_loop.py"])
        sys.modules.pop(f"rnmpc_b200.expe.main_{name}_closed_loop", None)        # run it as a script, not as the imported module
        runpy.run_module(f"rnmpc_b200.expe.main_{name}_closed_loop", run_name="__main__")
        assert "plot" in capsys.readouterr().out


def test_soft_nmpc_oracle_satisfies_the_nlp_optimality_conditions():
    """oracle/soft.py (restatement of solver/nlp_soft_constraints.py:175-276 solved by SQP): its fixed point satisfies the first-order
    conditions of the NLP evaluated from scratch (dynamics defect, stationarity with the returned multipliers, slack stationarity,
    complementarity), with inactive and with active slacks; and it reproduces the committed fixture (tests/golden/soft_pendulum.npz)."""
    from oracle import soft as osoft
    m, cfg = om.pendulum_config()
    N, Q, R, Qf = cfg["N"], cfg["Q"], cfg["R"], cfg["Qf"]
    g = load("soft_pendulum.npz")
    X0, U0 = g["X0"].T, g["U0"].T
    x_out = g["x0"]
    assert x_out[0] > m.x_ub[0]                                     # the fixture starts outside the box
    for xs in (om.PENDULUM_X0, x_out):
        sol = osoft.soft_nlp_solve(m, N, Q, R, Qf, xs, X0, U0)
        assert sol["success"] and sol["last_step"] < 1e-9
        k = osoft.kkt_nlp(m, N, Q, R, Qf, xs, sol)
        assert k["defect"] < 1e-10 and k["stationarity"] < 1e-6 and k["sign"] < 1e-8 and k["complementarity"] < 1e-6, k
        # the optimal slacks are the violations of the trajectory, the cost the reference's objective
        gam, gam_f = osoft.slacks_of(m, sol["primal_x"], sol["primal_u"])
        assert np.abs(gam - sol["gamma"]).max() < 1e-9 and np.abs(gam_f - sol["gamma_f"]).max() < 1e-9
    assert rel(sol["primal_x"], g["Xn"][0].T) < 1e-9 and abs(sol["cost"] / g["cost"][0] - 1) < 1e-9
    assert g["gamma"][0].max() > 0.5                                 # active slacks in the fixture


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_restated_osqp_polish_lands_on_the_exact_oracle_solution(name):
    """The QP boundary against a restatement of OSQP's published algorithm (oracle/osqp_restated.py; the library itself
    is absent): on every QP fixture, with the bounds exactly as the reference hands them to OSQP (solver/qp_jit.py:370-386),
    ADMM + polish returns the fixtures' exact solution (what the CUDA interior point is tested against at 1e-6) to
    <= 1e-6 relative in the primal and <= 1e-6 in the inequality multipliers as soon as eps is small enough for the
    ADMM iterate to identify the active set -- the default 1e-3 on most fixtures, 1e-4 / 1e-5 on the others
    (profiles/r3f_osqp_restated_gap.txt) -- while the un-polished ADMM point stays within a few eps of it."""
    from oracle import osqp_restated as osr
    from oracle.osqp_gap import fixture_qps, gaps
    n_default = n = 0
    for i, tight, qp, P, q, A, l, u, y, mu in fixture_qps(name):
        if name == "rocket" and i in (0, 2):
            continue            # these two need 3000+ ADMM iterations: covered by the committed table, not by every CI run
        for eps in (1e-3, 1e-4, 1e-5):
            x, lam, info = osr.solve(P, q, A, l, u, eps_abs=eps, eps_rel=eps)
            assert info["status"] == "solved"
            dx, dmu = gaps(qp, x, lam, y, mu)
            Ax = A @ x
            wrong = np.sum((lam > 1e-9) & (u - Ax > 1e-7)) + np.sum((lam < -1e-9) & (Ax - l > 1e-7))
            dxa, _ = gaps(qp, info["x_admm"], info["y_admm"], y, mu)
            assert dxa < 0.1                       # the ADMM point itself: eps-accurate only (rocket: 8e-2 at eps 1e-3)
            if info["polished"] and wrong == 0:
                assert dx < 1e-6 and dmu < 1e-6, (name, i, eps, dx, dmu)
                n_default += eps == 1e-3
                n += 1
                break
        else:
            raise AssertionError(f"{name} fixture {i}: the polish never identified the active set")
    assert n >= 2 and n_default >= 1


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_models_match_the_reference_sources_executed_numerically(name):
    """tests/golden/refdyn_*.npz hold the outputs of the reference's OWN dyn/*.py statements (imported unmodified with a numeric
    CasADi stand-in, oracle/ref_models.py).  Against them: the oracle's transcription (oracle/models.py), the product's model header
    compiled for the host (csrc/rnmpc_models.cuh through the C++ port), and the constant tables.  Where the reference tree is
    present the fixture is also regenerated and must reproduce bit for bit."""
    import ctypes as C
    from oracle import cpu_port, ref_models as rm
    g = load(f"refdyn_{name}.npz")
    m = om.MODELS[name]()
    xs, us = g["x"], g["u"]
    scale = lambda a: np.maximum(1.0, np.abs(a))
    ode = np.array([np.asarray(m.ode(x, u), float) for x, u in zip(xs, us)])
    assert (np.abs(ode - g["ode"]) / scale(g["ode"])).max() < 1e-14
    for key, h in (("xp", 0.05), ("xp_h002", 0.02)):
        xp = np.array([m.ddyn(x, u, h) for x, u in zip(xs, us)])
        assert (np.abs(xp - g[key]) / scale(g[key])).max() < 1e-14
    assert (np.abs(xs + 0.05 * ode - g["xp_euler"]) / scale(g["xp_euler"])).max() < 1e-14      # dyn/model.py:23-26
    for key in ("G", "g", "Gf", "gf", "E"):
        assert np.array_equal(np.asarray(getattr(m, key), float).reshape(g[key].shape), g[key]), key
    # A, B: complex-step derivative of the reference's own RK4 map (what ca.jacobian returns, SCP_SLS_jit.py:199-218)
    for x, u, A, B in zip(xs, us, g["A"], g["B"]):
        a, b = m.jacobians(x, u)
        assert np.abs(a - A).max() < 1e-12 * max(1.0, np.abs(A).max()) and np.abs(b - B).max() < 1e-12 * max(1.0, np.abs(B).max())
    # the product's model header on the host
    lib = cpu_port.load()
    out = np.empty_like(xs)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    xs_c, us_c = np.ascontiguousarray(xs), np.ascontiguousarray(us)
    for what, h, key in ((0, 0.0, "ode"), (1, 0.05, "xp"), (1, 0.02, "xp_h002")):
        assert lib.rnmpc_cpu_model_eval(m.model_id, what, len(xs), vp(xs_c), vp(us_c), C.c_double(h), vp(out)) == 0
        assert (np.abs(out - g[key]) / scale(g[key])).max() < 1e-13, (name, key)
    if rm.available():
        ref = rm.load()[name]()
        again = np.array([rm.ddyn(ref, x, u, 0.05) for x, u in zip(xs, us)])
        assert np.array_equal(again, g["xp"])


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_fast_sls_glue_matches_the_reference_methods_executed_here(name):
    """tests/golden/refglue_*.npz hold two consecutive passes of the reference's OWN fast_SLS.evaluate_dual_eta / backward_solve /
    update_tightening / check_convergence_socp (unbound methods of the reference class on an attribute bag, oracle/ref_glue.py).
    The oracle's FastSLS must reproduce every output: eta, gains, beta, back-offs, backoff_x / backoff_u (Q7), the tightened
    upper bounds handed to the QP (Q4: un-shifted gf) and the Q3 decisions."""
    g = load(f"refglue_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=2)
    s.set_nominal(g["X"], g["U"])
    s.update_jacobian()
    f = s.fast
    assert rel(np.array(f.A), g["A"]) < 1e-15 and rel(np.array(f.c_list), g["c"]) < 1e-15
    for p in range(2):
        c = f.cur
        c["dual_mu"], c["dual_mu_f"] = g[f"p{p}_mu"].copy(), g[f"p{p}_mu_f"].copy()
        c["beta"], c["beta_f"] = g[f"p{p}_beta_in"].copy(), g[f"p{p}_beta_f_in"].copy()
        f.evaluate_dual_eta(); f.backward_solve(); f.update_tightening()
        for key in ("eta", "eta_f", "K", "beta", "beta_f", "backoff", "backoff_f", "backoff_x", "backoff_u"):
            assert rel(c[key], g[f"p{p}_{key}"]) < 1e-12, (name, p, key)
        assert rel(f.ubg, g[f"p{p}_ubg"]) < 1e-12
    f._prev_primal_vec = None
    q3 = []
    for pv in g["q3_inputs"]:
        f.cur["primal_vec"] = pv
        q3.append(f.check_convergence_socp())
    assert q3 == list(g["p1_q3"])
    assert list(g["p1_q3"][-2:]) == [True, False]        # 5e-4 inside, 2e-3 outside the hard-coded 1e-3


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_qp_assembly_matches_the_reference_qp_class_executed_here(name):
    """tests/golden/refqp_*.npz hold what the reference's OWN QP class (solver/qp_jit.py, imported unmodified over a numeric CasADi
    stand-in and a recording `osqp` stand-in, oracle/ref_qp.py) hands to OSQP for the nominal and the tightened QP of one fast-SLS
    pass, and the dictionaries its solve() returns with the restated OSQP as the engine.  The oracle's dense assembly must be the same
    QP entry for entry (P = 2H, A, q, l, u -- u up to the documented stage-0 policy), and its exact solution must be what the
    reference's _pack_solution returned (polished): primal, dual_mu (ni, N), dual_mu_f, cost."""
    g = load(f"refqp_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    N, nx, nu = cfg["N"], m.nx, m.nu
    nz = nx + nu
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=2)
    s.set_nominal(g["X"], g["U"])
    s.update_jacobian()
    f = s.fast
    dx0 = g["X"][:, 0] - g["x0"]
    assert "'polishing': True" in str(g["setup_settings"]) and "eps" not in str(g["setup_settings"])     # quirk Q1: OSQP defaults
    for p in range(2):
        assert np.array_equal(g[f"p{p}_P"], 2 * f.qp.H)
        assert np.array_equal(g[f"p{p}_A"], f.qp.Acon)
        assert np.array_equal(g[f"p{p}_q"], np.asarray(f.q, float).ravel())
        l, u = f.qp.bounds(f.lbg, f.ubg, dx0)
        policy = np.zeros(len(u), bool)
        policy[nx:2 * nx] = policy[nx + nz:nx + nz + nx] = True          # stage-0 state rows of G (dropped: they restate the pinned x(0))
        assert np.array_equal(g[f"p{p}_l"], l)
        assert np.array_equal(g[f"p{p}_u"][~policy], u[~policy])
        assert (u[policy] >= oqp.BIG).all() and (g[f"p{p}_u"][policy] < oqp.BIG).all()
        assert bool(g[f"p{p}_polished"])
        assert f.forward_solve(dx0)
        c = f.cur
        for key, tol in (("primal_vec", 1e-6), ("primal_x", 1e-6), ("primal_u", 1e-6), ("dual_mu", 1e-6), ("dual_mu_f", 1e-6)):
            assert c[key].shape == g[f"p{p}_{key}"].shape
            assert relf(c[key], g[f"p{p}_{key}"]) < tol, (name, p, key, relf(c[key], g[f"p{p}_{key}"]))
        # the restated OSQP's polish sits on one end of the +-1e-10 equality intervals, the oracle on their midpoint: x costates
        assert abs(c["cost_nominal"] - float(g[f"p{p}_cost"])) < 1e-4 * max(1.0, abs(float(g[f"p{p}_cost"])))
        f.evaluate_dual_eta(); f.backward_solve(); f.update_tightening()


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_warm_start_shift_and_layout_match_the_reference_methods_executed_here(name):
    """tests/golden/refshift_*.npz: outputs of the reference's OWN SCP_SLS.reset_warm_start and OCP.pack / unpack_primal_nominal
    (unbound methods of the unmodified classes, model = the reference's dyn/*.py; oracle/ref_scp.py).  Against them: the oracle's
    shift (bit for bit up to the RK4 round-off of X_N), its packing, and the product's host mirror solver/ocp.py."""
    import rnmpc_b200  # noqa: F401
    from rnmpc_b200.solver.ocp import OCP
    g = load(f"refshift_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=1)
    s.set_nominal(g["X_in"], g["U_in"])
    s.reset_warm_start()
    X, U = s.cur["primal_x"], s.cur["primal_u"]
    assert np.array_equal(X[:, :N], g["X"][:, :N]) and np.array_equal(U, g["U"])
    assert np.abs(X[:, N] - g["X"][:, N]).max() < 1e-14 * max(1.0, np.abs(g["X"]).max())
    assert list(g["calls"]) == ["reset_solver_to_zeros"]
    # layout (solver/ocp.py:43-74) executed by the reference: oracle packing and the product's mirror
    assert np.array_equal(s.fast.qp.pack(g["X"], g["U"]), g["primal_vec"])
    o = OCP(N, cfg["Q"], cfg["R"], m, cfg["Qf"])
    assert np.array_equal(np.asarray(o.pack_primal_nominal(g["X"], g["U"]), float).ravel(), g["primal_vec"])
    X2, U2 = o.unpack_primal_nominal(g["primal_vec"])
    assert np.array_equal(np.asarray(X2, float), g["unpacked_X"]) and np.array_equal(np.asarray(U2, float), g["unpacked_U"])
    # the shifted inequality bounds the reference stashes (applied only with warm_start_ubg_shift): columns 1.. then the last repeated
    nx, ni = m.nx, m.ni
    stage = g["ubg_in"][:N * (nx + ni)].reshape((nx + ni, N), order="F")[nx:]
    want = np.concatenate([stage[:, 1:], stage[:, -1:]], axis=1).reshape(-1, order="F")
    assert np.array_equal(want, g["pending"])


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_scp_control_flow_matches_the_reference_solve_executed_here(name):
    """tests/golden/refsolve_*.npz: two MPC steps run by the reference's OWN SCP_SLS.solve / socp_step / check_convergence_scp /
    post_processing_solution / reset_warm_start (unbound methods of the unmodified class at the driver's rti / fast_sls_rti_steps;
    linearisation and the fast_SLS object supplied by the oracle, oracle/ref_scp.py).  The oracle's own SCPSLS.solve /
    reset_warm_start must walk the same path: every numeric key, `iterations`, `success`."""
    g = load(f"refsolve_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=int(g["rti"]),
                    fast_sls_rti_steps=int(g["sls_steps"]))
    s.set_nominal(g["X_in"], g["U_in"])
    for i, x0 in enumerate(g["x0s"]):
        if i > 0:
            s.reset_warm_start()
        sol = s.solve(x0)
        assert bool(sol["success"]) == bool(g[f"s{i}_success"]) and sol["iterations"] == int(g[f"s{i}_iterations"])
        for key in ("primal_x", "primal_u", "primal_vec", "delta_vec", "dual_mu", "dual_mu_f", "backoff", "backoff_x", "backoff_u", "dual_eta_f"):
            assert relf(sol[key], g[f"s{i}_{key}"].reshape(np.shape(sol[key]))) < 1e-9, (name, i, key)
        for key in ("cost", "cost_QP", "primal_infeasibility"):
            want = float(g[f"s{i}_{key}"].ravel()[0])
            assert abs(float(sol[key]) - want) < 1e-9 * max(1.0, abs(want)), (name, i, key, float(sol[key]), want)
        assert int(sol["SOCP_steps"]) == int(g[f"s{i}_SOCP_steps"])
        # the reference's result dictionary has the keys the drivers read
        assert {"primal_x", "primal_u", "backoff_x", "backoff_u", "t_qp_ms", "t_backward_ms", "success", "iterations", "it_data"} <= set(g[f"s{i}_keys"])


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_fast_sls_iteration_matches_the_reference_solve_executed_here(name):
    """tests/golden/reffast_*.npz: two consecutive fast_SLS.solve calls run END TO END by the reference's own statements (control flow,
    QP class, Numba kernels; only OSQP's ADMM is the restated one, polished; oracle/ref_fast_sls.py) -- pendulum in iterate-until-Q3
    mode, quadrotor with 2 and rocket with 1 RTI step.  The oracle's FastSLS.solve must return the same dictionaries: iteration counts
    and success exactly (incl. the Q3 early return of the second pendulum call), everything numeric to 1e-6."""
    g = load(f"reffast_{name}.npz")
    assert bool(g["all_polished"])
    m, cfg = om.CONFIGS[name]()
    N = cfg["N"]
    s = oscp.SCPSLS(N, cfg["Q"], cfg["R"], m, cfg["Qf"], cfg["Q_reg"], cfg["R_reg"], cfg["Q_reg_f"], rti=1, fast_sls_rti_steps=int(g["rti_steps"]) or None)
    s.set_nominal(g["X"], g["U"])
    s.update_jacobian()
    for i, dx0 in enumerate(g["dx0s"]):
        s.fast.cur["iteration_number"] = 0
        sol = s.fast.solve(dx0)
        assert bool(sol["success"]) == bool(g[f"s{i}_success"]) and int(sol["iteration_number"]) == int(g[f"s{i}_iteration_number"])
        for key in ("primal_x", "primal_u", "primal_vec", "dual_mu", "dual_mu_f", "eta_f", "K", "beta", "beta_f", "backoff", "backoff_f",
                    "backoff_x", "backoff_u"):
            assert relf(sol[key], g[f"s{i}_{key}"].reshape(np.shape(sol[key]))) < 1e-6, (name, i, key, relf(sol[key], g[f"s{i}_{key}"].reshape(np.shape(sol[key]))))
        assert relf(sol["eta"], g[f"s{i}_eta"]) < 2e-6
    if name == "pendulum":
        assert int(g["s0_iteration_number"]) >= 1 and int(g["s1_iteration_number"]) == 0        # Q3: the history carries over


@pytest.mark.parametrize("name", ["pendulum", "quadrotor", "rocket"])
def test_closed_loop_matches_the_reference_stack_executed_here(name):
    """tests/golden/refloop_*.npz: closed loops (pendulum 6, quadrotor 3, rocket 3 MPC steps, disturbed) run by the reference's OWN MPC
    stack -- SCP_SLS.solve / reset_warm_start over its own fast_SLS.solve, QP class, Numba kernels and Model.ddyn; only the values of the
    linearisation and OSQP's ADMM iteration are supplied (oracle/ref_scp.py).  The oracle's closed loop must follow: closed-loop states
    1e-5 (north star), nominal trajectories and back-offs 1e-6."""
    g = load(f"refloop_{name}.npz")
    m, cfg = om.CONFIGS[name]()
    cfg = dict(cfg, rti=int(g["rti"]), fast_sls_rti_steps=int(g["sls_steps"]))
    steps = len(g["W"])
    out = oscp.closed_loop(m, cfg, g["x0"], g["W"], steps, X0=g["X_in"], U0=g["U_in"])
    assert all(out["success"]) and len(out["X"]) == steps + 1
    assert relf(np.array(out["X"]), g["states"]) < 1e-5
    for i in range(steps):
        assert relf(out["Xn"][i], g["primal_x"][i]) < 1e-6 and relf(out["Un"][i], g["primal_u"][i]) < 1e-6
        assert relf(out["backoff_x"][i], g["backoff_x"][i]) < 1e-6 and relf(out["backoff_u"][i], g["backoff_u"][i]) < 1e-6
