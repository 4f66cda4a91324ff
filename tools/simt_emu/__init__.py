"""ctypes front end of the SIMT emulator build of the product's kernels (tools/simt_emu/librnmpc_emu.so).

TEST INFRASTRUCTURE ONLY: the product (robust-nonlinear-mpc_b200/) never imports this package.  It lets the CPU test-suite run
the unmodified kernel sources of csrc/ (one fiber per CUDA thread, warp collectives emulated) against the oracle in the
build container, where no GPU exists; the compiled sm_100a kernels are checked by the `-m gpu` tests through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "librnmpc_emu.so")
_lib = None
MODEL_IDS = {"pendulum": 0, "quadrotor": 1, "rocket": 2}


def build():
    res = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + res.stdout[-3000:] + res.stderr[-3000:])


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
    return _lib


def set_order(order):
    """Scheduling order of the emulated threads between barriers: 0 ascending, 1 descending, 2 pseudo-random.  Results of a
    correctly synchronised kernel do not depend on it."""
    load().emu_set_order(int(order))


def set_qp_kernel(k):
    """Kernel run by qp_mma(): 0 = qp_mma_kernel (one CTA of four warps per instance), 1 = qp_warp_kernel (one warp per instance),
    2 = qp_tm_kernel (one warp per instance, iterate in tensor memory)."""
    load().emu_set_qp_kernel(int(k))


def _cfg(name, m, cfg):
    from rnmpc_b200._lib import Config
    keep = dict(Q=cfg["Q"], R=cfg["R"], Qf=cfg["Qf"], Q_reg=cfg["Q_reg"], R_reg=cfg["R_reg"], Q_reg_f=cfg["Q_reg_f"], E=m.E,
                x_ub=m.x_ub, x_lb=m.x_lb, u_ub=m.u_ub, u_lb=m.u_lb, xf_ub=m.x_ub, xf_lb=m.x_lb)
    keep = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)) for k, v in keep.items()}
    c = Config(model_id=MODEL_IDS[name], N=int(cfg["N"]), rti=int(cfg["rti"]), sls_steps=int(cfg["fast_sls_rti_steps"] or 0), h=0.05,
               qp_max_iter=0, qp_tol_mu=0.0, qp_tol_res=0.0, **{k: v.ctypes.data_as(C.POINTER(C.c_double)) for k, v in keep.items()})
    return c, keep


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def pad_rows(A):
    """[..., nx, nx] -> [..., nx, ld] with ld = nx rounded up to even (the engine's resident layout of A_k)."""
    nx = A.shape[-1]
    ld = (nx + 1) & ~1
    out = np.zeros(A.shape[:-1] + (ld,))
    out[..., :nx] = A
    return np.ascontiguousarray(out)


def qp_mma(name, m, cfg, A, Bm, c, X, U, x0, backoff=None, backoff_f=None, max_iter=40, tol_mu=1e-14, tol_res=1e-9, mu0=-0.3):
    """qp_mma_kernel on the emulator.  A [B,N,nx,nx], Bm [B,N,nx,nu], c [B,N,nx], X [B,N+1,nx], U [B,N,nu], x0 [B,nx]."""
    lib = load()
    conf, keep = _cfg(name, m, cfg)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    Ap, Bm, c, X, U, x0 = pad_rows(f(A)), f(Bm), f(c), f(X), f(U), f(x0)
    B, N, nx, nu = X.shape[0], int(cfg["N"]), m.nx, m.nu
    ni, nif = 2 * (nx + nu), 2 * nx
    tight = backoff is not None
    bo = f(backoff) if tight else None
    bof = f(backoff_f) if tight else None
    out = dict(dX=np.zeros((B, N + 1, nx)), dU=np.zeros((B, N, nu)), mu=np.zeros((B, N, ni)), mu_f=np.zeros((B, nif)),
               cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32), status=np.zeros(B, dtype=np.int32))
    lib.emu_qp_mma.restype = C.c_int
    lib.emu_qp_mma.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double] + [C.c_void_p] * 7
    rc = lib.emu_qp_mma(C.byref(conf), B, int(tight), _p(Ap), _p(Bm), _p(c), _p(X), _p(U), _p(x0), _p(bo), _p(bof), None, 0,
                        int(max_iter), float(tol_mu), float(tol_res), float(mu0), _p(out["dX"]), _p(out["dU"]), _p(out["mu"]),
                        _p(out["mu_f"]), _p(out["cost"]), _p(out["iters"]), _p(out["status"]))
    if rc != 0:
        raise RuntimeError(f"emu_qp_mma failed: {rc}")
    return out


def sls_mma(name, m, cfg, A, Bm, eta=None, eta_f=None, mu=None, mu_f=None, beta_in=None, beta_f_in=None, want_S=True, want_phi=True, warps=3):
    """sls_mma_kernel on the emulator (one instance batch axis optional: A [B,N,nx,nx] or [N,nx,nx])."""
    lib = load()
    conf, keep = _cfg(name, m, cfg)
    f = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    A, Bm = f(A), f(Bm)
    if A.ndim == 3:
        A, Bm = A[None], Bm[None]
        eta, eta_f, mu, mu_f, beta_in, beta_f_in = (None if t is None else np.asarray(t)[None] for t in (eta, eta_f, mu, mu_f, beta_in, beta_f_in))
    Ap = pad_rows(A)
    B, N, nx, nu = A.shape[0], int(cfg["N"]), m.nx, m.nu
    ni, nif = 2 * (nx + nu), 2 * nx
    eta, eta_f, mu, mu_f, beta_in, beta_f_in = (f(t) for t in (eta, eta_f, mu, mu_f, beta_in, beta_f_in))
    out = dict(K=np.zeros((B, N, N + 1, nu, nx)), beta=np.zeros((B, N, N, ni)), beta_f=np.zeros((B, N + 1, nif)))
    out["S"] = np.zeros((B, N + 1, N + 1, nx, nx)) if want_S else None
    out["Phi_x"] = np.zeros((B, N + 1, N + 1, nx, nx)) if want_phi else None
    out["Phi_u"] = np.zeros((B, N, N + 1, nu, nx)) if want_phi else None
    lib.emu_sls_mma.restype = C.c_int
    lib.emu_sls_mma.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 14
    rc = lib.emu_sls_mma(C.byref(conf), B, int(warps), _p(Ap), _p(Bm), _p(eta), _p(eta_f), _p(mu), _p(mu_f), _p(beta_in), _p(beta_f_in),
                         _p(out["K"]), _p(out["S"]), _p(out["Phi_x"]), _p(out["Phi_u"]), _p(out["beta"]), _p(out["beta_f"]))
    if rc != 0:
        raise RuntimeError(f"emu_sls_mma failed: {rc}")
    return out


def qp_soft(name, m, cfg, A, Bm, c, X, U, x0, rho=1e6, rho1=1e6, max_iter=60, tol_mu=1e-16):
    """The soft-constrained comparator's QP (qp_warp_kernel, SOFT) on the emulator; shapes as qp_mma()."""
    lib = load()
    conf, keep = _cfg(name, m, cfg)
    f = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    Ap, Bm, c, X, U, x0 = pad_rows(f(A)), f(Bm), f(c), f(X), f(U), f(x0)
    B, N, nx, nu = X.shape[0], int(cfg["N"]), m.nx, m.nu
    ni, nif = 2 * (nx + nu), 2 * nx
    out = dict(dX=np.zeros((B, N + 1, nx)), dU=np.zeros((B, N, nu)), mu=np.zeros((B, N, ni)), mu_f=np.zeros((B, nif)),
               gamma=np.zeros((B, N, ni)), gamma_f=np.zeros((B, nif)), cost=np.zeros(B), iters=np.zeros(B, dtype=np.int32),
               status=np.zeros(B, dtype=np.int32))
    lib.emu_qp_soft.restype = C.c_int
    lib.emu_qp_soft.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_int, C.c_double] + [C.c_void_p] * 9
    rc = lib.emu_qp_soft(C.byref(conf), B, _p(Ap), _p(Bm), _p(c), _p(X), _p(U), _p(x0), float(rho), float(rho1), int(max_iter), float(tol_mu),
                         _p(out["dX"]), _p(out["dU"]), _p(out["mu"]), _p(out["mu_f"]), _p(out["gamma"]), _p(out["gamma_f"]), _p(out["cost"]),
                         _p(out["iters"]), _p(out["status"]))
    if rc != 0:
        raise RuntimeError(f"emu_qp_soft failed: {rc}")
    return out
