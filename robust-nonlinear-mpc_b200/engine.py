"""Batched device engine: thin Python layer over the C ABI; torch supplies device memory and streams only."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

MODEL_IDS = {"pendulum": 0, "quadrotor": 1, "rocket": 2}


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BatchSolver:
    """B independent SCP-SLS MPC instances resident on one GPU (one handle of the C library).

    Parameters mirror ``SCP_SLS(N, Q, R, m, Qf, Q_reg, R_reg, Q_reg_f, rti=, fast_sls_rti_steps=)`` of the
    reference (solver/SCP_SLS_jit.py:24-44); ``m`` needs ``nx, nu, E, g, gf`` and a ``model_id``/``name``.
    """

    def __init__(self, model_name, N, Q, R, Qf, Q_reg, R_reg, Q_reg_f, E, x_ub, x_lb, u_ub, u_lb, xf_ub=None, xf_lb=None,
                 rti=1, sls_steps=1, capacity=1, device=0, h=0.05, qp_max_iter=0, qp_tol_mu=0.0, qp_tol_res=0.0):
        self.lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("rnmpc_b200 needs a CUDA device (no CPU fallback)")
        self.model_name = model_name
        self.model_id = MODEL_IDS[model_name]
        nx, nu = C.c_int(), C.c_int()
        _lib.check(self.lib.rnmpc_model_dims(self.model_id, C.byref(nx), C.byref(nu)))
        self.nx, self.nu, self.N = nx.value, nu.value, int(N)
        self.nz, self.ni, self.ni_f = self.nx + self.nu, 2 * (self.nx + self.nu), 2 * self.nx
        self.capacity, self.device_index = int(capacity), int(device)
        self.device = torch.device("cuda", self.device_index)
        self.rti, self.sls_steps = int(rti), int(sls_steps)
        xf_ub = x_ub if xf_ub is None else xf_ub
        xf_lb = x_lb if xf_lb is None else xf_lb

        def mat(a, n):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
            if a.shape != (n, n):
                raise ValueError(f"expected ({n},{n}) matrix, got {a.shape}")
            return a

        def vec(a, n):
            a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
            if a.shape != (n,):
                raise ValueError(f"expected length-{n} vector, got {a.shape}")
            return a

        keep = dict(Q=mat(Q, self.nx), R=mat(R, self.nu), Qf=mat(Qf, self.nx), Q_reg=mat(Q_reg, self.nx),
                    R_reg=mat(R_reg, self.nu), Q_reg_f=mat(Q_reg_f, self.nx), E=mat(E, self.nx),
                    x_ub=vec(x_ub, self.nx), x_lb=vec(x_lb, self.nx), u_ub=vec(u_ub, self.nu), u_lb=vec(u_lb, self.nu),
                    xf_ub=vec(xf_ub, self.nx), xf_lb=vec(xf_lb, self.nx))
        self._host = keep
        cfg = _lib.Config(model_id=self.model_id, N=self.N, rti=self.rti, sls_steps=self.sls_steps, h=float(h),
                          qp_max_iter=int(qp_max_iter), qp_tol_mu=float(qp_tol_mu), qp_tol_res=float(qp_tol_res),
                          **{k: _dp(v) for k, v in keep.items()})
        self._h = C.c_void_p()
        _lib.check(self.lib.rnmpc_create(C.byref(self._h), C.byref(cfg), self.capacity, self.device_index))
        self._views = {}

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.rnmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dev(self, a, dtype=torch.float64):
        if isinstance(a, torch.Tensor):
            return a.to(device=self.device, dtype=dtype).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=self.device)

    def empty(self, *shape, dtype=torch.float64):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float64):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    _SHAPES = {
        "X": lambda s: (s.N + 1, s.nx), "U": lambda s: (s.N, s.nu), "dX": lambda s: (s.N + 1, s.nx), "dU": lambda s: (s.N, s.nu),
        "mu": lambda s: (s.N, s.ni), "mu_f": lambda s: (s.ni_f,), "K": lambda s: (s.N, s.N + 1, s.nu, s.nx),
        "beta": lambda s: (s.N, s.N, s.ni), "beta_f": lambda s: (s.N + 1, s.ni_f), "backoff": lambda s: (s.N, s.ni),
        "backoff_f": lambda s: (s.ni_f,), "A": lambda s: (s.N, s.nx, (s.nx + 1) & ~1), "B": lambda s: (s.N, s.nx, s.nu),
        "c": lambda s: (s.N, s.nx), "cost": lambda s: (), "Phi_x": lambda s: (s.N + 1, s.N + 1, s.nx, s.nx),
        "Phi_u": lambda s: (s.N, s.N + 1, s.nu, s.nx), "flags": lambda s: (), "status": lambda s: (), "qp_iters": lambda s: (),
        "q3_margin": lambda s: (), "iters_tmp": lambda s: (), "ws_valid": lambda s: (),      # last QP launch's iteration count per instance; workspace-valid flag
    }
    _INT = {"flags", "status", "qp_iters", "iters_tmp", "ws_valid"}

    def buffer(self, name, B=None):
        """Resident result buffer of the handle as a NumPy array (device -> host copy of the first B instances)."""
        B = self.capacity if B is None else B
        shape = (B,) + tuple(self._SHAPES[name](self))
        dt = np.int32 if name in self._INT else np.float64
        out = np.empty(shape, dtype=dt)
        _lib.check(self.lib.rnmpc_copy_buffer(self._h, name.encode(), C.c_void_p(out.ctypes.data), out.nbytes))
        return out

    def buffer_view(self, name, B=None):
        """Zero-copy view of a resident device buffer (object exposing __cuda_array_interface__; wrap with torch.as_tensor)."""
        B = self.capacity if B is None else B
        ptr, nbytes = C.c_void_p(), C.c_longlong()
        _lib.check(self.lib.rnmpc_get_buffer(self._h, name.encode(), C.byref(ptr), C.byref(nbytes)))
        shape = (B,) + tuple(self._SHAPES[name](self))
        typestr = "<i4" if name in self._INT else "<f8"

        class _View:
            __cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr.value, False), version=3, strides=None)
            _owner = self
        return _View()

    # -- stateless kernels -------------------------------------------------------------------------------------------
    def ddyn(self, x, u, w=None, h=0.0):
        x, u = self.dev(x).reshape(-1, self.nx), self.dev(u).reshape(-1, self.nu)
        w = None if w is None else self.dev(w).reshape(-1, self.nx)
        xp = torch.empty_like(x)
        _lib.check(self.lib.rnmpc_ddyn(self._h, x.shape[0], _ptr(x), _ptr(u), _ptr(w), float(h), _ptr(xp), self._stream()))
        return xp

    def linearize(self, X, U):
        """X [B, N+1, nx], U [B, N, nu] -> A [B, N, nx, nx], B [B, N, nx, nu], c [B, N, nx]."""
        X, U = self.dev(X), self.dev(U)
        B = X.shape[0]
        A, Bm, c = self.empty(B, self.N, self.nx, self.nx), self.empty(B, self.N, self.nx, self.nu), self.empty(B, self.N, self.nx)
        _lib.check(self.lib.rnmpc_linearize(self._h, B, _ptr(X), _ptr(U), _ptr(A), _ptr(Bm), _ptr(c), self._stream()))
        return A, Bm, c

    def sls(self, A, Bm, eta, eta_f, want_S=False, want_phi=False):
        A, Bm, eta, eta_f = self.dev(A), self.dev(Bm), self.dev(eta), self.dev(eta_f)
        B, N, nx, nu = A.shape[0], self.N, self.nx, self.nu
        out = dict(K=self.zeros(B, N, N + 1, nu, nx), beta=self.zeros(B, N, N, self.ni), beta_f=self.zeros(B, N + 1, self.ni_f),
                   backoff=self.zeros(B, N, self.ni), backoff_f=self.zeros(B, self.ni_f))
        out["S"] = self.zeros(B, N + 1, N + 1, nx, nx) if want_S else None
        out["Phi_x"] = self.zeros(B, N + 1, N + 1, nx, nx) if want_phi else None
        out["Phi_u"] = self.zeros(B, N, N + 1, nu, nx) if want_phi else None
        _lib.check(self.lib.rnmpc_sls(self._h, B, _ptr(A), _ptr(Bm), _ptr(eta), _ptr(eta_f), _ptr(out["K"]), _ptr(out["S"]),
                                      _ptr(out["Phi_x"]), _ptr(out["Phi_u"]), _ptr(out["beta"]), _ptr(out["beta_f"]),
                                      _ptr(out["backoff"]), _ptr(out["backoff_f"]), self._stream()))
        return out

    def qp(self, A, Bm, c, X, U, x0, backoff=None, backoff_f=None):
        A, Bm, c, X, U, x0 = (self.dev(t) for t in (A, Bm, c, X, U, x0))
        B, N = A.shape[0], self.N
        tight = backoff is not None
        bo = self.dev(backoff) if tight else None
        bof = self.dev(backoff_f) if tight else None
        out = dict(dX=self.zeros(B, N + 1, self.nx), dU=self.zeros(B, N, self.nu), mu=self.zeros(B, N, self.ni),
                   mu_f=self.zeros(B, self.ni_f), cost=self.zeros(B), iters=self.zeros(B, dtype=torch.int32),
                   status=self.zeros(B, dtype=torch.int32))
        _lib.check(self.lib.rnmpc_qp(self._h, B, int(tight), _ptr(A), _ptr(Bm), _ptr(c), _ptr(X), _ptr(U), _ptr(x0), _ptr(bo),
                                     _ptr(bof), _ptr(out["dX"]), _ptr(out["dU"]), _ptr(out["mu"]), _ptr(out["mu_f"]),
                                     _ptr(out["cost"]), _ptr(out["iters"]), _ptr(out["status"]), self._stream()))
        return out

    def qp_general(self, A, Bm, c, hi, lo, q, x0pin):
        """QP with explicit bounds / linear cost in entry layout [B, N+1, nz] (rnmpc_qp_general)."""
        A, Bm, c, hi, lo, x0pin = (self.dev(t) for t in (A, Bm, c, hi, lo, x0pin))
        q = None if q is None else self.dev(q)
        B, N = A.shape[0], self.N
        out = dict(dX=self.zeros(B, N + 1, self.nx), dU=self.zeros(B, N, self.nu), mu=self.zeros(B, N, self.ni),
                   mu_f=self.zeros(B, self.ni_f), cost=self.zeros(B), iters=self.zeros(B, dtype=torch.int32),
                   status=self.zeros(B, dtype=torch.int32))
        _lib.check(self.lib.rnmpc_qp_general(self._h, B, _ptr(A), _ptr(Bm), _ptr(c), _ptr(hi), _ptr(lo), _ptr(q), _ptr(x0pin),
                                             _ptr(out["dX"]), _ptr(out["dU"]), _ptr(out["mu"]), _ptr(out["mu_f"]), _ptr(out["cost"]),
                                             _ptr(out["iters"]), _ptr(out["status"]), self._stream()))
        return out

    def sls_duals(self, A, Bm, mu, mu_f, beta=None, beta_f=None, want_phi=True):
        """fast-SLS pass from QP duals (rnmpc_sls_duals); beta/beta_f None = epsilon (state after initialize_backoff)."""
        A, Bm, mu, mu_f = (self.dev(t) for t in (A, Bm, mu, mu_f))
        B, N, nx, nu = A.shape[0], self.N, self.nx, self.nu
        bi = None if beta is None else self.dev(beta)
        bfi = None if beta_f is None else self.dev(beta_f)
        out = dict(K=self.zeros(B, N, N + 1, nu, nx), beta=self.zeros(B, N, N, self.ni), beta_f=self.zeros(B, N + 1, self.ni_f),
                   backoff=self.zeros(B, N, self.ni), backoff_f=self.zeros(B, self.ni_f), S=self.zeros(B, N + 1, N + 1, nx, nx))
        out["Phi_x"] = self.zeros(B, N + 1, N + 1, nx, nx) if want_phi else None
        out["Phi_u"] = self.zeros(B, N, N + 1, nu, nx) if want_phi else None
        _lib.check(self.lib.rnmpc_sls_duals(self._h, B, _ptr(A), _ptr(Bm), _ptr(mu), _ptr(mu_f), _ptr(bi), _ptr(bfi), _ptr(out["K"]),
                                            _ptr(out["S"]), _ptr(out["Phi_x"]), _ptr(out["Phi_u"]), _ptr(out["beta"]), _ptr(out["beta_f"]),
                                            _ptr(out["backoff"]), _ptr(out["backoff_f"]), self._stream()))
        return out

    # -- stateful solver ---------------------------------------------------------------------------------------------
    def costates(self, B=None):
        """Multipliers of the dynamics rows of the last nominal QP of the resident solve, [B, N, nx] (rnmpc_costates)."""
        B = self.capacity if B is None else int(B)
        lam = self.empty(B, self.N, self.nx)
        _lib.check(self.lib.rnmpc_costates(self._h, B, _ptr(lam), self._stream()))
        return lam

    def soft_nmpc(self, x0, X, U, rho=1e6, rho1=1e6, max_iter=60, tol=1e-9):
        """Soft-constrained NMPC comparator (rnmpc_soft_nmpc): SQP from the guess X [B,N+1,nx], U [B,N,nu] -> dict(X, U, gamma, gamma_f,
        sqp_iters, status) of device tensors."""
        x0 = self.dev(x0).reshape(-1, self.nx)
        B = x0.shape[0]
        X, U = self.dev(X).reshape(B, self.N + 1, self.nx).clone(), self.dev(U).reshape(B, self.N, self.nu).clone()
        out = dict(X=X, U=U, gamma=self.zeros(B, self.N, self.ni), gamma_f=self.zeros(B, self.ni_f),
                   sqp_iters=self.zeros(B, dtype=torch.int32), status=self.zeros(B, dtype=torch.int32))
        _lib.check(self.lib.rnmpc_soft_nmpc(self._h, B, _ptr(x0), _ptr(X), _ptr(U), float(rho), float(rho1), int(max_iter), float(tol),
                                            _ptr(out["gamma"]), _ptr(out["gamma_f"]), _ptr(out["sqp_iters"]), _ptr(out["status"]), self._stream()))
        return out

    def set_nominal(self, X, U):
        X, U = self.dev(X), self.dev(U)
        _lib.check(self.lib.rnmpc_set_nominal(self._h, X.shape[0], _ptr(X), _ptr(U), self._stream()))

    def init_nominal(self, x0, iters=10):
        x0 = self.dev(x0).reshape(-1, self.nx)
        _lib.check(self.lib.rnmpc_init_nominal(self._h, x0.shape[0], _ptr(x0), int(iters), self._stream()))

    def solve(self, x0):
        x0 = self.dev(x0).reshape(-1, self.nx)
        _lib.check(self.lib.rnmpc_solve(self._h, x0.shape[0], _ptr(x0), self._stream()))

    def reset_warm_start(self, B):
        _lib.check(self.lib.rnmpc_reset_warm_start(self._h, int(B), self._stream()))

    def plant_step(self, x, w=None):
        """x [B, nx] device tensor, updated in place: x <- ddyn(x, U[:, 0]) + E w."""
        w = None if w is None else self.dev(w).reshape(-1, self.nx)
        _lib.check(self.lib.rnmpc_plant_step(self._h, x.shape[0], _ptr(x), _ptr(w), None, self._stream()))
        return x

    def enable_phi(self, enable=True):
        _lib.check(self.lib.rnmpc_enable_phi(self._h, int(bool(enable))))

    def set_option(self, name, value):
        _lib.check(self.lib.rnmpc_set_option(self._h, name.encode(), float(value)))

    def enable_timing(self, enable=True):
        """True / 1: events around the phases of the last solve (timings()); 2: an event pair around every launch of the hot
        kernels, summed by kernel_time() (no synchronisation in between)."""
        _lib.check(self.lib.rnmpc_enable_timing(self._h, int(enable)))

    def timings(self):
        out = (C.c_double * 3)()
        _lib.check(self.lib.rnmpc_get_timings(self._h, out))
        return dict(t_jac_ms=out[0], t_qp_ms=out[1], t_backward_ms=out[2])

    def kernel_time(self, kernel):
        """(total ms, launches) of every launch of `kernel` ("linearize" | "qp" | "sls") since the last call; needs enable_timing(2)."""
        tot, cnt = C.c_double(), C.c_longlong()
        _lib.check(self.lib.rnmpc_kernel_time(self._h, kernel.encode(), C.byref(tot), C.byref(cnt)))
        return float(tot.value), int(cnt.value)

    def ipm_iterations(self):
        """Interior-point iterations of all instances and QP launches of this handle since creation (device counter)."""
        v = np.zeros(1, dtype=np.uint64)
        _lib.check(self.lib.rnmpc_copy_buffer(self._h, b"iters_sum", C.c_void_p(v.ctypes.data), 8))
        return int(v[0])

    def live_solves(self):
        """Instances actually solved (not frozen by sticky_fail), summed over all solve() calls of this handle (device counter)."""
        v = np.zeros(1, dtype=np.uint64)
        _lib.check(self.lib.rnmpc_copy_buffer(self._h, b"live_sum", C.c_void_p(v.ctypes.data), 8))
        return int(v[0])

    def launch_count(self):
        return int(self.lib.rnmpc_launch_count(self._h))

    # -- host-buffer entry points ------------------------------------------------------------------------------------
    def solve_host(self, x0):
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1, self.nx))
        B, N = x0.shape[0], self.N
        X, U = np.empty((B, N + 1, self.nx)), np.empty((B, N, self.nu))
        bo, bof = np.empty((B, N, self.ni)), np.empty((B, self.ni_f))
        fl, st = np.empty(B, dtype=np.int32), np.empty(B, dtype=np.int32)
        vp = lambda a: C.c_void_p(a.ctypes.data)
        _lib.check(self.lib.rnmpc_solve_host(self._h, B, vp(x0), vp(X), vp(U), vp(bo), vp(bof), vp(fl), vp(st)))
        return dict(X=X, U=U, backoff=bo, backoff_f=bof, flags=fl, status=st)

    def closed_loop_host(self, x0, W, steps, X0=None, U0=None, init_iters=10):
        x0 = np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(-1, self.nx))
        B = x0.shape[0]
        vp = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
        Wc = None if W is None else np.ascontiguousarray(np.asarray(W, dtype=np.float64).reshape(B, steps, self.nx))
        X0c = None if X0 is None else np.ascontiguousarray(np.asarray(X0, dtype=np.float64).reshape(B, self.N + 1, self.nx))
        U0c = None if U0 is None else np.ascontiguousarray(np.asarray(U0, dtype=np.float64).reshape(B, self.N, self.nu))
        states, inputs = np.empty((B, steps + 1, self.nx)), np.empty((B, steps, self.nu))
        nfail = C.c_longlong(0)
        _lib.check(self.lib.rnmpc_closed_loop_host(self._h, B, int(steps), vp(x0), vp(Wc), vp(X0c), vp(U0c), int(init_iters),
                                                   vp(states), vp(inputs), C.cast(C.byref(nfail), C.c_void_p)))
        return dict(states=states, inputs=inputs, n_failed_solves=int(nfail.value))


def fp64_peak_tflops(device=0, pipe="dfma"):
    """Measured FP64 peak (TFLOP/s): pipe = "dfma" (scalar FMA) or "dmma" (FP64 tensor pipe, mma.sync.m8n8k4.f64)."""
    lib = _lib.load_library()
    v = C.c_double()
    fn = lib.rnmpc_fp64_mma_peak if pipe == "dmma" else lib.rnmpc_fp64_peak
    _lib.check(fn(int(device), C.byref(v)))
    return float(v.value)
