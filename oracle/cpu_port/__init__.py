"""ctypes wrapper of the C++ CPU port (oracle/cpu_port/rnmpc_cpu.cpp): timed CPU baseline of bench.py.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/__init__.py)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "librnmpc_cpu.so")
_lib = None


def build():
    subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.rnmpc_cpu_closed_loop.restype = C.c_longlong
        _lib.rnmpc_cpu_closed_loop_resume.restype = C.c_longlong
    return _lib


def closed_loop(m, cfg, x0, W, steps, X0, U0, threads=0, return_ok=False):
    """B closed loops on the host: x0 [B,nx], W [B,steps,nx] or None, X0 [B,N+1,nx], U0 [B,N,nu] -> states, inputs, n_failed
    (+ ok [B,steps] int32 with return_ok: 1 where the solve of that scenario and step succeeded; the port keeps stepping a
    scenario after a failed solve with the stale input, it does not freeze it)."""
    lib = load()
    d = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    x0, X0, U0 = d(x0).reshape(-1, m.nx), d(X0), d(U0)
    B = x0.shape[0]
    Wc = None if W is None else d(W).reshape(B, steps, m.nx)
    keep = [d(cfg[k]) for k in ("Q", "R", "Qf", "Q_reg", "R_reg", "Q_reg_f")] + [d(m.E), d(m.x_ub), d(m.x_lb), d(m.u_ub), d(m.u_lb),
                                                                            d(m.x_ub), d(m.x_lb)]
    states, inputs = np.empty((B, steps + 1, m.nx)), np.empty((B, steps, m.nu))
    ok = np.ones((B, steps), dtype=np.int32) if return_ok else None
    nfail = lib.rnmpc_cpu_closed_loop(C.c_int(m.model_id), C.c_int(cfg["N"]), C.c_int(cfg["rti"]), C.c_int(cfg["fast_sls_rti_steps"]),
                                      *[p(a) for a in keep], C.c_double(0.05), C.c_int(B), C.c_int(steps), p(x0), p(Wc), p(X0), p(U0),
                                      p(states), p(inputs), C.c_int(threads), p(ok))
    if return_ok:
        return states, inputs, int(nfail), ok
    return states, inputs, int(nfail)


class ResumableLoops:
    """B closed loops that can be advanced in several calls (bench.py: warm-up MPC steps in an untimed call, the timed steps in the
    next one) with the GPU arm's Monte-Carlo rule: a scenario whose solve failed is frozen (freeze_failed) and costs nothing afterwards."""

    def __init__(self, m, cfg, x0, X0, U0, freeze_failed=True, threads=0):
        d = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.float64))
        self.m, self.cfg, self.threads, self.freeze = m, cfg, int(threads), bool(freeze_failed)
        self.x = d(x0).reshape(-1, m.nx).copy()
        self.B = self.x.shape[0]
        self.X, self.U = d(X0).copy(), d(U0).copy()
        self.prevX, self.prevU = np.zeros_like(self.X), np.zeros_like(self.U)
        self.have_prev = np.zeros(self.B, dtype=np.int32)
        self.alive = np.ones(self.B, dtype=np.int32)
        self.started = False
        self._keep = [d(cfg[k]) for k in ("Q", "R", "Qf", "Q_reg", "R_reg", "Q_reg_f")] + [d(m.E), d(m.x_ub), d(m.x_lb), d(m.u_ub), d(m.u_lb),
                                                                                   d(m.x_ub), d(m.x_lb)]

    def advance(self, W, steps):
        """steps MPC steps of every live scenario with disturbances W [B,steps,nx] (None: noise free) -> (live solves, failed solves)."""
        lib = load()
        m, cfg = self.m, self.cfg
        p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        Wc = None if W is None else np.ascontiguousarray(np.asarray(W, dtype=np.float64)).reshape(self.B, steps, m.nx)
        states = np.empty((self.B, steps + 1, m.nx))
        live = C.c_longlong(0)
        nfail = lib.rnmpc_cpu_closed_loop_resume(C.c_int(m.model_id), C.c_int(cfg["N"]), C.c_int(cfg["rti"]), C.c_int(cfg["fast_sls_rti_steps"]),
                                                 *[p(a) for a in self._keep], C.c_double(0.05), C.c_int(self.B), C.c_int(steps), p(self.x), p(Wc),
                                                 p(self.X), p(self.U), p(self.prevX), p(self.prevU), p(self.have_prev), p(self.alive),
                                                 C.c_int(int(self.freeze)), C.c_int(int(self.started)), p(states), None, C.c_int(self.threads),
                                                 C.byref(live))
        self.x = np.ascontiguousarray(states[:, steps])
        self.started = True
        return int(live.value), int(nfail)
