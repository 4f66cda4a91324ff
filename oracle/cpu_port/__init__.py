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
