"""``SCP_SLS``: the reference's outer SCP / RTI solver class (solver/SCP_SLS_jit.py:15-551) on the CUDA library.

Same constructor, kwargs, methods and result-dict keys as the reference, so ``expe/main_*_robust_closed_loop.py`` run
unchanged with this package's directory on ``sys.path``.  One extra kwarg, ``batch`` (default 1): the number of
independent MPC instances solved per call; with ``batch > 1`` every array argument / result gains a leading batch axis.
All numerical work (linearisation, nominal QPs, Riccati sweeps, Phi propagation, back-offs, warm-start shift) happens
in ``librnmpc_b200.so``; this file only moves NumPy arrays in and out and keeps the reference's bookkeeping.
"""
from __future__ import annotations

import numpy as np

from ..engine import BatchSolver
from .ocp import OCP

_FAIL, _Q3, _TIGHT, _KVALID = 1, 2, 4, 16


class _FastSLSView:
    """Stands in for ``SCP_SLS.fast_SLS_solver`` (drivers poke ``verbose``, ``CONV_EPS``, ``MAX_ITER``, ...;
    expe/main_rocket_robust_closed_loop.py:86-93).  Settings that change the computation are forwarded to the engine or refused;
    none is silently ignored."""

    class _QPView:
        verbose = False
        export_standard_qp = False

    def __init__(self, owner):
        self._owner = owner
        self.verbose = False
        self.save_it_data = False
        # the reference stores CONV_EPS but check_convergence_socp tests against a hard-coded 1e-3 (fast_SLS_jit.py:594): kept, unused
        self.CONV_EPS = 1e-6
        self._max_iter = 30
        self.solver_forward = self._QPView()

    @property
    def MAX_ITER(self):
        """fast_SLS.MAX_ITER (fast_SLS_jit.py:206): cap of the iterate-to-convergence mode; forwarded to the engine."""
        return self._max_iter

    @MAX_ITER.setter
    def MAX_ITER(self, value):
        value = int(value)
        if value < 1:
            raise ValueError("MAX_ITER must be >= 1")
        self._max_iter = value
        if self._owner._engine is not None:
            self._owner._engine.set_option("sls_max_iter", value)

    @property
    def epsilon_backoff(self):
        return 1e-10

    @epsilon_backoff.setter
    def epsilon_backoff(self, value):
        # a compile-time constant of the kernels (rnmpc_common.cuh: kEpsBackoff = fast_SLS_jit.py:205)
        if float(value) != 1e-10:
            raise NotImplementedError("epsilon_backoff is fixed at the reference's 1e-10 in the CUDA kernels")

    def set_rti_steps(self, steps):
        self._owner.set_fast_sls_rti_steps(steps)


class LazySolution(dict):
    """Result dict of ``SCP_SLS.solve``: the keys every driver reads (trajectories, back-offs, success, timings) are plain
    entries; the rest of the reference's key set (duals, eta, K, beta, the linearisation lists, ...) is copied from the GPU on first
    access.  ``materialize()`` evaluates everything.  Once the solver has advanced (next ``solve`` / ``reset_warm_start``) keys that
    were never read are gone: they describe device buffers that have been overwritten."""

    def __init__(self, eager, thunks):
        super().__init__(eager)
        self._thunks = dict(thunks)
        self._expired = False

    def __missing__(self, key):
        th = self._thunks.pop(key, None)
        if th is None:
            raise KeyError(key)
        if self._expired:
            raise KeyError(f"result key {key!r} was not read before the solver advanced; read it right after solve() or call materialize()")
        for k, v in th().items():
            dict.__setitem__(self, k, v)
            self._thunks.pop(k, None)
        return dict.__getitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._thunks

    def get(self, key, default=None):
        return self[key] if key in self else default

    def keys(self):
        return list(dict.keys(self)) + [k for k in self._thunks if not dict.__contains__(self, k)]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def materialize(self):
        for k in list(self._thunks):
            if k in self._thunks:
                self[k]
        return self

    def items(self):
        self.materialize()
        return dict.items(self)

    def values(self):
        self.materialize()
        return dict.values(self)

    def copy(self):
        return dict(self.items())

    def _expire(self):
        self._expired = True


class SCP_SLS(OCP):
    def __init__(self, N, Q, R, m, Qf, Q_reg=None, R_reg=None, Q_reg_f=None, **kwargs):
        super().__init__(N, Q, R, m, Qf, Q_reg, R_reg, Q_reg_f)
        self.epsilon_convergence = 1e-10
        self.current_iteration_scp = {}
        self.convergence_data_scp = {}
        self.it_data = {}
        self.save_it_data = kwargs.get("save_it_data", True)
        self.verbose = kwargs.get("verbose", True)
        self.rti = kwargs.get("rti", -1)
        self.warm_start_ubg_shift_enabled = kwargs.get("warm_start_ubg_shift", True)   # dead in the reference (:82)
        self.fast_sls_rti_steps = kwargs.get("fast_sls_rti_steps", None)
        self.linearization_error = kwargs.get("linearization_error", False)
        if self.linearization_error:
            raise NotImplementedError("Linearization error is not implemented yet")       # SCP_SLS_jit.py:263-264
        self.batch = int(kwargs.get("batch", 1))
        self.device = int(kwargs.get("device", 0))
        self.nominal_init_iters = int(kwargs.get("nominal_init_iters", 10))
        self.lazy_results = bool(kwargs.get("lazy_results", True))      # False: every key is copied at solve time, as in the reference
        # phase timers behind the reference's t_jac_ms / t_qp_ms / t_backward_ms keys; False: the keys are NaN and a solve of a small
        # batch is replayed as one CUDA graph (rnmpc_solve_host)
        self.timing = bool(kwargs.get("timing", True))
        self._last_solution = None
        self.MAX_ITER_SCP = 100
        self.fast_SLS_solver = _FastSLSView(self)
        self._engine = None
        self._has_nominal = False
        self._build_engine()

    # ---------------------------------------------------------------------------------------------------- engine
    def _build_engine(self):
        # fast_sls_rti_steps None / <= 0: iterate fast-SLS until its convergence test (fast_SLS_jit.py:298-312) -> sls_steps = 0
        sls_steps = int(self.fast_sls_rti_steps) if (self.fast_sls_rti_steps is not None and self.fast_sls_rti_steps > 0) else 0
        m = self.m
        if getattr(m, "model_name", None) not in ("pendulum", "quadrotor", "rocket"):
            raise TypeError("model must be one of this package's dyn.Pendulum / dyn.Quadrotor / dyn.Rocket")
        if m.nw != m.nx:
            raise NotImplementedError("nw == nx is assumed (all reference models)")
        x_ub, x_lb, u_ub, u_lb, xf_ub, xf_lb = m.box()
        if self._engine is not None:
            self._engine.close()
        rti = int(self.rti) if (self.rti is not None and self.rti > 0) else 1
        self._engine = BatchSolver(m.model_name, self.N, self.Q, self.R, self.Qf, self.Q_reg, self.R_reg, self.Q_reg_f,
                                   np.asarray(m.E, float), x_ub, x_lb, u_ub, u_lb, xf_ub, xf_lb, rti=rti,
                                   sls_steps=sls_steps, capacity=self.batch, device=self.device)
        self._engine.enable_timing(self.timing)
        self._engine.set_option("sls_max_iter", self.fast_SLS_solver.MAX_ITER)
        self._has_nominal = False

    @property
    def engine(self) -> BatchSolver:
        return self._engine

    def set_fast_sls_rti_steps(self, steps):
        """solver/SCP_SLS_jit.py:394-401 -> fast_SLS.set_rti_steps: changes the step count only; the nominal trajectory, the Q3 history
        and the gains stay (engine option "sls_steps", no new handle)."""
        self.fast_sls_rti_steps = steps
        self._engine.set_option("sls_steps", int(steps) if (steps is not None and steps > 0) else 0)

    # ------------------------------------------------------------------------------------------------- nominal
    def _b(self, a, shape):
        a = np.asarray(a, dtype=float)
        return a.reshape((self.batch,) + shape)

    def set_nominal_trajectory(self, primal_x, primal_u):
        """Provide the initial nominal trajectory explicitly (primal_x (nx, N+1), primal_u (nu, N); leading batch axis
        when batch > 1).  Replaces the IPOPT guess of solve_nominal_trajectory (SCP_SLS_jit.py:161-188)."""
        nx, nu, N = self.m.nx, self.m.nu, self.N
        X = np.swapaxes(self._b(primal_x, (nx, N + 1)), 1, 2)
        U = np.swapaxes(self._b(primal_u, (nu, N)), 1, 2)
        self._engine.set_nominal(np.ascontiguousarray(X), np.ascontiguousarray(U))
        self._has_nominal = True
        self.current_iteration_scp["primal_x"] = self._out(np.swapaxes(X, 1, 2))
        self.current_iteration_scp["primal_u"] = self._out(np.swapaxes(U, 1, 2))

    def solve_nominal_trajectory(self, x0):
        """First-step nominal trajectory by the tube-free SQP of ``rnmpc_init_nominal`` (the reference calls IPOPT here)."""
        self._engine.init_nominal(self._b(x0, (self.m.nx,)), self.nominal_init_iters)
        ok = not bool((self._engine.buffer("flags", self.batch) & _FAIL).any())
        self._has_nominal = ok
        if self.verbose:
            print('SCP-SLS: Initial guess nominal trajectory converged!' if ok
                  else 'SCP-SLS: Initial guess nominal trajectory did not converge!')
        return ok

    def _out(self, a):
        return a[0] if self.batch == 1 else a

    # ---------------------------------------------------------------------------------------------------- solve
    def _advance(self):
        if self._last_solution is not None:
            self._last_solution._expire()
            self._last_solution = None

    def solve(self, x0):
        """solver/SCP_SLS_jit.py:65-152."""
        nx, nu, N, B = self.m.nx, self.m.nu, self.N, self.batch
        x0b = np.ascontiguousarray(self._b(x0, (nx,)))
        if not self._has_nominal:
            if not self.solve_nominal_trajectory(x0b):
                return {'success': False}
        self._advance()
        eng = self._engine
        if self.rti is not None and self.rti > 0:
            # one call: x0 in, the solve, trajectories / back-offs / flags out (rnmpc_solve_host: one CUDA graph for small batches)
            res = eng.solve_host(x0b)
            iterations = int(self.rti) - 1
            converged = None
        else:
            # run until max|delta| < epsilon_convergence (SCP_SLS_jit.py:108-133); each pass = linearise + fast-SLS step
            converged, iterations = False, 0
            for ii in range(int(self.MAX_ITER_SCP)):
                iterations = ii
                res = eng.solve_host(x0b)
                if (res["flags"] & _FAIL).any():
                    break
                delta = max(np.abs(eng.buffer("dX", B)).max(), np.abs(eng.buffer("dU", B)).max())
                if delta < self.epsilon_convergence:
                    converged = True
                    break
        flags = res["flags"]
        ok = (flags & _FAIL) == 0
        tight = (flags & _TIGHT) != 0
        kvalid = (flags & _KVALID) != 0
        ni = self.m.ni
        X, U, backoff, backoff_f = res["X"], res["U"], res["backoff"], res["backoff_f"]
        eps = self.fast_SLS_solver.epsilon_backoff
        # instances whose last fast-SLS solve took the Q3 early return hold the initial back-offs
        # (fast_SLS_jit.py:444-454): beta = eps, backoff = N*sqrt(eps), backoff_x = backoff_u = 0
        nt = ~tight
        if nt.any():
            backoff[nt], backoff_f[nt] = N * np.sqrt(eps), (N + 1) * np.sqrt(eps)
        backoff_x = np.concatenate([backoff[:, :, :nx], backoff_f[:, None, :nx]], axis=1)      # (B, N+1, nx)  Q7
        backoff_u = backoff[:, :, nx:nx + nu].copy()
        backoff_x[nt], backoff_u[nt] = 0.0, 0.0
        primal_x, primal_u = np.swapaxes(X, 1, 2), np.swapaxes(U, 1, 2)
        pack = lambda Xb, Ub: np.stack([self.pack_primal_nominal(Xb[b], Ub[b]) for b in range(B)])
        tm = eng.timings() if self.timing else dict(t_jac_ms=float('nan'), t_qp_ms=float('nan'), t_backward_ms=float('nan'))
        o = self._out
        if converged is None:
            success = ok
        else:
            success = ok & bool(converged)
        eager = {
            'primal_x': o(primal_x), 'primal_u': o(primal_u),
            'backoff': o(backoff), 'backoff_f': o(backoff_f), 'backoff_x': o(backoff_x), 'backoff_u': o(backoff_u),
            'SOCP_steps': o(np.where(tight, int(self.fast_sls_rti_steps or 0), 0)),
            'iterations': iterations,
            't_jac_ms': float(tm['t_jac_ms']), 't_qp_ms': float(tm['t_qp_ms']),
            't_backward_ms': float(tm['t_backward_ms']) if tight.all() else float('nan'),
            'E_list': [np.asarray(self.m.E, float) for _ in range(N + 1)],
            'status': o(res["status"]),
            'success': bool(success[0]) if B == 1 else success,
        }
        if B > 1:
            eager['tube_updated'] = tight

        # ---- the rest of the reference's key set, copied from the device on first access
        def t_primal_vec():
            return {'primal_vec': o(pack(primal_x, primal_u))}

        def t_delta():
            dX, dU = eng.buffer("dX", B), eng.buffer("dU", B)
            return {'delta_vec': o(pack(np.swapaxes(dX, 1, 2), np.swapaxes(dU, 1, 2)))}

        def t_duals():
            mu, mu_f = eng.buffer("mu", B), eng.buffer("mu_f", B)
            beta, beta_f = eng.buffer("beta", B), eng.buffer("beta_f", B)
            if nt.any():
                beta[nt], beta_f[nt] = eps, eps
            eta = np.zeros((B, N, N, ni))
            sq = 2.0 * np.sqrt(np.maximum(beta, eps))
            for k in range(N):
                eta[:, k, :k + 1, :] = mu[:, k, None, :] / sq[:, k, :k + 1, :]
            eta_f = mu_f[:, None, :] / (2.0 * np.sqrt(np.maximum(beta_f, eps)))
            lam = eng.costates(B).cpu().numpy()
            return {'dual_mu': o(np.swapaxes(mu, 1, 2)), 'dual_mu_f': o(mu_f), 'dual_vec': o(self._dual_vec(mu, mu_f, lam)),
                    'dual_eta': o(eta), 'dual_eta_f': o(eta_f), 'beta': o(beta), 'beta_f': o(beta_f)}

        def t_cost():
            cost_qp = eng.buffer("cost", B)
            Qs = np.einsum("bik,ij,bjk->b", primal_x[:, :, :N], self.Q, primal_x[:, :, :N]) \
                + np.einsum("bik,ij,bjk->b", primal_u, self.R, primal_u) \
                + np.einsum("bi,ij,bj->b", primal_x[:, :, N], self.Qf, primal_x[:, :, N])
            return {'cost_QP': o(cost_qp), 'cost': o(cost_qp + Qs), 'qp_iterations': o(eng.buffer("qp_iters", B))}

        def t_gains():
            # K / Phi are None when no tube was computed since the last reset (fast_SLS_jit.py:629-632 after the Q3 early
            # return); a stale K of an earlier SCP iteration of the same solve is returned as in the reference
            K = eng.buffer("K", B)
            out = {'K': (K[0] if kvalid[0] else None) if B == 1 else K}
            if eng._views.get("phi"):
                Px, Pu = eng.buffer("Phi_x", B), eng.buffer("Phi_u", B)
                out['Phi_x'] = (Px[0] if kvalid[0] else None) if B == 1 else Px
                out['Phi_u'] = (Pu[0] if kvalid[0] else None) if B == 1 else Pu
            else:
                out['Phi_x'] = out['Phi_u'] = None
            return out

        def t_lin():
            # what the trailing update_jacobian (SCP_SLS_jit.py:139) and the diagnostics (:451-455) leave in the reference's dict:
            # the linearisation of the UPDATED nominal trajectory and the signed maximum of its dynamics defect
            lin = self.linearization()
            c = lin['c_offset_list']
            cmax = np.max(np.stack([ck.reshape(-1) for ck in c])) if B == 1 else np.asarray(c).reshape(B, -1).max(axis=1)
            lin['primal_infeasibility'] = float(cmax) if B == 1 else cmax
            return lin

        groups = [(('primal_vec',), t_primal_vec), (('delta_vec',), t_delta),
                  (('dual_mu', 'dual_mu_f', 'dual_vec', 'dual_eta', 'dual_eta_f', 'beta', 'beta_f'), t_duals),
                  (('cost_QP', 'cost', 'qp_iterations'), t_cost), (('K', 'Phi_x', 'Phi_u'), t_gains),
                  (('A_list', 'B_list', 'c_offset_list', 'g_list', 'primal_infeasibility'), t_lin)]
        thunks = {k: th for keys, th in groups for k in keys}
        sol = LazySolution(eager, thunks)
        if self.save_it_data:
            self.it_data[iterations] = {k: sol[k] for k in ('primal_x', 'primal_u', 'backoff_x', 'backoff_u', 'cost')}
        sol.update(self.convergence_data_scp)
        dict.__setitem__(sol, 'it_data', self.it_data.copy())
        if not self.lazy_results:
            sol.materialize()
        else:
            self._last_solution = sol
        self.current_iteration_scp = sol
        return sol

    _keep_failed = True

    def _dual_vec(self, mu, mu_f, lam_dyn):
        """``dual_vec`` in the reference's row order (qp_jit.py:493-501): per stage the nx multipliers of the dynamics rows
        (rnmpc_costates) and the ni inequality multipliers, then the terminal ones (the x0 rows are dropped as in the reference)."""
        B, N, nx, ni = mu.shape[0], self.N, self.m.nx, self.m.ni
        lam = np.empty((B, N, nx + ni))
        lam[:, :, :nx] = lam_dyn
        lam[:, :, nx:] = mu
        return np.concatenate([lam.reshape(B, -1), mu_f], axis=1)

    def enable_phi(self, enable=True):
        """Also return the dense closed-loop responses ``Phi_x`` / ``Phi_u`` (0.7 MB per rocket instance)."""
        self._engine.enable_phi(enable)
        self._engine._views["phi"] = bool(enable)

    def linearization(self):
        """A_list, B_list, c_offset_list, g_list of the *current* nominal trajectory (what the trailing
        update_jacobian of SCP_SLS_jit.py:139 leaves in the result dict), computed on demand."""
        B, N = self.batch, self.N
        eng = self._engine
        X, U = eng.buffer("X", B), eng.buffer("U", B)
        A, Bm, c = (t.cpu().numpy() for t in eng.linearize(X, U))
        m = self.m
        g = np.asarray(m.g, float)[None, None, :] - np.einsum("ij,bkj->bki", np.asarray(m.G, float), np.concatenate([X[:, :N], U], axis=2))
        gN = np.asarray(m.gf, float)[None, :] - X[:, N] @ np.asarray(m.Gf, float).T
        if B == 1:
            return dict(A_list=list(A[0]), B_list=list(Bm[0]), c_offset_list=[ck.reshape(-1, 1) for ck in c[0]],
                        g_list=list(g[0]) + [gN[0]])
        return dict(A_list=A, B_list=Bm, c_offset_list=c, g_list=(g, gN))

    # ------------------------------------------------------------------------------------------------ warm start
    def reset_warm_start(self):
        """solver/SCP_SLS_jit.py:500-551."""
        if not self._has_nominal:
            return
        self._advance()
        self._engine.reset_warm_start(self.batch)
        self.current_iteration_scp = {}
        self.convergence_data_scp = {}
        self.it_data = {}
        X, U = self._engine.buffer("X", self.batch), self._engine.buffer("U", self.batch)
        self.current_iteration_scp['primal_x'] = self._out(np.swapaxes(X, 1, 2))
        self.current_iteration_scp['primal_u'] = self._out(np.swapaxes(U, 1, 2))

    def reset(self):
        """solver/SCP_SLS_jit.py:486-498."""
        self.current_iteration_scp = {}
        self.convergence_data_scp = {}
        self.it_data = {}
        self._build_engine()

    def check_convergence_scp(self):
        """solver/SCP_SLS_jit.py:610-622."""
        dv = self.current_iteration_scp.get('delta_vec')
        return dv is not None and float(np.max(np.abs(dv))) < self.epsilon_convergence
