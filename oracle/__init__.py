"""CPU oracle for the SCP / fast-SLS robust-MPC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and there only as the checker / the timed CPU baseline.  The product path (``robust-nonlinear-mpc_b200``)
never imports this package and fails loudly when its CUDA library is missing.

The oracle is a plain NumPy restatement of the reference algorithm
(``/root/reference/solver/{SCP_SLS_jit,fast_SLS_jit,qp_jit,ocp}.py`` and ``dyn/*.py``); every function cites the
reference file:line it follows.

Parity pinning:
  * Riccati backward sweep, Phi propagation and tube back-offs are pinned against the reference's OWN Numba
    kernels (``solver/fast_SLS_jit.py:43-188``), imported in the build container under import stubs by
    ``oracle/gen_golden.py``; their outputs are committed as ``tests/golden/sls_*.npz``.
  * The QP handed to OSQP (assembly, bounds, x0 rows, unpacking) is pinned against the reference's own ``QP``
    class driven over a recording ``osqp`` stand-in (``oracle/ref_qp.py`` -> ``tests/golden/refqp_*.npz``).
  * The fast-SLS glue (dual weights eta, tightening with quirks Q4 / Q7, the Q3 test) is pinned against the
    reference's own ``fast_SLS`` methods called unbound on an attribute bag (``oracle/ref_glue.py`` ->
    ``tests/golden/refglue_*.npz``).
  * Dynamics/Jacobians are pinned against closed-form SymPy differentiation of the same ODE and against
    finite differences (the reference evaluates them through CasADi 3.7.1, absent here).  The ODEs themselves
    are pinned against the reference's own ``dyn/*.py`` executed with a numeric CasADi stand-in
    (``oracle/ref_models.py`` -> ``tests/golden/refdyn_*.npz``).
  * The nominal QP is solved in the reference by OSQP 1.0.4 (absent here, un-vendored) with default
    tolerances plus polish: PARITY UNPINNED against the library at that boundary; ``oracle/osqp_restated.py``
    restates its published algorithm, and its polished answer coincides with this oracle's on every QP fixture
    (tests/test_oracle_cpu.py::test_restated_osqp_*).  The QP is strictly convex, so its primal
    solution is unique; the oracle solves it to KKT residual <= 1e-9 with an independent dense primal-dual
    interior-point method plus exact active-set refinement.  The MULTIPLIERS, which the algorithm also consumes
    (eta = mu / (2 sqrt(beta))), are unique only where the active-constraint gradients are linearly independent:
    on degenerate instances (about one in a thousand random-x0 quadrotor scenarios, tools/degeneracy_check.py)
    different solvers -- OSQP included -- return different multipliers, and closed-loop parity at 1e-5 is not well
    defined there.
"""
