/* rnmpc_b200 -- C ABI of the B200-native robust-nonlinear-MPC inner solve (SCP / fast-SLS / nominal QP).
 *
 * The reference (antoineleeman/robust-nonlinear-mpc) has no FFI: its boundary for this path is the Python object
 * API of solver/SCP_SLS_jit.py, solver/fast_SLS_jit.py and solver/qp_jit.py.  Each entry point below names the
 * reference interface it replaces (file:line).  The library is batch-first: every call processes `B` independent
 * MPC instances (Monte-Carlo scenarios); the Python mirror of the reference classes
 * (robust-nonlinear-mpc_b200/solver/*.py) binds these symbols through ctypes with B = 1 or B = batch.
 *
 * Conventions
 *   - all numeric buffers are float64, C-order, batch-major; `d_` = device pointer (cudaMalloc / torch CUDA tensor
 *     data_ptr), `h_` = host pointer.  The caller owns every buffer it passes.  A handle is bound to one CUDA device
 *     and is NOT thread-safe (the reference solver objects are not either).
 *   - return value: 0 = OK, < 0 = argument / CUDA error (rnmpc_last_error() gives the text).  Solver failure of an
 *     instance is NOT an error: it is reported per instance in `status` (0 ok, 1 QP not converged, 2 QP infeasible,
 *     3 numerical failure), mirroring the reference's {'success': False} (qp_jit.py:397-400, fast_SLS_jit.py:334-341).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous unless stated.
 *   - sizes: nx, nu from the model; nz = nx+nu; ni = 2*nz; ni_f = 2*nx; nw = nx.
 *   - box constraints only: G = [I; -I], Gf = [I; -I] (all reference models, dyn/rocket.py:129-139).
 */
#ifndef RNMPC_B200_H
#define RNMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rnmpc_handle rnmpc_handle;

enum { RNMPC_PENDULUM = 0, RNMPC_QUADROTOR = 1, RNMPC_ROCKET = 2 };

/* Host-side problem description (all pointers are host pointers, copied at create time).
 * Replaces the constructor arguments SCP_SLS(N, Q, R, m, Qf, Q_reg, R_reg, Q_reg_f, rti, fast_sls_rti_steps)
 * (solver/SCP_SLS_jit.py:24-44) and the model attributes m.E, m.g, m.gf (dyn/model.py:4-13, dyn/rocket.py:129-163). */
typedef struct rnmpc_config {
    int model_id;            /* RNMPC_PENDULUM / RNMPC_QUADROTOR / RNMPC_ROCKET */
    int N;                   /* horizon */
    int rti;                 /* SCP iterations per solve (>0); reference kwarg `rti` */
    int sls_steps;           /* fast-SLS iterations per SCP iteration; reference kwarg `fast_sls_rti_steps`.  0 = iterate until
                              * check_convergence_socp (cap MAX_ITER = 30, then failure: status 4), fast_SLS_jit.py:298-312 */
    const double* Q;         /* [nx*nx] */
    const double* R;         /* [nu*nu] */
    const double* Qf;        /* [nx*nx] */
    const double* Q_reg;     /* [nx*nx] */
    const double* R_reg;     /* [nu*nu] */
    const double* Q_reg_f;   /* [nx*nx] */
    const double* E;         /* [nx*nx] disturbance map (nw == nx) */
    const double* x_ub;      /* [nx]  g = [x_ub, u_ub, -x_lb, -u_lb] */
    const double* x_lb;      /* [nx] */
    const double* u_ub;      /* [nu] */
    const double* u_lb;      /* [nu] */
    const double* xf_ub;     /* [nx]  gf = [xf_ub, -xf_lb] */
    const double* xf_lb;     /* [nx] */
    double h;                /* RK4 step of ddyn (0.05 on the SCP path, dyn/model.py:15) */
    int qp_max_iter;         /* interior-point iteration cap (default 40 if <= 0) */
    double qp_tol_mu;        /* complementarity tolerance (default 1e-14 if <= 0) */
    double qp_tol_res;       /* dynamics-residual tolerance (default 1e-9 if <= 0) */
} rnmpc_config;

const char* rnmpc_last_error(void);
int rnmpc_model_dims(int model_id, int* nx, int* nu);

/* Lifetime.  `capacity` = largest batch the handle's internal workspace serves; `device` = CUDA ordinal. */
int rnmpc_create(rnmpc_handle** out, const rnmpc_config* cfg, int capacity, int device);
int rnmpc_destroy(rnmpc_handle* h);

/* ---- stateless kernels on caller-owned device buffers ------------------------------------------------------ */

/* x+ = ddyn(x, u) [+ E w]   (dyn/model.py:15-34; plant step expe/main_rocket_compare_closed_loop.py:159).
 * d_x [B][nx], d_u [B][nu], d_w [B][nx] or NULL, d_xp [B][nx]; step <= 0 selects the handle's h. */
int rnmpc_ddyn(rnmpc_handle* h, int B, const double* d_x, const double* d_u, const double* d_w, double step, double* d_xp,
               void* stream);

/* A_k, B_k, c_k along a trajectory (SCP_SLS.update_jacobian, solver/SCP_SLS_jit.py:251-277).
 * d_X [B][N+1][nx], d_U [B][N][nu] -> d_A [B][N][nx][nx], d_B [B][N][nx][nu], d_c [B][N][nx]. */
int rnmpc_linearize(rnmpc_handle* h, int B, const double* d_X, const double* d_U, double* d_A, double* d_B, double* d_c,
                    void* stream);

/* fast-SLS backward Riccati sweep + Phi propagation + back-offs with explicit weights
 * (_backward_solve_numba, _propagate, _backoff_from_phi: solver/fast_SLS_jit.py:65-188).
 * d_eta [B][N][N][ni], d_eta_f [B][N+1][ni_f] -> d_K [B][N][N+1][nu][nx], d_beta [B][N][N][ni], d_beta_f [B][N+1][ni_f],
 * d_backoff [B][N][ni], d_backoff_f [B][ni_f]; optional (NULL to skip) d_S [B][N+1][N+1][nx][nx],
 * d_Phi_x [B][N+1][N+1][nx][nx], d_Phi_u [B][N][N+1][nu][nx].  Entries with j > k are not written (zero them once). */
int rnmpc_sls(rnmpc_handle* h, int B, const double* d_A, const double* d_B, const double* d_eta, const double* d_eta_f,
              double* d_K, double* d_S, double* d_Phi_x, double* d_Phi_u, double* d_beta, double* d_beta_f,
              double* d_backoff, double* d_backoff_f, void* stream);

/* Nominal QP (QP.solve, solver/qp_jit.py:343-402 + _pack_solution :487-513).
 * tightened = 0: bounds g_k - G z_k + 1e-10, terminal gf - Gf x_N + 1e-10 (after update_dynamics/offset_constraints);
 * tightened = 1: g_k - G z_k - backoff_k, terminal gf - backoff_f (after fast_SLS.update_tightening, quirk Q4).
 * Inputs d_A, d_B, d_c, nominal d_X [B][N+1][nx], d_U [B][N][nu], measured state d_x0 [B][nx], d_backoff/d_backoff_f
 * (tightened only).  Outputs d_dX [B][N+1][nx], d_dU [B][N][nu], d_mu [B][N][ni] (= dual_mu^T), d_mu_f [B][ni_f],
 * d_cost [B], d_iters [B], d_status [B]. */
int rnmpc_qp(rnmpc_handle* h, int B, int tightened, const double* d_A, const double* d_B, const double* d_c,
             const double* d_X, const double* d_U, const double* d_x0, const double* d_backoff, const double* d_backoff_f,
             double* d_dX, double* d_dU, double* d_mu, double* d_mu_f, double* d_cost, int* d_iters, int* d_status,
             void* stream);

/* The same QP with explicit data, as the reference's QP object holds it after update_dynamics / offset_constraints /
 * update_ubg / update_q_cost_lin (solver/qp_jit.py:518-628): bounds d_hi, d_lo and linear cost d_q (NULL = 0) in "entry"
 * layout [B][N+1][nz] (stage k: x_k entries then u_k entries; stage N uses the first nx), dynamics offsets d_c, and the
 * value d_x0pin [B][nx] the first state is pinned to (QP.solve(x0) pins it to -x0, qp_jit.py:376-379).  The stage-0 state
 * bounds only serve the feasibility check.  Outputs as rnmpc_qp. */
int rnmpc_qp_general(rnmpc_handle* h, int B, const double* d_A, const double* d_B, const double* d_c, const double* d_hi,
                     const double* d_lo, const double* d_q, const double* d_x0pin, double* d_dX, double* d_dU, double* d_mu,
                     double* d_mu_f, double* d_cost, int* d_iters, int* d_status, void* stream);

/* fast-SLS pass from QP duals: evaluate_dual_eta + backward_solve + update_tightening kernels
 * (solver/fast_SLS_jit.py:475-571) in one launch.  d_mu [B][N][ni], d_mu_f [B][ni_f]; d_beta_in [B][N][N][ni] /
 * d_beta_f_in [B][N+1][ni_f] are the previous beta (both NULL = epsilon_backoff, the state after initialize_backoff).
 * d_beta_in may alias d_beta.  Outputs as rnmpc_sls. */
int rnmpc_sls_duals(rnmpc_handle* h, int B, const double* d_A, const double* d_B, const double* d_mu, const double* d_mu_f,
                    const double* d_beta_in, const double* d_beta_f_in, double* d_K, double* d_S, double* d_Phi_x, double* d_Phi_u,
                    double* d_beta, double* d_beta_f, double* d_backoff, double* d_backoff_f, void* stream);

/* ---- stateful solver on the handle's resident trajectories ----------------------------------------------- */

/* Multipliers of the dynamics rows of the last nominal QP of the resident solve -- the first nx entries of every stage block of the
 * reference's `dual_vec` (qp_jit.py:491-501; rows `A_k dx_k + B_k du_k - dx_{k+1} = -c_k`, qp_jit.py:194-273) -- recovered from the
 * stationarity conditions of the solution (OSQP sign convention P x + q + A' y = 0).  d_lam [B][N][nx].  On demand: not on the solve path. */
int rnmpc_costates(rnmpc_handle* h, int B, double* d_lam, void* stream);

/* Soft-constrained NMPC comparator (solver/nlp_soft_constraints.py:175-276 `NLPSoftConstraints`, handed to IPOPT by the reference;
 * expe/main_rocket_compare_closed_loop.py:176-222 `_run_soft`):  min sum x'Qx + u'Ru + x_N'Qf x_N + rho (|gamma|^2 + |gamma_f|^2) + rho1 sum(gamma)
 * s.t. x_{k+1} = ddyn(x_k, u_k), x_0 = x0, G [x_k; u_k] - g <= gamma_k, Gf x_N - gf <= gamma_f, gamma >= 0, by full-step SQP from the
 * trajectory guess in d_X [B][N+1][nx] / d_U [B][N][nu] (in / out) until max |step| < tol per instance, at most max_iter iterations
 * (status 4 = not converged; a QP failure gives its status).  Every iteration = the linearisation kernel + the interior-point QP kernel
 * with the slack variables eliminated entry-wise (rnmpc_qp_warp.cuh, SOFT).  Outputs: d_gamma [B][N][ni], d_gamma_f [B][ni_f] (both NULL to
 * skip), d_sqp_iters [B], d_status [B].  Uses the handle's workspace (not its resident robust trajectories).  Synchronises once per iteration. */
int rnmpc_soft_nmpc(rnmpc_handle* h, int B, const double* d_x0, double* d_X, double* d_U, double rho, double rho1, int max_iter, double tol,
                    double* d_gamma, double* d_gamma_f, int* d_sqp_iters, int* d_status, void* stream);

/* Load the nominal trajectories of B instances (replaces the IPOPT initial guess, SCP_SLS_jit.py:161-188, and is what
 * reset()/a fresh solver start from).  Clears the warm-start state incl. the Q3 history. */
int rnmpc_set_nominal(rnmpc_handle* h, int B, const double* d_X, const double* d_U, void* stream);

/* Tube-free SQP initialiser for the first MPC step (replaces NLP.solve / IPOPT, solver/nlp.py:158-217): states
 * interpolated from x0 to the origin, zero inputs, `iters` full-step nominal-QP iterations. */
int rnmpc_init_nominal(rnmpc_handle* h, int B, const double* d_x0, int iters, void* stream);

/* One SCP_SLS.solve(x0) (solver/SCP_SLS_jit.py:65-152, RTI policy): rti x [linearise, fast_SLS.solve = sls_steps x
 * (QP, eta, Q3 test, Riccati, Phi, back-offs, tightening) + final QP, nominal update].  d_x0 [B][nx]. */
int rnmpc_solve(rnmpc_handle* h, int B, const double* d_x0, void* stream);

/* SCP_SLS.reset_warm_start() (solver/SCP_SLS_jit.py:500-551): shift X, U, roll the terminal state. */
int rnmpc_reset_warm_start(rnmpc_handle* h, int B, void* stream);

/* Closed-loop plant step with the first nominal input: x <- ddyn(x, U[:,0]) + E w  (d_w may be NULL).
 * d_x [B][nx] in/out; d_u0 [B][nu] optional output. */
int rnmpc_plant_step(rnmpc_handle* h, int B, double* d_x, const double* d_w, double* d_u0, void* stream);

/* Device pointers of the resident results of the last solve (valid until the handle is destroyed).
 * name in: "X" [cap][N+1][nx], "U" [cap][N][nu], "dX", "dU", "mu" [cap][N][ni], "mu_f" [cap][ni_f], "K"
 * [cap][N][N+1][nu][nx], "beta" [cap][N][N][ni], "beta_f" [cap][N+1][ni_f], "backoff" [cap][N][ni], "backoff_f"
 * [cap][ni_f], "A" (padded rows: [cap][N][nx][nx rounded up to even]), "B", "c", "cost" [cap], "Phi_x", "Phi_u" (only after rnmpc_enable_phi), "flags" (int32 [cap]:
 * bit0 failed, bit1 Q3 early return in the last fast-SLS step, bit2 tube back-offs valid, bit3 Q3 history present,
 * bit4 K/Phi computed since the last warm-start reset), "status" (int32 [cap]),
 * "qp_iters" (int32 [cap], interior-point iterations summed over the QPs of the last solve). */
int rnmpc_get_buffer(rnmpc_handle* h, const char* name, void** d_ptr, long long* n_bytes);
int rnmpc_enable_phi(rnmpc_handle* h, int enable);
/* Synchronous device->host copy of the first n_bytes of a named resident buffer. */
int rnmpc_copy_buffer(rnmpc_handle* h, const char* name, void* h_dst, long long n_bytes);

/* ---- host-buffer entry points (what a NumPy caller such as expe/main_*_closed_loop.py uses) ------------------ */

/* solve from host x0 [B][nx]; copies back X [B][N+1][nx], U [B][N][nu], backoff [B][N][ni], backoff_f [B][ni_f],
 * flags [B], status [B] (any output pointer may be NULL).  Synchronous. */
int rnmpc_solve_host(rnmpc_handle* h, int B, const double* h_x0, double* h_X, double* h_U, double* h_backoff,
                     double* h_backoff_f, int* h_flags, int* h_status);

/* Whole closed loops (expe/main_rocket_compare_closed_loop.py:104-173 `_run_robust`) for B scenarios:
 * h_x0 [B][nx], h_W [B][steps][nx] (NULL = noise free), initial nominal h_X0/h_U0 ([B][N+1][nx], [B][N][nu]; NULL ->
 * rnmpc_init_nominal with init_iters).  Outputs (NULL to skip): h_states [B][steps+1][nx], h_inputs [B][steps][nu],
 * h_nfail [1] number of failed solves (= scenarios that died: a scenario whose solve failed is frozen for the rest of its loop and
 * is not solved again).  Synchronous; the handle's "sticky_fail" option is restored to the caller's value on every path. */
int rnmpc_closed_loop_host(rnmpc_handle* h, int B, int steps, const double* h_x0, const double* h_W, const double* h_X0,
                           const double* h_U0, int init_iters, double* h_states, double* h_inputs, long long* h_nfail);

/* Solver options: "warm_mask" (bit0: warm-start the first QP of a solve from the previous solve's multipliers, bit1: the later
 * QPs of a solve from the previous iterate; default 0 = every QP starts cold: with the centred cold start a cold QP needs 9-10
 * iterations with a maximum of 15-18, a warm-started one 10 with a tail to 30, and a launch lasts as long as its slowest
 * instance -- bit1 measured 8 % slower on the 29-step rocket loop, bit0 slower still), "sticky_fail" (1: an instance whose solve failed stays
 * frozen -- skipped by every kernel -- until rnmpc_set_nominal / rnmpc_init_nominal; default 0 = retried at every solve, as a
 * caller of the reference would), "qp_max_iter" (default 40), "qp_tol_mu" (default 1e-14), "qp_mu0" (cold start of the interior
 * point: 0 = multipliers 1; > 0 = centred, multipliers qp_mu0 / slack; < 0 = centred and scale-adaptive, multipliers
 * |qp_mu0| x mean slack / slack; default -0.3), "qp_block_max" (batches up to this size are solved by the one-CTA-per-instance QP
 * kernel with the working set in shared memory, rnmpc_qp_block.cuh: 2.9 x faster for one instance, crossover against the
 * instance-per-slot kernel near 1500 instances; default 1024, 0 = never; the two kernels agree to rounding, not bit for bit),
 * "qp_handoff_iter" (batches above qp_block_max: the instance-per-slot kernel stops after this many iterations and the instances
 * that are not done -- a launch lasts as long as its slowest instance -- are finished by the one-CTA-per-instance kernel from
 * the same iterate; default 9 = -5 % on the 29-step rocket loop, 0 = off),
 * "qp_warps" / "sls_warps" (warps per
 * CTA of the QP / fast-SLS kernel, 1..5; default 0 = automatic, currently 1: the warps are independent chains and small CTAs
 * fill the SMs evenly; 3 and 4 measured 20 % slower at batch 4096 because the grid no longer fits one wave).
 * "sls_steps" (fast-SLS iterations per SCP iteration at run time: SCP_SLS.set_fast_sls_rti_steps, SCP_SLS_jit.py:394-401, keeps the
 * handle and with it the Q3 history and K; 0 = iterate until check_convergence_socp), "sls_max_iter" (fast_SLS.MAX_ITER, fast_SLS_jit.py:206:
 * cap of that iteration, default 30), "qp_kernel" (0 automatic, 1 instance-per-slot, 2 round-1 block kernel, 3 one CTA per instance on
 * the tensor pipe, 4 one warp per instance, 5 one warp per instance with the iterate in tensor memory), "qp_warp_min" (smallest batch
 * the automatic choice gives to the one-warp-per-instance kernels, default 600), "sls_kernel" (0 tensor-pipe fast-SLS kernel, 1 slot kernel).
 * "sls_mma_warps" (warps = chains per CTA of the tensor-pipe fast-SLS kernel, 1..7, default 7), "qp_tm_ctas" (persistent CTAs of the
 * tensor-memory QP kernel per SM: 2 = all of the tensor memory, 1 for occupancy experiments: 4 instead of 8 resident instances per SM
 * measured 1.47 x slower), "qp_lpt" (1 = resident solves hand the instances to the tensor-memory kernel longest first; measured: no gain, off).
 * Diagnostics: "q3_eps" (threshold of the reference's check_convergence_socp, which hard-codes 1e-3, fast_SLS_jit.py:594), "q3_margin_reset" (resets the buffer
 * "q3_margin": per instance the closest the tested quantity came to that threshold since the reset).
 * Unknown names return -1. */
int rnmpc_set_option(rnmpc_handle* h, const char* name, double value);

/* Measured FP64 peaks of the device (TFLOP/s), for the roofline denominator: DFMA (16 independent chains per thread) and the FP64
 * tensor pipe (mma.sync.m8n8k4.f64, 8 independent chains per warp) -- the QP and fast-SLS kernels run their products on the latter. */
int rnmpc_fp64_peak(int device, double* tflops);
int rnmpc_fp64_mma_peak(int device, double* tflops);

/* Device times (ms) of the last rnmpc_solve when timing is enabled: out[0] linearisation (t_jac_ms), out[1] last QP
 * (t_qp_ms), out[2] last fast-SLS kernel (t_backward_ms); mirrors SCP_SLS_jit.py:339-341, qp_jit.py:392-395,
 * fast_SLS_jit.py:504-507.  Synchronises. */
int rnmpc_enable_timing(rnmpc_handle* h, int enable);
int rnmpc_get_timings(rnmpc_handle* h, double* out3);

/* rnmpc_enable_timing(h, 2): an event pair around EVERY launch of the three hot kernels, without any host synchronisation.
 * rnmpc_kernel_time sums (and forgets) the pairs recorded since its last call for kernel = "linearize" | "qp" | "sls":
 * total device time in ms and the number of launches (at most 65536 launches per kernel are kept between two calls).  Synchronises.  (bench.py: average launch duration of the QP kernel
 * over the timed region, for the roofline.) */
int rnmpc_kernel_time(rnmpc_handle* h, const char* kernel, double* total_ms, long long* launches);

/* Number of kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
long long rnmpc_launch_count(rnmpc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* RNMPC_B200_H */
