// Host-side engine: owns the device workspace of one handle and sequences the kernels of one SCP-SLS solve.
// One instantiation per model (explicitly instantiated in inst_*.cu so the three models compile in parallel).
#pragma once
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/rnmpc_b200.h"
#include "rnmpc_common.cuh"
#include "rnmpc_lin.cuh"
#include "rnmpc_qp.cuh"
#include "rnmpc_qp_block.cuh"
#include "rnmpc_qp_mma.cuh"
#include "rnmpc_qp_warp.cuh"
#include "rnmpc_qp_tmem.cuh"
#include "rnmpc_sls.cuh"
#include "rnmpc_sls_mma.cuh"

namespace rnmpc {

void set_error(const std::string& msg);

#define RN_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess) {                                                                        \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" +      \
                      std::to_string(__LINE__) + ")");                                                  \
            return -2;                                                                                  \
        }                                                                                               \
    } while (0)

struct IEngine {
    virtual ~IEngine() {}
    virtual int ddyn(int B, const double* x, const double* u, const double* w, double step, double* xp, cudaStream_t st) = 0;
    virtual int enable_timing(int enable) = 0;
    virtual int get_timings(double* out3) = 0;
    virtual int kernel_time(const char* kernel, double* total_ms, long long* count) = 0;
    virtual int linearize(int B, const double* X, const double* U, double* A, double* Bm, double* c, cudaStream_t st) = 0;
    virtual int sls(int B, const double* A, const double* Bm, const double* eta, const double* eta_f, double* K, double* S,
                    double* Phi_x, double* Phi_u, double* beta, double* beta_f, double* backoff, double* backoff_f,
                    cudaStream_t st) = 0;
    virtual int qp(int B, int tightened, const double* A, const double* Bm, const double* c, const double* X, const double* U,
                   const double* x0, const double* backoff, const double* backoff_f, double* dX, double* dU, double* mu,
                   double* mu_f, double* cost, int* iters, int* status, cudaStream_t st) = 0;
    virtual int qp_general(int B, const double* A, const double* Bm, const double* c, const double* hi, const double* lo,
                           const double* q, const double* x0pin, double* dX, double* dU, double* mu, double* mu_f, double* cost,
                           int* iters, int* status, cudaStream_t st) = 0;
    virtual int sls_duals(int B, const double* A, const double* Bm, const double* mu, const double* mu_f, const double* beta_in,
                          const double* beta_f_in, double* K, double* S, double* Phi_x, double* Phi_u, double* beta, double* beta_f,
                          double* backoff, double* backoff_f, cudaStream_t st) = 0;
    virtual int costates(int B, double* lam, cudaStream_t st) = 0;
    virtual int soft_nmpc(int B, const double* x0, double* Xs, double* Us, double rho, double rho1, int max_iter, double tol, double* gamma,
                          double* gamma_f, int* sqp_iters, int* status, cudaStream_t st) = 0;
    virtual int set_nominal(int B, const double* X, const double* U, cudaStream_t st) = 0;
    virtual int init_nominal(int B, const double* x0, int iters, cudaStream_t st) = 0;
    virtual int solve(int B, const double* x0, cudaStream_t st) = 0;
    virtual int reset_warm_start(int B, cudaStream_t st) = 0;
    virtual int plant_step(int B, double* x, const double* w, double* u0, cudaStream_t st) = 0;
    virtual int get_buffer(const char* name, void** ptr, long long* bytes) = 0;
    virtual int enable_phi(int enable) = 0;
    virtual int set_option(const char* name, double value) = 0;
    virtual double get_option(const char* name) = 0;
    virtual int dims(int* nx, int* nu, int* N, int* cap) = 0;
    virtual bool graphable() = 0;          // one solve is a fixed sequence of launches (no host decision inside): it can be captured in a CUDA graph
    long long launches = 0;
    long long epoch = 0;                   // bumped by every option / timing change: captured graphs of older epochs are stale
    int device = 0;
};

IEngine* make_pendulum(const rnmpc_config* cfg, int cap, int device);
IEngine* make_quadrotor(const rnmpc_config* cfg, int cap, int device);
IEngine* make_rocket(const rnmpc_config* cfg, int cap, int device);

// ---- small glue kernels ------------------------------------------------------------------------------------------
// check_convergence_socp (solver/fast_SLS_jit.py:581-600, quirk Q3): per instance max |delta - prev_delta| <= 1e-3 against
// the previous *first-in-step* QP solution, kept across MPC steps; only the primal test is used.
static __global__ void q3_check_kernel(int B, int nxs, int nus, const double* __restrict__ dX, const double* __restrict__ dU,
                                       double* __restrict__ prevX, double* __restrict__ prevU, int* __restrict__ flags, int skip_mask,
                                       double conv_eps, double* __restrict__ margin) {
    // one warp per instance: coalesced reads, shuffle max-reduction
    const int b = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    int f = flags[b];
    if (f & skip_mask) return;
    const double* x = dX + (size_t)b * nxs;
    const double* u = dU + (size_t)b * nus;
    double* px = prevX + (size_t)b * nxs;
    double* pu = prevU + (size_t)b * nus;
    double m = 0.0;
    for (int i = lane; i < nxs; i += 32) { const double v = x[i]; m = fmax(m, fabs(v - px[i])); px[i] = v; }
    for (int i = lane; i < nus; i += 32) { const double v = u[i]; m = fmax(m, fabs(v - pu[i])); pu[i] = v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        f &= ~FLAG_Q3;
        if ((f & FLAG_PREV) && m <= conv_eps) f |= FLAG_Q3;
        // closest approach of the tested quantity to its threshold since the last reset (diagnostic: a decision taken within
        // rounding of the threshold is not reproducible across implementations)
        if ((f & FLAG_PREV) && margin != nullptr) margin[b] = fmin(margin[b], fabs(m - conv_eps));
        f |= FLAG_PREV;
        flags[b] = f;
    }
}
// converge-to-tolerance mode of fast_SLS.solve (fast_SLS_jit.py:298-312): an instance whose _step returned True is finished
static __global__ void mark_sls_done_kernel(int B, int* flags, int* n_open) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int f = flags[b];
    if ((f & (FLAG_FAIL | FLAG_SLSDONE)) == 0) {
        if (f & FLAG_Q3) { f |= FLAG_SLSDONE; flags[b] = f; }
        else atomicAdd(n_open, 1);
    }
}
static __global__ void fail_open_kernel(int B, int* flags, int* status) {
    // MAX_ITER reached without convergence: the reference reports success False (fast_SLS_jit.py:310-312)
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int f = flags[b];
    if ((f & (FLAG_FAIL | FLAG_SLSDONE)) == 0) { flags[b] = f | FLAG_FAIL; status[b] = 4; }
}
// caller-provided dense A [n][nx][nx] -> padded [n][nx][ld] (pad zero)
static __global__ void pad_rows_kernel(long long n_rows, int nx, int ld, const double* __restrict__ src, double* __restrict__ dst) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n_rows * ld) return;
    const long long r = gid / ld;
    const int c = (int)(gid - r * ld);
    dst[gid] = (c < nx) ? src[r * nx + c] : 0.0;
}
static __global__ void clear_flag_kernel(int B, int* flags, int bits) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) flags[b] &= ~bits;
}
// Work order of the persistent QP kernel: instance indices sorted by a key (interior-point iterations of the previous solve),
// largest first -- a counting sort in one CTA (keys clamped to [0, 127]); the order inside a bin is arbitrary (it decides which
// warp solves an instance, never the result).
static __global__ void order_by_key_kernel(int B, const int* __restrict__ key, int* __restrict__ order) {
    __shared__ int cnt[128], start[128];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) atomicAdd(&cnt[min(max(key[b], 0), 127)], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 127; k >= 0; --k) { start[k] = acc; acc += cnt[k]; }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) order[atomicAdd(&start[min(max(key[b], 0), 127)], 1)] = b;
}
// start of a solve: clears the per-solve flag bits and counts the instances that will actually be solved (not frozen by sticky_fail)
static __global__ void begin_solve_kernel(int B, int* flags, int bits, unsigned long long* live_sum) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    int live = 0;
    if (b < B) { const int f = flags[b] & ~bits; flags[b] = f; live = (f & FLAG_FAIL) == 0; }
    const unsigned m = __ballot_sync(0xffffffffu, live);
    if ((threadIdx.x & 31) == 0 && m != 0) atomicAdd(live_sum, (unsigned long long)__popc(m));
}
static __global__ void set_tight_kernel(int B, int* flags, int skip_mask) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B && (flags[b] & skip_mask) == 0) flags[b] |= FLAG_TIGHT | FLAG_KVALID;
}
// socp_step (solver/SCP_SLS_jit.py:426-430): X += dX, U += dU for instances whose fast-SLS solve succeeded
static __global__ void apply_step_kernel(int B, int nxs, int nus, const double* __restrict__ dX, const double* __restrict__ dU,
                                         double* __restrict__ X, double* __restrict__ U, const int* __restrict__ flags) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = nxs + nus;
    if (gid >= (long long)B * per) return;
    const int b = (int)(gid / per), e = (int)(gid - (long long)b * per);
    if (flags[b] & FLAG_FAIL) return;
    if (e < nxs) X[(size_t)b * nxs + e] += dX[(size_t)b * nxs + e];
    else U[(size_t)b * nus + (e - nxs)] += dU[(size_t)b * nus + (e - nxs)];
}
static __global__ void fail_from_status_kernel(int B, const int* __restrict__ st_tmp, int* __restrict__ flags, int* __restrict__ iters_acc,
                                               const int* __restrict__ iters, int* __restrict__ st_acc, int skip_mask,
                                               unsigned long long* __restrict__ iters_sum) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (flags[b] & skip_mask) return;
    iters_acc[b] += iters[b];
    atomicAdd(iters_sum, (unsigned long long)iters[b]);      // interior-point iterations of all instances and launches since creation
    if (st_tmp[b] != 0) { flags[b] |= FLAG_FAIL; st_acc[b] = st_tmp[b]; }
}
static __global__ void interp_nominal_kernel(int B, int N, int nx, int nu, const double* __restrict__ x0, double* __restrict__ X,
                                             double* __restrict__ U) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int k = 0; k <= N; ++k)
        for (int i = 0; i < nx; ++i) X[((size_t)b * (N + 1) + k) * nx + i] = x0[(size_t)b * nx + i] * (1.0 - (double)k / (double)N);
    for (int k = 0; k < N; ++k)
        for (int p = 0; p < nu; ++p) U[((size_t)b * N + k) * nu + p] = 0.0;
}

// Multipliers of the dynamics rows of the last nominal QP (the rows `A_k dx_k + B_k du_k - dx_{k+1} = -c_k` of qp_jit.py:194-273, whose
// duals the reference returns inside `dual_vec`, qp_jit.py:491-501), recovered from the stationarity conditions of the solution the
// solve left behind: with y = x_nom + dx (the updated nominal), 2 H y + G' mu + A' lam = 0 gives
//   lam_{N-1} = 2 Qf y_N + mu_f,up - mu_f,lo ,   lam_{k-1} = 2 Q y_k + (mu_k,up - mu_k,lo)_x + A_k' lam_k   (k = N-1 .. 1).
// One thread per instance (an on-demand diagnostic output: never on the solve path).
template <int NX, int NU>
__global__ void costate_kernel(int B, int N, int lda, const double* __restrict__ A, const double* __restrict__ X,
                               const double* __restrict__ mu, const double* __restrict__ mu_f, const Params<NX, NU>* __restrict__ prm,
                               double* __restrict__ lam) {
    constexpr int NZ = NX + NU, NI = 2 * NZ;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double* Xb = X + (size_t)b * (N + 1) * NX;
    double y[NX], yn[NX];
    for (int i = 0; i < NX; ++i) {
        double g = mu_f[(size_t)b * 2 * NX + i] - mu_f[(size_t)b * 2 * NX + NX + i];
        for (int j = 0; j < NX; ++j) g = fma(2.0 * prm->Qf[i * NX + j], Xb[N * NX + j], g);
        y[i] = g;
    }
    for (int k = N - 1; k >= 0; --k) {
        double* out = lam + ((size_t)b * N + k) * NX;
        for (int i = 0; i < NX; ++i) out[i] = y[i];
        if (k == 0) break;
        const double* Ak = A + ((size_t)b * N + k) * NX * lda;
        const double* mk = mu + ((size_t)b * N + k) * NI;
        for (int i = 0; i < NX; ++i) {
            double g = mk[i] - mk[NZ + i];
            for (int j = 0; j < NX; ++j) g = fma(2.0 * prm->Q[i * NX + j], Xb[k * NX + j], g);
            for (int j = 0; j < NX; ++j) g = fma(Ak[j * lda + i], y[j], g);
            yn[i] = g;
        }
        for (int i = 0; i < NX; ++i) y[i] = yn[i];
    }
}

// One SQP iteration of the soft-constrained comparator is done: full step X += dX, U += dU of every open instance, convergence test
// max |delta| < tol (-> FLAG_SLSDONE), failed QP -> FLAG_FAIL.  One warp per instance.
static __global__ void soft_step_kernel(int B, int nxs, int nus, const double* __restrict__ dX, const double* __restrict__ dU,
                                        double* __restrict__ X, double* __restrict__ U, const int* __restrict__ qp_status, double tol,
                                        int* __restrict__ flags, int* __restrict__ sqp_iters, int* __restrict__ status, int* n_open,
                                        double* __restrict__ prox, double* __restrict__ prev_step) {
    const int b = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int f = flags[b];
    if (f & (FLAG_FAIL | FLAG_SLSDONE)) return;
    if (qp_status[b] != 0) {
        if (lane == 0) { flags[b] = f | FLAG_FAIL; status[b] = qp_status[b]; }
        return;
    }
    double m = 0.0;
    for (int i = lane; i < nxs; i += 32) { const double v = dX[(size_t)b * nxs + i]; m = fmax(m, fabs(v)); X[(size_t)b * nxs + i] += v; }
    for (int i = lane; i < nus; i += 32) { const double v = dU[(size_t)b * nus + i]; m = fmax(m, fabs(v)); U[(size_t)b * nus + i] += v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        // damping of the Gauss-Newton SQP (the QP's Hessian has no constraint curvature: with the large multipliers of violated soft
        // rows the undamped iteration can cycle): the proximal weight grows while the steps do not contract and decays while they do
        const int n_it = sqp_iters[b];
        double sg = prox[b];
        if (n_it > 0 && m > 0.8 * prev_step[b]) sg = fmax(4.0 * sg, 1.0);
        else { sg *= 0.25; if (sg < 1e-3) sg = 0.0; }
        prox[b] = sg;
        prev_step[b] = m;
        sqp_iters[b] = n_it + 1;
        if (m < tol) flags[b] = f | FLAG_SLSDONE;
        else atomicAdd(n_open, 1);
    }
}

template <class M>
__global__ void shift_kernel(int B, int N, double h, double* __restrict__ X, double* __restrict__ U) {
    // SCP_SLS.reset_warm_start (solver/SCP_SLS_jit.py:505-518): x_k <- x_{k+1}, u_k <- u_{k+1}, u_{N-1} kept,
    // x_N <- ddyn(x_N, u_{N-1}) evaluated on the un-shifted trajectory.  One warp per instance: every lane evaluates the (cheap) RK4 step,
    // the shift is a coalesced read-all-then-write of the trajectory (one thread per instance took 70 us for 4096 instances: 64 CTAs).
    constexpr int NX = M::NX, NU = M::NU;
    const int b = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    double* Xb = X + (size_t)b * (N + 1) * NX;
    double* Ub = U + (size_t)b * N * NU;
    double xN[NX], uN[NU], xo[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xN[i] = Xb[N * NX + i];
#pragma unroll
    for (int p = 0; p < NU; ++p) uN[p] = Ub[(N - 1) * NU + p];
    rk4_step<M, double>(xN, uN, h, xo);
    constexpr int CH = 4;                       // elements per lane and round
    const int nxs = N * NX, nus = (N - 1) * NU;
    for (int e0 = 0; e0 < nxs; e0 += 32 * CH) {
        double v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) { const int e = e0 + 32 * j + lane; v[j] = e < nxs ? Xb[e + NX] : 0.0; }
        __syncwarp();                           // a round's reads (up to e0 + 32 CH + NX) precede its writes (below e0 + 32 CH); later rounds read further up
#pragma unroll
        for (int j = 0; j < CH; ++j) { const int e = e0 + 32 * j + lane; if (e < nxs) Xb[e] = v[j]; }
        __syncwarp();
    }
    for (int e0 = 0; e0 < nus; e0 += 32 * CH) {
        double v[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) { const int e = e0 + 32 * j + lane; v[j] = e < nus ? Ub[e + NU] : 0.0; }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < CH; ++j) { const int e = e0 + 32 * j + lane; if (e < nus) Ub[e] = v[j]; }
        __syncwarp();
    }
#pragma unroll
    for (int i = 0; i < NX; ++i)
        if (lane == (i & 31)) Xb[N * NX + i] = xo[i];
}

template <class M>
struct Engine : IEngine {
    static constexpr int NX = M::NX, NU = M::NU, NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX;
    static constexpr int LDA = Slot<NX>::LD;       // row stride of the resident (padded) A_k
    using P = Params<NX, NU>;
    int N, cap, rti, sls_steps, qp_max_iter;
    double qp_tol_mu, qp_tol_res, h;
    int qp_block_max = 1024;       // batches up to this size use the one-CTA-per-instance QP kernel (rnmpc_qp_block.cuh); 0: never.  Measured
                                   // (tools/iter_stats.py, 29-step rocket loop, ms; block kernel vs qp_ipm_kernel with its hand-off):
                                   // 768 instances 157.4 vs 199.5, 1024: 198.3 vs 216.5, 1536: 276.5 vs 256.2
    size_t qp_block_smem = 0;
    double q3_eps = 1e-3;          // threshold of check_convergence_socp (fast_SLS_jit.py:594 hard-codes 1e-3; the drivers' CONV_EPS is not used
                                   // there): the Q3 early return fires when the nominal QP solution moved by less since the previous forward solve
    double* q3_margin = nullptr;   // [B] diagnostic: min |tested quantity - q3_eps| since "q3_margin_reset"
    int qp_handoff_iter = 9;       // large batches: qp_ipm_kernel stops after this many iterations and hands the unfinished instances to
                                   // qp_block_kernel (QpArgs::handoff_iter); 0: off.  29-step rocket loop at batch 4096 (ms): off 499.6,
                                   // 14: 501.3, 12: 497.3, 11: 487.2, 10: 479.3, 9: 473.1, 8: 483.3, 7: 503.9, 6: 529.0 (the mean iteration
                                   // count there is 9; tuned on that loop only)
    int *handoff_count = nullptr, *handoff_list = nullptr;
    double* handoff_state = nullptr;
    bool qp_block_ok = false;
    // QP kernel selection ("qp_kernel"): 0 automatic = a tensor-pipe kernel whenever one applies (diagonal cost, cold start, reference
    // bounds, working set fits shared memory): the one-warp-per-instance throughput kernel (rnmpc_qp_warp.cuh) for batches of at least
    // qp_warp_min instances, the one-CTA-per-instance latency kernel (rnmpc_qp_mma.cuh) below; 1 = instance-per-slot kernel (with its
    // hand-off), 2 = the round-1 one-CTA-per-instance kernel (rnmpc_qp_block.cuh), 3 = rnmpc_qp_mma.cuh at every batch size,
    // 4 = rnmpc_qp_warp.cuh at every batch size, 5 = rnmpc_qp_tmem.cuh (one warp per instance, iterate in tensor memory; the
    // automatic choice for large batches when the horizon fits its 16 stage columns, else rnmpc_qp_warp.cuh).  1 and 2 remain for comparison.
    int qp_kernel = 0;
    bool qp_tm_ok = false;
    size_t qp_tm_smem = 0;
    int* qp_work = nullptr;        // [1] instance counter of qp_tm_kernel's persistent warps
    int* qp_order = nullptr;       // [cap] work order of qp_tm_kernel (order_by_key_kernel), valid for the resident solve
    bool qp_order_valid = false;
    int qp_tm_ctas = 2;            // "qp_tm_ctas": persistent CTAs of qp_tm_kernel per SM (2 = all the tensor memory; 1 for occupancy experiments)
    int qp_lpt = 0;                // "qp_lpt": 1 = resident solves hand the instances to qp_tm_kernel longest first (previous solve's iterations).
                                   // Off: measured 285.0 ms with it against 278.7 ms without on the 29-step rocket loop at batch 4096 -- the previous
                                   // solve's count predicts little and the scattered instance order costs locality
    int n_sm = 148;
    int qp_warp_min = 600;         // "qp_warp_min": smallest batch the automatic selection gives to the one-warp-per-instance kernel
    bool qp_warp_ok = false;
    size_t qp_warp_smem = 0;
    double* Kw = nullptr;          // [B][N][NU][LDA] Riccati gains of qp_warp_kernel (its own padded layout)
    double* Wimg = nullptr;        // [B][N][SLOT] sweep images of qp_tm_kernel: A_k | B_k | K_k per stage, one bulk copy per ring slot
    int sls_kernel_sel = 0;        // "sls_kernel": 0 = tensor-pipe fast-SLS kernel (rnmpc_sls_mma.cuh, one warp per chain), 1 = round-1 slot kernel
    int sls_mma_warps = 7;         // warps (chains) per CTA of the tensor-pipe fast-SLS kernel ("sls_mma_warps")
    int dense_reg = 0;             // Q_reg has off-diagonal entries
    long long* qp_prof = nullptr;  // "qp_profile": per-phase cycle counters of instance 0 in qp_mma_kernel (buffer "qp_prof", 16 x int64)
    bool qp_mma_ok = false;
    size_t qp_mma_smem = 0;
    double qp_mu0 = -0.3;          // cold start of the interior point: centred, multipliers = 0.3 x mean slack / slack (see QpArgs::mu0)
    P* d_prm = nullptr;
    std::map<std::string, std::pair<void*, size_t>> bufs;
    bool phi_enabled = false;
    int diag_cost = 0;
    bool timing = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // timing mode 2: an event pair around EVERY launch of the three hot kernels, kept until kernel_time() sums them -- no host
    // synchronisation while a region is being timed
    int timing_mode = 0;
    struct EvPair { cudaEvent_t a, b; };
    std::vector<EvPair> ev_log[3];      // 0 linearise, 1 QP, 2 fast-SLS
    std::vector<EvPair> ev_pool;
    static constexpr size_t kMaxLoggedLaunches = 1 << 16;
    cudaEvent_t region_begin(int which, cudaStream_t st) {
        if (timing_mode != 2 || ev_log[which].size() >= kMaxLoggedLaunches) return nullptr;      // bounded: nobody may be collecting
        EvPair p;
        if (!ev_pool.empty()) { p = ev_pool.back(); ev_pool.pop_back(); }
        else if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return nullptr;
        ev_log[which].push_back(p);
        cudaEventRecord(p.a, st);
        return p.b;
    }
    static void region_end(cudaEvent_t e, cudaStream_t st) { if (e) cudaEventRecord(e, st); }

    double *X, *U, *A, *Bm, *c, *dX, *dU, *mu, *mu_f, *beta, *beta_f, *backoff, *backoff_f, *K, *cost, *prevX, *prevU;
    double *Phi_x = nullptr, *Phi_u = nullptr;
    double *wE, *wG, *rp, *Sc, *dff, *Kq, *Lq;
    int *flags, *status, *qp_iters, *iters_tmp, *status_tmp, *n_open, *ws_valid;
    unsigned long long* iters_sum;
    unsigned long long* live_sum;      // instances solved (not frozen) summed over all rnmpc_solve calls since creation
    bool have_solution = false, shifted_since_solve = false;   // host-side warm-start bookkeeping
    int warm_enabled = 1;
    int qp_warps = 0, sls_warps = 0;   // warps per CTA of the slot kernels (0: automatic, see cta_warps)
    int warm_mask = 0;             // warm starts are OFF by default.  Measured (tools/iter_stats.py / iter_tail.py, B = 4096, 29 steps): with the
                                   // centred cold start every QP needs 9-10 iterations (max 15-18); a warm-started tightened QP needs 10.1 on
                                   // average with a tail to 30 (slow or restarted warm starts), and a launch lasts as long as its slowest
                                   // instance: 572.8 ms with bit1, 530.6 ms with none.  Bit0 (across MPC steps) was slower still.
    bool sticky_fail = false;
    int sls_max_iter = 30;             // fast_SLS.MAX_ITER (fast_SLS_jit.py:206)

    template <class T>
    int alloc(const char* name, T*& ptr, size_t count) {
        RN_CUDA(cudaMalloc((void**)&ptr, count * sizeof(T)));
        RN_CUDA(cudaMemset(ptr, 0, count * sizeof(T)));
        bufs[name] = {(void*)ptr, count * sizeof(T)};
        return 0;
    }

    int init(const rnmpc_config* cfg, int cap_, int device_) {
        device = device_;
        RN_CUDA(cudaSetDevice(device));
        N = cfg->N; cap = cap_; rti = cfg->rti; sls_steps = cfg->sls_steps; h = cfg->h;
        qp_max_iter = cfg->qp_max_iter > 0 ? cfg->qp_max_iter : 40;
        qp_tol_mu = cfg->qp_tol_mu > 0 ? cfg->qp_tol_mu : 1e-14;
        qp_tol_res = cfg->qp_tol_res > 0 ? cfg->qp_tol_res : 1e-9;
        P hp;
        memcpy(hp.Q, cfg->Q, sizeof(hp.Q)); memcpy(hp.R, cfg->R, sizeof(hp.R)); memcpy(hp.Qf, cfg->Qf, sizeof(hp.Qf));
        for (int i = 0; i < NX * NX; ++i) hp.Q2[i] = 2.0 * hp.Q[i];
        for (int i = 0; i < NU * NU; ++i) hp.R2[i] = 2.0 * hp.R[i];
        memcpy(hp.Qreg, cfg->Q_reg, sizeof(hp.Qreg)); memcpy(hp.Rreg, cfg->R_reg, sizeof(hp.Rreg));
        memcpy(hp.Qregf, cfg->Q_reg_f, sizeof(hp.Qregf)); memcpy(hp.E, cfg->E, sizeof(hp.E));
        memcpy(hp.xub, cfg->x_ub, sizeof(hp.xub)); memcpy(hp.xlb, cfg->x_lb, sizeof(hp.xlb));
        memcpy(hp.uub, cfg->u_ub, sizeof(hp.uub)); memcpy(hp.ulb, cfg->u_lb, sizeof(hp.ulb));
        memcpy(hp.xubf, cfg->xf_ub, sizeof(hp.xubf)); memcpy(hp.xlbf, cfg->xf_lb, sizeof(hp.xlbf));
        hp.h = h;
        diag_cost = 1;
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j)
                if (i != j && (hp.Q[i * NX + j] != 0.0 || hp.Qf[i * NX + j] != 0.0)) diag_cost = 0;
        for (int i = 0; i < NU; ++i)
            for (int j = 0; j < NU; ++j)
                if (i != j && hp.R[i * NU + j] != 0.0) diag_cost = 0;
        for (int i = 0; i < NX; ++i)
            for (int j = 0; j < NX; ++j)
                if (i != j && hp.Qreg[i * NX + j] != 0.0) dense_reg = 1;
        RN_CUDA(cudaMalloc((void**)&d_prm, sizeof(P)));
        RN_CUDA(cudaMemcpy(d_prm, &hp, sizeof(P), cudaMemcpyHostToDevice));
        const size_t B = cap, n1 = N + 1;
        int rc = 0;
        rc |= alloc("X", X, B * n1 * NX); rc |= alloc("U", U, B * N * NU);
        rc |= alloc("A", A, B * N * NX * LDA); rc |= alloc("B", Bm, B * N * NX * NU); rc |= alloc("c", c, B * N * NX);
        rc |= alloc("dX", dX, B * n1 * NX); rc |= alloc("dU", dU, B * N * NU);
        rc |= alloc("mu", mu, B * N * NI); rc |= alloc("mu_f", mu_f, B * NIF);
        rc |= alloc("beta", beta, B * N * N * NI); rc |= alloc("beta_f", beta_f, B * n1 * NIF);
        rc |= alloc("backoff", backoff, B * N * NI); rc |= alloc("backoff_f", backoff_f, B * NIF);
        rc |= alloc("K", K, B * N * n1 * NU * NX); rc |= alloc("cost", cost, B);
        rc |= alloc("prevX", prevX, B * n1 * NX); rc |= alloc("prevU", prevU, B * N * NU);
        rc |= alloc("wE", wE, B * n1 * NZ * kRec); rc |= alloc("wG", wG, B * n1 * NZ);
        rc |= alloc("rp", rp, B * N * NX); rc |= alloc("Sc", Sc, B * N * NX); rc |= alloc("dff", dff, B * N * NU);
        rc |= alloc("Kq", Kq, B * N * NU * NX); rc |= alloc("Lq", Lq, B * N * NU * NU);
        rc |= alloc("flags", flags, B); rc |= alloc("status", status, B); rc |= alloc("qp_iters", qp_iters, B);
        rc |= alloc("iters_tmp", iters_tmp, B); rc |= alloc("status_tmp", status_tmp, B); rc |= alloc("n_open", n_open, 1); rc |= alloc("iters_sum", iters_sum, 1); rc |= alloc("live_sum", live_sum, 1);
        rc |= alloc("q3_margin", q3_margin, B); rc |= alloc("handoff_count", handoff_count, 1); rc |= alloc("handoff_list", handoff_list, B); rc |= alloc("handoff_state", handoff_state, B * 4); rc |= alloc("ws_valid", ws_valid, B);
        if (rc) return -2;
        const int smem_sls = (int)sizeof(SlotSmem<NX, NU>) * Slot<NX>::SPW * kWarps;
        RN_CUDA(cudaFuncSetAttribute(sls_kernel<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_sls));
        RN_CUDA(cudaFuncSetAttribute(sls_mma_kernel<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * SlsMma<NX, NU>::PER_WARP * (int)sizeof(double)));
        RN_CUDA(cudaFuncSetAttribute(sls_mma_kernel<NX, NU, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 7 * SlsMma<NX, NU>::PER_WARP * (int)sizeof(double)));
        qp_block_smem = qp_block_smem_doubles<NX, NU>(N) * sizeof(double);
        // the attribute belongs to the kernel, not to this handle (handles with different horizons share it): always the maximum
        qp_block_ok = qp_block_smem <= 227 * 1024 &&
                      cudaFuncSetAttribute(qp_block_kernel<NX, NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        cudaGetLastError();
        qp_mma_smem = qp_mma_smem_doubles<NX, NU>(N) * sizeof(double);
        qp_mma_ok = qp_mma_smem <= 227 * 1024 &&
                    cudaFuncSetAttribute(qp_mma_kernel<NX, NU, kHorizonFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                    cudaFuncSetAttribute(qp_mma_kernel<NX, NU, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                    cudaFuncSetAttribute(qp_mma_kernel<NX, NU, kHorizonFast, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                    cudaFuncSetAttribute(qp_mma_kernel<NX, NU, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        cudaGetLastError();
        {
            cudaDeviceProp prop;
            if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) n_sm = prop.multiProcessorCount;
        }
        qp_tm_smem = qp_tm_smem_bytes<NX, NU>(N);
        qp_tm_ok = N + 1 <= kTmMaxStages && qp_tm_smem <= 227 * 1024 && alloc("qp_work", qp_work, 1) == 0 && alloc("qp_order", qp_order, B) == 0 &&
                   alloc("Wimg", Wimg, (size_t)B * N * QpWarp<NX, NU>::SLOT) == 0 &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, kHorizonFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, kHorizonFast>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, kHorizonFast, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, kHorizonFast, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess &&
                   cudaFuncSetAttribute(qp_tm_kernel<NX, NU, 0, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100) == cudaSuccess;
        cudaGetLastError();
        qp_warp_smem = qp_warp_smem_doubles<NX, NU>(N) * sizeof(double);
        qp_warp_ok = qp_warp_smem <= 227 * 1024 && alloc("Kw", Kw, B * N * NU * LDA) == 0 &&
                     cudaFuncSetAttribute(qp_warp_kernel<NX, NU, kHorizonFast>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess &&
                     cudaFuncSetAttribute(qp_warp_kernel<NX, NU, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) == cudaSuccess;
        cudaGetLastError();
        RN_CUDA(cudaFuncSetAttribute(qp_ipm_kernel<NX, NU, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_sls));
        RN_CUDA(cudaFuncSetAttribute(qp_ipm_kernel<NX, NU, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_sls));
        RN_CUDA(cudaFuncSetAttribute(qp_ipm_kernel<NX, NU, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_sls));
        RN_CUDA(cudaFuncSetAttribute(qp_ipm_kernel<NX, NU, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_sls));
        return 0;
    }
    ~Engine() override {
        cudaSetDevice(device);
        for (auto& kv : bufs) cudaFree(kv.second.first);
        if (d_prm) cudaFree(d_prm);
        for (auto& e : ev) if (e) cudaEventDestroy(e);
        for (auto& log : ev_log) for (auto& p : log) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
        for (auto& p : ev_pool) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
    }

    static constexpr int kHorizonFast = 15;   // N of every reference driver (expe/main_*: N = 15): kernels instantiated with it as a constant
    static constexpr int kWarps = 5;   // warps per CTA of the slot kernels (160 threads; 2 CTAs = 10 warps per SM by shared memory)

    int dims(int* nx, int* nu, int* N_, int* cap_) override {
        if (nx) *nx = NX; if (nu) *nu = NU; if (N_) *N_ = N; if (cap_) *cap_ = cap;
        return 0;
    }
    int check_B(int B) {
        if (B <= 0 || B > cap) { set_error("batch " + std::to_string(B) + " outside (0, capacity=" + std::to_string(cap) + "]"); return -1; }
        return 0;
    }

    int ddyn(int B, const double* x, const double* u, const double* w, double step, double* xp, cudaStream_t st) override {
        if (B <= 0) return -1;
        ddyn_kernel<M><<<(B + 127) / 128, 128, 0, st>>>(B, x, NX, u, NU, w, d_prm->E, step > 0 ? step : h, xp);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    int linearize(int B, const double* X_, const double* U_, double* A_, double* B_, double* c_, cudaStream_t st) override {
        return linearize_ld(B, X_, U_, A_, NX, B_, c_, st);      // caller's dense layout
    }
    int linearize_ld(int B, const double* X_, const double* U_, double* A_, int lda, double* B_, double* c_, cudaStream_t st) {
        if (B <= 0) return -1;
        const long long n = (long long)B * N * (NZ - M::NPAS);      // one thread per differentiated column (rnmpc_lin.cuh)
        cudaEvent_t te = region_begin(0, st);
        linearize_kernel<M><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(B, N, X_, U_, h, A_, lda, B_, c_);
        region_end(te, st);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    // Warps per CTA of the slot kernels.  Every warp is an independent sequential chain (no CTA-level synchronisation), so the
    // CTA size only decides how the warps spread over the SMs: small CTAs put the warps of a batch that does not fill the
    // GPU on different SMs / scheduler partitions instead of stacking kWarps of them on a few SMs.
    int cta_warps(int option, long long warps) const {
        if (option >= 1 && option <= kWarps) return option;
        (void)warps;
        return 1;
    }
    int launch_sls(int B, const SlsArgs& a, cudaStream_t st) {
        // the kernel stages A_k / B_k with 16-byte asynchronous copies
        if ((reinterpret_cast<uintptr_t>(a.A) | reinterpret_cast<uintptr_t>(a.Bm)) & 15) { set_error("sls: A and B must be 16-byte aligned"); return -1; }
        const long long items = (long long)B * (N / 2 + 1);
        const long long warps = (items + Slot<NX>::SPW - 1) / Slot<NX>::SPW;
        const int w = cta_warps(sls_warps, warps);
        const unsigned grid = (unsigned)((warps + w - 1) / w);
        const int smem = (int)sizeof(SlotSmem<NX, NU>) * Slot<NX>::SPW * w;
        cudaEvent_t te = region_begin(2, st);
        if (sls_kernel_sel == 0) {
            // one warp per chain, every product of a stage on the FP64 tensor pipe (rnmpc_sls_mma.cuh)
            const int wpc = sls_mma_warps < 1 ? 1 : (sls_mma_warps > 7 ? 7 : sls_mma_warps);
            const unsigned gridm = (unsigned)((items + wpc - 1) / wpc);
            const bool lean = a.S == nullptr && a.Phi_x == nullptr && a.Phi_u == nullptr && a.eta == nullptr && a.eta_f == nullptr && !dense_reg;
            if (lean) sls_mma_kernel<NX, NU, true><<<gridm, wpc * 32, wpc * SlsMma<NX, NU>::PER_WARP * sizeof(double), st>>>(a, d_prm, dense_reg);
            else sls_mma_kernel<NX, NU><<<gridm, wpc * 32, wpc * SlsMma<NX, NU>::PER_WARP * sizeof(double), st>>>(a, d_prm, dense_reg);
        } else sls_kernel<NX, NU><<<grid, w * 32, smem, st>>>(a, d_prm);
        region_end(te, st);
        ++launches;
        RN_CUDA(cudaGetLastError());
        const long long n = (long long)B * ((long long)N * NI + NIF);
        return 0 * n;
    }
    int launch_backoff(int B, const double* beta_, const double* beta_f_, const int* fl, int skip_mask, double* bo, double* bof,
                       cudaStream_t st) {
        const long long n = (long long)B * ((long long)N * NI + NIF);
        backoff_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(B, N, NI, NIF, beta_, beta_f_, fl, skip_mask, bo, bof);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    double* Apad = nullptr;
    // caller-provided dense A -> padded scratch copy (no-op layout-wise when nx is even)
    const double* padded(int B, const double* A_, cudaStream_t st) {
        if (LDA == NX) return A_;
        if (Apad == nullptr) {
            if (cudaMalloc((void**)&Apad, (size_t)cap * N * NX * LDA * sizeof(double)) != cudaSuccess) return nullptr;
            bufs["Apad"] = {(void*)Apad, (size_t)cap * N * NX * LDA * sizeof(double)};
        }
        const long long n = (long long)B * N * NX * LDA;
        pad_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((long long)B * N * NX, NX, LDA, A_, Apad);
        ++launches;
        return Apad;
    }
    int sls(int B, const double* A_, const double* B_, const double* eta, const double* eta_f, double* K_, double* S_, double* Phi_x_,
            double* Phi_u_, double* beta_, double* beta_f_, double* backoff_, double* backoff_f_, cudaStream_t st) override {
        if (check_B(B)) return -1;
        A_ = padded(B, A_, st);
        if (A_ == nullptr) { set_error("out of memory"); return -2; }
        SlsArgs a{};
        a.B = B; a.N = N; a.A = A_; a.Bm = B_; a.eta = eta; a.eta_f = eta_f;
        a.K = K_; a.S = S_; a.Phi_x = Phi_x_; a.Phi_u = Phi_u_; a.beta = beta_; a.beta_f = beta_f_;
        int rc = launch_sls(B, a, st);
        if (rc) return rc;
        return launch_backoff(B, beta_, beta_f_, nullptr, 0, backoff_, backoff_f_, st);
    }
    int launch_qp(int B, int tightened, const double* A_, const double* B_, const double* c_, const double* X_, const double* U_,
                  const double* x0, const double* bo, const double* bof, const int* fl, int skip_mask, double* dX_, double* dU_,
                  double* mu_, double* mu_f_, double* cost_, int* iters_, int* status_, cudaStream_t st, const double* hi_in = nullptr,
                  const double* lo_in = nullptr, const double* q_lin = nullptr, int warm_mode = 0) {
        // the kernel stages A_k / B_k with 16-byte asynchronous copies
        if ((reinterpret_cast<uintptr_t>(A_) | reinterpret_cast<uintptr_t>(B_)) & 15) { set_error("qp: A and B must be 16-byte aligned"); return -1; }
        QpArgs a{};
        a.warm = warm_enabled ? warm_mode : 0; a.ws_valid = ws_valid;
        a.hi_in = hi_in; a.lo_in = lo_in; a.q_lin = q_lin;
        a.B = B; a.N = N; a.tightened = tightened; a.max_iter = qp_max_iter; a.tol_mu = qp_tol_mu; a.tol_res = qp_tol_res; a.mu0 = qp_mu0;
        a.A = A_; a.Bm = B_; a.c = c_; a.Xnom = X_; a.Unom = U_; a.x0 = x0; a.backoff = bo; a.backoff_f = bof;
        a.flags = fl; a.skip_mask = skip_mask;
        a.dX = dX_; a.dU = dU_; a.mu = mu_; a.mu_f = mu_f_; a.cost = cost_; a.iters = iters_; a.status = status_;
        a.E = wE; a.G = wG; a.diag_cost = diag_cost; a.rp = rp; a.Sc = Sc;
        a.dff = dff; a.Kq = Kq; a.Lq = Lq; a.prof = qp_prof;
        const long long warps = ((long long)B + Slot<NX>::SPW - 1) / Slot<NX>::SPW;
        const int w = cta_warps(qp_warps, warps);
        const unsigned grid = (unsigned)((warps + w - 1) / w);
        const int smem = (int)sizeof(SlotSmem<NX, NU>) * Slot<NX>::SPW * w;
        cudaEvent_t te = region_begin(1, st);
        const bool may_warm = a.warm != 0;
        const bool plain = diag_cost && !may_warm && hi_in == nullptr && lo_in == nullptr && q_lin == nullptr;
        if (plain && qp_tm_ok && qp_warp_ok && (qp_kernel == 5 || (qp_kernel == 0 && B >= qp_warp_min))) {
            // one warp per instance, iterate in tensor memory, persistent CTAs of four warps (rnmpc_qp_tmem.cuh)
            a.Kq = Wimg; a.work = qp_work;
            a.order = (qp_order_valid && fl == flags) ? qp_order : nullptr;      // only the resident solve has a history
            cudaMemsetAsync(qp_work, 0, sizeof(int), st);
            const unsigned gridt = (unsigned)std::min<long long>((long long)qp_tm_ctas * n_sm, ((long long)B + kTmWarps - 1) / kTmWarps);
            if (a.prof != nullptr) {          // phase-timer instantiation (option "qp_profile")
                if (N == kHorizonFast) qp_tm_kernel<NX, NU, kHorizonFast, true><<<gridt, 32 * kTmWarps, qp_tm_smem, st>>>(a, d_prm);
                else qp_tm_kernel<NX, NU, 0, true><<<gridt, 32 * kTmWarps, qp_tm_smem, st>>>(a, d_prm);
            } else if (N == kHorizonFast) qp_tm_kernel<NX, NU, kHorizonFast><<<gridt, 32 * kTmWarps, qp_tm_smem, st>>>(a, d_prm);
            else qp_tm_kernel<NX, NU, 0><<<gridt, 32 * kTmWarps, qp_tm_smem, st>>>(a, d_prm);
        } else if (plain && qp_warp_ok && (qp_kernel == 4 || qp_kernel == 5 || (qp_kernel == 0 && B >= qp_warp_min))) {
            // one warp (one CTA of 32 threads) per instance: the throughput kernel for large batches (rnmpc_qp_warp.cuh)
            a.Kq = Kw;
            if (N == kHorizonFast) qp_warp_kernel<NX, NU, kHorizonFast><<<(unsigned)B, 32, qp_warp_smem, st>>>(a, d_prm);
            else qp_warp_kernel<NX, NU, 0><<<(unsigned)B, 32, qp_warp_smem, st>>>(a, d_prm);
        } else if (plain && qp_mma_ok && (qp_kernel == 0 || qp_kernel == 3 || qp_kernel == 4)) {
            // one CTA per instance, factorisation on the FP64 tensor pipe, iterate resident in shared memory (rnmpc_qp_mma.cuh)
            // (the horizon of all reference drivers is a compile-time constant of the fast instantiation; any other horizon runs the generic one)
            if (a.prof != nullptr) {          // phase-timer instantiation (option "qp_profile")
                if (N == kHorizonFast) qp_mma_kernel<NX, NU, kHorizonFast, true><<<(unsigned)B, kMmaThreads, qp_mma_smem, st>>>(a, d_prm);
                else qp_mma_kernel<NX, NU, 0, true><<<(unsigned)B, kMmaThreads, qp_mma_smem, st>>>(a, d_prm);
            } else if (N == kHorizonFast) qp_mma_kernel<NX, NU, kHorizonFast><<<(unsigned)B, kMmaThreads, qp_mma_smem, st>>>(a, d_prm);
            else qp_mma_kernel<NX, NU, 0><<<(unsigned)B, kMmaThreads, qp_mma_smem, st>>>(a, d_prm);
        } else if (qp_block_ok && (B <= qp_block_max || qp_kernel == 2) && qp_kernel != 1 && diag_cost && !may_warm && hi_in == nullptr && lo_in == nullptr && q_lin == nullptr) {
            // small batch: one CTA per instance, working set in shared memory (rnmpc_qp_block.cuh)
            qp_block_kernel<NX, NU><<<(unsigned)B, kBlockThreads, qp_block_smem, st>>>(a, d_prm);
        } else if (diag_cost) {
            if (may_warm) qp_ipm_kernel<NX, NU, true, true><<<grid, w * 32, smem, st>>>(a, d_prm);
            else if (qp_block_ok && qp_handoff_iter > 0 && hi_in == nullptr && lo_in == nullptr && q_lin == nullptr) {
                // instance-per-slot kernel up to qp_handoff_iter iterations, then the stragglers on one CTA each (resume mode)
                cudaMemsetAsync(handoff_count, 0, sizeof(int), st);
                a.handoff_iter = qp_handoff_iter; a.handoff_count = handoff_count; a.handoff_list = handoff_list; a.handoff_state = handoff_state;
                qp_ipm_kernel<NX, NU, true, false><<<grid, w * 32, smem, st>>>(a, d_prm);
                a.resume = 1;
                qp_block_kernel<NX, NU><<<(unsigned)B, kBlockThreads, qp_block_smem, st>>>(a, d_prm);
                ++launches;
            } else qp_ipm_kernel<NX, NU, true, false><<<grid, w * 32, smem, st>>>(a, d_prm);
        } else {
            if (may_warm) qp_ipm_kernel<NX, NU, false, true><<<grid, w * 32, smem, st>>>(a, d_prm);
            else qp_ipm_kernel<NX, NU, false, false><<<grid, w * 32, smem, st>>>(a, d_prm);
        }
        region_end(te, st);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    int qp(int B, int tightened, const double* A_, const double* B_, const double* c_, const double* X_, const double* U_,
           const double* x0, const double* bo, const double* bof, double* dX_, double* dU_, double* mu_, double* mu_f_, double* cost_,
           int* iters_, int* status_, cudaStream_t st) override {
        if (check_B(B)) return -1;
        A_ = padded(B, A_, st);
        if (A_ == nullptr) { set_error("out of memory"); return -2; }
        int rc = launch_qp(B, tightened ? 1 : 0, A_, B_, c_, X_, U_, x0, bo, bof, nullptr, 0, dX_, dU_, mu_, mu_f_, cost_, iters_, status_, st);
        if (!rc) RN_CUDA(cudaMemsetAsync(ws_valid, 0, (size_t)B * sizeof(int), st));   // the workspace no longer belongs to the resident solve
        return rc;
    }

    int qp_general(int B, const double* A_, const double* B_, const double* c_, const double* hi_, const double* lo_, const double* q_,
                   const double* x0pin, double* dX_, double* dU_, double* mu_, double* mu_f_, double* cost_, int* iters_, int* status_,
                   cudaStream_t st) override {
        if (check_B(B)) return -1;
        A_ = padded(B, A_, st);
        if (A_ == nullptr) { set_error("out of memory"); return -2; }
        // zero nominal: dX is the decision variable itself, the pinned value is passed as "x0"
        RN_CUDA(cudaMemsetAsync(dXa_zero(), 0, (size_t)B * ((N + 1) * NX + N * NU) * sizeof(double), st));
        const double* Xz = dXa_zero();
        const double* Uz = Xz + (size_t)B * (N + 1) * NX;
        int rc = launch_qp(B, 0, A_, B_, c_, Xz, Uz, x0pin, nullptr, nullptr, nullptr, 0, dX_, dU_, mu_, mu_f_, cost_, iters_, status_, st,
                           hi_, lo_, q_);
        if (!rc) RN_CUDA(cudaMemsetAsync(ws_valid, 0, (size_t)B * sizeof(int), st));
        return rc;
    }
    double* zero_buf = nullptr;
    double* dXa_zero() {
        if (zero_buf == nullptr) {
            cudaMalloc((void**)&zero_buf, (size_t)cap * ((N + 1) * NX + N * NU) * sizeof(double));
            bufs["zero_buf"] = {(void*)zero_buf, (size_t)cap * ((N + 1) * NX + N * NU) * sizeof(double)};
        }
        return zero_buf;
    }
    int sls_duals(int B, const double* A_, const double* B_, const double* mu_, const double* mu_f_, const double* beta_in_,
                  const double* beta_f_in_, double* K_, double* S_, double* Phi_x_, double* Phi_u_, double* beta_, double* beta_f_,
                  double* backoff_, double* backoff_f_, cudaStream_t st) override {
        if (check_B(B)) return -1;
        A_ = padded(B, A_, st);
        if (A_ == nullptr) { set_error("out of memory"); return -2; }
        SlsArgs a{};
        a.B = B; a.N = N; a.A = A_; a.Bm = B_; a.mu = mu_; a.mu_f = mu_f_; a.beta_in = beta_in_; a.beta_f_in = beta_f_in_;
        a.K = K_; a.S = S_; a.Phi_x = Phi_x_; a.Phi_u = Phi_u_; a.beta = beta_; a.beta_f = beta_f_;
        int rc = launch_sls(B, a, st);
        if (rc) return rc;
        return launch_backoff(B, beta_, beta_f_, nullptr, 0, backoff_, backoff_f_, st);
    }
    int costates(int B, double* lam, cudaStream_t st) override {
        if (check_B(B)) return -1;
        costate_kernel<NX, NU><<<(B + 63) / 64, 64, 0, st>>>(B, N, LDA, A, X, mu, mu_f, d_prm, lam);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    // Soft-constrained NMPC comparator (solver/nlp_soft_constraints.py:175-276, expe/main_rocket_compare_closed_loop.py:176-222): the NLP with
    // every inequality row softened, solved by full-step SQP from the caller's trajectory guess (Xs, Us; in / out): linearise, soft QP
    // (qp_warp_kernel<.., SOFT>), step, until max |delta| < tol per instance (one 4-byte device -> host read per iteration).
    int* soft_flags = nullptr;
    double *soft_prox = nullptr, *soft_prev = nullptr;
    size_t soft_smem = 0;
    int soft_nmpc(int B, const double* x0, double* Xs, double* Us, double rho, double rho1, int max_iter, double tol, double* gamma,
                  double* gamma_f, int* sqp_iters, int* status_, cudaStream_t st) override {
        if (check_B(B)) return -1;
        if (!diag_cost) { set_error("soft_nmpc: diagonal cost matrices only"); return -1; }
        if (soft_flags == nullptr) {
            if (alloc("soft_flags", soft_flags, cap) || alloc("soft_prox", soft_prox, cap) || alloc("soft_prev", soft_prev, cap)) return -2;
            soft_smem = (qp_warp_smem_doubles<NX, NU>(N) + qp_warp_soft_extra_doubles<NX, NU>(N)) * sizeof(double);
            if (soft_smem > 227 * 1024 ||
                cudaFuncSetAttribute(qp_warp_kernel<NX, NU, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                set_error("soft_nmpc: the working set does not fit shared memory for this horizon"); return -1;
            }
        }
        RN_CUDA(cudaMemsetAsync(soft_flags, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(soft_prox, 0, (size_t)B * sizeof(double), st));
        RN_CUDA(cudaMemsetAsync(sqp_iters, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(status_, 0, (size_t)B * sizeof(int), st));
        for (int it = 0; it < max_iter; ++it) {
            int rc = linearize_ld(B, Xs, Us, A, LDA, Bm, c, st);
            if (rc) return rc;
            QpArgs a{};
            a.B = B; a.N = N; a.tightened = 0; a.max_iter = qp_max_iter + 20; a.tol_mu = 1e-16; a.tol_res = qp_tol_res; a.mu0 = qp_mu0;      // (terminal slacks sit at sqrt(mu / 2 rho) on the central path)
            a.A = A; a.Bm = Bm; a.c = c; a.Xnom = Xs; a.Unom = Us; a.x0 = x0;
            a.flags = soft_flags; a.skip_mask = FLAG_FAIL | FLAG_SLSDONE;
            a.dX = dX; a.dU = dU; a.mu = mu; a.mu_f = mu_f; a.cost = cost; a.iters = iters_tmp; a.status = status_tmp;
            a.diag_cost = 1; a.Kq = Kw; a.Lq = Lq; a.rho = rho; a.rho1 = rho1; a.gamma = gamma; a.gamma_f = gamma_f; a.prox = soft_prox;
            qp_warp_kernel<NX, NU, 0, true><<<(unsigned)B, 32, soft_smem, st>>>(a, d_prm);
            ++launches;
            RN_CUDA(cudaMemsetAsync(n_open, 0, sizeof(int), st));
            soft_step_kernel<<<(B + 3) / 4, 128, 0, st>>>(B, (N + 1) * NX, N * NU, dX, dU, Xs, Us, status_tmp, tol, soft_flags, sqp_iters, status_, n_open, soft_prox, soft_prev);
            ++launches;
            int open_host = 0;
            RN_CUDA(cudaMemcpyAsync(&open_host, n_open, sizeof(int), cudaMemcpyDeviceToHost, st));
            RN_CUDA(cudaStreamSynchronize(st));
            if (open_host == 0) break;
        }
        // instances still open after max_iter: not converged
        fail_open_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, soft_flags, status_);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    int set_nominal(int B, const double* X_, const double* U_, cudaStream_t st) override {
        if (check_B(B)) return -1;
        RN_CUDA(cudaMemcpyAsync(X, X_, (size_t)B * (N + 1) * NX * sizeof(double), cudaMemcpyDeviceToDevice, st));
        RN_CUDA(cudaMemcpyAsync(U, U_, (size_t)B * N * NU * sizeof(double), cudaMemcpyDeviceToDevice, st));
        RN_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(ws_valid, 0, (size_t)B * sizeof(int), st));
        have_solution = false; shifted_since_solve = false;
        return 0;
    }

    // one QP on the resident trajectories + failure bookkeeping
    int resident_qp(int B, int tightened, const double* x0, cudaStream_t st, int skip_mask = FLAG_FAIL, int warm_mode = 0) {
        int rc = launch_qp(B, tightened, A, Bm, c, X, U, x0, backoff, backoff_f, flags, skip_mask, dX, dU, mu, mu_f, cost, iters_tmp,
                           status_tmp, st, nullptr, nullptr, nullptr, warm_mode);
        if (rc) return rc;
        fail_from_status_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, status_tmp, flags, qp_iters, iters_tmp, status, skip_mask, iters_sum);
        ++launches;
        RN_CUDA(cudaGetLastError());
        return 0;
    }

    int init_nominal(int B, const double* x0, int iters, cudaStream_t st) override {
        if (check_B(B)) return -1;
        interp_nominal_kernel<<<(B + 127) / 128, 128, 0, st>>>(B, N, NX, NU, x0, X, U);
        ++launches;
        RN_CUDA(cudaMemsetAsync(flags, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(status, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(qp_iters, 0, (size_t)B * sizeof(int), st));
        RN_CUDA(cudaMemsetAsync(ws_valid, 0, (size_t)B * sizeof(int), st));
        have_solution = false; shifted_since_solve = false;
        for (int it = 0; it < iters; ++it) {
            int rc = linearize_ld(B, X, U, A, LDA, Bm, c, st);
            if (rc) return rc;
            rc = resident_qp(B, 0, x0, st, FLAG_FAIL, it > 0 ? 1 : 0);   // multipliers of the previous SQP iterate
            if (rc) return rc;
            apply_step_kernel<<<(unsigned)(((long long)B * ((N + 1) * NX + N * NU) + 255) / 256), 256, 0, st>>>(B, (N + 1) * NX, N * NU, dX, dU, X, U, flags);
            ++launches;
        }
        RN_CUDA(cudaGetLastError());
        return 0;
    }

    int solve(int B, const double* x0, cudaStream_t st) override {
        if (check_B(B)) return -1;
        // a new solve: failure flags of the previous solve are cleared, the Q3 history (FLAG_PREV) is kept
        begin_solve_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, (sticky_fail ? 0 : FLAG_FAIL) | FLAG_Q3 | FLAG_TIGHT, live_sum);
        ++launches;
        if (!sticky_fail) RN_CUDA(cudaMemsetAsync(status, 0, (size_t)B * sizeof(int), st));   // frozen instances keep their failure status
        qp_order_valid = false;
        if (qp_lpt && qp_tm_ok && have_solution && B >= qp_warp_min) {
            order_by_key_kernel<<<1, 1024, 0, st>>>(B, qp_iters, qp_order);      // before the counters of the previous solve are cleared
            ++launches;
            qp_order_valid = true;
        }
        RN_CUDA(cudaMemsetAsync(qp_iters, 0, (size_t)B * sizeof(int), st));
        if (timing) cudaEventRecord(ev[0], st);
        int rc = linearize_ld(B, X, U, A, LDA, Bm, c, st);             // update_jacobian (SCP_SLS_jit.py:80)
        if (timing) cudaEventRecord(ev[1], st);
        if (rc) return rc;
        const bool converge_mode = sls_steps <= 0;
        const int skipf = FLAG_FAIL | FLAG_SLSDONE;
        for (int ii = 0; ii < rti; ++ii) {
            // fast_SLS.solve (fast_SLS_jit.py:278-312): initialize_backoff, then either exactly sls_steps x _step + a final
            // forward_solve (RTI cap), or _step until check_convergence_socp returns True (cap MAX_ITER, then failure)
            clear_flag_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, FLAG_Q3 | FLAG_TIGHT | FLAG_SLSDONE);
            ++launches;
            const int n_steps = converge_mode ? sls_max_iter : sls_steps;
            bool all_done = false;
            for (int i = 0; i < n_steps && !all_done; ++i) {
                // warm start: first QP of a solve <- multipliers of the previous solve (shifted if reset_warm_start ran); later QPs of
                // the same linearisation <- previous iterate and multipliers (only the bounds moved)
                const int wm = (i > 0) ? ((warm_mask & 2) ? (1 | 4) : 0)
                                       : ((warm_mask & 1) ? (ii > 0 ? 1 : (have_solution ? (shifted_since_solve ? (1 | 2) : 1) : 0)) : 0);
                rc = resident_qp(B, -1, x0, st, skipf, wm);            // forward_solve
                if (rc) return rc;
                q3_check_kernel<<<(B + 3) / 4, 128, 0, st>>>(B, (N + 1) * NX, N * NU, dX, dU, prevX, prevU, flags, skipf, q3_eps, q3_margin);
                ++launches;
                if (converge_mode) {
                    RN_CUDA(cudaMemsetAsync(n_open, 0, sizeof(int), st));
                    mark_sls_done_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, n_open);
                    ++launches;
                    int open_host = 0;
                    RN_CUDA(cudaMemcpyAsync(&open_host, n_open, sizeof(int), cudaMemcpyDeviceToHost, st));
                    RN_CUDA(cudaStreamSynchronize(st));
                    if (open_host == 0) { all_done = true; break; }
                }
                SlsArgs a{};                                            // evaluate_dual_eta + backward_solve + update_tightening
                a.B = B; a.N = N; a.A = A; a.Bm = Bm; a.mu = mu; a.mu_f = mu_f; a.beta_in = beta; a.beta_f_in = beta_f;
                a.flags = flags; a.skip_mask = skipf | FLAG_Q3; a.beta_from_flags = 1;
                a.K = K; a.Phi_x = phi_enabled ? Phi_x : nullptr; a.Phi_u = phi_enabled ? Phi_u : nullptr;
                a.beta = beta; a.beta_f = beta_f;
                if (timing) cudaEventRecord(ev[4], st);
                rc = launch_sls(B, a, st);
                if (timing) cudaEventRecord(ev[5], st);
                if (rc) return rc;
                rc = launch_backoff(B, beta, beta_f, flags, skipf | FLAG_Q3, backoff, backoff_f, st);
                if (rc) return rc;
                set_tight_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, skipf | FLAG_Q3);
                ++launches;
            }
            if (timing) cudaEventRecord(ev[2], st);
            if (!converge_mode) {
                rc = resident_qp(B, -1, x0, st, FLAG_FAIL, (warm_mask & 2) ? (1 | 4) : 0);      // final forward_solve
            } else {
                // instances still open after MAX_ITER: one last QP with the latest bounds, then failure (fast_SLS_jit.py:310-312)
                rc = all_done ? 0 : resident_qp(B, -1, x0, st, skipf, (warm_mask & 2) ? (1 | 4) : 0);
                fail_open_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, status);
                ++launches;
            }
            if (timing) cudaEventRecord(ev[3], st);
            if (rc) return rc;
            apply_step_kernel<<<(unsigned)(((long long)B * ((N + 1) * NX + N * NU) + 255) / 256), 256, 0, st>>>(B, (N + 1) * NX, N * NU, dX, dU, X, U, flags);   // socp_step
            ++launches;
            if (ii + 1 < rti) {
                rc = linearize_ld(B, X, U, A, LDA, Bm, c, st);          // update_jacobian (SCP_SLS_jit.py:139); the one after
                if (rc) return rc;                                      // the last iteration is dead work and is not run
            }
        }
        have_solution = true; shifted_since_solve = false;
        RN_CUDA(cudaGetLastError());
        return 0;
    }

    int reset_warm_start(int B, cudaStream_t st) override {
        if (check_B(B)) return -1;
        shift_kernel<M><<<(unsigned)(((long long)B * 32 + 127) / 128), 128, 0, st>>>(B, N, h, X, U);
        ++launches;
        clear_flag_kernel<<<(B + 255) / 256, 256, 0, st>>>(B, flags, FLAG_KVALID);   // fast_SLS.reset_solver_to_zeros drops K / Phi
        ++launches;
        shifted_since_solve = true;
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    int plant_step(int B, double* x, const double* w, double* u0, cudaStream_t st) override {
        if (check_B(B)) return -1;
        ddyn_kernel<M><<<(B + 127) / 128, 128, 0, st>>>(B, x, NX, U, N * NU, w, d_prm->E, h, x);
        ++launches;
        if (u0 != nullptr)
            RN_CUDA(cudaMemcpy2DAsync(u0, NU * sizeof(double), U, (size_t)N * NU * sizeof(double), NU * sizeof(double), B,
                                      cudaMemcpyDeviceToDevice, st));
        RN_CUDA(cudaGetLastError());
        return 0;
    }
    int enable_timing(int enable) override {
        if (enable && ev[0] == nullptr)
            for (auto& e : ev) RN_CUDA(cudaEventCreate(&e));
        timing = enable != 0;
        timing_mode = enable;
        ++epoch;
        return 0;
    }
    int kernel_time(const char* kernel, double* total_ms, long long* count) override {
        const std::string n(kernel);
        const int which = n == "linearize" ? 0 : n == "qp" ? 1 : n == "sls" ? 2 : -1;
        if (which < 0) { set_error("kernel_time: unknown kernel " + n); return -1; }
        double tot = 0.0;
        long long cnt = 0;
        for (auto& p : ev_log[which]) {
            float ms = 0.f;
            if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { tot += ms; ++cnt; }
            ev_pool.push_back(p);
        }
        ev_log[which].clear();
        cudaGetLastError();
        *total_ms = tot; *count = cnt;
        return 0;
    }
    int get_timings(double* out3) override {
        out3[0] = out3[1] = out3[2] = -1.0;
        if (!timing) return 0;
        for (int i = 0; i < 3; ++i) {
            if (cudaEventQuery(ev[2 * i + 1]) == cudaErrorInvalidResourceHandle) continue;
            cudaEventSynchronize(ev[2 * i + 1]);
            float ms = -1.f;
            if (cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]) == cudaSuccess) out3[i] = ms;
        }
        cudaGetLastError();
        return 0;
    }
    int get_buffer(const char* name, void** ptr, long long* bytes) override {
        auto it = bufs.find(name);
        if (it == bufs.end()) { set_error(std::string("unknown buffer ") + name); return -1; }
        *ptr = it->second.first;
        if (bytes) *bytes = (long long)it->second.second;
        return 0;
    }
    // (not while phase timing is on: events recorded by the nodes of a replayed graph give no elapsed times -- measured: -1)
    bool graphable() override { return sls_steps > 0 && !timing && timing_mode == 0 && warm_mask == 0; }
    int set_option(const char* name, double value) override {
        const std::string n(name);
        ++epoch;
        if (n == "qp_warps") { qp_warps = (int)value; return 0; }
        if (n == "sls_warps") { sls_warps = (int)value; return 0; }
        if (n == "warm_mask") warm_mask = (int)value;            // bit0: first QP of a solve from the previous solve, bit1: later QPs of a solve
        else if (n == "sticky_fail") sticky_fail = value != 0;    // failed instances stay frozen until set_nominal / init_nominal
        else if (n == "qp_max_iter") qp_max_iter = (int)value;
        else if (n == "qp_tol_mu") qp_tol_mu = value;
        else if (n == "qp_mu0") qp_mu0 = value;
        else if (n == "qp_block_max") qp_block_max = (int)value;
        else if (n == "sls_max_iter") { if (value < 1) { set_error("sls_max_iter must be >= 1"); return -1; } sls_max_iter = (int)value; }
        else if (n == "sls_steps") { if (value < 0) { set_error("sls_steps must be >= 0"); return -1; } sls_steps = (int)value; }
        else if (n == "qp_kernel") qp_kernel = (int)value;
        else if (n == "qp_warp_min") qp_warp_min = (int)value;
        else if (n == "qp_lpt") qp_lpt = (int)value;
        else if (n == "qp_tm_ctas") qp_tm_ctas = value >= 2 ? 2 : 1;
        else if (n == "sls_kernel") sls_kernel_sel = (int)value;
        else if (n == "sls_mma_warps") sls_mma_warps = (int)value;
        else if (n == "qp_profile") {
            if (value != 0 && qp_prof == nullptr) { if (alloc("qp_prof", qp_prof, 16)) return -2; }
            else if (qp_prof != nullptr) cudaMemset(qp_prof, 0, 16 * sizeof(long long));
            if (value == 0) qp_prof = nullptr;
        }
        else if (n == "q3_eps") q3_eps = value;
        else if (n == "q3_margin_reset") { if (q3_margin) { std::vector<double> big((size_t)cap, 1e300); cudaMemcpy(q3_margin, big.data(), big.size() * sizeof(double), cudaMemcpyHostToDevice); } }
        else if (n == "qp_handoff_iter") qp_handoff_iter = (int)value;
        else { set_error("unknown option " + n); return -1; }
        return 0;
    }
    double get_option(const char* name) override {
        const std::string n(name);
        if (n == "sticky_fail") return sticky_fail ? 1.0 : 0.0;
        if (n == "sls_steps") return sls_steps;
        if (n == "sls_max_iter") return sls_max_iter;
        if (n == "qp_kernel") return qp_kernel;
        return 0.0;
    }
    int enable_phi(int enable) override {
        if (enable && Phi_x == nullptr) {
            RN_CUDA(cudaSetDevice(device));
            if (alloc("Phi_x", Phi_x, (size_t)cap * (N + 1) * (N + 1) * NX * NX)) return -2;
            if (alloc("Phi_u", Phi_u, (size_t)cap * N * (N + 1) * NU * NX)) return -2;
        }
        phi_enabled = enable != 0;
        return 0;
    }
};

template <class M>
IEngine* make_engine(const rnmpc_config* cfg, int cap, int device) {
    auto* e = new Engine<M>();
    if (e->init(cfg, cap, device) != 0) { delete e; return nullptr; }
    return e;
}

}  // namespace rnmpc
