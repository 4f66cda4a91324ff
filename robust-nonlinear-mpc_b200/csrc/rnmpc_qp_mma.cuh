// Nominal QP, tensor-pipe variant: ONE CTA (4 warps) PER INSTANCE, the Riccati factorisation of every interior-point iteration on
// the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA), the three vector sweeps as chains of small matrix-vector products on one warp,
// A_k, B_k, K_k, the Cholesky factors and the whole iterate resident in shared memory: NO global-memory access inside the iteration.  Same problem, same algorithm, same formulas and status codes as
// qp_ipm_kernel / qp_block_kernel (rnmpc_qp.cuh, rnmpc_qp_block.cuh; reference: solver/qp_jit.py:77-273, 362-402, 487-513).
//
// Why (measured, profiles/r1i_ncu_full_summary.csv, profiles/r2_dmma_probe.txt):
//   * qp_ipm_kernel keeps all instances of a batch in flight (4096 x ~100 KB = 3 x the L2) and re-streams that working set four
//     times per iteration: 11.8 GB of HBM traffic per launch for 0.26 GB of inputs + outputs, 80 % of the time on the long scoreboard.
//     One CTA per instance bounds the number of instances in flight to 2 per SM (30 MB of inputs: L2 resident) and keeps the
//     iterate on chip: HBM traffic is the compulsory traffic.
//   * qp_block_kernel does that already, but with one dot product per thread (2 shared-memory loads per DFMA, 4036 cycles per
//     Riccati stage).  Here a stage is two tile products on the tensor pipe:
//         W = S_{k+1} [A_k B_k rp_k]                       (M = NX, N = NX + NU + 1, K = NX)
//         M = [A_k B_k]^T [W  p_{k+1}]                     (upper triangle of tiles only: the result is symmetric)
//     which yield A'SA, A'SB, B'SB and, in the spare columns, every vector product of the predictor's backward recursion
//     (S rp, A'S rp, B'S rp, A'p, B'p).  A DMMA is 256 FMA per warp instruction, 16 cycles of the pipe, 26 cycles dependent latency.
//   * The operand fragments of [A_k B_k] are the same registers in both products (B operand of the first, A operand of the second:
//     Z[4s + t][8n + g]); they come straight from global memory (L2 resident), one stage ahead.
//   * The rank-NU update S_k = A'SA + (A'SB) K is one more k-step on the same pipe (accumulator initialised with the tile of A'SA).
//   * With 2 x 106 KB of shared memory per SM the L1 carve-out is nil: a global load inside the iteration costs an L2 round trip
//     (second version of this kernel, profiles/r2_qp_mma_v2_phases.txt: 1200-1600 cycles per phase of a stage where 200-450 were
//     expected).  A_k and B_k are therefore copied to shared memory once per QP (the launch's compulsory traffic) and every
//     fragment, tile and matrix-vector operand of the iteration comes from there.
//   * Instruction footprint: a first version of this kernel was 118 KB of SASS (every shared-memory array address recomputed from
//     the run-time horizon at each use, predictor and corrector code duplicated) and ran an interior-point iteration in 124 k cycles,
//     no faster than qp_block_kernel.  The horizon is a template parameter now (NC; 0 = run-time fallback), predictor and corrector
//     share one copy of the sweep code, loops over stages / entries are not unrolled.
// Padding: NX = 17 -> tiles of 8 rows / columns, K in steps of 4; padded rows and columns are kept at zero in shared memory.
#pragma once
#include "rnmpc_common.cuh"
#include "rnmpc_qp.cuh"

namespace rnmpc {

constexpr int kMmaThreads = 128;

// D (8 x 8) += A (8 x 4) B (4 x 8); lane (g = lane / 4, t = lane % 4) holds A[g][t], B[t][g], D[g][2t], D[g][2t + 1]
// (checked on the device: tools/dmma_probe.cu)
#ifndef RNMPC_EMU
__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

template <int NX, int NU>
struct QpMma {
    static constexpr int NZ = NX + NU;
    static constexpr int KP = (NX + 3) / 4 * 4, KS = KP / 4;      // inner dimension padded to k-steps of 4
    static constexpr int CB = (NX + 1) & ~1;                      // columns of Z = [A | B | rp]: A at 0, B at CB, rp at CRP
    static constexpr int CRP = CB + NU;
    static constexpr int CP = CRP + 1;                            // column of W that carries p_{k+1} (virtual: read from the vector)
    static constexpr int NT = (CP + 8) / 8;                       // column tiles covering 0 .. CP
    static constexpr int MTS = (NX + 7) / 8;                      // row tiles of S (first product)
    static constexpr int MTM = (CB + NU + 7) / 8;                 // row tiles of M (second product): rows = columns of [A | B]
    static constexpr int ld4(int n) { return n + ((4 - n % 8) + 8) % 8; }   // smallest L >= n with L = 4 (mod 8): fragment loads of a
                                                                            // row-major matrix with that stride are bank-conflict free
    static constexpr int LDS = ld4(KP), LDW = ld4(8 * NT);
    static constexpr int SROWS = 8 * MTS, WROWS = (8 * MTS > KP ? 8 * MTS : KP), MROWS = 8 * MTM;
    static constexpr int LDA = (NX + 1) & ~1;                     // row stride of the stored closed-loop matrices and gains
    static_assert(MTS <= 3, "one warp per row tile of S, one warp left for the stage scalars");
    static_assert(NU <= 4, "the rank-NU updates of a stage (S + (A'SB) K, A + B K) are one k-step of the tensor pipe");
    static_assert(MTM <= NT, "fragments of Z are shared by the two products");
};

__host__ __device__ inline size_t mma_ev(size_t n) { return (n + 1) & ~(size_t)1; }

// Operand fragments of Z_k = [A_k | B_k | rp_k] straight from memory: zf[s][n] = Z[4s + t][8n + g] (rows >= NX and unused columns
// zero).  A column tile that lies inside A_k is a strided read-only load.  A tile that mixes the last columns of A_k, the pad, B_k and
// the vector rp_k would make the lanes of a warp take three different branches per load (measured: 10 % of the stall samples of
// qp_warp_kernel, "branch resolving" 0.8 warps per issue); instead every lane keeps ONE generic pointer per mixed tile (global A_k or
// B_k, or the shared-memory vector) with its own stage and k-step strides, so a fragment is one predicated load on every lane.
// STAGE > 0: A_k and B_k are parts of one image per stage with that stride (qp_tm_kernel's sweep image, written by the kernel itself:
// coherent loads); STAGE = 0: the engine's separate arrays (read-only path).
// PLAIN: ordinary (coherent, any address space) loads instead of the read-only path -- sources the kernel wrote itself or keeps in
// shared memory (qp_mma_kernel's resident copies of A_k, B_k).
template <int NX, int NU, int NT, int KS, bool RP, int STAGE = 0, bool PLAIN = (STAGE > 0)>
struct ZFrag {
    static constexpr int CB = (NX + 1) & ~1, CRP = CB + NU, LDG = (NX + 1) & ~1;
    static constexpr int STA = STAGE > 0 ? STAGE : NX * LDG, STB = STAGE > 0 ? STAGE : NX * NU;
    static constexpr int NPURE = (NX / 8) < NT ? (NX / 8) : NT;      // tiles 0 .. NPURE-1 hold columns of A_k only
    static constexpr int NMIX = NT - NPURE > 0 ? NT - NPURE : 1;
    const double* base[NMIX];     // element (row t, this lane's column) of stage 0
    int ks[NMIX], ss[NMIX];       // doubles between stages / between k-steps (4 rows)
    bool ok[NMIX];
    int g, t4;
    __device__ __forceinline__ void init(const double* Ag, const double* Bg, const double* rp, const int lane) {
        g = lane >> 2; t4 = lane & 3;
#pragma unroll
        for (int j = 0; j < NMIX; ++j) {
            const int c = 8 * (NPURE + j) + g;
            base[j] = Ag; ks[j] = 0; ss[j] = 0; ok[j] = false;
            if (NPURE + j < NT) {
                if (c < NX) { base[j] = Ag + t4 * LDG + c; ks[j] = STA; ss[j] = 4 * LDG; ok[j] = true; }
                else if (c >= CB && c < CB + NU) { base[j] = Bg + t4 * NU + (c - CB); ks[j] = STB; ss[j] = 4 * NU; ok[j] = true; }
                else if (RP && c == CRP) { base[j] = rp + t4; ks[j] = NX; ss[j] = 4; ok[j] = true; }
            }
        }
    }
    __device__ __forceinline__ void load(const int k, const double* Ag, double (&zf)[KS][NT]) const {
        const double* Ak = Ag + (size_t)k * STA + t4 * LDG + g;
#pragma unroll
        for (int s = 0; s < KS; ++s) {
            const bool row = 4 * s + t4 < NX;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                double v = 0.0;
                if (n < NPURE) { if (row) v = PLAIN ? Ak[4 * s * LDG + 8 * n] : __ldg(Ak + 4 * s * LDG + 8 * n); }
                else { const int j = n - NPURE; if (row && ok[j]) v = base[j][(long long)k * ks[j] + s * ss[j]]; }
                zf[s][n] = v;
            }
        }
    }
};

// shared-memory footprint in doubles
template <int NX, int NU>
__host__ __device__ inline size_t qp_mma_smem_doubles(int N) {
    using D = QpMma<NX, NU>;
    const size_t NE = (size_t)(N + 1) * D::NZ;
    return mma_ev((size_t)N * NX * D::LDA) + mma_ev((size_t)N * NX * NU) + mma_ev((size_t)N * NU * D::LDA) + mma_ev((size_t)N * NU * NU)   // A, B, K, L
         + 10 * mma_ev(NE)                                                                                     // entry arrays
         + 3 * mma_ev((size_t)N * NX) + 2 * mma_ev((size_t)N * NU)                                             // rp, Sc, gv, dff, tu
         + mma_ev(D::SROWS * D::LDS) + mma_ev(D::WROWS * D::LDW) + mma_ev(D::MROWS * D::LDW)                   // S, W, M
         + mma_ev(D::KP) + 2 * mma_ev(D::NZ) + mma_ev(D::NZ + NX) + mma_ev(NU) + mma_ev(NU * NU) + 64;         // pv, rhs, dgn, qdt, uc, R2, red
}

template <int NX, int NU, int NC, bool PROF = false>          // PROF: phase timer (option "qp_profile") as its own instantiation
__global__ void __launch_bounds__(kMmaThreads, 2) qp_mma_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using D = QpMma<NX, NU>;
    constexpr int NZ = D::NZ, NI = 2 * NZ, NIF = 2 * NX, T = kMmaThreads, LDG = Slot<NX>::LD;
    constexpr int KP = D::KP, KS = D::KS, CB = D::CB, CRP = D::CRP, CP = D::CP, NT = D::NT, MTS = D::MTS, MTM = D::MTM;
    constexpr int LDS = D::LDS, LDW = D::LDW, LDA = D::LDA;
    static_assert(NZ <= 32, "one lane per state / input entry in the sweeps");
    static_assert(LDA == LDG, "the resident copy of A_k has the layout of the engine's padded A");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (b >= a.B) return;
    if (a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) return;
    // The warps' roles rotate with the instance index so that the DMMA-heavy roles of the CTAs sharing an SM land on
    // different scheduler partitions (warp w of every CTA sits on partition w).
    const int lane = threadIdx.x & 31;
    const int role = ((threadIdx.x >> 5) + b) & 3;
    const int rt = role * 32 + lane;               // logical thread index: work is distributed by rt, not threadIdx.x
    const int g = lane >> 2, t4 = lane & 3;
    const int N = NC > 0 ? NC : a.N;               // compile-time horizon: every shared-memory address below is a constant

    double* sp = reinterpret_cast<double*>(smem_raw);
    auto take = [&](size_t n) { double* p = sp; sp += mma_ev(n); return p; };
    const int NE = (N + 1) * NZ;
    double* As = take((size_t)N * NX * LDA); double* Bs = take((size_t)N * NX * NU);
    double* Kq = take((size_t)N * NU * LDA); double* Lq = take((size_t)N * NU * NU);
    double* e_su = take(NE); double* e_sl = take(NE); double* e_lu = take(NE); double* e_ll = take(NE);
    double* e_iu = take(NE); double* e_il = take(NE); double* e_z = take(NE); double* e_zn = take(NE);
    double* e_da = take(NE); double* e_dc = take(NE);
    double* rp = take((size_t)N * NX); double* Sc = take((size_t)N * NX); double* gv = take((size_t)N * NX);
    double* dff = take((size_t)N * NU); double* tu = take((size_t)N * NU);
    double* S = take(D::SROWS * LDS); double* W = take(D::WROWS * LDW); double* M = take(D::MROWS * LDW);
    double* pv = take(KP); double* rhs = take(NZ); double* dgn = take(NZ); double* qdt = take(NZ + NX); double* uc = take(NU);
    double* R2s = take(NU * NU);
    double* red = take(64);

    const double* Ag = a.A + (size_t)b * N * NX * LDG;
    const double* Bg = a.Bm + (size_t)b * N * NX * NU;
    const double* cg = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;

    // block-wide reduction of four values at once: v0, v1 by maximum, v2, v3 by sum (two barriers).  The partial results are
    // combined in the order of the roles: the same value in every thread, independent of the role rotation.
    auto block_reduce4 = [&](double& v0, double& v1, double& v2, double& v3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
            v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
        __syncthreads();
        if (lane == 0) { red[4 * role] = v0; red[4 * role + 1] = v1; red[4 * role + 2] = v2; red[4 * role + 3] = v3; }
        __syncthreads();
        v0 = fmax(fmax(red[0], red[4]), fmax(red[8], red[12]));
        v1 = fmax(fmax(red[1], red[5]), fmax(red[9], red[13]));
        v2 = ((red[2] + red[6]) + red[10]) + red[14];
        v3 = ((red[3] + red[7]) + red[11]) + red[15];
    };

    // diagnostic phase timer (option "qp_profile"): cycles of the logical thread 0 of instance 0 between consecutive marks
    const bool prof_on = PROF && a.prof != nullptr && b == 0 && rt == 0;
    long long* pacc = reinterpret_cast<long long*>(red + 40);
    long long t_last = 0;
    if (prof_on) {
        for (int i = 0; i < 16; ++i) pacc[i] = 0;
        t_last = clock64();
    }
    auto mark = [&](const int slot) {
        if (PROF && prof_on) { const long long t = clock64(); pacc[slot] += t - t_last; t_last = t; }
    };

    // ------------------------------------------------------------------ stage matrices -> shared memory (once per QP), padded buffers zeroed
    {
        const double2* A2 = reinterpret_cast<const double2*>(Ag);
        double2* As2 = reinterpret_cast<double2*>(As);
#pragma unroll 4
        for (int e = rt; e < N * NX * LDA / 2; e += T) As2[e] = __ldg(A2 + e);
#pragma unroll 4
        for (int e = rt; e < N * NX * NU; e += T) Bs[e] = __ldg(Bg + e);
    }
    for (int e = rt; e < D::SROWS * LDS; e += T) S[e] = 0.0;
    for (int e = rt; e < D::WROWS * LDW; e += T) W[e] = 0.0;
    for (int e = rt; e < D::MROWS * LDW; e += T) M[e] = 0.0;
    if (rt < KP) pv[rt] = 0.0;
    if (rt < NU * NU) R2s[rt] = prm->R2[rt];
    if (rt < NZ) qdt[rt] = 2.0 * (rt < NX ? prm->Q[rt * NX + rt] : prm->R[(rt - NX) * NU + (rt - NX)]);
    else if (rt < NZ + NX) qdt[rt] = 2.0 * prm->Qf[(rt - NZ) * NX + (rt - NZ)];
    if (rt == 0) red[32] = 0.0;        // "a pivot of the factorisation was not positive" flag of the current iteration
    auto qd_of = [&](const int k, const int i) -> double { return (k < N) ? qdt[i] : (i < NX ? qdt[NZ + i] : 0.0); };

    // ------------------------------------------------------------------ bounds, initial point, slacks (rnmpc_qp.cuh, cold start)
    const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);
    double bad = 0.0, n_box = 0.0, s_sum = 0.0, dmy = 0.0;
    {
        const double eps = tight ? 0.0 : kEpsQP;
#pragma unroll 1
        for (int idx = rt; idx < NE; idx += T) {
            const int k = idx / NZ, i = idx - k * NZ;
            double h_ = 1e300, l_ = -1e300, z = 0.0, zn = 0.0;
            bool box = false;
            if (i < NX) {
                const double xn = Xn[k * NX + i];
                zn = xn;
                if (k < N) {
                    const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                    const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                    h_ = (prm->xub[i] - xn) - bu + eps;
                    l_ = -((xn - prm->xlb[i]) - bl + eps);
                } else if (tight) {
                    h_ = prm->xubf[i] - a.backoff_f[(size_t)b * NIF + i];              // Q4: terminal bound not shifted by x_N
                    l_ = -(-prm->xlbf[i] - a.backoff_f[(size_t)b * NIF + NX + i]);
                } else {
                    h_ = (prm->xubf[i] - xn) + eps;
                    l_ = -((xn - prm->xlbf[i]) + eps);
                }
                if (k == 0) {
                    z = a.x0[(size_t)b * NX + i] - xn;                                  // x(0) is pinned to the measurement
                    if (z > h_ + kX0Tol || z < l_ - kX0Tol) bad = 1.0;
                } else box = true;
            } else if (k < N) {
                const int p = i - NX;
                const double un = Un[k * NU + p];
                zn = un;
                const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                h_ = (prm->uub[p] - un) - bu + eps;
                l_ = -((un - prm->ulb[p]) - bl + eps);
                box = true;
            }
            double su = 1e300, sl = 1e300;
            if (box) {
                const bool hu = h_ < kBig, hl = l_ > -kBig;
                if (hu && hl) {
                    const double m = 0.01 * (h_ - l_);
                    if (!(h_ - l_ > 1e-12)) bad = 1.0;
                    z = fmin(fmax(0.0, l_ + m), h_ - m);
                } else if (hu) z = fmin(0.0, h_ - 1.0);
                else if (hl) z = fmax(0.0, l_ + 1.0);
                if (hu) { su = h_ - z; n_box += 1.0; s_sum += su; }
                if (hl) { sl = z - l_; n_box += 1.0; s_sum += sl; }
            }
            e_su[idx] = su; e_sl[idx] = sl;
            e_iu[idx] = (su < kBig) ? 1.0 / su : 0.0; e_il[idx] = (sl < kBig) ? 1.0 / sl : 0.0;
            e_z[idx] = z; e_zn[idx] = zn; e_da[idx] = 0.0; e_dc[idx] = 0.0;
        }
    }
    block_reduce4(bad, dmy, n_box, s_sum);
    int status = bad > 0.0 ? 2 : 0;
    {   // centred cold start: multipliers mu_bar / slack (QpArgs::mu0)
        const double mu_bar = a.mu0 < 0.0 ? (n_box > 0.0 ? -a.mu0 * s_sum / n_box : 1.0) : a.mu0;
#pragma unroll 1
        for (int idx = rt; idx < NE; idx += T) {
            double lu, ll;
            if (a.mu0 != 0.0) { lu = mu_bar * e_iu[idx]; ll = mu_bar * e_il[idx]; }
            else { lu = e_iu[idx] > 0.0 ? 1.0 : 0.0; ll = e_il[idx] > 0.0 ? 1.0 : 0.0; }
            e_lu[idx] = lu; e_ll[idx] = ll;
        }
    }
    __syncthreads();
    // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
    double rp_inf = 0.0, mu_acc = 0.0;
#pragma unroll 1
    for (int e = rt; e < N * NX; e += T) {
        const int k = e / NX, r = e - k * NX;
        const double* Ar = As + ((size_t)k * NX + r) * LDA;
        const double* Br = Bs + ((size_t)k * NX + r) * NU;
        double acc = cg[e] - e_z[(k + 1) * NZ + r], acc2 = 0.0;
#pragma unroll
        for (int m = 0; m + 1 < NX; m += 2) { acc = fma(Ar[m], e_z[k * NZ + m], acc); acc2 = fma(Ar[m + 1], e_z[k * NZ + m + 1], acc2); }
        if (NX & 1) acc = fma(Ar[NX - 1], e_z[k * NZ + NX - 1], acc);
#pragma unroll
        for (int p = 0; p < NU; ++p) acc = fma(Br[p], e_z[k * NZ + NX + p], acc);
        acc += acc2;
        rp[e] = acc;
        rp_inf = fmax(rp_inf, fabs(acc));
    }
#pragma unroll 1
    for (int idx = rt; idx < NE; idx += T) {
        if (e_su[idx] < kBig) mu_acc = fma(e_su[idx], e_lu[idx], mu_acc);
        if (e_sl[idx] < kBig) mu_acc = fma(e_sl[idx], e_ll[idx], mu_acc);
    }
    {
        double d1 = 0.0, d3 = 0.0;
        block_reduce4(rp_inf, d1, mu_acc, d3);
    }
    double mu_c = n_box > 0.0 ? mu_acc / n_box : 0.0;
    double rp_scale = 1.0;

    auto sg_of = [&](int idx) -> double { return e_lu[idx] * e_iu[idx] + e_ll[idx] * e_il[idx]; };
    // right-hand side of one entry (rnmpc_qp.cuh rhs_eval): gradient (+ centring / second-order terms in the corrector pass)
    double sigma_mu = 0.0;
    auto rhs_entry = [&](const int pass, const int k, const int i) -> double {
        const int idx = k * NZ + i;
        double gq = qd_of(k, i) * (e_z[idx] + e_zn[idx]);
        if (pass == 1) {
            const double iu = e_iu[idx], il = e_il[idx], lu = e_lu[idx], ll = e_ll[idx], dza = e_da[idx];
            if (iu > 0.0) gq = fma(sigma_mu + dza * (lu * iu * dza - lu), iu, gq);
            if (il > 0.0) gq = fma(-(sigma_mu + dza * (ll * il * dza + ll)), il, gq);
        }
        return gq;
    };
    mark(0);

    int stalled = 0, it = 0;
    bool done = status != 0;
#pragma unroll 1
    for (it = 0; it < a.max_iter && !done; ++it) {
        bool chol_ok = true;
        double ap = 1.0, ad = 1.0, maxp0 = 0.0, maxd0 = 0.0;
        bool skip_corr = false;
        __syncthreads();
        // ================================================================ factorisation sweep + predictor backward recursion
#pragma unroll 1
        for (int e = rt; e < NX * NX; e += T) { const int i = e / NX, j = e - i * NX; S[i * LDS + j] = (i == j) ? (qdt[NZ + i] + sg_of(N * NZ + i)) : 0.0; }
        if (rt < NX) pv[rt] = rhs_entry(0, N, rt);
        __syncthreads();
        mark(1);
        ZFrag<NX, NU, NT, KS, true, 0, true> zsrc;
        zsrc.init(As, Bs, rp, lane);
        double zf[KS][NT];
        zsrc.load(N - 1, As, zf);
#pragma unroll 1
        for (int k = N - 1; k >= 0; --k) {
            double* Kk = Kq + (size_t)k * NU * LDA;
            // fragments of Z_k = [A_k | B_k | rp_k] (rows padded to KP with zeros): zf[s][n] = Z[4s + t][8n + g]
            // (branch-free: one pointer per lane and mixed column tile, see ZFrag -- the three-way branch per element had made this
            // load the largest part of phase A)
            // (loaded behind the second product of the previous stage: off the critical path of phase A)
            // ---- phase A: W = S [A B rp]  (one warp per row tile of S); the free warp prepares the stage scalars
            if (role < MTS) {
                double sa[KS], c0[NT], c1[NT];
#pragma unroll
                for (int s = 0; s < KS; ++s) sa[s] = S[(8 * role + g) * LDS + 4 * s + t4];
#pragma unroll
                for (int n = 0; n < NT; ++n) { c0[n] = 0.0; c1[n] = 0.0; }
#pragma unroll
                for (int s = 0; s < KS; ++s)
#pragma unroll
                    for (int n = 0; n < NT; ++n) dmma884(c0[n], c1[n], sa[s], zf[s][n]);       // NT independent accumulation chains
#pragma unroll
                for (int n = 0; n < NT; ++n) *reinterpret_cast<double2*>(W + (8 * role + g) * LDW + 8 * n + 2 * t4) = make_double2(c0[n], c1[n]);
            } else if (role == 3 && lane < NZ) {
                rhs[lane] = (lane < NX && k == 0) ? 0.0 : rhs_entry(0, k, lane);
                dgn[lane] = (lane < NX ? qdt[lane] : 0.0) + sg_of(k * NZ + lane);      // barrier weights (+ diagonal state cost); R2 is added to H below
            }
            __syncthreads();
            mark(2);
            // ---- phase B: M = [A B]^T [W p]  (upper triangle of tiles, round robin over the warps).  When H = B'SB + .. lies inside ONE tile
            //      (rocket, pendulum), that tile goes to warp 3 alone, which then factorises H while the other warps finish their tiles:
            //      the Cholesky chain (four dependent rsqrt, ~450 cycles) leaves the critical path of the stage.
            constexpr bool HT = (CB / 8 == (CB + NU - 1) / 8) && (CB / 8 < MTM);
            constexpr int TH = CB / 8;
            double L[NU][NU];
            bool chol_stage_ok = true;
            {
                int ti = 0, tj = 0;
#pragma unroll
                for (int mt = 0; mt < MTM; ++mt) {
#pragma unroll
                    for (int n = mt; n < NT; ++n, ++ti) {
                        const bool htile = HT && mt == TH && n == TH;
                        const int owner = HT ? (htile ? 3 : tj % 3) : (ti & 3);
                        if (!htile) ++tj;
                        if (owner != role) continue;
                        double c0 = 0.0, c1 = 0.0;
#pragma unroll
                        for (int s = 0; s < KS; ++s) {
                            const double bv = (8 * n + g == CP) ? pv[4 * s + t4] : W[(4 * s + t4) * LDW + 8 * n + g];
                            dmma884(c0, c1, zf[s][mt], bv);
                        }
                        *reinterpret_cast<double2*>(M + (8 * mt + g) * LDW + 8 * n + 2 * t4) = make_double2(c0, c1);
                    }
                }
            }
            auto factor_h = [&]() {
#pragma unroll
                for (int p = 0; p < NU; ++p)
#pragma unroll
                    for (int q = 0; q < NU; ++q) {
                        const int lo = p < q ? p : q, hi = p < q ? q : p;
                        L[p][q] = M[(CB + lo) * LDW + CB + hi] + R2s[p * NU + q] + (p == q ? dgn[NX + p] : 0.0);
                    }
                chol_stage_ok = chol_inplace<NU>(L);
            };
            if (HT && role == 3) {          // H's tile and dgn were written by this warp: a warp barrier makes them visible
                __syncwarp();
                factor_h();
            }
            if (k > 0) zsrc.load(k - 1, As, zf);      // next stage's operand fragments (rp, A, B do not change during the factorisation)
            __syncthreads();
            mark(3);
            // ---- phase C: H = R2 + diag + B'SB -> Cholesky (above when HT); K = -H^{-1} (A'SB)' column by column; t = ru + B'(p + rp_scale S rp),
            //      d = -H^{-1} t   (warp 3, one thread per column)
            if (role == 3 && lane < NX + 2) {
                if (!HT) factor_h();
                const bool ok = chol_stage_ok;
                const int rt = lane;             // (shadows the logical thread index: this phase is laid out on warp 3's lanes)
                // one triangular solve per thread on one code path: threads < NX a column of (A'SB)' -> column of K, thread NX the vector t -> d
                // (thread NX + 1 solves a copy of t and stores the factor instead)
                double f[NU];
#pragma unroll
                for (int p = 0; p < NU; ++p)
                    f[p] = rt < NX ? M[rt * LDW + CB + p] : rhs[NX + p] + M[(CB + p) * LDW + CP] + rp_scale * M[(CB + p) * LDW + CRP];
                if (rt == NX) {
#pragma unroll
                    for (int p = 0; p < NU; ++p) uc[p] = f[p];
                }
                chol_solve<NU>(L, f);
                double* fo = rt < NX ? Kk + rt : dff + k * NU;          // K_k[p][rt] (stride LDA) or d_k[p] (stride 1)
                const int fs = rt < NX ? LDA : 1;
                if (rt <= NX) {
#pragma unroll
                    for (int p = 0; p < NU; ++p) fo[p * fs] = -f[p];
                } else {
#pragma unroll
                    for (int p = 0; p < NU; ++p)
#pragma unroll
                        for (int q = 0; q < NU; ++q) Lq[k * NU * NU + p * NU + q] = L[p][q];
                    if (!ok) red[32] = 1.0;
                }
            }
            __syncthreads();
            mark(4);
            // ---- phase D: one more k-step per tile: S_k = A'SA + (A'SB) K + diag (upper triangle, mirrored: exactly symmetric);
            //      p_k = rx + A'g + K't and S rp_k on the last threads
            {
                if (k > 0) {
                    double kf[MTS];                               // B operand: K[t][8n + g]
#pragma unroll
                    for (int n = 0; n < MTS; ++n) kf[n] = (t4 < NU && 8 * n + g < NX) ? Kk[t4 * LDA + 8 * n + g] : 0.0;
                    int ti = 0;
#pragma unroll
                    for (int mt = 0; mt < MTS; ++mt) {
                        const int r = 8 * mt + g;
#pragma unroll
                        for (int n = mt; n < MTS; ++n, ++ti) {
                            if ((ti & 3) != role) continue;
                            const int c = 8 * n + 2 * t4;
                            const double2 mv = *reinterpret_cast<const double2*>(M + r * LDW + c);
                            double c0 = mv.x, c1 = mv.y;
                            const double gf = (t4 < NU) ? M[r * LDW + CB + t4] : 0.0;      // A operand: (A'SB)[r][t]
                            dmma884(c0, c1, gf, kf[n]);
                            if (r < NX) {
                                if (c < NX && r <= c) { const double v = c0 + (r == c ? dgn[r] : 0.0); S[r * LDS + c] = v; S[c * LDS + r] = v; }
                                if (c + 1 < NX && r <= c + 1) { const double v = c1 + (r == c + 1 ? dgn[r] : 0.0); S[r * LDS + c + 1] = v; S[(c + 1) * LDS + r] = v; }
                            }
                        }
                    }
                }
                if (rt >= T - NX) {
                    const int i = rt - (T - NX);
                    Sc[k * NX + i] = W[i * LDW + CRP];
                    if (k > 0) {
                        double acc = rhs[i] + M[i * LDW + CP] + rp_scale * M[i * LDW + CRP];
#pragma unroll
                        for (int p = 0; p < NU; ++p) acc = fma(Kk[p * LDA + i], uc[p], acc);
                        pv[i] = acc;                     // p_k: the p column of the next stage's second product (nobody reads pv in this phase)
                    }
                }
            }
            __syncthreads();
            mark(5);
        }
        chol_ok = red[32] == 0.0;

        // ================================================================ predictor (pass 0) and corrector (pass 1): one copy of the sweep code
        const double tau_c = mu_c > 1e-8 ? 0.995 : 0.9999;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            double* e_dz = pass ? e_dc : e_da;
            if (pass == 1) {
                // ---- corrector right-hand sides (e_dc doubles as scratch until the forward recursion overwrites it with the step)
#pragma unroll 1
                for (int idx = rt; idx < NE; idx += T) { const int k = idx / NZ; e_dc[idx] = rhs_entry(1, k, idx - k * NZ); }
                __syncthreads();
                mark(6);
                if (role == 0) {
                    // backward recursion on one warp: g_k = p_{k+1} + rp_scale Sc_k ; lanes < NX: p_k = rx_k + A_k' g_k + K_k' t_k,
                    // lanes NX .. NZ-1: t_k = ru_k + B_k' g_k (kept in tu for the feed-forward)
                    const int r = lane < NX ? lane : NX - 1;
                    const int p = (lane >= NX && lane < NZ) ? lane - NX : 0;
                    const double* bw_base = lane < NX ? As + r : Bs + p;                  // lanes >= NZ read A too; their result is not used
                    const int bw_kstride = lane < NX || lane >= NZ ? NX * LDA : NX * NU;
                    const int bw_stride = lane < NX || lane >= NZ ? LDA : NU;
                    double pl = e_dc[N * NZ + r];
#pragma unroll 1
                    for (int k = N - 1; k >= 0; --k) {
                        if (lane < NX) gv[k * NX + lane] = pl + rp_scale * Sc[k * NX + lane];
                        __syncwarp();
                        const double* gk = gv + k * NX;
                        // one code path for all lanes: column r of A_k (lanes < NX) or column p of B_k (lanes NX .. NZ-1), own base and strides
                        double a0 = lane < NZ ? e_dc[k * NZ + lane] : 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        {
                            const double* col = bw_base + (size_t)k * bw_kstride;
#pragma unroll
                            for (int m = 0; m + 3 < NX; m += 4) {
                                a0 = fma(col[m * bw_stride], gk[m], a0); a1 = fma(col[(m + 1) * bw_stride], gk[m + 1], a1);
                                a2 = fma(col[(m + 2) * bw_stride], gk[m + 2], a2); a3 = fma(col[(m + 3) * bw_stride], gk[m + 3], a3);
                            }
#pragma unroll
                            for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(col[m * bw_stride], gk[m], a0);
                        }
                        if (lane >= NX && lane < NZ) tu[k * NU + p] = (a0 + a1) + (a2 + a3);
                        __syncwarp();
                        if (lane < NX && k > 0) {
                            double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
                            for (int q = 0; q < NU; ++q) acc = fma(Kq[((size_t)k * NU + q) * LDA + r], tu[k * NU + q], acc);
                            pl = acc;
                        }
                    }
                }
                __syncthreads();
                mark(7);
                // ---- d_k = -H_k^{-1} t_k   (one thread per stage)
                if (rt < N) {
                    const int k = rt;
                    double d[NU], L[NU][NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        d[p] = tu[k * NU + p];
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = Lq[k * NU * NU + p * NU + q];
                    }
                    chol_solve<NU>(L, d);
#pragma unroll
                    for (int p = 0; p < NU; ++p) dff[k * NU + p] = -d[p];
                }
                __syncthreads();
                mark(8);
            }
            mark(9);
            // ---- forward recursion on one warp: lanes NX .. NZ-1: du_k = K_k dx_k + d_k ; lanes < NX: dx_{k+1} = A_k dx_k + B_k du_k +
            //      rp_scale rp_k, dx_0 = 0.  The steps land in the entries of e_dz.
            if (role == 0) {
                const int r = lane < NX ? lane : NX - 1;
                const int p = (lane >= NX && lane < NZ) ? lane - NX : 0;
                const double* fw_base = lane < NX || lane >= NZ ? As + (size_t)r * LDA : Kq + (size_t)p * LDA;      // lanes >= NZ: unused result
                const int fw_kstride = lane < NX || lane >= NZ ? NX * LDA : NU * LDA;
                const double* fw_init = lane < NX || lane >= NZ ? rp + r : dff + p;
                const int fw_istride = lane < NX || lane >= NZ ? NX : NU;
                const double fw_scale = lane < NX || lane >= NZ ? rp_scale : 1.0;
                if (lane < NX) e_dz[lane] = 0.0;
                __syncwarp();
#pragma unroll 1
                for (int k = 0; k < N; ++k) {
                    const double* xk = e_dz + k * NZ;
                    // one code path for all lanes: row r of A_k from rp_scale rp_k (lanes < NX) or row p of K_k from d_k (lanes NX .. NZ-1);
                    // dx_0 = 0, so stage 0 needs no special case
                    double a0 = fw_scale * fw_init[k * fw_istride], a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    {
                        const double* row = fw_base + (size_t)k * fw_kstride;
#pragma unroll
                        for (int m = 0; m + 3 < NX; m += 4) {
                            const double2 r01 = *reinterpret_cast<const double2*>(row + m), r23 = *reinterpret_cast<const double2*>(row + m + 2);
                            a0 = fma(r01.x, xk[m], a0); a1 = fma(r01.y, xk[m + 1], a1);
                            a2 = fma(r23.x, xk[m + 2], a2); a3 = fma(r23.y, xk[m + 3], a3);
                        }
#pragma unroll
                        for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(row[m], xk[m], a0);
                    }
                    if (lane >= NX && lane < NZ) e_dz[k * NZ + NX + p] = (a0 + a1) + (a2 + a3);
                    __syncwarp();
                    if (lane < NX) {
                        double acc = (a0 + a1) + (a2 + a3);
                        const double* Br = Bs + ((size_t)k * NX + r) * NU;
#pragma unroll
                        for (int q = 0; q < NU; ++q) acc = fma(Br[q], xk[NX + q], acc);
                        e_dz[(k + 1) * NZ + lane] = acc;
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
            mark(10);
            // ---- ratio test over all entries (rnmpc_qp.cuh ratio_entry)
            double rpm = 0.0, rdm = 0.0, S3 = 0.0, S4 = 0.0;
#pragma unroll 1
            for (int idx = rt; idx < NE; idx += T) {
                const double dz = e_dz[idx];
                const double iu = e_iu[idx], il = e_il[idx], lu = e_lu[idx], ll = e_ll[idx];
                const double dza = pass == 1 ? e_da[idx] : 0.0;
                const double smu = pass == 1 ? sigma_mu : 0.0;
                if (iu > 0.0) {
                    const double ru = lu * iu;
                    const double tgt = pass == 1 ? smu + dza * (ru * dza - lu) : 0.0;
                    const double dlu = tgt * iu - lu + ru * dz;
                    rpm = fmax(rpm, dz * iu);
                    if (dlu < 0.0) rdm = fmax(rdm, -dlu * fast_rcp(lu));
                    if (pass == 0) { S3 = fma(-dz, lu, S3); S4 = fma(-dz, dlu, S4); }
                }
                if (il > 0.0) {
                    const double rl = ll * il;
                    const double tgt = pass == 1 ? smu + dza * (rl * dza + ll) : 0.0;
                    const double dll = tgt * il - ll - rl * dz;
                    rpm = fmax(rpm, -dz * il);
                    if (dll < 0.0) rdm = fmax(rdm, -dll * fast_rcp(ll));
                    if (pass == 0) { S3 = fma(dz, ll, S3); S4 = fma(dz, dll, S4); }
                }
            }
            block_reduce4(rpm, rdm, S3, S4);
            mark(11);
            if (pass == 0) {
                maxp0 = rpm; maxd0 = rdm;
                ap = (rpm > 1.0) ? 1.0 / rpm : 1.0;
                ad = (rdm > 1.0) ? 1.0 / rdm : 1.0;
                const double mu_aff = n_box > 0.0 ? (mu_c * n_box * (1.0 - ad) + (ap - ad) * S3 + ap * ad * S4) / n_box : 0.0;
                skip_corr = (ap >= 0.999 && ad >= 0.999 && mu_aff <= 1e-3 * mu_c);
                const double ratio_mu = mu_c > 0.0 ? mu_aff / mu_c : 0.0;
                sigma_mu = ratio_mu * ratio_mu * ratio_mu * mu_c;
                if (skip_corr) {
                    // Newton phase: the predictor step with the fraction-to-boundary rule (rnmpc_qp.cuh)
                    ap = (maxp0 > tau_c) ? tau_c / maxp0 : 1.0;
                    ad = (maxd0 > tau_c) ? tau_c / maxd0 : 1.0;
                    break;
                }
            } else {
                ap = (rpm > tau_c) ? tau_c / rpm : 1.0;
                ad = (rdm > tau_c) ? tau_c / rdm : 1.0;
            }
        }
        // ================================================================ update, new reciprocals / complementarity
        double acc = 0.0;
        {
            const bool upd = chol_ok;
            const double so = skip_corr ? 0.0 : 1.0;
            const double* e_dz = skip_corr ? e_da : e_dc;
#pragma unroll 1
            for (int idx = rt; idx < NE; idx += T) {
                double su = e_su[idx], sl = e_sl[idx], lu = e_lu[idx], ll = e_ll[idx];
                if (upd) {
                    const double dza = e_da[idx], dz = e_dz[idx];
                    double iu = 0.0, il = 0.0;
                    if (su < kBig) {
                        const double ru = lu * e_iu[idx];
                        const double dlu = so * (sigma_mu + dza * (ru * dza - lu)) * e_iu[idx] - lu + ru * dz;
                        su -= ap * dz; lu += ad * dlu;
                        iu = fast_rcp(su);
                    }
                    if (sl < kBig) {
                        const double rl = ll * e_il[idx];
                        const double dll = so * (sigma_mu + dza * (rl * dza + ll)) * e_il[idx] - ll - rl * dz;
                        sl += ap * dz; ll += ad * dll;
                        il = fast_rcp(sl);
                    }
                    e_su[idx] = su; e_sl[idx] = sl; e_lu[idx] = lu; e_ll[idx] = ll; e_iu[idx] = iu; e_il[idx] = il;
                    e_z[idx] += ap * dz;
                }
                if (su < kBig) acc = fma(su, lu, acc);
                if (sl < kBig) acc = fma(sl, ll, acc);
            }
            if (upd) rp_scale *= (1.0 - ap);
        }
        {
            double d0 = 0.0, d1 = 0.0, d3 = 0.0;
            block_reduce4(d0, d1, acc, d3);
        }
        if (rt == 0) red[32] = 0.0;
        mark(12);
        const double mu_new = n_box > 0.0 ? acc / n_box : 0.0;
        if (!chol_ok) {
            if (mu_c > 1e-9) status = 3;           // a non-positive pivot at mu ~ 1e-14 means converged to working precision
            done = true;
        } else if (!(mu_new == mu_new)) { status = 3; done = true; }
        else {
            mu_c = mu_new;
            if (mu_c < a.tol_mu && rp_inf * rp_scale < a.tol_res && it >= 2) done = true;
            stalled = (ap < 1e-4 && ad < 1e-4) ? stalled + 1 : 0;
            if (!done && (stalled >= 3 || mu_c > 1e12)) { status = 2; done = true; }
        }
        if (done) { ++it; break; }
    }
    if (!done && status == 0) status = (mu_c < 1e-9 && rp_inf * rp_scale < 1e-7) ? 0 : 1;
    __syncthreads();

    // ---- outputs in the reference layouts + objective x'Hx + q'x (qp_jit.py:401, P = 2H, q = 2 H y_nom; diagonal H)
    {
        double* Xo = a.dX + (size_t)b * (N + 1) * NX;
        double* Uo = a.dU + (size_t)b * N * NU;
        double* mu = a.mu + (size_t)b * N * NI;
        double* muf = a.mu_f + (size_t)b * NIF;
        double acc = 0.0;
#pragma unroll 1
        for (int idx = rt; idx < NE; idx += T) {
            const int k = idx / NZ, i = idx - k * NZ;
            if (i >= NX && k == N) continue;
            const double z = e_z[idx];
            acc = fma(z, 0.5 * qd_of(k, i) * (z + 2.0 * e_zn[idx]), acc);
            if (i < NX) Xo[k * NX + i] = z; else Uo[k * NU + i - NX] = z;
            if (k < N) { mu[k * NI + i] = e_lu[idx]; mu[k * NI + NZ + i] = e_ll[idx]; }
            else { muf[i] = e_lu[idx]; muf[NX + i] = e_ll[idx]; }
        }
        double d0 = 0.0, d1 = 0.0, d3 = 0.0;
        block_reduce4(d0, d1, acc, d3);
        mark(13);
        if (prof_on) { for (int i = 0; i < 16; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + i, (unsigned long long)pacc[i]); }
        if (rt == 0) {
            a.cost[b] = acc;
            a.iters[b] = it;
            a.status[b] = status;
            if (a.ws_valid != nullptr) a.ws_valid[b] = 0;          // the records of qp_ipm_kernel's workspace were not written
        }
    }
}

}  // namespace rnmpc
