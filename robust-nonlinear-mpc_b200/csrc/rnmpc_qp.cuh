// Nominal QP kernel (reference: solver/qp_jit.py:77-273, 362-402, 487-513 handing the problem to OSQP 1.0.4).
//
//   min  sum_k dz_k' H_k dz_k + q_k' dz_k      (OSQP form: P = 2 blkdiag(Q,R,..,Qf), q = 2 H y_nom)
//   s.t. dx_{k+1} = A_k dx_k + B_k du_k + c_k,  dx_0 = x0 - x_nom0,  lo_k <= dz_k <= hi_k   (G = [I;-I])
//
// Solved per instance by a Mehrotra predictor-corrector primal-dual interior-point method whose every KKT solve is
// a Riccati factorisation / sweep over the block-tridiagonal structure (riccati_step of rnmpc_common.cuh: the same
// device function the fast-SLS kernel runs).  One instance per slot (LPS lanes), SPW instances per warp in
// lock-step.  Per-instance vectors live in global (L2-resident) workspace in one [N+1][nz] "entry" layout; the
// element-wise phases load groups of KU stages ahead of the arithmetic so that their latency overlaps.
#pragma once
#include "rnmpc_common.cuh"

namespace rnmpc {

struct QpArgs {
    int B, N;
    int tightened;          // 0: bounds g_k - G z_k + 1e-10 (first QP of a fast-SLS solve); 1: g_k - G z_k - backoff_k, terminal gf - backoff_f (Q4);
                            // -1: per instance, from flags[b] & FLAG_TIGHT
    int max_iter;
    double tol_mu, tol_res;
    const double* A;        // [B][N][nx][LD] (padded rows)
    const double* Bm;       // [B][N][nx][nu]
    const double* c;        // [B][N][nx]
    const double* Xnom;     // [B][N+1][nx]
    const double* Unom;     // [B][N][nu]
    const double* x0;       // [B][nx] measured state
    const double* backoff;  // [B][N][ni]   (tightened only)
    const double* backoff_f;// [B][ni_f]
    const int* flags;       // [B] or nullptr: instances with (flags & skip_mask) do no work
    int skip_mask;
    // general mode (QP class drop-in, qp_jit.py): explicit bounds and linear cost in entry layout [B][N+1][nz]; Xnom/Unom are zero
    const double* hi_in;
    const double* lo_in;
    const double* q_lin;
    // outputs
    double* dX;             // [B][N+1][nx]
    double* dU;             // [B][N][nu]
    double* mu;             // [B][N][ni]
    double* mu_f;           // [B][ni_f]
    double* cost;           // [B]
    int* iters;             // [B]
    int* status;            // [B] 0 ok, 1 not converged, 2 infeasible (measured state outside the stage-0 box by more than
                            //     kX0Tol, or crossed bounds), 3 numerical failure (non-positive pivot / NaN)
    // workspace, entry layout [B][N+1][nz] unless noted
    double* Z;              // primal iterate (x_k, u_k)
    double* SU;             // slack of the upper bound  (>= kBig: no bound)
    double* SL;             // slack of the lower bound
    double* LU;             // multiplier of the upper bound
    double* LL;             // multiplier of the lower bound
    double* DZA;            // affine (predictor) step
    double* DZC;            // combined step
    double* G;              // right-hand side of the current KKT solve
    double* SIG;            // barrier Hessian diagonal  LU/SU + LL/SL
    double* ZN;             // nominal trajectory in entry layout
    double* ISU;            // 1/SU (0 where there is no bound)
    double* ISL;            // 1/SL
    int diag_cost;          // 1: Q, R, Qf are diagonal (all reference drivers) -> element-wise gradient
    int warm;               // warm start from the workspace's previous solution of the same instance: bit0 multipliers, bit1 shifted by one
                            // stage (previous MPC step), bit2 primal iterate; a failed warm-started solve is repeated from the cold start
    int* ws_valid;          // [B] workspace holds a converged solution
    double* rp;             // [B][N][nx]     dynamics residual at the initial point
    double* Sc;             // [B][N][nx]     S_{k+1} rp_k
    double* dff;            // [B][N][nu]     feed-forward
    double* Kq;             // [B][N][nu][nx]
    double* Lq;             // [B][N][nu*nu]
};

template <int V>
struct IntTag { static constexpr int value = V; };

template <int NX, int NU>
__global__ void __launch_bounds__(160, 2) qp_ipm_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using SL_ = Slot<NX>;
    constexpr int LD = SL_::LD, LPS = SL_::LPS, SPW = SL_::SPW;
    constexpr int NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX;
    constexpr int KU = 4;                        // stages loaded ahead in the element-wise passes
    static_assert(LPS >= NU, "one lane per input entry");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SlotSmem<NX, NU>* smem = reinterpret_cast<SlotSmem<NX, NU>*>(smem_raw);

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int slot_w = lane / LPS;
    const int sl = lane - slot_w * LPS;
    const int N = a.N;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    const long long item = warp_global * SPW + slot_w;
    const int b = (int)item;
    bool active = (slot_w < SPW) && (item < a.B);
    if (active && a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) active = false;
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    SlotSmem<NX, NU>& sm = smem[warp_in_cta * SPW + slot_w];

    const int r0 = 2 * sl, r1 = 2 * sl + 1;
    const bool v1 = r1 < NX;
    const int r1c = v1 ? r1 : NX - 1;
    const bool vu = sl < NU;
    // the (up to) three entries of every stage this lane owns: x[r0], x[r1], u[sl]
    const int ei[3] = {r0, r1c, NX + (vu ? sl : 0)};
    const bool ev[3] = {true, v1, vu};

    const double* Ab = a.A + (size_t)b * N * NX * LD;
    const double* Bb = a.Bm + (size_t)b * N * NX * NU;
    const double* cb = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;
    const size_t eoff = (size_t)b * (N + 1) * NZ;
    double* __restrict__ Z = a.Z + eoff;
    double* __restrict__ SU = a.SU + eoff;
    double* __restrict__ SLk = a.SL + eoff;
    double* __restrict__ LU = a.LU + eoff;
    double* __restrict__ LL = a.LL + eoff;
    double* __restrict__ DZA = a.DZA + eoff;
    double* __restrict__ DZC = a.DZC + eoff;
    double* __restrict__ G = a.G + eoff;
    double* __restrict__ SIG = a.SIG + eoff;
    double* __restrict__ ZN = a.ZN + eoff;
    double* __restrict__ ISU = a.ISU + eoff;
    double* __restrict__ ISL = a.ISL + eoff;
    double* rp = a.rp + (size_t)b * N * NX;
    double* Sc = a.Sc + (size_t)b * N * NX;
    double* dff = a.dff + (size_t)b * N * NU;
    double* Kq = a.Kq + (size_t)b * N * NU * NX;
    double* Lq = a.Lq + (size_t)b * N * NU * NU;

    auto slot_sum2 = [&](double& x, double& y) {
        __syncwarp(mask);
        sm.red[2 * sl] = x; sm.red[2 * sl + 1] = y;
        __syncwarp(mask);
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int i = 0; i < LPS; ++i) { sx += sm.red[2 * i]; sy += sm.red[2 * i + 1]; }
        x = sx; y = sy;
    };
    auto slot_min2 = [&](double& x, double& y) {
        __syncwarp(mask);
        sm.red[2 * sl] = x; sm.red[2 * sl + 1] = y;
        __syncwarp(mask);
        double sx = sm.red[0], sy = sm.red[1];
#pragma unroll
        for (int i = 1; i < LPS; ++i) { sx = fmin(sx, sm.red[2 * i]); sy = fmin(sy, sm.red[2 * i + 1]); }
        x = sx; y = sy;
    };
    // element-wise pass over the entries of this lane: NA arrays are loaded for KU stages at a time, then
    // body(k, idx, vals) runs for every owned entry (idx = k*NZ + i)
    auto entry_pass = [&](auto na_tag, const double* const* arr, auto&& body) {
        constexpr int NA = decltype(na_tag)::value;
        for (int k0 = 0; k0 <= N; k0 += KU) {
            double v[KU][3][NA];
#pragma unroll
            for (int j = 0; j < KU; ++j) {
                const int k = (k0 + j <= N) ? (k0 + j) : N;
#pragma unroll
                for (int e = 0; e < 3; ++e)
#pragma unroll
                    for (int q = 0; q < NA; ++q) v[j][e][q] = arr[q][k * NZ + ei[e]];
            }
#pragma unroll
            for (int j = 0; j < KU; ++j) {
                const int k = k0 + j;
                if (k <= N) {
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                        if (ev[e]) body(k, k * NZ + ei[e], v[j][e]);
                }
            }
        }
    };
    using I4 = IntTag<4>;
    using I5 = IntTag<5>;
    using I6 = IntTag<6>;

    int status = 0;
    const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);
    double n_box = 0.0;
    bool warm = (a.warm != 0) && (a.ws_valid != nullptr) && (a.ws_valid[b] != 0);
    int it = 0, it_total = 0;
    double rp_scale = 1.0, rp_inf = 0.0, mu_c = 1.0;
    constexpr double kMuWarm = 1e-2, kFracWarm = 1e-3;
    constexpr int kWarmCap = 15;
    bool structural = false;       // infeasible by construction (crossed bounds / x(0) outside its box): no retry
    bool act = true;               // this slot (re)starts in the current attempt
  for (int attempt = 0; attempt < 2; ++attempt) {
    if (act) { status = 0; rp_scale = 1.0; }
    double n_box_new = 0.0;

    // ------------------------------------------------------------------ bounds, initial point, slacks, duals
    {
        const double e = tight ? 0.0 : kEpsQP;
        double bad = 0.0;
        for (int k = 0; k <= N; ++k) {
#pragma unroll
            for (int en = 0; en < 3; ++en) {
                if (!ev[en]) continue;
                const int i = ei[en];
                const int idx = k * NZ + i;
                double h_ = 1e300, l_ = -1e300, z = 0.0;
                bool box = false;
                if (en < 2) {                       // state entry
                    const double xn = Xn[k * NX + i];
                    if (k < N) {
                        const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                        const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                        h_ = (prm->xub[i] - xn) - bu + e;
                        l_ = -((xn - prm->xlb[i]) - bl + e);
                    } else if (tight) {
                        h_ = prm->xubf[i] - a.backoff_f[(size_t)b * NIF + i];          // Q4: terminal bound not shifted by x_N
                        l_ = -(-prm->xlbf[i] - a.backoff_f[(size_t)b * NIF + NX + i]);
                    } else {
                        h_ = (prm->xubf[i] - xn) + e;
                        l_ = -((xn - prm->xlbf[i]) + e);
                    }
                    if (k == 0) z = a.x0[(size_t)b * NX + i] - xn;      // x(0) is pinned to the measurement
                    else box = true;
                } else if (k < N) {                 // input entry
                    const int p = i - NX;
                    const double un = Un[k * NU + p];
                    const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                    const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                    h_ = (prm->uub[p] - un) - bu + e;
                    l_ = -((un - prm->ulb[p]) - bl + e);
                    box = true;
                }
                if (a.hi_in != nullptr && (en < 2 || k < N)) { h_ = a.hi_in[eoff + idx]; l_ = a.lo_in[eoff + idx]; }
                // the stage-0 state rows only restate the pinned x(0): they are dropped; a violation beyond kX0Tol (OSQP's default
                // eps_abs, which the reference's solver would tolerate) makes the instance infeasible
                if (k == 0 && en < 2 && (z > h_ + kX0Tol || z < l_ - kX0Tol)) bad = 1.0;
                double su = 1e300, sl_ = 1e300, lu = 0.0, ll = 0.0;
                if (box) {
                    const bool hu = h_ < kBig, hl = l_ > -kBig;
                    double z0 = 0.0, lu0 = 1.0, ll0 = 1.0;
                    if (warm) {
                        // previous solution of this instance (same workspace); bit1: the multipliers of the next stage (the
                        // trajectories were shifted by one stage since); reads run ahead of this lane's own writes
                        const int ks = (a.warm & 2) ? ((en < 2) ? (k + 1 <= N ? k + 1 : N) : (k + 1 < N ? k + 1 : N - 1)) : k;
                        lu0 = LU[ks * NZ + i]; ll0 = LL[ks * NZ + i];
                        if (a.warm & 4) z0 = Z[idx];
                    }
                    if (hu && hl) {
                        const double m = (warm ? kFracWarm : 0.01) * (h_ - l_);
                        if (!(h_ - l_ > 1e-12)) bad = 1.0;
                        z = fmin(fmax(z0, l_ + m), h_ - m);
                    } else if (hu) z = fmin(z0, h_ - 1.0);
                    else if (hl) z = fmax(z0, l_ + 1.0);
                    if (hu) { su = h_ - z; lu = warm ? fmax(lu0, kMuWarm / su) : 1.0; n_box_new += 1.0; }
                    if (hl) { sl_ = z - l_; ll = warm ? fmax(ll0, kMuWarm / sl_) : 1.0; n_box_new += 1.0; }
                }
                if (act) { Z[idx] = z; SU[idx] = su; SLk[idx] = sl_; LU[idx] = lu; LL[idx] = ll; }
                ZN[idx] = (en < 2) ? Xn[k * NX + i] : (k < N ? Un[k * NU + (i - NX)] : 0.0);
            }
        }
        slot_sum2(bad, n_box_new);
        if (act) {
            n_box = n_box_new;
            if (bad > 0.0) { status = 2; structural = true; }
        }
        __syncwarp(mask);
        // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
        for (int k = 0; k < N; ++k) {
            const double* Ak = Ab + (size_t)k * NX * LD;
            const double* Bk = Bb + (size_t)k * NX * NU;
            double acc0 = cb[k * NX + r0] - Z[(k + 1) * NZ + r0];
            double acc1 = cb[k * NX + r1c] - Z[(k + 1) * NZ + r1c];
#pragma unroll
            for (int m = 0; m < NX; ++m) {
                const double xm = Z[k * NZ + m];
                acc0 = fma(__ldg(Ak + r0 * LD + m), xm, acc0);
                acc1 = fma(__ldg(Ak + r1c * LD + m), xm, acc1);
            }
#pragma unroll
            for (int p = 0; p < NU; ++p) {
                const double up = Z[k * NZ + NX + p];
                acc0 = fma(__ldg(Bk + r0 * NU + p), up, acc0);
                acc1 = fma(__ldg(Bk + r1c * NU + p), up, acc1);
            }
            if (act) {
                rp[k * NX + r0] = acc0;
                if (v1) rp[k * NX + r1] = acc1;
            }
        }
        __syncwarp(mask);
    }

    {
        double m0 = 0.0, dmy = 0.0;
        for (int k = 0; k < N; ++k) {
            m0 = fmax(m0, fabs(rp[k * NX + r0]));
            if (v1) m0 = fmax(m0, fabs(rp[k * NX + r1]));
        }
        m0 = -m0;
        slot_min2(m0, dmy);
        if (act) rp_inf = -m0;
    }
    {
        double acc = 0.0, dmy = 0.0;
        const double* arr[4] = {SU, SLk, LU, LL};
        entry_pass(I4{}, arr, [&](int, int, const double (&v)[4]) {
            if (v[0] < kBig) acc = fma(v[0], v[2], acc);
            if (v[1] < kBig) acc = fma(v[1], v[3], acc);
        });
        slot_sum2(acc, dmy);
        if (act) mu_c = n_box > 0.0 ? acc / n_box : 0.0;
    }

    int stalled = 0;
    bool done = (status != 0) || !act;
    int resT = -1, resS = -1;      // stage whose A_k is resident (or in flight, cp.async) in sm.T / sm.S
    // a warm-started attempt that has not converged after kWarmCap iterations is abandoned for the cold start
    const int it_cap = (attempt == 0 && __any_sync(mask, warm)) ? min(a.max_iter, kWarmCap) : a.max_iter;
    for (it = 0; it < it_cap; ++it) {
        if (__all_sync(mask, done)) break;
        __syncwarp(mask);
        // ---- P1: reciprocal slacks, barrier diagonal SIG and gradient G = 2 H (z + z_nom)
        {
            const double* arr[6] = {SU, SLk, LU, LL, Z, ZN};
            const double qd[3] = {prm->Q[r0 * NX + r0], prm->Q[r1c * NX + r1c], vu ? prm->R[sl * NU + sl] : 0.0};
            const double qfd[2] = {prm->Qf[r0 * NX + r0], prm->Qf[r1c * NX + r1c]};
            const bool dg = a.diag_cost != 0;
            entry_pass(I6{}, arr, [&](int k, int idx, const double (&v)[6]) {
                double sg = 0.0, iu = 0.0, il = 0.0;
                if (v[0] < kBig) { iu = 1.0 / v[0]; sg = v[2] * iu; }
                if (v[1] < kBig) { il = 1.0 / v[1]; sg = fma(v[3], il, sg); }
                SIG[idx] = sg; ISU[idx] = iu; ISL[idx] = il;
                if (dg) {
                    const int i = idx - k * NZ;
                    const double q = (i >= NX) ? qd[2] : (i == r0 ? (k < N ? qd[0] : qfd[0]) : (k < N ? qd[1] : qfd[1]));
                    G[idx] = 2.0 * q * (v[4] + v[5]) + (a.q_lin != nullptr ? a.q_lin[eoff + idx] : 0.0);
                }
            });
            if (!dg) {
                for (int k = 0; k <= N; ++k) {
                    const double* Qm = (k < N) ? prm->Q : prm->Qf;
                    double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                    for (int m = 0; m < NX; ++m) {
                        const double y = Z[k * NZ + m] + ZN[k * NZ + m];
                        acc0 = fma(Qm[r0 * NX + m], y, acc0);
                        acc1 = fma(Qm[r1c * NX + m], y, acc1);
                    }
                    G[k * NZ + r0] = 2.0 * acc0 + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r0] : 0.0);
                    if (v1) G[k * NZ + r1] = 2.0 * acc1 + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r1] : 0.0);
                    if (vu && k < N) {
                        double acc = 0.0;
#pragma unroll
                        for (int m = 0; m < NU; ++m) acc = fma(prm->R[sl * NU + m], Z[k * NZ + NX + m] + ZN[k * NZ + NX + m], acc);
                        G[k * NZ + NX + sl] = 2.0 * acc + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + NX + sl] : 0.0);
                    }
                }
            }
        }
        __syncwarp(mask);

        double sigma_mu = 0.0, ap = 1.0, ad = 1.0;
        bool chol_ok = true;
        bool skip_corr = false, corr_skipped = false;
        double maxp0 = 0.0, maxd0 = 0.0;
        for (int pass = 0; pass < 2; ++pass) {
            // the corrector's two sweeps are skipped when every slot of the warp is in the Newton phase (see below)
            if (pass == 1 && __all_sync(mask, skip_corr)) break;
            // ---- backward: p_N = G_N ; pass 0 also builds S_N = 2 Qf + SIG_N
            sm.vec[r0] = G[N * NZ + r0];
            if (v1) sm.vec[r1] = G[N * NZ + r1];
            if (pass == 0) {
                resS = -1;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = rr ? r1 : r0;
                    if (rr && !v1) continue;
                    const double sg = SIG[N * NZ + r];
                    for (int c = 0; c < LD; ++c) sm.S[r * LD + c] = (c < NX) ? (2.0 * prm->Qf[r * NX + c] + (c == r ? sg : 0.0)) : 0.0;
                }
            }
            __syncwarp(mask);
            for (int k = N - 1; k >= 0; --k) {
                const double* Ak = Ab + (size_t)k * NX * LD;
                const double* Bk = Bb + (size_t)k * NX * NU;
                // prefetch this stage's scalars before the (long) matrix work
                const double gx0 = G[k * NZ + r0], gx1 = G[k * NZ + r1c];
                const double gu_ = vu ? G[k * NZ + NX + sl] : 0.0;
                const double rp0 = rp[k * NX + r0], rp1 = rp[k * NX + r1c];
                double sc0 = 0.0, sc1 = 0.0;
                // Sc_k = S_{k+1} rp_k while S_{k+1} is still in shared memory (pass 0); the corrector re-reads it
                auto pre_step = [&]() {
                    sm.vec[LD + r0] = rp0;
                    if (v1) sm.vec[LD + r1] = rp1;
                    __syncwarp(mask);
#pragma unroll 1
                    for (int c = 0; c < NX; ++c) {
                        const double cv = sm.vec[LD + c];
                        sc0 = fma(sm.S[r0 * LD + c], cv, sc0);
                        sc1 = fma(sm.S[r1c * LD + c], cv, sc1);
                    }
                    Sc[k * NX + r0] = sc0;
                    if (v1) Sc[k * NX + r1] = sc1;
                    __syncwarp(mask);
                };
                // vector recursion of the KKT solve; A_k is in sm.T, B_k in sm.Bm, K columns / factor in registers
                const double* Abuf = sm.T;
                auto vec_step = [&](const double (&k0)[NU], const double (&k1)[NU], const double (&L)[NU][NU]) {
                    // g = p_{k+1} + Sc
                    const double g0 = sm.vec[r0] + rp_scale * sc0;
                    const double g1 = v1 ? (sm.vec[r1] + rp_scale * sc1) : 0.0;
                    sm.vec[LD + r0] = g0;
                    if (v1) sm.vec[LD + r1] = g1;
                    __syncwarp(mask);
                    // t = ru + B' g
                    if (vu) {
                        double t = gu_;
#pragma unroll
                        for (int i = 0; i < NX; ++i) t = fma(sm.Bm[i * NU + sl], sm.vec[LD + i], t);
                        sm.vec[2 * LD + sl] = t;
                    }
                    __syncwarp(mask);
                    double tt[NU], dd[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) { tt[p] = sm.vec[2 * LD + p]; dd[p] = tt[p]; }
                    chol_solve<NU>(L, dd);
                    if (vu) {
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q < NU; ++q) if (q == sl) v = dd[q];
                        dff[k * NU + sl] = -v;
                    }
                    if (k > 0) {
                        // p_k = rx_k + A' g + K' t
                        double p0 = gx0, p1 = v1 ? gx1 : 0.0;
#pragma unroll 1
                        for (int i = 0; i < NX; ++i) {
                            const double gi = sm.vec[LD + i];
                            const double2 av = *reinterpret_cast<const double2*>(Abuf + i * LD + r0);
                            p0 = fma(av.x, gi, p0);
                            p1 = fma(av.y, gi, p1);
                        }
#pragma unroll
                        for (int p = 0; p < NU; ++p) { p0 = fma(k0[p], tt[p], p0); p1 = fma(k1[p], tt[p], p1); }
                        __syncwarp(mask);
                        sm.vec[r0] = p0;
                        if (v1) sm.vec[r1] = p1;
                    }
                    __syncwarp(mask);
                };
                if (pass == 0) {
                    const double cx0 = (k > 0) ? SIG[k * NZ + r0] : 0.0;
                    const double cx1 = (k > 0 && v1) ? SIG[k * NZ + r1] : 0.0;
                    if (vu) sm.cu[sl] = SIG[k * NZ + NX + sl];
                    const bool ok = riccati_step<NX, NU>(
                        sm, sl, mask, Ak, Bk, prm->Q2, prm->R2, cx0, cx1, Kq + (size_t)k * NU * NX, pre_step,
                        [&](const double (&k0)[NU], const double (&k1)[NU], const double (&L)[NU][NU]) {
                            if (sl == 0) {
#pragma unroll
                                for (int p = 0; p < NU; ++p)
#pragma unroll
                                    for (int q = 0; q < NU; ++q) Lq[k * NU * NU + p * NU + q] = L[p][q];
                            }
                            vec_step(k0, k1, L);
                        },
                        k > 0, k > 0 ? Ak - NX * LD : nullptr, k > 0 ? Bk - NX * NU : nullptr, resT == k, k > 0 ? Ak - NX * LD : nullptr);
                    resT = (k > 0) ? k - 1 : 0;
                    chol_ok = chol_ok && ok;
                } else {
                    double k0[NU], k1[NU], L[NU][NU];
                    cp_async_wait_all();
                    if (resT != k && resS != k) { load_padded_smem<NX, LD, LPS>(sm.T, Ak, sl); resT = k; }
                    const bool inT = (resT == k);
                    Abuf = inT ? sm.T : sm.S;
                    if (k > 0) {       // next stage's A into the other buffer while this stage computes
                        cp_async_matrix<NX, LD, LPS>(inT ? sm.S : sm.T, Ak - NX * LD, sl);
                        if (inT) resS = k - 1; else resT = k - 1;
                    }
                    load_matrix_smem<NX, NU, NU, LPS>(sm.Bm, Bk, sl);
                    if (k > 0) {
                        if (k > 1) prefetch_l2<LPS>(Ak - 2 * NX * LD, NX * LD * 8, sl);     // two stages ahead: the cp.async of the next stage hits L2
                        prefetch_l2<LPS>(Bk - NX * NU, NX * NU * 8, sl);
                        prefetch_l2<LPS>(Kq + (size_t)(k - 1) * NU * NX, NU * NX * 8, sl);
                    }
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        k0[p] = Kq[(size_t)k * NU * NX + p * NX + r0];
                        k1[p] = v1 ? Kq[(size_t)k * NU * NX + p * NX + r1c] : 0.0;
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = Lq[k * NU * NU + p * NU + q];
                    }
                    sc0 = Sc[k * NX + r0];
                    sc1 = Sc[k * NX + r1c];
                    __syncwarp(mask);
                    vec_step(k0, k1, L);
                }
            }
            // ---- forward: dx_0 = 0 ; du_k = K dx_k + d_k ; dx_{k+1} = A dx_k + B du_k + rp_k ; fused ratio test
            double* __restrict__ DZ = pass == 0 ? DZA : DZC;
            const double tau = pass == 0 ? 1.0 : (mu_c > 1e-8 ? 0.995 : 0.9999);
            // largest dz/s and -dlam/lam over the entries: alpha = min(1, tau / max)
            double rpm = 0.0, rdm = 0.0;
            auto ratio_entry = [&](double isu, double isl, double lu, double ll, double dz, double dza) {
                if (isu > 0.0) {
                    const double ru = lu * isu;
                    double tgt = 0.0;
                    if (pass == 1) tgt = sigma_mu + dza * (ru * dza - lu);
                    const double dlu = tgt * isu - lu + ru * dz;
                    rpm = fmax(rpm, dz * isu);
                    if (dlu < 0.0) rdm = fmax(rdm, -dlu / lu);
                }
                if (isl > 0.0) {
                    const double rl = ll * isl;
                    double tgt = 0.0;
                    if (pass == 1) tgt = sigma_mu + dza * (rl * dza + ll);
                    const double dll = tgt * isl - ll - rl * dz;
                    rpm = fmax(rpm, -dz * isl);
                    if (dll < 0.0) rdm = fmax(rdm, -dll / ll);
                }
            };
            __syncwarp(mask);
            sm.vec[r0] = 0.0;
            if (v1) sm.vec[r1] = 0.0;
            DZ[r0] = 0.0;
            if (v1) DZ[r1] = 0.0;
            __syncwarp(mask);
            for (int k = 0; k < N; ++k) {
                const double* Ak = Ab + (size_t)k * NX * LD;
                const double* Bk = Bb + (size_t)k * NX * NU;
                const double* Kk = Kq + (size_t)k * NU * NX;
                // stage A_k (already resident / in flight from the previous stage or sweep), B_k, K_k + this stage's entry data
                cp_async_wait_all();
                if (resT != k && resS != k) { load_padded_smem<NX, LD, LPS>(sm.T, Ak, sl); resT = k; }
                const bool inT = (resT == k);
                const double* Af = inT ? sm.T : sm.S;
                if (k + 1 < N) {
                    cp_async_matrix<NX, LD, LPS>(inT ? sm.S : sm.T, Ak + NX * LD, sl);
                    if (inT) resS = k + 1; else resT = k + 1;
                }
                load_matrix_smem<NX, NU, NU, LPS>(sm.Bm, Bk, sl);
                load_matrix_smem<NU, NX, NX, LPS>(sm.SB, Kk, sl);
                if (k + 1 < N) {
                    if (k + 2 < N) prefetch_l2<LPS>(Ak + 2 * NX * LD, NX * LD * 8, sl);
                    prefetch_l2<LPS>(Bk + NX * NU, NX * NU * 8, sl);
                    prefetch_l2<LPS>(Kk + NU * NX, NU * NX * 8, sl);
                }
                const int iu = k * NZ + NX + (vu ? sl : 0), ix0 = (k + 1) * NZ + r0, ix1 = (k + 1) * NZ + r1c;
                const double e_iu[3] = {ISU[ix0], ISU[ix1], ISU[iu]}, e_il[3] = {ISL[ix0], ISL[ix1], ISL[iu]};
                const double e_lu[3] = {LU[ix0], LU[ix1], LU[iu]}, e_ll[3] = {LL[ix0], LL[ix1], LL[iu]};
                double e_da[3] = {0.0, 0.0, 0.0};
                if (pass == 1) { e_da[0] = DZA[ix0]; e_da[1] = DZA[ix1]; e_da[2] = DZA[iu]; }
                double n0 = rp_scale * rp[k * NX + r0], n1 = rp_scale * rp[k * NX + r1c];
                double u = vu ? dff[k * NU + sl] : 0.0;
                __syncwarp(mask);
                if (vu) {
                    double u2 = 0.0;
#pragma unroll
                    for (int m = 0; m + 1 < NX; m += 2) {
                        u = fma(sm.SB[sl * NX + m], sm.vec[m], u);
                        u2 = fma(sm.SB[sl * NX + m + 1], sm.vec[m + 1], u2);
                    }
                    if (NX & 1) u = fma(sm.SB[sl * NX + NX - 1], sm.vec[NX - 1], u);
                    u += u2;
                    sm.vec[LD + sl] = u;
                    DZ[iu] = u;
                    ratio_entry(e_iu[2], e_il[2], e_lu[2], e_ll[2], u, e_da[2]);
                }
                {
                    double m0 = 0.0, m1 = 0.0;
#pragma unroll
                    for (int m = 0; m + 1 < NX; m += 2) {
                        const double xa = sm.vec[m], xb = sm.vec[m + 1];
                        n0 = fma(Af[r0 * LD + m], xa, n0);
                        n1 = fma(Af[r1c * LD + m], xa, n1);
                        m0 = fma(Af[r0 * LD + m + 1], xb, m0);
                        m1 = fma(Af[r1c * LD + m + 1], xb, m1);
                    }
                    if (NX & 1) {
                        const double xa = sm.vec[NX - 1];
                        n0 = fma(Af[r0 * LD + NX - 1], xa, n0);
                        n1 = fma(Af[r1c * LD + NX - 1], xa, n1);
                    }
                    n0 += m0; n1 += m1;
                }
                __syncwarp(mask);
#pragma unroll
                for (int p = 0; p < NU; ++p) {
                    const double up = sm.vec[LD + p];
                    n0 = fma(sm.Bm[r0 * NU + p], up, n0);
                    n1 = fma(sm.Bm[r1c * NU + p], up, n1);
                }
                __syncwarp(mask);
                sm.vec[r0] = n0;
                if (v1) sm.vec[r1] = n1;
                DZ[ix0] = n0;
                ratio_entry(e_iu[0], e_il[0], e_lu[0], e_ll[0], n0, e_da[0]);
                if (v1) {
                    DZ[ix1] = n1;
                    ratio_entry(e_iu[1], e_il[1], e_lu[1], e_ll[1], n1, e_da[1]);
                }
                __syncwarp(mask);
            }
            rpm = -rpm; rdm = -rdm;
            slot_min2(rpm, rdm);
            if (pass == 0) { maxp0 = -rpm; maxd0 = -rdm; }
            double apl = (-rpm > tau) ? tau / (-rpm) : 1.0;
            double adl = (-rdm > tau) ? tau / (-rdm) : 1.0;
            ap = apl; ad = adl;
            if (pass == 0) {
                // ---- P2: mu_aff and sigma ; P3: corrector right-hand side G += (sigma mu - corr_u)/su - (sigma mu - corr_l)/sl
                double acc = 0.0, dmy = 0.0;
                const double* arr[5] = {SU, SLk, LU, LL, DZA};
                entry_pass(I5{}, arr, [&](int, int idx, const double (&v)[5]) {
                    const double dz = v[4];
                    if (v[0] < kBig) { const double dl = v[2] * (ISU[idx] * dz - 1.0); acc = fma(v[0] - ap * dz, v[2] + ad * dl, acc); }
                    if (v[1] < kBig) { const double dl = -v[3] * (ISL[idx] * dz + 1.0); acc = fma(v[1] + ap * dz, v[3] + ad * dl, acc); }
                });
                slot_sum2(acc, dmy);
                const double mu_aff = n_box > 0.0 ? acc / n_box : 0.0;
                skip_corr = done || (ap >= 0.999 && ad >= 0.999 && mu_aff <= 1e-3 * mu_c);
                const double ratio_mu = mu_c > 0.0 ? mu_aff / mu_c : 0.0;
                sigma_mu = ratio_mu * ratio_mu * ratio_mu * mu_c;
                const double* arr6[6] = {ISU, ISL, LU, LL, DZA, G};
                entry_pass(I6{}, arr6, [&](int, int idx, const double (&v)[6]) {
                    const double dz = v[4];
                    double g = v[5];
                    if (v[0] > 0.0) { const double dl = v[2] * (v[0] * dz - 1.0); g = fma(sigma_mu + dz * dl, v[0], g); }
                    if (v[1] > 0.0) { const double dl = -v[3] * (v[1] * dz + 1.0); g = fma(-(sigma_mu - dz * dl), v[1], g); }
                    G[idx] = g;
                });
                __syncwarp(mask);
            }
        }
        if (skip_corr) {
            // Newton phase of this slot (full affine step, complementarity reduced 1000x): take the predictor step with the
            // fraction-to-boundary rule.  The decision is per slot -- whether the warp executed the corrector for a neighbour
            // does not change this instance's iterates (results do not depend on the position in the batch).
            corr_skipped = true;
            const double tau_c = mu_c > 1e-8 ? 0.995 : 0.9999;
            ap = (maxp0 > tau_c) ? tau_c / maxp0 : 1.0;
            ad = (maxd0 > tau_c) ? tau_c / maxd0 : 1.0;
        }
        // ---- P4: update (slots still iterating) and new complementarity
        double acc = 0.0, dmy = 0.0;
        {
            const double* arr6[6] = {SU, SLk, LU, LL, DZA, DZC};
            const bool upd = !done && chol_ok;
            entry_pass(I6{}, arr6, [&](int, int idx, const double (&v)[6]) {
                double su = v[0], sl_ = v[1], lu = v[2], ll = v[3];
                if (upd) {
                    const double dza = v[4], dz = corr_skipped ? v[4] : v[5];
                    const double so = corr_skipped ? 0.0 : 1.0;      // second-order + centring terms only with the corrector
                    if (su < kBig) {
                        const double iu = ISU[idx], ru = lu * iu;
                        const double dlu = so * (sigma_mu + dza * (ru * dza - lu)) * iu - lu + ru * dz;
                        su -= ap * dz; lu += ad * dlu;
                        SU[idx] = su; LU[idx] = lu;
                    }
                    if (sl_ < kBig) {
                        const double il = ISL[idx], rl = ll * il;
                        const double dll = so * (sigma_mu + dza * (rl * dza + ll)) * il - ll - rl * dz;
                        sl_ += ap * dz; ll += ad * dll;
                        SLk[idx] = sl_; LL[idx] = ll;
                    }
                    Z[idx] += ap * dz;
                }
                if (su < kBig) acc = fma(su, lu, acc);
                if (sl_ < kBig) acc = fma(sl_, ll, acc);
            });
            if (upd) rp_scale *= (1.0 - ap);
        }
        slot_sum2(acc, dmy);
        if (!done) {
            const double mu_new = n_box > 0.0 ? acc / n_box : 0.0;
            if (!chol_ok) {
                // a non-positive pivot late in the iteration (barrier weights ~ 1/mu) means the iterate is converged
                // to working precision: keep it; early on it is a genuine failure
                if (mu_c > 1e-9) status = 3;
                done = true;
            } else if (!(mu_new == mu_new)) { status = 3; done = true; }
            else {
                mu_c = mu_new;
                if (mu_c < a.tol_mu && rp_inf * rp_scale < a.tol_res && it >= 2) done = true;
                // infeasible / unbounded-dual detection: the iteration makes no progress (steps collapse) while the
                // dynamics residual or the complementarity stay large -> give up early instead of burning max_iter
                stalled = (ap < 1e-4 && ad < 1e-4) ? stalled + 1 : 0;
                if (!done && (stalled >= 3 || mu_c > 1e12)) { status = 2; done = true; }
            }
        }
    }
    if (!done && status == 0) status = (mu_c < 1e-9 && rp_inf * rp_scale < 1e-7) ? 0 : 1;
    if (act) it_total += it;
    // a warm start that did not converge is repeated from the cold start (whole warp: the slots run in lock-step)
    const bool retry = act && status != 0 && !structural && (warm || it_cap < a.max_iter);
    if (attempt == 1 || !__any_sync(mask, retry)) break;
    act = retry;                   // converged slots keep their solution and idle through the second attempt
    warm = false;
  }

    // ---- outputs in the reference layouts + objective  x'Hx + q'x  (qp_jit.py:401, P = 2H, q = 2 H y_nom)
    {
        __syncwarp(mask);
        double* Xo = a.dX + (size_t)b * (N + 1) * NX;
        double* Uo = a.dU + (size_t)b * N * NU;
        double* mu = a.mu + (size_t)b * N * NI;
        double* muf = a.mu_f + (size_t)b * NIF;
        double acc = 0.0, dmy = 0.0;
        for (int k = 0; k <= N; ++k) {
            const double* Qm = (k < N) ? prm->Q : prm->Qf;
            double q0 = 0.0, q1 = 0.0;
#pragma unroll
            for (int m = 0; m < NX; ++m) {
                const double y = Z[k * NZ + m] + 2.0 * Xn[k * NX + m];
                q0 = fma(Qm[r0 * NX + m], y, q0);
                q1 = fma(Qm[r1c * NX + m], y, q1);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr && !v1) continue;
                const int r = rr ? r1 : r0;
                const double z = Z[k * NZ + r];
                acc = fma(z, (rr ? q1 : q0) + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r] : 0.0), acc);
                Xo[k * NX + r] = z;
                const double lu = LU[k * NZ + r], ll = LL[k * NZ + r];
                if (k < N) { mu[k * NI + r] = lu; mu[k * NI + NZ + r] = ll; }
                else { muf[r] = lu; muf[NX + r] = ll; }
            }
            if (vu && k < N) {
                double q = 0.0;
#pragma unroll
                for (int m = 0; m < NU; ++m) q = fma(prm->R[sl * NU + m], Z[k * NZ + NX + m] + 2.0 * Un[k * NU + m], q);
                const double z = Z[k * NZ + NX + sl];
                acc = fma(z, q + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + NX + sl] : 0.0), acc);
                Uo[k * NU + sl] = z;
                mu[k * NI + NX + sl] = LU[k * NZ + NX + sl];
                mu[k * NI + NZ + NX + sl] = LL[k * NZ + NX + sl];
            }
        }
        slot_sum2(acc, dmy);
        if (sl == 0) {
            a.cost[b] = acc;
            a.iters[b] = it_total;
            a.status[b] = status;
            if (a.ws_valid != nullptr) a.ws_valid[b] = (status == 0);
        }
    }
}

}  // namespace rnmpc
