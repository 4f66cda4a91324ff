// Nominal QP kernel (reference: solver/qp_jit.py:77-273, 362-402, 487-513 handing the problem to OSQP 1.0.4).
//
//   min  sum_k dz_k' H_k dz_k + q_k' dz_k      (OSQP form: P = 2 blkdiag(Q,R,..,Qf), q = 2 H y_nom)
//   s.t. dx_{k+1} = A_k dx_k + B_k du_k + c_k,  dx_0 = x0 - x_nom0,  lo_k <= dz_k <= hi_k   (G = [I;-I])
//
// Solved per instance by a Mehrotra predictor-corrector primal-dual interior-point method whose every KKT solve is
// a Riccati factorisation / sweep over the block-tridiagonal structure (riccati_step of rnmpc_common.cuh: the same
// device function the fast-SLS kernel runs).  One instance per slot (LPS lanes), SPW instances per warp in lock-step.
//
// Per-instance iterate: one 96-byte record per box entry (stage k, component i) in global (L2-resident) workspace,
//   { s_up, s_lo, lam_up, lam_lo, 1/s_up, 1/s_lo, z, z_nom, dz_affine, dz_combined, sigma, 2*H_ii }
// so that the data of an entry arrives with six 16-byte loads from one or two cache lines (the first version kept twelve
// separate arrays: every entry touched twelve 32-byte sectors).  One interior-point iteration is four sweeps over the stages
// and one element-wise pass:
//   bwd0  Riccati factorisation + predictor right-hand side (formed on the fly from the records)
//   fwd0  predictor step, ratio test and the four sums that give mu_affine for any step lengths
//   bwd1  corrector right-hand side (on the fly) + backward vector recursion with the stored factors
//   fwd1  combined step + ratio test
//   P4    update of z, slacks, multipliers; new reciprocals, barrier diagonal and complementarity
#pragma once
#include "rnmpc_common.cuh"

namespace rnmpc {

constexpr int kRec = 12;    // doubles per entry record
constexpr int kStatusHandoff = 5;   // transient: handed to qp_block_kernel (never visible to callers)
enum { F_SU = 0, F_SL = 1, F_LU = 2, F_LL = 3, F_ISU = 4, F_ISL = 5, F_Z = 6, F_ZN = 7, F_DZA = 8, F_DZC = 9, F_SIG = 10, F_QD = 11 };

struct QpArgs {
    int B, N;
    int tightened;          // 0: bounds g_k - G z_k + 1e-10 (first QP of a fast-SLS solve); 1: g_k - G z_k - backoff_k, terminal gf - backoff_f (Q4);
                            // -1: per instance, from flags[b] & FLAG_TIGHT
    int max_iter;
    double tol_mu, tol_res;
    double mu0;             // cold start: 0 -> multipliers 1 (uncentred: s_i lam_i = s_i spreads over the box widths); > 0 -> centred,
                            // multipliers mu0 / slack; < 0 -> centred and scale-adaptive, multipliers |mu0| x mean slack / slack (default -0.3;
                            // -1 is the average complementarity of the uncentred start).  Rocket loop: 13.2 -> 9.6-10.1 iterations of the first QP
                            // for any centred variant with mu0 >= 1 or the adaptive one; mu0 <= 0.1 is worse than uncentred.
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
    // workspace
    double* E;              // [B][N+1][nz][kRec] entry records
    double* G;              // [B][N+1][nz]   gradient 2H(z + z_nom) + q, only for non-diagonal cost matrices
    double* rp;             // [B][N][nx]     dynamics residual at the initial point
    double* Sc;             // [B][N][nx]     S_{k+1} rp_k
    double* dff;            // [B][N][nu]     feed-forward
    double* Kq;             // [B][N][nu][nx]
    double* Lq;             // [B][N][nu*nu]
    int diag_cost;          // 1: Q, R, Qf are diagonal (all reference drivers) -> gradient formed on the fly from the records
    int warm;               // warm start from the workspace's previous solution of the same instance: bit0 multipliers, bit1 shifted by one
                            // stage (previous MPC step), bit2 primal iterate; a failed warm-started solve is repeated from the cold start
    int* ws_valid;          // [B] workspace holds a converged solution
    // straggler hand-off (large batches): an instance that is not done after `handoff_iter` iterations stops here with its iterate
    // in the record workspace; qp_block_kernel (resume mode, one CTA per instance) finishes it.  A launch lasts as long as its
    // slowest instance, and the slowest 2-3 % need 12-18 iterations where the mean is 9.
    int handoff_iter;       // 0: off
    int* handoff_count;     // [1] number of instances handed off (zeroed by the caller)
    int* handoff_list;      // [B] their indices
    double* handoff_state;  // [B][4] rp_scale, mu, rp_inf, n_box
    int resume;             // qp_block_kernel only: 1 = finish the instances of handoff_list
    // soft-constrained comparator (qp_warp_kernel<.., SOFT = true>): every bound row softened, cost rho gam^2 + rho1 gam per row
    // (terminal rows: rho only), solver/nlp_soft_constraints.py:222-232
    double rho, rho1;
    const double* prox;     // [B] or nullptr: proximal weight sigma_b of the SQP driver, cost + sigma/2 |dz|^2 (a damped Gauss-Newton step;
                            // vanishes at the fixed point, so the limit is unchanged)
    double* gamma;          // [B][N][ni] slack variables (output, may be nullptr)
    double* gamma_f;        // [B][ni_f]
    const int* order;       // qp_tm_kernel only: [B] instance index of the i-th work item (nullptr: identity); the engine sorts the instances
                            // by the interior-point iterations of their previous solve, longest first, so that the last instances a
                            // launch hands out are short ones (a launch ends with its slowest warp)
    int* work;              // qp_tm_kernel only: [1] next instance of the batch (zeroed by the caller): its warps draw instances from it
    long long* prof;        // qp_mma_kernel only, diagnostic ("qp_profile"): [16] cycles of instance 0 per phase, summed over the launches
};

#ifndef RNMPC_EMU
// DG: the cost matrices are diagonal (all reference drivers): the dense-cost code is compiled out of that instantiation, which
// keeps it out of the iteration loop's instruction footprint.
// WARM: the launch may warm-start (QpArgs::warm != 0); without it the second attempt, the warm branches of the initialisation and
// the checkpoints of the iteration are compiled out.
template <int NX, int NU, bool DG, bool WARM>
__global__ void __launch_bounds__(160, 2) qp_ipm_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using SL_ = Slot<NX>;
    constexpr int LD = SL_::LD, LPS = SL_::LPS, SPW = SL_::SPW;
    constexpr int NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX;
    constexpr int KU = 2;                        // stages loaded ahead in the element-wise pass
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
    const int iuc = NX + (vu ? sl : 0);
    // the (up to) three entries of every stage this lane owns: x[r0], x[r1], u[sl]
    const int ei[3] = {r0, r1c, iuc};
    const bool ev[3] = {true, v1, vu};

    const double* Ab = a.A + (size_t)b * N * NX * LD;
    const double* Bb = a.Bm + (size_t)b * N * NX * NU;
    const double* cb = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;
    const size_t eoff = (size_t)b * (N + 1) * NZ;
    double* __restrict__ E = a.E + eoff * kRec;
    double* __restrict__ G = a.G + eoff;
    double* rp = a.rp + (size_t)b * N * NX;
    double* Sc = a.Sc + (size_t)b * N * NX;
    double* dff = a.dff + (size_t)b * N * NU;
    double* Kq = a.Kq + (size_t)b * N * NU * NX;
    double* Lq = a.Lq + (size_t)b * N * NU * NU;
    constexpr bool dg = DG;

    auto ld2 = [&](int idx, int f) -> double2 { return *reinterpret_cast<const double2*>(E + (size_t)idx * kRec + f); };
    auto st2 = [&](int idx, int f, double x, double y) { *reinterpret_cast<double2*>(E + (size_t)idx * kRec + f) = make_double2(x, y); };

    auto slot_sum2 = [&](double& x, double& y) {
        __syncwarp(mask);
        sm.red[2 * sl] = x; sm.red[2 * sl + 1] = y;
        __syncwarp(mask);
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int i = 0; i < LPS; ++i) { sx += sm.red[2 * i]; sy += sm.red[2 * i + 1]; }
        x = sx; y = sy;
    };
    auto slot_max2 = [&](double& x, double& y) {
        __syncwarp(mask);
        sm.red[2 * sl] = x; sm.red[2 * sl + 1] = y;
        __syncwarp(mask);
        double sx = sm.red[0], sy = sm.red[1];
#pragma unroll
        for (int i = 1; i < LPS; ++i) { sx = fmax(sx, sm.red[2 * i]); sy = fmax(sy, sm.red[2 * i + 1]); }
        x = sx; y = sy;
    };

    int status = 0;
    const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);
    double n_box = 0.0;
    bool warm = WARM && (a.warm != 0) && (a.ws_valid != nullptr) && (a.ws_valid[b] != 0);
    int it = 0, it_total = 0;
    double rp_scale = 1.0, rp_inf = 0.0, mu_c = 1.0;
    constexpr double kMuWarm = 1e-2, kFracWarm = 1e-3;
    constexpr int kWarmCap = 15, kWarmChk = 5;
    bool structural = false;       // infeasible by construction (crossed bounds / x(0) outside its box): no retry
    bool act = true;               // this slot (re)starts in the current attempt
  for (int attempt = 0; attempt < (WARM ? 2 : 1); ++attempt) {
    if (act) { status = 0; rp_scale = 1.0; }
    double n_box_new = 0.0, s_sum = 0.0;

    // ------------------------------------------------------------------ bounds, initial point, slacks, duals -> records
    {
        const double e = tight ? 0.0 : kEpsQP;
        double bad = 0.0;
        for (int k = 0; k <= N; ++k) {
#pragma unroll
            for (int en = 0; en < 3; ++en) {
                if (!ev[en]) continue;
                const int i = ei[en];
                const int idx = k * NZ + i;
                double h_ = 1e300, l_ = -1e300, z = 0.0, zn = 0.0, qd = 0.0;
                bool box = false;
                if (en < 2) {                       // state entry
                    const double xn = Xn[k * NX + i];
                    zn = xn;
                    qd = 2.0 * ((k < N) ? prm->Q[i * NX + i] : prm->Qf[i * NX + i]);
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
                    zn = un;
                    qd = 2.0 * prm->R[p * NU + p];
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
                // the linear cost of the general mode is folded into the nominal: qd (z + zn) + q = qd (z + zn + q/qd)
                if (a.q_lin != nullptr && dg && qd != 0.0) zn += a.q_lin[eoff + idx] / qd;
                double su = 1e300, sl_ = 1e300, lu = 0.0, ll = 0.0;
                if (box) {
                    const bool hu = h_ < kBig, hl = l_ > -kBig;
                    double z0 = 0.0, lu0 = 1.0, ll0 = 1.0;
                    if (warm) {
                        // previous solution of this instance (same workspace); bit1: the multipliers of the next stage (the
                        // trajectories were shifted by one stage since); reads run ahead of this lane's own writes
                        const int ks = (a.warm & 2) ? ((en < 2) ? (k + 1 <= N ? k + 1 : N) : (k + 1 < N ? k + 1 : N - 1)) : k;
                        const double2 lw = ld2(ks * NZ + i, F_LU);
                        lu0 = lw.x; ll0 = lw.y;
                        if (a.warm & 4) z0 = E[(size_t)idx * kRec + F_Z];
                    }
                    if (hu && hl) {
                        const double m = (warm ? kFracWarm : 0.01) * (h_ - l_);
                        if (!(h_ - l_ > 1e-12)) bad = 1.0;
                        z = fmin(fmax(z0, l_ + m), h_ - m);
                    } else if (hu) z = fmin(z0, h_ - 1.0);
                    else if (hl) z = fmax(z0, l_ + 1.0);
                    if (hu) { su = h_ - z; lu = warm ? fmax(lu0, kMuWarm / su) : (a.mu0 > 0.0 ? a.mu0 / su : 1.0); n_box_new += 1.0; s_sum += su; }
                    if (hl) { sl_ = z - l_; ll = warm ? fmax(ll0, kMuWarm / sl_) : (a.mu0 > 0.0 ? a.mu0 / sl_ : 1.0); n_box_new += 1.0; s_sum += sl_; }
                }
                if (act) {
                    const double iu = (su < kBig) ? 1.0 / su : 0.0, il = (sl_ < kBig) ? 1.0 / sl_ : 0.0;
                    st2(idx, F_SU, su, sl_);
                    st2(idx, F_LU, lu, ll);
                    st2(idx, F_ISU, iu, il);
                    st2(idx, F_Z, z, zn);
                    st2(idx, F_DZA, 0.0, 0.0);
                    st2(idx, F_SIG, lu * iu + ll * il, qd);
                }
            }
        }
        slot_sum2(bad, n_box_new);
        if (act) {
            n_box = n_box_new;
            if (bad > 0.0) { status = 2; structural = true; }
        }
        __syncwarp(mask);
        if (a.mu0 < 0.0) {
            // scale-adaptive centred cold start: multipliers mu_bar / slack with mu_bar = |mu0| x the mean slack
            double dmy = 0.0;
            slot_sum2(s_sum, dmy);
            const double mu_bar = n_box_new > 0.0 ? -a.mu0 * s_sum / n_box_new : 1.0;
            if (act && !warm) {
                for (int k = 0; k <= N; ++k) {
#pragma unroll
                    for (int en = 0; en < 3; ++en) {
                        if (!ev[en] || (en == 2 && k == N)) continue;
                        const int idx = k * NZ + ei[en];
                        const double2 is = ld2(idx, F_ISU);
                        const double lu = mu_bar * is.x, ll = mu_bar * is.y;       // 1/slack is 0 for an absent bound
                        st2(idx, F_LU, lu, ll);
                        E[(size_t)idx * kRec + F_SIG] = lu * is.x + ll * is.y;
                    }
                }
            }
            __syncwarp(mask);
        }
        // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
        for (int k = 0; k < N; ++k) {
            const double* Ak = Ab + (size_t)k * NX * LD;
            const double* Bk = Bb + (size_t)k * NX * NU;
            double acc0 = cb[k * NX + r0] - E[(size_t)((k + 1) * NZ + r0) * kRec + F_Z];
            double acc1 = cb[k * NX + r1c] - E[(size_t)((k + 1) * NZ + r1c) * kRec + F_Z];
#pragma unroll
            for (int m = 0; m < NX; ++m) {
                const double xm = E[(size_t)(k * NZ + m) * kRec + F_Z];
                acc0 = fma(__ldg(Ak + r0 * LD + m), xm, acc0);
                acc1 = fma(__ldg(Ak + r1c * LD + m), xm, acc1);
            }
#pragma unroll
            for (int p = 0; p < NU; ++p) {
                const double up = E[(size_t)(k * NZ + NX + p) * kRec + F_Z];
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
        double m0 = 0.0, acc = 0.0;
        for (int k = 0; k < N; ++k) {
            m0 = fmax(m0, fabs(rp[k * NX + r0]));
            if (v1) m0 = fmax(m0, fabs(rp[k * NX + r1]));
        }
        for (int k = 0; k <= N; ++k) {
#pragma unroll
            for (int en = 0; en < 3; ++en) {
                if (!ev[en] || (en == 2 && k == N)) continue;
                const int idx = k * NZ + ei[en];
                const double2 s = ld2(idx, F_SU), l = ld2(idx, F_LU);
                if (s.x < kBig) acc = fma(s.x, l.x, acc);
                if (s.y < kBig) acc = fma(s.y, l.y, acc);
            }
        }
        double dmy = 0.0;
        slot_max2(m0, dmy);
        slot_sum2(acc, dmy);
        if (act) { rp_inf = m0; mu_c = n_box > 0.0 ? acc / n_box : 0.0; }
    }

    int stalled = 0;
    int it_slot = 0;               // iterations THIS slot took part in (the warp goes on until its slowest slot is done)
    double mu_chk = 1e300;         // complementarity at the last warm-start checkpoint
    bool done = (status != 0) || !act;
    int resT = -1, resS = -1;      // stage whose A_k is resident (or in flight, cp.async) in sm.T / sm.S
    int resB = -1, resK = -1, resL = -1;   // likewise B_k in sm.Bm, K_k in sm.SB, the Cholesky factor of stage k in sm.H
    for (it = 0; it < a.max_iter; ++it) {
        // A warm-started slot is checked every kWarmChk iterations from kWarmCap on: it goes on while it is at working precision
        // (mu < 1e-9) or still converging fast (mu fell 100x since the last checkpoint), otherwise it is abandoned for the cold
        // start.  Traced on the bench loop: the slow warm starts (0.1 % of the tightened QPs) all sat at mu = 1e-5 ... 1e-9 with
        // full steps at iteration 15, two to four iterations from done; restarting them cold cost 10-13 more.  Accepting them at
        // the cap instead left single solves accurate to ~1e-6 only (tools/failure_parity.py, profiles/r1_notes.md).
        if (warm && !done && it >= kWarmCap - kWarmChk && (it - kWarmCap) % kWarmChk == 0) {
            if (it >= kWarmCap && !(mu_c < 1e-9 && rp_inf * rp_scale < 1e-7) && !(mu_c < 1e-2 * mu_chk)) { status = 1; done = true; }
            mu_chk = mu_c;
        }
        if (__all_sync(mask, done)) break;
        if (!done) ++it_slot;
        __syncwarp(mask);
        if (!dg) {
            // dense cost matrices: gradient G = 2 H (z + z_nom) + q by explicit mat-vecs (not used by any reference driver)
            for (int k = 0; k <= N; ++k) {
                const double* Qm = (k < N) ? prm->Q : prm->Qf;
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int m = 0; m < NX; ++m) {
                    const double2 zz = ld2(k * NZ + m, F_Z);
                    acc0 = fma(Qm[r0 * NX + m], zz.x + zz.y, acc0);
                    acc1 = fma(Qm[r1c * NX + m], zz.x + zz.y, acc1);
                }
                G[k * NZ + r0] = 2.0 * acc0 + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r0] : 0.0);
                if (v1) G[k * NZ + r1] = 2.0 * acc1 + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r1] : 0.0);
                if (vu && k < N) {
                    double acc = 0.0;
#pragma unroll
                    for (int m = 0; m < NU; ++m) { const double2 zz = ld2(k * NZ + NX + m, F_Z); acc = fma(prm->R[sl * NU + m], zz.x + zz.y, acc); }
                    G[k * NZ + NX + sl] = 2.0 * acc + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + NX + sl] : 0.0);
                }
            }
            __syncwarp(mask);
        }

        double sigma_mu = 0.0, ap = 1.0, ad = 1.0;
        bool chol_ok = true;
        bool skip_corr = false, corr_skipped = false;
        double maxp0 = 0.0, maxd0 = 0.0;
        // right-hand side of one entry: gradient (+ centring / second-order terms in the corrector pass).  Split into the loads
        // (issued at the head of a stage) and the arithmetic (where the value is consumed, after the stage's matrix work), so that
        // the global round trip is off the warp's critical path.
        struct RhsRaw { double2 zz, sq, is, l; double dza; };
        auto rhs_load = [&](int pass, int idx, bool use) -> RhsRaw {
            RhsRaw r;
            r.zz = make_double2(0.0, 0.0); r.sq = r.zz; r.is = r.zz; r.l = r.zz; r.dza = 0.0;
            if (!use) return r;
            if (dg) { r.zz = ld2(idx, F_Z); r.sq = ld2(idx, F_SIG); }
            else r.zz.x = G[idx];
            if (pass == 1) { r.is = ld2(idx, F_ISU); r.l = ld2(idx, F_LU); r.dza = E[(size_t)idx * kRec + F_DZA]; }
            return r;
        };
        auto rhs_eval = [&](int pass, const RhsRaw& r) -> double {
            double g = dg ? r.sq.y * (r.zz.x + r.zz.y) : r.zz.x;
            if (pass == 1) {
                if (r.is.x > 0.0) g = fma(sigma_mu + r.dza * (r.l.x * r.is.x * r.dza - r.l.x), r.is.x, g);
                if (r.is.y > 0.0) g = fma(-(sigma_mu + r.dza * (r.l.y * r.is.y * r.dza + r.l.y)), r.is.y, g);
            }
            return g;
        };
        auto rhs_entry = [&](int pass, int idx, bool use) -> double { return rhs_eval(pass, rhs_load(pass, idx, use)); };
        for (int pass = 0; pass < 2; ++pass) {
            // the corrector's two sweeps are skipped when every slot of the warp is in the Newton phase (see below)
            if (pass == 1 && __all_sync(mask, skip_corr)) break;
            // ---- backward: p_N = rhs_N ; pass 0 also builds S_N = 2 Qf + SIG_N
            {
                const double gN0 = rhs_entry(pass, N * NZ + r0, true), gN1 = rhs_entry(pass, N * NZ + r1c, v1);
                sm.vec[r0] = gN0;
                if (v1) sm.vec[r1] = gN1;
            }
            if (pass == 0) {
                cp_async_wait_all();
                resS = -1;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = rr ? r1 : r0;
                    if (rr && !v1) continue;
                    const double sg = E[(size_t)(N * NZ + r) * kRec + F_SIG];
                    for (int c = 0; c < LD; ++c) sm.S[r * LD + c] = (c < NX) ? (2.0 * prm->Qf[r * NX + c] + (c == r ? sg : 0.0)) : 0.0;
                }
            }
            __syncwarp(mask);
            for (int k = N - 1; k >= 0; --k) {
                const double* Ak = Ab + (size_t)k * NX * LD;
                const double* Bk = Bb + (size_t)k * NX * NU;
                // this stage's scalars, fetched ahead of the (long) matrix work
                const RhsRaw raw0 = rhs_load(pass, k * NZ + r0, k > 0), raw1 = rhs_load(pass, k * NZ + r1c, k > 0 && v1);
                const RhsRaw rawu = rhs_load(pass, k * NZ + iuc, vu);
                if (k > 0) prefetch_l2<LPS>(E + (size_t)(k - 1) * NZ * kRec, NZ * kRec * 8, sl);      // next stage's records
                const double rp0 = rp[k * NX + r0], rp1 = rp[k * NX + r1c];
                double sc0 = 0.0, sc1 = 0.0;
                // Sc_k = S_{k+1} rp_k while S_{k+1} is still in shared memory (pass 0); the corrector re-reads it
                auto pre_step = [&]() {
                    sm.vec[LD + r0] = rp0;
                    if (v1) sm.vec[LD + r1] = rp1;
                    __syncwarp(mask);
                    double sc0b = 0.0, sc1b = 0.0;
#pragma unroll 4
                    for (int c = 0; c + 1 < NX; c += 2) {
                        const double2 cv = *reinterpret_cast<const double2*>(sm.vec + LD + c);
                        const double2 s0v = *reinterpret_cast<const double2*>(sm.S + r0 * LD + c);
                        const double2 s1v = *reinterpret_cast<const double2*>(sm.S + r1c * LD + c);
                        sc0 = fma(s0v.x, cv.x, sc0); sc0b = fma(s0v.y, cv.y, sc0b);
                        sc1 = fma(s1v.x, cv.x, sc1); sc1b = fma(s1v.y, cv.y, sc1b);
                    }
                    if (NX & 1) {
                        const double cv = sm.vec[LD + NX - 1];
                        sc0 = fma(sm.S[r0 * LD + NX - 1], cv, sc0);
                        sc1 = fma(sm.S[r1c * LD + NX - 1], cv, sc1);
                    }
                    sc0 += sc0b; sc1 += sc1b;
                    Sc[k * NX + r0] = sc0;
                    if (v1) Sc[k * NX + r1] = sc1;
                    __syncwarp(mask);
                };
                const double* Abuf = sm.T;
                // vector recursion of the KKT solve; A_k is in Abuf, B_k in sm.Bm, K columns / factor in registers
                auto vec_step = [&](const double (&k0)[NU], const double (&k1)[NU], const double (&L)[NU][NU]) {
                    // g = p_{k+1} + Sc
                    const double g0 = sm.vec[r0] + rp_scale * sc0;
                    const double g1 = v1 ? (sm.vec[r1] + rp_scale * sc1) : 0.0;
                    sm.vec[LD + r0] = g0;
                    if (v1) sm.vec[LD + r1] = g1;
                    if (pass == 1) { if (k == N - 1 || k == 0) cp_async_wait_all(); else cp_async_wait<3>(); }     // B_k has landed (A, K, L of the next stage may be in flight)
                    __syncwarp(mask);
                    // t = ru + B' g
                    if (vu) {
                        double t = rhs_eval(pass, rawu), t2 = 0.0;
#pragma unroll
                        for (int i = 0; i + 1 < NX; i += 2) {
                            t = fma(sm.Bm[i * NU + sl], sm.vec[LD + i], t);
                            t2 = fma(sm.Bm[(i + 1) * NU + sl], sm.vec[LD + i + 1], t2);
                        }
                        if (NX & 1) t = fma(sm.Bm[(NX - 1) * NU + sl], sm.vec[LD + NX - 1], t);
                        sm.vec[2 * LD + sl] = t + t2;
                    }
                    __syncwarp(mask);
                    if (pass == 1 && k > 0) { stage_block_async<NX * NU, LPS>(sm.Bm, Bk - NX * NU, sl); resB = k - 1; }
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
                        double p0 = rhs_eval(pass, raw0), p1 = v1 ? rhs_eval(pass, raw1) : 0.0, p0b = 0.0, p1b = 0.0;
#pragma unroll 4
                        for (int i = 0; i + 1 < NX; i += 2) {
                            const double2 gi = *reinterpret_cast<const double2*>(sm.vec + LD + i);
                            const double2 av = *reinterpret_cast<const double2*>(Abuf + i * LD + r0);
                            const double2 aw = *reinterpret_cast<const double2*>(Abuf + (i + 1) * LD + r0);
                            p0 = fma(av.x, gi.x, p0); p1 = fma(av.y, gi.x, p1);
                            p0b = fma(aw.x, gi.y, p0b); p1b = fma(aw.y, gi.y, p1b);
                        }
                        if (NX & 1) {
                            const double gi = sm.vec[LD + NX - 1];
                            const double2 av = *reinterpret_cast<const double2*>(Abuf + (NX - 1) * LD + r0);
                            p0 = fma(av.x, gi, p0); p1 = fma(av.y, gi, p1);
                        }
                        p0 += p0b; p1 += p1b;
#pragma unroll
                        for (int p = 0; p < NU; ++p) { p0 = fma(k0[p], tt[p], p0); p1 = fma(k1[p], tt[p], p1); }
                        __syncwarp(mask);
                        sm.vec[r0] = p0;
                        if (v1) sm.vec[r1] = p1;
                    }
                    __syncwarp(mask);
                };
                if (pass == 0) {
                    double cx0 = 0.0, cx1 = 0.0;
                    if (k > 0) {
                        cx0 = raw0.sq.x; cx1 = v1 ? raw1.sq.x : 0.0;
                        if (dg) { cx0 += raw0.sq.y; cx1 += v1 ? raw1.sq.y : 0.0; }       // diagonal stage cost folded in
                        else { cx0 = E[(size_t)(k * NZ + r0) * kRec + F_SIG]; if (v1) cx1 = E[(size_t)(k * NZ + r1) * kRec + F_SIG]; }
                    }
                    if (vu) sm.cu[sl] = dg ? rawu.sq.x : E[(size_t)(k * NZ + NX + sl) * kRec + F_SIG];
                    const bool ok = riccati_step<NX, NU>(
                        sm, sl, mask, Ak, Bk, dg ? nullptr : prm->Q2, prm->R2, cx0, cx1, Kq + (size_t)k * NU * NX, pre_step,
                        [&](const double (&k0)[NU], const double (&k1)[NU], const double (&L)[NU][NU]) {
                            if (sl == 0) {
#pragma unroll
                                for (int p = 0; p < NU; ++p)
#pragma unroll
                                    for (int q = 0; q < NU; ++q) Lq[k * NU * NU + p * NU + q] = L[p][q];
                            }
                            vec_step(k0, k1, L);
                        },
                        k > 0, k > 0 ? Ak - NX * LD : nullptr, k > 0 ? Bk - NX * NU : nullptr, resT == k, k > 0 ? Ak - NX * LD : nullptr,
                        resB == k, k > 0 ? Bk - NX * NU : nullptr);
                    resT = (k > 0) ? k - 1 : 0;
                    resB = (k > 0) ? k - 1 : 0;
                    resK = -1; resL = -1;          // sm.SB / sm.H were scratch of the factorisation
                    chol_ok = chol_ok && ok;
                } else {
                    double k0[NU], k1[NU], L[NU][NU];
                    // steady state: groups committed in the order A_k, K_k, L_k (head of the previous stage), B_k (its middle)
                    if (k == N - 1) cp_async_wait_all(); else cp_async_wait<1>();
                    if (resT != k && resS != k) { load_padded_smem<NX, LD, LPS>(sm.T, Ak, sl); resT = k; }
                    const bool inT = (resT == k);
                    Abuf = inT ? sm.T : sm.S;
                    if (resB != k) { load_matrix_smem<NX, NU, NU, LPS>(sm.Bm, Bk, sl); resB = k; }
                    if (resK != k) { load_matrix_smem<NU, NX, NX, LPS>(sm.SB, Kq + (size_t)k * NU * NX, sl); resK = k; }
                    if (resL != k) { load_matrix_smem<NU, NU, NU, LPS>(sm.H, Lq + k * NU * NU, sl); resL = k; }
                    if (k > 1) {
                        prefetch_l2<LPS>(Ak - 2 * NX * LD, NX * LD * 8, sl);     // two stages ahead: the asynchronous copies of the next stage hit L2
                        prefetch_l2<LPS>(Bk - 2 * NX * NU, NX * NU * 8, sl);
                        prefetch_l2<LPS>(Kq + (size_t)(k - 2) * NU * NX, NU * NX * 8, sl);
                    }
                    __syncwarp(mask);
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        k0[p] = sm.SB[p * NX + r0];
                        k1[p] = v1 ? sm.SB[p * NX + r1c] : 0.0;
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = sm.H[p * NU + q];
                    }
                    __syncwarp(mask);
                    if (k > 0) {       // next stage's A into the other buffer; the gain and the factor are in registers: fetch the next stage's
                        cp_async_matrix<NX, LD, LPS>(inT ? sm.S : sm.T, Ak - NX * LD, sl);
                        if (inT) resS = k - 1; else resT = k - 1;
                        stage_block_async<NU * NX, LPS>(sm.SB, Kq + (size_t)(k - 1) * NU * NX, sl);
                        stage_block_async<NU * NU, LPS>(sm.H, Lq + (k - 1) * NU * NU, sl);
                        resK = k - 1; resL = k - 1;
                    }
                    sc0 = Sc[k * NX + r0];
                    sc1 = Sc[k * NX + r1c];
                    __syncwarp(mask);
                    vec_step(k0, k1, L);
                }
            }
            // ---- forward: dx_0 = 0 ; du_k = K dx_k + d_k ; dx_{k+1} = A dx_k + B du_k + rp_k ; fused ratio test (+ mu_aff sums)
            const double tau = pass == 0 ? 1.0 : (mu_c > 1e-8 ? 0.995 : 0.9999);
            // largest dz/s and -dlam/lam over the entries: alpha = min(1, tau / max)
            double rpm = 0.0, rdm = 0.0;
            // (s + ap ds)(lam + ad dlam) summed over the entries = S1 + ad S2 + ap S3 + ap ad S4 for any step lengths, with
            // S1 = n mu, S3 = sum ds lam, S4 = sum ds dlam and (predictor: s dlam + lam ds = -s lam) S2 = -S1 - S3
            double S3 = 0.0, S4 = 0.0;
            auto ratio_entry = [&](double2 is, double2 l, double dz, double dza) {
                if (is.x > 0.0) {
                    const double ru = l.x * is.x;
                    double tgt = 0.0;
                    if (pass == 1) tgt = sigma_mu + dza * (ru * dza - l.x);
                    const double dlu = tgt * is.x - l.x + ru * dz;
                    rpm = fmax(rpm, dz * is.x);
                    if (dlu < 0.0) rdm = fmax(rdm, -dlu * fast_rcp(l.x));
                    if (pass == 0) { S3 = fma(-dz, l.x, S3); S4 = fma(-dz, dlu, S4); }
                }
                if (is.y > 0.0) {
                    const double rl = l.y * is.y;
                    double tgt = 0.0;
                    if (pass == 1) tgt = sigma_mu + dza * (rl * dza + l.y);
                    const double dll = tgt * is.y - l.y - rl * dz;
                    rpm = fmax(rpm, -dz * is.y);
                    if (dll < 0.0) rdm = fmax(rdm, -dll * fast_rcp(l.y));
                    if (pass == 0) { S3 = fma(dz, l.y, S3); S4 = fma(dz, dll, S4); }
                }
            };
            // entry data {x_{k+1}[r0], x_{k+1}[r1], u_k[sl]} of the current stage
            double2 q_is[3], q_l[3];
            double q_da[3] = {0.0, 0.0, 0.0};
            const int fdz = pass == 0 ? F_DZA : F_DZC;
            __syncwarp(mask);
            sm.vec[r0] = 0.0;
            if (v1) sm.vec[r1] = 0.0;
            E[(size_t)r0 * kRec + fdz] = 0.0;
            if (v1) E[(size_t)r1 * kRec + fdz] = 0.0;
            __syncwarp(mask);
            for (int k = 0; k < N; ++k) {
                const double* Ak = Ab + (size_t)k * NX * LD;
                const double* Bk = Bb + (size_t)k * NX * NU;
                const double* Kk = Kq + (size_t)k * NU * NX;
                const int iu = k * NZ + iuc, ix0 = (k + 1) * NZ + r0, ix1 = (k + 1) * NZ + r1c;
                q_is[0] = ld2(ix0, F_ISU); q_is[1] = ld2(ix1, F_ISU); q_is[2] = ld2(iu, F_ISU);
                q_l[0] = ld2(ix0, F_LU); q_l[1] = ld2(ix1, F_LU); q_l[2] = ld2(iu, F_LU);
                if (pass == 1) { q_da[0] = E[(size_t)ix0 * kRec + F_DZA]; q_da[1] = E[(size_t)ix1 * kRec + F_DZA]; q_da[2] = E[(size_t)iu * kRec + F_DZA]; }
                double n0 = rp_scale * rp[k * NX + r0], n1 = rp_scale * rp[k * NX + r1c];
                double u = vu ? dff[k * NU + sl] : 0.0;
                // A_k, B_k, K_k are resident or in flight.  In the steady state the groups were committed in the order A_k, K_k, B_k
                // (head, middle, end of the previous stage): B_k, needed last, may still be in flight here.
                if (k == 0) cp_async_wait_all(); else cp_async_wait<1>();
                if (resT != k && resS != k) { load_padded_smem<NX, LD, LPS>(sm.T, Ak, sl); resT = k; }
                const bool inT = (resT == k);
                const double* Af = inT ? sm.T : sm.S;
                if (resB != k) { load_matrix_smem<NX, NU, NU, LPS>(sm.Bm, Bk, sl); resB = k; }
                if (resK != k) { load_matrix_smem<NU, NX, NX, LPS>(sm.SB, Kk, sl); resK = k; }
                if (k + 1 < N) {
                    cp_async_matrix<NX, LD, LPS>(inT ? sm.S : sm.T, Ak + NX * LD, sl);
                    if (inT) resS = k + 1; else resT = k + 1;
                    prefetch_l2<LPS>(E + ((size_t)(k + 1) * NZ + NX) * kRec, NZ * kRec * 8, sl);      // records u_{k+1}, x_{k+2}
                    if (k + 2 < N) {
                        prefetch_l2<LPS>(Ak + 2 * NX * LD, NX * LD * 8, sl);
                        prefetch_l2<LPS>(Bk + 2 * NX * NU, NX * NU * 8, sl);
                        prefetch_l2<LPS>(Kk + 2 * NU * NX, NU * NX * 8, sl);
                    }
                }
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
                    E[(size_t)iu * kRec + fdz] = u;
                    ratio_entry(q_is[2], q_l[2], u, q_da[2]);
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
                if (k == 0) cp_async_wait_all(); else cp_async_wait<1>();      // B_k has landed (A_{k+1} may be in flight)
                __syncwarp(mask);
                // every lane is done with K_k (and dx_k): the next stage's gain streams into sm.SB
                if (k + 1 < N) { stage_block_async<NU * NX, LPS>(sm.SB, Kk + NU * NX, sl); resK = k + 1; }
#pragma unroll
                for (int p = 0; p < NU; ++p) {
                    const double up = sm.vec[LD + p];
                    n0 = fma(sm.Bm[r0 * NU + p], up, n0);
                    n1 = fma(sm.Bm[r1c * NU + p], up, n1);
                }
                __syncwarp(mask);
                if (k + 1 < N) { stage_block_async<NX * NU, LPS>(sm.Bm, Bk + NX * NU, sl); resB = k + 1; }
                sm.vec[r0] = n0;
                if (v1) sm.vec[r1] = n1;
                E[(size_t)ix0 * kRec + fdz] = n0;
                ratio_entry(q_is[0], q_l[0], n0, q_da[0]);
                if (v1) {
                    E[(size_t)ix1 * kRec + fdz] = n1;
                    ratio_entry(q_is[1], q_l[1], n1, q_da[1]);
                }
                __syncwarp(mask);
            }
            slot_max2(rpm, rdm);
            if (pass == 0) { maxp0 = rpm; maxd0 = rdm; }
            ap = (rpm > tau) ? tau / rpm : 1.0;
            ad = (rdm > tau) ? tau / rdm : 1.0;
            if (pass == 0) {
                // mu_aff from the four sums, sigma = (mu_aff/mu)^3
                slot_sum2(S3, S4);
                const double mu_aff = n_box > 0.0 ? (mu_c * n_box * (1.0 - ad) + (ap - ad) * S3 + ap * ad * S4) / n_box : 0.0;
                skip_corr = done || (ap >= 0.999 && ad >= 0.999 && mu_aff <= 1e-3 * mu_c);
                const double ratio_mu = mu_c > 0.0 ? mu_aff / mu_c : 0.0;
                sigma_mu = ratio_mu * ratio_mu * ratio_mu * mu_c;
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
        // ---- P4: update (slots still iterating), new reciprocals / barrier diagonal / complementarity
        double acc = 0.0, dmy = 0.0;
        {
            const bool upd = !done && chol_ok;
            const double so = corr_skipped ? 0.0 : 1.0;      // second-order + centring terms only with the corrector
            for (int k0 = 0; k0 <= N; k0 += KU) {
                double2 vs[KU][3], vl[KU][3], vi[KU][3], vz[KU][3], vd[KU][3];
#pragma unroll
                for (int j = 0; j < KU; ++j) {
                    const int k = (k0 + j <= N) ? (k0 + j) : N;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        const int idx = k * NZ + ei[e];
                        vs[j][e] = ld2(idx, F_SU); vl[j][e] = ld2(idx, F_LU); vi[j][e] = ld2(idx, F_ISU);
                        vz[j][e] = ld2(idx, F_Z); vd[j][e] = ld2(idx, F_DZA);
                    }
                }
#pragma unroll
                for (int j = 0; j < KU; ++j) {
                    const int k = k0 + j;
                    if (k > N) continue;
#pragma unroll
                    for (int e = 0; e < 3; ++e) {
                        if (!ev[e] || (e == 2 && k == N)) continue;
                        const int idx = k * NZ + ei[e];
                        double su = vs[j][e].x, sl_ = vs[j][e].y, lu = vl[j][e].x, ll = vl[j][e].y;
                        if (upd) {
                            const double dza = vd[j][e].x, dz = corr_skipped ? vd[j][e].x : vd[j][e].y;
                            double iu = 0.0, il = 0.0;
                            if (su < kBig) {
                                const double ru = lu * vi[j][e].x;
                                const double dlu = so * (sigma_mu + dza * (ru * dza - lu)) * vi[j][e].x - lu + ru * dz;
                                su -= ap * dz; lu += ad * dlu;
                                iu = fast_rcp(su);
                            }
                            if (sl_ < kBig) {
                                const double rl = ll * vi[j][e].y;
                                const double dll = so * (sigma_mu + dza * (rl * dza + ll)) * vi[j][e].y - ll - rl * dz;
                                sl_ += ap * dz; ll += ad * dll;
                                il = fast_rcp(sl_);
                            }
                            st2(idx, F_SU, su, sl_);
                            st2(idx, F_LU, lu, ll);
                            st2(idx, F_ISU, iu, il);
                            E[(size_t)idx * kRec + F_Z] = vz[j][e].x + ap * dz;
                            E[(size_t)idx * kRec + F_SIG] = lu * iu + ll * il;
                        }
                        if (su < kBig) acc = fma(su, lu, acc);
                        if (sl_ < kBig) acc = fma(sl_, ll, acc);
                    }
                }
            }
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
                if (!WARM && !done && a.handoff_iter > 0 && it + 1 >= a.handoff_iter) { status = kStatusHandoff; done = true; }
            }
        }
    }
    cp_async_wait_all();
    // out of iterations (the full budget): accepted at working precision
    if (!done && status == 0) status = (mu_c < 1e-9 && rp_inf * rp_scale < 1e-7) ? 0 : 1;
    if (act) it_total += it_slot;
    if (!WARM && status == kStatusHandoff) {
        // the iterate of this instance (records, rp) is complete in the workspace; the scalars of the iteration go with it
        __syncwarp(mask);
        if (sl == 0) {
            const int pos = atomicAdd(a.handoff_count, 1);
            a.handoff_list[pos] = b;
            double* hs = a.handoff_state + (size_t)b * 4;
            hs[0] = rp_scale; hs[1] = mu_c; hs[2] = rp_inf; hs[3] = n_box;
        }
    }
    // a warm start that did not converge is repeated from the cold start (whole warp: the slots run in lock-step)
    const bool retry = act && status != 0 && !structural && warm;
    if (!WARM || attempt == 1 || !__any_sync(mask, retry)) break;
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
                const double y = E[(size_t)(k * NZ + m) * kRec + F_Z] + 2.0 * Xn[k * NX + m];
                q0 = fma(Qm[r0 * NX + m], y, q0);
                q1 = fma(Qm[r1c * NX + m], y, q1);
            }
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (rr && !v1) continue;
                const int r = rr ? r1 : r0;
                const double z = E[(size_t)(k * NZ + r) * kRec + F_Z];
                acc = fma(z, (rr ? q1 : q0) + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + r] : 0.0), acc);
                Xo[k * NX + r] = z;
                const double2 l = ld2(k * NZ + r, F_LU);
                if (k < N) { mu[k * NI + r] = l.x; mu[k * NI + NZ + r] = l.y; }
                else { muf[r] = l.x; muf[NX + r] = l.y; }
            }
            if (vu && k < N) {
                double q = 0.0;
#pragma unroll
                for (int m = 0; m < NU; ++m) q = fma(prm->R[sl * NU + m], E[(size_t)(k * NZ + NX + m) * kRec + F_Z] + 2.0 * Un[k * NU + m], q);
                const double z = E[(size_t)(k * NZ + NX + sl) * kRec + F_Z];
                acc = fma(z, q + (a.q_lin != nullptr ? a.q_lin[eoff + k * NZ + NX + sl] : 0.0), acc);
                Uo[k * NU + sl] = z;
                const double2 l = ld2(k * NZ + NX + sl, F_LU);
                mu[k * NI + NX + sl] = l.x;
                mu[k * NI + NZ + NX + sl] = l.y;
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

#endif  // RNMPC_EMU

}  // namespace rnmpc
