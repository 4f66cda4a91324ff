// Nominal QP kernel (reference: solver/qp_jit.py:77-273, 362-402, 487-513 handing the problem to OSQP 1.0.4).
//
//   min  sum_k dz_k' H_k dz_k + q_k' dz_k      (OSQP form: P = 2 blkdiag(Q,R,..,Qf), q = 2 H y_nom)
//   s.t. dx_{k+1} = A_k dx_k + B_k du_k + c_k,  dx_0 = x0 - x_nom0,  lo_k <= dz_k <= hi_k   (G = [I;-I])
//
// Solved per instance by a Mehrotra predictor-corrector primal-dual interior-point method whose every KKT solve is
// a Riccati factorisation / sweep over the block-tridiagonal structure (riccati_step of rnmpc_common.cuh: the same
// device function the fast-SLS kernel runs).  One instance per slot (LPS lanes), SPW instances per warp in
// lock-step; dual variables are written straight into the reference's dual_mu / dual_mu_f layout.
#pragma once
#include "rnmpc_common.cuh"

namespace rnmpc {

struct QpArgs {
    int B, N;
    int tightened;          // 0: bounds g_k - G z_k + 1e-10 (first QP of a fast-SLS solve); 1: g_k - G z_k - backoff_k, terminal gf - backoff_f (Q4);
                            // -1: per instance, from flags[b] & FLAG_TIGHT
    int max_iter;
    double tol_mu, tol_res;
    const double* A;        // [B][N][nx][nx]
    const double* Bm;       // [B][N][nx][nu]
    const double* c;        // [B][N][nx]
    const double* Xnom;     // [B][N+1][nx]
    const double* Unom;     // [B][N][nu]
    const double* x0;       // [B][nx] measured state
    const double* backoff;  // [B][N][ni]   (tightened only)
    const double* backoff_f;// [B][ni_f]
    const int* flags;       // [B] or nullptr: instances with (flags & skip_mask) do no work
    int skip_mask;
    // outputs
    double* dX;             // [B][N+1][nx]
    double* dU;             // [B][N][nu]
    double* mu;             // [B][N][ni]
    double* mu_f;           // [B][ni_f]
    double* cost;           // [B]
    int* iters;             // [B]
    int* status;            // [B] 0 ok, 1 not converged, 2 infeasible / numerical failure
    // workspace
    double* hi;             // [B][N+1][nz]
    double* lo;             // [B][N+1][nz]
    double* gx;             // [B][N+1][nx]   2Q(x+xnom)
    double* gu;             // [B][N][nu]
    double* dXa;            // [B][N+1][nx]   affine (predictor) step
    double* dUa;            // [B][N][nu]
    double* dXc;            // [B][N+1][nx]   combined step
    double* dUc;            // [B][N][nu]
    double* rp;             // [B][N][nx]     dynamics residual at the initial point
    double* Sc;             // [B][N][nx]     S_{k+1} rp_k
    double* dff;            // [B][N][nu]     feed-forward
    double* Kq;             // [B][N][nu][nx]
    double* Lq;             // [B][N][nu*nu]
};

template <int NX, int NU>
__global__ void __launch_bounds__(128) qp_ipm_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using SL = Slot<NX>;
    constexpr int LD = SL::LD, LPS = SL::LPS, SPW = SL::SPW;
    constexpr int NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX;
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
    const double* Ab = a.A + (size_t)b * N * NX * NX;
    const double* Bb = a.Bm + (size_t)b * N * NX * NU;
    const double* cb = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;
    double* X = a.dX + (size_t)b * (N + 1) * NX;
    double* U = a.dU + (size_t)b * N * NU;
    double* mu = a.mu + (size_t)b * N * NI;
    double* muf = a.mu_f + (size_t)b * NIF;
    double* hi = a.hi + (size_t)b * (N + 1) * NZ;
    double* lo = a.lo + (size_t)b * (N + 1) * NZ;
    double* gx = a.gx + (size_t)b * (N + 1) * NX;
    double* gu = a.gu + (size_t)b * N * NU;
    double* dXa = a.dXa + (size_t)b * (N + 1) * NX;
    double* dUa = a.dUa + (size_t)b * N * NU;
    double* dXc = a.dXc + (size_t)b * (N + 1) * NX;
    double* dUc = a.dUc + (size_t)b * N * NU;
    double* rp = a.rp + (size_t)b * N * NX;
    double* Sc = a.Sc + (size_t)b * N * NX;
    double* dff = a.dff + (size_t)b * N * NU;
    double* Kq = a.Kq + (size_t)b * N * NU * NX;
    double* Lq = a.Lq + (size_t)b * N * NU * NU;

    // dual variable accessors in the reference layout: stage k<N: mu[k][i]; terminal: mu_f
    auto lamU_x = [&](int k, int r) -> double& { return (k < N) ? mu[k * NI + r] : muf[r]; };
    auto lamL_x = [&](int k, int r) -> double& { return (k < N) ? mu[k * NI + NZ + r] : muf[NX + r]; };
    auto lamU_u = [&](int k, int p) -> double& { return mu[k * NI + NX + p]; };
    auto lamL_u = [&](int k, int p) -> double& { return mu[k * NI + NZ + NX + p]; };

    // slot-wide reductions through shared memory
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
    // visit every box-constrained entry this lane owns: f(value index info)
    // x entries: stages 1..N rows r0,(r1); u entries: stages 0..N-1, p = sl, sl+LPS, ...
    int status = 0;
    const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);

    // ------------------------------------------------------------------ bounds + initial point
    {
        const double e = tight ? 0.0 : kEpsQP;
        for (int k = 0; k <= N; ++k) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = rr ? r1 : r0;
                if (rr && !v1) continue;
                double h_, l_;
                const double xn = Xn[k * NX + r];
                if (k < N) {
                    const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + r] : 0.0;
                    const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + r] : 0.0;
                    h_ = (prm->xub[r] - xn) - bu + e;
                    l_ = -((xn - prm->xlb[r]) - bl + e);
                } else if (tight) {
                    h_ = prm->xubf[r] - a.backoff_f[(size_t)b * NIF + r];          // Q4: terminal bound not shifted by x_N
                    l_ = -(-prm->xlbf[r] - a.backoff_f[(size_t)b * NIF + NX + r]);
                } else {
                    h_ = (prm->xubf[r] - xn) + e;
                    l_ = -((xn - prm->xlbf[r]) + e);
                }
                hi[k * NZ + r] = h_;
                lo[k * NZ + r] = l_;
            }
            if (k < N) {
                for (int p = sl; p < NU; p += LPS) {
                    const double un = Un[k * NU + p];
                    const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + NX + p] : 0.0;
                    const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + NX + p] : 0.0;
                    hi[k * NZ + NX + p] = (prm->uub[p] - un) - bu + e;
                    lo[k * NZ + NX + p] = -((un - prm->ulb[p]) - bl + e);
                }
            }
        }
        // initial point: x_0 = x0 - xnom_0 (pinned), others = 0 clipped 1 % inside the box, duals = 1
        double bad = 0.0, dummy = 0.0;
        for (int k = 0; k <= N; ++k) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = rr ? r1 : r0;
                if (rr && !v1) continue;
                if (k == 0) {
                    const double d0 = a.x0[(size_t)b * NX + r] - Xn[r];
                    X[r] = d0;
                    mu[r] = 0.0; mu[NZ + r] = 0.0;
                    // the stage-0 state rows are implied by the pinned x_0; violated rows mean an infeasible QP
                    if (d0 > hi[r] + 1e-9 || d0 < lo[r] - 1e-9) bad = 1.0;
                    continue;
                }
                const double h_ = hi[k * NZ + r], l_ = lo[k * NZ + r];
                const bool hu = h_ < kBig, hl = l_ > -kBig;
                double z = 0.0;
                if (hu && hl) {
                    const double m = 0.01 * (h_ - l_);
                    if (!(h_ - l_ > 1e-12)) bad = 1.0;
                    z = fmin(fmax(0.0, l_ + m), h_ - m);
                } else if (hu) z = fmin(0.0, h_ - 1.0);
                else if (hl) z = fmax(0.0, l_ + 1.0);
                X[k * NX + r] = z;
                lamU_x(k, r) = hu ? 1.0 : 0.0;
                lamL_x(k, r) = hl ? 1.0 : 0.0;
            }
            if (k < N) {
                for (int p = sl; p < NU; p += LPS) {
                    const double h_ = hi[k * NZ + NX + p], l_ = lo[k * NZ + NX + p];
                    const bool hu = h_ < kBig, hl = l_ > -kBig;
                    double z = 0.0;
                    if (hu && hl) {
                        const double m = 0.01 * (h_ - l_);
                        if (!(h_ - l_ > 1e-12)) bad = 1.0;
                        z = fmin(fmax(0.0, l_ + m), h_ - m);
                    } else if (hu) z = fmin(0.0, h_ - 1.0);
                    else if (hl) z = fmax(0.0, l_ + 1.0);
                    U[k * NU + p] = z;
                    lamU_u(k, p) = hu ? 1.0 : 0.0;
                    lamL_u(k, p) = hl ? 1.0 : 0.0;
                }
            }
        }
        slot_sum2(bad, dummy);
        if (bad > 0.0) status = 2;
        __syncwarp(mask);
        // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
        for (int k = 0; k < N; ++k) {
            const double* Ak = Ab + (size_t)k * NX * NX;
            const double* Bk = Bb + (size_t)k * NX * NU;
            double acc0 = cb[k * NX + r0] - X[(k + 1) * NX + r0];
            double acc1 = cb[k * NX + r1c] - X[(k + 1) * NX + r1c];
            for (int m = 0; m < NX; ++m) {
                const double xm = X[k * NX + m];
                acc0 = fma(__ldg(Ak + r0 * NX + m), xm, acc0);
                acc1 = fma(__ldg(Ak + r1c * NX + m), xm, acc1);
            }
            for (int p = 0; p < NU; ++p) {
                const double up = U[k * NU + p];
                acc0 = fma(__ldg(Bk + r0 * NU + p), up, acc0);
                acc1 = fma(__ldg(Bk + r1c * NU + p), up, acc1);
            }
            rp[k * NX + r0] = acc0;
            if (v1) rp[k * NX + r1] = acc1;
        }
        __syncwarp(mask);
    }

    double rp_scale = 1.0;
    double rp_inf = 0.0;
    {
        double m0 = 0.0, dmy = 0.0;
        for (int k = 0; k < N; ++k) {
            m0 = fmax(m0, fabs(rp[k * NX + r0]));
            if (v1) m0 = fmax(m0, fabs(rp[k * NX + r1]));
        }
        m0 = -m0;
        slot_min2(m0, dmy);
        rp_inf = -m0;
    }

    // complementarity measure
    auto compl_mu = [&]() -> double {
        double acc = 0.0, cnt = 0.0;
        for (int k = 1; k <= N; ++k) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int r = rr ? r1 : r0;
                if (rr && !v1) continue;
                const double z = X[k * NX + r], h_ = hi[k * NZ + r], l_ = lo[k * NZ + r];
                if (h_ < kBig) { acc = fma(h_ - z, lamU_x(k, r), acc); cnt += 1.0; }
                if (l_ > -kBig) { acc = fma(z - l_, lamL_x(k, r), acc); cnt += 1.0; }
            }
        }
        for (int k = 0; k < N; ++k)
            for (int p = sl; p < NU; p += LPS) {
                const double z = U[k * NU + p], h_ = hi[k * NZ + NX + p], l_ = lo[k * NZ + NX + p];
                if (h_ < kBig) { acc = fma(h_ - z, lamU_u(k, p), acc); cnt += 1.0; }
                if (l_ > -kBig) { acc = fma(z - l_, lamL_u(k, p), acc); cnt += 1.0; }
            }
        slot_sum2(acc, cnt);
        return cnt > 0.0 ? acc / cnt : 0.0;
    };

    // Newton direction pieces for one box entry; given step dz returns (ds_u, ds_l, dlu, dll)
    // target = sigma*mu - corr ;  dlu = (tu)/su - lu + (lu/su) dz ; dll = (tl)/sl - ll - (ll/sl) dz
    int it = 0;
    bool done = (status != 0);
    double mu_c = compl_mu();
    for (; it < a.max_iter; ++it) {
        if (__all_sync(mask, done)) break;
        // ---- gradient of the quadratic: gx_k = 2 Q_k (x_k + xnom_k), gu_k = 2 R (u_k + unom_k)
        __syncwarp(mask);
        for (int k = 1; k <= N; ++k) {
            const double* Qm = (k < N) ? prm->Q : prm->Qf;
            double acc0 = 0.0, acc1 = 0.0;
            for (int m = 0; m < NX; ++m) {
                const double y = X[k * NX + m] + Xn[k * NX + m];
                acc0 = fma(Qm[r0 * NX + m], y, acc0);
                acc1 = fma(Qm[r1c * NX + m], y, acc1);
            }
            gx[k * NX + r0] = 2.0 * acc0;
            if (v1) gx[k * NX + r1] = 2.0 * acc1;
        }
        for (int k = 0; k < N; ++k)
            for (int p = sl; p < NU; p += LPS) {
                double acc = 0.0;
                for (int m = 0; m < NU; ++m) acc = fma(prm->R[p * NU + m], U[k * NU + m] + Un[k * NU + m], acc);
                gu[k * NU + p] = 2.0 * acc;
            }
        __syncwarp(mask);

        // two Riccati solves per iteration: pass 0 = factorisation + predictor, pass 1 = corrector
        double sigma_mu = 0.0;
        double ap = 1.0, ad = 1.0;
        for (int pass = 0; pass < 2; ++pass) {
            // right-hand side of one box entry
            auto rhs_x = [&](int k, int r) -> double {
                double g = gx[k * NX + r];
                if (pass == 1) {
                    const double z = X[k * NX + r], h_ = hi[k * NZ + r], l_ = lo[k * NZ + r], dz = dXa[k * NX + r];
                    if (h_ < kBig) {
                        const double s = h_ - z, l = lamU_x(k, r);
                        const double dl = -l + (l / s) * dz;      // affine dual step
                        g += (sigma_mu - (-dz) * dl) / s;
                    }
                    if (l_ > -kBig) {
                        const double s = z - l_, l = lamL_x(k, r);
                        const double dl = -l - (l / s) * dz;
                        g -= (sigma_mu - dz * dl) / s;
                    }
                }
                return g;
            };
            auto rhs_u = [&](int k, int p) -> double {
                double g = gu[k * NU + p];
                if (pass == 1) {
                    const double z = U[k * NU + p], h_ = hi[k * NZ + NX + p], l_ = lo[k * NZ + NX + p], dz = dUa[k * NU + p];
                    if (h_ < kBig) {
                        const double s = h_ - z, l = lamU_u(k, p);
                        const double dl = -l + (l / s) * dz;
                        g += (sigma_mu - (-dz) * dl) / s;
                    }
                    if (l_ > -kBig) {
                        const double s = z - l_, l = lamL_u(k, p);
                        const double dl = -l - (l / s) * dz;
                        g -= (sigma_mu - dz * dl) / s;
                    }
                }
                return g;
            };
            // ---- backward: p_N = rx_N
            sm.vec[r0] = rhs_x(N, r0);
            if (v1) sm.vec[r1] = rhs_x(N, r1);
            if (pass == 0) {
                // S_N = 2 Qf + Sigma_N
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = rr ? r1 : r0;
                    if (rr && !v1) continue;
                    const double z = X[N * NX + r], h_ = hi[N * NZ + r], l_ = lo[N * NZ + r];
                    double sg = 0.0;
                    if (h_ < kBig) sg += lamU_x(N, r) / (h_ - z);
                    if (l_ > -kBig) sg += lamL_x(N, r) / (z - l_);
                    for (int c = 0; c < LD; ++c) sm.S[r * LD + c] = (c < NX) ? (2.0 * prm->Qf[r * NX + c] + (c == r ? sg : 0.0)) : 0.0;
                }
            }
            __syncwarp(mask);
            for (int k = N - 1; k >= 0; --k) {
                const double* Ak = Ab + (size_t)k * NX * NX;
                const double* Bk = Bb + (size_t)k * NX * NU;
                // vector recursion shared by both passes; A columns / K columns / factor in registers
                auto vec_step = [&](const double (&a0)[NX], const double (&a1)[NX], const double (&k0)[NU], const double (&k1)[NU],
                                    const double (&L)[NU][NU]) {
                    // Sc_k = S_{k+1} rp_k (pass 0) ; g = p + Sc
                    double sc0, sc1;
                    if (pass == 0) {
                        sm.vec[LD + r0] = rp[k * NX + r0];
                        if (v1) sm.vec[LD + r1] = rp[k * NX + r1];
                        __syncwarp(mask);
                        sc0 = 0.0; sc1 = 0.0;
#pragma unroll
                        for (int c = 0; c < NX; ++c) {
                            const double cv = sm.vec[LD + c];
                            sc0 = fma(sm.S[r0 * LD + c], cv, sc0);
                            sc1 = fma(sm.S[r1c * LD + c], cv, sc1);
                        }
                        Sc[k * NX + r0] = sc0;
                        if (v1) Sc[k * NX + r1] = sc1;
                        __syncwarp(mask);
                    } else {
                        sc0 = Sc[k * NX + r0];
                        sc1 = Sc[k * NX + r1c];
                    }
                    const double g0 = sm.vec[r0] + rp_scale * sc0;
                    const double g1 = v1 ? (sm.vec[r1] + rp_scale * sc1) : 0.0;
                    sm.vec[LD + r0] = g0;
                    if (v1) sm.vec[LD + r1] = g1;
                    __syncwarp(mask);
                    // t = ru + B' g
                    for (int p = sl; p < NU; p += LPS) {
                        double t = rhs_u(k, p);
#pragma unroll
                        for (int i = 0; i < NX; ++i) t = fma(sm.Bm[i * NU + p], sm.vec[LD + i], t);
                        sm.vec[2 * LD + p] = t;
                    }
                    __syncwarp(mask);
                    double tt[NU], dd[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) { tt[p] = sm.vec[2 * LD + p]; dd[p] = tt[p]; }
                    chol_solve<NU>(L, dd);
                    for (int p = sl; p < NU; p += LPS) {
                        // static-index select of dd[p]
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q < NU; ++q) if (q == p) v = dd[q];
                        dff[k * NU + p] = -v;
                    }
                    if (k > 0) {
                        // p_k = rx_k + A' g + K' t
                        double p0 = rhs_x(k, r0), p1 = v1 ? rhs_x(k, r1) : 0.0;
#pragma unroll
                        for (int i = 0; i < NX; ++i) {
                            const double gi = sm.vec[LD + i];
                            p0 = fma(a0[i], gi, p0);
                            p1 = fma(a1[i], gi, p1);
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
                    // Sigma_k (k >= 1) and Sigma_u
                    double cx0 = 0.0, cx1 = 0.0;
                    if (k > 0) {
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int r = rr ? r1 : r0;
                            if (rr && !v1) continue;
                            const double z = X[k * NX + r], h_ = hi[k * NZ + r], l_ = lo[k * NZ + r];
                            double sg = 0.0;
                            if (h_ < kBig) sg += lamU_x(k, r) / (h_ - z);
                            if (l_ > -kBig) sg += lamL_x(k, r) / (z - l_);
                            if (rr) cx1 = sg; else cx0 = sg;
                        }
                    }
                    for (int p = sl; p < NU; p += LPS) {
                        const double z = U[k * NU + p], h_ = hi[k * NZ + NX + p], l_ = lo[k * NZ + NX + p];
                        double sg = 0.0;
                        if (h_ < kBig) sg += lamU_u(k, p) / (h_ - z);
                        if (l_ > -kBig) sg += lamL_u(k, p) / (z - l_);
                        sm.cu[p] = sg;
                    }
                    const bool ok = riccati_step<NX, NU>(
                        sm, sl, mask, Ak, Bk, prm->Q2, prm->R2, cx0, cx1, Kq + (size_t)k * NU * NX,
                        [&](const double (&a0)[NX], const double (&a1)[NX], const double (&k0)[NU], const double (&k1)[NU],
                            const double (&L)[NU][NU]) {
                            if (sl == 0) {
#pragma unroll
                                for (int p = 0; p < NU; ++p)
#pragma unroll
                                    for (int q = 0; q < NU; ++q) Lq[k * NU * NU + p * NU + q] = L[p][q];
                            }
                            vec_step(a0, a1, k0, k1, L);
                        },
                        k > 0);
                    if (!ok) status = 2;
                } else {
                    // corrector: reload A columns, K columns, B and the factor
                    double a0[NX], a1[NX], k0[NU], k1[NU], L[NU][NU];
#pragma unroll
                    for (int i = 0; i < NX; ++i) {
                        a0[i] = __ldg(Ak + i * NX + r0);
                        a1[i] = v1 ? __ldg(Ak + i * NX + r1c) : 0.0;
                    }
                    for (int e = sl; e < NX * NU; e += LPS) sm.Bm[e] = __ldg(Bk + e);
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        k0[p] = Kq[(size_t)k * NU * NX + p * NX + r0];
                        k1[p] = v1 ? Kq[(size_t)k * NU * NX + p * NX + r1c] : 0.0;
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = Lq[k * NU * NU + p * NU + q];
                    }
                    __syncwarp(mask);
                    vec_step(a0, a1, k0, k1, L);
                }
            }
            // ---- forward: dx_0 = 0 ; du_k = K dx_k + d_k ; dx_{k+1} = A dx_k + B du_k + rp_k
            double* dXo = pass == 0 ? dXa : dXc;
            double* dUo = pass == 0 ? dUa : dUc;
            __syncwarp(mask);
            sm.vec[r0] = 0.0;
            if (v1) sm.vec[r1] = 0.0;
            dXo[r0] = 0.0;
            if (v1) dXo[r1] = 0.0;
            __syncwarp(mask);
            for (int k = 0; k < N; ++k) {
                const double* Ak = Ab + (size_t)k * NX * NX;
                const double* Bk = Bb + (size_t)k * NX * NU;
                const double* Kk = Kq + (size_t)k * NU * NX;
                for (int p = sl; p < NU; p += LPS) {
                    double u = dff[k * NU + p];
#pragma unroll
                    for (int m = 0; m < NX; ++m) u = fma(Kk[p * NX + m], sm.vec[m], u);
                    sm.vec[LD + p] = u;
                    dUo[k * NU + p] = u;
                }
                __syncwarp(mask);
                double n0 = rp_scale * rp[k * NX + r0], n1 = rp_scale * rp[k * NX + r1c];
#pragma unroll
                for (int m = 0; m < NX; ++m) {
                    const double xm = sm.vec[m];
                    n0 = fma(__ldg(Ak + r0 * NX + m), xm, n0);
                    n1 = fma(__ldg(Ak + r1c * NX + m), xm, n1);
                }
#pragma unroll
                for (int p = 0; p < NU; ++p) {
                    const double up = sm.vec[LD + p];
                    n0 = fma(__ldg(Bk + r0 * NU + p), up, n0);
                    n1 = fma(__ldg(Bk + r1c * NU + p), up, n1);
                }
                __syncwarp(mask);
                sm.vec[r0] = n0;
                if (v1) sm.vec[r1] = n1;
                dXo[(k + 1) * NX + r0] = n0;
                if (v1) dXo[(k + 1) * NX + r1] = n1;
                __syncwarp(mask);
            }
            // ---- step lengths (and mu_aff after the predictor)
            const double tau = pass == 0 ? 1.0 : (mu_c > 1e-8 ? 0.995 : 0.9999);
            double apl = 1.0, adl = 1.0;
            auto ratio = [&](double v, double dv, double& amin) {
                if (dv < 0.0) amin = fmin(amin, -tau * v / dv);
            };
            auto entry_step = [&](double z, double h_, double l_, double lu, double ll, double dz, double dza, double& dlu, double& dll) {
                dlu = 0.0; dll = 0.0;
                if (h_ < kBig) {
                    const double s = h_ - z;
                    double tgt = 0.0;
                    if (pass == 1) { const double dla = -lu + (lu / s) * dza; tgt = sigma_mu - (-dza) * dla; }
                    dlu = tgt / s - lu + (lu / s) * dz;
                    ratio(s, -dz, apl);
                    ratio(lu, dlu, adl);
                }
                if (l_ > -kBig) {
                    const double s = z - l_;
                    double tgt = 0.0;
                    if (pass == 1) { const double dla = -ll - (ll / s) * dza; tgt = sigma_mu - dza * dla; }
                    dll = tgt / s - ll - (ll / s) * dz;
                    ratio(s, dz, apl);
                    ratio(ll, dll, adl);
                }
            };
            for (int k = 1; k <= N; ++k) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = rr ? r1 : r0;
                    if (rr && !v1) continue;
                    double dlu, dll;
                    entry_step(X[k * NX + r], hi[k * NZ + r], lo[k * NZ + r], lamU_x(k, r), lamL_x(k, r), dXo[k * NX + r],
                               dXa[k * NX + r], dlu, dll);
                }
            }
            for (int k = 0; k < N; ++k)
                for (int p = sl; p < NU; p += LPS) {
                    double dlu, dll;
                    entry_step(U[k * NU + p], hi[k * NZ + NX + p], lo[k * NZ + NX + p], lamU_u(k, p), lamL_u(k, p), dUo[k * NU + p],
                               dUa[k * NU + p], dlu, dll);
                }
            slot_min2(apl, adl);
            ap = apl; ad = adl;
            if (pass == 0) {
                // mu_aff
                double acc = 0.0, cnt = 0.0;
                auto entry_mu = [&](double z, double h_, double l_, double lu, double ll, double dz) {
                    if (h_ < kBig) {
                        const double s = h_ - z;
                        const double dl = -lu + (lu / s) * dz;
                        acc = fma(s - ap * dz, lu + ad * dl, acc); cnt += 1.0;
                    }
                    if (l_ > -kBig) {
                        const double s = z - l_;
                        const double dl = -ll - (ll / s) * dz;
                        acc = fma(s + ap * dz, ll + ad * dl, acc); cnt += 1.0;
                    }
                };
                for (int k = 1; k <= N; ++k) {
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int r = rr ? r1 : r0;
                        if (rr && !v1) continue;
                        entry_mu(X[k * NX + r], hi[k * NZ + r], lo[k * NZ + r], lamU_x(k, r), lamL_x(k, r), dXa[k * NX + r]);
                    }
                }
                for (int k = 0; k < N; ++k)
                    for (int p = sl; p < NU; p += LPS)
                        entry_mu(U[k * NU + p], hi[k * NZ + NX + p], lo[k * NZ + NX + p], lamU_u(k, p), lamL_u(k, p), dUa[k * NU + p]);
                slot_sum2(acc, cnt);
                const double mu_aff = cnt > 0.0 ? acc / cnt : 0.0;
                const double ratio_mu = mu_c > 0.0 ? mu_aff / mu_c : 0.0;
                sigma_mu = ratio_mu * ratio_mu * ratio_mu * mu_c;
            }
        }
        // ---- update (only slots that are still iterating)
        if (!done) {
            auto entry_upd = [&](double& z, double h_, double l_, double& lu, double& ll, double dz, double dza) {
                double dlu = 0.0, dll = 0.0;
                if (h_ < kBig) {
                    const double s = h_ - z;
                    const double dla = -lu + (lu / s) * dza;
                    dlu = (sigma_mu - (-dza) * dla) / s - lu + (lu / s) * dz;
                }
                if (l_ > -kBig) {
                    const double s = z - l_;
                    const double dla = -ll - (ll / s) * dza;
                    dll = (sigma_mu - dza * dla) / s - ll - (ll / s) * dz;
                }
                z += ap * dz;
                lu += ad * dlu;
                ll += ad * dll;
            };
            for (int k = 1; k <= N; ++k) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = rr ? r1 : r0;
                    if (rr && !v1) continue;
                    entry_upd(X[k * NX + r], hi[k * NZ + r], lo[k * NZ + r], lamU_x(k, r), lamL_x(k, r), dXc[k * NX + r], dXa[k * NX + r]);
                }
            }
            for (int k = 0; k < N; ++k)
                for (int p = sl; p < NU; p += LPS)
                    entry_upd(U[k * NU + p], hi[k * NZ + NX + p], lo[k * NZ + NX + p], lamU_u(k, p), lamL_u(k, p), dUc[k * NU + p],
                              dUa[k * NU + p]);
            rp_scale *= (1.0 - ap);
        }
        __syncwarp(mask);
        const double mu_new = compl_mu();
        if (!done) {
            mu_c = mu_new;
            if (!(mu_c == mu_c) || status == 2) { status = 2; done = true; }
            else if (mu_c < a.tol_mu && rp_inf * rp_scale < a.tol_res && it >= 2) done = true;
        }
    }
    if (!done && status == 0) status = 1;

    // ---- objective value  x'Hx + q'x  (qp_jit.py:401 obj_val with P = 2H, q = 2 H y_nom)
    {
        double acc = 0.0, dmy = 0.0;
        for (int k = 0; k <= N; ++k) {
            const double* Qm = (k < N) ? prm->Q : prm->Qf;
            double q0 = 0.0, q1 = 0.0;
            for (int m = 0; m < NX; ++m) {
                const double y = X[k * NX + m] + 2.0 * Xn[k * NX + m];
                q0 = fma(Qm[r0 * NX + m], y, q0);
                q1 = fma(Qm[r1c * NX + m], y, q1);
            }
            acc = fma(X[k * NX + r0], q0, acc);
            if (v1) acc = fma(X[k * NX + r1], q1, acc);
        }
        for (int k = 0; k < N; ++k)
            for (int p = sl; p < NU; p += LPS) {
                double q = 0.0;
                for (int m = 0; m < NU; ++m) q = fma(prm->R[p * NU + m], U[k * NU + m] + 2.0 * Un[k * NU + m], q);
                acc = fma(U[k * NU + p], q, acc);
            }
        slot_sum2(acc, dmy);
        if (sl == 0) {
            a.cost[b] = acc;
            a.iters[b] = it;
            a.status[b] = status;
        }
    }
}

}  // namespace rnmpc
