// Nominal QP, latency variant: ONE CTA (4 warps) PER INSTANCE, the whole working set of the interior-point iteration resident in
// shared memory.  Same problem, same algorithm and same formulas as qp_ipm_kernel (rnmpc_qp.cuh; reference: solver/qp_jit.py) --
// Mehrotra predictor-corrector, every KKT solve a Riccati factorisation / sweep, centred cold start, same termination tests and
// status codes -- but parallel over the ENTRIES of the small matrices instead of over instances:
//   * the factorisation sweep runs on all 128 threads, one output entry of each matrix product per thread and round
//     (4 block barriers per stage; the Cholesky factorisation of the 4 x 4 block runs beside the largest product);
//   * the vector sweeps (predictor / corrector forward, corrector backward) run on warp 0 with one lane per vector entry;
//   * A_k, B_k, K_k, the Cholesky factors and the per-entry iterate (slacks, multipliers, steps) never leave shared memory.
// qp_ipm_kernel gives every instance 9 lanes and streams its ~100 KB working set from L2/HBM four times per iteration: right for
// thousands of instances, 3 x slower than one CPU core for a single one.  This kernel is for small batches.
// Restrictions (the engine falls back to qp_ipm_kernel otherwise): diagonal cost matrices, cold start, reference bounds
// (no explicit hi/lo/q arrays).  Results agree with qp_ipm_kernel to rounding, not bit for bit (different summation orders).
#pragma once
#include "rnmpc_common.cuh"
#include "rnmpc_qp.cuh"

namespace rnmpc {

constexpr int kBlockThreads = 128;

// shared-memory footprint in doubles (every array rounded up to an even count: 16-byte alignment)
template <int NX, int NU>
__host__ __device__ inline size_t qp_block_smem_doubles(int N) {
    constexpr int NZ = NX + NU;
    auto ev = [](size_t n) { return (n + 1) & ~(size_t)1; };
    const size_t NE = (size_t)(N + 1) * NZ;
    return ev((size_t)N * NX * NX) + ev((size_t)N * NX * NU) + ev((size_t)N * NU * NX) + ev((size_t)N * NU * NU)   // A, B, K, L
         + 2 * ev((size_t)N * NX) + ev((size_t)N * NU)                                                              // rp, Sc, dff
         + 12 * ev(NE)                                                                                               // entry arrays
         + 4 * ev(NX * NX) + 2 * ev(NX * NU) + ev(NU * NU)                                                           // S, Y, T, Sn, SB, Ft, H
         + 4 * ev(NX) + 3 * ev(NU) + ev(NX + NU) + 64;                                                               // p, pn, g, xc, tt, dd, uc, rhs, red
}

// One output entry of a small matrix product as a strided dot product: *out = add + sum_m a[m sa] b[m sb].
struct Dot { const double* a; const double* b; int sa, sb; double* out; double add; };

// The dot tasks e = t, t + TW, ... < count of thread t run side by side (R tasks, two accumulators each, loop over m unrolled): the
// loads of all of them are in flight together instead of one dependent load-FMA pair at a time.
template <int LEN, int R, int TW, class Decode>
__device__ __forceinline__ void run_dots_t(const int t, const int count, Decode&& dec) {
    Dot d[R];
    bool on[R];
    double acc[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) { const int e = t + r * TW; on[r] = e < count; d[r] = dec(on[r] ? e : 0); acc[r][0] = 0.0; acc[r][1] = 0.0; }
#pragma unroll
    for (int m = 0; m < LEN; ++m)
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][m & 1] = fma(d[r].a[m * d[r].sa], d[r].b[m * d[r].sb], acc[r][m & 1]);
#pragma unroll
    for (int r = 0; r < R; ++r) if (on[r]) *d[r].out = d[r].add + (acc[r][0] + acc[r][1]);
}
template <int LEN, int R, class Decode>
__device__ __forceinline__ void run_dots(const int t, const int count, Decode&& dec) { run_dots_t<LEN, R, kBlockThreads>(t, count, dec); }

template <int NX, int NU>
__global__ void __launch_bounds__(kBlockThreads, 2) qp_block_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    constexpr int NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX, LD = Slot<NX>::LD, T = kBlockThreads;
    static_assert(NX + NU <= 32 && NX <= 31, "one lane of warp 0 per vector entry, lane 31 free");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = a.N;
    const bool resume = a.resume != 0;
    int b = blockIdx.x;
    if (resume) {                   // finish the instances the instance-per-slot kernel handed off (QpArgs::handoff_*)
        if (b >= *a.handoff_count) return;
        b = a.handoff_list[b];
    } else {
        if (b >= a.B) return;
        if (a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) return;
    }

    double* sp = reinterpret_cast<double*>(smem_raw);
    auto take = [&](size_t n) { double* p = sp; sp += (n + 1) & ~(size_t)1; return p; };
    const int NE = (N + 1) * NZ;
    double* A = take((size_t)N * NX * NX); double* Bm = take((size_t)N * NX * NU);
    double* Kq = take((size_t)N * NU * NX); double* Lq = take((size_t)N * NU * NU);
    double* rp = take((size_t)N * NX); double* Sc = take((size_t)N * NX); double* dff = take((size_t)N * NU);
    double* e_su = take(NE); double* e_sl = take(NE); double* e_lu = take(NE); double* e_ll = take(NE);
    double* e_iu = take(NE); double* e_il = take(NE); double* e_z = take(NE); double* e_zn = take(NE);
    double* e_da = take(NE); double* e_dc = take(NE); double* e_sg = take(NE); double* e_qd = take(NE);
    double* S = take(NX * NX); double* Y = take(NX * NX); double* Tm = take(NX * NX); double* Sn = take(NX * NX);
    double* SB = take(NX * NU); double* Ft = take(NX * NU); double* H = take(NU * NU);
    double* pv = take(NX); double* pn = take(NX); double* gv = take(NX); double* xc = take(NX);
    double* tt = take(NU); double* dd = take(NU); double* uc = take(NU);
    double* rhs = take(NX + NU);       // right-hand sides of the current stage: rx (NX), ru (NU)
    double* red = take(64);

    const double* Ag = a.A + (size_t)b * N * NX * LD;
    const double* Bg = a.Bm + (size_t)b * N * NX * NU;
    const double* cg = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;

    // block-wide reductions (every thread receives the result; two barriers each)
    auto block_sum = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double r = 0.0;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) r += red[w];
        return r;
    };
    auto block_max = [&](double v) -> double {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if (lane == 0) red[warp] = v;
        __syncthreads();
        double r = red[0];
#pragma unroll
        for (int w = 1; w < T / 32; ++w) r = fmax(r, red[w]);
        return r;
    };

    if (tid == 0) red[32] = 0.0;       // "a pivot of the factorisation was not positive" flag of the current iteration
    // ------------------------------------------------------------------ stage matrices -> shared memory
    for (int e = tid; e < N * NX * NX; e += T) { const int k = e / (NX * NX), r = e - k * NX * NX; A[e] = __ldg(Ag + (size_t)k * NX * LD + (r / NX) * LD + (r % NX)); }
    for (int e = tid; e < N * NX * NU; e += T) Bm[e] = __ldg(Bg + e);

    // ------------------------------------------------------------------ bounds, initial point, slacks (rnmpc_qp.cuh, cold start)
    const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);
    double bad = 0.0, n_box = 0.0, s_sum = 0.0;
    if (!resume) {
        const double eps = tight ? 0.0 : kEpsQP;
        for (int idx = tid; idx < NE; idx += T) {
            const int k = idx / NZ, i = idx - k * NZ;
            double h_ = 1e300, l_ = -1e300, z = 0.0, zn = 0.0, qd = 0.0;
            bool box = false, present = true;
            if (i < NX) {
                const double xn = Xn[k * NX + i];
                zn = xn;
                qd = 2.0 * ((k < N) ? prm->Q[i * NX + i] : prm->Qf[i * NX + i]);
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
                qd = 2.0 * prm->R[p * NU + p];
                const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                h_ = (prm->uub[p] - un) - bu + eps;
                l_ = -((un - prm->ulb[p]) - bl + eps);
                box = true;
            } else present = false;                                                    // no input at the terminal stage
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
            (void)present;
            e_su[idx] = su; e_sl[idx] = sl;
            e_iu[idx] = (su < kBig) ? 1.0 / su : 0.0; e_il[idx] = (sl < kBig) ? 1.0 / sl : 0.0;
            e_z[idx] = z; e_zn[idx] = zn; e_da[idx] = 0.0; e_dc[idx] = 0.0; e_qd[idx] = qd;
        }
    }
    bad = block_max(bad);
    n_box = block_sum(n_box);
    s_sum = block_sum(s_sum);
    int status = bad > 0.0 ? 2 : 0;
    if (!resume) {   // centred cold start: multipliers mu_bar / slack (QpArgs::mu0)
        const double mu_bar = a.mu0 < 0.0 ? (n_box > 0.0 ? -a.mu0 * s_sum / n_box : 1.0) : a.mu0;
        for (int idx = tid; idx < NE; idx += T) {
            double lu, ll;
            if (a.mu0 != 0.0) { lu = mu_bar * e_iu[idx]; ll = mu_bar * e_il[idx]; }
            else { lu = e_iu[idx] > 0.0 ? 1.0 : 0.0; ll = e_il[idx] > 0.0 ? 1.0 : 0.0; }
            e_lu[idx] = lu; e_ll[idx] = ll;
            e_sg[idx] = lu * e_iu[idx] + ll * e_il[idx];
        }
    }
    __syncthreads();
    // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
    double rp_inf = 0.0;
    if (!resume)
    for (int e = tid; e < N * NX; e += T) {
        const int k = e / NX, r = e - k * NX;
        double acc = cg[e] - e_z[(k + 1) * NZ + r];
        for (int m = 0; m < NX; ++m) acc = fma(A[(k * NX + r) * NX + m], e_z[k * NZ + m], acc);
        for (int p = 0; p < NU; ++p) acc = fma(Bm[(k * NX + r) * NU + p], e_z[k * NZ + NX + p], acc);
        rp[e] = acc;
        rp_inf = fmax(rp_inf, fabs(acc));
    }
    rp_inf = block_max(rp_inf);
    double mu_c;
    if (!resume) {
        double acc = 0.0;
        for (int idx = tid; idx < NE; idx += T) {
            if (e_su[idx] < kBig) acc = fma(e_su[idx], e_lu[idx], acc);
            if (e_sl[idx] < kBig) acc = fma(e_sl[idx], e_ll[idx], acc);
        }
        acc = block_sum(acc);
        mu_c = n_box > 0.0 ? acc / n_box : 0.0;
    }
    double rp_scale = 1.0;
    int it0 = 0;                    // iterations already done by qp_ipm_kernel
    if (resume) {
        // iterate of the handed-off instance: entry records and dynamics residual from qp_ipm_kernel's workspace
        const double* E = a.E + (size_t)b * NE * kRec;
        for (int idx = tid; idx < NE; idx += T) {
            const double* r = E + (size_t)idx * kRec;
            e_su[idx] = r[F_SU]; e_sl[idx] = r[F_SL]; e_lu[idx] = r[F_LU]; e_ll[idx] = r[F_LL]; e_iu[idx] = r[F_ISU]; e_il[idx] = r[F_ISL];
            e_z[idx] = r[F_Z]; e_zn[idx] = r[F_ZN]; e_da[idx] = 0.0; e_dc[idx] = 0.0; e_sg[idx] = r[F_SIG]; e_qd[idx] = r[F_QD];
        }
        for (int e = tid; e < N * NX; e += T) rp[e] = a.rp[(size_t)b * N * NX + e];
        const double* hs = a.handoff_state + (size_t)b * 4;
        rp_scale = hs[0]; mu_c = hs[1]; rp_inf = hs[2]; n_box = hs[3];
        it0 = a.iters[b];
        status = 0;
        __syncthreads();
    }

    // right-hand side of one entry (rnmpc_qp.cuh rhs_eval): gradient (+ centring / second-order terms in the corrector pass)
    double sigma_mu = 0.0;
    auto rhs_entry = [&](int pass, int idx) -> double {
        double g = e_qd[idx] * (e_z[idx] + e_zn[idx]);
        if (pass == 1) {
            const double iu = e_iu[idx], il = e_il[idx], lu = e_lu[idx], ll = e_ll[idx], dza = e_da[idx];
            if (iu > 0.0) g = fma(sigma_mu + dza * (lu * iu * dza - lu), iu, g);
            if (il > 0.0) g = fma(-(sigma_mu + dza * (ll * il * dza + ll)), il, g);
        }
        return g;
    };

    int stalled = 0, it = 0;
    bool done = status != 0;
    for (it = it0; it < a.max_iter && !done; ++it) {
        bool chol_ok = true;
        double ap = 1.0, ad = 1.0, maxp0 = 0.0, maxd0 = 0.0;
        bool skip_corr = false;
        __syncthreads();
        // ================================================================ factorisation sweep + predictor right-hand side (all threads)
        for (int e = tid; e < NX * NX; e += T) { const int i = e / NX, j = e - i * NX; S[e] = (i == j) ? (e_qd[N * NZ + i] + e_sg[N * NZ + i]) : 0.0; }
        if (tid < NX) pv[tid] = rhs_entry(0, N * NZ + tid);
        __syncthreads();
        for (int k = N - 1; k >= 0; --k) {
            const double* Ak = A + (size_t)k * NX * NX;
            const double* Bk = Bm + (size_t)k * NX * NU;
            double* Kk = Kq + (size_t)k * NU * NX;
            // phase 1: Y = A'S, SB = S B, Sc_k = S rp_k (dot tasks), stage right-hand sides
            run_dots<NX, (NX * NX + NX * NU + NX + T - 1) / T>(tid, NX * NX + NX * NU + NX, [&](int e) -> Dot {
                if (e < NX * NX) { const int i = e / NX, j = e - i * NX; return Dot{Ak + i, S + j, NX, NX, Y + e, 0.0}; }
                if (e < NX * NX + NX * NU) { const int f = e - NX * NX, i = f / NU, p = f - i * NU; return Dot{S + i * NX, Bk + p, 1, NU, SB + f, 0.0}; }
                const int i = e - NX * NX - NX * NU;
                return Dot{S + i * NX, rp + k * NX, 1, 1, Sc + k * NX + i, 0.0};
            });
            if (tid >= T - NZ) { const int i = tid - (T - NZ); rhs[i] = (i < NX && k == 0) ? 0.0 : rhs_entry(0, k * NZ + i); }
            __syncthreads();
            // phase 2: H = R2 + diag(sig_u) + B'SB, Ft = Y B (= (B'SA)'), B'p and B'Sc (for t = ru + B'(p + rp_scale Sc)), g = p + rp_scale Sc
            run_dots<NX, (NU * NU + NX * NU + 2 * NU + T - 1) / T>(tid, NU * NU + NX * NU + 2 * NU, [&](int e) -> Dot {
                if (e < NU * NU) { const int p = e / NU, q = e - p * NU; return Dot{Bk + p, SB + q, NU, NU, H + e, prm->R2[e] + (p == q ? e_sg[k * NZ + NX + p] : 0.0)}; }
                if (e < NU * NU + NX * NU) { const int f = e - NU * NU, j = f / NU, p = f - j * NU; return Dot{Y + j * NX, Bk + p, 1, NU, Ft + f, 0.0}; }
                const int f = e - NU * NU - NX * NU;
                if (f < NU) return Dot{Bk + f, pv, NU, 1, tt + f, 0.0};
                return Dot{Bk + (f - NU), Sc + k * NX, NU, 1, dd + (f - NU), 0.0};
            });
            if (tid >= T - NX) { const int i = tid - (T - NX); gv[i] = pv[i] + rp_scale * Sc[k * NX + i]; }
            __syncthreads();
            // phase 3: threads 0 .. NX+1 factor H (each its own copy): K = -H^{-1} F column by column, t and d = -H^{-1} t, stored factor;
            // the other threads meanwhile form the part of S' that does not need K:  Sn = Y A
            if (tid < NX + 2) {
                double L[NU][NU];
#pragma unroll
                for (int p = 0; p < NU; ++p)
#pragma unroll
                    for (int q = 0; q < NU; ++q) L[p][q] = H[p * NU + q];
                const bool ok = chol_inplace<NU>(L);
                if (tid < NX) {
                    double f[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) f[p] = Ft[tid * NU + p];
                    chol_solve<NU>(L, f);
#pragma unroll
                    for (int p = 0; p < NU; ++p) Kk[p * NX + tid] = -f[p];
                } else if (tid == NX) {
                    double d[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) { d[p] = rhs[NX + p] + tt[p] + rp_scale * dd[p]; uc[p] = d[p]; }      // t (kept in uc for p')
                    chol_solve<NU>(L, d);
#pragma unroll
                    for (int p = 0; p < NU; ++p) dff[k * NU + p] = -d[p];
                } else {
#pragma unroll
                    for (int p = 0; p < NU; ++p)
#pragma unroll
                        for (int q = 0; q < NU; ++q) Lq[k * NU * NU + p * NU + q] = L[p][q];
                    if (!ok) red[32] = 1.0;
                }
            } else if (k > 0) {
                constexpr int TW = T - (NX + 2);
                run_dots_t<NX, (NX * NX + TW - 1) / TW, TW>(tid - (NX + 2), NX * NX, [&](int e) -> Dot {
                    const int i = e / NX, j = e - i * NX;
                    return Dot{Y + i * NX, Ak + j, 1, NX, Sn + e, 0.0};
                });
            }
            __syncthreads();
            if (k > 0) {
                // phase 4: S = sym(Sn + Ft K) + diagonal stage cost (2Q + barrier weights), p = rx + A'g + K't
                for (int e = tid; e < NX * NX; e += T) {
                    const int i = e / NX, j = e - i * NX;
                    double acc = Sn[e] + Sn[j * NX + i];
#pragma unroll
                    for (int p = 0; p < NU; ++p) acc = fma(Ft[i * NU + p], Kk[p * NX + j], fma(Ft[j * NU + p], Kk[p * NX + i], acc));
                    S[e] = 0.5 * acc + (i == j ? (e_qd[k * NZ + i] + e_sg[k * NZ + i]) : 0.0);
                }
                if (tid >= T - NX) {
                    const int i = tid - (T - NX);
                    double acc = rhs[i], acc2 = 0.0;
#pragma unroll
                    for (int m = 0; m + 1 < NX; m += 2) { acc = fma(Ak[m * NX + i], gv[m], acc); acc2 = fma(Ak[(m + 1) * NX + i], gv[m + 1], acc2); }
                    if (NX & 1) acc = fma(Ak[(NX - 1) * NX + i], gv[NX - 1], acc);
#pragma unroll
                    for (int p = 0; p < NU; ++p) acc = fma(Kk[p * NX + i], uc[p], acc);
                    pv[i] = acc + acc2;
                }
                __syncthreads();
            }
        }
        chol_ok = red[32] == 0.0;

        // ================================================================ vector sweeps on warp 0 (one lane per vector entry)
        // forward: dx_0 = 0 ; du_k = K dx_k + d_k ; dx_{k+1} = A dx_k + B du_k + rp_scale rp_k ; ratio test (+ sums for mu_affine)
        // NU-vector results (u = K x + d, t = ru + B'g) are split over LPO lanes per output and combined with shuffles
        constexpr int LPO = NU >= 3 ? 8 : (NU == 2 ? 16 : 32);
        auto forward = [&](const int pass) {
            const double tau = pass == 0 ? 1.0 : (mu_c > 1e-8 ? 0.995 : 0.9999);
            double rpm = 0.0, rdm = 0.0, S3 = 0.0, S4 = 0.0;
            double* e_dz = pass == 0 ? e_da : e_dc;
            const int up = lane / LPO, uj = lane - up * LPO;          // this lane's share of u[up]: terms m = uj, uj + LPO, ...
            if (lane < NX) { xc[lane] = 0.0; e_dz[lane] = 0.0; }
            __syncwarp();
            for (int k = 0; k < N; ++k) {
                const double* Ak = A + (size_t)k * NX * NX;
                const double* Bk = Bm + (size_t)k * NX * NU;
                const double* Kk = Kq + (size_t)k * NU * NX;
                // entry data of this stage's entries (x_{k+1}[lane] / u_k[lane - NX]) -- in flight during the products
                const int idx = lane < NX ? (k + 1) * NZ + lane : k * NZ + lane;
                double iu = 0.0, il = 0.0, lu = 0.0, ll = 0.0, dza = 0.0;
                if (lane < NZ) { iu = e_iu[idx]; il = e_il[idx]; lu = e_lu[idx]; ll = e_ll[idx]; dza = e_da[idx]; }
                double upart = 0.0;
                if (up < NU) {
#pragma unroll
                    for (int m = uj; m < NX; m += LPO) upart = fma(Kk[up * NX + m], xc[m], upart);
                }
                double n[4] = {0.0, 0.0, 0.0, 0.0};
                if (lane < NX) {
                    n[0] = rp_scale * rp[k * NX + lane];
#pragma unroll
                    for (int m = 0; m < NX; ++m) n[m & 3] = fma(Ak[lane * NX + m], xc[m], n[m & 3]);
                }
#pragma unroll
                for (int o = LPO / 2; o > 0; o >>= 1) upart += __shfl_xor_sync(0xffffffffu, upart, o);
                if (up < NU && uj == 0) uc[up] = upart + dff[k * NU + up];
                __syncwarp();
                double dz = 0.0;
                if (lane < NX) {
#pragma unroll
                    for (int p = 0; p < NU; ++p) n[p & 3] = fma(Bk[lane * NU + p], uc[p], n[p & 3]);
                    dz = (n[0] + n[1]) + (n[2] + n[3]);
                    xc[lane] = dz;
                } else if (lane < NZ) dz = uc[lane - NX];
                if (lane < NZ) {
                    e_dz[idx] = dz;
                    if (iu > 0.0) {
                        const double ru = lu * iu;
                        const double tgt = pass == 1 ? sigma_mu + dza * (ru * dza - lu) : 0.0;
                        const double dlu = tgt * iu - lu + ru * dz;
                        rpm = fmax(rpm, dz * iu);
                        if (dlu < 0.0) rdm = fmax(rdm, -dlu * fast_rcp(lu));
                        if (pass == 0) { S3 = fma(-dz, lu, S3); S4 = fma(-dz, dlu, S4); }
                    }
                    if (il > 0.0) {
                        const double rl = ll * il;
                        const double tgt = pass == 1 ? sigma_mu + dza * (rl * dza + ll) : 0.0;
                        const double dll = tgt * il - ll - rl * dz;
                        rpm = fmax(rpm, -dz * il);
                        if (dll < 0.0) rdm = fmax(rdm, -dll * fast_rcp(ll));
                        if (pass == 0) { S3 = fma(dz, ll, S3); S4 = fma(dz, dll, S4); }
                    }
                }
                __syncwarp();
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                rpm = fmax(rpm, __shfl_xor_sync(0xffffffffu, rpm, o)); rdm = fmax(rdm, __shfl_xor_sync(0xffffffffu, rdm, o));
                S3 += __shfl_xor_sync(0xffffffffu, S3, o); S4 += __shfl_xor_sync(0xffffffffu, S4, o);
            }
            if (lane == 0) { red[40] = rpm; red[41] = rdm; red[42] = S3; red[43] = S4; red[44] = tau; }
        };
        // corrector backward sweep with the stored factors: p_N = rhs_N ; g = p + rp_scale Sc ; t = ru + B'g ; d = -H^{-1} t ;
        // p_k = rx + A'g + K't.  p[i] only ever belongs to lane i: it lives in a register.
        auto backward_corrector = [&]() {
            const int up = lane / LPO, uj = lane - up * LPO;
            double pl = lane < NX ? rhs_entry(1, N * NZ + lane) : 0.0;
            for (int k = N - 1; k >= 0; --k) {
                const double* Ak = A + (size_t)k * NX * NX;
                const double* Bk = Bm + (size_t)k * NX * NU;
                const double* Kk = Kq + (size_t)k * NU * NX;
                double r = 0.0;
                if (lane < NX) {
                    gv[lane] = pl + rp_scale * Sc[k * NX + lane];
                    r = k > 0 ? rhs_entry(1, k * NZ + lane) : 0.0;
                }
                const double ru = up < NU && uj == 0 ? rhs_entry(1, k * NZ + NX + up) : 0.0;
                __syncwarp();
                double tpart = 0.0;
                if (up < NU) {
#pragma unroll
                    for (int i = uj; i < NX; i += LPO) tpart = fma(Bk[i * NU + up], gv[i], tpart);
                }
                double acc[4] = {r, 0.0, 0.0, 0.0};
                if (lane < NX && k > 0) {
#pragma unroll
                    for (int m = 0; m < NX; ++m) acc[m & 3] = fma(Ak[m * NX + lane], gv[m], acc[m & 3]);
                }
#pragma unroll
                for (int o = LPO / 2; o > 0; o >>= 1) tpart += __shfl_xor_sync(0xffffffffu, tpart, o);
                if (up < NU && uj == 0) tt[up] = tpart + ru;
                __syncwarp();
                if (lane == 31) {          // a lane without a p-entry (NX <= 31) solves for the feed-forward meanwhile
                    double L[NU][NU], d[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        d[p] = tt[p];
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = Lq[k * NU * NU + p * NU + q];
                    }
                    chol_solve<NU>(L, d);
#pragma unroll
                    for (int p = 0; p < NU; ++p) dff[k * NU + p] = -d[p];
                }
                if (lane < NX && k > 0) {
#pragma unroll
                    for (int p = 0; p < NU; ++p) acc[p & 3] = fma(Kk[p * NX + lane], tt[p], acc[p & 3]);
                    pl = (acc[0] + acc[1]) + (acc[2] + acc[3]);
                }
                __syncwarp();              // gv / tt are rewritten by the next stage
            }
        };

        __syncthreads();
        if (warp == 0) forward(0);
        __syncthreads();
        {
            const double rpm = red[40], rdm = red[41], S3 = red[42], S4 = red[43], tau = red[44];
            maxp0 = rpm; maxd0 = rdm;
            ap = (rpm > tau) ? tau / rpm : 1.0;
            ad = (rdm > tau) ? tau / rdm : 1.0;
            const double mu_aff = n_box > 0.0 ? (mu_c * n_box * (1.0 - ad) + (ap - ad) * S3 + ap * ad * S4) / n_box : 0.0;
            skip_corr = (ap >= 0.999 && ad >= 0.999 && mu_aff <= 1e-3 * mu_c);
            const double ratio_mu = mu_c > 0.0 ? mu_aff / mu_c : 0.0;
            sigma_mu = ratio_mu * ratio_mu * ratio_mu * mu_c;
        }
        if (!skip_corr) {
            __syncthreads();
            if (warp == 0) { backward_corrector(); forward(1); }
            __syncthreads();
            const double rpm = red[40], rdm = red[41], tau = red[44];
            ap = (rpm > tau) ? tau / rpm : 1.0;
            ad = (rdm > tau) ? tau / rdm : 1.0;
        } else {
            // Newton phase: the predictor step with the fraction-to-boundary rule (rnmpc_qp.cuh)
            const double tau_c = mu_c > 1e-8 ? 0.995 : 0.9999;
            ap = (maxp0 > tau_c) ? tau_c / maxp0 : 1.0;
            ad = (maxd0 > tau_c) ? tau_c / maxd0 : 1.0;
        }
        // ================================================================ update, new reciprocals / barrier weights / complementarity
        double acc = 0.0;
        {
            const bool upd = chol_ok;
            const double so = skip_corr ? 0.0 : 1.0;
            for (int idx = tid; idx < NE; idx += T) {
                double su = e_su[idx], sl = e_sl[idx], lu = e_lu[idx], ll = e_ll[idx];
                if (upd) {
                    const double dza = e_da[idx], dz = skip_corr ? dza : e_dc[idx];
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
                    e_sg[idx] = lu * iu + ll * il;
                }
                if (su < kBig) acc = fma(su, lu, acc);
                if (sl < kBig) acc = fma(sl, ll, acc);
            }
            if (upd) rp_scale *= (1.0 - ap);
        }
        acc = block_sum(acc);
        if (tid == 0) red[32] = 0.0;
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
        for (int idx = tid; idx < NE; idx += T) {
            const int k = idx / NZ, i = idx - k * NZ;
            if (i >= NX && k == N) continue;
            const double z = e_z[idx];
            acc = fma(z, 0.5 * e_qd[idx] * (z + 2.0 * e_zn[idx]), acc);
            if (i < NX) Xo[k * NX + i] = z; else Uo[k * NU + i - NX] = z;
            if (k < N) { mu[k * NI + i] = e_lu[idx]; mu[k * NI + NZ + i] = e_ll[idx]; }
            else { muf[i] = e_lu[idx]; muf[NX + i] = e_ll[idx]; }
        }
        acc = block_sum(acc);
        if (tid == 0) {
            a.cost[b] = acc;
            a.iters[b] = it;
            a.status[b] = status;
            if (a.ws_valid != nullptr) a.ws_valid[b] = 0;          // the records of qp_ipm_kernel's workspace were not written
        }
    }
}

}  // namespace rnmpc
