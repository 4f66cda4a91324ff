// Nominal QP, throughput variant for large batches: ONE WARP (one CTA of 32 threads) PER INSTANCE.  Same problem, algorithm, formulas
// and status codes as qp_mma_kernel (rnmpc_qp_mma.cuh; reference: solver/qp_jit.py:77-273, 362-402, 487-513): Mehrotra
// predictor-corrector interior point, every KKT solve a Riccati factorisation + vector sweeps, the tile products of a Riccati
// stage on the FP64 tensor pipe (mma.sync.m8n8k4.f64 -> DMMA).
//
// Why a second tensor-pipe kernel (measured, profiles/r2_notes.md):
//   * qp_mma_kernel gives an instance 4 warps and ~108 KB of shared memory (A_k, B_k, K_k and the iterate resident): 2 instances per
//     SM, four block barriers per Riccati stage, 5000 cycles per stage of which the DMMA pipe of a scheduler partition is busy for
//     ~450.  It is the latency kernel (one solve: 1.0 ms); at batch 4096 the tensor pipe is 9 % busy.
//   * sls_mma_kernel (one warp per chain, __syncwarp only, 11.5 KB per warp, 14 warps per SM) runs the same stage products 3 x
//     faster per stage at 48 % tensor-pipe utilisation: the batch supplies the parallelism, not the instance.
//   Here an instance therefore is one warp with ~36 KB of shared memory (6 instances per SM): the iterate (slacks, multipliers,
//   primal, both steps) and the per-stage vectors stay on chip; A_k, B_k come from global memory (at most 6 x 148 instances are in
//   flight: 38 MB of stage matrices, L2 resident) -- as operand fragments one stage ahead in the factorisation sweep, through a
//   three-slot ring of 1-D bulk copies (cp.async.bulk on the TMA unit, one lane issues a stage, completion on an mbarrier per slot;
//   the ring lives in the tile buffers, idle outside the factorisation) in the vector sweeps; K_k and the Cholesky factors
//   are written to a global (L2) workspace by the factorisation and come back through the same ring.  Reciprocal slacks are
//   recomputed (hardware seed + 2 Newton steps) instead of stored.  One CTA per instance lets the hardware scheduler replace a
//   finished instance at once (iteration counts spread from 8 to 25).
#pragma once
#include "rnmpc_common.cuh"
#include "rnmpc_qp.cuh"
#include "rnmpc_qp_mma.cuh"

namespace rnmpc {

template <int NX, int NU>
struct QpWarp {
    using Q = QpMma<NX, NU>;
    static constexpr int NZ = NX + NU;
    static constexpr int KP = Q::KP, KS = Q::KS, CB = Q::CB, CRP = Q::CRP, CP = Q::CP, NT = Q::NT, MTS = Q::MTS, MTM = Q::MTM;
    static constexpr int LDS = Q::LDS, LDW = Q::LDW, LDA = Q::LDA;
    static constexpr int mx(int a, int b) { return a > b ? a : b; }
    static constexpr int BUF1 = mx(8 * MTS * LDS, 8 * MTM * LDW);   // S (LDS layout) / M (LDW layout)
    static constexpr int BUF2 = 8 * MTS * LDW;                      // W
    static constexpr int MAT = BUF1 + BUF2;
    // one slot of the sweep ring: A_k [NX][LDA] | B_k [NX][NU] | K_k [NU][LDA]
    static constexpr int OFFB = NX * LDA, OFFK = OFFB + NX * NU, SLOT = OFFK + NU * LDA;
    static constexpr int NSLOT = 3;
    static_assert(NSLOT * SLOT <= MAT, "the sweep ring lives in the tile buffers");
    static_assert((NX * NU) % 2 == 0 && LDA % 2 == 0, "16-byte copies");
    static_assert(KP <= 8 * MTS, "B-operand rows of W are rows the first product writes");
    static_assert(NX + 2 <= 32 && NZ <= 32, "one lane per state / input entry");
};

// One softened bound row of the soft-constrained comparator (solver/nlp_soft_constraints.py:175-232): row slack s >= 0 (multiplier lam)
// of `+-z - bound <= gam`, slack variable gam >= 0 (multiplier nu), cost rho gam^2 + rho1 gam.  The slack variable only couples to its
// own row, so it is eliminated from the Newton system entry by entry: with ru = lam / s, rg = nu / gam, D = 2 rho + ru + rg
//   d gam = (ru (+-dz) + b) / D,  b = tgt / s + tg / gam - 2 rho gam - rho1      (tgt, tg: complementarity targets of the two pairs)
//   the row enters the Riccati system with weight ru (2 rho + rg) / D  (-> ru as rho -> inf: the hard constraint)
//   and right-hand side  tgt / s - ru b / D.
// The Riccati factorisation and the vector sweeps are those of the hard-constrained QP.
struct SoftSide {
    double s, lam, gam, nu, rho2, rho1;
    double is, ig, ru, rg, iD;
    __device__ __forceinline__ void prep() {
        is = fast_rcp(s); ig = fast_rcp(gam); ru = lam * is; rg = nu * ig; iD = fast_rcp(rho2 + ru + rg);
    }
    __device__ __forceinline__ double weight() const { return ru * (rho2 + rg) * iD; }
    __device__ __forceinline__ double bgam(const double tgt, const double tg) const { return tgt * is + tg * ig - rho2 * gam - rho1; }
    __device__ __forceinline__ double rhs(const double tgt, const double tg) const { return tgt * is - ru * bgam(tgt, tg) * iD; }
    // steps of the four quantities for a step sdz = +dz (upper bound) / -dz (lower bound) of the primal entry
    __device__ __forceinline__ void steps(const double sdz, const double tgt, const double tg, double& ds, double& dl, double& dg, double& dn) const {
        dg = (ru * sdz + bgam(tgt, tg)) * iD;
        ds = dg - sdz;
        dl = tgt * is - lam - ru * ds;
        dn = tg * ig - nu - rg * dg;
    }
};

// shared-memory footprint in doubles
template <int NX, int NU>
__host__ __device__ inline size_t qp_warp_soft_extra_doubles(int N) { return 4 * mma_ev((size_t)(N + 1) * (NX + NU)); }
template <int NX, int NU>
__host__ __device__ inline size_t qp_warp_smem_doubles(int N) {
    using D = QpWarp<NX, NU>;
    const size_t NE = (size_t)(N + 1) * D::NZ;
    return D::MAT + 7 * mma_ev(NE) + 2 * mma_ev((size_t)N * NX) + 2 * mma_ev((size_t)N * NU)      // tiles; iterate; rp, Sc; dff, tu
         + mma_ev(NX) + mma_ev(D::KP) + 2 * mma_ev(D::NZ) + mma_ev(NU) + mma_ev(NU * D::LDA)       // gv, pv, rhs, dgn, uc, kk
         + mma_ev(D::NZ + NX) + mma_ev(NU * NU) + 16 + 4;                                         // qdt, R2, profile counters, ring barriers
}

// SOFT: the soft-constrained comparator's QP (every bound row softened, QpArgs::rho / rho1; slack variables and their multipliers in
// four more entry arrays).  The hard-constrained instantiation compiles none of it.
template <int NX, int NU, int NC, bool SOFT = false>
__global__ void __launch_bounds__(32, SOFT ? 4 : 6) qp_warp_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using D = QpWarp<NX, NU>;
    constexpr int NZ = D::NZ, NI = 2 * NZ, NIF = 2 * NX, LDG = Slot<NX>::LD;
    constexpr int KP = D::KP, KS = D::KS, CB = D::CB, CRP = D::CRP, CP = D::CP, NT = D::NT, MTS = D::MTS, MTM = D::MTM;
    constexpr int LDS = D::LDS, LDW = D::LDW, LDA = D::LDA;
    static_assert(LDA == LDG, "the ring copies the engine's padded A_k as it is");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.x;
    if (b >= a.B) return;
    if (a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) return;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int N = NC > 0 ? NC : a.N;

    double* sp = reinterpret_cast<double*>(smem_raw);
    auto take = [&](size_t n) { double* p = sp; sp += mma_ev(n); return p; };
    const int NE = (N + 1) * NZ;
    double* buf1 = take(D::BUF1); double* buf2 = take(D::BUF2);
    double* ring = buf1;                                             // MAT contiguous doubles
    double* e_su = take(NE); double* e_sl = take(NE); double* e_lu = take(NE); double* e_ll = take(NE);
    double* e_zy = take(NE);                                         // z + z_nom (the gradient needs the sum, the outputs the difference)
    double* e_da = take(NE); double* e_dc = take(NE);
    double* e_gu = SOFT ? take(NE) : nullptr; double* e_gl = SOFT ? take(NE) : nullptr;      // slack variables of the upper / lower rows
    double* e_nu = SOFT ? take(NE) : nullptr; double* e_nl = SOFT ? take(NE) : nullptr;      // their multipliers
    double* rp = take((size_t)N * NX); double* Sc = take((size_t)N * NX);
    double* dff = take((size_t)N * NU); double* tu = take((size_t)N * NU);
    double* gv = take(NX); double* pv = take(KP); double* rhs = take(NZ); double* dgn = take(NZ); double* uc = take(NU);
    double* kk = take(NU * LDA);
    double* qdt = take(NZ + NX); double* R2s = take(NU * NU);
    long long* pacc = reinterpret_cast<long long*>(take(16));
    unsigned long long* rbar = reinterpret_cast<unsigned long long*>(take(4));      // one mbarrier per ring slot

    const double* Ag = a.A + (size_t)b * N * NX * LDG;
    const double* Bg = a.Bm + (size_t)b * N * NX * NU;
    const double* cg = a.c + (size_t)b * N * NX;
    const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
    const double* Un = a.Unom + (size_t)b * N * NU;
    double* Kg = a.Kq + (size_t)b * N * NU * LDA;                    // this kernel's layout: [N][NU][LDA]
    double* Lg = a.Lq + (size_t)b * N * NU * NU;

    // four values at once: v0, v1 by maximum, v2, v3 by sum; the butterfly leaves the same bits in every lane
    auto warp_reduce4 = [&](double& v0, double& v1, double& v2, double& v3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
            v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
    };
    auto zero_buf = [&](double* p, const int count) {
        for (int e = 2 * lane; e < count; e += 64) *reinterpret_cast<double2*>(p + e) = make_double2(0.0, 0.0);
    };
    auto rc = [&](const double s) -> double { return s < kBig ? fast_rcp(s) : 0.0; };

    // diagnostic phase timer (option "qp_profile"): cycles of lane 0 of instance 0 between consecutive marks
    const bool prof_on = a.prof != nullptr && b == 0 && lane == 0;
    long long t_last = 0;
    if (prof_on) {
        for (int i = 0; i < 16; ++i) pacc[i] = 0;
        t_last = clock64();
    }
    auto mark = [&](const int slot) {
        if (prof_on) { const long long t = clock64(); pacc[slot] += t - t_last; t_last = t; }
    };

    if (lane == 0) {
        for (int i = 0; i < D::NSLOT; ++i) mbar_init(rbar + i, 1);
        mbar_init_fence();
    }
    unsigned ring_par = 0;                     // bit i: parity of the next completion of slot i (the same in every lane)
    if (lane < KP) pv[lane] = 0.0;
    zero_buf(buf1, D::MAT);                        // tile buffers: finite contents from here on (see phase D)
    if (lane < NU * NU) R2s[lane] = prm->R2[lane];
    for (int e = lane; e < NZ + NX; e += 32)
        qdt[e] = 2.0 * (e < NX ? prm->Q[e * NX + e] : e < NZ ? prm->R[(e - NX) * NU + (e - NX)] : prm->Qf[(e - NZ) * NX + (e - NZ)]);
    __syncwarp();
    auto qd_of = [&](const int k, const int i) -> double { return (k < N) ? qdt[i] : (i < NX ? qdt[NZ + i] : 0.0); };

    // ------------------------------------------------------------------ bounds, initial point, slacks (rnmpc_qp_mma.cuh, cold start)
    const bool tight = SOFT ? false : (a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0));
    double bad = 0.0, n_box = 0.0, s_sum = 0.0, dmy = 0.0;
    {
        const double eps = (tight || SOFT) ? 0.0 : kEpsQP;
#pragma unroll 1
        for (int idx = lane; idx < NE; idx += 32) {
            const int k = idx / NZ, i = idx - k * NZ;
            double h_ = 1e300, l_ = -1e300, z = 0.0;
            bool box = false;
            if (i < NX) {
                const double xn = Xn[k * NX + i];
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
                    if (!SOFT && (z > h_ + kX0Tol || z < l_ - kX0Tol)) bad = 1.0;      // (softened: its slack absorbs the violation, a constant)
                } else box = true;
            } else if (k < N) {
                const int p = i - NX;
                const double un = Un[k * NU + p];
                const double bu = tight ? a.backoff[((size_t)b * N + k) * NI + i] : 0.0;
                const double bl = tight ? a.backoff[((size_t)b * N + k) * NI + NZ + i] : 0.0;
                h_ = (prm->uub[p] - un) - bu + eps;
                l_ = -((un - prm->ulb[p]) - bl + eps);
                box = true;
            }
            double su = 1e300, sl = 1e300;
            if (SOFT) {
                // start on the nominal trajectory (z = 0); e_su / e_sl carry the row violations v = +-z - bound until the second pass
                if (box) {
                    if (h_ < kBig) { su = z - h_; n_box += 2.0; s_sum += fmax(-su, 0.0); }
                    if (l_ > -kBig) { sl = l_ - z; n_box += 2.0; s_sum += fmax(-sl, 0.0); }
                }
            } else if (box) {
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
            e_zy[idx] = z;                                                             // z for now: the residual below needs it
            e_da[idx] = 0.0; e_dc[idx] = 0.0;
        }
    }
    warp_reduce4(bad, dmy, n_box, s_sum);
    int status = bad > 0.0 ? 2 : 0;
    {   // centred cold start: multipliers mu_bar / slack (QpArgs::mu0)
        const double mu_bar = a.mu0 < 0.0 ? (n_box > 0.0 ? -a.mu0 * s_sum / n_box * (SOFT ? 2.0 : 1.0) : 1.0) : a.mu0;
#pragma unroll 1
        for (int idx = lane; idx < NE; idx += 32) {
            if (SOFT) {
                // slack variables start as far from their bound as the row slacks do on average (the primal steps of the first
                // iterations are of that size: a start near the boundary, however well centred, blocks them): gam = max(v, 0) + g0,
                // row slack s = gam - v, multipliers mu_bar / s and mu_bar / gam
                const double g0 = n_box > 0.0 ? 2.0 * s_sum / n_box : 1.0;
                const double vu = e_su[idx], vl = e_sl[idx];
                double su = 1e300, sl = 1e300, lu = 0.0, ll = 0.0, gu = 1.0, gl = 1.0, nu_ = 0.0, nl_ = 0.0;
                if (vu < kBig) { gu = fmax(vu, 0.0) + g0; su = gu - vu; lu = mu_bar / su; nu_ = mu_bar / gu; }
                if (vl < kBig) { gl = fmax(vl, 0.0) + g0; sl = gl - vl; ll = mu_bar / sl; nl_ = mu_bar / gl; }
                e_su[idx] = su; e_sl[idx] = sl; e_lu[idx] = lu; e_ll[idx] = ll;
                e_gu[idx] = gu; e_gl[idx] = gl; e_nu[idx] = nu_; e_nl[idx] = nl_;
                continue;
            }
            const double su = e_su[idx], sl = e_sl[idx];
            double lu, ll;
            if (a.mu0 != 0.0) { lu = su < kBig ? mu_bar / su : 0.0; ll = sl < kBig ? mu_bar / sl : 0.0; }
            else { lu = su < kBig ? 1.0 : 0.0; ll = sl < kBig ? 1.0 : 0.0; }
            e_lu[idx] = lu; e_ll[idx] = ll;
        }
    }
    __syncwarp();
    // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}
    double rp_inf = 0.0, mu_acc = 0.0;
#pragma unroll 1
    for (int e = lane; e < N * NX; e += 32) {
        const int k = e / NX, r = e - k * NX;
        const double* Ar = Ag + ((size_t)k * NX + r) * LDG;
        const double* Br = Bg + ((size_t)k * NX + r) * NU;
        double acc = cg[e] - e_zy[(k + 1) * NZ + r], acc2 = 0.0;
#pragma unroll
        for (int m = 0; m + 1 < NX; m += 2) {
            const double2 av = __ldg(reinterpret_cast<const double2*>(Ar + m));
            acc = fma(av.x, e_zy[k * NZ + m], acc); acc2 = fma(av.y, e_zy[k * NZ + m + 1], acc2);
        }
        if (NX & 1) acc = fma(__ldg(Ar + NX - 1), e_zy[k * NZ + NX - 1], acc);
#pragma unroll
        for (int p = 0; p < NU; ++p) acc = fma(__ldg(Br + p), e_zy[k * NZ + NX + p], acc);
        acc += acc2;
        rp[e] = acc;
        rp_inf = fmax(rp_inf, fabs(acc));
    }
    __syncwarp();
#pragma unroll 1
    for (int idx = lane; idx < NE; idx += 32) {
        const int k = idx / NZ, i = idx - k * NZ;
        if (e_su[idx] < kBig) { mu_acc = fma(e_su[idx], e_lu[idx], mu_acc); if (SOFT) mu_acc = fma(e_gu[idx], e_nu[idx], mu_acc); }
        if (e_sl[idx] < kBig) { mu_acc = fma(e_sl[idx], e_ll[idx], mu_acc); if (SOFT) mu_acc = fma(e_gl[idx], e_nl[idx], mu_acc); }
        if (i < NX) e_zy[idx] += Xn[k * NX + i];                                       // from here on: z + z_nom
        else if (k < N) e_zy[idx] += Un[k * NU + i - NX];
    }
    {
        double d1 = 0.0, d3 = 0.0;
        warp_reduce4(rp_inf, d1, mu_acc, d3);
    }
    double mu_c = n_box > 0.0 ? mu_acc / n_box : 0.0;
    double rp_scale = 1.0;
    __syncwarp();

    double sigma_mu = 0.0;
    const double prox = (SOFT && a.prox != nullptr) ? a.prox[b] : 0.0;
    auto zn_of = [&](const int k, const int i) -> double { return i < NX ? Xn[k * NX + i] : Un[k * NU + i - NX]; };
    // the two softened rows of an entry (SOFT): state + derived coefficients
    auto soft_side = [&](const int idx, const bool upper, SoftSide& sd) -> bool {
        sd.s = upper ? e_su[idx] : e_sl[idx];
        if (!(sd.s < kBig)) return false;
        sd.lam = upper ? e_lu[idx] : e_ll[idx];
        sd.gam = upper ? e_gu[idx] : e_gl[idx];
        sd.nu = upper ? e_nu[idx] : e_nl[idx];
        sd.rho2 = 2.0 * a.rho;
        sd.rho1 = idx < N * NZ ? a.rho1 : 0.0;
        sd.prep();
        return true;
    };
    // complementarity targets of the corrector for one side: sigma mu - (affine step products), from the affine primal step dza
    auto soft_targets = [&](const SoftSide& sd, const double sdza, double& tgt, double& tg) {
        double ds, dl, dg, dn;
        sd.steps(sdza, 0.0, 0.0, ds, dl, dg, dn);
        tgt = sigma_mu - ds * dl;
        tg = sigma_mu - dg * dn;
    };
    auto sg_of = [&](const int idx) -> double {
        if (SOFT) {
            SoftSide sd;
            double w = prox;
            if (soft_side(idx, true, sd)) w += sd.weight();
            if (soft_side(idx, false, sd)) w += sd.weight();
            return w;
        }
        return e_lu[idx] * rc(e_su[idx]) + e_ll[idx] * rc(e_sl[idx]);
    };
    // right-hand side of one entry (rnmpc_qp.cuh rhs_eval): gradient (+ centring / second-order terms in the corrector pass)
    auto rhs_entry = [&](const int pass, const int k, const int i) -> double {
        const int idx = k * NZ + i;
        double gq = qd_of(k, i) * e_zy[idx];
        if (SOFT) {
            // the softened rows contribute in the predictor too: lam + d lam = w dz + (tgt / s - ru b / D)
            const double dza = e_da[idx];
            if (prox != 0.0 && (i < NX || k < N)) gq = fma(prox, e_zy[idx] - zn_of(k, i), gq);
            SoftSide sd;
            double tgt = 0.0, tg = 0.0;
            if (soft_side(idx, true, sd)) { if (pass == 1) soft_targets(sd, dza, tgt, tg); gq += sd.rhs(tgt, tg); }
            tgt = 0.0; tg = 0.0;
            if (soft_side(idx, false, sd)) { if (pass == 1) soft_targets(sd, -dza, tgt, tg); gq -= sd.rhs(tgt, tg); }
            return gq;
        }
        if (pass == 1) {
            const double su = e_su[idx], sl = e_sl[idx], lu = e_lu[idx], ll = e_ll[idx], dza = e_da[idx];
            if (su < kBig) { const double iu = fast_rcp(su); gq = fma(sigma_mu + dza * (lu * iu * dza - lu), iu, gq); }
            if (sl < kBig) { const double il = fast_rcp(sl); gq = fma(-(sigma_mu + dza * (ll * il * dza + ll)), il, gq); }
        }
        return gq;
    };
    // fragments of Z_k = [A_k | B_k | rp_k] (rows padded to KP with zeros): zf[s][n] = Z[4s + t][8n + g]; A_k, B_k from global
    // memory (L2), one stage ahead of their use
    ZFrag<NX, NU, NT, KS, true> zsrc;
    zsrc.init(Ag, Bg, rp, lane);
    auto load_z = [&](const int k, double (&zf)[KS][NT]) { zsrc.load(k, Ag, zf); };
    // stage k of the sweep ring: A_k, B_k, K_k -> slot k % 3 as three bulk copies (TMA unit) issued by one lane, completion on the
    // slot's mbarrier.  The caller has made sure (__syncwarp) that no lane still reads the slot.
    auto ring_issue = [&](const int k) {
        if (lane == 0 && k >= 0 && k < N) {
            const int si = k % D::NSLOT;
            double* sl = ring + si * D::SLOT;
            mbar_expect_tx(rbar + si, (unsigned)(D::SLOT * sizeof(double)));
            bulk_g2s(sl, Ag + (size_t)k * NX * LDG, (unsigned)(NX * LDA * sizeof(double)), rbar + si);
            bulk_g2s(sl + D::OFFB, Bg + (size_t)k * NX * NU, (unsigned)(NX * NU * sizeof(double)), rbar + si);
            bulk_g2s(sl + D::OFFK, Kg + (size_t)k * NU * LDA, (unsigned)(NU * LDA * sizeof(double)), rbar + si);
        }
    };
    auto ring_wait = [&](const int k) -> const double* {
        const int si = k % D::NSLOT;
        mbar_wait(rbar + si, (ring_par >> si) & 1u);
        ring_par ^= 1u << si;
        return ring + si * D::SLOT;
    };
    mark(0);

    int stalled = 0, it = 0;
    bool done = status != 0;
#pragma unroll 1
    for (it = 0; it < a.max_iter && !done; ++it) {
        bool chol_ok = true;
        double ap = 1.0, ad = 1.0, maxp0 = 0.0, maxd0 = 0.0;
        bool skip_corr = false;
        __syncwarp();
        // ================================================================ factorisation sweep + predictor backward recursion
        zero_buf(buf1, 8 * MTS * LDS);
        __syncwarp();
        if (lane < NX) { buf1[lane * LDS + lane] = qdt[NZ + lane] + sg_of(N * NZ + lane); pv[lane] = rhs_entry(0, N, lane); }
        {
            double zf[KS][NT];
            load_z(N - 1, zf);
            __syncwarp();
            mark(1);
#pragma unroll 1
            for (int k = N - 1; k >= 0; --k) {
                double zn_[KS][NT];
                if (k > 0) load_z(k - 1, zn_);
                if (lane < NZ) {
                    rhs[lane] = (lane < NX && k == 0) ? 0.0 : rhs_entry(0, k, lane);
                    dgn[lane] = (lane < NX ? qdt[lane] : 0.0) + sg_of(k * NZ + lane);   // barrier weights (+ diagonal state cost); R2 is added to H below
                }
                // ---- phase A: W = S [A B rp]
#pragma unroll
                for (int mt = 0; mt < MTS; ++mt) {
                    double sa[KS];
#pragma unroll
                    for (int s = 0; s < KS; ++s) sa[s] = buf1[(8 * mt + g) * LDS + 4 * s + t4];
#pragma unroll
                    for (int n = 0; n < NT; ++n) {
                        double c0 = 0.0, c1 = 0.0;
#pragma unroll
                        for (int s = 0; s < KS; ++s) dmma884(c0, c1, sa[s], zf[s][n]);
                        *reinterpret_cast<double2*>(buf2 + (8 * mt + g) * LDW + 8 * n + 2 * t4) = make_double2(c0, c1);
                    }
                }
                __syncwarp();
                mark(2);
                if (lane < NX) Sc[k * NX + lane] = buf2[lane * LDW + CRP];
                // ---- phase B: M = [A B]^T [W p]  (upper triangle of tiles) over the S buffer (S_{k+1} is no longer needed)
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    double wf[KS];
#pragma unroll
                    for (int s = 0; s < KS; ++s) wf[s] = (8 * n + g == CP) ? pv[4 * s + t4] : buf2[(4 * s + t4) * LDW + 8 * n + g];
#pragma unroll
                    for (int mt = 0; mt <= n && mt < MTM; ++mt) {
                        double c0 = 0.0, c1 = 0.0;
#pragma unroll
                        for (int s = 0; s < KS; ++s) dmma884(c0, c1, zf[s][mt], wf[s]);
                        *reinterpret_cast<double2*>(buf1 + (8 * mt + g) * LDW + 8 * n + 2 * t4) = make_double2(c0, c1);
                    }
                }
                __syncwarp();
                mark(3);
                // ---- phase C: H = R2 + diag + B'SB -> Cholesky (every lane its own copy); K = -H^{-1} (A'SB)' column by column;
                //      t = ru + B'(p + rp_scale S rp), d = -H^{-1} t
                {
                    double L[NU][NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p)
#pragma unroll
                        for (int q = 0; q < NU; ++q) {
                            const int lo = p < q ? p : q, hi = p < q ? q : p;
                            L[p][q] = buf1[(CB + lo) * LDW + CB + hi] + R2s[p * NU + q] + (p == q ? dgn[NX + p] : 0.0);
                        }
                    const bool ok = chol_inplace<NU>(L);
                    chol_ok = chol_ok && ok;
                    if (lane < NX) {
                        double f[NU];
#pragma unroll
                        for (int p = 0; p < NU; ++p) f[p] = buf1[lane * LDW + CB + p];
                        chol_solve<NU>(L, f);
#pragma unroll
                        for (int p = 0; p < NU; ++p) { kk[p * LDA + lane] = -f[p]; Kg[((size_t)k * NU + p) * LDA + lane] = -f[p]; }
                    } else if (lane == NX) {
                        double d[NU];
#pragma unroll
                        for (int p = 0; p < NU; ++p) { d[p] = rhs[NX + p] + buf1[(CB + p) * LDW + CP] + rp_scale * buf1[(CB + p) * LDW + CRP]; uc[p] = d[p]; }
                        chol_solve<NU>(L, d);
#pragma unroll
                        for (int p = 0; p < NU; ++p) dff[k * NU + p] = -d[p];
                    } else if (lane == NX + 1) {
#pragma unroll
                        for (int p = 0; p < NU; ++p)
#pragma unroll
                            for (int q = 0; q < NU; ++q) Lg[k * NU * NU + p * NU + q] = L[p][q];
                    }
                }
                __syncwarp();
                mark(4);
                // ---- phase D: one more k-step per tile: S_k = A'SA + (A'SB) K + diag (upper triangle, mirrored: exactly symmetric);
                //      p_k = rx + A'g + K't.  Operands to registers first: the buffer changes from the M layout back to the S layout.
                if (k > 0) {
                    double pk = 0.0;
                    if (lane < NX) {
                        pk = rhs[lane] + buf1[lane * LDW + CP] + rp_scale * buf1[lane * LDW + CRP];
#pragma unroll
                        for (int p = 0; p < NU; ++p) pk = fma(kk[p * LDA + lane], uc[p], pk);
                    }
                    double kf[MTS], mc0[MTS * (MTS + 1) / 2], mc1[MTS * (MTS + 1) / 2], gf[MTS];
#pragma unroll
                    for (int n = 0; n < MTS; ++n) kf[n] = (t4 < NU && 8 * n + g < NX) ? kk[t4 * LDA + 8 * n + g] : 0.0;      // B operand K[t][8n + g]
                    int ti = 0;
#pragma unroll
                    for (int mt = 0; mt < MTS; ++mt) {
                        gf[mt] = (t4 < NU) ? buf1[(8 * mt + g) * LDW + CB + t4] : 0.0;                                        // A operand (A'SB)[r][t]
#pragma unroll
                        for (int n = mt; n < MTS; ++n, ++ti) {
                            const double2 mv = *reinterpret_cast<const double2*>(buf1 + (8 * mt + g) * LDW + 8 * n + 2 * t4);
                            mc0[ti] = mv.x; mc1[ti] = mv.y;
                        }
                    }
                    // (no zeroing: the pads of the S layout may hold finite leftovers of the M tiles -- they only ever multiply the zero
                    // rows of Z_k; the buffers are cleared once per instance, so nothing non-finite can be left in them)
                    __syncwarp();
                    ti = 0;
#pragma unroll
                    for (int mt = 0; mt < MTS; ++mt) {
                        const int r = 8 * mt + g;
#pragma unroll
                        for (int n = mt; n < MTS; ++n, ++ti) {
                            const int c = 8 * n + 2 * t4;
                            double c0 = mc0[ti], c1 = mc1[ti];
                            dmma884(c0, c1, gf[mt], kf[n]);
                            if (r < NX) {
                                if (c < NX && r <= c) { const double v = c0 + (r == c ? dgn[r] : 0.0); buf1[r * LDS + c] = v; buf1[c * LDS + r] = v; }
                                if (c + 1 < NX && r <= c + 1) { const double v = c1 + (r == c + 1 ? dgn[r] : 0.0); buf1[r * LDS + c + 1] = v; buf1[(c + 1) * LDS + r] = v; }
                            }
                        }
                    }
                    if (lane < NX) pv[lane] = pk;          // p_k: the p column of the next stage's second product
#pragma unroll
                    for (int s = 0; s < KS; ++s)
#pragma unroll
                        for (int n = 0; n < NT; ++n) zf[s][n] = zn_[s][n];
                }
                __syncwarp();
                mark(5);
            }
        }
        fence_block();          // L_k written above is read back by other lanes
        fence_proxy_async();    // K_k (global) and the tile buffers (shared) written above are next touched by bulk copies
        __syncwarp();

        // ================================================================ predictor (pass 0) and corrector (pass 1): one copy of the sweep code
        const double tau_c = mu_c > 1e-8 ? 0.995 : 0.9999;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            double* e_dz = pass ? e_dc : e_da;
            if (pass == 1) {
                ring_issue(N - 1); ring_issue(N - 2);
                // ---- corrector right-hand sides (e_dc doubles as scratch until the forward recursion overwrites it with the step)
#pragma unroll 2
                for (int idx = lane; idx < NE; idx += 32) { const int k = idx / NZ; e_dc[idx] = rhs_entry(1, k, idx - k * NZ); }
                __syncwarp();
                mark(6);
                {
                    // backward recursion: g_k = p_{k+1} + rp_scale Sc_k ; lanes < NX: p_k = rx_k + A_k' g_k + K_k' t_k,
                    // lanes NX .. NZ-1: t_k = ru_k + B_k' g_k (kept in tu for the feed-forward)
                    const int r = lane < NX ? lane : NX - 1;
                    const int p = (lane >= NX && lane < NZ) ? lane - NX : 0;
                    double pl = e_dc[N * NZ + r];
#pragma unroll 1
                    for (int k = N - 1; k >= 0; --k) {
                        ring_issue(k - 2);                 // into the slot of stage k + 1 (finished: barrier at the end of the previous stage)
                        if (lane < NX) gv[lane] = pl + rp_scale * Sc[k * NX + lane];
                        const double* sl = ring_wait(k);
                        __syncwarp();
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        if (lane < NX) {
                            if (k > 0) {
                                const double* col = sl + r;
                                a0 = e_dc[k * NZ + r];
#pragma unroll
                                for (int m = 0; m + 3 < NX; m += 4) {
                                    a0 = fma(col[m * LDA], gv[m], a0); a1 = fma(col[(m + 1) * LDA], gv[m + 1], a1);
                                    a2 = fma(col[(m + 2) * LDA], gv[m + 2], a2); a3 = fma(col[(m + 3) * LDA], gv[m + 3], a3);
                                }
#pragma unroll
                                for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(col[m * LDA], gv[m], a0);
                            }
                        } else if (lane < NZ) {
                            const double* col = sl + D::OFFB + p;
                            a0 = e_dc[k * NZ + NX + p];
#pragma unroll
                            for (int m = 0; m + 3 < NX; m += 4) {
                                a0 = fma(col[m * NU], gv[m], a0); a1 = fma(col[(m + 1) * NU], gv[m + 1], a1);
                                a2 = fma(col[(m + 2) * NU], gv[m + 2], a2); a3 = fma(col[(m + 3) * NU], gv[m + 3], a3);
                            }
#pragma unroll
                            for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(col[m * NU], gv[m], a0);
                            tu[k * NU + p] = (a0 + a1) + (a2 + a3);
                        }
                        __syncwarp();
                        if (lane < NX && k > 0) {
                            double acc = (a0 + a1) + (a2 + a3);
#pragma unroll
                            for (int q = 0; q < NU; ++q) acc = fma(sl[D::OFFK + q * LDA + r], tu[k * NU + q], acc);
                            pl = acc;
                        }
                        __syncwarp();
                    }
                }
                mark(7);
                ring_issue(0); ring_issue(1);      // the forward recursion's first stages are in flight during the small solves
                // ---- d_k = -H_k^{-1} t_k   (one lane per stage)
                for (int k = lane; k < N; k += 32) {
                    double d[NU], L[NU][NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) {
                        d[p] = tu[k * NU + p];
#pragma unroll
                        for (int q = 0; q < NU; ++q) L[p][q] = Lg[k * NU * NU + p * NU + q];
                    }
                    chol_solve<NU>(L, d);
#pragma unroll
                    for (int p = 0; p < NU; ++p) dff[k * NU + p] = -d[p];
                }
                __syncwarp();
                mark(8);
            } else {
                ring_issue(0); ring_issue(1);
            }
            mark(9);
            // ---- forward recursion: lanes NX .. NZ-1: du_k = K_k dx_k + d_k ; lanes < NX: dx_{k+1} = A_k dx_k + B_k du_k +
            //      rp_scale rp_k, dx_0 = 0.  The steps land in the entries of e_dz.
            {
                const int r = lane < NX ? lane : NX - 1;
                const int p = (lane >= NX && lane < NZ) ? lane - NX : 0;
                if (lane < NX) e_dz[lane] = 0.0;
#pragma unroll 1
                for (int k = 0; k < N; ++k) {
                    ring_issue(k + 2);                     // into the slot of stage k - 1
                    const double* sl = ring_wait(k);
                    __syncwarp();
                    const double* xk = e_dz + k * NZ;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                    if (lane < NX) {
                        const double* row = sl + r * LDA;
                        a0 = rp_scale * rp[k * NX + r];
                        if (k > 0) {
#pragma unroll
                            for (int m = 0; m + 3 < NX; m += 4) {
                                const double2 r01 = *reinterpret_cast<const double2*>(row + m), r23 = *reinterpret_cast<const double2*>(row + m + 2);
                                a0 = fma(r01.x, xk[m], a0); a1 = fma(r01.y, xk[m + 1], a1);
                                a2 = fma(r23.x, xk[m + 2], a2); a3 = fma(r23.y, xk[m + 3], a3);
                            }
#pragma unroll
                            for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(row[m], xk[m], a0);
                        }
                    } else if (lane < NZ) {
                        const double* row = sl + D::OFFK + p * LDA;
                        a0 = dff[k * NU + p];
                        if (k > 0) {
#pragma unroll
                            for (int m = 0; m + 3 < NX; m += 4) {
                                const double2 r01 = *reinterpret_cast<const double2*>(row + m), r23 = *reinterpret_cast<const double2*>(row + m + 2);
                                a0 = fma(r01.x, xk[m], a0); a1 = fma(r01.y, xk[m + 1], a1);
                                a2 = fma(r23.x, xk[m + 2], a2); a3 = fma(r23.y, xk[m + 3], a3);
                            }
#pragma unroll
                            for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(row[m], xk[m], a0);
                        }
                        e_dz[k * NZ + NX + p] = (a0 + a1) + (a2 + a3);
                    }
                    __syncwarp();
                    if (lane < NX) {
                        double acc = (a0 + a1) + (a2 + a3);
                        const double* Br = sl + D::OFFB + r * NU;
#pragma unroll
                        for (int q = 0; q < NU; ++q) acc = fma(Br[q], xk[NX + q], acc);
                        e_dz[(k + 1) * NZ + lane] = acc;
                    }
                    __syncwarp();
                }
            }
            mark(10);
            // ---- ratio test over all entries (rnmpc_qp.cuh ratio_entry)
            double rpm = 0.0, rdm = 0.0, S3 = 0.0, S4 = 0.0;
#pragma unroll 2
            for (int idx = lane; idx < NE; idx += 32) {
                const double dz = e_dz[idx];
                const double su = e_su[idx], sl = e_sl[idx], lu = e_lu[idx], ll = e_ll[idx];
                const double dza = pass == 1 ? e_da[idx] : 0.0;
                const double smu = pass == 1 ? sigma_mu : 0.0;
                if (SOFT) {
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
                        SoftSide sd;
                        if (!soft_side(idx, side == 0, sd)) continue;
                        const double sg = side == 0 ? 1.0 : -1.0;
                        double tgt = 0.0, tg = 0.0, ds, dl, dg, dn;
                        if (pass == 1) soft_targets(sd, sg * e_da[idx], tgt, tg);
                        sd.steps(sg * dz, tgt, tg, ds, dl, dg, dn);
                        rpm = fmax(rpm, fmax(-ds * sd.is, -dg * sd.ig));
                        rdm = fmax(rdm, fmax(dl < 0.0 ? -dl * fast_rcp(sd.lam) : 0.0, dn < 0.0 ? -dn * fast_rcp(sd.nu) : 0.0));
                        if (pass == 0) { S3 = fma(ds, sd.lam, S3); S3 = fma(dg, sd.nu, S3); S4 = fma(ds, dl, S4); S4 = fma(dg, dn, S4); }
                    }
                    continue;
                }
                if (su < kBig) {
                    const double iu = fast_rcp(su);
                    const double ru = lu * iu;
                    const double tgt = pass == 1 ? smu + dza * (ru * dza - lu) : 0.0;
                    const double dlu = tgt * iu - lu + ru * dz;
                    rpm = fmax(rpm, dz * iu);
                    if (dlu < 0.0) rdm = fmax(rdm, -dlu * fast_rcp(lu));
                    if (pass == 0) { S3 = fma(-dz, lu, S3); S4 = fma(-dz, dlu, S4); }
                }
                if (sl < kBig) {
                    const double il = fast_rcp(sl);
                    const double rl = ll * il;
                    const double tgt = pass == 1 ? smu + dza * (rl * dza + ll) : 0.0;
                    const double dll = tgt * il - ll - rl * dz;
                    rpm = fmax(rpm, -dz * il);
                    if (dll < 0.0) rdm = fmax(rdm, -dll * fast_rcp(ll));
                    if (pass == 0) { S3 = fma(dz, ll, S3); S4 = fma(dz, dll, S4); }
                }
            }
            warp_reduce4(rpm, rdm, S3, S4);
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
        // ================================================================ update, new complementarity
        double acc = 0.0;
        {
            const bool upd = chol_ok;
            const double so = skip_corr ? 0.0 : 1.0;
            const double* e_dz = skip_corr ? e_da : e_dc;
#pragma unroll 2
            for (int idx = lane; idx < NE; idx += 32) {
                double su = e_su[idx], sl = e_sl[idx], lu = e_lu[idx], ll = e_ll[idx];
                if (SOFT) {
                    const double dza = e_da[idx], dz = e_dz[idx];
#pragma unroll
                    for (int side = 0; side < 2; ++side) {
                        SoftSide sd;
                        if (!soft_side(idx, side == 0, sd)) continue;
                        if (upd) {
                            const double sg = side == 0 ? 1.0 : -1.0;
                            double tgt = 0.0, tg = 0.0, ds, dl, dg, dn;
                            if (!skip_corr) soft_targets(sd, sg * dza, tgt, tg);
                            sd.steps(sg * dz, tgt, tg, ds, dl, dg, dn);
                            sd.s += ap * ds; sd.gam += ap * dg; sd.lam += ad * dl; sd.nu += ad * dn;
                            if (side == 0) { e_su[idx] = sd.s; e_lu[idx] = sd.lam; e_gu[idx] = sd.gam; e_nu[idx] = sd.nu; }
                            else { e_sl[idx] = sd.s; e_ll[idx] = sd.lam; e_gl[idx] = sd.gam; e_nl[idx] = sd.nu; }
                        }
                        acc = fma(sd.s, sd.lam, acc); acc = fma(sd.gam, sd.nu, acc);
                    }
                    if (upd) e_zy[idx] += ap * dz;
                    continue;
                }
                if (upd) {
                    const double dza = e_da[idx], dz = e_dz[idx];
                    if (su < kBig) {
                        const double iu = fast_rcp(su);
                        const double ru = lu * iu;
                        const double dlu = so * (sigma_mu + dza * (ru * dza - lu)) * iu - lu + ru * dz;
                        su -= ap * dz; lu += ad * dlu;
                    }
                    if (sl < kBig) {
                        const double il = fast_rcp(sl);
                        const double rl = ll * il;
                        const double dll = so * (sigma_mu + dza * (rl * dza + ll)) * il - ll - rl * dz;
                        sl += ap * dz; ll += ad * dll;
                    }
                    e_su[idx] = su; e_sl[idx] = sl; e_lu[idx] = lu; e_ll[idx] = ll;
                    e_zy[idx] += ap * dz;
                }
                if (su < kBig) acc = fma(su, lu, acc);
                if (sl < kBig) acc = fma(sl, ll, acc);
            }
            if (upd) rp_scale *= (1.0 - ap);
        }
        {
            double d0 = 0.0, d1 = 0.0, d3 = 0.0;
            warp_reduce4(d0, d1, acc, d3);
        }
        mark(12);
        const double mu_new = n_box > 0.0 ? acc / n_box : 0.0;
#ifdef RNMPC_EMU_DEBUG
        if (lane == 0) std::printf("it %d mu %.3e -> %.3e ap %.3e ad %.3e sigma_mu %.3e rp %.3e skip %d chol %d\n", it, mu_c, mu_new, ap, ad, sigma_mu, rp_inf * rp_scale, (int)skip_corr, (int)chol_ok);
#endif
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
    __syncwarp();

    // ---- outputs in the reference layouts + objective x'Hx + q'x (qp_jit.py:401, P = 2H, q = 2 H y_nom; diagonal H)
    {
        double* Xo = a.dX + (size_t)b * (N + 1) * NX;
        double* Uo = a.dU + (size_t)b * N * NU;
        double* mu = a.mu + (size_t)b * N * NI;
        double* muf = a.mu_f + (size_t)b * NIF;
        double acc = 0.0;
#pragma unroll 1
        for (int idx = lane; idx < NE; idx += 32) {
            const int k = idx / NZ, i = idx - k * NZ;
            if (i >= NX && k == N) continue;
            const double zn = i < NX ? Xn[k * NX + i] : Un[k * NU + i - NX];
            const double z = e_zy[idx] - zn;
            acc = fma(z, 0.5 * qd_of(k, i) * (z + 2.0 * zn), acc);
            if (SOFT) {
                // slack variables in the reference's row order [x_ub, u_ub, x_lb, u_lb] (+ their cost); the rows of the pinned x(0) carry
                // their violation
                double gu = e_su[idx] < kBig ? e_gu[idx] : 0.0, gl = e_sl[idx] < kBig ? e_gl[idx] : 0.0;
                if (k == 0 && i < NX) { gu = fmax(0.0, (z + zn) - prm->xub[i]); gl = fmax(0.0, prm->xlb[i] - (z + zn)); }
                const double r1 = k < N ? a.rho1 : 0.0;
                acc += a.rho * (gu * gu + gl * gl) + r1 * (gu + gl);
                if (a.gamma != nullptr) {
                    if (k < N) { a.gamma[((size_t)b * N + k) * NI + i] = gu; a.gamma[((size_t)b * N + k) * NI + NZ + i] = gl; }
                    else { a.gamma_f[(size_t)b * NIF + i] = gu; a.gamma_f[(size_t)b * NIF + NX + i] = gl; }
                }
            }
            if (i < NX) Xo[k * NX + i] = z; else Uo[k * NU + i - NX] = z;
            if (k < N) { mu[k * NI + i] = e_lu[idx]; mu[k * NI + NZ + i] = e_ll[idx]; }
            else { muf[i] = e_lu[idx]; muf[NX + i] = e_ll[idx]; }
        }
        double d0 = 0.0, d1 = 0.0, d3 = 0.0;
        warp_reduce4(d0, d1, acc, d3);
        mark(13);
        if (prof_on) { for (int i = 0; i < 16; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(a.prof) + i, (unsigned long long)pacc[i]); }
        if (lane == 0) {
            a.cost[b] = acc;
            a.iters[b] = it;
            a.status[b] = status;
            if (a.ws_valid != nullptr) a.ws_valid[b] = 0;          // the records of qp_ipm_kernel's workspace were not written
        }
    }
}

}  // namespace rnmpc
