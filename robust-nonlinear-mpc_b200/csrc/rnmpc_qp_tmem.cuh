// Nominal QP, throughput kernel for large batches, second version: one warp per instance as in qp_warp_kernel (rnmpc_qp_warp.cuh; same
// problem, algorithm, formulas and status codes; reference: solver/qp_jit.py:77-273, 362-402, 487-513), with the interior-point
// iterate in TENSOR MEMORY instead of shared memory.
//
// Why (measured on qp_warp_kernel, profiles/r2_notes.md):
//   * its 36 KB of shared memory per instance allow 6 warps per SM; the warps are latency bound (1.4 warps per scheduler, "wait"
//     stalls 2.9 per issue), so throughput follows the number of resident instances;
//   * half of that shared memory is the iterate (slacks, multipliers, primal, two steps: 7 arrays x (N+1) nz doubles), which is
//     only ever touched entry-wise: component i of stage k by one lane.  That is the access pattern of TMEM's 32x32b load / store
//     shape: lane i of a warp owns TMEM lane i, the column address selects the stage at run time (rnmpc_common.cuh, tm_ld / tm_st;
//     checked by tools/tmem_probe.cu).  All arrays of a stage sit in 16 consecutive columns: one tcgen05.ld fetches them together;
//   * its vector sweeps wait on the three-slot ring: a slot takes ~2400 cycles to arrive (L2 round trip under load + TMA), the lead
//     is two stages, so a stage lasts 1200 cycles where its arithmetic needs 250.  The shared memory freed by the iterate holds two more
//     slots (lead of four stages);
//   * with the iterate indexed by (stage, lane) every element-wise pass that only needs the entries of one stage is folded into the
//     sweep that visits the stage anyway -- barrier weights and gradients into the factorisation, the corrector's right-hand side
//     into its backward recursion, the ratio tests into the forward recursions -- where it runs while the next ring slot is in
//     flight.  Only the update of the iterate remains a pass of its own.
// A CTA is four independent warps (TMEM is allocated per CTA and a warp reaches one quarter of its lanes); the warps draw instances
// from a global counter until the batch is exhausted, so a warp that finishes early (8 to 25 iterations) starts the next instance.
// Requires N + 1 <= 16 (16 columns per stage, 256 columns per CTA, two CTAs per SM share the 512 columns).
#pragma once
#include <type_traits>
#include "rnmpc_common.cuh"
#include "rnmpc_qp.cuh"
#include "rnmpc_qp_mma.cuh"
#include "rnmpc_qp_warp.cuh"

namespace rnmpc {

constexpr int kTmWarps = 4;
constexpr int kTmMaxStages = 16;     // N + 1

template <int NX, int NU>
struct QpTm {
    using W = QpWarp<NX, NU>;
    static constexpr int NSLOT = 5;
    static constexpr int SLOTS = (W::MAT / 3) & ~1;            // slot stride: slots 0..2 fill the tile buffers, 3.. follow them contiguously
    static_assert(SLOTS >= W::SLOT, "three slots fit the tile buffers");
    static constexpr int EXTRA = (NSLOT - 3) * SLOTS + (W::MAT - 3 * SLOTS);        // ring slots 3, 4 (slots 0..2 live in the tile buffers)
};

// shared-memory footprint of one warp in doubles
template <int NX, int NU>
__host__ __device__ inline size_t qp_tm_warp_doubles(int N) {
    using D = QpWarp<NX, NU>;
    return D::MAT + QpTm<NX, NU>::EXTRA + 2 * mma_ev((size_t)N * NX) + 2 * mma_ev((size_t)N * NU)    // tiles, ring; rp, Sc; dff, tu
         + 2 * mma_ev(D::NZ) + mma_ev(NX) + mma_ev(D::KP) + 2 * mma_ev(D::NZ) + mma_ev(NU) + mma_ev(NU * D::LDA)   // xs[2], gv, pv, rhs, dgn, uc, kk
         + mma_ev(D::NZ + NX) + mma_ev(NU * NU) + 16 + 6;                                          // qdt, R2, profile counters, ring barriers
}
template <int NX, int NU>
__host__ __device__ inline size_t qp_tm_smem_bytes(int N) { return (kTmWarps * qp_tm_warp_doubles<NX, NU>(N) + 2) * sizeof(double); }

// TMEM columns of a stage: 8 doubles
enum { TM_SU = 0, TM_SL = 1, TM_LU = 2, TM_LL = 3, TM_ZY = 4, TM_DA = 5, TM_DC = 6, TM_ZN = 7 };    // ZN: nominal entry (constant)

// PROF: the phase timer (option "qp_profile") is its own instantiation -- its predicated clock reads between the phases cost the
// production kernel 1.2 % when they were merely switched off at run time.
template <int NX, int NU, int NC, bool PROF = false>
__global__ void __launch_bounds__(32 * kTmWarps, 2) qp_tm_kernel(const QpArgs a, const Params<NX, NU>* __restrict__ prm) {
    using D = QpWarp<NX, NU>;
    using T = QpTm<NX, NU>;
    constexpr int NZ = D::NZ, NI = 2 * NZ, NIF = 2 * NX, LDG = Slot<NX>::LD;
    constexpr int KP = D::KP, KS = D::KS, CB = D::CB, CRP = D::CRP, CP = D::CP, NT = D::NT, MTS = D::MTS, MTM = D::MTM;
    constexpr int LDS = D::LDS, LDW = D::LDW, LDA = D::LDA;
    static_assert(LDA == LDG, "the ring copies the engine's padded A_k as it is");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = (int)uniform_u32(threadIdx.x >> 5);      // warp index as a uniform value: shared-memory bases below are uniform
    const int g = lane >> 2, t4 = lane & 3;
    const int N = NC > 0 ? NC : a.N;

    // ---- tensor memory: one allocation per CTA, each warp its quarter of the lanes
    unsigned* tm_slot = reinterpret_cast<unsigned*>(smem_raw);
    if (warp == 0) tmem_alloc_256(tm_slot);
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    const unsigned tm_cta = *tm_slot;
    const unsigned tb = tmem_warp_base(tm_cta, warp);
    auto tcol = [&](const int k, const int arr) -> unsigned { return tb + (unsigned)(16 * k + 2 * arr); };

    double* sp = reinterpret_cast<double*>(smem_raw) + 2 + (size_t)warp * qp_tm_warp_doubles<NX, NU>(N);
    auto take = [&](size_t n) { double* p = sp; sp += mma_ev(n); return p; };
    double* buf1 = take(D::BUF1); double* buf2 = take(D::BUF2);
    double* ring_lo = buf1;                                          // slots 0..2
    double* ring_hi = take(T::EXTRA);                                // slots 3..
    double* rp = take((size_t)N * NX); double* Sc = take((size_t)N * NX);
    double* dff = take((size_t)N * NU); double* tu = take((size_t)N * NU);
    double* xs0 = take(NZ); double* xs1 = take(NZ);
    double* gv = take(NX); double* pv = take(KP); double* rhs = take(NZ); double* dgn = take(NZ); double* uc = take(NU);
    double* kk = take(NU * LDA);
    double* qdt = take(NZ + NX); double* R2s = take(NU * NU);
    long long* pacc = reinterpret_cast<long long*>(take(16));
    unsigned long long* rbar = reinterpret_cast<unsigned long long*>(take(6));      // one mbarrier per ring slot

    auto warp_reduce4 = [&](double& v0, double& v1, double& v2, double& v3) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 = fmax(v0, __shfl_xor_sync(0xffffffffu, v0, o)); v1 = fmax(v1, __shfl_xor_sync(0xffffffffu, v1, o));
            v2 += __shfl_xor_sync(0xffffffffu, v2, o); v3 += __shfl_xor_sync(0xffffffffu, v3, o);
        }
    };
    // Loops in front of the warp-collective tcgen05 loads / stores have WARP-UNIFORM trip counts and predicated bodies: with a
    // lane-dependent trip count here ptxas moved the loop's reconvergence point behind the first TMEM stores of an instance and
    // turned the bar.warp.sync in front of them into NOPs -- the lanes then ran the stores in two groups and the first group's were
    // lost (pendulum instantiation, found on the device; the CPU emulator cannot see it).
    auto zero_buf = [&](double* p, const int count) {
        for (int e0 = 0; e0 < count; e0 += 64) {
            const int e = e0 + 2 * lane;
            if (e < count) *reinterpret_cast<double2*>(p + e) = make_double2(0.0, 0.0);
        }
    };
    auto rc = [&](const double s) -> double { return s < kBig ? fast_rcp(s) : 0.0; };

    if (lane == 0) {
        for (int i = 0; i < T::NSLOT; ++i) mbar_init(rbar + i, 1);
        mbar_init_fence();
    }
    unsigned ring_par = 0;                     // bit i: parity of the next completion of slot i (the same in every lane)
    if (lane < NU * NU) R2s[lane] = prm->R2[lane];
    for (int e0 = 0; e0 < NZ + NX; e0 += 32) {
        const int e = e0 + lane;
        if (e < NZ + NX) qdt[e] = 2.0 * (e < NX ? prm->Q[e * NX + e] : e < NZ ? prm->R[(e - NX) * NU + (e - NX)] : prm->Qf[(e - NZ) * NX + (e - NZ)]);
    }
    __syncwarp();
    // 2 H_ii of entry (k, lane); 0 for lanes that hold no entry
    const double qd_s = lane < NZ ? qdt[lane] : 0.0, qd_t = lane < NX ? qdt[NZ + lane] : 0.0;      // this lane's entry: stage / terminal value
    auto qd_of = [&](const int k) -> double { return k < N ? qd_s : qd_t; };

#pragma unroll 1
    for (;;) {
        // ---- next instance of the batch
        int b = 0;
        if (lane == 0) b = atomicAdd(a.work, 1);
        b = (int)uniform_u32((unsigned)__shfl_sync(0xffffffffu, b, 0));      // uniform for the compiler too: global bases of the instance are uniform
        if (b >= a.B) break;
        if (a.order != nullptr) b = a.order[b];
        if (a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) continue;

        const double* Ag = a.A + (size_t)b * N * NX * LDG;
        const double* Bg = a.Bm + (size_t)b * N * NX * NU;
        const double* cg = a.c + (size_t)b * N * NX;
        const double* Xn = a.Xnom + (size_t)b * (N + 1) * NX;
        const double* Un = a.Unom + (size_t)b * N * NU;
        // sweep image of the instance in the workspace: [N][A_k (NX x LDA) | B_k (NX x NU) | K_k (NU x LDA)], one contiguous block per stage,
        // so that a ring slot is ONE bulk copy (three copies -- A_k, B_k, K_k from their own arrays -- cost the issuing lane ~300 cycles
        // per stage on the critical path of the sweeps: profiles/r2_notes.md).  A_k, B_k are copied in once per solve, K_k is written
        // by the factorisation.
        double* Wg = a.Kq + (size_t)b * N * D::SLOT;
        double* Lg = a.Lq + (size_t)b * N * NU * NU;

        // diagnostic phase timer (option "qp_profile"): cycles of lane 0 of instance 0 between consecutive marks
        const bool prof_on = PROF && a.prof != nullptr && b == 0 && lane == 0;
        long long t_last = 0;
        if (prof_on) {
            for (int i = 0; i < 16; ++i) pacc[i] = 0;
            t_last = clock64();
        }
        auto mark = [&](const int slot) {
            if (PROF && prof_on) { const long long t = clock64(); pacc[slot] += t - t_last; t_last = t; }
        };
        if (lane < KP) pv[lane] = 0.0;
        zero_buf(buf1, D::MAT + T::EXTRA);          // tile buffers + ring: finite contents from here on (see phase D)
        __syncwarp();

        // ------------------------------------------------------------------ bounds, initial point, slacks (rnmpc_qp_mma.cuh, cold start)
        // lane i handles component i of every stage; lanes without an entry hold "no bound" slacks and zero multipliers
        const bool tight = a.tightened < 0 ? ((a.flags[b] & FLAG_TIGHT) != 0) : (a.tightened != 0);
        double bad = 0.0, n_box = 0.0, s_sum = 0.0, dmy = 0.0;
        double* ztmp = buf1;                       // z of the initial point, [N+1][NZ], for the residual below (tile buffers are idle)
        unsigned nzmask = 0;                       // bit k: the initial point has a nonzero entry in stage k (the same in every lane)
        {
            const double eps = tight ? 0.0 : kEpsQP;
            const int i = lane;
            const bool isx = i < NX, isu = !isx && i < NZ;
            // The inputs of this pass are cold (first touch of the instance): everything that does not depend on the stage is loaded once,
            // the nominal entry and the back-offs of EIGHT stages are requested before the first of them is used (the TMEM stores are
            // volatile asm, so the compiler cannot overlap the misses of consecutive stages by itself: 84 k -> ~20 k cycles per QP).
            double ub = 1e300, lb = -1e300, ubf = 1e300, lbf = -1e300, x0i = 0.0, bfu = 0.0, bfl = 0.0;
            if (isx) {
                ub = prm->xub[i]; lb = prm->xlb[i]; ubf = prm->xubf[i]; lbf = prm->xlbf[i];
                x0i = a.x0[(size_t)b * NX + i];
                if (tight) { bfu = a.backoff_f[(size_t)b * NIF + i]; bfl = a.backoff_f[(size_t)b * NIF + NX + i]; }
            } else if (isu) { ub = prm->uub[i - NX]; lb = prm->ulb[i - NX]; }
            const double* znp = isx ? Xn + i : Un + (isu ? i - NX : 0);
            const int zn_stride = isx ? NX : NU;
            const double* bop = a.backoff + (size_t)b * N * NI + (i < NZ ? i : 0);
            constexpr int CH = 8;
#pragma unroll 1
            for (int k0 = 0; k0 <= N; k0 += CH) {
                double zn_[CH], bu_[CH], bl_[CH];
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int k = k0 + j;
                    const bool has = isx ? (k <= N) : (isu && k < N);
                    const bool hb = tight && i < NZ && k < N;
                    zn_[j] = has ? znp[k * zn_stride] : 0.0;
                    bu_[j] = hb ? bop[(size_t)k * NI] : 0.0;
                    bl_[j] = hb ? bop[(size_t)k * NI + NZ] : 0.0;
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const int k = k0 + j;
                    if (k > N) break;                      // warp-uniform
                    const double zn = zn_[j];
                    double h_ = 1e300, l_ = -1e300, z = 0.0;
                    bool box = false;
                    if (isx) {
                        if (k < N) {
                            h_ = (ub - zn) - bu_[j] + eps;
                            l_ = -((zn - lb) - bl_[j] + eps);
                        } else if (tight) {
                            h_ = ubf - bfu;                                                     // Q4: terminal bound not shifted by x_N
                            l_ = -(-lbf - bfl);
                        } else {
                            h_ = (ubf - zn) + eps;
                            l_ = -((zn - lbf) + eps);
                        }
                        if (k == 0) {
                            z = x0i - zn;                                                       // x(0) is pinned to the measurement
                            if (z > h_ + kX0Tol || z < l_ - kX0Tol) bad = 1.0;
                        } else box = true;
                    } else if (isu && k < N) {
                        h_ = (ub - zn) - bu_[j] + eps;
                        l_ = -((zn - lb) - bl_[j] + eps);
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
                    if (i < NZ) ztmp[k * NZ + i] = z;
                    if (i < NZ && z != 0.0) nzmask |= 1u << k;                   // this lane's entries; combined over the warp below
                    const double v[8] = {su, sl, 0.0, 0.0, z, 0.0, 0.0, zn};      // the nominal entry rides in the spare column
                    tm_st<8>(tcol(k, 0), v);
                }
            }
        }
        tm_wait_st();
        warp_reduce4(bad, dmy, n_box, s_sum);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nzmask |= __shfl_xor_sync(0xffffffffu, nzmask, o);
        int status = bad > 0.0 ? 2 : 0;
        __syncwarp();
        // dynamics residual at the initial point: rp_k = A x_k + B u_k + c_k - x_{k+1}.  The initial point is zero wherever zero lies inside
        // the bounds: only the stages with a nonzero entry (stage 0, clipped entries) read their A_k and B_k here.
        double rp_inf = 0.0, mu_acc = 0.0;
#pragma unroll 4
        for (int e0 = 0; e0 < N * NX; e0 += 32) {
            const int e = e0 + lane;
            if (e < N * NX) {
                const int k = e / NX, r = e - k * NX;
                rp[e] = cg[e] - ztmp[(k + 1) * NZ + r];
            }
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            if (!((nzmask >> k) & 1u)) continue;           // warp-uniform
            if (lane < NX) {
                const int r = lane;
                const double* Ar = Ag + ((size_t)k * NX + r) * LDG;
                const double* Br = Bg + ((size_t)k * NX + r) * NU;
                double acc = rp[k * NX + r], acc2 = 0.0;
#pragma unroll
                for (int m = 0; m + 1 < NX; m += 2) {
                    const double2 av = __ldg(reinterpret_cast<const double2*>(Ar + m));
                    acc = fma(av.x, ztmp[k * NZ + m], acc); acc2 = fma(av.y, ztmp[k * NZ + m + 1], acc2);
                }
                if (NX & 1) acc = fma(__ldg(Ar + NX - 1), ztmp[k * NZ + NX - 1], acc);
#pragma unroll
                for (int p = 0; p < NU; ++p) acc = fma(__ldg(Br + p), ztmp[k * NZ + NX + p], acc);
                rp[k * NX + r] = acc + acc2;
            }
        }
        __syncwarp();
        for (int e0 = 0; e0 < N * NX; e0 += 32) {
            const int e = e0 + lane;
            if (e < N * NX) rp_inf = fmax(rp_inf, fabs(rp[e]));
        }
        {   // A_k, B_k -> sweep image (16-byte copies, five stages in flight; the bulk copies that read it are behind the first factorisation
            // and its fence.proxy.async)
            constexpr int CA = NX * LDA / 2, CB2 = NX * NU / 2, CT = CA + CB2;      // double2 per stage: A part, B part
            static_assert((D::OFFB % 2) == 0 && (D::SLOT % 2) == 0, "16-byte aligned image parts");
#pragma unroll 5
            for (int k = 0; k < N; ++k) {
                const double2* As = reinterpret_cast<const double2*>(Ag + (size_t)k * NX * LDG);
                const double2* Bs = reinterpret_cast<const double2*>(Bg + (size_t)k * NX * NU);
                double2* Wd = reinterpret_cast<double2*>(Wg + (size_t)k * D::SLOT);
#pragma unroll
                for (int c0 = 0; c0 < CT; c0 += 32) {
                    const int c = c0 + lane;
                    if (c < CT) Wd[c] = c < CA ? __ldg(As + c) : __ldg(Bs + (c - CA));      // B part follows the A part (OFFB = NX * LDA)
                }
            }
            fence_block();          // the image is read back by other lanes (operand fragments) and by the bulk copies
            __syncwarp();
        }
        {   // centred cold start: multipliers mu_bar / slack (QpArgs::mu0); the primal entry becomes z + z_nom
            const double mu_bar = a.mu0 < 0.0 ? (n_box > 0.0 ? -a.mu0 * s_sum / n_box : 1.0) : a.mu0;
#pragma unroll 1
            for (int k = 0; k <= N; ++k) {
                double v[8];
                tm_ld<8>(tcol(k, 0), v);
                const double su = v[TM_SU], sl = v[TM_SL];
                double lu, ll;
                if (a.mu0 != 0.0) { lu = su < kBig ? mu_bar / su : 0.0; ll = sl < kBig ? mu_bar / sl : 0.0; }
                else { lu = su < kBig ? 1.0 : 0.0; ll = sl < kBig ? 1.0 : 0.0; }
                if (su < kBig) mu_acc = fma(su, lu, mu_acc);
                if (sl < kBig) mu_acc = fma(sl, ll, mu_acc);
                const double w[4] = {lu, ll, v[TM_ZY] + v[TM_ZN], 0.0};
                tm_st<4>(tcol(k, TM_LU), w);
            }
            tm_wait_st();
        }
        {
            double d1 = 0.0, d3 = 0.0;
            warp_reduce4(rp_inf, d1, mu_acc, d3);
        }
        double mu_c = n_box > 0.0 ? mu_acc / n_box : 0.0;
        double rp_scale = 1.0;
        __syncwarp();

        // operand fragments come from the sweep image too: the instance's hot set in L2 is the image alone (with A_k, B_k read from their
        // own arrays next to it, the resident instances' 114 MB no longer fit and phase A waited on HBM: +30 %)
        ZFrag<NX, NU, NT, KS, true, D::SLOT> zsrc;
        zsrc.init(Wg, Wg + D::OFFB, rp, lane);
        auto load_z = [&](const int k, double (&zf)[KS][NT]) { zsrc.load(k, Wg, zf); };
        // sweep ring: stage k -> slot k % NSLOT, one bulk copy of the stage image issued by one lane
        // (the slots are contiguous with one stride and the sweeps carry the slot index along instead of computing k % NSLOT: the
        // uniform-datapath address chain of the issue was a visible part of a stage)
        auto slot_ptr = [&](const int si) -> double* { return ring_lo + si * T::SLOTS; };
        auto slot_of = [&](const int k) -> int { return T::NSLOT == 5 ? k - 5 * ((k * 13) >> 6) : k % T::NSLOT; };      // k < 64
        auto slot_inc = [&](const int si) -> int { return si + 1 == T::NSLOT ? 0 : si + 1; };
        auto slot_dec = [&](const int si) -> int { return si == 0 ? T::NSLOT - 1 : si - 1; };
#ifdef RNMPC_RING_BULK       // one bulk copy per slot issued by an elected lane, completion on an mbarrier (measured: 1.7 % slower, below)
        auto ring_issue = [&](const int k, const int si) {        // all lanes, uniform operands; one elected lane issues (bulk_g2s_elect)
            if (k >= 0 && k < N)
                bulk_g2s_elect(slot_ptr(si), Wg + (size_t)k * D::SLOT, (unsigned)(D::SLOT * sizeof(double)), rbar + si);
        };
        auto ring_wait = [&](const int si) -> const double* {
            mbar_wait(rbar + si, (ring_par >> si) & 1u);
            ring_par ^= 1u << si;
            return slot_ptr(si);
        };
#else
        // default: the stage image as 16-byte cp.async of all lanes (7 per lane for the rocket), one commit group per stage (empty when
        // out of range, so the group count of the wait is a constant), completion by cp.async.wait_group.  Against the bulk copy +
        // mbarrier version: QP launch 2.670 -> 2.624 ms -- with one copy per slot the TMA issue was no longer the cost, the mbarrier round
        // trip of every wait (~125 cycles per stage) was.
        auto ring_issue = [&](const int k, const int si) {
            if (k >= 0 && k < N) {
                const double* src = Wg + (size_t)k * D::SLOT;
                double* dst = slot_ptr(si);
#pragma unroll
                for (int c0 = 0; c0 < D::SLOT / 2; c0 += 32) {
                    const int c = c0 + lane;
                    if (c < D::SLOT / 2) cpa16(dst + 2 * c, src + 2 * c);
                }
            }
            cpa_commit();
        };
        auto ring_wait = [&](const int si) -> const double* {
            cpa_wait<T::NSLOT - 1>();
            return slot_ptr(si);
        };
#endif
        double sigma_mu = 0.0;
        // corrector right-hand side of this lane's entry of a stage from its TMEM columns (rnmpc_qp.cuh rhs_eval, pass 1)
        auto rhs_corr = [&](const int k) -> double {
            double v[4], w[2];
            tm_ld<4>(tcol(k, TM_SU), v);
            tm_ld<2>(tcol(k, TM_ZY), w);
            const double su = v[0], sl = v[1], lu = v[2], ll = v[3], dza = w[1];
            // branch-free (a lane without a bound has reciprocal slack 0 and multiplier 0): the two reciprocals run side by side
            const double iu = rc(su), il = rc(sl);
            double gq = qd_of(k) * w[0];
            gq = fma(sigma_mu + dza * (lu * iu * dza - lu), iu, gq);
            gq = fma(-(sigma_mu + dza * (ll * il * dza + ll)), il, gq);
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
            __syncwarp();
            // ================================================================ factorisation sweep + predictor backward recursion
            zero_buf(buf1, 8 * MTS * LDS);
            __syncwarp();
            {
                double v[4], w[1];
                tm_ld<4>(tcol(N, TM_SU), v);
                tm_ld<1>(tcol(N, TM_ZY), w);
                if (lane < NX) {
                    buf1[lane * LDS + lane] = qd_t + (v[2] * rc(v[0]) + v[3] * rc(v[1]));
                    pv[lane] = qd_t * w[0];
                }
            }
            {
                double zf[KS][NT];
                load_z(N - 1, zf);
                double rhs_n = 0.0, dgn_n = 0.0;
                auto stage_scalars = [&](const int k, double& r_out, double& d_out) {
                    double v[4], w[1];
                    tm_ld<4>(tcol(k, TM_SU), v);
                    tm_ld<1>(tcol(k, TM_ZY), w);
                    r_out = (lane < NX && k == 0) ? 0.0 : qd_s * w[0];
                    d_out = (lane < NX ? qd_s : 0.0) + (v[2] * rc(v[0]) + v[3] * rc(v[1]));   // R2 is added to H below
                };
                stage_scalars(N - 1, rhs_n, dgn_n);
                __syncwarp();
                mark(1);
#pragma unroll 1
                for (int k = N - 1; k >= 0; --k) {
                    // gradient and barrier weight of this lane's entry (lanes >= NZ hold a neutral entry): the values of THIS stage were
                    // computed during the previous one (two TMEM loads and two reciprocals off the critical path); phase C is the first
                    // reader, behind the barrier that follows phase A
                    if (lane < NZ) { rhs[lane] = rhs_n; dgn[lane] = dgn_n; }
                    if (k > 0) stage_scalars(k - 1, rhs_n, dgn_n);
                    // ---- phase A: W = S [A B rp].  When the last row tile of S holds ONE row (NX = 8 j + 1: the rocket's 17), that row of W is
                    //      15 DFMA on the operand fragments already in registers plus two shuffle steps instead of 15 DMMA that are 1/8
                    //      useful: the tensor pipe is what phases A and B wait on (two warps per scheduler share it).
                    constexpr bool ROW1 = (NX - 8 * (MTS - 1)) == 1;
                    if (ROW1) {
                        constexpr int R = 8 * (MTS - 1);
                        double sr[KS];
#pragma unroll
                        for (int s = 0; s < KS; ++s) sr[s] = buf1[R * LDS + 4 * s + t4];          // S[R][4s + t]
#pragma unroll
                        for (int n = 0; n < NT; ++n) {
                            double acc = 0.0;
#pragma unroll
                            for (int s = 0; s < KS; ++s) acc = fma(sr[s], zf[s][n], acc);           // rows 4s + t of column 8n + g
                            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                            if (t4 == 0) buf2[R * LDW + 8 * n + g] = acc;
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < (ROW1 ? MTS - 1 : MTS); ++mt) {
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
                    // operand fragments of the NEXT stage straight into zf (its last use in this stage was the product above): the L2 round
                    // trip is covered by the Cholesky phase and the S update, no second fragment buffer and no copy
                    if (k > 0) load_z(k - 1, zf);
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
                        // one triangular solve per lane: lanes < NX a column of (A'SB)' -> column of K, lane NX the vector t -> d
                        double f[NU];
                        if (lane < NX) {
#pragma unroll
                            for (int p = 0; p < NU; ++p) f[p] = buf1[lane * LDW + CB + p];
                        } else {
#pragma unroll
                            for (int p = 0; p < NU; ++p) f[p] = rhs[NX + p] + buf1[(CB + p) * LDW + CP] + rp_scale * buf1[(CB + p) * LDW + CRP];
                            if (lane == NX) {
#pragma unroll
                                for (int p = 0; p < NU; ++p) uc[p] = f[p];
                            }
                        }
                        chol_solve<NU>(L, f);
                        if (lane < NX) {
#pragma unroll
                            for (int p = 0; p < NU; ++p) { kk[p * LDA + lane] = -f[p]; Wg[(size_t)k * D::SLOT + D::OFFK + p * LDA + lane] = -f[p]; }
                        } else if (lane == NX) {
#pragma unroll
                            for (int p = 0; p < NU; ++p) dff[k * NU + p] = -f[p];
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
                        // ROW1: the last column tile of S holds ONE column -- its NX entries S_k[r][R] are NU DFMA each on lane r instead of MTS
                        // tile updates on the tensor pipe with their predicated, mirrored stores
                        constexpr int MTE = ROW1 ? MTS - 1 : MTS;
                        double scol = 0.0;
                        if (ROW1 && lane < NX) {
                            constexpr int R = 8 * (MTS - 1);
                            scol = buf1[lane * LDW + R] + (lane == R ? dgn[R] : 0.0);          // M[r][R] (upper triangle: r <= R)
#pragma unroll
                            for (int p = 0; p < NU; ++p) scol = fma(buf1[lane * LDW + CB + p], kk[p * LDA + R], scol);
                        }
                        double kf[MTS], mc0[MTS * (MTS + 1) / 2], mc1[MTS * (MTS + 1) / 2], gf[MTS];
#pragma unroll
                        for (int n = 0; n < MTE; ++n) kf[n] = (t4 < NU && 8 * n + g < NX) ? kk[t4 * LDA + 8 * n + g] : 0.0;      // B operand K[t][8n + g]
                        int ti = 0;
#pragma unroll
                        for (int mt = 0; mt < MTE; ++mt) {
                            gf[mt] = (t4 < NU) ? buf1[(8 * mt + g) * LDW + CB + t4] : 0.0;                                        // A operand (A'SB)[r][t]
#pragma unroll
                            for (int n = mt; n < MTE; ++n, ++ti) {
                                const double2 mv = *reinterpret_cast<const double2*>(buf1 + (8 * mt + g) * LDW + 8 * n + 2 * t4);
                                mc0[ti] = mv.x; mc1[ti] = mv.y;
                            }
                        }
                        // (no zeroing: the pads of the S layout may hold finite leftovers of the M tiles -- they only ever multiply the zero
                        // rows of Z_k; the buffers are cleared once per instance, so nothing non-finite can be left in them)
                        __syncwarp();
                        if (ROW1 && lane < NX) {
                            constexpr int R = 8 * (MTS - 1);
                            buf1[lane * LDS + R] = scol; buf1[R * LDS + lane] = scol;
                        }
                        ti = 0;
#pragma unroll
                        for (int mt = 0; mt < MTE; ++mt) {
                            const int r = 8 * mt + g;
#pragma unroll
                            for (int n = mt; n < MTE; ++n, ++ti) {
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
                    }
                    __syncwarp();
                    mark(5);
                }
            }
#ifdef RNMPC_RING_BULK
            fence_block();          // L_k written above is read back by other lanes
            fence_proxy_async();    // K_k (global) and the tile buffers (shared) written above are next touched by bulk copies
#endif
            // (cp.async ring: the readers of K_k, L_k are lanes of this warp behind the warp barrier below, which orders memory among them)
            __syncwarp();

            // ================================================================ predictor (pass 0) and corrector (pass 1): one copy of the sweep code
            const double tau_c = mu_c > 1e-8 ? 0.995 : 0.9999;
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int arr_dz = pass ? TM_DC : TM_DA;
                if (pass == 1) {
#pragma unroll
                    for (int j = 1; j < T::NSLOT; ++j) ring_issue(N - j, slot_of(N - j));
                    mark(6);
                    {
                        // backward recursion: g_k = p_{k+1} + rp_scale Sc_k ; lanes < NX: p_k = rx_k + A_k' g_k + K_k' t_k,
                        // lanes NX .. NZ-1: t_k = ru_k + B_k' g_k (kept in tu for the feed-forward); rx, ru from the TMEM columns of the stage
                        const int r = lane < NX ? lane : NX - 1;
                        const int p = (lane >= NX && lane < NZ) ? lane - NX : 0;
                        const int bw_off = lane < NX ? lane : (lane < NZ ? D::OFFB + (lane - NX) : 0);
                        const int bw_stride = lane < NX ? LDA : NU;
                        double pl = rhs_corr(N);
                        int si = slot_of(N - 1);
#pragma unroll 1
                        for (int k = N - 1; k >= 0; --k) {
                            ring_issue(k - (T::NSLOT - 1), slot_inc(si));    // into the slot of stage k + 1 (finished: barrier at the end of the previous stage)
                            if (lane < NX) gv[lane] = pl + rp_scale * Sc[k * NX + lane];
                            const double rk = rhs_corr(k);
                            const double* sl = ring_wait(si);
                            si = slot_dec(si);
                            __syncwarp();
                            // one code path for all lanes: column r of A_k (lanes < NX) or column p of B_k (lanes NX .. NZ-1), own stride
                            double a0 = rk, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                            {
                                const double* col = sl + bw_off;
#pragma unroll
                                for (int m = 0; m + 3 < NX; m += 4) {
                                    a0 = fma(col[m * bw_stride], gv[m], a0); a1 = fma(col[(m + 1) * bw_stride], gv[m + 1], a1);
                                    a2 = fma(col[(m + 2) * bw_stride], gv[m + 2], a2); a3 = fma(col[(m + 3) * bw_stride], gv[m + 3], a3);
                                }
#pragma unroll
                                for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(col[m * bw_stride], gv[m], a0);
                            }
                            if (lane >= NX && lane < NZ) tu[k * NU + p] = (a0 + a1) + (a2 + a3);
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
#pragma unroll
                    for (int j = 0; j < T::NSLOT - 1; ++j) ring_issue(j, j);   // the forward recursion's first stages are in flight during the small solves
                    // ---- d_k = -H_k^{-1} t_k   (one lane per stage)
                    for (int k0 = 0; k0 < N; k0 += 32) {
                        const int k = k0 + lane;
                        if (k >= N) continue;
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
#pragma unroll
                    for (int j = 0; j < T::NSLOT - 1; ++j) ring_issue(j, j);
                }
                mark(9);
                // ---- forward recursion: lanes NX .. NZ-1: du_k = K_k dx_k + d_k ; lanes < NX: dx_{k+1} = A_k dx_k + B_k du_k +
                //      rp_scale rp_k, dx_0 = 0.  Lane i keeps its own entry of the step (dx_k[i] or du_k[i - NX]), stores it into the stage's
                //      TMEM column and runs the ratio test of that entry (rnmpc_qp.cuh ratio_entry) while the next ring slot is in flight.
                double rpm = 0.0, rdm = 0.0, S3 = 0.0, S4 = 0.0;
                // The ratio test of stage k - 1 is software-pipelined into stage k: its TMEM loads are issued at the head of the stage and its
                // arithmetic (four reciprocals) sits in the same basic block as the stage's matrix-vector product, so the two independent
                // dependency chains interleave instead of following each other.
                // (the forward recursion is instantiated per pass: the predictor carries no corrector terms, the corrector no predictor sums --
                //  with the pass as a run-time value both sets of instructions were issued in both)
                auto forward = [&](auto PASSC) {
                constexpr int PS = decltype(PASSC)::value;
                const int arr_dz = PS ? TM_DC : TM_DA;
                double rt_v[4], rt_w[1];
                auto ratio_load = [&](const int k, const double dz) {        // the warp is converged here (directly behind a __syncwarp)
                    tm_ld<4, false>(tcol(k, TM_SU), rt_v);
                    rt_w[0] = 0.0;
                    if (PS == 1) tm_ld<1, false>(tcol(k, TM_DA), rt_w);
                    const double dzs[1] = {dz};
                    tm_st<1, false>(tcol(k, arr_dz), dzs);
                };
                auto ratio_math = [&](const double dz) {
                    const double su = rt_v[0], sl = rt_v[1], lu = rt_v[2], ll = rt_v[3];
                    const double dza = rt_w[0];
                    const double smu = PS == 1 ? sigma_mu : 0.0;
                    // branch-free: a lane without a bound has reciprocal slack 0 and multiplier 0 (every term vanishes); the four
                    // reciprocals are independent and run side by side
                    // (these reciprocals only steer the iteration -- step lengths with a 0.5 % safety factor, the complementarity estimate of
                    // the predictor -- so the seed with one Newton step / the bare seed is enough: 12 DFMA less per entry and pass)
                    const double iu = su < kBig ? rough_rcp<1>(su) : 0.0, il = sl < kBig ? rough_rcp<1>(sl) : 0.0;
                    const double ilu = rough_rcp<0>(lu), ill = rough_rcp<0>(ll);  // inf for a zero multiplier: discarded by the selects below
                    const double ru = lu * iu, rl = ll * il;
                    const double tgu = PS == 1 ? smu + dza * (ru * dza - lu) : 0.0;
                    const double tgl = PS == 1 ? smu + dza * (rl * dza + ll) : 0.0;
                    const double dlu = tgu * iu - lu + ru * dz;
                    const double dll = tgl * il - ll - rl * dz;
                    rpm = fmax(rpm, fmax(dz * iu, -dz * il));
                    rdm = fmax(rdm, fmax(dlu < 0.0 ? -dlu * ilu : 0.0, dll < 0.0 ? -dll * ill : 0.0));
                    if (PS == 0) { S3 = fma(-dz, lu, S3); S4 = fma(-dz, dlu, S4); S3 = fma(dz, ll, S3); S4 = fma(dz, dll, S4); }
                };
                {
                    const int r = lane < NX ? lane : NX - 1;
                    const int fw_off = lane < NX ? lane * LDA : (lane < NZ ? D::OFFK + (lane - NX) * LDA : 0);
                    const double* fw_init = lane < NX ? rp + lane : (lane < NZ ? dff + (lane - NX) : rp);
                    const int fw_istride = lane < NX ? NX : (lane < NZ ? NU : 0);
                    const double fw_scale = lane < NX ? rp_scale : 1.0;
                    double dxk = 0.0;                          // lanes < NX: dx_k[lane]
                    double dz_prev = 0.0;                      // this lane's entry of the step of stage k - 1
                    if (lane < NZ) { xs0[lane] = 0.0; xs1[lane] = 0.0; }
                    int si = 0;
#pragma unroll 1
                    for (int k = 0; k < N; ++k) {
                        double* xk = (k & 1) ? xs1 : xs0;      // [dx_k ; du_k]
                        double* xn = (k & 1) ? xs0 : xs1;
                        ring_issue(k + (T::NSLOT - 1), slot_dec(si));        // into the slot of stage k - 1
                        const double* sl = ring_wait(si);
                        si = slot_inc(si);
                        __syncwarp();
                        if (k > 0) ratio_load(k - 1, dz_prev);
                        // one code path for all lanes: row r of A_k from rp_scale rp_k (lanes < NX) or row p of K_k from d_k (lanes NX .. NZ-1);
                        // dx_0 = 0, so stage 0 needs no special case
                        double a0 = fw_scale * fw_init[k * fw_istride], a1 = 0.0, a2 = 0.0, a3 = 0.0;
                        {
                            const double* row = sl + fw_off;
#pragma unroll
                            for (int m = 0; m + 3 < NX; m += 4) {
                                const double2 r01 = *reinterpret_cast<const double2*>(row + m), r23 = *reinterpret_cast<const double2*>(row + m + 2);
                                a0 = fma(r01.x, xk[m], a0); a1 = fma(r01.y, xk[m + 1], a1);
                                a2 = fma(r23.x, xk[m + 2], a2); a3 = fma(r23.y, xk[m + 3], a3);
                            }
#pragma unroll
                            for (int m = NX / 4 * 4; m < NX; ++m) a0 = fma(row[m], xk[m], a0);
                        }
                        if (k > 0) ratio_math(dz_prev);
                        if (lane >= NX && lane < NZ) xk[lane] = (a0 + a1) + (a2 + a3);
                        __syncwarp();
                        double dz = dxk;                        // this lane's entry of stage k
                        if (lane < NX) {
                            double acc = (a0 + a1) + (a2 + a3);
                            const double* Br = sl + D::OFFB + r * NU;
#pragma unroll
                            for (int q = 0; q < NU; ++q) acc = fma(Br[q], xk[NX + q], acc);
                            xn[lane] = acc;
                            dxk = acc;
                        } else dz = (a0 + a1) + (a2 + a3);
                        dz_prev = lane < NZ ? dz : 0.0;
                        __syncwarp();                           // dx_{k+1} visible; every lane is done with the slot of stage k
                    }
                    if (N > 0) { ratio_load(N - 1, dz_prev); ratio_math(dz_prev); }
                    __syncwarp();
                    dz_prev = lane < NX ? dxk : 0.0;
                    ratio_load(N, dz_prev); ratio_math(dz_prev);
                    tm_wait_st();
                }
                };
                if (pass == 0) forward(std::integral_constant<int, 0>{}); else forward(std::integral_constant<int, 1>{});
                mark(10);
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
#pragma unroll 2
                for (int k = 0; k <= N; ++k) {
                    double v[8];
                    tm_ld<8>(tcol(k, 0), v);
                    double su = v[TM_SU], sl = v[TM_SL], lu = v[TM_LU], ll = v[TM_LL];
                    if (upd) {
                        const double dza = v[TM_DA], dz = skip_corr ? v[TM_DA] : v[TM_DC];
                        // branch-free (no bound: slack 1e300 stays, reciprocal 0, multiplier 0)
                        const double iu = rc(su), il = rc(sl);
                        const double ru = lu * iu, rl = ll * il;
                        const double dlu = so * (sigma_mu + dza * (ru * dza - lu)) * iu - lu + ru * dz;
                        const double dll = so * (sigma_mu + dza * (rl * dza + ll)) * il - ll - rl * dz;
                        if (su < kBig) { su -= ap * dz; lu += ad * dlu; }
                        if (sl < kBig) { sl += ap * dz; ll += ad * dll; }
                        const double w[4] = {su, sl, lu, ll};
                        tm_st<4>(tcol(k, 0), w);
                        const double zy[1] = {v[TM_ZY] + ap * dz};
                        tm_st<1>(tcol(k, TM_ZY), zy);
                    }
                    if (su < kBig) acc = fma(su, lu, acc);
                    if (sl < kBig) acc = fma(sl, ll, acc);
                }
                tm_wait_st();
                if (upd) rp_scale *= (1.0 - ap);
            }
            {
                double d0 = 0.0, d1 = 0.0, d3 = 0.0;
                warp_reduce4(d0, d1, acc, d3);
            }
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
        __syncwarp();

        // ---- outputs in the reference layouts + objective x'Hx + q'x (qp_jit.py:401, P = 2H, q = 2 H y_nom; diagonal H)
        {
            double* Xo = a.dX + (size_t)b * (N + 1) * NX;
            double* Uo = a.dU + (size_t)b * N * NU;
            double* mu = a.mu + (size_t)b * N * NI;
            double* muf = a.mu_f + (size_t)b * NIF;
            double acc = 0.0;
#pragma unroll 1
            for (int k = 0; k <= N; ++k) {
                double v[4], zq[1];
                tm_ld<4>(tcol(k, TM_LU), v);           // lu, ll, zy, (da)
                tm_ld<1>(tcol(k, TM_ZN), zq);
                const int i = lane;
                if (i < NX || (i < NZ && k < N)) {
                    const double zn = zq[0];
                    const double z = v[2] - zn;
                    acc = fma(z, 0.5 * qd_of(k) * (z + 2.0 * zn), acc);
                    if (i < NX) Xo[k * NX + i] = z; else Uo[k * NU + i - NX] = z;
                    if (k < N) { mu[k * NI + i] = v[0]; mu[k * NI + NZ + i] = v[1]; }
                    else { muf[i] = v[0]; muf[NX + i] = v[1]; }
                }
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
        __syncwarp();
    }
    // ---- every warp of the CTA is done: free the tensor memory
    tmem_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc_256(tm_cta);
}

}  // namespace rnmpc
