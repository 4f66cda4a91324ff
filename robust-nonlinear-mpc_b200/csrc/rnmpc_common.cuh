// Shared device infrastructure: problem constants, the "slot" work decomposition and the register-blocked
// small-matrix primitives used by the fast-SLS and QP kernels.
//
// Work decomposition ("slot"): every sequential Riccati chain (one disturbance column pair of one MPC instance in
// the fast-SLS kernel, one whole instance in the QP kernel) is executed by LPS = ceil(NX/2) lanes of a warp; lane
// `sl` of a slot owns rows 2*sl and 2*sl+1 of every NX-row matrix of that chain in registers.  SPW = 32/LPS slots
// share a warp and run in lock-step; the only synchronisation is __syncwarp.  Small-matrix products are
//   out[my rows][:] = sum_m  in[my rows][m] * M[m][:]
// with `in` in registers and M broadcast from shared memory as 128-bit loads: 4 DFMA per LDS.128, which the
// microbenchmark in tools/fp64_microbench.cu measured at 83 % of the DFMA peak on B200 (1 row per lane: 54 %).
#pragma once
#ifdef RNMPC_EMU
#include "simt_emu.h"      // tools/simt_emu: host-side SIMT emulator (test infrastructure; never defined in the product build)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

namespace rnmpc {

constexpr double kEpsBackoff = 1e-10;   // solver/fast_SLS_jit.py:205
constexpr double kEpsQP = 1e-10;        // solver/qp_jit.py:19
// per-instance flag bits shared by the kernels
constexpr int FLAG_FAIL = 1;    // a QP of this solve failed: the instance is frozen (reference: success=False)
constexpr int FLAG_Q3 = 2;      // check_convergence_socp returned True: skip backward solve + tightening (Q3)
constexpr int FLAG_TIGHT = 4;   // tube back-offs were computed in this fast-SLS solve and are in the QP bounds
constexpr int FLAG_PREV = 8;    // _prev_primal_vec exists (Q3 state)
constexpr int FLAG_SLSDONE = 32; // converge-to-tolerance mode: this instance's fast-SLS iteration has terminated (fast_SLS_jit.py:305)
constexpr int FLAG_KVALID = 16; // K / Phi were computed since the last reset_solver_to_zeros (else the reference returns None)

constexpr double kX0Tol = 1e-3;         // tolerated violation of the (redundant) stage-0 state box by the measured state
constexpr double kBig = 1e19;           // bounds >= kBig are treated as infinite (qp_jit.py:382 uses 1e20)

template <int NX, int NU>
struct Params {
    double Q[NX * NX], R[NU * NU], Qf[NX * NX];           // nominal cost (OSQP sees P = 2*blkdiag)
    double Q2[NX * NX], R2[NU * NU];                      // 2Q, 2R (Hessian blocks of the QP)
    double Qreg[NX * NX], Rreg[NU * NU], Qregf[NX * NX];  // tube regularisers
    double E[NX * NX];                                    // disturbance map (nw == nx), same for every column (Q8)
    double xub[NX], xlb[NX], uub[NU], ulb[NU];            // stage box  g = [xub, uub, -xlb, -ulb]
    double xubf[NX], xlbf[NX];                            // terminal box gf = [xubf, -xlbf]
    double h;                                             // RK4 step (0.05, Q6)
};

template <int NX>
struct Slot {
    static constexpr int LPS = (NX + 1) / 2;   // lanes per slot
    static constexpr int SPW = 32 / LPS;       // slots per warp
    static constexpr int LD = (NX + 1) & ~1;   // even leading dimension of shared-memory matrices
    static constexpr int NC2 = LD / 2;         // double2 per row
};

__host__ __device__ constexpr int even_up(int n) { return (n + 1) & ~1; }

// acc{0,1}[c] += sum_m in(m).{x,y} * Mat[m*LD + c]   for c in [0, LD)   (Mat in shared memory, 16-byte aligned).
// The m-loop is deliberately NOT unrolled: one iteration is LD/2 LDS.128 + 2*LD DFMA (~50 instructions), so the hot loops
// of a Riccati step stay resident in the instruction cache (the fully unrolled version was 136 KB of SASS per kernel and
// ncu attributed 27 % of the stall samples to instruction fetch).
template <int M, int LD, class InFn>
__device__ __forceinline__ void rowmm2(InFn&& in, const double* __restrict__ Mat, double (&acc0)[LD], double (&acc1)[LD]) {
    // two-stage software pipeline: the LDS of row m+1 are in flight while the 2*LD DFMA of row m issue (ncu: without it the
    // DFMAs of this loop sat on the short scoreboard waiting for their own iteration's loads)
    constexpr int H = LD / 2;
    double2 va[H], vb[H];
    double2 ia = in(0), ib;
#pragma unroll
    for (int c = 0; c < H; ++c) va[c] = *reinterpret_cast<const double2*>(Mat + 2 * c);
#pragma unroll 1
    for (int m = 0; m + 1 < M; m += 2) {
        ib = in(m + 1);
#pragma unroll
        for (int c = 0; c < H; ++c) vb[c] = *reinterpret_cast<const double2*>(Mat + (m + 1) * LD + 2 * c);
#pragma unroll
        for (int c = 0; c < H; ++c) {
            acc0[2 * c] = fma(ia.x, va[c].x, acc0[2 * c]);
            acc0[2 * c + 1] = fma(ia.x, va[c].y, acc0[2 * c + 1]);
            acc1[2 * c] = fma(ia.y, va[c].x, acc1[2 * c]);
            acc1[2 * c + 1] = fma(ia.y, va[c].y, acc1[2 * c + 1]);
        }
        if (m + 2 < M) {
            ia = in(m + 2);
#pragma unroll
            for (int c = 0; c < H; ++c) va[c] = *reinterpret_cast<const double2*>(Mat + (m + 2) * LD + 2 * c);
        }
#pragma unroll
        for (int c = 0; c < H; ++c) {
            acc0[2 * c] = fma(ib.x, vb[c].x, acc0[2 * c]);
            acc0[2 * c + 1] = fma(ib.x, vb[c].y, acc0[2 * c + 1]);
            acc1[2 * c] = fma(ib.y, vb[c].x, acc1[2 * c]);
            acc1[2 * c + 1] = fma(ib.y, vb[c].y, acc1[2 * c + 1]);
        }
    }
    if (M & 1) {
#pragma unroll
        for (int c = 0; c < H; ++c) {
            acc0[2 * c] = fma(ia.x, va[c].x, acc0[2 * c]);
            acc0[2 * c + 1] = fma(ia.x, va[c].y, acc0[2 * c + 1]);
            acc1[2 * c] = fma(ia.y, va[c].x, acc1[2 * c]);
            acc1[2 * c + 1] = fma(ia.y, va[c].y, acc1[2 * c + 1]);
        }
    }
}

// Reciprocal to ~1 ulp without the IEEE division routine (~30 dependent instructions and a slow-path call): hardware
// seed (MUFU.RCP64H, 20+ bits) and two Newton steps.  Used where the quotient only scales a step or a barrier weight.
#ifndef RNMPC_EMU
__device__ __forceinline__ double fast_rcp(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
#endif

// Reciprocal for quantities that only steer the iteration (ratio tests, the predictor's complementarity estimate): seed + NEWTON
// Newton steps (0: 20+ bits, 1: 40+ bits).
#ifndef RNMPC_EMU
template <int NEWTON>
__device__ __forceinline__ double rough_rcp(const double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    if (NEWTON >= 1) { const double e = fma(-x, r, 1.0); r = fma(r, e, r); }
    return r;
}
#endif

// Cholesky factor of an SPD NU x NU matrix held in registers (lower triangle, in place; the diagonal is stored
// as its reciprocal).  Returns false if a pivot is not positive.
template <int NU>
__device__ __forceinline__ bool chol_inplace(double (&H)[NU][NU]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        double d = H[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-H[j][k], H[j][k], d);
        ok = ok && (d > 0.0);
        const double il = rsqrt(d);
        H[j][j] = il;
#pragma unroll
        for (int i = j + 1; i < NU; ++i) {
            double s = H[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s = fma(-H[i][k], H[j][k], s);
            H[i][j] = s * il;
        }
    }
    return ok;
}

// Solve (L L^T) z = f with the factor from chol_inplace.
template <int NU>
__device__ __forceinline__ void chol_solve(const double (&L)[NU][NU], double (&f)[NU]) {
#pragma unroll
    for (int i = 0; i < NU; ++i) {
        double s = f[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(-L[i][k], f[k], s);
        f[i] = s * L[i][i];
    }
#pragma unroll
    for (int i = NU - 1; i >= 0; --i) {
        double s = f[i];
#pragma unroll
        for (int k = i + 1; k < NU; ++k) s = fma(-L[k][i], f[k], s);
        f[i] = s * L[i][i];
    }
}

// bit views of a double (TMEM moves 32-bit columns)
#ifndef RNMPC_EMU
__device__ __forceinline__ unsigned double_lo(const double v) { return (unsigned)__double2loint(v); }
__device__ __forceinline__ unsigned double_hi(const double v) { return (unsigned)__double2hiint(v); }
__device__ __forceinline__ double hilo2double(const unsigned hi, const unsigned lo) { return __hiloint2double((int)hi, (int)lo); }
#endif

// 16-byte asynchronous global -> shared copies (LDGSTS) in a form the emulated build understands too (tools/simt_emu defers
// the copy until the matching wait, so a read before the wait / barrier shows up as a wrong result on the CPU already).
#ifndef RNMPC_EMU
__device__ __forceinline__ void cpa16(double* dst_smem, const double* __restrict__ src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src));
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;"); }
template <int PENDING>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void fence_block() { __threadfence_block(); }
// 1-D bulk copies global -> shared on the TMA unit (cp.async.bulk), completion counted in bytes on an mbarrier in shared memory:
// one lane issues a whole stage matrix with one instruction (bytes and addresses multiples of 16).
__device__ __forceinline__ void mbar_init(unsigned long long* bar, const int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, const unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(double* dst_smem, const double* __restrict__ src, const unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)),
                 "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
                 : "memory");
}
// Arm the barrier and issue the copy from ONE elected lane without a branch: both instructions are predicated on elect.sync's result.
// Called by all lanes with warp-uniform operands (see uniform_u32): the operands then sit in uniform registers and the copy is a
// handful of instructions.  (An `if (lane == 0)` around bulk_g2s with operands the compiler cannot prove uniform becomes a loop
// over the active lanes with four R2UR.BROADCAST per copy: ~200 cycles per call on the issuing warp, measured in qp_tm_kernel.)
__device__ __forceinline__ void bulk_g2s_elect(double* dst_smem, const double* __restrict__ src, const unsigned bytes, unsigned long long* bar) {
    asm volatile(
        "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n"
        "@p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n}"
        ::"r"((unsigned)__cvta_generic_to_shared(dst_smem)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}
// A value that is the same in every lane, in a form the compiler knows to be warp-uniform (a vote result lives in a uniform register):
// lane i contributes bit i.
__device__ __forceinline__ unsigned uniform_u32(const unsigned v) { return __ballot_sync(0xffffffffu, (v >> (threadIdx.x & 31)) & 1u); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, const unsigned parity) {
    const unsigned addr = (unsigned)__cvta_generic_to_shared(bar);
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    }
}
// orders earlier generic-proxy accesses (st.global / st.shared of this thread) before later asynchronous-proxy accesses (bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Tensor memory (TMEM, 256 KB per SM) as lane-private, column-indexed storage for fp64 state: thread i of warp w of a CTA owns TMEM lane
// 32 (w % 4) + i; a tcgen05.ld / st with the 32x32b shape moves N consecutive 32-bit columns of every thread's own lane (a double =
// 2 columns), at a warp-uniform, run-time column address -- registers that can be indexed.  tools/tmem_probe.cu checks this on the
// device (0 mismatches over 1184 CTAs, 40 cycles for a dependent store -> load round trip).  One warp of the CTA allocates and frees.
__device__ __forceinline__ void tmem_alloc_256(unsigned* slot_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((unsigned)__cvta_generic_to_shared(slot_smem)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_256(const unsigned base) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base) : "memory"); }
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ unsigned tmem_warp_base(const unsigned base, const int warp_in_cta) { return base + ((unsigned)(32 * (warp_in_cta & 3)) << 16); }
template <int NREG> __device__ __forceinline__ void tm_ld_raw(const unsigned taddr, unsigned (&r)[NREG]);      // load + wait::ld
template <int NREG> __device__ __forceinline__ void tm_st_raw(const unsigned taddr, const unsigned (&r)[NREG]);
template <> __device__ __forceinline__ void tm_ld_raw<2>(const unsigned taddr, unsigned (&r)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tm_ld_raw<4>(const unsigned taddr, unsigned (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tm_ld_raw<8>(const unsigned taddr, unsigned (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tm_ld_raw<16>(const unsigned taddr, unsigned (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\ttcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
template <> __device__ __forceinline__ void tm_st_raw<2>(const unsigned taddr, const unsigned (&r)[2]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(r[0]), "r"(r[1]) : "memory");
}
template <> __device__ __forceinline__ void tm_st_raw<4>(const unsigned taddr, const unsigned (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}
template <> __device__ __forceinline__ void tm_st_raw<8>(const unsigned taddr, const unsigned (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
template <> __device__ __forceinline__ void tm_st_raw<16>(const unsigned taddr, const unsigned (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { __syncwarp(); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#endif
// ND doubles of this thread's TMEM lane from / to 2 ND consecutive columns (both builds: tools/simt_emu provides tm_*_raw too)
// (the tcgen05 instructions are warp-collective and "aligned": the lanes re-converge first -- the compiler does not know that they must)
// SYNC = false: the caller knows the warp is converged (directly behind a __syncwarp)
template <int ND, bool SYNC = true>
__device__ __forceinline__ void tm_ld(const unsigned taddr, double (&v)[ND]) {
    unsigned r[2 * ND];
    if (SYNC) __syncwarp();
    tm_ld_raw<2 * ND>(taddr, r);
#pragma unroll
    for (int i = 0; i < ND; ++i) v[i] = hilo2double(r[2 * i + 1], r[2 * i]);
}
template <int ND, bool SYNC = true>
__device__ __forceinline__ void tm_st(const unsigned taddr, const double (&v)[ND]) {
    unsigned r[2 * ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) { r[2 * i] = double_lo(v[i]); r[2 * i + 1] = double_hi(v[i]); }
    if (SYNC) __syncwarp();
    tm_st_raw<2 * ND>(taddr, r);
}

#ifndef RNMPC_EMU      // the slot machinery below stages with cp.async: not emulated
// Shared-memory working set of one slot (see SlotSmem below for the bank-conflict padding).
template <int NX, int NU>
struct SlotSmemCore {
    static constexpr int LD = Slot<NX>::LD;
    double S[NX * LD];              // current cost-to-go S (backward) / Phi_x (forward), full matrix for broadcast
    double T[NX * LD];              // closed-loop matrix A+BK (backward)
    double Bm[even_up(NX * NU)];    // B_k
    double SB[even_up(NX * NU)];    // S B
    double H[even_up(NU * NU)];     // Cu + B'SB
    double cu[even_up(NU)];         // diagonal additions of the input cost
    union {                         // the fast-SLS forward pass and the QP sweeps never coexist in one kernel
        struct { double K[NU * LD]; double PU[NU * LD]; };                               // K_k^T / Phi_u (fast-SLS)
        struct { double vec[4 * LD]; double red[even_up(2 * Slot<NX>::LPS)]; };          // vector / reduction scratch (QP)
    };
};

// The slots of one warp read the same offset of their own SlotSmem in one LDS (a "broadcast" per slot).  With a slot
// stride that is a multiple of 128 B those 2-3 addresses fall into the same banks and every LDS.128 replays 3x (ncu:
// l1tex__data_bank_conflicts_pipe_lsu_mem_shared ~ 1.2e8 per QP launch).  Padding the stride to 16 (mod 128) bytes puts
// neighbouring slots into different 16-byte bank groups.
template <int NX, int NU>
struct SlotSmem : SlotSmemCore<NX, NU> {
    static constexpr int kCore = (int)sizeof(SlotSmemCore<NX, NU>);
    static constexpr int kPad = ((16 - kCore % 128) + 128) % 128;
    double pad_[kPad / 8 > 0 ? kPad / 8 : 1];
};

// Software prefetch of a global range into L2 (no registers held): issued one stage ahead of use, it turns the DRAM round
// trip at the head of every Riccati / sweep stage into an L2 hit.
template <int LPS>
__device__ __forceinline__ void prefetch_l2(const void* p, const int bytes, const int sl) {
    const char* c = reinterpret_cast<const char*>(p);
    for (int off = sl * 128; off < bytes; off += LPS * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
}

// A_k is stored in global memory in the same LD-strided layout as in shared memory ([NX][LD] doubles, pad column zero; the
// engine keeps its resident A that way and re-packs caller-provided dense matrices), so staging is a contiguous 16-byte copy.
// Asynchronous version (LDGSTS, no registers); completion with cp_async_wait_all() + __syncwarp.
template <int NX, int LD, int LPS>
__device__ __forceinline__ void cp_async_matrix(double* dst, const double* __restrict__ src, const int sl) {
    const unsigned base = (unsigned)__cvta_generic_to_shared(dst);
    constexpr int N16 = NX * LD / 2;
#pragma unroll 4
    for (int e = sl; e < N16; e += LPS)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (unsigned)(e * 16)), "l"(src + 2 * e));
    asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
// wait until at most PENDING of the most recently committed groups of this thread are still in flight (groups complete in order)
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// Stage COUNT contiguous doubles (a dense B_k, K_k or Cholesky factor) into shared memory for a LATER stage: asynchronous
// 16-byte copies when the block is a multiple of 16 bytes (source 16-byte aligned), a plain copy otherwise.  The caller has
// made sure (__syncwarp) that no lane still reads dst; completion with cp_async_wait_all() + __syncwarp.
template <int COUNT, int LPS>
__device__ __forceinline__ void stage_block_async(double* dst, const double* __restrict__ src, const int sl) {
    if constexpr ((COUNT % 2) == 0) {
        const unsigned base = (unsigned)__cvta_generic_to_shared(dst);
#pragma unroll
        for (int e = sl; e < COUNT / 2; e += LPS)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (unsigned)(e * 16)), "l"(src + 2 * e));
    } else {
        for (int e = sl; e < COUNT; e += LPS) dst[e] = __ldg(src + e);
    }
    cp_async_commit();          // always one group per call (possibly empty): callers count groups
}

// Synchronous staging of a padded A_k: all 16-byte loads of a lane in flight together.
template <int NX, int LD, int LPS>
__device__ __forceinline__ void load_padded_smem(double* __restrict__ dst, const double* __restrict__ src, const int sl) {
    constexpr int N16 = NX * LD / 2;
    constexpr int IT = (N16 + LPS - 1) / LPS;
    double2 v[IT];
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int e = sl + j * LPS;
        v[j] = (e < N16) ? __ldg(reinterpret_cast<const double2*>(src) + e) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int e = sl + j * LPS;
        if (e < N16) reinterpret_cast<double2*>(dst)[e] = v[j];
    }
}

// Cooperative load of an [R][C] global matrix into a shared LDS-strided matrix (zero pad columns).  All loads of a
// lane are issued before the first store so that their latencies overlap (one L2 round trip per matrix).
template <int R, int C, int LDS_, int LPS>
__device__ __forceinline__ void load_matrix_smem(double* __restrict__ dst, const double* __restrict__ src, const int sl) {
    constexpr int TOT = R * LDS_;
    constexpr int IT = (TOT + LPS - 1) / LPS;
    double v[IT];
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int e = sl + j * LPS;
        const int i = e / LDS_, c = e - i * LDS_;
        v[j] = (e < TOT && c < C) ? __ldg(src + i * C + c) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int e = sl + j * LPS;
        if (e < TOT) dst[e] = v[j];
    }
}

// One backward Riccati step of a chain (solver/fast_SLS_jit.py:43-62):
//   x = B'S, y = A'S, H = Cu + xB, F = xA, K = -H^{-1}F, S' = Cx + y(A + BK), S' <- (S'+S'^T)/2
// On entry sm.S holds S_{k+1}; on exit it holds S_k.  sm.T is used for A_k and then, in place, for A_k + B_k K_k.
// `pre()` runs while S_{k+1} is intact; `post(kcol0, kcol1, L)` runs once K_k and the Cholesky factor exist, with A_k
// still in sm.T (the QP fuses its vector recursion there).
template <int NX, int NU, class Pre, class Post>
__device__ __forceinline__ bool riccati_step(SlotSmem<NX, NU>& sm, const int sl, const unsigned mask,
                                             const double* __restrict__ Ag, const double* __restrict__ Bg,
                                             const double* __restrict__ Cx0, const double* __restrict__ Cu0,
                                             const double cxd0, const double cxd1, double* __restrict__ Kout, Pre&& pre,
                                             Post&& post, const bool need_S = true, const double* __restrict__ Ag_next = nullptr,
                                             const double* __restrict__ Bg_next = nullptr, const bool T_ready = false,
                                             const double* __restrict__ Ag_async = nullptr, const bool B_ready = false,
                                             const double* __restrict__ Bg_async = nullptr) {
    constexpr int LD = Slot<NX>::LD;
    constexpr int LPS = Slot<NX>::LPS;
    const int r0 = 2 * sl, r1 = 2 * sl + 1;
    const bool v1 = r1 < NX;
    const int r1c = v1 ? r1 : NX - 1;

    // (0) A_k -> sm.T, B_k -> sm.Bm
    // T_ready: A_k was copied into sm.T asynchronously while the previous stage computed
    // B_ready: likewise B_k in sm.Bm (Bg_async: the B to fetch for the next call once this one is done with sm.Bm)
    if (T_ready || B_ready) cp_async_wait_all();
    if (!T_ready) load_padded_smem<NX, LD, LPS>(sm.T, Ag, sl);
    if (!B_ready) load_matrix_smem<NX, NU, NU, LPS>(sm.Bm, Bg, sl);
    if (Ag_next != nullptr) {
        prefetch_l2<LPS>(Ag_next, NX * LD * 8, sl);
        prefetch_l2<LPS>(Bg_next, NX * NU * 8, sl);
    }
    __syncwarp(mask);
    pre();

    // (1) Y = A'S  (my rows r0, r1: columns r0, r1 of A)
    double y0[LD], y1[LD];
#pragma unroll
    for (int c = 0; c < LD; ++c) { y0[c] = 0.0; y1[c] = 0.0; }
    rowmm2<NX, LD>([&](int m) { return *reinterpret_cast<const double2*>(sm.T + m * LD + r0); }, sm.S, y0, y1);

    // (3) SB rows:  (S B)[r, :]
    double f0[NU], f1[NU];
    {
        double sb0[NU], sb1[NU];
#pragma unroll
        for (int p = 0; p < NU; ++p) { sb0[p] = 0.0; sb1[p] = 0.0; }
#pragma unroll 4
        for (int c = 0; c < NX; ++c) {
            const double s0 = sm.S[r0 * LD + c], s1 = sm.S[r1c * LD + c];
#pragma unroll
            for (int p = 0; p < NU; ++p) {
                const double bb = sm.Bm[c * NU + p];
                sb0[p] = fma(s0, bb, sb0[p]);
                sb1[p] = fma(s1, bb, sb1[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NU; ++p) {
            sm.SB[r0 * NU + p] = sb0[p];
            if (v1) sm.SB[r1 * NU + p] = sb1[p];
        }
    }
    __syncwarp(mask);          // every lane is done reading S_{k+1}; SB is complete
    // spill Y (my rows) over the S buffer: the remaining products read it back row-wise
#pragma unroll
    for (int c = 0; c < LD; c += 2) {
        *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(y0[c], y0[c + 1]);
        if (v1) *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(y1[c], y1[c + 1]);
    }
    // H = Cu + B'(SB): entry e of lane sl is e = sl + j LPS; the entries of a lane accumulate side by side (independent chains)
    {
        constexpr int HE = (NU * NU + LPS - 1) / LPS;
        double hs[HE], hc[HE];
        int hp[HE], hq[HE];
#pragma unroll
        for (int j = 0; j < HE; ++j) {
            const int e = sl + j * LPS, ec = e < NU * NU ? e : 0;
            hp[j] = ec / NU; hq[j] = ec - hp[j] * NU;
            hc[j] = __ldg(Cu0 + ec);
            hs[j] = 0.0;
        }
#pragma unroll 4
        for (int r = 0; r < NX; ++r) {
#pragma unroll
            for (int j = 0; j < HE; ++j) hs[j] = fma(sm.Bm[r * NU + hp[j]], sm.SB[r * NU + hq[j]], hs[j]);
        }
#pragma unroll
        for (int j = 0; j < HE; ++j) {
            const int e = sl + j * LPS;
            if (e < NU * NU) sm.H[e] = hc[j] + (hp[j] == hq[j] ? sm.cu[hp[j]] : 0.0) + hs[j];
        }
    }
    // (2) F[:, r] = (Y B)[r, :]
#pragma unroll
    for (int p = 0; p < NU; ++p) { f0[p] = 0.0; f1[p] = 0.0; }
#pragma unroll
    for (int c = 0; c < NX; ++c) {
#pragma unroll
        for (int p = 0; p < NU; ++p) {
            const double bb = sm.Bm[c * NU + p];
            f0[p] = fma(y0[c], bb, f0[p]);
            f1[p] = fma(y1[c], bb, f1[p]);
        }
    }
    __syncwarp(mask);

    // (4) K = -H^{-1} F, two columns per lane
    double L[NU][NU];
#pragma unroll
    for (int p = 0; p < NU; ++p)
#pragma unroll
        for (int q = 0; q < NU; ++q) L[p][q] = sm.H[p * NU + q];
    const bool ok = chol_inplace<NU>(L);
    chol_solve<NU>(L, f0);
    chol_solve<NU>(L, f1);
#pragma unroll
    for (int p = 0; p < NU; ++p) {
        f0[p] = -f0[p];
        f1[p] = -f1[p];
        if (Kout != nullptr) {
            Kout[p * NX + r0] = f0[p];
            if (v1) Kout[p * NX + r1] = f1[p];
        }
    }

    post(f0, f1, L);
    if (!need_S) return ok;

    // (5) closed loop in place:  T[:, r] = A[:, r] + B K[:, r]
#pragma unroll 4
    for (int i = 0; i < NX; ++i) {
        double2 t = *reinterpret_cast<const double2*>(sm.T + i * LD + r0);
#pragma unroll
        for (int p = 0; p < NU; ++p) {
            const double bb = sm.Bm[i * NU + p];
            t.x = fma(bb, f0[p], t.x);
            t.y = fma(bb, f1[p], t.y);
        }
        if (!v1) t.y = 0.0;
        *reinterpret_cast<double2*>(sm.T + i * LD + r0) = t;
    }
    __syncwarp(mask);
    if (Bg_async != nullptr) stage_block_async<NX * NU, LPS>(sm.Bm, Bg_async, sl);      // lands during (6)

    // (6) S' = Cx + Y Acl   (my rows; Y read back from the S buffer)
    double s0[LD], s1[LD];
#pragma unroll
    for (int c = 0; c < LD; ++c) { s0[c] = 0.0; s1[c] = 0.0; }
    rowmm2<NX, LD>([&](int m) { return make_double2(sm.S[r0 * LD + m], sm.S[r1c * LD + m]); }, sm.T, s0, s1);
    // Cx0 == nullptr: the stage cost is diagonal and already part of cxd0 / cxd1
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        if (Cx0 != nullptr) {
            s0[c] = __ldg(Cx0 + r0 * NX + c) + s0[c];
            s1[c] = __ldg(Cx0 + r1c * NX + c) + s1[c];
        }
        if (c == r0) s0[c] += cxd0;
        if (c == r1) s1[c] += cxd1;
    }

    // (7) symmetrise through shared memory (each lane only overwrites its own rows of the Y spill)
#pragma unroll
    for (int c = 0; c < LD; c += 2) {
        *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(s0[c], (c + 1 < NX) ? s0[c + 1] : 0.0);
        if (v1) *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(s1[c], (c + 1 < NX) ? s1[c + 1] : 0.0);
    }
    __syncwarp(mask);
    // every lane is done with A_k + B_k K_k: start copying the next stage's A into sm.T (lands during the symmetrisation)
    if (Ag_async != nullptr) cp_async_matrix<NX, LD, LPS>(sm.T, Ag_async, sl);
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        const double2 t = *reinterpret_cast<const double2*>(sm.S + c * LD + r0);
        s0[c] = 0.5 * (s0[c] + t.x);
        s1[c] = 0.5 * (s1[c] + t.y);
    }
    __syncwarp(mask);
#pragma unroll
    for (int c = 0; c < LD; c += 2) {
        *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(s0[c], (c + 1 < NX) ? s0[c + 1] : 0.0);
        if (v1) *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(s1[c], (c + 1 < NX) ? s1[c + 1] : 0.0);
    }
    __syncwarp(mask);
    return ok;
}

struct NoPre {
    __device__ __forceinline__ void operator()() const {}
};
struct NoPost {
    template <class K, class L>
    __device__ __forceinline__ void operator()(const K&, const K&, const L&) const {}
};

#endif  // RNMPC_EMU

}  // namespace rnmpc
