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
#include <cuda_runtime.h>
#include <stdint.h>

namespace rnmpc {

constexpr double kEpsBackoff = 1e-10;   // solver/fast_SLS_jit.py:205
constexpr double kEpsQP = 1e-10;        // solver/qp_jit.py:19
// per-instance flag bits shared by the kernels
constexpr int FLAG_FAIL = 1;    // a QP of this solve failed: the instance is frozen (reference: success=False)
constexpr int FLAG_Q3 = 2;      // check_convergence_socp returned True: skip backward solve + tightening (Q3)
constexpr int FLAG_TIGHT = 4;   // tube back-offs were computed in this fast-SLS solve and are in the QP bounds
constexpr int FLAG_PREV = 8;    // _prev_primal_vec exists (Q3 state)
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

// acc{0,1}[c] += sum_m in{0,1}[m] * Mat[m*LD + c]   for c in [0, LD)   (Mat in shared memory, 16-byte aligned)
template <int M, int LD>
__device__ __forceinline__ void rowmm2(const double (&in0)[M], const double (&in1)[M], const double* __restrict__ Mat,
                                       double (&acc0)[LD], double (&acc1)[LD]) {
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int c = 0; c < LD; c += 2) {
            const double2 v = *reinterpret_cast<const double2*>(Mat + m * LD + c);
            acc0[c] = fma(in0[m], v.x, acc0[c]);
            acc0[c + 1] = fma(in0[m], v.y, acc0[c + 1]);
            acc1[c] = fma(in1[m], v.x, acc1[c]);
            acc1[c + 1] = fma(in1[m], v.y, acc1[c + 1]);
        }
    }
}

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
        const double l = sqrt(d);
        const double il = 1.0 / l;
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

// Shared-memory working set of one slot.
template <int NX, int NU>
struct SlotSmem {
    static constexpr int LD = Slot<NX>::LD;
    double S[NX * LD];              // current cost-to-go S (backward) / Phi_x (forward), full matrix for broadcast
    double T[NX * LD];              // closed-loop matrix A+BK (backward)
    double Bm[even_up(NX * NU)];    // B_k
    double SB[even_up(NX * NU)];    // S B
    double H[even_up(NU * NU)];     // Cu + B'SB
    double K[NU * LD];              // K_k (forward) / Phi_u
    double PU[NU * LD];
    double cu[even_up(NU)];         // diagonal additions of the input cost
    double vec[4 * LD];             // vector scratch (QP sweeps)
    double red[even_up(4 * Slot<NX>::LPS)];   // reduction scratch
};

// One backward Riccati step of a chain (solver/fast_SLS_jit.py:43-62):
//   x = B'S, y = A'S, H = Cu + xB, F = xA, K = -H^{-1}F, S' = Cx + y(A + BK), S' <- (S'+S'^T)/2
// On entry sm.S holds S_{k+1}; on exit it holds S_k.  `acol` (columns r0,r1 of A_k), `kcol` (columns of K_k) and the
// Cholesky factor stay in registers for the caller (the QP fuses its vector recursion here through `Hook`).
template <int NX, int NU, class Hook>
__device__ __forceinline__ bool riccati_step(SlotSmem<NX, NU>& sm, const int sl, const unsigned mask,
                                             const double* __restrict__ Ag, const double* __restrict__ Bg,
                                             const double* __restrict__ Cx0, const double* __restrict__ Cu0,
                                             const double cxd0, const double cxd1, double* __restrict__ Kout, Hook&& hook,
                                             const bool need_S = true) {
    constexpr int LD = Slot<NX>::LD;
    constexpr int LPS = Slot<NX>::LPS;
    const int r0 = 2 * sl, r1 = 2 * sl + 1;
    const bool v1 = r1 < NX;
    const int r1c = v1 ? r1 : NX - 1;

    // (0) columns r0,r1 of A_k into registers, B_k into shared memory
    double a0[NX], a1[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        a0[i] = __ldg(Ag + i * NX + r0);
        a1[i] = v1 ? __ldg(Ag + i * NX + r1c) : 0.0;
    }
    for (int e = sl; e < NX * NU; e += LPS) sm.Bm[e] = __ldg(Bg + e);
    __syncwarp(mask);

    // (1) Y = A'S  (my rows)
    double y0[LD], y1[LD];
#pragma unroll
    for (int c = 0; c < LD; ++c) { y0[c] = 0.0; y1[c] = 0.0; }
    rowmm2<NX, LD>(a0, a1, sm.S, y0, y1);

    // (2) F[:, r] = (Y B)[r, :]      (3) SB rows
    double f0[NU], f1[NU];
    {
        double sb0[NU], sb1[NU];
#pragma unroll
        for (int p = 0; p < NU; ++p) { f0[p] = 0.0; f1[p] = 0.0; sb0[p] = 0.0; sb1[p] = 0.0; }
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            const double s0 = sm.S[r0 * LD + c], s1 = sm.S[r1c * LD + c];
#pragma unroll
            for (int p = 0; p < NU; ++p) {
                const double b = sm.Bm[c * NU + p];
                f0[p] = fma(y0[c], b, f0[p]);
                f1[p] = fma(y1[c], b, f1[p]);
                sb0[p] = fma(s0, b, sb0[p]);
                sb1[p] = fma(s1, b, sb1[p]);
            }
        }
#pragma unroll
        for (int p = 0; p < NU; ++p) {
            sm.SB[r0 * NU + p] = sb0[p];
            if (v1) sm.SB[r1 * NU + p] = sb1[p];
        }
    }
    __syncwarp(mask);
    for (int e = sl; e < NU * NU; e += LPS) {
        const int p = e / NU, q = e % NU;
        double hsum = 0.0;
#pragma unroll
        for (int r = 0; r < NX; ++r) hsum = fma(sm.Bm[r * NU + p], sm.SB[r * NU + q], hsum);
        sm.H[e] = __ldg(Cu0 + e) + (p == q ? sm.cu[p] : 0.0) + hsum;
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

    // caller hook: vector recursion that needs A columns, K columns, the factor and S_{k+1}
    hook(a0, a1, f0, f1, L);
    if (!need_S) return ok;

    // (5) closed-loop columns  Acl[:, r] = A[:, r] + B K[:, r]  -> sm.T
#pragma unroll
    for (int i = 0; i < NX; ++i) {
        double c0 = a0[i], c1 = a1[i];
#pragma unroll
        for (int p = 0; p < NU; ++p) {
            const double b = sm.Bm[i * NU + p];
            c0 = fma(b, f0[p], c0);
            c1 = fma(b, f1[p], c1);
        }
        *reinterpret_cast<double2*>(sm.T + i * LD + r0) = make_double2(c0, v1 ? c1 : 0.0);
    }
    __syncwarp(mask);

    // (6) S' = Cx + Y Acl   (my rows)
    double s0[LD], s1[LD];
#pragma unroll
    for (int c = 0; c < LD; ++c) { s0[c] = 0.0; s1[c] = 0.0; }
    {
        double yy0[NX], yy1[NX];
#pragma unroll
        for (int m = 0; m < NX; ++m) { yy0[m] = y0[m]; yy1[m] = y1[m]; }
        rowmm2<NX, LD>(yy0, yy1, sm.T, s0, s1);
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        s0[c] = __ldg(Cx0 + r0 * NX + c) + s0[c];
        s1[c] = __ldg(Cx0 + r1c * NX + c) + s1[c];
    }
#pragma unroll
    for (int c = 0; c < NX; ++c) {
        if (c == r0) s0[c] += cxd0;
        if (c == r1) s1[c] += cxd1;
    }
    __syncwarp(mask);

    // (7) symmetrise through shared memory
#pragma unroll
    for (int c = 0; c < LD; c += 2) {
        *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(s0[c], (c + 1 < NX) ? s0[c + 1] : 0.0);
        if (v1) *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(s1[c], (c + 1 < NX) ? s1[c + 1] : 0.0);
    }
    __syncwarp(mask);
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

struct NoHook {
    template <class A, class K, class L>
    __device__ __forceinline__ void operator()(const A&, const A&, const K&, const K&, const L&) const {}
};

}  // namespace rnmpc
