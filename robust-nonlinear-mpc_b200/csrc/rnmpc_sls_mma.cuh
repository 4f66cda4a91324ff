// fast-SLS kernel, tensor-pipe variant: per disturbance column j the Riccati backward sweep, the closed-loop response propagation
// and the tube back-off terms (reference: solver/fast_SLS_jit.py:43-188, 475-487) with every (NX x NX) product on the FP64 tensor
// pipe (mma.sync.m8n8k4.f64 -> DMMA).  Same inputs, outputs and chain decomposition as sls_kernel (rnmpc_sls.cuh): a chain covers
// column s (N - s stages) and then column N - s (s stages); ONE WARP PER CHAIN, no block-level synchronisation.
//
// Why (profiles/r1i_ncu_full_summary.csv): sls_kernel gives a chain 9 lanes and feeds its DFMAs with broadcast LDS.128 of three
// chains per warp: 0.8 shared-memory wavefronts per clock per SM (the LSU data pipe is the binding unit), FP64 pipe 25 % busy.
// Here a Riccati stage of a chain is
//     W = S_{k+1} [A_k B_k],   M = [A_k B_k]^T W  (upper triangle of tiles),   S_k = Cx + M_xx + (A'SB) K  (one more k-step)
// and a propagation stage is  Phi_u = K Phi,  Phi+ = (A + B K) Phi  with the row norms (beta) taken from the accumulator fragments:
// two shared-memory loads per DMMA (256 FMA) instead of one per four DFMA.  The fragments of [A_k B_k] come straight from global
// memory (L1 / L2: the chains of an instance read the same stages) and serve both products of a stage.
#pragma once
#include "rnmpc_common.cuh"
#include "rnmpc_qp_mma.cuh"
#include "rnmpc_sls.cuh"

namespace rnmpc {

template <int NX, int NU>
struct SlsMma {
    using Q = QpMma<NX, NU>;
    static constexpr int NZ = NX + NU;
    static constexpr int KP = Q::KP, KS = Q::KS, CB = Q::CB, MTS = Q::MTS, LDS = Q::LDS, LDA = Q::LDA;
    static constexpr int NTZ = (CB + NU + 7) / 8;               // column tiles of Z = [A | B]  (= row tiles of M)
    static constexpr int LDW = Q::ld4(8 * NTZ);
    static constexpr int LDF = Q::ld4(8 * MTS);                 // row stride of Phi as B operand
    static constexpr int ROWS = (8 * MTS > KP ? 8 * MTS : KP);
    static constexpr int mx(int a, int b) { return a > b ? a : b; }
    static constexpr int BUF1 = mx(8 * MTS * LDS, 8 * NTZ * LDW);   // S (LDS layout) / M (LDW layout) / A + B K (LDS layout)
    static constexpr int BUF2 = mx(ROWS * LDW, ROWS * LDF);         // W / Phi
    static constexpr int SMALL = ((NU * LDA + NZ + 1) & ~1);        // K_k (Acl build), diagonal cost additions
    static constexpr int PER_WARP = BUF1 + BUF2 + SMALL;            // doubles
    static_assert(NU <= 4, "rank-NU updates are one k-step");
};

// LEAN: the resident solve's configuration (no S / Phi outputs, diagonal regularisation, weights from mu and beta) as its own
// instantiation: those run-time switches sit inside the stage loops (predicated stores per tile, selects per element).
template <int NX, int NU, bool LEAN = false>
__global__ void __launch_bounds__(224, 2) sls_mma_kernel(const SlsArgs a, const Params<NX, NU>* __restrict__ prm, const int dense_reg) {
    double* const aS = LEAN ? nullptr : a.S;
    double* const aPhi_x = LEAN ? nullptr : a.Phi_x;
    double* const aPhi_u = LEAN ? nullptr : a.Phi_u;
    const double* const aeta = LEAN ? nullptr : a.eta;
    const double* const aeta_f = LEAN ? nullptr : a.eta_f;
    const int dreg = LEAN ? 0 : dense_reg;
    using D = SlsMma<NX, NU>;
    constexpr int NZ = NX + NU, NI = 2 * NZ, NIF = 2 * NX, LDG = Slot<NX>::LD;
    constexpr int KP = D::KP, KS = D::KS, CB = D::CB, MTS = D::MTS, NTZ = D::NTZ, LDS = D::LDS, LDW = D::LDW, LDF = D::LDF, LDA = D::LDA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int N = a.N;
    const int n_chain = N / 2 + 1;
    const long long item = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    if (item >= (long long)a.B * n_chain) return;
    const int b = (int)(item / n_chain);
    const int s = (int)(item - (long long)b * n_chain);
    if (a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) return;

    // When the last row tile of an NX-row matrix holds ONE row (NX = 8 j + 1: the rocket's 17), that row of a product is 15 DFMA on
    // the operand fragments already in registers plus two shuffle steps instead of 15 DMMA that are 1/8 useful -- this kernel waits
    // on the tensor pipe (50 % busy, the top stall reason): 150 -> 117 DMMA per (Riccati, propagation) pair.
    constexpr bool ROW1 = (NX - 8 * (MTS - 1)) == 1;
    constexpr int RL = 8 * (MTS - 1);              // the single row
    constexpr int MTD = ROW1 ? MTS - 1 : MTS;       // row tiles that go to the tensor pipe
    double* buf1 = reinterpret_cast<double*>(smem_raw) + (size_t)warp_in_cta * D::PER_WARP;
    double* buf2 = buf1 + D::BUF1;
    double* kk = buf2 + D::BUF2;
    double* dg = kk + NU * LDA;

    const bool use_beta = (a.beta_in != nullptr) && (!a.beta_from_flags || (a.flags[b] & FLAG_TIGHT) != 0);
    const double* Ab = a.A + (size_t)b * N * NX * LDG;
    const double* Bb = a.Bm + (size_t)b * N * NX * NU;
    const size_t kstride = (size_t)NU * NX;

    auto weight = [&](int k, int col, int i) -> double {   // eta[k][col][i]
        if (aeta != nullptr) return aeta[(((size_t)b * N + k) * N + col) * NI + i];
        const double m = a.mu[((size_t)b * N + k) * NI + i];
        double be = kEpsBackoff;
        if (use_beta) be = fmax(a.beta_in[(((size_t)b * N + k) * N + col) * NI + i], kEpsBackoff);
        return 0.5 * m * rsqrt(be);        // mu / (2 sqrt(beta)): one reciprocal square root instead of an IEEE square root and division
    };                                     // (the two were 20 % of this kernel's instructions, profiles/r2_notes.md)
    // eta[k][col][i] + eta[k][col][i + NZ], the diagonal addition of entry i.  LEAN: beta comes from this kernel's own previous launch,
    // which stores the same value for the two sides of a box row -- one load and ONE rsqrt for both (the weights were 7 % of the
    // instructions).
    auto weight2 = [&](int k, int col, int i) -> double {
        if (!LEAN) return weight(k, col, i) + weight(k, col, i + NZ);
        const double* mp = a.mu + ((size_t)b * N + k) * NI + i;
        double be = kEpsBackoff;
        if (use_beta) be = fmax(a.beta_in[(((size_t)b * N + k) * N + col) * NI + i], kEpsBackoff);
        return 0.5 * (mp[0] + mp[NZ]) * rsqrt(be);
    };
    auto weight_f = [&](int col, int i) -> double {        // eta_f[col][i]
        if (aeta_f != nullptr) return aeta_f[((size_t)b * (N + 1) + col) * NIF + i];
        const double m = a.mu_f[(size_t)b * NIF + i];
        double be = kEpsBackoff;
        if (use_beta) be = fmax(a.beta_f_in[((size_t)b * (N + 1) + col) * NIF + i], kEpsBackoff);
        return 0.5 * m * rsqrt(be);
    };
    // fragments of Z_k = [A_k | B_k] (rows padded to KP with zeros): zf[s][n] = Z[4s + t][8n + g]  (branch-free: ZFrag, rnmpc_qp_mma.cuh)
    ZFrag<NX, NU, NTZ, KS, false> zsrc;
    zsrc.init(Ab, Bb, nullptr, lane);
    auto load_z = [&](const int k, double (&zf)[KS][NTZ]) { zsrc.load(k, Ab, zf); };
    auto zero_buf = [&](double* p, const int count) {
        for (int e = 2 * lane; e < count; e += 64) *reinterpret_cast<double2*>(p + e) = make_double2(0.0, 0.0);
    };

    // chain s: column s (N - s stages), then column N - s (s stages; none when it is the same column: even N, s = N/2)
    const int n_first = N - s;
    const int n_second = (N - s == s) ? 0 : s;
    const int n_steps = n_first + n_second;

    // ---------------- backward: Riccati recursion (solver/fast_SLS_jit.py:65-84) ----------------
    {
        if (ROW1) {            // rows RL + 1 .. of W are no longer written by a tile product: they only meet zero operand rows, but must be finite
            zero_buf(buf2, D::BUF2);
            __syncwarp();
        }
        double zf[KS][NTZ];
        if (n_steps > 0) load_z(N - 1, zf);
#pragma unroll 1
        for (int t = 0; t < n_steps; ++t) {
            const bool first = t < n_first;
            const int col = first ? s : N - s;
            const int k = first ? (N - 1 - t) : (N - 1 - (t - n_first));
            if (t == 0 || t == n_first) {
                // S[N, col] = Gf' diag(eta_f[col]) Gf + Q_reg_f   (G = [I; -I] -> diagonal), padded rows / columns zero
                __syncwarp();
                zero_buf(buf1, 8 * MTS * LDS);
                __syncwarp();
                {
                    // the terminal weights go on the diagonal: one pair per lane, started before the copy of Q_reg_f
                    const double wd = lane < NX ? weight_f(col, lane) + weight_f(col, lane + NX) : 0.0;
                    for (int e = lane; e < NX * NX; e += 32) { const int i = e / NX, j = e - i * NX; buf1[i * LDS + j] = prm->Qregf[e]; }
                    __syncwarp();
                    if (lane < NX) buf1[lane * LDS + lane] += wd;
                }
                __syncwarp();
                if (aS != nullptr) {
                    double* So = aS + ((((size_t)b * (N + 1) + N) * (N + 1)) + col) * NX * NX;
                    for (int e = lane; e < NX * NX; e += 32) { const int i = e / NX, j = e - i * NX; So[e] = buf1[i * LDS + j]; }
                }
            }
            // next stage of this chain (same column: k - 1; first stage of the second column: N - 1): fragments in flight during this stage
            double zn_[KS][NTZ];
            if (t + 1 < n_steps) load_z((t + 1 == n_first) ? (N - 1) : (k - 1), zn_);
            // C_kj = G' diag(eta[k, col]) G : diagonal additions to Q_reg / R_reg.  The weight of the NEXT stage (two global loads, a
            // square root and a division per bound) is started here and lands in dg at the end of this stage: off the critical path.
            if (t == 0 && lane < NZ) dg[lane] = weight2(k, col, lane);
            double dg_next = 0.0;
            if (t + 1 < n_steps && lane < NZ) {
                const bool first_n = t + 1 < n_first;
                const int col_n = first_n ? s : N - s;
                const int k_n = first_n ? (N - 2 - t) : (N - 1 - (t + 1 - n_first));
                dg_next = weight2(k_n, col_n, lane);
            }
            // ---- W = S [A B]
            if (ROW1) {
                double sr[KS];
#pragma unroll
                for (int q = 0; q < KS; ++q) sr[q] = buf1[RL * LDS + 4 * q + t4];             // S[RL][4q + t]
#pragma unroll
                for (int n = 0; n < NTZ; ++n) {
                    double acc = 0.0;
#pragma unroll
                    for (int q = 0; q < KS; ++q) acc = fma(sr[q], zf[q][n], acc);              // rows 4q + t of column 8n + g
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    if (t4 == 0) buf2[RL * LDW + 8 * n + g] = acc;
                }
            }
#pragma unroll
            for (int mt = 0; mt < MTD; ++mt) {
                double sa[KS];
#pragma unroll
                for (int q = 0; q < KS; ++q) sa[q] = buf1[(8 * mt + g) * LDS + 4 * q + t4];
#pragma unroll
                for (int n = 0; n < NTZ; ++n) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int q = 0; q < KS; ++q) dmma884(c0, c1, sa[q], zf[q][n]);
                    *reinterpret_cast<double2*>(buf2 + (8 * mt + g) * LDW + 8 * n + 2 * t4) = make_double2(c0, c1);
                }
            }
            __syncwarp();
            // ---- M = [A B]' W  (upper triangle of tiles) over the S buffer (S_{k+1} is no longer needed)
#pragma unroll
            for (int n = 0; n < NTZ; ++n) {
                double wf[KS];
#pragma unroll
                for (int q = 0; q < KS; ++q) wf[q] = buf2[(4 * q + t4) * LDW + 8 * n + g];
#pragma unroll
                for (int mt = 0; mt <= n; ++mt) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int q = 0; q < KS; ++q) dmma884(c0, c1, zf[q][mt], wf[q]);
                    *reinterpret_cast<double2*>(buf1 + (8 * mt + g) * LDW + 8 * n + 2 * t4) = make_double2(c0, c1);
                }
            }
            __syncwarp();
            // ---- H = Cu + B'SB -> Cholesky (every lane its own copy), K = -H^{-1} (A'SB)' : lane i owns column i
            double* Kout = a.K + (((size_t)b * N + k) * (N + 1) + col) * kstride;
            {
                double L[NU][NU];
#pragma unroll
                for (int p = 0; p < NU; ++p)
#pragma unroll
                    for (int q = 0; q < NU; ++q) {
                        const int lo = p < q ? p : q, hi = p < q ? q : p;
                        L[p][q] = buf1[(CB + lo) * LDW + CB + hi] + prm->Rreg[p * NU + q] + (p == q ? dg[NX + p] : 0.0);
                    }
                chol_inplace<NU>(L);
                if (lane < NX) {
                    double f[NU];
#pragma unroll
                    for (int p = 0; p < NU; ++p) f[p] = buf1[lane * LDW + CB + p];
                    chol_solve<NU>(L, f);
#pragma unroll
                    for (int p = 0; p < NU; ++p) { kk[p * LDA + lane] = -f[p]; Kout[p * NX + lane] = -f[p]; }
                }
            }
            __syncwarp();
            // ---- S_k = Cx + A'SA + (A'SB) K : one k-step per upper tile on top of the A'SA tile; operands to registers first, the
            //      buffer then changes from the M layout back to the (zero padded) S layout
            {
                // ROW1: the single column of the last tile as NU DFMA per lane (as in qp_tm_kernel's phase D)
                double scol = 0.0;
                if (ROW1 && lane < NX) {
                    scol = buf1[lane * LDW + RL] + (lane == RL ? dg[RL] : 0.0);
                    scol += dreg ? 0.5 * (prm->Qreg[lane * NX + RL] + prm->Qreg[RL * NX + lane]) : (lane == RL ? prm->Qreg[RL * NX + RL] : 0.0);
#pragma unroll
                    for (int p = 0; p < NU; ++p) scol = fma(buf1[lane * LDW + CB + p], kk[p * LDA + RL], scol);
                }
                double kf[MTS], mc0[MTS * (MTS + 1) / 2], mc1[MTS * (MTS + 1) / 2], gf[MTS];
#pragma unroll
                for (int n = 0; n < MTD; ++n) kf[n] = (t4 < NU && 8 * n + g < NX) ? kk[t4 * LDA + 8 * n + g] : 0.0;
                int ti = 0;
#pragma unroll
                for (int mt = 0; mt < MTD; ++mt) {
                    gf[mt] = (t4 < NU) ? buf1[(8 * mt + g) * LDW + CB + t4] : 0.0;
#pragma unroll
                    for (int n = mt; n < MTD; ++n, ++ti) {
                        const double2 mv = *reinterpret_cast<const double2*>(buf1 + (8 * mt + g) * LDW + 8 * n + 2 * t4);
                        mc0[ti] = mv.x; mc1[ti] = mv.y;
                    }
                }
                // (no zeroing: the pads of the S layout may hold finite leftovers of the M tiles -- they only multiply the zero rows of
                // Z_k; the buffer was cleared when the column started)
                __syncwarp();
                if (ROW1 && lane < NX) { buf1[lane * LDS + RL] = scol; buf1[RL * LDS + lane] = scol; }
                ti = 0;
#pragma unroll
                for (int mt = 0; mt < MTD; ++mt) {
                    const int r = 8 * mt + g;
#pragma unroll
                    for (int n = mt; n < MTD; ++n, ++ti) {
                        const int c = 8 * n + 2 * t4;
                        double c0 = mc0[ti], c1 = mc1[ti];
                        dmma884(c0, c1, gf[mt], kf[n]);
                        if (r < NX) {
                            if (c < NX && r <= c) {
                                double v = c0 + (r == c ? dg[r] : 0.0);
                                v += dreg ? 0.5 * (prm->Qreg[r * NX + c] + prm->Qreg[c * NX + r]) : (r == c ? prm->Qreg[r * NX + r] : 0.0);
                                buf1[r * LDS + c] = v; buf1[c * LDS + r] = v;
                            }
                            if (c + 1 < NX && r <= c + 1) {
                                double v = c1 + (r == c + 1 ? dg[r] : 0.0);
                                v += dreg ? 0.5 * (prm->Qreg[r * NX + c + 1] + prm->Qreg[(c + 1) * NX + r]) : (r == c + 1 ? prm->Qreg[r * NX + r] : 0.0);
                                buf1[r * LDS + c + 1] = v; buf1[(c + 1) * LDS + r] = v;
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (aS != nullptr) {
                double* So = aS + ((((size_t)b * (N + 1) + k) * (N + 1)) + col) * NX * NX;
                for (int e = lane; e < NX * NX; e += 32) { const int i = e / NX, j = e - i * NX; So[e] = buf1[i * LDS + j]; }
            }
#pragma unroll
            for (int q = 0; q < KS; ++q)
#pragma unroll
                for (int n = 0; n < NTZ; ++n) zf[q][n] = zn_[q][n];
            if (lane < NZ) dg[lane] = dg_next;          // every lane is past its last read of dg (barrier above)
        }
    }
    __syncwarp();

    // ---------------- forward: Phi propagation + back-off terms (solver/fast_SLS_jit.py:87-188) ----------------
    // row norms of E (Phi_x[j, j] = E for every column)
    double e_norm = 0.0;
    if (lane < NX) {
#pragma unroll 1
        for (int c = 0; c < NX; ++c) { const double v = prm->E[lane * NX + c]; e_norm = fma(v, v, e_norm); }
    }
    double nx_cur[MTS];                          // lanes with t4 == 0: squared norm of row 8 mt + g of the current Phi_x
    // buf1 holds a zero padded (S layout) matrix from the backward pass (or nothing if the chain had no stage): make the pads zero
    __syncwarp();
    zero_buf(buf1, 8 * MTS * LDS);
#pragma unroll 1
    for (int t = 0; t < n_steps; ++t) {
        const bool first = t < n_first;
        const int col = first ? s : N - s;
        const int k = first ? (s + t) : (N - s) + (t - n_first);      // stages col .. N-1 of the chain's current column
        if (t == 0 || t == n_first) {
            // Phi_x[col, col] = E (zero padded)
            __syncwarp();
            zero_buf(buf2, D::ROWS * LDF);
            __syncwarp();
            for (int e = lane; e < NX * NX; e += 32) { const int i = e / NX, j = e - i * NX; buf2[i * LDF + j] = prm->E[e]; }
#pragma unroll
            for (int mt = 0; mt < MTS; ++mt) nx_cur[mt] = __shfl_sync(0xffffffffu, e_norm, (8 * mt + g) & 31);
            if (aPhi_x != nullptr) {
                double* Po = aPhi_x + ((((size_t)b * (N + 1) + col) * (N + 1)) + col) * NX * NX;
                for (int e = lane; e < NX * NX; e += 32) Po[e] = prm->E[e];
            }
        }
        const double* Ak = Ab + (size_t)k * NX * LDG;
        const double* Bk = Bb + (size_t)k * NX * NU;
        const double* Kin = a.K + (((size_t)b * N + k) * (N + 1) + col) * kstride;
        // ---- A + B K : one k-step per tile on top of the A tile, into buf1 (S layout, pads stay zero)
        double kf[MTS], ka[KS];
#pragma unroll
        for (int n = 0; n < MTS; ++n) kf[n] = (t4 < NU && 8 * n + g < NX) ? Kin[t4 * NX + 8 * n + g] : 0.0;      // B operand K[t][8n + g]
#pragma unroll
        for (int q = 0; q < KS; ++q) ka[q] = (g < NU && 4 * q + t4 < NX) ? Kin[g * NX + 4 * q + t4] : 0.0;       // A operand K[g][4q + t]
        if (ROW1 && lane < NX) {          // row RL of A + B K: NU DFMA per column
            double acc = __ldg(Ak + RL * LDG + lane);
#pragma unroll
            for (int p = 0; p < NU; ++p) acc = fma(__ldg(Bk + RL * NU + p), Kin[p * NX + lane], acc);
            buf1[RL * LDS + lane] = acc;
        }
#pragma unroll
        for (int mt = 0; mt < MTD; ++mt) {
            const int r = 8 * mt + g;
            const double bf = (t4 < NU && r < NX) ? __ldg(Bk + r * NU + t4) : 0.0;
#pragma unroll
            for (int n = 0; n < MTS; ++n) {
                const int c = 8 * n + 2 * t4;
                double c0 = 0.0, c1 = 0.0;
                if (r < NX && c < NX) { c0 = __ldg(Ak + r * LDG + c); if (c + 1 < NX) c1 = __ldg(Ak + r * LDG + c + 1); }
                dmma884(c0, c1, bf, kf[n]);
                if (r < NX && c < NX) { buf1[r * LDS + c] = c0; if (c + 1 < NX) buf1[r * LDS + c + 1] = c1; }
            }
        }
        // ---- fragments of Phi (B operand of both products)
        double pf[KS][MTS];
#pragma unroll
        for (int q = 0; q < KS; ++q)
#pragma unroll
            for (int n = 0; n < MTS; ++n) pf[q][n] = buf2[(4 * q + t4) * LDF + 8 * n + g];
        __syncwarp();          // A + B K complete; every lane holds its fragments of Phi
        // ---- Phi_u = K Phi (rows g < NU), its row norms
        double bu = 0.0;
        {
            double* Po = aPhi_u != nullptr ? aPhi_u + (((size_t)b * N + k) * (N + 1) + col) * kstride : nullptr;
#pragma unroll
            for (int n = 0; n < MTS; ++n) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int q = 0; q < KS; ++q) dmma884(c0, c1, ka[q], pf[q][n]);
                bu = fma(c0, c0, fma(c1, c1, bu));
                const int c = 8 * n + 2 * t4;
                if (Po != nullptr && g < NU) { if (c < NX) Po[g * NX + c] = c0; if (c + 1 < NX) Po[g * NX + c + 1] = c1; }
            }
            bu += __shfl_xor_sync(0xffffffffu, bu, 1);
            bu += __shfl_xor_sync(0xffffffffu, bu, 2);
        }
        // ---- Phi+ = (A + B K) Phi, row norms of the result from the accumulator fragments
        double nc0[MTS][MTS], nc1[MTS][MTS], nx_new[MTS], rowl[MTS];
        if (ROW1) {                          // row RL of Phi+: lanes with t4 == 0 end up with columns g, 8 + g, 16 + g
            double sr[KS], nn = 0.0;
#pragma unroll
            for (int q = 0; q < KS; ++q) sr[q] = buf1[RL * LDS + 4 * q + t4];
#pragma unroll
            for (int n = 0; n < MTS; ++n) {
                double acc = 0.0;
#pragma unroll
                for (int q = 0; q < KS; ++q) acc = fma(sr[q], pf[q][n], acc);
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                rowl[n] = acc;
                nn = fma(acc, acc, nn);
            }
            nn += __shfl_xor_sync(0xffffffffu, nn, 4);
            nn += __shfl_xor_sync(0xffffffffu, nn, 8);
            nn += __shfl_xor_sync(0xffffffffu, nn, 16);
            nx_new[MTS - 1] = nn;            // squared norm of the row (pad columns of Phi are zero); every lane has it
        }
#pragma unroll
        for (int mt = 0; mt < MTD; ++mt) {
            double af[KS];
#pragma unroll
            for (int q = 0; q < KS; ++q) af[q] = buf1[(8 * mt + g) * LDS + 4 * q + t4];
            double nn = 0.0;
#pragma unroll
            for (int n = 0; n < MTS; ++n) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int q = 0; q < KS; ++q) dmma884(c0, c1, af[q], pf[q][n]);
                nc0[mt][n] = c0; nc1[mt][n] = c1;
                nn = fma(c0, c0, fma(c1, c1, nn));
            }
            nn += __shfl_xor_sync(0xffffffffu, nn, 1);
            nn += __shfl_xor_sync(0xffffffffu, nn, 2);
            nx_new[mt] = nn;
        }
        // beta[k, col, :] = max(rowwise ||[Phi_x; Phi_u]||^2, eps)   (G = [I;-I]: rows i and i+nz coincide)
        if (t4 == 0) {
            double* bo = a.beta + (((size_t)b * N + k) * N + col) * NI;
#pragma unroll
            for (int mt = 0; mt < MTS; ++mt) {
                const int r = 8 * mt + g;
                if (r < NX) { const double v = fmax(nx_cur[mt], kEpsBackoff); bo[r] = v; bo[r + NZ] = v; }
            }
            if (g < NU) { const double v = fmax(bu, kEpsBackoff); bo[NX + g] = v; bo[NX + g + NZ] = v; }
            if (k == N - 1) {
                double* bf = a.beta_f + ((size_t)b * (N + 1) + col) * NIF;
#pragma unroll
                for (int mt = 0; mt < MTS; ++mt) {
                    const int r = 8 * mt + g;
                    if (r < NX) { const double v = fmax(nx_new[mt], kEpsBackoff); bf[r] = v; bf[r + NX] = v; }
                }
            }
        }
        // ---- Phi+ over Phi (every lane has its fragments of the old one in registers)
        {
            double* Po = aPhi_x != nullptr ? aPhi_x + ((((size_t)b * (N + 1) + k + 1) * (N + 1)) + col) * NX * NX : nullptr;
#pragma unroll
            for (int mt = 0; mt < MTD; ++mt) {
                const int r = 8 * mt + g;
#pragma unroll
                for (int n = 0; n < MTS; ++n) {
                    const int c = 8 * n + 2 * t4;
                    *reinterpret_cast<double2*>(buf2 + r * LDF + c) = make_double2(nc0[mt][n], nc1[mt][n]);
                    if (Po != nullptr && r < NX) { if (c < NX) Po[r * NX + c] = nc0[mt][n]; if (c + 1 < NX) Po[r * NX + c + 1] = nc1[mt][n]; }
                }
                nx_cur[mt] = nx_new[mt];
            }
            if (ROW1) {
                if (t4 == 0) {
#pragma unroll
                    for (int n = 0; n < MTS; ++n) {
                        const int c = 8 * n + g;
                        buf2[RL * LDF + c] = rowl[n];
                        if (Po != nullptr && c < NX) Po[RL * NX + c] = rowl[n];
                    }
                }
                nx_cur[MTS - 1] = nx_new[MTS - 1];
            }
        }
        __syncwarp();
    }
    if (s == 0) {
        // column N: Phi_x[N, N] = E, no feedback (Q8)
        if (lane < NX) {
            const double v = fmax(e_norm, kEpsBackoff);
            double* bf = a.beta_f + ((size_t)b * (N + 1) + N) * NIF;
            bf[lane] = v; bf[lane + NX] = v;
        }
        if (aPhi_x != nullptr) {
            double* Po = aPhi_x + ((((size_t)b * (N + 1) + N) * (N + 1)) + N) * NX * NX;
            for (int e = lane; e < NX * NX; e += 32) Po[e] = prm->E[e];
        }
        if (aS != nullptr) {
            // S[N, N] = Gf' diag(eta_f[N]) Gf + Q_reg_f (set by the reference although column N has no Riccati step)
            double* So = aS + ((((size_t)b * (N + 1) + N) * (N + 1)) + N) * NX * NX;
            for (int e = lane; e < NX * NX; e += 32) {
                const int i = e / NX, j = e - i * NX;
                So[e] = prm->Qregf[e] + (i == j ? weight_f(N, i) + weight_f(N, i + NX) : 0.0);
            }
        }
    }
}

}  // namespace rnmpc
