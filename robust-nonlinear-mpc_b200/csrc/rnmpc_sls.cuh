// fast-SLS kernel: per disturbance column j the Riccati backward sweep, the closed-loop response propagation
// and the tube back-off terms (reference: solver/fast_SLS_jit.py:43-188, 475-487).
#pragma once
#include "rnmpc_common.cuh"

namespace rnmpc {

struct SlsArgs {
    int B, N;
    const double* A;       // [B][N][nx][LD] (padded rows, see Slot<NX>::LD)
    const double* Bm;      // [B][N][nx][nu]
    // weights: either explicit eta/eta_f (reference layout [B][N][N][ni], [B][N+1][ni_f]) or mu/mu_f + beta/beta_f
    const double* eta;
    const double* eta_f;
    const double* mu;      // [B][N][ni]      (dual_mu transposed: stage-major)
    const double* mu_f;    // [B][ni_f]
    const double* beta_in; // [B][N][N][ni]   or nullptr -> epsilon (first pass of every fast-SLS solve, Q5)
    const double* beta_f_in;  // [B][N+1][ni_f]
    const int* flags;      // [B] or nullptr: instances with (flags & skip_mask) do no work (Q3 early return, failed QP)
    int skip_mask;
    int beta_from_flags;   // 1: use beta_in only for instances with FLAG_TIGHT set (else epsilon)
    double* K;             // [B][N][N+1][nu][nx]
    double* S;             // [B][N+1][N+1][nx][nx] or nullptr
    double* Phi_x;         // [B][N+1][N+1][nx][nx] or nullptr
    double* Phi_u;         // [B][N][N+1][nu][nx]   or nullptr
    double* beta;          // [B][N][N][ni]
    double* beta_f;        // [B][N+1][ni_f]
};

#ifndef RNMPC_EMU
// Chain s of an instance covers column s (N-s steps) followed by column N-s (s steps): N Riccati steps and N
// propagation steps for every chain, so all slots of a warp run the same trip counts.
template <int NX, int NU>
__global__ void __launch_bounds__(160, 2) sls_kernel(const SlsArgs a, const Params<NX, NU>* __restrict__ prm) {
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
    const int n_chain = N / 2 + 1;
    const long long warp_global = (long long)blockIdx.x * (blockDim.x >> 5) + warp_in_cta;
    const long long item = warp_global * SPW + slot_w;
    const int b = (int)(item / n_chain);
    const int s = (int)(item - (long long)b * n_chain);
    bool active = (slot_w < SPW) && (item < (long long)a.B * n_chain);
    if (active && a.flags != nullptr && (a.flags[b] & a.skip_mask) != 0) active = false;
    const unsigned mask = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    SlotSmem<NX, NU>& sm = smem[warp_in_cta * SPW + slot_w];

    const int r0 = 2 * sl, r1 = 2 * sl + 1;
    const bool v1 = r1 < NX;
    const int r1c = v1 ? r1 : NX - 1;
    const bool use_beta = (a.beta_in != nullptr) && (!a.beta_from_flags || (a.flags[b] & FLAG_TIGHT) != 0);
    const double* Ab = a.A + (size_t)b * N * NX * LD;
    const double* Bb = a.Bm + (size_t)b * N * NX * NU;
    const size_t kstride = (size_t)NU * NX;

    auto weight = [&](int k, int col, int i) -> double {   // eta[k][col][i]
        if (a.eta != nullptr) return a.eta[(((size_t)b * N + k) * N + col) * NI + i];
        const double m = a.mu[((size_t)b * N + k) * NI + i];
        double be = kEpsBackoff;
        if (use_beta) be = fmax(a.beta_in[(((size_t)b * N + k) * N + col) * NI + i], kEpsBackoff);
        return m / (2.0 * sqrt(be));
    };
    auto weight_f = [&](int col, int i) -> double {        // eta_f[col][i]
        if (a.eta_f != nullptr) return a.eta_f[((size_t)b * (N + 1) + col) * NIF + i];
        const double m = a.mu_f[(size_t)b * NIF + i];
        double be = kEpsBackoff;
        if (use_beta) be = fmax(a.beta_f_in[((size_t)b * (N + 1) + col) * NIF + i], kEpsBackoff);
        return m / (2.0 * sqrt(be));
    };

    // ---------------- backward: Riccati recursion (solver/fast_SLS_jit.py:65-84) ----------------
    const int n_first = N - s;     // steps of column s
    bool t_ready = false;
    for (int t = 0; t < N; ++t) {
        const bool first = t < n_first;
        const int col = first ? s : N - s;
        const int k = first ? (N - 1 - t) : (N - 1 - (t - n_first));
        if (t == 0 || t == n_first) {
            // S[N, col] = Gf' diag(eta_f[col]) Gf + Q_reg_f   (G = [I; -I] -> diagonal)
            const double d0 = weight_f(col, r0) + weight_f(col, r0 + NX);
            const double d1 = v1 ? (weight_f(col, r1) + weight_f(col, r1 + NX)) : 0.0;
#pragma unroll
            for (int c = 0; c < LD; c += 2) {
                double x0 = (c < NX) ? prm->Qregf[r0 * NX + c] : 0.0, x1 = (c + 1 < NX) ? prm->Qregf[r0 * NX + c + 1] : 0.0;
                if (c == r0) x0 += d0;
                if (c + 1 == r0) x1 += d0;
                *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(x0, x1);
                if (v1) {
                    double z0 = (c < NX) ? prm->Qregf[r1 * NX + c] : 0.0, z1 = (c + 1 < NX) ? prm->Qregf[r1 * NX + c + 1] : 0.0;
                    if (c == r1) z0 += d1;
                    if (c + 1 == r1) z1 += d1;
                    *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(z0, z1);
                }
            }
            if (a.S != nullptr) {
                double* So = a.S + ((((size_t)b * (N + 1) + N) * (N + 1)) + col) * NX * NX;
                for (int c = 0; c < NX; ++c) {
                    So[r0 * NX + c] = sm.S[r0 * LD + c];
                    if (v1) So[r1 * NX + c] = sm.S[r1 * LD + c];
                }
            }
        }
        __syncwarp(mask);      // outside the branch: slots of one warp start their second column at different t
        // C_kj = G' diag(eta[k, col]) G : diagonal additions to Q_reg / R_reg
        const double cx0 = weight(k, col, r0) + weight(k, col, r0 + NZ);
        const double cx1 = v1 ? (weight(k, col, r1) + weight(k, col, r1 + NZ)) : 0.0;
        if (sl < NU) sm.cu[sl] = weight(k, col, NX + sl) + weight(k, col, NX + sl + NZ);
        if (LPS < NU) { for (int p = LPS + sl; p < NU; p += LPS) sm.cu[p] = weight(k, col, NX + p) + weight(k, col, NX + p + NZ); }
        double* Kout = a.K + (((size_t)b * N + k) * (N + 1) + col) * kstride;
        // stage of the next Riccati step of this chain (same column: k-1; first step of the second column: N-1)
        const double* a_next = nullptr;
        const double* b_next = nullptr;
        if (t + 1 < N) {
            const int kn = (t + 1 < n_first) ? (k - 1) : ((t + 1 == n_first) ? (N - 1) : (k - 1));
            a_next = Ab + (size_t)kn * NX * LD;
            b_next = Bb + (size_t)kn * NX * NU;
        }
        riccati_step<NX, NU>(sm, sl, mask, Ab + (size_t)k * NX * LD, Bb + (size_t)k * NX * NU, prm->Qreg, prm->Rreg, cx0, cx1,
                             Kout, NoPre(), NoPost(), true, k > 0 ? Ab + (size_t)(k - 1) * NX * LD : nullptr,
                             k > 0 ? Bb + (size_t)(k - 1) * NX * NU : nullptr, t_ready, a_next, t_ready, b_next);
        t_ready = a_next != nullptr;
        if (a.S != nullptr) {
            double* So = a.S + ((((size_t)b * (N + 1) + k) * (N + 1)) + col) * NX * NX;
            for (int c = 0; c < NX; ++c) {
                So[r0 * NX + c] = sm.S[r0 * LD + c];
                if (v1) So[r1 * NX + c] = sm.S[r1 * LD + c];
            }
        }
    }
    __syncwarp(mask);

    // ---------------- forward: Phi propagation + back-off terms (solver/fast_SLS_jit.py:87-188) ----------------
    for (int t = 0; t < N; ++t) {
        const bool first = t < n_first;
        const int col = first ? s : N - s;
        const int k = first ? (s + t) : t;
        if (t == 0 || t == n_first) {
            // Phi_x[col, col] = E
#pragma unroll
            for (int c = 0; c < LD; c += 2) {
                const double x0 = (c < NX) ? prm->E[r0 * NX + c] : 0.0, x1 = (c + 1 < NX) ? prm->E[r0 * NX + c + 1] : 0.0;
                *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(x0, x1);
                if (v1) {
                    const double z0 = (c < NX) ? prm->E[r1 * NX + c] : 0.0, z1 = (c + 1 < NX) ? prm->E[r1 * NX + c + 1] : 0.0;
                    *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(z0, z1);
                }
            }
            if (a.Phi_x != nullptr) {
                double* Po = a.Phi_x + ((((size_t)b * (N + 1) + col) * (N + 1)) + col) * NX * NX;
                for (int c = 0; c < NX; ++c) {
                    Po[r0 * NX + c] = prm->E[r0 * NX + c];
                    if (v1) Po[r1 * NX + c] = prm->E[r1 * NX + c];
                }
            }
        }
        // K[k, col] -> shared, transposed (Kt[m][p]); A_k -> sm.T; loads batched so their latencies overlap
        const double* Kin = a.K + (((size_t)b * N + k) * (N + 1) + col) * kstride;
        {
            constexpr int ITK = (NU * NX + LPS - 1) / LPS;
            double kv[ITK];
#pragma unroll
            for (int j = 0; j < ITK; ++j) { const int e = sl + j * LPS; kv[j] = (e < NU * NX) ? Kin[e] : 0.0; }
            load_padded_smem<NX, LD, LPS>(sm.T, Ab + (size_t)k * NX * LD, sl);
#pragma unroll
            for (int j = 0; j < ITK; ++j) {
                const int e = sl + j * LPS;
                if (e < NU * NX) { const int p = e / NX, m = e - p * NX; sm.K[m * NU + p] = kv[j]; }
            }
        }
        __syncwarp(mask);
        // Phi_u[k, col] = K Phi_x : my two columns
        {
            double pu0[NU], pu1[NU];
#pragma unroll
            for (int p = 0; p < NU; ++p) { pu0[p] = 0.0; pu1[p] = 0.0; }
#pragma unroll 1
            for (int m = 0; m < NX; ++m) {
                const double2 ph = *reinterpret_cast<const double2*>(sm.S + m * LD + r0);
#pragma unroll
                for (int p = 0; p < NU; ++p) {
                    const double kk = sm.K[m * NU + p];
                    pu0[p] = fma(kk, ph.x, pu0[p]);
                    pu1[p] = fma(kk, ph.y, pu1[p]);
                }
            }
#pragma unroll
            for (int p = 0; p < NU; ++p) {
                *reinterpret_cast<double2*>(sm.PU + p * LD + r0) = make_double2(pu0[p], v1 ? pu1[p] : 0.0);
                if (a.Phi_u != nullptr) {
                    double* Po = a.Phi_u + (((size_t)b * N + k) * (N + 1) + col) * kstride;
                    Po[p * NX + r0] = pu0[p];
                    if (v1) Po[p * NX + r1] = pu1[p];
                }
            }
        }
        // closed-loop rows  (A + B K)[r, :] -> sm.T (my rows), and beta of the state rows
        {
            const double* Bk = Bb + (size_t)k * NX * NU;
            double b0[NU], b1[NU];
#pragma unroll
            for (int p = 0; p < NU; ++p) { b0[p] = __ldg(Bk + r0 * NU + p); b1[p] = __ldg(Bk + r1c * NU + p); }
            double n0 = 0.0, n1 = 0.0;
#pragma unroll 1
            for (int m = 0; m < NX; ++m) {
                double x0 = sm.T[r0 * LD + m], x1 = sm.T[r1c * LD + m];      // A_k rows (in place -> closed loop rows)
#pragma unroll
                for (int p = 0; p < NU; ++p) {
                    const double kk = sm.K[m * NU + p];
                    x0 = fma(b0[p], kk, x0);
                    x1 = fma(b1[p], kk, x1);
                }
                sm.T[r0 * LD + m] = x0;
                if (v1) sm.T[r1 * LD + m] = x1;
                const double p0 = sm.S[r0 * LD + m], p1 = sm.S[r1c * LD + m];
                n0 = fma(p0, p0, n0);
                n1 = fma(p1, p1, n1);
            }
            // beta[k, col, :] = max(rowwise ||[Phi_x; Phi_u]||^2, eps)   (G = [I;-I]: rows i and i+nz coincide)
            double* bo = a.beta + (((size_t)b * N + k) * N + col) * NI;
            n0 = fmax(n0, kEpsBackoff);
            n1 = fmax(n1, kEpsBackoff);
            bo[r0] = n0; bo[r0 + NZ] = n0;
            if (v1) { bo[r1] = n1; bo[r1 + NZ] = n1; }
            __syncwarp(mask);      // PU complete
            for (int p = sl; p < NU; p += LPS) {
                double nn = 0.0;
#pragma unroll
                for (int c = 0; c < NX; ++c) { const double q = sm.PU[p * LD + c]; nn = fma(q, q, nn); }
                nn = fmax(nn, kEpsBackoff);
                bo[NX + p] = nn; bo[NX + p + NZ] = nn;
            }
        }
        // Phi_x[k+1, col] = (A + B K) Phi_x[k, col] : my rows
        double ph0[LD], ph1[LD];
#pragma unroll
        for (int c = 0; c < LD; ++c) { ph0[c] = 0.0; ph1[c] = 0.0; }
        rowmm2<NX, LD>([&](int m) { return make_double2(sm.T[r0 * LD + m], sm.T[r1c * LD + m]); }, sm.S, ph0, ph1);
        __syncwarp(mask);
#pragma unroll
        for (int c = 0; c < LD; c += 2) {
            *reinterpret_cast<double2*>(sm.S + r0 * LD + c) = make_double2(ph0[c], (c + 1 < NX) ? ph0[c + 1] : 0.0);
            if (v1) *reinterpret_cast<double2*>(sm.S + r1 * LD + c) = make_double2(ph1[c], (c + 1 < NX) ? ph1[c + 1] : 0.0);
        }
        if (a.Phi_x != nullptr) {
            double* Po = a.Phi_x + ((((size_t)b * (N + 1) + k + 1) * (N + 1)) + col) * NX * NX;
#pragma unroll
            for (int c = 0; c < NX; ++c) {
                Po[r0 * NX + c] = ph0[c];
                if (v1) Po[r1 * NX + c] = ph1[c];
            }
        }
        if (k == N - 1) {
            // beta_f[col, :] from Phi_x[N, col]
            double n0 = 0.0, n1 = 0.0;
#pragma unroll
            for (int c = 0; c < NX; ++c) { n0 = fma(ph0[c], ph0[c], n0); n1 = fma(ph1[c], ph1[c], n1); }
            n0 = fmax(n0, kEpsBackoff);
            n1 = fmax(n1, kEpsBackoff);
            double* bf = a.beta_f + ((size_t)b * (N + 1) + col) * NIF;
            bf[r0] = n0; bf[r0 + NX] = n0;
            if (v1) { bf[r1] = n1; bf[r1 + NX] = n1; }
        }
        __syncwarp(mask);
    }
    if (s == 0) {
        // column N: Phi_x[N, N] = E, no feedback (Q8)
        double n0 = 0.0, n1 = 0.0;
#pragma unroll
        for (int c = 0; c < NX; ++c) {
            const double e0 = prm->E[r0 * NX + c], e1 = prm->E[r1c * NX + c];
            n0 = fma(e0, e0, n0);
            n1 = fma(e1, e1, n1);
            if (a.Phi_x != nullptr) {
                double* Po = a.Phi_x + ((((size_t)b * (N + 1) + N) * (N + 1)) + N) * NX * NX;
                Po[r0 * NX + c] = e0;
                if (v1) Po[r1 * NX + c] = e1;
            }
        }
        n0 = fmax(n0, kEpsBackoff);
        n1 = fmax(n1, kEpsBackoff);
        double* bf = a.beta_f + ((size_t)b * (N + 1) + N) * NIF;
        bf[r0] = n0; bf[r0 + NX] = n0;
        if (v1) { bf[r1] = n1; bf[r1 + NX] = n1; }
        if (a.S != nullptr) {
            // S[N, N] = Gf' diag(eta_f[N]) Gf + Q_reg_f (set by the reference although column N has no Riccati step)
            double* So = a.S + ((((size_t)b * (N + 1) + N) * (N + 1)) + N) * NX * NX;
            const double d0 = weight_f(N, r0) + weight_f(N, r0 + NX);
            const double d1 = v1 ? (weight_f(N, r1) + weight_f(N, r1 + NX)) : 0.0;
            for (int c = 0; c < NX; ++c) {
                So[r0 * NX + c] = prm->Qregf[r0 * NX + c] + (c == r0 ? d0 : 0.0);
                if (v1) So[r1 * NX + c] = prm->Qregf[r1 * NX + c] + (c == r1 ? d1 : 0.0);
            }
        }
    }
}

#endif  // RNMPC_EMU

// backoff[k][i] = sum_{j<=k} sqrt(beta[k][j][i]),  backoff_f[i] = sum_{j<=N} sqrt(beta_f[j][i])
// (solver/fast_SLS_jit.py:172-186); one thread per (instance, k, i).
static __global__ void backoff_reduce_kernel(int B, int N, int ni, int ni_f, const double* __restrict__ beta,
                                      const double* __restrict__ beta_f, const int* __restrict__ flags, int skip_mask,
                                      double* __restrict__ backoff, double* __restrict__ backoff_f) {
    const long long per = (long long)N * ni + ni_f;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * per) return;
    const int b = (int)(gid / per);
    if (flags != nullptr && (flags[b] & skip_mask) != 0) return;
    const int e = (int)(gid - (long long)b * per);
    if (e < N * ni) {
        const int k = e / ni, i = e - k * ni;
        // eight loads in flight, summed in the reference's order (j ascending)
        const double* bp = beta + (((size_t)b * N + k) * N) * ni + i;
        double acc = 0.0;
        for (int j0 = 0; j0 <= k; j0 += 8) {
            double v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] = (j0 + jj <= k) ? bp[(size_t)(j0 + jj) * ni] : 0.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) if (j0 + jj <= k) acc += sqrt(v[jj]);
        }
        backoff[((size_t)b * N + k) * ni + i] = acc;
    } else {
        const int i = e - N * ni;
        const double* bp = beta_f + ((size_t)b * (N + 1)) * ni_f + i;
        double acc = 0.0;
        for (int j0 = 0; j0 <= N; j0 += 8) {
            double v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) v[jj] = (j0 + jj <= N) ? bp[(size_t)(j0 + jj) * ni_f] : 0.0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) if (j0 + jj <= N) acc += sqrt(v[jj]);
        }
        backoff_f[(size_t)b * ni_f + i] = acc;
    }
}

}  // namespace rnmpc
