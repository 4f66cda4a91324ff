// Linearisation kernel: A_k = d f/dx, B_k = d f/du of the RK4 map and the multiple-shooting defect
// c_k = f(x_k,u_k) - x_{k+1} along the nominal trajectory (reference: solver/SCP_SLS_jit.py:199-218, 274-277,
// evaluated there through CasADi).  One thread per (instance, stage, tangent direction): forward-mode dual numbers
// through the four ODE evaluations give one column of [A_k B_k] exactly; the columns of the kinematic-chain states
// (position, velocity) are closed-form and are not differentiated.
#pragma once
#include "rnmpc_common.cuh"
#include "rnmpc_models.cuh"

namespace rnmpc {

template <class M>
__global__ void __launch_bounds__(128) linearize_kernel(int B, int N, const double* __restrict__ X, const double* __restrict__ U,
                                                        double h, double* __restrict__ A, int lda, double* __restrict__ Bm,
                                                        double* __restrict__ c) {
    constexpr int NX = M::NX, NU = M::NU, NZ = NX + NU;
    constexpr int NPAS = M::NPAS, CHAIN = M::CHAIN, NACT = NZ - NPAS;     // kinematic-chain columns are closed-form (rnmpc_models.cuh)
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)B * N * NACT) return;
    const int ta = (int)(gid % NACT);
    const int t = NPAS + ta;
    const long long bk = gid / NACT;
    const int k = (int)(bk % N);
    const long long b = bk / N;
    const double* x = X + (b * (N + 1) + k) * NX;
    const double* u = U + (b * N + k) * NU;
    double* Ao = A + bk * NX * lda;          // lda = NX (caller's dense layout) or Slot<NX>::LD (resident padded layout)
    // the closed-form columns, spread over the stage's threads
    for (int e = ta; e < NPAS * NX; e += NACT) {
        const int i = e / NPAS, tp = e - i * NPAS;
        Ao[i * lda + tp] = (i == tp ? 1.0 : 0.0) + ((tp >= CHAIN && i == tp - CHAIN) ? h : 0.0);
    }
    Dual xd[NX], ud[NU], xp[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xd[i] = Dual(x[i], i == t ? 1.0 : 0.0);
#pragma unroll
    for (int p = 0; p < NU; ++p) ud[p] = Dual(u[p], (NX + p) == t ? 1.0 : 0.0);
    rk4_step<M, Dual>(xd, ud, h, xp);
    if (t < NX) {
#pragma unroll
        for (int i = 0; i < NX; ++i) Ao[i * lda + t] = xp[i].d;
    } else {
        double* Bo = Bm + bk * NX * NU;
#pragma unroll
        for (int i = 0; i < NX; ++i) Bo[i * NU + (t - NX)] = xp[i].d;
    }
    if (ta == 0) {
        const double* xn = x + NX;
        double* co = c + bk * NX;
#pragma unroll
        for (int i = 0; i < NX; ++i) co[i] = xp[i].v - xn[i];
    }
}

// x+ = f(x, u) (+ E w): plant step / roll-out (dyn/model.py:15-34; expe/main_rocket_compare_closed_loop.py:159)
template <class M>
__global__ void ddyn_kernel(int B, const double* x, int xstride, const double* __restrict__ u, int ustride,
                            const double* __restrict__ w, const double* __restrict__ E, double h, double* xp) {
    constexpr int NX = M::NX, NU = M::NU;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double xs[NX], us[NU], xo[NX];
#pragma unroll
    for (int i = 0; i < NX; ++i) xs[i] = x[(size_t)b * xstride + i];
#pragma unroll
    for (int p = 0; p < NU; ++p) us[p] = u[(size_t)b * ustride + p];
    rk4_step<M, double>(xs, us, h, xo);
    if (w != nullptr) {
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            double acc = 0.0;
            for (int j = 0; j < NX; ++j) acc += E[i * NX + j] * w[(size_t)b * NX + j];   // m.E @ w
            xo[i] += acc;
        }
    }
#pragma unroll
    for (int i = 0; i < NX; ++i) xp[(size_t)b * xstride + i] = xo[i];
}

}  // namespace rnmpc
