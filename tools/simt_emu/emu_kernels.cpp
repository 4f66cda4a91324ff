// C entry points that run the product's kernels (csrc/*.cuh, unmodified source, compiled with -DRNMPC_EMU) block by block on the
// host through the SIMT emulator.  Test infrastructure only: the CPU suite checks kernel logic against the oracle here, the GPU
// suite checks the compiled sm_100a kernels through the C ABI.
#define RNMPC_EMU 1
#include "simt_emu.h"

#include "../../include/rnmpc_b200.h"
#include "../../robust-nonlinear-mpc_b200/csrc/rnmpc_qp_mma.cuh"
#include "../../robust-nonlinear-mpc_b200/csrc/rnmpc_qp_warp.cuh"
#include "../../robust-nonlinear-mpc_b200/csrc/rnmpc_qp_tmem.cuh"
#include "../../robust-nonlinear-mpc_b200/csrc/rnmpc_sls_mma.cuh"

namespace rnmpc {
alignas(128) unsigned char smem_raw[232448];
}

using namespace rnmpc;

static int g_generic = 0;      // 1: run the run-time-horizon instantiation even for N = 15
extern "C" void emu_set_order(int order) { simt::g_order = order; }
extern "C" void emu_set_generic(int on) { g_generic = on; }

template <int NX, int NU>
static void fill_params(const rnmpc_config* cfg, Params<NX, NU>& hp) {
    memcpy(hp.Q, cfg->Q, sizeof(hp.Q)); memcpy(hp.R, cfg->R, sizeof(hp.R)); memcpy(hp.Qf, cfg->Qf, sizeof(hp.Qf));
    for (int i = 0; i < NX * NX; ++i) hp.Q2[i] = 2.0 * hp.Q[i];
    for (int i = 0; i < NU * NU; ++i) hp.R2[i] = 2.0 * hp.R[i];
    memcpy(hp.Qreg, cfg->Q_reg, sizeof(hp.Qreg)); memcpy(hp.Rreg, cfg->R_reg, sizeof(hp.Rreg));
    memcpy(hp.Qregf, cfg->Q_reg_f, sizeof(hp.Qregf)); memcpy(hp.E, cfg->E, sizeof(hp.E));
    memcpy(hp.xub, cfg->x_ub, sizeof(hp.xub)); memcpy(hp.xlb, cfg->x_lb, sizeof(hp.xlb));
    memcpy(hp.uub, cfg->u_ub, sizeof(hp.uub)); memcpy(hp.ulb, cfg->u_lb, sizeof(hp.ulb));
    memcpy(hp.xubf, cfg->xf_ub, sizeof(hp.xubf)); memcpy(hp.xlbf, cfg->xf_lb, sizeof(hp.xlbf));
    hp.h = cfg->h;
}

static int g_qp_kernel = 0;    // 0: qp_mma_kernel (one CTA of 4 warps per instance), 1: qp_warp_kernel (one warp per instance),
                               // 2: qp_tm_kernel (one warp per instance, iterate in tensor memory, persistent CTAs of 4 warps)
extern "C" void emu_set_qp_kernel(int k) { g_qp_kernel = k; }

template <int NX, int NU>
static int run_qp_mma(const rnmpc_config* cfg, QpArgs a) {
    Params<NX, NU> hp;
    fill_params<NX, NU>(cfg, hp);
    if (g_qp_kernel == 2) {
        if (qp_tm_smem_bytes<NX, NU>(a.N) > sizeof(smem_raw) || a.N + 1 > kTmMaxStages) return -4;
        std::vector<double> Kw((size_t)a.B * a.N * QpWarp<NX, NU>::SLOT, 0.0), Lw((size_t)a.B * a.N * NU * NU, 0.0);      // sweep images
        a.Kq = Kw.data(); a.Lq = Lw.data();
        int work = 0;
        a.work = &work;
        const int grid = a.B >= 6 ? 2 : 1;          // two persistent CTAs: the warps of the second find the counter advanced by the first
        for (int blk = 0; blk < grid; ++blk) {
            simt::cpa_reset();
            if (blk == 1) work = a.B / 2;            // (the emulator runs the blocks one after the other: hand the second CTA its share)
            const int Bsave = a.B;
            if (blk == 0 && grid == 2) a.B = Bsave / 2;
            if (a.N == 15 && !g_generic) simt::run_block(32 * kTmWarps, dim3(blk), dim3(grid), [&]() { qp_tm_kernel<NX, NU, 15>(a, &hp); });
            else simt::run_block(32 * kTmWarps, dim3(blk), dim3(grid), [&]() { qp_tm_kernel<NX, NU, 0>(a, &hp); });
            a.B = Bsave;
        }
        return 0;
    }
    if (g_qp_kernel == 1) {
        if (qp_warp_smem_doubles<NX, NU>(a.N) * sizeof(double) > sizeof(smem_raw)) return -4;
        std::vector<double> Kw((size_t)a.B * a.N * NU * QpWarp<NX, NU>::LDA, 0.0), Lw((size_t)a.B * a.N * NU * NU, 0.0);
        a.Kq = Kw.data(); a.Lq = Lw.data();
        for (int b = 0; b < a.B; ++b) {
            simt::cpa_reset();
            if (a.N == 15 && !g_generic) simt::run_block(32, dim3(b), dim3(a.B), [&]() { qp_warp_kernel<NX, NU, 15>(a, &hp); });
            else simt::run_block(32, dim3(b), dim3(a.B), [&]() { qp_warp_kernel<NX, NU, 0>(a, &hp); });
        }
        return 0;
    }
    if (qp_mma_smem_doubles<NX, NU>(a.N) * sizeof(double) > sizeof(smem_raw)) return -4;
    for (int b = 0; b < a.B; ++b) {
        if (a.N == 15 && !g_generic) simt::run_block(kMmaThreads, dim3(b), dim3(a.B), [&]() { qp_mma_kernel<NX, NU, 15>(a, &hp); });
        else simt::run_block(kMmaThreads, dim3(b), dim3(a.B), [&]() { qp_mma_kernel<NX, NU, 0>(a, &hp); });
    }
    return 0;
}

extern "C" int emu_qp_mma(const rnmpc_config* cfg, int B, int tightened, const double* A_padded, const double* Bm, const double* c,
                          const double* X, const double* U, const double* x0, const double* backoff, const double* backoff_f,
                          const int* flags, int skip_mask, int max_iter, double tol_mu, double tol_res, double mu0, double* dX,
                          double* dU, double* mu, double* mu_f, double* cost, int* iters, int* status) {
    QpArgs a{};
    a.B = B; a.N = cfg->N; a.tightened = tightened; a.max_iter = max_iter; a.tol_mu = tol_mu; a.tol_res = tol_res; a.mu0 = mu0;
    a.A = A_padded; a.Bm = Bm; a.c = c; a.Xnom = X; a.Unom = U; a.x0 = x0; a.backoff = backoff; a.backoff_f = backoff_f;
    a.flags = flags; a.skip_mask = skip_mask;
    a.dX = dX; a.dU = dU; a.mu = mu; a.mu_f = mu_f; a.cost = cost; a.iters = iters; a.status = status;
    a.diag_cost = 1;
    switch (cfg->model_id) {
        case RNMPC_PENDULUM: return run_qp_mma<4, 1>(cfg, a);
        case RNMPC_QUADROTOR: return run_qp_mma<13, 4>(cfg, a);
        case RNMPC_ROCKET: return run_qp_mma<17, 4>(cfg, a);
    }
    return -1;
}

template <int NX, int NU>
static int run_sls_mma(const rnmpc_config* cfg, SlsArgs a, int warps) {
    Params<NX, NU> hp;
    fill_params<NX, NU>(cfg, hp);
    int dense = 0;
    for (int i = 0; i < NX; ++i)
        for (int j = 0; j < NX; ++j)
            if (i != j && hp.Qreg[i * NX + j] != 0.0) dense = 1;
    const long long items = (long long)a.B * (a.N / 2 + 1);
    const int grid = (int)((items + warps - 1) / warps);
    if ((size_t)warps * SlsMma<NX, NU>::PER_WARP * sizeof(double) > sizeof(smem_raw)) return -4;
    for (int blk = 0; blk < grid; ++blk)
        simt::run_block(warps * 32, dim3(blk), dim3(grid), [&]() { sls_mma_kernel<NX, NU>(a, &hp, dense); });
    return 0;
}

// soft-constrained comparator QP (qp_warp_kernel<.., SOFT = true>) on given linearisation data
template <int NX, int NU>
static int run_qp_soft(const rnmpc_config* cfg, QpArgs a) {
    Params<NX, NU> hp;
    fill_params<NX, NU>(cfg, hp);
    if ((qp_warp_smem_doubles<NX, NU>(a.N) + qp_warp_soft_extra_doubles<NX, NU>(a.N)) * sizeof(double) > sizeof(smem_raw)) return -4;
    std::vector<double> Kw((size_t)a.B * a.N * NU * QpWarp<NX, NU>::LDA, 0.0), Lw((size_t)a.B * a.N * NU * NU, 0.0);
    a.Kq = Kw.data(); a.Lq = Lw.data();
    for (int b = 0; b < a.B; ++b) {
        simt::cpa_reset();
        simt::run_block(32, dim3(b), dim3(a.B), [&]() { qp_warp_kernel<NX, NU, 0, true>(a, &hp); });
    }
    return 0;
}
extern "C" int emu_qp_soft(const rnmpc_config* cfg, int B, const double* A_padded, const double* Bm, const double* c, const double* X,
                           const double* U, const double* x0, double rho, double rho1, int max_iter, double tol_mu, double* dX, double* dU,
                           double* mu, double* mu_f, double* gamma, double* gamma_f, double* cost, int* iters, int* status) {
    QpArgs a{};
    a.B = B; a.N = cfg->N; a.tightened = 0; a.max_iter = max_iter; a.tol_mu = tol_mu; a.tol_res = 1e-9; a.mu0 = -0.3;
    a.A = A_padded; a.Bm = Bm; a.c = c; a.Xnom = X; a.Unom = U; a.x0 = x0;
    a.dX = dX; a.dU = dU; a.mu = mu; a.mu_f = mu_f; a.cost = cost; a.iters = iters; a.status = status;
    a.diag_cost = 1; a.rho = rho; a.rho1 = rho1; a.gamma = gamma; a.gamma_f = gamma_f;
    switch (cfg->model_id) {
        case RNMPC_PENDULUM: return run_qp_soft<4, 1>(cfg, a);
        case RNMPC_QUADROTOR: return run_qp_soft<13, 4>(cfg, a);
        case RNMPC_ROCKET: return run_qp_soft<17, 4>(cfg, a);
    }
    return -1;
}

// fast-SLS pass with explicit weights (eta, eta_f) or from duals (mu, mu_f [+ beta_in, beta_f_in]); A padded [B][N][nx][ld]
extern "C" int emu_sls_mma(const rnmpc_config* cfg, int B, int warps, const double* A_padded, const double* Bm, const double* eta,
                           const double* eta_f, const double* mu, const double* mu_f, const double* beta_in, const double* beta_f_in,
                           double* K, double* S, double* Phi_x, double* Phi_u, double* beta, double* beta_f) {
    SlsArgs a{};
    a.B = B; a.N = cfg->N; a.A = A_padded; a.Bm = Bm; a.eta = eta; a.eta_f = eta_f; a.mu = mu; a.mu_f = mu_f;
    a.beta_in = beta_in; a.beta_f_in = beta_f_in;
    a.K = K; a.S = S; a.Phi_x = Phi_x; a.Phi_u = Phi_u; a.beta = beta; a.beta_f = beta_f;
    switch (cfg->model_id) {
        case RNMPC_PENDULUM: return run_sls_mma<4, 1>(cfg, a, warps);
        case RNMPC_QUADROTOR: return run_sls_mma<13, 4>(cfg, a, warps);
        case RNMPC_ROCKET: return run_sls_mma<17, 4>(cfg, a, warps);
    }
    return -1;
}
