// C entry points that run the product's kernels (csrc/*.cuh, unmodified source, compiled with -DRNMPC_EMU) block by block on the
// host through the SIMT emulator.  Test infrastructure only: the CPU suite checks kernel logic against the oracle here, the GPU
// suite checks the compiled sm_100a kernels through the C ABI.
#define RNMPC_EMU 1
#include "simt_emu.h"

#include "../../include/rnmpc_b200.h"
#include "../../robust-nonlinear-mpc_b200/csrc/rnmpc_qp_mma.cuh"

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

template <int NX, int NU>
static int run_qp_mma(const rnmpc_config* cfg, QpArgs a) {
    Params<NX, NU> hp;
    fill_params<NX, NU>(cfg, hp);
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
