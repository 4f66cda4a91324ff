// C ABI of librnmpc_b200.so (declared in include/rnmpc_b200.h).
#include <chrono>
#include <mutex>
#include <vector>

#include "rnmpc_engine.cuh"

namespace rnmpc {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

// DFMA peak probe: 16 independent accumulators per thread
__global__ void fp64_peak_kernel(double* out, int iters, double a, double b) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// FP64 tensor-pipe probe: 8 independent mma.sync.m8n8k4.f64 accumulation chains per warp
__global__ void fp64_mma_peak_kernel(double* out, int iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace rnmpc

using namespace rnmpc;

// Host-buffer path of small batches (rnmpc_solve_host, the single-solve latency path): persistent device / pinned-host staging, and
// the whole call -- input copy, every launch of the solve, output copies -- replayed as ONE CUDA graph on the handle's own stream.
struct HostPath {
    static constexpr int kMaxB = 64;
    cudaStream_t hs = nullptr;
    double* d_x0 = nullptr;
    char* h_pin = nullptr;          // [x0 | X | U | backoff | backoff_f | flags | status] for kMaxB instances
    size_t off[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaGraphExec_t exec = nullptr;
    int exec_B = 0;
    long long exec_epoch = -1, exec_launches = 0;
};
struct rnmpc_handle {
    IEngine* e;
    int nx, nu, N, cap;
    HostPath hp;
};

#define RN_API_CUDA(call)                                                                      \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                     \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

extern "C" {

const char* rnmpc_last_error(void) { return g_err.c_str(); }

int rnmpc_model_dims(int model_id, int* nx, int* nu) {
    switch (model_id) {
        case RNMPC_PENDULUM: *nx = PendulumModel::NX; *nu = PendulumModel::NU; return 0;
        case RNMPC_QUADROTOR: *nx = QuadrotorModel::NX; *nu = QuadrotorModel::NU; return 0;
        case RNMPC_ROCKET: *nx = RocketModel::NX; *nu = RocketModel::NU; return 0;
    }
    set_error("unknown model id");
    return -1;
}

int rnmpc_create(rnmpc_handle** out, const rnmpc_config* cfg, int capacity, int device) {
    if (!out || !cfg) { set_error("null argument"); return -1; }
    if (cfg->N < 1 || cfg->rti < 1 || cfg->sls_steps < 0 || capacity < 1) {
        set_error("N, rti and capacity must be >= 1, sls_steps >= 0");
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: rnmpc_b200 has no CPU fallback");
        return -3;
    }
    if (device < 0 || device >= ndev) { set_error("bad device ordinal"); return -1; }
    IEngine* e = nullptr;
    switch (cfg->model_id) {
        case RNMPC_PENDULUM: e = make_pendulum(cfg, capacity, device); break;
        case RNMPC_QUADROTOR: e = make_quadrotor(cfg, capacity, device); break;
        case RNMPC_ROCKET: e = make_rocket(cfg, capacity, device); break;
        default: set_error("unknown model id"); return -1;
    }
    if (!e) return -2;
    auto* h = new rnmpc_handle();
    h->e = e;
    e->dims(&h->nx, &h->nu, &h->N, &h->cap);
    *out = h;
    return 0;
}

int rnmpc_destroy(rnmpc_handle* h) {
    if (!h) return -1;
    cudaSetDevice(h->e->device);
    if (h->hp.exec) cudaGraphExecDestroy(h->hp.exec);
    if (h->hp.hs) cudaStreamDestroy(h->hp.hs);
    if (h->hp.d_x0) cudaFree(h->hp.d_x0);
    if (h->hp.h_pin) cudaFreeHost(h->hp.h_pin);
    delete h->e;
    delete h;
    return 0;
}

#define RN_ENTER(h)                                  \
    if (!(h)) { set_error("null handle"); return -1; } \
    RN_API_CUDA(cudaSetDevice((h)->e->device));

int rnmpc_ddyn(rnmpc_handle* h, int B, const double* x, const double* u, const double* w, double step, double* xp, void* stream) {
    RN_ENTER(h);
    return h->e->ddyn(B, x, u, w, step, xp, (cudaStream_t)stream);
}
int rnmpc_linearize(rnmpc_handle* h, int B, const double* X, const double* U, double* A, double* Bm, double* c, void* stream) {
    RN_ENTER(h);
    return h->e->linearize(B, X, U, A, Bm, c, (cudaStream_t)stream);
}
int rnmpc_sls(rnmpc_handle* h, int B, const double* A, const double* Bm, const double* eta, const double* eta_f, double* K, double* S,
              double* Phi_x, double* Phi_u, double* beta, double* beta_f, double* backoff, double* backoff_f, void* stream) {
    RN_ENTER(h);
    if (!A || !Bm || !eta || !eta_f || !K || !beta || !beta_f || !backoff || !backoff_f) { set_error("null buffer"); return -1; }
    return h->e->sls(B, A, Bm, eta, eta_f, K, S, Phi_x, Phi_u, beta, beta_f, backoff, backoff_f, (cudaStream_t)stream);
}
int rnmpc_qp(rnmpc_handle* h, int B, int tightened, const double* A, const double* Bm, const double* c, const double* X, const double* U,
             const double* x0, const double* backoff, const double* backoff_f, double* dX, double* dU, double* mu, double* mu_f,
             double* cost, int* iters, int* status, void* stream) {
    RN_ENTER(h);
    if (tightened && (!backoff || !backoff_f)) { set_error("tightened QP needs back-offs"); return -1; }
    return h->e->qp(B, tightened, A, Bm, c, X, U, x0, backoff, backoff_f, dX, dU, mu, mu_f, cost, iters, status, (cudaStream_t)stream);
}
int rnmpc_qp_general(rnmpc_handle* h, int B, const double* A, const double* Bm, const double* c, const double* hi, const double* lo,
                     const double* q, const double* x0pin, double* dX, double* dU, double* mu, double* mu_f, double* cost, int* iters,
                     int* status, void* stream) {
    RN_ENTER(h);
    if (!A || !Bm || !c || !hi || !lo || !x0pin || !dX || !dU || !mu || !mu_f || !cost || !iters || !status) { set_error("null buffer"); return -1; }
    return h->e->qp_general(B, A, Bm, c, hi, lo, q, x0pin, dX, dU, mu, mu_f, cost, iters, status, (cudaStream_t)stream);
}
int rnmpc_sls_duals(rnmpc_handle* h, int B, const double* A, const double* Bm, const double* mu, const double* mu_f, const double* beta_in,
                    const double* beta_f_in, double* K, double* S, double* Phi_x, double* Phi_u, double* beta, double* beta_f,
                    double* backoff, double* backoff_f, void* stream) {
    RN_ENTER(h);
    if (!A || !Bm || !mu || !mu_f || !K || !beta || !beta_f || !backoff || !backoff_f) { set_error("null buffer"); return -1; }
    if ((beta_in == nullptr) != (beta_f_in == nullptr)) { set_error("beta_in and beta_f_in go together"); return -1; }
    return h->e->sls_duals(B, A, Bm, mu, mu_f, beta_in, beta_f_in, K, S, Phi_x, Phi_u, beta, beta_f, backoff, backoff_f, (cudaStream_t)stream);
}
int rnmpc_costates(rnmpc_handle* h, int B, double* lam, void* stream) {
    RN_ENTER(h);
    if (!lam) { set_error("null buffer"); return -1; }
    return h->e->costates(B, lam, (cudaStream_t)stream);
}
int rnmpc_soft_nmpc(rnmpc_handle* h, int B, const double* x0, double* X, double* U, double rho, double rho1, int max_iter, double tol,
                    double* gamma, double* gamma_f, int* sqp_iters, int* status, void* stream) {
    RN_ENTER(h);
    if (!x0 || !X || !U || !sqp_iters || !status) { set_error("null buffer"); return -1; }
    if (!(rho > 0.0) || rho1 < 0.0 || max_iter < 1 || !(tol > 0.0)) { set_error("soft_nmpc: rho > 0, rho1 >= 0, max_iter >= 1, tol > 0"); return -1; }
    if ((gamma == nullptr) != (gamma_f == nullptr)) { set_error("gamma and gamma_f go together"); return -1; }
    return h->e->soft_nmpc(B, x0, X, U, rho, rho1, max_iter, tol, gamma, gamma_f, sqp_iters, status, (cudaStream_t)stream);
}
int rnmpc_set_nominal(rnmpc_handle* h, int B, const double* X, const double* U, void* stream) {
    RN_ENTER(h);
    return h->e->set_nominal(B, X, U, (cudaStream_t)stream);
}
int rnmpc_init_nominal(rnmpc_handle* h, int B, const double* x0, int iters, void* stream) {
    RN_ENTER(h);
    return h->e->init_nominal(B, x0, iters, (cudaStream_t)stream);
}
int rnmpc_solve(rnmpc_handle* h, int B, const double* x0, void* stream) {
    RN_ENTER(h);
    return h->e->solve(B, x0, (cudaStream_t)stream);
}
int rnmpc_reset_warm_start(rnmpc_handle* h, int B, void* stream) {
    RN_ENTER(h);
    return h->e->reset_warm_start(B, (cudaStream_t)stream);
}
int rnmpc_plant_step(rnmpc_handle* h, int B, double* x, const double* w, double* u0, void* stream) {
    RN_ENTER(h);
    return h->e->plant_step(B, x, w, u0, (cudaStream_t)stream);
}
int rnmpc_get_buffer(rnmpc_handle* h, const char* name, void** ptr, long long* bytes) {
    if (!h || !name || !ptr) { set_error("null argument"); return -1; }
    return h->e->get_buffer(name, ptr, bytes);
}
int rnmpc_enable_phi(rnmpc_handle* h, int enable) {
    RN_ENTER(h);
    return h->e->enable_phi(enable);
}
int rnmpc_copy_buffer(rnmpc_handle* h, const char* name, void* dst, long long n_bytes) {
    RN_ENTER(h);
    void* p; long long n;
    if (!dst || h->e->get_buffer(name, &p, &n)) return -1;
    if (n_bytes > n) { set_error("copy larger than buffer"); return -1; }
    RN_API_CUDA(cudaDeviceSynchronize());
    RN_API_CUDA(cudaMemcpy(dst, p, (size_t)n_bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int rnmpc_enable_timing(rnmpc_handle* h, int enable) {
    RN_ENTER(h);
    return h->e->enable_timing(enable);
}
int rnmpc_get_timings(rnmpc_handle* h, double* out3) {
    RN_ENTER(h);
    if (!out3) return -1;
    return h->e->get_timings(out3);
}
int rnmpc_kernel_time(rnmpc_handle* h, const char* kernel, double* total_ms, long long* launches) {
    RN_ENTER(h);
    if (!kernel || !total_ms || !launches) return -1;
    return h->e->kernel_time(kernel, total_ms, launches);
}
int rnmpc_set_option(rnmpc_handle* h, const char* name, double value) {
    RN_ENTER(h);
    if (!name) return -1;
    return h->e->set_option(name, value);
}
long long rnmpc_launch_count(rnmpc_handle* h) { return h ? h->e->launches : -1; }

static int copy_out(rnmpc_handle* h, const char* name, void* dst, size_t bytes, cudaStream_t st) {
    if (!dst) return 0;
    void* p; long long n;
    if (h->e->get_buffer(name, &p, &n)) return -1;
    RN_API_CUDA(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyDeviceToHost, st));
    return 0;
}

static int host_path_init(rnmpc_handle* h) {
    HostPath& hp = h->hp;
    if (hp.hs) return 0;
    const int nx = h->nx, nu = h->nu, N = h->N, ni = 2 * (nx + nu), K = HostPath::kMaxB;
    const size_t sz[7] = {(size_t)K * nx * 8, (size_t)K * (N + 1) * nx * 8, (size_t)K * N * nu * 8, (size_t)K * N * ni * 8,
                          (size_t)K * 2 * nx * 8, (size_t)K * 4, (size_t)K * 4};
    for (int i = 0; i < 7; ++i) hp.off[i + 1] = hp.off[i] + ((sz[i] + 15) & ~(size_t)15);
    RN_API_CUDA(cudaStreamCreate(&hp.hs));        // a blocking stream: ordered with the caller's work on the default stream
    RN_API_CUDA(cudaMalloc((void**)&hp.d_x0, sz[0]));
    RN_API_CUDA(cudaMallocHost((void**)&hp.h_pin, hp.off[7]));
    return 0;
}

// input copy + solve + output copies of B <= kMaxB instances, enqueued on the handle's stream (also the body of the captured graph)
static int host_path_enqueue(rnmpc_handle* h, int B) {
    HostPath& hp = h->hp;
    const int nx = h->nx, nu = h->nu, N = h->N, ni = 2 * (nx + nu);
    RN_API_CUDA(cudaMemcpyAsync(hp.d_x0, hp.h_pin + hp.off[0], (size_t)B * nx * 8, cudaMemcpyHostToDevice, hp.hs));
    int rc = h->e->solve(B, hp.d_x0, hp.hs);
    if (rc) return rc;
    const char* names[6] = {"X", "U", "backoff", "backoff_f", "flags", "status"};
    const size_t bytes[6] = {(size_t)B * (N + 1) * nx * 8, (size_t)B * N * nu * 8, (size_t)B * N * ni * 8, (size_t)B * 2 * nx * 8, (size_t)B * 4, (size_t)B * 4};
    for (int i = 0; i < 6; ++i) {
        void* p; long long n;
        if (h->e->get_buffer(names[i], &p, &n)) return -1;
        RN_API_CUDA(cudaMemcpyAsync(hp.h_pin + hp.off[i + 1], p, bytes[i], cudaMemcpyDeviceToHost, hp.hs));
    }
    return 0;
}

int rnmpc_solve_host(rnmpc_handle* h, int B, const double* h_x0, double* h_X, double* h_U, double* h_backoff, double* h_backoff_f,
                     int* h_flags, int* h_status) {
    RN_ENTER(h);
    if (!h_x0 || B <= 0 || B > h->cap) { set_error("bad arguments"); return -1; }
    const int nx = h->nx, nu = h->nu, N = h->N, ni = 2 * (nx + nu);
    if (B <= HostPath::kMaxB) {
        HostPath& hp = h->hp;
        int rc = host_path_init(h);
        if (rc) return rc;
        memcpy(hp.h_pin + hp.off[0], h_x0, (size_t)B * nx * 8);
        if (h->e->graphable()) {
            if (hp.exec == nullptr || hp.exec_B != B || hp.exec_epoch != h->e->epoch) {
                if (hp.exec) { cudaGraphExecDestroy(hp.exec); hp.exec = nullptr; }
                const long long l0 = h->e->launches;
                cudaGraph_t graph = nullptr;
                RN_API_CUDA(cudaStreamBeginCapture(hp.hs, cudaStreamCaptureModeThreadLocal));
                rc = host_path_enqueue(h, B);
                cudaError_t ce = cudaStreamEndCapture(hp.hs, &graph);
                hp.exec_launches = h->e->launches - l0;
                h->e->launches = l0;
                if (rc || ce != cudaSuccess || graph == nullptr) {
                    if (graph) cudaGraphDestroy(graph);
                    cudaGetLastError();
                    if (rc) return rc;
                    set_error(std::string("solve_host: graph capture failed: ") + cudaGetErrorString(ce));
                    return -2;
                }
                ce = cudaGraphInstantiate(&hp.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess) { hp.exec = nullptr; set_error(std::string("solve_host: cudaGraphInstantiate: ") + cudaGetErrorString(ce)); return -2; }
                hp.exec_B = B; hp.exec_epoch = h->e->epoch;
            }
            RN_API_CUDA(cudaGraphLaunch(hp.exec, hp.hs));
            h->e->launches += hp.exec_launches;
        } else {
            rc = host_path_enqueue(h, B);
            if (rc) { cudaStreamSynchronize(hp.hs); return rc; }
        }
        cudaError_t e = cudaStreamSynchronize(hp.hs);
        if (e != cudaSuccess) { set_error(std::string("solve: ") + cudaGetErrorString(e)); return -2; }
        if (h_X) memcpy(h_X, hp.h_pin + hp.off[1], (size_t)B * (N + 1) * nx * 8);
        if (h_U) memcpy(h_U, hp.h_pin + hp.off[2], (size_t)B * N * nu * 8);
        if (h_backoff) memcpy(h_backoff, hp.h_pin + hp.off[3], (size_t)B * N * ni * 8);
        if (h_backoff_f) memcpy(h_backoff_f, hp.h_pin + hp.off[4], (size_t)B * 2 * nx * 8);
        if (h_flags) memcpy(h_flags, hp.h_pin + hp.off[5], (size_t)B * 4);
        if (h_status) memcpy(h_status, hp.h_pin + hp.off[6], (size_t)B * 4);
        return 0;
    }
    // large batches: direct copies (the staging would only add a host copy)
    double* d_x0;
    RN_API_CUDA(cudaMalloc((void**)&d_x0, (size_t)B * nx * sizeof(double)));
    RN_API_CUDA(cudaMemcpy(d_x0, h_x0, (size_t)B * nx * sizeof(double), cudaMemcpyHostToDevice));
    int rc = h->e->solve(B, d_x0, 0);
    if (!rc) rc = copy_out(h, "X", h_X, (size_t)B * (N + 1) * nx * sizeof(double), 0);
    if (!rc) rc = copy_out(h, "U", h_U, (size_t)B * N * nu * sizeof(double), 0);
    if (!rc) rc = copy_out(h, "backoff", h_backoff, (size_t)B * N * ni * sizeof(double), 0);
    if (!rc) rc = copy_out(h, "backoff_f", h_backoff_f, (size_t)B * 2 * nx * sizeof(double), 0);
    if (!rc) rc = copy_out(h, "flags", h_flags, (size_t)B * sizeof(int), 0);
    if (!rc) rc = copy_out(h, "status", h_status, (size_t)B * sizeof(int), 0);
    cudaError_t e = cudaDeviceSynchronize();
    cudaFree(d_x0);
    if (e != cudaSuccess) { set_error(std::string("solve: ") + cudaGetErrorString(e)); return -2; }
    return rc;
}

int rnmpc_closed_loop_host(rnmpc_handle* h, int B, int steps, const double* h_x0, const double* h_W, const double* h_X0,
                           const double* h_U0, int init_iters, double* h_states, double* h_inputs, long long* h_nfail) {
    RN_ENTER(h);
    if (!h_x0 || B <= 0 || B > h->cap || steps < 1) { set_error("bad arguments"); return -1; }
    const int nx = h->nx, nu = h->nu, N = h->N;
    double *d_x = nullptr, *d_W = nullptr, *d_states = nullptr, *d_inputs = nullptr, *d_tmp = nullptr, *dX0 = nullptr, *dU0 = nullptr;
    cudaStream_t st = 0;
    long long nfail = 0;
    const double sticky_before = h->e->get_option("sticky_fail");
    // one exit: every device buffer is freed and the caller's sticky_fail setting restored on success and on every error path
    auto finish = [&](int rc) {
        cudaStreamSynchronize(st);
        for (double* p : {d_x, d_W, d_states, d_inputs, d_tmp, dX0, dU0}) if (p) cudaFree(p);
        h->e->set_option("sticky_fail", sticky_before);
        if (h_nfail) *h_nfail = nfail;
        return rc;
    };
#define RN_LOOP_CUDA(call)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) { set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return finish(-2); } \
    } while (0)
    RN_LOOP_CUDA(cudaMalloc((void**)&d_x, (size_t)B * nx * sizeof(double)));
    RN_LOOP_CUDA(cudaMalloc((void**)&d_states, (size_t)B * (steps + 1) * nx * sizeof(double)));
    RN_LOOP_CUDA(cudaMalloc((void**)&d_inputs, (size_t)B * steps * nu * sizeof(double)));
    RN_LOOP_CUDA(cudaMemcpyAsync(d_x, h_x0, (size_t)B * nx * sizeof(double), cudaMemcpyHostToDevice, st));
    if (h_W) {
        // [B][steps][nx] on the host -> step-major [steps][B][nx] slices are gathered with a strided 2-D copy per step
        RN_LOOP_CUDA(cudaMalloc((void**)&d_W, (size_t)B * steps * nx * sizeof(double)));
        RN_LOOP_CUDA(cudaMemcpyAsync(d_W, h_W, (size_t)B * steps * nx * sizeof(double), cudaMemcpyHostToDevice, st));
        RN_LOOP_CUDA(cudaMalloc((void**)&d_tmp, (size_t)B * nx * sizeof(double)));
    }
    int rc = 0;
    if (h_X0 && h_U0) {
        RN_LOOP_CUDA(cudaMalloc((void**)&dX0, (size_t)B * (N + 1) * nx * sizeof(double)));
        RN_LOOP_CUDA(cudaMalloc((void**)&dU0, (size_t)B * N * nu * sizeof(double)));
        RN_LOOP_CUDA(cudaMemcpyAsync(dX0, h_X0, (size_t)B * (N + 1) * nx * sizeof(double), cudaMemcpyHostToDevice, st));
        RN_LOOP_CUDA(cudaMemcpyAsync(dU0, h_U0, (size_t)B * N * nu * sizeof(double), cudaMemcpyHostToDevice, st));
        rc = h->e->set_nominal(B, dX0, dU0, st);
    } else {
        rc = h->e->init_nominal(B, d_x, init_iters > 0 ? init_iters : 10, st);
    }
    if (rc) return finish(rc);
    std::vector<int> fl(B), dead(B, 0);
    h->e->set_option("sticky_fail", 1.0);      // Monte-Carlo semantics: a scenario whose solve failed is dead (the reference's driver would stop)
    void* d_flags; long long nb;
    h->e->get_buffer("flags", &d_flags, &nb);
    const size_t srow = (size_t)(steps + 1) * nx * sizeof(double);
    for (int i = 0; i < steps; ++i) {
        RN_LOOP_CUDA(cudaMemcpy2DAsync(d_states + (size_t)i * nx, srow, d_x, nx * sizeof(double), nx * sizeof(double), B,
                                       cudaMemcpyDeviceToDevice, st));
        if (i > 0 && (rc = h->e->reset_warm_start(B, st)) != 0) return finish(rc);
        if ((rc = h->e->solve(B, d_x, st)) != 0) return finish(rc);
        const double* w = nullptr;
        if (h_W) {
            RN_LOOP_CUDA(cudaMemcpy2DAsync(d_tmp, nx * sizeof(double), d_W + (size_t)i * nx, (size_t)steps * nx * sizeof(double),
                                           nx * sizeof(double), B, cudaMemcpyDeviceToDevice, st));
            w = d_tmp;
        }
        if ((rc = h->e->plant_step(B, d_x, w, nullptr, st)) != 0) return finish(rc);
        void* dU; h->e->get_buffer("U", &dU, &nb);
        RN_LOOP_CUDA(cudaMemcpy2DAsync(d_inputs + (size_t)i * nu, (size_t)steps * nu * sizeof(double), dU,
                                       (size_t)N * nu * sizeof(double), nu * sizeof(double), B, cudaMemcpyDeviceToDevice, st));
        if (h_nfail) {
            // failed solves = scenarios that died at THIS step (a dead scenario is not solved again)
            RN_LOOP_CUDA(cudaMemcpyAsync(fl.data(), d_flags, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
            RN_LOOP_CUDA(cudaStreamSynchronize(st));
            for (int b = 0; b < B; ++b)
                if ((fl[b] & 1) && !dead[b]) { dead[b] = 1; ++nfail; }
        }
    }
    RN_LOOP_CUDA(cudaMemcpy2DAsync(d_states + (size_t)steps * nx, srow, d_x, nx * sizeof(double), nx * sizeof(double), B,
                                   cudaMemcpyDeviceToDevice, st));
    if (h_states) RN_LOOP_CUDA(cudaMemcpyAsync(h_states, d_states, (size_t)B * srow, cudaMemcpyDeviceToHost, st));
    if (h_inputs) RN_LOOP_CUDA(cudaMemcpyAsync(h_inputs, d_inputs, (size_t)B * steps * nu * sizeof(double), cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error(std::string("closed_loop: ") + cudaGetErrorString(e)); return finish(-2); }
#undef RN_LOOP_CUDA
    return finish(0);
}

int rnmpc_fp64_peak(int device, double* tflops) {
    if (!tflops) return -1;
    RN_API_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    RN_API_CUDA(cudaGetDeviceProperties(&p, device));
    const int blocks = p.multiProcessorCount * 8, thr = 256, iters = 4096;
    double* out;
    RN_API_CUDA(cudaMalloc((void**)&out, sizeof(double) * blocks * thr));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, thr>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    RN_API_CUDA(cudaGetLastError());
    *tflops = 2.0 * 16 * iters * (double)blocks * thr / (best * 1e-3) * 1e-12;
    return 0;
}

int rnmpc_fp64_mma_peak(int device, double* tflops) {
    if (!tflops) return -1;
    RN_API_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    RN_API_CUDA(cudaGetDeviceProperties(&p, device));
    const int blocks = p.multiProcessorCount * 8, thr = 256, iters = 2048;
    double* out;
    RN_API_CUDA(cudaMalloc((void**)&out, sizeof(double) * blocks * thr));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        fp64_mma_peak_kernel<<<blocks, thr>>>(out, iters, 1e-3, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    RN_API_CUDA(cudaGetLastError());
    // one m8n8k4 = 8*8*4 multiply-adds per warp
    *tflops = 2.0 * 256.0 * 8 * iters * (double)blocks * (thr / 32) / (best * 1e-3) * 1e-12;
    return 0;
}

}  // extern "C"
