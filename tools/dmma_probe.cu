// Probes for the DMMA-based Riccati kernels (sm_100a): fragment layout of mma.sync.m8n8k4.f64, dependent-issue
// latencies (DMMA, DFMA, rsqrt, reciprocal seed + Newton, shuffle, LDS, block barrier) measured with clock64 on one warp,
// and a 1-D bulk asynchronous copy (cp.async.bulk + mbarrier) smoke test.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// C (8x8) = A (8x4, row-major) * B (4x8, row-major) with the assumed fragment layout
__global__ void layout_kernel(const double* A, const double* B, double* C) {
    const int l = threadIdx.x, g = l >> 2, t = l & 3;
    double c0 = 0.0, c1 = 0.0;
    dmma(c0, c1, A[g * 4 + t], B[t * 8 + g]);
    C[g * 8 + 2 * t] = c0;
    C[g * 8 + 2 * t + 1] = c1;
}

__global__ void latency_kernel(long long* out, double* sink, int n) {
    __shared__ double sm[1024];
    const int l = threadIdx.x;
    for (int i = l; i < 1024; i += blockDim.x) sm[i] = 1.0 + 1e-3 * i;
    __syncthreads();
    double a = 1.0 + 1e-6 * l, b = 1.0 - 1e-6 * l, c0 = 0.0, c1 = 0.0;
    long long t0, t1;
    // 1. dependent DMMA chain
    t0 = clock64();
    for (int i = 0; i < n; ++i) dmma(c0, c1, a, b);
    t1 = clock64();
    if (l == 0) out[0] = t1 - t0;
    // 2. four independent DMMA chains
    double d0 = 0, d1 = 0, e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { dmma(c0, c1, a, b); dmma(d0, d1, a, b); dmma(e0, e1, a, b); dmma(f0, f1, a, b); }
    t1 = clock64();
    if (l == 0) out[1] = t1 - t0;
    // 3. dependent DFMA chain
    double x = a;
    t0 = clock64();
    for (int i = 0; i < n; ++i) x = fma(x, b, a);
    t1 = clock64();
    if (l == 0) out[2] = t1 - t0;
    // 4. dependent rsqrt chain
    double y = 2.0 + a;
    t0 = clock64();
    for (int i = 0; i < n; ++i) y = rsqrt(y) + 1.5;
    t1 = clock64();
    if (l == 0) out[3] = t1 - t0;
    // 5. reciprocal seed + 2 Newton steps, dependent
    double z = 2.0 + b;
    t0 = clock64();
    for (int i = 0; i < n; ++i) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(z));
        double e = fma(-z, r, 1.0);
        r = fma(r, e, r);
        e = fma(-z, r, 1.0);
        r = fma(r, e, r);
        z = r + 1.5;
    }
    t1 = clock64();
    if (l == 0) out[4] = t1 - t0;
    // 6. dependent shuffle chain (64-bit)
    double s = a;
    t0 = clock64();
    for (int i = 0; i < n; ++i) s = __shfl_xor_sync(0xffffffffu, s, 1) + 1.0;
    t1 = clock64();
    if (l == 0) out[5] = t1 - t0;
    // 7. dependent LDS chain (pointer chasing through shared memory)
    __shared__ int nxt[256];
    for (int i = l; i < 256; i += blockDim.x) nxt[i] = (i * 7 + 3) & 255;
    __syncthreads();
    int p = l;
    t0 = clock64();
    for (int i = 0; i < n; ++i) p = nxt[p];
    t1 = clock64();
    if (l == 0) out[6] = t1 - t0;
    // 8. __syncthreads round trips
    t0 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    t1 = clock64();
    if (l == 0) out[7] = t1 - t0;
    // 9. STS + __syncwarp + LDS broadcast round trip
    double v = a;
    t0 = clock64();
    for (int i = 0; i < n; ++i) { sm[l & 31] = v; __syncwarp(); v = sm[(l + 1) & 31] + 1.0; __syncwarp(); }
    t1 = clock64();
    if (l == 0) out[8] = t1 - t0;
    // 10. sqrt + IEEE division, dependent
    double q = 3.0 + a;
    t0 = clock64();
    for (int i = 0; i < n; ++i) q = 1.0 / sqrt(q) + 1.5;
    t1 = clock64();
    if (l == 0) out[9] = t1 - t0;
    sink[blockIdx.x * blockDim.x + l] = c0 + c1 + d0 + d1 + e0 + e1 + f0 + f1 + x + y + z + s + p + v + q;
}

// 1-D bulk copy global -> shared completing on an mbarrier (TMA unit, no tensor map)
__global__ void bulk_kernel(const double* src, double* dst, int n_doubles) {
    extern __shared__ __align__(128) unsigned char smraw[];
    double* buf = reinterpret_cast<double*>(smraw);
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_s = (unsigned)__cvta_generic_to_shared(&bar);
    const unsigned buf_s = (unsigned)__cvta_generic_to_shared(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned bytes = (unsigned)(n_doubles * 8);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes));
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(buf_s), "l"(src), "r"(bytes), "r"(bar_s) : "memory");
    }
    unsigned ok = 0;
    while (!ok) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_s), "r"(0u) : "memory");
    }
    for (int i = threadIdx.x; i < n_doubles; i += blockDim.x) dst[i] = buf[i] * 2.0;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sms %d\n", p.name, p.multiProcessorCount);
    // layout
    double hA[32], hB[32], hC[64], *dA, *dB, *dC;
    for (int i = 0; i < 32; ++i) { hA[i] = 1.0 + i; hB[i] = 0.5 - 0.25 * i; }
    CK(cudaMalloc(&dA, sizeof hA)); CK(cudaMalloc(&dB, sizeof hB)); CK(cudaMalloc(&dC, sizeof hC));
    CK(cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice));
    layout_kernel<<<1, 32>>>(dA, dB, dC);
    CK(cudaMemcpy(hC, dC, sizeof hC, cudaMemcpyDeviceToHost));
    double err = 0;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 8; ++j) {
            double r = 0;
            for (int k = 0; k < 4; ++k) r += hA[i * 4 + k] * hB[k * 8 + j];
            err = fmax(err, fabs(r - hC[i * 8 + j]));
        }
    printf("dmma layout check: max err %.3e (%s)\n", err, err < 1e-9 ? "OK: A[g][t], B[t][g], C[g][2t..2t+1]" : "MISMATCH");
    // latencies
    long long *dout, hout[10];
    double* sink;
    const int n = 2000;
    CK(cudaMalloc(&dout, sizeof hout)); CK(cudaMalloc(&sink, 8 * 1024));
    const char* names[10] = {"dependent DMMA", "4 independent DMMA chains (per round of 4)", "dependent DFMA", "dependent rsqrt(+add)",
                             "rcp seed + 2 Newton (+add)", "dependent 64-bit shuffle (+add)", "dependent LDS (pointer chase)", "__syncthreads",
                             "STS + syncwarp + LDS + syncwarp (+add)", "1/sqrt IEEE (+add)"};
    for (int threads : {32, 128}) {
        latency_kernel<<<1, threads>>>(dout, sink, n);
        latency_kernel<<<1, threads>>>(dout, sink, n);
        CK(cudaMemcpy(hout, dout, sizeof hout, cudaMemcpyDeviceToHost));
        printf("-- one CTA of %d threads, cycles per operation:\n", threads);
        for (int i = 0; i < 10; ++i) printf("   %-48s %8.1f\n", names[i], (double)hout[i] / n);
    }
    // bulk copy
    {
        const int nd = 306;      // one padded A_k: 17 x 18 doubles = 2448 bytes (multiple of 16)
        double h[512], *src, *dst;
        for (int i = 0; i < nd; ++i) h[i] = i + 0.5;
        CK(cudaMalloc(&src, nd * 8)); CK(cudaMalloc(&dst, nd * 8));
        CK(cudaMemcpy(src, h, nd * 8, cudaMemcpyHostToDevice));
        bulk_kernel<<<1, 128, nd * 8>>>(src, dst, nd);
        CK(cudaDeviceSynchronize());
        double h2[512];
        CK(cudaMemcpy(h2, dst, nd * 8, cudaMemcpyDeviceToHost));
        double e = 0;
        for (int i = 0; i < nd; ++i) e = fmax(e, fabs(h2[i] - 2.0 * h[i]));
        printf("cp.async.bulk 2448 B: max err %.3e (%s)\n", e, e == 0.0 ? "OK" : "MISMATCH");
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
