// FP64 micro-benchmarks for B200 (sm_100a): DFMA peak, DMMA m8n8k4 rate, and DFMA fed by
// broadcast shared-memory loads (the "lane owns a row, operand broadcast from smem" pattern
// used by the fast-SLS / Riccati kernels).  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC, int ACTIVE>
__global__ void dfma_peak(double* out, int iters, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    if ((threadIdx.x & 31) < ACTIVE) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA with operands broadcast from shared memory: per k-step load NLD doubles (as LDS.128 pairs) and
// do ROWS*NLD FMAs (ROWS register-resident "rows" per lane).
template <int ROWS, int NC>
__global__ void dfma_smem_bcast(double* out, int iters) {
    __shared__ __align__(16) double sm[17 * 24];
    for (int i = threadIdx.x; i < 17 * 24; i += blockDim.x) sm[i] = 1.0 / (1 + i);
    __syncthreads();
    unsigned smbase = (unsigned)__cvta_generic_to_shared(sm);
    double acc[ROWS][NC];
    double row[ROWS][17];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
#pragma unroll
        for (int c = 0; c < NC; ++c) acc[r][c] = 0;
#pragma unroll
        for (int m = 0; m < 17; ++m) row[r][m] = threadIdx.x * 1e-3 + m + r;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < 17; ++m) {
#pragma unroll
            for (int c = 0; c < NC; c += 2) {
                double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(smbase + (unsigned)((m * 24 + c) * 8)));
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    acc[r][c] = fma(row[r][m], v.x, acc[r][c]);
                    acc[r][c + 1] = fma(row[r][m], v.y, acc[r][c + 1]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) row[r][0] += 1e-9 * acc[r][0];
    }
    double s = 0;
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int c = 0; c < NC; ++c) s += acc[r][c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_peak(double* out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-4;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 4096;
    for (int bps : {1, 2, 4, 8}) {
        int blocks = sms * bps, thr = 256;
        float ms = time_ms([&] { dfma_peak<16, 32><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * iters * (double)blocks * thr;
        printf("dfma_peak acc16 all-lanes blocks/SM=%d thr=%d: %.3f ms %.2f TFLOP/s\n", bps, thr, ms, fl / ms * 1e-9);
    }
    {
        int blocks = sms * 4, thr = 256;
        float ms = time_ms([&] { dfma_peak<16, 17><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); }, 5);
        double fl = 2.0 * 16 * iters * (double)blocks * thr * 17.0 / 32.0;
        printf("dfma_peak 17/32 lanes: %.3f ms %.2f useful TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { dfma_peak<8, 32><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); }, 5);
        fl = 2.0 * 8 * iters * (double)blocks * thr;
        printf("dfma_peak acc8: %.3f ms %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
        ms = time_ms([&] { dfma_peak<4, 32><<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); }, 5);
        fl = 2.0 * 4 * iters * (double)blocks * thr;
        printf("dfma_peak acc4: %.3f ms %.2f TFLOP/s\n", ms, fl / ms * 1e-9);
    }
    for (int bps : {1, 2, 4}) {
        int blocks = sms * bps, thr = 128;
        int it2 = 256;
        float ms = time_ms([&] { dfma_smem_bcast<1, 22><<<blocks, thr>>>(out, it2); }, 5);
        double fl = 2.0 * 17 * 22 * it2 * (double)blocks * thr;
        printf("dfma_smem_bcast rows=1 nc=22 blocks/SM=%d thr=%d: %.3f ms %.2f TFLOP/s\n", bps, thr, ms, fl / ms * 1e-9);
        ms = time_ms([&] { dfma_smem_bcast<2, 22><<<blocks, thr>>>(out, it2); }, 5);
        fl = 2.0 * 2 * 17 * 22 * it2 * (double)blocks * thr;
        printf("dfma_smem_bcast rows=2 nc=22 blocks/SM=%d thr=%d: %.3f ms %.2f TFLOP/s\n", bps, thr, ms, fl / ms * 1e-9);
    }
    for (int bps : {1, 2, 4, 8}) {
        int blocks = sms * bps, thr = 128;
        float ms = time_ms([&] { dmma_peak<<<blocks, thr>>>(out, iters); }, 5);
        double fl = 2.0 * 8 * 8 * 4 * 8 * iters * (double)blocks * (thr / 32);
        printf("dmma m8n8k4 blocks/SM=%d thr=%d: %.3f ms %.2f TFLOP/s\n", bps, thr, ms, fl / ms * 1e-9);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
