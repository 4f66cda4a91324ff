// Probe: tensor memory (TMEM) as lane-private, column-indexed storage for fp64 state (tcgen05.alloc / st / ld / dealloc on sm_100a).
//   * a CTA of 4 warps allocates 256 columns; warp w owns lanes 32w..32w+31; lane i of the warp stores doubles at dynamic
//     columns and reads them back (32x32b shape: thread i <-> TMEM lane 32w + i)
//   * two CTAs per SM (100 KB of dynamic shared memory each) both allocate 256 columns
//   * latency of a dependent st -> wait -> ld -> wait round trip and of a dependent ld chain
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tm_st(unsigned taddr, double v) {
    const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ double tm_ld(unsigned taddr) {
    unsigned lo, hi;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128) probe(double* out, long long* cyc, int* err, int ncols_used) {
    extern __shared__ unsigned char dyn[];
    __shared__ unsigned base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((unsigned)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned base = base_s + ((unsigned)(32 * warp) << 16);
    // write pattern: column pair j holds value f(cta, warp, lane, j)
    const int npair = ncols_used / 2;
    for (int j = 0; j < npair; ++j) tm_st(base + 2 * j, 1000.0 * blockIdx.x + 100.0 * warp + lane + 1e-3 * j);
    tm_wait_st();
    int bad = 0;
    // read back in a different (dynamic) order
    for (int jj = 0; jj < npair; ++jj) {
        const int j = (jj * 37 + lane * 0 + 5) % npair;      // warp-uniform dynamic column
        const double v = tm_ld(base + 2 * j);
        if (v != 1000.0 * blockIdx.x + 100.0 * warp + lane + 1e-3 * j) ++bad;
    }
    // read-modify-write of one column, then re-read
    {
        const double v = tm_ld(base + 2 * 3);
        tm_st(base + 2 * 3, v * 2.0);
        tm_wait_st();
        const double w = tm_ld(base + 2 * 3);
        if (w != 2.0 * (1000.0 * blockIdx.x + 100.0 * warp + lane + 1e-3 * 3)) ++bad;
    }
    if (bad) atomicAdd(err, bad);
    // latencies (warp 0 of block 0)
    if (blockIdx.x == 0 && warp == 0) {
        double x = 1.0;
        long long t0 = clock64();
        for (int i = 0; i < 256; ++i) { tm_st(base + 2 * (i & 63), x); tm_wait_st(); x = tm_ld(base + 2 * (i & 63)) + 1.0; }
        long long t1 = clock64();
        unsigned col = 0;
        for (int i = 0; i < 64; ++i) tm_st(base + 2 * i, (double)((i * 7 + 3) & 63));
        tm_wait_st();
        long long t2 = clock64();
        for (int i = 0; i < 256; ++i) { const double v = tm_ld(base + 2 * col); col = (unsigned)v; }
        long long t3 = clock64();
        if (lane == 0) { cyc[0] = (t1 - t0) / 256; cyc[1] = (t3 - t2) / 256; out[0] = x + col; }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(base_s) : "memory");
    if (threadIdx.x == 0) dyn[0] = 1;
}

int main() {
    double* out; long long* cyc; int* err;
    CK(cudaMalloc(&out, 64)); CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&err, 4));
    CK(cudaMemset(err, 0, 4)); CK(cudaMemset(cyc, 0, 64));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    int nb = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, probe, 128, 100 * 1024));
    printf("occupancy (smem/regs only): %d CTAs per SM\n", nb);
    probe<<<148 * 2 * 4, 128, 100 * 1024>>>(out, cyc, err, 224);
    CK(cudaDeviceSynchronize());
    int herr; long long hc[2];
    CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hc, cyc, 16, cudaMemcpyDeviceToHost));
    printf("tmem st/ld round trips: %d mismatches over %d CTAs x 4 warps x 112 column pairs\n", herr, 148 * 8);
    printf("dependent st -> wait::st -> ld -> wait::ld (+add): %lld cycles;  dependent ld chain (ld + wait::ld + cvt): %lld cycles\n", hc[0], hc[1]);
    printf("done\n");
    return 0;
}
