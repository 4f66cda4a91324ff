// Minimal SIMT emulator (host, test infrastructure only): runs ONE thread block of an unmodified CUDA kernel source on the CPU,
// one fiber (ucontext) per thread, so that the kernels of csrc/ can be checked against the oracle where no GPU exists (the
// build container).  Supported: threadIdx / blockIdx / blockDim / gridDim, __syncthreads, __syncwarp, __shfl_xor_sync,
// __shfl_sync, __ballot_sync, __all_sync, __any_sync, atomicAdd, extern __shared__ (one block at a time), the warp-collective
// mma.sync.m8n8k4.f64 through the wrapper rnmpc::dmma884, cp.async through rnmpc::cpa16 / cpa_commit / cpa_wait (deferred until the wait), the
// usual math functions.  Not supported: static __shared__.
// Threads of a warp that call a collective block until every lane of the mask has arrived (two-phase exchange), so divergent
// control flow around a collective is detected as a deadlock instead of being silently accepted.
// The product never includes this file: csrc/*.cuh see it only under -DRNMPC_EMU (tools/simt_emu/Makefile).
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; } __attribute__((aligned(16)));
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
typedef void* cudaStream_t;

namespace simt {
struct Thread {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    bool done = false;
    int wait_kind = 0;            // 0 runnable, 1 block barrier, 2 warp barrier
    unsigned wait_mask = 0;
};
struct Block {
    std::vector<Thread> th;
    ucontext_t sched;
    int cur = 0;
    int nthreads = 0;
    // warp exchange slots
    struct Slot { double d[4]; unsigned u; int pred; };
    std::vector<Slot> slots;
    std::function<void()> body;
};
extern Block* g_blk;
extern dim3 g_blockIdx, g_gridDim, g_blockDim;
extern int g_order;

inline int tid() { return g_blk->cur; }
void yield_block_barrier();
void yield_warp_barrier(unsigned mask);
void run_block(int nthreads, dim3 bidx, dim3 gdim, const std::function<void()>& body, size_t stack_bytes = (size_t)1 << 20);
}  // namespace simt

// threadIdx.x etc. as objects with implicit conversion
struct EmuThreadIdx { struct { operator unsigned() const { return (unsigned)simt::tid(); } } x; unsigned y = 0, z = 0; };
struct EmuBlockIdx { struct { operator unsigned() const { return simt::g_blockIdx.x; } } x; unsigned y = 0, z = 0; };
struct EmuBlockDim { struct { operator unsigned() const { return simt::g_blockDim.x; } } x; unsigned y = 1, z = 1; };
struct EmuGridDim { struct { operator unsigned() const { return simt::g_gridDim.x; } } x; unsigned y = 1, z = 1; };
static EmuThreadIdx threadIdx;
static EmuBlockIdx blockIdx;
static EmuBlockDim blockDim;
static EmuGridDim gridDim;

// the kernels declare `extern __shared__ unsigned char smem_raw[];` at block scope inside namespace rnmpc: the driver defines
// rnmpc::smem_raw (one block runs at a time)

static inline void __syncthreads() { simt::yield_block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::yield_warp_barrier(mask); }

template <class T>
static inline T emu_exchange(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "");
    simt::Block* b = simt::g_blk;
    const int me = b->cur, w0 = me & ~31;
    std::memcpy(&b->slots[me].d[0], &v, sizeof(T));
    simt::yield_warp_barrier(mask);
    T r;
    std::memcpy(&r, &b->slots[w0 + (src_lane & 31)].d[0], sizeof(T));
    simt::yield_warp_barrier(mask);
    return r;
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o) { return emu_exchange(mask, v, (simt::tid() & 31) ^ o); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { return emu_exchange(mask, v, src); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    simt::Block* b = simt::g_blk;
    const int me = b->cur, w0 = me & ~31;
    b->slots[me].pred = pred ? 1 : 0;
    simt::yield_warp_barrier(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) { if (w0 + l < b->nthreads && !b->th[w0 + l].done && b->slots[w0 + l].pred) r |= 1u << l; }
    simt::yield_warp_barrier(mask);
    return r;
}
static inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == (mask & __ballot_sync(mask, 1)); }
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int atomicAdd(int* p, int v) { int o = *p; *p += v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline long long clock64() { return 0; }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::fma; using std::fmax; using std::fmin; using std::fabs; using std::sqrt;

// cp.async emulation: a copy is performed when its group is waited for (not when it is issued), per emulated thread
namespace simt {
struct CpaOp { double* d; const double* s; };
inline std::vector<std::vector<CpaOp>> g_cpa_groups[1024];
inline std::vector<CpaOp> g_cpa_open[1024];
inline void cpa_reset() { for (int i = 0; i < 1024; ++i) { g_cpa_groups[i].clear(); g_cpa_open[i].clear(); } }
}  // namespace simt
namespace rnmpc {
static inline void cpa16(double* dst, const double* src) { simt::g_cpa_open[simt::tid()].push_back({dst, src}); }
static inline void cpa_commit() { const int t = simt::tid(); simt::g_cpa_groups[t].push_back(simt::g_cpa_open[t]); simt::g_cpa_open[t].clear(); }
template <int PENDING>
static inline void cpa_wait() {
    auto& g = simt::g_cpa_groups[simt::tid()];
    while ((int)g.size() > PENDING) {
        for (auto& op : g.front()) { op.d[0] = op.s[0]; op.d[1] = op.s[1]; }
        g.erase(g.begin());
    }
}
static inline void fence_block() {}
// bulk copies + mbarrier: the copy happens at issue, the wait is a no-op (the protocol's barriers are not checked by the emulator)
static inline void mbar_init(unsigned long long* bar, int) { *bar = 0; }
static inline void mbar_init_fence() {}
static inline void mbar_expect_tx(unsigned long long*, unsigned) {}
static inline void bulk_g2s(double* dst, const double* src, unsigned bytes, unsigned long long*) { std::memcpy(dst, src, bytes); }
static inline void mbar_wait(unsigned long long*, unsigned) {}
static inline void bulk_g2s_elect(double* dst, const double* src, unsigned bytes, unsigned long long*) { if ((threadIdx.x & 31) == 0) std::memcpy(dst, src, bytes); }
static inline unsigned uniform_u32(const unsigned v) { return v; }
static inline void fence_proxy_async() {}
// tensor memory: 128 lanes x 512 columns of 32 bits per (emulated) SM = per block; address = (lane << 16) | column
}  // namespace rnmpc
namespace simt { inline unsigned g_tmem[128][512]; }
namespace rnmpc {
static inline unsigned double_lo(const double v) { unsigned long long u; std::memcpy(&u, &v, 8); return (unsigned)(u & 0xffffffffu); }
static inline unsigned double_hi(const double v) { unsigned long long u; std::memcpy(&u, &v, 8); return (unsigned)(u >> 32); }
static inline double hilo2double(const unsigned hi, const unsigned lo) { unsigned long long u = ((unsigned long long)hi << 32) | lo; double v; std::memcpy(&v, &u, 8); return v; }
static inline void tmem_alloc_256(unsigned* slot) { *slot = 0; }
static inline void tmem_dealloc_256(unsigned) {}
static inline void tmem_fence_before_sync() {}
static inline void tmem_fence_after_sync() {}
static inline unsigned tmem_warp_base(const unsigned base, const int warp_in_cta) { return base + ((unsigned)(32 * (warp_in_cta & 3)) << 16); }
template <int NREG> static inline void tm_ld_raw(const unsigned taddr, unsigned (&r)[NREG]) {
    const int lane = (int)(taddr >> 16) + (simt::tid() & 31), col = (int)(taddr & 0xffffu);
    if (lane >= 128 || col + NREG > 512) { std::fprintf(stderr, "tmem load out of range\n"); std::abort(); }
    for (int i = 0; i < NREG; ++i) r[i] = simt::g_tmem[lane][col + i];
}
template <int NREG> static inline void tm_st_raw(const unsigned taddr, const unsigned (&r)[NREG]) {
    const int lane = (int)(taddr >> 16) + (simt::tid() & 31), col = (int)(taddr & 0xffffu);
    if (lane >= 128 || col + NREG > 512) { std::fprintf(stderr, "tmem store out of range\n"); std::abort(); }
    for (int i = 0; i < NREG; ++i) simt::g_tmem[lane][col + i] = r[i];
}
static inline void tm_wait_st() {}
}  // namespace rnmpc
namespace rnmpc {
// D (8x8) += A (8x4) B (4x8) across the 32 lanes of the calling warp (fragment layout of mma.sync.m8n8k4.f64)
static inline void dmma884(double& c0, double& c1, const double a, const double bb) {
    simt::Block* b = simt::g_blk;
    const int me = b->cur, w0 = me & ~31, lane = me & 31, g = lane >> 2, t = lane & 3;
    b->slots[me].d[0] = a; b->slots[me].d[1] = bb;
    simt::yield_warp_barrier(0xffffffffu);
    double r0 = c0, r1 = c1;
    for (int k = 0; k < 4; ++k) {
        const double av = b->slots[w0 + g * 4 + k].d[0];                 // A[g][k]
        r0 = std::fma(av, b->slots[w0 + (2 * t) * 4 + k].d[1], r0);      // B[k][2t]
        r1 = std::fma(av, b->slots[w0 + (2 * t + 1) * 4 + k].d[1], r1);  // B[k][2t+1]
    }
    simt::yield_warp_barrier(0xffffffffu);
    c0 = r0; c1 = r1;
}
static inline double fast_rcp(const double x) { return 1.0 / x; }
template <int NEWTON> static inline double rough_rcp(const double x) { return 1.0 / x; }
}  // namespace rnmpc
