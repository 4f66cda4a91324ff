// Scheduler of the SIMT emulator (see simt_emu.h).
#include "simt_emu.h"

namespace simt {
Block* g_blk = nullptr;
dim3 g_blockIdx, g_gridDim, g_blockDim;
int g_order = 0;          // scheduling order of the runnable threads: 0 ascending, 1 descending, 2 pseudo-random per round
static unsigned g_rng = 12345u;

static void trampoline() {
    Block* b = g_blk;
    b->body();
    b->th[b->cur].done = true;
    swapcontext(&b->th[b->cur].ctx, &b->sched);
}

void yield_block_barrier() {
    Block* b = g_blk;
    Thread& t = b->th[b->cur];
    t.wait_kind = 1;
    swapcontext(&t.ctx, &b->sched);
}
void yield_warp_barrier(unsigned mask) {
    Block* b = g_blk;
    Thread& t = b->th[b->cur];
    t.wait_kind = 2;
    t.wait_mask = mask;
    swapcontext(&t.ctx, &b->sched);
}

void run_block(int nthreads, dim3 bidx, dim3 gdim, const std::function<void()>& body, size_t stack_bytes) {
    Block blk;
    blk.nthreads = nthreads;
    blk.th.resize(nthreads);
    blk.slots.resize((nthreads + 31) / 32 * 32);
    blk.body = body;
    g_blk = &blk;
    g_blockIdx = bidx; g_gridDim = gdim; g_blockDim = dim3(nthreads);
    for (int i = 0; i < nthreads; ++i) {
        Thread& t = blk.th[i];
        t.stack.resize(stack_bytes);
        getcontext(&t.ctx);
        t.ctx.uc_stack.ss_sp = t.stack.data();
        t.ctx.uc_stack.ss_size = t.stack.size();
        t.ctx.uc_link = &blk.sched;
        makecontext(&t.ctx, trampoline, 0);
    }
    for (;;) {
        bool progressed = false, all_done = true;
        // run every runnable thread until it blocks or finishes
        unsigned off = 0;
        if (g_order == 2) { g_rng = g_rng * 1664525u + 1013904223u; off = g_rng >> 8; }
        for (int j = 0; j < nthreads; ++j) {
            // a missing barrier shows up as a result that depends on this order (tests run all three)
            const int i = g_order == 0 ? j : g_order == 1 ? nthreads - 1 - j : (int)((j * 37u + off) % (unsigned)nthreads);
            Thread& t = blk.th[i];
            if (t.done) continue;
            all_done = false;
            if (t.wait_kind != 0) continue;
            blk.cur = i;
            swapcontext(&blk.sched, &t.ctx);
            progressed = true;
        }
        if (all_done) break;
        // release warp barriers whose masked live lanes have all arrived with the same mask
        for (int w0 = 0; w0 < nthreads; w0 += 32) {
            for (int l = 0; l < 32 && w0 + l < nthreads; ++l) {
                Thread& t = blk.th[w0 + l];
                if (t.done || t.wait_kind != 2) continue;
                const unsigned mask = t.wait_mask;
                bool ok = true;
                for (int m = 0; m < 32 && w0 + m < nthreads; ++m) {
                    if (!((mask >> m) & 1u)) continue;
                    Thread& o = blk.th[w0 + m];
                    if (o.done) continue;
                    if (o.wait_kind != 2 || o.wait_mask != mask) { ok = false; break; }
                }
                if (ok) {
                    for (int m = 0; m < 32 && w0 + m < nthreads; ++m)
                        if (((mask >> m) & 1u) && !blk.th[w0 + m].done) { blk.th[w0 + m].wait_kind = 0; }
                    progressed = true;
                }
            }
        }
        // release the block barrier when every live thread waits on it
        {
            bool ok = true, any = false;
            for (int i = 0; i < nthreads; ++i) {
                Thread& t = blk.th[i];
                if (t.done) continue;
                any = true;
                if (t.wait_kind != 1) { ok = false; break; }
            }
            if (ok && any) {
                for (int i = 0; i < nthreads; ++i) blk.th[i].wait_kind = 0;
                progressed = true;
            }
        }
        if (!progressed) {
            fprintf(stderr, "simt_emu: deadlock (divergent barrier / collective) in block %u\n", bidx.x);
            for (int i = 0; i < nthreads; ++i)
                if (!blk.th[i].done) fprintf(stderr, "  thread %d waits kind %d mask %08x\n", i, blk.th[i].wait_kind, blk.th[i].wait_mask);
            abort();
        }
    }
    g_blk = nullptr;
}
}  // namespace simt
