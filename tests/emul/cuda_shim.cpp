// TEST INFRASTRUCTURE ONLY -- fiber scheduler behind tests/emul/cuda_shim.h.
#include "cuda_shim.h"

uint3_ threadIdx, blockIdx;
dim3 blockDim, gridDim;
char *emul_dynamic_smem;

namespace {
struct Fiber {
    ucontext_t ctx;
    char *stack;
    bool done;
};
std::vector<Fiber> fibers;
ucontext_t sched_ctx;
int current = -1;
std::function<void()> *cur_body;
const size_t STACK = 512 * 1024;

void trampoline()
{
    (*cur_body)();
    fibers[current].done = true;
    swapcontext(&fibers[current].ctx, &sched_ctx);
}
}

double emul_xchg[64][32];
namespace {
struct Bar { int count; int gen; };
Bar bars[128];
}

// counting barrier over `nthreads` fibers; arrivals after the last one wait for
// the generation to advance (all participants must call with the same id/count)
void emul_barrier(int id, int nthreads)
{
    Bar &b = bars[id];
    int g = b.gen;
    if (++b.count >= nthreads) {
        b.count = 0;
        b.gen++;
        return;
    }
    while (b.gen == g) emul_yield();
}

void emul_yield()
{
    swapcontext(&fibers[current].ctx, &sched_ctx);
}

void emul_run_block(std::function<void()> body, unsigned nthreads)
{
    cur_body = &body;
    memset(bars, 0, sizeof(bars));
    if (fibers.size() < nthreads) {
        size_t old = fibers.size();
        fibers.resize(nthreads);
        for (size_t i = old; i < nthreads; ++i) {
            fibers[i].stack = (char *)malloc(STACK);
        }
    }
    for (unsigned i = 0; i < nthreads; ++i) {
        getcontext(&fibers[i].ctx);
        fibers[i].ctx.uc_stack.ss_sp = fibers[i].stack;
        fibers[i].ctx.uc_stack.ss_size = STACK;
        fibers[i].ctx.uc_link = &sched_ctx;
        fibers[i].done = false;
        makecontext(&fibers[i].ctx, trampoline, 0);
    }
    unsigned remaining = nthreads;
    while (remaining) {
        remaining = 0;
        for (unsigned i = 0; i < nthreads; ++i) {
            if (fibers[i].done) continue;
            current = i;
            threadIdx.x = i % blockDim.x;
            threadIdx.y = (i / blockDim.x) % blockDim.y;
            threadIdx.z = i / (blockDim.x * blockDim.y);
            swapcontext(&sched_ctx, &fibers[i].ctx);
            if (!fibers[i].done) ++remaining;
        }
    }
}
