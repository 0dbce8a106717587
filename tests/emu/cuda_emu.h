// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: runs CUDA kernel SOURCE on the CPU so that kernel logic (indexing, warp
// collectives, shared-memory carve-ups) can be checked against the oracle without a GPU.  Nothing under polarcub_b200/
// includes, links or executes this; the product path has no CPU fallback.  Used by tests/emu/*.cpp only.
//
// Model: one OS thread; every CUDA thread of a block is a ucontext coroutine that runs until it reaches a collective
// (__syncwarp / __syncthreads / __shfl* / __ballot / __reduce*), where it parks until all lanes named by the mask have
// arrived.  Blocks run one after the other.  A collective that can never complete (a lane named in the mask has exited or
// waits elsewhere) aborts with a message -- the emulator is stricter than the hardware on purpose.
//
// A kernel source is made emulable by
//   * launching through PC_LAUNCH(kernel, grid, block, smem, stream, args...),
//   * declaring dynamic shared memory with PC_DYN_SMEM(name),
//   * keeping inline PTX behind `#ifdef PC_EMU` alternatives.
#pragma once
#ifndef PC_EMU
#error "cuda_emu.h is for -DPC_EMU host builds of the kernel sources (tests only)"
#endif

#include <cuda_runtime.h>  // vector types, cudaError_t, cudaStream_t (declarations only; nothing of libcudart is called)
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <vector>

#undef __shared__
#define __shared__ static  // static shared arrays: one block runs at a time
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__ __attribute__((noinline))

namespace emu {

struct Lane;
struct Bar {
    unsigned count = 0, gen = 0;
};
struct Block {
    std::vector<Lane *> lanes;
    std::vector<std::map<uint32_t, Bar>> warp_bars;  // per warp, keyed by member mask
    Bar block_bar;
    Bar named[16];
    std::vector<unsigned char> smem;
    uint64_t slots[1024][2];  // exchange slots, one per thread
    bool progress = false;
};
struct Lane {
    ucontext_t ctx;
    std::vector<unsigned char> stack;
    uint3 tid;
    int linear = 0;
    bool done = false;
    Block *blk = nullptr;
};

struct State {
    Lane *cur = nullptr;
    ucontext_t sched;
    uint3 bid{0, 0, 0};
    dim3 bdim{1, 1, 1}, gdim{1, 1, 1};
    const std::function<void()> *body = nullptr;
};
inline State &S() {
    static State s;
    return s;
}

inline void yield() { swapcontext(&S().cur->ctx, &S().sched); }

inline void trampoline() {
    (*S().body)();
    S().cur->done = true;
    S().cur->blk->progress = true;
    swapcontext(&S().cur->ctx, &S().sched);
}

inline void arrive_wait(Bar &b, unsigned members) {
    Block *blk = S().cur->blk;
    const unsigned my = b.gen;
    if (++b.count == members) {
        b.count = 0;
        ++b.gen;
        blk->progress = true;
        return;
    }
    while (b.gen == my) yield();
}

inline void named_bar(int id, int threads) { arrive_wait(S().cur->blk->named[id & 15], (unsigned)threads); }
inline int lane_id() { return S().cur->linear & 31; }
inline int warp_id() { return S().cur->linear >> 5; }

inline void warp_bar(uint32_t mask) {
    Block *blk = S().cur->blk;
    if (!((mask >> lane_id()) & 1u)) {
        fprintf(stderr, "cuda_emu: lane %d executes a collective whose mask %08x does not name it\n", lane_id(), mask);
        abort();
    }
    arrive_wait(blk->warp_bars[warp_id()][mask], (unsigned)__builtin_popcount(mask));
}

template <class T>
inline T exchange(uint32_t mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 16, "exchange slot");
    Block *blk = S().cur->blk;
    const int base = warp_id() * 32;
    memcpy(blk->slots[base + lane_id()], &v, sizeof(T));
    warp_bar(mask);
    T r;
    memcpy(&r, blk->slots[base + (src_lane & 31)], sizeof(T));
    warp_bar(mask);
    return r;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()> &body) {
    State &st = S();
    st.gdim = grid;
    st.bdim = block;
    st.body = &body;
    const int nthreads = (int)(block.x * block.y * block.z);
    const size_t stack_bytes = 256 * 1024;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                st.bid = uint3{bx, by, bz};
                Block blk;
                blk.smem.assign(smem_bytes + 64, 0xCD);  // poisoned: reads of unwritten shared memory show up
                blk.warp_bars.resize((nthreads + 31) / 32);
                std::vector<Lane> lanes(nthreads);
                for (int t = 0; t < nthreads; ++t) {
                    Lane &l = lanes[t];
                    l.linear = t;
                    l.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
                    l.blk = &blk;
                    l.stack.resize(stack_bytes);
                    getcontext(&l.ctx);
                    l.ctx.uc_stack.ss_sp = l.stack.data();
                    l.ctx.uc_stack.ss_size = stack_bytes;
                    l.ctx.uc_link = &st.sched;
                    makecontext(&l.ctx, (void (*)())trampoline, 0);
                    blk.lanes.push_back(&l);
                }
                int alive = nthreads;
                while (alive > 0) {
                    blk.progress = false;
                    for (int t = 0; t < nthreads; ++t) {
                        Lane &l = lanes[t];
                        if (l.done) continue;
                        st.cur = &l;
                        swapcontext(&st.sched, &l.ctx);
                        if (l.done) --alive;
                    }
                    if (!blk.progress && alive > 0) {
                        fprintf(stderr, "cuda_emu: deadlock in block (%u,%u,%u): %d threads wait on a collective that cannot complete\n",
                                bx, by, bz, alive);
                        abort();
                    }
                }
            }
    st.cur = nullptr;
}

inline unsigned char *dyn_smem() {
    unsigned char *p = S().cur->blk->smem.data();
    return (unsigned char *)(((uintptr_t)p + 63) & ~(uintptr_t)63);
}

}  // namespace emu

// ---- built-in variables ---------------------------------------------------------------------------------------
#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)

#define PC_DYN_SMEM(name) unsigned char *name = emu::dyn_smem()
#define PC_LAUNCH(kernel, grid, block, smem, stream, ...) emu::launch(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })

// ---- collectives ------------------------------------------------------------------------------------------------
inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::warp_bar(mask); }
inline void __syncthreads() {
    emu::Block *b = emu::S().cur->blk;
    emu::arrive_wait(b->block_bar, (unsigned)b->lanes.size());
}
template <class T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = emu::lane_id();
    const int s = (lane & ~(width - 1)) | (src & (width - 1));
    return emu::exchange(mask, v, s);
}
template <class T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32) {
    const int lane = emu::lane_id();
    int s = lane ^ lanemask;
    if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
    return emu::exchange(mask, v, s);
}
template <class T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = emu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane & ~(width - 1))) s = lane;
    return emu::exchange(mask, v, s);
}
template <class T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = emu::lane_id();
    int s = lane + (int)delta;
    if (s > (lane | (width - 1))) s = lane;
    return emu::exchange(mask, v, s);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    unsigned r = 0;
    // every lane of the mask contributes one bit: gather with 32 exchanges' worth of slots in one round
    emu::Block *blk = emu::S().cur->blk;
    const int base = emu::warp_id() * 32;
    uint64_t v = pred ? 1 : 0;
    memcpy(blk->slots[base + emu::lane_id()], &v, 8);
    emu::warp_bar(mask);
    for (int l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) {
            uint64_t o;
            memcpy(&o, blk->slots[base + l], 8);
            r |= (unsigned)(o & 1) << l;
        }
    emu::warp_bar(mask);
    return r;
}
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
template <class T, class F>
inline T emu_reduce(unsigned mask, T v, F f) {
    emu::Block *blk = emu::S().cur->blk;
    const int base = emu::warp_id() * 32;
    uint64_t w = 0;
    memcpy(&w, &v, sizeof(T));
    memcpy(blk->slots[base + emu::lane_id()], &w, 8);
    emu::warp_bar(mask);
    bool first = true;
    T acc{};
    for (int l = 0; l < 32; ++l)
        if ((mask >> l) & 1u) {
            T o;
            memcpy(&o, blk->slots[base + l], sizeof(T));
            acc = first ? o : f(acc, o);
            first = false;
        }
    emu::warp_bar(mask);
    return acc;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a > b ? a : b; }); }
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a < b ? a : b; }); }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a + b; }); }
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a | b; }); }
inline unsigned __reduce_and_sync(unsigned mask, unsigned v) { return emu_reduce(mask, v, [](unsigned a, unsigned b) { return a & b; }); }
inline int __reduce_max_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a > b ? a : b; }); }
inline int __reduce_min_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a < b ? a : b; }); }
inline int __reduce_add_sync(unsigned mask, int v) { return emu_reduce(mask, v, [](int a, int b) { return a + b; }); }

// ---- arithmetic intrinsics (compile with -ffp-contract=off) -------------------------------------------------------
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline int __double2hiint(double d) {
    uint64_t u;
    memcpy(&u, &d, 8);
    return (int)(u >> 32);
}
inline int __double2loint(double d) {
    uint64_t u;
    memcpy(&u, &d, 8);
    return (int)(u & 0xffffffffu);
}
inline double __hiloint2double(int hi, int lo) {
    uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
inline long long __double_as_longlong(double d) {
    long long u;
    memcpy(&u, &d, 8);
    return u;
}
inline double __longlong_as_double(long long u) {
    double d;
    memcpy(&d, &u, 8);
    return d;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline unsigned __brev(unsigned x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
    return (x >> 16) | (x << 16);
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    const uint64_t v = ((uint64_t)hi << 32) | lo;
    return (unsigned)(v >> (sh & 31));
}
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) {
        const unsigned sel = (s >> (4 * i)) & 0xf;
        unsigned b = (unsigned)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) b = (b & 0x80) ? 0xff : 0;
        r |= b << (8 * i);
    }
    return r;
}
template <class T>
inline T atomicAdd(T *p, T v) {
    T o = *p;
    *p = o + v;
    return o;
}
template <class T>
inline T atomicMax(T *p, T v) {
    T o = *p;
    if (v > o) *p = v;
    return o;
}
template <class T>
inline T atomicOr(T *p, T v) {
    T o = *p;
    *p = o | v;
    return o;
}
inline void __threadfence() {}
inline void __trap() { abort(); }

// ---- the few runtime calls the host side of a kernel file makes ---------------------------------------------------
namespace emu {
inline cudaError_t Malloc(void **p, size_t n) {
    *p = malloc(n ? n : 1);
    if (*p) memset(*p, 0xCD, n);  // poisoned like fresh device memory is arbitrary
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t Free(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t Memcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t MemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
    memcpy(d, s, n);
    return cudaSuccess;
}
inline cudaError_t MemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) {
    memset(d, v, n);
    return cudaSuccess;
}
template <class K>
inline cudaError_t FuncSetAttribute(K, cudaFuncAttribute, int) {
    return cudaSuccess;
}
inline cudaError_t GetLastError() { return cudaSuccess; }
}  // namespace emu
#define cudaMalloc(p, n) emu::Malloc((void **)(p), (n))
#define cudaFree(p) emu::Free((void *)(p))
#define cudaMemcpy emu::Memcpy
#define cudaMemcpyAsync emu::MemcpyAsync
#define cudaMemsetAsync emu::MemsetAsync
#define cudaFuncSetAttribute emu::FuncSetAttribute
#define cudaGetLastError emu::GetLastError
