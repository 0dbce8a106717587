// common.cuh -- shared host/device helpers for the polarcub_b200 CUDA library (sm_100a).
#pragma once
#ifdef PC_EMU
#include "../../tests/emu/cuda_emu.h"  // CPU emulation of the kernel sources: test builds only (tests/emu)
#else
// dynamic shared memory and kernel launches, spelled so that tests/emu can run the kernel sources on the CPU
#define PC_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define PC_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <vector>

#include "../../include/polarcub_b200.h"

struct pc_plan;

namespace pc {

// ---- error plumbing (no exceptions cross the C-ABI) ------------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<unsigned long long> g_launches;

#define PC_CUDA(call)                                                                            \
    do {                                                                                         \
        cudaError_t _e = (call);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            pc::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));  \
            return PC_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define PC_LAUNCH_CHECK()                                                                        \
    do {                                                                                         \
        pc::g_launches.fetch_add(1, std::memory_order_relaxed);                                  \
        cudaError_t _e = cudaGetLastError();                                                     \
        if (_e != cudaSuccess) {                                                                 \
            pc::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return PC_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

#define PC_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) {                                          \
            pc::set_error("%s (%s)", msg, #cond);               \
            return PC_ERR_INVALID;                              \
        }                                                       \
    } while (0)

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
static inline size_t align256(size_t a) { return (a + 255) & ~(size_t)255; }

int num_sms();  // SM count of the current device (cached)
void prof_mark(cudaStream_t st);  // bench.py timing hook: call before and after the dominant kernel launch
void prof_suspend(int delta);     // +1 / -1: a composite operation that brackets itself silences the marks of its parts
struct ProfScope {                // brackets a multi-launch operation as ONE measured unit
    cudaStream_t st;
    explicit ProfScope(cudaStream_t s) : st(s) {
        prof_mark(st);
        prof_suspend(1);
    }
    ~ProfScope() {
        prof_suspend(-1);
        prof_mark(st);
    }
};

// ---- schedule of the SC tree walk -------------------------------------------------------------------
// The decoder never recurses: the host flattens the tree walk of
// BinaryPolarEncoderDecoder.recursiveEncodeDecode (BinaryPolarEncoderDecoder.py:223-325) into a list of
// nodes visited left to right.  A node is either a single information leaf or a maximal all-frozen
// (rate-0) sub-tree whose codeword is known in advance, so its probabilities are never computed.
enum : int { NODE_INFO = 0, NODE_RATE0 = 1, NODE_GENIE = 2, NODE_RATE1 = 3, NODE_BLOCK = 4 };  // GENIE: known leaf bit, leaf probabilities captured
// BLOCK (binary decoder's block schedule, pc_plan::sched_b): a sub-tree of 2^SC_LB leaves that is not all-frozen, decoded by one unrolled
// routine -- the same node updates in the same order as the leaf-by-leaf walk; `bits` = frozen mask (low 2^SC_LB bits, bit j = leaf
// i + j is frozen) | frozen values << 2^SC_LB
#ifndef SC_LEAF_BLOCK_LOG
#define SC_LEAF_BLOCK_LOG 4
#endif
constexpr int SC_LB = SC_LEAF_BLOCK_LOG;  // <= 4: the block's vector lives in the walk's shared-memory levels and its masks in 32 bits
// RATE1 (hybrid decoder's sub-block schedules only): all-information sub-tree, `bits` = number of entries it spans after this one
struct SchedEntry {
    int32_t i;      // first u index covered by the node
    int8_t l;       // log2 of the node size
    int8_t kind;    // NODE_INFO / NODE_RATE0
    int8_t top;     // highest level whose vector must be recomputed before this node: ctz(i), or n for i == 0
    int8_t pad;
    uint32_t bits;  // rate-0, l < 5: the node codeword (natural order) in the low 2^l bits; l >= 5: word offset into r0_words
};

int bitrev_words_launch(int n, int64_t B, const uint32_t *in, uint32_t *out, cudaStream_t st);  // encode.cu
void scl_tables_release(const pc_plan *p);
void stream_tables_release(const pc_plan *p);
void trellis_tables_release(const pc_plan *p);
void genie_tables_release(const pc_plan *p);
void hybrid_tables_release(const pc_plan *p);

}  // namespace pc

struct pc_plan {
    int q, n, N, k, device;
    std::vector<uint8_t> frozen_mask, frozen_vals;
    std::vector<pc::SchedEntry> sched;
    std::vector<pc::SchedEntry> sched_b;  // binary, n > SC_LB: the same walk with NODE_BLOCK entries for the leaf blocks (sc_binary.cu)
    pc::SchedEntry *d_sched_b = nullptr;
    // device copies
    pc::SchedEntry *d_sched = nullptr;
    uint32_t *d_r0_words = nullptr;   // natural-order codewords of rate-0 nodes with l >= 5
    int32_t *d_src = nullptr;         // [N] u index -> information index, or -1 when frozen
    uint32_t *d_frozen_words = nullptr;  // [ceil(N/32)] frozen values as bits (binary)
    uint8_t *d_frozen_mask = nullptr;    // [N]
    uint8_t *d_frozen_vals = nullptr;    // [N]
    // warp-per-frame encoder (encode.cu): per u word w, 8 words {information-position mask, number of information bits
    // before the word, the five bit-deposit (expand) masks of that mask, frozen-value bits}
    uint32_t *d_enc_tab = nullptr;       // [ceil(N/32)][8]
    // block encoder (2^16 <= N <= 2^20): the same table as [2][32][N/1024] uint4 -- half h of the entry of word t * 32 + j
    // at ((h * 32 + j) * N/1024 + t), so that the threads t of a warp read consecutive 16-byte chunks
    uint32_t *d_enc_tab_t = nullptr;
};

// ---- device helpers ---------------------------------------------------------------------------------
namespace pc {

__device__ __forceinline__ uint32_t bitrev_n(uint32_t i, int n) { return n == 0 ? 0u : (__brev(i) >> (32 - n)); }

// Gathers output word j of the bit-reversal permutation out[i] = nat[rev_n(i)] from a word array.
// `ld(w)` returns natural-order word w.
template <class Ld>
__device__ __forceinline__ uint32_t bitrev_gather_word(Ld ld, int n, uint32_t j) {
    uint32_t out = 0;
    const uint32_t N = 1u << n;
#pragma unroll 4
    for (uint32_t b = 0; b < 32; ++b) {
        uint32_t i = 32u * j + b;
        if (i < N) {
            uint32_t r = bitrev_n(i, n);
            out |= ((ld(r >> 5) >> (r & 31u)) & 1u) << b;
        }
    }
    return out;
}

}  // namespace pc
