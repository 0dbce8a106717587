// scl_tables.cuh -- host-flattened op list of the SC-list decoders (shared by scl.cu and scl_bin.cu).
//
// The recursion of QaryPolarEncoderDecoder.recursiveListDecode (QaryPolarEncoderDecoder.py:403-757) depends on the
// frozen set only, so it is flattened once per plan into a list of ops that every frame of a batch executes.
#pragma once
#include "common.cuh"

namespace pc {

enum : int { OP_MINUS = 0, OP_PLUS = 1, OP_COMBINE = 2, OP_RATE0 = 3, OP_REP = 4, OP_RATE1 = 5, OP_SPC = 6 };
constexpr int SCL_LMAX = 32;
constexpr int SCL_THREADS = 64;
// scl_path.cu op flags
enum : uint32_t {
    SCLP_SSRC = 1u << 8,   // the source vector is in the shared ("single path") layout: written before the first fork
    SCLP_SDST = 1u << 9,   // the op runs before the first fork: one path, the lanes of a frame share its elements
    SCLP_FUSED = 1u << 10, // MINUS / PLUS (l) with the following MINUS (l-1) folded in
    SCLP_CHAN = 1u << 11,  // the source is the channel level (l == n)
    SCLP_DUAL = 1u << 12,  // PLUS after the first fork whose source is one shared vector (or the channel): its output has two
                           // variants per element only, g(a, b, 0) and g(a, b, 1); lists of 4+ store those instead of L vectors
    SCLP_DSRC = 1u << 13   // the source level was written by a SCLP_DUAL op
};

struct alignas(16) SclOp {
    int8_t kind, l, c, pad;
    int32_t i;         // first u index of the node
    int32_t info_idx;  // informationVectorIndex when the node starts
    int32_t fv_idx;    // position of the frozen-values iterator when the node starts
    int32_t kpos;      // Rep: offset of the single information index inside the segment
    int32_t coef_off;  // Rep: offset into rep_coef (natural-order T(e_kpos) mod q)
    int32_t pad2;      // 32-byte record: two 16-byte loads
    int32_t coefw_off; // Rep, q = 2: offset into rep_coef_words (REFERENCE-order T(e_kpos), bit-packed, max(1, size/32) words)
};

struct SclTables {
    std::vector<SclOp> ops;
    std::vector<int32_t> a_src, f_src, info_src;
    std::vector<int8_t> node_level;
    std::vector<uint8_t> rep_coef;
    // q = 2 only (scl_bin.cu): bit-packed tables
    std::vector<uint32_t> rep_coef_words;  // per Rep op, reference order
    std::vector<uint32_t> stage_mask;      // [n][max(1,N/32)] natural-order masked-butterfly masks: bit pos set iff
                                           // (pos & 2^t) == 0 and the fast node holding pos is larger than 2^t
    std::vector<uint2> ops2;               // packed ops: x = kind | l << 3 | c << 7 | i << 8, y = fv_idx | coefw_off << 16
    std::vector<uint2> ops3;               // scl_warp.cu: ops2 with PLUS/MINUS (l >= 7) + MINUS (l-1) pairs fused (x bit 30)
    std::vector<int32_t> perm;             // [N] reference position -> natural position inside its fast node
    // scl_path.cu: x = kind | l << 3 | c << 7 | flags (SCLP_*), y = first u index, z = fv_idx, w = coefw_off (Rep)
    std::vector<uint4> opsP;
    int n_leaf = 0;                        // number of fast nodes
    SclOp *d_ops = nullptr;
    int32_t *d_a_src = nullptr, *d_f_src = nullptr, *d_info_src = nullptr;
    int8_t *d_node_level = nullptr;
    uint8_t *d_rep_coef = nullptr;
    uint32_t *d_rep_coef_words = nullptr, *d_stage_mask = nullptr;
    int32_t *d_perm = nullptr;
    uint2 *d_ops2 = nullptr, *d_ops3 = nullptr;
    uint4 *d_opsP = nullptr;
};


// builds (once per plan) and returns the device tables; nullptr + pc_last_error() on failure
SclTables *scl_tables(const pc_plan *p);

}  // namespace pc
