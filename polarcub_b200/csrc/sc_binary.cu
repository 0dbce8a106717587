// sc_binary.cu -- batched binary successive-cancellation decoding, float64, bit-identical to the reference.
//
// Replaces BinaryPolarEncoderDecoder.decode / recursiveEncodeDecode (BinaryPolarEncoderDecoder.py:71-99, :223-325)
// and the arithmetic of BinaryMemorylessVectorDistribution (minusTransform :15-29, plusTransform :31-47,
// calcNormalizationVector / normalize :71-87, calcMarginalizedProbabilities :52-69) for a uniform prior.
//
// Design (B200-first, not a port of the recursion):
//  * one FRAME PER LANE: all frames share the frozen set, so the tree walk is identical for every lane
//    of a warp (no divergence, no lane idles in the small stages that dominate SC) and every node vector
//    is laid out [element][lane] so loads and stores are fully coalesced 256-byte rows;
//  * the walk is a host-flattened schedule (plan.cu) with maximal rate-0 sub-trees pruned;
//  * node state is ONE float64 per element: after the reference's max-normalisation a probability pair is
//    (1, r) or (r, 1) with r = min/max, or (0, 0) after contradicting hard knowledge.  We store r with the
//    sign bit saying which side is 1, and NaN for (0, 0).  Every product / sum / quotient the reference
//    performs on the pair is reproduced on r with the same IEEE-754 roundings (multiplications by exactly
//    1.0 are dropped -- they are exact), so decisions and probabilities are bit-identical;
//  * levels 0..LS of the tree live in shared memory (one private column per thread), the larger levels in
//    an L2-resident global scratch private to the warp; partial sums are bit-packed, the low five levels in
//    a register;
//  * the channel level is read through a transposed, bit-reversed copy made by the ingest kernel, and the
//    decoded words are transposed back (codeword bit-reversed to the reference order) by the egress kernel.
#include <map>
#include <mutex>

#include "async_copy.cuh"
#include "sc_arith.cuh"

namespace pc {

// sc_stream.cu: one frame per CTA, upper stages streamed through HBM (large blocks)
bool sc_stream_supported(const pc_plan *plan);
size_t sc_stream_workspace_bytes(const pc_plan *plan, int64_t B);
int64_t sc_stream_wave_frames(const pc_plan *plan);
int sc_stream_decode(const pc_plan *plan, int kind, const void *d_in, int64_t B, const double *h_table, int Y, uint32_t *d_cw,
                     uint32_t *d_info, void *ws, size_t ws_bytes, cudaStream_t st);

#ifndef SC_SYNC
#define SC_SYNC 2                // barriers among the warps of a block (see the kernel's group loop)
#endif
constexpr int LS = 4;            // levels 0..LS in shared memory: 2^(LS+1)-1 doubles per thread
constexpr int SC_THREADS = 256;  // 8 warps = 256 frames per block
constexpr int SMEM_VALS = (1 << (LS + 1)) - 1;
constexpr int SC_BLOCKS_PER_SM = 3;
constexpr int SC_MAX_N = 16;     // frame-per-lane kernel; larger blocks use the hybrid / streamed decoders
constexpr int SC_INPUT_PACKED = 2;  // internal input kind: the top level is a packed vector [2^n][Bpad] (hybrid decoder)

struct ScParams {
    int n, k, n_sched, Y;
    int64_t frames;  // frames in this chunk
    int64_t Bpad;    // row pitch (frames, multiple of 32) of the transposed buffers
    const SchedEntry *sched;
    const uint32_t *r0_words;
    const void *in_t;   // [N][Bpad] uint8 symbols or double2 probability pairs, natural (bit-reversed) order
    double *vals;       // [warps][N - 2^(LS+1)][32] scratch for levels > LS
    uint32_t *cw_t;     // [Nw][Bpad] natural-order codeword words (also the partial-sum store)
    uint32_t *info_t;   // [Kw][Bpad]
    const uint32_t *u_t;  // genie / prior-encode: [Nw][Bpad] the known u bits (natural u order)
    double *marg_t;       // optional capture: [N][Bpad] packed level-0 value of every leaf
    const uint8_t *fmask; // dual / prior modes: [N] 1 = frozen
    const double *rnd;    // dual / prior modes: randomlyGeneratedNumbers, row (frame) pitch rnd_stride doubles (0: one shared vector)
    int64_t rnd_stride;
    double table[32];   // symbols: [Y][2] joint probabilities
};

// One f / g node update on packed values with a single division site: the g node's same-side case (1 * 1, ra * rb) is
// already normalised and divides by exactly 1.0 (exact), so both node kinds end in one `mn / mx`.  Same products, sums,
// comparisons and quotient as f_packed / g_packed (sc_arith.cuh): identical bits.
__device__ __forceinline__ double node_packed(double a, double b, bool isg, uint32_t u) {
    const double ra = d_abs(a), rb = d_abs(b);
    const double prod = __dmul_rn(ra, rb);
    double mx, mn;
    bool side;
    if (!isg) {
        const uint32_t s = d_sign(a) ^ d_sign(b);
        const double A = __dadd_rn(1.0, prod), Bv = __dadd_rn(ra, rb);
        const bool c = Bv > A, d = A > Bv;
        mx = c ? Bv : A;
        mn = c ? A : Bv;
        side = s ? d : c;
    } else {
        const uint32_t sa = d_sign(a) ^ u, sb = d_sign(b);
        const bool same = sa == sb, c = ra > rb, e = rb > ra, lt1 = prod < 1.0;
        mx = same ? 1.0 : (c ? ra : rb);
        mn = same ? prod : (c ? rb : ra);
        const uint32_t s1 = sb & (lt1 ? 1u : 0u), s2 = sb ? (c ? 1u : 0u) : (e ? 1u : 0u);
        side = (same ? s1 : s2) != 0u;
    }
    return d_pack(mn / mx, side ? 1u : 0u);
}

// ---- leaf blocks (NODE_BLOCK, common.cuh): the 2^SC_LB leaves of a sub-tree that is not all-frozen, decoded by one routine from the
// level-SC_LB vector.  The leaf-by-leaf walk pays ~200 instructions of schedule handling per entry (entry decode, level-loop set-up for levels
// of 1 / 2 / 4 elements, partial sums, information word bookkeeping) on top of its node updates; here the SAME node updates run in
// the same order on the same operands (frozen leaves take their values, all-frozen halves are skipped exactly as the walk skips its
// rate-0 nodes), so decisions and codewords are identical, and the bookkeeping is paid once per block.  Loops over the two halves of a
// (sub-)block are not unrolled: one copy of the node routine per tree level keeps the code small.
// natural-order polar transform of the low 2^L bits: T([a; b]) = [T(a) ^ T(b); T(b)]
template <int L>
__device__ __forceinline__ uint32_t xf_bits(uint32_t w) {
    if (L >= 1) w ^= (w >> 1) & 0x55555555u;
    if (L >= 2) w ^= (w >> 2) & 0x33333333u;
    if (L >= 3) w ^= (w >> 4) & 0x0f0f0f0fu;
    if (L >= 4) w ^= (w >> 8) & 0x00ff00ffu;
    return w & ((1u << (1 << L)) - 1u);
}
// The sub-tree of 2^L leaves whose level-L vector is in the thread's shared-memory slots (element h of level l at sv[((1 << l) + h) * STR],
// the walk's own layout; the levels below L are written there as the walk does).  fm / fv: frozen mask / values of the leaves;
// u: decisions, x: their codeword.  The two halves run through ONE copy of the code per level (the half loop is not unrolled) with a
// warp-uniform branch around the specialised f / g updates.
template <int L, int STR>
struct LeafBlock {
    static __device__ __forceinline__ void run(double *sv, uint32_t fm, uint32_t fv, uint32_t &u, uint32_t &x) {
        constexpr int H = 1 << (L - 1);
        constexpr uint32_t HM = (1u << H) - 1u;
        const double *vL = sv + (STR << L);
        double *vl = sv + (STR << (L - 1));
        uint32_t uu = 0, xl = 0, xr = 0;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const uint32_t mH = (fm >> (H * half)) & HM, fH = (fv >> (H * half)) & HM;
            uint32_t uH = fH, xH = xf_bits<L - 1>(fH);
            if (mH != HM) {  // not all-frozen: f (first half) or g with the first half's codeword, then the half's own sub-tree
                if (half == 0) {
#pragma unroll 4
                    for (int h = 0; h < H; ++h) vl[h * STR] = node_packed(vL[h * STR], vL[(h + H) * STR], false, 0u);
                } else {
#pragma unroll 4
                    for (int h = 0; h < H; ++h) vl[h * STR] = node_packed(vL[h * STR], vL[(h + H) * STR], true, (xl >> h) & 1u);
                }
                LeafBlock<L - 1, STR>::run(sv, mH, fH, uH, xH);
            }
            uu |= uH << (H * half);
            if (half == 0)
                xl = xH;
            else
                xr = xH;
        }
        u = uu;
        x = (xl ^ xr) | (xr << H);
    }
};
template <int STR>
struct LeafBlock<1, STR> {  // two leaves from their level-1 vector
    static __device__ __forceinline__ void run(double *sv, uint32_t fm, uint32_t fv, uint32_t &u, uint32_t &x) {
        const double v0 = sv[2 * STR], v1 = sv[3 * STR];
        uint32_t u0 = fv & 1u, u1 = (fv >> 1) & 1u;
        if (!(fm & 1u)) u0 = d_sign(node_packed(v0, v1, false, 0u));  // level 0: p0 >= p1 -> 0 (ties and (0,0) -> 0)
        if (!(fm & 2u)) u1 = d_sign(node_packed(v0, v1, true, u0));
        u = u0 | (u1 << 1);
        x = (u0 ^ u1) | (u1 << 1);
    }
};

// Elements [0, size) of a level from the level above (size >= 4), batches of 4: every load of a batch is issued before the
// first division.  SSTR / DSTR: element strides (doubles) of the source and destination levels -- SC_THREADS for the
// shared-memory levels, 32 for the warp's global scratch -- compile-time so that the batch's addresses are immediates.
// g: the decision bits come from `uw` (one word per 32 elements, pitch Bpad) or, for levels below 32, from `ureg`.
// Erasure-type frames (every value of the sub-block's input vector is hard knowledge r = 0, an erasure r = 1 or the (0,0)
// contradiction state -- all a BEC frame can hold): the reference's products, sums and quotients on such pairs are exact
// and stay in the same four states, so a node update is a selection on the high words -- no division, no branch:
//   f: (0,0) if either is; erasure if either is; else hard with side sa ^ sb
//   g: (0,0) if either is; both hard: side sb if sa ^ u == sb, else the (0,0) contradiction; one erased: the other one's
//      (bit-adjusted) knowledge; both erased: erasure
// Same decisions as node_packed on these operands (f_packed01 / g_packed01 in sc_arith.cuh spell out the cases).
// One-byte state codes of the hybrid decoder's upper stages: bit 0 side, bit 1 erasure, bit 2 contradiction.
__device__ __forceinline__ uint32_t state_code(double v) {
    if (v != v) return 4u;
    return d_abs(v) == 1.0 ? 2u : d_sign(v);
}
// f / g on four frames at once (one code per byte), the same case table as node01
__device__ __forceinline__ uint32_t f8(uint32_t a, uint32_t b) {
    const uint32_t o = a | b;
    const uint32_t c = o & 0x04040404u;
    const uint32_t e = (o & 0x02020202u) & ~(c >> 1);
    const uint32_t s = ((a ^ b) & 0x01010101u) & ~(e >> 1) & ~(c >> 2);
    return c | e | s;
}
__device__ __forceinline__ uint32_t g8(uint32_t a, uint32_t b, uint32_t u) {  // u: the decision bit in bit 0 of each byte
    const uint32_t m = 0x01010101u;
    const uint32_t ea = (a >> 1) & m, eb = (b >> 1) & m;
    const uint32_t sa = (a ^ u) & m, sb = b & m;
    const uint32_t c = ((((a | b) >> 2) & m) | (~(ea | eb) & (sa ^ sb))) & m;
    const uint32_t e = ea & eb & ~c;
    const uint32_t s = ((eb & sa) | (~eb & sb)) & m & ~c & ~e;
    return (c << 2) | (e << 1) | s;
}

__device__ __forceinline__ double node01(double a, double b, bool isg, uint32_t u) {
    const uint32_t ha = (uint32_t)__double2hiint(a), hb = (uint32_t)__double2hiint(b);
    const bool ca = (ha & 0x7ff00000u) == 0x7ff00000u, cb = (hb & 0x7ff00000u) == 0x7ff00000u;
    const bool ea = ha == 0x3ff00000u, eb = hb == 0x3ff00000u;
    uint32_t r;
    if (!isg) {
        r = (ea || eb) ? 0x3ff00000u : ((ha ^ hb) & 0x80000000u);
    } else {
        const uint32_t sa = (ha ^ (u << 31)) & 0x80000000u, sb = hb & 0x80000000u;
        const uint32_t hard = sa != sb ? 0x7ff80000u : sb;
        r = ea ? (eb ? 0x3ff00000u : sb) : (eb ? sa : hard);
    }
    if (ca || cb) r = 0x7ff80000u;
    return __hiloint2double((int)r, 0);
}
__device__ __forceinline__ bool d_is01c(double x) {  // +-0, 1.0 or NaN
    const uint32_t h = (uint32_t)__double2hiint(x) & 0x7fffffffu;
    return (h == 0u && __double2loint(x) == 0) || (h == 0x3ff00000u && __double2loint(x) == 0) || x != x;
}

// F01: `w01` (warp-uniform) selects node01 for the whole batch
template <int SSTR, int DSTR, bool F01 = false, int BW = 4>
__device__ __forceinline__ void level_batches(const double *sp, double *dp, int size, bool isg, const uint32_t *uw, int64_t Bpad,
                                              uint32_t ureg, bool w01 = false) {
    const double *sp2 = sp + (int64_t)size * SSTR;
    const int per_word = size < 32 ? size : 32;
#pragma unroll 1
    for (int w0 = 0; w0 < size; w0 += 32) {
        uint32_t ubw = ureg;
        if (isg && size >= 32) {
            ubw = *uw;
            uw += Bpad;
        }
#pragma unroll 1
        for (int hh = 0; hh < per_word; hh += BW) {
            double a[BW], b[BW];
#pragma unroll
            for (int u = 0; u < BW; ++u) {
                a[u] = sp[u * SSTR];
                b[u] = sp2[u * SSTR];
            }
            if (F01 && w01) {
#pragma unroll
                for (int u = 0; u < BW; ++u) dp[u * DSTR] = node01(a[u], b[u], isg, (ubw >> (hh + u)) & 1u);
            } else {
#pragma unroll
                for (int u = 0; u < BW; ++u) dp[u * DSTR] = node_packed(a[u], b[u], isg, (ubw >> (hh + u)) & 1u);
            }
            sp += BW * SSTR;
            sp2 += BW * SSTR;
            dp += BW * DSTR;
        }
    }
}

// Levels lev and lev-1 from level lev+1 in ONE sweep over the global scratch: the two level-lev elements an element of level lev-1
// needs are produced (and stored: the g pass of level lev reads them a sub-tree later) and combined while still in registers, so
// level lev is not re-read for the f pass below it.  Two elements of level lev-1 per iteration: eight loads in flight, four + two node
// updates.  sp: level lev+1 (global, pitch 32), d0: level lev (global), d1: level lev-1 (pitch DSTR1: global or shared memory).
#ifndef SC_FUSE2
#define SC_FUSE2 1
#endif
template <int DSTR1>
__device__ __forceinline__ void fused2_levels(const double *sp, double *d0, double *d1, int lev, bool isg, const uint32_t *uw, int64_t Bpad) {
    const int S = 1 << (lev - 1);  // elements of level lev-1 (>= 16); level lev has 2 S, the source 4 S
    uint32_t ub0 = 0, ub1 = 0;     // decision words of the level-lev elements h .. and h + S ..
    if (isg && S < 32) {           // S == 16: level lev is one word
        ub0 = uw[0];
        ub1 = ub0 >> S;
    }
    const double *s0 = sp, *s1 = sp + (int64_t)S * 32, *s2 = s1 + (int64_t)S * 32, *s3 = s2 + (int64_t)S * 32;
    double *e0 = d0, *e1 = d0 + (int64_t)S * 32;
#pragma unroll 1
    for (int h = 0; h < S; h += 2) {
        if (isg && S >= 32 && (h & 31) == 0) {
            ub0 = uw[(int64_t)(h >> 5) * Bpad];
            ub1 = uw[(int64_t)((h + S) >> 5) * Bpad];
        }
        double a[2], b[2], c[2], d[2], pq[2], qq[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            a[u] = s0[u * 32];
            b[u] = s2[u * 32];
            c[u] = s1[u * 32];
            d[u] = s3[u * 32];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            pq[u] = node_packed(a[u], b[u], isg, (ub0 >> ((h + u) & 31)) & 1u);
            qq[u] = node_packed(c[u], d[u], isg, (ub1 >> ((h + u) & 31)) & 1u);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            e0[u * 32] = pq[u];
            e1[u * 32] = qq[u];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) d1[u * DSTR1] = node_packed(pq[u], qq[u], false, 0u);
        s0 += 64, s1 += 64, s2 += 64, s3 += 64, e0 += 64, e1 += 64, d1 += 2 * DSTR1;
    }
}

// MODE 0: the decoder.  The other modes walk the UNPRUNED schedule (one NODE_GENIE entry per leaf):
//   MODE_GENIE  every leaf bit is known (u_t); leaf probabilities are captured (marg_t)
//   MODE_DUAL   non-uniform a-priori distribution: lanes 2j / 2j+1 hold the a-posteriori (xy) and a-priori (x) trees of ONE
//               frame in lock step (BinaryPolarEncoderDecoder.py:277-317); an information leaf takes the xy lane's decision
//               (:248-252), a frozen leaf the x lane's rule `0 iff P(u_i = 0 | past) >= r_i` (:258-262); the pair exchanges
//               the bit with one shuffle, so both lanes carry the same partial sums
//   MODE_PRIOR  encoding under a non-uniform prior: the lane holds the x tree; information leaves take the given bit (u_t),
//               frozen leaves the rule
enum : int { MODE_DECODE = 0, MODE_GENIE = 1, MODE_DUAL = 2, MODE_PRIOR = 3 };

// P(u = 0 | past) of a packed level-0 value, as calcMarginalizedProbabilities computes it (NaN = (0,0) -> 0.5)
__device__ __forceinline__ double leaf_m0(double v) {
    if (v != v) return 0.5;
    const double rr = d_abs(v);
    const double p0 = d_sign(v) ? rr : 1.0, p1 = d_sign(v) ? 1.0 : rr;
    return p0 / __dadd_rn(__dadd_rn(0.0, p0), p1);
}

// R1 (hybrid decoder, channels with hard output symbols): the schedule carries a NODE_RATE1 entry in front of every all-information
// sub-tree.  When the node's vector is hard knowledge in all 32 frames of the warp (r = 0: every value of a BEC frame that is
// not an erasure), SC decoding of the sub-tree returns the hard decisions themselves -- the reference's products and sums
// on (1,0) / (0,1) pairs are exact and no leaf is a tie -- so the node's codeword is read off the sign bits and the
// sub-tree's entries are skipped; otherwise the walk continues into the sub-tree.  The information words are NOT produced
// by this variant (the hybrid decoder takes them from the transform of the codeword).
template <int KIND, int MODE = MODE_DECODE, bool R1 = false>
__global__ void __launch_bounds__(SC_THREADS, R1 ? 1 : SC_BLOCKS_PER_SM) sc_decode_kernel(const ScParams p) {
    extern __shared__ double sm_vals[];  // [SMEM_VALS][SC_THREADS]
    __shared__ double s_table[32];
    __shared__ double s_lut[KIND == PC_INPUT_SYMBOLS ? 768 : 1];
    if (threadIdx.x < 32) s_table[threadIdx.x] = p.table[threadIdx.x];
    __syncthreads();
    const int n = p.n, N = 1 << n;
    // Discrete channel outputs: an element of level n-1 is a function of two channel symbols and at most one decision bit,
    // so it is looked up (table built here with the same f_raw / g_raw arithmetic: identical bits) instead of being computed,
    // stored and re-read -- the largest level never touches the global scratch.  [mode][y_a][y_b], mode 0: f, 1 + u: g.
    const bool lut = KIND == PC_INPUT_SYMBOLS && n >= 7;
    const int Y = p.Y;
    if (lut) {
        for (int idx = threadIdx.x; idx < 3 * Y * Y; idx += SC_THREADS) {
            const int m = idx / (Y * Y), ya = (idx / Y) % Y, yb = idx % Y;
            const double a0 = s_table[2 * ya], a1 = s_table[2 * ya + 1], b0 = s_table[2 * yb], b1 = s_table[2 * yb + 1];
            s_lut[idx] = m == 0 ? f_raw(a0, a1, b0, b1) : g_raw(a0, a1, b0, b1, (uint32_t)(m - 1));
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (SC_THREADS / 32) + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * (SC_THREADS / 32);
    const int64_t groups = (p.frames + 31) / 32;
    const int64_t gvals = N > (1 << (LS + 1)) ? (int64_t)N - (1 << (LS + 1)) : 0;
    // element h of level lev: shared memory (lev <= LS) at sv[((1 << lev) - 1 + h) * SC_THREADS], else the warp's global
    // scratch at gv[((1 << lev) - 2^(LS+1) + h) * 32]; both reached through generic pointers (one code path)
    double *gv = p.vals + (int64_t)warp_global * gvals * 32 + lane - (int64_t)(1 << (LS + 1)) * 32;
    double *sv = sm_vals + threadIdx.x - SC_THREADS;
    auto lvl_ptr = [&](int lev) -> double * { return lev <= LS ? sv + (int64_t)(SC_THREADS << lev) : gv + ((int64_t)32 << lev); };
    auto lvl_stride = [&](int lev) -> int { return lev <= LS ? SC_THREADS : 32; };

    // SYNC: the warps of a block walk the same schedule on different frames; meeting at a barrier keeps them fetching the same
    // instructions at the same time (the kernel is ~100 KB of code: warps that drift apart over a long grid-stride loop thrash the
    // instruction cache).  1: at the start of every frame group, 2: and before every schedule entry.  A warp that has no group left
    // still attends the block's barriers.
    constexpr int SYNC = (R1 || MODE != MODE_DECODE) ? 0 : SC_SYNC;
    const int64_t iters = (groups + warps_total - 1) / warps_total;
#pragma unroll 1
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t grp = warp_global + it * warps_total;
        if (SYNC) __syncthreads();
        if (grp >= groups) {
            if (SYNC >= 2 && n > 0)
                for (int ei = 0; ei < p.n_sched; ++ei) __syncthreads();
            continue;
        }
        const int64_t col = grp * 32 + lane;  // always < Bpad; columns >= frames hold padding
        auto root = [&](int h, double &v0, double &v1) {
            if (KIND == PC_INPUT_SYMBOLS) {
                const uint32_t y = ((const uint8_t *)p.in_t)[(int64_t)h * p.Bpad + col];
                v0 = s_table[2 * y];
                v1 = s_table[2 * y + 1];
            } else if (KIND == SC_INPUT_PACKED) {
                v0 = v1 = 1.0;  // n == 0 only: not used by the hybrid decoder
            } else {
                const double2 t = ((const double2 *)p.in_t)[(int64_t)h * p.Bpad + col];
                v0 = t.x;
                v1 = t.y;
            }
        };
        uint32_t *xw = p.cw_t + col;  // word w at xw[w * Bpad]
        uint32_t *iw = p.info_t + col;
        uint32_t cwreg = 0, infoacc = 0, gword = 0;
        int icount = 0;
        int top_mode = 0;  // lut: 0 while level n-1 is f of the channel pairs (first half), 1 when it is g with x[0, N/2)

        int resume = -1;  // R1: a rate-1 node was not hard; the next entry (its first child) continues with f at this level
        bool w01 = false;  // R1: every value of this group's input vector is +-0, 1 or NaN (set by the first f pass): node01

        if (n == 0) {  // no transform: leaf rule on the raw pair (BinaryPolarEncoderDecoder.py:250-252)
            const SchedEntry e = p.sched[0];
            uint32_t bit = e.bits & 1u;
            if (e.kind == NODE_INFO) {
                double p0, p1;
                root(0, p0, p1);
                const double s = __dadd_rn(__dadd_rn(0.0, p0), p1);
                bit = 0;
                if (s > 0.0) bit = (p0 / s >= p1 / s) ? 0u : 1u;
                iw[0] = bit;
            }
            xw[0] = bit;
            continue;
        }

#pragma unroll 1
        for (int ei = 0; ei < p.n_sched; ++ei) {
            if (SYNC >= 2) __syncthreads();
            const SchedEntry e = p.sched[ei];
            const int i = e.i, l = e.l, top = e.top;
            const int stop = e.kind == NODE_RATE0 ? l + 1 : l;
            // levels to (re)compute for this node: g at level `top` with the sibling's partial sums x[i - 2^top, i), then f
            // down to the node; i == 0 starts with f at level n-1; a rate-0 node that is a whole plus child needs nothing
            int lev = i == 0 ? n - 1 : (top >= stop ? top : -1);
            bool isg = i != 0;
            if (R1 && resume >= 0) {
                lev = resume;
                isg = false;
                resume = -1;
            }
            int step = 1;  // levels produced by this iteration
#pragma unroll 1
            for (; lev >= stop; lev -= step, isg = false) {
                step = 1;
                const int size = 1 << lev;
                if (lut && lev == n - 1) {  // looked up on demand by the level below
                    top_mode = isg ? 1 : 0;
                    continue;
                }
                const uint32_t *uw = xw + (int64_t)((i - size) >> 5) * p.Bpad;  // g: decision words of whole-word levels
                const uint32_t ureg = cwreg >> ((i - size) & 31);               // g: decision bits of levels below 32
                double *dp = lvl_ptr(lev);
                const int dstr = lvl_stride(lev);
                if (SC_FUSE2 && !R1 && MODE == MODE_DECODE && lev > stop && lev > LS && lev + 1 < n && !(lut && lev == n - 2)) {
                    // source (level lev+1) and level lev in the global scratch, level lev-1 there or (lev - 1 == LS) in shared memory
                    if (lev - 1 > LS)
                        fused2_levels<32>(lvl_ptr(lev + 1), dp, lvl_ptr(lev - 1), lev, isg, uw, p.Bpad);
                    else
                        fused2_levels<SC_THREADS>(lvl_ptr(lev + 1), dp, lvl_ptr(lev - 1), lev, isg, uw, p.Bpad);
                    step = 2;
                    continue;
                }
                if (KIND == SC_INPUT_PACKED && lev + 1 == n) {
                    // the top level is a packed vector [2^n][Bpad] produced by the element-parallel upper stages (sc hybrid)
                    const double *sp = (const double *)p.in_t + col, *sp2 = sp + (int64_t)size * p.Bpad;
                    constexpr int BW = R1 ? 16 : 4;  // R1: few resident warps, more loads in flight per warp
#pragma unroll 1
                    for (int h0 = 0; h0 < size; h0 += BW) {
                        uint32_t ub = 0;
                        if (isg) ub = (size >= 32 ? uw[(int64_t)(h0 >> 5) * p.Bpad] : ureg) >> (h0 & 31);
                        double a[BW], b[BW];
#pragma unroll
                        for (int u = 0; u < BW; ++u) {
                            a[u] = (h0 + u < size) ? sp[(int64_t)u * p.Bpad] : 1.0;
                            b[u] = (h0 + u < size) ? sp2[(int64_t)u * p.Bpad] : 1.0;
                        }
                        bool fast = false;
                        if (R1) {
                            // erasure-type test of the sub-block's input vector: its first f pass sees every element once
                            bool ok = true;
#pragma unroll
                            for (int u = 0; u < BW; ++u) ok = ok && d_is01c(a[u]) && d_is01c(b[u]);
                            fast = __all_sync(0xffffffffu, ok);
                            if (!isg && i == 0) w01 = (h0 == 0 || w01) && fast;
                        }
                        if (fast) {
#pragma unroll
                            for (int u = 0; u < BW; ++u)
                                if (h0 + u < size) dp[(int64_t)(h0 + u) * dstr] = node01(a[u], b[u], isg, (ub >> u) & 1u);
                        } else {
#pragma unroll
                            for (int u = 0; u < BW; ++u)
                                if (h0 + u < size) dp[(int64_t)(h0 + u) * dstr] = node_packed(a[u], b[u], isg, (ub >> u) & 1u);
                        }
                        sp += BW * p.Bpad;
                        sp2 += BW * p.Bpad;
                    }
                    continue;
                }
                if (lev + 1 == n) {
                    // from the channel pairs (continuous inputs, or block lengths below the lookup-table threshold)
#pragma unroll 1
                    for (int h = 0; h < size; ++h) {
                        double a0, a1, b0, b1;
                        root(h, a0, a1);
                        root(h + size, b0, b1);
                        uint32_t u = 0;
                        if (isg) u = ((size >= 32 ? uw[(int64_t)(h >> 5) * p.Bpad] : ureg) >> (h & 31)) & 1u;
                        dp[(int64_t)h * dstr] = isg ? g_raw(a0, a1, b0, b1, u) : f_raw(a0, a1, b0, b1);
                    }
                    continue;
                }
                if (size < 4) {  // levels 0, 1: shared memory on both sides
                    const double *sp = sv + (SC_THREADS << (lev + 1));
#pragma unroll 1
                    for (int h = 0; h < size; ++h)
                        dp[h * SC_THREADS] = (R1 && w01) ? node01(sp[h * SC_THREADS], sp[(h + size) * SC_THREADS], isg, (ureg >> h) & 1u)
                                                         : node_packed(sp[h * SC_THREADS], sp[(h + size) * SC_THREADS], isg, (ureg >> h) & 1u);
                    continue;
                }
                if (lut && lev == n - 2) {
                    // straight from the channel symbols through the level n-1 table (destination: global scratch)
                    const uint8_t *y0p = (const uint8_t *)p.in_t + col, *y1p = y0p + (int64_t)(N >> 1) * p.Bpad;
                    const uint8_t *y2p = y0p + (int64_t)(N >> 2) * p.Bpad, *y3p = y2p + (int64_t)(N >> 1) * p.Bpad;
                    const uint32_t *xa = xw, *xb = xw + (int64_t)(N >> 7) * p.Bpad;  // x words of elements h and h + N/4
                    const uint32_t *uwp = uw;
#pragma unroll 1
                    for (int w0 = 0; w0 < size; w0 += 32) {
                        uint32_t ubw = 0, wa = 0, wb = 0;
                        if (isg) ubw = *uwp, uwp += p.Bpad;
                        if (top_mode) wa = *xa, wb = *xb, xa += p.Bpad, xb += p.Bpad;
#pragma unroll 1
                        for (int hh = 0; hh < 32; hh += 4) {
                            double a[4], b[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const uint32_t y0 = y0p[u * p.Bpad], y1 = y1p[u * p.Bpad], y2 = y2p[u * p.Bpad], y3 = y3p[u * p.Bpad];
                                const uint32_t ma = top_mode ? 1u + ((wa >> (hh + u)) & 1u) : 0u;
                                const uint32_t mb = top_mode ? 1u + ((wb >> (hh + u)) & 1u) : 0u;
                                a[u] = s_lut[(ma * Y + y0) * Y + y1];
                                b[u] = s_lut[(mb * Y + y2) * Y + y3];
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) dp[u * 32] = node_packed(a[u], b[u], isg, (ubw >> (hh + u)) & 1u);
                            y0p += 4 * p.Bpad, y1p += 4 * p.Bpad, y2p += 4 * p.Bpad, y3p += 4 * p.Bpad;
                            dp += 4 * 32;
                        }
                    }
                    continue;
                }
                if (lev > LS)
                    level_batches<32, 32, R1, R1 ? 16 : 4>(gv + ((int64_t)32 << (lev + 1)), dp, size, isg, uw, p.Bpad, ureg, w01);
                else if (lev == LS)
                    level_batches<32, SC_THREADS, R1, R1 ? 16 : 4>(gv + ((int64_t)32 << (lev + 1)), dp, size, isg, uw, p.Bpad, ureg, w01);
                else
                    level_batches<SC_THREADS, SC_THREADS, R1>(sv + (SC_THREADS << (lev + 1)), dp, size, isg, uw, p.Bpad, ureg, w01);
            }
            // ---- the node itself -----------------------------------------------------------------------
            if (R1 && e.kind == NODE_RATE1) {
                // element h of the node's vector: level l of the tree, or the kernel's input when the node is the whole block
                const double *np_;
                int64_t nstr;
                if (l == n) {
                    np_ = (const double *)p.in_t + col;
                    nstr = p.Bpad;
                } else {
                    np_ = lvl_ptr(l);
                    nstr = lvl_stride(l);
                }
                unsigned long long soft = 0ULL;  // OR of the magnitudes: 0 iff every element is +-0
                if (l < 5) {
                    uint32_t w = 0;
#pragma unroll 1
                    for (int h = 0; h < (1 << l); ++h) {
                        const double v = np_[h * nstr];
                        soft |= (unsigned long long)__double_as_longlong(v) << 1;
                        w |= d_sign(v) << h;
                    }
                    if (!__all_sync(0xffffffffu, soft == 0ULL)) {
                        resume = l - 1;
                        continue;
                    }
                    cwreg |= w << (i & 31);
                } else {
#pragma unroll 1
                    for (int w0 = 0; w0 < (1 << (l - 5)); ++w0) {
                        uint32_t w = 0;
#pragma unroll 8
                        for (int h = 0; h < 32; ++h) {
                            const double v = np_[(int64_t)(32 * w0 + h) * nstr];
                            soft |= (unsigned long long)__double_as_longlong(v) << 1;
                            w |= d_sign(v) << h;
                        }
                        xw[(int64_t)((i >> 5) + w0) * p.Bpad] = w;  // overwritten by the leaf walk if the node is not hard
                    }
                    if (!__all_sync(0xffffffffu, soft == 0ULL)) {
                        resume = l - 1;
                        continue;
                    }
                }
                ei += (int)e.bits;  // skip the sub-tree's entries
            } else if (MODE != MODE_DECODE) {
                // genie pass (BinaryPolarEncoderDecoder.py:114-178): every index is frozen to a known bit; the leaf's
                // probabilities P(u_i | u_0^{i-1}, y) are captured (:268-273) as the packed level-0 value
                const double v0 = sv[SC_THREADS];
                if (p.marg_t) p.marg_t[(int64_t)i * p.Bpad + col] = v0;
                if (MODE != MODE_DUAL && (i & 31) == 0) gword = p.u_t[(int64_t)(i >> 5) * p.Bpad + col];
                uint32_t bit = (gword >> (i & 31)) & 1u;
                if (MODE == MODE_DUAL || MODE == MODE_PRIOR) {
                    const bool frozen = p.fmask[i] != 0;
                    const int64_t rrow = col < p.frames ? col : p.frames - 1;
                    const uint32_t rule = leaf_m0(v0) >= p.rnd[rrow * p.rnd_stride + i] ? 0u : 1u;
                    if (MODE == MODE_PRIOR) {
                        if (frozen) bit = rule;
                    } else {
                        const uint32_t mine = (lane & 1) ? rule : d_sign(v0);
                        bit = __shfl_sync(0xffffffffu, mine, frozen ? (lane | 1) : (lane & ~1));
                        if (!frozen) {
                            infoacc |= bit << (icount & 31);
                            if ((++icount & 31) == 0) {
                                iw[(int64_t)((icount >> 5) - 1) * p.Bpad] = infoacc;
                                infoacc = 0;
                            }
                        }
                    }
                }
                cwreg |= bit << (i & 31);
            } else if (!R1 && e.kind == NODE_BLOCK) {
                constexpr uint32_t BM = (1u << (1 << SC_LB)) - 1u;
                uint32_t u8, x8;
                LeafBlock<SC_LB, SC_THREADS>::run(sv, e.bits & BM, (e.bits >> (1 << SC_LB)) & BM, u8, x8);
                uint32_t m = ~e.bits & BM, packed = 0;  // the block's information bits, in u order
                int cnt = 0;
                while (m) {
                    packed |= ((u8 >> (__ffs(m) - 1)) & 1u) << cnt;
                    ++cnt;
                    m &= m - 1;
                }
                const int sh = icount & 31;
                infoacc |= packed << sh;
                if (sh + cnt >= 32) {
                    iw[(int64_t)(icount >> 5) * p.Bpad] = infoacc;
                    infoacc = sh + cnt > 32 ? packed >> (32 - sh) : 0u;
                }
                icount += cnt;
                cwreg |= x8 << (i & 31);
            } else if (e.kind == NODE_INFO) {
                const uint32_t bit = d_sign(sv[SC_THREADS]);  // level 0: p0 >= p1 -> 0 (ties and (0,0) -> 0), :252
                infoacc |= bit << (icount & 31);
                if ((++icount & 31) == 0) {
                    iw[(int64_t)((icount >> 5) - 1) * p.Bpad] = infoacc;
                    infoacc = 0;
                }
                cwreg |= bit << (i & 31);
            } else if (l < 5) {
                cwreg |= e.bits << (i & 31);
            } else {
#pragma unroll 1
                for (int w = 0; w < (1 << (l - 5)); ++w) xw[(int64_t)((i >> 5) + w) * p.Bpad] = p.r0_words[e.bits + w];
            }
            // ---- partial sums: x[lo, lo+s) ^= x[lo+s, lo+2s) whenever a plus child completes -----------
            int lv = l, ii = i;
            while (lv < 5 && lv < n && ((ii >> lv) & 1)) {  // inside the register word
                const int s = 1 << lv;
                const int sh = (ii - s) & 31;
                cwreg ^= ((cwreg >> (sh + s)) & ((1u << s) - 1u)) << sh;
                ii -= s;
                ++lv;
            }
            const int end = i + (1 << l);
            if (l < 5 && ((end & 31) == 0 || end == N)) {  // the register word is complete
                xw[(int64_t)((end - 1) >> 5) * p.Bpad] = cwreg;
                cwreg = 0;
            }
            while (lv < n && ((ii >> lv) & 1)) {  // whole words (lv >= 5 here)
                const int s = 1 << lv;
                uint32_t *lo = xw + (int64_t)((ii - s) >> 5) * p.Bpad;
#pragma unroll 1
                for (int w = 0; w < (s >> 5); ++w) lo[(int64_t)w * p.Bpad] ^= lo[(int64_t)(w + (s >> 5)) * p.Bpad];
                ii -= s;
                ++lv;
            }
        }
        if ((MODE == MODE_DECODE || MODE == MODE_DUAL) && (icount & 31)) iw[(int64_t)(icount >> 5) * p.Bpad] = infoacc;
    }
}

// ---- sub-block decoder on state codes (hybrid decoder, erasure-type channels) ----------------------------------------
// The same schedule walk as sc_decode_kernel<.., R1> (frame per lane, rate-0 pruning, exact rate-1 shortcut), but a node
// vector is one BYTE per element, four consecutive elements per 32-bit word, updated four at a time by f8 / g8.  The whole
// tree below level 9 is 516 bytes per frame and lives in shared memory ([word][thread]); level 9 (128 words per frame) in a
// per-warp global scratch that stays in L2.  Input: the level-n vector as [2^n / 4][Bpad] words (hy_level8_kernel<PACK4>).
// Output: the codeword words (partial sums) only -- the hybrid decoder takes the information bits from their transform.
constexpr int S8_LS = 8;                                  // levels <= S8_LS in shared memory
constexpr int S8_WORDS = 2 + ((1 << (S8_LS - 1)) - 1);    // level 0, level 1, then 2^(l-2) words for l = 2..S8_LS
__device__ __forceinline__ uint32_t expand4(uint32_t b) { return (b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21); }
__device__ __forceinline__ uint32_t signs4(uint32_t w) {  // bit 0 of each byte -> 4 bits
    const uint32_t x = w & 0x01010101u;
    return (x | (x >> 7) | (x >> 14) | (x >> 21)) & 15u;
}

__global__ void __launch_bounds__(SC_THREADS, 1) sc_decode8_kernel(const ScParams p) {
    extern __shared__ uint32_t sm_w8[];  // [S8_WORDS][SC_THREADS]
    const int n = p.n, N = 1 << n;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * (SC_THREADS / 32) + (threadIdx.x >> 5);
    const int warps_total = gridDim.x * (SC_THREADS / 32);
    const int64_t groups = (p.frames + 31) / 32;
    uint32_t *sv = sm_w8 + threadIdx.x;
    uint32_t *gv = (uint32_t *)p.vals + (int64_t)warp_global * (N >> 3) * 32 + lane;  // level n-1: 2^(n-3) words per lane
    // words of level lev (2 <= lev < n): base pointer and word stride
    auto lvl_ptr = [&](int lev) -> uint32_t * {
        return lev > S8_LS ? gv : sv + (lev >= 2 ? (1 + (1 << (lev - 2))) : lev) * SC_THREADS;
    };
    auto lvl_stride = [&](int lev) -> int64_t { return lev > S8_LS ? 32 : SC_THREADS; };

#pragma unroll 1
    for (int64_t grp = warp_global; grp < groups; grp += warps_total) {
        const int64_t col = grp * 32 + lane;
        const uint32_t *in_w = (const uint32_t *)p.in_t + col;  // word g of the input at in_w[g * Bpad]
        uint32_t *xw = p.cw_t + col;
        uint32_t cwreg = 0;
        int resume = -1;
#pragma unroll 1
        for (int ei = 0; ei < p.n_sched; ++ei) {
            const SchedEntry e = p.sched[ei];
            const int i = e.i, l = e.l, top = e.top;
            const int stop = e.kind == NODE_RATE0 ? l + 1 : l;
            int lev = i == 0 ? n - 1 : (top >= stop ? top : -1);
            bool isg = i != 0;
            if (resume >= 0) {
                lev = resume;
                isg = false;
                resume = -1;
            }
#pragma unroll 1
            for (; lev >= stop; --lev, isg = false) {
                const int size = 1 << lev;
                const uint32_t *uw = xw + (int64_t)((i - size) >> 5) * p.Bpad;  // g: decision words of whole-word levels
                const uint32_t ureg = cwreg >> ((i - size) & 31);               // g: decision bits of levels below 32
                // one code path per (source, destination) memory pair, so that shared-memory levels use shared-memory
                // instructions with compile-time strides: words of level lev from level lev + 1
                auto run = [&](const uint32_t *sp, const int64_t ss, uint32_t *dp, const int64_t ds) {
                    if (lev < 2) {  // the source level is a single word: elements (0,2),(1,3) of level 2 / (0,1) of level 1
                        const uint32_t a = sp[0];
                        const uint32_t b = a >> (lev == 1 ? 16 : 8), u = expand4(ureg & 15u);
                        dp[0] = isg ? g8(a, b, u) : f8(a, b);
                        return;
                    }
                    const int nw = size >> 2;
                    const uint32_t *sp2 = sp + (int64_t)nw * ss;
                    if (nw < 8) {
#pragma unroll 1
                        for (int w = 0; w < nw; ++w) {
                            const uint32_t a = sp[w * ss], b = sp2[w * ss], u = expand4((ureg >> (4 * w)) & 15u);
                            dp[w * ds] = isg ? g8(a, b, u) : f8(a, b);
                        }
                        return;
                    }
#pragma unroll 1
                    for (int w0 = 0; w0 < nw; w0 += 8) {  // eight words = 32 elements = one word of decision bits
                        uint32_t ub = 0;
                        if (isg) ub = size >= 32 ? uw[(int64_t)(w0 >> 3) * p.Bpad] : ureg;
                        uint32_t a[8], b[8];
#pragma unroll
                        for (int w = 0; w < 8; ++w) {
                            a[w] = sp[(w0 + w) * ss];
                            b[w] = sp2[(w0 + w) * ss];
                        }
#pragma unroll
                        for (int w = 0; w < 8; ++w)
                            dp[(w0 + w) * ds] = isg ? g8(a[w], b[w], expand4((ub >> (4 * w)) & 15u)) : f8(a[w], b[w]);
                    }
                };
                auto soff = [](int l) { return (l >= 2 ? (1 + (1 << (l - 2))) : l) * SC_THREADS; };  // shared-memory word offset of level l
                if (lev + 1 == n)
                    run(in_w, p.Bpad, lev > S8_LS ? gv : sv + soff(lev), lev > S8_LS ? 32 : SC_THREADS);
                else if (lev + 1 > S8_LS)
                    run(gv, 32, sv + soff(lev), SC_THREADS);
                else
                    run(sv + soff(lev + 1), SC_THREADS, sv + soff(lev), SC_THREADS);
            }
            // ---- the node itself -----------------------------------------------------------------------
            if (e.kind == NODE_RATE1) {
                const uint32_t *np_;
                int64_t nstr;
                if (l == n) {
                    np_ = in_w;
                    nstr = p.Bpad;
                } else {
                    np_ = lvl_ptr(l);
                    nstr = lvl_stride(l);
                }
                uint32_t soft = 0;  // erasure / contradiction bits of any element
                if (l < 5) {
                    uint32_t w = 0;
#pragma unroll 1
                    for (int q = 0; q < (1 << (l - 2)); ++q) {
                        const uint32_t v = np_[q * nstr];
                        soft |= v & 0x06060606u;
                        w |= signs4(v) << (4 * q);
                    }
                    if (!__all_sync(0xffffffffu, soft == 0u)) {
                        resume = l - 1;
                        continue;
                    }
                    cwreg |= w << (i & 31);
                } else {
#pragma unroll 1
                    for (int w0 = 0; w0 < (1 << (l - 5)); ++w0) {
                        uint32_t v[8], w = 0;
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = np_[(int64_t)(8 * w0 + q) * nstr];
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            soft |= v[q] & 0x06060606u;
                            w |= signs4(v[q]) << (4 * q);
                        }
                        xw[(int64_t)((i >> 5) + w0) * p.Bpad] = w;  // overwritten by the leaf walk if the node is not hard
                    }
                    if (!__all_sync(0xffffffffu, soft == 0u)) {
                        resume = l - 1;
                        continue;
                    }
                }
                ei += (int)e.bits;  // skip the sub-tree's entries
            } else if (e.kind == NODE_INFO) {
                cwreg |= (sv[0] & 1u) << (i & 31);  // level 0: side bit; erasure (tie) and (0,0) -> 0
            } else if (l < 5) {
                cwreg |= e.bits << (i & 31);
            } else {
#pragma unroll 1
                for (int w = 0; w < (1 << (l - 5)); ++w) xw[(int64_t)((i >> 5) + w) * p.Bpad] = p.r0_words[e.bits + w];
            }
            // ---- partial sums: x[lo, lo+s) ^= x[lo+s, lo+2s) whenever a plus child completes -----------
            int lv = l, ii = i;
            while (lv < 5 && lv < n && ((ii >> lv) & 1)) {
                const int s = 1 << lv;
                const int sh = (ii - s) & 31;
                cwreg ^= ((cwreg >> (sh + s)) & ((1u << s) - 1u)) << sh;
                ii -= s;
                ++lv;
            }
            const int end = i + (1 << l);
            if (l < 5 && ((end & 31) == 0 || end == N)) {
                xw[(int64_t)((end - 1) >> 5) * p.Bpad] = cwreg;
                cwreg = 0;
            }
            while (lv < n && ((ii >> lv) & 1)) {
                const int s = 1 << lv;
                uint32_t *lo = xw + (int64_t)((ii - s) >> 5) * p.Bpad;
#pragma unroll 1
                for (int w = 0; w < (s >> 5); ++w) lo[(int64_t)w * p.Bpad] ^= lo[(int64_t)(w + (s >> 5)) * p.Bpad];
                ii -= s;
                ++lv;
            }
        }
    }
}

// ---- ingest: caller layout [frames][N] -> transposed, bit-reversed [N][Bpad] --------------------------
template <class T>
__global__ void __launch_bounds__(256) ingest_kernel(int n, int64_t frames, int64_t Bpad, const T *__restrict__ in,
                                                     T *__restrict__ out, T pad_value, int ymax = 255) {
    __shared__ T tile[32][33];
    const int N = 1 << n;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int ptiles = (N + 31) / 32;
    for (int pt = blockIdx.y; pt < ptiles; pt += gridDim.y) {
        const int i0 = pt * 32;
        for (int r = ty; r < 32; r += 8) {
            const int64_t f = f0 + r;
            const int pos = i0 + tx;
            T v = (f < frames && pos < N) ? in[f * N + pos] : pad_value;
            if constexpr (sizeof(T) == 1) v = v > (T)ymax ? (T)ymax : v;  // symbols index a Y-row table: out-of-range bytes read row Y-1
            tile[r][tx] = v;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int pos = i0 + r;
            if (pos < N) out[(int64_t)bitrev_n((uint32_t)pos, n) * Bpad + f0 + tx] = tile[tx][r];
        }
        __syncthreads();
    }
}

// uint8 symbols, N >= 128, 4-byte aligned input: 128 frames x 128 positions per tile, 4-byte global accesses on both
// sides (a frame row is read as words, an output row is written as 4 frames per lane).
__global__ void __launch_bounds__(256) ingest_u8_kernel(int n, int64_t frames, int64_t Bpad, const uint8_t *__restrict__ in,
                                                        uint8_t *__restrict__ out, int ymax) {
    __shared__ uint32_t tile[128][33];  // row (f & 3) * 32 + (f >> 2): the four frames of a lane sit 32 rows apart
    const int N = 1 << n;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * 128;
    const uint32_t *in32 = (const uint32_t *)in;
    for (int pt = blockIdx.y; pt < (N >> 7); pt += gridDim.y) {
        const int i0 = pt * 128;
#pragma unroll 4
        for (int r = wrp; r < 128; r += 8) {
            const int64_t f = f0 + r;
            // symbols index a Y-row table: out-of-range bytes read row Y-1 (the ABI requires y < Y; this keeps the lookups in bounds)
            tile[(r & 3) * 32 + (r >> 2)][lane] = f < frames ? __vminu4(in32[(f * N + i0) / 4 + lane], (uint32_t)ymax * 0x01010101u) : 0u;
        }
        __syncthreads();
        if (f0 + 4 * lane < Bpad) {
#pragma unroll 4
            for (int q = wrp; q < 128; q += 8) {
                const int w = q >> 2, sh = 8 * (q & 3);
                const uint32_t v = ((tile[lane][w] >> sh) & 0xffu) | (((tile[32 + lane][w] >> sh) & 0xffu) << 8) |
                                   (((tile[64 + lane][w] >> sh) & 0xffu) << 16) | (((tile[96 + lane][w] >> sh) & 0xffu) << 24);
                *(uint32_t *)(out + (int64_t)bitrev_n((uint32_t)(i0 + q), n) * Bpad + f0 + 4 * lane) = v;
            }
        }
        __syncthreads();
    }
}

// ---- egress of the codeword with the bit-reversal permutation, n >= 10: out[i] = nat[rev_n(i)] splits into 2^(n-10)
// independent 32 x 32 bit-matrix transposes (word index and bit index swap roles, each 5-bit field reversed), done in
// registers: 5 x 16 masked swaps instead of 1024 single-bit gathers per group of 32 words.
__global__ void __launch_bounds__(256) egress_bitrev_kernel(int n, int64_t frames, int64_t Bpad, const uint32_t *__restrict__ in_t,
                                                            uint32_t *__restrict__ out) {
    __shared__ uint32_t tile[8][32][33];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int G = 1 << (n - 10), W = 1 << (n - 5);
    const int64_t col = f0 + lane;
    for (int g = blockIdx.y * 8 + wrp; g < G; g += gridDim.y * 8) {
        uint32_t A[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) A[k] = in_t[(int64_t)((31 - (int)(__brev((uint32_t)k) >> 27)) * G + g) * Bpad + col];
        // Hacker's Delight transpose32 (anti-transpose in LSB-first numbering: T[c] bit r = A[31-r] bit 31-c)
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) {
            const int j = 16 >> jj;
            const uint32_t m = jj == 0 ? 0x0000ffffu : jj == 1 ? 0x00ff00ffu : jj == 2 ? 0x0f0f0f0fu : jj == 3 ? 0x33333333u : 0x55555555u;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                if (!(k & j)) {
                    const uint32_t t = (A[k] ^ (A[k + j] >> j)) & m;
                    A[k] ^= t;
                    A[k + j] ^= t << j;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) tile[wrp][lane][c] = A[c];
        __syncwarp();
        const int gw = n > 10 ? (int)(__brev((uint32_t)g) >> (32 - (n - 10))) : 0;
        const int ow = (31 - (int)(__brev((uint32_t)lane) >> 27)) * G + gw;  // output word written by this lane
        for (int r = 0; r < 32; ++r) {
            const int64_t f = f0 + r;
            if (f < frames) out[f * W + ow] = tile[wrp][r][lane];
        }
        __syncwarp();
    }
}

// ---- egress: [W][Bpad] words -> [frames][W], optionally applying the bit-reversal permutation ---------
template <bool BITREV>
__global__ void __launch_bounds__(256) egress_kernel(int n, int W, int64_t frames, int64_t Bpad,
                                                     const uint32_t *__restrict__ in_t, uint32_t *__restrict__ out) {
    __shared__ uint32_t tile[8][32][33];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    const int jtiles = (W + 31) / 32;
    for (int jt = blockIdx.y * 8 + wrp; jt < jtiles; jt += gridDim.y * 8) {
        const int j0 = jt * 32;
        const int64_t col = f0 + lane;
        for (int t = 0; t < 32; ++t) {
            const int j = j0 + t;
            uint32_t v = 0;
            if (j < W) {
                if (BITREV)
                    v = bitrev_gather_word([&](uint32_t wi) { return in_t[(int64_t)wi * Bpad + col]; }, n, (uint32_t)j);
                else
                    v = in_t[(int64_t)j * Bpad + col];
            }
            tile[wrp][lane][t] = v;
        }
        __syncwarp();
        for (int r = 0; r < 32; ++r) {
            const int64_t f = f0 + r;
            if (f < frames && j0 + lane < W) out[f * W + j0 + lane] = tile[wrp][r][lane];
        }
        __syncwarp();
    }
}

struct ScLayout {
    int64_t chunk, Bpad;
    size_t off_in, off_cw, off_info, off_vals, total;
    int grid;
};

static int sc_grid_max() { return num_sms() * SC_BLOCKS_PER_SM; }

static ScLayout sc_layout(const pc_plan *plan, int64_t chunk, int kind) {
    ScLayout L;
    const int64_t N = plan->N, Nw = (N + 31) / 32, Kw = (plan->k + 31) / 32 > 0 ? (plan->k + 31) / 32 : 1;
    L.chunk = chunk;
    L.Bpad = round_up(chunk, 32);
    const int64_t blocks = (L.Bpad + SC_THREADS - 1) / SC_THREADS;
    L.grid = (int)(blocks < sc_grid_max() ? blocks : sc_grid_max());
    const int64_t gvals = N > (1 << (LS + 1)) ? N - (1 << (LS + 1)) : 0;
    size_t o = 0;
    L.off_in = o;
    o += align256((size_t)N * L.Bpad * (kind == PC_INPUT_SYMBOLS ? 1 : 16));
    L.off_cw = o;
    o += align256((size_t)Nw * L.Bpad * 4);
    L.off_info = o;
    o += align256((size_t)Kw * L.Bpad * 4);
    L.off_vals = o;
    o += align256((size_t)L.grid * (SC_THREADS / 32) * gvals * 32 * 8 + 256);
    L.total = o;
    return L;
}

// Frames per launch: a whole multiple of the resident lane count (so every persistent warp walks the same
// number of 32-frame groups), capped so that the transposed input stays around 1-4 GiB.
static int64_t sc_pick_chunk(int64_t B, int kind) {
    const int64_t wave = (int64_t)sc_grid_max() * SC_THREADS;
    const int64_t cap = kind == PC_INPUT_SYMBOLS ? (1 << 20) : (1 << 18);
    int64_t chunk = round_up(B, 32);
    if (chunk > cap + cap / 2) chunk = cap >= wave ? cap / wave * wave : cap;  // up to 1.5 cap goes in one launch
    return chunk;
}

// Large blocks (and, on request, any block of at least 64 symbols) take the frame-per-CTA streamed decoder:
// PC_SC_STREAM=1 forces it, PC_SC_STREAM=0 forbids it below the frame-per-lane limit.
static bool sc_use_hybrid(const pc_plan *plan, int64_t B, int kind);
static size_t sc_hybrid_workspace_bytes(const pc_plan *plan, int64_t B, int et = -1);
static int sc_hybrid_decode(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_table, int Y, uint32_t *d_cw,
                            uint32_t *d_info, void *ws, size_t ws_bytes, cudaStream_t st);

static bool sc_use_stream(const pc_plan *plan, int64_t B) {
    if (!sc_stream_supported(plan)) return false;
    if (plan->n > SC_MAX_N) return true;
    const char *s = getenv("PC_SC_STREAM");
    if (s && *s) return atoi(s) != 0;
    return plan->n >= 14 && B < 2048;  // few long frames: not enough frames to fill 32-frame warps
}

static int sc_decode_common(const pc_plan *plan, int kind, const void *d_in, int64_t B, const double *h_table, int Y,
                            uint32_t *d_cw, uint32_t *d_info, void *ws, size_t ws_bytes, cudaStream_t st) {
    PC_REQUIRE(plan && plan->q == 2, "binary plan required");
    PC_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_in && d_cw && (d_info || plan->k == 0) && ws, "null buffer");
    PC_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    if (kind == PC_INPUT_SYMBOLS) PC_REQUIRE(h_table && Y >= 1 && Y <= 16, "symbol table must have 1..16 rows");
    if (sc_use_hybrid(plan, B, kind))
        return sc_hybrid_decode(plan, (const uint8_t *)d_in, B, h_table, Y, d_cw, d_info, ws, ws_bytes, st);
    if (sc_use_stream(plan, B)) return sc_stream_decode(plan, kind, d_in, B, h_table, Y, d_cw, d_info, ws, ws_bytes, st);
    PC_REQUIRE(plan->n <= SC_MAX_N, "block length too large for the frame-per-lane SC decoder");
    // largest chunk (multiple of 32 frames) that fits the workspace
    int64_t chunk = sc_pick_chunk(B, kind);
    while (chunk > 32 && sc_layout(plan, chunk, kind).total > ws_bytes) chunk = round_up(chunk / 2, 32);
    ScLayout L = sc_layout(plan, chunk, kind);
    if (L.total > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed for a 32-frame chunk", ws_bytes, L.total);
        return PC_ERR_NOMEM;
    }
    const int N = plan->N, Nw = (N + 31) / 32, Kw = (plan->k + 31) / 32;
    char *base = (char *)ws;
    ScParams p{};
    p.n = plan->n;
    p.k = plan->k;
    p.n_sched = (int)plan->sched.size();
    p.Y = Y;
    p.Bpad = L.Bpad;
    p.sched = plan->d_sched;
    const char *bl = getenv("PC_SC_BLOCK");  // 0: the leaf-by-leaf schedule (tests compare the two)
    if (plan->d_sched_b && !(bl && atoi(bl) == 0)) {
        p.sched = plan->d_sched_b;
        p.n_sched = (int)plan->sched_b.size();
    }
    p.r0_words = plan->d_r0_words;
    p.in_t = base + L.off_in;
    p.vals = (double *)(base + L.off_vals);
    p.cw_t = (uint32_t *)(base + L.off_cw);
    p.info_t = (uint32_t *)(base + L.off_info);
    for (int i = 0; i < 32; ++i) p.table[i] = (kind == PC_INPUT_SYMBOLS && i < 2 * Y) ? h_table[i] : 0.0;
    const size_t smem = (size_t)SMEM_VALS * SC_THREADS * sizeof(double);
    if (kind == PC_INPUT_SYMBOLS)
        PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<PC_INPUT_SYMBOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
        PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<PC_INPUT_PROBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t f0 = 0; f0 < B; f0 += chunk) {
        const int64_t frames = (B - f0) < chunk ? (B - f0) : chunk;
        const int64_t tiles = (frames + 31) / 32;
        p.frames = frames;
        const int ptiles = (N + 31) / 32;
        dim3 ig((unsigned)tiles, (unsigned)(ptiles < 64 ? ptiles : 64));
        if (kind == PC_INPUT_SYMBOLS && N >= 128 && (((uintptr_t)d_in + (size_t)f0 * N) & 3) == 0) {
            dim3 ig8((unsigned)((frames + 127) / 128), (unsigned)((N >> 7) < 16 ? (N >> 7) : 16));
            ingest_u8_kernel<<<ig8, 256, 0, st>>>(plan->n, frames, L.Bpad, (const uint8_t *)d_in + f0 * N, (uint8_t *)p.in_t, Y - 1);
        } else if (kind == PC_INPUT_SYMBOLS) {
            ingest_kernel<uint8_t><<<ig, 256, 0, st>>>(plan->n, frames, L.Bpad, (const uint8_t *)d_in + f0 * N,
                                                       (uint8_t *)p.in_t, (uint8_t)0, Y - 1);
        } else {
            ingest_kernel<double2><<<ig, 256, 0, st>>>(plan->n, frames, L.Bpad, (const double2 *)d_in + f0 * N,
                                                       (double2 *)p.in_t, make_double2(0.5, 0.5));
        }
        PC_LAUNCH_CHECK();
        const int64_t blocks = (tiles * 32 + SC_THREADS - 1) / SC_THREADS;
        const int grid = (int)(blocks < L.grid ? blocks : L.grid);
        prof_mark(st);
        if (kind == PC_INPUT_SYMBOLS)
            sc_decode_kernel<PC_INPUT_SYMBOLS><<<grid, SC_THREADS, smem, st>>>(p);
        else
            sc_decode_kernel<PC_INPUT_PROBS><<<grid, SC_THREADS, smem, st>>>(p);
        prof_mark(st);
        PC_LAUNCH_CHECK();
        const int jt_cw = (Nw + 255) / 256;
        if (plan->n >= 10) {
            const int G = 1 << (plan->n - 10);
            egress_bitrev_kernel<<<dim3((unsigned)tiles, (unsigned)((G + 7) / 8 < 8 ? (G + 7) / 8 : 8)), 256, 0, st>>>(
                plan->n, frames, L.Bpad, p.cw_t, d_cw + f0 * Nw);
        } else {
            egress_kernel<true><<<dim3((unsigned)tiles, (unsigned)jt_cw), 256, 0, st>>>(plan->n, Nw, frames, L.Bpad, p.cw_t,
                                                                                         d_cw + f0 * Nw);
        }
        PC_LAUNCH_CHECK();
        if (Kw > 0) {
            const int jt_i = (Kw + 255) / 256;
            egress_kernel<false><<<dim3((unsigned)tiles, (unsigned)jt_i), 256, 0, st>>>(plan->n, Kw, frames, L.Bpad,
                                                                                         p.info_t, d_info + f0 * Kw);
            PC_LAUNCH_CHECK();
        }
    }
    return PC_OK;
}

// ======================================================================================================================
// Hybrid decoder for large blocks (N = 2^17 .. 2^24) over discrete channels: the SAME natural-order algorithm as
// sc_decode_kernel, but the levels >= HY_L0 -- where a level has thousands of elements per frame -- run as batch-wide,
// element-parallel kernels over [element][frame] vectors streamed through HBM (coalesced over the frames, every lane busy),
// and the 2^(n - HY_L0) sub-blocks of 2^HY_L0 leaves are decoded by sc_decode_kernel<SC_INPUT_PACKED> (frame per lane),
// which reads its top level straight from the level-HY_L0 vector.  The host walks the top of the tree exactly like the
// kernel's schedule loop does (g at level ctz, f down, leaf block, partial sums) -- every frame of the batch executes the
// same sequence.  Versus the frame-per-CTA streamed decoder (sc_stream.cu) the leaf walk is no longer one warp per frame
// with 1-16 active lanes: throughput is set by how many frames fit in memory (5.4 MB of float64 state per 2^20 frame).
// Three walks share this host loop:
//   * float64 states, any discrete channel: hy_level_sym_kernel / hy_level_kernel + sc_decode_kernel<packed>;
//   * channels with a hard output symbol: the sub-block schedules carry NODE_RATE1 markers (exact rate-1 shortcut,
//     sc_decode_kernel<packed, decode, R1>), warps whose sub-block input is hard / erased / (0,0) only use node01;
//   * erasure-type channels (every table row hard knowledge or an exact erasure; n >= 13): one-byte state codes all the
//     way -- hy_level_sym8_kernel / hy_level8_kernel + sc_decode8_kernel -- 1.9 MB per 2^20 frame, batches of one sub-block
//     wave.  PC_SC_R1=0 / PC_SC_HY8=0 select the plainer walks (the tests compare all three).
constexpr int HY_L0 = 10;

struct HybridTables {
    std::vector<pc_plan *> sub;       // one plan per sub-block of 2^HY_L0 leaves
    std::vector<uint8_t> all_frozen;  // the sub-block is a rate-0 node
    int32_t *d_info_pos = nullptr;    // [k] u index of information bit j
    // schedules with a NODE_RATE1 entry in front of every all-information sub-tree of 4+ leaves (sc_decode_kernel<.., R1>)
    SchedEntry *d_sched_r1 = nullptr;
    std::vector<int32_t> r1_off, r1_len;  // per sub-block: first entry / number of entries in d_sched_r1
};

// sub-plan schedule -> the same walk with NODE_RATE1 markers; `at` walks sp->sched in step with the recursion
static void build_r1_schedule(const pc_plan *sp, int i, int l, size_t &at, std::vector<SchedEntry> &out) {
    const SchedEntry &e = sp->sched[at];
    if (e.i == i && e.l == l) {  // a leaf or a maximal rate-0 node
        out.push_back(e);
        ++at;
        return;
    }
    bool all_info = true;
    for (int j = i; j < i + (1 << l) && all_info; ++j) all_info = !sp->frozen_mask[j];
    const size_t mark = out.size();
    if (all_info && l >= 2) {
        SchedEntry m{};
        m.i = i;
        m.l = (int8_t)l;
        m.kind = NODE_RATE1;
        m.top = (int8_t)(i == 0 ? sp->n : __builtin_ctz((unsigned)i));
        out.push_back(m);
    }
    build_r1_schedule(sp, i, l - 1, at, out);
    build_r1_schedule(sp, i + (1 << (l - 1)), l - 1, at, out);
    if (all_info && l >= 2) out[mark].bits = (uint32_t)(out.size() - mark - 1);
}
static std::mutex g_hy_mu;
static std::map<const pc_plan *, HybridTables *> g_hy_tables;

static HybridTables *hybrid_tables(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_hy_mu);
    auto it = g_hy_tables.find(p);
    if (it != g_hy_tables.end()) return it->second;
    HybridTables *T = new HybridTables();
    // every failure path below releases what was built so far (sub-plans, device buffers): nothing is cached half-built
    auto fail = [&](const char *why) -> HybridTables * {
        if (why) set_error("%s", why);
        for (pc_plan *q : T->sub) pc_plan_destroy(q);
        if (T->d_info_pos) cudaFree(T->d_info_pos);
        if (T->d_sched_r1) cudaFree(T->d_sched_r1);
        delete T;
        return nullptr;
    };
    const int Ns = 1 << HY_L0, NS = p->N >> HY_L0;
    for (int j = 0; j < NS; ++j) {
        pc_plan *sp = nullptr;
        if (pc_plan_create(2, HY_L0, p->frozen_mask.data() + (size_t)j * Ns, p->frozen_vals.data() + (size_t)j * Ns, &sp) != PC_OK)
            return fail(nullptr);
        T->sub.push_back(sp);
        T->all_frozen.push_back(sp->k == 0 ? 1 : 0);
    }
    std::vector<SchedEntry> r1;
    for (int j = 0; j < NS; ++j) {
        size_t at = 0;
        T->r1_off.push_back((int32_t)r1.size());
        build_r1_schedule(T->sub[j], 0, HY_L0, at, r1);
        T->r1_len.push_back((int32_t)r1.size() - T->r1_off.back());
    }
    if (cudaMalloc((void **)&T->d_sched_r1, r1.size() * sizeof(SchedEntry)) != cudaSuccess ||
        cudaMemcpy(T->d_sched_r1, r1.data(), r1.size() * sizeof(SchedEntry), cudaMemcpyHostToDevice) != cudaSuccess)
        return fail("hybrid tables: device upload failed");
    std::vector<int32_t> pos;
    for (int i = 0; i < p->N; ++i)
        if (!p->frozen_mask[i]) pos.push_back(i);
    if (pos.empty()) pos.push_back(0);
    if (cudaMalloc((void **)&T->d_info_pos, pos.size() * 4) != cudaSuccess ||
        cudaMemcpy(T->d_info_pos, pos.data(), pos.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess)
        return fail("hybrid tables: device upload failed");
    g_hy_tables[p] = T;
    return T;
}

void hybrid_tables_release(const pc_plan *p) {
    HybridTables *T = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_hy_mu);
        auto it = g_hy_tables.find(p);
        if (it == g_hy_tables.end()) return;
        T = it->second;
        g_hy_tables.erase(it);
    }
    for (pc_plan *q : T->sub) pc_plan_destroy(q);
    cudaFree(T->d_info_pos);
    cudaFree(T->d_sched_r1);
    delete T;
}

// the channel's joint-probability table, passed by value to the lookup-table kernel
struct HyRootParams {
    double table[32];
};

// Level n-1 is never stored: an element of it is a function of two channel symbols and (in the second half of the frame) one
// decision bit, looked up in lut[mode][y_a][y_b] (mode 0: f, 1 + u: g) -- built on the device with f_raw / g_raw, so the
// bits are those of hy_root_kernel.  Level n-2 is then computed straight from four symbols per element.
__global__ void hy_lut_kernel(int Y, const HyRootParams tp, double *__restrict__ lut) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= 3 * Y * Y) return;
    const int m = idx / (Y * Y), ya = (idx / Y) % Y, yb = idx % Y;
    const double a0 = tp.table[2 * ya], a1 = tp.table[2 * ya + 1], b0 = tp.table[2 * yb], b1 = tp.table[2 * yb + 1];
    const double v = m == 0 ? f_raw(a0, a1, b0, b1) : g_raw(a0, a1, b0, b1, (uint32_t)(m - 1));
    lut[idx] = v;
    ((uint8_t *)(lut + 768))[idx] = (uint8_t)state_code(v);  // erasure-type channels: the same table as state codes
}

// ---- erasure-type channels (every table row is hard knowledge, an exact erasure or (0,0)): every value of the tree is one of
// four states, so the upper stages keep ONE BYTE per element instead of a float64 and update four frames per 32-bit
// operation (f8 / g8); 8 x less HBM traffic and workspace than the float64 stages, identical decisions. ----

// level n-2 from the channel symbols: a thread owns 32 consecutive elements (one word of decision bits per operand) of four
// adjacent frames, so the decision words of levels n-1 / n-2 are loaded once per 32 elements and the symbol rows stream through
__global__ void __launch_bounds__(256) hy_level_sym8_kernel(int64_t quarter, int64_t Bpad, const uint8_t *__restrict__ sym, int Y,
                                                            const uint8_t *__restrict__ lut8, const uint32_t *__restrict__ x0,
                                                            int top_g, const uint32_t *__restrict__ xw, int isg,
                                                            uint8_t *__restrict__ out) {
    __shared__ uint8_t s_lut[768];
    for (int i = threadIdx.x; i < 3 * Y * Y; i += blockDim.x) s_lut[i] = lut8[i];
    __syncthreads();
    const int64_t Bq = Bpad >> 2;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (quarter >> 5) * Bq) return;
    const int64_t hb = gid / Bq, fq = gid - hb * Bq, half = 2 * quarter;
    const uint32_t *p0 = (const uint32_t *)sym + hb * 32 * Bq + fq;  // symbols of element h = 32 hb + k, frames 4 fq .. 4 fq + 3
    const uint32_t *p1 = p0 + half * Bq, *p2 = p0 + quarter * Bq, *p3 = p2 + half * Bq;
    uint32_t *o = (uint32_t *)out + hb * 32 * Bq + fq;
    uint4 wa = make_uint4(0, 0, 0, 0), wb = wa, wu = wa;
    if (top_g) {
        wa = *(const uint4 *)(x0 + hb * Bpad + 4 * fq);
        wb = *(const uint4 *)(x0 + (hb + (quarter >> 5)) * Bpad + 4 * fq);
    }
    if (isg) wu = *(const uint4 *)(xw + hb * Bpad + 4 * fq);
    const uint32_t wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w}, wuv[4] = {wu.x, wu.y, wu.z, wu.w};
    const uint32_t YY = (uint32_t)(Y * Y);
#pragma unroll 4
    for (int k = 0; k < 32; ++k) {
        const uint32_t y0 = p0[k * Bq], y1 = p1[k * Bq], y2 = p2[k * Bq], y3 = p3[k * Bq];
        uint32_t A = 0, Bv = 0, U = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t ma = top_g ? YY + YY * ((wav[j] >> k) & 1u) : 0u, mb = top_g ? YY + YY * ((wbv[j] >> k) & 1u) : 0u;
            A |= (uint32_t)s_lut[ma + ((y0 >> (8 * j)) & 255u) * Y + ((y1 >> (8 * j)) & 255u)] << (8 * j);
            Bv |= (uint32_t)s_lut[mb + ((y2 >> (8 * j)) & 255u) * Y + ((y3 >> (8 * j)) & 255u)] << (8 * j);
            U |= ((wuv[j] >> k) & 1u) << (8 * j);
        }
        o[k * Bq] = isg ? g8(A, Bv, U) : f8(A, Bv);
    }
}

// out[h][f] = node(in[h][f], in[h + size][f]): a thread owns 32 consecutive elements (one word of decision bits) of four
// adjacent frames.  PACK4 (the level the sub-block kernel reads): the output is [h / 4][f] WORDS of four consecutive elements
// (a 4 x 4 byte transpose in registers), the sub-block kernel's own packing.
template <bool PACK4>
__global__ void __launch_bounds__(256) hy_level8_kernel(int64_t size, int64_t Bpad, const uint8_t *__restrict__ in,
                                                        const uint32_t *__restrict__ xw, int isg, uint8_t *__restrict__ out) {
    const int64_t Bq = Bpad >> 2;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (size >> 5) * Bq) return;
    const int64_t hb = gid / Bq, fq = gid - hb * Bq;
    const uint32_t *a = (const uint32_t *)in + hb * 32 * Bq + fq, *b = a + size * Bq;
    uint32_t *o = (uint32_t *)out + hb * 32 * Bq + fq;
    uint4 wu = make_uint4(0, 0, 0, 0);
    if (isg) wu = *(const uint4 *)(xw + hb * Bpad + 4 * fq);
#pragma unroll 2
    for (int k0 = 0; k0 < 32; k0 += 4) {
        uint32_t r[4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = k0 + kk;
            const uint32_t av = a[k * Bq], bv = b[k * Bq];
            const uint32_t u = ((wu.x >> k) & 1u) | (((wu.y >> k) & 1u) << 8) | (((wu.z >> k) & 1u) << 16) | (((wu.w >> k) & 1u) << 24);
            r[kk] = isg ? g8(av, bv, u) : f8(av, bv);
        }
        if (PACK4) {
            // r[e] holds element k0 + e of frames 0..3 (byte j = frame j) -> t[j] holds frame j's elements k0..k0+3
            const uint32_t lo01 = __byte_perm(r[0], r[1], 0x5140), hi01 = __byte_perm(r[0], r[1], 0x7362);
            const uint32_t lo23 = __byte_perm(r[2], r[3], 0x5140), hi23 = __byte_perm(r[2], r[3], 0x7362);
            uint4 t;
            t.x = __byte_perm(lo01, lo23, 0x5410);
            t.y = __byte_perm(lo01, lo23, 0x7632);
            t.z = __byte_perm(hi01, hi23, 0x5410);
            t.w = __byte_perm(hi01, hi23, 0x7632);
            *(uint4 *)((uint32_t *)out + (hb * 8 + (k0 >> 2)) * Bpad + 4 * fq) = t;
        } else {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) o[(k0 + kk) * Bq] = r[kk];
        }
    }
}

// Level n-2 from the channel symbols with the four symbol rows of every element staged by bulk asynchronous copies: a ring of two
// slots (8 elements x 4 rows x 1 KB = 32 KB each); the slot an 8-element step has consumed is re-armed for the step after next.
__global__ void __launch_bounds__(256) hy_level_sym8_bulk_kernel(int64_t quarter, int64_t Bpad, const uint8_t *__restrict__ sym, int Y,
                                                                 const uint8_t *__restrict__ lut8, const uint32_t *__restrict__ x0,
                                                                 int top_g, const uint32_t *__restrict__ xw, int isg,
                                                                 uint8_t *__restrict__ out) {
    extern __shared__ __align__(128) uint32_t sm_rows[];  // [slot][stream][row][256]
    __shared__ uint64_t bars[2];
    __shared__ uint8_t s_lut[768];
    const int t = threadIdx.x;
    for (int i = t; i < 3 * Y * Y; i += blockDim.x) s_lut[i] = lut8[i];
    if (t == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int64_t Bq = Bpad >> 2, half = 2 * quarter;
    const int64_t hb = blockIdx.y, fq0 = (int64_t)blockIdx.x * 256;
    const int cnt = (int)((Bq - fq0) < 256 ? (Bq - fq0) : 256);
    const uint32_t bytes = (uint32_t)cnt * 4u;
    const uint8_t *b0 = sym + (hb * 32) * Bpad + fq0 * 4;
    auto issue = [&](int s) {  // rows 8 s .. 8 s + 7 of the four streams into slot s & 1
        const int slot = s & 1;
        mbar_expect_tx(&bars[slot], 32u * bytes);
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            const uint8_t *src = b0 + ((j & 1) ? half : 0) * Bpad + ((j & 2) ? quarter : 0) * Bpad + (int64_t)(8 * s) * Bpad;
#pragma unroll 1
            for (int r = 0; r < 8; ++r) bulk_g2s(sm_rows + ((slot * 4 + j) * 8 + r) * 256, src + (int64_t)r * Bpad, bytes, &bars[slot]);
        }
    };
    if (t == 0) {
        issue(0);
        issue(1);
    }
    const bool act = t < cnt;
    const int64_t fq = fq0 + t;
    uint4 wa = make_uint4(0, 0, 0, 0), wb = wa, wu = wa;
    if (act && top_g) {
        wa = *(const uint4 *)(x0 + hb * Bpad + 4 * fq);
        wb = *(const uint4 *)(x0 + (hb + (quarter >> 5)) * Bpad + 4 * fq);
    }
    if (act && isg) wu = *(const uint4 *)(xw + hb * Bpad + 4 * fq);
    const uint32_t wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w}, wuv[4] = {wu.x, wu.y, wu.z, wu.w};
    const uint32_t YY = (uint32_t)(Y * Y);
    uint32_t *o = (uint32_t *)out + hb * 32 * Bq + fq;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        const int slot = s & 1;
        mbar_wait(&bars[slot], (uint32_t)((s >> 1) & 1));
        if (act) {
#pragma unroll 4
            for (int r = 0; r < 8; ++r) {
                const int k = 8 * s + r;
                const uint32_t *row = sm_rows + (slot * 4 * 8 + r) * 256 + t;  // stream j at row[j * 8 * 256]
                const uint32_t y0 = row[0], y1 = row[8 * 256], y2 = row[16 * 256], y3 = row[24 * 256];  // streams: h, h + half, h + quarter, h + quarter + half
                uint32_t A = 0, Bv = 0, U = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t ma = top_g ? YY + YY * ((wav[j] >> k) & 1u) : 0u, mb = top_g ? YY + YY * ((wbv[j] >> k) & 1u) : 0u;
                    A |= (uint32_t)s_lut[ma + ((y0 >> (8 * j)) & 255u) * Y + ((y1 >> (8 * j)) & 255u)] << (8 * j);
                    Bv |= (uint32_t)s_lut[mb + ((y2 >> (8 * j)) & 255u) * Y + ((y3 >> (8 * j)) & 255u)] << (8 * j);
                    U |= ((wuv[j] >> k) & 1u) << (8 * j);
                }
                o[k * Bq] = isg ? g8(A, Bv, U) : f8(A, Bv);
            }
        }
        if (s + 2 < 4) {
            __syncthreads();  // every thread has read the slot
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(s + 2);
            }
        }
    }
}

// The same level update with the input rows staged by BULK ASYNCHRONOUS COPIES (cp.async.bulk completing on mbarriers, async_copy.cuh).
// With plain loads the bytes in flight are bounded by the register file (one register per 4-byte load: ~64 KB per SM at 50 %
// occupancy, long_scoreboard 26 warps per issue, 69 % of the copy peak in ncu); staged rows cost no registers.  A block owns 32
// consecutive elements x 256 frame quads (1 KB per row): one thread arms four barriers (8 rows of both operands each, 16 KB) and
// issues all 64 row copies at once; the block consumes stage after stage.  64 KB of shared memory per block, three blocks per SM.
template <bool PACK4>
__global__ void __launch_bounds__(256) hy_level8_bulk_kernel(int64_t size, int64_t Bpad, const uint8_t *__restrict__ in,
                                                             const uint32_t *__restrict__ xw, int isg, uint8_t *__restrict__ out) {
    extern __shared__ __align__(128) uint32_t sm_rows[];  // [2][32][256]: rows of operand a, then of operand b
    __shared__ uint64_t bars[4];
    const int64_t Bq = Bpad >> 2;
    const int64_t hb = blockIdx.y, fq0 = (int64_t)blockIdx.x * 256;
    const int cnt = (int)((Bq - fq0) < 256 ? (Bq - fq0) : 256);  // frame quads of this tile (a multiple of 8: Bpad is a multiple of 32)
    const uint32_t bytes = (uint32_t)cnt * 4u;
    const int t = threadIdx.x;
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < 4; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (t == 0) {
        const uint8_t *a = in + (hb * 32) * Bpad + fq0 * 4, *b = a + size * Bpad;
#pragma unroll 1
        for (int s = 0; s < 4; ++s) {
            mbar_expect_tx(&bars[s], 16u * bytes);
#pragma unroll 1
            for (int r = 8 * s; r < 8 * s + 8; ++r) {
                bulk_g2s(sm_rows + r * 256, a + (int64_t)r * Bpad, bytes, &bars[s]);
                bulk_g2s(sm_rows + (32 + r) * 256, b + (int64_t)r * Bpad, bytes, &bars[s]);
            }
        }
    }
    const bool act = t < cnt;
    const int64_t fq = fq0 + t;
    uint4 wu = make_uint4(0, 0, 0, 0);
    if (isg && act) wu = *(const uint4 *)(xw + hb * Bpad + 4 * fq);
    uint32_t *o = (uint32_t *)out + hb * 32 * Bq + fq;
#pragma unroll 1
    for (int s = 0; s < 4; ++s) {
        mbar_wait(&bars[s], 0);
        if (!act) continue;
#pragma unroll
        for (int k0 = 8 * s; k0 < 8 * s + 8; k0 += 4) {
            uint32_t r[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const int k = k0 + kk;
                const uint32_t av = sm_rows[k * 256 + t], bv = sm_rows[(32 + k) * 256 + t];
                const uint32_t u = ((wu.x >> k) & 1u) | (((wu.y >> k) & 1u) << 8) | (((wu.z >> k) & 1u) << 16) | (((wu.w >> k) & 1u) << 24);
                r[kk] = isg ? g8(av, bv, u) : f8(av, bv);
            }
            if (PACK4) {
                const uint32_t lo01 = __byte_perm(r[0], r[1], 0x5140), hi01 = __byte_perm(r[0], r[1], 0x7362);
                const uint32_t lo23 = __byte_perm(r[2], r[3], 0x5140), hi23 = __byte_perm(r[2], r[3], 0x7362);
                uint4 tt;
                tt.x = __byte_perm(lo01, lo23, 0x5410);
                tt.y = __byte_perm(lo01, lo23, 0x7632);
                tt.z = __byte_perm(hi01, hi23, 0x5410);
                tt.w = __byte_perm(hi01, hi23, 0x7632);
                *(uint4 *)((uint32_t *)out + (hb * 8 + (k0 >> 2)) * Bpad + 4 * fq) = tt;
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) o[(k0 + kk) * Bq] = r[kk];
            }
        }
    }
}

// out[h][f], h < quarter = N/4: node(L[h], L[h + quarter]) with L[e] = lut[mode(e)][sym[e]][sym[e + N/2]];
// top_g: level n-1 is in its g phase, its decision bits are x[0, N/2) = the words at x0; xw: this level's decision words
__global__ void __launch_bounds__(256) hy_level_sym_kernel(int64_t quarter, int64_t Bpad, const uint8_t *__restrict__ sym, int Y,
                                                           const double *__restrict__ lut, const uint32_t *__restrict__ x0,
                                                           int top_g, const uint32_t *__restrict__ xw, int isg,
                                                           double *__restrict__ out) {
    __shared__ double s_lut[768];
    for (int i = threadIdx.x; i < 3 * Y * Y; i += blockDim.x) s_lut[i] = lut[i];
    __syncthreads();
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= quarter * Bpad) return;
    const int64_t h = gid / Bpad, f = gid - h * Bpad, hb = h + quarter, half = 2 * quarter;
    const uint32_t y0 = sym[h * Bpad + f], y1 = sym[(h + half) * Bpad + f];
    const uint32_t y2 = sym[hb * Bpad + f], y3 = sym[(hb + half) * Bpad + f];
    uint32_t ma = 0, mb = 0;
    if (top_g) {
        ma = 1u + ((x0[(h >> 5) * Bpad + f] >> (h & 31)) & 1u);
        mb = 1u + ((x0[(hb >> 5) * Bpad + f] >> (hb & 31)) & 1u);
    }
    uint32_t u = 0;
    if (isg) u = (xw[(h >> 5) * Bpad + f] >> (h & 31)) & 1u;
    out[gid] = node_packed(s_lut[(ma * Y + y0) * Y + y1], s_lut[(mb * Y + y2) * Y + y3], isg != 0, u);
}

// out[h][f] = node(in[h][f], in[h + size][f]) for h < size; g: decision bit h of the words at xw (element h -> word h / 32).
// Two adjacent frames per thread (Bpad is a multiple of 32): 16-byte loads and stores, twice the bytes in flight per SM.
__global__ void __launch_bounds__(256) hy_level_kernel(int64_t size, int64_t Bpad, const double *__restrict__ in,
                                                       const uint32_t *__restrict__ xw, int isg, double *__restrict__ out) {
    const int64_t gid = 2 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (gid >= size * Bpad) return;
    const int64_t h = gid / Bpad, f = gid - h * Bpad;
    uint32_t u0 = 0, u1 = 0;
    if (isg) {
        const uint2 uw = *(const uint2 *)(xw + (h >> 5) * Bpad + f);
        u0 = (uw.x >> (h & 31)) & 1u;
        u1 = (uw.y >> (h & 31)) & 1u;
    }
    const double2 a = *(const double2 *)(in + gid), b = *(const double2 *)(in + gid + size * Bpad);
    double2 r;
    r.x = node_packed(a.x, b.x, isg != 0, u0);
    r.y = node_packed(a.y, b.y, isg != 0, u1);
    *(double2 *)(out + gid) = r;
}

// partial sums of a completed plus child: lo[w][f] ^= lo[w + words][f]
__global__ void __launch_bounds__(256) hy_xor_kernel(int64_t words, int64_t Bpad, uint32_t *__restrict__ lo) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= words * Bpad) return;
    lo[gid] ^= lo[gid + words * Bpad];
}

// information bits: word w of frame f gathers the bits of u (packed, [frames][Nw]) at the information positions
__global__ void __launch_bounds__(256) hy_info_kernel(int64_t frames, int k, int Nw, const int32_t *__restrict__ info_pos,
                                                      const uint32_t *__restrict__ u, uint32_t *__restrict__ info) {
    const int Kw = (k + 31) >> 5;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= frames * Kw) return;
    const int64_t f = wid / Kw;
    const int w = (int)(wid - f * Kw);
    const int j = 32 * w + lane;
    uint32_t b = 0;
    if (j < k) {
        const int pos = info_pos[j];
        b = (u[f * Nw + (pos >> 5)] >> (pos & 31)) & 1u;
    }
    const uint32_t wv = __ballot_sync(0xffffffffu, b);
    if (lane == 0) info[wid] = wv;
}

// The same gather, one thread per u WORD: the word's information bits are compressed to the low end (Hacker's Delight
// "compress" with the plan's bit-deposit masks, the inverse of the encoder's deposit) and OR-ed into the frame's information
// words at bit offset `before`.  A block covers 256 consecutive u words = one contiguous bit range of the output, assembled
// in shared memory; fully covered output words are stored, the two partial ones OR-ed (the output is zeroed first).
__global__ void __launch_bounds__(256) hy_info_words_kernel(int Nw, int Kw, const uint32_t *__restrict__ tab,
                                                            const uint32_t *__restrict__ u, uint32_t *__restrict__ info) {
    __shared__ uint32_t win[264];
    __shared__ uint32_t s_first, s_end;
    const int64_t f = blockIdx.x;
    const int wb = blockIdx.y * 256, w = wb + threadIdx.x;
    for (int i = threadIdx.x; i < 264; i += 256) win[i] = 0u;
    uint32_t m = 0, before = 0, x = 0;
    if (w < Nw) {
        const uint4 t0 = __ldg((const uint4 *)(tab + (size_t)w * 8)), t1 = __ldg((const uint4 *)(tab + (size_t)w * 8) + 1);
        m = t0.x, before = t0.y;
        x = u[f * Nw + w] & m;
        uint32_t t;
        t = x & t0.z, x = (x ^ t) | (t >> 1);
        t = x & t0.w, x = (x ^ t) | (t >> 2);
        t = x & t1.x, x = (x ^ t) | (t >> 4);
        t = x & t1.y, x = (x ^ t) | (t >> 8);
        t = x & t1.z, x = (x ^ t) | (t >> 16);
    }
    if (threadIdx.x == 0) s_first = before;
    const int last = (Nw - wb < 256 ? Nw - wb : 256) - 1;
    if (threadIdx.x == last) s_end = before + (uint32_t)__popc(m);
    __syncthreads();
    const uint32_t base = s_first >> 5;  // first output word this block touches
    if (m) {
        const uint32_t rel = before - 32u * base, sh = rel & 31u;
        atomicOr(&win[rel >> 5], x << sh);
        if (sh && (x >> (32u - sh))) atomicOr(&win[(rel >> 5) + 1], x >> (32u - sh));
    }
    __syncthreads();
    const uint32_t first = s_first, end = s_end;
    if (end == first) return;
    const uint32_t nwords = ((end + 31u) >> 5) - base;
    uint32_t *o = info + f * Kw + base;
    for (uint32_t i = threadIdx.x; i < nwords; i += 256) {
        const bool full = 32u * (base + i) >= first && 32u * (base + i + 1) <= end;
        if (full)
            o[i] = win[i];
        else if (win[i])
            atomicOr(o + i, win[i]);
    }
}

struct HyLayout {
    int64_t chunk, Bpad;
    size_t off_sym, off_cw, off_u, off_lut, off_lev[32], off_sub, sub_bytes, total;
};

static int64_t hy_frames_cap(bool et = false) {
    const char *s = getenv("PC_SC_HYBRID_FRAMES");
    // ~5.4 MB of workspace per 2^20-symbol frame; erasure-type channels (one byte per element): ~1.9 MB, and one
    // resident wave of the sub-block kernel (256 frames per SM) is the natural batch
    const int64_t v = s && *s ? atoll(s) : (et ? (int64_t)num_sms() * SC_THREADS : 16384);
    return v >= 32 ? round_up(v, 32) : 32;
}

// erasure-type channel table: every row is hard knowledge, an exact erasure, or (0,0)
static bool hy_erasure_type(const double *h_table, int Y) {
    if (const char *s = getenv("PC_SC_HY8"))
        if (atoi(s) == 0) return false;
    bool any_hard = false;
    for (int y = 0; y < Y; ++y) {
        const double p0 = h_table[2 * y], p1 = h_table[2 * y + 1];
        const bool hard = (p0 == 0.0) != (p1 == 0.0);
        any_hard = any_hard || hard;
        if (!(hard || p0 == p1)) return false;
    }
    return any_hard;
}

static HyLayout hy_layout(const pc_plan *plan, HybridTables *T, int64_t chunk, bool et = false) {
    HyLayout L{};
    const int64_t N = plan->N, Nw = N >> 5;
    L.chunk = chunk;
    L.Bpad = round_up(chunk, 32);
    size_t o = 0;
    L.off_sym = o;
    o += align256((size_t)N * L.Bpad);
    L.off_cw = o;
    o += align256((size_t)Nw * L.Bpad * 4);
    L.off_u = o;  // [chunk][Nw] reference-order codeword, then its transform, for the information gather
    o += align256((size_t)2 * Nw * L.Bpad * 4);
    L.off_lut = o;
    o += align256(768 * 8 + 768);  // float64 table, then the same table as state codes
    for (int l = HY_L0; l < plan->n - 1; ++l) {  // level n-1 is looked up from the symbols, never stored
        L.off_lev[l] = o;
        o += align256(((size_t)1 << l) * L.Bpad * (et ? 1 : 8));
    }
    L.off_sub = o;
    size_t sb = 256;
    if (T)
        for (size_t j = 0; j < T->sub.size(); ++j) {  // EVERY sub-plan: the layout depends on the sub-plan's k (host arithmetic only)
            const size_t b = sc_layout(T->sub[j], chunk, PC_INPUT_SYMBOLS).total;
            if (b > sb) sb = b;
        }
    L.sub_bytes = align256(sb);
    o += L.sub_bytes;
    L.total = o;
    return L;
}

static bool sc_use_hybrid(const pc_plan *plan, int64_t B, int kind) {
    if (plan->q != 2 || plan->n < HY_L0 + 2 || plan->n > 20 || kind != PC_INPUT_SYMBOLS) return false;
    const char *s = getenv("PC_SC_HYBRID");  // 1 forces it (any block above 2^10, for tests), 0 forbids it
    if (s && *s) return atoi(s) != 0;
    if (plan->n <= SC_MAX_N) return false;
    return B >= 6144;  // a batch takes ~N x 3.4 us whatever its size (the leaf blocks are frame per lane): below ~6 k
                       // frames the frame-per-CTA streamed decoder is faster
}

extern "C" int pc_polar_transform_bits(int n, const uint32_t *d_cw_packed, uint32_t *d_u_packed, int64_t B, void *stream);

// et: 1 erasure-type table (byte states), 0 float64 states, -1 table unknown (the larger of the two)
static size_t sc_hybrid_workspace_bytes(const pc_plan *plan, int64_t B, int et) {
    HybridTables *T = hybrid_tables(plan);
    size_t need = 0;
    for (int e = 0; e < 2; ++e) {
        if (et >= 0 && et != e) continue;
        int64_t chunk = round_up(B, 32);
        if (chunk > hy_frames_cap(e != 0)) chunk = hy_frames_cap(e != 0);
        const size_t b = hy_layout(plan, T, chunk, e != 0).total;
        if (b > need) need = b;
    }
    return need;
}

static int sc_hybrid_decode(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_table, int Y, uint32_t *d_cw,
                            uint32_t *d_info, void *ws, size_t ws_bytes, cudaStream_t st) {
    HybridTables *T = hybrid_tables(plan);
    if (!T) return PC_ERR_CUDA;
    // rate-1 shortcut: only channels with a hard output symbol (a table row with exactly one zero) can produce r = 0
    bool r1 = false;
    for (int y = 0; y < Y; ++y) r1 = r1 || ((h_table[2 * y] == 0.0) != (h_table[2 * y + 1] == 0.0));
    if (const char *s = getenv("PC_SC_R1")) r1 = r1 && atoi(s) != 0;
    const bool et = r1 && plan->n >= HY_L0 + 3 && hy_erasure_type(h_table, Y);
    const char *bk = getenv("PC_HY_BULK");  // 0: the plain-load level kernel (kept for comparison)
    const bool bulk8 = et && !(bk && atoi(bk) == 0);
    if (bulk8) {
        PC_CUDA(cudaFuncSetAttribute(hy_level8_bulk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        PC_CUDA(cudaFuncSetAttribute(hy_level8_bulk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        PC_CUDA(cudaFuncSetAttribute(hy_level_sym8_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    }
    int64_t chunk = round_up(B, 32);
    if (chunk > hy_frames_cap(et)) chunk = hy_frames_cap(et);
    while (chunk > 32 && hy_layout(plan, T, chunk, et).total > ws_bytes) chunk = round_up(chunk / 2, 32);
    const HyLayout L = hy_layout(plan, T, chunk, et);
    if (L.total > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed for a 32-frame chunk", ws_bytes, L.total);
        return PC_ERR_NOMEM;
    }
    const int n = plan->n, N = plan->N, Nw = N >> 5, Kw = (plan->k + 31) / 32, Ns = 1 << HY_L0, NS = N >> HY_L0;
    char *base = (char *)ws;
    uint8_t *sym = (uint8_t *)(base + L.off_sym);
    uint32_t *cw_t = (uint32_t *)(base + L.off_cw);
    uint32_t *cw_ref = (uint32_t *)(base + L.off_u), *u_ref = cw_ref + (size_t)Nw * L.Bpad;
    auto V = [&](int l) -> double * { return (double *)(base + L.off_lev[l]); };
    auto V8 = [&](int l) -> uint8_t * { return (uint8_t *)(base + L.off_lev[l]); };
    HyRootParams tp{};
    for (int i = 0; i < 32; ++i) tp.table[i] = i < 2 * Y ? h_table[i] : 0.0;
    const size_t smem = (size_t)SMEM_VALS * SC_THREADS * sizeof(double);
    PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<SC_INPUT_PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<SC_INPUT_PACKED, MODE_DECODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA(cudaFuncSetAttribute(sc_decode8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S8_WORDS * SC_THREADS * 4));
    ProfScope prof_scope(st);  // the whole walk is the measured unit
    double *lut = (double *)(base + L.off_lut);
    hy_lut_kernel<<<3, 256, 0, st>>>(Y, tp, lut);
    PC_LAUNCH_CHECK();
    const int64_t Bp = L.Bpad;
    auto blocks_of = [](int64_t items) { return (unsigned)((items + 255) / 256); };
    for (int64_t f0 = 0; f0 < B; f0 += chunk) {
        const int64_t frames = (B - f0) < chunk ? (B - f0) : chunk;
        const int64_t tiles = (frames + 31) / 32;
        if ((((uintptr_t)d_y + (size_t)f0 * N) & 3) == 0) {
            dim3 ig8((unsigned)((frames + 127) / 128), (unsigned)((N >> 7) < 64 ? (N >> 7) : 64));
            ingest_u8_kernel<<<ig8, 256, 0, st>>>(n, frames, Bp, d_y + f0 * N, sym, Y - 1);
        } else {
            dim3 ig((unsigned)tiles, 64u);
            ingest_kernel<uint8_t><<<ig, 256, 0, st>>>(n, frames, Bp, d_y + f0 * N, sym, (uint8_t)0, Y - 1);
        }
        PC_LAUNCH_CHECK();
        // level `lev` (>= HY_L0) of the current path: f, or g with the decision words of x[i - 2^lev, i)
        auto level_op = [&](int lev, bool isg, int64_t i) -> int {
            const int64_t size = (int64_t)1 << lev;
            const uint32_t *xw = isg ? cw_t + ((i - size) >> 5) * Bp : cw_t;
            if (lev == n - 1) return PC_OK;  // looked up on demand by the level below
            if (et && lev == n - 2 && bulk8 && (size >> 5) <= 65535)
                hy_level_sym8_bulk_kernel<<<dim3((unsigned)((Bp / 4 + 255) / 256), (unsigned)(size >> 5)), 256, 65536, st>>>(
                    size, Bp, sym, Y, (const uint8_t *)(lut + 768), cw_t, i >= N / 2 ? 1 : 0, xw, isg ? 1 : 0, V8(lev));
            else if (et && lev == n - 2)
                hy_level_sym8_kernel<<<blocks_of((size >> 5) * (Bp / 4)), 256, 0, st>>>(size, Bp, sym, Y, (const uint8_t *)(lut + 768), cw_t,
                                                                              i >= N / 2 ? 1 : 0, xw, isg ? 1 : 0, V8(lev));
            else if (et && bulk8 && (size >> 5) <= 65535) {
                const dim3 g((unsigned)((Bp / 4 + 255) / 256), (unsigned)(size >> 5));
                if (lev == HY_L0)
                    hy_level8_bulk_kernel<true><<<g, 256, 65536, st>>>(size, Bp, V8(lev + 1), xw, isg ? 1 : 0, V8(lev));
                else
                    hy_level8_bulk_kernel<false><<<g, 256, 65536, st>>>(size, Bp, V8(lev + 1), xw, isg ? 1 : 0, V8(lev));
            } else if (et && lev == HY_L0)
                hy_level8_kernel<true><<<blocks_of((size >> 5) * (Bp / 4)), 256, 0, st>>>(size, Bp, V8(lev + 1), xw, isg ? 1 : 0, V8(lev));
            else if (et)
                hy_level8_kernel<false><<<blocks_of((size >> 5) * (Bp / 4)), 256, 0, st>>>(size, Bp, V8(lev + 1), xw, isg ? 1 : 0, V8(lev));
            else if (lev == n - 2)
                hy_level_sym_kernel<<<blocks_of(size * Bp), 256, 0, st>>>(size, Bp, sym, Y, lut, cw_t, i >= N / 2 ? 1 : 0, xw,
                                                                         isg ? 1 : 0, V(lev));
            else
                hy_level_kernel<<<blocks_of(size * Bp / 2), 256, 0, st>>>(size, Bp, V(lev + 1), xw, isg ? 1 : 0, V(lev));
            PC_LAUNCH_CHECK();
            return PC_OK;
        };
        for (int j = 0; j < NS; ++j) {
            const int64_t i = (int64_t)j * Ns;
            const int stop = T->all_frozen[j] ? HY_L0 + 1 : HY_L0;  // a rate-0 sub-block never needs its own level
            int lev;
            bool isg = false;
            if (j == 0) {
                lev = n - 1;
            } else {
                lev = HY_L0 + __builtin_ctz((unsigned)j);  // g at this level, then f down to the sub-block
                isg = true;
            }
            for (; lev >= stop; --lev, isg = false) {
                const int rc = level_op(lev, isg, i);
                if (rc) return rc;
            }
            // the sub-block: frame per lane, top level read from V(HY_L0), partial sums written into the frame's words
            const pc_plan *sp = T->sub[j];
            const ScLayout SL = sc_layout(sp, chunk, PC_INPUT_SYMBOLS);
            ScParams p{};
            p.n = HY_L0;
            p.k = sp->k;
            p.n_sched = (int)sp->sched.size();
            p.Y = Y;
            p.frames = frames;
            p.Bpad = Bp;
            p.sched = sp->d_sched;
            p.r0_words = sp->d_r0_words;
            p.in_t = V(HY_L0);
            p.vals = (double *)(base + L.off_sub + SL.off_vals);
            p.cw_t = cw_t + (i >> 5) * Bp;
            p.info_t = (uint32_t *)(base + L.off_sub + SL.off_info);
            const int64_t blocks = (tiles * 32 + SC_THREADS - 1) / SC_THREADS;
            const int grid = (int)(blocks < SL.grid ? blocks : SL.grid);
            if (et) {
                p.sched = T->d_sched_r1 + T->r1_off[j];
                p.n_sched = T->r1_len[j];
                p.in_t = V8(HY_L0);
                const int g8 = (int)(blocks < num_sms() ? blocks : num_sms());
                sc_decode8_kernel<<<g8, SC_THREADS, (size_t)S8_WORDS * SC_THREADS * 4, st>>>(p);
            } else if (r1) {
                p.sched = T->d_sched_r1 + T->r1_off[j];
                p.n_sched = T->r1_len[j];
                sc_decode_kernel<SC_INPUT_PACKED, MODE_DECODE, true><<<grid, SC_THREADS, smem, st>>>(p);
            } else {
                const char *bl = getenv("PC_SC_BLOCK");
                if (sp->d_sched_b && !(bl && atoi(bl) == 0)) {  // leaf blocks, as in sc_decode_common
                    p.sched = sp->d_sched_b;
                    p.n_sched = (int)sp->sched_b.size();
                }
                sc_decode_kernel<SC_INPUT_PACKED><<<grid, SC_THREADS, smem, st>>>(p);
            }
            PC_LAUNCH_CHECK();
            // partial sums above the sub-block: x[ii - s, ii) ^= x[ii, ii + s) whenever a plus child completes
            int lv = HY_L0;
            int64_t ii = i;
            while (lv < n && ((ii >> lv) & 1)) {
                const int64_t s = (int64_t)1 << lv;
                hy_xor_kernel<<<blocks_of((s >> 5) * Bp), 256, 0, st>>>(s >> 5, Bp, cw_t + ((ii - s) >> 5) * Bp);
                PC_LAUNCH_CHECK();
                ii -= s;
                ++lv;
            }
        }
        // outputs: codeword to the reference order; information = gather of its transform at the information positions
        {
            const int G = 1 << (n - 10);
            if (n >= 16 && !getenv("PC_HY_EGRESS_GATHER")) {
                // word transpose to frame rows (coalesced on both sides), then the bit reversal with the block encoder's
                // shared-memory 32 x 32 bit transposes: the one-pass egress below scatters 4-byte stores G words apart
                const int jt = (Nw + 255) / 256;
                egress_kernel<false><<<dim3((unsigned)tiles, (unsigned)(jt < 64 ? jt : 64)), 256, 0, st>>>(n, Nw, frames, Bp, cw_t, cw_ref);
                PC_LAUNCH_CHECK();
                const int rc = bitrev_words_launch(n, frames, cw_ref, d_cw + f0 * Nw, st);
                if (rc) return rc;
            } else {
                egress_bitrev_kernel<<<dim3((unsigned)tiles, (unsigned)((G + 7) / 8 < 64 ? (G + 7) / 8 : 64)), 256, 0, st>>>(
                    n, frames, Bp, cw_t, d_cw + f0 * Nw);
                PC_LAUNCH_CHECK();
            }
            if (Kw > 0) {
                const int rc = pc_polar_transform_bits(n, d_cw + f0 * Nw, u_ref, frames, st);
                if (rc) return rc;
                if (getenv("PC_HY_INFO_GATHER")) {  // the per-bit gather (kept for comparison)
                    hy_info_kernel<<<blocks_of(frames * Kw * 32), 256, 0, st>>>(frames, plan->k, Nw, T->d_info_pos, u_ref, d_info + f0 * Kw);
                } else {
                    PC_CUDA(cudaMemsetAsync(d_info + f0 * Kw, 0, (size_t)frames * Kw * 4, st));
                    hy_info_words_kernel<<<dim3((unsigned)frames, (unsigned)((Nw + 255) / 256)), 256, 0, st>>>(Nw, Kw, plan->d_enc_tab, u_ref,
                                                                                                        d_info + f0 * Kw);
                }
                PC_LAUNCH_CHECK();
            }
        }
    }
    return PC_OK;
}

// ---- genie pass: all indices frozen to per-frame known bits, leaf probabilities captured -------------------------------
// (BinaryPolarEncoderDecoder.genieSingleDecodeSimulatioan, BinaryPolarEncoderDecoder.py:114-178, capture :268-273)
struct GenieTables {
    SchedEntry *d_sched = nullptr;
    int n_sched = 0;
};
static std::mutex g_genie_mu;
static std::map<const pc_plan *, GenieTables *> g_genie_tables;

static GenieTables *genie_tables(const pc_plan *plan) {
    std::lock_guard<std::mutex> lk(g_genie_mu);
    auto it = g_genie_tables.find(plan);
    if (it != g_genie_tables.end()) return it->second;
    std::vector<SchedEntry> sched((size_t)plan->N);
    for (int i = 0; i < plan->N; ++i) {
        SchedEntry e{};
        e.i = i;
        e.l = 0;
        e.kind = NODE_GENIE;
        e.top = (int8_t)(i == 0 ? plan->n : __builtin_ctz((unsigned)i));
        sched[i] = e;
    }
    GenieTables *T = new GenieTables();
    T->n_sched = plan->N;
    if (cudaMalloc((void **)&T->d_sched, sizeof(SchedEntry) * sched.size()) != cudaSuccess ||
        cudaMemcpy(T->d_sched, sched.data(), sizeof(SchedEntry) * sched.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("genie schedule: device upload failed");
        delete T;
        return nullptr;
    }
    g_genie_tables[plan] = T;
    return T;
}

void genie_tables_release(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_genie_mu);
    auto it = g_genie_tables.find(p);
    if (it == g_genie_tables.end()) return;
    cudaFree(it->second->d_sched);
    delete it->second;
    g_genie_tables.erase(it);
}

// u_t[w][f] = bits [bit_off + 32 w, bit_off + 32 w + 32) of row f of `u` (row pitch in words), masked to N bits
__global__ void __launch_bounds__(256) genie_u_ingest_kernel(int N, int64_t frames, int64_t Bpad, const uint32_t *__restrict__ u,
                                                             int64_t pitch_words, int bit_off, uint32_t *__restrict__ u_t) {
    const int Nw = N >= 32 ? N >> 5 : 1;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= Bpad * Nw) return;
    const int w = (int)(gid / Bpad);
    const int64_t f = gid - (int64_t)w * Bpad;
    uint32_t v = 0;
    if (f < frames) {
        const int64_t b0 = (int64_t)bit_off + 32 * (int64_t)w;
        const uint32_t *row = u + f * pitch_words;
        const int sh = (int)(b0 & 31);
        v = row[b0 >> 5] >> sh;
        if (sh && N - 32 * w > 32 - sh) v |= row[(b0 >> 5) + 1] << (32 - sh);
        if (N < 32) v &= (1u << N) - 1u;
    }
    u_t[gid] = v;
}

// marg[f][off + 2 i .. + 1] = calcMarginalizedProbabilities of leaf i (BinaryMemorylessVectorDistribution.py:52-69) from the
// packed level-0 value: pair (1, r) or (r, 1) by the sign bit, NaN = (0, 0) -> [0.5, 0.5]
__global__ void __launch_bounds__(256) genie_marg_egress_kernel(int N, int64_t frames, int64_t Bpad, const double *__restrict__ marg_t,
                                                                double *__restrict__ out, int64_t pitch, int64_t off) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t f0 = (int64_t)blockIdx.x * 32;
    for (int it = blockIdx.y; it < (N + 31) / 32; it += gridDim.y) {
        const int i0 = it * 32;
        for (int r = ty; r < 32; r += 8) tile[r][tx] = (i0 + r < N) ? marg_t[(int64_t)(i0 + r) * Bpad + f0 + tx] : 0.0;
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int64_t f = f0 + r;
            const int i = i0 + tx;
            if (f < frames && i < N) {
                const double v = tile[tx][r];
                double m0 = 0.5, m1 = 0.5;
                if (v == v) {
                    const double rr = d_abs(v);
                    const double p0 = d_sign(v) ? rr : 1.0, p1 = d_sign(v) ? 1.0 : rr;
                    const double s = __dadd_rn(__dadd_rn(0.0, p0), p1);
                    m0 = p0 / s;
                    m1 = p1 / s;
                }
                double2 *o = (double2 *)(out + f * pitch + off) + i;
                *o = make_double2(m0, m1);
            }
        }
        __syncthreads();
    }
}

struct GenieLayout {
    ScLayout L;
    size_t off_u, off_marg, total;
};
static GenieLayout genie_layout(const pc_plan *plan, int64_t chunk) {
    GenieLayout G;
    G.L = sc_layout(plan, chunk, PC_INPUT_PROBS);
    const int64_t N = plan->N, Nw = N >= 32 ? N >> 5 : 1;
    size_t o = G.L.total;
    G.off_u = o;
    o += align256((size_t)Nw * G.L.Bpad * 4);
    G.off_marg = o;
    o += align256((size_t)N * G.L.Bpad * 8);
    G.total = o;
    return G;
}

size_t sc_genie_workspace_bytes(const pc_plan *plan, int64_t B) {
    return genie_layout(plan, sc_pick_chunk(B, PC_INPUT_PROBS)).total;
}

// Walk of the UNPRUNED tree over B rows of probability pairs (see the MODE_* comment above the kernel).
// d_u: known u bits (genie, prior-encode), row f at d_u + f * u_pitch_words, first bit u_bit_off; d_marg (optional): row f
// at d_marg + f * marg_pitch + marg_off; d_rnd: randomlyGeneratedNumbers (dual / prior modes), row pitch rnd_stride doubles.
static int sc_walk_common(int mode, const pc_plan *plan, const double *d_in, const uint32_t *d_u, int64_t u_pitch_words,
                          int u_bit_off, const double *d_rnd, int64_t rnd_stride, int64_t B, uint32_t *d_cw, uint32_t *d_info, double *d_marg,
                          int64_t marg_pitch, int64_t marg_off, void *ws, size_t ws_bytes, cudaStream_t st) {
    PC_REQUIRE(plan && plan->q == 2, "binary plan required");
    PC_REQUIRE(plan->n >= 1 && plan->n <= SC_MAX_N, "this pass needs 2 <= N <= 65536");
    PC_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_in && d_cw && ws, "null buffer");
    PC_REQUIRE(mode == MODE_DUAL || d_u, "known u bits missing");
    PC_REQUIRE(mode == MODE_GENIE || (d_rnd && (rnd_stride == 0 || rnd_stride >= plan->N)), "common randomness missing");
    PC_REQUIRE(mode != MODE_DUAL || (B % 2 == 0 && (d_info || plan->k == 0)), "dual pass: rows come in (xy, x) pairs");
    PC_REQUIRE(mode != MODE_GENIE || d_marg, "genie pass needs the marginal output");
    PC_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
    PC_REQUIRE(((uintptr_t)d_marg & 15) == 0 && (marg_pitch & 1) == 0 && (marg_off & 1) == 0, "marginal output must be 16-byte aligned");
    GenieTables *T = genie_tables(plan);
    if (!T) return PC_ERR_CUDA;
    int64_t chunk = sc_pick_chunk(B, PC_INPUT_PROBS);
    while (chunk > 32 && genie_layout(plan, chunk).total > ws_bytes) chunk = round_up(chunk / 2, 32);
    GenieLayout G = genie_layout(plan, chunk);
    if (G.total > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed for a 32-frame chunk", ws_bytes, G.total);
        return PC_ERR_NOMEM;
    }
    const ScLayout &L = G.L;
    const int N = plan->N, Nw = (N + 31) / 32, Kw = (plan->k + 31) / 32;
    char *base = (char *)ws;
    ScParams p{};
    p.n = plan->n;
    p.k = plan->k;
    p.n_sched = T->n_sched;
    p.Bpad = L.Bpad;
    p.sched = T->d_sched;
    p.r0_words = plan->d_r0_words;
    p.in_t = base + L.off_in;
    p.vals = (double *)(base + L.off_vals);
    p.cw_t = (uint32_t *)(base + L.off_cw);
    p.info_t = (uint32_t *)(base + L.off_info);
    p.u_t = (const uint32_t *)(base + G.off_u);
    p.marg_t = d_marg ? (double *)(base + G.off_marg) : nullptr;
    p.fmask = plan->d_frozen_mask;
    p.rnd_stride = rnd_stride;
    const size_t smem = (size_t)SMEM_VALS * SC_THREADS * sizeof(double);
    PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<PC_INPUT_PROBS, MODE_GENIE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<PC_INPUT_PROBS, MODE_DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA(cudaFuncSetAttribute(sc_decode_kernel<PC_INPUT_PROBS, MODE_PRIOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int64_t f0 = 0; f0 < B; f0 += chunk) {
        const int64_t frames = (B - f0) < chunk ? (B - f0) : chunk;
        const int64_t tiles = (frames + 31) / 32;
        p.frames = frames;
        p.rnd = d_rnd ? d_rnd + f0 * rnd_stride : nullptr;
        const int ptiles = (N + 31) / 32;
        dim3 ig((unsigned)tiles, (unsigned)(ptiles < 64 ? ptiles : 64));
        ingest_kernel<double2><<<ig, 256, 0, st>>>(plan->n, frames, L.Bpad, (const double2 *)d_in + f0 * N, (double2 *)p.in_t,
                                                   make_double2(0.5, 0.5));
        PC_LAUNCH_CHECK();
        if (d_u) {
            const int64_t uitems = L.Bpad * Nw;
            genie_u_ingest_kernel<<<(unsigned)((uitems + 255) / 256), 256, 0, st>>>(N, frames, L.Bpad, d_u + f0 * u_pitch_words,
                                                                                     u_pitch_words, u_bit_off, (uint32_t *)p.u_t);
            PC_LAUNCH_CHECK();
        }
        const int64_t blocks = (tiles * 32 + SC_THREADS - 1) / SC_THREADS;
        const int grid = (int)(blocks < L.grid ? blocks : L.grid);
        if (mode == MODE_GENIE)
            sc_decode_kernel<PC_INPUT_PROBS, MODE_GENIE><<<grid, SC_THREADS, smem, st>>>(p);
        else if (mode == MODE_DUAL)
            sc_decode_kernel<PC_INPUT_PROBS, MODE_DUAL><<<grid, SC_THREADS, smem, st>>>(p);
        else
            sc_decode_kernel<PC_INPUT_PROBS, MODE_PRIOR><<<grid, SC_THREADS, smem, st>>>(p);
        PC_LAUNCH_CHECK();
        if (plan->n >= 10) {
            const int Gq = 1 << (plan->n - 10);
            egress_bitrev_kernel<<<dim3((unsigned)tiles, (unsigned)((Gq + 7) / 8 < 8 ? (Gq + 7) / 8 : 8)), 256, 0, st>>>(
                plan->n, frames, L.Bpad, p.cw_t, d_cw + f0 * Nw);
        } else {
            egress_kernel<true><<<dim3((unsigned)tiles, 1u), 256, 0, st>>>(plan->n, Nw, frames, L.Bpad, p.cw_t, d_cw + f0 * Nw);
        }
        PC_LAUNCH_CHECK();
        if (mode == MODE_DUAL && Kw > 0) {
            egress_kernel<false><<<dim3((unsigned)tiles, (unsigned)((Kw + 255) / 256)), 256, 0, st>>>(plan->n, Kw, frames, L.Bpad,
                                                                                                      p.info_t, d_info + f0 * Kw);
            PC_LAUNCH_CHECK();
        }
        if (d_marg) {
            genie_marg_egress_kernel<<<dim3((unsigned)tiles, (unsigned)(ptiles < 64 ? ptiles : 64)), 256, 0, st>>>(
                N, frames, L.Bpad, p.marg_t, d_marg + f0 * marg_pitch, marg_pitch, marg_off);
            PC_LAUNCH_CHECK();
        }
    }
    return PC_OK;
}

int sc_genie_common(const pc_plan *plan, const double *d_xy, const uint32_t *d_u, int64_t u_pitch_words, int u_bit_off, int64_t B,
                    uint32_t *d_cw, double *d_marg, int64_t marg_pitch, int64_t marg_off, void *ws, size_t ws_bytes,
                    cudaStream_t st) {
    return sc_walk_common(MODE_GENIE, plan, d_xy, d_u, u_pitch_words, u_bit_off, nullptr, 0, B, d_cw, nullptr, d_marg, marg_pitch,
                          marg_off, ws, ws_bytes, st);
}

int sc_dual_common(const pc_plan *plan, const double *d_pairs, const double *d_rnd, int64_t rnd_stride, int64_t rows, uint32_t *d_cw,
                   uint32_t *d_info, double *d_marg, void *ws, size_t ws_bytes, cudaStream_t st) {
    return sc_walk_common(MODE_DUAL, plan, d_pairs, nullptr, 0, 0, d_rnd, rnd_stride, rows, d_cw, d_info, d_marg, (int64_t)2 * plan->N, 0, ws,
                          ws_bytes, st);
}

int sc_prior_common(const pc_plan *plan, const double *d_x, const uint32_t *d_u, const double *d_rnd, int64_t rnd_stride, int64_t B,
                    uint32_t *d_cw, double *d_marg, void *ws, size_t ws_bytes, cudaStream_t st) {
    return sc_walk_common(MODE_PRIOR, plan, d_x, d_u, (plan->N + 31) / 32, 0, d_rnd, rnd_stride, B, d_cw, nullptr, d_marg, (int64_t)2 * plan->N, 0,
                          ws, ws_bytes, st);
}

}  // namespace pc

extern "C" {

int64_t pc_sc_wave_frames(const pc_plan *plan) {
    if (!plan) return 0;
    if (plan->n > pc::SC_MAX_N) return pc::sc_stream_wave_frames(plan);
    return (int64_t)pc::sc_grid_max() * pc::SC_THREADS;
}

size_t pc_sc_workspace_bytes_symbols(const pc_plan *plan, int64_t B, const double *h_table, int Y) {
    if (!plan || B <= 0) return 256;
    if (h_table && Y >= 1 && Y <= 16 && pc::sc_use_hybrid(plan, B, PC_INPUT_SYMBOLS))
    {
        const char *s = getenv("PC_SC_R1");
        const bool et = !(s && atoi(s) == 0) && plan->n >= pc::HY_L0 + 3 && pc::hy_erasure_type(h_table, Y);
        return pc::sc_hybrid_workspace_bytes(plan, B, et ? 1 : 0);
    }
    return pc_sc_workspace_bytes(plan, B, PC_INPUT_SYMBOLS);
}

size_t pc_sc_workspace_bytes(const pc_plan *plan, int64_t B, int input_kind) {
    if (!plan || B <= 0) return 256;
    if (pc::sc_use_hybrid(plan, B, input_kind)) return pc::sc_hybrid_workspace_bytes(plan, B);
    if (pc::sc_use_stream(plan, B)) return pc::sc_stream_workspace_bytes(plan, B);
    return pc::sc_layout(plan, pc::sc_pick_chunk(B, input_kind), input_kind).total;
}

size_t pc_sc_genie_workspace_bytes(const pc_plan *plan, int64_t B) {
    if (!plan || B <= 0) return 256;
    return pc::sc_genie_workspace_bytes(plan, B);
}

int pc_sc_genie_probs(const pc_plan *plan, const double *d_xy, const uint32_t *d_u_packed, int64_t B, uint32_t *d_cw_packed,
                      double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!plan) {
        pc::set_error("plan is null");
        return PC_ERR_INVALID;
    }
    const int Nw = (plan->N + 31) / 32;
    return pc::sc_genie_common(plan, d_xy, d_u_packed, Nw, 0, B, d_cw_packed, d_marg, (int64_t)2 * plan->N, 0, d_workspace,
                               workspace_bytes, (cudaStream_t)stream);
}

int pc_sc_decode_probs_prior(const pc_plan *plan, const double *d_pairs, const double *d_rnd, int64_t rnd_row_stride, int64_t rows,
                             uint32_t *d_cw_packed, uint32_t *d_info_packed, double *d_marg, void *d_workspace,
                             size_t workspace_bytes, void *stream) {
    if (!plan) {
        pc::set_error("plan is null");
        return PC_ERR_INVALID;
    }
    return pc::sc_dual_common(plan, d_pairs, d_rnd, rnd_row_stride, rows, d_cw_packed, d_info_packed, d_marg, d_workspace, workspace_bytes,
                              (cudaStream_t)stream);
}

int pc_sc_encode_prior(const pc_plan *plan, const double *d_x, const uint32_t *d_u_packed, const double *d_rnd,
                       int64_t rnd_row_stride, int64_t B,
                       uint32_t *d_cw_packed, double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!plan) {
        pc::set_error("plan is null");
        return PC_ERR_INVALID;
    }
    return pc::sc_prior_common(plan, d_x, d_u_packed, d_rnd, rnd_row_stride, B, d_cw_packed, d_marg, d_workspace, workspace_bytes,
                               (cudaStream_t)stream);
}

int pc_sc_decode_probs(const pc_plan *plan, const double *d_xy, int64_t B, uint32_t *d_cw_packed, uint32_t *d_info_packed,
                       void *d_workspace, size_t workspace_bytes, void *stream) {
    return pc::sc_decode_common(plan, PC_INPUT_PROBS, d_xy, B, nullptr, 0, d_cw_packed, d_info_packed, d_workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

int pc_sc_decode_symbols(const pc_plan *plan, const uint8_t *d_y, int64_t B, const double *h_table, int Y,
                         uint32_t *d_cw_packed, uint32_t *d_info_packed, void *d_workspace, size_t workspace_bytes,
                         void *stream) {
    return pc::sc_decode_common(plan, PC_INPUT_SYMBOLS, d_y, B, h_table, Y, d_cw_packed, d_info_packed, d_workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
