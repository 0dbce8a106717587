// scl_path.cu -- binary (q = 2) SC-list decoding with ONE PATH PER LANE, float64 linear domain, arithmetic and decisions
// identical to QaryPolarEncoderDecoder.listDecode with q = 2 (QaryPolarEncoderDecoder.py:118-227, recursiveListDecode
// :403-757, helpers :759-820, normalize :867-872, QaryMemorylessVectorDistribution.py:26-118).
//
// Mapping.  A warp decodes 32 / G frames at once, G = the list size rounded up to a power of two (L = 8: four frames per
// warp); lane g * G + t owns path slot t of frame g.  Every frame of a batch walks the same op list (the recursion depends
// on the frozen set only), so the warp never diverges:
//  * f / g passes: a lane walks the elements of ITS path one after the other; every node vector is stored
//    [element][lane], so a warp access is one 512-byte row whatever the lazy-copy permutation is (a permuted read stays
//    inside the frame's 128-byte quarter of the row); two to six independent node updates are in flight per lane;
//  * the order-dependent float64 products of the fast nodes (np.product is a left-to-right product), the reliabilities, the
//    least-reliable top-k and the candidate metrics are per-path work: one lane each, all 32 lanes busy;
//  * list bookkeeping (metric, parent maps, counts) lives in registers and a few bytes of shared memory per lane; op decode,
//    loop control and the prune are paid once per FOUR frames instead of once per frame (frame-per-warp, scl_warp.cu);
//  * the prune is L rounds of an arg-max inside the frame's lane group (shuffles of width G), giving the oracle's ascending
//    (metric, index) order; lists that need no pruning go through the same rounds keyed by the candidate index;
//  * before the first forking node the list holds one path: those ops run in a cooperative mode (the G lanes of a frame
//    share the elements of the single path, vectors in a packed "shared" layout that later ops read as a broadcast);
//  * path vectors of levels <= lsm are in shared memory, larger levels in a per-warp global scratch;
//  * the channel level is read in the caller's layout: float64 probability pairs (xyVectorDistribution.probs) or uint8 output
//    symbols looked up in the channel table (makeQaryMemorylessVectorDistribution, QaryMemorylessDistribution.py:757-776).
// Three kernels per batch: sclp_prep_kernel (codewords of the actual word and of the frozen values, bit-packed),
// sclp_kernel (the list decoder), sclp_final_kernel (selection, ProbResult, information bits; the genie path that listDecode
// tracks is replayed there, and only for frames whose actual word is not in the final list or when the caller asks for the
// list outputs -- ProbResult needs actual_prob in no other case).
#include <mutex>

#include "async_copy.cuh"
#include "scl_arith.cuh"

namespace pc {

struct SclpParams {
    int n, k, L, G, gsh, n_ops, lsm, rgl, NW, n_leaf, sym, want_list, nfw;
    int nst;              // stages of the bulk-copy ring (0: plain loads)
    int l2lvl;            // L2 policies: global levels <= l2lvl are kept (evict_last), larger ones stream (evict_first); 0 = no hints
    int alias;            // diagnostic (PC_SCLP_ALIAS): every CTA uses the scratch of CTA 0 -- wrong results, shows the DRAM-free speed
    int wpc, sync, smem_per_warp;  // warps per CTA, per-op barrier among the warps of a scheduler, shared-memory bytes per warp
    int64_t frames;
    const uint4 *ops;
    const uint32_t *coef_words;
    const double2 *xy;    // [frames][N] caller layout (reference order)
    const uint8_t *y;     // [frames][N] channel output symbols
    const double2 *tab;   // [256] channel table rows
    const uint32_t *fvp;  // [frames][nfw] packed frozen values (u order) or null = all zero
    const uint32_t *Fb;   // [frames][NW] node-local codewords of the frozen values (reference order) or null
    const uint32_t *Acw;  // [frames][NW] codeword of the actual word
    double2 *vg;          // [grid][vg_stride] path vectors of levels > lsm
    int64_t vg_stride;
    uint32_t *rg;         // [grid][rg_stride] path codewords of levels >= rgl
    int64_t rg_stride;
    uint32_t *selcw;      // [frames][NW] root codeword of the selected path ([frames][L][NW]: every path, with want_list)
    int32_t *found;       // [frames] index of the actual word in the final list or -1
    int32_t *lsize;       // [frames]
    double *lprob;        // [frames][L] normalised metrics in list order
    double *mxs;          // [frames][n_leaf] list maximum at every fast node (the genie replay divides by them)
    unsigned long long *timing;  // optional [8][16] cycles per (op kind, level), summed over warps (tuning runs: PC_SCLP_TIMING)
};

static size_t sclp_smem_bytes(int n, int lsm, int rgl, int nst) {
    size_t b = (size_t)((2 << lsm) - 1) * 32 * 16;  // path vectors
    b += (size_t)2 * scl2_wsum(rgl) * 32 * 4;       // path codewords of levels < rgl
    b += (size_t)(n + 1) * 2 * 32;                  // parent maps
    b = (b + 15) & ~(size_t)15;
    b += (size_t)nst * (8 * 32 * 16 + 8);           // bulk-copy ring: eight rows per stage, one mbarrier per stage
    return (b + 15) & ~(size_t)15;
}

// ---- the f / g pass over the lane's own path ------------------------------------------------------------------------------------
// ONE loop serves every source kind through warp-uniform switches, so that the kernel's hot code stays resident in the
// instruction cache while 12-24 warps per SM sit at different ops of the list.  A step takes EIGHT consecutive source elements:
// four independent level l-1 updates and (fused) two level l-2 updates -- the float64 division of the normalisation is a chain of
// ~14 dependent operations, and interleaved chains are what keeps the FP64 pipe fed with 3-4 warps per scheduler.  The raw
// elements of step k+1 are fetched before step k is computed.
enum : int { SK_PATH = 0, SK_STAGED = 1, SK_SYM = 2, SK_DUAL = 3, SK_SHARED = 4 };
struct Src {
    int kind;
    // SK_PATH: float64 pairs at element stride ss: a path vector ([element][lane], stride 32) or the channel probabilities (1)
    // SK_STAGED: an HBM level through the bulk-copy ring: a stage holds the eight 512-byte rows of a step (4 KB contiguous in
    //   HBM), requested by lane 0 `nst` steps ahead; every lane reads its path's column `col` of the staged rows
    // SK_DUAL / SK_SHARED: pairs in the shared layout from p = level base + the frame's first lane; SK_DUAL picks variant
    //   v = bit e of `bits` (the codeword that selected the variants, this path's column): index 2 e + v
    // SK_SYM: channel symbols through the channel table
    const double2 *p;
    int ss, gsh;
    const uint32_t *bits;
    const uint8_t *y;
    const double2 *tab;
    double2 *stg;
    uint64_t *bars;
    uint32_t *ph;
    int nst, nsteps, col, lane;
    uint64_t pol;  // L2 policy of the bulk copies of an SK_STAGED source (0: none)
};
struct Raw8 {
    double2 v[8];
};
__device__ __forceinline__ void src_issue(const Src &s, int k) {
    const int st = k & (s.nst - 1);
    mbar_expect_tx(s.bars + st, 4096u);
    if (s.pol)
        bulk_g2s_hint(s.stg + st * 256, s.p + (int64_t)k * 256, 4096u, s.bars + st, s.pol);
    else
        bulk_g2s(s.stg + st * 256, s.p + (int64_t)k * 256, 4096u, s.bars + st);
}
__device__ __forceinline__ double2 src_at(const Src &s, int i) { return s.p[((i >> s.gsh) << 5) + (i & ((1 << s.gsh) - 1))]; }
// raw elements [8k, 8k + CNT) of the source (CNT = 8, or 4 for the two-element levels)
template <int CNT>
__device__ __forceinline__ void src_fetch(const Src &s, int k, Raw8 &r) {
    switch (s.kind) {
        case SK_PATH: {
            const double2 *q = s.p + (int64_t)(8 * k) * s.ss;
#pragma unroll
            for (int i = 0; i < CNT; ++i) r.v[i] = q[i * s.ss];
            break;
        }
        case SK_STAGED: {
            const int st = k & (s.nst - 1);
            mbar_wait(s.bars + st, (*s.ph >> st) & 1u);
            *s.ph ^= 1u << st;
            const double2 *q = s.stg + st * 256 + s.col;
#pragma unroll
            for (int i = 0; i < CNT; ++i) r.v[i] = q[i << 5];
            __syncwarp();
            if (s.lane == 0 && k + s.nst < s.nsteps) src_issue(s, k + s.nst);
            break;
        }
        case SK_SYM: {
            const uint8_t *q = s.y + 8 * k;
#pragma unroll
            for (int i = 0; i < CNT; ++i) r.v[i] = s.tab[q[i]];
            break;
        }
        case SK_DUAL: {
            const uint32_t w = s.bits[((8 * k) >> 5) << 5] >> ((8 * k) & 31);
#pragma unroll
            for (int i = 0; i < CNT; ++i) r.v[i] = src_at(s, 16 * k + 2 * i + (int)((w >> i) & 1u));
            break;
        }
        default:
#pragma unroll
            for (int i = 0; i < CNT; ++i) r.v[i] = src_at(s, 8 * k + i);
    }
}

// level l (the source, 2 * half elements) -> level l-1 at dp[e * 32], and with `fused` also level l-2 at dp2[e * 32] (f of adjacent
// pairs of level l-1: the MINUS (l-1) that follows).  u bits of the left child's codeword at rp[word * 32].  All lanes run the
// loop (the staged source is warp-wide); lanes without a path skip the arithmetic.  PLUS is a compile-time parameter and the
// four (+ two) node updates of a step run in lock step (node_lockstep), so their dependency chains overlap; the rare operands
// outside the fast division's range are redone after.  The loop is unrolled by two over a pair of element buffers: the raw
// elements of step k+1 are fetched before step k is computed and no buffer is ever copied.
// destination policies: pd / pd2 != 0 -> the level is in GLOBAL memory and stored with that L2 policy, 0 -> generic store (shared memory)
struct DstPol {
    uint64_t pd, pd2;
};
__device__ __forceinline__ void vstore(double2 *p, const double2 v, const uint64_t pol) {
    if (pol)
        st_global_hint(p, v, pol);
    else
        *p = v;
}
template <bool PLUS>
__device__ __forceinline__ void fg_step(const Raw8 &r, const uint32_t u4, double2 *__restrict__ &dp, double2 *__restrict__ &dp2,
                                        const bool fused, const DstPol dpol) {
    double2 y[4];
    const uint32_t slow = node_lockstep<4, PLUS>(r.v, u4, y);
    if (slow) {  // rare: an operand below 1e-291 (or an un-normalised channel pair): the IEEE division
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if ((slow >> i) & 1u) y[i] = node_fg(r.v[2 * i], r.v[2 * i + 1], PLUS, (u4 >> i) & 1u);
    }
    vstore(dp, y[0], dpol.pd), vstore(dp + 32, y[1], dpol.pd), vstore(dp + 64, y[2], dpol.pd), vstore(dp + 96, y[3], dpol.pd);
    dp += 128;
    if (fused) {
        double2 z[2];
        const uint32_t s2 = node_lockstep<2, false>(y, 0u, z);
        if (s2) {
            if (s2 & 1u) z[0] = node_fg(y[0], y[1], false, 0u);
            if (s2 & 2u) z[1] = node_fg(y[2], y[3], false, 0u);
        }
        vstore(dp2, z[0], dpol.pd2), vstore(dp2 + 32, z[1], dpol.pd2);
        dp2 += 64;
    }
}

template <bool PLUS, bool PREFETCH>
__device__ __forceinline__ void fg_pass(const Src &src, double2 *__restrict__ dp, double2 *__restrict__ dp2, const bool fused,
                                        const uint32_t *rp, const int half, const bool valid, const DstPol dpol) {
    Raw8 A, B;
    if (half == 2) {  // four source elements: one half step
        src_fetch<4>(src, 0, A);
        if (valid) {
            const uint32_t w = PLUS ? rp[0] : 0u;
            const double2 y0 = node_fg(A.v[0], A.v[1], PLUS, w & 1u), y1 = node_fg(A.v[2], A.v[3], PLUS, (w >> 1) & 1u);
            dp[0] = y0, dp[32] = y1;
            if (fused) dp2[0] = node_fg(y0, y1, false, 0u);
        }
        return;
    }
    const int nsteps = half >> 2;  // four elements of level l-1 per step: 1, 2, 4, ...
    if (src.kind == SK_STAGED && src.lane == 0) {
        const int pre = nsteps < src.nst ? nsteps : src.nst;
        for (int k = 0; k < pre; ++k) src_issue(src, k);
    }
    // u bits: word w of the codeword covers eight steps; the next word is requested as soon as a word is taken into use
    const bool pv = PLUS && valid;
    uint32_t w = pv ? rp[0] : 0u, wn = (pv && nsteps > 8) ? rp[32] : 0u;
    if (!PREFETCH) {
        // register-lean variant (16 warps per SM): one element buffer, the operands of a step are fetched right before it; the
        // HBM levels are still prefetched by the bulk-copy ring
#pragma unroll 1
        for (int k = 0; k < nsteps; ++k) {
            src_fetch<8>(src, k, A);
            if (valid) {
                const int j = (4 * k) & 31;
                fg_step<PLUS>(A, (w >> j) & 15u, dp, dp2, fused, dpol);
                if (PLUS && j == 28) {
                    w = wn;
                    if (k + 9 < nsteps) wn = rp[((4 * k + 36) >> 5) << 5];
                }
            }
        }
        return;
    }
    src_fetch<8>(src, 0, A);
    if (nsteps == 1) {
        if (valid) fg_step<PLUS>(A, w & 15u, dp, dp2, fused, dpol);
        return;
    }
#pragma unroll 1
    for (int k = 0; k < nsteps; k += 2) {
        src_fetch<8>(src, k + 1, B);
        if (valid) fg_step<PLUS>(A, (w >> ((4 * k) & 31)) & 15u, dp, dp2, fused, dpol);
        if (k + 2 < nsteps) src_fetch<8>(src, k + 2, A);
        if (valid) {
            const int j = (4 * k + 4) & 31;
            fg_step<PLUS>(B, (w >> j) & 15u, dp, dp2, fused, dpol);
            if (PLUS && j == 28) {
                w = wn;
                if (k + 10 < nsteps) wn = rp[((4 * k + 40) >> 5) << 5];
            }
        }
    }
}

__device__ __forceinline__ double group_max(double v, const int G) {
    for (int o = 1; o < G; o <<= 1) {
        const double x = __shfl_xor_sync(0xffffffffu, v, o);
        v = x > v ? x : v;
    }
    return v;
}
__device__ __forceinline__ int group_sum(int v, const int G) {
    for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The prune of recursiveListDecode (:446-451 etc.) over the M <= 8 candidates each lane of a frame holds: keep the
// ns = min(#nonzero, L) largest under the total order (metric, index), listed ascending; lists with at most L candidates keep
// all of them in index order (the same rounds keyed by the index alone).  Candidate i of the lane has index ibase + i * istep.
// On return lane t < nout holds its new path: the metric and the candidate index.
// The metrics are non-negative float64, so their bit patterns order like unsigned integers: a lane first sorts its own
// candidates (descending (key, index); a sorting network in registers), then every round is one arg-max of the lanes' heads
// over the frame's lane group -- three redux.sync (high word, low word among the equals, index among the equals) -- and the
// winning lane pops its head.
#ifndef SCLP_SELECT_SHFL
#define SCLP_SELECT_SHFL 1
#endif
struct Cand {
    uint32_t hi, lo;
    int idx;  // -1: no candidate
};
__device__ __forceinline__ void cand_cx(Cand &a, Cand &b) {  // a >= b after
    const bool sw = b.hi > a.hi || (b.hi == a.hi && (b.lo > a.lo || (b.lo == a.lo && b.idx > a.idx)));
    const Cand x = a;
    a.hi = sw ? b.hi : a.hi, a.lo = sw ? b.lo : a.lo, a.idx = sw ? b.idx : a.idx;
    b.hi = sw ? x.hi : b.hi, b.lo = sw ? x.lo : b.lo, b.idx = sw ? x.idx : b.idx;
}
template <int M>
__device__ __forceinline__ void select_paths(const double (&cv)[8], const int ibase, const int istep, const bool valid, const int cnt,
                                             const int L, const int G, const int t, const uint32_t gmask, int &nout,
                                             double &newprob, int &ci) {
    const int C = cnt * M;
    int nzl = 0;
#pragma unroll
    for (int i = 0; i < M; ++i) nzl += (valid && cv[i] != 0.0) ? 1 : 0;
    const int nz = G == 1 ? nzl : (int)__reduce_add_sync(gmask, (unsigned)nzl);
    const bool byidx = C <= L;
    const int ns = byidx ? C : (nz < L ? nz : L);
    nout = ns;
    newprob = 0.0;
    ci = 0;
    if (byidx) {  // list growth: new path t is candidate t (the candidate indices of a frame are 0 .. C-1)
        const int own = istep == 1 ? t / M : (cnt ? t % cnt : 0), slot = istep == 1 ? t % M : (cnt ? t / cnt : 0);
#pragma unroll
        for (int i = 0; i < M; ++i) {
            const double v = __shfl_sync(gmask, cv[i], (own & (G - 1)) + (int)(__ffs((int)gmask) - 1));
            if (t < C && slot == i) newprob = v;
        }
        ci = t < C ? t : 0;
        return;
    }
    Cand c[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        c[i].hi = valid ? (uint32_t)__double2hiint(cv[i]) : 0u;
        c[i].lo = valid ? (uint32_t)__double2loint(cv[i]) : 0u;
        c[i].idx = valid ? ibase + i * istep : -1;
    }
    if (M == 2) {
        cand_cx(c[0], c[1]);
    } else if (M == 4) {
        cand_cx(c[0], c[1]), cand_cx(c[2], c[3]), cand_cx(c[0], c[2]), cand_cx(c[1], c[3]), cand_cx(c[1], c[2]);
    } else {  // 19 compare-exchanges
        cand_cx(c[0], c[1]), cand_cx(c[2], c[3]), cand_cx(c[4], c[5]), cand_cx(c[6], c[7]);
        cand_cx(c[0], c[2]), cand_cx(c[1], c[3]), cand_cx(c[4], c[6]), cand_cx(c[5], c[7]);
        cand_cx(c[1], c[2]), cand_cx(c[5], c[6]), cand_cx(c[0], c[4]), cand_cx(c[3], c[7]);
        cand_cx(c[1], c[5]), cand_cx(c[2], c[6]);
        cand_cx(c[1], c[4]), cand_cx(c[3], c[6]);
        cand_cx(c[2], c[4]), cand_cx(c[3], c[5]);
        cand_cx(c[3], c[4]);
    }
#pragma unroll 1
    for (int r = 0; r < ns; ++r) {
        uint32_t mh, ml;
        int bi;
        if (SCLP_SELECT_SHFL) {  // butterfly arg-max inside the lane group: 3 shuffles per stage
            mh = c[0].hi, ml = c[0].lo, bi = c[0].idx;
            for (int o = 1; o < G; o <<= 1) {
                const uint32_t oh = __shfl_xor_sync(gmask, mh, o), ol = __shfl_xor_sync(gmask, ml, o);
                const int oi = __shfl_xor_sync(gmask, bi, o);
                const bool better = oh > mh || (oh == mh && (ol > ml || (ol == ml && oi > bi)));
                mh = better ? oh : mh, ml = better ? ol : ml, bi = better ? oi : bi;
            }
        } else {
            mh = __reduce_max_sync(gmask, c[0].hi);
            const bool e1 = c[0].hi == mh;
            ml = __reduce_max_sync(gmask, e1 ? c[0].lo : 0u);
            const bool e2 = e1 && c[0].lo == ml;
            bi = (int)__reduce_max_sync(gmask, e2 ? (uint32_t)(c[0].idx + 1) : 0u) - 1;
        }
        if (t == ns - 1 - r) {
            newprob = __hiloint2double((int)mh, (int)ml);
            ci = bi;
        }
        const bool pop = c[0].idx == bi && bi >= 0;  // candidate indices are unique inside a frame
#pragma unroll
        for (int i = 0; i + 1 < M; ++i) {
            c[i].hi = pop ? c[i + 1].hi : c[i].hi;
            c[i].lo = pop ? c[i + 1].lo : c[i].lo;
            c[i].idx = pop ? c[i + 1].idx : c[i].idx;
        }
        if (pop) c[M - 1].hi = 0u, c[M - 1].lo = 0u, c[M - 1].idx = -1;
    }
}

// MB = resident warps per SM the build is register-capped for.  The decoder is bound by per-warp latency (profiles/r2_a): its
// throughput is linear in the resident warps as long as nothing spills.  Two builds: 12 warps (168 registers: eight-element steps
// with the next step's operands prefetched into registers) and 16 warps (128 registers: the same steps, operands fetched right
// before use -- the HBM levels still arrive through the bulk-copy ring): 16 warps decode 18 % more frames per second.
constexpr int SCLP_MAX_WARPS_PER_SM = 16;
constexpr int SCLP_DEFAULT_WARPS_PER_SM = 16;
template <int MB>
__global__ void __launch_bounds__(32 * MB, 1) sclp_kernel(const SclpParams p) {
    PC_DYN_SMEM(smem_all);
    // a CTA is p.wpc independent warps (one CTA per SM); each warp owns its slice of the shared memory and its frames.  With
    // p.sync the warps that share a scheduler (warp index mod 4) meet at a named barrier before every op, so that they fetch
    // the same instructions at the same time.
    const int wic = threadIdx.x >> 5;
    unsigned char *smem_raw = smem_all + (size_t)wic * p.smem_per_warp;
    const int64_t gw = (int64_t)blockIdx.x * p.wpc + wic, nwarps = (int64_t)gridDim.x * p.wpc;
    const int n = p.n, N = 1 << n, L = p.L, G = p.G, gsh = p.gsh, lsm = p.lsm, NW = p.NW, rgl = p.rgl;
    const int lane = threadIdx.x & 31, t = lane & (G - 1), gbase = lane & ~(G - 1);
    constexpr uint32_t FULL = 0xffffffffu;
    // ---- shared-memory carve-up (sclp_smem_bytes mirrors this) ----
    const int vrows = (2 << lsm) - 1;
    double2 *Vs = (double2 *)smem_raw;
    uint32_t *Rs = (uint32_t *)(Vs + vrows * 32);
    uint8_t *om = (uint8_t *)(Rs + 2 * scl2_wsum(rgl) * 32);
    double2 *stg = (double2 *)(smem_raw + (((size_t)((char *)(om + (n + 1) * 2 * 32) - (char *)smem_raw) + 15) & ~(size_t)15));
    uint64_t *bars = (uint64_t *)(stg + p.nst * 256);
    uint32_t ph = 0;
    if (p.nst) {
        if (lane == 0) {
            for (int s = 0; s < p.nst; ++s) mbar_init(bars + s, 1u);
            mbar_fence_init();
        }
        __syncwarp();
    }
    const int64_t gws = p.alias ? wic : gw;
    double2 *Vg = p.vg + gws * p.vg_stride - (int64_t)vrows * 32;
    uint32_t *Rg = p.rg + gw * p.rg_stride - (int64_t)2 * scl2_wsum(rgl) * 32;
    // level l of the path vectors: rows (2^l - 1) .. ; per-path layout: element e of slot s at [e][gbase + s]; shared layout
    // (one path, written before the first fork): element e at [e / G][gbase + e % G]
    auto vbase = [&](int l) -> double2 * { return (l <= lsm ? Vs : Vg) + (((int64_t)1 << l) - 1) * 32; };
    auto rbase = [&](int l, int c) -> uint32_t * { return (l < rgl ? Rs : Rg) + (int64_t)(2 * scl2_wsum(l) + c * scl2_W(l)) * 32; };
    auto OM = [&](int l, int c) -> uint8_t * { return om + (l * 2 + c) * 32; };
    // L2 policy of level l's vectors: 0 for the shared-memory levels (generic accesses) and when the hints are off
    const uint64_t pol_first = p.l2lvl ? l2_policy_evict_first() : 0, pol_last = p.l2lvl ? l2_policy_evict_last() : 0;
    auto lvl_pol = [&](int l) -> uint64_t { return (l <= lsm || !p.l2lvl) ? (uint64_t)0 : (l <= p.l2lvl ? pol_last : pol_first); };
    const int fpw = 32 >> gsh;
    const bool dualon = G >= 4;  // two-variant storage of SCLP_DUAL outputs pays with four or more paths

#pragma unroll 1
    // every warp of the grid runs the same number of rounds (the barriers below need that); surplus rounds decode a copy of
    // the last frame and write nothing
    const int64_t nwaves = (p.frames + fpw - 1) / fpw, rounds = (nwaves + nwarps - 1) / nwarps;
#pragma unroll 1
    for (int64_t rd = 0; rd < rounds; ++rd) {
        const int64_t wave = rd * nwarps + gw;
        int64_t f = wave * fpw + (lane >> gsh);
        const bool fvalid = f < p.frames;
        if (!fvalid) f = p.frames - 1;  // idle lane groups decode a copy of the last frame and write nothing
        const double2 *xyf = p.xy ? p.xy + f * N : nullptr;
        const uint8_t *yf = p.y ? p.y + f * N : nullptr;
        const uint32_t *Ff = p.Fb ? p.Fb + f * NW : nullptr;
        double prob = 1.0;
        int cnt = 1, leaf = 0;
        long long tprev = 0;
        int tkey = 0;
        __syncwarp();
#pragma unroll 1
        for (int oi = 0; oi < p.n_ops; ++oi) {
            if (p.sync == 1)
                pc_named_barrier(1 + (wic & 3), 32 * ((p.wpc - (wic & 3) + 3) >> 2));
            else if (p.sync == 2)
                pc_named_barrier(1, 32 * p.wpc);
            const uint4 op = p.ops[oi];
            const int kind = op.x & 7, l = (op.x >> 3) & 15, c = (op.x >> 7) & 1, i0 = (int)op.y;
#ifndef PC_EMU
            if (p.timing) {
                const long long now = clock64();
                if (oi > 0 && lane == 0) atomicAdd(p.timing + tkey, (unsigned long long)(now - tprev));
                tprev = now;
                tkey = kind * 16 + l + (((op.x & SCLP_FUSED) && kind <= OP_PLUS) ? 0 : 0);
            }
#endif
            const bool ssrc = op.x & SCLP_SSRC, sdst = op.x & SCLP_SDST, chan = op.x & SCLP_CHAN;
            const bool dsrc = dualon && (op.x & SCLP_DSRC), ddst = dualon && (op.x & SCLP_DUAL);
            const int size = 1 << l;
            const bool valid = t < cnt;
            // element j of the source vector of this op, for the slot `slot` of the lane's frame
            auto ldsrc = [&](int e, int slot) -> double2 {
                if (chan) return p.sym ? p.tab[yf[e]] : xyf[e];
                const double2 *b = vbase(l);
                if (dsrc) {
                    const int i = 2 * e + (int)((rbase(l, 0)[((e >> 5) << 5) + gbase + slot] >> (e & 31)) & 1u);
                    return b[((i >> gsh) << 5) + gbase + (i & (G - 1))];
                }
                return ssrc ? b[((e >> gsh) << 5) + gbase + (e & (G - 1))] : b[(e << 5) + gbase + slot];
            };
            if (kind == OP_MINUS || kind == OP_PLUS) {
                const bool plus = kind == OP_PLUS, fused = op.x & SCLP_FUSED;
                const int half = size >> 1;
                int srcslot = t;
                if (plus && !sdst && valid) srcslot = OM(l - 1, 0)[lane];
                const uint32_t *rp = rbase(l - 1, 0) + (sdst ? gbase : lane);  // u bits: the left child's codeword of this path
                if (ddst || sdst) {
                    // The lanes of a frame share the items of ONE vector, float64 pairs in the shared layout:
                    //  * cooperative mode before the first fork (sdst): the single path's f / g pass, one level per round;
                    //  * two-variant outputs (ddst): one shared parent (the channel or a shared-layout vector) and four or more
                    //    paths -- item 2 e + v = g(a_e, b_e, v); the paths pick their variant by their codeword bit when they read
                    //    the vector (SK_DUAL / ldsrc).
                    const int npass = (sdst && fused) ? 2 : 1;
#pragma unroll 1
                    for (int ps = 0; ps < npass; ++ps) {
                        const int lv = l - ps;
                        const bool pl = plus && ps == 0, ch = chan && ps == 0;
                        const double2 *sb = ch ? nullptr : vbase(lv);
                        double2 *db = vbase(lv - 1);
                        const int items = ddst ? (1 << lv) : (1 << (lv - 1));
                        // four items per lane and step: the eight loads go out together, the four node updates interleave
#pragma unroll 1
                        for (int i0 = t; i0 < items; i0 += 4 * G) {
                            double2 a[4], b[4];
                            uint32_t u[4];
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int i = i0 + r * G;
                                a[r] = b[r] = make_double2(1.0, 1.0);
                                u[r] = 0;
                                if (i < items) {
                                    const int e = ddst ? i >> 1 : i;
                                    if (ch) {
                                        if (p.sym)
                                            a[r] = p.tab[yf[2 * e]], b[r] = p.tab[yf[2 * e + 1]];
                                        else
                                            a[r] = xyf[2 * e], b[r] = xyf[2 * e + 1];
                                    } else {
                                        const double2 *q = sb + (((2 * e) >> gsh) << 5) + gbase + ((2 * e) & (G - 1));
                                        a[r] = q[0], b[r] = G >= 2 ? q[1] : q[32];  // one lane per frame: consecutive elements sit in consecutive rows
                                    }
                                    u[r] = ddst ? (uint32_t)(i & 1) : (pl ? (rp[(e >> 5) << 5] >> (e & 31)) & 1u : 0u);
                                }
                            }
                            double2 y[4], in[8];
#pragma unroll
                            for (int r = 0; r < 4; ++r) in[2 * r] = a[r], in[2 * r + 1] = b[r];
                            const uint32_t ub = u[0] | u[1] << 1 | u[2] << 2 | u[3] << 3;
                            const uint32_t sl = pl ? node_lockstep<4, true>(in, ub, y) : node_lockstep<4, false>(in, 0u, y);
                            if (sl) {
#pragma unroll
                                for (int r = 0; r < 4; ++r)
                                    if ((sl >> r) & 1u) y[r] = node_fg(a[r], b[r], pl, u[r]);
                            }
#pragma unroll
                            for (int r = 0; r < 4; ++r) {
                                const int i = i0 + r * G;
                                if (i < items) db[((i >> gsh) << 5) + gbase + (i & (G - 1))] = y[r];
                            }
                        }
                        if (ps + 1 < npass) __syncwarp();
                    }
                } else {
                    // the lane's own path: per-path destination(s)
                    Src sr;
                    sr.gsh = gsh, sr.ss = 32, sr.pol = 0;
                    if (chan) {
                        sr.kind = p.sym ? SK_SYM : SK_PATH;
                        sr.y = yf, sr.tab = p.tab, sr.p = xyf, sr.ss = 1;
                    } else if (dsrc || ssrc) {
                        sr.kind = dsrc ? SK_DUAL : SK_SHARED;
                        sr.p = vbase(l) + gbase, sr.bits = rbase(l, 0) + gbase + srcslot;
                    } else if (l > lsm && p.nst && half >= 4) {
                        sr.kind = SK_STAGED;
                        sr.p = vbase(l), sr.stg = stg, sr.bars = bars, sr.ph = &ph, sr.nst = p.nst, sr.nsteps = half >> 2;
                        sr.col = gbase + srcslot, sr.lane = lane;
                        sr.pol = lvl_pol(l);
                    } else {
                        sr.kind = SK_PATH;
                        sr.p = vbase(l) + gbase + srcslot;
                    }
                    const DstPol dpol{lvl_pol(l - 1), lvl_pol(l >= 2 ? l - 2 : 0)};
                    if (plus)
                        fg_pass<true, (MB <= 12)>(sr, vbase(l - 1) + lane, vbase(l >= 2 ? l - 2 : 0) + lane, fused, rp, half, valid, dpol);
                    else
                        fg_pass<false, (MB <= 12)>(sr, vbase(l - 1) + lane, vbase(l >= 2 ? l - 2 : 0) + lane, fused, rp, half, valid, dpol);
                }
                __syncwarp();
                continue;
            }
            if (kind == OP_COMBINE) {  // :726-754 in reference order: out[2h] = m[h] + p[h], out[2h+1] = -p[h]
                if (valid) {
                    const int mi = OM(l - 1, 1)[lane];
                    const uint32_t *rm = rbase(l - 1, 0) + gbase + mi, *rq = rbase(l - 1, 1) + lane;
                    uint32_t *ro = rbase(l, c) + lane;
                    const int Wo = scl2_W(l);
#pragma unroll 1
                    for (int w = 0; w < Wo; ++w) {
                        const int sh = (w & 1) * 16;
                        const uint32_t m16 = (rm[(w >> 1) << 5] >> sh) & 0xffffu, p16 = (rq[(w >> 1) << 5] >> sh) & 0xffffu;
                        ro[w << 5] = spread16(m16 ^ p16) | (spread16(p16) << 1);
                    }
                    OM(l, c)[lane] = OM(l - 1, 0)[gbase + mi];
                }
                __syncwarp();
                continue;
            }
            // ------------------------------- fast nodes ------------------------------------------------------
            // Rate-0 :495-518, Rep :521-578, Rate-1 :581-628, SPC :631-682.  One element loop serves all four kinds (code size);
            // the order-dependent float64 products (np.product is a left-to-right product) run one path per lane.
            const int li = leaf++;
            const int Wl = scl2_W(l);
            const uint32_t smask = size >= 32 ? 0xffffffffu : ((1u << size) - 1u);
            const int bsh = size >= 32 ? 0 : (i0 & 31);  // node-local slice of the frame-wide bit arrays: word i0/32 + w, shifted by bsh
            auto fslice = [&](int w) -> uint32_t { return Ff ? (Ff[(i0 >> 5) + w] >> bsh) & smask : 0u; };
            const bool rep = kind == OP_REP, fork = kind == OP_RATE1 || kind == OP_SPC, spc = kind == OP_SPC;
            const bool plain = !chan && !dsrc && !ssrc;  // the lane's own vector, per-path layout
            const uint32_t *coefw = p.coef_words + op.w;
            uint32_t *ro = rbase(l, c) + lane;
            uint32_t *hdp = (fork && l >= 6) ? rbase(l - 1, 0) + lane : nullptr;  // hard-decision words of large nodes: the dead child areas
            const int nb = size < 4 ? size : 4;
            int nout = cnt, src = t;
            double newprob = 0.0;
            double cv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) cv[i] = 0.0;
            uint32_t hdw = 0, par = 0, delta = 0;
            int pk0 = -1, pk1 = -1, pk2 = -1, pk3 = -1;
            if (valid) {
                double pr0 = 1.0, pr1 = 1.0;                         // Rate-0 / Rep products; pass 2: product of the non-forked maxima
                double s0 = -1.0, s1 = -1.0, s2 = -1.0, s3 = -1.0;    // pass 1 of Rate-1 / SPC: the four largest (score, j), ascending
                int j0 = 0, j1 = 0, j2 = 0, j3 = 0;
                uint32_t fb = 0, cb = 0;
                const int npass = fork ? 2 : 1;
#pragma unroll 1
                for (int ps = 0; ps < npass; ++ps) {
#pragma unroll 1
                    for (int jb = 0; jb < size; jb += 4) {
                        // four consecutive elements of the lane's own vector: independent loads ahead of the sequential chains
                        double2 v[4];
                        if (plain) {
                            const double2 *q = vbase(l) + (jb << 5) + lane;
                            const uint64_t pl = lvl_pol(l);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                v[i] = i < nb ? (pl ? ld_global_hint(q + (i << 5), pl) : q[i << 5]) : make_double2(1.0, 1.0);
                        } else {
#pragma unroll 1
                            for (int i = 0; i < nb; ++i) {  // the channel, shared-layout and two-variant vectors: few nodes per frame
                                const double2 x = ldsrc(jb + i, t);
                                if (i == 0) v[0] = x;
                                if (i == 1) v[1] = x;
                                if (i == 2) v[2] = x;
                                if (i == 3) v[3] = x;
                            }
                        }
                        if (!fork && (jb & 31) == 0) fb = fslice(jb >> 5), cb = rep ? fb ^ coefw[jb >> 5] : 0u;
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (i < nb) {
                                const int j = jb + i;
                                const double2 v2 = v[i];
                                if (!fork) {
                                    // products of P[j].{x|y} selected by the candidate codeword's bits
                                    pr0 = __dmul_rn(pr0, (fb >> (j & 31)) & 1u ? v2.y : v2.x);
                                    if (rep) pr1 = __dmul_rn(pr1, (cb >> (j & 31)) & 1u ? v2.y : v2.x);
                                } else if (ps == 0) {
                                    // hard decisions, reliabilities (second-largest / largest, :763-768) and the 2 (Rate-1) or 4 (SPC)
                                    // largest (score, j), ties to the later index (pickLeastReliableIndices :759-761), kept ascending
                                    // s0 <= .. <= s3 by a branch-free insertion
                                    const bool one = v2.y > v2.x;
                                    const double s = (one ? v2.x : v2.y) / (one ? v2.y : v2.x);
                                    const bool g0 = s >= s0, g1 = s >= s1, g2 = s >= s2, g3 = s >= s3;
                                    s0 = g1 ? s1 : (g0 ? s : s0), j0 = g1 ? j1 : (g0 ? j : j0);
                                    s1 = g2 ? s2 : (g1 ? s : s1), j1 = g2 ? j2 : (g1 ? j : j1);
                                    s2 = g3 ? s3 : (g2 ? s : s2), j2 = g3 ? j3 : (g2 ? j : j2);
                                    s3 = g3 ? s : s3, j3 = g3 ? j : j3;
                                    hdw |= (one ? 1u : 0u) << (j & 31);
                                } else {
                                    // in element order: product of the non-forked maxima (:785-788, :814-817)
                                    const bool forked = j == pk0 || j == pk1 || j == pk2 || j == pk3;
                                    pr0 = forked ? pr0 : __dmul_rn(pr0, v2.y > v2.x ? v2.y : v2.x);
                                }
                            }
                        if (fork && ps == 0 && (((jb + 3) & 31) == 31 || jb + 4 >= size)) {
                            par ^= hdw;
                            if (hdp) {
                                hdp[(jb >> 5) << 5] = hdw;
                                hdw = 0;
                            }
                        }
                    }
                    if (fork && ps == 0) {  // picks in ascending (score, index) order
                        if (spc)
                            pk0 = j0, pk1 = j1, pk2 = j2, pk3 = j3;
                        else
                            pk0 = j2, pk1 = j3;
                    }
                }
                if (!fork) {
                    cv[0] = __dmul_rn(prob, pr0);
                    cv[1] = __dmul_rn(prob, pr1);
                } else {
                    const double basep = __dmul_rn(prob, pr0);
                    // parity of the non-forked hard decisions
                    auto hbit = [&](int j) -> uint32_t {
                        const uint32_t w = hdp ? hdp[(j >> 5) << 5] : hdw;
                        return (w >> (j & 31)) & 1u;
                    };
                    uint32_t pp = __popc(par) & 1u;
                    pp ^= hbit(pk0) ^ hbit(pk1);
                    if (spc) pp ^= hbit(pk2) ^ hbit(pk3);
                    const uint32_t fval = spc && p.fvp ? (p.fvp[f * p.nfw + (op.z >> 5)] >> (op.z & 31)) & 1u : 0u;
                    delta = (fval ^ pp) & 1u;
                    // candidate metrics (forkIndices / forkIndicesSpc, :770-820): digits of the fork index, first pick outermost.
                    // The partial products are built pick by pick: c[fk] = ((f0 * f1) * f2) * f3 in that order.
                    const int npk = spc ? 4 : 2;
#pragma unroll 1
                    for (int w = 0; w < npk; ++w) {
                        const int pj = w == 0 ? pk0 : (w == 1 ? pk1 : (w == 2 ? pk2 : pk3));
                        const double2 fw = ldsrc(pj, t);
                        if (w == 0) {
#pragma unroll
                            for (int fk = 0; fk < 8; ++fk) cv[fk] = (spc ? (fk >> 2) : (fk >> 1)) & 1 ? fw.y : fw.x;
                        } else {
#pragma unroll
                            for (int fk = 0; fk < 8; ++fk) {
                                // digit of pick w in fork index fk: Rate-1 has two digits (w = 0, 1), SPC three plus the dependent pick
                                const uint32_t dg = w == 3 ? (delta ^ (uint32_t)__popc((uint32_t)fk)) & 1u
                                                           : (uint32_t)(fk >> (spc ? 2 - w : 1 - w)) & 1u;
                                cv[fk] = __dmul_rn(cv[fk], dg ? fw.y : fw.x);
                            }
                        }
                    }
#pragma unroll
                    for (int fk = 0; fk < 8; ++fk) cv[fk] = (spc || fk < 4) ? __dmul_rn(cv[fk], basep) : 0.0;
                }
            }
            // prune / order the candidates (Rate-0 keeps its paths as they are)
            int sel = 0, fk = 0;
            if (kind == OP_RATE0) {
                newprob = cv[0];
            } else {
                int ci;
                const int M = rep ? 2 : (spc ? 8 : 4);
                const uint32_t gmask = G == 32 ? 0xffffffffu : ((1u << G) - 1u) << gbase;
                if (rep)
                    select_paths<2>(cv, t, cnt, valid, cnt, L, G, t, gmask, nout, newprob, ci);
                else if (spc)
                    select_paths<8>(cv, t * 8, 1, valid, cnt, L, G, t, gmask, nout, newprob, ci);
                else
                    select_paths<4>(cv, t * 4, 1, valid, cnt, L, G, t, gmask, nout, newprob, ci);
                if (rep)
                    sel = ci >= cnt ? 1 : 0, src = ci - sel * cnt;
                else
                    src = ci / M, fk = ci - src * M;
            }
            if (!fork) {
                if (t < nout) {
#pragma unroll 1
                    for (int w = 0; w < Wl; ++w) ro[w << 5] = fslice(w) ^ (sel ? coefw[w] : 0u);
                    OM(l, c)[lane] = (uint8_t)src;
                }
            } else {
                if (hdp) __syncwarp();  // the hard-decision words of the source paths
                // the source path's picks, parity and hard decisions; forked positions overwritten
                const int sl = gbase + (t < nout ? src : t);
                const uint32_t pa = (uint32_t)(pk0 & 0xffff) | (uint32_t)(pk1 & 0xffff) << 16;
                const uint32_t pb = (uint32_t)(pk2 & 0xffff) | (uint32_t)(pk3 & 0xffff) << 16;
                const uint32_t qa = __shfl_sync(FULL, pa, sl), qb = __shfl_sync(FULL, pb, sl);
                const uint32_t sdelta = __shfl_sync(FULL, delta, sl), shd = __shfl_sync(FULL, hdw, sl);
                if (t < nout) {
                    const int q0 = qa & 0xffff, q1 = qa >> 16, q2 = qb & 0xffff, q3 = qb >> 16;
                    const int nfork = spc ? 3 : 2;
                    const uint32_t *hs = l >= 6 ? rbase(l - 1, 0) + sl : nullptr;
                    // hard-decision words are read before any word of this node is written: for l >= 6 the node's codeword area
                    // (level l) and the temporaries (level l-1) are distinct
#pragma unroll 1
                    for (int w = 0; w < Wl; ++w) {
                        uint32_t v = hs ? hs[w << 5] : shd;
                        const int qq[3] = {q0, q1, q2};
#pragma unroll
                        for (int i = 0; i < 3; ++i)
                            if (i < nfork && (qq[i] >> 5) == w)
                                v = (v & ~(1u << (qq[i] & 31))) | ((uint32_t)((fk >> (nfork - 1 - i)) & 1) << (qq[i] & 31));
                        if (spc && (q3 >> 5) == w) v = (v & ~(1u << (q3 & 31))) | (((sdelta ^ (uint32_t)__popc((uint32_t)fk)) & 1u) << (q3 & 31));
                        ro[w << 5] = v;
                    }
                    OM(l, c)[lane] = (uint8_t)src;
                }
            }
            // normalise (:867-872), count
            {
                const double mx = group_max(t < nout ? newprob : -1.0, G);
                prob = newprob / mx;
                if (t == 0 && fvalid) p.mxs[f * p.n_leaf + li] = mx;
                cnt = nout;
            }
            __syncwarp();
        }
#ifndef PC_EMU
        if (p.timing && lane == 0) atomicAdd(p.timing + tkey, (unsigned long long)(clock64() - tprev));
#endif
        // ---- final list: is the actual word in it (listDecode :172-213 compares the information vectors; codewords here) ----
        {
            const bool valid = t < cnt;
            const uint32_t *root = rbase(n, 0) + lane, *aw = p.Acw + f * NW;
            bool eq = valid;
#pragma unroll 1
            for (int w = 0; w < NW; ++w) eq = eq && root[w << 5] == aw[w];
            const uint32_t bal = __ballot_sync(FULL, eq);
            const uint32_t grp = G == 32 ? bal : (bal >> gbase) & ((1u << G) - 1u);
            const int found = grp ? __ffs((int)grp) - 1 : -1;
            if (fvalid) {
                if (t == 0) {
                    p.found[f] = found;
                    p.lsize[f] = cnt;
                }
                if (t < L) p.lprob[f * L + t] = valid ? prob : 0.0;
                const int sel = found >= 0 ? found : 0;
                if (p.want_list) {
                    if (valid) {
                        uint32_t *o = p.selcw + (f * L + t) * NW;
#pragma unroll 1
                        for (int w = 0; w < NW; ++w) o[w] = root[w << 5];
                    }
                } else if (t == sel) {
                    uint32_t *o = p.selcw + f * NW;
#pragma unroll 1
                    for (int w = 0; w < NW; ++w) o[w] = root[w << 5];
                }
            }
        }
        __syncwarp();
    }
}

// ---- prep: codewords of the actual word and of the frozen values ------------------------------------------------------------
struct SclpPrepParams {
    int n, k, NW, nfw, kw;
    int64_t frames;
    const int32_t *a_src, *f_src, *perm;
    const uint32_t *stage_mask;
    const uint32_t *ainfo;  // [frames][kw] packed actual information
    const uint32_t *fvp;    // [frames][nfw] packed frozen values or null
    uint32_t *Acw, *Ab, *Fb;  // [frames][NW]: root codeword of the actual word; node-local codewords (reference order) of the
                              // actual word and of the frozen values (Fb null when fvp is null)
};

constexpr int SCLP_AUX_WARPS = 4;  // warps per block of the prep / final kernels: independent frames, own shared-memory slices
__global__ void __launch_bounds__(32 * SCLP_AUX_WARPS) sclp_prep_kernel(const SclpPrepParams p) {
    PC_DYN_SMEM(smem_raw);
    const int n = p.n, N = 1 << n, NW = p.NW, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr uint32_t FULL = 0xffffffffu;
    uint32_t *T0 = (uint32_t *)smem_raw + (size_t)wib * (3 * NW + 4), *T1 = T0 + NW, *T2 = T1 + NW;
#pragma unroll 1
    for (int64_t f = (int64_t)blockIdx.x * SCLP_AUX_WARPS + wib; f < p.frames; f += (int64_t)gridDim.x * SCLP_AUX_WARPS) {
        const uint32_t *ai = p.ainfo + f * p.kw, *fv = p.fvp ? p.fvp + f * p.nfw : nullptr;
        __syncwarp();
        // u-domain bits in natural order: the actual word (A) and the frozen values alone (F)
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
            const int pos = 32 * w + lane;
            uint32_t a = 0, fb = 0;
            if (pos < N) {
                const int sa = p.a_src[pos], sf = p.f_src[pos];
                if (sa >= 0)
                    a = (ai[sa >> 5] >> (sa & 31)) & 1u;
                else if (fv)
                    a = (fv[(~sa) >> 5] >> ((~sa) & 31)) & 1u;
                if (sf >= 0 && fv) fb = (fv[sf >> 5] >> (sf & 31)) & 1u;
            }
            const uint32_t wa = __ballot_sync(FULL, a), wf = __ballot_sync(FULL, fb);
            if (lane == 0) {
                T0[w] = wa;
                T1[w] = wf;
                T2[w] = wa;
            }
        }
        __syncwarp();
        // root codeword of the actual word: full transform, then the bit reversal to the reference's order
#pragma unroll 1
        for (int st = 0; st < n; ++st) {
            const int s = 1 << st;
#pragma unroll 1
            for (int w = lane; w < NW; w += 32) {
                if (s < 32) {
                    const uint32_t m = s == 1 ? 0x55555555u : s == 2 ? 0x33333333u : s == 4 ? 0x0f0f0f0fu : s == 8 ? 0x00ff00ffu : 0x0000ffffu;
                    T2[w] ^= (T2[w] >> s) & m;
                } else if (!(w & (s >> 5))) {
                    T2[w] ^= T2[w + (s >> 5)];
                }
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
            const int pos = 32 * w + lane;
            uint32_t b = 0;
            if (pos < N) {
                const uint32_t r = bitrev_n((uint32_t)pos, n);
                b = (T2[r >> 5] >> (r & 31)) & 1u;
            }
            const uint32_t wv = __ballot_sync(FULL, b);
            if (lane == 0) p.Acw[f * NW + w] = wv;
        }
        // node-local codewords: masked butterfly up to each fast node's size, then the per-node bit reversal
#pragma unroll 1
        for (int st = 0; st < n; ++st) {
            const int s = 1 << st;
#pragma unroll 1
            for (int w = lane; w < NW; w += 32) {
                const uint32_t m = p.stage_mask[st * NW + w];
                if (m) {
                    if (s < 32) {
                        T0[w] ^= (T0[w] >> s) & m;
                        T1[w] ^= (T1[w] >> s) & m;
                    } else {
                        T0[w] ^= T0[w + (s >> 5)] & m;
                        T1[w] ^= T1[w + (s >> 5)] & m;
                    }
                }
            }
            __syncwarp();
        }
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
            const int i = 32 * w + lane;
            uint32_t a = 0, fb = 0;
            if (i < N) {
                const int sp = p.perm[i];
                a = (T0[sp >> 5] >> (sp & 31)) & 1u;
                fb = (T1[sp >> 5] >> (sp & 31)) & 1u;
            }
            const uint32_t wa = __ballot_sync(FULL, a), wf = __ballot_sync(FULL, fb);
            if (lane == 0) {
                p.Ab[f * NW + w] = wa;
                if (p.Fb) p.Fb[f * NW + w] = wf;
            }
        }
    }
}

// ---- final: selection, ProbResult, information bits; genie replay where listDecode's actual_prob is needed ------------------
struct SclpFinalParams {
    int n, k, L, NW, kw, n_ops, n_leaf, sym, want_list, lsmA, smem_per_warp;
    int64_t frames;
    const uint4 *ops;
    const int32_t *info_src;
    const double2 *xy;
    const uint8_t *y;
    const double2 *tab;
    const uint32_t *Ab, *selcw;
    const int32_t *found, *lsize;
    const double *lprob, *mxs;
    double2 *va;          // [grid][va_stride] the actual path's vectors of levels > lsmA
    int64_t va_stride;
    uint32_t *info;       // [frames][kw] packed information of the selected word
    int32_t *result;      // [frames] ProbResult
    int32_t *out_lsize;   // optional final-list outputs (caller layouts)
    double *out_lprob, *out_aprob;
    uint32_t *out_linfo;  // [frames][L][kw]
};

static size_t sclp_final_smem_bytes(int n, int lsmA) {
    const int NW = (1 << n) >= 32 ? (1 << n) >> 5 : 1;
    size_t b = (size_t)((2 << lsmA) - 1) * 16;          // actual path vectors of levels <= lsmA
    b += (size_t)(2 * scl2_wsum(n + 1) + 2) * 4;        // its codewords per (level, child)
    b += (size_t)NW * 4 + 16;                           // T0
    return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(32 * SCLP_AUX_WARPS) sclp_final_kernel(const SclpFinalParams p) {
    PC_DYN_SMEM(smem_all);
    const int n = p.n, N = 1 << n, NW = p.NW, L = p.L, k = p.k, lane = threadIdx.x & 31, wib = threadIdx.x >> 5, lsmA = p.lsmA;
    constexpr uint32_t FULL = 0xffffffffu;
    const int vrows = (2 << lsmA) - 1;
    unsigned char *smem_raw = smem_all + (size_t)wib * p.smem_per_warp;
    double2 *VAs = (double2 *)smem_raw;
    uint32_t *RA = (uint32_t *)(VAs + vrows);
    uint32_t *T0 = RA + 2 * scl2_wsum(n + 1) + 2;
    const int64_t gwarp = (int64_t)blockIdx.x * SCLP_AUX_WARPS + wib, gwarps = (int64_t)gridDim.x * SCLP_AUX_WARPS;
    double2 *VAg = p.va + gwarp * p.va_stride - vrows;
    auto vbase = [&](int l) -> double2 * { return (l <= lsmA ? VAs : VAg) + (((int64_t)1 << l) - 1); };
    auto rbase = [&](int l, int c) -> uint32_t * { return RA + 2 * scl2_wsum(l) + c * scl2_W(l); };
#pragma unroll 1
    for (int64_t f = gwarp; f < p.frames; f += gwarps) {
        const int found = p.found[f], cnt = p.lsize[f];
        const double *lp = p.lprob + f * L;
        double maxp = lp[0], minp = lp[0];
#pragma unroll 1
        for (int t = 1; t < cnt; ++t) {
            maxp = lp[t] > maxp ? lp[t] : maxp;
            minp = lp[t] < minp ? lp[t] : minp;
        }
        double ap = 1.0;
        __syncwarp();
        if (found < 0 || p.want_list) {
            // replay the op list for the actual path alone (the path listDecode tracks as actualXyVectorDistribution, :484-509 etc.)
            const double2 *xyf = p.xy ? p.xy + f * N : nullptr;
            const uint8_t *yf = p.y ? p.y + f * N : nullptr;
            const uint32_t *Af = p.Ab + f * NW;
            int leaf = 0;
#pragma unroll 1
            for (int oi = 0; oi < p.n_ops; ++oi) {
                const uint4 op = p.ops[oi];
                const int kind = op.x & 7, l = (op.x >> 3) & 15, c = (op.x >> 7) & 1, i0 = (int)op.y;
                const bool chan = op.x & SCLP_CHAN;
                const int size = 1 << l;
                auto ldsrc = [&](int lv, int e) -> double2 {
                    if (lv == n) return p.sym ? p.tab[yf[e]] : xyf[e];
                    return vbase(lv)[e];
                };
                (void)chan;
                if (kind == OP_MINUS || kind == OP_PLUS) {
                    const bool plus = kind == OP_PLUS;
                    const uint32_t *rp = rbase(l - 1, 0);
                    double2 *d1 = vbase(l - 1);
#pragma unroll 1
                    for (int h = lane; h < (size >> 1); h += 32) {
                        const uint32_t u = plus ? (rp[h >> 5] >> (h & 31)) & 1u : 0u;
                        d1[h] = node_update(ldsrc(l, 2 * h), ldsrc(l, 2 * h + 1), plus, u);
                    }
                    __syncwarp();
                    if (op.x & SCLP_FUSED) {
                        double2 *d2 = vbase(l - 2);
#pragma unroll 1
                        for (int h = lane; h < (size >> 2); h += 32) d2[h] = node_update(d1[2 * h], d1[2 * h + 1], false, 0u);
                        __syncwarp();
                    }
                    continue;
                }
                if (kind == OP_COMBINE) {
                    const uint32_t *rm = rbase(l - 1, 0), *rq = rbase(l - 1, 1);
                    uint32_t *ro = rbase(l, c);
                    const int Wo = scl2_W(l);
#pragma unroll 1
                    for (int w = lane; w < Wo; w += 32) {
                        const int sh = (w & 1) * 16;
                        const uint32_t m16 = (rm[w >> 1] >> sh) & 0xffffu, p16 = (rq[w >> 1] >> sh) & 0xffffu;
                        ro[w] = spread16(m16 ^ p16) | (spread16(p16) << 1);
                    }
                    __syncwarp();
                    continue;
                }
                // fast node: the actual path's product over the node (:503-509 etc.), normalised like the list was
                const int li = leaf++;
                const uint32_t smask = size >= 32 ? 0xffffffffu : ((1u << size) - 1u);
                const int bsh = size >= 32 ? 0 : (i0 & 31);
                const uint32_t *aw = Af + (i0 >> 5);
                if (lane == 0) {
                    double pr = 1.0;
#pragma unroll 1
                    for (int j = 0; j < size; ++j) {
                        const double2 v2 = ldsrc(l, j);
                        pr = __dmul_rn(pr, (aw[j >> 5] >> (bsh + (j & 31))) & 1u ? v2.y : v2.x);
                    }
                    ap = __dmul_rn(ap, pr / p.mxs[f * p.n_leaf + li]);
                }
                const int Wl = scl2_W(l);
#pragma unroll 1
                for (int w = lane; w < Wl; w += 32) rbase(l, c)[w] = (aw[w] >> bsh) & smask;
                __syncwarp();
            }
            ap = __shfl_sync(FULL, ap, 0);
        }
        if (lane == 0) {
            int res;
            if (found >= 0)
                res = lp[found] == maxp ? 0 : 1;
            else
                res = ap > maxp ? 2 : (ap == maxp ? 3 : (ap >= minp ? 4 : 5));
            p.result[f] = res;
            if (p.out_lsize) {
                p.out_lsize[f] = cnt;
                p.out_aprob[f] = ap;
#pragma unroll 1
                for (int t = 0; t < L; ++t) p.out_lprob[f * L + t] = t < cnt ? lp[t] : 0.0;
            }
        }
        // information of a path = gather of T(root codeword): bit-reverse to natural order, butterfly, gather
        const int sel = found >= 0 ? found : 0;
        const int npaths = p.out_linfo ? cnt : 1;
#pragma unroll 1
        for (int pi = 0; pi < npaths; ++pi) {
            const int t = p.out_linfo ? pi : sel;
            const uint32_t *root = p.want_list ? p.selcw + (f * L + t) * NW : p.selcw + f * NW;
            __syncwarp();
#pragma unroll 1
            for (int w = 0; w < NW; ++w) {
                const int pos = 32 * w + lane;
                uint32_t b = 0;
                if (pos < N) {
                    const uint32_t r = bitrev_n((uint32_t)pos, n);
                    b = (root[r >> 5] >> (r & 31)) & 1u;
                }
                const uint32_t wv = __ballot_sync(FULL, b);
                if (lane == 0) T0[w] = wv;
            }
            __syncwarp();
#pragma unroll 1
            for (int st = 0; st < n; ++st) {
                const int s = 1 << st;
#pragma unroll 1
                for (int w = lane; w < NW; w += 32) {
                    if (s < 32) {
                        const uint32_t m = s == 1 ? 0x55555555u : s == 2 ? 0x33333333u : s == 4 ? 0x0f0f0f0fu : s == 8 ? 0x00ff00ffu : 0x0000ffffu;
                        T0[w] ^= (T0[w] >> s) & m;
                    } else if (!(w & (s >> 5))) {
                        T0[w] ^= T0[w + (s >> 5)];
                    }
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int w = 0; w < p.kw; ++w) {
                const int j = 32 * w + lane;
                uint32_t b = 0;
                if (j < k) {
                    const int pos = p.info_src[j];
                    b = (T0[pos >> 5] >> (pos & 31)) & 1u;
                }
                const uint32_t wv = __ballot_sync(FULL, b);
                if (lane == 0) {
                    if (p.out_linfo) p.out_linfo[(f * L + pi) * p.kw + w] = wv;
                    if (t == sel) p.info[f * p.kw + w] = wv;
                }
            }
        }
    }
}

// ---- bit packing between the byte-per-symbol ABI (pc_scl_decode_probs) and the packed one --------------------------------
__global__ void __launch_bounds__(256) pack_rows_kernel(const uint8_t *in, uint32_t *out, int64_t rows, int nbits, int W) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = wid; i < rows * W; i += nw) {
        const int64_t r = i / W;
        const int w = (int)(i - r * W), j = 32 * w + lane;
        const uint32_t b = j < nbits ? in[r * nbits + j] & 1u : 0u;
        const uint32_t v = __ballot_sync(0xffffffffu, b);
        if (lane == 0) out[i] = v;
    }
}
__global__ void __launch_bounds__(256) unpack_rows_kernel(const uint32_t *in, uint8_t *out, int64_t rows, int nbits, int W) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rows * nbits; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / nbits;
        const int j = (int)(i - r * nbits);
        out[i] = (uint8_t)((in[r * W + (j >> 5)] >> (j & 31)) & 1u);
    }
}

struct SclpTab {
    double2 v[256];
};
__global__ void __launch_bounds__(256) sclp_tab_kernel(const SclpTab t, double2 *out) { out[threadIdx.x] = t.v[threadIdx.x]; }

// ---- host side ------------------------------------------------------------------------------------------------
struct SclpConfig {
    int G, gsh, lsm, rgl, lsmA, grid, grid2, per_sm, nst;
    size_t smem, smem2, vg_stride, rg_stride, va_stride;
    bool ok;
};

static int envp_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

// tuning knobs are read ONCE per process (a changed environment between the workspace query and the decode cannot break
// the sizing contract)
struct SclpKnobs {
    int warps_per_sm, lsm, rgl, off, nst, timing, skew, sync, alias, l2lvl;
    SclpKnobs() {
        warps_per_sm = envp_int("PC_SCLP_WARPS_PER_SM", SCLP_DEFAULT_WARPS_PER_SM);
        if (warps_per_sm > SCLP_MAX_WARPS_PER_SM) warps_per_sm = SCLP_MAX_WARPS_PER_SM;
        lsm = envp_int("PC_SCLP_LSM", -1);
        rgl = envp_int("PC_SCLP_RGL", 7);
        nst = envp_int("PC_SCLP_STAGES", 2);
        timing = envp_int("PC_SCLP_TIMING", 0);
        skew = envp_int("PC_SCLP_SKEW", 1);
        sync = envp_int("PC_SCLP_SYNC", 1);
        alias = envp_int("PC_SCLP_ALIAS", 0);
        l2lvl = envp_int("PC_SCLP_L2LVL", 5);  // +3 % (sweep: 0 / 4 / 5 / 6 / 7 / 12 -> 1.698 / 1.737 / 1.753 / 1.751 / 1.727 / 1.698 Gbit/s)
        off = envp_int("PC_SCL_GENERIC", 0);  // tests: force the generic (q <= 5, frame per lane) decoder for q = 2
    }
};
#ifdef PC_EMU
static SclpKnobs sclp_knobs() { return SclpKnobs(); }  // emulated tests vary the knobs between calls
#else
static const SclpKnobs &sclp_knobs() {
    static const SclpKnobs k;
    return k;
}
#endif

static SclpConfig sclp_config(const pc_plan *plan, int L, int64_t B) {
    SclpConfig c{};
    const int n = plan->n;
    c.ok = false;
    if (plan->q != 2 || n < 1 || n > 13 || L < 1 || L > 32) return c;
    const SclpKnobs kn = sclp_knobs();
    int G = 1, gsh = 0;
    while (G < L) G <<= 1, ++gsh;
    c.G = G;
    c.gsh = gsh;
    int rgl = kn.rgl < 1 ? 1 : kn.rgl;
    if (rgl > n + 1) rgl = n + 1;
    c.rgl = rgl;
    const int target = kn.warps_per_sm > 0 ? kn.warps_per_sm : SCLP_DEFAULT_WARPS_PER_SM;
    const size_t budget = (size_t)(227 * 1024 - 1024) / (size_t)target;  // 1 KB per CTA is the system's
    int lsm = n - 1 < 6 ? n - 1 : 6;
    if (lsm < 1) lsm = 1;
    int nst = kn.nst;
    nst = nst >= 8 ? 8 : nst >= 4 ? 4 : nst >= 2 ? 2 : 0;  // a power of two, or 0 = plain loads
    c.nst = nst;
    while (lsm > 1 && sclp_smem_bytes(n, lsm, rgl, nst) > budget) --lsm;
    if (kn.lsm >= 1 && kn.lsm <= 7) lsm = kn.lsm;
    c.lsm = lsm;
    c.smem = sclp_smem_bytes(n, lsm, rgl, nst);
    if (c.smem > 220 * 1024) return c;
    int per_sm = (int)((227 * 1024 - 1024) / c.smem);
    if (per_sm > target) per_sm = target;
    if (per_sm > 32) per_sm = 32;
    if (per_sm < 1) per_sm = 1;
    c.per_sm = per_sm;
    const int fpw = 32 / G;
    int64_t waves = (B + fpw - 1) / fpw;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > waves) grid = (waves + per_sm - 1) / per_sm * per_sm;  // whole CTAs of per_sm warps
    c.grid = (int)(grid > 0 ? grid : per_sm);
    const int64_t vrows_all = ((int64_t)1 << n) - 1, vrows_s = ((int64_t)2 << lsm) - 1;
    c.vg_stride = (size_t)(vrows_all > vrows_s ? vrows_all - vrows_s : 0) * 32 + 32 * (size_t)(kn.skew > 0 ? kn.skew : 1);
    c.rg_stride = (size_t)2 * (scl2_wsum(n + 1) - scl2_wsum(rgl)) * 32 + 32;
    c.lsmA = n - 1 < 7 ? (n - 1 < 0 ? 0 : n - 1) : 7;
    c.smem2 = sclp_final_smem_bytes(n, c.lsmA);
    int64_t grid2 = (int64_t)num_sms() * 32;  // warps of the final kernel (each with its own replay scratch)
    if (grid2 > B) grid2 = (B + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS * SCLP_AUX_WARPS;
    c.grid2 = (int)(grid2 > 0 ? grid2 : SCLP_AUX_WARPS);
    const int64_t arows = ((int64_t)1 << n) - 1, arows_s = ((int64_t)2 << c.lsmA) - 1;
    c.va_stride = (size_t)(arows > arows_s ? arows - arows_s : 0) + 2;
    c.ok = true;
    return c;
}

bool sclp_supported(const pc_plan *plan, int L) {
    if (sclp_knobs().off) return false;
    return sclp_config(plan, L, 1).ok;
}

int64_t sclp_wave_frames(const pc_plan *plan, int L) {
    const SclpConfig c = sclp_config(plan, L, (int64_t)1 << 40);
    return (int64_t)c.grid * (32 / c.G);
}

static const int64_t SCLP_CHUNK = 1 << 16;  // frames per prep / decode / final round (bounds the per-frame buffers)

struct SclpLayout {
    size_t off_vg, off_rg, off_va, off_tab, off_acw, off_ab, off_fb, off_sel, off_found, off_lsize, off_lprob, off_mxs;
    size_t off_ai, off_fv, off_info, off_linfo;  // packed staging of the byte-per-symbol ABI
    size_t total;
};

static SclpLayout sclp_layout(const pc_plan *plan, const SclTables *T, int L, int64_t B, bool want_list, bool bytes_abi) {
    const SclpConfig c = sclp_config(plan, L, B);
    const int64_t F = B < SCLP_CHUNK ? B : SCLP_CHUNK;
    const int N = plan->N, NW = N >= 32 ? N >> 5 : 1, k = plan->k, kw = (k + 31) / 32, nfw = (N - k + 31) / 32;
    SclpLayout y{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o = align256(o + bytes);
        return at;
    };
    y.off_vg = take((size_t)c.grid * c.vg_stride * sizeof(double2));
    y.off_rg = take((size_t)c.grid * c.rg_stride * 4);
    y.off_va = take((size_t)c.grid2 * c.va_stride * sizeof(double2));
    y.off_tab = take(256 * sizeof(double2));
    y.off_acw = take((size_t)F * NW * 4);
    y.off_ab = take((size_t)F * NW * 4);
    y.off_fb = take((size_t)F * NW * 4);
    y.off_sel = take((size_t)F * NW * 4 * (want_list ? L : 1));
    y.off_found = take((size_t)F * 4);
    y.off_lsize = take((size_t)F * 4);
    y.off_lprob = take((size_t)F * L * 8);
    y.off_mxs = take((size_t)F * (T ? T->n_leaf : N) * 8);
    if (bytes_abi) {
        y.off_ai = take((size_t)F * (kw ? kw : 1) * 4);
        y.off_fv = take((size_t)F * (nfw ? nfw : 1) * 4);
        y.off_info = take((size_t)F * (kw ? kw : 1) * 4);
        y.off_linfo = take(want_list ? (size_t)F * L * (kw ? kw : 1) * 4 : 4);
    }
    y.total = o;
    return y;
}

size_t sclp_workspace_bytes(const pc_plan *plan, int L, int64_t B, bool want_list) {
    SclTables *T = scl_tables(plan);
    return sclp_layout(plan, T, L, B, want_list, true).total;
}

template <int MB>
static int sclp_launch(const SclpParams &p, int ctas, size_t smem, cudaStream_t st) {
    PC_CUDA(cudaFuncSetAttribute(sclp_kernel<MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_LAUNCH(sclp_kernel<MB>, ctas, 32 * p.wpc, smem, st, p);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

// The decoder proper on packed inputs.  d_xy (float64 pairs) or d_y + d_tab (symbols); d_fvp may be null (all-zero frozen
// values); outputs packed information [B][kw], ProbResult [B] and the optional final-list outputs.
struct SclpIo {
    const double *d_xy;
    const uint8_t *d_y;
    const double *h_table;
    int Y;
    const uint32_t *d_fvp, *d_ainfo;
    uint32_t *d_info;
    int32_t *d_res, *d_lsize;
    double *d_lprob, *d_aprob;
    uint32_t *d_linfo;
};

static int sclp_run(const pc_plan *plan, const SclTables *T, int L, const SclpIo &io, int64_t B, char *base, const SclpLayout &y,
                    bool want_list, cudaStream_t st) {
    const SclpConfig c = sclp_config(plan, L, B);
    const int N = plan->N, NW = N >= 32 ? N >> 5 : 1, k = plan->k, kw = (k + 31) / 32, nfw = (N - k + 31) / 32;
    double2 *d_tab = (double2 *)(base + y.off_tab);
    if (io.d_y) {
        // the channel table travels as a kernel argument (no host buffer has to outlive the call, no synchronisation)
        SclpTab tb;
        for (int i = 0; i < 256; ++i) {
            const int r = i < io.Y ? i : io.Y - 1;  // symbols beyond the table read its last row (the ABI requires y < Y)
            tb.v[i] = make_double2(io.h_table[2 * r], io.h_table[2 * r + 1]);
        }
        PC_LAUNCH(sclp_tab_kernel, 1, 256, 0, st, tb, d_tab);
        PC_LAUNCH_CHECK();
    }
    for (int64_t f0 = 0; f0 < B; f0 += SCLP_CHUNK) {
        const int64_t F = B - f0 < SCLP_CHUNK ? B - f0 : SCLP_CHUNK;
        SclpPrepParams q{};
        q.n = plan->n, q.k = k, q.NW = NW, q.nfw = nfw, q.kw = kw;
        q.frames = F;
        q.a_src = T->d_a_src, q.f_src = T->d_f_src, q.perm = T->d_perm, q.stage_mask = T->d_stage_mask;
        q.ainfo = io.d_ainfo + f0 * kw;
        q.fvp = io.d_fvp ? io.d_fvp + f0 * nfw : nullptr;
        q.Acw = (uint32_t *)(base + y.off_acw);
        q.Ab = (uint32_t *)(base + y.off_ab);
        q.Fb = io.d_fvp ? (uint32_t *)(base + y.off_fb) : nullptr;
        int64_t g0 = (int64_t)num_sms() * 16;
        if (g0 > (F + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS) g0 = (F + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS;
        PC_LAUNCH(sclp_prep_kernel, (int)g0, 32 * SCLP_AUX_WARPS, (size_t)SCLP_AUX_WARPS * (3 * NW + 4) * 4, st, q);
        PC_LAUNCH_CHECK();

        SclpParams p{};
        p.n = plan->n, p.k = k, p.L = L, p.G = c.G, p.gsh = c.gsh, p.lsm = c.lsm, p.rgl = c.rgl, p.NW = NW;
        p.nst = c.nst;
        p.n_ops = (int)T->opsP.size(), p.n_leaf = T->n_leaf, p.sym = io.d_y ? 1 : 0, p.want_list = want_list ? 1 : 0, p.nfw = nfw;
        p.frames = F;
        p.ops = T->d_opsP;
        p.coef_words = T->d_rep_coef_words;
        p.xy = io.d_xy ? (const double2 *)io.d_xy + f0 * N : nullptr;
        p.y = io.d_y ? io.d_y + f0 * N : nullptr;
        p.tab = d_tab;
        p.fvp = q.fvp;
        p.Fb = q.Fb;
        p.Acw = q.Acw;
        p.vg = (double2 *)(base + y.off_vg);
        p.vg_stride = (int64_t)c.vg_stride;
        p.rg = (uint32_t *)(base + y.off_rg);
        p.rg_stride = (int64_t)c.rg_stride;
        p.selcw = (uint32_t *)(base + y.off_sel);
        p.found = (int32_t *)(base + y.off_found);
        p.lsize = (int32_t *)(base + y.off_lsize);
        p.lprob = (double *)(base + y.off_lprob);
        p.mxs = (double *)(base + y.off_mxs);
        const int fpw = 32 / c.G;
        int64_t grid = ((F + fpw - 1) / fpw + c.per_sm - 1) / c.per_sm * c.per_sm;
        if (grid > c.grid) grid = c.grid;
        static unsigned long long *d_timing = nullptr;
#ifndef PC_EMU
        if (sclp_knobs().timing && !d_timing) {
            cudaMalloc((void **)&d_timing, 128 * 8);
            cudaMemset(d_timing, 0, 128 * 8);
        }
#endif
        p.timing = d_timing;
        prof_mark(st);
        {
            // grid warps in CTAs of per_sm warps (one CTA per SM)
            p.alias = sclp_knobs().alias;
            p.l2lvl = sclp_knobs().l2lvl;
            p.wpc = c.per_sm, p.sync = sclp_knobs().sync, p.smem_per_warp = (int)c.smem;
            const int ctas = (int)((grid + c.per_sm - 1) / c.per_sm);
            const int rc = c.per_sm <= 12 ? sclp_launch<12>(p, ctas, c.smem * (size_t)c.per_sm, st)
                                          : sclp_launch<16>(p, ctas, c.smem * (size_t)c.per_sm, st);
            if (rc) return rc;
        }
        prof_mark(st);
#ifndef PC_EMU
        if (d_timing) {  // tuning runs only: cycles per (op kind, level) of the launch just made, summed over warps
            unsigned long long h[128];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d_timing, sizeof h, cudaMemcpyDeviceToHost);
            cudaMemset(d_timing, 0, sizeof h);
            static const char *names[8] = {"MINUS", "PLUS", "COMBINE", "RATE0", "REP", "RATE1", "SPC", "?"};
            unsigned long long tot = 0;
            for (int i = 0; i < 128; ++i) tot += h[i];
            fprintf(stderr, "SCLP_TIMING total warp-cycles %llu over %d warps\n", tot, (int)grid);
            for (int kk = 0; kk < 7; ++kk) {
                unsigned long long ks = 0;
                for (int l2 = 0; l2 < 16; ++l2) ks += h[kk * 16 + l2];
                fprintf(stderr, "SCLP_TIMING %-8s %5.1f%% :", names[kk], 100.0 * ks / (tot ? tot : 1));
                for (int l2 = 0; l2 < 16; ++l2)
                    if (h[kk * 16 + l2]) fprintf(stderr, " l%d=%.1f%%", l2, 100.0 * h[kk * 16 + l2] / (tot ? tot : 1));
                fprintf(stderr, "\n");
            }
        }
#endif

        SclpFinalParams r{};
        r.n = plan->n, r.k = k, r.L = L, r.NW = NW, r.kw = kw, r.n_ops = p.n_ops, r.n_leaf = T->n_leaf, r.sym = p.sym;
        r.want_list = p.want_list, r.lsmA = c.lsmA;
        r.frames = F;
        r.ops = T->d_opsP;
        r.info_src = T->d_info_src;
        r.xy = p.xy, r.y = p.y, r.tab = d_tab;
        r.Ab = q.Ab, r.selcw = p.selcw, r.found = p.found, r.lsize = p.lsize, r.lprob = p.lprob, r.mxs = p.mxs;
        r.va = (double2 *)(base + y.off_va);
        r.va_stride = (int64_t)c.va_stride;
        r.info = io.d_info + f0 * kw;
        r.result = io.d_res + f0;
        if (want_list) {
            r.out_lsize = io.d_lsize + f0;
            r.out_lprob = io.d_lprob + f0 * L;
            r.out_aprob = io.d_aprob + f0;
            r.out_linfo = io.d_linfo ? io.d_linfo + f0 * L * kw : nullptr;
        }
        int64_t g2 = (c.grid2 + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS;  // c.grid2 warps (the va scratch is sized per warp)
        if (g2 > (F + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS) g2 = (F + SCLP_AUX_WARPS - 1) / SCLP_AUX_WARPS;
        r.smem_per_warp = (int)c.smem2;
        PC_CUDA(cudaFuncSetAttribute(sclp_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(c.smem2 * SCLP_AUX_WARPS)));
        PC_LAUNCH(sclp_final_kernel, (int)g2, 32 * SCLP_AUX_WARPS, c.smem2 * SCLP_AUX_WARPS, st, r);
        PC_LAUNCH_CHECK();
    }
    return PC_OK;
}

static int pack_rows(const uint8_t *in, uint32_t *out, int64_t rows, int nbits, cudaStream_t st) {
    if (rows == 0 || nbits == 0) return PC_OK;
    const int W = (nbits + 31) / 32;
    int64_t blocks = (rows * W * 32 + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
    PC_LAUNCH(pack_rows_kernel, (int)blocks, 256, 0, st, in, out, rows, nbits, W);
    PC_LAUNCH_CHECK();
    return PC_OK;
}
static int unpack_rows(const uint32_t *in, uint8_t *out, int64_t rows, int nbits, cudaStream_t st) {
    if (rows == 0 || nbits == 0) return PC_OK;
    const int W = (nbits + 31) / 32;
    int64_t blocks = (rows * nbits + 255) / 256;
    if (blocks > (int64_t)num_sms() * 16) blocks = (int64_t)num_sms() * 16;
    PC_LAUNCH(unpack_rows_kernel, (int)blocks, 256, 0, st, in, out, rows, nbits, W);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

// pc_scl_decode_probs (byte per symbol) on this decoder: pack, decode, unpack -- chunk by chunk
int sclp_decode_bytes(const pc_plan *plan, const SclTables *T, int L, const double *d_xy, const uint8_t *d_fv, const uint8_t *d_ainfo,
                      int64_t B, uint8_t *d_info, int32_t *d_res, int32_t *d_lsize, double *d_lprob, double *d_aprob,
                      uint8_t *d_linfo, void *ws, size_t ws_bytes, cudaStream_t st) {
    const bool want_list = d_lsize != nullptr;
    const SclpLayout y = sclp_layout(plan, T, L, B, want_list, true);
    if (y.total > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, y.total);
        return PC_ERR_NOMEM;
    }
    char *base = (char *)ws;
    const int N = plan->N, k = plan->k, kw = (k + 31) / 32, nf = N - k;
    for (int64_t f0 = 0; f0 < B; f0 += SCLP_CHUNK) {
        const int64_t F = B - f0 < SCLP_CHUNK ? B - f0 : SCLP_CHUNK;
        SclpIo io{};
        io.d_xy = d_xy + f0 * N * 2;
        uint32_t *ai = (uint32_t *)(base + y.off_ai), *fv = (uint32_t *)(base + y.off_fv);
        int rc = pack_rows(d_ainfo + f0 * k, ai, F, k, st);
        if (rc) return rc;
        if (d_fv && nf > 0) {
            rc = pack_rows(d_fv + f0 * nf, fv, F, nf, st);
            if (rc) return rc;
            io.d_fvp = fv;
        }
        io.d_ainfo = ai;
        io.d_info = (uint32_t *)(base + y.off_info);
        io.d_res = d_res + f0;
        if (want_list) {
            io.d_lsize = d_lsize + f0;
            io.d_lprob = d_lprob + f0 * L;
            io.d_aprob = d_aprob + f0;
            io.d_linfo = d_linfo ? (uint32_t *)(base + y.off_linfo) : nullptr;
        }
        rc = sclp_run(plan, T, L, io, F, base, y, want_list, st);
        if (rc) return rc;
        if (k > 0) {
            rc = unpack_rows(io.d_info, d_info + f0 * k, F, k, st);
            if (rc) return rc;
            if (io.d_linfo) {
                rc = unpack_rows(io.d_linfo, d_linfo + f0 * L * k, F * L, k, st);
                if (rc) return rc;
            }
        }
    }
    (void)kw;
    return PC_OK;
}

size_t sclp_workspace_bytes_packed(const pc_plan *plan, int L, int64_t B, bool want_list) {
    SclTables *T = scl_tables(plan);
    return sclp_layout(plan, T, L, B, want_list, false).total;
}

int sclp_decode_packed(const pc_plan *plan, const SclTables *T, int L, const double *d_xy, const uint8_t *d_y, const double *h_table,
                       int Y, const uint32_t *d_fvp, const uint32_t *d_ainfo, int64_t B, uint32_t *d_info, int32_t *d_res,
                       int32_t *d_lsize, double *d_lprob, double *d_aprob, uint32_t *d_linfo, void *ws, size_t ws_bytes,
                       cudaStream_t st) {
    const bool want_list = d_lsize != nullptr;
    const SclpLayout y = sclp_layout(plan, T, L, B, want_list, false);
    if (y.total > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, y.total);
        return PC_ERR_NOMEM;
    }
    SclpIo io{};
    io.d_xy = d_xy, io.d_y = d_y, io.h_table = h_table, io.Y = Y;
    io.d_fvp = d_fvp, io.d_ainfo = d_ainfo, io.d_info = d_info, io.d_res = d_res;
    io.d_lsize = d_lsize, io.d_lprob = d_lprob, io.d_aprob = d_aprob, io.d_linfo = d_linfo;
    return sclp_run(plan, T, L, io, B, (char *)ws, y, want_list, st);
}

}  // namespace pc
