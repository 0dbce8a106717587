// trellis.cu -- SC decoding over the deletion channel: BinaryPolarEncoderDecoder.decode with a
// CollectionOfBinaryTrellises as xyVectorDistribution (BinaryPolarEncoderDecoder.py:71-99, :223-325;
// VectorDistributions/BinaryTrellis.py:206-438; VectorDistributions/CollectionOfBinaryTrellises.py:55-129), uniform prior.
//
// Structure of the path: the top n0 levels of the decoding tree operate on 2^(n-n0) independent trellises of length 2^n0
// (one per guard-band-delimited sub-word); after n0 transforms each trellis has collapsed to one probability pair and the
// remaining n-n0 levels are ordinary memoryless SC decoding of 2^n0 sub-blocks of length T = 2^(n-n0).  Here:
//  * the top tree (2^(n0+1) - 2 trellis transforms, 2^n0 sub-block decodes, 2^n0 - 1 combines) is walked by the host side
//    of this library with batch-wide launches -- every frame of the batch executes the same sequence;
//  * trellis kernels map one (frame, trellis) per thread.  The reference keeps vertices and edges in Python dicts and
//    sums edge products in dict INSERTION order; float64 results depend on that order, so each thread keeps the same
//    insertion-ordered vertex / edge lists in a small fixed-capacity record (capacities follow from the sub-word length
//    and the level) and takes every sum in the reference's order: bit-identical probabilities;
//  * trellis construction from the trimmed sub-word (buildTrellis_uniformInput_deletion) is fused into the first kernel;
//  * the collapsed pairs are max-normalised (BinaryMemorylessVectorDistribution.normalize) and handed to the frame-per-lane
//    SC decoder of sc_binary.cu with a per-sub-block plan; partial sums come back bit-packed and are combined with the
//    16 -> 32 bit interleave (x[2h] = m[h] ^ p[h], x[2h+1] = p[h], BinaryPolarEncoderDecoder.py:321-323).
#include <cmath>
#include <map>
#include <mutex>

#include "common.cuh"

namespace pc {

// record layout of one trellis with LEN edge layers, VC vertex slots per layer, LC list entries per vertex
struct TLayout {
    int LEN, VC, LC, EC;
    int off_nv, off_ne, off_vpos, off_nout, off_nin, off_efrom, off_eto, off_elab, off_out, off_in, off_vprob, off_ep, bytes;
};

static TLayout t_layout(int LEN, int VC, int D) {
    TLayout L;
    L.LEN = LEN;
    L.VC = VC;
    L.LC = 2 * (D + 1);
    L.EC = VC * L.LC;
    int o = 0;
    auto take = [&](int bytes, int align) {
        o = (o + align - 1) / align * align;
        const int at = o;
        o += bytes;
        return at;
    };
    L.off_vprob = take((LEN + 1) * VC * 8, 8);
    L.off_ep = take(LEN * L.EC * 8, 8);
    L.off_nv = take((LEN + 1) * 4, 4);
    L.off_ne = take(LEN * 4, 4);
    L.off_out = take((LEN + 1) * VC * L.LC * 2, 2);
    L.off_in = take((LEN + 1) * VC * L.LC * 2, 2);
    L.off_vpos = take((LEN + 1) * VC, 1);
    L.off_nout = take((LEN + 1) * VC, 1);
    L.off_nin = take((LEN + 1) * VC, 1);
    L.off_efrom = take(LEN * L.EC, 1);
    L.off_eto = take(LEN * L.EC, 1);
    L.off_elab = take(LEN * L.EC, 1);
    L.bytes = (o + 15) / 16 * 16;
    return L;
}

// Records are stored field-major across the launch's records (structure of arrays): element `idx` of a field of record
// `tid` lives at base + field_offset * nt + (idx * nt + tid) * sizeof(T), so the threads of a warp -- which walk trellises
// of the same shape in near lock step -- touch consecutive addresses.
struct TRef {
    char *base;
    int64_t nt, tid;  // records in this launch, this thread's record
    TLayout L;
    template <class T>
    __device__ __forceinline__ T &at(int off, int idx) const {
        return *(T *)(base + (int64_t)off * nt + ((int64_t)idx * nt + tid) * (int64_t)sizeof(T));
    }
    __device__ int &nv(int l) const { return at<int>(L.off_nv, l); }
    __device__ int &ne(int l) const { return at<int>(L.off_ne, l); }
    __device__ uint8_t &vpos(int l, int k) const { return at<uint8_t>(L.off_vpos, l * L.VC + k); }
    __device__ double &vprob(int l, int k) const { return at<double>(L.off_vprob, l * L.VC + k); }
    __device__ uint8_t &nout(int l, int k) const { return at<uint8_t>(L.off_nout, l * L.VC + k); }
    __device__ uint8_t &nin(int l, int k) const { return at<uint8_t>(L.off_nin, l * L.VC + k); }
    __device__ uint16_t &out(int l, int k, int a) const { return at<uint16_t>(L.off_out, (l * L.VC + k) * L.LC + a); }
    __device__ uint16_t &in(int l, int k, int a) const { return at<uint16_t>(L.off_in, (l * L.VC + k) * L.LC + a); }
    __device__ uint8_t &efrom(int l, int e) const { return at<uint8_t>(L.off_efrom, l * L.EC + e); }
    __device__ uint8_t &eto(int l, int e) const { return at<uint8_t>(L.off_eto, l * L.EC + e); }
    __device__ uint8_t &elab(int l, int e) const { return at<uint8_t>(L.off_elab, l * L.EC + e); }
    __device__ double &ep(int l, int e) const { return at<double>(L.off_ep, l * L.EC + e); }
};

__device__ void t_clear(const TRef &t) {
    for (int l = 0; l <= t.L.LEN; ++l) t.nv(l) = 0;
    for (int l = 0; l < t.L.LEN; ++l) t.ne(l) = 0;
}

// __getVertexAndAddIfNeeded (BinaryTrellis.py:154-161): index of the vertex in its layer's insertion-ordered list
__device__ int t_vertex(const TRef &t, int layer, int vp) {
    const int n = t.nv(layer);
    for (int k = 0; k < n; ++k)
        if (t.vpos(layer, k) == vp) return k;
    if (n >= t.L.VC) return t.L.VC - 1;  // cannot happen: VC covers every vertical position of the sub-word
    t.vpos(layer, n) = (uint8_t)vp;
    t.vprob(layer, n) = -1.0;
    t.nout(layer, n) = 0;
    t.nin(layer, n) = 0;
    t.nv(layer) = n + 1;
    return n;
}

__device__ void t_set_vertex_prob(const TRef &t, int layer, int vp, double p) { t.vprob(layer, t_vertex(t, layer, vp)) = p; }

// addToEdgeProb + __getEdgeAndAddIfNeeded (BinaryTrellis.py:128-136, :163-175)
__device__ void t_add_edge_prob(const TRef &t, int layer, int from_vp, int to_vp, int label, double p) {
    const int fk = t_vertex(t, layer, from_vp);
    const int tk = t_vertex(t, layer + 1, to_vp);
    int id = -1;
    const int no = t.nout(layer, fk);
    for (int a = 0; a < no; ++a) {
        const int e = t.out(layer, fk, a);
        if (t.eto(layer, e) == tk && t.elab(layer, e) == label) {
            id = e;
            break;
        }
    }
    if (id < 0) {
        id = t.ne(layer);
        const int ni = t.nin(layer + 1, tk);
        if (id >= t.L.EC || no >= t.L.LC || ni >= t.L.LC) return;  // cannot happen (capacities cover the structure)
        t.ne(layer) = id + 1;
        t.efrom(layer, id) = (uint8_t)fk;
        t.eto(layer, id) = (uint8_t)tk;
        t.elab(layer, id) = (uint8_t)label;
        t.ep(layer, id) = 0.0;
        t.out(layer, fk, no) = (uint16_t)id;
        t.nout(layer, fk) = (uint8_t)(no + 1);
        t.in(layer + 1, tk, ni) = (uint16_t)id;
        t.nin(layer + 1, tk) = (uint8_t)(ni + 1);
    }
    t.ep(layer, id) = __dadd_rn(t.ep(layer, id), p);
}

// calcNormalizationVector + normalize (BinaryTrellis.py:280-306)
__device__ void t_normalize(const TRef &t) {
    for (int i = 0; i < t.L.LEN; ++i) {
        double tp0 = 0.0, tp1 = 0.0;
        const int n = t.nv(i);
        for (int k = 0; k < n; ++k) {
            const int no = t.nout(i, k);
            for (int a = 0; a < no; ++a) {
                const int e = t.out(i, k, a);
                if (t.elab(i, e))
                    tp1 = __dadd_rn(tp1, t.ep(i, e));
                else
                    tp0 = __dadd_rn(tp0, t.ep(i, e));
            }
        }
        double nrm = tp0 > tp1 ? tp0 : tp1;
        if (nrm == 0.0) nrm = 1.0;
        const int ne = t.ne(i);
        for (int e = 0; e < ne; ++e) t.ep(i, e) = t.ep(i, e) / nrm;
    }
}

struct TrellisParams {
    int n0, T, maxlen, ones, trimmed;
    int64_t frames;
    double d;                 // deletion probability
    double pw1[17], pw0[17];  // ones > 0: (1-d)^i and d^i, i <= ones (computed on the host with libm, like the reference)
    double comb[17];
    const uint8_t *sub_bits;  // [frames][T][maxlen]
    const int32_t *sub_len;   // [frames][T]
    char *blob_in, *blob_out;
    TLayout Lin, Lout;
    const uint32_t *decision;  // plus transform: the minus child's codeword, [frames][dec_words]; null: minus transform
    int dec_words, sub;        // decision bits per trellis
    double *probs_out;         // collapse: [frames][T][2] max-normalised pairs (input of the sub-block SC decoder)
    double *raw_out;           // optional: the unnormalised collapsed pairs (tests)
    const int32_t *perm;       // [frames * T] record handled by each thread: (frame, trellis) pairs bucketed by sub-word length
};

// buildTrellis_uniformInput_deletion (BinaryTrellis.py:309-438)
__device__ void t_build(const TRef &t, const uint8_t *rw, int rlen, const TrellisParams &p) {
    const int L = t.L.LEN, ones = p.ones;
    const double d = p.d;
    t_clear(t);
    const int deletionCount = L + 2 * ones - rlen;
    if (ones > 0) {
        const int m = ones < rlen ? ones : rlen;
        for (int i = 0; i < 1 + m; ++i) t_set_vertex_prob(t, 0, i, __dmul_rn(__dmul_rn(p.comb[i], p.pw1[i]), p.pw0[ones - i]));
        for (int i = rlen; i > rlen - m - 1; --i) {
            const int j = rlen - i;
            t_set_vertex_prob(t, L, i, __dmul_rn(__dmul_rn(p.comb[j], p.pw1[j]), p.pw0[ones - j]));
        }
    } else {
        t_set_vertex_prob(t, 0, 0, 1.0);
        t_set_vertex_prob(t, L, rlen, 1.0);
    }
    const double pnd = __dmul_rn(0.5, __dadd_rn(1.0, -d)), pdel = __dmul_rn(0.5, d);
    for (int l = 0; l < L; ++l) {
        int vmin, vmax;
        if (ones > 0) {
            vmin = l + ones - deletionCount > 0 ? l + ones - deletionCount : 0;
            vmax = l + ones < rlen ? l + ones : rlen;
        } else {
            vmin = l - deletionCount > 0 ? l - deletionCount : 0;
            vmax = l < rlen ? l : rlen;
        }
        for (int vp = vmin; vp <= vmax; ++vp) {
            if (vp < rlen) t_add_edge_prob(t, l, vp, vp + 1, rw[vp] & 1, pnd);
            if (l + 1 + ones - deletionCount <= vp) {
                for (int label = 0; label < 2; ++label) {
                    const double pr = (!p.trimmed || label == 1 || (vp > 0 && vp < rlen)) ? pdel : 0.5;
                    t_add_edge_prob(t, l, vp, vp, label, pr);
                }
            }
        }
    }
}

// __miusPlusTransform (BinaryTrellis.py:206-258): `in` of length LEN -> `out` of length LEN/2
__device__ void t_transform(const TRef &in, const TRef &out, bool plus, uint32_t dec_bits) {
    const int LEN = in.L.LEN;
    t_clear(out);
    for (int k = 0; k < in.nv(0); ++k) t_set_vertex_prob(out, 0, in.vpos(0, k), in.vprob(0, k));
    for (int k = 0; k < in.nv(LEN); ++k) t_set_vertex_prob(out, LEN / 2, in.vpos(LEN, k), in.vprob(LEN, k));
    for (int ml = 1; ml <= LEN; ml += 2) {
        const int nw = in.nv(ml);
        for (int wk = 0; wk < nw; ++wk) {
            const int ni = in.nin(ml, wk), no = in.nout(ml, wk);
            for (int a = 0; a < ni; ++a) {
                const int ei = in.in(ml, wk, a);
                const int x0 = in.elab(ml - 1, ei);
                const double p0 = in.ep(ml - 1, ei);
                const int u_vp = in.vpos(ml - 1, in.efrom(ml - 1, ei));
                for (int b = 0; b < no; ++b) {
                    const int eo = in.out(ml, wk, b);
                    const int x1 = in.elab(ml, eo);
                    const int v_vp = in.vpos(ml + 1, in.eto(ml, eo));
                    const double np = __dmul_rn(p0, in.ep(ml, eo));
                    const int mlabel = x0 != x1 ? 1 : 0;
                    if (!plus) {
                        t_add_edge_prob(out, (ml - 1) / 2, u_vp, v_vp, mlabel, np);
                    } else {
                        if (mlabel != (int)((dec_bits >> ((ml - 1) / 2)) & 1u)) continue;
                        t_add_edge_prob(out, (ml - 1) / 2, u_vp, v_vp, x1, np);
                    }
                }
            }
        }
    }
}

// one (frame, trellis) per thread.  BUILD: construct the level-0 trellis into blob_in first.  The transform output is
// normalised (LEN_out >= 1 trellis levels) or, when it has length 1, collapsed to a probability pair.
template <bool BUILD>
__global__ void __launch_bounds__(128) trellis_step_kernel(const TrellisParams p) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= p.frames * p.T) return;
    // Thread gid works on record rec = perm[gid] and keeps that trellis in slot gid of the blobs in every launch of the walk.
    // The permutation lists the (frame, trellis) pairs by sub-word length (= number of deletions = the shape of the trellis:
    // vertices per layer, edges per vertex), so the threads of a warp run the same loops instead of waiting for the widest one.
    const int64_t rec = p.perm ? (int64_t)p.perm[gid] : gid;
    const int64_t f = rec / p.T;
    const int i = (int)(rec - f * p.T);
    const int64_t nt = p.frames * p.T;
    TRef in{p.blob_in, nt, gid, p.Lin};
    if (BUILD) {
        const int rlen = p.sub_len[rec];
        t_build(in, p.sub_bits + rec * p.maxlen, rlen < p.maxlen ? rlen : p.maxlen, p);
        if (!p.blob_out && !p.probs_out) return;
    }
    const bool plus = p.decision != nullptr;
    uint32_t dec = 0;
    if (plus) {
        const int bit0 = i * p.sub;  // `sub` is a power of two <= 16: the bits never straddle a word
        dec = (p.decision[f * p.dec_words + (bit0 >> 5)] >> (bit0 & 31)) & ((1u << p.sub) - 1u);
    }
    if (p.Lin.LEN > 2) {
        TRef out{p.blob_out, nt, gid, p.Lout};
        t_transform(in, out, plus, dec);
        t_normalize(out);
        return;
    }
    // LEN == 2: the child has length 1 -> calcMarginalizedProbabilities(normalize=False) (BinaryTrellis.py:260-278), then
    // BinaryMemorylessVectorDistribution.normalize (:79-87).  The child is built in the output record, then reduced.
    TRef out{p.blob_out, nt, gid, p.Lout};
    t_transform(in, out, plus, dec);
    double m0 = 0.0, m1 = 0.0;
    const int n = out.nv(0);
    for (int k = 0; k < n; ++k) {
        const int no = out.nout(0, k);
        for (int a = 0; a < no; ++a) {
            const int e = out.out(0, k, a);
            const double term = __dmul_rn(__dmul_rn(out.vprob(0, k), out.ep(0, e)), out.vprob(1, out.eto(0, e))) / 1.0;
            if (out.elab(0, e))
                m1 = __dadd_rn(m1, term);
            else
                m0 = __dadd_rn(m0, term);
        }
    }
    if (p.raw_out) {
        p.raw_out[rec * 2] = m0;
        p.raw_out[rec * 2 + 1] = m1;
    }
    double nrm = m0 > m1 ? m0 : m1;
    if (nrm == 0.0) nrm = 1.0;
    p.probs_out[rec * 2] = m0 / nrm;
    p.probs_out[rec * 2 + 1] = m1 / nrm;
}

// ---- bucketing of the records by sub-word length: a counting sort (keys 0 .. 255) -------------------------------------------
__global__ void __launch_bounds__(256) tr_hist_kernel(const int32_t *__restrict__ len, int64_t nt, int maxlen, int32_t *hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (int64_t)gridDim.x * blockDim.x) {
        int k = len[i];
        k = k < 0 ? 0 : (k > maxlen ? maxlen : k);
        atomicAdd(hist + k, 1);
    }
}
__global__ void __launch_bounds__(256) tr_scan_kernel(int32_t *hist) {  // exclusive prefix sums of 256 bins, in place
    __shared__ int32_t s[256];
    s[threadIdx.x] = hist[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        int32_t acc = 0;
        for (int i = 0; i < 256; ++i) {
            const int32_t v = s[i];
            s[i] = acc;
            acc += v;
        }
    }
    __syncthreads();
    hist[threadIdx.x] = s[threadIdx.x];
}
__global__ void __launch_bounds__(256) tr_scatter_kernel(const int32_t *__restrict__ len, int64_t nt, int maxlen, int32_t *cursor,
                                                         int32_t *perm) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += (int64_t)gridDim.x * blockDim.x) {
        int k = len[i];
        k = k < 0 ? 0 : (k > maxlen ? maxlen : k);
        perm[atomicAdd(cursor + k, 1)] = (int32_t)i;  // the order inside a bucket does not matter: records are independent
    }
}

__device__ __forceinline__ uint32_t tr_spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// out[f][w] = interleave of the minus / plus children's codewords (BinaryPolarEncoderDecoder.py:321-323)
__global__ void combine_words_kernel(int64_t frames, int Wc, int Wo, const uint32_t *__restrict__ m, const uint32_t *__restrict__ pp,
                                     uint32_t *__restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= frames * Wo) return;
    const int64_t f = gid / Wo;
    const int w = (int)(gid - f * Wo);
    const int sh = (w & 1) * 16;
    const uint32_t m16 = (m[f * Wc + (w >> 1)] >> sh) & 0xffffu, p16 = (pp[f * Wc + (w >> 1)] >> sh) & 0xffffu;
    out[gid] = tr_spread16(m16 ^ p16) | (tr_spread16(p16) << 1);
}

// information bits of the sub-blocks, concatenated in u order: out bit q of frame f comes from sub-block j, bit q - koff[j]
struct MergeParams {
    int nsub, k, Kw;
    int koff[17];
    int kw[16];
    const uint32_t *src[16];
};
__global__ void merge_info_kernel(int64_t frames, const MergeParams p, uint32_t *__restrict__ out) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one warp per output word
    const int64_t wid = gid >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= frames * p.Kw) return;
    const int64_t f = wid / p.Kw;
    const int w = (int)(wid - f * p.Kw);
    const int q = 32 * w + lane;
    uint32_t bit = 0;
    if (q < p.k) {
        int j = 0;
        while (j + 1 < p.nsub && q >= p.koff[j + 1]) ++j;
        const int b = q - p.koff[j];
        bit = (p.src[j][f * p.kw[j] + (b >> 5)] >> (b & 31)) & 1u;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) out[wid] = word;
}

// ---- host side: sub-block plans and the top-tree walk ----------------------------------------------------------------
struct TrellisTables {
    int n0 = 0;
    std::vector<pc_plan *> sub;  // one plan per sub-block of length T
};
static std::mutex g_tr_mu;
static std::map<std::pair<const pc_plan *, int>, TrellisTables *> g_tr_tables;

static TrellisTables *trellis_tables(const pc_plan *p, int n0) {
    std::lock_guard<std::mutex> lk(g_tr_mu);
    auto key = std::make_pair(p, n0);
    auto it = g_tr_tables.find(key);
    if (it != g_tr_tables.end()) return it->second;
    TrellisTables *T = new TrellisTables();
    T->n0 = n0;
    const int Tn = p->N >> n0;
    for (int j = 0; j < (1 << n0); ++j) {
        pc_plan *sp = nullptr;
        if (pc_plan_create(2, p->n - n0, p->frozen_mask.data() + (size_t)j * Tn, p->frozen_vals.data() + (size_t)j * Tn, &sp) != PC_OK) {
            for (pc_plan *q : T->sub) pc_plan_destroy(q);
            delete T;
            return nullptr;
        }
        T->sub.push_back(sp);
    }
    g_tr_tables[key] = T;
    return T;
}

void trellis_tables_release(const pc_plan *p) {
    std::vector<TrellisTables *> dead;
    {
        std::lock_guard<std::mutex> lk(g_tr_mu);
        for (auto it = g_tr_tables.begin(); it != g_tr_tables.end();) {
            if (it->first.first == p) {
                dead.push_back(it->second);
                it = g_tr_tables.erase(it);
            } else {
                ++it;
            }
        }
    }
    for (TrellisTables *T : dead) {
        for (pc_plan *q : T->sub) pc_plan_destroy(q);
        delete T;
    }
}

struct TrellisWs {
    int64_t chunk;
    std::vector<TLayout> lay;       // level k = 0 .. n0 (level n0: the length-1 children)
    std::vector<size_t> off_blob;   // per level
    std::vector<size_t> off_cw;     // per level k: two buffers (minus, plus child codewords of a level-k node) + result
    size_t off_probs, off_raw, off_info, off_sc, off_perm, off_hist, sc_bytes, total;
    std::vector<size_t> off_subinfo;
};

static int words_of(int bits) { return bits >= 32 ? bits >> 5 : 1; }

static TrellisWs trellis_ws(const pc_plan *plan, int n0, int maxlen, int64_t chunk) {
    TrellisWs W;
    W.chunk = chunk;
    const int Tn = plan->N >> n0, tl = 1 << n0, VC = maxlen + 1;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += align256(bytes + 256);
        return at;
    };
    for (int k = 0; k <= n0; ++k) {
        W.lay.push_back(t_layout(tl >> k > 0 ? tl >> k : 1, VC, 1 << k));
        W.off_blob.push_back(take((size_t)chunk * Tn * W.lay[k].bytes));
    }
    for (int k = 0; k <= n0; ++k) W.off_cw.push_back(take((size_t)2 * chunk * words_of(plan->N >> k) * 4));
    W.off_probs = take((size_t)chunk * Tn * 16);
    W.off_raw = take((size_t)chunk * Tn * 16);
    W.off_perm = take((size_t)chunk * Tn * 4);
    W.off_hist = take(256 * 4);
    W.off_info = 0;
    for (int j = 0; j < (1 << n0); ++j) W.off_subinfo.push_back(take((size_t)chunk * words_of(Tn) * 4));
    W.sc_bytes = 0;
    W.off_sc = o;
    W.total = o;
    return W;
}

// frames per batch-wide launch of the top-tree walk (read once: the workspace query and the decode must agree)
static int64_t trellis_chunk() {
    static const int64_t c = [] {
        const char *s = getenv("PC_TRELLIS_CHUNK");
        const int64_t v = s && *s ? atoll(s) : 32768;  // 8192 -> 32768: +15 % (fewer, longer launches; tails amortised)
        return v < 32 ? (int64_t)32 : v;
    }();
    return c;
}

extern "C" size_t pc_sc_workspace_bytes(const pc_plan *plan, int64_t B, int input_kind);
extern "C" int pc_sc_decode_probs(const pc_plan *plan, const double *d_xy, int64_t B, uint32_t *d_cw_packed, uint32_t *d_info_packed,
                                  void *d_workspace, size_t workspace_bytes, void *stream);

// sc_binary.cu: genie pass over a (sub-)block
size_t sc_genie_workspace_bytes(const pc_plan *plan, int64_t B);
int sc_genie_common(const pc_plan *plan, const double *d_xy, const uint32_t *d_u, int64_t u_pitch_words, int u_bit_off, int64_t B,
                    uint32_t *d_cw, double *d_marg, int64_t marg_pitch, int64_t marg_off, void *ws, size_t ws_bytes,
                    cudaStream_t st);

struct TrellisRun {
    const uint32_t *genie_u = nullptr;  // genie pass: known u bits [frames][Nw] and captured leaf probabilities [frames][N][2]
    double *genie_marg = nullptr;
    const pc_plan *plan;
    TrellisTables *tabs;
    TrellisWs W;
    TrellisParams base;
    char *ws;
    size_t sc_bytes;
    int64_t frames;
    cudaStream_t st;
    double *raw_first;  // optional capture of the first collapsed vector
    bool raw_done;
};

// decode the node `node` of level k (collection length N >> k, trellis length tl >> k); its codeword goes to `cw_out`
static int trellis_descend(TrellisRun &R, int k, int node, uint32_t *cw_out) {
    const pc_plan *plan = R.plan;
    const int n0 = R.tabs->n0, Tn = plan->N >> n0;
    const int64_t F = R.frames;
    const int child_bits = plan->N >> (k + 1);
    const int Wc = words_of(child_bits), Wo = words_of(plan->N >> k);
    // the two children's codewords live in the level-(k+1) pair of slots; deeper nodes use deeper levels' slots
    uint32_t *cwA = (uint32_t *)(R.ws + R.W.off_cw[k + 1]);
    uint32_t *cwB = cwA + (size_t)R.W.chunk * Wc;
    for (int pass = 0; pass < 2; ++pass) {
        TrellisParams p = R.base;
        p.frames = F;
        p.blob_in = R.ws + R.W.off_blob[k];
        p.Lin = R.W.lay[k];
        p.blob_out = R.ws + R.W.off_blob[k + 1];
        p.Lout = R.W.lay[k + 1];
        p.decision = pass == 0 ? nullptr : cwA;
        p.dec_words = Wc;
        p.sub = (1 << n0 >> k) / 2;
        p.probs_out = (double *)(R.ws + R.W.off_probs);
        p.raw_out = nullptr;
        const bool collapse = k + 1 == n0;
        if (collapse && R.raw_first && !R.raw_done) {
            p.raw_out = R.raw_first;
            R.raw_done = true;
        }
        const int64_t threads = F * Tn;
        const unsigned grid = (unsigned)((threads + 127) / 128);
        if (k == 0 && pass == 0)
            trellis_step_kernel<true><<<grid, 128, 0, R.st>>>(p);
        else
            trellis_step_kernel<false><<<grid, 128, 0, R.st>>>(p);
        PC_LAUNCH_CHECK();
        uint32_t *dst = pass == 0 ? cwA : cwB;
        const int child = 2 * node + pass;
        if (collapse) {
            const pc_plan *sp = R.tabs->sub[child];
            uint32_t *sinfo = (uint32_t *)(R.ws + R.W.off_subinfo[child]);
            int rc;
            if (R.genie_u)
                rc = sc_genie_common(sp, p.probs_out, R.genie_u, words_of(plan->N), child * Tn, F, dst, R.genie_marg,
                                     (int64_t)2 * plan->N, (int64_t)2 * child * Tn, R.ws + R.W.off_sc, R.sc_bytes, R.st);
            else
                rc = pc_sc_decode_probs(sp, p.probs_out, F, dst, sinfo, R.ws + R.W.off_sc, R.sc_bytes, R.st);
            if (rc) return rc;
        } else {
            const int rc = trellis_descend(R, k + 1, child, dst);
            if (rc) return rc;
        }
    }
    const int64_t items = F * Wo;
    combine_words_kernel<<<(unsigned)((items + 255) / 256), 256, 0, R.st>>>(F, Wc, Wo, cwA, cwB, cw_out);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // namespace pc

extern "C" {

static size_t trellis_sc_bytes(pc::TrellisTables *T, int64_t chunk) {
    size_t sc = 256;
    if (T)
        for (pc_plan *sp : T->sub) {
            size_t b = pc_sc_workspace_bytes(sp, chunk, PC_INPUT_PROBS);
            if (sp->n >= 1) {
                const size_t g = pc::sc_genie_workspace_bytes(sp, chunk);
                if (g > b) b = g;
            }
            if (b > sc) sc = b;
        }
    return sc;
}

size_t pc_trellis_workspace_bytes(const pc_plan *plan, int n0, int maxlen, int64_t B) {
    // the same limits as pc_trellis_decode, BEFORE any table is built (2^n0 sub-plans)
    if (!plan || plan->q != 2 || n0 < 1 || n0 > 4 || n0 > plan->n || plan->n - n0 > 16 || B <= 0 || maxlen < 1 || maxlen > 250) return 256;
    int64_t chunk = B < pc::trellis_chunk() ? B : pc::trellis_chunk();
    pc::TrellisWs W = pc::trellis_ws(plan, n0, maxlen, chunk);
    pc::TrellisTables *T = pc::trellis_tables(plan, n0);
    return W.total + pc::align256(trellis_sc_bytes(T, chunk) + 256);
}

static int trellis_common(const pc_plan *plan, int n0, double deletion_prob, int ones, const uint8_t *d_sub_bits,
                          const int32_t *d_sub_len, int maxlen, int64_t B, uint32_t *d_cw_packed, uint32_t *d_info_packed,
                          double *d_first_collapse, const uint32_t *d_genie_u, double *d_genie_marg, void *d_workspace,
                          size_t workspace_bytes, void *stream) {
    using namespace pc;
    const bool genie = d_genie_u != nullptr;
    PC_REQUIRE(plan && plan->q == 2, "binary plan required");
    PC_REQUIRE(n0 >= 1 && n0 <= 4 && n0 <= plan->n, "n0 must be in [1, min(4, n)]");
    PC_REQUIRE(plan->n - n0 <= 16, "sub-blocks longer than 2^16 are not supported");
    PC_REQUIRE(ones >= 0 && ones <= 16 && maxlen >= 1 && maxlen <= 250, "ones must be in [0,16], maxlen in [1,250]");
    PC_REQUIRE(deletion_prob >= 0.0 && deletion_prob <= 1.0, "deletion probability out of range");
    PC_REQUIRE(B >= 0, "negative batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_sub_bits && d_sub_len && d_cw_packed && (genie || d_info_packed || plan->k == 0) && d_workspace, "null buffer");
    if (genie) PC_REQUIRE(d_genie_marg && plan->n - n0 >= 1, "genie pass needs the marginal output and sub-blocks of at least 2 symbols");
    PC_REQUIRE(((uintptr_t)d_workspace & 255) == 0, "workspace must be 256-byte aligned");
    TrellisTables *T = trellis_tables(plan, n0);
    if (!T) return PC_ERR_CUDA;
    ProfScope prof_scope((cudaStream_t)stream);  // the whole top-tree walk is the measured unit
    int64_t chunk = B < pc::trellis_chunk() ? B : pc::trellis_chunk();
    const size_t sc = trellis_sc_bytes(T, chunk);
    TrellisWs W = trellis_ws(plan, n0, maxlen, chunk);
    if (W.total + align256(sc + 256) > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, W.total + align256(sc + 256));
        return PC_ERR_NOMEM;
    }
    const int Tn = plan->N >> n0, Nw = words_of(plan->N), Kw = (plan->k + 31) / 32;
    TrellisRun R{};
    R.plan = plan;
    R.tabs = T;
    R.W = W;
    R.ws = (char *)d_workspace;
    R.sc_bytes = align256(sc + 256);
    R.st = (cudaStream_t)stream;
    R.base = TrellisParams{};
    R.base.n0 = n0;
    R.base.T = Tn;
    R.base.maxlen = maxlen;
    R.base.ones = ones;
    R.base.trimmed = 1;  // buildCollectionOfBinaryTrellises_uniformInput_deletion always trims (CollectionOfBinaryTrellises.py:117)
    R.base.d = deletion_prob;
    for (int i = 0; i <= 16; ++i) {
        R.base.pw1[i] = std::pow(1.0 - deletion_prob, i);
        R.base.pw0[i] = std::pow(deletion_prob, i);
        double c = 1.0;
        for (int j = 1; j <= i && i <= ones; ++j) c = c * (double)(ones - i + j) / (double)j;
        R.base.comb[i] = i <= ones ? std::floor(c + 0.5) : 0.0;
    }
    // information offsets of the sub-blocks
    MergeParams M{};
    M.nsub = 1 << n0;
    M.k = plan->k;
    M.Kw = Kw;
    int acc = 0;
    for (int j = 0; j < M.nsub; ++j) {
        M.koff[j] = acc;
        acc += T->sub[j]->k;
        M.kw[j] = (T->sub[j]->k + 31) / 32;
        M.src[j] = (const uint32_t *)(R.ws + W.off_subinfo[j]);
    }
    M.koff[M.nsub] = acc;
    for (int64_t f0 = 0; f0 < B; f0 += chunk) {
        R.frames = (B - f0) < chunk ? (B - f0) : chunk;
        R.base.sub_bits = d_sub_bits + f0 * Tn * maxlen;
        R.base.sub_len = d_sub_len + f0 * Tn;
        R.raw_first = d_first_collapse ? d_first_collapse + f0 * Tn * 2 : nullptr;
        R.raw_done = false;
        {
            // bucket this chunk's (frame, trellis) records by sub-word length (PC_TRELLIS_BUCKET=0 keeps the natural order)
            static const bool bucket = !(getenv("PC_TRELLIS_BUCKET") && atoi(getenv("PC_TRELLIS_BUCKET")) == 0);
            R.base.perm = nullptr;
            if (bucket && maxlen <= 255) {
                const int64_t nt = R.frames * Tn;
                int32_t *hist = (int32_t *)(R.ws + W.off_hist), *perm = (int32_t *)(R.ws + W.off_perm);
                int64_t g = (nt + 255) / 256;
                if (g > (int64_t)num_sms() * 8) g = (int64_t)num_sms() * 8;
                PC_CUDA(cudaMemsetAsync(hist, 0, 256 * 4, R.st));
                tr_hist_kernel<<<(unsigned)g, 256, 0, R.st>>>(R.base.sub_len, nt, maxlen, hist);
                tr_scan_kernel<<<1, 256, 0, R.st>>>(hist);
                tr_scatter_kernel<<<(unsigned)g, 256, 0, R.st>>>(R.base.sub_len, nt, maxlen, hist, perm);
                PC_LAUNCH_CHECK();
                R.base.perm = perm;
            }
        }
        R.genie_u = genie ? d_genie_u + f0 * Nw : nullptr;
        R.genie_marg = genie ? d_genie_marg + f0 * 2 * plan->N : nullptr;
        uint32_t *root = (uint32_t *)(R.ws + W.off_cw[0]);
        const int rc = trellis_descend(R, 0, 0, root);
        if (rc) return rc;
        PC_CUDA(cudaMemcpyAsync(d_cw_packed + f0 * Nw, root, (size_t)R.frames * Nw * 4, cudaMemcpyDeviceToDevice, R.st));
        if (Kw > 0 && !genie) {
            const int64_t warps = R.frames * Kw;
            merge_info_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, R.st>>>(R.frames, M, d_info_packed + f0 * Kw);
            PC_LAUNCH_CHECK();
        }
    }
    return PC_OK;
}


int pc_trellis_decode(const pc_plan *plan, int n0, double deletion_prob, int ones, const uint8_t *d_sub_bits,
                      const int32_t *d_sub_len, int maxlen, int64_t B, uint32_t *d_cw_packed, uint32_t *d_info_packed,
                      double *d_first_collapse, void *d_workspace, size_t workspace_bytes, void *stream) {
    return trellis_common(plan, n0, deletion_prob, ones, d_sub_bits, d_sub_len, maxlen, B, d_cw_packed, d_info_packed,
                          d_first_collapse, nullptr, nullptr, d_workspace, workspace_bytes, stream);
}

int pc_trellis_genie(const pc_plan *plan, int n0, double deletion_prob, int ones, const uint8_t *d_sub_bits,
                     const int32_t *d_sub_len, int maxlen, const uint32_t *d_u_packed, int64_t B, uint32_t *d_cw_packed,
                     double *d_marg, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_u_packed || !d_marg) {
        pc::set_error("genie pass needs the known u bits and the marginal output");
        return PC_ERR_INVALID;
    }
    return trellis_common(plan, n0, deletion_prob, ones, d_sub_bits, d_sub_len, maxlen, B, d_cw_packed, nullptr, nullptr,
                          d_u_packed, d_marg, d_workspace, workspace_bytes, stream);
}

}  // extern "C"
