// scl_warp.cu -- binary (q = 2) SC-list decoding, ONE FRAME PER WARP, float64 linear domain, arithmetic and decisions
// identical to QaryPolarEncoderDecoder.listDecode with q = 2 (QaryPolarEncoderDecoder.py:118-227, recursiveListDecode
// :403-757, helpers :759-820, normalize :867-872, QaryMemorylessVectorDistribution.py:26-118).
//
// Why this mapping (profiles/r1_c_scl_ncu_summary.md): a list decoder walks ~1000 dependent ops per N = 4096 frame and most
// of them touch fewer than 300 (path, element) items.  With a frame per CTA (scl_bin.cu) every op costs CTA barriers, every
// warp re-decodes the op and re-derives the same addresses, the code footprint thrashes the instruction cache, and only
// ~13 % of the issued instructions are node arithmetic.  Here a warp owns a frame:
//  * no CTA barrier anywhere -- phases are separated by __syncwarp; a CTA is one warp, up to 32 resident per SM, each at its
//    own point of its own frame, so the SM always has independent work to issue;
//  * per-frame state is small enough for that residency: path vectors of levels <= lsm, the short path codewords, the genie
//    / frozen codeword bits and the list bookkeeping live in ~8 KB of shared memory; the larger levels stream through a
//    per-warp global scratch (written and re-read with 16-byte coalesced accesses, mostly L2 hits);
//  * vectors stay in the REFERENCE's index order (children of a node are elements (2h, 2h+1)), so the channel
//    probabilities are read in the caller's layout -- no ingest / transpose pass;
//  * lazy path copy: a pruned list is a permutation table per (level, child) (omap), never a copy of a vector;
//  * pruning is L rounds of a warp arg-max over the candidate metrics (three redux.sync per round on the order-preserving
//    integer image of the non-negative float64 metrics), giving the oracle's ascending (metric, index) order;
//  * the order-dependent float64 products of the fast nodes (np.product is a left-to-right product) run one path per lane;
//  * the genie ("actual") path that listDecode tracks is slot L of every array.
#include "scl_arith.cuh"

namespace pc {

struct SclwParams {
    int n, k, L, n_ops, nfrozen, lsm;
    int rgl;               // path codewords of levels >= rgl live in the global scratch `rg`, smaller levels in shared memory
    int xy_al32;           // the channel input is 32-byte aligned
    int tx_words;          // extra shared words behind the path vectors so that the prologue's two N-bit temporaries fit
    int64_t frames;
    const uint2 *ops;      // packed ops (SclTables::ops3): x = kind | l << 3 | c << 7 | i << 8 | fused << 30, y = fv_idx | coefw_off << 16
    const int32_t *a_src, *f_src, *info_src, *perm;
    const uint32_t *stage_mask, *coef_words;
    const double2 *xy;     // [frames][N] caller layout (reference order)
    const uint8_t *fv;     // [frames][N-k]
    const uint8_t *ainfo;  // [frames][k]
    double2 *vg;           // [grid][vg_stride] scratch for levels > lsm
    int64_t vg_stride;
    uint32_t *rg;          // [grid][rg_stride]
    int64_t rg_stride;
    uint8_t *info;         // [frames][k]
    int32_t *result;       // [frames]
    int32_t *list_size;    // optional, caller layouts
    double *list_prob, *actual_prob;
    uint8_t *list_info;
};

// level l of the path vectors starts at double2 index ((1 << l) - 1) * S; slot stride 1 << l; levels 0 .. n-1
static size_t sclw_smem_bytes(int n, int L, int lsm, int rgl, int *tx_words) {
    const int S = L + 1, N = 1 << n, NW = N >= 32 ? N >> 5 : 1;
    size_t v = (size_t)((2 << lsm) - 1) * S * 16;
    int tx = 0;
    if (v < (size_t)3 * NW * 4) tx = (int)(((size_t)3 * NW * 4 - v + 15) / 16 * 4);
    if (tx_words) *tx_words = tx;
    size_t b = v + (size_t)tx * 4;
    b += (size_t)(11 * L + 4) * 8;                                  // prob, newprob, basep, cand, misc
    b += (size_t)2 * (n + 1) * 8;                                   // vptr, rptr
    b += (size_t)2 * S * scl2_wsum(rgl < n + 1 ? rgl : n + 1) * 4;  // Rw (levels < rgl)
    b += (size_t)2 * NW * 4;                                        // Abits, Fbits
    b += (size_t)(4 * L + 3 * (n + 1) + 4) * 4;                     // keep, selsrc, selfk, hds, nl, nin, ivars
    b += (size_t)4 * L * 2 + L + (size_t)(n + 1) * 2 * L + L;       // pick, delta, omap, eqf
    return (b + 15) & ~(size_t)15;
}

// 32-byte global accesses (two adjacent float64 pairs): one LDG.256 / STG.256 when the address is 32-byte aligned
__device__ __forceinline__ void ld32(const double2 *p, bool aligned32, double2 &a, double2 &b) {
    if (aligned32) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
    } else {
        a = p[0];
        b = p[1];
    }
}
__device__ __forceinline__ void st32(double2 *p, const double2 a, const double2 b) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}

__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int sh = 16; sh > 0; sh >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, v, sh);
        v = o > v ? o : v;
    }
    return v;
}

__global__ void __launch_bounds__(32, 24) sclw_kernel(const SclwParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, N = 1 << n, L = p.L, S = L + 1, k = p.k, lsm = p.lsm;
    const int NW = N >= 32 ? N >> 5 : 1;
    const int lane = threadIdx.x;
    constexpr uint32_t FULL = 0xffffffffu;
    // ---- shared-memory carve-up (sclw_smem_bytes mirrors this) -----------------------------------------
    const int VS = ((2 << lsm) - 1) * S;
    double2 *Vs = (double2 *)smem_raw;
    double *prob = (double *)(Vs + VS) + p.tx_words / 2;
    double *newprob = prob + L;
    double *basep = newprob + L;   // Rate-1 / SPC: prob[t] * product of the non-forked maxima
    double *cand = basep + L;      // [8 L]
    double *misc = cand + 8 * L;   // [0] genie product of the node, [1] actual_prob
    double2 **vptr = (double2 **)(misc + 4);        // [n+1] base of the level-l path vectors ([n]: the frame's channel input)
    uint32_t **rptr = (uint32_t **)(vptr + n + 1);  // [n+1] base of the level-l path codewords
    uint32_t *Rw = (uint32_t *)(rptr + n + 1);
    const int rgl = p.rgl < n + 1 ? p.rgl : n + 1;
    uint32_t *Abits = Rw + 2 * S * scl2_wsum(rgl);
    uint32_t *Fbits = Abits + NW;
    int *keep = (int *)(Fbits + NW);
    int *selsrc = keep + L, *selfk = selsrc + L;
    uint32_t *hds = (uint32_t *)(selfk + L);  // [L] hard-decision words of fast nodes smaller than 32
    int *nl = (int *)(hds + L);   // [(n+1)][2]
    int *nin = nl + 2 * (n + 1);  // [n+1]
    int *ivars = nin + (n + 1);   // [4]
    int16_t *pick = (int16_t *)(ivars + 4);  // [L][4]
    uint8_t *delta = (uint8_t *)(pick + 4 * L);
    uint8_t *omap = delta + L;               // [(n+1)][2][L]
    uint32_t *T0 = (uint32_t *)smem_raw, *T1 = T0 + NW, *T2 = T1 + NW;  // prologue / epilogue temporaries over the (then dead) path vectors

    {
        double2 *vg = p.vg + (int64_t)blockIdx.x * p.vg_stride - VS;
        uint32_t *rgc = p.rg + (int64_t)blockIdx.x * p.rg_stride - 2 * S * scl2_wsum(rgl);
        for (int l = lane; l <= n; l += 32) {
            vptr[l] = (l <= lsm ? Vs : vg) + ((1 << l) - 1) * S;
            rptr[l] = (l < rgl ? Rw : rgc) + 2 * S * scl2_wsum(l);
        }
    }
    double *mxs = (double *)(p.vg + (int64_t)(blockIdx.x + 1) * p.vg_stride) - (N / 2 + 2);  // per-leaf list maxima (<= N/2 leaves)
    auto R = [&](int l, int c, int slot) -> uint32_t * { return rptr[l] + (c * S + slot) * scl2_W(l); };
    auto OM = [&](int l, int c) -> uint8_t * { return omap + (l * 2 + c) * L; };

#pragma unroll 1
    for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
        const double2 *xyf = p.xy + f * N;
        const uint8_t *fvf = p.fv ? p.fv + f * p.nfrozen : nullptr;
        const uint8_t *aif = p.ainfo + f * k;
        __syncwarp();
        // ---- node-local codewords of the genie path (A) and of the frozen values (F) ------------------------
        // u-domain bits in natural order, masked butterfly up to each fast node's size, then the per-node bit reversal
        // that turns natural positions into the reference's order.
#pragma unroll 1
        for (int w = 0; w < NW; ++w) {
            const int pos = 32 * w + lane;
            uint32_t a = 0, fb = 0;
            if (pos < N) {
                const int sa = p.a_src[pos], sf = p.f_src[pos];
                a = sa >= 0 ? aif[sa] : fvf[~sa];
                fb = sf >= 0 ? fvf[sf] : 0u;
            }
            const uint32_t wa = __ballot_sync(FULL, a & 1u), wf = __ballot_sync(FULL, fb & 1u);
            if (lane == 0) {
                T0[w] = wa;
                T1[w] = wf;
            }
        }
        __syncwarp();
        // root codeword of the actual word (reference order = bit reversal of the natural-order transform): R(n, 0, L)
#pragma unroll 1
        for (int w = lane; w < NW; w += 32) T2[w] = T0[w];
        __syncwarp();
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
        {
            uint32_t *groot = R(n, 0, L);
#pragma unroll 1
            for (int w = 0; w < NW; ++w) {
                const int pos = 32 * w + lane;
                uint32_t b = 0;
                if (pos < N) {
                    const uint32_t r = bitrev_n((uint32_t)pos, n);
                    b = (T2[r >> 5] >> (r & 31)) & 1u;
                }
                const uint32_t wv = __ballot_sync(FULL, b);
                if (lane == 0) groot[w] = wv;
            }
        }
#pragma unroll 1
        for (int t = 0; t < n; ++t) {
            const int s = 1 << t;
#pragma unroll 1
            for (int w = lane; w < NW; w += 32) {
                const uint32_t m = p.stage_mask[t * NW + w];
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
                const int src = p.perm[i];
                a = (T0[src >> 5] >> (src & 31)) & 1u;
                fb = (T1[src >> 5] >> (src & 31)) & 1u;
            }
            const uint32_t wa = __ballot_sync(FULL, a), wf = __ballot_sync(FULL, fb);
            if (lane == 0) {
                Abits[w] = wa;
                Fbits[w] = wf;
            }
        }
        if (lane == 0) {
            prob[0] = 1.0;
            misc[1] = 1.0;
            nin[n] = 1;
            nl[n * 2 + 0] = 1;
            vptr[n] = (double2 *)xyf;
        }
        __syncwarp();

        // pass 0 decodes the list; pass 1 (genie) replays the ops for the actual path alone (slot L) to get listDecode's
        // actual_prob.  ProbResult needs it only when the actual word is not in the final list, so pass 1 is skipped for
        // frames that decoded correctly unless the caller asked for the final-list outputs.
        int found = -1;
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
        const bool genie = pass == 1;
        if (genie) {
            // the actual word is in the list iff a root codeword equals its codeword: T(u) with the same frozen values, so
            // compare the information bits instead -- bits of A at information positions vs the path's u
            const int cntf = nl[n * 2 + 0];
#pragma unroll 1
            for (int t = 0; t < cntf; ++t) {
                const uint32_t *a = R(n, 0, t), *b = R(n, 0, L);
                bool eq = true;
#pragma unroll 1
                for (int w = lane; w < NW; w += 32) eq &= a[w] == b[w];
                if (__all_sync(FULL, eq) && found < 0) found = t;
            }
            if (found >= 0 && !p.list_size) break;
            if (lane == 0) misc[1] = 1.0;
            __syncwarp();
        }
        int leaf_idx = 0;
        uint2 opn = p.ops[0];
#pragma unroll 1
        for (int oi = 0; oi < p.n_ops; ++oi) {
            const uint2 opk = opn;
            if (oi + 1 < p.n_ops) opn = p.ops[oi + 1];  // the next op's load overlaps this op
            const int kind = opk.x & 7, l = (opk.x >> 3) & 15, c = (opk.x >> 7) & 1, i0 = (int)((opk.x >> 8) & 0x3fffffu);
            const int size = 1 << l;
            if (kind == OP_MINUS || kind == OP_PLUS) {
                const bool plus = kind == OP_PLUS;
                const int cnt = genie ? 0 : (plus ? nl[(l - 1) * 2 + 0] : nin[l]);
                const int nt = genie ? 1 : cnt;  // paths handled by this pass: the list (slots 0 .. cnt-1) or the genie path (slot L)
                const uint8_t *om = OM(l - 1, 0);
                const double2 *sbase = vptr[l];
                const int sstride = l == n ? 0 : size;  // the channel level is shared by all paths
                double2 *dbase = vptr[l - 1];
                const uint32_t *rb = R(l - 1, 0, 0);
                const int rw = scl2_W(l - 1);
                const bool fused = (opk.x >> 30) & 1u;  // PLUS / MINUS (l) followed by MINUS (l-1): one pass, two levels out
                if (l >= 7) {
                    // large level: iteration `it` covers 128 consecutive source elements of one path; the loads of iteration
                    // it+1 are issued before the arithmetic of iteration it.  Plain: lane h and h+32 of the 64 outputs.
                    // Fused: the lane's 4 consecutive elements give 2 outputs of level l-1 and 1 of level l-2.
                    // (levels >= 6 are always in the global scratch: 32-byte vector loads and stores)
                    const int nblk = 1 << (l - 7);
                    double2 *d2base = vptr[l - 2];
                    const int o2 = fused ? 2 : 64;
                    const bool al = l < n || p.xy_al32;
                    const int loff = fused ? 4 * lane : 2 * lane;
                    auto path_src = [&](int t) -> const double2 * {
                        const int src = genie ? L : (plus ? (int)om[t] : t);
                        return sbase + src * sstride + loff;
                    };
                    const double2 *q = path_src(0);
                    double2 e0, e1, e2, e3;
                    ld32(q, al, e0, e1);
                    ld32(q + o2, al, e2, e3);
#pragma unroll 1
                    for (int t = 0; t < nt; ++t) {
                        const int slot = genie ? L : t;
                        const double2 *qnp = t + 1 < nt ? path_src(t + 1) : nullptr;  // first block of the next path
                        double2 *D = dbase + (slot << (l - 1)) + (fused ? 2 * lane : lane);
                        double2 *D2 = d2base + (slot << (l - 2)) + lane;
                        const uint32_t *rp = rb + slot * rw + (fused ? lane >> 4 : 0);
                        const int sh0 = fused ? (2 * lane) & 31 : lane;
#pragma unroll 1
                        for (int blk = 0; blk < nblk; ++blk) {
                            const double2 *qn = blk + 1 < nblk ? q + 128 : qnp;
                            double2 f0 = e0, f1 = e1, f2 = e2, f3 = e3;
                            if (qn) {
                                ld32(qn, al, f0, f1);
                                ld32(qn + o2, al, f2, f3);
                            }
                            uint32_t u0 = 0, u1 = 0;
                            if (plus) {
                                const uint32_t w0 = rp[0], w1 = fused ? w0 >> 1 : rp[1];
                                u0 = (w0 >> sh0) & 1u;
                                u1 = (w1 >> sh0) & 1u;
                            }
                            const double2 y0 = node_update(e0, e1, plus, u0), y1 = node_update(e2, e3, plus, u1);
                            if (fused) {
                                st32(D, y0, y1);
                                *D2 = node_update(y0, y1, false, 0u);
                            } else {
                                D[0] = y0;
                                D[32] = y1;
                            }
                            D += 64;
                            D2 += 32;
                            rp += 2;
                            e0 = f0, e1 = f1, e2 = f2, e3 = f3;
                            q = qn;
                        }
                    }
                    if (lane == 0 && fused && !genie) nin[l - 2] = cnt;
                } else {
                    const int half = size >> 1;
                    const int total = nt << (l - 1);
#pragma unroll 1
                    for (int idx = lane; idx < total; idx += 32) {
                        const int t = idx >> (l - 1), h = idx & (half - 1);
                        const int slot = genie ? L : t;
                        const int src = genie ? L : (plus ? (int)om[t] : t);
                        const double2 *P = sbase + src * sstride;
                        const double2 a = P[2 * h], b = P[2 * h + 1];
                        const uint32_t u1 = plus ? (rb[slot * rw + (h >> 5)] >> (h & 31)) & 1u : 0u;
                        dbase[(slot << (l - 1)) + h] = node_update(a, b, plus, u1);
                    }
                }
                if (lane == 0 && !genie) nin[l - 1] = cnt;
                __syncwarp();
                continue;
            }
            if (kind == OP_COMBINE) {  // :726-754 in reference order: out[2h] = m[h] + p[h], out[2h+1] = -p[h]
                const int cnt = genie ? 0 : nl[(l - 1) * 2 + 1];
                const int nt = genie ? 1 : cnt;
                const int Wo = scl2_W(l), wsh = l <= 5 ? 0 : l - 5;
                const uint8_t *om1 = OM(l - 1, 1), *om0 = OM(l - 1, 0);
                const uint32_t *rm = R(l - 1, 0, 0), *rp = R(l - 1, 1, 0);
                uint32_t *ro = R(l, c, 0);
                const int Wc = scl2_W(l - 1);
#pragma unroll 1
                for (int idx = lane; idx < nt << wsh; idx += 32) {
                    const int t = idx >> wsh, w = idx & (Wo - 1);
                    const int slot = genie ? L : t;
                    const int mi = genie ? L : (int)om1[t];
                    const int sh = (w & 1) * 16;
                    const uint32_t m16 = (rm[mi * Wc + (w >> 1)] >> sh) & 0xffffu;
                    const uint32_t p16 = (rp[slot * Wc + (w >> 1)] >> sh) & 0xffffu;
                    ro[slot * Wo + w] = spread16(m16 ^ p16) | (spread16(p16) << 1);
                }
                if (!genie) {
#pragma unroll 1
                    for (int t = lane; t < cnt; t += 32) OM(l, c)[t] = om0[om1[t]];
                    if (lane == 0) nl[l * 2 + c] = cnt;
                }
                __syncwarp();
                continue;
            }
            // ------------------------------- fast nodes ------------------------------------------------------
            const int li = leaf_idx++;
            if (genie) {
                // genie pass: only the actual path's product over the node (:503-509 etc.), normalised like the list was
                const double2 *P = vptr[l] + (l == n ? 0 : L * size);
                const uint32_t *aw = Abits + (i0 >> 5);
                const int bsh = size >= 32 ? 0 : (i0 & 31);
                const int Wl = scl2_W(l);
                const uint32_t smask = size >= 32 ? 0xffffffffu : ((1u << size) - 1u);
                if (lane == 0) {
                    double pr = 1.0;
#pragma unroll 1
                    for (int j = 0; j < size; ++j) {
                        const double2 v2 = P[j];
                        pr = __dmul_rn(pr, (aw[j >> 5] >> (bsh + (j & 31))) & 1u ? v2.y : v2.x);
                    }
                    misc[1] = __dmul_rn(misc[1], pr / mxs[li]);
                }
#pragma unroll 1
                for (int w = lane; w < Wl; w += 32) R(l, c, L)[w] = (aw[w] >> bsh) & smask;
                __syncwarp();
                continue;
            }
            const int cnt = nin[l];
            const bool spc = kind == OP_SPC;
            const int nfork = spc ? 3 : 2;
            const int fs = kind == OP_REP ? 2 : (spc ? 8 : 4);
            const int Wl = scl2_W(l);
            const uint32_t smask = size >= 32 ? 0xffffffffu : ((1u << size) - 1u);
            const uint32_t *coefw = p.coef_words + (opk.y >> 16);
            const double2 *Vl = vptr[l];
            const int vstride = l == n ? 0 : size;
            // node-local codeword slices (reference order; i0 is a multiple of the node size): bit j is word j/32, bit j%32 + bsh
            const uint32_t *aw = Abits + (i0 >> 5), *fw = Fbits + (i0 >> 5);
            const int bsh = size >= 32 ? 0 : (i0 & 31);
            int nout = cnt;
            if (kind == OP_RATE0 || kind == OP_REP) {
                // one job per lane: left-to-right product of P[j].{x|y} selected by the candidate codeword's bits
                // (np.product order; Rate-0 :495-518, Rep :521-578)
                const int njobs = kind == OP_REP ? 2 * cnt : cnt;
#pragma unroll 1
                for (int job = lane; job < njobs; job += 32) {
                    const int s = job < cnt ? 0 : 1, t = job - s * cnt;
                    const double2 *P = Vl + t * vstride;
                    const uint32_t *bw = fw;
                    double pr = 1.0;
#pragma unroll 1
                    for (int w0 = 0; w0 < size; w0 += 32) {
                        const uint32_t bits = (bw[w0 >> 5] >> bsh) ^ (s ? coefw[w0 >> 5] : 0u);
                        const int m = size - w0 < 32 ? size - w0 : 32;
#pragma unroll 1
                        for (int b = 0; b < m; ++b) {
                            const double2 v2 = P[w0 + b];
                            pr = __dmul_rn(pr, (bits >> b) & 1u ? v2.y : v2.x);
                        }
                    }
                    if (kind == OP_RATE0)
                        newprob[t] = __dmul_rn(prob[t], pr);
                    else
                        cand[s * cnt + t] = __dmul_rn(prob[t], pr);
                }
            } else {  // Rate-1 :581-628 and SPC :631-682
                // phase 0, all lanes over (path, element): reliabilities (second-largest / largest, :763-768) and hard decisions.
                // Both are parked in the dead level l-1 region of the path vectors (S 2^l doubles): reliabilities in the
                // first cnt 2^l doubles, hard-decision words in the genie slot's share.
                double *scr = (double *)vptr[l - 1];
                uint32_t *hd = l >= 5 ? (uint32_t *)(scr + ((size_t)L << l)) : hds;
                const int total = cnt << l;
#pragma unroll 1
                for (int base = 0; base < total; base += 32) {
                    const int idx = base + lane;
                    const bool valid = idx < total;
                    const int t = idx >> l, j = idx & (size - 1);
                    double2 v2 = make_double2(1.0, 1.0);
                    if (valid) v2 = Vl[t * vstride + j];
                    const bool one = v2.y > v2.x;
                    const double m1 = one ? v2.y : v2.x, m2 = one ? v2.x : v2.y;
                    const double s = m2 / m1;
                    if (valid) scr[idx] = s;
                    const uint32_t bal = __ballot_sync(FULL, valid && one);
                    if (size >= 32) {
                        if (lane == 0) hd[idx >> 5] = bal;
                    } else if (valid && j == 0) {
                        hd[t] = (bal >> (lane & ~(size - 1))) & smask;
                    }
                }
                __syncwarp();
                // pickLeastReliableIndices (:759-768): per path the 2 (Rate-1) or 4 (SPC) largest (score, j), ties to the later
                // index.  G lanes share a path: each keeps the top four of its strided elements (ascending s0 <= .. <= s3,
                // branch-free insertion), then the lists are merged by arg-max rounds inside the lane group.
                {
                    int cl = 0;
                    while ((1 << cl) < cnt) ++cl;
                    const int gsh = 5 - cl, G = 1 << gsh;  // cnt <= 32
                    const int t = lane >> gsh, qq = lane & (G - 1);
                    double s0 = -1.0, s1 = -1.0, s2 = -1.0, s3 = -1.0;
                    int j0 = 0, j1 = 0, j2 = 0, j3 = 0;
                    if (t < cnt) {
                        const double *sp = scr + (t << l);
#pragma unroll 1
                        for (int j = qq; j < size; j += G) {
                            const double s = sp[j];
                            const bool g0 = s >= s0, g1 = s >= s1, g2 = s >= s2, g3 = s >= s3;
                            s0 = g1 ? s1 : (g0 ? s : s0), j0 = g1 ? j1 : (g0 ? j : j0);
                            s1 = g2 ? s2 : (g1 ? s : s1), j1 = g2 ? j2 : (g1 ? j : j1);
                            s2 = g3 ? s3 : (g2 ? s : s2), j2 = g3 ? j3 : (g2 ? j : j2);
                            s3 = g3 ? s : s3, j3 = g3 ? j : j3;
                        }
                    }
                    const int npick = spc ? 4 : 2;
#pragma unroll 1
                    for (int r = 0; r < npick; ++r) {
                        double bs = s3;
                        int bj = j3;
#pragma unroll 1
                        for (int o = 1; o < G; o <<= 1) {
                            const double os = __shfl_xor_sync(FULL, bs, o);
                            const int oj = __shfl_xor_sync(FULL, bj, o);
                            if (os > bs || (os == bs && oj > bj)) bs = os, bj = oj;
                        }
                        if (s3 == bs && j3 == bj) {  // this lane's head was taken
                            s3 = s2, j3 = j2;
                            s2 = s1, j2 = j1;
                            s1 = s0, j1 = j0;
                            s0 = -1.0, j0 = 0;
                        }
                        if (qq == 0 && t < cnt) pick[t * 4 + npick - 1 - r] = (int16_t)bj;
                    }
                    if (!spc && qq == 0 && t < cnt) pick[t * 4 + 2] = pick[t * 4 + 3] = -1;
                }
                __syncwarp();
                // per path, in element order: product of the non-forked maxima
                const int fval = spc ? fvf[opk.y & 0xffffu] : 0;
#pragma unroll 1
                for (int job = lane; job < cnt; job += 32) {
                    const double2 *P = Vl + job * vstride;
                    const int p0 = pick[job * 4], p1 = pick[job * 4 + 1], p2 = pick[job * 4 + 2], p3 = pick[job * 4 + 3];
                    double pr = 1.0;
#pragma unroll 1
                    for (int j = 0; j < size; ++j) {
                        const double2 v2 = P[j];
                        const double v = v2.y > v2.x ? v2.y : v2.x;
                        const bool forked = j == p0 || j == p1 || j == p2 || j == p3;
                        pr = forked ? pr : __dmul_rn(pr, v);
                    }
                    {
                        basep[job] = __dmul_rn(prob[job], pr);
                        // parity of the non-forked hard decisions
                        uint32_t par = 0;
                        const uint32_t *hw = hd + job * Wl;
#pragma unroll 1
                        for (int w = 0; w < Wl; ++w) par ^= hw[w];
                        par = __popc(par);
                        par ^= (hw[p0 >> 5] >> (p0 & 31)) ^ (hw[p1 >> 5] >> (p1 & 31));
                        if (spc) par ^= (hw[p2 >> 5] >> (p2 & 31)) ^ (hw[p3 >> 5] >> (p3 & 31));
                        delta[job] = (uint8_t)((fval ^ par) & 1);
                    }
                }
                __syncwarp();
                // candidate metrics (forkIndices / forkIndicesSpc, :770-820), one (path, fork) per lane
                const int C = cnt * fs;
#pragma unroll 1
                for (int idx = lane; idx < C; idx += 32) {
                    const int t = idx / fs, fk = idx - t * fs;
                    const double2 *P = Vl + t * vstride;
                    const int16_t *pk = pick + t * 4;
                    double pf = 1.0;
                    int sf = 0;
#pragma unroll 1
                    for (int w = 0; w < nfork; ++w) {
                        const int dg = (fk >> (nfork - 1 - w)) & 1;
                        const double2 v2 = P[pk[w]];
                        pf = __dmul_rn(pf, dg ? v2.y : v2.x);
                        sf += dg;
                    }
                    if (spc) {
                        const int dep = (delta[t] ^ sf) & 1;
                        const double2 v2 = P[pk[3]];
                        pf = __dmul_rn(pf, dep ? v2.y : v2.x);
                    }
                    cand[idx] = __dmul_rn(pf, basep[t]);
                }
            }
            __syncwarp();
            // ---- prune (:446-451 etc.): keep the ns = min(#nonzero, L) largest candidates under the total order
            // (metric, index), listed ascending.  Round r extracts the r-th largest: non-negative float64 compare like their
            // bit patterns, so the arg-max is three integer redux.sync (high word, low word, index); a candidate that has
            // been taken is zeroed in place (zeros are never taken: r < ns <= #nonzero).
            if (kind != OP_RATE0) {
                const int C = cnt * fs;
                nout = C;
                if (C > L) {
                    double v0 = lane < C ? cand[lane] : 0.0, v1 = lane + 32 < C ? cand[lane + 32] : 0.0;
                    int nzc = (v0 != 0.0 ? 1 : 0) + (v1 != 0.0 ? 1 : 0);
#pragma unroll 1
                    for (int cc = lane + 64; cc < C; cc += 32) nzc += cand[cc] != 0.0 ? 1 : 0;
                    nzc = __reduce_add_sync(FULL, nzc);
                    const int ns = nzc < L ? nzc : L;
#pragma unroll 1
                    for (int r = 0; r < ns; ++r) {
                        // this lane's best candidate; among equals the later index wins
                        double bv = v1 >= v0 ? v1 : v0;
                        int bi = v1 >= v0 ? lane + 32 : lane;
#pragma unroll 1
                        for (int cc = lane + 64; cc < C; cc += 32) {
                            const double v = cand[cc];
                            if (v >= bv) bv = v, bi = cc;
                        }
                        const uint32_t hi = (uint32_t)__double2hiint(bv), lo = (uint32_t)__double2loint(bv);
                        const uint32_t mh = __reduce_max_sync(FULL, hi);
                        const uint32_t ml = __reduce_max_sync(FULL, hi == mh ? lo : 0u);
                        const uint32_t mi = __reduce_max_sync(FULL, (hi == mh && lo == ml) ? (uint32_t)bi + 1u : 0u) - 1u;
                        if ((mi & 31u) == (uint32_t)lane) {
                            keep[ns - 1 - r] = (int)mi;
                            newprob[ns - 1 - r] = bv;
                            if (mi < 32u)
                                v0 = 0.0;
                            else if (mi < 64u)
                                v1 = 0.0;
                            else
                                cand[mi] = 0.0;
                        }
                    }
                    nout = ns;
                } else {
#pragma unroll 1
                    for (int t = lane; t < C; t += 32) {
                        keep[t] = t;
                        newprob[t] = cand[t];
                    }
                }
                __syncwarp();
            }
            // ---- phase 2: lazy copy (omap), normalise (:867-872) ---------------------------------------------------
            {
                if (kind == OP_RATE0) {
#pragma unroll 1
                    for (int t = lane; t < cnt; t += 32) OM(l, c)[t] = (uint8_t)t;
                } else {
#pragma unroll 1
                    for (int t = lane; t < nout; t += 32) {
                        const int cidx = keep[t];
                        int src, sel;
                        if (kind == OP_REP) {
                            sel = cidx >= cnt ? 1 : 0;
                            src = cidx - sel * cnt;
                        } else {
                            src = cidx / fs;
                            sel = cidx - src * fs;
                        }
                        selsrc[t] = src;
                        selfk[t] = sel;
                        OM(l, c)[t] = (uint8_t)src;
                    }
                }
                double mx = lane < nout ? newprob[lane] : newprob[0];
#pragma unroll 1
                for (int t = lane + 32; t < nout; t += 32) {
                    const double v = newprob[t];
                    if (v > mx) mx = v;
                }
                mx = warp_max_f64(mx);
#pragma unroll 1
                for (int t = lane; t < nout; t += 32) prob[t] = newprob[t] / mx;
                if (lane == 0) {
                    mxs[li] = mx;
                    nl[l * 2 + c] = nout;
                }
            }
            __syncwarp();
            // ---- phase 3: node codewords of the surviving paths and of the genie path ---------------------------
            {
                uint32_t *ro = R(l, c, 0);
                const int wsh = l <= 5 ? 0 : l - 5;
#pragma unroll 1
                for (int idx = lane; idx < nout << wsh; idx += 32) {
                    const int t = idx >> wsh, w = idx & (Wl - 1);
                    uint32_t v;
                    if (kind == OP_RATE0 || kind == OP_REP) {
                        v = (fw[w] >> bsh) & smask;
                        if (kind == OP_REP && selfk[t]) v ^= coefw[w];
                    } else {
                        // the source path's hard decisions with the forked positions overwritten
                        const uint32_t *hd = l >= 5 ? (uint32_t *)((double *)vptr[l - 1] + ((size_t)L << l)) : hds;
                        const int src = selsrc[t], fk = selfk[t];
                        const int16_t *pk = pick + src * 4;
                        v = hd[src * Wl + w];
#pragma unroll 1
                        for (int i = 0; i < nfork; ++i) {
                            const int j = pk[i];
                            if ((j >> 5) == w) v = (v & ~(1u << (j & 31))) | ((uint32_t)((fk >> (nfork - 1 - i)) & 1) << (j & 31));
                        }
                        if (spc) {
                            const int j = pk[3];
                            if ((j >> 5) == w) v = (v & ~(1u << (j & 31))) | ((uint32_t)((delta[src] ^ __popc(fk)) & 1) << (j & 31));
                        }
                    }
                    ro[t * Wl + w] = v;
                }
            }
            __syncwarp();
        }

        }  // passes
        // ---- final selection (listDecode :172-213): the actual word is in the list iff a root codeword equals its codeword ----
        const int cnt = nl[n * 2 + 0];
        if (lane == 0) {
            double maxp = prob[0], minp = prob[0];
#pragma unroll 1
            for (int t = 1; t < cnt; ++t) {
                maxp = prob[t] > maxp ? prob[t] : maxp;
                minp = prob[t] < minp ? prob[t] : minp;
            }
            const double ap = misc[1];  // meaningful (and needed) only when the genie pass ran
            int res;
            if (found >= 0)
                res = prob[found] == maxp ? 0 : 1;
            else
                res = ap > maxp ? 2 : (ap == maxp ? 3 : (ap >= minp ? 4 : 5));
            p.result[f] = res;
            if (p.list_size) {
                p.list_size[f] = cnt;
                p.actual_prob[f] = ap;
#pragma unroll 1
                for (int t = 0; t < L; ++t) p.list_prob[f * L + t] = t < cnt ? prob[t] : 0.0;
            }
        }
        __syncwarp();
        // information of a path = gather of T(root codeword): bit-reverse to natural order, butterfly, gather
        const int sel = found >= 0 ? found : 0;
        const int npaths = p.list_info ? cnt : 1;
#pragma unroll 1
        for (int pi = 0; pi < npaths; ++pi) {
            const int t = p.list_info ? pi : sel;
            const uint32_t *root = R(n, 0, t);
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
            for (int j = lane; j < k; j += 32) {
                const int pos = p.info_src[j];
                const uint8_t v = (uint8_t)((T0[pos >> 5] >> (pos & 31)) & 1u);
                if (p.list_info) p.list_info[(f * L + pi) * k + j] = v;
                if (t == sel) p.info[f * k + j] = v;
            }
            __syncwarp();
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct SclwConfig {
    int lsm, rgl, grid, tx_words;
    size_t smem, vg_stride, rg_stride;  // vg_stride in double2 elements, rg_stride in words, per warp
    bool ok;
};

static int envw_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

static SclwConfig sclw_config(const pc_plan *plan, int L, int64_t B) {
    SclwConfig c{};
    const int n = plan->n, S = L + 1;
    c.ok = false;
    if (plan->q != 2 || n < 1 || n > 13 || L > 32) return c;
    int rgl = envw_int("PC_SCLW_RGL", 9);
    if (rgl < 1) rgl = 1;
    if (rgl > n + 1) rgl = n + 1;
    c.rgl = rgl;
    // shared-memory budget per warp: aim at `target` resident warps per SM
    const int target = envw_int("PC_SCLW_WARPS_PER_SM", 24);
    const size_t budget = (size_t)(227 * 1024) / (size_t)(target > 0 ? target : 1) - 1024;
    int lsm = n - 1 < 5 ? n - 1 : 5;  // levels >= 6 stay in the global scratch (the large-level loops use 32-byte global accesses)
    while (lsm > 0 && sclw_smem_bytes(n, L, lsm, rgl, nullptr) > budget) --lsm;
    const int forced = envw_int("PC_SCLW_LSM", -1);
    if (forced >= 0 && forced <= n - 1 && forced <= 5) lsm = forced;
    c.lsm = lsm;
    c.smem = sclw_smem_bytes(n, L, lsm, rgl, &c.tx_words);
    if (c.smem > 220 * 1024) return c;
    int per_sm = (int)((227 * 1024) / (c.smem + 1024));
    if (per_sm > 32) per_sm = 32;
    const int cap = envw_int("PC_SCLW_MAX_PER_SM", 0);  // tuning: fewer resident warps than the shared memory allows
    if (cap > 0 && per_sm > cap) per_sm = cap;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > B) grid = B;
    c.grid = (int)(grid > 0 ? grid : 1);
    const int64_t vtot = (int64_t)((1 << n) - 1) * S, vs = (int64_t)((2 << lsm) - 1) * S;
    c.vg_stride = (size_t)(vtot > vs ? vtot - vs : 0) + 2 + ((size_t)(1 << n) / 2 + 2 + 1) / 2 + 1;  // + per-leaf maxima (<= N/2 leaves)
    c.rg_stride = (size_t)2 * S * (scl2_wsum(n + 1) - scl2_wsum(rgl)) + 4;
    c.vg_stride = (c.vg_stride + 1) & ~(size_t)1;  // 32-byte aligned per-warp regions
    c.ok = true;
    return c;
}

bool sclw_supported(const pc_plan *plan, int L) {
    static const bool on = envw_int("PC_SCL_WARP", 0) != 0;  // round 1's kernel: comparison runs only (read once)
    if (!on) return false;
    return sclw_config(plan, L, 1).ok;
}

int64_t sclw_wave_frames(const pc_plan *plan, int L) { return sclw_config(plan, L, (int64_t)1 << 40).grid; }

size_t sclw_workspace_bytes(const pc_plan *plan, int L, int64_t B) {
    const SclwConfig c = sclw_config(plan, L, B);
    return align256(align256((size_t)c.grid * c.vg_stride * sizeof(double2) + 256) + (size_t)c.grid * c.rg_stride * 4 + 256);
}

int sclw_decode(const pc_plan *plan, const SclTables *T, int L, const double *d_xy, const uint8_t *d_fv, const uint8_t *d_ainfo,
                int64_t B, uint8_t *d_info, int32_t *d_res, int32_t *d_lsize, double *d_lprob, double *d_aprob,
                uint8_t *d_linfo, void *ws, size_t ws_bytes, cudaStream_t st) {
    const SclwConfig c = sclw_config(plan, L, B);
    if (!c.ok) {
        set_error("sclw: unsupported configuration");
        return PC_ERR_UNSUPPORTED;
    }
    const size_t need = sclw_workspace_bytes(plan, L, B);
    if (need > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", ws_bytes, need);
        return PC_ERR_NOMEM;
    }
    SclwParams p{};
    p.n = plan->n;
    p.k = plan->k;
    p.L = L;
    p.n_ops = (int)T->ops.size();
    p.nfrozen = plan->N - plan->k;
    p.lsm = c.lsm;
    p.rgl = c.rgl;
    p.tx_words = c.tx_words;
    p.xy_al32 = ((uintptr_t)d_xy & 31) == 0 ? 1 : 0;
    p.frames = B;
    p.ops = T->d_ops3;
    p.n_ops = (int)T->ops3.size();
    p.a_src = T->d_a_src;
    p.f_src = T->d_f_src;
    p.info_src = T->d_info_src;
    p.perm = T->d_perm;
    p.stage_mask = T->d_stage_mask;
    p.coef_words = T->d_rep_coef_words;
    p.xy = (const double2 *)d_xy;
    p.fv = d_fv;
    p.ainfo = d_ainfo;
    p.vg = (double2 *)ws;
    p.vg_stride = (int64_t)c.vg_stride;
    p.rg = (uint32_t *)((char *)ws + align256((size_t)c.grid * c.vg_stride * sizeof(double2) + 256));
    p.rg_stride = (int64_t)c.rg_stride;
    p.info = d_info;
    p.result = d_res;
    p.list_size = d_lsize;
    p.list_prob = d_lprob;
    p.actual_prob = d_aprob;
    p.list_info = d_linfo;
    PC_CUDA(cudaFuncSetAttribute(sclw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    prof_mark(st);
    sclw_kernel<<<c.grid, 32, c.smem, st>>>(p);
    prof_mark(st);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // namespace pc
