// sc_stream.cu -- binary SC decoding of LARGE blocks (N up to 2^24), one frame per CTA, float64, bit-identical to the
// reference (BinaryPolarEncoderDecoder.recursiveEncodeDecode, BinaryPolarEncoderDecoder.py:223-325, with the
// BinaryMemorylessVectorDistribution arithmetic of sc_arith.cuh).
//
// The frame-per-lane decoder of sc_binary.cu needs 32 frames per warp and N float64 of private state per frame; at
// N = 2^20 that is neither available (a frame is 16 MiB of channel probabilities) nor parallel enough.  Here:
//  * a CTA owns one frame and walks the pruned tree; the f / g node updates of the upper stages run element-parallel over
//    the CTA and STREAM through HBM (level vectors above LSM live in a per-CTA global scratch, read and written as
//    coalesced 16-byte pairs in the reference's index order -- element h of a child comes from elements 2h, 2h+1 of its
//    parent, so the channel probabilities are consumed in the caller's layout with no transpose pass);
//  * levels <= LSM finish in shared memory;
//  * the last five stages (aligned blocks of 32 leaves) run in ONE WARP with the level vectors in registers, __shfl for
//    the pair gathers and the partial sums of the block in a register word -- no block-wide barrier inside a block;
//  * partial sums of the larger nodes are bit-packed per level (shared memory below XSL, global above); the combine step
//    x[2h] = m[h] ^ p[h], x[2h+1] = p[h] (BinaryPolarEncoderDecoder.py:321-323) is a 16 -> 32 bit interleave;
//  * all-frozen (rate-0) sub-trees are pruned on the host: their codeword is known, their probabilities never computed.
#include <map>
#include <mutex>

#include "sc_arith.cuh"

namespace pc {

enum : int { SOP_F = 0, SOP_G = 1, SOP_COMBINE = 2, SOP_R0 = 3, SOP_BLOCK = 4 };
constexpr int STREAM_BLOCK_L = 5;  // leaves per warp-resident block = 2^5

struct StreamTables {
    std::vector<uint2> ops;            // x = kind | l << 3 | c << 8 | count << 9, y = argument
    std::vector<SchedEntry> sched;     // plan->sched with rate-0 codewords in REFERENCE order
    std::vector<uint32_t> r0_words;    // reference-order codewords of rate-0 nodes with l >= 5
    std::vector<int32_t> info_pos;     // information index -> u index
    uint2 *d_ops = nullptr;
    SchedEntry *d_sched = nullptr;
    uint32_t *d_r0_words = nullptr;
    int32_t *d_info_pos = nullptr;
};

static std::mutex g_stream_mu;
static std::map<const pc_plan *, StreamTables *> g_stream_tables;

static uint32_t rev_bits(uint32_t j, int l) {
    uint32_t r = 0;
    for (int t = 0; t < l; ++t) r |= ((j >> t) & 1u) << (l - 1 - t);
    return r;
}

// reference-order codeword of the all-frozen node (i, l): natural-order butterfly, then per-node bit reversal
static void r0_codeword_ref(const pc_plan *p, int i, int l, std::vector<uint8_t> &out) {
    const int size = 1 << l;
    std::vector<uint8_t> c(p->frozen_vals.begin() + i, p->frozen_vals.begin() + i + size);
    for (int s = 1; s < size; s <<= 1)
        for (int b = 0; b < size; b += 2 * s)
            for (int j = b; j < b + s; ++j) c[j] ^= c[j + s];
    out.resize(size);
    for (int j = 0; j < size; ++j) out[j] = c[rev_bits((uint32_t)j, l)] & 1;
}

static bool all_frozen(const std::vector<int32_t> &pre, int i, int l) { return pre[i + (1 << l)] - pre[i] == (1 << l); }

static void stream_walk(const pc_plan *p, StreamTables &T, const std::vector<int32_t> &pre, const std::vector<int32_t> &first_entry,
                        int i, int l) {
    const int c = (i >> l) & 1;
    if (all_frozen(pre, i, l)) {
        std::vector<uint8_t> cw;
        r0_codeword_ref(p, i, l, cw);
        const uint32_t off = (uint32_t)T.r0_words.size();
        const int words = l >= 5 ? 1 << (l - 5) : 1;
        for (int w = 0; w < words; ++w) {
            uint32_t v = 0;
            for (int b = 0; b < 32 && 32 * w + b < (1 << l); ++b) v |= (uint32_t)cw[32 * w + b] << b;
            T.r0_words.push_back(v);
        }
        T.ops.push_back(make_uint2((uint32_t)SOP_R0 | (uint32_t)l << 3 | (uint32_t)c << 8, off));
        return;
    }
    if (l == STREAM_BLOCK_L) {
        const int e0 = first_entry[i >> STREAM_BLOCK_L];
        int cnt = 0;
        while (e0 + cnt < (int)T.sched.size() && T.sched[e0 + cnt].i < i + (1 << l)) ++cnt;
        T.ops.push_back(make_uint2((uint32_t)SOP_BLOCK | (uint32_t)l << 3 | (uint32_t)c << 8 | (uint32_t)cnt << 9, (uint32_t)e0));
        return;
    }
    const int half = 1 << (l - 1);
    if (!all_frozen(pre, i, l - 1)) T.ops.push_back(make_uint2((uint32_t)SOP_F | (uint32_t)(l - 1) << 3, 0u));
    stream_walk(p, T, pre, first_entry, i, l - 1);
    if (!all_frozen(pre, i + half, l - 1)) T.ops.push_back(make_uint2((uint32_t)SOP_G | (uint32_t)(l - 1) << 3, 0u));
    stream_walk(p, T, pre, first_entry, i + half, l - 1);
    T.ops.push_back(make_uint2((uint32_t)SOP_COMBINE | (uint32_t)l << 3 | (uint32_t)c << 8, 0u));
}

template <class T>
static cudaError_t upload(T *&dst, const std::vector<T> &v) {
    cudaError_t e = cudaMalloc((void **)&dst, sizeof(T) * (v.size() ? v.size() : 1));
    if (e != cudaSuccess) return e;
    if (v.size()) e = cudaMemcpy(dst, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return e;
}

static StreamTables *stream_tables(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_stream_tables.find(p);
    if (it != g_stream_tables.end()) return it->second;
    StreamTables *T = new StreamTables();
    const int N = p->N;
    std::vector<int32_t> pre(N + 1, 0);
    for (int i = 0; i < N; ++i) pre[i + 1] = pre[i] + (p->frozen_mask[i] ? 1 : 0);
    T->sched = p->sched;
    std::vector<int32_t> first_entry((N >> STREAM_BLOCK_L) + 1, -1);
    for (size_t e = 0; e < T->sched.size(); ++e) {
        SchedEntry &s = T->sched[e];
        if ((s.i & 31) == 0 && first_entry[s.i >> 5] < 0) first_entry[s.i >> 5] = (int32_t)e;
        if (s.kind == NODE_RATE0 && s.l < 5) {  // natural -> reference order inside the node
            uint32_t v = 0;
            for (int j = 0; j < (1 << s.l); ++j) v |= ((s.bits >> rev_bits((uint32_t)j, s.l)) & 1u) << j;
            s.bits = v;
        }
    }
    T->info_pos.assign(p->k > 0 ? p->k : 1, 0);
    for (int i = 0, q = 0; i < N; ++i)
        if (!p->frozen_mask[i]) T->info_pos[q++] = i;
    stream_walk(p, *T, pre, first_entry, 0, p->n);
    if (upload(T->d_ops, T->ops) != cudaSuccess || upload(T->d_sched, T->sched) != cudaSuccess ||
        upload(T->d_r0_words, T->r0_words) != cudaSuccess || upload(T->d_info_pos, T->info_pos) != cudaSuccess) {
        set_error("stream tables: device upload failed");
        return nullptr;
    }
    g_stream_tables[p] = T;
    return T;
}

void stream_tables_release(const pc_plan *p) {
    std::lock_guard<std::mutex> lk(g_stream_mu);
    auto it = g_stream_tables.find(p);
    if (it == g_stream_tables.end()) return;
    StreamTables *T = it->second;
    cudaFree(T->d_ops), cudaFree(T->d_sched), cudaFree(T->d_r0_words), cudaFree(T->d_info_pos);
    delete T;
    g_stream_tables.erase(it);
}

__host__ __device__ inline int st_W(int l) { return l <= 5 ? 1 : 1 << (l - 5); }
__host__ __device__ inline int64_t st_wsum(int l) { return l <= 6 ? l : 4 + ((int64_t)1 << (l - 5)); }  // sum of W(0..l-1)

struct StreamParams {
    int n, k, n_ops, lsm, xsl, Y;
    int64_t frames;
    const uint2 *ops;
    const SchedEntry *sched;
    const uint32_t *r0_words;
    const int32_t *info_pos;
    const void *in;       // [frames][N] double2 probability pairs or uint8 symbols (reference order)
    double *vg;           // [grid][vg_stride] level vectors above lsm
    uint32_t *xg;         // [grid][xg_stride] partial sums of levels >= xsl, then the u-decision words
    int64_t vg_stride, xg_stride;
    uint32_t *cw_out, *info_out;
    double table[32];
};

__device__ __forceinline__ uint32_t st_spread16(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

template <int KIND>
__global__ void __launch_bounds__(256) sc_stream_kernel(const StreamParams p) {
    extern __shared__ __align__(16) unsigned char st_smem[];
    const int n = p.n, N = 1 << n, lsm = p.lsm, xsl = p.xsl;
    const int NW = N >> 5;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    double *Vs = (double *)st_smem;                       // level l at [2^l, 2^(l+1)), l <= lsm
    uint32_t *Xs = (uint32_t *)(Vs + (2 << lsm));         // partial sums of levels < xsl: [2 * wsum(xsl)]
    uint2 *opw = (uint2 *)(Xs + ((2 * st_wsum(xsl) + 1) & ~(int64_t)1));  // [2][64]
    double *s_table = (double *)(opw + 128);              // [32]
    if (tid < 32) s_table[tid] = p.table[tid];
    double *vg = p.vg + (int64_t)blockIdx.x * p.vg_stride - (2 << lsm);
    uint32_t *xg = p.xg + (int64_t)blockIdx.x * p.xg_stride - 2 * st_wsum(xsl);
    uint32_t *U = p.xg + (int64_t)blockIdx.x * p.xg_stride + 2 * (st_wsum(n + 1) - st_wsum(xsl));
    auto V = [&](int l) -> double * { return (l <= lsm ? Vs : vg) + ((int64_t)1 << l); };
    auto X = [&](int l, int c) -> uint32_t * { return (l < xsl ? Xs : xg) + 2 * st_wsum(l) + c * st_W(l); };

    for (int64_t f = blockIdx.x; f < p.frames; f += gridDim.x) {
        __syncthreads();
        if (tid < 64 && tid < p.n_ops) opw[tid] = p.ops[tid];
        __syncthreads();
        uint2 op_pre = make_uint2(0u, 0u);
        for (int oi = 0; oi < p.n_ops; ++oi) {
            const int wi = oi & 63, buf = (oi >> 6) & 1;
            if (tid < 64) {
                if (wi == 0 && oi + 64 + tid < p.n_ops) op_pre = p.ops[oi + 64 + tid];
                if (wi == 63) opw[(buf ^ 1) * 64 + tid] = op_pre;
            }
            const uint2 opk = opw[buf * 64 + wi];
            const int kind = opk.x & 7, l = (opk.x >> 3) & 31, c = (opk.x >> 8) & 1;
            if (kind == SOP_F || kind == SOP_G) {
                // V[l][h] from elements 2h, 2h+1 of level l+1 (the channel when l + 1 == n)
                const int size = 1 << l;
                double *dst = V(l);
                const uint32_t *ub = X(l, 0);
                if (l + 1 == n) {
                    for (int h = tid; h < size; h += T) {
                        double a0, a1, b0, b1;
                        if (KIND == PC_INPUT_SYMBOLS) {
                            const uchar2 y = ((const uchar2 *)p.in)[f * (N >> 1) + h];
                            a0 = s_table[2 * y.x], a1 = s_table[2 * y.x + 1];
                            b0 = s_table[2 * y.y], b1 = s_table[2 * y.y + 1];
                        } else {
                            const double2 *src = (const double2 *)p.in + f * N;
                            const double2 a = src[2 * h], b = src[2 * h + 1];
                            a0 = a.x, a1 = a.y, b0 = b.x, b1 = b.y;
                        }
                        dst[h] = kind == SOP_F ? f_raw(a0, a1, b0, b1) : g_raw(a0, a1, b0, b1, (ub[h >> 5] >> (h & 31)) & 1u);
                    }
                } else {
                    const double2 *src = (const double2 *)V(l + 1);
                    for (int h = tid; h < size; h += T) {
                        const double2 ab = src[h];
                        dst[h] = kind == SOP_F ? f_packed01(ab.x, ab.y) : g_packed01(ab.x, ab.y, (ub[h >> 5] >> (h & 31)) & 1u);
                    }
                }
                __syncthreads();
                continue;
            }
            if (kind == SOP_COMBINE) {
                const int Wo = st_W(l);
                const uint32_t *m = X(l - 1, 0), *pp = X(l - 1, 1);
                uint32_t *out = X(l, c);
                for (int w = tid; w < Wo; w += T) {
                    const int sh = (w & 1) * 16;
                    const uint32_t m16 = (m[w >> 1] >> sh) & 0xffffu, p16 = (pp[w >> 1] >> sh) & 0xffffu;
                    out[w] = st_spread16(m16 ^ p16) | (st_spread16(p16) << 1);
                }
                __syncthreads();
                continue;
            }
            if (kind == SOP_R0) {
                const int Wo = st_W(l);
                uint32_t *out = X(l, c);
                for (int w = tid; w < Wo; w += T) out[w] = p.r0_words[opk.y + w];
                __syncthreads();
                continue;
            }
            // ---- SOP_BLOCK: 32 leaves in warp 0; level vectors 0..5 in shared memory, warp-level barriers only ---------
            if (warp == 0) {
                const int e0 = (int)opk.y, cnt = (opk.x >> 9) & 63;
                // the block's schedule entries (<= 32) are fetched with one coalesced load and broadcast by shuffle
                int my_i = 0;
                uint32_t my_meta = 0, my_bits = 0;
                if (lane < cnt) {
                    const uint32_t *se = (const uint32_t *)(p.sched + e0 + lane);
                    my_i = (int)se[0];
                    my_meta = se[1];
                    my_bits = se[2];
                }
                uint32_t x5 = 0, ubits = 0;
                const int i_block = __shfl_sync(0xffffffffu, my_i, 0);
                for (int q = 0; q < cnt; ++q) {
                    const int e_i = __shfl_sync(0xffffffffu, my_i, q);
                    const uint32_t e_meta = __shfl_sync(0xffffffffu, my_meta, q), e_bits = __shfl_sync(0xffffffffu, my_bits, q);
                    const int li = e_i & 31, el = (int)(e_meta & 0xff), e_kind = (int)((e_meta >> 8) & 0xff),
                              top = (int)((e_meta >> 16) & 0xff);
                    const int stop = e_kind == NODE_RATE0 ? el + 1 : el;
                    int lev;
                    if (li == 0) {
                        lev = 4;
                    } else if (top >= stop) {  // g at level top with the sibling's partial sums
                        if (lane < (1 << top)) {
                            const double2 ab = ((const double2 *)(Vs + (2 << top)))[lane];
                            Vs[(1 << top) + lane] = g_packed01(ab.x, ab.y, (Xs[2 * top] >> lane) & 1u);
                        }
                        __syncwarp();
                        lev = top - 1;
                    } else {
                        lev = -1;
                    }
                    for (; lev >= stop; --lev) {
                        if (lane < (1 << lev)) {
                            const double2 ab = ((const double2 *)(Vs + (2 << lev)))[lane];
                            Vs[(1 << lev) + lane] = f_packed01(ab.x, ab.y);
                        }
                        __syncwarp();
                    }
                    // every lane keeps the (warp-uniform) partial-sum words of the block: same value to the same address
                    if (e_kind == NODE_INFO) {  // p0 >= p1 -> 0 (ties and (0,0) -> 0), BinaryPolarEncoderDecoder.py:252
                        const uint32_t bit = d_sign(Vs[1]);
                        ubits |= bit << li;
                        Xs[li & 1] = bit;
                    } else {
                        Xs[2 * el + ((li >> el) & 1)] = e_bits;
                    }
                    int lv = el, ii = li;
                    while (lv < 5 && ((ii >> lv) & 1)) {  // a plus child completed: x[2h] = m[h] ^ p[h], x[2h+1] = p[h]
                        const uint32_t m = Xs[2 * lv], pw = Xs[2 * lv + 1];
                        const uint32_t cw = st_spread16(m ^ pw) | (st_spread16(pw) << 1);
                        ii -= 1 << lv;
                        ++lv;
                        if (lv == 5)
                            x5 = cw;
                        else
                            Xs[2 * lv + ((ii >> lv) & 1)] = cw;
                    }
                }
                if (lane == 0) {
                    X(5, c)[0] = x5;
                    U[i_block >> 5] = ubits;
                }
            }
            __syncthreads();
        }
        // ---- outputs: the root codeword (already in the reference's order) and the information bits -------------------
        const uint32_t *root = X(n, 0);
        uint32_t *cw = p.cw_out + f * NW;
        for (int w = tid; w < NW; w += T) cw[w] = root[w];
        const int Kw = (p.k + 31) >> 5;
        uint32_t *io = p.info_out + f * Kw;
        for (int w = warp; w < Kw; w += nwarps) {
            const int j = 32 * w + lane;
            uint32_t b = 0;
            if (j < p.k) {
                const int pos = p.info_pos[j];
                b = (U[pos >> 5] >> (pos & 31)) & 1u;
            }
            const uint32_t wv = __ballot_sync(0xffffffffu, b);
            if (lane == 0) io[w] = wv;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct StreamConfig {
    int lsm, xsl, threads, grid;
    size_t smem, vg_stride, xg_stride;
};

static int st_env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

static StreamConfig stream_config(const pc_plan *plan, int64_t B) {
    StreamConfig c{};
    const int n = plan->n;
    int lsm = st_env_int("PC_STREAM_LSM", 10);
    if (lsm < STREAM_BLOCK_L) lsm = STREAM_BLOCK_L;
    if (lsm > n - 1) lsm = n - 1;
    if (lsm > 13) lsm = 13;
    int xsl = lsm + 3 < n + 1 ? lsm + 3 : n + 1;
    c.lsm = lsm;
    c.xsl = xsl;
    c.smem = (size_t)(2 << lsm) * 8 + (size_t)((2 * st_wsum(xsl) + 1) & ~(int64_t)1) * 4 + 128 * 8 + 32 * 8;
    c.threads = st_env_int("PC_STREAM_THREADS", 128);
    if (c.threads < 64 || c.threads > 256 || (c.threads & 31)) c.threads = 128;
    int per_sm = (int)((227 * 1024) / (c.smem + 1024));
    if (per_sm > 2048 / c.threads) per_sm = 2048 / c.threads;
    const int forced = st_env_int("PC_STREAM_CTAS_PER_SM", 0);
    if (forced > 0 && forced < per_sm) per_sm = forced;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)num_sms() * per_sm;
    if (grid > B) grid = B;
    c.grid = (int)(grid > 0 ? grid : 1);
    c.vg_stride = (size_t)(((int64_t)1 << n) - (2 << lsm) > 0 ? ((int64_t)1 << n) - (2 << lsm) : 0) + 2;
    c.xg_stride = (size_t)(2 * (st_wsum(n + 1) - st_wsum(xsl))) + ((size_t)1 << n) / 32 + 4;
    return c;
}

bool sc_stream_supported(const pc_plan *plan) { return plan && plan->q == 2 && plan->n > STREAM_BLOCK_L && plan->n <= 24; }

int64_t sc_stream_wave_frames(const pc_plan *plan) { return stream_config(plan, (int64_t)1 << 40).grid; }

size_t sc_stream_workspace_bytes(const pc_plan *plan, int64_t B) {
    const StreamConfig c = stream_config(plan, B);
    return align256((size_t)c.grid * c.vg_stride * 8 + 256) + align256((size_t)c.grid * c.xg_stride * 4 + 256);
}

int sc_stream_decode(const pc_plan *plan, int kind, const void *d_in, int64_t B, const double *h_table, int Y, uint32_t *d_cw,
                     uint32_t *d_info, void *ws, size_t ws_bytes, cudaStream_t st) {
    PC_REQUIRE(sc_stream_supported(plan), "streamed SC decoder needs a binary plan with 6 <= n <= 24");
    if (B == 0) return PC_OK;
    StreamConfig c = stream_config(plan, B);
    // fewer resident CTAs when the caller's workspace is smaller than the full grid needs
    while (c.grid > 1 && align256((size_t)c.grid * c.vg_stride * 8 + 256) + align256((size_t)c.grid * c.xg_stride * 4 + 256) > ws_bytes)
        c.grid = (c.grid + 1) / 2;
    const size_t vbytes = align256((size_t)c.grid * c.vg_stride * 8 + 256);
    if (vbytes + align256((size_t)c.grid * c.xg_stride * 4 + 256) > ws_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed for one frame", ws_bytes,
                  vbytes + align256((size_t)c.grid * c.xg_stride * 4 + 256));
        return PC_ERR_NOMEM;
    }
    StreamTables *T = stream_tables(plan);
    if (!T) return PC_ERR_CUDA;
    StreamParams p{};
    p.n = plan->n;
    p.k = plan->k;
    p.n_ops = (int)T->ops.size();
    p.lsm = c.lsm;
    p.xsl = c.xsl;
    p.Y = Y;
    p.frames = B;
    p.ops = T->d_ops;
    p.sched = T->d_sched;
    p.r0_words = T->d_r0_words;
    p.info_pos = T->d_info_pos;
    p.in = d_in;
    p.vg = (double *)ws;
    p.xg = (uint32_t *)((char *)ws + vbytes);
    p.vg_stride = (int64_t)c.vg_stride;
    p.xg_stride = (int64_t)c.xg_stride;
    p.cw_out = d_cw;
    p.info_out = d_info;
    for (int i = 0; i < 32; ++i) p.table[i] = (kind == PC_INPUT_SYMBOLS && i < 2 * Y) ? h_table[i] : 0.0;
    prof_mark(st);
    if (kind == PC_INPUT_SYMBOLS) {
        PC_CUDA(cudaFuncSetAttribute(sc_stream_kernel<PC_INPUT_SYMBOLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        sc_stream_kernel<PC_INPUT_SYMBOLS><<<c.grid, c.threads, c.smem, st>>>(p);
    } else {
        PC_CUDA(cudaFuncSetAttribute(sc_stream_kernel<PC_INPUT_PROBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
        sc_stream_kernel<PC_INPUT_PROBS><<<c.grid, c.threads, c.smem, st>>>(p);
    }
    prof_mark(st);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // namespace pc
