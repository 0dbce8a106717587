// channel.cu -- the steps either side of the decoders on the device: channel simulation for Monte-Carlo runs
// and the guard-band plumbing of the deletion channel.
//
//  * discrete memoryless channels (BSC / BEC / QSC / any P(y | x) table): the inverse-CDF walk of the reference's simulators
//    (test3.py:35-54 `if probSum + P(y | x) >= rand`, test2.py:29-65) with a COUNTER-BASED generator (Philox4x32-10) keyed by
//    (seed, global frame index, position): results do not depend on the batch split or the number of ranks;
//  * BI-AWGN: y = (1 - 2 x) + sigma * N(0, 1) (Box-Muller on two Philox uniforms), optionally quantised to Y uniform levels
//    on [-ymax, ymax] -- the symbol input of pc_sc_decode_symbols / pc_scl_decode_symbols;
//  * addDeletionGuardBands (Guardbands.py:4-44), deletionChannelSimulation (VectorDistributions/BinaryTrellis.py:441-461: keep
//    a symbol iff r >= p), removeDeletionGuardBands / trimZerosAtEdges (Guardbands.py:47-93) producing the fixed-width sub-word
//    arrays pc_trellis_decode ingests.
// The deterministic parts (guard bands in / out) are bit-exact against the reference; the noise is statistically equivalent
// (the reference draws from CPython's Mersenne Twister one symbol after the other, which no parallel generator reproduces).
#include <math.h>

#include <vector>

#include "common.cuh"

namespace pc {

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// two uniforms in [0, 1) with 53 random bits each for (seed, stream, frame, position)
__device__ __forceinline__ void uniform2(uint64_t seed, uint32_t stream, uint64_t frame, uint32_t pos, double &u0, double &u1) {
    const uint4 r = philox4x32(make_uint4(pos, (uint32_t)frame, (uint32_t)(frame >> 32), stream),
                               make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    u0 = (double)(((uint64_t)(r.x >> 5) << 26) | (uint64_t)(r.y >> 6)) * (1.0 / 9007199254740992.0);
    u1 = (double)(((uint64_t)(r.z >> 5) << 26) | (uint64_t)(r.w >> 6)) * (1.0 / 9007199254740992.0);
}

constexpr int CH_MAXTAB = 1024;  // X * Y entries of the conditional table
struct DmcTab {
    double p[CH_MAXTAB];
};

// x from bytes or from packed bits (binary); y = the first output symbol whose cumulated P(y | x) reaches the uniform draw
template <bool PACKED>
__global__ void __launch_bounds__(256) dmc_kernel(const uint8_t *xb, const uint32_t *xp, int N, int X, int Y, int64_t B,
                                                  int64_t frame0, uint64_t seed, const double *tab, uint8_t *y) {
    extern __shared__ double s_tab[];
    for (int i = threadIdx.x; i < X * Y; i += blockDim.x) s_tab[i] = tab[i];
    __syncthreads();
    const int NW = (N + 31) >> 5;
    const int64_t total = B * (int64_t)N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / N;
        const int pos = (int)(i - f * N);
        int x = PACKED ? (int)((xp[f * NW + (pos >> 5)] >> (pos & 31)) & 1u) : (int)xb[i];
        x = x < X ? x : X - 1;
        double u, unused;
        uniform2(seed, 1u, (uint64_t)(frame0 + f), (uint32_t)pos, u, unused);
        const double *row = s_tab + x * Y;
        double acc = 0.0;
        int out = Y - 1;
        for (int yy = 0; yy < Y; ++yy) {
            if (acc + row[yy] >= u) {
                out = yy;
                break;
            }
            acc += row[yy];
        }
        y[i] = (uint8_t)out;
    }
}

template <bool PACKED>
__global__ void __launch_bounds__(256) awgn_kernel(const uint8_t *xb, const uint32_t *xp, int N, int64_t B, int64_t frame0,
                                                   uint64_t seed, double sigma, int Y, double ymax, uint8_t *yq, double *yf) {
    const int NW = (N + 31) >> 5;
    const int64_t total = B * (int64_t)N;
    const double step = Y > 0 ? 2.0 * ymax / (double)Y : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / N;
        const int pos = (int)(i - f * N);
        const int x = PACKED ? (int)((xp[f * NW + (pos >> 5)] >> (pos & 31)) & 1u) : (int)(xb[i] & 1u);
        double u0, u1;
        uniform2(seed, 2u, (uint64_t)(frame0 + f), (uint32_t)pos, u0, u1);
        // Box-Muller; 1 - u0 is in (0, 1]
        const double g = sqrt(-2.0 * log(1.0 - u0)) * cospi(2.0 * u1);
        const double v = (1.0 - 2.0 * (double)x) + sigma * g;
        if (yf) yf[i] = v;
        if (yq) {
            double q = floor((v + ymax) / step);
            q = q < 0.0 ? 0.0 : (q > (double)(Y - 1) ? (double)(Y - 1) : q);
            yq[i] = (uint8_t)(int)q;
        }
    }
}

// ---- guard bands ---------------------------------------------------------------------------------------------
constexpr int GB_MAXT = 4096;
// out[f][starts[t] .. ) = ones x 1, the t-th sub-word of enc (sub symbols), ones x 1; zeros elsewhere
__global__ void __launch_bounds__(256) add_guard_bands_kernel(const uint8_t *enc, int N, int T, int sub, int ones, const int32_t *starts,
                                                              int total, int64_t B, uint8_t *out) {
    const int64_t all = B * (int64_t)total;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < all; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = i / total;
        const int pos = (int)(i - f * total);
        // the sub-word slot that contains pos, if any: binary search over starts (ascending)
        int lo = 0, hi = T - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (starts[mid] <= pos)
                lo = mid;
            else
                hi = mid - 1;
        }
        const int r = pos - starts[lo];
        uint8_t v = 0;
        if (r >= 0 && r < sub + 2 * ones) v = (r < ones || r >= ones + sub) ? (uint8_t)1 : enc[f * N + (int64_t)lo * sub + (r - ones)];
        out[i] = v;
    }
}

// one warp per frame: symbol i survives iff its uniform draw is >= p (BinaryTrellis.py:455-458); survivors are compacted in order
__global__ void __launch_bounds__(256) deletion_kernel(const uint8_t *in, int len, double p, uint64_t seed, int64_t frame0, int64_t B,
                                                       uint8_t *out, int32_t *out_len) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t f = warp; f < B; f += nwarps) {
        int written = 0;
        for (int base = 0; base < len; base += 32) {
            const int i = base + lane;
            bool keep = false;
            uint8_t v = 0;
            if (i < len) {
                double u, unused;
                uniform2(seed, 3u, (uint64_t)(frame0 + f), (uint32_t)i, u, unused);
                keep = !(u < p);
                v = in[f * len + i];
            }
            const uint32_t m = __ballot_sync(0xffffffffu, keep);
            if (keep) out[f * len + written + __popc(m & ((1u << lane) - 1u))] = v;
            written += __popc(m);
        }
        for (int i = written + lane; i < len; i += 32) out[f * len + i] = 0;
        if (lane == 0) out_len[f] = written;
    }
}

// removeDeletionGuardBands (Guardbands.py:47-63) for one frame per thread: trim the zeros at both edges, split in two halves,
// recurse down to n0; the 2^(n-n0) trimmed sub-words go to sub_bits[f][t][0 .. maxlen), their lengths to sub_len[f][t].
// A sub-word longer than maxlen sets *overflow (the caller sizes maxlen; the trellis decoder rejects longer words anyway).
__global__ void __launch_bounds__(128) remove_guard_bands_kernel(const uint8_t *recv, const int32_t *recv_len, int stride, int n, int n0,
                                                                 int maxlen, int64_t B, uint8_t *sub_bits, int32_t *sub_len,
                                                                 int32_t *overflow) {
    const int depth = n > n0 ? n - n0 : 0, T = 1 << depth;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *a = recv + f * stride;
        int len = recv_len ? recv_len[f] : stride;
        len = len < 0 ? 0 : (len > stride ? stride : len);
        // explicit stack of (lo, hi, level): the right half is pushed first so that the sub-words come out left to right
        int lo_s[24], hi_s[24], lv_s[24];
        int sp = 0, t = 0;
        lo_s[0] = 0, hi_s[0] = len, lv_s[0] = n;
        sp = 1;
        while (sp > 0) {
            --sp;
            int lo = lo_s[sp], hi = hi_s[sp];
            const int lv = lv_s[sp];
            // trimZerosAtEdges, :66-93
            while (lo < hi && a[lo] != 1) ++lo;
            if (lo == hi) {
                hi = lo;
            } else {
                while (a[hi - 1] != 1) --hi;
            }
            if (lv <= n0) {
                const int l = hi - lo;
                if (l > maxlen) atomicExch(overflow, 1);
                const int w = l < maxlen ? l : maxlen;
                uint8_t *o = sub_bits + (f * T + t) * (int64_t)maxlen;
                for (int i = 0; i < w; ++i) o[i] = a[lo + i];
                for (int i = w; i < maxlen; ++i) o[i] = 0;
                sub_len[f * T + t] = l;
                ++t;
            } else {
                const int mid = lo + (hi - lo) / 2;
                lo_s[sp] = mid, hi_s[sp] = hi, lv_s[sp] = lv - 1;
                ++sp;
                lo_s[sp] = lo, hi_s[sp] = mid, lv_s[sp] = lv - 1;
                ++sp;
            }
        }
    }
}

// ---- packed channel symbols (host <-> device traffic): `bits` in {1, 2, 4} bits per symbol, little-endian inside a byte ---------
// one thread per packed 32-bit word: 32 / bits symbols, written as 16-byte stores
__global__ void __launch_bounds__(256) unpack_symbols_kernel(const uint32_t *__restrict__ in, int bits, int64_t words, uint8_t *__restrict__ out) {
    const int per = 32 / bits;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t v = in[w];
        uint4 *o = (uint4 *)(out + w * per);
        for (int g = 0; g < per / 16; ++g) {
            uint32_t q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t b = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) b |= ((v >> ((g * 16 + j * 4 + t) * bits)) & mask) << (8 * t);
                q[j] = b;
            }
            o[g] = make_uint4(q[0], q[1], q[2], q[3]);
        }
        if (per == 8) {  // 4 bits per symbol: eight symbols = two 32-bit stores
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                lo |= ((v >> (t * 4)) & 15u) << (8 * t);
                hi |= ((v >> ((t + 4) * 4)) & 15u) << (8 * t);
            }
            ((uint2 *)(out + w * 8))[0] = make_uint2(lo, hi);
        }
    }
}
__global__ void __launch_bounds__(256) pack_symbols_kernel(const uint8_t *__restrict__ in, int bits, int64_t words, uint32_t *__restrict__ out) {
    const int per = 32 / bits;
    const uint32_t mask = (1u << bits) - 1u;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int t = 0; t < per; ++t) v |= ((uint32_t)in[w * per + t] & mask) << (t * bits);
        out[w] = v;
    }
}

static int grid_for(int64_t items, int block) {
    int64_t g = (items + block - 1) / block;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    return (int)(g > 0 ? g : 1);
}

// positions of the sub-words inside the guarded word (Guardbands.py:23: floor(2^((1 - xi)(m - 1))) zeros between the halves of
// a 2^m block) and the total length
static int guard_layout(int n, int n0, double xi, int ones, std::vector<int32_t> &starts) {
    const int T = 1 << (n > n0 ? n - n0 : 0), sub = (1 << (n < n0 ? n : n0)) + 2 * ones;
    starts.assign(T, 0);
    int64_t pos = 0;
    for (int t = 0; t < T; ++t) {
        if (t) {
            int tz = 0;
            while (!((t >> tz) & 1)) ++tz;
            const int m = n0 + 1 + tz;
            pos += (int64_t)floor(pow(2.0, (1.0 - xi) * (double)(m - 1)));
        }
        starts[t] = (int32_t)pos;
        pos += sub;
    }
    return (int)pos;
}

}  // namespace pc

extern "C" {

int pc_channel_simulate_dmc(const uint8_t *d_x, const uint32_t *d_x_packed, int64_t B, int N, int X, int Y, const double *h_cond,
                            uint64_t seed, int64_t frame0, uint8_t *d_y, void *d_workspace, size_t workspace_bytes, void *stream) {
    using namespace pc;
    PC_REQUIRE(B >= 0 && N >= 1, "bad batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE((d_x != nullptr) != (d_x_packed != nullptr), "exactly one of d_x / d_x_packed must be given");
    PC_REQUIRE(h_cond && d_y, "null buffer");
    PC_REQUIRE(X >= 1 && Y >= 1 && Y <= 256 && X * Y <= CH_MAXTAB, "conditional table too large (X * Y <= 1024, Y <= 256)");
    if (d_x_packed) PC_REQUIRE(X == 2, "packed input is binary");
    PC_REQUIRE(d_workspace && workspace_bytes >= (size_t)X * Y * 8 && ((uintptr_t)d_workspace & 7) == 0, "workspace of X * Y * 8 bytes needed");
    cudaStream_t st = (cudaStream_t)stream;
    PC_CUDA(cudaMemcpyAsync(d_workspace, h_cond, (size_t)X * Y * 8, cudaMemcpyHostToDevice, st));
    const int grid = grid_for(B * (int64_t)N, 256);
    if (d_x_packed)
        dmc_kernel<true><<<grid, 256, (size_t)X * Y * 8, st>>>(nullptr, d_x_packed, N, X, Y, B, frame0, seed, (const double *)d_workspace, d_y);
    else
        dmc_kernel<false><<<grid, 256, (size_t)X * Y * 8, st>>>(d_x, nullptr, N, X, Y, B, frame0, seed, (const double *)d_workspace, d_y);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_channel_simulate_biawgn(const uint8_t *d_x, const uint32_t *d_x_packed, int64_t B, int N, double sigma, uint64_t seed,
                               int64_t frame0, int Y, double ymax, uint8_t *d_y_quantised, double *d_y_real, void *stream) {
    using namespace pc;
    PC_REQUIRE(B >= 0 && N >= 1, "bad batch");
    if (B == 0) return PC_OK;
    PC_REQUIRE((d_x != nullptr) != (d_x_packed != nullptr), "exactly one of d_x / d_x_packed must be given");
    PC_REQUIRE(d_y_quantised || d_y_real, "no output buffer");
    PC_REQUIRE(sigma >= 0.0, "negative sigma");
    if (d_y_quantised) PC_REQUIRE(Y >= 2 && Y <= 256 && ymax > 0.0, "quantiser needs 2 <= Y <= 256 levels and ymax > 0");
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = grid_for(B * (int64_t)N, 256);
    if (d_x_packed)
        awgn_kernel<true><<<grid, 256, 0, st>>>(nullptr, d_x_packed, N, B, frame0, seed, sigma, Y, ymax, d_y_quantised, d_y_real);
    else
        awgn_kernel<false><<<grid, 256, 0, st>>>(d_x, nullptr, N, B, frame0, seed, sigma, Y, ymax, d_y_quantised, d_y_real);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_unpack_symbols(const void *d_packed, int bits, int64_t count, uint8_t *d_symbols, void *stream) {
    using namespace pc;
    PC_REQUIRE(bits == 1 || bits == 2 || bits == 4, "1, 2 or 4 bits per symbol");
    PC_REQUIRE(count >= 0 && (count * bits) % 32 == 0, "the symbol count must fill whole 32-bit words");
    if (count == 0) return PC_OK;
    PC_REQUIRE(d_packed && d_symbols && ((uintptr_t)d_packed & 3) == 0 && ((uintptr_t)d_symbols & 15) == 0, "null or misaligned buffer");
    const int64_t words = count * bits / 32;
    unpack_symbols_kernel<<<grid_for(words, 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t *)d_packed, bits, words, d_symbols);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_pack_symbols(const uint8_t *d_symbols, int bits, int64_t count, void *d_packed, void *stream) {
    using namespace pc;
    PC_REQUIRE(bits == 1 || bits == 2 || bits == 4, "1, 2 or 4 bits per symbol");
    PC_REQUIRE(count >= 0 && (count * bits) % 32 == 0, "the symbol count must fill whole 32-bit words");
    if (count == 0) return PC_OK;
    PC_REQUIRE(d_packed && d_symbols && ((uintptr_t)d_packed & 3) == 0, "null or misaligned buffer");
    const int64_t words = count * bits / 32;
    pack_symbols_kernel<<<grid_for(words, 256), 256, 0, (cudaStream_t)stream>>>(d_symbols, bits, words, (uint32_t *)d_packed);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_guard_band_length(int n, int n0, double xi, int ones) {
    if (n < 0 || n > 24 || n0 < 0 || ones < 0 || n - n0 > 12) return -1;
    std::vector<int32_t> starts;
    return pc::guard_layout(n, n0, xi, ones, starts);
}

int pc_add_guard_bands(const uint8_t *d_encoded, int64_t B, int n, int n0, double xi, int ones, uint8_t *d_out, void *d_workspace,
                       size_t workspace_bytes, void *stream) {
    using namespace pc;
    PC_REQUIRE(B >= 0 && n >= 0 && n <= 24 && n0 >= 0 && ones >= 0 && n - n0 <= 12, "bad guard-band parameters");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_encoded && d_out, "null buffer");
    std::vector<int32_t> starts;
    const int total = guard_layout(n, n0, xi, ones, starts);
    const int T = (int)starts.size(), N = 1 << n, sub = N / T;
    PC_REQUIRE(d_workspace && workspace_bytes >= (size_t)T * 4 && ((uintptr_t)d_workspace & 3) == 0, "workspace of 4 * 2^(n-n0) bytes needed");
    cudaStream_t st = (cudaStream_t)stream;
    // the layout travels through pageable memory of this call's frame: copy synchronously with respect to the host buffer
    PC_CUDA(cudaMemcpyAsync(d_workspace, starts.data(), (size_t)T * 4, cudaMemcpyHostToDevice, st));
    PC_CUDA(cudaStreamSynchronize(st));
    add_guard_bands_kernel<<<grid_for(B * (int64_t)total, 256), 256, 0, st>>>(d_encoded, N, T, sub, ones, (const int32_t *)d_workspace, total, B, d_out);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_deletion_channel(const uint8_t *d_in, int64_t B, int len, double deletion_prob, uint64_t seed, int64_t frame0, uint8_t *d_out,
                        int32_t *d_out_len, void *stream) {
    using namespace pc;
    PC_REQUIRE(B >= 0 && len >= 0, "bad batch");
    if (B == 0 || len == 0) return PC_OK;
    PC_REQUIRE(d_in && d_out && d_out_len, "null buffer");
    PC_REQUIRE(d_in != d_out, "in-place deletion is not supported");
    deletion_kernel<<<grid_for(B * 32, 256), 256, 0, (cudaStream_t)stream>>>(d_in, len, deletion_prob, seed, frame0, B, d_out, d_out_len);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_remove_guard_bands(const uint8_t *d_received, const int32_t *d_received_len, int64_t B, int stride, int n, int n0, int maxlen,
                          uint8_t *d_sub_bits, int32_t *d_sub_len, int32_t *d_overflow, void *stream) {
    using namespace pc;
    PC_REQUIRE(B >= 0 && stride >= 0 && n >= 0 && n <= 24 && n0 >= 0 && n - n0 <= 12 && maxlen >= 1, "bad guard-band parameters");
    if (B == 0) return PC_OK;
    PC_REQUIRE(d_received && d_sub_bits && d_sub_len && d_overflow, "null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    PC_CUDA(cudaMemsetAsync(d_overflow, 0, 4, st));
    remove_guard_bands_kernel<<<grid_for(B, 128), 128, 0, st>>>(d_received, d_received_len, stride, n, n0, maxlen, B, d_sub_bits, d_sub_len,
                                                                 d_overflow);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // extern "C"
