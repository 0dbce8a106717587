// encode.cu -- polar encode butterflies (binary bit-packed, q-ary byte symbols) for sm_100a.
//
// Reference semantics (BinaryPolarEncoderDecoder.py:321-323): x[2h] = m[h] ^ p[h], x[2h+1] = p[h] with m, p the
// encodings of the first / second half of u, i.e. x = u B_N F^{(x)n}.  Internally the butterfly runs in
// natural order (T([a;b]) = [T(a)^T(b), T(b)], contiguous halves, friendly to packed words) and the
// bit-reversal permutation B_N is applied once while storing: x_ref[i] = x_nat[rev_n(i)].
// The q-ary map (QaryPolarEncoderDecoder.py:397-399) is T([a;b]) = [T(a)+T(b), -T(b)] mod q.
#include "common.cuh"

namespace pc {

enum : int { SRC_INFO = 0, SRC_WORDS = 1 };

// One frame per group of `tpf` threads (power of two >= 32), `fpb` groups per block, words in shared memory.
template <int SRC>
__global__ void __launch_bounds__(256) encode_bits_kernel(int n, int k, int64_t B, int tpf, const uint32_t *__restrict__ in,
                                                          const int32_t *__restrict__ src,
                                                          const uint32_t *__restrict__ frozen_words,
                                                          uint32_t *__restrict__ out) {
    extern __shared__ uint32_t sm_words[];
    const int N = 1 << n;
    const int Nw = (N + 31) >> 5;
    const int Kw = (k + 31) >> 5;
    const int fpb = blockDim.x / tpf;
    const int grp = threadIdx.x / tpf;
    const int t = threadIdx.x % tpf;
    const int lane = threadIdx.x & 31;
    uint32_t *w = sm_words + (size_t)grp * Nw;
    const int64_t frames_per_iter = (int64_t)gridDim.x * fpb;
    const int64_t iters = (B + frames_per_iter - 1) / frames_per_iter;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t f = it * frames_per_iter + (int64_t)blockIdx.x * fpb + grp;
        const bool live = f < B;
        // ---- 1. u words -----------------------------------------------------------------------------
        if (SRC == SRC_INFO) {
            const uint32_t *info = in + (live ? f : 0) * Kw;
            for (int wi = t >> 5; wi < Nw; wi += tpf >> 5) {
                const int pos = 32 * wi + lane;
                uint32_t bit = 0;
                if (pos < N) {
                    const int s = __ldg(src + pos);
                    bit = s >= 0 ? ((__ldg(info + (s >> 5)) >> (s & 31)) & 1u) : ((__ldg(frozen_words + wi) >> lane) & 1u);
                }
                const uint32_t word = __ballot_sync(0xffffffffu, bit);
                if (lane == 0) w[wi] = word;
            }
        } else {
            const uint32_t *xin = in + (live ? f : 0) * Nw;
            for (int wi = t; wi < Nw; wi += tpf) w[wi] = __ldg(xin + wi);
        }
        __syncthreads();
        // ---- 2. butterfly: strides below 32 inside each word ----------------------------------------
        for (int wi = t; wi < Nw; wi += tpf) {
            uint32_t x = w[wi];
            if (n > 0) x ^= (x >> 1) & 0x55555555u;
            if (n > 1) x ^= (x >> 2) & 0x33333333u;
            if (n > 2) x ^= (x >> 4) & 0x0f0f0f0fu;
            if (n > 3) x ^= (x >> 8) & 0x00ff00ffu;
            if (n > 4) x ^= (x >> 16) & 0x0000ffffu;
            w[wi] = x;
        }
        // strides of whole words
        for (int d = 1; d < Nw; d <<= 1) {
            __syncthreads();
            for (int idx = t; idx < (Nw >> 1); idx += tpf) {
                const int lo = ((idx & ~(d - 1)) << 1) | (idx & (d - 1));
                w[lo] ^= w[lo + d];
            }
        }
        __syncthreads();
        // ---- 3. bit-reversal permutation while storing ----------------------------------------------
        if (live) {
            uint32_t *o = out + f * Nw;
            for (int j = t; j < Nw; j += tpf) o[j] = bitrev_gather_word([&](uint32_t wi) { return w[wi]; }, n, (uint32_t)j);
        }
        __syncthreads();
    }
}

// ---- warp per frame, 2^10 <= N <= 2^15: the whole frame lives in registers (G = N / 1024 words per lane, word lane * G + j
// in register j), 128-bit loads / stores for G >= 4 --------------------------------------------------------------------
//  1. u words: the information bits are deposited at the information positions of each word with a 5-step bit-deposit
//     (Hacker's Delight "expand", masks precomputed per plan) from a funnel-shifted window of the packed information;
//  2. butterfly T([a;b]) = [T(a)^T(b), T(b)]: strides below 32 by shifts and masks, word strides below G between registers,
//     above by __shfl_xor;
//  3. x_ref[i] = x_nat[rev_n(i)]: N / 1024 independent 32 x 32 bit-matrix transposes ACROSS THE LANES of the warp
//     (5 shuffle steps), preceded by one lane permutation; lane c ends up owning G consecutive output words.
template <int SRC, int LOGG>
__global__ void __launch_bounds__(256) encode_warp_kernel(int k, int64_t B, const uint32_t *__restrict__ in,
                                                          const uint32_t *__restrict__ tab, uint32_t *__restrict__ out) {
    constexpr int G = 1 << LOGG, Nw = 32 * G;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int Kw = (k + 31) >> 5;
    const int src_lane = 31 - (int)(__brev((uint32_t)lane) >> 27);
    for (int64_t f = warp; f < B; f += nwarps) {
        uint32_t x[G];
        if (SRC == SRC_INFO) {
            const uint32_t *info = in + f * Kw;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const uint4 t0 = __ldg((const uint4 *)(tab + (size_t)(lane * G + j) * 8));
                const uint4 t1 = __ldg((const uint4 *)(tab + (size_t)(lane * G + j) * 8) + 1);
                const uint32_t m = t0.x, before = t0.y;
                uint32_t v = 0;
                if (m) {
                    const int wi = (int)(before >> 5), sh = (int)(before & 31);
                    const uint32_t lo = __ldg(info + wi), hi = (sh && wi + 1 < Kw) ? __ldg(info + wi + 1) : 0u;
                    v = __funnelshift_r(lo, hi, sh);  // the next 32 information bits
                    // expand: deposit the low popc(m) bits of v at the set bits of m
                    uint32_t tt;
                    tt = v << 16, v = (v & ~t1.z) | (tt & t1.z);
                    tt = v << 8, v = (v & ~t1.y) | (tt & t1.y);
                    tt = v << 4, v = (v & ~t1.x) | (tt & t1.x);
                    tt = v << 2, v = (v & ~t0.w) | (tt & t0.w);
                    tt = v << 1, v = (v & ~t0.z) | (tt & t0.z);
                    v &= m;
                }
                x[j] = v | t1.w;
            }
        } else {
            const uint32_t *xin = in + f * Nw + lane * G;
            if (G >= 4) {
#pragma unroll
                for (int j = 0; j < G; j += 4) {
                    const uint4 q = __ldg((const uint4 *)(xin + j));
                    x[j] = q.x, x[j + 1 < G ? j + 1 : j] = q.y, x[j + 2 < G ? j + 2 : j] = q.z, x[j + 3 < G ? j + 3 : j] = q.w;
                }
            } else {
#pragma unroll
                for (int j = 0; j < G; ++j) x[j] = __ldg(xin + j);
            }
        }
        // butterfly inside the words, between registers, between lanes
#pragma unroll
        for (int j = 0; j < G; ++j) {
            uint32_t v = x[j];
            v ^= (v >> 1) & 0x55555555u;
            v ^= (v >> 2) & 0x33333333u;
            v ^= (v >> 4) & 0x0f0f0f0fu;
            v ^= (v >> 8) & 0x00ff00ffu;
            v ^= (v >> 16) & 0x0000ffffu;
            x[j] = v;
        }
#pragma unroll
        for (int d = 1; d < G; d <<= 1)
#pragma unroll
            for (int j = 0; j < G; ++j)
                if (!(j & d)) x[j] ^= x[j + d < G ? j + d : j];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const uint32_t pv = __shfl_xor_sync(0xffffffffu, x[j], d);
                x[j] ^= (lane & d) ? 0u : pv;
            }
        // bit reversal: lane permutation, then a 32 x 32 bit transpose across the lanes per register
        uint32_t y[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            uint32_t v = __shfl_sync(0xffffffffu, x[j], src_lane);
#pragma unroll
            for (int jj = 0; jj < 5; ++jj) {
                const int s = 16 >> jj;
                const uint32_t msk = jj == 0 ? 0x0000ffffu : jj == 1 ? 0x00ff00ffu : jj == 2 ? 0x0f0f0f0fu : jj == 3 ? 0x33333333u : 0x55555555u;
                const uint32_t pv = __shfl_xor_sync(0xffffffffu, v, s);
                if (lane & s) {
                    const uint32_t t = (pv ^ (v >> s)) & msk;  // this lane is row k + s, the partner row k
                    v ^= t << s;
                } else {
                    const uint32_t t = (v ^ (pv >> s)) & msk;
                    v ^= t;
                }
            }
            // lane c now holds T_j[c]: output word (31 - rev5(c)) * G + rev_LOGG(j)
            constexpr int dummy = 0;
            (void)dummy;
            y[LOGG ? (int)(__brev((uint32_t)j) >> (32 - (LOGG ? LOGG : 1))) : 0] = v;
        }
        uint32_t *o = out + f * Nw + src_lane * G;
        if (G >= 4) {
#pragma unroll
            for (int j = 0; j < G; j += 4)
                *(uint4 *)(o + j) = make_uint4(y[j], y[j + 1 < G ? j + 1 : j], y[j + 2 < G ? j + 2 : j], y[j + 3 < G ? j + 3 : j]);
        } else {
#pragma unroll
            for (int j = 0; j < G; ++j) o[j] = y[j];
        }
    }
}

template <int SRC>
static int launch_warp_encoder(int n, int k, int64_t B, const uint32_t *in, const uint32_t *tab, uint32_t *out, cudaStream_t st) {
    const int64_t want = (B + 7) / 8;
    const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    switch (n) {
        case 10: encode_warp_kernel<SRC, 0><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
        case 11: encode_warp_kernel<SRC, 1><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
        case 12: encode_warp_kernel<SRC, 2><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
        case 13: encode_warp_kernel<SRC, 3><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
        case 14: encode_warp_kernel<SRC, 4><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
        default: encode_warp_kernel<SRC, 5><<<grid, 256, 0, st>>>(k, B, in, tab, out); break;
    }
    PC_LAUNCH_CHECK();
    return PC_OK;
}

// ---- 2^16 <= N <= 2^20: 2^(n-10) threads per frame, 1024 / 2^(n-10) frames per 1024-thread block, the block's 2^20 bits in
// 128 KB of shared memory.  Word index bits: [T (top 5) | M (n-10)] = [thread t (n-10) | j (5)].
//  A. thread t owns the 32 consecutive words t * 32 + j in registers: u words by the same bit-deposit as the warp kernel (or
//     the packed input words); the frame's packed input is first staged in shared memory with coalesced loads and the
//     deposit table is read in a transposed layout (plan->d_enc_tab_t), because per-thread 128-byte windows cost one LSU
//     wavefront per lane; butterfly strides inside the words, between registers (word bits 0-4) and between lanes
//     (word bits 5 .. n-11, shuffles); the words go to shared memory;
//  B. thread M re-reads the 32 words (T, M), T = 0..31: butterfly over the top five word bits between registers, then the
//     bit reversal x_ref[i] = x_nat[rev_n(i)] -- for every M a 32 x 32 bit-matrix transpose (in registers) whose rows land
//     in the words (31 - rev5(c), rev(M)); the rows are staged in shared memory and copied out coalesced.
// Shared-memory word w lives at w ^ ((w >> 5) & 31): phase A's stride-32 stores, phase B's loads and the bit-reversed row
// stores are all bank-conflict free.
// BFLY = false: the bit-reversal permutation alone (x_ref = B_N x_nat), used by the hybrid decoder's egress.
template <int SRC, bool BFLY = true>
__global__ void __launch_bounds__(1024, 1) encode_block_kernel(int n, int k, int64_t B, const uint32_t *__restrict__ in,
                                                               const uint32_t *__restrict__ tab, uint32_t *__restrict__ out) {
    extern __shared__ uint32_t sm_words[];
    const int lg = n - 10, tpf = 1 << lg, Nw = 32 << lg, fpb = 1024 >> lg;
    const int slot = threadIdx.x >> lg, t = threadIdx.x & (tpf - 1), lane = threadIdx.x & 31;
    const int Kw = (k + 31) >> 5;
    uint32_t *w = sm_words + (size_t)slot * Nw;
    const uint32_t rev_m = __brev((uint32_t)t) >> (32 - lg);
    for (int64_t f0 = (int64_t)blockIdx.x * fpb; f0 < B; f0 += (int64_t)gridDim.x * fpb) {
        const int64_t f = f0 + slot;
        const bool live = f < B;
        uint32_t x[32];
        // ---- A ---------------------------------------------------------------------------------------
        // the frame's packed input goes through shared memory first, so that the global reads are coalesced (a thread's own
        // 32 words / information window are 128 bytes apart from its neighbour's)
        if (SRC == SRC_INFO) {
            const uint32_t *info = in + (live ? f : 0) * Kw;
            {   // the next frame of this slot: pull its information words into L2 while this one is encoded
                const int64_t fn = f + (int64_t)gridDim.x * fpb;
                if (fn < B && t * 32 < Kw) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + fn * Kw + t * 32));
            }
#pragma unroll 1
            for (int i0 = t; i0 < Kw; i0 += 8 * tpf) {  // eight independent loads in flight per thread
                uint32_t v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = i0 + u * tpf < Kw ? __ldg(info + i0 + u * tpf) : 0u;
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (i0 + u * tpf < Kw) w[i0 + u * tpf] = v[u];
            }
            __syncthreads();
            const uint4 *tp = (const uint4 *)tab + t;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint4 t0 = __ldg(tp + (size_t)j * tpf), t1 = __ldg(tp + (size_t)(32 + j) * tpf);
                const uint32_t m = t0.x, before = t0.y;
                uint32_t v = 0;
                if (m) {
                    const int wi = (int)(before >> 5), sh = (int)(before & 31);
                    const uint32_t lo = w[wi], hi = (sh && wi + 1 < Kw) ? w[wi + 1] : 0u;
                    v = __funnelshift_r(lo, hi, sh);
                    uint32_t tt;
                    tt = v << 16, v = (v & ~t1.z) | (tt & t1.z);
                    tt = v << 8, v = (v & ~t1.y) | (tt & t1.y);
                    tt = v << 4, v = (v & ~t1.x) | (tt & t1.x);
                    tt = v << 2, v = (v & ~t0.w) | (tt & t0.w);
                    tt = v << 1, v = (v & ~t0.z) | (tt & t0.z);
                    v &= m;
                }
                x[j] = v | t1.w;
            }
            __syncthreads();  // every information window has been read: the buffer now takes the u words
        } else {
            const uint32_t *xin = in + (live ? f : 0) * Nw;
            {
                const int64_t fn = f + (int64_t)gridDim.x * fpb;
                if (fn < B) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + fn * Nw + t * 32));
            }
#pragma unroll 8
            for (int i = t; i < Nw; i += tpf) w[i ^ ((i >> 5) & 31)] = __ldg(xin + i);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = w[(t * 32 + j) ^ lane];
        }
        if (BFLY) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                uint32_t v = x[j];
                v ^= (v >> 1) & 0x55555555u;
                v ^= (v >> 2) & 0x33333333u;
                v ^= (v >> 4) & 0x0f0f0f0fu;
                v ^= (v >> 8) & 0x00ff00ffu;
                v ^= (v >> 16) & 0x0000ffffu;
                x[j] = v;
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (!(j & d)) x[j] ^= x[j | d];
#pragma unroll 1
            for (int d = 1; d < (1 << (n - 15)); d <<= 1) {  // word bits 5 .. n-11 (the lanes); the higher bits belong to phase B
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const uint32_t pv = __shfl_xor_sync(0xffffffffu, x[j], d);
                    x[j] ^= (lane & d) ? 0u : pv;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) w[(t * 32 + j) ^ lane] = x[j];
        __syncthreads();
        // ---- B ---------------------------------------------------------------------------------------
#pragma unroll
        for (int T = 0; T < 32; ++T) {
            const int idx = T * tpf + t;
            x[T] = w[idx ^ ((idx >> 5) & 31)];
        }
        __syncthreads();
        if (BFLY) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
#pragma unroll
                for (int T = 0; T < 32; ++T)
                    if (!(T & d)) x[T] ^= x[T | d];
        }
        uint32_t A[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) A[c] = x[31 - (int)(__brev((uint32_t)c) >> 27)];
        // Hacker's Delight transpose32 (anti-transpose in LSB-first numbering: T[c] bit r = A[31-r] bit 31-c)
#pragma unroll
        for (int jj = 0; jj < 5; ++jj) {
            const int j = 16 >> jj;
            const uint32_t m = jj == 0 ? 0x0000ffffu : jj == 1 ? 0x00ff00ffu : jj == 2 ? 0x0f0f0f0fu : jj == 3 ? 0x33333333u : 0x55555555u;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                if (!(c & j)) {
                    const uint32_t tt = (A[c] ^ (A[c + j] >> j)) & m;
                    A[c] ^= tt;
                    A[c + j] ^= tt << j;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int ow = (31 - (int)(__brev((uint32_t)c) >> 27)) * tpf + (int)rev_m;
            w[ow ^ ((ow >> 5) & 31)] = A[c];
        }
        __syncthreads();
        if (live) {
            uint32_t *o = out + f * Nw;
#pragma unroll 8
            for (int i = t; i < Nw; i += tpf) o[i] = w[i ^ ((i >> 5) & 31)];
        }
        __syncthreads();
    }
}

// q-ary: one frame per block, one byte per symbol in shared memory.
__global__ void __launch_bounds__(256) encode_qary_kernel(int q, int n, int k, int64_t B, const uint8_t *__restrict__ info,
                                                          const int32_t *__restrict__ src,
                                                          const uint8_t *__restrict__ frozen_vals,
                                                          uint8_t *__restrict__ out) {
    extern __shared__ uint8_t sm_sym[];
    const int N = 1 << n;
    for (int64_t f = blockIdx.x; f < B; f += gridDim.x) {
        for (int i = threadIdx.x; i < N; i += blockDim.x) {
            const int s = __ldg(src + i);
            sm_sym[i] = s >= 0 ? __ldg(info + f * k + s) : __ldg(frozen_vals + i);
        }
        for (int d = 1; d < N; d <<= 1) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < (N >> 1); idx += blockDim.x) {
                const int lo = ((idx & ~(d - 1)) << 1) | (idx & (d - 1));
                const int a = sm_sym[lo], b = sm_sym[lo + d];
                int s = a + b;
                s = s >= q ? s - q : s;
                sm_sym[lo] = (uint8_t)s;
                sm_sym[lo + d] = (uint8_t)(b ? q - b : 0);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < N; i += blockDim.x) out[f * N + i] = sm_sym[bitrev_n((uint32_t)i, n)];
        __syncthreads();
    }
}

static int launch_bits(int src_kind, int n, int k, int64_t B, const uint32_t *in, const int32_t *src,
                       const uint32_t *frozen_words, uint32_t *out, cudaStream_t st, const uint32_t *enc_tab = nullptr,
                       const uint32_t *enc_tab_t = nullptr) {
    if (B == 0) return PC_OK;
    const char *ev = getenv("PC_ENCODE_CTA");  // 1: keep the frame-per-CTA kernel (tests compare the two)
    if (n >= 10 && n <= 15 && !(ev && *ev == '1') && ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0) {
        if (src_kind == SRC_WORDS) return launch_warp_encoder<SRC_WORDS>(n, k, B, in, nullptr, out, st);
        if (enc_tab) return launch_warp_encoder<SRC_INFO>(n, k, B, in, enc_tab, out, st);
    }
    if (n >= 16 && n <= 20 && !(ev && *ev == '1') && (src_kind == SRC_WORDS || enc_tab_t)) {
        const int fpb = 1024 >> (n - 10);
        const int64_t want = (B + fpb - 1) / fpb;
        const int grid = (int)(want < (int64_t)num_sms() ? want : (int64_t)num_sms());
        const int smem = 128 * 1024;
        if (src_kind == SRC_WORDS) {
            PC_CUDA(cudaFuncSetAttribute(encode_block_kernel<SRC_WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            encode_block_kernel<SRC_WORDS><<<grid, 1024, smem, st>>>(n, k, B, in, nullptr, out);
        } else {
            PC_CUDA(cudaFuncSetAttribute(encode_block_kernel<SRC_INFO>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            encode_block_kernel<SRC_INFO><<<grid, 1024, smem, st>>>(n, k, B, in, enc_tab_t, out);
        }
        PC_LAUNCH_CHECK();
        return PC_OK;
    }
    const int Nw = ((1 << n) + 31) >> 5;
    const int threads = 256;
    int tpf = 32;
    while (tpf < Nw && tpf < threads) tpf <<= 1;
    const int fpb = threads / tpf;
    const size_t smem = (size_t)fpb * Nw * sizeof(uint32_t);
    PC_REQUIRE(smem <= 200 * 1024, "block length too large for the shared-memory encoder (n <= 20)");
    const int64_t want = (B + fpb - 1) / fpb;
    const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    if (src_kind == SRC_INFO) {
        if (smem > 48 * 1024)
            PC_CUDA(cudaFuncSetAttribute(encode_bits_kernel<SRC_INFO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_bits_kernel<SRC_INFO><<<grid, threads, smem, st>>>(n, k, B, tpf, in, src, frozen_words, out);
    } else {
        if (smem > 48 * 1024)
            PC_CUDA(cudaFuncSetAttribute(encode_bits_kernel<SRC_WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encode_bits_kernel<SRC_WORDS><<<grid, threads, smem, st>>>(n, k, B, tpf, in, src, frozen_words, out);
    }
    PC_LAUNCH_CHECK();
    return PC_OK;
}

// x_ref = B_N x_nat on packed words, [B][N/32] -> [B][N/32], 16 <= n <= 20 (the hybrid decoder's codeword egress)
int bitrev_words_launch(int n, int64_t B, const uint32_t *in, uint32_t *out, cudaStream_t st) {
    if (B == 0) return PC_OK;
    PC_REQUIRE(n >= 16 && n <= 20, "bitrev_words_launch: 16 <= n <= 20");
    const int fpb = 1024 >> (n - 10);
    const int64_t want = (B + fpb - 1) / fpb;
    const int grid = (int)(want < (int64_t)num_sms() ? want : (int64_t)num_sms());
    const int smem = 128 * 1024;
    PC_CUDA(cudaFuncSetAttribute(encode_block_kernel<SRC_WORDS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    encode_block_kernel<SRC_WORDS, false><<<grid, 1024, smem, st>>>(n, 0, B, in, nullptr, out);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // namespace pc

extern "C" {

int pc_encode_bits(const pc_plan *plan, const uint32_t *d_info_packed, uint32_t *d_cw_packed, int64_t B, void *stream) {
    PC_REQUIRE(plan && plan->q == 2, "binary plan required");
    PC_REQUIRE(B >= 0 && (d_cw_packed || B == 0) && (d_info_packed || plan->k == 0 || B == 0), "null buffer");
    return pc::launch_bits(pc::SRC_INFO, plan->n, plan->k, B, d_info_packed, plan->d_src, plan->d_frozen_words,
                           d_cw_packed, (cudaStream_t)stream, plan->d_enc_tab, plan->d_enc_tab_t);
}

int pc_polar_transform_bits(int n, const uint32_t *d_cw_packed, uint32_t *d_u_packed, int64_t B, void *stream) {
    PC_REQUIRE(n >= 0 && n <= 20, "n must be in [0,20]");
    PC_REQUIRE(B >= 0 && ((d_cw_packed && d_u_packed) || B == 0), "null buffer");
    return pc::launch_bits(pc::SRC_WORDS, n, 0, B, d_cw_packed, nullptr, nullptr, d_u_packed, (cudaStream_t)stream);
}

int pc_qsc_encode(const pc_plan *plan, const uint8_t *d_info, uint8_t *d_cw, int64_t B, void *stream) {
    PC_REQUIRE(plan != nullptr, "plan is null");
    PC_REQUIRE(B >= 0 && (d_cw || B == 0) && (d_info || plan->k == 0 || B == 0), "null buffer");
    PC_REQUIRE(plan->n <= 16, "q-ary encoder supports n <= 16");
    if (B == 0) return PC_OK;
    const size_t smem = (size_t)plan->N;
    if (smem > 48 * 1024)
        PC_CUDA(cudaFuncSetAttribute(pc::encode_qary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t cap = (int64_t)pc::num_sms() * 8;
    const int grid = (int)(B < cap ? B : cap);
    pc::encode_qary_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(plan->q, plan->n, plan->k, B, d_info, plan->d_src,
                                                                      plan->d_frozen_vals, d_cw);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

}  // extern "C"
