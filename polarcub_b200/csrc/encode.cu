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
                       const uint32_t *frozen_words, uint32_t *out, cudaStream_t st) {
    if (B == 0) return PC_OK;
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

}  // namespace pc

extern "C" {

int pc_encode_bits(const pc_plan *plan, const uint32_t *d_info_packed, uint32_t *d_cw_packed, int64_t B, void *stream) {
    PC_REQUIRE(plan && plan->q == 2, "binary plan required");
    PC_REQUIRE(B >= 0 && (d_cw_packed || B == 0) && (d_info_packed || plan->k == 0 || B == 0), "null buffer");
    return pc::launch_bits(pc::SRC_INFO, plan->n, plan->k, B, d_info_packed, plan->d_src, plan->d_frozen_words,
                           d_cw_packed, (cudaStream_t)stream);
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
