// plan.cu -- host-side plan: frozen set, frozen values and the flattened SC tree walk.
//
// Replaces BinaryPolarEncoderDecoder.__init__ / initializeFrozenOrInformationAndRandomlyGeneratedNumbers
// (BinaryPolarEncoderDecoder.py:16-44) and QaryPolarEncoderDecoder.__init__ (QaryPolarEncoderDecoder.py:27-63).
// The Mersenne-Twister stream that defines the frozen bits is drawn by the Python host with the CPython
// stdlib (it must be CPython's random.Random); this file only consumes the resulting values.
#include <cstdarg>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace pc {

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

// out: p->sched (lb < 0: leaf by leaf) or p->sched_b (lb = SC_LB: NODE_BLOCK entries for the mixed / information sub-trees of 2^lb leaves);
// both reference the same r0_words (the block pass re-uses the offsets the first pass assigned, in the same order)
static void build_schedule(pc_plan *p, int i, int l, std::vector<uint32_t> &r0_words, int lb = -1) {
    std::vector<SchedEntry> &out = lb < 0 ? p->sched : p->sched_b;
    const int size = 1 << l;
    bool all_frozen = true;
    for (int j = i; j < i + size; ++j)
        if (!p->frozen_mask[j]) {
            all_frozen = false;
            break;
        }
    SchedEntry e{};
    e.i = i;
    e.l = (int8_t)l;
    e.top = (int8_t)(i == 0 ? p->n : __builtin_ctz((unsigned)i));
    if (all_frozen) {
        e.kind = NODE_RATE0;
        // natural-order codeword of the node: T([a;b]) = [T(a)^T(b), T(b)]
        std::vector<uint8_t> c(p->frozen_vals.begin() + i, p->frozen_vals.begin() + i + size);
        if (p->q == 2) {
            for (int s = 1; s < size; s <<= 1)
                for (int b = 0; b < size; b += 2 * s)
                    for (int j = b; j < b + s; ++j) c[j] ^= c[j + s];
            if (l < 5) {
                uint32_t w = 0;
                for (int j = 0; j < size; ++j) w |= (uint32_t)(c[j] & 1) << j;
                e.bits = w;
            } else {
                e.bits = (uint32_t)r0_words.size();
                for (int wi = 0; wi < size / 32; ++wi) {
                    uint32_t w = 0;
                    for (int j = 0; j < 32; ++j) w |= (uint32_t)(c[wi * 32 + j] & 1) << j;
                    r0_words.push_back(w);
                }
            }
        }
        out.push_back(e);
        return;
    }
    if (l == lb) {
        e.kind = NODE_BLOCK;
        uint32_t m = 0, v = 0;
        for (int j = 0; j < size; ++j) {
            if (p->frozen_mask[i + j]) m |= 1u << j;
            if (p->frozen_mask[i + j] && (p->frozen_vals[i + j] & 1)) v |= 1u << j;
        }
        e.bits = m | (v << size);
        out.push_back(e);
        return;
    }
    if (l == 0) {
        e.kind = NODE_INFO;
        out.push_back(e);
        return;
    }
    build_schedule(p, i, l - 1, r0_words, lb);
    build_schedule(p, i + size / 2, l - 1, r0_words, lb);
}

}  // namespace pc

extern "C" {

int pc_version(void) { return 100; }

const char *pc_last_error(void) { return pc::t_err; }

unsigned long long pc_kernel_launch_count(void) { return pc::g_launches.load(); }

int pc_plan_create(int q, int n, const uint8_t *h_frozen_mask, const uint8_t *h_frozen_vals, pc_plan **out) {
    PC_REQUIRE(out != nullptr, "out is null");
    *out = nullptr;
    PC_REQUIRE(q >= 2 && q <= 16, "alphabet size must be in [2,16]");
    PC_REQUIRE(n >= 0 && n <= 24, "n must be in [0,24]");
    PC_REQUIRE(h_frozen_mask != nullptr && h_frozen_vals != nullptr, "frozen arrays are null");
    pc_plan *p = new (std::nothrow) pc_plan();
    if (!p) {
        pc::set_error("out of host memory");
        return PC_ERR_NOMEM;
    }
    p->q = q;
    p->n = n;
    p->N = 1 << n;
    const int N = p->N;
    p->frozen_mask.assign(h_frozen_mask, h_frozen_mask + N);
    p->frozen_vals.assign(h_frozen_vals, h_frozen_vals + N);
    p->k = 0;
    std::vector<int32_t> src(N);
    for (int i = 0; i < N; ++i) {
        p->frozen_mask[i] = p->frozen_mask[i] ? 1 : 0;
        if (p->frozen_mask[i]) {
            if (p->frozen_vals[i] >= q) {
                delete p;
                pc::set_error("frozen value out of range at index %d", i);
                return PC_ERR_INVALID;
            }
            src[i] = -1;
        } else {
            p->frozen_vals[i] = 0;
            src[i] = p->k++;
        }
    }
    std::vector<uint32_t> r0_words;
    pc::build_schedule(p, 0, n, r0_words);
    if (q == 2 && n > pc::SC_LB) {  // the block schedule visits the same rate-0 nodes in the same order: identical r0_words offsets
        std::vector<uint32_t> r0_again;
        pc::build_schedule(p, 0, n, r0_again, pc::SC_LB);
    }
    const int Nw = (N + 31) / 32;
    std::vector<uint32_t> fw(Nw, 0u);
    for (int i = 0; i < N; ++i)
        if (p->frozen_mask[i] && (p->frozen_vals[i] & 1)) fw[i >> 5] |= 1u << (i & 31);

    // bit-deposit tables of the warp-per-frame encoder: Hacker's Delight 7-5 "expand" masks of each word's information mask
    std::vector<uint32_t> enc_tab((size_t)Nw * 8, 0u);
    {
        uint32_t before = 0;
        for (int w = 0; w < Nw; ++w) {
            uint32_t m = 0;
            for (int b = 0; b < 32 && 32 * w + b < N; ++b)
                if (!p->frozen_mask[32 * w + b]) m |= 1u << b;
            uint32_t *t = &enc_tab[(size_t)w * 8];
            t[0] = m;
            t[1] = before;
            uint32_t mm = m, mk = ~m << 1;
            for (int i = 0; i < 5; ++i) {
                uint32_t mp = mk ^ (mk << 1);
                mp ^= mp << 2;
                mp ^= mp << 4;
                mp ^= mp << 8;
                mp ^= mp << 16;
                const uint32_t mv = mp & mm;
                t[2 + i] = mv;
                mm = (mm ^ mv) | (mv >> (1 << i));
                mk &= ~mp;
            }
            t[7] = fw[w];
            before += (uint32_t)__builtin_popcount(m);
        }
    }

    cudaError_t e = cudaGetDevice(&p->device);
    auto fail = [&](const char *what) {
        pc::set_error("pc_plan_create: %s -> %s", what, cudaGetErrorString(e));
        pc_plan_destroy(p);
        return PC_ERR_CUDA;
    };
    if (e != cudaSuccess) return fail("cudaGetDevice");
#define UP(dst, vec, T)                                                                       \
    do {                                                                                      \
        size_t bytes = sizeof(T) * ((vec).size() ? (vec).size() : 1);                         \
        e = cudaMalloc((void **)&(dst), bytes);                                               \
        if (e != cudaSuccess) return fail("cudaMalloc");                                      \
        if ((vec).size()) {                                                                   \
            e = cudaMemcpy((dst), (vec).data(), sizeof(T) * (vec).size(), cudaMemcpyHostToDevice); \
            if (e != cudaSuccess) return fail("cudaMemcpy");                                  \
        }                                                                                     \
    } while (0)
    UP(p->d_sched, p->sched, pc::SchedEntry);
    if (p->sched_b.size()) UP(p->d_sched_b, p->sched_b, pc::SchedEntry);
    UP(p->d_r0_words, r0_words, uint32_t);
    UP(p->d_src, src, int32_t);
    UP(p->d_frozen_words, fw, uint32_t);
    UP(p->d_frozen_mask, p->frozen_mask, uint8_t);
    UP(p->d_frozen_vals, p->frozen_vals, uint8_t);
    UP(p->d_enc_tab, enc_tab, uint32_t);
    if (q == 2 && n >= 16 && n <= 20) {
        const int G = N >> 10;
        std::vector<uint32_t> tt(enc_tab.size());
        for (int t = 0; t < G; ++t)
            for (int j = 0; j < 32; ++j)
                for (int h = 0; h < 2; ++h)
                    for (int c = 0; c < 4; ++c)
                        tt[((size_t)(h * 32 + j) * G + t) * 4 + c] = enc_tab[(size_t)(t * 32 + j) * 8 + 4 * h + c];
        UP(p->d_enc_tab_t, tt, uint32_t);
    }
#undef UP
    *out = p;
    return PC_OK;
}

void pc_plan_destroy(pc_plan *p) {
    if (!p) return;
    pc::scl_tables_release(p);
    pc::stream_tables_release(p);
    pc::trellis_tables_release(p);
    pc::genie_tables_release(p);
    pc::hybrid_tables_release(p);
    cudaFree(p->d_sched);
    cudaFree(p->d_sched_b);
    cudaFree(p->d_r0_words);
    cudaFree(p->d_src);
    cudaFree(p->d_frozen_words);
    cudaFree(p->d_enc_tab);
    cudaFree(p->d_enc_tab_t);
    cudaFree(p->d_frozen_mask);
    cudaFree(p->d_frozen_vals);
    delete p;
}

int pc_plan_k(const pc_plan *p) { return p ? p->k : PC_ERR_INVALID; }
int pc_plan_length(const pc_plan *p) { return p ? p->N : PC_ERR_INVALID; }
int pc_plan_schedule_len(const pc_plan *p) { return p ? (int)p->sched.size() : PC_ERR_INVALID; }

}  // extern "C"
