// util.cu -- error counters for Monte-Carlo runs and the dominant-kernel timing hook used by bench.py.
//
// The reference counts frame errors serially on the host (BinaryPolarEncoderDecoder.py:374-387,
// QaryPolarEncoderDecoder.py:907-909); here one kernel reduces {frames, frame errors, bit errors} per rank
// and the host all-reduces the three int64 over NCCL.
#include <algorithm>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace pc {

__global__ void __launch_bounds__(256) count_errors_kernel(const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                                                           int64_t B, int W, int nbits, unsigned long long *out) {
    unsigned long long ferr = 0, berr = 0;
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < B; f += (int64_t)gridDim.x * blockDim.x) {
        int bits = 0;
        for (int w = 0; w < W; ++w) {
            uint32_t x = a[f * W + w] ^ b[f * W + w];
            if (w == W - 1 && (nbits & 31)) x &= (1u << (nbits & 31)) - 1u;
            bits += __popc(x);
        }
        berr += bits;
        ferr += bits != 0;
    }
    for (int o = 16; o > 0; o >>= 1) {
        ferr += __shfl_xor_sync(0xffffffffu, ferr, o);
        berr += __shfl_xor_sync(0xffffffffu, berr, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 1, ferr);
        atomicAdd(out + 2, berr);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out + 0, (unsigned long long)B);
}

// ---- timing hook: CUDA events around each launch of the dominant kernel, on the launching stream --------
static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_prof_ev;  // begin/end pairs
static size_t g_prof_used = 0;

static std::atomic<int> g_prof_suspend{0};
void prof_suspend(int delta) { g_prof_suspend.fetch_add(delta); }

void prof_mark(cudaStream_t st) {
    if (!g_prof_on || g_prof_suspend.load() > 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (g_prof_used == g_prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        g_prof_ev.push_back(e);
    }
    cudaEventRecord(g_prof_ev[g_prof_used++], st);
}

}  // namespace pc

extern "C" {

int pc_count_errors(const uint32_t *d_a, const uint32_t *d_b, int64_t B, int nbits, unsigned long long *d_out3,
                    void *stream) {
    PC_REQUIRE(B >= 0 && nbits >= 0 && d_out3, "bad arguments");
    if (B == 0 || nbits == 0) return PC_OK;
    PC_REQUIRE(d_a && d_b, "null buffer");
    const int W = (nbits + 31) / 32;
    int64_t blocks = (B + 255) / 256;
    if (blocks > pc::num_sms() * 8) blocks = pc::num_sms() * 8;
    pc::count_errors_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, B, W, nbits, d_out3);
    PC_LAUNCH_CHECK();
    return PC_OK;
}

int pc_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(pc::g_prof_mu);
    pc::g_prof_on = on != 0;
    pc::g_prof_used = 0;
    return PC_OK;
}

// total_ms: the time during which at least one of the marked launches was running -- the union of their [begin, end] intervals.
// Launches on one stream follow each other and the union is the sum of their durations; launches that a caller spreads over
// several streams overlap (the CTAs of the next one start on the SMs the previous one has left) and are not counted twice.
int pc_profile_read(double *total_ms, unsigned long long *launches) {
    std::lock_guard<std::mutex> lk(pc::g_prof_mu);
    std::vector<std::pair<double, double>> iv;
    for (size_t i = 0; i + 1 < pc::g_prof_used; i += 2) {
        float b = 0, e = 0;
        PC_CUDA(cudaEventSynchronize(pc::g_prof_ev[i + 1]));
        PC_CUDA(cudaEventSynchronize(pc::g_prof_ev[i]));
        PC_CUDA(cudaEventElapsedTime(&b, pc::g_prof_ev[0], pc::g_prof_ev[i]));  // may be negative: another stream started earlier
        PC_CUDA(cudaEventElapsedTime(&e, pc::g_prof_ev[i], pc::g_prof_ev[i + 1]));
        iv.emplace_back((double)b, (double)b + (double)e);
    }
    std::sort(iv.begin(), iv.end());
    double tot = 0, hi = -1e300;
    for (const auto &v : iv) {
        const double lo = v.first > hi ? v.first : hi;
        if (v.second > lo) tot += v.second - lo;
        if (v.second > hi) hi = v.second;
    }
    if (total_ms) *total_ms = tot;
    if (launches) *launches = (unsigned long long)iv.size();
    return PC_OK;
}

}  // extern "C"
