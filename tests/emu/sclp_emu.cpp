// sclp_emu.cpp -- TEST INFRASTRUCTURE ONLY: the SC-list kernel sources (polarcub_b200/csrc/scl_path.cu + scl_tables.cu)
// compiled for the CPU through tests/emu/cuda_emu.h, so that their logic can be checked against the oracle without a GPU
// (tests/test_emu_scl.py).  Build: tests/emu/build.py.  The product never loads this library.
#include <cstdarg>

#include "../../polarcub_b200/csrc/scl_tables.cu"
#include "../../polarcub_b200/csrc/scl_path.cu"

namespace pc {
static char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
}
static int g_sms = 2;
int num_sms() { return g_sms; }
void prof_mark(cudaStream_t) {}
void prof_suspend(int) {}
}  // namespace pc

extern "C" {

const char *emu_last_error(void) { return pc::t_err; }
void emu_set_sms(int v) { pc::g_sms = v; }

// listDecode of a batch through the emulated kernels; byte-per-symbol buffers like pc_scl_decode_probs, or symbols + table
int emu_sclp_decode(int n, int L, const uint8_t *frozen_mask, const double *xy, const uint8_t *y, const double *table, int Y,
                    const uint8_t *fv, const uint8_t *ainfo, int64_t B, uint8_t *info, int32_t *res, int32_t *lsize, double *lprob,
                    double *aprob, uint8_t *linfo) {
    pc_plan plan;
    plan.q = 2, plan.n = n, plan.N = 1 << n, plan.k = 0, plan.device = 0;
    plan.frozen_mask.assign(frozen_mask, frozen_mask + plan.N);
    plan.frozen_vals.assign(plan.N, 0);
    for (int i = 0; i < plan.N; ++i) plan.k += !plan.frozen_mask[i];
    pc::SclTables *T = pc::scl_tables(&plan);
    if (!T) return -100;
    int rc;
    const int k = plan.k, kw = (k + 31) / 32, nf = plan.N - k, nfw = (nf + 31) / 32;
    if (xy) {
        const size_t need = pc::sclp_workspace_bytes(&plan, L, B, lsize != nullptr);
        std::vector<char> ws(need + 512);
        char *base = (char *)(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
        rc = pc::sclp_decode_bytes(&plan, T, L, xy, fv, ainfo, B, info, res, lsize, lprob, aprob, linfo, base, need, 0);
    } else {
        // packed ABI with channel symbols
        std::vector<uint32_t> ai((size_t)B * (kw ? kw : 1), 0), fp((size_t)B * (nfw ? nfw : 1), 0), oi((size_t)B * (kw ? kw : 1), 0);
        std::vector<uint32_t> ol((size_t)(lsize ? B * L : 1) * (kw ? kw : 1), 0);
        for (int64_t b = 0; b < B; ++b) {
            for (int j = 0; j < k; ++j) ai[b * kw + (j >> 5)] |= (uint32_t)(ainfo[b * k + j] & 1) << (j & 31);
            if (fv)
                for (int j = 0; j < nf; ++j) fp[b * nfw + (j >> 5)] |= (uint32_t)(fv[b * nf + j] & 1) << (j & 31);
        }
        const size_t need = pc::sclp_workspace_bytes_packed(&plan, L, B, lsize != nullptr);
        std::vector<char> ws(need + 512);
        char *base = (char *)(((uintptr_t)ws.data() + 255) & ~(uintptr_t)255);
        rc = pc::sclp_decode_packed(&plan, T, L, nullptr, y, table, Y, fv ? fp.data() : nullptr, ai.data(), B, oi.data(), res, lsize,
                                    lprob, aprob, (lsize && linfo) ? ol.data() : nullptr, base, need, 0);
        for (int64_t b = 0; b < B; ++b) {
            for (int j = 0; j < k; ++j) info[b * k + j] = (oi[b * kw + (j >> 5)] >> (j & 31)) & 1;
            if (lsize && linfo)
                for (int t = 0; t < L; ++t)
                    for (int j = 0; j < k; ++j) linfo[(b * L + t) * k + j] = (ol[(b * L + t) * kw + (j >> 5)] >> (j & 31)) & 1;
        }
    }
    pc::scl_tables_release(&plan);
    return rc;
}
}
