// qlog_arith.cuh -- log-domain arithmetic of the q-ary decoders (sc_qary.cu, scl.cu): the `if self.use_log` branches of
// QaryMemorylessVectorDistribution (VectorDistributions/QaryMemorylessVectorDistribution.py:31-42, :50-62, :74-118) and of
// QaryPolarEncoderDecoder.listDecode / recursiveListDecode (QaryPolarEncoderDecoder.py:140, :436-566, :765, :784, :813, :869).
#pragma once
#include <math.h>

#include "common.cuh"

namespace pc {

// ---- log domain (QaryPolarEncoderDecoder(..., use_log=True)): numpy.logaddexp and scipy.special.logsumexp restated with the
// device's exp / log1p / log.  Those differ from the host libm in the last bit now and then, so the log domain is checked to a
// tolerance (decisions identical on the golden vectors, metrics within 1e-12), not bit for bit.
__device__ __forceinline__ double d_logaddexp(const double x, const double y) {  // npy_logaddexp
    if (x == y) return x + 0.693147180559945309417232121458176568;
    const double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp;
}
__device__ __forceinline__ double d_log_small_int(const int m) {  // log(m), m = 1 .. 8, correctly rounded
    switch (m) {
        case 1: return 0.0;
        case 2: return 0.6931471805599453;
        case 3: return 1.0986122886681098;
        case 4: return 1.3862943611198906;
        case 5: return 1.6094379124341003;
        case 6: return 1.791759469228055;
        case 7: return 1.9459101090932196;
        default: return 2.0794415416798357;
    }
}
// scipy.special.logsumexp of Q values (scipy >= 1.15): the m maximal elements leave the sum,
// log1p(sum_{others} exp(a - max) / m) + log(m) + max; an infinite result falls back to log(sum(exp(a)))
template <int Q>
__device__ __forceinline__ double q_logsumexp(const double (&a)[Q]) {
    double mx = a[0];
#pragma unroll
    for (int x = 1; x < Q; ++x) mx = a[x] > mx ? a[x] : mx;
    int m = 0;
    double s = 0.0;
#pragma unroll
    for (int x = 0; x < Q; ++x) {
        if (a[x] == mx)
            ++m;
        else
            s += exp(a[x] - mx);
    }
    if (s != 0.0) s = s / (double)m;
    double out = log1p(s) + d_log_small_int(m) + mx;
    if (!isfinite(out)) {
        double t = 0.0;
#pragma unroll
        for (int x = 0; x < Q; ++x) t += exp(a[x]);
        out = log(t);
    }
    return out;
}
// f / g / normalise in the log domain, QaryMemorylessVectorDistribution.py:31-42, :50-62, :104-118 (`if self.use_log` branches)
template <int Q>
__device__ __forceinline__ void q_node_log(const double (&a)[Q], const double (&b)[Q], bool isg, int u1, double (&d)[Q]) {
    if (!isg) {
#pragma unroll
        for (int x = 0; x < Q; ++x) d[x] = -INFINITY;
#pragma unroll
        for (int x1 = 0; x1 < Q; ++x1)
#pragma unroll
            for (int x2 = 0; x2 < Q; ++x2) d[(x1 + x2) % Q] = d_logaddexp(d[(x1 + x2) % Q], a[x1] + b[x2]);
    } else {
#pragma unroll
        for (int u2 = 0; u2 < Q; ++u2) {
            double av = a[u2 % Q];
#pragma unroll
            for (int r = 1; r < Q; ++r) av = u1 == r ? a[(u2 + r) % Q] : av;
            d[u2] = d_logaddexp(-INFINITY, av + b[(Q - u2) % Q]);
        }
    }
    const double t = q_logsumexp<Q>(d);
    if (t != -INFINITY) {
#pragma unroll
        for (int x = 0; x < Q; ++x) d[x] -= t;
    }
}
// np.sum of n float64 values val(0) .. val(n-1), n a power of two: numpy's pairwise summation (umath loops_utils.h.src):
// below 8 values a running sum from 0, up to 128 values eight interleaved accumulators combined as
// ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7)), above that the two halves recursively (a binary tree over 128-blocks)
template <class F>
__device__ __forceinline__ double np_sum_pow2(F val, const int n) {
    if (n < 8) {
        double r = 0.0;
        for (int j = 0; j < n; ++j) r += val(j);
        return r;
    }
    const int bs = n < 128 ? n : 128, nb = n / bs;
    double stack[16];
    int sp = 0;
    for (int b = 0; b < nb; ++b) {
        double r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = val(b * bs + j);
        for (int i = 8; i < bs; i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] += val(b * bs + i + j);
        }
        stack[sp++] = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (int t = b + 1; (t & 1) == 0 && sp >= 2; t >>= 1) {
            stack[sp - 2] = stack[sp - 2] + stack[sp - 1];
            --sp;
        }
    }
    return stack[0];
}

}  // namespace pc
