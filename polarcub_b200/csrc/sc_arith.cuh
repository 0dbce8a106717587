// sc_arith.cuh -- the binary SC node arithmetic on the packed (ratio, side) representation, float64, bit-identical to
// BinaryMemorylessVectorDistribution (minusTransform :15-29, plusTransform :31-47, normalize :71-87).
//
// After the reference's max-normalisation a probability pair is (1, r) or (r, 1) with r = min/max, or (0, 0) after
// contradicting hard knowledge.  One float64 holds r, its sign bit says which side is 1, NaN encodes (0, 0).  Every
// product / sum / quotient the reference performs on the pair is reproduced on r with the same IEEE-754 roundings
// (multiplications by exactly 1.0 are dropped -- they are exact), so decisions and probabilities are bit-identical.
#pragma once
#include "common.cuh"

namespace pc {

__device__ __forceinline__ double d_abs(double x) { return __longlong_as_double(__double_as_longlong(x) & 0x7fffffffffffffffLL); }
__device__ __forceinline__ uint32_t d_sign(double x) { return (uint32_t)(__double2hiint(x)) >> 31; }
__device__ __forceinline__ double d_pack(double r, uint32_t side) {
    return __hiloint2double((__double2hiint(r) & 0x7fffffff) | (int)(side << 31), __double2loint(r));
}

// normalise a raw pair by its maximum (BinaryMemorylessVectorDistribution.py:71-87) and pack it.
// 0/0 -> NaN encodes the (0,0) state the reference keeps when the maximum is 0.
__device__ __forceinline__ double pack_pair(double o0, double o1) {
    const bool gt = o1 > o0;
    const double mx = gt ? o1 : o0, mn = gt ? o0 : o1;
    return d_pack(mn / mx, gt ? 1u : 0u);
}

// f on raw pairs (channel level), BinaryMemorylessVectorDistribution.py:21-26
__device__ __forceinline__ double f_raw(double a0, double a1, double b0, double b1) {
    const double o0 = __dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1));
    const double o1 = __dadd_rn(__dmul_rn(a0, b1), __dmul_rn(a1, b0));
    return pack_pair(o0, o1);
}
// g on raw pairs, BinaryMemorylessVectorDistribution.py:37-44
__device__ __forceinline__ double g_raw(double a0, double a1, double b0, double b1, uint32_t u) {
    const double o0 = __dmul_rn(u ? a1 : a0, b0);
    const double o1 = __dmul_rn(u ? a0 : a1, b1);
    return pack_pair(o0, o1);
}
// f on packed normalised values: (o0,o1) is (1 + ra*rb, ra + rb), or swapped when the sides differ
__device__ __forceinline__ double f_packed(double a, double b) {
    const double ra = d_abs(a), rb = d_abs(b);
    const uint32_t s = d_sign(a) ^ d_sign(b);
    const double A = __dadd_rn(1.0, __dmul_rn(ra, rb));
    const double Bv = __dadd_rn(ra, rb);
    const bool c = Bv > A, d = A > Bv;
    const double mx = c ? Bv : A, mn = c ? A : Bv;
    return d_pack(mn / mx, (s ? d : c) ? 1u : 0u);
}
// g on packed normalised values
__device__ __forceinline__ double g_packed(double a, double b, uint32_t u) {
    const double ra = d_abs(a), rb = d_abs(b);
    const uint32_t sa = d_sign(a) ^ u, sb = d_sign(b);
    if (sa == sb) {  // (1*1, ra*rb): already normalised (division by 1.0 is exact)
        const double r = __dmul_rn(ra, rb);
        return d_pack(r, (sb && r < 1.0) ? 1u : 0u);
    }
    // out[sb] = ra, out[1-sb] = rb
    const bool c = ra > rb;
    const double mx = c ? ra : rb, mn = c ? rb : ra;
    const bool gt = sb ? (ra > rb) : (rb > ra);
    return d_pack(mn / mx, gt ? 1u : 0u);
}


// ---- exact shortcuts for hard / erased values --------------------------------------------------------------------
// When both operands are r = 0 (hard knowledge) or r = 1 (erasure) -- every value of a BEC frame -- the reference's
// products, sums and quotients are exact small integers, so the result is selected without the division.  Same
// bits as f_packed / g_packed; anything else (including the 0/0 contradiction case) takes the general path.
// out-of-line general paths: a real call keeps the compiler from evaluating the division speculatively
static __device__ __noinline__ double f_packed_call(double a, double b) { return f_packed(a, b); }
static __device__ __noinline__ double g_packed_call(double a, double b, uint32_t u) { return g_packed(a, b, u); }
__device__ __forceinline__ bool d_is01(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x) & 0x7fffffffffffffffULL;
    return b == 0ULL || b == 0x3ff0000000000000ULL;
}
__device__ __forceinline__ double f_packed01(double a, double b) {
    if (d_is01(a) && d_is01(b)) {
        const double ra = d_abs(a), rb = d_abs(b);
        // (A, B) = (1 + ra rb, ra + rb) is (1,0), (1,1) or (2,2): min/max = 0, 1, 1; A > B only for ra = rb = 0
        const bool both0 = ra == 0.0 && rb == 0.0;
        return d_pack(both0 ? 0.0 : 1.0, ((d_sign(a) ^ d_sign(b)) && both0) ? 1u : 0u);
    }
    return f_packed_call(a, b);
}
__device__ __forceinline__ double g_packed01(double a, double b, uint32_t u) {
    if (d_is01(a) && d_is01(b)) {
        const double ra = d_abs(a), rb = d_abs(b);
        const uint32_t sa = d_sign(a) ^ u, sb = d_sign(b);
        if (sa == sb) {
            const double r = (ra == 1.0 && rb == 1.0) ? 1.0 : 0.0;  // ra * rb
            return d_pack(r, (sb && r < 1.0) ? 1u : 0u);
        }
        if (ra != 0.0 || rb != 0.0) {  // (ra, rb) = (0,0) with opposite sides is the 0/0 contradiction: general path
            const bool gt = sb ? (ra > rb) : (rb > ra);
            return d_pack(ra == rb ? 1.0 : 0.0, gt ? 1u : 0u);
        }
    }
    return g_packed_call(a, b, u);
}

}  // namespace pc
