// Natural logarithm with a fully specified operation sequence: the main path of fdlibm's e_log (argument
// reduction x = 2^k (1 + f), s = f / (2 + f), degree-14 polynomial in s) written with IEEE-754 +, -, *, / only, in a
// fixed order, without fused multiply-adds.  Two builds that keep this order (gcc for x86-64 without FMA
// contraction, nvcc with -fmad=false) return bit-identical results, which is what makes the "exact" superpixel mode
// reproduce the scalar oracle's decisions bit for bit (the reference itself calls CUDA's device log(), whose bits no
// CPU library reproduces).  Domain: positive, finite, normal x (here x >= 2 pi / 12).  Error < 1 ulp.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define DETLOG_HD __host__ __device__ __forceinline__
#else
#define DETLOG_HD inline
#endif

DETLOG_HD double det_log(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint64_t bits;
#if defined(__CUDA_ARCH__)
    bits = (uint64_t)__double_as_longlong(x);
#else
    std::memcpy(&bits, &x, 8);
#endif
    int32_t hx = (int32_t)(bits >> 32);
    const uint32_t lx = (uint32_t)bits;
    int32_t k = (hx >> 20) - 1023;
    hx &= 0x000fffff;
    int32_t i = (hx + 0x95f64) & 0x100000;
    const uint64_t mbits = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | lx;  // normalise x or x / 2
    double m;
#if defined(__CUDA_ARCH__)
    m = __longlong_as_double((long long)mbits);
#else
    std::memcpy(&m, &mbits, 8);
#endif
    k += i >> 20;
    const double f = m - 1.0;
#if defined(__CUDA_ARCH__)
    // correctly rounded quotient without the full division sequence: y = RN(1 / t), q = RN(f y), r = f - t q (exact),
    // RN(q + r y) = RN(f / t) (Markstein) - the same bits as the IEEE division of the host build
    const double t = 2.0 + f, y = __drcp_rn(t), q0 = f * y;
    const double s = fma(fma(-t, q0, f), y, q0);
#else
    const double s = f / (2.0 + f);
#endif
    const double dk = (double)k;
    const double z = s * s;
    i = hx - 0x6147a;
    const double w = z * z;
    const int32_t j = 0x6b851 - hx;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    const double R = t2 + t1;
    if (i > 0) {
        const double hfsq = (0.5 * f) * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}
