/*
 * oracle/port/port_expf.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * The reference's YOLO post-processing (detectors/yolo3.cpp:157-170) calls the C library's single-precision exp, and its
 * results feed a threshold, a sort and integer truncations, so a GPU version can only be bit-exact if it reproduces that
 * function.  glibc (2.27 and later) evaluates expf in double precision with a 32-entry table of 2^(i/32) and a cubic;
 * restated here in plain IEEE double operations (no fused multiply-add), which is exactly what a CUDA kernel can execute
 * with FP64 instructions.  tests/test_oracle_yolo.py checks it against the C library on tens of millions of arguments.
 * Published algorithm: glibc sysdeps/ieee754/flt-32/e_expf.c + e_exp2f_data.c (ARM optimized-routines), N = 32.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint64_t EXP2F_TAB[32] = {          /* bits(2^(i/32)) - ((i << 52) / 32) */
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,
};

__attribute__((visibility("default")))
float port_expf(float x)
{
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
    if (x != x) return x;
    if (x > 0x1.62e42ep6f) return INFINITY;              /* overflow: x > log(0x1p128) */
    if (x < -0x1.9fe368p6f) return 0.0f;                 /* underflow to zero: x < log(0x1p-150) */
    double z = InvLn2N * (double)x;
    double kd = z + SHIFT;                               /* round to nearest integer, kept in the low mantissa bits */
    uint64_t ki; memcpy(&ki, &kd, 8);
    kd -= SHIFT;
    const double r = z - kd;
    uint64_t t = EXP2F_TAB[ki % 32] + (ki << (52 - 5));
    double s; memcpy(&s, &t, 8);
    z = C0 * r + C1;
    const double r2 = r * r;
    double y = C2 * r + 1.0;
    y = z * r2 + y;
    y = y * s;
    return (float)y;
}

/* number of arguments (every `stride`-th float bit pattern with |x| <= 87) on which port_expf differs from the C library */
__attribute__((visibility("default")))
long port_expf_mismatches(uint32_t stride, long *tested)
{
    long bad = 0, n = 0;
    for (uint64_t u = 0; u < 0xFFFFFFFFull; u += stride) {
        const uint32_t b = (uint32_t)u; float x; memcpy(&x, &b, 4);
        if (!(x == x) || fabsf(x) > 87.0f) continue;
        const float a = expf(x), c = port_expf(x);
        uint32_t ua, uc; memcpy(&ua, &a, 4); memcpy(&uc, &c, 4);
        ++n; if (ua != uc) ++bad;
    }
    if (tested) *tested = n;
    return bad;
}
