// fhog_common.cuh -- per-pixel pieces shared by the fused kernels (kcf_fused.cuh) and the any-size path (kcf_generic.cu).
#pragma once
#include "mot_internal.h"

namespace mot {

// gray = (float)(0.144*B + 0.587*G + 0.299*R), evaluated in double by the reference (top/drawlib.c:234; yes, 0.144).
// For 8-bit channels the double result is never within 3e-11 (relative) of a float rounding boundary, while the double
// evaluation is within 5e-16 of N/1000 with N = 144B + 587G + 299R, so the reference value IS the correctly rounded float
// of N/1000 (checked against the C expression for all 2^24 colours in tests/test_cabi_exports.py).  N < 2^24 is exact in
// float; one Newton step on q0 = N * RN(1/1000) gives the correctly rounded quotient without touching the FP64 pipe.
__device__ __forceinline__ float bgr_gray(const uint8_t *p)
{
    // three integer multiply-adds (FMA pipe); written as PTX because the compiler otherwise packs the bytes with three
    // permutes to feed a dot-product instruction, which costs two instructions more on the busier ALU pipe
    unsigned n = 144u * p[0];
    asm("mad.lo.u32 %0, %1, 587, %0;" : "+r"(n) : "r"((unsigned)p[1]));
    asm("mad.lo.u32 %0, %1, 299, %0;" : "+r"(n) : "r"((unsigned)p[2]));
    const float nf = (float)n;
    const float rcp = 1.0f / 1000.0f;
    const float q0 = __fmul_rn(nf, rcp);
    const float rem = __fmaf_rn(-q0, 1000.0f, nf);
    return __fmaf_rn(rem, rcp, q0);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }


// One pixel of gradMag + gradQuantize (libhog/gradientMex.cpp:59-100, :112-145) given its two gradient components:
// magnitude through the host-harvested rsqrtps / rcpps tables, orientation bin through the acos/quantiser step table.
// Returns M0 = M/16 and writes the bin.  rs_tab / rc_tab / bn_tab may live in shared or global memory.
__device__ __forceinline__ float grad_pixel(float gx, float gy, const float *rs_tab, const float *rc_tab, const uint32_t *bn_tab,
                                            const FhogTablesDev &tab, int *bin_out)
{
    const float m2 = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
    float m;
    {
        const uint32_t u = __float_as_uint(m2);
        if (u < 0x00800000u) m = 1e10f;                       // rsqrtps(+0 / denormal) = +inf -> MIN(.,1e10f)
        else {
            const int e = (int)(u >> 23) - 127;
            const uint32_t key = (u & 0x7FFFFFu) >> (23 - tab.rsqrt_bits);
            const uint32_t T = __float_as_uint(rs_tab[((e & 1) << tab.rsqrt_bits) + key]);
            m = __uint_as_float(T - ((uint32_t)(e >> 1) << 23));
            m = (m < 1e10f) ? m : 1e10f;
        }
    }
    float M;
    {
        const uint32_t u = __float_as_uint(m);
        const int e = (int)(u >> 23) - 127;
        const uint32_t U = __float_as_uint(rc_tab[(u & 0x7FFFFFu) >> (23 - tab.rcp_bits)]);
        M = __uint_as_float(U - ((uint32_t)e << 23));
    }
    float gn = __fmul_rn(__fmul_rn(gx, m), 10000.0f);
    gn = __uint_as_float(__float_as_uint(gn) ^ (__float_as_uint(gy) & 0x80000000u));
    int ai = __float2int_rz(gn) + 10010;
    ai = clampi(ai, 0, 20019);
    const uint32_t ent = bn_tab[(gy < 0.f ? tab.bin_nseg : 0) + (ai >> tab.bin_shift)];
    int bb = (int)(ent & 0xFFu) - ((uint32_t)ai >= (ent >> 8) ? 1 : 0);
    if (bb >= 18) bb = 0;
    *bin_out = bb;
    return __fmul_rn(M, 0.0625f);                             // M0 = M * (1/bin^2), gradientMex.cpp:143
}

// ---- the same pixel step with the table indexing folded into a few integer ops (used by the fused kernels) ---------------
struct LutConsts { int sh_rs; uint32_t mask_rs, flip_rs; int bin_shift, bin_nseg; uint32_t u_cap, m0_cap_bits; };

__device__ __forceinline__ LutConsts make_lut_consts(const FhogTablesDev &t)
{
    LutConsts k;
    k.sh_rs = 23 - t.rsqrt_bits; k.mask_rs = (2u << t.rsqrt_bits) - 1u; k.flip_rs = 1u << t.rsqrt_bits;   // [exponent LSB | top mantissa bits], parity of (e - 127)
    k.bin_shift = t.bin_shift; k.bin_nseg = t.bin_nseg;
    k.u_cap = t.u_cap; k.m0_cap_bits = __float_as_uint(__fmul_rn(t.rcp_cap, 0.0625f));
    return k;
}

// One pixel of gradMag + gradQuantize for the fused kernel: returns M/16 (gradientMex.cpp:83-84, :192) with the orientation
// bin (:90-97, :130-131) in its five low mantissa bits, which the rcpps table leaves zero (fhog_tables.cpp).
__device__ __forceinline__ uint32_t grad_pixel_k(float gx, float gy, const float2 *rsrc_tab, const uint32_t *bn2_tab, const LutConsts &k)
{
    const float m2 = __fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy));
    // rsqrtps: table[parity(e)][top mantissa bits] * 2^-q, q = e >> 1 = ((biased + 1) >> 1) - 64 (e = unbiased exponent).
    // rcpps of that value: rcp(T * 2^-q) = rcp(T) * 2^q exactly, and rcp(T) / 16 sits next to T in the fused table.
    const uint32_t u = __float_as_uint(m2);
    const float2 tr = rsrc_tab[((u >> k.sh_rs) & k.mask_rs) ^ k.flip_rs];
    // q << 23 = qs - 0x20000000 with qs as below; the constant is folded into the device copy of the table (mot_capi.cu)
    const uint32_t qs = ((u + 0x00800000u) >> 1) & 0x7F800000u;
    float m = __uint_as_float(__float_as_uint(tr.x) - qs);
    uint32_t M0 = __float_as_uint(tr.y) + qs;
    if (u <= k.u_cap) { m = 1e10f; M0 = k.m0_cap_bits; }                  // MIN(rsqrtps(M2), 1e10f) saturates (zero / denormal / tiny M2)
    float gn = __fmul_rn(__fmul_rn(gx, m), 10000.0f);
    gn = __uint_as_float(__float_as_uint(gn) ^ (__float_as_uint(gy) & 0x80000000u));
    // |gn| <= 10004 for finite input, so the index needs no clamp; the unsigned min only keeps garbage input inside the table
    const uint32_t ai = min((uint32_t)(__float2int_rz(gn) + 10010), 20019u);
    const uint32_t ent = bn2_tab[(gy < 0.f ? k.bin_nseg : 0) + (ai >> k.bin_shift)];
    const uint32_t bb = (ai >= (ent >> 10)) ? ((ent >> 5) & 31u) : (ent & 31u);
    return M0 | bb;
}

}  // namespace mot
