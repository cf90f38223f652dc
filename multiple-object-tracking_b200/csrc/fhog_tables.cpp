// fhog_tables.cpp -- see fhog_tables.h.  Host only; needs SSE (every x86-64 CPU).
#include "fhog_tables.h"

#include <emmintrin.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace mot {
namespace {

inline float hw_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
inline float hw_rcp(float x) { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// Smallest key width for which fn is constant on every run of 2^(23-K) consecutive mantissas of binade `expo`.
template <class F> int find_key_bits(F fn, int expo_lo, int expo_hi)
{
    for (int K = 6; K <= 14; ++K) {
        bool ok = true;
        const uint32_t run = 1u << (23 - K);
        for (int e = expo_lo; e <= expo_hi && ok; ++e) {
            const uint32_t base = (uint32_t)(e + 127) << 23;
            for (uint32_t key = 0; key < (1u << K) && ok; ++key) {
                const uint32_t first = f2u(fn(u2f(base | (key * run))));
                for (uint32_t m = key * run; m < (key + 1) * run; ++m)
                    if (f2u(fn(u2f(base | m))) != first) { ok = false; break; }
            }
        }
        if (ok) return K;
    }
    return -1;
}

void build(FhogTables &t)
{
    // ---- rsqrtps -------------------------------------------------------------------------
    t.rsqrt_bits = find_key_bits(hw_rsqrt, 0, 1);
    if (t.rsqrt_bits < 0) { t.error = "rsqrtps is not a <=14-bit mantissa table on this CPU; cannot reproduce the reference fHOG"; return; }
    {
        const int K = t.rsqrt_bits;
        t.rsqrt_tab.resize(2u << K);
        for (int p = 0; p < 2; ++p)
            for (uint32_t key = 0; key < (1u << K); ++key)
                t.rsqrt_tab[((size_t)p << K) + key] = hw_rsqrt(u2f(((uint32_t)(p + 127) << 23) | (key << (23 - K))));
        for (int e = -100; e <= 100; ++e)
            for (uint32_t key = 0; key < (1u << K); ++key)
                for (int rep = 0; rep < 2; ++rep) {
                    const uint32_t m = rep ? (((key + 1) << (23 - K)) - 1) : (key << (23 - K));
                    const float x = u2f(((uint32_t)(e + 127) << 23) | m);
                    if (f2u(hw_rsqrt(x)) != f2u(emu_rsqrt(t, x))) { t.error = "rsqrtps exponent scaling is not exact on this CPU"; return; }
                }
    }
    // ---- rcpps ---------------------------------------------------------------------------
    t.rcp_bits = find_key_bits(hw_rcp, 0, 0);
    if (t.rcp_bits < 0) { t.error = "rcpps is not a <=14-bit mantissa table on this CPU; cannot reproduce the reference fHOG"; return; }
    {
        const int K = t.rcp_bits;
        t.rcp_tab.resize(1u << K);
        for (uint32_t key = 0; key < (1u << K); ++key) t.rcp_tab[key] = hw_rcp(u2f((127u << 23) | (key << (23 - K))));
        for (int e = -60; e <= 60; ++e)
            for (uint32_t key = 0; key < (1u << K); ++key)
                for (int rep = 0; rep < 2; ++rep) {
                    const uint32_t m = rep ? (((key + 1) << (23 - K)) - 1) : (key << (23 - K));
                    const float x = u2f(((uint32_t)(e + 127) << 23) | m);
                    if (f2u(hw_rcp(x)) != f2u(emu_rcp(t, x))) { t.error = "rcpps exponent scaling is not exact on this CPU"; return; }
                }
        if (f2u(hw_rcp(1e10f)) != f2u(emu_rcp(t, 1e10f))) { t.error = "rcpps(1e10) mismatch"; return; }
    }
    // ---- acos table (gradientMex.cpp:47-56; C++ float overload of acos) --------------------
    {
        const int n = 10000, b = 10;
        const float PI = 3.14159265f;
        t.acos_tab.resize(2 * (n + b));
        float *a1 = t.acos_tab.data() + n + b;
        for (int i = -n - b; i < -n; ++i) a1[i] = PI;
        for (int i = -n; i < n; ++i) a1[i] = std::acos(i / float(n));
        for (int i = n; i < n + b; ++i) a1[i] = 0;
        for (int i = -n - b; i < n / 10; ++i) if (a1[i] > PI - 1e-6f) a1[i] = PI - 1e-6f;
    }
    // ---- orientation bin step table (gradMag :90-97 + gradQuantize :130-131 / :143-144) -----
    {
        const float PI = 3.14159265f;
        const float oMult = (float)18 / (2 * PI);
        const int NI = 20020;
        std::vector<int> pre(2 * NI);
        for (int s = 0; s < 2; ++s)
            for (int i = 0; i < NI; ++i) {
                float o = t.acos_tab[i];
                if (s) o += PI;
                const float oo = o * oMult;
                pre[s * NI + i] = (int)(oo + .5f);      // 0..18, 18 wraps to 0 later
            }
        bool done = false;
        for (int shift = 8; shift >= 4 && !done; --shift) {
            const int nseg = ((NI - 1) >> shift) + 1;
            std::vector<uint32_t> tab(2 * nseg);
            bool ok = true;
            for (int s = 0; s < 2 && ok; ++s)
                for (int g = 0; g < nseg && ok; ++g) {
                    const int lo = g << shift, hi = std::min(NI, (g + 1) << shift);
                    const int base = pre[s * NI + lo];
                    uint32_t thr = 0xFFFFFF;
                    for (int i = lo + 1; i < hi; ++i) {
                        const int d = base - pre[s * NI + i];
                        if (d == 0 && thr == 0xFFFFFF) continue;
                        if (d == 1 && thr == 0xFFFFFF) { thr = (uint32_t)i; continue; }
                        if (d == 1) continue;
                        ok = false; break;                // more than one step (or a non-monotone step) inside the segment
                    }
                    tab[s * nseg + g] = (thr << 8) | (uint32_t)base;
                }
            if (ok) { t.bin_shift = shift; t.bin_nseg = nseg; t.bin_tab.swap(tab); done = true; }
        }
        if (!done) { t.error = "orientation-bin table is not a step function of the acos index (libm acosf not monotone?)"; return; }
        for (int s = 0; s < 2; ++s)
            for (int i = 0; i < NI; ++i) {
                int want = pre[s * NI + i]; if (want >= 18) want = 0;
                if (emu_bin(t, i - 10010, s) != want) { t.error = "orientation-bin table self-check failed"; return; }
            }
    }
    t.rsrc_tab.resize(2 * t.rsqrt_tab.size());
    for (size_t i = 0; i < t.rsqrt_tab.size(); ++i) { t.rsrc_tab[2 * i] = t.rsqrt_tab[i]; t.rsrc_tab[2 * i + 1] = emu_rcp(t, t.rsqrt_tab[i]) * 0.0625f; }
    t.rcp_cap = emu_rcp(t, 1e10f);
    {
        // saturation MIN(rsqrt(M2), 1e10f) as a threshold on the bits of M2: bisect, then verify the equivalence on both ends
        // of every table run of the binades around 1e-20
        uint32_t lo = 0x007FFFFFu, hi = f2u(1.0f);            // lo saturates (denormal -> +inf), hi does not
        while (hi - lo > 1) { const uint32_t mid = lo + (hi - lo) / 2; if (!(emu_rsqrt(t, u2f(mid)) < 1e10f)) lo = mid; else hi = mid; }
        t.u_cap = lo;
        const int K = t.rsqrt_bits;
        for (int e = -75; e <= -55; ++e)
            for (uint32_t key = 0; key < (1u << K); ++key)
                for (int rep = 0; rep < 2; ++rep) {
                    const uint32_t u = ((uint32_t)(e + 127) << 23) | (rep ? (((key + 1) << (23 - K)) - 1) : (key << (23 - K)));
                    if ((!(emu_rsqrt(t, u2f(u)) < 1e10f)) != (u <= t.u_cap)) { t.error = "rsqrt saturation is not a single threshold"; return; }
                }
    }
    t.bin2_tab.resize(t.bin_tab.size());
    for (size_t i = 0; i < t.bin_tab.size(); ++i) {
        const uint32_t base = t.bin_tab[i] & 0xFFu, thr = t.bin_tab[i] >> 8;
        const uint32_t before = base >= 18 ? 0 : base, after = base == 0 ? 0 : ((base - 1) >= 18 ? 0 : base - 1);
        if (thr != 0xFFFFFFu && (thr > 20019u || base == 0)) { t.error = "orientation-bin table: unexpected step"; return; }
        t.bin2_tab[i] = ((thr == 0xFFFFFFu ? 0x3FFFFFu : thr) << 10) | (after << 5) | before;
    }
    // the fused kernel carries the orientation bin (0..17) in the five low mantissa bits of M/16; rcpps results are
    // short-mantissa table values, so those bits are zero -- verified here rather than assumed
    for (size_t i = 0; i < t.rsqrt_tab.size(); ++i)
        if (f2u(t.rsrc_tab[2 * i + 1]) & 31u) { t.error = "rcpps returns more than 18 mantissa bits on this CPU; unsupported"; return; }
    if (f2u(t.rcp_cap) & 31u) { t.error = "rcpps(1e10) has low mantissa bits set; unsupported"; return; }
    t.ok = true;
}

}  // namespace

float emu_rsqrt(const FhogTables &t, float x)
{
    const uint32_t u = f2u(x);
    if (u < 0x00800000u) return u2f(0x7F800000u);          // +0 / denormal -> +inf (DAZ behaviour of rsqrtps)
    const int e = (int)(u >> 23) - 127;
    const int p = e & 1, q = e >> 1;
    const uint32_t key = (u & 0x7FFFFFu) >> (23 - t.rsqrt_bits);
    const uint32_t T = f2u(t.rsqrt_tab[((size_t)p << t.rsqrt_bits) + key]);
    return u2f(T - ((uint32_t)q << 23));
}

float emu_rcp(const FhogTables &t, float x)
{
    const uint32_t u = f2u(x);
    const int e = (int)(u >> 23) - 127;
    const uint32_t key = (u & 0x7FFFFFu) >> (23 - t.rcp_bits);
    const uint32_t U = f2u(t.rcp_tab[key]);
    return u2f(U - ((uint32_t)e << 23));
}

int emu_bin(const FhogTables &t, int idx, int gy_negative)
{
    int i = idx + 10010;
    if (i < 0) i = 0;
    if (i > 20019) i = 20019;
    const uint32_t ent = t.bin_tab[(size_t)(gy_negative ? t.bin_nseg : 0) + (i >> t.bin_shift)];
    int b = (int)(ent & 0xFF) - ((uint32_t)i >= (ent >> 8) ? 1 : 0);
    return b >= 18 ? 0 : b;
}

const FhogTables &fhog_tables()
{
    static FhogTables t;
    static std::once_flag once;
    std::call_once(once, [] { build(t); });
    return t;
}

}  // namespace mot
