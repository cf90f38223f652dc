// fhog_tables.h -- host-built lookup tables that let the GPU reproduce the reference fHOG bit for bit.
//
// libhog/gradientMex.cpp:83-84 computes the gradient magnitude with the SSE approximations
// _mm_rsqrt_ps / _mm_rcp_ps (libhog/sse.hpp:40-41).  Their ~3e-4 relative error is larger than the
// 1e-4 parity tolerance and feeds a truncating table index (:90) and a nearest-bin quantiser (:130-131),
// so exact math on the GPU FAILS parity.  Both instructions are pure
// table look-ups on the top mantissa bits with exact power-of-two exponent scaling; the tables are
// harvested from the host CPU the process runs on (Intel and AMD differ), verified exhaustively, and
// uploaded.  The orientation path (acosTable :47-56 + gradQuantize :130-131, built with the host's
// libm) is folded into one exact "orientation bin" step table.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mot {

struct FhogTables {
    bool ok = false;
    std::string error;
    // rsqrt: value for x in [1,4): index = parity * (1 << rsqrt_bits) + (mantissa >> (23 - rsqrt_bits))
    int rsqrt_bits = 0;
    std::vector<float> rsqrt_tab;
    // rcp: value for x in [1,2): index = mantissa >> (23 - rcp_bits)
    int rcp_bits = 0;
    std::vector<float> rcp_tab;
    // fused form for the kernels: rsrc_tab[2i] = rsqrt_tab[i], rsrc_tab[2i+1] = rcp(rsqrt_tab[i]) / 16 (both lookups share one
    // index; the exact power-of-two factor is the M * 0.0625 of the gather, gradientMex.cpp:192, folded in);
    // rcp_cap = rcp(1e10f), the value taken when MIN(RCPSQRT(M2), 1e10f) saturates (zero gradient);
    // u_cap = the largest float bit pattern of M2 for which it saturates (rsqrt is non-increasing, so "saturates" is one compare)
    std::vector<float> rsrc_tab;
    float rcp_cap = 0.f;
    uint32_t u_cap = 0;
    // orientation bin: segment = (idx + 10010) >> bin_shift, entry[sign * nseg + segment] = (thr << 8) | base,
    // bin = base - (idx + 10010 >= thr), then 18 wraps to 0.  idx = (int)(Gx * m * 10000) in [-10010, 10010).
    int bin_shift = 0, bin_nseg = 0;
    std::vector<uint32_t> bin_tab;
    // the same step table for the fused kernel with the wrap 18 -> 0 folded in: (thr << 10) | (bin at/after thr) << 5 | (bin before thr)
    std::vector<uint32_t> bin2_tab;
    // the raw acos table (20020 floats, index idx + 10010) kept for tests / debugging dumps
    std::vector<float> acos_tab;
};

// Harvest + verify (about 50 ms, once per process).  Never throws; check .ok / .error.
const FhogTables &fhog_tables();

// Host emulation of the two instructions from the tables (used by the self-check and the CPU-side tests).
float emu_rsqrt(const FhogTables &t, float x);
float emu_rcp(const FhogTables &t, float x);
int emu_bin(const FhogTables &t, int idx, int gy_negative);

}  // namespace mot
