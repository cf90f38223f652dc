// kcf_any.cuh -- geometry of the fused any-size KCF kernel (kcf_any_kernel.cuh; instantiated in kcf_any_inst.cu, launched by kcf_any.cu),
// shared with the host layer that sizes its launches.
#pragma once
#include "mot_internal.h"

namespace mot {

// Shared-memory plan of one job with an hr x wc cell grid, in floats.  Depends on (hr, wc) only, so that the host can size a
// launch and the kernel can lay out every job of a mixed-size launch by itself.
//   A  (M/16 | bin) per pixel in the zero-bordered, y-de-interleaved layout of the gather  ->  spectra of one channel tile (two
//      ping-pong buffers of the Stockham passes, batch-major)  ->  zf / response
//   B  SSE tables + staged frame rows + gray strip (gradient phase)  ->  18-bin cell histograms R1
//   C  block normalisers, cell energies, channel-sum accumulators, twiddles, Hann vectors, reduction scratch
struct AnyGeo {
    int hr, wc, nb, sk, S, h0, w0;
    int rs, os;             // cell-histogram strides: column (hr) and orientation plane (multiple of 32: the gather's updates are conflict-free)
    int ps, pc, padm;       // (M | bin) layout: sub-column pitch, column pitch, words
    int jp;                 // column pairs of the two-for-one row pass
    int tc, xbuf;           // channels per spectral tile; float2 per ping-pong buffer
    int xw;                 // pixel columns per gradient strip
    int gs;                 // gray strip column stride (odd)
    int raw_pitch;          // bytes per staged frame row
    int lutp;               // floats reserved for the SSE tables
    int aF, bF;             // region sizes
    int oB, oN, oE, oACC, oTWR, oTWC, oWY, oWX, oRED, oLAB, total;   // offsets / total (floats)
    int ok;                 // fits the budget
    // Windows whose (M | bin) map and cell histograms do not fit one CTA's shared memory together run in STRIPS of `cs` cell columns:
    // frame rows, gray, (M | bin) and the histograms of one strip at a time in shared memory, the finished histograms parked in a
    // per-CTA scratch area in global memory (L2-resident), from which the spectral phase generates the channels.
    int strips, cs, oss;    // strip mode; cell columns per strip; orientation-plane stride of the strip's histograms (multiple of 32)
};

constexpr int ANY_SMEM_BUDGET_FLOATS = (227 * 1024 - 2048) / 4;      // dynamic shared memory one CTA may use (static part left out)

__host__ __device__ inline int any_max(int a, int b) { return a > b ? a : b; }
__host__ __device__ inline int any_min(int a, int b) { return a < b ? a : b; }

// Spectra of a tile of `tc` channels are held batch-major: element e of sequence b at [e * pitch + b], pitch odd, so that the 32
// lanes of a warp run the same butterfly on 32 different sequences (conflict-free, warp-uniform twiddles).  Row pass: hr elements x
// (tc * jp) sequences; column pass: wc elements x (tc * sk) sequences.
__host__ __device__ inline int any_xbuf(int hr, int wc, int jp, int sk, int tc) { return any_max(hr * ((tc * jp) | 1), wc * ((tc * sk) | 1)); }

__host__ __device__ inline AnyGeo any_geo(int hr, int wc, int lut_floats)
{
    AnyGeo g;
    g.hr = hr; g.wc = wc; g.nb = hr * wc; g.sk = hr / 2 + 1; g.S = wc * g.sk; g.h0 = 4 * hr; g.w0 = 4 * wc;
    g.rs = hr; g.os = (g.nb + 31) & ~31;
    g.jp = (wc + 1) / 2;
    g.gs = (g.h0 + 2) | 1;
    const int rmax = 4 * hr + 3, cmax = 4 * wc + 3;
    g.raw_pitch = ((3 * cmax + 30) + 15) & ~15;
    const int raw_floats = rmax * (g.raw_pitch / 4);
    g.lutp = (lut_floats + 3) & ~3;
    const int r1f = 18 * g.os;
    // region C, offsets relative to its start
    int o = 0;
    const int cN = o;   o += ((wc + 1) * (hr + 1) + 3) & ~3;
    const int cE = o;   o += (g.nb + 3) & ~3;
    const int cACC = o; o += (2 * g.S + 3) & ~3;
    const int cTWR = o; o += (2 * hr + 3) & ~3;
    const int cTWC = o; o += (2 * wc + 3) & ~3;
    const int cWY = o;  o += (hr + 3) & ~3;
    const int cWX = o;  o += (wc + 3) & ~3;
    const int cRED = o; o += 64;
    const int cLAB = o; o += 4 * (g.sk + wc);               // 1-D label spectra of a per-track label sigma (double2), extensions only
    const int cF = o;
    g.ok = 0; g.xw = 0; g.bF = 0; g.aF = 0; g.ps = 0; g.pc = 0; g.padm = 0;
    // Two choices, most comfortable first: the sub-column pitch of the (M | bin) layout = 8 (mod 16), which makes the gradient phase's
    // stores conflict-free, or the bare minimum hr + 1; the gradient strip as wide as fits (whole window when narrow, else 62 / 30 / 14 / 6 pixel columns).
    for (int pad = 1; pad >= 0 && !g.ok; --pad) {
        const int ps = pad ? (((hr + 1 - 8 + 15) / 16) * 16 + 8) : hr + 1;
        const int pc = 4 * ps + 2, padm = (g.w0 + 4) * pc;
        const int aF = (any_max(padm, 4 * any_xbuf(hr, wc, g.jp, g.sk, 1)) + 3) & ~3;
        const int cand[4] = { g.w0 <= 64 ? g.w0 : 62, 30, 14, 6 };      // + 2 apron columns = whole warps of the gray conversion
        for (int q = 0; q < 4 && !g.ok; ++q) {
            const int xw = any_min(cand[q], g.w0);
            const int bF = (any_max(r1f, g.lutp + raw_floats + (xw + 2) * g.gs) + 3) & ~3;
            g.ps = ps; g.pc = pc; g.padm = padm; g.aF = aF; g.xw = xw; g.bF = bF;      // the last one tried stays when nothing fits
            g.ok = aF + bF + cF <= ANY_SMEM_BUDGET_FLOATS;
        }
    }
    g.strips = 0; g.cs = wc; g.oss = g.os;
    if (!g.ok) {
        // strip mode: the widest strip that still leaves room for tiles of >= 4 channels, else >= 2, else 1
        const int ps = hr + 1, pc = 4 * ps + 2;
        for (int want_tc = 4; want_tc >= 1 && !g.ok; want_tc = want_tc > 2 ? 2 : want_tc - 1) {
            for (int cs = 16; cs >= 1 && !g.ok; cs >>= 1) {
                if (cs > wc) continue;
                const int pxw = 4 * cs + 6;                                   // gray columns of a strip: its pixels, +-2 for the gather, +-1 for the gradient
                const int pitch = ((3 * pxw + 30) + 15) & ~15;
                const int raw_s = rmax * (pitch / 4), oss = (cs * hr + 31) & ~31;
                const int bF = (g.lutp + any_max(raw_s + pxw * g.gs, 18 * oss) + 3) & ~3;
                const int a_min = any_max((4 * cs + 4) * pc, 4 * any_xbuf(hr, wc, g.jp, g.sk, want_tc));
                const int left = ANY_SMEM_BUDGET_FLOATS - bF - cF;
                if (left < a_min) continue;
                g.strips = 1; g.cs = cs; g.oss = oss; g.ps = ps; g.pc = pc; g.padm = (4 * cs + 4) * pc; g.raw_pitch = pitch; g.xw = 4 * cs + 4;
                g.bF = bF; g.aF = any_min(left & ~3, (any_max(g.padm, 4 * any_xbuf(hr, wc, g.jp, g.sk, KCF_CHAN)) + 3) & ~3);
                g.ok = 1;
            }
        }
    }
    // channels per tile: as many as two ping-pong buffers in region A hold, then balanced over the resulting number of tiles
    int tcmax = 1;
    for (int tc = 2; tc <= KCF_CHAN; ++tc) if (4 * any_xbuf(hr, wc, g.jp, g.sk, tc) <= g.aF) tcmax = tc;
    const int ntiles = (KCF_CHAN + tcmax - 1) / tcmax;
    g.tc = (KCF_CHAN + ntiles - 1) / ntiles;
    g.xbuf = any_xbuf(hr, wc, g.jp, g.sk, g.tc);
    g.oB = g.aF;
    const int c0 = g.aF + g.bF;
    g.oN = c0 + cN; g.oE = c0 + cE; g.oACC = c0 + cACC; g.oTWR = c0 + cTWR; g.oTWC = c0 + cTWC; g.oWY = c0 + cWY; g.oWX = c0 + cWX; g.oRED = c0 + cRED; g.oLAB = c0 + cLAB;
    g.total = c0 + cF;
    if (hr < 2 || wc < 2) g.ok = 0;
    return g;
}

// Per-N tables (N = side of a cell grid, 2..nmax), built once per context on the host in the reference's precision:
//   hann[off(N) + i]  float   0.5 - 0.5 cos(2 pi i / (N - 1))           include/sigpack/window/window.h:34-48, 83-89
//   tw  [off(N) + t]  float2  exp(-2 pi i t / N)
//   lab [off(N) + k]  double2 DFT_N of the circularly shifted 1-D Gaussian label   trackers/kcf.cpp:96-122, 78-94
//   plan[N]           radices of the mixed-radix transform of length N
// off(N) = N (N - 1) / 2 - 1.  Every (hr, wc) combination is served from them, so a tracker of a never-seen size can be born on
// the device (the device-resident frame loop) without a host round trip.
struct AnyPlan { unsigned short nf, r[7]; };
struct AnyTablesDev {
    const float *hann; const float2 *tw; const double2 *lab; const AnyPlan *plan; int nmax;
};
__host__ __device__ inline int any_off(int N) { return N * (N - 1) / 2 - 1; }

// 0 when the fused any-size kernel cannot hold an hr x wc window even in strips (the unfused path kcf_generic.cu serves it then)
size_t kcf_any_smem_bytes(int hr, int wc, int lut_floats);
// bytes of global scratch one CTA needs for this window (0: everything stays in shared memory)
size_t kcf_any_scratch_bytes(int hr, int wc, int lut_floats);
// One launch over jobs of ANY mix of sizes whose plans fit `smem_bytes`; threads = CTA size to use.  A job whose plan does not fit
// (a host-side sizing bug) is skipped and *err_flag (device-visible, may be null) set, never run out of bounds.
// scratch / scratch_stride: per-CTA global scratch for strip-mode windows (null: the launch holds none; such jobs are skipped and flagged).
int kcf_launch_any(int mode, const KcfLaunch &p, const AnyTablesDev &at, size_t smem_bytes, int threads, int ctas_per_sm, int *err_flag,
                   float *scratch, long scratch_stride_floats, int max_ctas, cudaStream_t s);

}  // namespace mot
