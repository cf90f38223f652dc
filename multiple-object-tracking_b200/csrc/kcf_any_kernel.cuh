// kcf_any_kernel.cuh -- the fused KCF/DCF kernel for ANY window size: everything between the BGR frame bytes and the updated model in
// one launch, like kcf_fused.cuh, but with the cell grid (hr x wc) a run-time property of each job.
//
// The reference plans FFTW for whatever size the detection box has (trackers/kcf.cpp:146-195) and runs, per track and frame,
//   rgb2Gray + bilinearInterpolationGray      top/drawlib.c:192-240, 542-637   (called top/td.cpp:348-364)
//   FHoG::extract -> gradMag -> fhog          libhog/fhog.h:16-38, libhog/gradientMex.cpp:59-100, 148-317
//   kcf_get_features / kcf_fft2_features      trackers/kcf.cpp:245-267 (31 r2c plans of n0 = f_cols, n1 = f_rows)
//   predict: kcf_linear_correlation_zf + kcf_predict_ifft2 (+ the clamp of top/td.cpp:378-381)   kcf.cpp:306-362, 397-439
//   update : kcf_linear_correlation_kf + kcf_update_alpha + kcf_update_xf                         kcf.cpp:269-304, 364-395, 441-476
// Here one persistent CTA runs one job at a time with all intermediates in shared memory (plan: kcf_any.cuh):
//   P0  frame rows of the crop through the bulk-copy engine; SSE tables; per-N twiddles and Hann vectors
//   P1  per strip of pixel columns: bytes -> gray (or the reference's scrambled resample) -> gradient magnitude and orientation bin,
//       table-emulated SSE arithmetic, bit-exact (fhog_common.cuh)
//   P2  18-bin cell histograms by ordered gather (the reference's summation order), boundary scaling, cell energies
//   P3  2x2 block normalisers
//   P4..P6 per tile of channels: features x Hann window -> row transform (two real columns per complex transform) -> split and
//       transpose -> column transform -> x conj(model) (predict) or |.|^2 and model lerp (update) -> channel sum, in channel order
//   P7  predict: x alpha x norm -> inverse column transform -> Hermitian rebuild -> inverse row transform -> first-max argmax ->
//       box shift (float, truncated) -> optional clamp;   update: kf, alpha lerp, tracker_update bookkeeping
// Transforms are Stockham mixed-radix passes in shared memory (radices 4, 2, 3, 5, 7 in registers; any other prime factor by its
// definition), driven by a per-length plan, so every length works.  Spectra are held batch-major (element e of sequence b at
// [e * pitch + b]): the 32 lanes of a warp run the SAME butterfly on 32 different sequences, so index arithmetic and twiddles are
// warp-uniform and every shared-memory access is conflict-free.  The model is read once and written once per update, coalesced.
#pragma once
#include "kcf_any.cuh"
#include "fhog_common.cuh"
#include "copy_async.cuh"
#include <cstdlib>

namespace mot {

namespace {

constexpr int ANY_MAX_STAGES = 7;

__device__ __forceinline__ uint32_t magic_of(uint32_t d) { return d <= 1 ? 0u : (0xFFFFFFFFu / d + 1u); }
// n / d for n * d < 2^32 (m = magic_of(d))
__device__ __forceinline__ int fdiv(int n, uint32_t m) { return m ? (int)__umulhi((uint32_t)n, m) : n; }

// Everything about the current window size that costs a division: computed by one warp when the size changes between jobs
struct AnyJobConst {
    int hr, wc;
    AnyGeo g;
    int nst_r, nst_c;                                   // stages of the length-hr / length-wc transforms
    int rad_r[ANY_MAX_STAGES], ns_r[ANY_MAX_STAGES]; uint32_t mg_ns_r[ANY_MAX_STAGES];
    int rad_c[ANY_MAX_STAGES], ns_c[ANY_MAX_STAGES]; uint32_t mg_ns_c[ANY_MAX_STAGES];
    uint32_t mg_small[64];                              // magic_of(d), d < 64 (chunk counts of the passes)
    uint32_t mg_hr, mg_hr1, mg_sk, mg_jp, mg_nych;
    float norm;                                         // feature_norm_ratio, kcf.cpp:197
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <int DIR> __device__ __forceinline__ float2 rot90(float2 a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }

// R-point DFT in registers, X[q] = sum_r v[r] exp(DIR 2 pi i q r / R)
template <int R, int DIR> struct Bfly;
template <int DIR> struct Bfly<2, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[2]) { const float2 a = v[0], b = v[1]; v[0] = cadd(a, b); v[1] = csub(a, b); }
};
template <int DIR> struct Bfly<3, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[3])
    {
        const float2 t = cadd(v[1], v[2]);
        const float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
        const float2 d = csub(v[1], v[2]);
        const float2 r = rot90<DIR>(make_float2(0.8660254038f * d.x, 0.8660254038f * d.y));
        v[0] = cadd(v[0], t); v[1] = cadd(m, r); v[2] = csub(m, r);
    }
};
template <int DIR> struct Bfly<4, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[4])
    {
        const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = rot90<DIR>(csub(v[1], v[3]));
        v[0] = cadd(a, c); v[2] = csub(a, c); v[1] = cadd(b, d); v[3] = csub(b, d);
    }
};
template <int DIR> struct Bfly<5, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[5])
    {
        const float c1 = 0.3090169944f, c2 = -0.8090169944f, s1 = 0.9510565163f, s2 = 0.5877852523f;
        const float2 t1 = cadd(v[1], v[4]), t2 = cadd(v[2], v[3]), t3 = csub(v[1], v[4]), t4 = csub(v[2], v[3]);
        const float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x, v[0].y + c1 * t1.y + c2 * t2.y);
        const float2 a2 = make_float2(v[0].x + c2 * t1.x + c1 * t2.x, v[0].y + c2 * t1.y + c1 * t2.y);
        const float2 b1 = rot90<DIR>(make_float2(s1 * t3.x + s2 * t4.x, s1 * t3.y + s2 * t4.y));
        const float2 b2 = rot90<DIR>(make_float2(s2 * t3.x - s1 * t4.x, s2 * t3.y - s1 * t4.y));
        v[0] = make_float2(v[0].x + t1.x + t2.x, v[0].y + t1.y + t2.y);
        v[1] = cadd(a1, b1); v[4] = csub(a1, b1); v[2] = cadd(a2, b2); v[3] = csub(a2, b2);
    }
};
template <int DIR> struct Bfly<7, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[7])
    {
        const float c1 = 0.6234898019f, c2 = -0.2225209340f, c3 = -0.9009688679f, s1 = 0.7818314825f, s2 = 0.9749279122f, s3 = 0.4338837391f;
        const float2 t1 = cadd(v[1], v[6]), t2 = cadd(v[2], v[5]), t3 = cadd(v[3], v[4]);
        const float2 u1 = csub(v[1], v[6]), u2 = csub(v[2], v[5]), u3 = csub(v[3], v[4]);
        const float2 a1 = make_float2(v[0].x + c1 * t1.x + c2 * t2.x + c3 * t3.x, v[0].y + c1 * t1.y + c2 * t2.y + c3 * t3.y);
        const float2 a2 = make_float2(v[0].x + c2 * t1.x + c3 * t2.x + c1 * t3.x, v[0].y + c2 * t1.y + c3 * t2.y + c1 * t3.y);
        const float2 a3 = make_float2(v[0].x + c3 * t1.x + c1 * t2.x + c2 * t3.x, v[0].y + c3 * t1.y + c1 * t2.y + c2 * t3.y);
        const float2 b1 = rot90<DIR>(make_float2(s1 * u1.x + s2 * u2.x + s3 * u3.x, s1 * u1.y + s2 * u2.y + s3 * u3.y));
        const float2 b2 = rot90<DIR>(make_float2(s2 * u1.x - s3 * u2.x - s1 * u3.x, s2 * u1.y - s3 * u2.y - s1 * u3.y));
        const float2 b3 = rot90<DIR>(make_float2(s3 * u1.x - s1 * u2.x + s2 * u3.x, s3 * u1.y - s1 * u2.y + s2 * u3.y));
        v[0] = make_float2(v[0].x + t1.x + t2.x + t3.x, v[0].y + t1.y + t2.y + t3.y);
        v[1] = cadd(a1, b1); v[6] = csub(a1, b1); v[2] = cadd(a2, b2); v[5] = csub(a2, b2); v[3] = cadd(a3, b3); v[4] = csub(a3, b3);
    }
};

// composite radices (decimation in time over the smaller factor): fewer passes through shared memory
template <int DIR> struct Bfly<6, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[6])
    {
        float2 e[3] = { v[0], v[2], v[4] }, o[3] = { v[1], v[3], v[5] };
        Bfly<3, DIR>::run(e); Bfly<3, DIR>::run(o);
        const float sg = DIR < 0 ? -1.f : 1.f;
        const float2 t1 = cmul(o[1], make_float2(0.5f, sg * 0.8660254038f)), t2 = cmul(o[2], make_float2(-0.5f, sg * 0.8660254038f));
        v[0] = cadd(e[0], o[0]); v[3] = csub(e[0], o[0]);
        v[1] = cadd(e[1], t1);   v[4] = csub(e[1], t1);
        v[2] = cadd(e[2], t2);   v[5] = csub(e[2], t2);
    }
};
template <int DIR> struct Bfly<8, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[8])
    {
        float2 e[4] = { v[0], v[2], v[4], v[6] }, o[4] = { v[1], v[3], v[5], v[7] };
        Bfly<4, DIR>::run(e); Bfly<4, DIR>::run(o);
        const float h = 0.7071067812f;
        // w8^1 = h (1 -+ i), w8^2 = -+ i, w8^3 = h (-1 -+ i)   (upper sign: forward)
        const float2 r1 = rot90<DIR>(o[1]), r3 = rot90<DIR>(o[3]);
        const float2 t1 = make_float2(h * (o[1].x + r1.x), h * (o[1].y + r1.y));
        const float2 t2 = rot90<DIR>(o[2]);
        const float2 t3 = make_float2(h * (r3.x - o[3].x), h * (r3.y - o[3].y));
        v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
        v[1] = cadd(e[1], t1);   v[5] = csub(e[1], t1);
        v[2] = cadd(e[2], t2);   v[6] = csub(e[2], t2);
        v[3] = cadd(e[3], t3);   v[7] = csub(e[3], t3);
    }
};
template <int DIR> struct Bfly<9, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[9])
    {
        float2 a0[3] = { v[0], v[3], v[6] }, a1[3] = { v[1], v[4], v[7] }, a2[3] = { v[2], v[5], v[8] };
        Bfly<3, DIR>::run(a0); Bfly<3, DIR>::run(a1); Bfly<3, DIR>::run(a2);
        const float sg = DIR < 0 ? -1.f : 1.f;
        const float2 w1 = make_float2(0.7660444431f, sg * 0.6427876097f), w2 = make_float2(0.1736481777f, sg * 0.9848077530f),
                     w4 = make_float2(-0.9396926208f, sg * 0.3420201433f);
        // X[k1 + 3 k2] = DFT3 over j of w9^(j k1) A_j[k1]
        float2 b0[3] = { a0[0], a1[0], a2[0] };
        float2 b1[3] = { a0[1], cmul(a1[1], w1), cmul(a2[1], w2) };
        float2 b2[3] = { a0[2], cmul(a1[2], w2), cmul(a2[2], w4) };
        Bfly<3, DIR>::run(b0); Bfly<3, DIR>::run(b1); Bfly<3, DIR>::run(b2);
        v[0] = b0[0]; v[3] = b0[1]; v[6] = b0[2];
        v[1] = b1[0]; v[4] = b1[1]; v[7] = b1[2];
        v[2] = b2[0]; v[5] = b2[1]; v[8] = b2[2];
    }
};
template <int DIR> struct Bfly<10, DIR> {
    static __device__ __forceinline__ void run(float2 (&v)[10])
    {
        float2 e[5] = { v[0], v[2], v[4], v[6], v[8] }, o[5] = { v[1], v[3], v[5], v[7], v[9] };
        Bfly<5, DIR>::run(e); Bfly<5, DIR>::run(o);
        const float sg = DIR < 0 ? -1.f : 1.f;
        const float2 t1 = cmul(o[1], make_float2(0.8090169944f, sg * 0.5877852523f)), t2 = cmul(o[2], make_float2(0.3090169944f, sg * 0.9510565163f));
        const float2 t3 = cmul(o[3], make_float2(-0.3090169944f, sg * 0.9510565163f)), t4 = cmul(o[4], make_float2(-0.8090169944f, sg * 0.5877852523f));
        v[0] = cadd(e[0], o[0]); v[5] = csub(e[0], o[0]);
        v[1] = cadd(e[1], t1);   v[6] = csub(e[1], t1);
        v[2] = cadd(e[2], t2);   v[7] = csub(e[2], t2);
        v[3] = cadd(e[3], t3);   v[8] = csub(e[3], t3);
        v[4] = cadd(e[4], t4);   v[9] = csub(e[4], t4);
    }
};

// One Stockham pass of radix R over `nbatch` sequences of length n held batch-major (element e of sequence b at [e * pitch + b]).
// A warp task = (butterfly j, chunk of 32 sequences): j, its twiddles and every address term are warp-uniform, the lanes differ
// only in b.  Ns = product of the radices of the passes before this one; tw[t] = exp(-2 pi i t / n).
template <int R, int DIR>
__device__ __forceinline__ void fft_pass(const float2 *__restrict__ in, float2 *__restrict__ out, int n, int Ns, uint32_t mg_ns, const uint32_t *mg_small,
                                         int nbatch, int pitch, const float2 *__restrict__ tw, int warp, int lane, int nwarps)
{
    const int m = n / R, step = fdiv(m, mg_ns);
    const int nchunk = (nbatch + 31) >> 5, ntask = m * nchunk;
    const uint32_t mg_chunk = mg_small[nchunk & 63];
    const int mp = m * pitch, np = Ns * pitch;
    for (int task = warp; task < ntask; task += nwarps) {
        const int j = fdiv(task, mg_chunk), b = ((task - j * nchunk) << 5) + lane;
        const int k = j - fdiv(j, mg_ns) * Ns;
        if (b >= nbatch) continue;
        const float2 *src = in + j * pitch + b;
        float2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[r * mp];
        if (Ns > 1) {
            const int ks = k * step;                      // k r step < n for every r < R
#pragma unroll
            for (int r = 1; r < R; ++r) {
                float2 w = tw[ks * r];                    // same word for every lane: broadcast
                if (DIR > 0) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        Bfly<R, DIR>::run(v);
        float2 *dst = out + ((j - k) * R + k) * pitch + b;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * np] = v[r];
    }
}

// The same pass for any other (prime) radix, straight from the definition: a warp task = (output o, chunk of 32 sequences)
template <int DIR>
__device__ __forceinline__ void fft_pass_prime(const float2 *__restrict__ in, float2 *__restrict__ out, int n, int R, int Ns, uint32_t mg_ns, const uint32_t *mg_small,
                                               int nbatch, int pitch, const float2 *__restrict__ tw, int warp, int lane, int nwarps)
{
    const uint32_t mg_R = magic_of(R);
    const int m = fdiv(n, mg_R), step = fdiv(m, mg_ns);
    const int nchunk = (nbatch + 31) >> 5, ntask = n * nchunk;
    const uint32_t mg_chunk = mg_small[nchunk & 63];
    const int mp = m * pitch;
    for (int task = warp; task < ntask; task += nwarps) {
        const int o = fdiv(task, mg_chunk), b = ((task - o * nchunk) << 5) + lane;
        const int oq = fdiv(o, mg_ns), k = o - oq * Ns;            // o = (jhi * R + q) * Ns + k
        const int jhi = fdiv(oq, mg_R), q = oq - jhi * R;
        const int j = jhi * Ns + k;
        int inc = k * step + q * m;                                  // phase advance per r, < 2n
        if (inc >= n) inc -= n;
        if (b >= nbatch) continue;
        const float2 *src = in + j * pitch + b;
        float2 acc = src[0];
        int ph = 0;
        for (int r = 1; r < R; ++r) {
            ph += inc; if (ph >= n) ph -= n;
            float2 w = tw[ph];
            if (DIR > 0) w.y = -w.y;
            const float2 x = src[r * mp];
            acc.x = fmaf(x.x, w.x, fmaf(-x.y, w.y, acc.x));
            acc.y = fmaf(x.x, w.y, fmaf(x.y, w.x, acc.y));
        }
        out[o * pitch + b] = acc;
    }
}

// All passes of a batch of transforms; ping-pongs between buf0 (input) and buf1; returns the buffer holding the result.
// Ends with a barrier after every pass (the next pass, or the caller, reads what other threads wrote).
template <int DIR>
__device__ __forceinline__ float2 *fft_batch(float2 *buf0, float2 *buf1, int n, int nst, const int *rad, const int *ns, const uint32_t *mg_ns, const uint32_t *mg_small,
                                             int nbatch, int pitch, const float2 *tw, int warp, int lane, int nwarps)
{
    float2 *in = buf0, *out = buf1;
    for (int s = 0; s < nst; ++s) {
        const int R = rad[s], Ns = ns[s];
        switch (R) {
        case 4: fft_pass<4, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 2: fft_pass<2, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 3: fft_pass<3, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 5: fft_pass<5, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 7: fft_pass<7, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 8: fft_pass<8, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 6: fft_pass<6, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 10: fft_pass<10, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        case 9: fft_pass<9, DIR>(in, out, n, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        default: fft_pass_prime<DIR>(in, out, n, R, Ns, mg_ns[s], mg_small, nbatch, pitch, tw, warp, lane, nwarps); break;
        }
        __syncthreads();
        float2 *q = in; in = out; out = q;
    }
    return in;
}

// ordered gather of the 18-bin histograms, CPT cells per thread interleaved (libhog/gradientMex.cpp:183-230, 308-309)
template <int CPT>
// Cells are numbered column-major inside the strip (ncell of them, first cell column cx0 of the window); MB is the strip's map
// (local pixel column 0 = window pixel 4 cx0 - 2), R1 its histograms with plane stride OS.
__device__ __forceinline__ void gather_cells(const AnyGeo &g, uint32_t mg_hr, const uint32_t *__restrict__ MB, float *__restrict__ R1, int OS, float *__restrict__ Es,
                                             int cell0, int ncell, int cx0, int NT)
{
    const int PC = g.pc, PS = g.ps;
    int ccx[CPT], ccy[CPT]; bool live[CPT]; float *h[CPT]; const uint32_t *mb0[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
        const int cell = cell0 + u * NT;
        live[u] = cell < ncell;
        const int cc = live[u] ? cell : 0;
        ccx[u] = fdiv(cc, mg_hr); ccy[u] = cc - ccx[u] * g.hr;
        h[u] = R1 + cc;
        mb0[u] = MB + (4 * ccx[u]) * PC + ccy[u];
        if (live[u])
            for (int o = 0; o < 18; ++o) h[u][o * OS] = 0.f;
    }
    if (!live[0]) return;
#pragma unroll
    for (int dx = 0; dx < 8; ++dx) {
        const float wxv = 0.125f + 0.25f * (float)(dx < 4 ? dx : 7 - dx);
#pragma unroll
        for (int dy = 0; dy < 8; ++dy) {
            const float w = wxv * (0.125f + 0.25f * (float)(dy < 4 ? dy : 7 - dy));     // dyadic weights: exact product
#pragma unroll
            for (int u = 0; u < CPT; ++u) {
                if (!live[u]) continue;
                const uint32_t mb = mb0[u][dx * PC + (dy & 3) * PS + (dy >> 2)];
                const float v = __fmul_rn(w, __uint_as_float(mb & ~31u));
                float *const hb = h[u] + (int)(mb & 31u) * OS;
                *hb = __fadd_rn(*hb, v);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < CPT; ++u) {
        if (!live[u]) continue;
        const int cx = ccx[u] + cx0, cy = ccy[u];
        // boundary cells x 8/7 per touching side (gradientMex.cpp:226-229): x first, then y; multiplying by 1.0f is the identity
        const float sx0 = (cx == 0) ? 8.f / 7.f : 1.f, sy0 = (cy == 0) ? 8.f / 7.f : 1.f;
        const float sx1 = (cx == g.wc - 1) ? 8.f / 7.f : 1.f, sy1 = (cy == g.hr - 1) ? 8.f / 7.f : 1.f;
        float e = 0.f, r[18];
#pragma unroll
        for (int o = 0; o < 18; ++o) {
            float v = h[u][o * OS];
            v = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(v, sx0), sy0), sx1), sy1);
            h[u][o * OS] = v; r[o] = v;
        }
#pragma unroll
        for (int o = 0; o < 9; ++o) { const float r2 = __fadd_rn(r[o], r[o + 9]); e = __fadd_rn(e, __fmul_rn(r2, r2)); }
        Es[cx * g.hr + cy] = e;
    }
}

}  // namespace

// EXT: the north-star extensions are compiled in (Gaussian kernel, sub-pixel peak, padding, per-track label sigma); without it the
// kernel is exactly the reference's filter and carries none of their registers
template <int MODE, bool DUMP, int NTMAX, bool STRIPS, bool EXT>
__global__ void __launch_bounds__(NTMAX, 1) kcf_any_kernel(const KcfLaunch p, const AnyTablesDev at, const int lut_floats, const int smem_floats, int *err_flag,
                                                           float *r1g_base, long r1g_stride)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ __align__(8) uint64_t mbar_lut, mbar_raw;
    __shared__ AnyJobConst jc;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    if (tid == 0) { mbar_init(&mbar_lut, 1); mbar_init(&mbar_raw, 1); jc.hr = -1; jc.wc = -1; }
    if (tid < 64) jc.mg_small[tid] = magic_of(tid);
    __syncthreads();
    const bool lut_smem = lut_floats > 0;
    const int n_rs = 2 * (2 << p.tab.rsqrt_bits), n_bn = (2 * p.tab.bin_nseg + 3) & ~3;
    const int Wm = p.frame_w - 1, Hm = p.frame_h - 1;
    const int n_jobs = p.n_jobs_dev ? min(*p.n_jobs_dev, p.n_jobs) : p.n_jobs;
    uint32_t phase = 0, phase_raw = 0;

    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        __syncthreads();                                       // the previous job is done with shared memory (and with jc)
        // ---------------------------------------------------------------- job descriptor (every thread reads the same words: broadcast)
        const int slot = p.slots[job];
        KcfMeta *const meta = p.meta + slot;
        const int bi = p.box_index ? p.box_index[job] : job;
        // extensions: the caller's box is the TARGET, the tracked window is the box scaled by the padding about its centre
        const mot_bbox_t tbox = p.boxes[bi];
        const bool padded = EXT && p.ext.padding > 1.0f;
        const mot_bbox_t box = padded ? kcf_pad_box(tbox, p.ext.padding) : tbox;
        const bool gauss = EXT && p.ext.gaussian != 0;
        const int rows = meta->rows, cols = meta->cols, hr = meta->hr, wc = meta->wc;
        const bool first_update = meta->first_update != 0;
        const uint8_t *const frame = (p.gray == nullptr) ? p.frame_ptr[p.frames[job]] : nullptr;
        float2 *const model = meta->model_ptr ? meta->model_ptr : p.model + (long)slot * p.model_stride;
        float *const alpha = meta->alpha_ptr ? meta->alpha_ptr : p.alpha + (long)slot * p.alpha_stride;

        if (jc.hr != hr || jc.wc != wc) {                      // block-uniform: the window size changed (first job, mixed launches)
            __syncthreads();
            if (warp == 0) {
                if (lane == 0) {
                    jc.g = any_geo(hr, wc, lut_floats);
                    jc.norm = (float)(1.0 / (double)((float)(wc * hr * KCF_CHAN)));
                    jc.mg_hr = magic_of(hr); jc.mg_hr1 = magic_of(hr - 1); jc.mg_sk = magic_of(hr / 2 + 1);
                    jc.mg_jp = magic_of((wc + 1) / 2); jc.mg_nych = magic_of((4 * hr + 31) >> 5);
                }
                if (lane == 1 || lane == 2) {
                    const int n = lane == 1 ? hr : wc;
                    const AnyPlan pl = at.plan[n];
                    int *rad = lane == 1 ? jc.rad_r : jc.rad_c, *ns = lane == 1 ? jc.ns_r : jc.ns_c;
                    uint32_t *mgn = lane == 1 ? jc.mg_ns_r : jc.mg_ns_c;
                    int acc = 1;
                    for (int s = 0; s < (int)pl.nf; ++s) { rad[s] = pl.r[s]; ns[s] = acc; mgn[s] = magic_of(acc); acc *= pl.r[s]; }
                    if (lane == 1) jc.nst_r = pl.nf; else jc.nst_c = pl.nf;
                }
                __syncwarp();
                if (lane == 0) { jc.hr = hr; jc.wc = wc; }
            }
            __syncthreads();
        }
        const AnyGeo &g = jc.g;
        if (g.total > smem_floats || !g.ok || (g.strips != 0) != STRIPS || (STRIPS && 18L * g.os > r1g_stride)) {     // cannot happen when the host sized the launch; never run out of bounds
            if (tid == 0 && err_flag) *reinterpret_cast<volatile int *>(err_flag) = 1;
            continue;
        }
        const int H0 = g.h0, W0 = g.w0, SK = g.sk, S = g.S, JP = g.jp, NB = g.nb;
        uint32_t *const MB = reinterpret_cast<uint32_t *>(smem);
        float *const Bf = smem + g.oB;
        // histograms: all in shared memory (on top of the tables and the staging area, dead by then), or -- strip mode -- one strip
        // at a time behind the tables, parked in this CTA's global scratch from which the spectral phase reads them
        float *const R1s = STRIPS ? Bf + g.lutp : Bf;
        const float *const R1 = STRIPS ? r1g_base + (long)blockIdx.x * r1g_stride : Bf;
        unsigned char *const raw = reinterpret_cast<unsigned char *>(Bf + g.lutp);
        float *const GSB = Bf + g.lutp + (4 * hr + 3) * (g.raw_pitch / 4);
        float *const Ns = smem + g.oN, *const Es = smem + g.oE;
        float2 *const ACC = reinterpret_cast<float2 *>(smem + g.oACC);
        float2 *const TWR = reinterpret_cast<float2 *>(smem + g.oTWR), *const TWC = reinterpret_cast<float2 *>(smem + g.oTWC);
        float *const wy_s = smem + g.oWY, *const wx_s = smem + g.oWX, *const red = smem + g.oRED;

        // ---------------------------------------------------------------- P0: staging
        int l = box.l, t = box.t, r = box.r, b = box.b;
        if (t > b) { const int q = t; t = b; b = q; }          // top/drawlib.c:203-215
        if (l > r) { const int q = l; l = r; r = q; }
        const int rows_s = b - t + 1, cols_s = r - l + 1;
        const bool identity = (p.gray == nullptr) && rows_s == rows && cols_s == cols;
        const bool can_stage = identity && (((uintptr_t)frame | (uintptr_t)p.frame_stride) & 15) == 0 && rows_s <= 4 * hr + 3;
        if (warp == 0) {
            if (lane == 0) {
                mbar_expect_tx(&mbar_lut, lut_smem ? (uint32_t)(n_rs + n_bn) * 4u : 0u);
                if (lut_smem) { bulk_g2s(Bf, p.tab.rsrc_tab, n_rs * 4, &mbar_lut); bulk_g2s(Bf + n_rs, p.tab.bin2_tab, n_bn * 4, &mbar_lut); }
            }
        } else if (warp == 1) {
            // the model is not needed before the spectral phase: pull it (and alpha) into L2 now
            if (lane == 0 && (MODE == KCF_MODE_PREDICT || !first_update)) prefetch_l2_bulk(model, (uint32_t)(KCF_CHAN * S * 8) & ~15u);
            if (lane == 1) prefetch_l2_bulk(alpha, (uint32_t)(S * 4) & ~15u);
        }
        {
            const float2 *twr = at.tw + any_off(hr), *twc = at.tw + any_off(wc);
            const float *hy = at.hann + any_off(hr), *hx = at.hann + any_off(wc);
            for (int i = tid; i < hr; i += NT) { TWR[i] = twr[i]; wy_s[i] = 0.5f * hy[i]; }     // halved: carries the x0.5 of hogChannels (exact scaling)
            for (int i = tid; i < wc; i += NT) { TWC[i] = twc[i]; wx_s[i] = hx[i]; }
        }

        // ---------------------------------------------------------------- P1 + P2, per strip of cell columns (one strip = the whole window unless STRIPS)
        {
            const LutConsts lk = make_lut_consts(p.tab);
            const float2 *const rsrc = lut_smem ? reinterpret_cast<const float2 *>(Bf) : p.tab.rsrc_tab;
            const uint32_t *const bn = lut_smem ? reinterpret_cast<const uint32_t *>(Bf) + n_rs : p.tab.bin2_tab;
            const int GS = g.gs, PC = g.pc, PS = g.ps;
            const float xs_f = __fdiv_rn((float)cols_s, (float)cols), ys_f = __fdiv_rn((float)rows_s, (float)rows);
            int a0 = 0;                                                     // frame-row byte offset of the staged span of the current strip
            // gray value of template pixel (x, y): staged frame bytes, the caller's gray patch, plain loads (unaligned frames), or
            // the reference's resample when the crop has another size than the template
            auto gray_at = [&](int x, int y, bool from_stage) -> float {
                if (from_stage) return bgr_gray(raw + y * g.raw_pitch + clampi(l + x, 0, Wm) * 3 - a0);
                if (p.gray != nullptr) return p.gray[(long)job * p.gray_stride + x * rows + y];
                if (identity) return bgr_gray(frame + (long)clampi(t + y, 0, Hm) * p.frame_stride + clampi(l + x, 0, Wm) * 3);
                // the reference resamples a column-major crop as if it were row-major height x width; reproduced through
                // linear indices (top/drawlib.c:542-637, called as (dst, src, rows_s, cols_s, rows_d, cols_d), top/td.cpp:357-364)
                const int k = x * rows + y;                                  // column-major template element
                const int ky = k / cols, kx = k - ky * cols;
                const float sx = __fmul_rn((float)kx, xs_f), sy = __fmul_rn((float)ky, ys_f);
                const int x0 = __float2int_rz(sx), y0 = __float2int_rz(sy);
                const float fx = __fsub_rn(sx, (float)x0), fy = __fsub_rn(sy, (float)y0);
                const float ifx = __fsub_rn(1.0f, fx), ify = __fsub_rn(1.0f, fy);
                const int x1 = (x0 + 1 >= cols_s) ? x0 : x0 + 1, y1 = (y0 + 1 >= rows_s) ? y0 : y0 + 1;
                float c[4];
                const int sidx[4] = { y0 * cols_s + x0, y0 * cols_s + x1, y1 * cols_s + x0, y1 * cols_s + x1 };
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int sc = sidx[q] / rows_s, sr = sidx[q] - sc * rows_s;     // column-major crop element
                    c[q] = bgr_gray(frame + (long)clampi(t + sr, 0, Hm) * p.frame_stride + clampi(l + sc, 0, Wm) * 3);
                }
                const float l0 = __fadd_rn(__fmul_rn(ifx, c[0]), __fmul_rn(fx, c[1]));
                const float l1 = __fadd_rn(__fmul_rn(ifx, c[2]), __fmul_rn(fx, c[3]));
                return __fadd_rn(__fmul_rn(ify, l0), __fmul_rn(fy, l1));
            };
            if (DUMP && p.dump.gray)                                        // (test hook) the whole rows x cols patch, column-major
                for (int k = tid; k < rows * cols; k += NT) { const int x = k / rows; p.dump.gray[(long)job * p.dump.stride_px + k] = gray_at(x, k - x * rows, false); }
            const bool inside = l >= 0 && l + cols - 1 <= Wm;               // no horizontal clamping needed (the common case)
            // gradient mapping: a warp keeps one chunk of 32 rows (its y, border factor and the y part of the store address are fixed)
            // and walks over pixel columns; with fewer warps than chunks it takes several chunks in turn
            const int nych = (H0 + 31) >> 5;
            const int ywarps = min(nwarps, nych), ngrp = max(1, fdiv(nwarps, jc.mg_nych));
            const int wyc = nwarps >= nych ? warp - fdiv(warp, jc.mg_nych) * nych : warp, xg = nwarps >= nych ? fdiv(warp, jc.mg_nych) : 0;
            const int CSW = STRIPS ? g.cs : wc, OSS = STRIPS ? g.oss : g.os;
            for (int cs0v = 0; cs0v < (STRIPS ? wc : 1); cs0v += CSW) {
                const int cs0 = STRIPS ? cs0v : 0, cs1 = STRIPS ? min(wc, cs0v + CSW) : wc;      // compile-time 0 / wc without strips
                // pixel columns [pxb, pxe) of the window live in the strip's (M | bin) map, local column x - pxb; [gxb, gxe) of them exist
                const int pxb = 4 * cs0 - 2, pxe = 4 * cs1 + 2, gxb = max(pxb, 0), gxe = min(pxe, W0);
                // template columns whose gray the strip needs (one more on each side for the gradient), and their bytes in a frame row
                const int ta = max(gxb - 1, 0), tb = min(gxe, cols - 1);
                const int xa = clampi(l + ta, 0, Wm), xb = clampi(l + tb, 0, Wm);
                const int a1 = ((xb + 1) * 3 + 15) & ~15;
                a0 = (xa * 3) & ~15;
                const bool staged = can_stage && (a1 - a0) <= g.raw_pitch;
                if (STRIPS) __syncthreads();                           // the previous strip is done with the staging area, its histograms and its map
                if (warp == 0) {
                    if (lane == 0) mbar_expect_tx(&mbar_raw, staged ? (uint32_t)rows_s * (uint32_t)(a1 - a0) : 0u);
                    __syncwarp();
                    if (staged)
                        for (int y = lane; y < rows_s; y += 32)
                            bulk_g2s(raw + y * g.raw_pitch, frame + (long)clampi(t + y, 0, Hm) * p.frame_stride + a0, a1 - a0, &mbar_raw);
                }
                // zero border of the map: the columns outside the window entirely, rows y+2 in {0, 1, H0+2, H0+3} of the others
                {
                    const int ncol = pxe - pxb;
                    for (int k = tid; k < 4 * PC; k += NT) {
                        const int cq = k / PC, o = k - cq * PC, lc = cq < 2 ? cq : ncol - 4 + cq;      // local columns 0, 1, ncol-2, ncol-1
                        const int x = pxb + lc;
                        if (x < 0 || x >= W0) MB[lc * PC + o] = 0u;
                    }
                    for (int k = tid; k < 4 * (gxe - gxb); k += NT) {
                        const int lc = gxb - pxb + (k >> 2), q = k & 3;
                        MB[lc * PC + q * PS + (q < 2 ? 0 : hr)] = 0u;        // y+2 = 0, 1 -> (sub 0, 1; idx 0);  y+2 = H0+2, H0+3 -> (sub 2, 3; idx hr)
                    }
                }
                if (warp == 0) { mbar_wait(&mbar_raw, phase_raw); if (cs0 == 0) mbar_wait(&mbar_lut, phase); }
                phase_raw ^= 1u;                                       // one arrival per strip, bytes or not
                __syncthreads();
                for (int xs = gxb; xs < gxe; xs += g.xw) {
                    const int xe = min(xs + g.xw, gxe), nx = xe - xs + 2;
                    // gray of template pixels (x, y), x in [xs-1, xe], y in [-1, H0], coordinates clamped into the template: the clamped
                    // apron turns grad1's one-sided border differences (gradientMex.cpp:15-37) into plain differences
                    for (int yy = warp; yy < H0 + 2; yy += nwarps) {
                        const int y = clampi(yy - 1, 0, rows - 1);
                        if (staged && inside) {
                            const unsigned char *rrow = raw + y * g.raw_pitch + l * 3 - a0;
                            for (int lx = lane; lx < nx; lx += 32) GSB[lx * GS + yy] = bgr_gray(rrow + clampi(xs - 1 + lx, 0, cols - 1) * 3);
                        } else {
                            for (int lx = lane; lx < nx; lx += 32) GSB[lx * GS + yy] = gray_at(clampi(xs - 1 + lx, 0, cols - 1), y, staged);
                        }
                    }
                    __syncthreads();
                    // libhog/gradientMex.cpp:15-37 (grad1), :59-100 (gradMag, d=1, full=true), :112-145 (gradQuantize, nearest bin)
                    if (xg < ngrp)
                        for (int yc = wyc; yc < nych; yc += ywarps) {
                            const int y = (yc << 5) + lane;
                            if (y >= H0) continue;
                            const float ry = (y == 0 || y == rows - 1) ? 1.f : .5f;
                            const float *gp = GSB + (xg + 1) * GS + y + 1;
                            uint32_t *mp = MB + (xs + xg - pxb) * PC + ((y + 2) & 3) * PS + ((y + 2) >> 2);
                            for (int lx = xg; lx < xe - xs; lx += ngrp, gp += ngrp * GS, mp += ngrp * PC) {
                                const int x = xs + lx;
                                const float rx = (x == 0 || x == cols - 1) ? 1.f : .5f;
                                const float gx = __fmul_rn(__fsub_rn(gp[GS], gp[-GS]), rx);
                                const float gy = __fmul_rn(__fsub_rn(gp[1], gp[-1]), ry);
                                const uint32_t mb = grad_pixel_k(gx, gy, rsrc, bn, lk);
                                *mp = mb;
                                if (DUMP && p.dump.m0) {
                                    const long idx = (long)job * p.dump.stride_px + x * H0 + y;
                                    p.dump.m0[idx] = __uint_as_float(mb & ~31u); p.dump.bin[idx] = (int)(mb & 31u);
                                }
                            }
                        }
                    __syncthreads();
                }
                // ---- P2: cell histograms of the strip (ordered gather); up to four cells per thread interleaved (independent
                // dependency chains), larger strips go round again
                const int ncell = (cs1 - cs0) * hr;
                for (int cb = 0; cb < ncell; cb += 4 * NT) {
                    const int left = ncell - cb;
                    if (left <= NT) gather_cells<1>(g, jc.mg_hr, MB, R1s, OSS, Es, cb + tid, ncell, cs0, NT);
                    else if (left <= 2 * NT) gather_cells<2>(g, jc.mg_hr, MB, R1s, OSS, Es, cb + tid, ncell, cs0, NT);
                    else if (left <= 3 * NT) gather_cells<3>(g, jc.mg_hr, MB, R1s, OSS, Es, cb + tid, ncell, cs0, NT);
                    else gather_cells<4>(g, jc.mg_hr, MB, R1s, OSS, Es, cb + tid, ncell, cs0, NT);
                }
                if (STRIPS) {
                    // park the strip's histograms in the CTA's global scratch: plane o of the window at o * g.os, cells column-major
                    __syncthreads();
                    float *const R1w = r1g_base + (long)blockIdx.x * r1g_stride + cs0 * hr;
                    for (int o = warp; o < 18; o += nwarps)
                        for (int c2 = lane; c2 < ncell; c2 += 32) R1w[o * g.os + c2] = R1s[o * OSS + c2];
                }
            }
            phase ^= 1u;                                                   // the tables arrive once per job
        }
        __syncthreads();
        const int OS = g.os, RS = g.rs;
        if (DUMP && p.dump.r1) {
            float *d = p.dump.r1 + (long)job * p.dump.stride_cell * 18;
            for (int i = tid; i < 18 * NB; i += NT) { const int o = i / NB, c2 = i - o * NB; d[i] = R1[o * OS + c2]; }
        }

        // ---------------------------------------------------------------- P3: 2x2 block normalisers (hogNormMatrix, gradientMex.cpp:236-253)
        for (int i = tid; i < (wc - 1) * (hr - 1); i += NT) {
            const int x = fdiv(i, jc.mg_hr1), y = i - x * (hr - 1);
            const float eps = 1e-4f / 4 / 4 / 4 / 4 / 4;
            float e = __fadd_rn(Es[x * hr + y], Es[x * hr + y + 1]);
            e = __fadd_rn(e, Es[(x + 1) * hr + y]);
            e = __fadd_rn(e, Es[(x + 1) * hr + y + 1]);
            e = __fadd_rn(e, eps);
            const float nv = __fdiv_rn(1.0f, __fsqrt_rn(e));
            float *const q = Ns + (x + 1) * (hr + 1) + y + 1;
            const bool xl = x == 0, xh = x == wc - 2, yl = y == 0, yh = y == hr - 2;
            q[0] = nv;
            if (yl) q[-1] = nv;
            if (yh) q[1] = nv;
            if (xl) { q[-(hr + 1)] = nv; if (yl) q[-(hr + 1) - 1] = nv; if (yh) q[-(hr + 1) + 1] = nv; }
            if (xh) { q[(hr + 1)] = nv; if (yl) q[(hr + 1) - 1] = nv; if (yh) q[(hr + 1) + 1] = nv; }
        }
        __syncthreads();
        if (DUMP && p.dump.nrm) {
            float *d = p.dump.nrm + (long)job * ((wc + 1) * (hr + 1));
            for (int i = tid; i < (wc + 1) * (hr + 1); i += NT) d[i] = Ns[i];
        }

        // ---------------------------------------------------------------- P4..P6: spectral phase, one tile of channels at a time
        const bool first = (MODE == KCF_MODE_UPDATE) && first_update;
        const float fac = first ? 1.0f : p.factor;                         // kcf.cpp:443
        const float omf = __fsub_rn(1.0f, fac);
        const bool need_model = (MODE == KCF_MODE_PREDICT) || !first;
        float2 *const X0 = reinterpret_cast<float2 *>(smem);
        float2 *const X1 = X0 + g.xbuf;
        double sxx = 0.0, syy = 0.0;         // Gaussian kernel: this thread's share of sum |xf|^2 and sum |model|^2 over the FULL spectrum (FP64: they are subtracted from each other later)
        for (int c0 = 0; c0 < KCF_CHAN; c0 += g.tc) {
            const int c1 = min(KCF_CHAN, c0 + g.tc), nch = c1 - c0;
            const int nbr = nch * JP, BPr = nbr | 1;                      // row pass: sequences (channel, column pair), batch pitch
            const int nbc = nch * SK, BPc = nbc | 1;                      // column pass: sequences (channel, bin k)
            // ---- features: two adjacent cell columns (2jp, 2jp+1) as the real / imaginary part of one sequence along the rows.
            // One thread per (column pair, row) loops over the channels of the tile: the six block normalisers around its two
            // cells and the window weights stay in registers.
            for (int item = tid; item < JP * hr; item += NT) {
                const int jp = fdiv(item, jc.mg_hr), i = item - jp * hr, j0 = 2 * jp;
                const bool two = j0 + 1 < wc;
                const float *const np0 = Ns + j0 * (hr + 1) + i;
                // GETT(0), GETT(1), GETT(hb1), GETT(hb1 + 1) of cell (j, i): N[j+1][i+1], N[j+1][i], N[j][i+1], N[j][i]
                const float a00 = np0[0], a01 = np0[1], a10 = np0[hr + 1], a11 = np0[hr + 2];
                const float a20 = two ? np0[2 * (hr + 1)] : 0.f, a21 = two ? np0[2 * (hr + 1) + 1] : 0.f;
                const float wyi = wy_s[i];
                const float w0 = __fmul_rn(wyi, wx_s[j0]), w1 = two ? __fmul_rn(wyi, wx_s[j0 + 1]) : 0.f;
                const float *const r0 = R1 + j0 * RS + i, *const r1p = r0 + (two ? RS : 0);
                float2 *const dst = X0 + i * BPr + jp;
                const int n1 = max(0, min(c1, 27) - c0);                  // type-1 channels of this tile (gradientMex.cpp:266-270)
                for (int cc = 0; cc < n1; ++cc) {
                    const int c = c0 + cc;
                    const int ro = (c < 18 ? c : c - 18) * OS;
                    float rv0 = r0[ro], rv1 = r1p[ro];
                    if (c >= 18) { rv0 = __fadd_rn(rv0, r0[ro + 9 * OS]); rv1 = __fadd_rn(rv1, r1p[ro + 9 * OS]); }   // R2 = R1[o] + R1[o+9] (gradientMex.cpp:308-309)
                    float h0 = __fadd_rn(fminf(__fmul_rn(rv0, a11), 0.2f), fminf(__fmul_rn(rv0, a10), 0.2f));
                    h0 = __fadd_rn(h0, fminf(__fmul_rn(rv0, a01), 0.2f));
                    h0 = __fadd_rn(h0, fminf(__fmul_rn(rv0, a00), 0.2f));
                    float h1 = __fadd_rn(fminf(__fmul_rn(rv1, a21), 0.2f), fminf(__fmul_rn(rv1, a20), 0.2f));
                    h1 = __fadd_rn(h1, fminf(__fmul_rn(rv1, a11), 0.2f));
                    h1 = __fadd_rn(h1, fminf(__fmul_rn(rv1, a10), 0.2f));
                    const float f0 = __fmul_rn(h0, w0), f1 = two ? __fmul_rn(h1, w1) : 0.f;       // (hsum * 0.5) * (wy * wx), kcf.cpp:251-258
                    dst[cc * JP] = make_float2(f0, f1);
                    if (DUMP && p.dump.feat) {
                        float *d = p.dump.feat + (long)job * p.dump.stride_cell * KCF_CHAN + c * NB + j0 * hr + i;
                        d[0] = f0; if (two) d[hr] = f1;
                    }
                }
                if (c1 > 27) {
                    // texture channels 27..30 (hogChannels type 2, gradientMex.cpp:271-275): the 18 orientation loads of a cell are
                    // shared by its four block normalisers; the results are doubled (exact) because the window rows are stored halved
                    float t0[4] = { 0.f, 0.f, 0.f, 0.f }, t1[4] = { 0.f, 0.f, 0.f, 0.f };
#pragma unroll 6
                    for (int o = 0; o < 18; ++o) {
                        const float rv0 = r0[o * OS], rv1 = r1p[o * OS];
                        t0[0] = __fadd_rn(t0[0], __fmul_rn(fminf(__fmul_rn(rv0, a11), 0.2f), .2357f));
                        t0[1] = __fadd_rn(t0[1], __fmul_rn(fminf(__fmul_rn(rv0, a10), 0.2f), .2357f));
                        t0[2] = __fadd_rn(t0[2], __fmul_rn(fminf(__fmul_rn(rv0, a01), 0.2f), .2357f));
                        t0[3] = __fadd_rn(t0[3], __fmul_rn(fminf(__fmul_rn(rv0, a00), 0.2f), .2357f));
                        t1[0] = __fadd_rn(t1[0], __fmul_rn(fminf(__fmul_rn(rv1, a21), 0.2f), .2357f));
                        t1[1] = __fadd_rn(t1[1], __fmul_rn(fminf(__fmul_rn(rv1, a20), 0.2f), .2357f));
                        t1[2] = __fadd_rn(t1[2], __fmul_rn(fminf(__fmul_rn(rv1, a11), 0.2f), .2357f));
                        t1[3] = __fadd_rn(t1[3], __fmul_rn(fminf(__fmul_rn(rv1, a10), 0.2f), .2357f));
                    }
#pragma unroll
                    for (int bq = 0; bq < 4; ++bq) {
                        const int c = 27 + bq;
                        const float f0 = __fmul_rn(t0[bq] + t0[bq], w0), f1 = two ? __fmul_rn(t1[bq] + t1[bq], w1) : 0.f;
                        if (c >= c0 && c < c1) dst[(c - c0) * JP] = make_float2(f0, f1);
                        if (DUMP && p.dump.feat && c >= c0 && c < c1) {
                            float *d = p.dump.feat + (long)job * p.dump.stride_cell * KCF_CHAN + c * NB + j0 * hr + i;
                            d[0] = f0; if (two) d[hr] = f1;
                        }
                    }
                }
            }
            __syncthreads();
            // ---- transform along the rows (length hr), nch * JP sequences
            float2 *const Zr = fft_batch<-1>(X0, X1, hr, jc.nst_r, jc.rad_r, jc.ns_r, jc.mg_ns_r, jc.mg_small, nbr, BPr, TWR, warp, lane, nwarps);
            float2 *const Rb = (Zr == X0) ? X1 : X0;
            // ---- split the two real columns of each sequence, re-batch for the column pass: Rb[j * BPc + cc * SK + k]
            for (int q = tid; q < nbr * SK; q += NT) {
                const int bq = fdiv(q, jc.mg_sk), k = q - bq * SK;        // k fastest: conflict-free on both sides (odd pitches)
                const int cc = fdiv(bq, jc.mg_jp), jp = bq - cc * JP;
                const float2 zk = Zr[k * BPr + bq], zn = Zr[(k == 0 ? 0 : hr - k) * BPr + bq];
                // A = (Z[k] + conj(Z[n-k])) / 2,  B = (Z[k] - conj(Z[n-k])) / (2i)
                float2 *d = Rb + (2 * jp) * BPc + cc * SK + k;
                d[0] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                if (2 * jp + 1 < wc) d[BPc] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
            }
            __syncthreads();
            // ---- transform along the columns (length wc), nch * SK sequences
            float2 *const Xf = fft_batch<-1>(Rb, Zr, wc, jc.nst_c, jc.rad_c, jc.ns_c, jc.mg_ns_c, jc.mg_small, nbc, BPc, TWC, warp, lane, nwarps);
            // ---- spectral products and the channel sum, in channel order; e = j' * SK + k is the FFTW half-spectrum index (kcf.cpp:180-186)
            for (int e = tid; e < S; e += NT) {
                const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK;
                float2 acc = (c0 == 0) ? make_float2(0.f, 0.f) : ACC[e];
                const float hw = (k == 0 || 2 * k == hr) ? 1.f : 2.f;     // weight of a half-spectrum bin in a sum over the full spectrum
                const float2 *xp = Xf + jq * BPc + k;
                float2 *mp = model + (long)c0 * S + e;
                for (int cb = 0; cb < nch; cb += 8) {
                    // up to eight model values in flight per thread before the first one is used
                    float2 mv[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) mv[u] = (need_model && cb + u < nch) ? mp[(long)(cb + u) * S] : make_float2(0.f, 0.f);
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int cc = cb + u;
                        if (cc >= nch) break;
                        const float2 v = xp[cc * SK];
                        if (DUMP && p.dump.spec) p.dump.spec[(long)job * p.dump.stride_spec * KCF_CHAN + (c0 + cc) * S + e] = v;
                        if (gauss) { sxx += (double)(hw * (v.x * v.x + v.y * v.y)); syy += (double)(hw * (mv[u].x * mv[u].x + mv[u].y * mv[u].y)); }
                        float2 o;
                        if (MODE == KCF_MODE_PREDICT) {
                            o = make_float2(v.x * mv[u].x + v.y * mv[u].y, v.y * mv[u].x - v.x * mv[u].y);       // xf * conj(model), kcf.cpp:306-345
                        } else {
                            o = make_float2(v.x * v.x + v.y * v.y, 0.f);                                         // |xf|^2, kcf.cpp:269-293
                            // model = (1-f) model + f xf, kcf.cpp:380-395 (f = 1 on the first update: the old model drops out)
                            mp[(long)cc * S] = first ? v : make_float2(__fadd_rn(__fmul_rn(omf, mv[u].x), __fmul_rn(fac, v.x)),
                                                                       __fadd_rn(__fmul_rn(omf, mv[u].y), __fmul_rn(fac, v.y)));
                        }
                        if (c0 + cc == 0) acc = o; else { acc.x = __fadd_rn(acc.x, o.x); acc.y = __fadd_rn(acc.y, o.y); }
                    }
                }
                ACC[e] = acc;
            }
            __syncthreads();
        }

        // ---------------------------------------------------------------- P6 / P7
        const int BPk = SK | 1, BPj = JP | 1;
        // single-channel 2-D transforms on the two ping-pong buffers.  Spectrum layout [j' * BPk + k] (k <= hr/2), spatial layout packed:
        // [i * BPj + jp] = (value of column 2jp, value of column 2jp+1) at row i.  Both return the buffer holding the result.
        auto inv2d = [&](float2 *in, float2 *other) -> float2 * {
            float2 *const Yc = fft_batch<+1>(in, other, wc, jc.nst_c, jc.rad_c, jc.ns_c, jc.mg_ns_c, jc.mg_small, SK, BPk, TWC, warp, lane, nwarps);
            float2 *const Zb = (Yc == in) ? other : in;
            // c2r along the rows (kcf.cpp:397-399): rebuild the Hermitian sequences of two columns, A + iB, and transform them together.
            // FFTW's c2r takes the DC and (even length) Nyquist bins as real.
            for (int q = tid; q < JP * hr; q += NT) {
                const int jp = fdiv(q, jc.mg_hr), i = q - jp * hr;
                const bool up = 2 * i > hr;
                const int ks = up ? hr - i : i;
                const bool realbin = (i == 0) || (2 * i == hr);
                float2 A = Yc[(2 * jp) * BPk + ks];
                float2 B = (2 * jp + 1 < wc) ? Yc[(2 * jp + 1) * BPk + ks] : make_float2(0.f, 0.f);
                if (up) { A.y = -A.y; B.y = -B.y; }
                if (realbin) { A.y = 0.f; B.y = 0.f; }
                Zb[i * BPj + jp] = make_float2(A.x - B.y, A.y + B.x);
            }
            __syncthreads();
            return fft_batch<+1>(Zb, Yc, hr, jc.nst_r, jc.rad_r, jc.ns_r, jc.mg_ns_r, jc.mg_small, JP, BPj, TWR, warp, lane, nwarps);
        };
        auto fwd2d = [&](float2 *in, float2 *other) -> float2 * {
            float2 *const Zr = fft_batch<-1>(in, other, hr, jc.nst_r, jc.rad_r, jc.ns_r, jc.mg_ns_r, jc.mg_small, JP, BPj, TWR, warp, lane, nwarps);
            float2 *const Rb = (Zr == in) ? other : in;
            for (int q = tid; q < JP * SK; q += NT) {
                const int jp = fdiv(q, jc.mg_sk), k = q - jp * SK;
                const float2 zk = Zr[k * BPj + jp], zn = Zr[(k == 0 ? 0 : hr - k) * BPj + jp];
                float2 *d = Rb + (2 * jp) * BPk + k;
                d[0] = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y - zn.y));
                if (2 * jp + 1 < wc) d[BPk] = make_float2(0.5f * (zk.y + zn.y), 0.5f * (zn.x - zk.x));
            }
            __syncthreads();
            return fft_batch<-1>(Rb, Zr, wc, jc.nst_c, jc.rad_c, jc.ns_c, jc.mg_ns_c, jc.mg_small, SK, BPk, TWC, warp, lane, nwarps);
        };
        // label spectrum yf[j'][k] = LY[j'] * LX[k]: the per-N tables hold the 1-D transforms for the reference's fixed label sigma; a
        // per-track sigma (extension: sqrt(target area) x osf / cell) gets its two 1-D transforms computed here, by the definition
        const double2 *ly = at.lab + any_off(wc), *lx = at.lab + any_off(hr);
        if (EXT && MODE == KCF_MODE_UPDATE && p.ext.osf > 0.f) {
            double2 *const LAB = reinterpret_cast<double2 *>(smem + g.oLAB);          // [0, SK): rows (length hr); [SK, SK + wc): columns (length wc)
            const int tw_ = tbox.r >= tbox.l ? tbox.r - tbox.l + 1 : tbox.l - tbox.r + 1, th_ = tbox.b >= tbox.t ? tbox.b - tbox.t + 1 : tbox.t - tbox.b + 1;
            const float lsig = sqrtf((float)tw_ * (float)th_) * p.ext.osf * (1.0f / KCF_CELL);
            const float sinv = (float)(1.0 / ((double)lsig * (double)lsig));
            for (int q = tid; q < SK + wc; q += NT) {
                const bool isrow = q < SK;
                const int n = isrow ? hr : wc, kk = isrow ? q : q - SK;
                const float2 *tw = isrow ? TWR : TWC;
                double sr = 0.0, si = 0.0;
                int ph = 0;
                for (int b2 = 0; b2 < n; ++b2) {
                    const int x = b2 <= n - 1 - n / 2 ? b2 : b2 - n;                 // position of index b2 after the circular shift of the label
                    const double gv = (double)(float)exp(-0.5 * (double)(x * x) * (double)sinv);
                    const float2 w = tw[ph];
                    sr += gv * (double)w.x; si += gv * (double)w.y;
                    ph += kk; if (ph >= n) ph -= n;
                }
                LAB[q] = make_double2(sr, si);
            }
            __syncthreads();
            lx = LAB; ly = LAB + SK;
        }
        float *const alpha_im = gauss ? p.alpha_im + (long)slot * p.alpha_stride : nullptr;
        float2 *Rz = nullptr;
        if (gauss) {
            // ---- Gaussian kernel correlation: k = exp(-max(0, |x|^2 + |z|^2 - 2 x*z) / (sigma^2 numel)), kf = fft2(k)
            // |x|^2, |z|^2 by Parseval from the half spectra (deterministic: per-thread partial sums, fixed reduction tree)
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { sxx += __shfl_down_sync(0xFFFFFFFFu, sxx, off); syy += __shfl_down_sync(0xFFFFFFFFu, syy, off); }
            double *const redd = reinterpret_cast<double *>(red);                    // 64 floats = 32 doubles: nwarps <= 16 here or 32 with one value each
            if (lane == 0) { if (nwarps <= 16) { redd[warp] = sxx; redd[16 + warp] = syy; } else { redd[warp] = sxx; } }
            for (int e = tid; e < S; e += NT) { const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK; X0[jq * BPk + k] = ACC[e]; }
            __syncthreads();
            double xxd = 0.0, yyd = 0.0;
            if (nwarps <= 16) { for (int w = 0; w < nwarps; ++w) { xxd += redd[w]; yyd += redd[16 + w]; } }
            else {
                for (int w = 0; w < nwarps; ++w) xxd += redd[w];
                __syncthreads();
                if (lane == 0) redd[warp] = syy;
                __syncthreads();
                for (int w = 0; w < nwarps; ++w) yyd += redd[w];
            }
            const float inv_n = 1.0f / (float)NB;
            if (MODE == KCF_MODE_UPDATE) yyd = xxd;                                   // autocorrelation
            const float xx = 0.f, yy = (float)((xxd + yyd) * (double)inv_n);          // (|x|^2 + |z|^2) / N, summed in FP64
            float2 *const Rk = inv2d(X0, X1);
            float2 *const Ro = (Rk == X0) ? X1 : X0;
            const float escale = 1.0f / (p.ext.sigma * p.ext.sigma * (float)(NB * KCF_CHAN));
            for (int q = tid; q < JP * hr; q += NT) {
                const int jp = fdiv(q, jc.mg_hr), i = q - jp * hr;
                const float2 z = Rk[i * BPj + jp];
                // k - 1 is transformed, not k: with HOG features the kernel is flat (k = 1 - small), and alpha = yf / (kf + lambda) divides by
                // the transform of that small part -- expm1 keeps its relative precision, the "1" is added to the DC bin afterwards (= N)
                const float k0 = expm1f(-fmaxf(0.f, xx + yy - 2.f * z.x * inv_n) * escale);
                const float k1 = (2 * jp + 1 < wc) ? expm1f(-fmaxf(0.f, xx + yy - 2.f * z.y * inv_n) * escale) : 0.f;
                Rk[i * BPj + jp] = make_float2(k0, k1);
            }
            __syncthreads();
            float2 *const KF = fwd2d(Rk, Ro);
            if (tid == 0) KF[0].x += (float)NB;
            __syncthreads();
            if (MODE == KCF_MODE_UPDATE) {
                for (int e = tid; e < S; e += NT) {
                    const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK;
                    const float2 kf = KF[jq * BPk + k];
                    if (DUMP && p.dump.kf) p.dump.kf[(long)job * p.dump.stride_spec + e] = kf.x;
                    const double2 gy = ly[jq], gx = lx[k];
                    const float yr = (float)(gy.x * gx.x - gy.y * gx.y), yi = (float)(gy.x * gx.y + gy.y * gx.x);
                    // alphaf = yf / (kf + lambda), complex
                    const float dr = kf.x + p.lamda, di = kf.y, dn = 1.0f / (dr * dr + di * di);
                    const float ar = (yr * dr + yi * di) * dn, ai = (yi * dr - yr * di) * dn;
                    alpha[e] = first ? ar : omf * alpha[e] + fac * ar;
                    alpha_im[e] = first ? ai : omf * alpha_im[e] + fac * ai;
                }
            } else {
                for (int e = tid; e < S; e += NT) {
                    const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK;
                    const float2 kf = KF[jq * BPk + k];
                    const float ar = alpha[e], ai = alpha_im[e];
                    const float2 zf = make_float2(ar * kf.x - ai * kf.y, ar * kf.y + ai * kf.x);
                    KF[jq * BPk + k] = zf;
                    if (DUMP && p.dump.zf) p.dump.zf[(long)job * p.dump.stride_spec + e] = zf;
                }
                __syncthreads();
                Rz = inv2d(KF, (KF == X0) ? X1 : X0);
            }
        } else if (MODE == KCF_MODE_UPDATE) {
            for (int e = tid; e < S; e += NT) {
                const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK;
                const float kf = __fmul_rn(ACC[e].x, jc.norm);                                   // kcf.cpp:295-303
                if (DUMP && p.dump.kf) p.dump.kf[(long)job * p.dump.stride_spec + e] = kf;
                // Re(yf): the label is an outer product of two 1-D Gaussians, so its 2-D transform is the product of their 1-D ones
                const double2 gy = ly[jq], gx = lx[k];
                const float yf = (float)(gy.x * gx.x - gy.y * gx.y);
                const float an = __fdiv_rn(yf, __fadd_rn(kf, p.lamda));                          // kcf.cpp:373
                alpha[e] = first ? an : __fadd_rn(__fmul_rn(omf, alpha[e]), __fmul_rn(fac, an)); // kcf.cpp:374
            }
        } else {
            // predict: zf = (sum) * alpha * norm (kcf.cpp:356-357), batch-major for the inverse column transform: [j' * BPk + k]
            for (int e = tid; e < S; e += NT) {
                const int jq = fdiv(e, jc.mg_sk), k = e - jq * SK;
                const float al = alpha[e];
                float2 acc = ACC[e];
                acc.x = __fmul_rn(__fmul_rn(acc.x, al), jc.norm);
                acc.y = __fmul_rn(__fmul_rn(acc.y, al), jc.norm);
                X0[jq * BPk + k] = acc;
                if (DUMP && p.dump.zf) p.dump.zf[(long)job * p.dump.stride_spec + e] = acc;
            }
            __syncthreads();
            Rz = inv2d(X0, X1);
        }
        if (MODE == KCF_MODE_UPDATE) {
            if (tid == 0) {
                // tracker_update, kcf.cpp:462-476 (box = the window; with padding the target's size is remembered for the way back)
                int bl = box.l, bt = box.t, br = box.r, bb = box.b;
                meta->pos = box;
                meta->scale_horiz = __fdiv_rn((float)(br - bl + 1), (float)cols);
                meta->scale_vert = __fdiv_rn((float)(bb - bt + 1), (float)rows);
                meta->first_update = 0;
                if (EXT) {
                    meta->tw = tbox.r >= tbox.l ? tbox.r - tbox.l + 1 : tbox.l - tbox.r + 1;
                    meta->th = tbox.b >= tbox.t ? tbox.b - tbox.t + 1 : tbox.t - tbox.b + 1;
                }
            }
            continue;
        }
        // response[j][i] = Re / Im of Rz[i][j >> 1]; first maximum in memory order (j outer, i inner), strict '>' from -99999 (kcf.cpp:402-417)
        float best = -99999.0f; int besti = 0x7FFFFFFF;
        for (int idx = tid; idx < NB; idx += NT) {
            const int j = fdiv(idx, jc.mg_hr), i = idx - j * hr;
            const float2 zz = Rz[i * BPj + (j >> 1)];
            const float v = (j & 1) ? zz.y : zz.x;
            if (DUMP && p.dump.resp) p.dump.resp[(long)job * p.dump.stride_cell + idx] = v;
            if (v > best) { best = v; besti = idx; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_down_sync(0xFFFFFFFFu, best, off);
            const int oi = __shfl_down_sync(0xFFFFFFFFu, besti, off);
            if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
        }
        int *const redi = reinterpret_cast<int *>(red) + 32;
        if (lane == 0) { red[warp] = best; redi[warp] = besti; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < nwarps; ++w) { const float ov = red[w]; const int oi = redi[w]; if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; } }
            int vd = 1, hd = 1;                                            // the reference leaves these uninitialised when nothing beats -99999
            if (besti != 0x7FFFFFFF) { hd = besti / hr + 1; vd = besti - (hd - 1) * hr + 1; }
            if (DUMP && p.dump.peak) { p.dump.peak[2 * job] = vd; p.dump.peak[2 * job + 1] = hd; }
            // extension: parabolic refinement of the peak from its circular neighbours, in cells
            float sdv = 0.f, sdh = 0.f;
            if (EXT && p.ext.subpixel && besti != 0x7FFFFFFF) {
                const int i0 = vd - 1, j0 = hd - 1;
                auto rz = [&](int j, int i) { const float2 zz = Rz[i * BPj + (j >> 1)]; return (j & 1) ? zz.y : zz.x; };
                const float c = rz(j0, i0);
                const float up = rz(j0, i0 == 0 ? hr - 1 : i0 - 1), dn = rz(j0, i0 == hr - 1 ? 0 : i0 + 1);
                const float lf = rz(j0 == 0 ? wc - 1 : j0 - 1, i0), rt = rz(j0 == wc - 1 ? 0 : j0 + 1, i0);
                const float d1 = 2.f * c - dn - up, d2 = 2.f * c - rt - lf;
                sdv = d1 != 0.f ? 0.5f * (dn - up) / d1 : 0.f;
                sdh = d2 != 0.f ? 0.5f * (rt - lf) / d2 : 0.f;
            }
            if (EXT) { meta->sub_dv = sdv; meta->sub_dh = sdh; }
            if (vd > hr / 2) vd -= hr;                                     // kcf.cpp:419-420
            if (hd > wc / 2) hd -= wc;
            mot_bbox_t pos = meta->pos;
            const float dv = __fmul_rn((float)KCF_CELL * ((float)(vd - 1) + sdv), meta->scale_vert);      // sdv = sdh = 0 without the extension: the reference's expression
            const float dh = __fmul_rn((float)KCF_CELL * ((float)(hd - 1) + sdh), meta->scale_horiz);
            pos.t = __float2int_rz(__fadd_rn((float)pos.t, dv));           // kcf.cpp:423-426 (float math, truncation)
            pos.b = __float2int_rz(__fadd_rn((float)pos.b, dv));
            pos.l = __float2int_rz(__fadd_rn((float)pos.l, dh));
            pos.r = __float2int_rz(__fadd_rn((float)pos.r, dh));
            meta->pos = pos;
            if (padded) pos = kcf_unpad_box(pos, meta->tw, meta->th);      // extension: report the target, not the window
            if (p.clamp_to_frame) {                                        // top/td.cpp:378-381
                pos.l = clampi(pos.l, 0, p.frame_w - 1); pos.r = clampi(pos.r, 0, p.frame_w - 1);
                pos.t = clampi(pos.t, 0, p.frame_h - 1); pos.b = clampi(pos.b, 0, p.frame_h - 1);
            }
            p.boxes[bi] = pos;
        }
    }
}

}  // namespace mot
