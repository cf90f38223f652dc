// fft_reg.cuh -- small power-of-two FFTs held entirely in registers (one transform per thread).
//
// Replaces the FFTW 3.3.5 plans the reference creates per tracker (trackers/kcf.cpp:178-195,
// executed at :265 and :399).  A 2-D r2c of a 32x32 cell grid is 32 length-32 real transforms
// followed by 17 length-32 complex ones: far too small for a library call per track, so every
// thread owns one 1-D transform, fully unrolled, twiddles folded into immediates; the 2-D
// structure (and the transposes between the two passes) lives in shared memory (kcf_fused.cu).
//
// Decimation-in-frequency radix-2; the result is left in bit-reversed order and read back
// through the compile-time permutation brev<N>(k), which costs nothing once unrolled.
#pragma once
#include <cuda_runtime.h>
#include "twiddles64.h"

namespace mot {

template <int N> __host__ __device__ constexpr int log2c() { return N <= 1 ? 0 : 1 + log2c<N / 2>(); }

template <int N> __host__ __device__ constexpr int brev(int k)
{
    int r = 0;
    for (int b = 0; b < log2c<N>(); ++b) r |= ((k >> b) & 1) << (log2c<N>() - 1 - b);
    return r;
}

// w = exp(DIR * 2*pi*i * num / den), den | 64, evaluated at compile time after unrolling.
template <int DIR> __device__ __forceinline__ float2 cmul_tw(float2 v, int num, int den)
{
    const int idx = (num * (64 / den)) & 63;
    if (idx == 0) return v;
    if (idx == 32) return make_float2(-v.x, -v.y);
    if (idx == 16) return DIR < 0 ? make_float2(v.y, -v.x) : make_float2(-v.y, v.x);   // * (-i) or * (+i)
    if (idx == 48) return DIR < 0 ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
    const float c = tw::C64[idx];
    const float s = DIR < 0 ? -tw::S64[idx] : tw::S64[idx];
    return make_float2(v.x * c - v.y * s, v.x * s + v.y * c);
}

// In-place complex FFT of N points (N = 2,4,8,16,32,64).  DIR = -1 forward, +1 inverse (un-normalised).
// On return a[brev<N>(k)] holds X[k].
template <int N, int DIR> __device__ __forceinline__ void fft_dif(float2 (&a)[N])
{
#pragma unroll
    for (int len = N; len >= 2; len >>= 1) {
        const int half = len >> 1;
#pragma unroll
        for (int b = 0; b < N; b += len) {
#pragma unroll
            for (int i = 0; i < half; ++i) {
                const float2 u = a[b + i], v = a[b + i + half];
                a[b + i] = make_float2(u.x + v.x, u.y + v.y);
                a[b + i + half] = cmul_tw<DIR>(make_float2(u.x - v.x, u.y - v.y), i, len);
            }
        }
    }
}

// N-point FFT shared by a pair of adjacent lanes (lane parity `half`): lane 0 holds points 0..N/2-1, lane 1 holds points
// N/2..N-1.  The first radix-2 (DIF) stage crosses the pair through warp shuffles; each lane then transforms its N/2
// points on its own.  On return lane `half` holds X[2m + half] in a[brev<N/2>(m)].  `mask` = lanes taking part.
template <int N, int DIR> __device__ __forceinline__ void fft_pair(float2 (&a)[N / 2], int half, unsigned mask)
{
    constexpr int H = N / 2;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float ox = __shfl_xor_sync(mask, a[i].x, 1), oy = __shfl_xor_sync(mask, a[i].y, 1);
        const float2 sum = make_float2(a[i].x + ox, a[i].y + oy);                       // lane 0: u + v
        const float2 dif = cmul_tw<DIR>(make_float2(ox - a[i].x, oy - a[i].y), i, N);   // lane 1: (u - v) w^i, u = partner's point
        a[i] = half ? dif : sum;
    }
    __syncwarp(mask);       // callers transform shared-memory columns in place: the pair's loads precede either lane's stores
    fft_dif<H, DIR>(a);
}

// The same transform with the first stage fed straight from memory: ld(i) returns point i, both lanes of a pair read all N
// points (identical addresses inside a pair: one shared-memory broadcast, no shuffles) and each keeps its half of the
// butterflies: lane 0 u + v, lane 1 (u - v) w^i, written branch-free (sign and twiddle selected per lane).
template <int N, int DIR, class LD, class MID> __device__ __forceinline__ void fft_pair_ld(float2 (&a)[N / 2], int half, unsigned mask, LD ld, MID after_loads)
{
    constexpr int H = N / 2;
    const float sg = half ? -1.f : 1.f;
#pragma unroll
    for (int i = 0; i < H; ++i) {
        const float2 u = ld(i), v = ld(i + H);
        const float2 r = make_float2(fmaf(sg, v.x, u.x), fmaf(sg, v.y, u.y));       // exactly u + v or u - v
        const int idx = (i * (64 / N)) & 63;
        if (idx == 0) a[i] = r;
        else {
            const float c = half ? tw::C64[idx] : 1.f;
            const float sn = half ? (DIR < 0 ? -tw::S64[idx] : tw::S64[idx]) : 0.f;
            a[i] = make_float2(r.x * c - r.y * sn, r.x * sn + r.y * c);
        }
    }
    __syncwarp(mask);       // in-place callers: every load of the pair precedes either lane's stores
    after_loads();          // the points are in registers: the memory they came from may be overwritten from here on
    fft_dif<H, DIR>(a);
}
template <int N, int DIR, class LD> __device__ __forceinline__ void fft_pair_ld(float2 (&a)[N / 2], int half, unsigned mask, LD ld)
{
    fft_pair_ld<N, DIR>(a, half, mask, ld, [] {});
}

// position of packed bin k in row j of the spectrum buffer: k XOR f(j), with f chosen so that every access pattern of the
// kernels is bank-conflict free for 64-bit words: (a) the row pass writes rows j = 0..15 / 16..31 at a fixed k (f is a
// bijection on each half), (b) the column pass reads rows i and i + WC/2 in one instruction (f differs in the top bit),
// (c) it writes rows 2m and 2m + 1 in one instruction (f differs in the top bit again), (d) the channel sum reads along k.
template <int HK, int WC> __device__ __forceinline__ int fpos(int k, int j)
{
    const int fj = (((j & 1) * (HK >> 1)) | ((j >> 1) & ((HK >> 1) - 1))) ^ ((j >= WC / 2) ? (HK >> 1) : 0);
    return k ^ fj;
}

}  // namespace mot
