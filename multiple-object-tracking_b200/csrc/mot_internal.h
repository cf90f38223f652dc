// mot_internal.h -- device-side data model shared by the kernels and the C-ABI host layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/mot_b200.h"

namespace mot {

constexpr int KCF_CHAN = 31;          // trackers/kcf.cpp:157 f_chan = 32 - 1
constexpr int KCF_CELL = 4;           // trackers/kcf.cpp:488
constexpr int KCF_THREADS = 1024;     // one CTA per track job, 32 warps, 64 registers per thread
constexpr int NB_MAX = 1152;          // cells per window the fused kernel can hold in shared memory (32x32 = 1024 named shape)

// Per-track persistent state (one per slot), trackers/kcf.cpp:27-76 minus everything derivable.
struct KcfMeta {
    int rows, cols;                   // template size frozen at tracker_new (kcf.cpp:148-152)
    int hr, wc;                       // f_rows, f_cols
    mot_bbox_t pos;                   // pkcf->pos
    float scale_horiz, scale_vert;    // kcf.cpp:470-472
    int first_update;                 // kcf.cpp:209
    int size_class;                   // index into the per-size constant tables
    float2 *model_ptr;                // any-size tracks own their model / alpha (null: slot arena, see KcfLaunch)
    float *alpha_ptr;
    int tw, th;                       // extensions: size of the TARGET box at the last update (the window is tw x th times the padding)
    float sub_dv, sub_dh;             // extensions: sub-cell refinement of the last predicted peak (0 without sub-pixel mode)
};

// North-star extensions of the filter.  All off = the reference's filter (linear kernel, integer peak, window =
// box, label sigma 0.7289 cells), which is the only mode with a reference oracle; anything else is validated against the NumPy
// restatement tests/kcf_ext_numpy.py (parity unpinned by the reference) and served by the any-size kernel.
struct KcfExt {
    int gaussian;                     // Gaussian kernel correlation (Henriques et al., KCF) instead of the linear one
    float sigma;                      // its bandwidth
    int subpixel;                     // parabolic refinement of the response peak
    float padding;                    // window = target box x padding about its centre (<= 1: none)
    float osf;                        // label sigma = sqrt(target w x h) x osf / cell (0: the reference's fixed 0.7289 cells)
};

// window of a target box and back; integer arithmetic shared by host and device (unpad(pad(b)) == b)
__host__ __device__ inline mot_bbox_t kcf_pad_box(mot_bbox_t b, float p)
{
    if (b.t > b.b) { const int q = b.t; b.t = b.b; b.b = q; }
    if (b.l > b.r) { const int q = b.l; b.l = b.r; b.r = q; }
    const int w = b.r - b.l + 1, h = b.b - b.t + 1;
    const int pw = (int)((float)w * p), ph = (int)((float)h * p);
    const int nl = (b.l + b.r - pw + 1) >> 1, nt = (b.t + b.b - ph + 1) >> 1;
    b.l = nl; b.r = nl + pw - 1; b.t = nt; b.b = nt + ph - 1;
    return b;
}
__host__ __device__ inline mot_bbox_t kcf_unpad_box(mot_bbox_t wdw, int tw, int th)
{
    const int l = (wdw.l + wdw.r - tw + 2) >> 1, t = (wdw.t + wdw.b - th + 2) >> 1;
    wdw.l = l; wdw.r = l + tw - 1; wdw.t = t; wdw.b = t + th - 1;
    return wdw;
}

// Per-size constants shared by every track of the same window (kcf.cpp:203-207 are size-only).
struct KcfClassDev {
    int hr, wc;
    const float *wy, *wx;             // hann_f(hr), hann_f(wc): cos_win = wy * wx^T (kcf.cpp:124-130)
    const float *yf_re;               // Re(fft2(labels)), S floats (only the real part is ever used, kcf.cpp:373)
    float norm;                       // feature_norm_ratio = 1/(wc*hr*31) (kcf.cpp:197)
    const double2 *tw_hr, *tw_wc;     // exp(-2 pi i t / n) for n = hr, wc (any-size DFT path)
};

struct FhogTablesDev {
    const float *rsqrt_tab; int rsqrt_bits;
    const float *rcp_tab; int rcp_bits;
    const uint32_t *bin_tab; int bin_shift, bin_nseg;
    const float2 *rsrc_tab;           // {rsqrt_tab[i], rcp(rsqrt_tab[i]) / 16}: one look-up serves RCPSQRT and RCP (and the gather's 1/16)
    float rcp_cap;                    // rcp(1e10f)
    const uint32_t *bin2_tab;         // bin_tab with the wrap folded in: (thr << 10) | after << 5 | before
    uint32_t u_cap;                   // MIN(rsqrt(M2), 1e10f) saturates iff bits(M2) <= u_cap
};

// Optional stage dumps (all may be null).  Index = job * stride of that stage.
struct KcfDump {
    float *gray;        // rows*cols, column-major
    float *m0;          // w0*h0 (x-major, y fastest): M * 1/16
    int   *bin;         // w0*h0
    float *r1;          // 18*wc*hr
    float *nrm;         // (wc+1)*(hr+1)
    float *feat;        // 31*wc*hr  (windowed features = xf_tm)
    float2 *spec;       // 31*S      (xf_fq)
    float2 *zf;         // S
    float *resp;        // wc*hr
    float *kf;          // S (real part)
    int   *peak;        // 2 ints: vert_delta, horiz_delta (1-based, before wrap)
    float *margin;      // 2 floats: best, second best response
    long stride_px, stride_cell, stride_spec;   // per-job strides (elements) for px-sized / cell-sized / S-sized dumps
};

struct KcfLaunch {
    int n_jobs;
    const int *n_jobs_dev;            // optional: the job count lives on the device (min'ed with n_jobs, which then sizes the grid)
    const int *box_index;             // optional: job j's box is boxes[box_index[j]] instead of boxes[j]
    const int *slots;                 // [n] track slot of each job
    const int *frames;                // [n] frame slot of each job (ignored when gray != null)
    mot_bbox_t *boxes;                // [n] predict: in = crop box, out = predicted box; update: in = new position / crop box
    const uint8_t *const *frame_ptr;  // [n_frame_slots] device base pointers of BGR u8 frames
    const void *frame_tmaps;          // optional: [n_frame_slots][3] tensor maps (128 bytes each) of the frames as 2-D arrays of 32-bit words,
                                      // box = 108 words x {35, 67, 131} rows: the staged crop of a fixed-size fused kernel in ONE TMA load
    int frame_w, frame_h, frame_stride;
    const float *gray;                // optional: [n] pre-cropped gray patches (rows*cols col-major, stride gray_stride floats)
    long gray_stride;
    KcfMeta *meta;
    float2 *model; long model_stride; // xf_md, slot stride in float2
    float *alpha; long alpha_stride;  // slot stride in floats
    const KcfClassDev *classes;
    FhogTablesDev tab;
    int clamp_to_frame;               // fold top/td.cpp:378-381 into predict
    float factor, lamda;              // kcf.cpp:211-212
    KcfExt ext;                       // extensions (any-size kernel only); all zero = the reference's filter
    float *alpha_im;                  // Gaussian kernel: imaginary part of the (then complex) alpha, same slot stride as alpha
    KcfDump dump;
};

enum { KCF_MODE_PREDICT = 0, KCF_MODE_UPDATE = 1 };

// returns 0 when (hr, wc) has a register-FFT instantiation
int kcf_launch_fast(int mode, int hr, int wc, const KcfLaunch &p, cudaStream_t s);
size_t kcf_fast_smem_bytes(int hr, int wc);

// Any window size (hr, wc >= 2): unfused multi-kernel pipeline with per-job scratch in global memory.
// rows_max / cols_max bound the template sizes of the jobs (4*hr+3 / 4*wc+3).  Returns the number of kernels launched
// (> 0) or a negative cudaError.
size_t kcf_generic_scratch_bytes(int hr, int wc);
int kcf_launch_generic(int mode, int hr, int wc, const KcfLaunch &p, void *scratch, size_t scratch_bytes, cudaStream_t s);

}  // namespace mot
