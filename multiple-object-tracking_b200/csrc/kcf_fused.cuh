// kcf_fused.cuh -- the fused KCF/DCF kernels: persistent CTAs (one per SM) looping over (track, frame) jobs, everything
// between the BGR frame bytes and the updated model in ONE launch, so a track touches HBM once per predict and once per
// update.  Across jobs a CTA is software-pipelined: the next job's descriptor and crop rows are fetched during the spectral
// phases of the current one, and (predict) the response stage of a job runs on two warps underneath the next crop conversion.
//
// Replaces, for one track (reference paths relative to its root):
//   rgb2Gray + bilinearInterpolationGray      top/drawlib.c:192-240, 542-637   (called top/td.cpp:348-364)
//   FHoG::extract -> gradMag -> fhog          libhog/fhog.h:16-38, libhog/gradientMex.cpp:59-100, 148-317
//   kcf_get_features / kcf_fft2_features      trackers/kcf.cpp:245-267 (31 FFTW r2c plans)
//   predict: kcf_linear_correlation_zf + kcf_predict_ifft2 (+ the clamp of top/td.cpp:378-381)   kcf.cpp:306-362, 397-439
//   update : kcf_linear_correlation_kf + kcf_update_alpha + kcf_update_xf                         kcf.cpp:269-304, 364-395, 441-476
//
// Shared-memory plan for HR x WC cells (floats; 32x32 -> 223 KB, one CTA of 1024 threads per SM):
//   F   [31*NB]              gray patch -> (M/16, orientation bin) -> windowed features of all 31 channels -> their packed half spectra, in place
//   R1  [18*WC*(HR+1)]       SSE look-up tables (gradient phase) -> 18-bin cell histograms -> zf, response
//   MQ  [2 float2 / thread]  cp.async landing slots of the two model values of P5b -> Nyquist column of the spectra
//   N   [(WC+1)*(HR+1)]      block normalisers;  E [NB] cell energies;  wy[HR], wx[WC] Hann vectors
// The 31-channel feature tensor never leaves shared memory: one thread per cell generates its 31 windowed features from R1 and N
// (P4x), one thread per (channel, column) transforms them in place (P4y: real FFT of HR points as a complex FFT of HR/2) and stores
// them PACKED (DC.re, Nyquist.re share one complex slot), so the column pass is exactly HR/2 complex FFTs per channel: 31*16 = 496
// thread-sized transforms at 32x32.
//
// Arithmetic follows the reference operation by operation where its rounding is observable (gray: the double expression
// reproduced exactly in integers + three f32 operations, unfused
// f32 MUL/ADD in fHOG via __fmul_rn/__fadd_rn, SSE rsqrt/rcp through host-harvested tables, the histogram summed in the
// reference's pixel order); the FFTs and the spectral products are ordinary f32.
#pragma once
#include "mot_internal.h"
#include "fhog_common.cuh"
#include "fft_reg.cuh"
#include "copy_async.cuh"
#include <type_traits>

namespace mot {

template <int HR, int WC> struct Geo {
    static constexpr int NB = HR * WC;
    static constexpr int HK = HR / 2;                    // packed complex bins per column
    static constexpr int SK = HR / 2 + 1;                // FFTW half-spectrum length along rows
    static constexpr int S = WC * SK;
    static constexpr int H0 = 4 * HR, W0 = 4 * WC;
    static constexpr int RMAX = 4 * HR + 3, CMAX = 4 * WC + 3;
    static constexpr int GS = RMAX + 2;                  // gray column stride incl. a (now unused) 1-pixel border (odd -> conflict-free both ways)
    static constexpr int RS = HR + 1;                    // histogram / normaliser column stride
    // (M/16, bin) per pixel with a 2..5-pixel zero border, de-interleaved along y: [x+2][(y+2)&3][(y+2)>>2]
    static constexpr int PS = ((HR + 2 - 8 + 15) / 16) * 16 + 8;   // sub-column pitch: >= HR + 2 and = 8 (mod 16), so that a warp's P1 stores spread over the banks
    static constexpr int PC = 4 * PS + 2;                  // column pitch of the padded layout (= 2 mod 32: the zero-border stores spread over the banks)
    static constexpr int PADM = (W0 + 8) * PC;             // words of (M0 | bin): the bin rides in the 5 low mantissa bits, which are zero (fhog_tables.cpp)
    static constexpr int RAW_PITCH = 432;                  // bytes per staged frame row: 3*(4*32+3) + 2*15 alignment slack, 16-byte multiple
    static constexpr int RAW_FLOATS = RMAX * RAW_PITCH / 4;
    static constexpr int F_MIN = 31 * NB;
    static constexpr int F_A = (CMAX + 2) * GS > PADM ? (CMAX + 2) * GS : PADM;
    static constexpr int F_FLOATS = ((F_MIN > F_A ? F_MIN : F_A) + 31) & ~31;       // the regions behind it stay 128-byte aligned (2-D TMA destination)
    static constexpr int R1_MIN = 18 * WC * RS;
    static constexpr int N_FLOATS = (WC + 1) * (HR + 1);
    static constexpr int MQ_FLOATS = 4 * KCF_THREADS;      // [2][NT] float2: cp.async landing slots of the two model values of P5b; the first row then holds the Nyquist column [31][WC]
    static constexpr int ZB_FLOATS = 0;                    // (the raw transform of the packed (DC, Nyquist) column stays in its own spectrum slots)
    // Square grids: channel c is produced (P4b) and column-transformed (P5) by the same aligned group of WC threads, so the barrier
    // between the two phases is a warp-level one and the fast warps' model streaming overlaps the slow warps' feature generation.
    static constexpr bool FUSE45 = HR == WC;
    static constexpr int PIX_PER_THREAD = (H0 * W0 + KCF_THREADS - 1) / KCF_THREADS;
    static_assert(3 * CMAX + 30 <= RAW_PITCH, "staged row pitch");
    static_assert((F_FLOATS & 1) == 0, "float2 alignment of the R1 region");
};

// R1 region: cell histograms later, but first the SSE tables (lut_floats) and the staged frame rows (raw_floats)
__host__ __device__ inline int r1_region_floats(int r1_min, int lut_floats, int raw_floats)
{
    const int early = ((lut_floats + 31) & ~31) + raw_floats;
    const int v = r1_min > early ? r1_min : early;
    return (v + 3) & ~3;
}

template <int HR, int WC> size_t smem_bytes(int lut_floats)
{
    using G = Geo<HR, WC>;
    return sizeof(float) * (size_t)(G::F_FLOATS + r1_region_floats(G::R1_MIN, lut_floats, G::RAW_FLOATS) + G::MQ_FLOATS + G::ZB_FLOATS + G::N_FLOATS + G::NB + HR + WC + 64);
}

// Everything a job needs from the per-job / per-track arrays, gathered two jobs ahead by one lane of the staging warp
// so that no phase sits on a chain of dependent global loads (slot -> meta -> ..., frame index -> frame pointer).
struct JobDesc {
    mot_bbox_t box, pos;              // p.boxes[job]; meta->pos
    const uint8_t *frame;             // p.frame_ptr[p.frames[job]] (null with pre-cropped gray input)
    int slot, rows, cols, size_class, first_update, fslot;
    float scale_horiz, scale_vert;
    int box_idx;                      // where the job's box lives in p.boxes (job index, or p.box_index[job])
};

template <int HR, int WC, int MODE, bool DUMP>
__global__ void __launch_bounds__(KCF_THREADS, 1) kcf_fused_kernel(const KcfLaunch p, const int lut_floats)
{
    using G = Geo<HR, WC>;
    constexpr int NT = KCF_THREADS, NB = G::NB, HK = G::HK, SK = G::SK, S = G::S, GS = G::GS, RS = G::RS;
    constexpr int H0 = G::H0, W0 = G::W0;
    extern __shared__ __align__(128) float smem[];
    float *const F = smem;
    float *const R1 = F + G::F_FLOATS;
    float2 *const MQ = reinterpret_cast<float2 *>(R1 + r1_region_floats(G::R1_MIN, lut_floats, G::RAW_FLOATS));
    float *const Ns = reinterpret_cast<float *>(MQ) + G::MQ_FLOATS + G::ZB_FLOATS;
    float *const Es = Ns + G::N_FLOATS;
    float *const wy_s = Es + NB;
    float *const wx_s = wy_s + HR;
    float *const red = wx_s + WC;

    __shared__ __align__(8) uint64_t mbar, mbar2;     // SSE tables / staged frame rows
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&mbar, 1); mbar_init(&mbar2, 1); }
    __syncthreads();
    const bool lut_smem = lut_floats > 0;
    const int n_rs = 2 * (2 << p.tab.rsqrt_bits), n_bn = (2 * p.tab.bin_nseg + 3) & ~3;      // floats of the fused {rsqrt, rcp} table; bin entries
    unsigned char *const raw = reinterpret_cast<unsigned char *>(R1 + ((lut_floats + 31) & ~31));
    const int Wm = p.frame_w - 1, Hm = p.frame_h - 1;

    // Fetch the frame rows of job `jb`'s crop into the staging area with the bulk-copy engine (cp.async.bulk on mbar2).
    // Rows start on arbitrary byte offsets, so each row is fetched as the enclosing 16-byte-aligned span; that needs a
    // 16-byte aligned frame base and stride (true for every common frame width).  Whether the crop has the template size
    // (the only case that uses the staged rows) is known only when the metadata arrives: a crop that fits the staging area
    // is fetched speculatively.  Executed by the last warp; exactly one arrival on mbar2 per job, with or without bytes.
    __shared__ __align__(8) JobDesc s_desc[4];        // ring: job iteration it lives in s_desc[it & 3]
    __shared__ KcfClassDev s_cls;                     // constants of the current window class (reloaded when the class changes)
    __shared__ int s_cls_id;
    constexpr int ROI_WARP = NT / 32 - 1;             // the last warp has the lightest P5 (bin-0 tasks)
    auto fetch_desc = [&](int jb, JobDesc *d) {
        const int sl = p.slots[jb];
        const KcfMeta *m = p.meta + sl;
        const int bi = p.box_index ? p.box_index[jb] : jb;
        d->box = p.boxes[bi]; d->box_idx = bi;
        d->fslot = (p.gray == nullptr) ? p.frames[jb] : 0;
        d->frame = (p.gray == nullptr) ? p.frame_ptr[d->fslot] : nullptr;
        d->slot = sl; d->rows = m->rows; d->cols = m->cols; d->size_class = m->size_class; d->first_update = m->first_update;
        d->pos = m->pos; d->scale_horiz = m->scale_horiz; d->scale_vert = m->scale_vert;
    };
    // A crop that lies inside the frame vertically is fetched by ONE 2-D TMA load through the frame's tensor map (box = RAW_PITCH
    // bytes x RMAX rows, rows past the crop are simply not used, columns past the frame row are zero-filled and never read); crops
    // that reach over the top / bottom edge need the edge row replicated and take the row-by-row path.
    const bool tma_ok = p.frame_tmaps != nullptr && (smem_u32(raw) & 127u) == 0;
    constexpr int TMAP_ROWCLASS = HR == 8 ? 0 : HR == 16 ? 1 : 2;
    auto issue_roi = [&](const mot_bbox_t bx, const uint8_t *frame_of_job, int fslot) {
        const int lane = tid & 31;
        int l = bx.l, t = bx.t, r = bx.r, b = bx.b;
        if (t > b) { const int q = t; t = b; b = q; }
        if (l > r) { const int q = l; l = r; r = q; }
        const int rows_s = b - t + 1, cols_s = r - l + 1;
        const uint8_t *frame = frame_of_job;
        const bool fetch = (p.gray == nullptr) && rows_s <= G::RMAX && cols_s <= G::CMAX && (((uintptr_t)frame | (uintptr_t)p.frame_stride) & 15) == 0;
        const int x_lo = clampi(l, 0, Wm), x_hi = clampi(l + cols_s - 1, 0, Wm);
        const int a0 = (x_lo * 3) & ~15, a1 = ((x_hi + 1) * 3 + 15) & ~15;
        const bool tma = fetch && tma_ok && t >= 0 && t + rows_s - 1 <= Hm;
        if (lane == 0) {
            mbar_expect_tx(&mbar2, tma ? (uint32_t)(G::RMAX * G::RAW_PITCH) : fetch ? (uint32_t)rows_s * (uint32_t)(a1 - a0) : 0u);
            if (tma) tma_load_2d(raw, reinterpret_cast<const char *>(p.frame_tmaps) + ((long)fslot * 3 + TMAP_ROWCLASS) * 128, a0 >> 2, t, &mbar2);
        }
        __syncwarp();
        if (fetch && !tma)
            for (int y = lane; y < rows_s; y += 32)
                bulk_g2s(raw + y * G::RAW_PITCH, frame + (long)clampi(t + y, 0, Hm) * p.frame_stride + a0, a1 - a0, &mbar2);
    };

    // Persistent CTA: jobs blockIdx.x, blockIdx.x + gridDim.x, ...; the crop of the NEXT job streams into shared memory
    // while the spectral phases of the current one run (the staging area is free from P5 on).
    // ------------------------------------------------------------------ P7: response = c2r(zf), first-max argmax, box shift
    // Run by the first two warps only (it is a 34-thread and a 32-thread transform plus a reduction), synchronised with a
    // named barrier, so that the other thirty warps can already convert the next job's crop (see P0): zf_s, resp and red
    // are not touched by anything else until the next job's tables are staged, which these two warps do themselves.
    float2 *const zf_s = reinterpret_cast<float2 *>(R1) + KCF_CHAN * WC;      // [WC][SK], behind the Nyquist column FN
    float *const resp = reinterpret_cast<float *>(zf_s + S);                  // [WC][HR]
    auto p7_tail = [&](const JobDesc &pj, int pjob) {
        {
        // inverse complex FFT along the WC columns of every half-spectrum bin k, one lane pair per bin
        constexpr int HW = WC / 2;
        const int kk = tid >> 1, half = tid & 1;
        const bool active = kk < SK;
        const unsigned msk = __ballot_sync(0xFFFFFFFFu, active);
        if (active) {
            float2 a[HW];
            fft_pair_ld<WC, +1>(a, half, msk, [&](int j) { return zf_s[j * SK + kk]; });
#pragma unroll
            for (int m = 0; m < HW; ++m) zf_s[(2 * m + half) * SK + kk] = a[brev<HW>(m)];
        }
    }
    bar_sync_named(1, 64);
    if (tid < WC) {
        // c2r along the HR rows of column j = tid: Hermitian half spectrum -> HR reals through one complex FFT of HR/2
        const float2 *Y = zf_s + tid * SK;
        float2 q[HK];
        q[0] = make_float2(Y[0].x + Y[HK].x, Y[0].x - Y[HK].x);      // c2r ignores Im(DC), Im(Nyquist)
#pragma unroll
        for (int k = 1; k < HK; ++k) {
            const float2 A = Y[k], B = Y[HK - k];
            const int ti = (k * (64 / HR)) & 63;
            const float cs = tw::C64[ti], sn = tw::S64[ti];
            const float sx_ = A.x + B.x, sy_ = A.y - B.y, dx_ = A.x - B.x, dy_ = A.y + B.y;
            // Q[k] = (A + conj(B)) + i exp(+2 pi i k / HR) (A - conj(B))
            q[k] = make_float2(sx_ - dx_ * sn - dy_ * cs, sy_ + dx_ * cs - dy_ * sn);
        }
        fft_dif<HK, +1>(q);
#pragma unroll
        for (int m = 0; m < HK; ++m) { const float2 v = q[brev<HK>(m)]; resp[tid * HR + 2 * m] = v.x; resp[tid * HR + 2 * m + 1] = v.y; }
    }
    bar_sync_named(1, 64);
    if (DUMP && p.dump.resp) {
        float *d = p.dump.resp + (long)pjob * p.dump.stride_cell;
        for (int i = tid; i < NB; i += 64) d[i] = resp[i];
    }
    // first maximum in memory order (j outer, i inner), strict '>' from -99999 (kcf.cpp:402-417)
    float best = -99999.0f; int besti = 0x7FFFFFFF;
    for (int i = tid; i < NB; i += 64) { const float v = resp[i]; if (v > best) { best = v; besti = i; } }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const float ov = __shfl_down_sync(0xFFFFFFFFu, best, off);
        const int oi = __shfl_down_sync(0xFFFFFFFFu, besti, off);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
    }
    int *const redi = reinterpret_cast<int *>(red) + 32;
    if ((tid & 31) == 0) { red[tid >> 5] = best; redi[tid >> 5] = besti; }
    bar_sync_named(1, 64);
    if (tid == 0) {
        for (int w = 1; w < 2; ++w) { const float ov = red[w]; const int oi = redi[w]; if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; } }
        int vd = 1, hd = 1;                                            // reference leaves these uninitialised when nothing beats -99999
        if (besti != 0x7FFFFFFF) { hd = besti / HR + 1; vd = besti - (hd - 1) * HR + 1; }
        if (DUMP && p.dump.peak) { p.dump.peak[2 * pjob] = vd; p.dump.peak[2 * pjob + 1] = hd; }
        if (vd > HR / 2) vd -= HR;                                     // kcf.cpp:419-420
        if (hd > WC / 2) hd -= WC;
        mot_bbox_t pos = pj.pos;
        const float dv = __fmul_rn((float)(KCF_CELL * (vd - 1)), pj.scale_vert);
        const float dh = __fmul_rn((float)(KCF_CELL * (hd - 1)), pj.scale_horiz);
        pos.t = __float2int_rz(__fadd_rn((float)pos.t, dv));           // kcf.cpp:423-426 (float math, truncation)
        pos.b = __float2int_rz(__fadd_rn((float)pos.b, dv));
        pos.l = __float2int_rz(__fadd_rn((float)pos.l, dh));
        pos.r = __float2int_rz(__fadd_rn((float)pos.r, dh));
        (p.meta + pj.slot)->pos = pos;
        if (p.clamp_to_frame) {                                        // top/td.cpp:378-381
            pos.l = clampi(pos.l, 0, p.frame_w - 1); pos.r = clampi(pos.r, 0, p.frame_w - 1);
            pos.t = clampi(pos.t, 0, p.frame_h - 1); pos.b = clampi(pos.b, 0, p.frame_h - 1);
        }
        p.boxes[pj.box_idx] = pos;
    }
    };
    bool p7_pending = false;
    uint32_t phase = 0;
    // the job count may live on the device (job lists built by a previous kernel, csrc/td_device.cu); p.n_jobs then bounds it
    const int n_jobs = p.n_jobs_dev ? min(*p.n_jobs_dev, p.n_jobs) : p.n_jobs;
    if (tid == 0) s_cls_id = -1;
    if ((tid >> 5) == ROI_WARP && (int)blockIdx.x < n_jobs) {
        if ((tid & 31) == 0) {
            fetch_desc(blockIdx.x, &s_desc[0]);
            if ((int)(blockIdx.x + gridDim.x) < n_jobs) fetch_desc(blockIdx.x + gridDim.x, &s_desc[1]);
        }
        __syncwarp();
        issue_roi(s_desc[0].box, s_desc[0].frame, s_desc[0].fslot);
    }
    int it = 0;
    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x, phase ^= 1u, ++it) {
    __syncthreads();                                   // the previous job is done with every shared-memory region
    const JobDesc &jd = s_desc[it & 3];
    // ------------------------------------------------------------------ P0: tables, Hann vectors, ROI -> gray
    // SSE tables -> shared memory (one arrival on mbar per job); issued by threads 0 and 1 once P7 of the previous job, which
    // still works in that area, is done
    auto stage_tables = [&]() {
        if (tid == 0) { mbar_expect_tx(&mbar, lut_smem ? (uint32_t)(n_rs + n_bn) * 4u : 0u); if (lut_smem) bulk_g2s(R1, p.tab.rsrc_tab, n_rs * 4, &mbar); }
        if (tid == 1 && lut_smem) bulk_g2s(R1 + n_rs, p.tab.bin2_tab, n_bn * 4, &mbar);
    };

    const int slot = jd.slot;
    mot_bbox_t box = jd.box;
    const uint8_t *frame = jd.frame;
    int l = box.l, t = box.t, r = box.r, b = box.b;
    if (t > b) { const int q = t; t = b; b = q; }              // top/drawlib.c:203-215
    if (l > r) { const int q = l; l = r; r = q; }
    const int rows_s = b - t + 1, cols_s = r - l + 1;
    const bool fetch = (p.gray == nullptr) && rows_s <= G::RMAX && cols_s <= G::CMAX && (((uintptr_t)frame | (uintptr_t)p.frame_stride) & 15) == 0;
    const int x_lo = clampi(l, 0, Wm), x_hi = clampi(l + cols_s - 1, 0, Wm);
    const int a0 = (x_lo * 3) & ~15;                           // aligned start of the staged byte span of a frame row

    const int rows = jd.rows, cols = jd.cols;
    const bool first_update = jd.first_update != 0;
    if (jd.size_class != s_cls_id) {                   // block-uniform; in a launch of one class: the first job only
        const int sc = jd.size_class;
        __syncthreads();
        if (tid == 0) { s_cls = p.classes[sc]; s_cls_id = sc; }
        __syncthreads();
        if (tid >= 32 && tid < 32 + HR) wy_s[tid - 32] = 0.5f * s_cls.wy[tid - 32];   // halved: carries the x0.5 of hogChannels (exact scaling)
        if (tid >= 64 && tid < 64 + WC) wx_s[tid - 64] = s_cls.wx[tid - 64];
    }
    const KcfClassDev &cls = s_cls;
    const bool identity = (p.gray == nullptr) && rows_s == rows && cols_s == cols;
    const bool staged = fetch && identity;

    // The model (31*S complex) is not needed before P5, tens of microseconds from now: pull it into L2 now so that P5
    // sees L2 latency instead of HBM latency (there is no shared memory left to stage it in).
    if (MODE == KCF_MODE_PREDICT || !first_update) {
        const char *mp = reinterpret_cast<const char *>(p.model + (long)slot * p.model_stride);
        if (tid == 96) prefetch_l2_bulk(mp, KCF_CHAN * S * 8);
    }
    if (tid == 97) prefetch_l2_bulk(p.alpha + (long)slot * p.alpha_stride, (S * 4 + 15) & ~15);

    // P7 of the previous job (predict only): underneath the crop conversion when the staged fast path is taken -- warps 0-1
    // finish the previous job while warps 2-31 convert -- otherwise first, by itself
    const bool p7_now = (MODE == KCF_MODE_PREDICT) && p7_pending;
    const bool overlap = p7_now && staged;
    const JobDesc &pjd = s_desc[(it - 1) & 3];
    if (p7_now && !overlap) {
        if (tid < 64) p7_tail(pjd, job - (int)gridDim.x);
        __syncthreads();
    }
    p7_pending = false;
    if (!overlap) stage_tables();
    if (p.gray != nullptr) {
        const float *src = p.gray + (long)job * p.gray_stride;
        for (int idx = tid; idx < rows * cols; idx += NT) { const int x = idx / rows, y = idx - x * rows; F[(x + 1) * GS + y + 1] = src[idx]; }
    } else if (staged && overlap && tid < 64) {
        p7_tail(pjd, job - (int)gridDim.x);
        stage_tables();
    } else if (staged) {
        // one warp polls for the staged rows, the others sleep at the barrier (all warps, or warps 2-31 next to a running P7)
        const int w0 = overlap ? 2 : 0, nw = NT / 32 - w0;
        if ((tid >> 5) == w0) mbar_wait(&mbar2, phase);
        if (overlap) bar_sync_named(2, NT - 64); else __syncthreads();
        // staged bytes -> gray: one warp per row, lanes along x (3-byte pixels: conflict-free shared loads)
        const int warp = (tid >> 5) - w0, lane = tid & 31;
        if (l >= 0 && l + cols - 1 <= Wm) {
            // crop inside the frame horizontally (the common case): no per-pixel clamping
            for (int y = warp; y < rows; y += nw) {
                const unsigned char *rrow = raw + y * G::RAW_PITCH + (l * 3 - a0);
                // WC full 32-pixel chunks (cols >= 4 * WC) with compile-time offsets, then the 0..3 leftover pixels
                const unsigned char *px = rrow + lane * 3;
                float *dst = F + (lane + 1) * GS + y + 1;
#pragma unroll
                for (int c = 0; c < W0 / 32; ++c) dst[c * 32 * GS] = bgr_gray(px + c * 96);
                if (lane + W0 < cols) dst[W0 * GS] = bgr_gray(px + W0 * 3);
            }
        } else {
            for (int y = warp; y < rows; y += nw) {
                const unsigned char *rrow = raw + y * G::RAW_PITCH - a0;
#pragma unroll 4
                for (int x = lane; x < cols; x += 32) F[(x + 1) * GS + y + 1] = bgr_gray(rrow + clampi(l + x, 0, Wm) * 3);
            }
        }
    } else if (identity) {
        // unaligned frames: plain loads.  Equal sizes: bilinearInterpolationGray is an exact copy (top/drawlib.c:610-633, xs = ys = 1)
        const int warp = tid >> 5, lane = tid & 31;
        for (int y = warp; y < rows; y += NT / 32) {
            const uint8_t *r0 = frame + (long)clampi(t + y, 0, Hm) * p.frame_stride;
#pragma unroll 4
            for (int x = lane; x < cols; x += 32) F[(x + 1) * GS + y + 1] = bgr_gray(r0 + clampi(l + x, 0, Wm) * 3);
        }
    } else {
        // the reference resamples a column-major crop as if it were row-major height x width; reproduced through
        // linear indices (top/drawlib.c:542-637, called as (dst, src, rows_s, cols_s, rows_d, cols_d), top/td.cpp:357-364)
        const float xs = __fdiv_rn((float)cols_s, (float)cols), ys = __fdiv_rn((float)rows_s, (float)rows);
        for (int k = tid; k < rows * cols; k += NT) {
            const int yy = k / cols, xx = k - yy * cols;
            const float sx = __fmul_rn((float)xx, xs), sy = __fmul_rn((float)yy, ys);
            const int x0 = __float2int_rz(sx), y0 = __float2int_rz(sy);
            const float fx = __fsub_rn(sx, (float)x0), fy = __fsub_rn(sy, (float)y0);
            const float ifx = __fsub_rn(1.0f, fx), ify = __fsub_rn(1.0f, fy);
            const int x1 = (x0 + 1 >= cols_s) ? x0 : x0 + 1, y1 = (y0 + 1 >= rows_s) ? y0 : y0 + 1;
            float c[4];
            const int sidx[4] = { y0 * cols_s + x0, y0 * cols_s + x1, y1 * cols_s + x0, y1 * cols_s + x1 };
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int sc = sidx[q] / rows_s, sr = sidx[q] - sc * rows_s;   // column-major crop element
                c[q] = bgr_gray(frame + (long)clampi(t + sr, 0, Hm) * p.frame_stride + clampi(l + sc, 0, Wm) * 3);
            }
            const float l0 = __fadd_rn(__fmul_rn(ifx, c[0]), __fmul_rn(fx, c[1]));
            const float l1 = __fadd_rn(__fmul_rn(ifx, c[2]), __fmul_rn(fx, c[3]));
            const float o = __fadd_rn(__fmul_rn(ify, l0), __fmul_rn(fy, l1));
            const int dc = k / rows, dr = k - dc * rows;                         // column-major template element
            F[(dc + 1) * GS + dr + 1] = o;
        }
    }
    if (tid < 32) {
        if (!staged) mbar_wait(&mbar2, phase);                 // (speculative) rows must have landed before the region is reused
        mbar_wait(&mbar, phase);
    }
    __syncthreads();
    if (DUMP && p.dump.gray) {
        float *d = p.dump.gray + (long)job * p.dump.stride_px;
        for (int idx = tid; idx < rows * cols; idx += NT) { const int x = idx / rows, y = idx - x * rows; d[idx] = F[(x + 1) * GS + y + 1]; }
    }

    // ------------------------------------------------------------------ P1: gradient magnitude + orientation bin
    // libhog/gradientMex.cpp:15-37 (grad1), :59-100 (gradMag, d=1, full=true), :112-145 (gradQuantize, nearest bin)
    uint32_t mbr[G::PIX_PER_THREAD];                   // M/16 with the bin in its five low (zero) mantissa bits
    {
        const LutConsts lk = make_lut_consts(p.tab);
        // thread -> pixels (x0 + q * XS, y): y is fixed per thread and the addresses of the 16 pixels differ by compile-time offsets
        static_assert(NT % H0 == 0 && (H0 * W0) % NT == 0, "P1 pixel mapping");
        constexpr int XS = NT / H0;
        const int y = tid & (H0 - 1), x0 = tid / H0;
        const float ry = (y == 0 || y == rows - 1) ? 1.f : .5f;
        const float *const g0 = F + (x0 + 1) * GS + y + 1;
        // grad1's one-sided differences on the border (factor 1) are central differences against the pixel itself: the neighbour
        // offsets of a border pixel collapse to 0.  y is fixed per thread, so the vertical pair is two per-thread pointers.
        const float *const gup = g0 - (y == 0 ? 0 : 1), *const gdn = g0 + (y == rows - 1 ? 0 : 1);
        auto p1_pixels = [&](const float2 *rsrc, const uint32_t *bn) {
#pragma unroll
            for (int q = 0; q < G::PIX_PER_THREAD; ++q) {
                const int x = x0 + q * XS;
                const float *g = g0 + q * XS * GS;
                // grad1: one-sided difference (x1) on the border, central difference (x0.5) inside
                // only the first and the last pixel of a thread can sit on the left / right border (cols >= W0): the selects of
                // the others fold away
                const float rx = (q == 0 || q == G::PIX_PER_THREAD - 1) ? ((x == 0 || x == cols - 1) ? 1.f : .5f) : .5f;
                const int oL = (q == 0 && x == 0) ? 0 : -GS, oR = (q == G::PIX_PER_THREAD - 1 && x == cols - 1) ? 0 : GS;
                const float gx = __fmul_rn(__fsub_rn(g[oR], g[oL]), rx);
                const float gy = __fmul_rn(__fsub_rn(gdn[q * XS * GS], gup[q * XS * GS]), ry);
                mbr[q] = grad_pixel_k(gx, gy, rsrc, bn, lk);
            }
        };
        if (lut_smem) p1_pixels(reinterpret_cast<const float2 *>(R1), reinterpret_cast<const uint32_t *>(R1) + n_rs);   // shared-memory tables (LDS)
        else p1_pixels(p.tab.rsrc_tab, p.tab.bin2_tab);                                                               // oversized tables stay in global memory
    }
    __syncthreads();
    // (M0, bin) overwrite the gray patch in a zero-bordered layout de-interleaved along y, [x+2][(y+2)&3][(y+2)>>2], so
    // that the cell-parallel gather below reads consecutive words and needs no bounds checks (a zero magnitude adds +0)
    constexpr int PC = G::PC;
    uint32_t *const MB = reinterpret_cast<uint32_t *>(F);
    {
        constexpr int XS = NT / H0;
        const int y = tid & (H0 - 1), x0 = tid / H0;
        uint32_t *const mb0 = MB + (x0 + 2) * PC + ((y + 2) & 3) * G::PS + ((y + 2) >> 2);
#pragma unroll
        for (int q = 0; q < G::PIX_PER_THREAD; ++q) {
            mb0[q * XS * PC] = mbr[q];
            if (DUMP && p.dump.m0) {
                const int idx = (x0 + q * XS) * H0 + y;
                p.dump.m0[(long)job * p.dump.stride_px + idx] = __uint_as_float(mbr[q] & ~31u); p.dump.bin[(long)job * p.dump.stride_px + idx] = (unsigned char)(mbr[q] & 31u);
            }
        }
    }
    // zero border: 8 full columns (x+2 in {0,1,W0+2..W0+7}) and 8 rows of the interior columns (y+2 in {0,1,H0+2..H0+7})
    for (int k = tid; k < 8 * (H0 + 8) + 8 * W0; k += NT) {
        int xs, ys;
        if (k < 8 * (H0 + 8)) { const int cxx = k / (H0 + 8); ys = k - cxx * (H0 + 8); xs = cxx < 2 ? cxx : W0 + cxx; }
        else { const int k2 = k - 8 * (H0 + 8); const int ry = k2 & 7; xs = 2 + (k2 >> 3); ys = ry < 2 ? ry : H0 + ry; }
        const int a = xs * PC + (ys & 3) * G::PS + (ys >> 2);
        MB[a] = 0u;
    }
    __syncthreads();

    // ------------------------------------------------------------------ P2: 18-bin cell histograms by ordered gather
    // libhog/gradientMex.cpp:183-221 (bilinear spatial interpolation, softBin<0) + :225-230 (boundary x 8/7).
    // One thread owns a cell (two, interleaved, at 32x32) and adds its 8x8 contributing pixels in the reference's
    // (x outer, y inner) order, so the sums are deterministic and equal to the reference's sequential scatter.
    constexpr int CPT = (NB + NT - 1) / NT;                    // cells per thread
    {
        int ccx[CPT], ccy[CPT]; bool live[CPT]; float *h[CPT]; int base[CPT];
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            const int cell = tid + u * NT;
            live[u] = cell < NB;
            const int cc = live[u] ? cell : 0;
            ccx[u] = cc / HR; ccy[u] = cc - ccx[u] * HR;
            h[u] = R1 + ccx[u] * RS + ccy[u];
            base[u] = (4 * ccx[u]) * PC + ccy[u];
            if (live[u]) {
#pragma unroll
                for (int o = 0; o < 18; ++o) h[u][o * (WC * RS)] = 0.f;
            }
        }
#pragma unroll
        for (int dx = 0; dx < 8; ++dx) {                                                     // fully unrolled: weights and offsets are immediates
            const float wxv = 0.125f + 0.25f * (float)(dx < 4 ? dx : 7 - dx);
#pragma unroll
            for (int dy = 0; dy < 8; ++dy) {
                const float w = wxv * (0.125f + 0.25f * (float)(dy < 4 ? dy : 7 - dy));     // dyadic weights: exact product
#pragma unroll
                for (int u = 0; u < CPT; ++u) {
                    if (!live[u]) continue;
                    const int a = base[u] + dx * PC + (dy & 3) * G::PS + (dy >> 2);
                    const uint32_t mb = MB[a];
                    const float v = __fmul_rn(w, __uint_as_float(mb & ~31u));
                    float *const hb = h[u] + (int)(mb & 31u) * (WC * RS);
                    *hb = __fadd_rn(*hb, v);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            if (!live[u]) continue;
            const int cx = ccx[u], cy = ccy[u];
            // boundary cells x 8/7 per touching side (gradientMex.cpp:226-229); a cell touches at most one side per axis, and
            // multiplying by 1.0f is the identity, so two unconditional products reproduce the four conditional ones
            const float sx = (cx == 0 || cx == WC - 1) ? 8.f / 7.f : 1.f, sy = (cy == 0 || cy == HR - 1) ? 8.f / 7.f : 1.f;
            float e = 0.f;
            float r[18];
#pragma unroll
            for (int o = 0; o < 18; ++o) {
                float v = h[u][o * (WC * RS)];
                v = __fmul_rn(__fmul_rn(v, sx), sy);
                h[u][o * (WC * RS)] = v; r[o] = v;
            }
            // cell energy over the 9 contrast-insensitive bins, R2 = R1[o] + R1[o+9] (gradientMex.cpp:308-309, :242-243)
#pragma unroll
            for (int o = 0; o < 9; ++o) { const float r2 = __fadd_rn(r[o], r[o + 9]); e = __fadd_rn(e, __fmul_rn(r2, r2)); }
            Es[cx * HR + cy] = e;
        }
    }
    __syncthreads();
    if (DUMP && p.dump.r1) {
        float *d = p.dump.r1 + (long)job * p.dump.stride_cell * 18;
        for (int i = tid; i < 18 * NB; i += NT) { const int o = i / NB, c2 = i - o * NB, x = c2 / HR, y = c2 - x * HR; d[i] = R1[o * (WC * RS) + x * RS + y]; }
    }

    // ------------------------------------------------------------------ P3: 2x2 block normalisers (hogNormMatrix, :236-253)
    // (WC-1) x (HR-1) distinct block sums; the border rows / columns of the (WC+1) x (HR+1) matrix replicate them, so each
    // thread computes one value and stores it to its 1, 2 or 4 positions (one pass instead of 1 + 1/16)
    for (int i = tid; i < (WC - 1) * (HR - 1); i += NT) {
        const int x = i / (HR - 1), y = i - x * (HR - 1);
        const float eps = 1e-4f / 4 / 4 / 4 / 4 / 4;
        float e = __fadd_rn(Es[x * HR + y], Es[x * HR + y + 1]);
        e = __fadd_rn(e, Es[(x + 1) * HR + y]);
        e = __fadd_rn(e, Es[(x + 1) * HR + y + 1]);
        e = __fadd_rn(e, eps);
        const float nv = __fdiv_rn(1.0f, __fsqrt_rn(e));
        float *const q = Ns + (x + 1) * (HR + 1) + y + 1;
        const bool xl = x == 0, xh = x == WC - 2, yl = y == 0, yh = y == HR - 2;
        q[0] = nv;
        if (yl) q[-1] = nv;
        if (yh) q[1] = nv;
        if (xl) { q[-(HR + 1)] = nv; if (yl) q[-(HR + 1) - 1] = nv; if (yh) q[-(HR + 1) + 1] = nv; }
        if (xh) { q[(HR + 1)] = nv; if (yl) q[(HR + 1) - 1] = nv; if (yh) q[(HR + 1) + 1] = nv; }
    }
    __syncthreads();
    if (DUMP && p.dump.nrm) {
        float *d = p.dump.nrm + (long)job * G::N_FLOATS;
        for (int i = tid; i < G::N_FLOATS; i += NT) d[i] = Ns[i];
    }

    // ------------------------------------------------------------------ P4x: all 31 windowed features of a cell, one thread per cell
    // hogChannels (gradientMex.cpp:256-280): the clipped products min(R1[o] * N_blk, 0.2) of the 18 orientations feed BOTH the
    // contrast-sensitive channels (type 1, :266-270: the four are added in block order; the x0.5 commutes exactly with the additions
    // and sits in the stored window rows) and the four texture channels (type 2, :271-275: x0.2357, accumulated in orientation
    // order), so a cell computes them once; the contrast-insensitive channels take R2 = R1[o] + R1[o+9] (:308-309).  Every feature
    // is multiplied by cos_win (kcf.cpp:251-258) here and parked as a plain float in the spectrum slot of its (channel, column),
    // rotated by the column index so that the column-wise reader below is bank-conflict free.
    static_assert(NB <= NT, "P4x: one cell per thread");
    if (tid < NB) {
        const int j = tid / HR, i = tid - j * HR;
        const float *const n0 = Ns + j * (HR + 1) + i, *const n1 = n0 + (HR + 1);
        const float nv0 = n1[1], nv1 = n1[0], nv2 = n0[1], nv3 = n0[0];      // GETT(0), GETT(1), GETT(hb1), GETT(hb1+1)
        const float w = __fmul_rn(wy_s[i], wx_s[j]);                         // (0.5 wy) * wx
        const float *const rp = R1 + j * RS + i;
        float *const fp = F + j * HR + ((i + j) & (HR - 1));                 // channel c: fp[c * NB]
        float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;
        float rlo[9];
#pragma unroll
        for (int o = 0; o < 18; ++o) {
            const float rv = rp[o * (WC * RS)];
            const float t0 = fminf(__fmul_rn(rv, nv0), 0.2f), t1 = fminf(__fmul_rn(rv, nv1), 0.2f);
            const float t2 = fminf(__fmul_rn(rv, nv2), 0.2f), t3 = fminf(__fmul_rn(rv, nv3), 0.2f);
            fp[o * NB] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(t0, t1), t2), t3), w);
            h0 = __fadd_rn(h0, __fmul_rn(t0, .2357f)); h1 = __fadd_rn(h1, __fmul_rn(t1, .2357f));
            h2 = __fadd_rn(h2, __fmul_rn(t2, .2357f)); h3 = __fadd_rn(h3, __fmul_rn(t3, .2357f));
            if (o < 9) rlo[o] = rv;
            else {
                const float r2 = __fadd_rn(rlo[o - 9], rv);
                const float u0 = fminf(__fmul_rn(r2, nv0), 0.2f), u1 = fminf(__fmul_rn(r2, nv1), 0.2f);
                const float u2 = fminf(__fmul_rn(r2, nv2), 0.2f), u3 = fminf(__fmul_rn(r2, nv3), 0.2f);
                fp[(9 + o) * NB] = __fmul_rn(__fadd_rn(__fadd_rn(__fadd_rn(u0, u1), u2), u3), w);
            }
        }
        // doubled (exact) because the window rows are stored halved
        fp[27 * NB] = __fmul_rn(h0 + h0, w); fp[28 * NB] = __fmul_rn(h1 + h1, w);
        fp[29 * NB] = __fmul_rn(h2 + h2, w); fp[30 * NB] = __fmul_rn(h3 + h3, w);
    }
    __syncthreads();
    // The histograms are consumed: the staging area is free again -> start streaming the next job's crop underneath P4y..P7
    if ((tid >> 5) == ROI_WARP && job + (int)gridDim.x < n_jobs) {
        const JobDesc &nd = s_desc[(it + 1) & 3];
        issue_roi(nd.box, nd.frame, nd.fslot);
        if ((tid & 31) == 0 && job + 2 * (int)gridDim.x < n_jobs) fetch_desc(job + 2 * gridDim.x, &s_desc[(it + 2) & 3]);
    }

    // ------------------------------------------------------------------ P4y: real FFT along the rows of every (channel, column)
    // r2c along the HR rows as ONE complex FFT of HR/2 points, in place; packed store (DC and Nyquist, both real, share slot 0).
    float2 *const F2 = reinterpret_cast<float2 *>(F);
    for (int task = tid; task < KCF_CHAN * WC; task += NT) {
        const int c = task / WC, j = task - c * WC;
        float2 z[HK];
        {
            const float *const tp = F + (c * WC + j) * HR;
#pragma unroll
            for (int i = 0; i < HR; ++i) {
                const float f = tp[(i + j) & (HR - 1)];
                if (i & 1) z[i >> 1].y = f; else z[i >> 1].x = f;
            }
        }
        if (DUMP && p.dump.feat) {
#pragma unroll
            for (int i = 0; i < HR; ++i) p.dump.feat[(long)job * p.dump.stride_cell * KCF_CHAN + c * NB + j * HR + i] = (i & 1) ? z[i >> 1].y : z[i >> 1].x;
        }
        fft_dif<HK, -1>(z);
        const int sw = fpos<HK, WC>(0, j);
        float2 *const dst = F2 + (c * WC + j) * HK;
        {
            const float2 z0 = z[0];
            dst[0 ^ sw] = make_float2(z0.x + z0.y, z0.x - z0.y);      // (DC, Nyquist), both real
        }
#pragma unroll
        for (int k = 1; k <= HK / 2; ++k) {
            const float2 A = z[brev<HK>(k)], Bc = z[brev<HK>(HK - k)];
            const int ti = (k * (64 / HR)) & 63;
            const float cs = tw::C64[ti], sn = tw::S64[ti];
            // X[k] = 1/2 [ (A + conj(B)) - i w^k (A - conj(B)) ],  w = exp(-2 pi i / HR)
            {
                const float sx_ = A.x + Bc.x, sy_ = A.y - Bc.y, dx_ = A.x - Bc.x, dy_ = A.y + Bc.y;
                dst[k ^ sw] = make_float2(0.5f * (sx_ + dy_ * cs - dx_ * sn), 0.5f * (sy_ - dx_ * cs - dy_ * sn));
            }
            if (k != HK - k) {
                // same formula for k' = HK - k: A' = B, B' = A, cos' = -cos, sin' = sin
                const float sx_ = Bc.x + A.x, sy_ = Bc.y - A.y, dx_ = Bc.x - A.x, dy_ = Bc.y + A.y;
                dst[(HK - k) ^ sw] = make_float2(0.5f * (sx_ - dy_ * cs - dx_ * sn), 0.5f * (sy_ + dx_ * cs - dy_ * sn));
            }
        }
    }
    // square grids: the rows of channel c were written by the WC threads that transform its columns next -- a warp-level barrier
    if (G::FUSE45) __syncwarp(); else __syncthreads();

    // ------------------------------------------------------------------ P5: complex FFT along the WC columns + spectral work
    // Tasks (channel c, packed bin k), each run by a PAIR of adjacent lanes that hold half of the WC points each (the
    // first radix-2 stage goes through warp shuffles), so the transform fits the 64-register budget of a 1024-thread CTA.
    // Lane `half` of a pair ends up with the output rows j' = 2m + half.  Bins k >= 1 are ordinary columns; bin 0 carries
    // two real-input columns (DC and Nyquist): its raw transform is parked in ZB and P5b, one row per thread on all threads,
    // separates and finishes both columns (the bin-0 lanes only park 16 values while the other lanes of their warp work).
    float2 *const FN = MQ;                                             // Nyquist column [31][WC]: row e replaces the model value thread e fetched into its own first slot
    // (job-level values are re-read from the descriptor where they are needed instead of living in registers across the phases)
    float2 *const model = p.model + (long)jd.slot * p.model_stride;
    const bool first = (MODE == KCF_MODE_UPDATE) && jd.first_update != 0;
    const float fac = first ? 1.0f : p.factor;                         // kcf.cpp:443
    const float omf = __fsub_rn(1.0f, fac);
    const bool need_model = (MODE == KCF_MODE_PREDICT) || !first;
    // the two model values of P5b (DC and Nyquist bin of this thread's row) start their way into the thread's own slots now, through
    // cp.async: they arrive underneath the column pass, without holding registers across it
    if (tid < KCF_CHAN * WC && need_model) {
        const float2 *mr = model + (tid / WC) * S + (tid % WC) * SK;
        cp_async8(MQ + tid, mr); cp_async8(MQ + NT + tid, mr + HK);
        cp_async_commit();
    }
    {
        constexpr int HW = WC / 2;
        // task = (channel, bin): HK bins per channel, so a half-warp holds an aligned run of bins (bank-conflict free with
        // fpos) and every warp has the same mix of work; the bin-0 pair of a warp only parks its transform (see P5b)
        static_assert(2 * KCF_CHAN * HK <= NT, "P5 task layout");
        const int task = tid >> 1, half = tid & 1;
        const bool active = task < KCF_CHAN * HK;
        const unsigned msk = __ballot_sync(0xFFFFFFFFu, active);
        if (active) {
            const int c = task / HK, k = task - c * HK;
            const bool k0 = k == 0;
            float2 *const mrow = model + c * S + half * SK + k;        // FFTW layout [c][wc][hr/2+1] (kcf.cpp:180-186): row j' = 2m + half
            // The model column of this bin (HW values per lane) streams straight into the spectrum slots the lane will overwrite with
            // its results: once the pair has the column in registers those slots are dead, so all HW values are in flight through
            // cp.async (SASS LDGSTS) during the transform, without a register or an extra byte of shared memory.
            const bool stream = need_model && !k0;
            float2 a[HW];
            {
                const float2 *const col = F2 + c * WC * HK;
                fft_pair_ld<WC, -1>(a, half, msk, [&](int j) { return col[j * HK + fpos<HK, WC>(k, j)]; }, [&] {
                    if (stream) {
#pragma unroll
                        for (int m = 0; m < HW; ++m) { const int jp = 2 * m + half; cp_async8(F2 + (c * WC + jp) * HK + fpos<HK, WC>(k, jp), mrow + m * 2 * SK); }
                        cp_async_commit();
                    }
                });
            }
            if (k0) {
                // slot 0 = FFT(DC_j + i Nyq_j): parked as it is, in place (row j' of the transform in the bin-0 slot of row j');
                // P5b separates the two columns with all threads
#pragma unroll
                for (int m = 0; m < HW; ++m) { const int jp = 2 * m + half; F2[(c * WC + jp) * HK + fpos<HK, WC>(0, jp)] = a[brev<HW>(m)]; }
            } else {
                if (stream) cp_async_wait_all();
#pragma unroll
                for (int m = 0; m < HW; ++m) {
                    const int jp = 2 * m + half;
                    float2 *const slot = F2 + (c * WC + jp) * HK + fpos<HK, WC>(k, jp);
                    const float2 mv = stream ? *slot : make_float2(0.f, 0.f);
                    const float2 v = a[brev<HW>(m)];
                    if (DUMP && p.dump.spec) p.dump.spec[(long)job * p.dump.stride_spec * KCF_CHAN + c * S + jp * SK + k] = v;
                    float2 o;
                    if (MODE == KCF_MODE_PREDICT) {
                        o = make_float2(v.x * mv.x + v.y * mv.y, v.y * mv.x - v.x * mv.y);                 // xf * conj(model), kcf.cpp:306-345
                    } else {
                        o = make_float2(v.x * v.x + v.y * v.y, 0.f);                                         // |xf|^2, kcf.cpp:269-293
                        // model = (1-f) model + f xf, kcf.cpp:380-395 (f = 1 on the first update: the old model drops out)
                        mrow[m * 2 * SK] = first ? v : make_float2(__fadd_rn(__fmul_rn(omf, mv.x), __fmul_rn(fac, v.x)),
                                                                   __fadd_rn(__fmul_rn(omf, mv.y), __fmul_rn(fac, v.y)));
                    }
                    *slot = o;
                }
            }
        }
    }
    static_assert(KCF_CHAN * WC <= NT, "P5b: one row per thread");
    // square grids: P5b's row (c, j') reads what the bin-0 pair of the SAME channel group parked, and writes only its own slots
    if (G::FUSE45) __syncwarp(); else __syncthreads();
    // ---- P5b: the DC (k = 0) and Nyquist (k = HR/2) columns, one row j' per thread: Z[j'] and Z[-j'] give both
    const unsigned p5b_mask = __ballot_sync(0xFFFFFFFFu, tid < KCF_CHAN * WC);
    if (tid < KCF_CHAN * WC) {
        const int e = tid, c = e / WC, jp = e - c * WC, jm = (WC - jp) & (WC - 1);
        const float2 v = F2[(c * WC + jp) * HK + fpos<HK, WC>(0, jp)], y = F2[(c * WC + jm) * HK + fpos<HK, WC>(0, jm)];
        cp_async_wait_all();
        const float2 m0n = need_model ? MQ[tid] : make_float2(0.f, 0.f), m1n = need_model ? MQ[NT + tid] : make_float2(0.f, 0.f);
        __syncwarp(p5b_mask);          // rows j' and -j' of a channel sit in the same warp: both have read before either slot is overwritten
        const float2 dc = make_float2(0.5f * (v.x + y.x), 0.5f * (v.y - y.y));
        const float2 nq = make_float2(0.5f * (v.y + y.y), 0.5f * (y.x - v.x));
        const int sp0 = c * S + jp * SK, sp1 = sp0 + HK;
        if (DUMP && p.dump.spec) { float2 *d = p.dump.spec + (long)job * p.dump.stride_spec * KCF_CHAN; d[sp0] = dc; d[sp1] = nq; }
        float2 o0, o1;
        if (MODE == KCF_MODE_PREDICT) {
            o0 = make_float2(dc.x * m0n.x + dc.y * m0n.y, dc.y * m0n.x - dc.x * m0n.y);
            o1 = make_float2(nq.x * m1n.x + nq.y * m1n.y, nq.y * m1n.x - nq.x * m1n.y);
        } else {
            o0 = make_float2(dc.x * dc.x + dc.y * dc.y, 0.f);
            o1 = make_float2(nq.x * nq.x + nq.y * nq.y, 0.f);
            model[sp0] = first ? dc : make_float2(__fadd_rn(__fmul_rn(omf, m0n.x), __fmul_rn(fac, dc.x)), __fadd_rn(__fmul_rn(omf, m0n.y), __fmul_rn(fac, dc.y)));
            model[sp1] = first ? nq : make_float2(__fadd_rn(__fmul_rn(omf, m1n.x), __fmul_rn(fac, nq.x)), __fadd_rn(__fmul_rn(omf, m1n.y), __fmul_rn(fac, nq.y)));
        }
        F2[(c * WC + jp) * HK + fpos<HK, WC>(0, jp)] = o0;
        FN[e] = o1;
    }
    __syncthreads();

    // ------------------------------------------------------------------ P6: sum over the 31 channels (in channel order)
    float *const alpha = p.alpha + (long)jd.slot * p.alpha_stride;
    static_assert(S <= NT, "P6: one bin per thread");
    if (tid < S) {
        // threads [0, WC*HK): packed bins, a half-warp along k of one row (conflict-free with fpos); the rest: the Nyquist bins
        const int j = tid < WC * HK ? tid / HK : tid - WC * HK, k = tid < WC * HK ? tid - j * HK : HK;
        const int e = j * SK + k;
        // per-bin inputs first: their global-memory latency hides behind the 31-channel sum
        const float al0 = alpha[e];
        const float yf0 = (MODE == KCF_MODE_UPDATE) ? cls.yf_re[e] : 0.f;
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < KCF_CHAN; ++c) {
            const float2 v = (k < HK) ? F2[(c * WC + j) * HK + fpos<HK, WC>(k, j)] : FN[c * WC + j];
            if (c == 0) acc = v; else { acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); }
        }
        if (MODE == KCF_MODE_PREDICT) {
            const float al = al0;
            acc.x = __fmul_rn(__fmul_rn(acc.x, al), cls.norm);         // kcf.cpp:356-357
            acc.y = __fmul_rn(__fmul_rn(acc.y, al), cls.norm);
            zf_s[e] = acc;
            if (DUMP && p.dump.zf) p.dump.zf[(long)job * p.dump.stride_spec + e] = acc;
        } else {
            const float kf = __fmul_rn(acc.x, cls.norm);               // kcf.cpp:295-303
            if (DUMP && p.dump.kf) p.dump.kf[(long)job * p.dump.stride_spec + e] = kf;
            const float an = __fdiv_rn(yf0, __fadd_rn(kf, p.lamda));                         // kcf.cpp:373
            alpha[e] = first ? an : __fadd_rn(__fmul_rn(omf, al0), __fmul_rn(fac, an));      // kcf.cpp:374 (factor 1 on the first update: the old value drops out)
        }
    }
    if (MODE == KCF_MODE_UPDATE) {
        if (tid == 0) {
            // tracker_update, kcf.cpp:462-476
            KcfMeta *const mt = p.meta + jd.slot;
            const mot_bbox_t bx = jd.box;
            mt->pos = bx;
            mt->scale_horiz = __fdiv_rn((float)(bx.r - bx.l + 1), (float)jd.cols);
            mt->scale_vert = __fdiv_rn((float)(bx.b - bx.t + 1), (float)jd.rows);
            mt->first_update = 0;
        }
        continue;                                      // next job (uniform: MODE is a template parameter)
    }
    p7_pending = true;                                 // P7 of this job runs underneath P0 of the next one (or after the loop)
    }   // persistent job loop
    if (MODE == KCF_MODE_PREDICT && p7_pending) {
        __syncthreads();
        if (tid < 64) p7_tail(s_desc[(it - 1) & 3], (int)blockIdx.x + (it - 1) * (int)gridDim.x);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
template <int HR, int WC> int kcf_launch_size(int mode, const KcfLaunch &p, cudaStream_t s)
{
    const FhogTablesDev &t = p.tab;
    int lut_floats = 2 * (2 << t.rsqrt_bits) + ((2 * t.bin_nseg + 3) & ~3);      // fused {rsqrt, rcp} table + bin step table
    if (lut_floats > 8192) lut_floats = 0;                                      // too large: read the tables from global memory
    const size_t bytes = smem_bytes<HR, WC>(lut_floats);
    const bool dump = p.dump.gray != nullptr;
    const void *fn;
    if (mode == KCF_MODE_PREDICT) fn = dump ? (const void *)kcf_fused_kernel<HR, WC, KCF_MODE_PREDICT, true> : (const void *)kcf_fused_kernel<HR, WC, KCF_MODE_PREDICT, false>;
    else                          fn = dump ? (const void *)kcf_fused_kernel<HR, WC, KCF_MODE_UPDATE, true> : (const void *)kcf_fused_kernel<HR, WC, KCF_MODE_UPDATE, false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    // persistent CTAs: one per SM (223 KB of shared memory each at 32x32), looping over the jobs
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = p.n_jobs < sms ? p.n_jobs : sms;
    void *args[2] = { (void *)&p, (void *)&lut_floats };
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(KCF_THREADS), args, bytes, s);
    return (int)e;
}

}  // namespace mot
