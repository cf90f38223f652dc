// mot_ctx.h -- private: the context object behind mot_ctx_t, shared by the C-ABI translation units.
#pragma once
#include "mot_internal.h"
#include "kalman.h"
#include "assoc.h"
#include "kcf_any.cuh"

#include <algorithm>
#include <string>
#include <vector>

int mot_fail(int code, const char *fmt, ...);      // records the message for mot_last_error() and returns code
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return mot_fail(MOT_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)


namespace mot {

// fast: a fixed-size fused kernel exists (kcf_fused.cuh); any_smem > 0: the fused any-size kernel (kcf_any.cu) holds the window in
// that much shared memory; neither: the unfused path (kcf_generic.cu)
struct SizeClass { int hr, wc, live; bool fast; size_t any_smem; float *d_wy, *d_wx, *d_yf; double2 *d_twh, *d_tww; float norm; };

template <class T> struct DevBuf {
    T *p = nullptr; size_t n = 0;
    // grows geometrically; the old block is released only after the new one exists, so a failed allocation leaves the buffer
    // as it was (contents are NOT carried over: every user refills the buffer after ensure())
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        const size_t nn = std::max(want, n * 2);
        T *q = nullptr;
        const cudaError_t e = cudaMalloc(&q, sizeof(T) * nn);
        if (e != cudaSuccess) return e;
        if (p) cudaFree(p);
        p = q; n = nn;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};
template <class T> struct PinBuf {
    T *p = nullptr; size_t n = 0;
    cudaError_t ensure(size_t want) {
        if (want <= n) return cudaSuccess;
        const size_t nn = std::max(want, n * 2);
        T *q = nullptr;
        const cudaError_t e = cudaMallocHost(&q, sizeof(T) * nn);
        if (e != cudaSuccess) return e;
        if (p) cudaFreeHost(p);
        p = q; n = nn;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; n = 0; }
};

}  // namespace mot

using namespace mot;

struct mot_ctx_s {
    int device = 0, W = 0, H = 0, max_tracks = 0, n_frames = 0, kind = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // frame uploads run here so that they overlap with kernels of the compute stream
    std::vector<cudaEvent_t> slot_uploaded;   // per frame slot: recorded after its last upload
    std::vector<char> slot_pending;           // an upload of this slot has not been waited for by the compute stream yet
    long launches = 0;
    // frames
    std::vector<uint8_t *> frame_owned;
    std::vector<const uint8_t *> frame_ptr_h;
    const uint8_t **d_frame_ptr = nullptr;
    void *d_frame_tmaps = nullptr;        // [n_frames][3] tensor maps of the frame slots (2-D TMA crop loads of the fixed-size fused kernels)
    bool tmaps_ok = false;
    bool frame_ptr_dirty = true;
    int frame_stride = 0;
    // track slots
    std::vector<int> free_slots;
    std::vector<char> used;
    std::vector<KcfMeta> meta_h;          // host mirror of the immutable part (rows, cols, hr, wc, size_class)
    KcfMeta *d_meta = nullptr;
    float2 *d_model = nullptr; long model_stride = 0;
    float *d_alpha = nullptr; long alpha_stride = 0;
    std::vector<SizeClass> classes;
    KcfClassDev *d_classes = nullptr;
    FhogTablesDev tab{};
    AnyTablesDev any{};                   // per-N tables of the any-size kernel (Hann, twiddles, label spectra, radix plans)
    int lut_floats = 0;                   // floats of the SSE tables a fused kernel stages in shared memory
    int sm_count = 0;
    KcfExt ext{};                         // north-star extensions (mot_ctx_set_kcf_options); all zero = the reference's filter
    bool ext_on = false;
    float *d_alpha_im = nullptr;          // imaginary part of alpha (Gaussian kernel), same slot stride as d_alpha
    int *h_any_err = nullptr, *d_any_err = nullptr;   // mapped pinned word the any-size kernel sets when a job does not fit its launch
    float *d_tab_rsqrt = nullptr, *d_tab_rcp = nullptr, *d_tab_rsrc = nullptr; uint32_t *d_tab_bin = nullptr, *d_tab_bin2 = nullptr;
    KalmanState kal{};
    // staging
    DevBuf<int> d_slots, d_frames, d_TD, d_assign;
    DevBuf<mot_bbox_t> d_boxes, d_trk, d_det;
    DevBuf<KcfMeta> d_meta_stage;
    DevBuf<float> d_gray;
    DevBuf<char> d_scratch;               // per-job intermediates of the any-size KCF path
    DevBuf<double> d_dist, d_work, d_cost;
    PinBuf<int> h_slots, h_frames, h_TD, h_assign;
    PinBuf<mot_bbox_t> h_boxes, h_trk, h_det;
    PinBuf<KcfMeta> h_meta_stage;
    PinBuf<double> h_dist, h_cost;
    // dumps
    bool dumps = false;
    KcfDump dump{};
    int dump_hr = 0, dump_wc = 0, dump_rows = 0, dump_cols = 0;
};

// the same over an any-size job list: every job's shared-memory plan fits `smem_bytes` (csrc/kcf_any.cuh); scratch != null: a list
// of STRIP-MODE windows, each of the (at most scratch_ctas) CTAs parking its histograms in scratch_stride_floats of global memory
int mot_ctx_kcf_launch_any(mot_ctx_t *c, int mode, size_t smem_bytes, int n_max, const int *n_dev, const int *slots, const int *frames,
                           mot_bbox_t *boxes, const int *box_index, int clamp, float *scratch, long scratch_stride_floats, int scratch_ctas);
// ---- internal services of mot_capi.cu used by the device-resident frame loop (csrc/td_device.cu) ---------------------------
int mot_ctx_kcf_class(mot_ctx_t *c, int hr, int wc, int *cls_out);       // index of the (created on demand) per-size constant tables
int mot_ctx_frames_ready(mot_ctx_t *c);                                  // frame pointer table on the device, pending uploads waited for
int mot_ctx_associate_dev(mot_ctx_t *c, int n_mat, const int *d_T, const int *d_D, const mot_bbox_t *d_trk, long trk_stride,
                          const mot_bbox_t *d_det, long det_stride, int cost_mode, double *d_dist, long dist_stride,
                          int *d_assign, long assign_stride, double *d_cost, int max_dim, double *d_work);
// one fused launch for class `cls` over a device-resident job list: n_dev jobs (at most n_max), job j = (slots[j], frames[j],
// boxes[box_index[j]])
int mot_ctx_kcf_launch(mot_ctx_t *c, int mode, int cls, int n_max, const int *n_dev, const int *slots, const int *frames,
                       mot_bbox_t *boxes, const int *box_index, int clamp);
