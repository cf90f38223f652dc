/*
 * mot_b200.h -- C ABI of the B200-native tracking back end (libmot_b200.so).
 *
 * Drop-in boundary for the data-parallel hot path of huangfcn/multiple-object-tracking:
 *   fHOG  ->  KCF/DCF correlation filter  ->  Kalman predict/update  ->  cost matrix + Hungarian.
 * Every entry point names the reference interface it replaces (paths relative to the reference root).
 * Plain pointers and sizes only; no C++ or torch types.  All functions return 0 on success and a
 * negative code on failure; mot_last_error() gives the text.  Nothing here ever falls back to a CPU
 * implementation: without a CUDA device (or with an unsupported shape) the call FAILS.
 *
 * The reference's own per-object plugin signatures (top/td.cpp:229-261, C++ linkage)
 *     void* tracker_new(bbox_t*); void tracker_predict(void*, float*, bbox_t*);
 *     void tracker_update(void*, float*, bbox_t*); void tracker_delete(void*);
 *     void assignmentoptimal(int*, double*, double*, int, int);
 *     extern "C" rgb2Gray(...), bilinearInterpolationGray(...)
 * are provided on top of this ABI by multiple-object-tracking_b200/host/tracker_shim.cpp, with bbox_t = struct _bbox_pos_s as in
 * top/cnntype.h:36-41 so that the mangled names are the ones top/td.cpp links against (tests/test_link_plugin.py links a client
 * written against the reference's declarations with -Wl,--no-undefined).
 */
#ifndef MOT_B200_H
#define MOT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)      /* the library itself is built with -fvisibility=hidden */
#endif

/* top/cnntype.h:36-41 -- layout-identical to bbox_t: inclusive pixel coordinates, field order l,t,b,r. */
typedef struct mot_bbox_s { int l, t, b, r; int type; float score; } mot_bbox_t;

/* top/cnntype.h:43-47 -- layout-identical to bbox_chain_t (the detector's per-frame output). */
typedef struct mot_bbox_chain_s { int nbox; mot_bbox_t bbox[128]; } mot_bbox_chain_t;

/* Which of the reference's two link-time tracker plugins a context emulates
 * (yolo3tracker.vcxproj:134-138 links trackers/kalman.cpp; #define KCF_TRACKER, top/td.cpp:47, selects trackers/kcf.cpp). */
enum { MOT_TRACKER_KALMAN = 0, MOT_TRACKER_KCF = 1 };

/* Association cost (top/td.cpp:386-457).  REF_CENTROID is the shipped expression; IOU_CLAMPED is the finite form of
 * the commented-out IoU cost (:404, :439), whose literal form goes negative / -inf and hangs the reference solver. */
enum { MOT_COST_REF_CENTROID = 0, MOT_COST_IOU_CLAMPED = 1 };

enum {
    MOT_OK = 0, MOT_ERR_CUDA = -1, MOT_ERR_ARG = -2, MOT_ERR_SHAPE = -3, MOT_ERR_CAPACITY = -4,
    MOT_ERR_TABLES = -5, MOT_ERR_KIND = -6
};

typedef struct mot_ctx_s mot_ctx_t;

/* ---- context ------------------------------------------------------------------------------------------ */

/* One context per GPU and per host thread (the reference calls the plugin from exactly one thread, top/td.cpp:306).
 * frame_w/frame_h: the reference hard-codes 1280x720 (top/cnntype.h:5-6); here they are parameters.
 * max_tracks: track slots (reference cap 256, top/td.cpp:12); n_frame_slots: frames resident at once (one per stream). */
int mot_ctx_create(mot_ctx_t **out, int device, int frame_w, int frame_h, int max_tracks, int n_frame_slots, int tracker_kind);
void mot_ctx_destroy(mot_ctx_t *ctx);
const char *mot_last_error(void);
/* Run all work of this context on the caller's CUDA stream (a cudaStream_t passed as void*); NULL = the context's own. */
int mot_ctx_set_stream(mot_ctx_t *ctx, void *cuda_stream);
int mot_sync(mot_ctx_t *ctx);
/* MOT_TRACKER_KALMAN or MOT_TRACKER_KCF */
int mot_ctx_kind(mot_ctx_t *ctx);
/* Kernels launched by this context since creation (for bench.py's gpu_launches). */
long mot_launch_count(mot_ctx_t *ctx);

/* ---- north-star extensions of the correlation filter -------------------------------------------------------------
 * The reference's filter is the LINEAR-kernel DCF with an integer peak, no padding and a fixed label sigma (trackers/kcf.cpp:269-304,
 * 397-428, 205); that is the default here and the only mode with a reference oracle.  These options add what the reference lacks:
 * the Gaussian kernel correlation of KCF (Henriques et al.), a parabolic sub-pixel refinement of the response peak, a tracking window
 * larger than the target box (padding), and a label width that follows the target size.  Parity UNPINNED by the reference: validated
 * against the NumPy restatement tests/kcf_ext_numpy.py.  Set before the first tracker exists; every window then runs in the fused
 * any-size kernel (up to 1152 spectrum bins).  All zero = reference behaviour. */
typedef struct mot_kcf_options_s {
    int gaussian_kernel;          /* 1: k = exp(-max(0, |x|^2 + |z|^2 - 2 x.z) / (sigma^2 numel)), alpha complex */
    float kernel_sigma;           /* its bandwidth (KCF on HOG: 0.5) */
    int subpixel_peak;            /* 1: refine the peak by a parabola through its circular neighbours; the box moves by 4 (peak + delta) scale */
    float padding;                /* window = target box x padding about its centre (0 or 1: none; north-star: 2.5); boxes in/out stay TARGET boxes */
    float output_sigma_factor;    /* label sigma = sqrt(target w x h) x factor / cell (0: the reference's fixed 0.7289 cells; KCF: 0.1) */
} mot_kcf_options_t;
int mot_ctx_set_kcf_options(mot_ctx_t *ctx, const mot_kcf_options_t *opt);

/* KCF contexts: size the slots of the model arena for windows up to max_rows x max_cols pixels (either orientation).  The reference
 * allocates every tracker's spectra for whatever size its detection box has (trackers/kcf.cpp:146-195); here a slot holds 1152
 * half-spectrum bins by default (the named 128x128 shape: 544), host-managed trackers beyond that own individually allocated storage,
 * and the device-resident loop can only spawn windows a slot holds.  Call before the first tracker exists; memory = 252 bytes per bin
 * and track slot. */
int mot_ctx_reserve_window(mot_ctx_t *ctx, int max_rows, int max_cols);

/* ---- frames (replaces the cv::Mat the tracking thread pops, top/td.cpp:330-331) --------------------------- */

/* Copy a host BGR u8 frame (rows of stride_bytes) into frame slot `slot` (async on the context stream). */
int mot_frame_upload(mot_ctx_t *ctx, int slot, const uint8_t *host_bgr, int stride_bytes);
/* Zero-copy: make frame slot `slot` refer to a caller-owned DEVICE buffer (BGR u8, rows of stride_bytes). */
int mot_frame_bind_device(mot_ctx_t *ctx, int slot, const uint8_t *dev_bgr, int stride_bytes);
/* Copy frame slot `slot` back to the host (rows of stride_bytes >= 3 * frame_w); synchronous. */
int mot_frame_download(mot_ctx_t *ctx, int slot, uint8_t *host_bgr, int stride_bytes);

/* ---- overlay (replaces the rectangle loop of the tracking thread, top/td.cpp:647-733, and drawRect,
 *      top/drawlib.c:97-151) ------------------------------------------------------------------------------------ */

/* For entry i, in order: `thickness` nested one-pixel rectangles (box shrunk by 0, 1, ... pixels; the reference draws 3)
 * of colour rgb[i] (0xRRGGBB, stored as bytes R, G, B like drawRect does) into frame slot frame_slots[i].  Later entries
 * overwrite earlier ones; reversed boxes are swapped per rectangle; addressing is linear (y * stride + 3 * x) with no
 * clipping, exactly as drawRect, except that bytes outside the frame buffer are not written.  A frame bound with
 * mot_frame_bind_device is written in place. */
int mot_overlay_batch(mot_ctx_t *ctx, int n, const int *frame_slots, const mot_bbox_t *boxes, const uint32_t *rgb, int thickness);
/* The colour the reference gives track `tid`: colormap[hashcolor(tid + 1) & 255] -- top/td.cpp:619-620 hashes the id counter
 * after `tid = tracker_id++` (hashcolor :295-305, palette :652-699). */
uint32_t mot_track_color(uint32_t tid);

/* ---- detector post-processing (replaces decode_netout / correct_yolo_boxes / sort / do_nms, detectors/yolo3.cpp:141-356, and
 *      the per-image driver in tensorRunB, :487-527) ---------------------------------------------------------------------- */

/* out0/out1/out2: the three YOLO3 output maps of ONE image (host memory; grids (tensor/32) << 0, 1, 2; 3 anchors x (5 + classes)
 * values per cell), anchors18 as in yolo3_options_t (:88-94).  Writes the detections exactly as the reference would put them into
 * bbox_chain_t (same boxes, classes, scores and order: the C library's expf, the exchange sort and the never-cleared suppression
 * flags are reproduced) and returns their number (>= 0), or a negative error (MOT_ERR_CAPACITY beyond 4096 candidates). */
int mot_yolo_post(mot_ctx_t *ctx, const float *out0, const float *out1, const float *out2, const int *anchors18, float obj_thresh, float nms_thresh,
                  int tensor_h, int tensor_w, int image_h, int image_w, int num_classes, mot_bbox_t *out, int max_out);

/* ---- tracker plugin, batched (replaces tracker_new/predict/update/delete, trackers/kcf.cpp:455-491 and
 *      trackers/kalman.cpp:131-163; one call = the reference's loop over tracks, top/td.cpp:344-384, 512-582) --- */

/* tracker_new for n boxes; writes n handles (slot numbers >= 0). */
int mot_tracker_new_batch(mot_ctx_t *ctx, int n, const mot_bbox_t *boxes, int *handles_out);
int mot_tracker_delete_batch(mot_ctx_t *ctx, int n, const int *handles);
/* 1 when mot_tracker_new_batch can build a tracker for this box (Kalman: always; KCF: at least 2x2 cells = 8x8 px and not larger
 * than the frame), else 0.  The frame loops use it to count-and-skip such detections instead of failing the step. */
int mot_tracker_spawnable(mot_ctx_t *ctx, const mot_bbox_t *box);

/* tracker_predict for n tracks, fused with the crop + gray + resize in front of it (top/td.cpp:348-364):
 * boxes[i] in  = crop rectangle in frame frame_slots[i] (the caller's tracker_info.bbox),
 * boxes[i] out = predicted box.  clamp != 0 also applies make_x/y_in_range (top/td.cpp:378-381).
 * HOST arrays; blocks until the result is back.  For the Kalman kind frame_slots may be NULL. */
int mot_predict_batch(mot_ctx_t *ctx, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp);
/* tracker_update for n tracks, fused with crop + gray + resize (top/td.cpp:517-540, 558-580): boxes[i] is both the
 * crop rectangle and the measurement / new position.  HOST arrays; asynchronous on the context stream. */
int mot_update_batch(mot_ctx_t *ctx, int n, const int *handles, const int *frame_slots, const mot_bbox_t *boxes);

/* tracker_predict (+ clamp) followed at once by tracker_update with the predicted box -- the reference's treatment of a track
 * without a detection (top/td.cpp:344-384 then :550-582), and the whole per-frame work of single-target tracking (BASELINE config 1) --
 * as ONE call with one synchronisation.  boxes[i] in = crop rectangle, out = predicted (clamped) box.  HOST arrays. */
int mot_track_batch(mot_ctx_t *ctx, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp);

/* Same two calls with every array already resident on the DEVICE (no copies, no sync): the steady-state path. */
int mot_predict_batch_dev(mot_ctx_t *ctx, int n, const int *d_handles, const int *d_frame_slots, mot_bbox_t *d_boxes, int clamp);
int mot_update_batch_dev(mot_ctx_t *ctx, int n, const int *d_handles, const int *d_frame_slots, const mot_bbox_t *d_boxes);

/* The literal plugin form: the caller already cropped/resized a gray patch (column-major rows x cols f32, HOST),
 * exactly what tracker_predict/tracker_update receive as `rgb` (trackers/kcf.cpp:455-476). */
int mot_predict_gray(mot_ctx_t *ctx, int handle, const float *gray_host, mot_bbox_t *box_out);
int mot_update_gray(mot_ctx_t *ctx, int handle, const float *gray_host, const mot_bbox_t *box);

/* ---- patch preprocessing on its own (replaces rgb2Gray + bilinearInterpolationGray, top/drawlib.c:192-240, 542-637) */
int mot_crop_gray_resize(mot_ctx_t *ctx, int frame_slot, const mot_bbox_t *box, int rows_d, int cols_d, float *gray_host_out);

/* The reference's two patch helpers with their own signatures' semantics, HOST pointers in and out (one upload, one kernel, one
 * download): what host/tracker_shim.cpp exports as the C-linkage symbols rgb2Gray / bilinearInterpolationGray of top/td.cpp:245-261.
 * mot_rgb2gray_host: crop [l..r] x [t..b] (swapped when reversed, top/drawlib.c:203-215) of a BGR u8 frame with rows of stride_bytes
 * (the reference hard-codes 3840, :9-10) -> gray f32, column-major rows x cols; reads exactly the bytes the reference reads (no clipping).
 * mot_resize_gray_host: top/drawlib.c:542-637 literally -- both buffers row-major height x width. */
int mot_rgb2gray_host(mot_ctx_t *ctx, const uint8_t *host_bgr, int stride_bytes, int l, int t, int r, int b, float *gray_host_out);
int mot_resize_gray_host(mot_ctx_t *ctx, float *dst_host, const float *src_host, int h_src, int w_src, int h, int w);

/* ---- association (replaces the cost loops top/td.cpp:386-457 and assignmentoptimal, trackers/hungarian/hungarian.cpp:29) */

/* n_mat independent problems.  Problem m has T[m] tracker boxes and D[m] detection boxes stored at trk + m*trk_stride,
 * det + m*det_stride.  Cost matrix m is written column-major with rows = the smaller side (trackers if T<D else
 * detections) at dist + m*dist_stride (may be NULL to skip the dump); assignment m (one int per ROW, column or -1) at
 * assign + m*assign_stride; total cost at cost[m].  HOST arrays. */
int mot_associate_batch(mot_ctx_t *ctx, int n_mat, const int *T, const int *D,
                        const mot_bbox_t *trk, long trk_stride, const mot_bbox_t *det, long det_stride,
                        int cost_mode, double *dist, long dist_stride, int *assign, long assign_stride, double *cost);
/* assignmentoptimal on caller-supplied matrices (column-major nrows x ncols doubles), n_mat of them, HOST arrays. */
int mot_assign_batch(mot_ctx_t *ctx, int n_mat, const int *nrows, const int *ncols, const double *dist, long dist_stride,
                     int *assign, long assign_stride, double *cost);
/* Device-resident forms (no copies, no sync). */
int mot_associate_batch_dev(mot_ctx_t *ctx, int n_mat, const int *d_T, const int *d_D,
                            const mot_bbox_t *d_trk, long trk_stride, const mot_bbox_t *d_det, long det_stride,
                            int cost_mode, double *d_dist, long dist_stride, int *d_assign, long assign_stride, double *d_cost,
                            int max_dim);

/* ---- frame loop (replaces one iteration of pthread_mtcnn_trkn, top/td.cpp:343-644) -------------------------- */
typedef struct mot_td_s mot_td_t;
/* One loop state per stream; frame_slot = where this stream's frames are uploaded. */
int mot_td_create(mot_td_t **out, mot_ctx_t *ctx, int frame_slot, int cap, int cost_mode);
void mot_td_destroy(mot_td_t *td);
/* host_bgr may be NULL when the frame slot was already filled (or for the Kalman kind, which never reads pixels). */
int mot_td_step(mot_td_t *td, const uint8_t *host_bgr, int stride_bytes, const mot_bbox_t *dets, int ndet);
/* The same step fed with the detector's wire format: one bbox_chain_t per frame (top/cnntype.h:43-47; the tracking thread reads
 * pdetected->nbox and ->bbox[], top/td.cpp:326-335), and a detector batch as tensorRunB delivers it (top/td.cpp:178-204: up to
 * MAX_GPU_BATCH = 4 frames, one chain each), consumed frame by frame in queue order. */
int mot_td_step_chain(mot_td_t *td, const uint8_t *host_bgr, int stride_bytes, const mot_bbox_chain_t *chain);
int mot_td_step_chain_batch(mot_td_t *td, int nbatch, const uint8_t *const *host_bgr, int stride_bytes, const mot_bbox_chain_t *const *chains);
/* Lock-step over several streams: one batched predict / associate / update per frame instead of one per stream. */
int mot_td_step_multi(mot_td_t **tds, int n_streams, const uint8_t *const *host_bgr, int stride_bytes,
                      const mot_bbox_t *const *dets, const int *ndet);
int mot_td_ntracks(mot_td_t *td);
/* detections skipped so far because no tracker can be built for their window (see mot_tracker_spawnable) */
long mot_td_dropped(mot_td_t *td);
void mot_td_get(mot_td_t *td, uint32_t *tid, mot_bbox_t *boxes, int *age, int *vis, int *invis);
/* The overlay loop of top/td.cpp:647-733 for the current track table, drawn on the device into the loop's frame slot
 * (three nested rectangles per track in colour mot_track_color(tid)); read the frame back with mot_frame_download. */
int mot_td_overlay(mot_td_t *td);
int mot_td_last(mot_td_t *td, mot_bbox_t *predicted, int *assigned_trackers);

/* ---- the same frame loop with the track tables resident on the DEVICE: for the Kalman kind ONE launch per frame for all
 *      streams (a CTA per stream; six launches beyond 256 tracks or detections per stream), for the KCF kind the fused kernels over per-class job lists built on the device; no host synchronisation (replaces top/td.cpp:343-644 including the bookkeeping,
 *      the stable compaction of lost tracks :585-609 and the spawn order :612-644) --------------------------------- */
typedef struct mot_tdd_s mot_tdd_t;
/* n_streams independent streams, at most cap tracks (reference: 256, top/td.cpp:12) and max_det detections
 * (reference: 128, top/cnntype.h:46) each, both <= 1024.  The context needs >= n_streams*cap free slots and no host-managed
 * trackers.  KCF contexts: stream s reads frame slot base + s (mot_tdd_frame_base; mot_frame_upload / mot_frame_bind_device
 * before each step).  A detection of any window size a fused kernel serves spawns a track on the device: the fixed-size kernels
 * (cell grid sides 8, 16, 32), the any-size kernel in shared memory (up to about 1400 cells) and in strip mode (up to about 8000
 * cells = e.g. 360 x 360 px), provided its half spectrum fits a slot of the model arena (1152 bins by default: call
 * mot_ctx_reserve_window before mot_tdd_create for larger windows).  Detections beyond that, smaller than 8x8 px or larger than the
 * frame are counted (mot_tdd_dropped) -- the host-side loop mot_td_step serves every size. */
int mot_tdd_create(mot_tdd_t **out, mot_ctx_t *ctx, int n_streams, int cap, int max_det, int cost_mode);
void mot_tdd_destroy(mot_tdd_t *tdd);
/* detections already on the device: d_dets[n_streams][max_det], d_ndet[n_streams]; asynchronous on the context stream */
int mot_tdd_step_dev(mot_tdd_t *tdd, const mot_bbox_t *d_dets, const int *d_ndet);
/* detections in host arrays (dets[s] points at ndet[s] boxes); staged, uploaded and stepped; asynchronous */
int mot_tdd_step(mot_tdd_t *tdd, const mot_bbox_t *const *dets, const int *ndet);
/* one bbox_chain_t per stream (NULL = no detections): the detector's wire format; max_det must cover the chains' nbox */
int mot_tdd_step_chains(mot_tdd_t *tdd, const mot_bbox_chain_t *const *chains);
/* KCF kind: from the next step on, stream s reads frame slot base + s (default 0).  Steps are asynchronous: do not upload
 * into a slot that a step still in flight reads -- alternate two bases (2 * n_streams frame slots) or call mot_sync first. */
int mot_tdd_frame_base(mot_tdd_t *tdd, int base);
/* KCF kind: declare the window sizes (pixels) the detections will have; the kernels of all other classes are not launched */
int mot_tdd_kcf_windows(mot_tdd_t *tdd, int n, const int *rows, const int *cols);
/* KCF kind: number of detections of stream s that could not spawn a track so far (synchronises) */
int mot_tdd_dropped(mot_tdd_t *tdd, int stream);
/* snapshot of stream s (synchronises); returns the number of live tracks or a negative error */
int mot_tdd_read(mot_tdd_t *tdd, int stream, uint32_t *tid, mot_bbox_t *boxes, int *age, int *vis, int *invis);

/* ---- test hooks: stage dumps of the fused KCF kernel ------------------------------------------------------- */
/* Stages: 0 gray | 1 m0 | 2 bin(int) | 3 r1 | 4 norm | 5 feat(xf_tm) | 6 spec(xf_fq, float pairs) | 7 zf | 8 response |
 *         9 kf | 10 peak(int x2) | 11 margin(float x2).  Enable before a predict/update of ONE track, then fetch. */
int mot_debug_enable_dumps(mot_ctx_t *ctx, int enable);
long mot_debug_fetch(mot_ctx_t *ctx, int stage, void *host_out, long max_bytes);
/* Raw per-slot state: which = 0 xf_md (31*S float pairs) | 1 alpha (S floats) | 2 Kalman x[6] | 3 Kalman P[36] col-major |
 * 4 extension: sub-cell refinement {vertical, horizontal} of the last predicted peak | 5 extension: Im(alpha) (S floats). */
long mot_debug_state(mot_ctx_t *ctx, int handle, int which, void *host_out, long max_bytes);
/* Host-harvested SSE tables (no GPU needed): which = 0 rsqrt | 1 rcp | 2 acos(20020) | 3 fused {rsqrt, rcp(rsqrt)/16} pairs |
 * 4 orientation-bin step table (u32 bits) | 5 the same with the wrap folded in (u32 bits) | 6 {saturation threshold bits, rcp(1e10f)};
 * info[0..3] = rsqrt_bits, rcp_bits, bin_shift, bin_nseg */
long mot_debug_tables(int which, float *out, long max_floats, int *info);

/* How the fused any-size kernel would run an hr x wc cell grid (host-only, no GPU needed): out[16] = ok, strip mode, shared-memory
 * bytes, channels per tile, gradient strip width, cell columns per strip, spectral buffer (complex), region A / B floats, scratch bytes
 * per CTA, CTAs per SM, threads per CTA, passes of the length-hr / length-wc transforms; radices[14] = their radices (7 slots each). */
int mot_debug_any_plan(int hr, int wc, int *out, int *radices);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
