// td_device.cu -- the frame loop of the reference with the track table RESIDENT ON THE DEVICE (both tracker kinds).
//
// One iteration of pthread_mtcnn_trkn (top/td.cpp:343-644) for S independent streams is ONE launch (td_frame_kalman_kernel, a CTA per
// stream, up to 256 tracks / detections per stream) or six and no host
// synchronisation: kalman_predict (+clamp, :344-384) -> cost matrices + Munkres (:386-470) -> td_scatter (:472-547,
// :550-556 bookkeeping) -> kalman_update (:539, :581) -> td_lifecycle (delete lost with stable compaction :585-609, spawn
// per unassigned detection in ascending order :612-644).  The bookkeeping, the order of the track table and the ids are
// the reference's, per stream; host/td_loop.cpp is the same loop with the table on the host.
// Stream s owns the tracker slots [s*cap, (s+1)*cap).
//
// KCF kind: the same loop with the fused KCF kernels in place of the Kalman ones.  The track table does not say which
// window size a track has in a form the host could read without synchronising, so every frame a small kernel sorts the
// live tracks into job lists: one per fixed-size fused kernel (cell grids with sides 8/16/32) and three for the fused
// any-size kernel (kcf_any.cu), by the shared memory a window needs (four, two or one CTA per SM); the kernels take their
// job count from the device and exit at once when their list is empty; a fourth any-size list holds the STRIP-MODE windows (those
// beyond one CTA's shared memory, about 1400 to 8000 cells), launched with a per-CTA scratch area owned by the loop object.  A tracker
// of ANY size the fused any-size kernel serves is therefore
// born on the device: its constants come from the per-N tables, nothing is computed on the host.  tracker_new (trackers/kcf.cpp:484-491, :139-213) runs
// inside the lifecycle kernel (metadata only: model and alpha are fully written by the first update), followed by one
// more update launch over the tracks spawned in this frame (the reference's first update, top/td.cpp:629-641).  Stream s
// reads frame slot frame_base + s (mot_tdd_frame_base: alternate two bases to upload frame k+1 under the kernels of frame k).
// Only detections no tracker can be built for here -- smaller than 2x2 cells, larger than the frame, beyond the any-size kernel's
// strip mode (about 8000 cells), or with more half-spectrum bins than a slot of the context's arena holds (1152 by default,
// mot_ctx_reserve_window raises it) -- are skipped and counted (mot_tdd_dropped); the host-side loop (host/td_loop.cpp) serves the
// last two kinds through individually allocated models and the unfused path.
#include "mot_ctx.h"
#include "kalman.cuh"
#include "assoc.cuh"
#include <cstdlib>

namespace mot {

struct TddState {
    int S, cap, max_det;
    int *ntracks; uint32_t *tracker_id;          // [S]
    uint32_t *tid; int *slot, *age, *vis, *invis; mot_bbox_t *bbox;      // [S][cap]; slot = -1 beyond ntracks
    int *assign;                                 // [S][md] rows of the cost matrix -> column
    int *assigned_detected;                      // [S][max_det]
    int md;
    // KCF kind
    int kcf, frame_w, frame_h;
    int frame_base;                              // stream s reads frame slot frame_base + s
    int cls_id[9];                               // context class index of fused class 3*hi + wi (cell sides 8, 16, 32); -1: disabled
    int any_on[4];                               // any-size job lists (list 9 + b) enabled; b = 3: strip-mode windows
    int bins_max;                                // half-spectrum bins one slot of the context's model arena holds
    long strip_floats;                           // per-CTA scratch of the strip-mode list (floats)
    int lut_floats;                              // what any_geo needs to size a window's shared memory
    KcfMeta *meta;
    int *jl_slot, *jl_frame, *jl_box, *jl_count; // [TDD_LISTS][S*cap] x 3, [TDD_LISTS]: live tracks grouped by kernel
    int *sp_slot, *sp_frame, *sp_box, *sp_count; // the same for the tracks spawned in this frame
    int *dropped;                                // [S] detections that could not spawn (no fused kernel for their window)
};

constexpr int TDD_LISTS = 13;                    // 9 fixed-size fused classes + 3 any-size lists by shared-memory bucket + the strip-mode list

// any-size list of a window that needs `floats` of shared memory: 0 -> four CTAs per SM, 1 -> two, 2 -> one (mot_capi.cu: any_launch_shape)
__host__ __device__ inline size_t tdd_any_list_bytes(int b) { return b == 0 ? (size_t)(227 * 1024) / 4 - 1024 : b == 1 ? (size_t)(227 * 1024) / 2 - 1024 : (size_t)ANY_SMEM_BUDGET_FLOATS * 4; }

__device__ __forceinline__ int fused_side(int cells) { return cells == 8 ? 0 : cells == 16 ? 1 : cells == 32 ? 2 : -1; }
__device__ __forceinline__ int fused_class_of(const TddState &st, int rows, int cols)
{
    if (rows > st.frame_h || cols > st.frame_w) return -1;
    const int hi = fused_side(rows / KCF_CELL), wi = fused_side(cols / KCF_CELL);
    if (hi < 0 || wi < 0) return -1;
    return st.cls_id[3 * hi + wi] >= 0 ? 3 * hi + wi : -1;
}
// job list of a window: a fixed-size fused class, else the any-size list its shared-memory plan falls into, else -1 (no tracker)
__device__ __forceinline__ int kcf_list_of(const TddState &st, int rows, int cols)
{
    if (rows > st.frame_h || cols > st.frame_w || rows / KCF_CELL < 2 || cols / KCF_CELL < 2) return -1;
    const int k = fused_class_of(st, rows, cols);
    if (k >= 0) return k;
    if (fused_side(rows / KCF_CELL) >= 0 && fused_side(cols / KCF_CELL) >= 0) return -1;      // a fixed-size class that was switched off
    const AnyGeo g = any_geo(rows / KCF_CELL, cols / KCF_CELL, st.lut_floats);
    if (!g.ok || g.S > st.bins_max) return -1;             // no fused kernel, or the model does not fit a slot of the arena
    if (g.strips) return (st.any_on[3] && 18L * g.os <= st.strip_floats) ? 12 : -1;
    const size_t bytes = (size_t)g.total * 4;
    const int b = bytes <= tdd_any_list_bytes(0) ? 0 : bytes <= tdd_any_list_bytes(1) ? 1 : 2;
    return st.any_on[b] ? 9 + b : -1;
}

// live tracks -> one job list per window class (order inside a list is irrelevant: jobs are independent)
// (host-array steps: the CTA of stream s first fetches the stream's detections from the pinned staging set into the device arrays
// that the later kernels of the frame read -- no copy-engine transfer in front of the frame)
__global__ void td_joblist_kernel(TddState st, const mot_bbox_t *host_dets, const int *host_ndet, mot_bbox_t *dets_w, int *ndet_w)
{
    const int s = blockIdx.x, T = st.ntracks[s];
    if (host_dets) {
        const int D0 = host_ndet[s];
        const long long *src = reinterpret_cast<const long long *>(host_dets + (long)s * st.max_det);
        long long *dst = reinterpret_cast<long long *>(dets_w + (long)s * st.max_det);
        for (int i = threadIdx.x; i < D0 * 3; i += blockDim.x) dst[i] = src[i];
        if (threadIdx.x == 0) ndet_w[s] = D0;
    }
    const long N = (long)st.S * st.cap;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const long e = (long)s * st.cap + i;
        const int slot = st.slot[e];
        const KcfMeta *m = st.meta + slot;
        const int k = m->size_class >= 1000 ? 9 + (m->size_class - 1000) : 3 * fused_side(m->hr) + fused_side(m->wc);
        const int at = atomicAdd(&st.jl_count[k], 1);
        st.jl_slot[k * N + at] = slot; st.jl_frame[k * N + at] = st.frame_base + s; st.jl_box[k * N + at] = (int)e;
    }
}

// scatter of the assignment + bookkeeping of assigned / unassigned tracks (top/td.cpp:472-502, 512-556)
__device__ __forceinline__ void td_scatter_cta(const TddState &st, const mot_bbox_t *dets, const int *ndet, int *at /* [1024] shared */)
{
    const int s = blockIdx.x, T = st.ntracks[s], D = ndet[s];
    for (int i = threadIdx.x; i < T; i += blockDim.x) at[i] = -1;
    for (int j = threadIdx.x; j < D; j += blockDim.x) st.assigned_detected[(long)s * st.max_det + j] = -1;
    __syncthreads();
    if (T && D) {
        const int *a = st.assign + (long)s * st.md;
        if (T < D) { for (int i = threadIdx.x; i < T; i += blockDim.x) { const int j = a[i]; at[i] = j; if (j >= 0) st.assigned_detected[(long)s * st.max_det + j] = i; } }
        else       { for (int j = threadIdx.x; j < D; j += blockDim.x) { const int i = a[j]; if (i >= 0) at[i] = j; st.assigned_detected[(long)s * st.max_det + j] = i; } }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const long o = (long)s * st.cap + i;
        const int j = at[i];
        if (j >= 0) { st.bbox[o] = dets[(long)s * st.max_det + j]; st.vis[o]++; st.age[o]++; st.invis[o] = 0; }     // :541-546
        else        { st.age[o]++; st.invis[o]++; }                                                               // :555-556
    }
}

__global__ void td_scatter_kernel(TddState st, const mot_bbox_t *dets, const int *ndet)
{
    __shared__ int at[1024];
    td_scatter_cta(st, dets, ndet, at);
}

// Exclusive scan of 0/1 flags v[0..n), n <= 1024, by a CTA of 256..1024 threads (four consecutive entries per thread, warp shuffles,
// one pass over the warp totals): on return v[i] = number of set flags before i; returns the number of set flags.
__device__ int block_scan_flags(int *v, int n, int *wsum /* [33] shared */)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int a[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int i = 4 * tid + q; a[q] = i < n ? v[i] : 0; }
    const int e1 = a[0], e2 = e1 + a[1], e3 = e2 + a[2], sum = e3 + a[3];
    int inc = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, inc, off); if (lane >= off) inc += t; }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        const int w = lane < nw ? wsum[lane] : 0;
        int winc = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, winc, off); if (lane >= off) winc += t; }
        if (lane < nw) wsum[lane] = winc - w;
        if (lane == 31) wsum[32] = winc;
    }
    __syncthreads();
    const int base = wsum[warp] + inc - sum;
    const int ex[4] = { 0, e1, e2, e3 };
#pragma unroll
    for (int q = 0; q < 4; ++q) { const int i = 4 * tid + q; if (i < n) v[i] = base + ex[q]; }
    __syncthreads();
    return wsum[32];
}

// delete lost tracks with stable compaction (top/td.cpp:585-609), then spawn (top/td.cpp:612-644)
__device__ __forceinline__ void td_lifecycle_cta(const TddState &st, const KalmanState &kal, const mot_bbox_t *dets, const int *ndet, int *sm /* [3072] shared */)
{
    const int s = blockIdx.x, T = st.ntracks[s], D = ndet[s], cap = st.cap, tid = threadIdx.x, NTH = blockDim.x;
    int *pos = sm;                    // [cap] new position of a kept track / [max_det] rank of a spawning detection
    int *freeid = sm + 1024;          // [cap] free local slot ids in ascending order
    int *flag = sm + 2048;            // [1024] scratch flags of the scans
    __shared__ int wsum[33];
    // ---- keep flags and their exclusive scan -------------------------------------------------------------------------------
    for (int i = tid; i < T; i += NTH) {
        const long o = (long)s * cap + i;
        const bool lost = ((st.age[o] < 10) && (st.vis[o] * 5 < 3 * st.age[o])) || (st.invis[o] >= 20);        // :587-590
        flag[i] = lost ? 0 : 1; pos[i] = lost ? -1 : 0;
    }
    __syncthreads();
    const int n2 = block_scan_flags(flag, T, wsum);
    for (int i = tid; i < T; i += NTH) if (pos[i] == 0) pos[i] = flag[i];
    __syncthreads();
    // stable compaction through registers
    uint32_t r_tid[4]; int r_slot[4], r_age[4], r_vis[4], r_inv[4]; mot_bbox_t r_box[4]; int r_pos[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = tid + q * NTH; r_pos[q] = -1;
        if (i < T) { const long o = (long)s * cap + i; r_pos[q] = pos[i]; r_tid[q] = st.tid[o]; r_slot[q] = st.slot[o]; r_age[q] = st.age[o]; r_vis[q] = st.vis[o]; r_inv[q] = st.invis[o]; r_box[q] = st.bbox[o]; }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (r_pos[q] >= 0) { const long o = (long)s * cap + r_pos[q]; st.tid[o] = r_tid[q]; st.slot[o] = r_slot[q]; st.age[o] = r_age[q]; st.vis[o] = r_vis[q]; st.invis[o] = r_inv[q]; st.bbox[o] = r_box[q]; }
    }
    for (int i = n2 + tid; i < T; i += NTH) st.slot[(long)s * cap + i] = -1;
    __syncthreads();
    // ---- free local slot ids (ascending) + ranks of the spawning detections ------------------------------------------------
    __shared__ uint32_t used[32];                       // cap <= 1024
    if (tid < 32) used[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n2; i += NTH) { const int ls = st.slot[(long)s * cap + i] - s * cap; atomicOr(&used[ls >> 5], 1u << (ls & 31)); }
    for (int j = tid; j < D; j += NTH) {
        bool cand = st.assigned_detected[(long)s * st.max_det + j] < 0;
        if (cand && st.kcf) {
            const mot_bbox_t b = dets[(long)s * st.max_det + j];
            if (kcf_list_of(st, b.b - b.t + 1, b.r - b.l + 1) < 0) { cand = false; atomicAdd(&st.dropped[s], 1); }
        }
        pos[j] = cand ? 1 : 0;
    }
    __syncthreads();
    // free local slot ids in ascending order
    for (int ls = tid; ls < cap; ls += NTH) flag[ls] = ((used[ls >> 5] >> (ls & 31)) & 1u) ? 0 : 1;
    __syncthreads();
    {
        int isfree[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int ls = tid + q * NTH; isfree[q] = ls < cap ? flag[ls] : 0; }
        __syncthreads();
        block_scan_flags(flag, cap, wsum);
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int ls = tid + q * NTH; if (isfree[q]) freeid[flag[ls]] = ls; }
    }
    __syncthreads();
    // ranks of the spawning detections (ascending detection order), cut off where the table is full
    for (int j = tid; j < D; j += NTH) flag[j] = pos[j];
    __syncthreads();
    const int cand = block_scan_flags(flag, D, wsum);
    const int room = cap - n2;
    for (int j = tid; j < D; j += NTH) pos[j] = (pos[j] && flag[j] < room) ? flag[j] : -1;
    __syncthreads();
    const int spawned = cand < room ? cand : (room > 0 ? room : 0);
    const uint32_t id0 = st.tracker_id[s];
    for (int j = tid; j < D; j += NTH) {
        const int rk = pos[j];
        if (rk < 0) continue;
        const long o = (long)s * cap + n2 + rk;
        const mot_bbox_t b = dets[(long)s * st.max_det + j];
        const int sl = s * cap + freeid[rk];
        st.tid[o] = id0 + (uint32_t)rk; st.slot[o] = sl; st.age[o] = 0; st.vis[o] = 0; st.invis[o] = 0; st.bbox[o] = b;      // :618-627
        if (st.kcf) {
            // tracker_new, trackers/kcf.cpp:484-491 + kcf_init :139-213: sizes frozen, pos = box, scales 1, first_update set
            KcfMeta m{};
            m.rows = b.b - b.t + 1; m.cols = b.r - b.l + 1; m.hr = m.rows / KCF_CELL; m.wc = m.cols / KCF_CELL;
            m.pos = b; m.scale_horiz = 1.0f; m.scale_vert = 1.0f; m.first_update = 1;
            const int k = kcf_list_of(st, m.rows, m.cols);
            m.size_class = k < 9 ? st.cls_id[k] : 1000 + (k - 9);        // any-size tracks remember their list, not a class
            m.model_ptr = nullptr; m.alpha_ptr = nullptr;
            st.meta[sl] = m;
            const long N = (long)st.S * cap;
            const int at = atomicAdd(&st.sp_count[k], 1);
            st.sp_slot[k * N + at] = sl; st.sp_frame[k * N + at] = st.frame_base + s; st.sp_box[k * N + at] = (int)o;
            continue;
        }
        // tracker_new, trackers/kalman.cpp:147-163: x0 = [l,t,r,b,0,0], P0 = 1e4 I
        const double x0[6] = { (double)b.l, (double)b.t, (double)b.r, (double)b.b, 0.0, 0.0 };
        for (int k = 0; k < 6; ++k) kal.x[(long)k * kal.cap + sl] = x0[k];
        for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) kal.P[(long)(c * 6 + r) * kal.cap + sl] = (r == c) ? 1e4 : 0.0;
    }
    __syncthreads();
    if (tid == 0) { st.ntracks[s] = n2 + spawned; st.tracker_id[s] = id0 + (uint32_t)spawned; }
}

__global__ void td_lifecycle_kernel(TddState st, KalmanState kal, const mot_bbox_t *dets, const int *ndet)
{
    extern __shared__ int sm[];
    td_lifecycle_cta(st, kal, dets, ndet, sm);
}

// ---- the Kalman kind's whole frame in ONE launch: a CTA per stream runs predict (+clamp), the cost matrix, the Munkres solver, the
// scatter + bookkeeping, the update and the lifecycle back to back (the phases of one stream only depend on each other, and a few
// hundred tracks are a CTA's worth of work), so a frame costs one launch latency instead of six.  Same device code as the separate
// kernels (kalman.cuh, assoc.cuh, the *_cta functions above): the results are the same bit for bit.
constexpr int TDF_THREADS = 256;      // 255 registers per thread: the Kalman update keeps its 6x6 FP64 algebra in registers (no spills)
struct TdFrameArgs {
    double *dist, *work, *cost; int cost_mode; double screen_dis; int mat_doubles;
    // host-array steps: the detections are read by the kernel itself from the pinned staging set (device-accessible under unified
    // addressing) into the device arrays it then works on -- no copy-engine transfers in front of a one-launch frame (null: device arrays given)
    const mot_bbox_t *host_dets; const int *host_ndet; mot_bbox_t *dets_w; int *ndet_w;
};

__global__ void __launch_bounds__(TDF_THREADS, 1) td_frame_kalman_kernel(TddState st, KalmanState kal, const mot_bbox_t *dets, const int *ndet, TdFrameArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int at[1024];
    const int s = blockIdx.x, tid = threadIdx.x, cap = st.cap, md = st.md;
    if (a.host_dets) {
        // this stream's detections: pinned host memory -> the device arrays (dets == a.dets_w, ndet == a.ndet_w), 8 bytes per thread and trip
        const int D0 = a.host_ndet[s];
        const long long *src = reinterpret_cast<const long long *>(a.host_dets + (long)s * st.max_det);
        long long *dst = reinterpret_cast<long long *>(a.dets_w + (long)s * st.max_det);
        for (int i = tid; i < D0 * 3; i += TDF_THREADS) dst[i] = src[i];
        if (tid == 0) a.ndet_w[s] = D0;
        __syncthreads();
    }
    const int T = st.ntracks[s], D = ndet[s];
    mot_bbox_t *const trk = st.bbox + (long)s * cap;
    const mot_bbox_t *const det = dets + (long)s * st.max_det;
    // predict + clamp (top/td.cpp:344-384)
    for (int i = tid; i < cap; i += TDF_THREADS) { const int sl = st.slot[(long)s * cap + i]; if (sl >= 0) kalman_predict_one(kal, sl, trk + i, 1, st.frame_w, st.frame_h); }
    __syncthreads();
    // cost matrix (top/td.cpp:386-457), column-major with rows = the smaller side.  The boxes are staged in shared memory (the
    // solver's area, not yet in use) and a thread keeps ONE row, so a cell costs no division and no global load, and four cells
    // are in flight per thread (the FP64 square root is a long dependent chain).
    double *const dist = a.dist + (long)s * md * md;
    {
        mot_bbox_t *const sb = reinterpret_cast<mot_bbox_t *>(smem_raw);            // [T] track boxes, then [D] detections: <= 2 * 256 * 24 bytes
        for (int i = tid; i < T; i += TDF_THREADS) sb[i] = trk[i];
        for (int j = tid; j < D; j += TDF_THREADS) sb[T + j] = det[j];
        __syncthreads();
        const int nR = T < D ? T : D, nC = T < D ? D : T;
        const mot_bbox_t *const rowb = T < D ? sb : sb + T, *const colb = T < D ? sb + T : sb;      // rows = trackers when T < D, else detections
        if (nR > 0) {                                                                                  // nR <= md <= TDF_THREADS (checked at creation)
            const int G = TDF_THREADS / nR, r = tid % nR, g = tid / nR;                                 // column groups when there are threads to spare
            if (g < G) {
                const mot_bbox_t rb = rowb[r];
                for (int c0 = g; c0 < nC; c0 += 4 * G) {
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int c = c0 + u * G; if (c < nC) v[u] = T < D ? cost_cell(rb, colb[c], a.cost_mode, a.screen_dis) : cost_cell(colb[c], rb, a.cost_mode, a.screen_dis); }
#pragma unroll
                    for (int u = 0; u < 4; ++u) { const int c = c0 + u * G; if (c < nC) dist[r + (long)nR * c] = v[u]; }
                }
            }
        }
    }
    __syncthreads();
    // assignment (trackers/hungarian/hungarian.cpp:29-368)
    munkres_cta<TDF_THREADS>(dist, a.work + (long)s * md * md, T < D ? T : D, T < D ? D : T, md, st.assign + (long)s * md, a.cost + s, smem_raw, a.mat_doubles);
    __syncthreads();
    td_scatter_cta(st, dets, ndet, at);
    __syncthreads();
    // update with the assigned detection, or with the predicted box itself (top/td.cpp:539, 581)
    // (one thread per track here: the update is a chain of ~90 dependent FP64 operations whatever the number of lanes that share it, so
    // the eight-lane form of kalman.cuh, which needs two passes over 64 tracks in a 256-thread CTA, is slower in this kernel -- measured)
    for (int i = tid; i < cap; i += TDF_THREADS) { const int sl = st.slot[(long)s * cap + i]; if (sl >= 0) kalman_update_one(kal, sl, trk[i]); }
    __syncthreads();
    td_lifecycle_cta(st, kal, dets, ndet, reinterpret_cast<int *>(smem_raw));
}

}  // namespace mot

using namespace mot;

struct mot_tdd_s {
    mot_ctx_t *ctx;
    TddState st{};
    int cost_mode;
    double *d_dist = nullptr, *d_cost = nullptr, *d_work = nullptr;      // cost matrices, totals, the solver's working copy (owned: graph-safe)
    bool fused_frame = false;                                            // Kalman kind: the whole frame in one launch (td_frame_kalman_kernel)
    const mot_bbox_t *host_dets_next = nullptr; const int *host_ndet_next = nullptr;   // set by mot_tdd_step for the launch it makes next
    float *d_strip = nullptr; int strip_ctas = 0;                        // strip-mode list: per-CTA histogram scratch (L2-resident)
    // Host-array steps: detections are staged in pinned memory and uploaded.  Two staging sets used alternately, each with its own
    // "consumed" event, so that the host prepares step k+1 while step k runs and only ever waits for step k-1.
    DevBuf<mot_bbox_t> d_dets[2]; DevBuf<int> d_ndet[2];
    PinBuf<mot_bbox_t> h_dets[2]; PinBuf<int> h_ndet[2];
    cudaEvent_t consumed[2] = { nullptr, nullptr };
    int flip = 0;
    // The host-array step of the Kalman kind is launch-latency bound (two copies + six kernels for a few hundred tracks): its fixed
    // sequence is captured once per staging set into a CUDA graph and replayed.
    cudaGraphExec_t graph[2] = { nullptr, nullptr };
    cudaStream_t graph_stream = nullptr;
};

extern "C" {

static int tdd_step_kcf(mot_tdd_t *t, const mot_bbox_t *d_dets, const int *d_ndet);

static void tdd_release(mot_tdd_t *t)
{
    TddState &st = t->st;
    cudaFree(st.ntracks); cudaFree(st.tracker_id); cudaFree(st.tid); cudaFree(st.slot); cudaFree(st.age); cudaFree(st.vis); cudaFree(st.invis);
    cudaFree(st.bbox); cudaFree(st.assign); cudaFree(st.assigned_detected); cudaFree(t->d_dist); cudaFree(t->d_cost); cudaFree(t->d_work);
    cudaFree(st.jl_slot); cudaFree(st.jl_frame); cudaFree(st.jl_box); cudaFree(st.sp_slot); cudaFree(st.sp_frame); cudaFree(st.sp_box);
    cudaFree(st.jl_count); cudaFree(st.dropped); cudaFree(t->d_strip);
    for (int q = 0; q < 2; ++q) {
        if (t->graph[q]) cudaGraphExecDestroy(t->graph[q]);
        if (t->consumed[q]) cudaEventDestroy(t->consumed[q]);
        t->d_dets[q].release(); t->d_ndet[q].release(); t->h_dets[q].release(); t->h_ndet[q].release();
    }
}

// every allocation of a loop object; on failure the caller releases whatever was obtained (all pointers start out null)
static int tdd_alloc(mot_tdd_t *t, mot_ctx_t *c, int n_streams, int cap, int max_det, bool kcf)
{
    TddState &st = t->st;
    st.S = n_streams; st.cap = cap; st.max_det = max_det; st.md = cap > max_det ? cap : max_det;
    const size_t n = (size_t)n_streams * cap;
    CU(cudaMalloc(&st.ntracks, sizeof(int) * n_streams)); CU(cudaMalloc(&st.tracker_id, sizeof(uint32_t) * n_streams));
    CU(cudaMalloc(&st.tid, sizeof(uint32_t) * n)); CU(cudaMalloc(&st.slot, sizeof(int) * n)); CU(cudaMalloc(&st.age, sizeof(int) * n));
    CU(cudaMalloc(&st.vis, sizeof(int) * n)); CU(cudaMalloc(&st.invis, sizeof(int) * n)); CU(cudaMalloc(&st.bbox, sizeof(mot_bbox_t) * n));
    CU(cudaMalloc(&st.assign, sizeof(int) * (size_t)n_streams * st.md)); CU(cudaMalloc(&st.assigned_detected, sizeof(int) * (size_t)n_streams * max_det));
    CU(cudaMalloc(&t->d_dist, sizeof(double) * (size_t)n_streams * st.md * st.md)); CU(cudaMalloc(&t->d_cost, sizeof(double) * n_streams));
    CU(cudaMalloc(&t->d_work, sizeof(double) * (size_t)n_streams * st.md * st.md));
    CU(cudaMemsetAsync(st.ntracks, 0, sizeof(int) * n_streams, c->stream)); CU(cudaMemsetAsync(st.tracker_id, 0, sizeof(uint32_t) * n_streams, c->stream));
    CU(cudaMemsetAsync(st.slot, 0xFF, sizeof(int) * n, c->stream));
    CU(cudaMemsetAsync(st.age, 0, sizeof(int) * n, c->stream)); CU(cudaMemsetAsync(st.vis, 0, sizeof(int) * n, c->stream)); CU(cudaMemsetAsync(st.invis, 0, sizeof(int) * n, c->stream));
    CU(cudaMemsetAsync(st.bbox, 0, sizeof(mot_bbox_t) * n, c->stream)); CU(cudaMemsetAsync(st.tid, 0, sizeof(uint32_t) * n, c->stream));
    st.kcf = kcf ? 1 : 0; st.frame_w = c->W; st.frame_h = c->H; st.meta = c->d_meta; st.frame_base = 0;
    for (int k = 0; k < 9; ++k) st.cls_id[k] = -1;
    for (int b = 0; b < 4; ++b) st.any_on[b] = kcf ? 1 : 0;
    st.lut_floats = c->lut_floats; st.bins_max = (int)c->alpha_stride; st.strip_floats = 0;
    if (kcf) {
        static const int side[3] = { 8, 16, 32 };
        for (int hi = 0; hi < 3; ++hi)
            for (int wi = 0; wi < 3; ++wi) { const int rc = mot_ctx_kcf_class(c, side[hi], side[wi], &st.cls_id[3 * hi + wi]); if (rc) return rc; }
        CU(cudaMalloc(&st.jl_slot, sizeof(int) * TDD_LISTS * n)); CU(cudaMalloc(&st.jl_frame, sizeof(int) * TDD_LISTS * n)); CU(cudaMalloc(&st.jl_box, sizeof(int) * TDD_LISTS * n));
        CU(cudaMalloc(&st.sp_slot, sizeof(int) * TDD_LISTS * n)); CU(cudaMalloc(&st.sp_frame, sizeof(int) * TDD_LISTS * n)); CU(cudaMalloc(&st.sp_box, sizeof(int) * TDD_LISTS * n));
        CU(cudaMalloc(&st.jl_count, sizeof(int) * 2 * TDD_LISTS)); st.sp_count = st.jl_count + TDD_LISTS;
        CU(cudaMalloc(&st.dropped, sizeof(int) * n_streams));
        // strip-mode windows: 18 orientation planes of the cell grid per resident CTA (one CTA per SM); a window has fewer than
        // 2 * bins cells, and the any-size kernel ends at about 8200
        t->strip_ctas = c->sm_count;
        st.strip_floats = 18L * ((std::min(2 * st.bins_max, 8448) + 31) & ~31);
        CU(cudaMalloc(&t->d_strip, sizeof(float) * st.strip_floats * t->strip_ctas));
        CU(cudaMemsetAsync(st.dropped, 0, sizeof(int) * n_streams, c->stream));
    }
    return 0;
}

int mot_tdd_create(mot_tdd_t **out, mot_ctx_t *c, int n_streams, int cap, int max_det, int cost_mode)
{
    if (!out || !c || n_streams <= 0 || cap <= 0 || cap > 1024 || max_det <= 0 || max_det > 1024) return mot_fail(MOT_ERR_ARG, "mot_tdd_create: bad argument (cap and max_det must be in 1..1024)");
    const bool kcf = c->kind == MOT_TRACKER_KCF;
    if (kcf && c->ext_on) return mot_fail(MOT_ERR_ARG, "the device-resident loop runs the reference's filter only; use mot_td_step with the KCF extensions");
    if (kcf && c->n_frames < n_streams) return mot_fail(MOT_ERR_ARG, "the KCF frame loop reads stream s from frame slot s: the context has %d frame slots, %d are needed", c->n_frames, n_streams);
    if ((long)n_streams * cap > c->max_tracks) return mot_fail(MOT_ERR_CAPACITY, "context has %d track slots, %d streams x %d are needed", c->max_tracks, n_streams, cap);
    for (char u : c->used) if (u) return mot_fail(MOT_ERR_ARG, "the context already holds host-managed trackers");
    CU(cudaSetDevice(c->device));
    mot_tdd_t *t = new mot_tdd_s();
    t->ctx = c; t->cost_mode = cost_mode;
    const int rc = tdd_alloc(t, c, n_streams, cap, max_det, kcf);
    if (rc) { tdd_release(t); delete t; return rc; }
    if (!kcf && t->st.md <= TDF_THREADS && !getenv("MOT_TDD_UNFUSED")) {      // larger problems want the 1024-thread solver: separate launches
        const size_t bytes = std::max(munkres_smem_bytes(t->st.md, munkres_mat_doubles(t->st.md)), sizeof(int) * 3072);
        t->fused_frame = cudaFuncSetAttribute((const void *)td_frame_kalman_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
    }
    for (long i = 0; i < (long)n_streams * cap; ++i) c->used[i] = 1;        // these slots now belong to the device-side tables
    c->free_slots.erase(std::remove_if(c->free_slots.begin(), c->free_slots.end(), [&](int s) { return s < n_streams * cap; }), c->free_slots.end());
    *out = t;
    return 0;
}

void mot_tdd_destroy(mot_tdd_t *t)
{
    if (!t) return;
    cudaSetDevice(t->ctx->device); cudaStreamSynchronize(t->ctx->stream);
    const TddState &st = t->st;
    for (long i = 0; i < (long)st.S * st.cap; ++i) { t->ctx->used[i] = 0; t->ctx->free_slots.push_back((int)i); }
    tdd_release(t);
    delete t;
}

// KCF kind: job lists -> predict per class -> association -> scatter -> update per class -> lifecycle -> first update of the spawned
static int tdd_step_kcf(mot_tdd_t *t, const mot_bbox_t *d_dets, const int *d_ndet)
{
    mot_ctx_t *c = t->ctx; TddState &st = t->st;
    const int n = st.S * st.cap;
    int rc = mot_ctx_frames_ready(c); if (rc) return rc;
    CU(cudaMemsetAsync(st.jl_count, 0, sizeof(int) * 2 * TDD_LISTS, c->stream));
    td_joblist_kernel<<<st.S, 256, 0, c->stream>>>(st, t->host_dets_next, t->host_ndet_next, const_cast<mot_bbox_t *>(d_dets), const_cast<int *>(d_ndet));
    t->host_dets_next = nullptr; t->host_ndet_next = nullptr;
    auto per_class = [&](int mode, const int *count, const int *slot, const int *frame, const int *box, int clamp) {
        for (int k = 0; k < 9; ++k) {
            if (st.cls_id[k] < 0) continue;
            const int r = mot_ctx_kcf_launch(c, mode, st.cls_id[k], n, count + k, slot + (long)k * n, frame + (long)k * n, st.bbox, box + (long)k * n, clamp);
            if (r) return r;
        }
        for (int b = 0; b < 3; ++b) {
            if (!st.any_on[b]) continue;
            const int k = 9 + b;
            const int r = mot_ctx_kcf_launch_any(c, mode, tdd_any_list_bytes(b), n, count + k, slot + (long)k * n, frame + (long)k * n, st.bbox, box + (long)k * n, clamp, nullptr, 0, 0);
            if (r) return r;
        }
        if (st.any_on[3]) {
            const int k = 12;
            const int r = mot_ctx_kcf_launch_any(c, mode, tdd_any_list_bytes(2), n, count + k, slot + (long)k * n, frame + (long)k * n, st.bbox, box + (long)k * n, clamp,
                                                 t->d_strip, st.strip_floats, t->strip_ctas);
            if (r) return r;
        }
        return 0;
    };
    rc = per_class(KCF_MODE_PREDICT, st.jl_count, st.jl_slot, st.jl_frame, st.jl_box, 1); if (rc) return rc;       // top/td.cpp:344-384
    rc = mot_ctx_associate_dev(c, st.S, st.ntracks, d_ndet, st.bbox, st.cap, d_dets, st.max_det, t->cost_mode,
                               t->d_dist, (long)st.md * st.md, st.assign, st.md, t->d_cost, st.md, t->d_work);
    if (rc) return rc;
    td_scatter_kernel<<<st.S, 256, 0, c->stream>>>(st, d_dets, d_ndet);
    rc = per_class(KCF_MODE_UPDATE, st.jl_count, st.jl_slot, st.jl_frame, st.jl_box, 0); if (rc) return rc;        // top/td.cpp:512-582
    td_lifecycle_kernel<<<st.S, 256, sizeof(int) * 3072, c->stream>>>(st, c->kal, d_dets, d_ndet);
    rc = per_class(KCF_MODE_UPDATE, st.sp_count, st.sp_slot, st.sp_frame, st.sp_box, 0); if (rc) return rc;        // top/td.cpp:629-641
    CU(cudaGetLastError());
    c->launches += 3;           // job lists, scatter, lifecycle (the fused launches and the association count themselves)
    return 0;
}

/* detections: device arrays dets[S][max_det], ndet[S]; everything is enqueued on the context stream, nothing synchronises */
int mot_tdd_step_dev(mot_tdd_t *t, const mot_bbox_t *d_dets, const int *d_ndet)
{
    if (!t || !d_dets || !d_ndet) return mot_fail(MOT_ERR_ARG, "mot_tdd_step_dev: null argument");
    mot_ctx_t *c = t->ctx; TddState &st = t->st;
    CU(cudaSetDevice(c->device));
    const int n = st.S * st.cap;
    if (st.kcf) return tdd_step_kcf(t, d_dets, d_ndet);
    if (t->fused_frame) {
        TdFrameArgs a{ t->d_dist, t->d_work, t->d_cost, t->cost_mode, 1.0 / (double)c->W, munkres_mat_doubles(st.md), t->host_dets_next, t->host_ndet_next,
                       const_cast<mot_bbox_t *>(d_dets), const_cast<int *>(d_ndet) };
        t->host_dets_next = nullptr; t->host_ndet_next = nullptr;
        const size_t bytes = std::max(munkres_smem_bytes(st.md, a.mat_doubles), sizeof(int) * 3072);
        td_frame_kalman_kernel<<<st.S, TDF_THREADS, bytes, c->stream>>>(st, c->kal, d_dets, d_ndet, a);
        CU(cudaGetLastError());
        c->launches += 1;
        return 0;
    }
    int rc = kalman_predict(c->kal, n, st.slot, st.bbox, 1, c->W, c->H, c->stream);
    if (rc) return mot_fail(MOT_ERR_CUDA, "kalman_predict launch failed (%d)", rc);
    rc = mot_ctx_associate_dev(c, st.S, st.ntracks, d_ndet, st.bbox, st.cap, d_dets, st.max_det, t->cost_mode,
                               t->d_dist, (long)st.md * st.md, st.assign, st.md, t->d_cost, st.md, t->d_work);
    if (rc) return rc;
    td_scatter_kernel<<<st.S, 256, 0, c->stream>>>(st, d_dets, d_ndet);
    rc = kalman_update(c->kal, n, st.slot, st.bbox, c->stream);
    if (rc) return mot_fail(MOT_ERR_CUDA, "kalman_update launch failed (%d)", rc);
    td_lifecycle_kernel<<<st.S, 256, sizeof(int) * 3072, c->stream>>>(st, c->kal, d_dets, d_ndet);
    CU(cudaGetLastError());
    c->launches += 4;           // predict, scatter, update, lifecycle (+2 counted by the association call)
    return 0;
}

/* host-array convenience: detections of every stream are staged and uploaded, then mot_tdd_step_dev; asynchronous */
int mot_tdd_step(mot_tdd_t *t, const mot_bbox_t *const *dets, const int *ndet)
{
    if (!t || !dets || !ndet) return mot_fail(MOT_ERR_ARG, "mot_tdd_step: null argument");
    mot_ctx_t *c = t->ctx; TddState &st = t->st;
    CU(cudaSetDevice(c->device));
    for (int s = 0; s < st.S; ++s)
        if (ndet[s] < 0 || ndet[s] > st.max_det) return mot_fail(MOT_ERR_ARG, "stream %d has %d detections (max %d)", s, ndet[s], st.max_det);
    const int q = t->flip; t->flip ^= 1;
    const bool fresh = t->h_dets[q].n < (size_t)st.S * st.max_det;      // first use of this staging set: its graph (if any) is stale
    CU(t->h_dets[q].ensure((size_t)st.S * st.max_det)); CU(t->d_dets[q].ensure((size_t)st.S * st.max_det)); CU(t->h_ndet[q].ensure(st.S)); CU(t->d_ndet[q].ensure(st.S));
    if (!t->consumed[q]) CU(cudaEventCreateWithFlags(&t->consumed[q], cudaEventDisableTiming));
    else CU(cudaEventSynchronize(t->consumed[q]));   // the step that last used this staging set (two steps ago) has read it
    for (int s = 0; s < st.S; ++s) {
        t->h_ndet[q].p[s] = ndet[s];
        if (ndet[s]) memcpy(t->h_dets[q].p + (size_t)s * st.max_det, dets[s], sizeof(mot_bbox_t) * ndet[s]);
    }
    const size_t det_bytes = sizeof(mot_bbox_t) * (size_t)st.S * st.max_det;
    if (st.kcf) {
        // plain launches (the sequence waits for frame uploads recorded on another stream); the first kernel of the frame fetches the
        // detections from the pinned staging set itself
        t->host_dets_next = t->h_dets[q].p; t->host_ndet_next = t->h_ndet[q].p;
        const int rc = mot_tdd_step_dev(t, t->d_dets[q].p, t->d_ndet[q].p);
        t->host_dets_next = nullptr; t->host_ndet_next = nullptr;      // (consumed by the launch; cleared here too in case the step failed before it)
        if (rc) return rc;
        CU(cudaEventRecord(t->consumed[q], c->stream));
        return 0;
    }
    if (t->fused_frame) {
        // one launch, and the kernel fetches the detections from the pinned staging set itself
        t->host_dets_next = t->h_dets[q].p; t->host_ndet_next = t->h_ndet[q].p;
        const int rc = mot_tdd_step_dev(t, t->d_dets[q].p, t->d_ndet[q].p);
        t->host_dets_next = nullptr; t->host_ndet_next = nullptr;      // (consumed by the launch; cleared here too in case the step failed before it)
        if (rc) return rc;
        CU(cudaEventRecord(t->consumed[q], c->stream));
        return 0;
    }
    if (t->graph_stream != c->stream || fresh) {
        for (int k = 0; k < 2; ++k) if ((t->graph_stream != c->stream || k == q) && t->graph[k]) { cudaGraphExecDestroy(t->graph[k]); t->graph[k] = nullptr; }
        t->graph_stream = c->stream;
    }
    if (!t->graph[q]) {
        // first call with this staging set (or the stream changed): record the sequence.  Every buffer the captured kernels touch
        // belongs to this loop object (tables, cost matrices, the solver's working copy, the staging set) or to the context's fixed
        // Kalman state: nothing a later call on the context can reallocate.
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        cudaError_t e1 = cudaMemcpyAsync(t->d_dets[q].p, t->h_dets[q].p, det_bytes, cudaMemcpyHostToDevice, c->stream);
        cudaError_t e2 = cudaMemcpyAsync(t->d_ndet[q].p, t->h_ndet[q].p, sizeof(int) * st.S, cudaMemcpyHostToDevice, c->stream);
        const long l0 = c->launches;
        const int rc = (e1 == cudaSuccess && e2 == cudaSuccess) ? mot_tdd_step_dev(t, t->d_dets[q].p, t->d_ndet[q].p) : MOT_ERR_CUDA;
        c->launches = l0;
        const cudaError_t e3 = cudaStreamEndCapture(c->stream, &g);
        if (rc || e3 != cudaSuccess || !g) { if (g) cudaGraphDestroy(g); return rc ? rc : mot_fail(MOT_ERR_CUDA, "graph capture of the frame loop failed: %s", cudaGetErrorString(e3)); }
        const cudaError_t e4 = cudaGraphInstantiate(&t->graph[q], g, 0);
        cudaGraphDestroy(g);
        if (e4 != cudaSuccess) { t->graph[q] = nullptr; return mot_fail(MOT_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e4)); }
    }
    CU(cudaGraphLaunch(t->graph[q], c->stream));
    CU(cudaEventRecord(t->consumed[q], c->stream));
    c->launches += t->fused_frame ? 1 : 6;
    return 0;
}

/* the same with one bbox_chain_t per stream, the detector's wire format (top/cnntype.h:43-47; filled by tensorRunB, top/td.cpp:204) */
int mot_tdd_step_chains(mot_tdd_t *t, const mot_bbox_chain_t *const *chains)
{
    if (!t || !chains) return mot_fail(MOT_ERR_ARG, "mot_tdd_step_chains: null argument");
    const int S = t->st.S;
    std::vector<const mot_bbox_t *> dets(S);
    std::vector<int> nd(S);
    for (int s = 0; s < S; ++s) {
        if (!chains[s]) { dets[s] = nullptr; nd[s] = 0; continue; }
        if (chains[s]->nbox < 0 || chains[s]->nbox > 128) return mot_fail(MOT_ERR_ARG, "stream %d: chain with %d boxes (0..128)", s, chains[s]->nbox);
        dets[s] = chains[s]->bbox; nd[s] = chains[s]->nbox;
    }
    return mot_tdd_step(t, dets.data(), nd.data());
}

/* KCF kind: from the next step on, stream s reads frame slot base + s.  The steps are asynchronous, so a slot must not be
 * uploaded again while a step that reads it is still in flight: alternate two bases (2 * n_streams slots), or mot_sync first. */
int mot_tdd_frame_base(mot_tdd_t *t, int base)
{
    if (!t || base < 0) return mot_fail(MOT_ERR_ARG, "mot_tdd_frame_base: bad argument");
    if (!t->st.kcf) return mot_fail(MOT_ERR_KIND, "not a KCF frame loop");
    if (base + t->st.S > t->ctx->n_frames) return mot_fail(MOT_ERR_ARG, "frame slots %d..%d do not exist (the context has %d)", base, base + t->st.S - 1, t->ctx->n_frames);
    t->st.frame_base = base;
    return 0;
}

/* KCF kind: restrict the loop to the listed window sizes (pixels): the kernels of all other job lists are not launched, and
 * detections of any other kind of size are counted as dropped */
int mot_tdd_kcf_windows(mot_tdd_t *t, int n, const int *rows, const int *cols)
{
    if (!t || n < 0 || (n && (!rows || !cols))) return mot_fail(MOT_ERR_ARG, "mot_tdd_kcf_windows: bad argument");
    if (!t->st.kcf) return mot_fail(MOT_ERR_KIND, "not a KCF frame loop");
    static const int side[3] = { 8, 16, 32 };
    int keep[9] = { 0 }, any_keep[4] = { 0 };
    for (int i = 0; i < n; ++i) {
        const int hr = rows[i] / KCF_CELL, wc = cols[i] / KCF_CELL;
        int hi = -1, wi = -1;
        for (int q = 0; q < 3; ++q) { if (hr == side[q]) hi = q; if (wc == side[q]) wi = q; }
        if (hi >= 0 && wi >= 0) { keep[3 * hi + wi] = 1; continue; }
        const AnyGeo g = (hr >= 2 && wc >= 2) ? any_geo(hr, wc, t->st.lut_floats) : AnyGeo{};
        if (hr < 2 || wc < 2 || !g.ok || g.S > t->st.bins_max || (g.strips && 18L * g.os > t->st.strip_floats))
            return mot_fail(MOT_ERR_SHAPE, "window %dx%d px: no fused kernel holds it (2x2 cells up to about 8000 cells) or its %d spectrum bins exceed the slot size %d (mot_ctx_reserve_window)", rows[i], cols[i], g.S, t->st.bins_max);
        if (g.strips) { any_keep[3] = 1; continue; }
        const size_t bytes = (size_t)g.total * 4;
        any_keep[bytes <= tdd_any_list_bytes(0) ? 0 : bytes <= tdd_any_list_bytes(1) ? 1 : 2] = 1;
    }
    for (int hi = 0; hi < 3; ++hi)
        for (int wi = 0; wi < 3; ++wi) {
            int &id = t->st.cls_id[3 * hi + wi];
            if (!keep[3 * hi + wi]) id = -1;
            else if (id < 0) { const int rc = mot_ctx_kcf_class(t->ctx, side[hi], side[wi], &id); if (rc) return rc; }
        }
    for (int b = 0; b < 4; ++b) t->st.any_on[b] = any_keep[b];
    return 0;
}

/* KCF kind: detections of stream s that could not spawn a track so far (window without a fused kernel); synchronises */
int mot_tdd_dropped(mot_tdd_t *t, int s)
{
    if (!t || s < 0 || s >= t->st.S) return mot_fail(MOT_ERR_ARG, "mot_tdd_dropped: bad argument");
    if (!t->st.kcf) return 0;
    mot_ctx_t *c = t->ctx;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    int n = 0;
    CU(cudaMemcpy(&n, t->st.dropped + s, sizeof(int), cudaMemcpyDeviceToHost));
    return n;
}

/* snapshot of one stream's track table (synchronises); returns the number of tracks */
int mot_tdd_read(mot_tdd_t *t, int s, uint32_t *tid, mot_bbox_t *boxes, int *age, int *vis, int *invis)
{
    if (!t || s < 0 || s >= t->st.S) return mot_fail(MOT_ERR_ARG, "mot_tdd_read: bad argument");
    mot_ctx_t *c = t->ctx; TddState &st = t->st;
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    int n = 0;
    CU(cudaMemcpy(&n, st.ntracks + s, sizeof(int), cudaMemcpyDeviceToHost));
    const long o = (long)s * st.cap;
    if (tid) CU(cudaMemcpy(tid, st.tid + o, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    if (boxes) CU(cudaMemcpy(boxes, st.bbox + o, sizeof(mot_bbox_t) * n, cudaMemcpyDeviceToHost));
    if (age) CU(cudaMemcpy(age, st.age + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (vis) CU(cudaMemcpy(vis, st.vis + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
    if (invis) CU(cudaMemcpy(invis, st.invis + o, sizeof(int) * n, cudaMemcpyDeviceToHost));
    return n;
}

}  // extern "C"
