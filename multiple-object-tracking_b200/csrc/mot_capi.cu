// mot_capi.cu -- the C ABI declared in include/mot_b200.h: context, frames, tracker slots, batched launches.
// Host-side only bookkeeping; every numeric step of the hot path runs in the CUDA kernels of this directory.
// There is deliberately NO CPU fallback: a missing device, a failed launch or an unsupported shape is an error.
#include "mot_ctx.h"
#include "fhog_tables.h"
#include "overlay.h"
#include "yolo_post.h"

#include <cuda.h>            // CUtensorMap and its enums (the encoder itself is fetched through cudaGetDriverEntryPoint: no libcuda link dependency)

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace mot;

static thread_local std::string g_err;
int mot_fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define fail mot_fail

static constexpr int MAX_CLASSES = 1024;
static constexpr long DUMP_NB_MAX = 4096;       // stage dumps (test hook) are available for windows up to 64x64 cells

// Small host arrays (slots, frame indices, boxes) reach the device through this kernel reading the pinned staging
// buffers directly (zero-copy over PCIe) instead of through cudaMemcpyAsync: H2D copies of every stream share one copy
// engine, so a 256 KB memcpy would queue behind the 400 MB of next-frame uploads and serialise the pipeline.
__global__ void stage_in_kernel(int *d_slots, const int *h_slots, int *d_frames, const int *h_frames, mot_bbox_t *d_boxes, const mot_bbox_t *h_boxes, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    d_slots[i] = h_slots[i];
    if (d_frames) d_frames[i] = h_frames[i];
    d_boxes[i] = h_boxes[i];
}

// pinned host arrays -> device arrays of an association call: two box arrays (as 8-byte words) and the problem sizes
__global__ void stage_assoc_kernel(long long *d_a, const long long *h_a, long na, long long *d_b, const long long *h_b, long nb, int *d_c, const int *h_c, long nc)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) d_a[i] = h_a[i];
    if (i < nb) d_b[i] = h_b[i];
    if (i < nc) d_c[i] = h_c[i];
}

__global__ void meta_scatter_kernel(KcfMeta *meta, const KcfMeta *src, const int *slots, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) meta[slots[i]] = src[i];
}

__global__ void zero_state_kernel(float2 *model, long model_stride, float *alpha, long alpha_stride, const int *slots, long n_model, long n_alpha)
{
    const int s = slots[blockIdx.y];
    float2 *m = model + (long)s * model_stride;
    float *a = alpha + (long)s * alpha_stride;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_model; i += (long)gridDim.x * blockDim.x) m[i] = make_float2(0.f, 0.f);
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n_alpha; i += (long)gridDim.x * blockDim.x) a[i] = 0.f;
}

// ---- per-size constants (trackers/kcf.cpp:96-144, 203-207) computed on the host in the reference's precision -------------
static void hann_f(int N, std::vector<float> &h)
{
    const double PI_2 = 6.28318530717958647692;                     // include/sigpack/base/base.h:14
    h.resize(N);
    for (int i = 0; i < N; ++i) h[i] = (float)(0.5 - 0.5 * std::cos(1.0 * PI_2 * i / (N - 1)));   // window.h:34-48, 83-89
}

static void label_spectrum_re(int hr, int wc, std::vector<float> &yf_re)
{
    // gaussian_shaped_labels(0.7289, f_rows, f_cols) + circshift (kcf.cpp:96-122, 78-94), then Re(fft2) (kcf.cpp:132-144)
    const float sigma = 0.7289f;
    const float sinv = (float)(1.0 / (sigma * sigma));
    std::vector<float> gx(hr), gy(wc), lab((size_t)hr * wc);
    const int x0 = -hr / 2, y0 = -wc / 2;
    for (int i = 0; i < hr; ++i) { const int x = x0 + i; gx[i] = (float)std::exp(-0.5 * x * x * sinv); }
    for (int j = 0; j < wc; ++j) { const int y = y0 + j; gy[j] = (float)std::exp(-0.5 * y * y * sinv); }
    for (int j = 0; j < wc; ++j) {
        int jj = (j + y0) % wc; if (jj < 0) jj += wc;
        for (int i = 0; i < hr; ++i) { int ii = (i + x0) % hr; if (ii < 0) ii += hr; lab[(size_t)jj * hr + ii] = gx[i] * gy[j]; }
    }
    // Y[j][k] = sum_a sum_b lab[a][b] exp(-2 pi i (j a / wc + k b / hr)); separable evaluation in double
    const int sk = hr / 2 + 1;
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> tr((size_t)wc * sk), ti((size_t)wc * sk);
    for (int a = 0; a < wc; ++a)
        for (int k = 0; k < sk; ++k) {
            double sr = 0, si = 0;
            for (int b = 0; b < hr; ++b) {
                const double ang = two_pi * (double)((long)k * b % hr) / hr;
                sr += lab[(size_t)a * hr + b] * std::cos(ang); si -= lab[(size_t)a * hr + b] * std::sin(ang);
            }
            tr[(size_t)a * sk + k] = sr; ti[(size_t)a * sk + k] = si;
        }
    yf_re.resize((size_t)wc * sk);
    for (int j = 0; j < wc; ++j)
        for (int k = 0; k < sk; ++k) {
            double sr = 0;
            for (int a = 0; a < wc; ++a) {
                const double ang = two_pi * (double)((long)j * a % wc) / wc;
                sr += tr[(size_t)a * sk + k] * std::cos(ang) + ti[(size_t)a * sk + k] * std::sin(ang);
            }
            yf_re[(size_t)j * sk + k] = (float)sr;
        }
}

// ---- per-N tables of the any-size kernel (kcf_any.cuh): every cell-grid side 2..nmax -------------------------------------------
static void any_plan(int n, AnyPlan &pl)
{
    // fewest passes over the radices that have a register butterfly (2..10); a prime factor above 10 becomes a pass of its own,
    // evaluated from the definition.  Largest radix first: the first pass needs no twiddles.
    static const int allowed[] = { 10, 9, 8, 7, 6, 5, 4, 3, 2 };
    std::vector<int> best(n + 1, 1 << 20), pick(n + 1, 0);
    best[1] = 0;
    for (int v = 2; v <= n; ++v) {
        if (n % v) continue;
        for (int r : allowed) if (v % r == 0 && best[v / r] + 1 < best[v]) { best[v] = best[v / r] + 1; pick[v] = r; }
        if (!pick[v] || best[v] >= (1 << 20)) {
            int pr = 0;
            for (int q = 11; q <= v; q += 2) if (v % q == 0) { pr = q; break; }
            if (pr && best[v / pr] + 1 < best[v]) { best[v] = best[v / pr] + 1; pick[v] = pr; }
        }
    }
    pl = AnyPlan{};
    int nf = 0, fac[16];
    for (int v = n; v > 1 && pick[v] && nf < 16; v /= pick[v]) fac[nf++] = pick[v];
    std::sort(fac, fac + nf, [](int a, int b) { return a > b; });
    // a prime radix (> 10) goes last so that the passes before it stay cheap; keep at most 7 passes (n <= 2^14 always fits)
    std::stable_partition(fac, fac + nf, [](int r) { return r <= 10; });
    for (int i = 0; i < nf && i < 7; ++i) pl.r[i] = (unsigned short)fac[i];
    pl.nf = (unsigned short)std::min(nf, 7);
}

static int build_any_tables(mot_ctx_t *c, int nmax)
{
    const long total = any_off(nmax + 1);
    std::vector<float> hann(total); std::vector<float2> tw(total); std::vector<double2> lab(total); std::vector<AnyPlan> plan(nmax + 1);
    const double two_pi = 6.283185307179586476925286766559;
    const float sigma = 0.7289f;                                       // kcf.cpp:205
    const float sinv = (float)(1.0 / (sigma * sigma));
    std::vector<float> hn, g; std::vector<double> cs, sn;
    for (int N = 2; N <= nmax; ++N) {
        const long o = any_off(N);
        hann_f(N, hn);
        cs.resize(N); sn.resize(N);
        for (int t = 0; t < N; ++t) { cs[t] = std::cos(two_pi * t / N); sn[t] = std::sin(two_pi * t / N); tw[o + t] = make_float2((float)cs[t], (float)-sn[t]); hann[o + t] = hn[t]; }
        // 1-D label exp(-0.5 x^2 / sigma^2), x = -N/2 + i, circularly shifted so that x = 0 sits at index 0 (kcf.cpp:96-122, 78-94)
        g.assign(N, 0.f);
        const int x0 = -N / 2;
        for (int i = 0; i < N; ++i) { const int x = x0 + i; int ii = (i + x0) % N; if (ii < 0) ii += N; g[ii] = (float)std::exp(-0.5 * x * x * sinv); }
        for (int k = 0; k < N; ++k) {
            double sr = 0, si = 0;
            for (int b = 0; b < N; ++b) { const int tt = (int)((long)k * b % N); sr += g[b] * cs[tt]; si -= g[b] * sn[tt]; }
            lab[o + k] = make_double2(sr, si);
        }
        any_plan(N, plan[N]);
    }
    float *d_h; float2 *d_t; double2 *d_l; AnyPlan *d_p;
    CU(cudaMalloc(&d_h, sizeof(float) * total)); CU(cudaMalloc(&d_t, sizeof(float2) * total)); CU(cudaMalloc(&d_l, sizeof(double2) * total)); CU(cudaMalloc(&d_p, sizeof(AnyPlan) * (nmax + 1)));
    CU(cudaMemcpy(d_h, hann.data(), sizeof(float) * total, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_t, tw.data(), sizeof(float2) * total, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_l, lab.data(), sizeof(double2) * total, cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_p, plan.data(), sizeof(AnyPlan) * (nmax + 1), cudaMemcpyHostToDevice));
    c->any = AnyTablesDev{ d_h, d_t, d_l, d_p, nmax };
    return 0;
}

static int get_class(mot_ctx_t *c, int hr, int wc, int *out)
{
    for (size_t i = 0; i < c->classes.size(); ++i) if (c->classes[i].hr == hr && c->classes[i].wc == wc) { *out = (int)i; return 0; }
    if ((int)c->classes.size() >= MAX_CLASSES) return fail(MOT_ERR_CAPACITY, "too many distinct window sizes (%d)", MAX_CLASSES);
    SizeClass sc{}; sc.hr = hr; sc.wc = wc; sc.live = 0;
    sc.norm = (float)(1.0 / ((float)(wc * hr * 31)));                // kcf.cpp:197
    std::vector<float> wy, wx, yf;
    hann_f(hr, wy); hann_f(wc, wx); label_spectrum_re(hr, wc, yf);
    CU(cudaMalloc(&sc.d_wy, sizeof(float) * hr)); CU(cudaMalloc(&sc.d_wx, sizeof(float) * wc)); CU(cudaMalloc(&sc.d_yf, sizeof(float) * yf.size()));
    CU(cudaMemcpy(sc.d_wy, wy.data(), sizeof(float) * hr, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc.d_wx, wx.data(), sizeof(float) * wc, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(sc.d_yf, yf.data(), sizeof(float) * yf.size(), cudaMemcpyHostToDevice));
    sc.fast = !c->ext_on && kcf_fast_smem_bytes(hr, wc) != 0;          // the extensions live in the any-size kernel only
    sc.any_smem = (!sc.fast && hr >= 2 && wc >= 2 && hr <= c->any.nmax && wc <= c->any.nmax) ? kcf_any_smem_bytes(hr, wc, c->lut_floats) : 0;
    {
        // exp(-2 pi i t / n) tables for the any-size DFT path, evaluated in double with exact argument reduction
        const double two_pi = 6.283185307179586476925286766559;
        std::vector<double2> th(hr), tw(wc);
        for (int t = 0; t < hr; ++t) th[t] = make_double2(std::cos(two_pi * t / hr), -std::sin(two_pi * t / hr));
        for (int t = 0; t < wc; ++t) tw[t] = make_double2(std::cos(two_pi * t / wc), -std::sin(two_pi * t / wc));
        CU(cudaMalloc(&sc.d_twh, sizeof(double2) * hr)); CU(cudaMalloc(&sc.d_tww, sizeof(double2) * wc));
        CU(cudaMemcpy(sc.d_twh, th.data(), sizeof(double2) * hr, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(sc.d_tww, tw.data(), sizeof(double2) * wc, cudaMemcpyHostToDevice));
    }
    KcfClassDev kd{ hr, wc, sc.d_wy, sc.d_wx, sc.d_yf, sc.norm, sc.d_twh, sc.d_tww };
    CU(cudaMemcpy(c->d_classes + c->classes.size(), &kd, sizeof(kd), cudaMemcpyHostToDevice));
    c->classes.push_back(sc);
    *out = (int)c->classes.size() - 1;
    return 0;
}

// The compute stream waits for the uploads of the frame slots a launch is going to read (all pending ones when the slots
// are only known on the device).  Uploads of OTHER slots (the next frames) keep streaming underneath the kernels.
static int wait_frames(mot_ctx_t *c, int n, const int *frame_slots)
{
    if (frame_slots) {
        for (int i = 0; i < n; ++i) {
            const int s = frame_slots[i];
            if (s >= 0 && s < c->n_frames && c->slot_pending[s]) { CU(cudaStreamWaitEvent(c->stream, c->slot_uploaded[s], 0)); c->slot_pending[s] = 0; }
        }
    } else {
        for (int s = 0; s < c->n_frames; ++s)
            if (c->slot_pending[s]) { CU(cudaStreamWaitEvent(c->stream, c->slot_uploaded[s], 0)); c->slot_pending[s] = 0; }
    }
    return 0;
}

// Tensor maps of the frame slots: every frame as a 2-D array of 32-bit words (rows of frame_stride bytes), box = 108 words (432 bytes,
// the staged row pitch of kcf_fused.cuh) x 35 / 67 / 131 rows (the tallest crop of an 8 / 16 / 32-cell window).  One
// cp.async.bulk.tensor.2d then fetches a whole crop.  Any failure just leaves the kernels on their row-by-row bulk copies.
static void build_frame_tmaps(mot_ctx_t *c)
{
    c->tmaps_ok = false;
    if (c->kind != MOT_TRACKER_KCF || (c->frame_stride & 15) || c->frame_stride < 432) return;
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = nullptr;
    static bool looked = false;
    if (!looked) {
        looked = true;
        void *fp = nullptr; cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) encode = (encode_fn)fp;
    }
    if (!encode) return;
    static const cuuint32_t box_rows[3] = { 35, 67, 131 };
    std::vector<CUtensorMap> maps((size_t)c->n_frames * 3);
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    for (int s = 0; s < c->n_frames; ++s)
        for (int k = 0; k < 3; ++k) {
            CUtensorMap &m = maps[(size_t)s * 3 + k];
            memset(&m, 0, sizeof(m));
            if (!c->frame_ptr_h[s] || ((uintptr_t)c->frame_ptr_h[s] & 15)) { if (c->frame_ptr_h[s]) return; else continue; }
            const cuuint64_t dims[2] = { (cuuint64_t)c->frame_stride / 4, (cuuint64_t)c->H };
            const cuuint64_t strides[1] = { (cuuint64_t)c->frame_stride };
            const cuuint32_t box[2] = { 108, box_rows[k] }, estr[2] = { 1, 1 };
            if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint8_t *>(c->frame_ptr_h[s]), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return;
        }
    if (!c->d_frame_tmaps && cudaMalloc(&c->d_frame_tmaps, sizeof(CUtensorMap) * maps.size()) != cudaSuccess) { c->d_frame_tmaps = nullptr; return; }
    if (cudaMemcpyAsync(c->d_frame_tmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess) return;
    c->tmaps_ok = getenv("MOT_NO_TMA2D") == nullptr;
}

static int sync_frame_ptrs(mot_ctx_t *c)
{
    if (!c->frame_ptr_dirty) return 0;
    build_frame_tmaps(c);
    CU(cudaMemcpyAsync(c->d_frame_ptr, c->frame_ptr_h.data(), sizeof(void *) * c->n_frames, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));      // the host vector may change right after
    c->frame_ptr_dirty = false;
    return 0;
}

extern "C" { static void fill_launch(mot_ctx_t *c, KcfLaunch &L, int n, const int *d_slots, const int *d_frames, mot_bbox_t *d_boxes, int clamp); }

namespace {
struct DevScratch {                      // frees whatever was allocated when the call returns, on every path
    std::vector<void *> p;
    ~DevScratch() { for (void *q : p) cudaFree(q); }
    template <class T> cudaError_t get(T **out, size_t bytes) { void *q = nullptr; const cudaError_t e = cudaMalloc(&q, bytes); if (e == cudaSuccess) { p.push_back(q); *out = (T *)q; } return e; }
};
}

int mot_ctx_kcf_class(mot_ctx_t *c, int hr, int wc, int *cls_out) { return get_class(c, hr, wc, cls_out); }

int mot_ctx_frames_ready(mot_ctx_t *c)
{
    { const int rc = sync_frame_ptrs(c); if (rc) return rc; }
    return wait_frames(c, 0, nullptr);
}

int mot_ctx_kcf_launch(mot_ctx_t *c, int mode, int cls, int n_max, const int *n_dev, const int *slots, const int *frames,
                       mot_bbox_t *boxes, const int *box_index, int clamp)
{
    if (cls < 0 || cls >= (int)c->classes.size() || !c->classes[cls].fast) return fail(MOT_ERR_SHAPE, "class %d has no fused kernel", cls);
    KcfLaunch L; fill_launch(c, L, n_max, slots, frames, boxes, clamp);
    L.n_jobs_dev = n_dev; L.box_index = box_index; L.dump = KcfDump{};
    const int rc = kcf_launch_fast(mode, c->classes[cls].hr, c->classes[cls].wc, L, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "KCF launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

// cost matrices + assignment over device arrays with a caller-owned working copy (n_mat * max_dim^2 doubles): the device-resident
// frame loop passes its own, so that a captured graph never holds a pointer into a context buffer that may be reallocated
int mot_ctx_associate_dev(mot_ctx_t *c, int n_mat, const int *d_T, const int *d_D, const mot_bbox_t *d_trk, long trk_stride,
                          const mot_bbox_t *d_det, long det_stride, int cost_mode, double *d_dist, long dist_stride,
                          int *d_assign, long assign_stride, double *d_cost, int max_dim, double *d_work)
{
    if (n_mat == 0) return 0;
    if (!d_dist || dist_stride < (long)max_dim * max_dim) return fail(MOT_ERR_ARG, "the cost matrices need a device buffer of max_dim^2 doubles per problem");
    AssocLaunch A{};
    A.n_mat = n_mat; A.T = d_T; A.D = d_D; A.trk = d_trk; A.trk_stride = trk_stride; A.det = d_det; A.det_stride = det_stride;
    A.cost_mode = cost_mode; A.screen_dis = 1.0 / (double)c->W;
    A.dist = d_dist; A.dist_stride = dist_stride; A.assign = d_assign; A.assign_stride = assign_stride; A.cost = d_cost; A.max_dim = max_dim;
    A.work = d_work; A.work_stride = (long)max_dim * max_dim;
    int rc = assoc_cost(A, c->stream); if (rc) return fail(MOT_ERR_CUDA, "cost kernel launch failed (%d)", rc);
    rc = assoc_solve(A, nullptr, c->stream); if (rc) return fail(MOT_ERR_CUDA, "munkres kernel launch failed (%d)", rc);
    c->launches += 2;
    return 0;
}

// CTA size and CTAs per SM of an any-size launch with `smem` bytes of shared memory per CTA.  The register file (64 K) allows 1024
// threads per SM at 64 registers or 512 at 128; the kernel is much shorter with 128 (no spills, nothing rematerialised), and two
// or more CTAs per SM hide each other's barriers.  Measured on B200 (profiles/ab_any.sh): one CTA per SM -> 512 threads x 128
// registers; two or three -> 2 x 256 x 128; four or more (small windows) -> 4 x 256 x 64.
static void any_launch_shape(size_t smem, int *threads, int *ctas)
{
    const int k = (int)((227 * 1024) / (smem + 1024));                // per-CTA reservation included
    if (k >= 4) { *threads = 256; *ctas = 4; }
    else if (k >= 2) { *threads = 256; *ctas = 2; }
    else { *threads = 512; *ctas = 1; }
}


int mot_ctx_kcf_launch_any(mot_ctx_t *c, int mode, size_t smem_bytes, int n_max, const int *n_dev, const int *slots, const int *frames,
                           mot_bbox_t *boxes, const int *box_index, int clamp, float *scratch, long scratch_stride_floats, int scratch_ctas)
{
    KcfLaunch L; fill_launch(c, L, n_max, slots, frames, boxes, clamp);
    L.n_jobs_dev = n_dev; L.box_index = box_index; L.dump = KcfDump{};
    int threads, ctas; any_launch_shape(smem_bytes, &threads, &ctas);
    const int rc = kcf_launch_any(mode, L, c->any, smem_bytes, threads, ctas, c->d_any_err, scratch, scratch_stride_floats, scratch ? scratch_ctas : 0, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "KCF (any-size kernel) launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 1;
    return 0;
}

// =====================================================================================================================
extern "C" {

const char *mot_last_error(void) { return g_err.c_str(); }

int mot_ctx_create(mot_ctx_t **out, int device, int frame_w, int frame_h, int max_tracks, int n_frame_slots, int tracker_kind)
{
    if (!out || frame_w <= 0 || frame_h <= 0 || max_tracks <= 0 || n_frame_slots <= 0) return fail(MOT_ERR_ARG, "mot_ctx_create: bad argument");
    if (tracker_kind != MOT_TRACKER_KALMAN && tracker_kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "unknown tracker kind %d", tracker_kind);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) return fail(MOT_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(MOT_ERR_ARG, "device %d out of range (%d visible)", device, ndev);
    CU(cudaSetDevice(device));
    const FhogTables &ft = fhog_tables();
    if (!ft.ok) return fail(MOT_ERR_TABLES, "%s", ft.error.c_str());
    mot_ctx_t *c = new mot_ctx_t();
    c->device = device; c->W = frame_w; c->H = frame_h; c->max_tracks = max_tracks; c->n_frames = n_frame_slots; c->kind = tracker_kind;
    CU(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    c->frame_owned.assign(n_frame_slots, nullptr);
    c->frame_ptr_h.assign(n_frame_slots, nullptr);
    c->slot_uploaded.assign(n_frame_slots, nullptr); c->slot_pending.assign(n_frame_slots, 0);
    for (auto &e : c->slot_uploaded) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->frame_stride = frame_w * 3;
    CU(cudaMalloc(&c->d_frame_ptr, sizeof(void *) * n_frame_slots));
    c->used.assign(max_tracks, 0);
    c->meta_h.assign(max_tracks, KcfMeta{});
    for (int i = max_tracks - 1; i >= 0; --i) c->free_slots.push_back(i);
    if (tracker_kind == MOT_TRACKER_KCF) {
        c->model_stride = (long)KCF_CHAN * NB_MAX;                   // S <= nb <= NB_MAX
        c->alpha_stride = NB_MAX;
        CU(cudaMalloc(&c->d_meta, sizeof(KcfMeta) * max_tracks));
        CU(cudaMalloc(&c->d_model, sizeof(float2) * c->model_stride * max_tracks));
        CU(cudaMalloc(&c->d_alpha, sizeof(float) * c->alpha_stride * max_tracks));
        CU(cudaMalloc(&c->d_classes, sizeof(KcfClassDev) * MAX_CLASSES));
        CU(cudaMalloc(&c->d_tab_rsqrt, sizeof(float) * ft.rsqrt_tab.size()));
        CU(cudaMalloc(&c->d_tab_rcp, sizeof(float) * ft.rcp_tab.size()));
        CU(cudaMalloc(&c->d_tab_bin, sizeof(uint32_t) * (ft.bin_tab.size() + 4)));      // bulk copies move 16-byte multiples
        CU(cudaMemset(c->d_tab_bin, 0, sizeof(uint32_t) * (ft.bin_tab.size() + 4)));
        CU(cudaMemcpy(c->d_tab_rsqrt, ft.rsqrt_tab.data(), sizeof(float) * ft.rsqrt_tab.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_tab_rcp, ft.rcp_tab.data(), sizeof(float) * ft.rcp_tab.size(), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(c->d_tab_bin, ft.bin_tab.data(), sizeof(uint32_t) * ft.bin_tab.size(), cudaMemcpyHostToDevice));
        {
            // device copy of the fused {rsqrt, rcp/16} table with the constant part of the exponent arithmetic folded in:
            // the kernel computes m = T * 2^-q and M/16 = R16 * 2^q as bits(T) - qs and bits(R16) + qs with
            // qs = (q << 23) = qs' - 0x20000000, qs' = ((bits(M2) + 0x00800000) >> 1) & 0x7F800000; storing
            // bits(T) + 0x20000000 and bits(R16) - 0x20000000 saves the subtraction per pixel (fhog_common.cuh)
            std::vector<uint32_t> biased(ft.rsrc_tab.size());
            for (size_t i = 0; i < biased.size(); ++i) {
                uint32_t u; memcpy(&u, &ft.rsrc_tab[i], 4);
                biased[i] = (i & 1) ? u - 0x20000000u : u + 0x20000000u;
            }
            CU(cudaMalloc(&c->d_tab_rsrc, sizeof(float) * biased.size()));
            CU(cudaMemcpy(c->d_tab_rsrc, biased.data(), sizeof(uint32_t) * biased.size(), cudaMemcpyHostToDevice));
        }
        CU(cudaMalloc(&c->d_tab_bin2, sizeof(uint32_t) * (ft.bin2_tab.size() + 4)));
        CU(cudaMemset(c->d_tab_bin2, 0, sizeof(uint32_t) * (ft.bin2_tab.size() + 4)));
        CU(cudaMemcpy(c->d_tab_bin2, ft.bin2_tab.data(), sizeof(uint32_t) * ft.bin2_tab.size(), cudaMemcpyHostToDevice));
        c->tab = FhogTablesDev{ c->d_tab_rsqrt, ft.rsqrt_bits, c->d_tab_rcp, ft.rcp_bits, c->d_tab_bin, ft.bin_shift, ft.bin_nseg,
                                reinterpret_cast<const float2 *>(c->d_tab_rsrc), ft.rcp_cap, c->d_tab_bin2, ft.u_cap };
        c->lut_floats = 2 * (2 << ft.rsqrt_bits) + ((2 * ft.bin_nseg + 3) & ~3);
        if (c->lut_floats > 8192) c->lut_floats = 0;
        CU(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
        { const int rc = build_any_tables(c, std::max(32, std::max(frame_w, frame_h) / KCF_CELL)); if (rc) return rc; }
        CU(cudaHostAlloc(&c->h_any_err, sizeof(int), cudaHostAllocMapped)); *c->h_any_err = 0;
        CU(cudaHostGetDevicePointer(&c->d_any_err, c->h_any_err, 0));
    } else {
        c->kal.cap = max_tracks;
        CU(cudaMalloc(&c->kal.x, sizeof(double) * 6 * max_tracks));
        CU(cudaMalloc(&c->kal.P, sizeof(double) * 36 * max_tracks));
    }
    *out = c;
    return 0;
}

void mot_ctx_destroy(mot_ctx_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto p : c->frame_owned) if (p) cudaFree(p);
    cudaFree(c->d_alpha_im); cudaFree(c->d_frame_tmaps); cudaFree(c->d_frame_ptr); cudaFree(c->d_meta); cudaFree(c->d_model); cudaFree(c->d_alpha); cudaFree(c->d_classes);
    cudaFree(c->d_tab_rsqrt); cudaFree(c->d_tab_rcp); cudaFree(c->d_tab_rsrc); cudaFree(c->d_tab_bin); cudaFree(c->d_tab_bin2); cudaFree(c->kal.x); cudaFree(c->kal.P);
    if (c->h_any_err) cudaFreeHost(c->h_any_err);
    cudaFree((void *)c->any.hann); cudaFree((void *)c->any.tw); cudaFree((void *)c->any.lab); cudaFree((void *)c->any.plan);
    for (auto &sc : c->classes) { cudaFree(sc.d_wy); cudaFree(sc.d_wx); cudaFree(sc.d_yf); cudaFree(sc.d_twh); cudaFree(sc.d_tww); }
    for (auto &m : c->meta_h) { if (m.model_ptr) cudaFree(m.model_ptr); if (m.alpha_ptr) cudaFree(m.alpha_ptr); }
    c->d_scratch.release();
    c->d_slots.release(); c->d_frames.release(); c->d_TD.release(); c->d_assign.release(); c->d_boxes.release(); c->d_trk.release();
    c->d_det.release(); c->d_meta_stage.release(); c->d_gray.release(); c->d_dist.release(); c->d_work.release(); c->d_cost.release();
    c->h_slots.release(); c->h_frames.release(); c->h_TD.release(); c->h_assign.release(); c->h_boxes.release(); c->h_trk.release();
    c->h_det.release(); c->h_meta_stage.release(); c->h_dist.release(); c->h_cost.release();
    cudaFree(c->dump.gray); cudaFree(c->dump.m0); cudaFree(c->dump.bin); cudaFree(c->dump.r1); cudaFree(c->dump.nrm); cudaFree(c->dump.feat);
    cudaFree(c->dump.spec); cudaFree(c->dump.zf); cudaFree(c->dump.resp); cudaFree(c->dump.kf); cudaFree(c->dump.peak); cudaFree(c->dump.margin);
    cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); for (auto e : c->slot_uploaded) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int mot_ctx_set_kcf_options(mot_ctx_t *c, const mot_kcf_options_t *o)
{
    if (!c || !o) return fail(MOT_ERR_ARG, "mot_ctx_set_kcf_options: null argument");
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    for (char u : c->used) if (u) return fail(MOT_ERR_ARG, "the KCF options must be set before the first tracker is created");
    if (o->gaussian_kernel && !(o->kernel_sigma > 0.f)) return fail(MOT_ERR_ARG, "the Gaussian kernel needs kernel_sigma > 0");
    if (o->padding != 0.f && (o->padding < 1.f || o->padding > 4.f)) return fail(MOT_ERR_ARG, "padding must be in [1, 4]");
    if (o->output_sigma_factor < 0.f) return fail(MOT_ERR_ARG, "output_sigma_factor must be >= 0");
    CU(cudaSetDevice(c->device));
    c->ext = KcfExt{ o->gaussian_kernel ? 1 : 0, o->kernel_sigma, o->subpixel_peak ? 1 : 0, o->padding > 1.f ? o->padding : 0.f, o->output_sigma_factor };
    c->ext_on = c->ext.gaussian || c->ext.subpixel || c->ext.padding > 1.f || c->ext.osf > 0.f;
    if (c->ext.gaussian && !c->d_alpha_im) CU(cudaMalloc(&c->d_alpha_im, sizeof(float) * c->alpha_stride * c->max_tracks));
    // the per-size classes remember which kernel serves them: start afresh
    for (auto &sc : c->classes) { sc.fast = !c->ext_on && kcf_fast_smem_bytes(sc.hr, sc.wc) != 0;
                                  sc.any_smem = (!sc.fast && sc.hr >= 2 && sc.wc >= 2 && sc.hr <= c->any.nmax && sc.wc <= c->any.nmax) ? kcf_any_smem_bytes(sc.hr, sc.wc, c->lut_floats) : 0; }
    return 0;
}

// The slot arena holds model and alpha of every window up to `alpha_stride` half-spectrum bins (default NB_MAX: the named 32x32-cell
// shape with room to spare).  Larger windows otherwise own individually allocated storage (host-managed trackers) or cannot be born on
// the device at all (device-resident loop): this call re-sizes the arena for windows up to max_rows x max_cols pixels, either orientation.
int mot_ctx_reserve_window(mot_ctx_t *c, int max_rows, int max_cols)
{
    if (!c || max_rows < 2 * KCF_CELL || max_cols < 2 * KCF_CELL) return fail(MOT_ERR_ARG, "mot_ctx_reserve_window: bad argument");
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    for (char u : c->used) if (u) return fail(MOT_ERR_ARG, "the window reservation must be made before the first tracker is created");
    CU(cudaSetDevice(c->device));
    const long hr = max_rows / KCF_CELL, wc = max_cols / KCF_CELL;
    const long bins = std::max<long>(NB_MAX, std::max(wc * (hr / 2 + 1), hr * (wc / 2 + 1)));
    if (bins == c->alpha_stride) return 0;
    float2 *nm = nullptr; float *na = nullptr, *ni = nullptr;
    cudaError_t e = cudaMalloc(&nm, sizeof(float2) * KCF_CHAN * bins * c->max_tracks);
    if (e == cudaSuccess) e = cudaMalloc(&na, sizeof(float) * bins * c->max_tracks);
    if (e == cudaSuccess && c->d_alpha_im) e = cudaMalloc(&ni, sizeof(float) * bins * c->max_tracks);
    if (e != cudaSuccess) { cudaFree(nm); cudaFree(na); cudaFree(ni); cudaGetLastError(); return fail(MOT_ERR_CAPACITY, "no memory for %d slots of %ld bins: %s", c->max_tracks, bins, cudaGetErrorString(e)); }
    CU(cudaStreamSynchronize(c->stream));
    cudaFree(c->d_model); cudaFree(c->d_alpha); if (c->d_alpha_im) cudaFree(c->d_alpha_im);
    c->d_model = nm; c->d_alpha = na; c->d_alpha_im = ni;
    c->model_stride = (long)KCF_CHAN * bins; c->alpha_stride = bins;
    return 0;
}

int mot_ctx_set_stream(mot_ctx_t *c, void *s) { if (!c) return fail(MOT_ERR_ARG, "null ctx"); c->stream = s ? (cudaStream_t)s : c->own_stream; return 0; }
int mot_sync(mot_ctx_t *c)
{
    if (!c) return fail(MOT_ERR_ARG, "null ctx");
    CU(cudaSetDevice(c->device)); CU(cudaStreamSynchronize(c->copy_stream)); CU(cudaStreamSynchronize(c->stream));
    if (c->h_any_err && *c->h_any_err) { *c->h_any_err = 0; return fail(MOT_ERR_SHAPE, "a KCF job did not fit the shared memory of its launch and was skipped (internal sizing error)"); }
    return 0;
}
long mot_launch_count(mot_ctx_t *c) { return c ? c->launches : 0; }
int mot_ctx_kind(mot_ctx_t *c) { return c ? c->kind : MOT_ERR_ARG; }

// ---- frames ----------------------------------------------------------------------------------------------------------
int mot_frame_upload(mot_ctx_t *c, int slot, const uint8_t *host_bgr, int stride_bytes)
{
    if (!c || slot < 0 || slot >= c->n_frames || !host_bgr) return fail(MOT_ERR_ARG, "mot_frame_upload: bad argument");
    CU(cudaSetDevice(c->device));
    if (!c->frame_owned[slot]) CU(cudaMalloc(&c->frame_owned[slot], (size_t)c->W * 3 * c->H));
    if (c->frame_ptr_h[slot] != c->frame_owned[slot]) { c->frame_ptr_h[slot] = c->frame_owned[slot]; c->frame_ptr_dirty = true; }
    if (c->frame_stride != c->W * 3) return fail(MOT_ERR_ARG, "mixing bound device frames of another stride with uploaded frames");
    // On the copy stream: the upload of the next frame overlaps with kernels still running on the compute stream.  The next
    // launch waits for it (frames_ready); the caller must not overwrite a slot that enqueued, unfinished work still reads
    // (the host-array calls synchronise before returning, so alternating two slots per stream is always safe).
    if (stride_bytes == c->W * 3) CU(cudaMemcpyAsync(c->frame_owned[slot], host_bgr, (size_t)c->W * 3 * c->H, cudaMemcpyHostToDevice, c->copy_stream));
    else CU(cudaMemcpy2DAsync(c->frame_owned[slot], (size_t)c->W * 3, host_bgr, (size_t)stride_bytes, (size_t)c->W * 3, c->H, cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->slot_uploaded[slot], c->copy_stream));
    c->slot_pending[slot] = 1;
    return 0;
}

int mot_frame_bind_device(mot_ctx_t *c, int slot, const uint8_t *dev_bgr, int stride_bytes)
{
    if (!c || slot < 0 || slot >= c->n_frames || !dev_bgr || stride_bytes < c->W * 3) return fail(MOT_ERR_ARG, "mot_frame_bind_device: bad argument");
    bool any_other = false;
    for (int i = 0; i < c->n_frames; ++i) if (i != slot && c->frame_ptr_h[i]) any_other = true;
    if (any_other && stride_bytes != c->frame_stride) return fail(MOT_ERR_ARG, "all frame slots must share one stride (%d vs %d)", stride_bytes, c->frame_stride);
    c->frame_stride = stride_bytes;
    if (c->frame_ptr_h[slot] != dev_bgr) { c->frame_ptr_h[slot] = dev_bgr; c->frame_ptr_dirty = true; }
    return 0;
}

int mot_frame_download(mot_ctx_t *c, int slot, uint8_t *host_bgr, int stride_bytes)
{
    if (!c || slot < 0 || slot >= c->n_frames || !host_bgr || stride_bytes < c->W * 3) return fail(MOT_ERR_ARG, "mot_frame_download: bad argument");
    if (!c->frame_ptr_h[slot]) return fail(MOT_ERR_ARG, "frame slot %d holds no frame", slot);
    CU(cudaSetDevice(c->device));
    { const int rc = wait_frames(c, 1, &slot); if (rc) return rc; }
    CU(cudaMemcpy2DAsync(host_bgr, (size_t)stride_bytes, c->frame_ptr_h[slot], (size_t)c->frame_stride, (size_t)c->W * 3, c->H, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

// ---- overlay (top/td.cpp:647-733) -----------------------------------------------------------------------------------------
uint32_t mot_track_color(uint32_t tid) { return track_color(tid); }

int mot_overlay_batch(mot_ctx_t *c, int n, const int *frame_slots, const mot_bbox_t *boxes, const uint32_t *rgb, int thickness)
{
    if (!c || n < 0 || thickness < 0 || (n && (!frame_slots || !boxes || !rgb))) return fail(MOT_ERR_ARG, "mot_overlay_batch: bad argument");
    if (n == 0 || thickness == 0) return 0;
    CU(cudaSetDevice(c->device));
    // entries grouped by frame slot, order kept inside a slot (later entries overwrite earlier ones, like the reference's loop)
    const int ns = c->n_frames;
    CU(c->h_slots.ensure((size_t)ns + 1)); CU(c->h_frames.ensure(n)); CU(c->h_boxes.ensure(n)); CU(c->h_TD.ensure(n));
    CU(c->d_slots.ensure((size_t)ns + 1)); CU(c->d_frames.ensure(n)); CU(c->d_boxes.ensure(n)); CU(c->d_TD.ensure(n));
    int *begin = c->h_slots.p, *order = c->h_frames.p;
    for (int s = 0; s <= ns; ++s) begin[s] = 0;
    for (int i = 0; i < n; ++i) {
        const int s = frame_slots[i];
        if (s < 0 || s >= ns || !c->frame_ptr_h[s]) return fail(MOT_ERR_ARG, "entry %d: frame slot %d holds no frame", i, s);
        begin[s + 1]++;
    }
    for (int s = 0; s < ns; ++s) begin[s + 1] += begin[s];
    {
        std::vector<int> cur(begin, begin + ns);
        for (int i = 0; i < n; ++i) order[cur[frame_slots[i]]++] = i;
    }
    std::memcpy(c->h_boxes.p, boxes, sizeof(mot_bbox_t) * n);
    std::memcpy(c->h_TD.p, rgb, sizeof(uint32_t) * n);
    { const int rc = wait_frames(c, n, frame_slots); if (rc) return rc; }
    { const int rc = sync_frame_ptrs(c); if (rc) return rc; }
    CU(cudaMemcpyAsync(c->d_slots.p, begin, sizeof(int) * (ns + 1), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_frames.p, order, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_boxes.p, c->h_boxes.p, sizeof(mot_bbox_t) * n, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_TD.p, c->h_TD.p, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, c->stream));
    const int rc = overlay_draw(const_cast<uint8_t *const *>(reinterpret_cast<const uint8_t *const *>(c->d_frame_ptr)), ns, c->frame_stride,
                                (long)c->frame_stride * c->H, c->d_slots.p, c->d_frames.p, c->d_boxes.p, reinterpret_cast<const uint32_t *>(c->d_TD.p), thickness, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "overlay launch failed (%d)", rc);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->stream));      // the pinned staging arrays are reused by the next call
    return 0;
}

// ---- detector post-processing (detectors/yolo3.cpp:141-356, 487-527) ---------------------------------------------------------
int mot_yolo_post(mot_ctx_t *c, const float *out0, const float *out1, const float *out2, const int *anchors18, float obj_thresh, float nms_thresh,
                  int tensor_h, int tensor_w, int image_h, int image_w, int num_classes, mot_bbox_t *out, int max_out)
{
    if (!c || !out0 || !out1 || !out2 || !anchors18 || !out || max_out <= 0) return fail(MOT_ERR_ARG, "mot_yolo_post: null argument");
    if (tensor_h < 32 || tensor_w < 32 || tensor_h % 32 || tensor_w % 32 || image_h <= 0 || image_w <= 0 || num_classes < 1 || num_classes > 1024)
        return fail(MOT_ERR_SHAPE, "mot_yolo_post: tensor %dx%d (multiples of 32), image %dx%d, %d classes (1..1024)", tensor_h, tensor_w, image_h, image_w, num_classes);
    CU(cudaSetDevice(c->device));
    DevScratch ds;
    const float *host[3] = { out0, out1, out2 };
    float *d_o[3]; int *d_anch, *d_cnt; void *d_cand; mot_bbox_t *d_boxes;
    const size_t per = (size_t)3 * (5 + num_classes);
    for (int k = 0; k < 3; ++k) {
        const size_t nfl = (size_t)((tensor_h / 32) << k) * ((tensor_w / 32) << k) * per;
        CU(ds.get(&d_o[k], sizeof(float) * nfl));
        CU(cudaMemcpyAsync(d_o[k], host[k], sizeof(float) * nfl, cudaMemcpyHostToDevice, c->stream));
    }
    CU(ds.get(&d_anch, sizeof(int) * 18)); CU(ds.get(&d_cnt, sizeof(int) * 2)); CU(ds.get(&d_cand, yolo_cand_bytes())); CU(ds.get(&d_boxes, sizeof(mot_bbox_t) * max_out));
    CU(cudaMemcpyAsync(d_anch, anchors18, sizeof(int) * 18, cudaMemcpyHostToDevice, c->stream));
    const float *d_oc[3] = { d_o[0], d_o[1], d_o[2] };
    const int rc = yolo_post_launch(d_oc, d_anch, obj_thresh, nms_thresh, tensor_h, tensor_w, image_h, image_w, num_classes, d_cand, d_cnt, d_boxes, max_out, d_cnt + 1, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "YOLO post-processing launch failed: %s", cudaGetErrorString((cudaError_t)rc));
    c->launches += 4;
    int h_cnt[2] = { 0, 0 };
    CU(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int) * 2, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (h_cnt[0] > yolo_cap()) return fail(MOT_ERR_CAPACITY, "%d candidates pass the objectness threshold (capacity %d)", h_cnt[0], yolo_cap());
    if (h_cnt[1] > 0) CU(cudaMemcpy(out, d_boxes, sizeof(mot_bbox_t) * h_cnt[1], cudaMemcpyDeviceToHost));
    return h_cnt[1];
}

// ---- tracker slots -----------------------------------------------------------------------------------------------------
int mot_tracker_new_batch(mot_ctx_t *c, int n, const mot_bbox_t *boxes, int *handles_out)
{
    if (!c || n < 0 || (n && (!boxes || !handles_out))) return fail(MOT_ERR_ARG, "mot_tracker_new_batch: bad argument");
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    if ((int)c->free_slots.size() < n) return fail(MOT_ERR_CAPACITY, "out of track slots (%d requested, %zu free)", n, c->free_slots.size());
    CU(c->h_slots.ensure(n)); CU(c->d_slots.ensure(n));
    if (c->kind == MOT_TRACKER_KCF) {
        CU(c->h_meta_stage.ensure(n)); CU(c->d_meta_stage.ensure(n));
        // validate everything first so that a failure leaves no half-created trackers
        const bool padded = c->ext.padding > 1.0f;
        for (int i = 0; i < n; ++i) {
            const mot_bbox_t wb = padded ? kcf_pad_box(boxes[i], c->ext.padding) : boxes[i];           // extension: the window is the padded target
            const int rows = wb.b - wb.t + 1, cols = wb.r - wb.l + 1;                                  // kcf.cpp:148-149
            const int hr = rows / KCF_CELL, wc = cols / KCF_CELL;
            if (hr < 2 || wc < 2) return fail(MOT_ERR_SHAPE, "window %dx%d px is smaller than 2x2 cells (8x8 px)", rows, cols);
            if (!padded && (rows > c->H || cols > c->W)) return fail(MOT_ERR_SHAPE, "window %dx%d px is larger than the %dx%d frame", rows, cols, c->H, c->W);
            if (c->ext_on && (!kcf_any_smem_bytes(hr, wc, c->lut_floats) || hr > c->any.nmax || wc > c->any.nmax || (long)wc * (hr / 2 + 1) > c->alpha_stride))
                return fail(MOT_ERR_SHAPE, "window %dx%d px: the KCF extensions are served by the any-size kernel up to %ld spectrum bins (mot_ctx_reserve_window raises it)", rows, cols, c->alpha_stride);
        }
        long max_model = 0, max_alpha = 0;
        for (int i = 0; i < n; ++i) {
            KcfMeta m{};
            const mot_bbox_t wb = padded ? kcf_pad_box(boxes[i], c->ext.padding) : boxes[i];
            m.rows = wb.b - wb.t + 1; m.cols = wb.r - wb.l + 1;
            m.hr = m.rows / KCF_CELL; m.wc = m.cols / KCF_CELL;
            m.pos = wb; m.scale_horiz = 1.0f; m.scale_vert = 1.0f; m.first_update = 1;                // kcf.cpp:200-209
            m.tw = std::abs(boxes[i].r - boxes[i].l) + 1; m.th = std::abs(boxes[i].b - boxes[i].t) + 1;
            int cls = 0; const int rc = get_class(c, m.hr, m.wc, &cls); if (rc) return rc;
            m.size_class = cls;
            const long S_ = (long)m.wc * (m.hr / 2 + 1);
            if (!c->classes[cls].fast && (!c->classes[cls].any_smem || S_ > c->alpha_stride)) {
                // sizes that no fused kernel holds own their model / alpha (they can exceed the fixed slot stride)
                CU(cudaMalloc(&m.model_ptr, sizeof(float2) * KCF_CHAN * S_)); CU(cudaMalloc(&m.alpha_ptr, sizeof(float) * S_));
                CU(cudaMemsetAsync(m.model_ptr, 0, sizeof(float2) * KCF_CHAN * S_, c->stream)); CU(cudaMemsetAsync(m.alpha_ptr, 0, sizeof(float) * S_, c->stream));
            }
            const int slot = c->free_slots.back(); c->free_slots.pop_back();
            c->used[slot] = 1; c->meta_h[slot] = m; c->classes[cls].live++;
            c->h_slots.p[i] = slot; c->h_meta_stage.p[i] = m; handles_out[i] = slot;
            if (c->classes[cls].fast || (c->classes[cls].any_smem && S_ <= c->alpha_stride)) { max_model = std::max(max_model, KCF_CHAN * S_); max_alpha = std::max(max_alpha, S_); }
        }
        CU(cudaMemcpyAsync(c->d_slots.p, c->h_slots.p, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_meta_stage.p, c->h_meta_stage.p, sizeof(KcfMeta) * n, cudaMemcpyHostToDevice, c->stream));
        meta_scatter_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(c->d_meta, c->d_meta_stage.p, c->d_slots.p, n);
        // xf_md and alpha start at zero (kcf.cpp:172-174)
        zero_state_kernel<<<dim3(8, n), 256, 0, c->stream>>>(c->d_model, c->model_stride, c->d_alpha, c->alpha_stride, c->d_slots.p, max_model, max_alpha);
        c->launches += 2;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(c->stream));      // staging buffers are reused by the next call
    } else {
        CU(c->h_boxes.ensure(n)); CU(c->d_boxes.ensure(n));
        for (int i = 0; i < n; ++i) {
            const int slot = c->free_slots.back(); c->free_slots.pop_back();
            c->used[slot] = 1; c->h_slots.p[i] = slot; c->h_boxes.p[i] = boxes[i]; handles_out[i] = slot;
        }
        CU(cudaMemcpyAsync(c->d_slots.p, c->h_slots.p, sizeof(int) * n, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(c->d_boxes.p, c->h_boxes.p, sizeof(mot_bbox_t) * n, cudaMemcpyHostToDevice, c->stream));
        const int rc = kalman_init(c->kal, n, c->d_slots.p, c->d_boxes.p, c->stream);
        if (rc) return fail(MOT_ERR_CUDA, "kalman_init launch failed (%d)", rc);
        c->launches += 1;
        CU(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int mot_tracker_spawnable(mot_ctx_t *c, const mot_bbox_t *box)
{
    if (!c || !box) return 0;
    if (c->kind != MOT_TRACKER_KCF) return 1;
    const bool padded = c->ext.padding > 1.0f;
    const mot_bbox_t wb = padded ? kcf_pad_box(*box, c->ext.padding) : *box;
    const int rows = wb.b - wb.t + 1, cols = wb.r - wb.l + 1, hr = rows / KCF_CELL, wc = cols / KCF_CELL;
    if (hr < 2 || wc < 2) return 0;
    if (c->ext_on) return kcf_any_smem_bytes(hr, wc, c->lut_floats) != 0 && hr <= c->any.nmax && wc <= c->any.nmax && (long)wc * (hr / 2 + 1) <= c->alpha_stride;
    return rows <= c->H && cols <= c->W;
}

int mot_tracker_delete_batch(mot_ctx_t *c, int n, const int *handles)
{
    if (!c || n < 0 || (n && !handles)) return fail(MOT_ERR_ARG, "mot_tracker_delete_batch: bad argument");
    for (int i = 0; i < n; ++i) {
        const int s = handles[i];
        if (s < 0 || s >= c->max_tracks || !c->used[s]) return fail(MOT_ERR_ARG, "delete of invalid handle %d", s);
        c->used[s] = 0;
        if (c->kind == MOT_TRACKER_KCF) {
            KcfMeta &m = c->meta_h[s];
            c->classes[m.size_class].live--;
            if (m.model_ptr || m.alpha_ptr) { cudaStreamSynchronize(c->stream); cudaFree(m.model_ptr); cudaFree(m.alpha_ptr); m.model_ptr = nullptr; m.alpha_ptr = nullptr; }
        }
        c->free_slots.push_back(s);
    }
    return 0;
}

// ---- KCF launches --------------------------------------------------------------------------------------------------------
static void fill_launch(mot_ctx_t *c, KcfLaunch &L, int n, const int *d_slots, const int *d_frames, mot_bbox_t *d_boxes, int clamp)
{
    L = KcfLaunch{};
    L.n_jobs = n; L.slots = d_slots; L.frames = d_frames; L.boxes = d_boxes;
    L.frame_tmaps = c->tmaps_ok ? c->d_frame_tmaps : nullptr;
    L.frame_ptr = c->d_frame_ptr; L.frame_w = c->W; L.frame_h = c->H; L.frame_stride = c->frame_stride;
    L.gray = nullptr; L.gray_stride = 0;
    L.meta = c->d_meta; L.model = c->d_model; L.model_stride = c->model_stride; L.alpha = c->d_alpha; L.alpha_stride = c->alpha_stride;
    L.classes = c->d_classes; L.tab = c->tab; L.clamp_to_frame = clamp;
    L.factor = 0.05f; L.lamda = 0.0001f;                                 // kcf.cpp:211-212
    L.ext = c->ext; L.alpha_im = c->d_alpha_im;
    if (c->dumps) L.dump = c->dump;
}

static int kcf_run(mot_ctx_t *c, int mode, int hr, int wc, KcfLaunch &L)
{
    if (!c->ext_on && kcf_fast_smem_bytes(hr, wc) != 0) {
        const int rc = kcf_launch_fast(mode, hr, wc, L, c->stream);
        if (rc) return fail(MOT_ERR_CUDA, "KCF launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
        return 0;
    }
    {
        // any other size that one CTA can hold: the fused any-size kernel; as many CTAs per SM as its shared memory allows
        const size_t smem = kcf_any_smem_bytes(hr, wc, c->lut_floats);
        if (smem && hr <= c->any.nmax && wc <= c->any.nmax) {
            int threads, ctas; any_launch_shape(smem, &threads, &ctas);
            // windows that run in strips park their cell histograms in a per-CTA scratch area (L2-resident: 18 planes of the cell grid)
            const size_t scr = kcf_any_scratch_bytes(hr, wc, c->lut_floats);
            float *scratch = nullptr;
            const int grid_cap = c->sm_count * ctas;
            if (scr) { CU(c->d_scratch.ensure(scr * (size_t)grid_cap)); scratch = reinterpret_cast<float *>(c->d_scratch.p); }
            const int rc = kcf_launch_any(mode, L, c->any, smem, threads, ctas, c->d_any_err, scratch, (long)(scr / sizeof(float)), grid_cap, c->stream);
            if (rc) return fail(MOT_ERR_CUDA, "KCF (any-size kernel) launch failed: %s", cudaGetErrorString((cudaError_t)rc));
            c->launches += 1;
            return 0;
        }
    }
    if (c->ext_on) return fail(MOT_ERR_SHAPE, "the KCF extensions need a window the any-size kernel holds (%dx%d cells does not fit)", hr, wc);
    // whatever is left (windows too large for one CTA): unfused pipeline with per-job scratch, processed in chunks of at most ~1 GB of scratch
    const size_t per_job = kcf_generic_scratch_bytes(hr, wc);
    size_t jobs = std::min<size_t>((size_t)L.n_jobs, std::max<size_t>(1, ((size_t)1 << 30) / per_job));
    CU(c->d_scratch.ensure(jobs * per_job));
    const int rc = kcf_launch_generic(mode, hr, wc, L, c->d_scratch.p, c->d_scratch.n, c->stream);
    if (rc < 0) return fail(MOT_ERR_CUDA, "KCF (any-size path) launch failed: %s", cudaGetErrorString((cudaError_t)(-rc)));
    c->launches += rc;
    return 0;
}

// host arrays: group the jobs by window size, one launch per size
static int kcf_batch_host(mot_ctx_t *c, int mode, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp, bool read_back)
{
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    if (n == 0) return 0;
    if (!handles || !frame_slots || !boxes) return fail(MOT_ERR_ARG, "null array");
    CU(cudaSetDevice(c->device));
    { const int rc = sync_frame_ptrs(c); if (rc) return rc; }
    { const int rc = wait_frames(c, n, frame_slots); if (rc) return rc; }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) {
        const int s = handles[i];
        if (s < 0 || s >= c->max_tracks || !c->used[s]) return fail(MOT_ERR_ARG, "invalid handle %d", s);
        if (frame_slots[i] < 0 || frame_slots[i] >= c->n_frames || !c->frame_ptr_h[frame_slots[i]]) return fail(MOT_ERR_ARG, "frame slot %d is empty", frame_slots[i]);
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c->meta_h[handles[a]].size_class < c->meta_h[handles[b]].size_class; });
    CU(c->h_slots.ensure(n)); CU(c->h_frames.ensure(n)); CU(c->h_boxes.ensure(n));
    CU(c->d_slots.ensure(n)); CU(c->d_frames.ensure(n)); CU(c->d_boxes.ensure(n));
    for (int i = 0; i < n; ++i) { c->h_slots.p[i] = handles[order[i]]; c->h_frames.p[i] = frame_slots[order[i]]; c->h_boxes.p[i] = boxes[order[i]]; }
    stage_in_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_slots.p, c->h_slots.p, c->d_frames.p, c->h_frames.p, c->d_boxes.p, c->h_boxes.p, n);
    CU(cudaGetLastError());
    c->launches += 1;
    for (int a = 0; a < n;) {
        const int cls = c->meta_h[c->h_slots.p[a]].size_class;
        int b = a; while (b < n && c->meta_h[c->h_slots.p[b]].size_class == cls) ++b;
        KcfLaunch L; fill_launch(c, L, b - a, c->d_slots.p + a, c->d_frames.p + a, c->d_boxes.p + a, clamp);
        if (c->dumps && (long)c->classes[cls].hr * c->classes[cls].wc > DUMP_NB_MAX) return fail(MOT_ERR_SHAPE, "stage dumps are limited to %ld cells", DUMP_NB_MAX);
        if (c->dumps) { c->dump_hr = c->classes[cls].hr; c->dump_wc = c->classes[cls].wc; c->dump_rows = c->meta_h[c->h_slots.p[a]].rows; c->dump_cols = c->meta_h[c->h_slots.p[a]].cols; if (b - a != 1 || n != 1) return fail(MOT_ERR_ARG, "stage dumps need a batch of exactly one job"); }
        const int rc = kcf_run(c, mode, c->classes[cls].hr, c->classes[cls].wc, L); if (rc) return rc;
        a = b;
    }
    if (read_back) {
        CU(cudaMemcpyAsync(c->h_boxes.p, c->d_boxes.p, sizeof(mot_bbox_t) * n, cudaMemcpyDeviceToHost, c->stream));
        CU(cudaStreamSynchronize(c->stream));
        for (int i = 0; i < n; ++i) boxes[order[i]] = c->h_boxes.p[i];
    } else {
        CU(cudaStreamSynchronize(c->stream));     // staging buffers are reused by the next call
    }
    if (c->h_any_err && *c->h_any_err) { *c->h_any_err = 0; return fail(MOT_ERR_SHAPE, "a KCF job did not fit the shared memory of its launch and was skipped (internal sizing error)"); }
    return 0;
}

// predict (+ clamp) followed at once by update with the predicted box -- what the reference does for a track without a detection
// (top/td.cpp:344-384 then :550-582) -- as ONE call: one staging kernel, predict and update launches back to back per window size,
// one synchronisation.  Host arrays; boxes in = crop rectangles, out = predicted (clamped) boxes.
static int kcf_track_host(mot_ctx_t *c, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp)
{
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    if (n == 0) return 0;
    if (!handles || !frame_slots || !boxes) return fail(MOT_ERR_ARG, "null array");
    if (c->dumps) return fail(MOT_ERR_ARG, "stage dumps are per call of mot_predict_batch / mot_update_batch");
    CU(cudaSetDevice(c->device));
    { const int rc = sync_frame_ptrs(c); if (rc) return rc; }
    { const int rc = wait_frames(c, n, frame_slots); if (rc) return rc; }
    std::vector<int> order(n);
    for (int i = 0; i < n; ++i) {
        const int s = handles[i];
        if (s < 0 || s >= c->max_tracks || !c->used[s]) return fail(MOT_ERR_ARG, "invalid handle %d", s);
        if (frame_slots[i] < 0 || frame_slots[i] >= c->n_frames || !c->frame_ptr_h[frame_slots[i]]) return fail(MOT_ERR_ARG, "frame slot %d is empty", frame_slots[i]);
        order[i] = i;
    }
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return c->meta_h[handles[a]].size_class < c->meta_h[handles[b]].size_class; });
    CU(c->h_slots.ensure(n)); CU(c->h_frames.ensure(n)); CU(c->h_boxes.ensure(n));
    CU(c->d_slots.ensure(n)); CU(c->d_frames.ensure(n)); CU(c->d_boxes.ensure(n));
    for (int i = 0; i < n; ++i) { c->h_slots.p[i] = handles[order[i]]; c->h_frames.p[i] = frame_slots[order[i]]; c->h_boxes.p[i] = boxes[order[i]]; }
    stage_in_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->d_slots.p, c->h_slots.p, c->d_frames.p, c->h_frames.p, c->d_boxes.p, c->h_boxes.p, n);
    CU(cudaGetLastError());
    c->launches += 1;
    for (int mode = KCF_MODE_PREDICT; mode <= KCF_MODE_UPDATE; ++mode)
        for (int a = 0; a < n;) {
            const int cls = c->meta_h[c->h_slots.p[a]].size_class;
            int b = a; while (b < n && c->meta_h[c->h_slots.p[b]].size_class == cls) ++b;
            KcfLaunch L; fill_launch(c, L, b - a, c->d_slots.p + a, c->d_frames.p + a, c->d_boxes.p + a, mode == KCF_MODE_PREDICT ? clamp : 0);
            const int rc = kcf_run(c, mode, c->classes[cls].hr, c->classes[cls].wc, L); if (rc) return rc;
            a = b;
        }
    CU(cudaMemcpyAsync(c->h_boxes.p, c->d_boxes.p, sizeof(mot_bbox_t) * n, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) boxes[order[i]] = c->h_boxes.p[i];
    if (c->h_any_err && *c->h_any_err) { *c->h_any_err = 0; return fail(MOT_ERR_SHAPE, "a KCF job did not fit the shared memory of its launch and was skipped (internal sizing error)"); }
    return 0;
}

static int kcf_batch_dev(mot_ctx_t *c, int mode, int n, const int *d_handles, const int *d_frame_slots, mot_bbox_t *d_boxes, int clamp)
{
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    { const int rc = sync_frame_ptrs(c); if (rc) return rc; }
    { const int rc = wait_frames(c, 0, nullptr); if (rc) return rc; }
    int cls = -1;
    for (size_t i = 0; i < c->classes.size(); ++i) if (c->classes[i].live > 0) { if (cls >= 0) return fail(MOT_ERR_SHAPE, "device-array batches need all live trackers to share one window size; use the host-array call"); cls = (int)i; }
    if (cls < 0) return fail(MOT_ERR_ARG, "no live trackers");
    KcfLaunch L; fill_launch(c, L, n, d_handles, d_frame_slots, d_boxes, clamp);
    L.dump = KcfDump{};
    return kcf_run(c, mode, c->classes[cls].hr, c->classes[cls].wc, L);
}

static int kalman_batch_host(mot_ctx_t *c, int mode, int n, const int *handles, mot_bbox_t *boxes, int clamp)
{
    if (n == 0) return 0;
    if (!handles || !boxes) return fail(MOT_ERR_ARG, "null array");
    CU(cudaSetDevice(c->device));
    for (int i = 0; i < n; ++i) if (handles[i] < 0 || handles[i] >= c->max_tracks || !c->used[handles[i]]) return fail(MOT_ERR_ARG, "invalid handle %d", handles[i]);
    CU(c->h_slots.ensure(n)); CU(c->h_boxes.ensure(n));
    memcpy(c->h_slots.p, handles, sizeof(int) * n); memcpy(c->h_boxes.p, boxes, sizeof(mot_bbox_t) * n);
    // The kernels read the slots and read / write the boxes in the PINNED staging arrays themselves (device-accessible under unified
    // addressing): 28 bytes per track each way over PCIe instead of three copy-engine transfers around a microsecond-sized kernel.
    int rc;
    if (mode == KCF_MODE_PREDICT) rc = kalman_predict(c->kal, n, c->h_slots.p, c->h_boxes.p, clamp, c->W, c->H, c->stream);
    else rc = kalman_update(c->kal, n, c->h_slots.p, c->h_boxes.p, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "Kalman launch failed (%d)", rc);
    c->launches += 1;
    CU(cudaStreamSynchronize(c->stream));
    if (mode == KCF_MODE_PREDICT) memcpy(boxes, c->h_boxes.p, sizeof(mot_bbox_t) * n);
    return 0;
}

int mot_predict_batch(mot_ctx_t *c, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp)
{
    if (!c || n < 0) return fail(MOT_ERR_ARG, "mot_predict_batch: bad argument");
    if (c->kind == MOT_TRACKER_KCF) return kcf_batch_host(c, KCF_MODE_PREDICT, n, handles, frame_slots, boxes, clamp, true);
    return kalman_batch_host(c, KCF_MODE_PREDICT, n, handles, boxes, clamp);
}

int mot_track_batch(mot_ctx_t *c, int n, const int *handles, const int *frame_slots, mot_bbox_t *boxes, int clamp)
{
    if (!c || n < 0) return fail(MOT_ERR_ARG, "mot_track_batch: bad argument");
    if (c->kind == MOT_TRACKER_KCF) return kcf_track_host(c, n, handles, frame_slots, boxes, clamp);
    // Kalman: the update of an unassigned track takes its own predicted box as the measurement (top/td.cpp:581)
    int rc = kalman_batch_host(c, KCF_MODE_PREDICT, n, handles, boxes, clamp);
    if (!rc) rc = kalman_batch_host(c, KCF_MODE_UPDATE, n, handles, boxes, 0);
    return rc;
}

int mot_update_batch(mot_ctx_t *c, int n, const int *handles, const int *frame_slots, const mot_bbox_t *boxes)
{
    if (!c || n < 0) return fail(MOT_ERR_ARG, "mot_update_batch: bad argument");
    if (c->kind == MOT_TRACKER_KCF) return kcf_batch_host(c, KCF_MODE_UPDATE, n, handles, frame_slots, const_cast<mot_bbox_t *>(boxes), 0, false);
    return kalman_batch_host(c, KCF_MODE_UPDATE, n, handles, const_cast<mot_bbox_t *>(boxes), 0);
}

int mot_predict_batch_dev(mot_ctx_t *c, int n, const int *d_handles, const int *d_frame_slots, mot_bbox_t *d_boxes, int clamp)
{
    if (!c || n < 0) return fail(MOT_ERR_ARG, "mot_predict_batch_dev: bad argument");
    if (c->kind == MOT_TRACKER_KCF) return kcf_batch_dev(c, KCF_MODE_PREDICT, n, d_handles, d_frame_slots, d_boxes, clamp);
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    const int rc = kalman_predict(c->kal, n, d_handles, d_boxes, clamp, c->W, c->H, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "Kalman launch failed (%d)", rc);
    c->launches += 1;
    return 0;
}

int mot_update_batch_dev(mot_ctx_t *c, int n, const int *d_handles, const int *d_frame_slots, const mot_bbox_t *d_boxes)
{
    if (!c || n < 0) return fail(MOT_ERR_ARG, "mot_update_batch_dev: bad argument");
    if (c->kind == MOT_TRACKER_KCF) return kcf_batch_dev(c, KCF_MODE_UPDATE, n, d_handles, d_frame_slots, const_cast<mot_bbox_t *>(d_boxes), 0);
    if (n == 0) return 0;
    CU(cudaSetDevice(c->device));
    const int rc = kalman_update(c->kal, n, d_handles, d_boxes, c->stream);
    if (rc) return fail(MOT_ERR_CUDA, "Kalman launch failed (%d)", rc);
    c->launches += 1;
    return 0;
}

static int kcf_gray(mot_ctx_t *c, int mode, int handle, const float *gray_host, mot_bbox_t *box)
{
    if (!c || !box) return fail(MOT_ERR_ARG, "null argument");
    if (c->kind != MOT_TRACKER_KCF) return kalman_batch_host(c, mode, 1, &handle, box, 0);     // Kalman ignores the patch (kalman.cpp:131-139)
    if (!gray_host) return fail(MOT_ERR_ARG, "null gray patch");
    if (handle < 0 || handle >= c->max_tracks || !c->used[handle]) return fail(MOT_ERR_ARG, "invalid handle %d", handle);
    CU(cudaSetDevice(c->device));
    const KcfMeta &m = c->meta_h[handle];
    const size_t npx = (size_t)m.rows * m.cols;
    CU(c->d_gray.ensure(npx)); CU(c->d_slots.ensure(1)); CU(c->d_boxes.ensure(1)); CU(c->h_boxes.ensure(1)); CU(c->h_slots.ensure(1));
    c->h_slots.p[0] = handle; c->h_boxes.p[0] = *box;
    CU(cudaMemcpyAsync(c->d_gray.p, gray_host, sizeof(float) * npx, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_slots.p, c->h_slots.p, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_boxes.p, c->h_boxes.p, sizeof(mot_bbox_t), cudaMemcpyHostToDevice, c->stream));
    KcfLaunch L; fill_launch(c, L, 1, c->d_slots.p, nullptr, c->d_boxes.p, 0);
    L.gray = c->d_gray.p; L.gray_stride = (long)npx;
    if (c->dumps && (long)m.hr * m.wc > DUMP_NB_MAX) return fail(MOT_ERR_SHAPE, "stage dumps are limited to %ld cells", DUMP_NB_MAX);
    if (c->dumps) { c->dump_hr = m.hr; c->dump_wc = m.wc; c->dump_rows = m.rows; c->dump_cols = m.cols; }
    const int rc = kcf_run(c, mode, m.hr, m.wc, L); if (rc) return rc;
    if (mode == KCF_MODE_PREDICT) CU(cudaMemcpyAsync(c->h_boxes.p, c->d_boxes.p, sizeof(mot_bbox_t), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (mode == KCF_MODE_PREDICT) *box = c->h_boxes.p[0];
    return 0;
}

int mot_predict_gray(mot_ctx_t *c, int handle, const float *gray_host, mot_bbox_t *box_out) { return kcf_gray(c, KCF_MODE_PREDICT, handle, gray_host, box_out); }
int mot_update_gray(mot_ctx_t *c, int handle, const float *gray_host, const mot_bbox_t *box) { mot_bbox_t b = *box; return kcf_gray(c, KCF_MODE_UPDATE, handle, gray_host, &b); }

// ---- association ---------------------------------------------------------------------------------------------------------
int mot_associate_batch_dev(mot_ctx_t *c, int n_mat, const int *d_T, const int *d_D, const mot_bbox_t *d_trk, long trk_stride,
                            const mot_bbox_t *d_det, long det_stride, int cost_mode, double *d_dist, long dist_stride,
                            int *d_assign, long assign_stride, double *d_cost, int max_dim)
{
    if (!c || n_mat < 0 || max_dim < 1 || max_dim > 1024) return fail(MOT_ERR_ARG, "mot_associate_batch_dev: bad argument (max_dim %d)", max_dim);
    if (n_mat == 0) return 0;
    CU(cudaSetDevice(c->device));
    CU(c->d_work.ensure((size_t)n_mat * max_dim * max_dim));
    return mot_ctx_associate_dev(c, n_mat, d_T, d_D, d_trk, trk_stride, d_det, det_stride, cost_mode, d_dist, dist_stride, d_assign, assign_stride, d_cost, max_dim, c->d_work.p);
}

int mot_associate_batch(mot_ctx_t *c, int n_mat, const int *T, const int *D, const mot_bbox_t *trk, long trk_stride,
                        const mot_bbox_t *det, long det_stride, int cost_mode, double *dist, long dist_stride,
                        int *assign, long assign_stride, double *cost)
{
    if (!c || n_mat < 0 || (n_mat && (!T || !D || !trk || !det || !assign))) return fail(MOT_ERR_ARG, "mot_associate_batch: bad argument");
    if (n_mat == 0) return 0;
    CU(cudaSetDevice(c->device));
    int md = 1; long nt = 0, nd = 0;
    for (int m = 0; m < n_mat; ++m) { md = std::max(md, std::max(T[m], D[m])); nt = std::max<long>(nt, T[m]); nd = std::max<long>(nd, D[m]); }
    if (md > 1024) return fail(MOT_ERR_SHAPE, "association problems larger than 1024 are not supported (%d)", md);
    nt = std::max<long>(nt, 1); nd = std::max<long>(nd, 1);
    const long ds = (long)md * md;
    CU(c->h_TD.ensure(2 * n_mat)); CU(c->d_TD.ensure(2 * n_mat));
    CU(c->h_trk.ensure(nt * n_mat)); CU(c->d_trk.ensure(nt * n_mat)); CU(c->h_det.ensure(nd * n_mat)); CU(c->d_det.ensure(nd * n_mat));
    CU(c->d_dist.ensure(ds * n_mat)); CU(c->d_assign.ensure((size_t)md * n_mat)); CU(c->d_cost.ensure(n_mat));
    CU(c->h_assign.ensure((size_t)md * n_mat)); CU(c->h_cost.ensure(n_mat));
    for (int m = 0; m < n_mat; ++m) {
        c->h_TD.p[m] = T[m]; c->h_TD.p[n_mat + m] = D[m];
        memcpy(c->h_trk.p + nt * m, trk + trk_stride * m, sizeof(mot_bbox_t) * T[m]);
        memcpy(c->h_det.p + nd * m, det + det_stride * m, sizeof(mot_bbox_t) * D[m]);
    }
    // in: one staging kernel reads the three pinned arrays (the boxes are read once per cost cell, so they are brought over first);
    // out: the solver writes assignments and totals straight into the pinned result arrays -- no copy-engine transfer either way
    {
        const long nb1 = nt * n_mat, nb2 = nd * n_mat, nmax = std::max<long>(std::max(nb1, nb2) * 3, 2L * n_mat);
        stage_assoc_kernel<<<(unsigned)((nmax + 255) / 256), 256, 0, c->stream>>>(reinterpret_cast<long long *>(c->d_trk.p), reinterpret_cast<const long long *>(c->h_trk.p), nb1 * 3,
                                                                               reinterpret_cast<long long *>(c->d_det.p), reinterpret_cast<const long long *>(c->h_det.p), nb2 * 3,
                                                                               c->d_TD.p, c->h_TD.p, 2L * n_mat);
        CU(cudaGetLastError());
        c->launches += 1;
    }
    const int rc = mot_associate_batch_dev(c, n_mat, c->d_TD.p, c->d_TD.p + n_mat, c->d_trk.p, nt, c->d_det.p, nd, cost_mode,
                                           c->d_dist.p, ds, c->h_assign.p, md, c->h_cost.p, md);
    if (rc) return rc;
    if (dist) { CU(c->h_dist.ensure(ds * n_mat)); CU(cudaMemcpyAsync(c->h_dist.p, c->d_dist.p, sizeof(double) * ds * n_mat, cudaMemcpyDeviceToHost, c->stream)); }
    CU(cudaStreamSynchronize(c->stream));
    for (int m = 0; m < n_mat; ++m) {
        const int nr = T[m] < D[m] ? T[m] : D[m];
        memcpy(assign + assign_stride * m, c->h_assign.p + (size_t)md * m, sizeof(int) * nr);
        if (cost) cost[m] = c->h_cost.p[m];
        if (dist) memcpy(dist + dist_stride * m, c->h_dist.p + ds * m, sizeof(double) * (size_t)T[m] * D[m]);
    }
    return 0;
}

int mot_assign_batch(mot_ctx_t *c, int n_mat, const int *nrows, const int *ncols, const double *dist, long dist_stride,
                     int *assign, long assign_stride, double *cost)
{
    if (!c || n_mat < 0 || (n_mat && (!nrows || !ncols || !dist || !assign))) return fail(MOT_ERR_ARG, "mot_assign_batch: bad argument");
    if (n_mat == 0) return 0;
    CU(cudaSetDevice(c->device));
    int md = 1;
    for (int m = 0; m < n_mat; ++m) md = std::max(md, std::max(nrows[m], ncols[m]));
    if (md > 1024) return fail(MOT_ERR_SHAPE, "association problems larger than 1024 are not supported (%d)", md);
    const long ds = (long)md * md;
    CU(c->h_TD.ensure(2 * n_mat)); CU(c->d_TD.ensure(2 * n_mat));
    CU(c->d_dist.ensure(ds * n_mat)); CU(c->h_dist.ensure(ds * n_mat)); CU(c->d_work.ensure(ds * n_mat));
    CU(c->d_assign.ensure((size_t)md * n_mat)); CU(c->d_cost.ensure(n_mat)); CU(c->h_assign.ensure((size_t)md * n_mat)); CU(c->h_cost.ensure(n_mat));
    for (int m = 0; m < n_mat; ++m) {
        c->h_TD.p[2 * m] = nrows[m]; c->h_TD.p[2 * m + 1] = ncols[m];
        memcpy(c->h_dist.p + ds * m, dist + dist_stride * m, sizeof(double) * (size_t)nrows[m] * ncols[m]);
    }
    CU(cudaMemcpyAsync(c->d_TD.p, c->h_TD.p, sizeof(int) * 2 * n_mat, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(c->d_dist.p, c->h_dist.p, sizeof(double) * ds * n_mat, cudaMemcpyHostToDevice, c->stream));
    AssocLaunch A{};
    A.n_mat = n_mat; A.dist = c->d_dist.p; A.dist_stride = ds; A.work = c->d_work.p; A.work_stride = ds;
    A.assign = c->d_assign.p; A.assign_stride = md; A.cost = c->d_cost.p; A.max_dim = md;
    const int rc = assoc_solve(A, c->d_TD.p, c->stream); if (rc) return fail(MOT_ERR_CUDA, "munkres kernel launch failed (%d)", rc);
    c->launches += 1;
    CU(cudaMemcpyAsync(c->h_assign.p, c->d_assign.p, sizeof(int) * (size_t)md * n_mat, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_cost.p, c->d_cost.p, sizeof(double) * n_mat, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    for (int m = 0; m < n_mat; ++m) {
        memcpy(assign + assign_stride * m, c->h_assign.p + (size_t)md * m, sizeof(int) * nrows[m]);
        if (cost) cost[m] = c->h_cost.p[m];
    }
    return 0;
}

// ---- stand-alone crop + gray + resize ----------------------------------------------------------------------------------------
int mot_crop_gray_resize(mot_ctx_t *c, int frame_slot, const mot_bbox_t *box, int rows_d, int cols_d, float *gray_host_out)
{
    // Served by the fused kernel's own front end through the stage dump, so there is exactly one implementation of it.
    if (!c || !box || !gray_host_out) return fail(MOT_ERR_ARG, "null argument");
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    mot_bbox_t tb{ 0, 0, rows_d - 1, cols_d - 1, 0, 0.f };
    int h = -1;
    int rc = mot_tracker_new_batch(c, 1, &tb, &h); if (rc) return rc;
    const bool was = c->dumps;
    rc = mot_debug_enable_dumps(c, 1);
    mot_bbox_t b = *box;
    if (!rc) rc = mot_predict_batch(c, 1, &h, &frame_slot, &b, 0);
    if (!rc) { const long got = mot_debug_fetch(c, 0, gray_host_out, (long)sizeof(float) * rows_d * cols_d); if (got < 0) rc = (int)got; }
    if (!was) mot_debug_enable_dumps(c, 0);
    mot_tracker_delete_batch(c, 1, &h);
    return rc;
}

// ---- test hooks ------------------------------------------------------------------------------------------------------------
int mot_debug_enable_dumps(mot_ctx_t *c, int enable)
{
    if (!c) return fail(MOT_ERR_ARG, "null ctx");
    if (c->kind != MOT_TRACKER_KCF) return fail(MOT_ERR_KIND, "context is not a KCF context");
    CU(cudaSetDevice(c->device));
    if (enable && !c->dump.gray) {
        const long cell = DUMP_NB_MAX, px = 16L * cell + 12L * 2 * cell + 64, spec = cell;     // sized for windows up to DUMP_NB_MAX cells
        KcfDump &d = c->dump;
        d.stride_px = px; d.stride_cell = cell; d.stride_spec = spec;
        CU(cudaMalloc(&d.gray, sizeof(float) * px)); CU(cudaMalloc(&d.m0, sizeof(float) * px)); CU(cudaMalloc(&d.bin, sizeof(int) * px));
        CU(cudaMalloc(&d.r1, sizeof(float) * cell * 18)); CU(cudaMalloc(&d.nrm, sizeof(float) * (3 * cell + 8)));
        CU(cudaMalloc(&d.feat, sizeof(float) * cell * KCF_CHAN)); CU(cudaMalloc(&d.spec, sizeof(float2) * spec * KCF_CHAN));
        CU(cudaMalloc(&d.zf, sizeof(float2) * spec)); CU(cudaMalloc(&d.resp, sizeof(float) * cell)); CU(cudaMalloc(&d.kf, sizeof(float) * spec));
        CU(cudaMalloc(&d.peak, sizeof(int) * 2)); CU(cudaMalloc(&d.margin, sizeof(float) * 2));
    }
    c->dumps = enable != 0;
    return 0;
}

long mot_debug_fetch(mot_ctx_t *c, int stage, void *host_out, long max_bytes)
{
    if (!c || !host_out || !c->dump.gray) return fail(MOT_ERR_ARG, "dumps are not enabled");
    cudaSetDevice(c->device);
    const long hr = c->dump_hr, wc = c->dump_wc, nb = hr * wc, S = wc * (hr / 2 + 1), px = (long)c->dump_rows * c->dump_cols, px0 = 16 * nb;
    const void *src = nullptr; long bytes = 0;
    switch (stage) {
    case 0: src = c->dump.gray; bytes = 4 * px; break;
    case 1: src = c->dump.m0; bytes = 4 * px0; break;
    case 2: src = c->dump.bin; bytes = 4 * px0; break;
    case 3: src = c->dump.r1; bytes = 4 * 18 * nb; break;
    case 4: src = c->dump.nrm; bytes = 4 * (wc + 1) * (hr + 1); break;
    case 5: src = c->dump.feat; bytes = 4 * KCF_CHAN * nb; break;
    case 6: src = c->dump.spec; bytes = 8 * KCF_CHAN * S; break;
    case 7: src = c->dump.zf; bytes = 8 * S; break;
    case 8: src = c->dump.resp; bytes = 4 * nb; break;
    case 9: src = c->dump.kf; bytes = 4 * S; break;
    case 10: src = c->dump.peak; bytes = 8; break;
    default: return fail(MOT_ERR_ARG, "unknown stage %d", stage);
    }
    if (bytes > max_bytes) return fail(MOT_ERR_ARG, "buffer too small (%ld > %ld)", bytes, max_bytes);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(MOT_ERR_CUDA, "dump fetch: %s", cudaGetErrorString(e));
    return bytes;
}

long mot_debug_state(mot_ctx_t *c, int handle, int which, void *host_out, long max_bytes)
{
    if (!c || !host_out || handle < 0 || handle >= c->max_tracks || !c->used[handle]) return fail(MOT_ERR_ARG, "mot_debug_state: bad argument");
    cudaSetDevice(c->device);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return fail(MOT_ERR_CUDA, "sync: %s", cudaGetErrorString(e));
    if (c->kind == MOT_TRACKER_KCF) {
        const KcfMeta &m = c->meta_h[handle];
        const long S = (long)m.wc * (m.hr / 2 + 1);
        if (which == 0) {
            const long bytes = 8 * KCF_CHAN * S; if (bytes > max_bytes) return fail(MOT_ERR_ARG, "buffer too small");
            e = cudaMemcpy(host_out, m.model_ptr ? m.model_ptr : c->d_model + (long)handle * c->model_stride, bytes, cudaMemcpyDeviceToHost);
            return e == cudaSuccess ? bytes : fail(MOT_ERR_CUDA, "%s", cudaGetErrorString(e));
        }
        if (which == 1) {
            const long bytes = 4 * S; if (bytes > max_bytes) return fail(MOT_ERR_ARG, "buffer too small");
            e = cudaMemcpy(host_out, m.alpha_ptr ? m.alpha_ptr : c->d_alpha + (long)handle * c->alpha_stride, bytes, cudaMemcpyDeviceToHost);
            return e == cudaSuccess ? bytes : fail(MOT_ERR_CUDA, "%s", cudaGetErrorString(e));
        }
        if (which == 4) {                               // extension: sub-cell refinement (vertical, horizontal) of the last predicted peak
            if (max_bytes < 8) return fail(MOT_ERR_ARG, "buffer too small");
            KcfMeta mt;
            e = cudaMemcpy(&mt, c->d_meta + handle, sizeof(KcfMeta), cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) return fail(MOT_ERR_CUDA, "%s", cudaGetErrorString(e));
            float v[2] = { mt.sub_dv, mt.sub_dh }; memcpy(host_out, v, 8);
            return 8;
        }
        if (which == 5) {                               // extension: imaginary part of alpha (Gaussian kernel)
            const long bytes = 4 * S; if (bytes > max_bytes || !c->d_alpha_im) return fail(MOT_ERR_ARG, "no complex alpha / buffer too small");
            e = cudaMemcpy(host_out, c->d_alpha_im + (long)handle * c->alpha_stride, bytes, cudaMemcpyDeviceToHost);
            return e == cudaSuccess ? bytes : fail(MOT_ERR_CUDA, "%s", cudaGetErrorString(e));
        }
        return fail(MOT_ERR_ARG, "unknown state %d for a KCF context", which);
    }
    if (which == 2 || which == 3) {
        const int n = which == 2 ? 6 : 36;
        if ((long)sizeof(double) * n > max_bytes) return fail(MOT_ERR_ARG, "buffer too small");
        const double *base = which == 2 ? c->kal.x : c->kal.P;
        e = cudaMemcpy2D(host_out, sizeof(double), base + handle, sizeof(double) * c->kal.cap, sizeof(double), n, cudaMemcpyDeviceToHost);
        return e == cudaSuccess ? (long)sizeof(double) * n : fail(MOT_ERR_CUDA, "%s", cudaGetErrorString(e));
    }
    return fail(MOT_ERR_ARG, "unknown state %d for a Kalman context", which);
}

// Host-only view of how the any-size kernel would run an hr x wc window (no GPU needed): out[0..15] = ok, strips, total shared-memory
// bytes, channels per tile, gradient strip width (pixel columns), cell columns per strip, spectral buffer (float2), A / B region floats,
// scratch bytes per CTA, CTAs per SM, threads per CTA, passes of the length-hr / length-wc transforms, reserved;
// radices[0..13] = the radices of the two transforms (7 slots each, zero padded).
int mot_debug_any_plan(int hr, int wc, int *out, int *radices)
{
    if (hr < 1 || wc < 1 || !out) return fail(MOT_ERR_ARG, "mot_debug_any_plan: bad argument");
    const FhogTables &t = fhog_tables();
    if (!t.ok) return fail(MOT_ERR_TABLES, "%s", t.error.c_str());
    int lut = 2 * (2 << t.rsqrt_bits) + ((2 * t.bin_nseg + 3) & ~3);
    if (lut > 8192) lut = 0;
    const AnyGeo g = any_geo(hr, wc, lut);
    int threads = 0, ctas = 0;
    if (g.ok) any_launch_shape((size_t)g.total * 4, &threads, &ctas);
    AnyPlan pr{}, pc{};
    any_plan(hr, pr); any_plan(wc, pc);
    const int v[16] = { g.ok, g.strips, g.total * 4, g.tc, g.xw, g.cs, g.xbuf, g.aF, g.bF, (int)kcf_any_scratch_bytes(hr, wc, lut), ctas, threads, pr.nf, pc.nf, 0, 0 };
    for (int i = 0; i < 16; ++i) out[i] = v[i];
    if (radices) for (int i = 0; i < 7; ++i) { radices[i] = pr.r[i]; radices[7 + i] = pc.r[i]; }
    return 0;
}

long mot_debug_tables(int which, float *out, long max_floats, int *info)
{
    const FhogTables &t = fhog_tables();
    if (!t.ok) return fail(MOT_ERR_TABLES, "%s", t.error.c_str());
    if (info) { info[0] = t.rsqrt_bits; info[1] = t.rcp_bits; info[2] = t.bin_shift; info[3] = t.bin_nseg; }
    if (which == 4 || which == 5) {                     // 32-bit entries handed out through the float buffer, bit for bit
        const std::vector<uint32_t> &u = which == 4 ? t.bin_tab : t.bin2_tab;
        if (out) { if ((long)u.size() > max_floats) return fail(MOT_ERR_ARG, "buffer too small"); memcpy(out, u.data(), sizeof(uint32_t) * u.size()); }
        return (long)u.size();
    }
    if (which == 6) {                                   // constants: bits of M2 up to which MIN(rsqrt(M2), 1e10f) saturates; rcp(1e10f)
        if (out) { if (max_floats < 2) return fail(MOT_ERR_ARG, "buffer too small"); memcpy(out, &t.u_cap, 4); memcpy(out + 1, &t.rcp_cap, 4); }
        return 2;
    }
    const std::vector<float> *v = which == 0 ? &t.rsqrt_tab : which == 1 ? &t.rcp_tab : which == 2 ? &t.acos_tab : which == 3 ? &t.rsrc_tab : nullptr;
    if (!v) return fail(MOT_ERR_ARG, "unknown table %d", which);
    if (out) { if ((long)v->size() > max_floats) return fail(MOT_ERR_ARG, "buffer too small"); memcpy(out, v->data(), sizeof(float) * v->size()); }
    return (long)v->size();
}

}  // extern "C"
