// yolo_post.cu -- YOLO3 post-processing on the device: network outputs of one image -> the tracker's detections.
//
// Replaces decode_netout, correct_yolo_boxes, sort, do_nms (detectors/yolo3.cpp:141-356) and the per-image driver inside
// tensorRunB (:487-527).  The detections are integer boxes made by thresholding, sorting and truncating single-precision
// values, so every floating-point operation is done exactly as the reference's compiler does it: this file is compiled
// with --fmad=false, the logistic / exponential use the C library's expf algorithm restated in FP64 (glibc: 32-entry table
// of 2^(i/32) and a cubic in double; oracle/port/port_expf.c pins the recipe against the library), and the order-dependent
// parts keep the reference's order:
//   * candidates are numbered the way the reference's loops push them (scale, cell, anchor, class) and gathered by rank;
//   * per class the reference's exchange sort is reproduced (by rank when all scores differ, literally otherwise);
//   * `is_suppressed` is never cleared between classes (:287, :300-303): flags raised for one class stay raised at the
//     same list positions for all later classes -- kept;
//   * the output order is class by class, descending score, clipped to the image (:505-526).
// One CTA per image for everything after the decode; this is a correctness-first kernel (a few thousand candidates).
#include "mot_internal.h"
#include "yolo_post.h"

namespace mot {

constexpr int YOLO_CAP = 4096;           // candidates per image
constexpr int YOLO_THREADS = 1024;

__constant__ unsigned long long c_exp2f_tab[32] = {      // bits(2^(i/32)) - ((i << 52) / 32): glibc e_exp2f_data.c
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,
};

// expf as glibc >= 2.27 evaluates it (sysdeps/ieee754/flt-32/e_expf.c), in separate IEEE double operations
// x86 cvttss2si: truncation toward zero, and INT_MIN for NaN or |v| >= 2^31
__device__ __forceinline__ int cvttss2si(float v) { return (v >= -2147483648.0f && v < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000; }

__device__ float expf_glibc(float x)
{
    const double N = 32.0;
    const double InvLn2N = 0x1.71547652b82fep+0 * N, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / N / N / N, C1 = 0x1.ebfce50fac4f3p-3 / N / N, C2 = 0x1.62e42ff0c52d6p-1 / N;
    if (x != x) return x;
    if (x > 0x1.62e42ep6f) return __int_as_float(0x7F800000);
    if (x < -0x1.9fe368p6f) return 0.0f;
    double z = __dmul_rn(InvLn2N, (double)x);
    double kd = __dadd_rn(z, SHIFT);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    const unsigned long long t = c_exp2f_tab[ki % 32] + (ki << (52 - 5));
    const double s = __longlong_as_double((long long)t);
    z = __dadd_rn(__dmul_rn(C0, r), C1);
    const double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(z, r2), y);
    y = __dmul_rn(y, s);
    return (float)y;
}

__device__ __forceinline__ float act(const float *o, int k, int per)          // yolo3.cpp:157-170
{
    const int r = k % per;
    if (r == 2 || r == 3) return expf_glibc(o[k]);
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_glibc(-o[k])));
}

struct YoloCand { unsigned key; float x, y, u, w, s; int c; };

// one thread per (cell, anchor) of one scale: candidates with score >= obj_thresh, numbered in the reference's push order
__global__ void yolo_decode_kernel(const float *out, int gh, int gw, int nc, const int *anch, float obj_thresh, int th, int tw,
                                   unsigned key_base, YoloCand *cand, int *count)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= gh * gw * 3) return;
    const int cellidx = t / 3, b = t - cellidx * 3, per = 5 + nc;
    const int row = cellidx / gw, col = cellidx - row * gw;
    const float *cell = out + (long)cellidx * 3 * per;
    const float objectness = act(cell, b * per + 4, per);
    for (int j = 0; j < nc; ++j) {
        const float scores = __fmul_rn(act(cell, b * per + 5 + j, per), objectness);
        if (scores >= obj_thresh) {
            const int at = atomicAdd(count, 1);
            if (at < YOLO_CAP) {
                YoloCand c;
                c.key = key_base + (unsigned)(t * nc + j);
                c.x = __fdiv_rn(__fadd_rn((float)col, act(cell, b * per + 0, per)), (float)gw);
                c.y = __fdiv_rn(__fadd_rn((float)row, act(cell, b * per + 1, per)), (float)gh);
                c.u = __fdiv_rn(__fmul_rn((float)anch[2 * b + 0], act(cell, b * per + 2, per)), (float)tw);
                c.w = __fdiv_rn(__fmul_rn((float)anch[2 * b + 1], act(cell, b * per + 3, per)), (float)th);
                c.s = scores; c.c = j;
                cand[at] = c;
            }
        }
    }
}

struct YoloDet { int xmin, ymin, xmax, ymax, cls; float s; };

__global__ void __launch_bounds__(YOLO_THREADS) yolo_nms_kernel(const YoloCand *cand, const int *count, int nc, float nms_thresh,
                                                                int th, int tw, int ih, int iw, mot_bbox_t *out, int max_out, int *n_out)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    YoloDet *det = reinterpret_cast<YoloDet *>(smraw);                       // [YOLO_CAP] sorted by (class, push order)
    int *idx = reinterpret_cast<int *>(det + YOLO_CAP);                      // [YOLO_CAP] sort result of the current class
    unsigned char *flag = reinterpret_cast<unsigned char *>(idx + YOLO_CAP); // [YOLO_CAP] is_suppressed, never cleared between classes
    __shared__ int seg[1025];                                                // class segments (nc <= 1024)
    __shared__ int tie, nbox;
    const int tid = threadIdx.x;
    const int n = min(*count, YOLO_CAP);
    if (tid == 0) { nbox = 0; }
    for (int i = tid; i <= nc; i += YOLO_THREADS) seg[i] = 0;
    for (int i = tid; i < YOLO_CAP; i += YOLO_THREADS) flag[i] = 0;
    __syncthreads();
    if (n == 0) { if (tid == 0) *n_out = 0; return; }
    // ---- correct_yolo_boxes (:203-251) + gather by (class, push order) -------------------------------------------------------
    float new_w, new_h;
    if (__fdiv_rn((float)tw, (float)iw) < __fdiv_rn((float)th, (float)ih)) { new_w = (float)tw; new_h = roundf(__fdiv_rn(__fmul_rn((float)ih, (float)tw), (float)iw)); }
    else { new_h = (float)th; new_w = roundf(__fdiv_rn(__fmul_rn((float)iw, (float)th), (float)ih)); }
    const float x_offset = (float)__ddiv_rn(__ddiv_rn((double)__fsub_rn((float)tw, new_w), 2.0), (double)tw);
    const float x_scale = __fdiv_rn(new_w, (float)tw);
    const float y_offset = (float)__ddiv_rn(__ddiv_rn((double)__fsub_rn((float)th, new_h), 2.0), (double)th);
    const float y_scale = __fdiv_rn(new_h, (float)th);
    for (int i = tid; i < n; i += YOLO_THREADS) atomicAdd(&seg[cand[i].c + 1], 1);
    __syncthreads();
    if (tid == 0) for (int c = 0; c < nc; ++c) seg[c + 1] += seg[c];          // seg[c] = first position of class c
    __syncthreads();
    for (int i = tid; i < n; i += YOLO_THREADS) {
        const YoloCand ci = cand[i];
        int rank = 0;                                                        // candidates of the same class pushed earlier
        for (int k = 0; k < n; ++k) { const YoloCand ck = cand[k]; rank += (ck.c == ci.c && ck.key < ci.key) ? 1 : 0; }
        const float x = __fmul_rn(__fdiv_rn(__fsub_rn(ci.x, x_offset), x_scale), (float)iw);
        const float y = __fmul_rn(__fdiv_rn(__fsub_rn(ci.y, y_offset), y_scale), (float)ih);
        const float w = __fmul_rn(__fdiv_rn(ci.u, x_scale), (float)iw);
        const float h = __fmul_rn(__fdiv_rn(ci.w, y_scale), (float)ih);
        YoloDet d;
        // (int) of a float on the reference's x86 build is cvttss2si: out-of-range, infinite and NaN values all give INT_MIN
        // ("integer indefinite"), whereas CUDA's conversion saturates (+inf -> INT_MAX, NaN -> 0).  A box whose raw t_w / t_h
        // overflows expf must come out as the reference's, so the x86 result is reproduced.
        d.xmin = cvttss2si(__fsub_rn(x, __fdiv_rn(w, 2.0f))); d.xmax = cvttss2si(__fadd_rn(x, __fdiv_rn(w, 2.0f)));
        d.ymin = cvttss2si(__fsub_rn(y, __fdiv_rn(h, 2.0f))); d.ymax = cvttss2si(__fadd_rn(y, __fdiv_rn(h, 2.0f)));
        d.cls = ci.c; d.s = ci.s;
        det[seg[ci.c] + rank] = d;
    }
    __syncthreads();
    // ---- per class: sort (:253-276), NMS (:305-338), output (:340-354 + :505-526) -----------------------------------------------
    for (int c = 0; c < nc; ++c) {
        const int s0 = seg[c], m = seg[c + 1] - s0;
        if (m == 0) continue;                                               // block-uniform
        const YoloDet *cb = det + s0;
        if (tid == 0) tie = 0;
        __syncthreads();
        for (int i = tid; i < m; i += YOLO_THREADS) {
            int rank = 0, eq = 0;
            for (int k = 0; k < m; ++k) { rank += (cb[k].s > cb[i].s) ? 1 : 0; eq += (cb[k].s == cb[i].s) ? 1 : 0; }
            if (eq > 1) tie = 1;
            idx[rank] = i;                                                  // valid when all scores differ
        }
        __syncthreads();
        if (tie) {
            if (tid == 0) {                                                 // the reference's exchange sort, literally
                for (int i = 0; i < m; ++i) idx[i] = i;
                for (int i = 0; i < m; ++i)
                    for (int j = i + 1; j < m; ++j)
                        if (cb[idx[j]].s > cb[idx[i]].s) { const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
            }
            __syncthreads();
        }
        for (int i = 0; i < m; ++i) {
            const int bi = idx[i];
            if (!flag[bi]) {                                                // block-uniform (shared flag, barrier below)
                const YoloDet b = cb[bi];
                for (int j = i + 1 + tid; j < m; j += YOLO_THREADS) {
                    const YoloDet a = cb[idx[j]];
                    const float maxX = (float)min(a.xmax, b.xmax), maxY = (float)min(a.ymax, b.ymax);
                    const float minX = (float)max(a.xmin, b.xmin), minY = (float)max(a.ymin, b.ymin);
                    const float oW = __fadd_rn(__fsub_rn(maxX, minX), 1.0f), oH = __fadd_rn(__fsub_rn(maxY, minY), 1.0f);
                    if ((oW > 0) & (oH > 0)) {
                        const float a1 = (float)((a.xmax - a.xmin + 1) * (a.ymax - a.ymin + 1));
                        const float a2 = (float)((b.xmax - b.xmin + 1) * (b.ymax - b.ymin + 1));
                        const float inter = __fmul_rn(oW, oH);
                        const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a1, a2), inter));
                        if (iou > nms_thresh) flag[idx[j]] = 1;
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) {
            for (int i = 0; i < m; ++i) {
                if (flag[idx[i]]) continue;
                YoloDet d = cb[idx[i]];
                d.ymin = max(d.ymin, 0); d.xmin = max(d.xmin, 0); d.ymax = min(d.ymax, ih - 1); d.xmax = min(d.xmax, iw - 1);
                if (d.ymin > d.ymax || d.xmin > d.xmax || d.ymin < 0 || d.xmin < 0 || d.xmax >= iw || d.ymax >= ih) continue;
                if (nbox >= max_out) break;
                mot_bbox_t o; o.t = d.ymin; o.l = d.xmin; o.b = d.ymax; o.r = d.xmax; o.type = d.cls; o.score = d.s;
                out[nbox++] = o;
            }
        }
        __syncthreads();
    }
    if (tid == 0) *n_out = nbox;
}

size_t yolo_nms_smem_bytes() { return (size_t)YOLO_CAP * (sizeof(YoloDet) + sizeof(int) + 1); }
int yolo_cap() { return YOLO_CAP; }

int yolo_post_launch(const float *const d_out[3], const int *d_anchors, float obj_thresh, float nms_thresh, int th, int tw, int ih, int iw, int nc,
                     void *d_cand, int *d_count, mot_bbox_t *d_boxes, int max_out, int *d_nout, cudaStream_t s)
{
    const int gh = th / 32, gw = tw / 32;
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return (int)e;
    unsigned key_base = 0;
    for (int k = 0; k < 3; ++k) {                                           // anchors + 12, + 6, + 0 (yolo3.cpp:496-498)
        const int g_h = gh << k, g_w = gw << k, nthr = g_h * g_w * 3;
        yolo_decode_kernel<<<(nthr + 127) / 128, 128, 0, s>>>(d_out[k], g_h, g_w, nc, d_anchors + (12 - 6 * k), obj_thresh, th, tw, key_base,
                                                           reinterpret_cast<YoloCand *>(d_cand), d_count);
        key_base += (unsigned)(nthr * nc);
    }
    e = cudaFuncSetAttribute(yolo_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)yolo_nms_smem_bytes());
    if (e != cudaSuccess) return (int)e;
    yolo_nms_kernel<<<1, YOLO_THREADS, yolo_nms_smem_bytes(), s>>>(reinterpret_cast<const YoloCand *>(d_cand), d_count, nc, nms_thresh, th, tw, ih, iw,
                                                                   d_boxes, max_out, d_nout);
    return (int)cudaGetLastError();
}

size_t yolo_cand_bytes() { return sizeof(YoloCand) * (size_t)YOLO_CAP; }

}  // namespace mot
