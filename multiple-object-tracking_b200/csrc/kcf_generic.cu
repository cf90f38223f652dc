// kcf_generic.cu -- the KCF path for ANY window size (FFTW accepts any n; detections come in any size).
//
// Same stages and the same reference arithmetic as the fused kernels (kcf_fused.cuh), but unfused: one small kernel
// per stage, intermediates in a per-job scratch area in global memory, and the 2-D transforms evaluated as separable
// DFTs straight from the definition with FP64 accumulation (O(n^2) per dimension; no radix restriction, more accurate
// than an f32 FFT).  This is the correctness path for sizes that have no register-FFT instantiation; it is not tuned.
//   gray      rgb2Gray + bilinearInterpolationGray          top/drawlib.c:192-240, 542-637
//   grad      grad1 / gradMag / gradQuantize                libhog/gradientMex.cpp:15-145
//   hist      gradHist (softBin<0 branch) + energies        libhog/gradientMex.cpp:183-230, 308-309
//   norm      hogNormMatrix                                 libhog/gradientMex.cpp:236-253
//   feat      hogChannels types 1/2 x cos_win               libhog/gradientMex.cpp:256-280, trackers/kcf.cpp:245-259
//   rowdft / coldft   fftwf r2c 2-D                         trackers/kcf.cpp:261-267
//   coldft epilogue + chansum   zf / kf / alpha / model     trackers/kcf.cpp:269-395
//   icol / irow / peak          c2r, argmax, box shift      trackers/kcf.cpp:397-428, top/td.cpp:378-381
#include "mot_internal.h"
#include "fhog_common.cuh"

namespace mot {

struct GenGeo {
    int hr, wc, nb, sk, S, rmax, cmax;
    size_t o_gray, o_m0, o_bin, o_r1, o_e, o_n, o_feat, o_row, o_prod, o_zf, o_resp, job_bytes;
};

static GenGeo make_geo(int hr, int wc)
{
    GenGeo g{};
    g.hr = hr; g.wc = wc; g.nb = hr * wc; g.sk = hr / 2 + 1; g.S = wc * g.sk; g.rmax = 4 * hr + 3; g.cmax = 4 * wc + 3;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
    g.o_gray = take(sizeof(float) * g.rmax * g.cmax);
    g.o_m0 = take(sizeof(float) * 16 * g.nb);
    g.o_bin = take(16 * g.nb);
    g.o_r1 = take(sizeof(float) * 18 * g.nb);
    g.o_e = take(sizeof(float) * g.nb);
    g.o_n = take(sizeof(float) * (hr + 1) * (wc + 1));
    g.o_feat = take(sizeof(float) * 31 * g.nb);
    g.o_row = take(sizeof(float2) * 31 * g.S);
    g.o_prod = take(sizeof(float2) * 31 * g.S);
    g.o_zf = take(sizeof(float2) * g.S);
    g.o_resp = take(sizeof(float) * g.nb);
    g.job_bytes = o;
    return g;
}

struct GenArgs {
    KcfLaunch p;
    GenGeo g;
    char *scratch;
    int job0;                  // first job of this chunk
};

#define JOB_PROLOGUE                                                     \
    const int jl = blockIdx.y;                                           \
    const int job = a.job0 + jl;                                         \
    char *const sc = a.scratch + (size_t)jl * a.g.job_bytes;             \
    const int slot = a.p.slots[job];                                     \
    KcfMeta *const meta = a.p.meta + slot;                               \
    const int rows = meta->rows, cols = meta->cols;                      \
    const int hr = a.g.hr, wc = a.g.wc;                                  \
    (void)sc; (void)rows; (void)cols; (void)hr; (void)wc;

__global__ void gen_gray(const GenArgs a)
{
    JOB_PROLOGUE
    float *gray = reinterpret_cast<float *>(sc + a.g.o_gray);              // column-major rows x cols
    const KcfLaunch &p = a.p;
    if (p.gray) {
        const float *src = p.gray + (long)job * p.gray_stride;
        for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < rows * cols; k += gridDim.x * blockDim.x) gray[k] = src[k];
        return;
    }
    const uint8_t *frame = p.frame_ptr[p.frames[job]];
    const mot_bbox_t box = p.boxes[job];
    int l = box.l, t = box.t, r = box.r, b = box.b;
    if (t > b) { const int q = t; t = b; b = q; }
    if (l > r) { const int q = l; l = r; r = q; }
    const int rows_s = b - t + 1, cols_s = r - l + 1, Wm = p.frame_w - 1, Hm = p.frame_h - 1;
    const float xs = __fdiv_rn((float)cols_s, (float)cols), ys = __fdiv_rn((float)rows_s, (float)rows);
    const bool same = rows_s == rows && cols_s == cols;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < rows * cols; k += gridDim.x * blockDim.x) {
        if (same) {
            const int x = k / rows, y = k - x * rows;
            gray[k] = bgr_gray(frame + (long)clampi(t + y, 0, Hm) * p.frame_stride + clampi(l + x, 0, Wm) * 3);
            continue;
        }
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
            const int scx = sidx[q] / rows_s, sr = sidx[q] - scx * rows_s;
            c[q] = bgr_gray(frame + (long)clampi(t + sr, 0, Hm) * p.frame_stride + clampi(l + scx, 0, Wm) * 3);
        }
        const float l0 = __fadd_rn(__fmul_rn(ifx, c[0]), __fmul_rn(fx, c[1]));
        const float l1 = __fadd_rn(__fmul_rn(ifx, c[2]), __fmul_rn(fx, c[3]));
        gray[k] = __fadd_rn(__fmul_rn(ify, l0), __fmul_rn(fy, l1));      // linear index k is the column-major template element
    }
}

__global__ void gen_grad(const GenArgs a)
{
    JOB_PROLOGUE
    const float *gray = reinterpret_cast<const float *>(sc + a.g.o_gray);
    float *m0 = reinterpret_cast<float *>(sc + a.g.o_m0);
    unsigned char *bins = reinterpret_cast<unsigned char *>(sc + a.g.o_bin);
    const int h0 = 4 * hr, w0 = 4 * wc;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < h0 * w0; idx += gridDim.x * blockDim.x) {
        const int x = idx / h0, y = idx - x * h0;
        const float *g = gray + x * rows + y;
        const int xm = x > 0 ? -rows : 0, xp = x < cols - 1 ? rows : 0, ym = y > 0 ? -1 : 0, yp = y < rows - 1 ? 1 : 0;
        const float rx = (x == 0 || x == cols - 1) ? 1.f : .5f, ry = (y == 0 || y == rows - 1) ? 1.f : .5f;
        const float gx = __fmul_rn(__fsub_rn(g[xp], g[xm]), rx), gy = __fmul_rn(__fsub_rn(g[yp], g[ym]), ry);
        int bb;
        m0[idx] = grad_pixel(gx, gy, a.p.tab.rsqrt_tab, a.p.tab.rcp_tab, a.p.tab.bin_tab, a.p.tab, &bb);
        bins[idx] = (unsigned char)bb;
    }
}

__global__ void gen_hist(const GenArgs a)
{
    JOB_PROLOGUE
    const float *m0 = reinterpret_cast<const float *>(sc + a.g.o_m0);
    const unsigned char *bins = reinterpret_cast<const unsigned char *>(sc + a.g.o_bin);
    float *r1 = reinterpret_cast<float *>(sc + a.g.o_r1);                  // [18][wc][hr]
    float *E = reinterpret_cast<float *>(sc + a.g.o_e);
    const int h0 = 4 * hr, w0 = 4 * wc, nb = a.g.nb;
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < nb; cell += gridDim.x * blockDim.x) {
        const int cx = cell / hr, cy = cell - cx * hr;
        float acc[18];
#pragma unroll
        for (int o = 0; o < 18; ++o) acc[o] = 0.f;
        for (int dx = 0; dx < 8; ++dx) {
            const int px = 4 * cx - 2 + dx;
            if (px < 0 || px >= w0) continue;
            const float wxv = 0.125f + 0.25f * (float)(dx < 4 ? dx : 7 - dx);
            for (int dy = 0; dy < 8; ++dy) {
                const int py = 4 * cy - 2 + dy;
                if (py < 0 || py >= h0) continue;
                const float w = wxv * (0.125f + 0.25f * (float)(dy < 4 ? dy : 7 - dy));
                const float v = __fmul_rn(w, m0[px * h0 + py]);
                const int bb = bins[px * h0 + py];
#pragma unroll
                for (int o = 0; o < 18; ++o) if (o == bb) acc[o] = __fadd_rn(acc[o], v);     // keeps acc[] in registers
            }
        }
        float e = 0.f;
#pragma unroll
        for (int o = 0; o < 18; ++o) {
            float v = acc[o];
            if (cx == 0) v = __fmul_rn(v, 8.f / 7.f);
            if (cy == 0) v = __fmul_rn(v, 8.f / 7.f);
            if (cx == wc - 1) v = __fmul_rn(v, 8.f / 7.f);
            if (cy == hr - 1) v = __fmul_rn(v, 8.f / 7.f);
            acc[o] = v; r1[o * nb + cell] = v;
        }
#pragma unroll
        for (int o = 0; o < 9; ++o) { const float r2 = __fadd_rn(acc[o], acc[o + 9]); e = __fadd_rn(e, __fmul_rn(r2, r2)); }
        E[cell] = e;
    }
}

__global__ void gen_norm(const GenArgs a)
{
    JOB_PROLOGUE
    const float *E = reinterpret_cast<const float *>(sc + a.g.o_e);
    float *N = reinterpret_cast<float *>(sc + a.g.o_n);
    const int n = (hr + 1) * (wc + 1);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int X = i / (hr + 1), Y = i - X * (hr + 1);
        const int x = clampi(X, 1, wc - 1) - 1, y = clampi(Y, 1, hr - 1) - 1;
        const float eps = 1e-4f / 4 / 4 / 4 / 4 / 4;
        float e = __fadd_rn(E[x * hr + y], E[x * hr + y + 1]);
        e = __fadd_rn(e, E[(x + 1) * hr + y]);
        e = __fadd_rn(e, E[(x + 1) * hr + y + 1]);
        e = __fadd_rn(e, eps);
        N[i] = __fdiv_rn(1.0f, __fsqrt_rn(e));
    }
}

__global__ void gen_feat(const GenArgs a)
{
    JOB_PROLOGUE
    const float *r1 = reinterpret_cast<const float *>(sc + a.g.o_r1);
    const float *N = reinterpret_cast<const float *>(sc + a.g.o_n);
    float *feat = reinterpret_cast<float *>(sc + a.g.o_feat);             // [31][wc][hr]
    const KcfClassDev cls = a.p.classes[meta->size_class];
    const int nb = a.g.nb;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 31 * nb; idx += gridDim.x * blockDim.x) {
        const int c = idx / nb, cell = idx - c * nb, j = cell / hr, i = cell - j * hr;
        const float *n0 = N + j * (hr + 1) + i, *n1 = n0 + (hr + 1);
        float h;
        if (c < 27) {
            const float rv = (c < 18) ? r1[c * nb + cell] : __fadd_rn(r1[(c - 18) * nb + cell], r1[(c - 9) * nb + cell]);
            h = __fmul_rn(fminf(__fmul_rn(rv, n1[1]), 0.2f), .5f);
            h = __fadd_rn(h, __fmul_rn(fminf(__fmul_rn(rv, n1[0]), 0.2f), .5f));
            h = __fadd_rn(h, __fmul_rn(fminf(__fmul_rn(rv, n0[1]), 0.2f), .5f));
            h = __fadd_rn(h, __fmul_rn(fminf(__fmul_rn(rv, n0[0]), 0.2f), .5f));
        } else {
            const int blk = c - 27;
            const float nv = (blk == 0) ? n1[1] : (blk == 1) ? n1[0] : (blk == 2) ? n0[1] : n0[0];
            h = 0.f;
            for (int o = 0; o < 18; ++o) h = __fadd_rn(h, __fmul_rn(fminf(__fmul_rn(r1[o * nb + cell], nv), 0.2f), .2357f));
        }
        feat[idx] = __fmul_rn(h, __fmul_rn(cls.wy[i], cls.wx[j]));
    }
}

// row pass: R[c][j][k] = sum_i f[c][j][i] exp(-2 pi i ik/hr), k <= hr/2
__global__ void gen_rowdft(const GenArgs a)
{
    JOB_PROLOGUE
    const float *feat = reinterpret_cast<const float *>(sc + a.g.o_feat);
    float2 *row = reinterpret_cast<float2 *>(sc + a.g.o_row);             // [31][wc][sk]
    const KcfClassDev cls = a.p.classes[meta->size_class];
    const int sk = a.g.sk, S = a.g.S;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 31 * S; idx += gridDim.x * blockDim.x) {
        const int cj = idx / sk, k = idx - cj * sk;                        // cj = c*wc + j
        const float *f = feat + (size_t)cj * hr;
        double sr = 0.0, si = 0.0;
        int t = 0;
        for (int i = 0; i < hr; ++i) {
            const double2 w = cls.tw_hr[t];
            sr += (double)f[i] * w.x; si += (double)f[i] * w.y;
            t += k; if (t >= hr) t -= hr;
        }
        row[idx] = make_float2((float)sr, (float)si);
    }
}

// column pass + spectral work: X[c][j'][k] = sum_j R[c][j][k] exp(-2 pi i j j'/wc)
template <int MODE> __global__ void gen_coldft(const GenArgs a)
{
    JOB_PROLOGUE
    const float2 *row = reinterpret_cast<const float2 *>(sc + a.g.o_row);
    float2 *prod = reinterpret_cast<float2 *>(sc + a.g.o_prod);           // [31][wc][sk]
    const KcfClassDev cls = a.p.classes[meta->size_class];
    float2 *model = meta->model_ptr ? meta->model_ptr : a.p.model + (long)slot * a.p.model_stride;
    const int sk = a.g.sk, S = a.g.S;
    const bool first = (MODE == KCF_MODE_UPDATE) && meta->first_update != 0;
    const float fac = first ? 1.0f : a.p.factor, omf = __fsub_rn(1.0f, fac);
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 31 * S; idx += gridDim.x * blockDim.x) {
        const int c = idx / S, rem = idx - c * S, jp = rem / sk, k = rem - jp * sk;
        const float2 *src = row + (size_t)c * S + k;
        double sr = 0.0, si = 0.0;
        int t = 0;
        for (int j = 0; j < wc; ++j) {
            const double2 w = cls.tw_wc[t];
            const float2 v = src[(size_t)j * sk];
            sr += (double)v.x * w.x - (double)v.y * w.y; si += (double)v.x * w.y + (double)v.y * w.x;
            t += jp; if (t >= wc) t -= wc;
        }
        const float2 v = make_float2((float)sr, (float)si);
        if (a.p.dump.spec) a.p.dump.spec[idx] = v;
        if (MODE == KCF_MODE_PREDICT) {
            const float2 m = model[idx];
            prod[idx] = make_float2(v.x * m.x + v.y * m.y, v.y * m.x - v.x * m.y);
        } else {
            prod[idx] = make_float2(v.x * v.x + v.y * v.y, 0.f);
            if (first) model[idx] = v;
            else { const float2 m = model[idx]; model[idx] = make_float2(__fadd_rn(__fmul_rn(omf, m.x), __fmul_rn(fac, v.x)), __fadd_rn(__fmul_rn(omf, m.y), __fmul_rn(fac, v.y))); }
        }
    }
}

template <int MODE> __global__ void gen_chansum(const GenArgs a)
{
    JOB_PROLOGUE
    const float2 *prod = reinterpret_cast<const float2 *>(sc + a.g.o_prod);
    float2 *zf = reinterpret_cast<float2 *>(sc + a.g.o_zf);
    const KcfClassDev cls = a.p.classes[meta->size_class];
    float *alpha = meta->alpha_ptr ? meta->alpha_ptr : a.p.alpha + (long)slot * a.p.alpha_stride;
    const int S = a.g.S;
    const bool first = (MODE == KCF_MODE_UPDATE) && meta->first_update != 0;
    const float fac = first ? 1.0f : a.p.factor, omf = __fsub_rn(1.0f, fac);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < S; e += gridDim.x * blockDim.x) {
        float2 acc = prod[e];
        for (int c = 1; c < 31; ++c) { const float2 v = prod[(size_t)c * S + e]; acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); }
        if (MODE == KCF_MODE_PREDICT) {
            const float al = alpha[e];
            acc.x = __fmul_rn(__fmul_rn(acc.x, al), cls.norm); acc.y = __fmul_rn(__fmul_rn(acc.y, al), cls.norm);
            zf[e] = acc;
            if (a.p.dump.zf) a.p.dump.zf[e] = acc;
        } else {
            const float kf = __fmul_rn(acc.x, cls.norm);
            if (a.p.dump.kf) a.p.dump.kf[e] = kf;
            const float an = __fdiv_rn(cls.yf_re[e], __fadd_rn(kf, a.p.lamda));
            alpha[e] = __fadd_rn(__fmul_rn(omf, alpha[e]), __fmul_rn(fac, an));
        }
    }
}

// inverse column pass: Y[j][k] = sum_j' zf[j'][k] exp(+2 pi i j j'/wc), written to the (now free) row buffer
__global__ void gen_icol(const GenArgs a)
{
    JOB_PROLOGUE
    const float2 *zf = reinterpret_cast<const float2 *>(sc + a.g.o_zf);
    float2 *Y = reinterpret_cast<float2 *>(sc + a.g.o_row);
    const KcfClassDev cls = a.p.classes[meta->size_class];
    const int sk = a.g.sk, S = a.g.S;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < S; idx += gridDim.x * blockDim.x) {
        const int j = idx / sk, k = idx - j * sk;
        double sr = 0.0, si = 0.0;
        int t = 0;
        for (int jp = 0; jp < wc; ++jp) {
            const double2 w = cls.tw_wc[t];                                // conj: exp(+i)
            const float2 v = zf[(size_t)jp * sk + k];
            sr += (double)v.x * w.x + (double)v.y * w.y; si += (double)v.y * w.x - (double)v.x * w.y;
            t += j; if (t >= wc) t -= wc;
        }
        Y[idx] = make_float2((float)sr, (float)si);
    }
}

// c2r along the rows: r[j][i] = Re(Y0) + [hr even] (-1)^i Re(Y_{hr/2}) + 2 sum_{0<k<hr/2} Re(Y_k exp(+2 pi i ik/hr))
__global__ void gen_irow(const GenArgs a)
{
    JOB_PROLOGUE
    const float2 *Y = reinterpret_cast<const float2 *>(sc + a.g.o_row);
    float *resp = reinterpret_cast<float *>(sc + a.g.o_resp);
    const KcfClassDev cls = a.p.classes[meta->size_class];
    const int sk = a.g.sk, nb = a.g.nb;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nb; idx += gridDim.x * blockDim.x) {
        const int j = idx / hr, i = idx - j * hr;
        const float2 *y = Y + (size_t)j * sk;
        double s = (double)y[0].x;
        const int kmax = (hr & 1) ? sk : sk - 1;                           // exclusive bound of the doubled terms
        int t = i % hr;
        for (int k = 1; k < kmax; ++k) {
            const double2 w = cls.tw_hr[t];                                // exp(-i th): Re(Y e^{+i th}) = Yr cos + Yi sin(-(-)) 
            s += 2.0 * ((double)y[k].x * w.x + (double)y[k].y * w.y);
            t += i; if (t >= hr) t -= hr;
        }
        if (!(hr & 1)) s += ((i & 1) ? -1.0 : 1.0) * (double)y[sk - 1].x;
        resp[idx] = (float)s;
        if (a.p.dump.resp) a.p.dump.resp[idx] = (float)s;
    }
}

template <int MODE> __global__ void gen_finish(const GenArgs a)
{
    JOB_PROLOGUE
    const KcfLaunch &p = a.p;
    if (MODE == KCF_MODE_UPDATE) {
        if (threadIdx.x == 0) {
            const mot_bbox_t box = p.boxes[job];
            meta->pos = box;
            meta->scale_horiz = __fdiv_rn((float)(box.r - box.l + 1), (float)cols);
            meta->scale_vert = __fdiv_rn((float)(box.b - box.t + 1), (float)rows);
            meta->first_update = 0;
        }
        return;
    }
    const float *resp = reinterpret_cast<const float *>(sc + a.g.o_resp);
    __shared__ float sv[256]; __shared__ int si[256];
    float best = -99999.0f; int besti = 0x7FFFFFFF;
    for (int i = threadIdx.x; i < a.g.nb; i += blockDim.x) { const float v = resp[i]; if (v > best) { best = v; besti = i; } }
    sv[threadIdx.x] = best; si[threadIdx.x] = besti;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)blockDim.x; ++w) if (sv[w] > best || (sv[w] == best && si[w] < besti)) { best = sv[w]; besti = si[w]; }
        int vd = 1, hd = 1;
        if (besti != 0x7FFFFFFF) { hd = besti / hr + 1; vd = besti - (hd - 1) * hr + 1; }
        if (p.dump.peak) { p.dump.peak[0] = vd; p.dump.peak[1] = hd; }
        if (vd > hr / 2) vd -= hr;
        if (hd > wc / 2) hd -= wc;
        mot_bbox_t pos = meta->pos;
        const float dv = __fmul_rn((float)(KCF_CELL * (vd - 1)), meta->scale_vert);
        const float dh = __fmul_rn((float)(KCF_CELL * (hd - 1)), meta->scale_horiz);
        pos.t = __float2int_rz(__fadd_rn((float)pos.t, dv)); pos.b = __float2int_rz(__fadd_rn((float)pos.b, dv));
        pos.l = __float2int_rz(__fadd_rn((float)pos.l, dh)); pos.r = __float2int_rz(__fadd_rn((float)pos.r, dh));
        meta->pos = pos;
        if (p.clamp_to_frame) {
            pos.l = clampi(pos.l, 0, p.frame_w - 1); pos.r = clampi(pos.r, 0, p.frame_w - 1);
            pos.t = clampi(pos.t, 0, p.frame_h - 1); pos.b = clampi(pos.b, 0, p.frame_h - 1);
        }
        p.boxes[job] = pos;
    }
}

__global__ void gen_dump(const GenArgs a)
{
    JOB_PROLOGUE
    const KcfDump &d = a.p.dump;
    const int nb = a.g.nb, tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    const float *gray = reinterpret_cast<const float *>(sc + a.g.o_gray), *m0 = reinterpret_cast<const float *>(sc + a.g.o_m0);
    const unsigned char *bins = reinterpret_cast<const unsigned char *>(sc + a.g.o_bin);
    const float *r1 = reinterpret_cast<const float *>(sc + a.g.o_r1), *N = reinterpret_cast<const float *>(sc + a.g.o_n);
    const float *feat = reinterpret_cast<const float *>(sc + a.g.o_feat);
    for (int i = tid; i < rows * cols; i += nt) d.gray[i] = gray[i];
    for (int i = tid; i < 16 * nb; i += nt) { d.m0[i] = m0[i]; d.bin[i] = bins[i]; }
    for (int i = tid; i < 18 * nb; i += nt) d.r1[i] = r1[i];
    for (int i = tid; i < (hr + 1) * (wc + 1); i += nt) d.nrm[i] = N[i];
    for (int i = tid; i < 31 * nb; i += nt) d.feat[i] = feat[i];
}

size_t kcf_generic_scratch_bytes(int hr, int wc) { return make_geo(hr, wc).job_bytes; }

int kcf_launch_generic(int mode, int hr, int wc, const KcfLaunch &p, void *scratch, size_t scratch_bytes, cudaStream_t s)
{
    GenArgs a{};
    a.p = p; a.g = make_geo(hr, wc); a.scratch = static_cast<char *>(scratch);
    const int chunk = (int)(scratch_bytes / a.g.job_bytes);
    if (chunk < 1) return -(int)cudaErrorMemoryAllocation;
    int launches = 0;
    auto blocks = [](long work) { long b = (work + 255) / 256; return (unsigned)(b < 1 ? 1 : (b > 1024 ? 1024 : b)); };
    for (int j0 = 0; j0 < p.n_jobs; j0 += chunk) {
        const int nj = (p.n_jobs - j0 < chunk) ? p.n_jobs - j0 : chunk;
        a.job0 = j0;
        const GenGeo &g = a.g;
        gen_gray<<<dim3(blocks((long)g.rmax * g.cmax), nj), 256, 0, s>>>(a);
        gen_grad<<<dim3(blocks(16L * g.nb), nj), 256, 0, s>>>(a);
        gen_hist<<<dim3(blocks(g.nb), nj), 256, 0, s>>>(a);
        gen_norm<<<dim3(blocks((long)(hr + 1) * (wc + 1)), nj), 256, 0, s>>>(a);
        gen_feat<<<dim3(blocks(31L * g.nb), nj), 256, 0, s>>>(a);
        if (p.dump.gray) { gen_dump<<<dim3(blocks(31L * g.nb), nj), 256, 0, s>>>(a); ++launches; }
        gen_rowdft<<<dim3(blocks(31L * g.S), nj), 256, 0, s>>>(a);
        launches += 6;
        if (mode == KCF_MODE_PREDICT) {
            gen_coldft<KCF_MODE_PREDICT><<<dim3(blocks(31L * g.S), nj), 256, 0, s>>>(a);
            gen_chansum<KCF_MODE_PREDICT><<<dim3(blocks(g.S), nj), 256, 0, s>>>(a);
            gen_icol<<<dim3(blocks(g.S), nj), 256, 0, s>>>(a);
            gen_irow<<<dim3(blocks(g.nb), nj), 256, 0, s>>>(a);
            gen_finish<KCF_MODE_PREDICT><<<dim3(1, nj), 256, 0, s>>>(a);
            launches += 5;
        } else {
            gen_coldft<KCF_MODE_UPDATE><<<dim3(blocks(31L * g.S), nj), 256, 0, s>>>(a);
            gen_chansum<KCF_MODE_UPDATE><<<dim3(blocks(g.S), nj), 256, 0, s>>>(a);
            gen_finish<KCF_MODE_UPDATE><<<dim3(1, nj), 32, 0, s>>>(a);
            launches += 3;
        }
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return -(int)e;
    }
    return launches;
}

}  // namespace mot
