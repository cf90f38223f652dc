/*
 * oracle/port/port_kcf.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the reference KCF/DCF tracker, trackers/kcf.cpp:
 *   port_kcf_new      <- kcf_initialize + gaussian_shaped_labels + circshift + cosine_window_function
 *                        + kcf_fft2_label                        kcf.cpp:146-214, 96-122, 78-94, 124-130, 132-144
 *   features()        <- kcf_get_features + kcf_fft2_features    kcf.cpp:245-267
 *   port_kcf_predict  <- kcf_linear_correlation_zf + kcf_predict_ifft2 + kcf_predict + tracker_predict
 *                                                                kcf.cpp:306-362, 397-439, 455-460
 *   port_kcf_update   <- kcf_linear_correlation_kf + kcf_update_alpha + kcf_update_xf + kcf_update
 *                        + tracker_update                        kcf.cpp:269-304, 364-395, 441-476
 * Note what the reference really is: a LINEAR-kernel correlation filter with a REAL alpha; only the
 * labels are Gaussian shaped; integer-cell argmax, no sub-pixel step, no padding.
 * FFTs: see port_fft.h (FFTW 3.3.5 is an absent third-party dependency; parity at that boundary is unpinned).
 * Layout: features f32[31][wc][hr] (hr fastest); spectra c64[31][wc][hr/2+1] (FFTW n0=wc, n1=hr).
 */
#include "port_types.h"
#include "port_fft.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void port_fhog_extract(const float *I, int h, int w, float *H);

typedef struct {
    int rows, cols, hr, wc, nb, S;
    float *xf_tm, *xf_fq, *xf_md;      /* 32*nb | 31*S*2 | 31*S*2 */
    float *yf, *kf, *zf;               /* S*2 each */
    float *alpha, *response;           /* S | nb */
    float *labels, *cos_win;           /* nb each */
    float feature_norm_ratio;
    bbox_t pos; float scale_vert, scale_horiz;
    int first_update; float factor, lamda;
} port_kcf_t;

/* include/sigpack/window/window.h:34-48, 83-89: hann_f(N)[i] = (float)(0.5 - 0.5 cos(2 pi i/(N-1)) [+ zero terms]) */
static void hann_f(int N, float *h)
{
    const double PI_2 = 6.28318530717958647692;          /* include/sigpack/base/base.h:14 */
    for (int i = 0; i < N; ++i) {
        double ha = 0.5 - 0.5 * cos(1.0 * PI_2 * i / (N - 1)) + 0.0 * cos(2.0 * PI_2 * i / (N - 1))
                  - 0.0 * cos(3.0 * PI_2 * i / (N - 1)) + 0.0 * cos(4.0 * PI_2 * i / (N - 1));
        h[i] = (float)ha;
    }
}

__attribute__((visibility("default")))
void *port_kcf_new(const bbox_t *pbox)
{
    port_kcf_t *k = (port_kcf_t *)calloc(1, sizeof(*k));
    k->rows = pbox->b - pbox->t + 1; k->cols = pbox->r - pbox->l + 1;           /* kcf.cpp:148-149 */
    k->hr = k->rows / 4; k->wc = k->cols / 4;                                    /* :155-156, cell 4 (:488) */
    k->nb = k->hr * k->wc; k->S = k->wc * (k->hr / 2 + 1);
    int nb = k->nb, S = k->S, hr = k->hr, wc = k->wc;
    k->xf_tm = (float *)calloc((size_t)nb * 32, sizeof(float));
    k->xf_fq = (float *)calloc((size_t)S * 2 * 31, sizeof(float));
    k->xf_md = (float *)calloc((size_t)S * 2 * 31, sizeof(float));              /* zeroed, :174 */
    k->yf = (float *)calloc((size_t)S * 2, sizeof(float));
    k->kf = (float *)calloc((size_t)S * 2, sizeof(float));
    k->zf = (float *)calloc((size_t)S * 2, sizeof(float));
    k->alpha = (float *)calloc((size_t)S, sizeof(float));
    k->response = (float *)calloc((size_t)nb, sizeof(float));
    k->labels = (float *)calloc((size_t)nb, sizeof(float));
    k->cos_win = (float *)calloc((size_t)nb, sizeof(float));
    k->feature_norm_ratio = (float)(1.0 / ((float)(wc * hr * 31)));              /* :197 */
    k->scale_vert = 1.0f; k->scale_horiz = 1.0f; k->pos = *pbox;                /* :200-202 */

    /* gaussian_shaped_labels(0.7289, f_rows, f_cols), :96-122 */
    {
        float sigma = 0.7289f;
        float sigma_s_inv = (float)(1.0 / (sigma * sigma));
        float *gx = (float *)malloc(sizeof(float) * hr), *gy = (float *)malloc(sizeof(float) * wc);
        float *lab = (float *)malloc(sizeof(float) * nb);
        int x0 = -hr / 2, y0 = -wc / 2;
        for (int i = 0; i < hr; ++i) { int x = x0 + i; gx[i] = (float)exp(-0.5 * x * x * sigma_s_inv); }
        for (int j = 0; j < wc; ++j) { int y = y0 + j; gy[j] = (float)exp(-0.5 * y * y * sigma_s_inv); }
        for (int j = 0; j < wc; ++j) for (int i = 0; i < hr; ++i) lab[j * hr + i] = gx[i] * gy[j];
        /* circshift(out, in, xshift=-hr/2, yshift=-wc/2), :78-94 */
        for (int j = 0; j < wc; ++j) {
            int jj = (j + y0) % wc; if (jj < 0) jj += wc;
            for (int i = 0; i < hr; ++i) {
                int ii = (i + x0) % hr; if (ii < 0) ii += hr;
                k->labels[jj * hr + ii] = lab[j * hr + i];
            }
        }
        free(gx); free(gy); free(lab);
    }
    /* cosine_window_function(f_rows, f_cols) = hann(hr) * hann(wc)^T, :124-130 */
    {
        float *wy = (float *)malloc(sizeof(float) * hr), *wx = (float *)malloc(sizeof(float) * wc);
        hann_f(hr, wy); hann_f(wc, wx);
        for (int j = 0; j < wc; ++j) for (int i = 0; i < hr; ++i) k->cos_win[j * hr + i] = wy[i] * wx[j];
        free(wy); free(wx);
    }
    pf_r2c_2d(wc, hr, k->labels, k->yf);                                         /* kcf_fft2_label :132-144 */
    k->first_update = 1; k->factor = 0.05f; k->lamda = 0.0001f;                  /* :210-212 */
    return k;
}

__attribute__((visibility("default")))
void port_kcf_delete(void *p)
{
    port_kcf_t *k = (port_kcf_t *)p;
    free(k->xf_tm); free(k->xf_fq); free(k->xf_md); free(k->yf); free(k->kf); free(k->zf);
    free(k->alpha); free(k->response); free(k->labels); free(k->cos_win); free(k);
}

/* kcf_get_features (:245-259) + kcf_fft2_features (:261-267) */
static void features(port_kcf_t *k, const float *gray)
{
    port_fhog_extract(gray, k->rows, k->cols, k->xf_tm);
    for (int c = 0; c < 31; ++c) for (int i = 0; i < k->nb; ++i) k->xf_tm[c * k->nb + i] *= k->cos_win[i];
    for (int c = 0; c < 31; ++c) pf_r2c_2d(k->wc, k->hr, k->xf_tm + (long)c * k->nb, k->xf_fq + (long)c * k->S * 2);
}

__attribute__((visibility("default")))
void port_kcf_predict(void *p, const float *gray, bbox_t *pbox)
{
    port_kcf_t *k = (port_kcf_t *)p;
    const int S = k->S;
    features(k, gray);
    /* kcf_linear_correlation_zf (:306-362): zf = sum_c xf * conj(md), then * alpha * norm */
    for (int c = 0; c < 31; ++c) {
        const float *a = k->xf_fq + (long)c * S * 2, *b = k->xf_md + (long)c * S * 2;
        for (int i = 0; i < S; ++i) {
            float ia = a[2 * i], qa = a[2 * i + 1], ib = b[2 * i], qb = b[2 * i + 1];
            float ic = ia * ib + qa * qb;
            float qc = qa * ib - ia * qb;
            if (c == 0) { k->zf[2 * i] = ic; k->zf[2 * i + 1] = qc; }
            else        { k->zf[2 * i] += ic; k->zf[2 * i + 1] += qc; }
        }
    }
    for (int i = 0; i < S; ++i) {
        k->zf[2 * i]     = k->zf[2 * i]     * k->alpha[i] * k->feature_norm_ratio;
        k->zf[2 * i + 1] = k->zf[2 * i + 1] * k->alpha[i] * k->feature_norm_ratio;
    }
    /* kcf_predict_ifft2 (:397-428) */
    pf_c2r_2d(k->wc, k->hr, k->zf, k->response);
    float max_val = -99999.0f; int vd = 0, hd = 0;
    const float *pr = k->response;
    for (int j = 1; j <= k->wc; ++j) for (int i = 1; i <= k->hr; ++i) {
        if (pr[0] > max_val) { max_val = pr[0]; vd = i; hd = j; }
        ++pr;
    }
    if (vd > k->hr / 2) vd -= k->hr;
    if (hd > k->wc / 2) hd -= k->wc;
    k->pos.t = (int)(k->pos.t + 4 * (vd - 1) * k->scale_vert);
    k->pos.b = (int)(k->pos.b + 4 * (vd - 1) * k->scale_vert);
    k->pos.l = (int)(k->pos.l + 4 * (hd - 1) * k->scale_horiz);
    k->pos.r = (int)(k->pos.r + 4 * (hd - 1) * k->scale_horiz);
    *pbox = k->pos;                                                              /* :438 */
}

__attribute__((visibility("default")))
void port_kcf_update(void *p, const float *gray, const bbox_t *pbox)
{
    port_kcf_t *k = (port_kcf_t *)p;
    const int S = k->S;
    k->pos = *pbox;                                                              /* :470-472 */
    k->scale_horiz = ((float)(pbox->r - pbox->l + 1)) / ((float)k->cols);
    k->scale_vert  = ((float)(pbox->b - pbox->t + 1)) / ((float)k->rows);
    float factor = k->first_update ? 1.0f : k->factor;                           /* :443 */
    float lamda = k->lamda;
    k->first_update = 0;
    features(k, gray);
    /* kcf_linear_correlation_kf (:269-304) */
    for (int c = 0; c < 31; ++c) {
        const float *a = k->xf_fq + (long)c * S * 2;
        for (int i = 0; i < S; ++i) {
            float e = (a[2 * i] * a[2 * i]) + (a[2 * i + 1] * a[2 * i + 1]);
            if (c == 0) k->kf[2 * i] = e; else k->kf[2 * i] = e + k->kf[2 * i];
        }
    }
    for (int i = 0; i < S; ++i) k->kf[2 * i] = k->kf[2 * i] * k->feature_norm_ratio;
    /* kcf_update_alpha (:364-378) */
    for (int i = 0; i < S; ++i) {
        float a = k->yf[2 * i] / (k->kf[2 * i] + lamda);
        k->alpha[i] = (1 - factor) * k->alpha[i] + factor * a;
    }
    /* kcf_update_xf (:380-395) */
    for (long i = 0; i < (long)31 * S * 2; ++i) k->xf_md[i] = (1 - factor) * k->xf_md[i] + factor * k->xf_fq[i];
}

/* same selector as oracle/capi/kcf_capi.cpp:ref_kcf_get */
__attribute__((visibility("default")))
long port_kcf_get(void *p, int which, float *out)
{
    port_kcf_t *k = (port_kcf_t *)p;
    long nb = k->nb, S = k->S, n = 0; const float *src = 0;
    switch (which) {
    case 0: src = k->xf_tm; n = 31 * nb; break;
    case 1: src = k->xf_fq; n = 31 * S * 2; break;
    case 2: src = k->xf_md; n = 31 * S * 2; break;
    case 3: src = k->yf; n = S * 2; break;
    case 4: src = k->kf; n = S * 2; break;
    case 5: src = k->zf; n = S * 2; break;
    case 6: src = k->alpha; n = S; break;
    case 7: src = k->response; n = nb; break;
    case 8: src = k->labels; n = nb; break;
    case 9: src = k->cos_win; n = nb; break;
    default: return -1;
    }
    memcpy(out, src, sizeof(float) * (size_t)n);
    return n;
}

__attribute__((visibility("default")))
void port_kcf_dims(void *p, int *dims)
{
    port_kcf_t *k = (port_kcf_t *)p;
    dims[0] = k->rows; dims[1] = k->cols; dims[2] = k->hr; dims[3] = k->wc; dims[4] = 31; dims[5] = k->S;
}
