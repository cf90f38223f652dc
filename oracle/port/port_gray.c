/*
 * oracle/port/port_gray.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the patch preprocessing in front of fHOG:
 *   port_rgb2gray        <- rgb2Gray                  top/drawlib.c:192-240
 *   port_resize_gray     <- bilinearInterpolationGray top/drawlib.c:542-637
 *   port_cost_matrix     <- the cost loops            top/td.cpp:386-457
 * The reference hard-codes a 1280-pixel frame (byte stride 3840, top/drawlib.c:9-10); the
 * restatement takes the byte stride as a parameter and is cross-checked against the compiled
 * original at stride 3840 (tests/test_oracle_port.py).
 */
#include "port_types.h"
#include <math.h>

/* top/drawlib.c:192-240.  Output is column-major rows x cols (pdst[c*rows + r]);
 * gray = 0.144*B + 0.587*G + 0.299*R evaluated in double, rounded to float (:234). */
__attribute__((visibility("default")))
void port_rgb2gray(float *pgra, const uint8_t *pbgr, int stride_bytes, int left, int top, int right, int bottom)
{
    if (top > bottom) { int t = top; top = bottom; bottom = t; }       /* :203-208 */
    if (left > right) { int t = left; left = right; right = t; }       /* :210-215 */
    int cols = right - left + 1, rows = bottom - top + 1;
    const uint8_t *prow = pbgr + (long)top * stride_bytes + (long)left * 3;
    for (int r = 0; r < rows; ++r) {
        const uint8_t *px = prow;
        for (int c = 0; c < cols; ++c) {
            double B = px[0], G = px[1], R = px[2];
            pgra[(long)c * rows + r] = (float)(0.144 * B + 0.587 * G + 0.299 * R);
            px += 3;
        }
        prow += stride_bytes;
    }
}

/* top/drawlib.c:542-637.  The callee treats both buffers as ROW-major (height x width) while the
 * caller's data is column-major; reproduced literally through linear indexing, so equal sizes give
 * an exact copy and unequal sizes give the reference's index-scrambled resample.
 * Called as (dst, src, rows_s, cols_s, rows_d, cols_d) (top/td.cpp:357-364). */
__attribute__((visibility("default")))
void port_resize_gray(float *pdst, const float *psrc, int heightSource, int widthSource, int height, int width)
{
    float xs = ((float)widthSource) / ((float)width);
    float ys = ((float)heightSource) / ((float)height);
    for (int y = 0; y < height; ++y) {
        float sy = y * ys;
        int y0 = (int)sy;
        float fy = sy - y0, ify = 1.0f - fy;
        int y1 = y0 + 1; if (y1 >= heightSource) y1 = y0;
        for (int x = 0; x < width; ++x) {
            float sx = x * xs;
            int x0 = (int)sx;
            float fx = sx - x0, ifx = 1.0f - fx;
            int x1 = x0 + 1; if (x1 >= widthSource) x1 = x0;
            float c1 = psrc[(long)y0 * widthSource + x0], c2 = psrc[(long)y0 * widthSource + x1];
            float c3 = psrc[(long)y1 * widthSource + x0], c4 = psrc[(long)y1 * widthSource + x1];
            float l0 = ifx * c1 + fx * c2;
            float l1 = ifx * c3 + fx * c4;
            pdst[(long)y * width + x] = ify * l0 + fy * l1;
        }
    }
}

/* One cost cell, top/td.cpp:394-419.  REF_CENTROID is the shipped expression (centroid distance
 * / frame_w + 1.0 on class mismatch; the reference's SCREEN_DIS is 1/1280, td.cpp:50).
 * IOU_CLAMPED is the finite form of the commented-out IoU cost. */
__attribute__((visibility("default")))
double port_cost_cell(const bbox_t *t, const bbox_t *d, int mode, double screen_dis)
{
    int maxl = t->l > d->l ? t->l : d->l, maxt = t->t > d->t ? t->t : d->t;
    int minr = t->r < d->r ? t->r : d->r, minb = t->b < d->b ? t->b : d->b;
    double dista = 0.0;
    if (mode == PORT_COST_IOU_CLAMPED) {
        int iw = minr - maxl, ih = minb - maxt;
        if (iw < 0) iw = 0;
        if (ih < 0) ih = 0;
        double inter = (double)(iw * ih);
        double uni = (double)((t->b - t->t) * (t->r - t->l) + (d->b - d->t) * (d->r - d->l)) - inter;
        dista = (uni > 0.0) ? 1.0 - inter / uni : 1.0;
    } else {
        int cxi = (t->l + t->r) >> 1, cyi = (t->t + t->b) >> 1;
        int cxj = (d->l + d->r) >> 1, cyj = (d->t + d->b) >> 1;
        dista += sqrt((double)((cxi - cxj) * (cxi - cxj) + (cyi - cyj) * (cyi - cyj))) * screen_dis;
    }
    if (t->type != d->type) dista += 1.0;
    return dista;
}

/* top/td.cpp:386-457: column-major, rows = the smaller side (trackers if T < D, else detections). */
__attribute__((visibility("default")))
void port_cost_matrix(double *dist, const bbox_t *trk, int T, const bbox_t *det, int D, int mode, double screen_dis)
{
    double *p = dist;
    if (T < D) {
        for (int j = 0; j < D; ++j) for (int i = 0; i < T; ++i) *p++ = port_cost_cell(&trk[i], &det[j], mode, screen_dis);
    } else {
        for (int i = 0; i < T; ++i) for (int j = 0; j < D; ++j) *p++ = port_cost_cell(&trk[i], &det[j], mode, screen_dis);
    }
}
