/*
 * oracle/port/port_yolo.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the YOLO3 post-processing that produces the tracker's detections:
 *   decode    <- decode_netout        detectors/yolo3.cpp:141-201
 *   correct   <- correct_yolo_boxes   detectors/yolo3.cpp:203-251
 *   sort_idx  <- sort                 detectors/yolo3.cpp:253-276   (in-place exchange sort on an index vector, not stable)
 *   nms       <- do_nms               detectors/yolo3.cpp:278-356
 *   driver    <- tensorRunB           detectors/yolo3.cpp:487-527   (three scales, clip to the image, bbox_chain_t output)
 * Quirks kept on purpose: the logistic / exponential are single precision (C++ float overloads of exp); the box corners are
 * truncated to int; `is_suppressed` is never cleared between classes (:287, :300-303), so flags raised while class c was
 * processed stay raised at the same positions for every later class; the exchange sort's order among equal scores.
 * Not reproduced: with no candidate at all the reference pushes one uninitialised box (:214-217, undefined behaviour); the
 * restatement (and the wrapper of the compiled original) return no detections.
 * Pinned bit-exactly against the compiled original (oracle/_ref/libref_yolo.so) in tests/test_oracle_yolo.py.
 */
#include "port_types.h"
#include <math.h>
#include <stdlib.h>

typedef struct { float x, y, u, w; int c; float s; } predecode_t;                       /* :96-100 */
typedef struct { int xmin, ymin, xmax, ymax, classes; float objectness; } detection_t;  /* :102-109 */

typedef struct { predecode_t *p; int n, cap; } pre_vec;
static void pre_push(pre_vec *v, predecode_t b)
{
    if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 256; v->p = (predecode_t *)realloc(v->p, sizeof(predecode_t) * (size_t)v->cap); }
    v->p[v->n++] = b;
}

static float act(const float *o, int k, int per)          /* :157-170: exp for w/h, logistic for everything else */
{
    const int r = k % per;
    if (r == 2 || r == 3) return expf(o[k]);
    return 1.0f / (1.0f + expf(-o[k]));
}

static void decode(pre_vec *boxes, const float *out, const int *anch, float obj_thresh, int th, int tw, int gh, int gw, int nc)
{
    const int per = 5 + nc, nb_box = 3 * per;
    for (int i = 0; i < gh * gw; ++i) {                                                   /* :180-200 */
        const int row = i / gw, col = i % gw;
        const float *cell = out + (long)(row * gw + col) * nb_box;
        for (int b = 0; b < 3; ++b) {
            const float objectness = act(cell, b * per + 4, per);
            for (int j = 0; j < nc; ++j) {
                const float scores = act(cell, b * per + 5 + j, per) * objectness;
                if (scores >= obj_thresh) {
                    predecode_t box;
                    box.x = ((float)col + act(cell, b * per + 0, per)) / (float)gw;
                    box.y = ((float)row + act(cell, b * per + 1, per)) / (float)gh;
                    box.u = (float)anch[2 * b + 0] * act(cell, b * per + 2, per) / (float)tw;
                    box.w = (float)anch[2 * b + 1] * act(cell, b * per + 3, per) / (float)th;
                    box.s = scores; box.c = j;
                    pre_push(boxes, box);
                }
            }
        }
    }
}

static void correct(detection_t *cb, const predecode_t *boxes, int n, int th, int tw, int ih, int iw)
{
    float new_w, new_h;                                                                   /* :219-230 */
    if (((float)tw / (float)iw) < ((float)th / (float)ih)) { new_w = (float)tw; new_h = roundf((float)ih * (float)tw / (float)iw); }
    else { new_h = (float)th; new_w = roundf((float)iw * (float)th / (float)ih); }
    for (int i = 0; i < n; ++i) {                                                         /* :232-249 */
        const float x_offset = (float)((double)((float)tw - new_w) / 2.0 / (double)tw);
        const float x_scale = new_w / (float)tw;
        const float y_offset = (float)((double)((float)th - new_h) / 2.0 / (double)th);
        const float y_scale = new_h / (float)th;
        const float x = (boxes[i].x - x_offset) / x_scale * (float)iw;
        const float y = (boxes[i].y - y_offset) / y_scale * (float)ih;
        const float w = boxes[i].u / x_scale * (float)iw;
        const float h = boxes[i].w / y_scale * (float)ih;
        cb[i].xmin = (int)(x - w / 2); cb[i].xmax = (int)(x + w / 2);
        cb[i].ymin = (int)(y - h / 2); cb[i].ymax = (int)(y + h / 2);
        cb[i].objectness = boxes[i].s; cb[i].classes = boxes[i].c;
    }
}

static void sort_idx(const detection_t *cb, int n, int *idx)                             /* :253-276 */
{
    for (int i = 0; i < n; ++i) idx[i] = i;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (cb[idx[j]].objectness > cb[idx[i]].objectness) { const int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
}

__attribute__((visibility("default")))
int port_yolo_post(const float *out0, const float *out1, const float *out2, const int *anchors, float obj_thresh, float nms_thresh,
                   int th, int tw, int ih, int iw, int nc, bbox_t *out, int max_out)
{
    const int gh = th / 32, gw = tw / 32;                                                 /* :405-406 */
    pre_vec pv = { 0, 0, 0 };
    decode(&pv, out0, anchors + 12, obj_thresh, th, tw, gh << 0, gw << 0, nc);            /* :496-498 */
    decode(&pv, out1, anchors + 6, obj_thresh, th, tw, gh << 1, gw << 1, nc);
    decode(&pv, out2, anchors + 0, obj_thresh, th, tw, gh << 2, gw << 2, nc);
    const int n = pv.n;
    if (n == 0) { free(pv.p); return 0; }
    detection_t *boxes = (detection_t *)malloc(sizeof(detection_t) * (size_t)n);
    detection_t *cb = (detection_t *)malloc(sizeof(detection_t) * (size_t)n);
    detection_t *nb = (detection_t *)malloc(sizeof(detection_t) * (size_t)n);
    int *idx = (int *)malloc(sizeof(int) * (size_t)n);
    char *is_suppressed = (char *)calloc((size_t)n, 1);      /* grows by appended zeros in the reference; only [0, n_class) is ever touched */
    correct(boxes, pv.p, n, th, tw, ih, iw);
    int nn = 0;
    for (int c = 0; c < nc; ++c) {                                                        /* :292-354 */
        int m = 0;
        for (int j = 0; j < n; ++j) if (boxes[j].classes == c) cb[m++] = boxes[j];
        sort_idx(cb, m, idx);
        for (int i = 0; i < m; ++i) {
            if (is_suppressed[idx[i]]) continue;
            for (int j = i + 1; j < m; ++j) {
                const detection_t *a = &cb[idx[j]], *b = &cb[idx[i]];
                const float maxX = (float)(a->xmax < b->xmax ? a->xmax : b->xmax), maxY = (float)(a->ymax < b->ymax ? a->ymax : b->ymax);
                const float minX = (float)(a->xmin > b->xmin ? a->xmin : b->xmin), minY = (float)(a->ymin > b->ymin ? a->ymin : b->ymin);
                const float overWidth = maxX - minX + 1, overHeight = maxY - minY + 1;
                if ((overWidth > 0) & (overHeight > 0)) {
                    const float area1 = (float)((a->xmax - a->xmin + 1) * (a->ymax - a->ymin + 1));
                    const float area2 = (float)((b->xmax - b->xmin + 1) * (b->ymax - b->ymin + 1));
                    const float IOU = (overWidth * overHeight) / (area1 + area2 - overWidth * overHeight);
                    if (IOU > nms_thresh) is_suppressed[idx[j]] = 1;
                }
            }
        }
        for (int i = 0; i < m; ++i) if (!is_suppressed[idx[i]]) nb[nn++] = cb[idx[i]];
    }
    int nbox = 0;
    for (int i = 0; i < nn; ++i) {                                                        /* :505-526 */
        detection_t d = nb[i];
        d.ymin = d.ymin > 0 ? d.ymin : 0; d.xmin = d.xmin > 0 ? d.xmin : 0;
        d.ymax = d.ymax < ih - 1 ? d.ymax : ih - 1; d.xmax = d.xmax < iw - 1 ? d.xmax : iw - 1;
        if (d.ymin > d.ymax || d.xmin > d.xmax || d.ymin < 0 || d.xmin < 0 || d.xmax >= iw || d.ymax >= ih) continue;
        if (nbox >= max_out) break;
        out[nbox].t = d.ymin; out[nbox].l = d.xmin; out[nbox].b = d.ymax; out[nbox].r = d.xmax; out[nbox].type = d.classes; out[nbox].score = d.objectness;
        ++nbox;
    }
    free(pv.p); free(boxes); free(cb); free(nb); free(idx); free(is_suppressed);
    return nbox;
}
