/*
 * oracle/capi/kcf_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference KCF tracker by including trackers/kcf.cpp by path (its
 * kcf_t and the kcf_* stage functions are file-local, kcf.cpp:27-76,245-453, so the wrapper has
 * to live in the same translation unit) and exposes extern "C" entry points for ctypes:
 * the four tracker_* plugin functions (kcf.cpp:455-491), getters for every intermediate the
 * parity tests compare (xf_tm, xf_fq, xf_md, yf, kf, zf, alpha, response, labels, cos_win), and a
 * frame-loop driver (oracle/port/port_tdloop.inc) that calls the compiled reference functions.
 * libhog/gradientMex.cpp, trackers/hungarian/hungarian.cpp and fftw_shim.cpp are linked alongside.
 */
#include "ref_common.h"
#include "trackers/kcf.cpp"          /* by path: -I/root/reference */
#undef max
#undef min

void assignmentoptimal(int *assignment, double *cost, double *distMatrixIn, int nOfRows, int nOfColumns);
extern "C" {
void port_rgb2gray(float *, const uint8_t *, int, int, int, int, int);
void port_resize_gray(float *, const float *, int, int, int, int);
void port_cost_matrix(double *, const bbox_t *, int, const bbox_t *, int, int, double);
}

REF_API void *ref_kcf_new(bbox_t *pbox) { return tracker_new(pbox); }
REF_API void ref_kcf_predict(void *p, float *gray, bbox_t *pbox) { tracker_predict(p, gray, pbox); }
REF_API void ref_kcf_update(void *p, float *gray, bbox_t *pbox) { tracker_update(p, gray, pbox); }
REF_API void ref_kcf_delete(void *p) { tracker_delete(p); }

/* dims[0..5] = rows, cols, f_rows, f_cols, f_chan, S = f_cols*(f_rows/2+1) */
REF_API void ref_kcf_dims(void *p, int *dims)
{
    kcf_t *k = (kcf_t *)p;
    dims[0] = k->rows; dims[1] = k->cols; dims[2] = k->f_rows; dims[3] = k->f_cols; dims[4] = k->f_chan;
    dims[5] = k->f_cols * (k->f_rows / 2 + 1);
}

/* which: 0 xf_tm f32[31*nb] | 1 xf_fq c64[31*S] | 2 xf_md c64[31*S] | 3 yf c64[S] | 4 kf c64[S] |
 *        5 zf c64[S] | 6 alpha f32[S] | 7 response f32[nb] | 8 labels f32[nb] | 9 cos_win f32[nb]
 * returns the number of floats copied */
REF_API long ref_kcf_get(void *p, int which, float *out)
{
    kcf_t *k = (kcf_t *)p;
    long nb = (long)k->f_cols * k->f_rows, S = (long)k->f_cols * (k->f_rows / 2 + 1), n = 0;
    const float *src = 0;
    switch (which) {
    case 0: src = k->xf_tm; n = 31 * nb; break;
    case 1: src = (const float *)k->xf_fq; n = 31 * S * 2; break;
    case 2: src = (const float *)k->xf_md; n = 31 * S * 2; break;
    case 3: src = (const float *)k->yf; n = S * 2; break;
    case 4: src = (const float *)k->kf; n = S * 2; break;
    case 5: src = (const float *)k->zf; n = S * 2; break;
    case 6: src = k->alpha; n = S; break;
    case 7: src = k->response; n = nb; break;
    case 8: src = k->labels.memptr(); n = nb; break;
    case 9: src = k->cos_win.memptr(); n = nb; break;
    default: return -1;
    }
    memcpy(out, src, sizeof(float) * (size_t)n);
    return n;
}

REF_API void ref_kcf_state(void *p, bbox_t *pos, float *scales, int *first_update)
{
    kcf_t *k = (kcf_t *)p;
    *pos = k->pos; scales[0] = k->scale_horiz; scales[1] = k->scale_vert; *first_update = k->first_update;
}

/* stand-alone fHOG through the reference's own driver, libhog/fhog.h:16-38 */
REF_API void ref_fhog_extract(float *I, int h, int w, float *H) { FHoG::extract(I, h, w, H); }
REF_API void ref_gradmag(float *I, float *M, float *O, int h, int w) { gradMag(I, M, O, h, w, 1, true); }
REF_API void ref_gradhist18(float *M, float *O, float *H, int h, int w) { gradHist(M, O, H, h, w, 4, 18, -1, true); }

/* frame loop over the compiled reference */
#define TDL_PREFIX(n) ref_kcf_##n
#define TDL_EXPORT REF_API
#define TDL_IS_KCF 1
#define TDL_TRK_NEW(pb) tracker_new(pb)
#define TDL_TRK_PREDICT(p, g, pb) tracker_predict(p, g, pb)
#define TDL_TRK_UPDATE(p, g, pb) tracker_update(p, g, pb)
#define TDL_TRK_DELETE(p) tracker_delete(p)
#define TDL_ASSIGN(a, c, d, nr, nc) assignmentoptimal(a, c, d, nr, nc)
#include "../port/port_tdloop.inc"
