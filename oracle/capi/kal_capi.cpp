/*
 * oracle/capi/kal_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Compiles the UNMODIFIED reference Kalman tracker by including trackers/kalman.cpp by path
 * (kalman.cpp:131-163 exports the same four tracker_* symbols as kcf.cpp, hence a separate
 * shared object) together with the vendored header-only SigPack 1.2.4 sp::KF
 * (include/sigpack/kalman/kalman.h:120-267) and Armadillo 9.100.5.
 */
#include "ref_common.h"
#include "trackers/kalman.cpp"       /* by path: -I/root/reference */

void assignmentoptimal(int *assignment, double *cost, double *distMatrixIn, int nOfRows, int nOfColumns);
extern "C" {
void port_rgb2gray(float *, const uint8_t *, int, int, int, int, int);
void port_resize_gray(float *, const float *, int, int, int, int);
void port_cost_matrix(double *, const bbox_t *, int, const bbox_t *, int, int, double);
}

REF_API void *ref_kal_new(bbox_t *pbox) { return tracker_new(pbox); }
REF_API void ref_kal_predict(void *p, float *gray, bbox_t *pbox) { tracker_predict(p, gray, pbox); }
REF_API void ref_kal_update(void *p, float *gray, bbox_t *pbox) { tracker_update(p, gray, pbox); }
REF_API void ref_kal_delete(void *p) { tracker_delete(p); }

/* x[6], P[36] (column-major, as arma stores it), K[24] (6x4 column-major) */
REF_API void ref_kal_state(void *p, double *x, double *P, double *K)
{
    kalman_tracker_t *t = (kalman_tracker_t *)p;
    arma::mat xs = t->pkalman->get_state_vec(), Ps = t->pkalman->get_err_cov(), Ks = t->pkalman->get_kalman_gain();
    if (x) memcpy(x, xs.memptr(), sizeof(double) * 6);
    if (P) memcpy(P, Ps.memptr(), sizeof(double) * 36);
    if (K) memcpy(K, Ks.memptr(), sizeof(double) * 24);
}

/* a whole sequence in one call (stress tests): n tracks x nframes, measurements meas[f][i], predicted boxes out[f][i] */
REF_API void ref_kal_run(int n, int nframes, bbox_t *init, bbox_t *meas, bbox_t *out)
{
    for (int i = 0; i < n; ++i) {
        void *t = tracker_new(&init[i]);
        for (int f = 0; f < nframes; ++f) {
            bbox_t b = init[i];
            tracker_predict(t, nullptr, &b);
            out[(long)f * n + i] = b;
            tracker_update(t, nullptr, &meas[(long)f * n + i]);
        }
        tracker_delete(t);
    }
}

#define TDL_PREFIX(n) ref_kal_##n
#define TDL_EXPORT REF_API
#define TDL_IS_KCF 0
#define TDL_TRK_NEW(pb) tracker_new(pb)
#define TDL_TRK_PREDICT(p, g, pb) tracker_predict(p, g, pb)
#define TDL_TRK_UPDATE(p, g, pb) tracker_update(p, g, pb)
#define TDL_TRK_DELETE(p) tracker_delete(p)
#define TDL_ASSIGN(a, c, d, nr, nc) assignmentoptimal(a, c, d, nr, nc)
#include "../port/port_tdloop.inc"
