/*
 * oracle/port/port_kalman.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the reference constant-velocity Kalman tracker:
 *   port_kal_new      <- tracker_new + kalman_tracker_initialize   trackers/kalman.cpp:147-163, 29-97
 *   port_kal_predict  <- kalman_tracker_predict + sp::KF::predict  trackers/kalman.cpp:105-116,
 *                                                                  include/sigpack/kalman/kalman.h:207-220
 *   port_kal_update   <- kalman_tracker_update + sp::KF::update    trackers/kalman.cpp:118-128,
 *                                                                  include/sigpack/kalman/kalman.h:225-237
 * State [l,t,r,b,vx,vy], N=6, M=4, double precision, dT=1, P0=1e4, Q0=1e-2, R0=512 (kalman.cpp:46-51).
 * The 4x4 inverse is Armadillo's closed-form inv_tiny (include/armadillo_bits/op_inv_meat.hpp:69-72);
 * here: adjugate / determinant.  Matrices are column-major like arma::mat.
 */
#include "port_types.h"
#include <stdlib.h>
#include <string.h>

#define A_(M_, r, c, ld) (M_)[(c) * (ld) + (r)]

typedef struct { double x[6], P[36], K[24]; } port_kal_t;

static const double A6[36] = {   /* column-major of kalman.cpp:55-63 */
    1,0,0,0,0,0,  0,1,0,0,0,0,  0,0,1,0,0,0,  0,0,0,1,0,0,  1,0,1,0,1,0,  0,1,0,1,0,1 };
static const double Qt[36] = {   /* kalman.cpp:75-83 (symmetric) */
    0.25,0,0,0,0.5,0,  0,0.25,0,0,0,0.5,  0,0,0.25,0,0.5,0,  0,0,0,0.25,0,0.5,  0.5,0,0.5,0,1,0,  0,0.5,0,0.5,0,1 };

static void mm(const double *X, const double *Y, double *Z, int n, int k, int m)   /* Z(n x m) = X(n x k) Y(k x m) */
{
    for (int c = 0; c < m; ++c) for (int r = 0; r < n; ++r) {
        double s = 0.0;
        for (int i = 0; i < k; ++i) s += A_(X, r, i, n) * A_(Y, i, c, k);
        A_(Z, r, c, n) = s;
    }
}
static void tr(const double *X, double *Y, int n, int m)    /* Y(m x n) = X(n x m)^T */
{
    for (int c = 0; c < m; ++c) for (int r = 0; r < n; ++r) A_(Y, c, r, m) = A_(X, r, c, n);
}

static void inv4(const double *S, double *Si)
{
    double cof[16];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) {
        double m[9]; int k = 0;
        for (int i = 0; i < 4; ++i) if (i != r) for (int j = 0; j < 4; ++j) if (j != c) m[k++] = A_(S, i, j, 4);
        double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        cof[r * 4 + c] = ((r + c) & 1) ? -d : d;
    }
    double det = 0.0;
    for (int c = 0; c < 4; ++c) det += A_(S, 0, c, 4) * cof[0 * 4 + c];
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) A_(Si, r, c, 4) = cof[c * 4 + r] / det;
}

__attribute__((visibility("default")))
void *port_kal_new(const bbox_t *pbox)
{
    port_kal_t *k = (port_kal_t *)calloc(1, sizeof(*k));
    k->x[0] = pbox->l; k->x[1] = pbox->t; k->x[2] = pbox->r; k->x[3] = pbox->b;   /* kalman.cpp:152-157 */
    for (int i = 0; i < 6; ++i) A_(k->P, i, i, 6) = 1e4;                           /* :90-91 */
    return k;
}
__attribute__((visibility("default")))
void port_kal_delete(void *p) { free(p); }

__attribute__((visibility("default")))
void port_kal_predict(void *p, const float *gray, bbox_t *pbox)
{
    (void)gray;
    port_kal_t *k = (port_kal_t *)p;
    double x2[6], AP[36], At[36], APA[36];
    mm(A6, k->x, x2, 6, 6, 1);                      /* x = A x (+ B u with L = 0) */
    memcpy(k->x, x2, sizeof(x2));
    mm(A6, k->P, AP, 6, 6, 6); tr(A6, At, 6, 6); mm(AP, At, APA, 6, 6, 6);
    for (int i = 0; i < 36; ++i) k->P[i] = APA[i] + 1e-2 * Qt[i];   /* P = A P A^T + Q, Q = Q0 * Qt (:84) */
    pbox->l = (int)k->x[0]; pbox->t = (int)k->x[1]; pbox->r = (int)k->x[2]; pbox->b = (int)k->x[3];   /* :112-115 */
}

__attribute__((visibility("default")))
void port_kal_update(void *p, const float *gray, const bbox_t *pbox)
{
    (void)gray;
    port_kal_t *k = (port_kal_t *)p;
    double z[4] = { (double)pbox->l, (double)pbox->t, (double)pbox->r, (double)pbox->b };   /* :122-125 */
    double H[24] = { 0 }, Ht[24], PHt[24], HP[24], S[16], Si[16], zerr[4], Jf[36], JfP[36], Jft[36], JPJ[36], KR[24], Kt[24], KRK[36];
    for (int i = 0; i < 4; ++i) A_(H, i, i, 4) = 1.0;                  /* H = [I4 0], kalman.cpp:66-72 */
    tr(H, Ht, 4, 6);
    mm(k->P, Ht, PHt, 6, 6, 4);
    mm(H, k->P, HP, 4, 6, 6); mm(HP, Ht, S, 4, 6, 4);
    for (int i = 0; i < 4; ++i) A_(S, i, i, 4) += 512.0;               /* + R, R = 512 I4 (:87-88) */
    inv4(S, Si);
    mm(PHt, Si, k->K, 6, 4, 4);                                        /* K = P H^T inv(H P H^T + R) */
    for (int i = 0; i < 4; ++i) zerr[i] = z[i] - k->x[i];              /* z_err = z - H x */
    for (int r = 0; r < 6; ++r) { double s = 0.0; for (int i = 0; i < 4; ++i) s += A_(k->K, r, i, 6) * zerr[i]; k->x[r] += s; }
    /* Joseph form: P = (I-KH) P (I-KH)^T + K R K^T */
    mm(k->K, H, Jf, 6, 4, 6);
    for (int i = 0; i < 36; ++i) Jf[i] = -Jf[i];
    for (int i = 0; i < 6; ++i) A_(Jf, i, i, 6) += 1.0;
    mm(Jf, k->P, JfP, 6, 6, 6); tr(Jf, Jft, 6, 6); mm(JfP, Jft, JPJ, 6, 6, 6);
    for (int i = 0; i < 24; ++i) KR[i] = k->K[i] * 512.0;
    tr(k->K, Kt, 6, 4); mm(KR, Kt, KRK, 6, 4, 6);
    for (int i = 0; i < 36; ++i) k->P[i] = JPJ[i] + KRK[i];
}

__attribute__((visibility("default")))
void port_kal_state(void *p, double *x, double *P, double *K)
{
    port_kal_t *k = (port_kal_t *)p;
    if (x) memcpy(x, k->x, sizeof(k->x));
    if (P) memcpy(P, k->P, sizeof(k->P));
    if (K) memcpy(K, k->K, sizeof(k->K));
}

/* a whole sequence in one call (stress tests): n tracks x nframes, measurements meas[f][i], predicted boxes out[f][i] */
__attribute__((visibility("default")))
void port_kal_run(int n, int nframes, const bbox_t *init, const bbox_t *meas, bbox_t *out)
{
    for (int i = 0; i < n; ++i) {
        void *t = port_kal_new(&init[i]);
        for (int f = 0; f < nframes; ++f) {
            bbox_t b = init[i];
            port_kal_predict(t, 0, &b);
            out[(long)f * n + i] = b;
            port_kal_update(t, 0, &meas[(long)f * n + i]);
        }
        port_kal_delete(t);
    }
}
