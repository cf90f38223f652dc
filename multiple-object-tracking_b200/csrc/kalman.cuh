// kalman.cuh -- device code of the batched constant-velocity Kalman tracker: one thread per track, FP64 registers.  Included by
// kalman.cu (the batched launches) and td_device.cu (the one-launch frame loop); both are compiled with --fmad=false so that every
// product and sum is individually rounded like the reference's -ffp-contract=off build.
//
// Replaces trackers/kalman.cpp:105-128 (kalman_tracker_predict / kalman_tracker_update) and the SigPack sp::KF
// algebra behind them (include/sigpack/kalman/kalman.h:207-237): x = A x, P = A P A^T + Q; K = P H^T inv(H P H^T + R),
// x += K (z - H x), Joseph form P = (I-KH) P (I-KH)^T + K R K^T.  N = 6 states [l,t,r,b,vx,vy], M = 4 measurements.
// A and H are 0/1 matrices (kalman.cpp:55-72), so the products with them are written out as the additions they are;
// adding the reference's explicit zeros would not change a single bit.  State is stored SoA ([component][slot]) so the
// batch reads and writes are coalesced; 336 B of state per track makes this launch-latency bound, not bandwidth bound.
#pragma once
#include "mot_internal.h"
#include "kalman.h"

namespace mot {

__device__ __forceinline__ void inv4_adj(const double (&S)[4][4], double (&Si)[4][4])
{
    // closed form, like Armadillo's inv_tiny (include/armadillo_bits/op_inv_meat.hpp:69-72): adjugate / determinant
    double cof[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double m[9]; int k = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) if (i != r)
#pragma unroll
                for (int j = 0; j < 4; ++j) if (j != c) m[k++] = S[i][j];
            const double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
            cof[r][c] = ((r + c) & 1) ? -d : d;
        }
    double det = 0.0;
#pragma unroll
    for (int c = 0; c < 4; ++c) det += S[0][c] * cof[0][c];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) Si[r][c] = cof[c][r] / det;
}

// predict of the track in slot s; *box: its entry of the caller's box array (type / score stay untouched)
__device__ __forceinline__ void kalman_predict_one(const KalmanState &st, int s, mot_bbox_t *box, int clamp, int fw, int fh)
{
    double x[6], P[6][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = st.x[(long)k * st.cap + s];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) P[r][c] = st.P[(long)(c * 6 + r) * st.cap + s];
    // x = A x  (kalman.h:209): rows 0,2 add vx, rows 1,3 add vy
    x[0] += x[4]; x[1] += x[5]; x[2] += x[4]; x[3] += x[5];
    // P = A P A^T + Q  (kalman.h:210)
    double AP[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) AP[r][c] = (r < 4) ? P[r][c] + P[4 + (r & 1)][c] : P[r][c];
    const double q25 = 1e-2 * 0.25, q50 = 1e-2 * 0.50, q100 = 1e-2 * 1.00;      // Q = Q0 * Qt, kalman.cpp:75-85
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double v = (c < 4) ? AP[r][c] + AP[r][4 + (c & 1)] : AP[r][c];
            double q = 0.0;
            if (r == c) q = (r < 4) ? q25 : q100;
            else if ((r < 4 && c == 4 + (r & 1)) || (c < 4 && r == 4 + (c & 1))) q = q50;
            P[r][c] = v + q;
        }
#pragma unroll
    for (int k = 0; k < 6; ++k) st.x[(long)k * st.cap + s] = x[k];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) st.P[(long)(c * 6 + r) * st.cap + s] = P[r][c];
    // kalman.cpp:112-115: double -> int truncation; type / score are left untouched
    mot_bbox_t b = *box;
    b.l = __double2int_rz(x[0]); b.t = __double2int_rz(x[1]); b.r = __double2int_rz(x[2]); b.b = __double2int_rz(x[3]);
    if (clamp) {                                                                  // top/td.cpp:378-381
        b.l = min(max(0, b.l), fw - 1); b.r = min(max(0, b.r), fw - 1);
        b.t = min(max(0, b.t), fh - 1); b.b = min(max(0, b.b), fh - 1);
    }
    *box = b;
}

// update of the track in slot s with the measurement b
__device__ __forceinline__ void kalman_update_one(const KalmanState &st, int s, const mot_bbox_t b)
{
    double x[6], P[6][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) x[k] = st.x[(long)k * st.cap + s];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) P[r][c] = st.P[(long)(c * 6 + r) * st.cap + s];
    const double z[4] = { (double)b.l, (double)b.t, (double)b.r, (double)b.b };   // kalman.cpp:122-125
    // S = H P H^T + R = P[0:4,0:4] + 512 I  (kalman.cpp:87-88)
    double S[4][4], Si[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) S[r][c] = P[r][c] + (r == c ? 512.0 : 0.0);
    inv4_adj(S, Si);
    // K = P H^T inv(S) = P[:,0:4] Si  (kalman.h:228)
    double K[6][4];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += P[r][k] * Si[k][c];
            K[r][c] = a;
        }
    // x += K (z - H x)  (kalman.h:231-232)
    double ze[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) ze[k] = z[k] - x[k];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) a += K[r][k] * ze[k];
        x[r] += a;
    }
    // Joseph form (kalman.h:235-236): Jf = I - K H; P = Jf P Jf^T + K R K^T
    double Jf[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) Jf[r][c] = (r == c ? 1.0 : 0.0) - (c < 4 ? K[r][c] : 0.0);
    double JP[6][6];
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a += Jf[r][k] * P[k][c];
            JP[r][c] = a;
        }
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a += JP[r][k] * Jf[c][k];
            double kr = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) kr += (K[r][k] * 512.0) * K[c][k];
            P[r][c] = a + kr;
        }
#pragma unroll
    for (int k = 0; k < 6; ++k) st.x[(long)k * st.cap + s] = x[k];
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int r = 0; r < 6; ++r) st.P[(long)(c * 6 + r) * st.cap + s] = P[r][c];
}

// ---- the same update spread over the lanes of an 8-lane group (lanes 0..5 = the six rows of P / K / Jf, lanes 0..3 also the four
// rows of the cofactor matrix): every element is produced by exactly the expression of kalman_update_one above -- same operands, same
// order of the sums -- so the result is the same bit for bit, but a lane's chain of dependent FP64 operations is a sixth as long.
// sm = 132 doubles of shared memory owned by the group; all 32 lanes of the warp must call (s < 0: a group without a track).
constexpr int KALMAN_COOP_DOUBLES = 132;

template <int R> __device__ __forceinline__ void kalman_cof_row(const double (&S)[4][4], double *out)
{
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double m[9]; int k = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) if (i != R)
#pragma unroll
            for (int j = 0; j < 4; ++j) if (j != c) m[k++] = S[i][j];
        const double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        out[c] = ((R + c) & 1) ? -d : d;
    }
}

__device__ __forceinline__ void kalman_update_coop(const KalmanState &st, int s, const mot_bbox_t b, double *sm, int q)
{
    double *const Ps = sm, *const cof = sm + 36, *const Sis = sm + 52, *const Ks = sm + 68, *const Jfs = sm + 92, *const zes = sm + 128;
    const bool row = s >= 0 && q < 6, row4 = s >= 0 && q < 4;
    double xq = 0.0;
    if (row) {
#pragma unroll
        for (int c = 0; c < 6; ++c) Ps[q * 6 + c] = st.P[(long)(c * 6 + q) * st.cap + s];
        xq = st.x[(long)q * st.cap + s];
    }
    if (row4) {
        const double z = (double)(q == 0 ? b.l : q == 1 ? b.t : q == 2 ? b.r : b.b);     // kalman.cpp:122-125
        zes[q] = z - xq;
    }
    __syncwarp();
    // S = H P H^T + R = P[0:4,0:4] + 512 I (kalman.cpp:87-88); cofactors of row q
    double S[4][4];
    if (row4) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) S[r][c] = Ps[r * 6 + c] + (r == c ? 512.0 : 0.0);
        switch (q) {
        case 0: kalman_cof_row<0>(S, cof + 0); break;
        case 1: kalman_cof_row<1>(S, cof + 4); break;
        case 2: kalman_cof_row<2>(S, cof + 8); break;
        default: kalman_cof_row<3>(S, cof + 12); break;
        }
    }
    __syncwarp();
    if (row4) {
        double det = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c) det += S[0][c] * cof[c];
#pragma unroll
        for (int c = 0; c < 4; ++c) Sis[q * 4 + c] = cof[c * 4 + q] / det;       // Si[q][c] = cof[c][q] / det
    }
    __syncwarp();
    double Kr[4], JPr[6];
    if (row) {
        // K = P H^T inv(S) = P[:,0:4] Si  (kalman.h:228)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += Ps[q * 6 + k] * Sis[k * 4 + c];
            Kr[c] = a; Ks[q * 4 + c] = a;
        }
        // x += K (z - H x)  (kalman.h:231-232)
        {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) a += Kr[k] * zes[k];
            xq += a;
            st.x[(long)q * st.cap + s] = xq;
        }
        // Joseph form (kalman.h:235-236): Jf = I - K H; P = Jf P Jf^T + K R K^T
        double Jfr[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) { Jfr[c] = (q == c ? 1.0 : 0.0) - (c < 4 ? Kr[c] : 0.0); Jfs[q * 6 + c] = Jfr[c]; }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a += Jfr[k] * Ps[k * 6 + c];
            JPr[c] = a;
        }
    }
    __syncwarp();
    if (row) {
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) a += JPr[k] * Jfs[c * 6 + k];
            double kr = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) kr += (Kr[k] * 512.0) * Ks[c * 4 + k];
            st.P[(long)(c * 6 + q) * st.cap + s] = a + kr;
        }
    }
    __syncwarp();                          // the group's shared memory may be reused by the caller's next track
}

}  // namespace mot
