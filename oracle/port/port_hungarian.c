/*
 * oracle/port/port_hungarian.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the reference's Munkres solver, trackers/hungarian/hungarian.cpp:29-368
 * (assignmentoptimal + step2a/2b/3/4/5 + buildassignmentvector + computeassignmentcost), as one
 * loop over an explicit state instead of mutual recursion, and with the three n^2 bool matrices
 * replaced by index vectors (a row or a column never holds more than one star, a row never more
 * than one prime, so "first starred/primed entry in row/column" is a single look-up).  Every
 * DECISION is the reference's:
 *   - row-min reduction when rows <= cols, column-min otherwise (:65-125)
 *   - greedy initial stars in the reference's scan order (:93-101, :127-140)
 *   - zero test fabs(x) < DBL_EPSILON (:95, :249)
 *   - step 3 scans columns ascending / rows ascending, and after covering a row it CONTINUES with the
 *     next column of the same sweep (:246-272)
 *   - step 5 adds h to every covered row, THEN subtracts h from every uncovered column, so a cell that
 *     is both sees (d+h)-h with two roundings (:355-364)
 * which is what bit-exact assignments on tied / degenerate matrices require.
 * dist is column-major nrows x ncols; assignment[row] = col or -1.
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>

__attribute__((visibility("default")))
long port_assignmentoptimal(int *assignment, double *cost, const double *distIn, int nR, int nC)
{
    long n = (long)nR * nC, iters = 0;
    double *d = (double *)malloc(sizeof(double) * (size_t)(n ? n : 1));
    int *starOfRow = (int *)malloc(sizeof(int) * (size_t)(nR + 1)), *starOfCol = (int *)malloc(sizeof(int) * (size_t)(nC + 1));
    int *primeOfRow = (int *)malloc(sizeof(int) * (size_t)(nR + 1));
    char *covR = (char *)calloc((size_t)nR + 1, 1), *covC = (char *)calloc((size_t)nC + 1, 1);
    int minDim;
    *cost = 0;
    for (int r = 0; r < nR; ++r) { assignment[r] = -1; starOfRow[r] = -1; primeOfRow[r] = -1; }
    for (int c = 0; c < nC; ++c) starOfCol[c] = -1;
    for (long i = 0; i < n; ++i) d[i] = distIn[i];

    if (nR <= nC) {
        minDim = nR;
        for (int r = 0; r < nR; ++r) {
            double mn = d[r];
            for (int c = 1; c < nC; ++c) { double v = d[r + (long)nR * c]; if (v < mn) mn = v; }
            for (int c = 0; c < nC; ++c) d[r + (long)nR * c] -= mn;
        }
        for (int r = 0; r < nR; ++r) for (int c = 0; c < nC; ++c)
            if (fabs(d[r + (long)nR * c]) < DBL_EPSILON && !covC[c]) { starOfRow[r] = c; starOfCol[c] = r; covC[c] = 1; break; }
    } else {
        minDim = nC;
        for (int c = 0; c < nC; ++c) {
            double *col = d + (long)nR * c, mn = col[0];
            for (int r = 1; r < nR; ++r) if (col[r] < mn) mn = col[r];
            for (int r = 0; r < nR; ++r) col[r] -= mn;
        }
        for (int c = 0; c < nC; ++c) for (int r = 0; r < nR; ++r)
            if (fabs(d[r + (long)nR * c]) < DBL_EPSILON && !covR[r]) { starOfRow[r] = c; starOfCol[c] = r; covC[c] = 1; covR[r] = 1; break; }
        for (int r = 0; r < nR; ++r) covR[r] = 0;
    }

    enum { S2B, S3, S5, DONE } st = S2B;
    while (st != DONE) {
        ++iters;
        if (st == S2B) {                                   /* :213-236 */
            int ncov = 0;
            for (int c = 0; c < nC; ++c) if (covC[c]) ++ncov;
            st = (ncov == minDim) ? DONE : S3;
        } else if (st == S3) {                             /* :239-279 */
            int zerosFound = 1, aug_r = -1, aug_c = -1;
            while (zerosFound && aug_r < 0) {
                zerosFound = 0;
                for (int c = 0; c < nC && aug_r < 0; ++c) {
                    if (covC[c]) continue;
                    for (int r = 0; r < nR; ++r) {
                        if (covR[r] || !(fabs(d[r + (long)nR * c]) < DBL_EPSILON)) continue;
                        primeOfRow[r] = c;
                        int sc = starOfRow[r];
                        if (sc < 0) { aug_r = r; aug_c = c; }
                        else { covR[r] = 1; covC[sc] = 0; zerosFound = 1; }
                        break;
                    }
                }
            }
            if (aug_r < 0) { st = S5; continue; }
            /* step 4 (:282-334): alternate along starred column / primed row, using the OLD stars */
            {
                int r = aug_r, c = aug_c;
                for (;;) {
                    int sr = starOfCol[c];                 /* old star in this column (or none) */
                    starOfRow[r] = c; starOfCol[c] = r;    /* star the primed zero */
                    if (sr < 0) break;
                    r = sr; c = primeOfRow[r];             /* its row's prime takes over; the old star (r, old c) is dropped */
                }
                for (int i = 0; i < nR; ++i) { primeOfRow[i] = -1; covR[i] = 0; }
                /* step 2a (:193-210): cover every column holding a star (covers are never cleared here) */
                for (int cc = 0; cc < nC; ++cc) if (starOfCol[cc] >= 0) covC[cc] = 1;
                st = S2B;
            }
        } else {                                           /* step 5 (:337-368) */
            double h = DBL_MAX;
            for (int r = 0; r < nR; ++r) if (!covR[r]) for (int c = 0; c < nC; ++c) if (!covC[c]) {
                double v = d[r + (long)nR * c]; if (v < h) h = v;
            }
            for (int r = 0; r < nR; ++r) if (covR[r]) for (int c = 0; c < nC; ++c) d[r + (long)nR * c] += h;
            for (int c = 0; c < nC; ++c) if (!covC[c]) for (int r = 0; r < nR; ++r) d[r + (long)nR * c] -= h;
            st = S3;
        }
    }
    for (int r = 0; r < nR; ++r) {                         /* :161-176, :179-189 */
        assignment[r] = starOfRow[r];
        if (assignment[r] >= 0) *cost += distIn[r + (long)nR * assignment[r]];
    }
    free(d); free(starOfRow); free(starOfCol); free(primeOfRow); free(covR); free(covC);
    return iters;
}
