// assoc.cuh -- batched association: cost matrices and the reference's Munkres solver, one CTA per problem.
//
// Replaces the cost loops of top/td.cpp:386-457 and assignmentoptimal + step2a/2b/3/4/5 of
// trackers/hungarian/hungarian.cpp:29-368.  Assignments must be BIT-EXACT with the reference, ties included, so the
// solver is not "a" Hungarian algorithm but the reference's, decision for decision:
//   * the same reductions (row minima when rows <= cols, column minima otherwise) and greedy initial stars,
//   * the same zero test fabs(x) < DBL_EPSILON on the same doubles (step 5 adds h to covered rows, then subtracts h
//     from uncovered columns, so doubly-qualified cells see (d+h)-h with both roundings),
//   * step 3's sweep order: columns ascending, first uncovered zero row of the column, and after covering that row the
//     sweep CONTINUES with the next column (uncovered columns to the right are visited in the same sweep).
// What changes is the machinery: the three n^2 bool matrices become index vectors (a row/column holds at most one
// star, a row at most one prime) plus a column-major ZERO BITMAP in shared memory (n^2 bits), so step 3 is a walk over
// set bits by one warp instead of an n^2 scan, and step 5 -- the only full-matrix pass -- is done by the whole CTA with
// coalesced 256-byte rows.  Matrices up to ~160x160 live in shared memory; larger ones stay in global memory / L2.
#pragma once
#include "assoc.h"
#include <cfloat>

namespace mot {

__device__ __forceinline__ double cost_cell(const mot_bbox_t &t, const mot_bbox_t &d, int mode, double screen_dis)
{
    const int maxl = max(t.l, d.l), maxt = max(t.t, d.t), minr = min(t.r, d.r), minb = min(t.b, d.b);
    double dista = 0.0;
    if (mode == MOT_COST_IOU_CLAMPED) {
        const int iw = max(0, minr - maxl), ih = max(0, minb - maxt);
        const double inter = (double)(iw * ih);
        const double uni = __dsub_rn((double)((t.b - t.t) * (t.r - t.l) + (d.b - d.t) * (d.r - d.l)), inter);
        dista = (uni > 0.0) ? __dsub_rn(1.0, __ddiv_rn(inter, uni)) : 1.0;
    } else {
        // top/td.cpp:406-415: centroid distance * SCREEN_DIS
        const int cxi = (t.l + t.r) >> 1, cyi = (t.t + t.b) >> 1, cxj = (d.l + d.r) >> 1, cyj = (d.t + d.b) >> 1;
        dista = __dmul_rn(__dsqrt_rn((double)((cxi - cxj) * (cxi - cxj) + (cyi - cyj) * (cyi - cyj))), screen_dis);
    }
    if (t.type != d.type) dista = __dadd_rn(dista, 1.0);                   // top/td.cpp:416-419
    return dista;
}

// ---------------------------------------------------------------------------------------------------------------------
struct MunkresSmem {
    double *mat;            // optional shared-memory copy of the working matrix
    uint32_t *Zc;           // [nC][nWr] zero bitmap, column-major: bit r of word (c, r>>5)
    uint32_t *covR, *covC;  // cover bit masks
    uint32_t *cand, *candAll; // columns that may hold an UNCOVERED zero (superset) / that hold any zero
    uint32_t *Zr;           // [nR][nWc] row-major copy of the zero bitmap for the greedy start (null when it does not fit)
    int *starOfRow, *starOfCol, *primeOfRow;
    double *redd;           // [32] reduction scratch
    double *rmin;           // [1024] partial row minima of the reduction pass
    int *ctrl;              // [4]: 0 = control word, 1 = aug row, 2 = aug col
};

__device__ __forceinline__ bool tst(const uint32_t *w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }

// next column >= from that is uncovered and flagged as a candidate, or n if none; executed uniformly by a warp
__device__ __forceinline__ int next_candidate(const uint32_t *cov, const uint32_t *cand, int from, int n)
{
    while (from < n) {
        const int w = from >> 5;
        uint32_t bits = ~cov[w] & cand[w] & (0xFFFFFFFFu << (from & 31));
        if (bits) { const int i = (w << 5) + __ffs(bits) - 1; return i < n ? i : n; }
        from = (w + 1) << 5;
    }
    return n;
}

// next index >= from whose bit in `mask` is CLEAR (i.e. next uncovered), or n if none; executed uniformly by a warp
__device__ __forceinline__ int next_clear(const uint32_t *mask, int from, int n)
{
    while (from < n) {
        const int w = from >> 5;
        uint32_t bits = ~mask[w] & (0xFFFFFFFFu << (from & 31));
        if (bits) { const int i = (w << 5) + __ffs(bits) - 1; return i < n ? i : n; }
        from = (w + 1) << 5;
    }
    return n;
}

enum { CTRL_DONE = 1, CTRL_STEP5 = 2, CTRL_FAIL = 3 };

// The solver for ONE problem, run by a whole CTA of NT threads (a multiple of 32, at most 1024): distIn = column-major nR x nC costs,
// work = global working copy (used when the matrix does not fit the smem_mat_doubles of shared memory), assign[nR], *cost_out.
// md = the max_dim the shared-memory carve-up was sized for (munkres_smem_bytes).  Ends with every thread past its last barrier.
template <int NT>
__device__ void munkres_cta(const double *distIn, double *work, const int nR, const int nC, const int md, int *assign, double *cost_out,
                            unsigned char *smem_raw, const int smem_mat_doubles)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    if (nR <= 0 || nC <= 0) { if (tid == 0) *cost_out = 0.0; return; }      // the reference skips the call (top/td.cpp:460)
    const int minDim = nR <= nC ? nR : nC;
    const int nWr = (nR + 31) >> 5, nWc = (nC + 31) >> 5;
    const int mdW = (md + 31) >> 5;
    const int zs = nWr | 1;                                 // odd row stride of the zero bitmap

    MunkresSmem s;
    {
        unsigned char *q = smem_raw;
        s.mat = reinterpret_cast<double *>(q); q += sizeof(double) * (size_t)smem_mat_doubles;
        s.redd = reinterpret_cast<double *>(q); q += sizeof(double) * 32;
        s.rmin = reinterpret_cast<double *>(q); q += sizeof(double) * 1024;
        s.Zc = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * (size_t)md * (mdW | 1);
        s.covR = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * mdW;
        s.covC = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * mdW;
        s.cand = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * mdW;
        s.candAll = reinterpret_cast<uint32_t *>(q); q += sizeof(uint32_t) * mdW;
        s.Zr = (md <= 512) ? reinterpret_cast<uint32_t *>(q) : nullptr; q += (md <= 512) ? sizeof(uint32_t) * (size_t)md * mdW : 0;
        s.starOfRow = reinterpret_cast<int *>(q); q += sizeof(int) * md;
        s.starOfCol = reinterpret_cast<int *>(q); q += sizeof(int) * md;
        s.primeOfRow = reinterpret_cast<int *>(q); q += sizeof(int) * md;
        s.ctrl = reinterpret_cast<int *>(q);
    }
    double *const d = ((long)nR * nC <= smem_mat_doubles) ? s.mat : work;

    // state
    for (int i = tid; i < nR; i += NT) { s.starOfRow[i] = -1; s.primeOfRow[i] = -1; assign[i] = -1; }
    for (int i = tid; i < nC; i += NT) s.starOfCol[i] = -1;
    for (int i = tid; i < mdW; i += NT) { s.covR[i] = 0; s.covC[i] = 0; }

    // Working copy (hungarian.cpp:41-54) + reduction (:65-89 rows, :104-124 columns) + zero bitmap in TWO passes over the input with
    // eight independent loads in flight per thread (the passes are latency-bound: a 256 x 256 matrix is 512 KB in L2).  The minimum
    // of a row / column is taken in any order: `v < mn` over all entries gives the same value whatever the order (the sign of a
    // zero minimum may differ, which no later comparison can see).
    constexpr int UR = NT >= 512 ? 8 : 16;          // fewer threads: more loads in flight per thread (there are registers to spare)
    if (nR <= nC) {
        // warp = (32-row word w, column group g): a contiguous run of columns per group; per-lane running minimum, groups combined
        // through shared memory.  The second pass also collects, per lane = per row, the zero bits of the columns it visits and
        // ORs them word by word into the row-major bitmap Zr that the greedy start reads (no separate transposition pass).
        // (with fewer warps than 32-row words -- more than 32 * NW rows -- a warp takes several words in turn, G = 1)
        const int G = NW >= nWr ? NW / nWr : 1, per = (nC + G - 1) / G;
        if (s.Zr) for (int i = tid; i < nR * nWc; i += NT) s.Zr[i] = 0u;
        for (int wt = warp; wt < nWr * G; wt += NW) {
            const int w = wt % nWr, g = wt / nWr, r = (w << 5) + lane;
            const int cbeg = g * per, cend = min(nC, cbeg + per);
            const bool on = r < nR;
            double mn = INFINITY;
            for (int c0 = cbeg; c0 < cend; c0 += UR) {
                double v[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) { const int c = c0 + u; v[u] = (on && c < cend) ? distIn[r + (long)nR * c] : INFINITY; }
#pragma unroll
                for (int u = 0; u < UR; ++u) if (v[u] < mn) mn = v[u];
            }
            s.rmin[g * (nWr << 5) + (w << 5) + lane] = mn;
        }
        __syncthreads();
        for (int wt = warp; wt < nWr * G; wt += NW) {
            const int w = wt % nWr, g = wt / nWr, r = (w << 5) + lane;
            const int cbeg = g * per, cend = min(nC, cbeg + per);
            const bool on = r < nR;
            double mn = s.rmin[(w << 5) + lane];
            for (int q = 1; q < G; ++q) { const double o = s.rmin[q * (nWr << 5) + (w << 5) + lane]; if (o < mn) mn = o; }
            uint32_t acc = 0; int cw = cbeg >> 5;
            for (int c0 = cbeg; c0 < cend; c0 += UR) {
                double v[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) { const int c = c0 + u; v[u] = (on && c < cend) ? distIn[r + (long)nR * c] : 1.0; }
#pragma unroll
                for (int u = 0; u < UR; ++u) {
                    const int c = c0 + u;
                    if (c >= cend) break;                                  // warp-uniform
                    const double x = __dsub_rn(v[u], mn);
                    if (on) d[r + (long)nR * c] = x;
                    const bool z = on && fabs(x) < DBL_EPSILON;
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, z);
                    if (lane == 0) s.Zc[c * zs + w] = word;
                    if ((c >> 5) != cw) { if (s.Zr && acc) atomicOr(&s.Zr[r * nWc + cw], acc); acc = 0; cw = c >> 5; }
                    acc |= (z ? 1u : 0u) << (c & 31);
                }
            }
            if (s.Zr && acc) atomicOr(&s.Zr[r * nWc + cw], acc);
        }
    } else {
        // one warp per column, lanes along the rows
        for (int c = warp; c < nC; c += NW) {
            const double *col = distIn + (long)nR * c;
            double mn = INFINITY;
            for (int r0 = lane; r0 < nR; r0 += 32 * UR) {
                double v[UR];
#pragma unroll
                for (int u = 0; u < UR; ++u) { const int r = r0 + 32 * u; v[u] = r < nR ? col[r] : INFINITY; }
#pragma unroll
                for (int u = 0; u < UR; ++u) if (v[u] < mn) mn = v[u];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xFFFFFFFFu, mn, off); if (o < mn) mn = o; }
            for (int w = 0; w < nWr; ++w) {
                const int r = (w << 5) + lane;
                bool z = false;
                if (r < nR) { const double x = __dsub_rn(col[r], mn); d[r + (long)nR * c] = x; z = fabs(x) < DBL_EPSILON; }
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, z);
                if (lane == 0) s.Zc[c * zs + w] = word;
            }
        }
    }
    __syncthreads();
    // Column filters for step 3: candAll = columns with a zero, cand = columns with a zero in an uncovered row.  While step 3
    // runs, rows only get covered, so `cand` computed before it stays a superset; it is refreshed after step 5 (new zeros) and
    // reset to candAll after step 4 (all rows uncovered).  Pure acceleration: the sweep still checks every column it visits.
    auto refresh_candidates = [&]() {
        for (int cb = warp * 32; cb < nC; cb += NT) {
            const int c = cb + lane;
            uint32_t any = 0, unc = 0;
            if (c < nC) for (int w = 0; w < nWr; ++w) { const uint32_t z = s.Zc[c * zs + w]; any |= z; unc |= z & ~s.covR[w]; }
            const uint32_t wa = __ballot_sync(0xFFFFFFFFu, any != 0), wu = __ballot_sync(0xFFFFFFFFu, unc != 0);
            if (lane == 0) { s.candAll[cb >> 5] = wa; s.cand[cb >> 5] = wu; }
        }
    };
    refresh_candidates();
    __syncthreads();

    // Greedy initial stars (hungarian.cpp:91-101 / :126-140) are sequential by nature: row r takes its first zero column that no
    // earlier row took.  When the FIRST zero columns of all rows are distinct -- the normal case of tracking, where every track has
    // its own nearest detection -- that is what the greedy does for every row (by induction no earlier row can have covered it), so
    // all rows star their first zero in parallel and only a collision falls back to the one-warp sequential form.
    bool stars_done = false;
    if (nR <= nC && s.Zr) {
        if (tid == 0) s.ctrl[3] = 0;
        __syncthreads();
        for (int r = tid; r < nR; r += NT) {
            int f = -1;
            for (int w = 0; w < nWc; ++w) { const uint32_t word = s.Zr[r * nWc + w]; if (word) { f = (w << 5) + __ffs(word) - 1; break; } }
            if (f >= 0) {
                if (atomicCAS(&s.starOfCol[f], -1, r) != -1) s.ctrl[3] = 1;
                else s.starOfRow[r] = f;
            }
        }
        __syncthreads();
        stars_done = s.ctrl[3] == 0;
        if (stars_done) {
            for (int cb = warp * 32; cb < nC; cb += NT) {
                const int c = cb + lane;
                const uint32_t word = __ballot_sync(0xFFFFFFFFu, c < nC && s.starOfCol[c] >= 0);
                if (lane == 0) s.covC[cb >> 5] = word;
            }
        } else {
            for (int i = tid; i < nR; i += NT) s.starOfRow[i] = -1;
            for (int i = tid; i < nC; i += NT) s.starOfCol[i] = -1;
        }
        __syncthreads();
    }
    if (warp == 0 && !stars_done) {
        if (nR <= nC && s.Zr) {
            // first uncovered zero column of each row, rows ascending (hungarian.cpp:91-101): one ballot per row
            for (int r = 0; r < nR; ++r) {
                const uint32_t mw = (lane < nWc) ? (s.Zr[r * nWc + lane] & ~s.covC[lane]) : 0u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mw != 0);
                if (bal) {
                    const int fl = __ffs(bal) - 1;
                    const int cs = (fl << 5) + __ffs(__shfl_sync(0xFFFFFFFFu, mw, fl)) - 1;
                    if (lane == 0) { s.starOfRow[r] = cs; s.starOfCol[cs] = r; s.covC[cs >> 5] |= 1u << (cs & 31); }
                    __syncwarp();
                }
            }
        } else if (nR <= nC) {
            for (int r = 0; r < nR; ++r) {
                for (int cb = 0; cb < nC; cb += 32) {
                    const int c = cb + lane;
                    const bool ok = (c < nC) && ((s.Zc[c * zs + (r >> 5)] >> (r & 31)) & 1u) && !tst(s.covC, c);
                    const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
                    if (bal) {
                        const int cs = cb + __ffs(bal) - 1;
                        if (lane == 0) { s.starOfRow[r] = cs; s.starOfCol[cs] = r; s.covC[cs >> 5] |= 1u << (cs & 31); }
                        __syncwarp();
                        break;
                    }
                }
            }
        } else {
            for (int c = 0; c < nC; ++c) {
                const uint32_t mw = (lane < nWr) ? (s.Zc[c * zs + lane] & ~s.covR[lane]) : 0u;
                const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mw != 0);
                if (bal) {
                    const int fl = __ffs(bal) - 1;
                    const int r = (fl << 5) + __ffs(__shfl_sync(0xFFFFFFFFu, mw, fl)) - 1;
                    if (lane == 0) { s.starOfRow[r] = c; s.starOfCol[c] = r; s.covC[c >> 5] |= 1u << (c & 31); s.covR[r >> 5] |= 1u << (r & 31); }
                    __syncwarp();
                }
            }
            if (lane < mdW) s.covR[lane] = 0;
            __syncwarp();
        }
    }
    __syncthreads();

    long guard = 0;
    const long guard_max = 8L * md * md + 1024;             // the reference never terminates on -inf / NaN costs; we do
    bool after_step5 = false;
    for (;;) {
        if (warp == 0) {
            int ctrl = 0;
            for (;;) {
                if (!after_step5) {
                    // step 2b (hungarian.cpp:213-236)
                    int cnt = (lane < nWc) ? __popc(s.covC[lane]) : 0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, off);
                    if (cnt == minDim) { ctrl = CTRL_DONE; break; }
                }
                after_step5 = false;
                // step 3 (hungarian.cpp:239-279)
                int aug_r = -1, aug_c = -1;
                bool zerosFound = true;
                while (zerosFound && aug_r < 0) {
                    zerosFound = false;
                    for (int c = next_candidate(s.covC, s.cand, 0, nC); c < nC; c = next_candidate(s.covC, s.cand, c + 1, nC)) {
                        const uint32_t mw = (lane < nWr) ? (s.Zc[c * zs + lane] & ~s.covR[lane]) : 0u;
                        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, mw != 0);
                        if (!bal) { if (lane == 0) s.cand[c >> 5] &= ~(1u << (c & 31)); __syncwarp(); continue; }   // stays empty until step 4 / 5
                        const int fl = __ffs(bal) - 1;
                        const int r = (fl << 5) + __ffs(__shfl_sync(0xFFFFFFFFu, mw, fl)) - 1;
                        const int sc = s.starOfRow[r];
                        if (lane == 0) s.primeOfRow[r] = c;
                        if (sc < 0) { aug_r = r; aug_c = c; __syncwarp(); break; }
                        if (lane == 0) { s.covR[r >> 5] |= 1u << (r & 31); s.covC[sc >> 5] &= ~(1u << (sc & 31)); }
                        zerosFound = true;
                        __syncwarp();
                    }
                }
                if (aug_r < 0) { ctrl = CTRL_STEP5; break; }
                // step 4 (hungarian.cpp:282-334): flip stars along the alternating path (old stars drive the walk)
                if (lane == 0) {
                    int r = aug_r, c = aug_c;
                    for (;;) {
                        const int sr = s.starOfCol[c];
                        s.starOfRow[r] = c; s.starOfCol[c] = r;
                        if (sr < 0) break;
                        r = sr; c = s.primeOfRow[r];
                    }
                }
                __syncwarp();
                for (int i = lane; i < nR; i += 32) s.primeOfRow[i] = -1;
                if (lane < mdW) { s.covR[lane] = 0; s.cand[lane] = s.candAll[lane]; }      // all rows uncovered again
                // step 2a (hungarian.cpp:193-210): cover every column that holds a star
                for (int w = lane; w < nWc; w += 32) {
                    uint32_t bits = 0;
                    for (int b = 0; b < 32; ++b) { const int c = (w << 5) + b; if (c < nC && s.starOfCol[c] >= 0) bits |= 1u << b; }
                    s.covC[w] |= bits;
                }
                __syncwarp();
            }
            if (lane == 0) s.ctrl[0] = ctrl;
        }
        __syncthreads();
        const int ctrl = s.ctrl[0];
        if (ctrl != CTRL_STEP5) break;
        if (++guard > guard_max) { if (tid == 0) s.ctrl[0] = CTRL_FAIL; __syncthreads(); break; }

        // step 5 (hungarian.cpp:337-368), whole CTA.  h = smallest uncovered element.  A warp keeps ONE 32-row word of the matrix (its
        // row-cover bits live in a register, nothing is divided per cell) and walks over a contiguous run of columns, eight at a time so
        // that eight independent loads are in flight; words without an uncovered row (first pass) and covered columns without a covered
        // row in the word (second pass) are skipped by the whole warp.  The passes are bound by instruction issue, not by memory.
        constexpr int U = 8;
        const int G5 = NW >= nWr ? NW / nWr : 1, per5 = (nC + G5 - 1) / G5;
        double h = DBL_MAX;
        for (int wt = warp; wt < nWr * G5; wt += NW) {
            const int w = wt % nWr, g = wt / nWr, r = (w << 5) + lane;
            const bool rowu = r < nR && !((s.covR[w] >> lane) & 1u);
            if (__ballot_sync(0xFFFFFFFFu, rowu) == 0u) continue;             // warp-uniform
            const int cbeg = g * per5, cend = min(nC, cbeg + per5);
            const double *const dr = d + r;
            for (int c0 = cbeg; c0 < cend; c0 += U) {
                double v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u;
                    v[u] = (rowu && c < cend && !tst(s.covC, c)) ? dr[(long)nR * c] : DBL_MAX;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) if (v[u] < h) h = v[u];
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) { const double o = __shfl_xor_sync(0xFFFFFFFFu, h, off); if (o < h) h = o; }
        if (lane == 0) s.redd[warp] = h;
        __syncthreads();
        h = s.redd[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) { const double o = s.redd[w]; if (o < h) h = o; }
        // add h to covered rows, then subtract h from uncovered columns; refresh the zero bits of touched cells
        for (int wt = warp; wt < nWr * G5; wt += NW) {
            const int w = wt % nWr, g = wt / nWr, r = (w << 5) + lane;
            const uint32_t crw = s.covR[w];
            const bool inr = r < nR, rc = inr && ((crw >> lane) & 1u);
            const int cbeg = g * per5, cend = min(nC, cbeg + per5);
            double *const dr = d + r;
            for (int c0 = cbeg; c0 < cend; c0 += U) {
                double v[U]; uint32_t st5[U];                                  // bit 0: column covered, bit 1: task live, bit 2: this cell is touched
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c = c0 + u;
                    const bool in = c < cend, cc = in && tst(s.covC, c);
                    const bool live = in && !(cc && crw == 0u);                // covered column, no covered row in this word: untouched
                    const bool touch = live && inr && (rc || !cc);
                    st5[u] = (cc ? 1u : 0u) | (live ? 2u : 0u) | (touch ? 4u : 0u);
                    v[u] = touch ? dr[(long)nR * c] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!(st5[u] & 2u)) continue;                              // warp-uniform (depends on the column and the word only)
                    const int c = c0 + u;
                    bool z = (s.Zc[c * zs + w] >> lane) & 1u;
                    if (st5[u] & 4u) {
                        double x = v[u];
                        if (rc) x = __dadd_rn(x, h);
                        if (!(st5[u] & 1u)) x = __dsub_rn(x, h);
                        dr[(long)nR * c] = x;
                        z = fabs(x) < DBL_EPSILON;
                    }
                    const uint32_t word = __ballot_sync(0xFFFFFFFFu, z);
                    __syncwarp();                                             // every lane has read the old word
                    if (lane == 0) s.Zc[c * zs + w] = word;
                }
            }
        }
        after_step5 = true;
        __syncthreads();
        refresh_candidates();                                     // new zeros appeared, old ones vanished
        __syncthreads();
    }

    // buildassignmentvector + computeassignmentcost (hungarian.cpp:161-189); rows summed in ascending order.  The per-row costs
    // are fetched in parallel into shared memory (the zero bitmap is dead by now), then added up by one thread in row order.
    const bool fail = s.ctrl[0] == CTRL_FAIL;
    double *const rowcost = reinterpret_cast<double *>(s.Zc);      // md*(mdW|1) words >= nR doubles for every md >= 1? guarded below
    const bool stage = (size_t)nR * sizeof(double) <= sizeof(uint32_t) * (size_t)md * (mdW | 1);
    __syncthreads();
    for (int r = tid; r < nR; r += NT) {
        const int c = s.starOfRow[r];
        assign[r] = fail ? -1 : c;
        if (stage) rowcost[r] = (!fail && c >= 0) ? distIn[r + (long)nR * c] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double cst = 0.0;
        if (fail) cst = nan("");
        else if (stage) {
            // unassigned rows hold +0.0, and x + 0.0 == x for every x this sum can reach (it starts at +0.0, so it is never -0.0):
            // no branch, loads ahead of the chain of additions
            int r = 0;
            for (; r + 8 <= nR; r += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = rowcost[r + u];
#pragma unroll
                for (int u = 0; u < 8; ++u) cst = __dadd_rn(cst, v[u]);
            }
            for (; r < nR; ++r) cst = __dadd_rn(cst, rowcost[r]);
        } else for (int r = 0; r < nR; ++r) {
            const int c = s.starOfRow[r];
            if (c >= 0) cst = __dadd_rn(cst, distIn[r + (long)nR * c]);
        }
        *cost_out = cst;
    }
}


// dynamic shared memory of munkres_cta for problems up to md x md with mat_doubles of the matrix held in shared memory
__host__ __device__ inline size_t munkres_smem_bytes(int md, int mat_doubles)
{
    const int mdW = (md + 31) >> 5;
    return sizeof(double) * (size_t)mat_doubles + sizeof(double) * (32 + 1024) + sizeof(uint32_t) * ((size_t)md * (mdW | 1) + 4 * mdW + (md <= 512 ? (size_t)md * mdW : 0)) +
           sizeof(int) * (3 * (size_t)md + 4);
}
// how much of an md x md matrix to keep in shared memory: all of it when the whole carve-up stays under 200 KB, else nothing
__host__ __device__ inline int munkres_mat_doubles(int md) { return munkres_smem_bytes(md, md * md) > 200 * 1024 ? 0 : md * md; }

}  // namespace mot
