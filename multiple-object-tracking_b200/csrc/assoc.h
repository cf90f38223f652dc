// assoc.h -- batched association: cost matrices + Munkres (see assoc.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/mot_b200.h"

namespace mot {

struct AssocLaunch {
    int n_mat;
    const int *T, *D;                                  // [n_mat] tracker / detection counts
    const mot_bbox_t *trk; long trk_stride;            // problem m: trk + m*trk_stride
    const mot_bbox_t *det; long det_stride;
    int cost_mode; double screen_dis;                  // 1 / frame_w (the reference's SCREEN_DIS, top/td.cpp:50)
    double *dist; long dist_stride;                    // cost matrices out (column-major, rows = smaller side); required
    double *work; long work_stride;                    // Munkres working copies (global), >= max_dim^2 doubles each
    int *assign; long assign_stride;                   // one int per ROW
    double *cost;                                      // [n_mat]
    int max_dim;                                       // upper bound of max(T, D) over the batch (sizes shared memory)
};

// cost matrices only (top/td.cpp:386-457)
int assoc_cost(const AssocLaunch &p, cudaStream_t s);
// Munkres on p.dist (nrows = min side, see header) -> p.assign / p.cost.  nrows/ncols given explicitly when
// rows_cols != nullptr ([n_mat][2]); otherwise derived from T, D with the reference's rule (rows = T if T < D else D).
int assoc_solve(const AssocLaunch &p, const int *rows_cols, cudaStream_t s);

}  // namespace mot
