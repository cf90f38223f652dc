// kalman.h -- batched Kalman tracker state and launchers (see kalman.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/mot_b200.h"

namespace mot {

// SoA over track slots: x[k*cap + slot] (k < 6), P[(c*6 + r)*cap + slot] (column-major like arma::mat)
struct KalmanState { double *x; double *P; int cap; };

int kalman_init(const KalmanState &st, int n, const int *d_slots, const mot_bbox_t *d_boxes, cudaStream_t s);
int kalman_predict(const KalmanState &st, int n, const int *d_slots, mot_bbox_t *d_boxes, int clamp, int fw, int fh, cudaStream_t s);
int kalman_update(const KalmanState &st, int n, const int *d_slots, const mot_bbox_t *d_boxes, cudaStream_t s);

}  // namespace mot
