// yolo_post.h -- see yolo_post.cu
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "../../include/mot_b200.h"

namespace mot {

size_t yolo_cand_bytes();
size_t yolo_nms_smem_bytes();
int yolo_cap();
// d_out[k]: network output of scale k (grid (th/32 << k) x (tw/32 << k), 3 anchors, 5 + nc values), device memory.
// Four launches on stream s; d_nout receives the number of detections written to d_boxes (at most max_out).
int yolo_post_launch(const float *const d_out[3], const int *d_anchors, float obj_thresh, float nms_thresh, int th, int tw, int ih, int iw, int nc,
                     void *d_cand, int *d_count, mot_bbox_t *d_boxes, int max_out, int *d_nout, cudaStream_t s);

}  // namespace mot
