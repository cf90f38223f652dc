// oracle/capi/yolo_capi.cpp -- TEST INFRASTRUCTURE ONLY.
//
// C wrapper around the reference's YOLO3 post-processing (detectors/yolo3.cpp:141-356: decode_netout, correct_yolo_boxes,
// sort, do_nms), compiled from the reference's own text (YOLO_EXTRACT_INC, produced by oracle/tools/extract_yolo.py in a
// scratch directory because the file as a whole needs TensorFlow / OpenCV / Windows).  The per-image driver around those
// four functions is part of tensorRunB (:487-527), which cannot be extracted; it is restated here line by line.
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
using namespace std;

#include "ref_common.h"
#include "top/cnntype.h"          // bbox_t (the reference file, by path)
#include YOLO_EXTRACT_INC

extern "C" __attribute__((visibility("default")))
int ref_yolo_post(const float *out0, const float *out1, const float *out2, const int *anchors, float obj_thresh, float nms_thresh,
                  int tensor_height, int tensor_width, int image_height, int image_width, int num_classes, bbox_t *out, int max_out)
{
    const int grid_h = tensor_height / 32, grid_w = tensor_width / 32;                       // :405-406
    std::vector<predecode_t> boxes; std::vector<detection_t> nboxes; std::vector<detection_t> cboxes;
    decode_netout(boxes, out0, anchors + 12, obj_thresh, tensor_height, tensor_width, grid_h << 0, grid_w << 0, num_classes);   // :496-498
    decode_netout(boxes, out1, anchors + 6, obj_thresh, tensor_height, tensor_width, grid_h << 1, grid_w << 1, num_classes);
    decode_netout(boxes, out2, anchors + 0, obj_thresh, tensor_height, tensor_width, grid_h << 2, grid_w << 2, num_classes);
    if (boxes.empty()) return 0;                 // the reference pushes an uninitialised box here (:214-217): undefined, not reproduced
    correct_yolo_boxes(cboxes, boxes, tensor_height, tensor_width, image_height, image_width, num_classes);                  // :500
    do_nms(nboxes, cboxes, nms_thresh, num_classes);                                                                         // :501
    int nbox = 0;
    for (size_t i = 0; i < nboxes.size(); i++) {                                                                             // :505-526
        nboxes[i].ymin = max(nboxes[i].ymin, 0);
        nboxes[i].xmin = max(nboxes[i].xmin, 0);
        nboxes[i].ymax = min(nboxes[i].ymax, (image_height - 1));
        nboxes[i].xmax = min(nboxes[i].xmax, (image_width - 1));
        if ((nboxes[i].ymin > nboxes[i].ymax) || (nboxes[i].xmin > nboxes[i].xmax) || (nboxes[i].ymin < 0) || (nboxes[i].xmin < 0) ||
            (nboxes[i].xmax >= image_width) || (nboxes[i].ymax >= image_height))
            continue;
        if (nbox >= max_out) break;              // bbox_chain_t holds 128 (top/cnntype.h:46); the reference does not check
        out[nbox].t = nboxes[i].ymin; out[nbox].l = nboxes[i].xmin; out[nbox].b = nboxes[i].ymax; out[nbox].r = nboxes[i].xmax;
        out[nbox].type = nboxes[i].classes; out[nbox].score = nboxes[i].objectness;
        ++nbox;
    }
    return nbox;
}
