// overlay.h -- see overlay.cu
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "../../include/mot_b200.h"

namespace mot {

// Entries order[slot_begin[s] .. slot_begin[s+1]) are drawn into frame slot s, in that order.
int overlay_draw(uint8_t *const *d_frame_ptr, int n_slots, int stride, long frame_bytes, const int *d_slot_begin, const int *d_order,
                 const mot_bbox_t *d_boxes, const uint32_t *d_rgb, int thickness, cudaStream_t s);

// colormap[hashcolor(tid + 1) & 255]: the colour the reference fixes at spawn (top/td.cpp:619-620 hashes the counter after
// `tid = tracker_id++`); palette top/td.cpp:652-699 = the xterm 256-colour table as the reference spells it
uint32_t track_color(uint32_t tid);

}  // namespace mot
