// overlay.cu -- tracking rectangles drawn into the BGR frames on the device.
//
// Replaces the overlay loop of the tracking thread (top/td.cpp:647-733): for every track, in table order, three nested
// one-pixel rectangles (box, box shrunk by 1, by 2) in the track's colour, each drawn by drawRect (top/drawlib.c:97-151):
// swap top/bottom and left/right if reversed, then the two horizontal runs (bytes R, G, B per pixel) and the two
// vertical runs, addressed linearly as y * stride + 3 * x with no clipping.  Later tracks overwrite earlier ones where
// rectangles overlap, so the order is part of the result: one CTA per frame walks that frame's entries in order, all
// threads draw one entry's pixels, a barrier separates entries.  Writes that fall outside the frame buffer (the reference
// would corrupt memory there) are dropped; an x beyond the row end lands in the next row exactly as in the reference.
#include "mot_internal.h"
#include "overlay.h"

namespace mot {

constexpr int OVERLAY_THREADS = 256;

__global__ void __launch_bounds__(OVERLAY_THREADS) overlay_kernel(uint8_t *const *frame_ptr, int stride, long frame_bytes, const int *slot_begin,
                                                                   const int *order, const mot_bbox_t *boxes, const uint32_t *rgb, int thickness)
{
    const int slot = blockIdx.x;
    const int e0 = slot_begin[slot], e1 = slot_begin[slot + 1];
    if (e0 == e1) return;
    uint8_t *const fb = frame_ptr[slot];
    for (int e = e0; e < e1; ++e) {
        const int i = order[e];
        const mot_bbox_t bx = boxes[i];
        const uint32_t col = rgb[i];
        const uint8_t R = (col >> 16) & 0xff, G = (col >> 8) & 0xff, B = col & 0xff;      // drawlib.c:106-108
        for (int k = 0; k < thickness; ++k) {                                              // td.cpp:701-732
            int left = bx.l + k, top = bx.t + k, right = bx.r - k, bottom = bx.b - k;
            if (top > bottom) { const int q = top; top = bottom; bottom = q; }             // drawlib.c:112-124
            if (left > right) { const int q = left; left = right; right = q; }
            const int nh = right - left + 1, nv = bottom - top + 1;
            for (int q = threadIdx.x; q < 2 * (nh + nv); q += OVERLAY_THREADS) {
                int x, y;
                if (q < 2 * nh) { x = left + (q >> 1); y = (q & 1) ? bottom : top; }       // drawlib.c:126-136
                else { const int v = q - 2 * nh; y = top + (v >> 1); x = (v & 1) ? right : left; }   // drawlib.c:138-150
                const long off = (long)y * stride + 3L * x;
                if (off >= 0 && off + 2 < frame_bytes) { fb[off] = R; fb[off + 1] = G; fb[off + 2] = B; }
            }
        }
        __syncthreads();
    }
}

int overlay_draw(uint8_t *const *d_frame_ptr, int n_slots, int stride, long frame_bytes, const int *d_slot_begin, const int *d_order,
                 const mot_bbox_t *d_boxes, const uint32_t *d_rgb, int thickness, cudaStream_t s)
{
    if (n_slots <= 0) return 0;
    overlay_kernel<<<n_slots, OVERLAY_THREADS, 0, s>>>(d_frame_ptr, stride, frame_bytes, d_slot_begin, d_order, d_boxes, d_rgb, thickness);
    return (int)cudaGetLastError();
}

// The reference's table is the xterm 256-colour palette -- 16 system colours, the 6x6x6 cube on levels {00,5f,87,af,d7,ff},
// a 24-step gray ramp 08 + 0a*i -- except for two gray entries that it spells 0x606060 and 0x666666 (top/td.cpp:693).
uint32_t track_color(uint32_t tid)
{
    // top/td.cpp:619-620: `tid = tracker_id++; color = hashcolor(tracker_id) & 255` -- the hash is taken of the counter AFTER
    // the increment, i.e. of tid + 1 (wrapping like the reference's uint32_t counter)
    uint32_t a = tid + 1u;                              // hashcolor, top/td.cpp:295-305
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    const uint32_t idx = a & 255u;
    static const uint32_t sys16[16] = { 0x000000, 0x800000, 0x008000, 0x808000, 0x000080, 0x800080, 0x008080, 0xc0c0c0,
                                        0x808080, 0xff0000, 0x00ff00, 0xffff00, 0x0000ff, 0xff00ff, 0x00ffff, 0xffffff };
    static const uint32_t lv[6] = { 0x00, 0x5f, 0x87, 0xaf, 0xd7, 0xff };
    if (idx < 16) return sys16[idx];
    if (idx < 232) { const uint32_t q = idx - 16; return (lv[q / 36] << 16) | (lv[(q / 6) % 6] << 8) | lv[q % 6]; }
    if (idx == 241) return 0x606060;
    if (idx == 242) return 0x666666;
    return (8u + 10u * (idx - 232)) * 0x010101u;
}

}  // namespace mot
