/*
 * oracle/port/port_overlay.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the overlay step of the tracking thread:
 *   port_draw_rect     <- drawRect                       top/drawlib.c:97-151
 *   port_overlay       <- the three-rectangle loop       top/td.cpp:647-733
 *   port_hashcolor     <- hashcolor                      top/td.cpp:295-305
 *   port_track_color   <- colormap[hashcolor(tid + 1) & 255] top/td.cpp:619-620, 652-699
 * drawRect addresses pixels linearly (PIXEL_AT(y, x) = 3840 y + 3 x for the hard-coded 1280-pixel frame,
 * top/drawlib.c:9-10) and does no clipping; the restatement takes the byte stride and the buffer size and skips bytes
 * that fall outside the buffer (the original would write out of bounds there).  Cross-checked against the compiled
 * original at stride 3840 and against a palette / hash fixture extracted from the reference (tests/golden).
 */
#include "port_types.h"

__attribute__((visibility("default")))
void port_draw_rect(uint8_t *fbuf, int stride_bytes, long nbytes, int left, int top, int right, int bottom, uint32_t rgb)
{
    const uint8_t R = (rgb >> 16) & 0xff, G = (rgb >> 8) & 0xff, B = rgb & 0xff;      /* :106-108 */
    if (top > bottom) { int q = top; top = bottom; bottom = q; }                      /* :112-117 */
    if (left > right) { int q = left; left = right; right = q; }                      /* :119-124 */
    long ot = (long)top * stride_bytes + 3L * left, ob = (long)bottom * stride_bytes + 3L * left;
    for (int i = left; i <= right; ++i, ot += 3, ob += 3) {                           /* :132-136 */
        if (ot >= 0 && ot + 2 < nbytes) { fbuf[ot] = R; fbuf[ot + 1] = G; fbuf[ot + 2] = B; }
        if (ob >= 0 && ob + 2 < nbytes) { fbuf[ob] = R; fbuf[ob + 1] = G; fbuf[ob + 2] = B; }
    }
    long ol = (long)top * stride_bytes + 3L * left, orr = (long)top * stride_bytes + 3L * right;
    for (int i = top; i <= bottom; ++i, ol += stride_bytes, orr += stride_bytes) {    /* :144-150 */
        if (ol >= 0 && ol + 2 < nbytes) { fbuf[ol] = R; fbuf[ol + 1] = G; fbuf[ol + 2] = B; }
        if (orr >= 0 && orr + 2 < nbytes) { fbuf[orr] = R; fbuf[orr + 1] = G; fbuf[orr + 2] = B; }
    }
}

/* top/td.cpp:647-733: for every track in table order, the box and the box shrunk by 1 and by 2 pixels */
__attribute__((visibility("default")))
void port_overlay(uint8_t *fbuf, int stride_bytes, long nbytes, int n, const bbox_t *boxes, const uint32_t *rgb, int thickness)
{
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < thickness; ++k)
            port_draw_rect(fbuf, stride_bytes, nbytes, boxes[j].l + k, boxes[j].t + k, boxes[j].r - k, boxes[j].b - k, rgb[j]);
}

__attribute__((visibility("default")))
uint32_t port_hashcolor(uint32_t a)                                                   /* top/td.cpp:295-305 */
{
    a = (a + 0x7ed55d16u) + (a << 12);
    a = (a ^ 0xc761c23cu) ^ (a >> 19);
    a = (a + 0x165667b1u) + (a << 5);
    a = (a + 0xd3a2646cu) ^ (a << 9);
    a = (a + 0xfd7046c5u) + (a << 3);
    a = (a ^ 0xb55a4f09u) ^ (a >> 16);
    return a;
}

/* the table of top/td.cpp:652-697 is the xterm 256-colour palette with gray entries 241 and 242 spelt 0x606060, 0x666666 */
__attribute__((visibility("default")))
uint32_t port_colormap(int idx)
{
    static const uint32_t sys16[16] = { 0x000000, 0x800000, 0x008000, 0x808000, 0x000080, 0x800080, 0x008080, 0xc0c0c0,
                                        0x808080, 0xff0000, 0x00ff00, 0xffff00, 0x0000ff, 0xff00ff, 0x00ffff, 0xffffff };
    static const uint32_t lv[6] = { 0x00, 0x5f, 0x87, 0xaf, 0xd7, 0xff };
    idx &= 255;
    if (idx < 16) return sys16[idx];
    if (idx < 232) { int q = idx - 16; return (lv[q / 36] << 16) | (lv[(q / 6) % 6] << 8) | lv[q % 6]; }
    if (idx == 241) return 0x606060;
    if (idx == 242) return 0x666666;
    return (8u + 10u * (uint32_t)(idx - 232)) * 0x010101u;
}

__attribute__((visibility("default")))
/* top/td.cpp:619-620: tid = tracker_id++; color = hashcolor(tracker_id) & 255 -> the hash of tid + 1; drawn as colormap[color] (:699) */
uint32_t port_track_color(uint32_t tid) { return port_colormap((int)(port_hashcolor(tid + 1u) & 255u)); }
