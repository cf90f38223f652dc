// tracker_shim.cpp -- the reference's per-object plugin symbols on top of the batched C ABI (batch of one).
//
// The reference selects its tracker at LINK time: trackers/kalman.cpp:131-163 and trackers/kcf.cpp:455-491 export the same
// four C++-linkage functions (declared top/td.cpp:229-232), and trackers/hungarian/hungarian.cpp:29 exports
// assignmentoptimal (top/td.cpp:234).  Linking this file (plus libmot_b200.so) instead of kalman.cpp / kcf.cpp /
// hungarian.cpp gives top/td.cpp the same symbols, with every call executed on the GPU.  The tracker kind and the frame
// size, compile-time choices in the reference (#define KCF_TRACKER top/td.cpp:47; MTCNN_IMGW/H top/cnntype.h:5-6), come
// from mot_shim_configure() or the environment (MOT_TRACKER=kcf|kalman, MOT_FRAME_W, MOT_FRAME_H, MOT_MAX_TRACKS).
// Like the reference (top/td.cpp:622) there is no error channel in these signatures: failures are printed and the call
// becomes a no-op; mot_last_error() keeps the text.
#include "../../include/mot_b200.h"

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// top/cnntype.h:36-41 -- the reference's own struct TAG, because it is part of the mangled names top/td.cpp links against
// (_Z11tracker_newP11_bbox_pos_s, ...); layout-identical to mot_bbox_t, which is all the C ABI below ever sees.
typedef struct _bbox_pos_s { int l, t, b, r; int type; float score; } bbox_t;
static_assert(sizeof(bbox_t) == sizeof(mot_bbox_t) && offsetof(bbox_t, b) == offsetof(mot_bbox_t, b) && offsetof(bbox_t, r) == offsetof(mot_bbox_t, r) &&
              offsetof(bbox_t, type) == offsetof(mot_bbox_t, type) && offsetof(bbox_t, score) == offsetof(mot_bbox_t, score), "bbox_t layout");
static inline mot_bbox_t *mb(bbox_t *p) { return reinterpret_cast<mot_bbox_t *>(p); }

static mot_ctx_t *g_ctx = nullptr;
static int g_kind = -1, g_w = 1280, g_h = 720, g_cap = 256;     // reference defaults: cnntype.h:5-6, td.cpp:12

extern "C" int mot_shim_configure(int tracker_kind, int frame_w, int frame_h, int max_tracks)
{
    if (g_ctx) { mot_ctx_destroy(g_ctx); g_ctx = nullptr; }
    g_kind = tracker_kind; g_w = frame_w; g_h = frame_h; g_cap = max_tracks;
    return mot_ctx_create(&g_ctx, 0, g_w, g_h, g_cap, 1, g_kind);
}

extern "C" mot_ctx_t *mot_shim_context(void)
{
    if (g_ctx) return g_ctx;
    if (g_kind < 0) {
        const char *k = getenv("MOT_TRACKER");
        g_kind = (k && !strcmp(k, "kcf")) ? MOT_TRACKER_KCF : MOT_TRACKER_KALMAN;     // the shipped project links Kalman
        if (getenv("MOT_FRAME_W")) g_w = atoi(getenv("MOT_FRAME_W"));
        if (getenv("MOT_FRAME_H")) g_h = atoi(getenv("MOT_FRAME_H"));
        if (getenv("MOT_MAX_TRACKS")) g_cap = atoi(getenv("MOT_MAX_TRACKS"));
    }
    if (mot_ctx_create(&g_ctx, 0, g_w, g_h, g_cap, 1, g_kind)) { fprintf(stderr, "[mot_b200] %s\n", mot_last_error()); g_ctx = nullptr; }
    return g_ctx;
}

static inline int handle_of(void *p) { return (int)(reinterpret_cast<intptr_t>(p)) - 1; }

void *tracker_new(bbox_t *pbox)
{
    mot_ctx_t *c = mot_shim_context();
    int h = -1;
    if (!c || mot_tracker_new_batch(c, 1, mb(pbox), &h)) { fprintf(stderr, "[mot_b200] tracker_new: %s\n", mot_last_error()); return nullptr; }
    return reinterpret_cast<void *>(static_cast<intptr_t>(h + 1));
}

void tracker_predict(void *ptracker, float *rgb, bbox_t *pbox)
{
    if (!ptracker || !g_ctx) return;
    if (mot_predict_gray(g_ctx, handle_of(ptracker), rgb, mb(pbox))) fprintf(stderr, "[mot_b200] tracker_predict: %s\n", mot_last_error());
}

void tracker_update(void *ptracker, float *rgb, bbox_t *pbox)
{
    if (!ptracker || !g_ctx) return;
    if (mot_update_gray(g_ctx, handle_of(ptracker), rgb, mb(pbox))) fprintf(stderr, "[mot_b200] tracker_update: %s\n", mot_last_error());
}

void tracker_delete(void *ptracker)
{
    if (!ptracker || !g_ctx) return;
    const int h = handle_of(ptracker);
    mot_tracker_delete_batch(g_ctx, 1, &h);
}

void assignmentoptimal(int *assignment, double *cost, double *distMatrixIn, int nOfRows, int nOfColumns)
{
    mot_ctx_t *c = mot_shim_context();
    *cost = 0;
    for (int r = 0; r < nOfRows; ++r) assignment[r] = -1;
    if (!c || mot_assign_batch(c, 1, &nOfRows, &nOfColumns, distMatrixIn, 0, assignment, 0, cost))
        fprintf(stderr, "[mot_b200] assignmentoptimal: %s\n", mot_last_error());
}

// ---- the C-linkage patch helpers top/td.cpp declares at :245-261 (implemented by top/drawlib.c:192-240, 542-637 in the reference).
// The reference hard-codes a 1280-pixel frame row (PIXEL_AT / SCREEN_WIDTH, top/drawlib.c:9-10); here the row is the configured
// frame width, which is the same thing for the reference's 1280x720 default.
extern "C" void rgb2Gray(float *pgra, uint8_t *prgb, int32_t left, int32_t top, int32_t right, int32_t bottom)
{
    mot_ctx_t *c = mot_shim_context();
    if (!c || mot_rgb2gray_host(c, prgb, 3 * g_w, left, top, right, bottom, pgra)) fprintf(stderr, "[mot_b200] rgb2Gray: %s\n", mot_last_error());
}

extern "C" void bilinearInterpolationGray(float *pdst, const float *psrc, int rows_s, int cols_s, int rows_d, int cols_d)
{
    mot_ctx_t *c = mot_shim_context();
    if (!c || mot_resize_gray_host(c, pdst, psrc, rows_s, cols_s, rows_d, cols_d)) fprintf(stderr, "[mot_b200] bilinearInterpolationGray: %s\n", mot_last_error());
}

// extern "C" doors onto the C++-linkage symbols above, for ctypes-driven tests
extern "C" void *mot_shim_tracker_new(bbox_t *b) { return tracker_new(b); }
extern "C" void mot_shim_tracker_predict(void *p, float *g, bbox_t *b) { tracker_predict(p, g, b); }
extern "C" void mot_shim_tracker_update(void *p, float *g, bbox_t *b) { tracker_update(p, g, b); }
extern "C" void mot_shim_tracker_delete(void *p) { tracker_delete(p); }
extern "C" void mot_shim_assignmentoptimal(int *a, double *c, double *d, int nr, int nc) { assignmentoptimal(a, c, d, nr, nc); }
