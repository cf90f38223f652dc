// plugin_client.cpp -- TEST INFRASTRUCTURE: a translation unit written against the REFERENCE's declarations, the way
// top/td.cpp is, to prove that libmot_b200.so is link-compatible with it.
//
// With -DREF_CNNTYPE_H='"/root/reference/top/cnntype.h"' the reference's own header supplies bbox_t / bbox_chain_t; where the
// reference is absent (the GPU box) the two typedefs are restated from top/cnntype.h:36-47 -- the struct TAGS matter, they are
// part of the mangled names.  The prototypes are the ones top/td.cpp:229-261 declares (C++ linkage for the tracker plugin and
// assignmentoptimal, C linkage for the patch helpers).  client_kcf_sequence() then makes the calls the tracking thread makes
// for one track: spawn + first update (top/td.cpp:612-644), and per frame crop / resize / predict / clamp (:344-384) and
// crop / resize / update (:550-582).
#include <stdint.h>
#include <stdlib.h>
#ifdef REF_CNNTYPE_H
#include REF_CNNTYPE_H
#else
typedef struct _bbox_pos_s { int l, t, b, r; int type; float score; } bbox_t;
typedef struct _bbox_chain_s { int nbox; bbox_t bbox[128]; } bbox_chain_t;
#define MTCNN_IMGW (1280)
#define MTCNN_IMGH ( 720)
#endif

void tracker_predict(void * ptracker, float * rgb, bbox_t * pbox);
void tracker_update (void * ptracker, float * rgb, bbox_t * pbox);
void tracker_delete (void * ptracker                            );
void * tracker_new  (bbox_t * pbox                              );

void assignmentoptimal(int *assignment, double *cost, double *distMatrixIn, int nOfRows, int nOfColumns);

extern "C" void rgb2Gray(float * pgra, uint8_t * prgb, int32_t left, int32_t top, int32_t right, int32_t bottom);
extern "C" void bilinearInterpolationGray(float * pdst, const float * psrc, int rows_s, int cols_s, int rows_d, int cols_d);

#define cmin(x, y)  (((x) < (y)) ? (x) : (y))
#define cmax(x, y)  (((x) > (y)) ? (x) : (y))
#define make_x_in_range(x)  (cmin(cmax(0, (x)), (MTCNN_IMGW-1)))
#define make_y_in_range(y)  (cmin(cmax(0, (y)), (MTCNN_IMGH-1)))

// frames: nframes BGR 1280x720 images back to back; boxes_out[k] = the track's box after frame k (predict, clamp, update with
// the predicted box = the unassigned branch).  Returns 0, or -1 when tracker_new failed.
extern "C" int client_kcf_sequence(uint8_t *frames, int nframes, const bbox_t *first, bbox_t *boxes_out)
{
    const long fbytes = 3L * MTCNN_IMGW * MTCNN_IMGH;
    bbox_t bbox = *first;
    void *trk = tracker_new(&bbox);
    if (!trk) return -1;
    float *gray = (float *)malloc(sizeof(float) * 2 * MTCNN_IMGW * MTCNN_IMGH);
    const int rows = bbox.b - bbox.t + 1, cols = bbox.r - bbox.l + 1;
    rgb2Gray(gray, frames, bbox.l, bbox.t, bbox.r, bbox.b);
    tracker_update(trk, gray, &bbox);
    for (int k = 1; k < nframes; ++k) {
        uint8_t *img = frames + k * fbytes;
        for (int pass = 0; pass < 2; ++pass) {
            float *src = gray + MTCNN_IMGH * MTCNN_IMGH;
            const int rows_b = bbox.b - bbox.t + 1, cols_b = bbox.r - bbox.l + 1;
            if (rows_b == rows && cols_b == cols) rgb2Gray(gray, img, bbox.l, bbox.t, bbox.r, bbox.b);
            else {
                rgb2Gray(src, img, bbox.l, bbox.t, bbox.r, bbox.b);
                bilinearInterpolationGray(gray, src, rows_b, cols_b, rows, cols);
            }
            if (pass == 0) {
                tracker_predict(trk, gray, &bbox);
                bbox.l = make_x_in_range(bbox.l); bbox.r = make_x_in_range(bbox.r);
                bbox.t = make_y_in_range(bbox.t); bbox.b = make_y_in_range(bbox.b);
            } else tracker_update(trk, gray, &bbox);
        }
        boxes_out[k] = bbox;
    }
    boxes_out[0] = *first;
    free(gray);
    tracker_delete(trk);
    return 0;
}

extern "C" void client_assign(int *assignment, double *cost, double *dist, int nr, int nc) { assignmentoptimal(assignment, cost, dist, nr, nc); }
extern "C" int client_sizes(void) { return (int)sizeof(bbox_t) * 1000 + (int)(sizeof(bbox_chain_t) % 1000); }
