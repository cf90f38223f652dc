/*
 * oracle/port/port_td_kcf.c -- TEST INFRASTRUCTURE ONLY.
 * The frame loop of top/td.cpp:343-644 (restated in port_tdloop.inc) built with KCF_TRACKER defined
 * (top/td.cpp:47), over the C restatements of the KCF plugin and of assignmentoptimal.
 */
#include "port_types.h"
void port_rgb2gray(float *, const uint8_t *, int, int, int, int, int);
void port_resize_gray(float *, const float *, int, int, int, int);
void port_cost_matrix(double *, const bbox_t *, int, const bbox_t *, int, int, double);
void *port_kcf_new(const bbox_t *);
void port_kcf_predict(void *, const float *, bbox_t *);
void port_kcf_update(void *, const float *, const bbox_t *);
void port_kcf_delete(void *);
long port_assignmentoptimal(int *, double *, const double *, int, int);

#define TDL_PREFIX(n) port_kcf_##n
#define TDL_EXPORT __attribute__((visibility("default")))
#define TDL_IS_KCF 1
#define TDL_TRK_NEW(pb) port_kcf_new(pb)
#define TDL_TRK_PREDICT(p, g, pb) port_kcf_predict(p, g, pb)
#define TDL_TRK_UPDATE(p, g, pb) port_kcf_update(p, g, pb)
#define TDL_TRK_DELETE(p) port_kcf_delete(p)
#define TDL_ASSIGN(a, c, d, nr, nc) port_assignmentoptimal(a, c, d, nr, nc)
#include "port_tdloop.inc"
