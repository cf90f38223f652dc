/*
 * oracle/port/port_types.h -- TEST INFRASTRUCTURE ONLY.
 * POD types of the path, layout-identical to top/cnntype.h:36-47 (field order l,t,b,r).
 */
#ifndef PORT_TYPES_H
#define PORT_TYPES_H
#include <stdint.h>

#ifndef __CNN_TYPE_H__          /* the compiled-reference wrappers already have the original */
typedef struct port_bbox_s { int l, t, b, r; int type; float score; } bbox_t;
#endif

/* cost modes of the association step */
enum { PORT_COST_REF_CENTROID = 0, PORT_COST_IOU_CLAMPED = 1 };

/* tracker kinds: which of the two link-time plugins of the reference is emulated */
enum { PORT_TRACKER_KALMAN = 0, PORT_TRACKER_KCF = 1 };

#endif
