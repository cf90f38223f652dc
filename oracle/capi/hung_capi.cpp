/*
 * oracle/capi/hung_capi.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 * extern "C" door onto the compiled, unmodified trackers/hungarian/hungarian.cpp:29
 * (assignmentoptimal has C++ linkage in the reference, top/td.cpp:234).
 */
#include "ref_common.h"
void assignmentoptimal(int *assignment, double *cost, double *distMatrixIn, int nOfRows, int nOfColumns);
REF_API void ref_assignmentoptimal(int *assignment, double *cost, double *dist, int nrows, int ncols)
{
    assignmentoptimal(assignment, cost, dist, nrows, ncols);
}
