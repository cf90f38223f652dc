// kalman.cu -- batched launches of the Kalman predict / update (device code: kalman.cuh).
#include "kalman.cuh"

namespace mot {

__global__ void kalman_init_kernel(KalmanState st, int n, const int *slots, const mot_bbox_t *boxes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
    const mot_bbox_t b = boxes[i];
    // tracker_new, kalman.cpp:147-163: x0 = [l,t,r,b,0,0]; P0 = 1e4 I (kalman.cpp:90-91)
    const double x0[6] = { (double)b.l, (double)b.t, (double)b.r, (double)b.b, 0.0, 0.0 };
    for (int k = 0; k < 6; ++k) st.x[(long)k * st.cap + s] = x0[k];
    for (int c = 0; c < 6; ++c) for (int r = 0; r < 6; ++r) st.P[(long)(c * 6 + r) * st.cap + s] = (r == c) ? 1e4 : 0.0;
}

__global__ void kalman_predict_kernel(KalmanState st, int n, const int *slots, mot_bbox_t *boxes, int clamp, int fw, int fh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int s = slots[i];
    if (s < 0) return;                         // inactive entry of a device-resident track table
    kalman_predict_one(st, s, boxes + i, clamp, fw, fh);
}

// eight lanes per track (kalman_update_coop), eight tracks per 64-thread CTA
__global__ void kalman_update_kernel(KalmanState st, int n, const int *slots, const mot_bbox_t *boxes)
{
    __shared__ double sm[8 * KALMAN_COOP_DOUBLES];
    const int i = blockIdx.x * 8 + (threadIdx.x >> 3);
    const int s = i < n ? slots[i] : -1;
    kalman_update_coop(st, s, s >= 0 ? boxes[i] : mot_bbox_t{}, sm + (threadIdx.x >> 3) * KALMAN_COOP_DOUBLES, threadIdx.x & 7);
}

int kalman_init(const KalmanState &st, int n, const int *d_slots, const mot_bbox_t *d_boxes, cudaStream_t s)
{
    if (n <= 0) return 0;
    kalman_init_kernel<<<(n + 127) / 128, 128, 0, s>>>(st, n, d_slots, d_boxes);
    return (int)cudaGetLastError();
}
int kalman_predict(const KalmanState &st, int n, const int *d_slots, mot_bbox_t *d_boxes, int clamp, int fw, int fh, cudaStream_t s)
{
    if (n <= 0) return 0;
    kalman_predict_kernel<<<(n + 63) / 64, 64, 0, s>>>(st, n, d_slots, d_boxes, clamp, fw, fh);
    return (int)cudaGetLastError();
}
int kalman_update(const KalmanState &st, int n, const int *d_slots, const mot_bbox_t *d_boxes, cudaStream_t s)
{
    if (n <= 0) return 0;
    kalman_update_kernel<<<(n + 7) / 8, 64, 0, s>>>(st, n, d_slots, d_boxes);
    return (int)cudaGetLastError();
}

}  // namespace mot
