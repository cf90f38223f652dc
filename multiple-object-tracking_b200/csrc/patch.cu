// patch.cu -- the two patch helpers of the reference as stand-alone device kernels behind host-pointer entry points.
//
//   rgb2Gray                    top/drawlib.c:192-240   crop of a BGR u8 frame -> gray f32, column-major rows x cols
//   bilinearInterpolationGray   top/drawlib.c:542-637   literal: BOTH buffers are treated as row-major height x width
//
// The batched path never calls these (crop, gray and resize are the front end of the fused KCF kernels); they exist so that
// the reference's C-linkage symbols (declared top/td.cpp:245-261) can be exported by host/tracker_shim.cpp and an unmodified
// top/td.cpp runs every numeric step on the GPU.  Host pointers in, host pointers out: one upload, one kernel, one download.
#include "mot_ctx.h"
#include "fhog_common.cuh"

namespace mot {

// crop[c * rows + r] = gray(src[r][c]), src = packed crop rows (3 * cols bytes each)
__global__ void gray_crop_kernel(const uint8_t *src, int src_pitch, int rows, int cols, float *dst)
{
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < rows * cols; k += gridDim.x * blockDim.x) {
        const int r = k / cols, c = k - r * cols;                   // consecutive threads: consecutive pixels of a frame row
        dst[c * rows + r] = bgr_gray(src + (long)r * src_pitch + 3 * c);          // drawlib.c:234-235
    }
}

// drawlib.c:551-634, unfused f32 operations in the reference's order
__global__ void bilinear_gray_kernel(float *dst, const float *src, int hs, int ws, int h, int w)
{
    const float xs = __fdiv_rn((float)ws, (float)w), ys = __fdiv_rn((float)hs, (float)h);
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < h * w; k += gridDim.x * blockDim.x) {
        const int y = k / w, x = k - y * w;
        const float sx = __fmul_rn((float)x, xs), sy = __fmul_rn((float)y, ys);
        const int x0 = __float2int_rz(sx), y0 = __float2int_rz(sy);
        const float fx = __fsub_rn(sx, (float)x0), fy = __fsub_rn(sy, (float)y0);
        const float ifx = __fsub_rn(1.0f, fx), ify = __fsub_rn(1.0f, fy);
        const int x1 = (x0 + 1 >= ws) ? x0 : x0 + 1, y1 = (y0 + 1 >= hs) ? y0 : y0 + 1;
        const float c1 = src[y0 * ws + x0], c2 = src[y0 * ws + x1], c3 = src[y1 * ws + x0], c4 = src[y1 * ws + x1];
        const float l0 = __fadd_rn(__fmul_rn(ifx, c1), __fmul_rn(fx, c2));
        const float l1 = __fadd_rn(__fmul_rn(ifx, c3), __fmul_rn(fx, c4));
        dst[k] = __fadd_rn(__fmul_rn(ify, l0), __fmul_rn(fy, l1));
    }
}

}  // namespace mot

extern "C" {

int mot_rgb2gray_host(mot_ctx_t *c, const uint8_t *host_bgr, int stride_bytes, int l, int t, int r, int b, float *gray_host_out)
{
    if (!c || !host_bgr || !gray_host_out || stride_bytes <= 0) return mot_fail(MOT_ERR_ARG, "mot_rgb2gray_host: bad argument");
    if (t > b) { const int q = t; t = b; b = q; }                   // drawlib.c:203-215
    if (l > r) { const int q = l; l = r; r = q; }
    const int rows = b - t + 1, cols = r - l + 1;
    if ((long)rows * cols > (1L << 28)) return mot_fail(MOT_ERR_SHAPE, "mot_rgb2gray_host: crop %dx%d is too large", rows, cols);
    CU(cudaSetDevice(c->device));
    const size_t pitch = ((size_t)cols * 3 + 15) & ~(size_t)15, npx = (size_t)rows * cols;
    CU(c->d_scratch.ensure(pitch * rows)); CU(c->d_gray.ensure(npx));
    // exactly the bytes the reference reads: rows x (3 * cols) starting at PIXEL_AT(top, left) (drawlib.c:220-222), no clipping
    CU(cudaMemcpy2DAsync(c->d_scratch.p, pitch, host_bgr + (long)t * stride_bytes + 3L * l, (size_t)stride_bytes, (size_t)cols * 3, rows,
                         cudaMemcpyHostToDevice, c->stream));
    const unsigned blocks = (unsigned)std::min<size_t>((npx + 255) / 256, 2048);
    gray_crop_kernel<<<blocks, 256, 0, c->stream>>>(reinterpret_cast<const uint8_t *>(c->d_scratch.p), (int)pitch, rows, cols, c->d_gray.p);
    CU(cudaGetLastError());
    c->launches += 1;
    CU(cudaMemcpyAsync(gray_host_out, c->d_gray.p, sizeof(float) * npx, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

int mot_resize_gray_host(mot_ctx_t *c, float *dst_host, const float *src_host, int h_src, int w_src, int h, int w)
{
    if (!c || !dst_host || !src_host || h_src <= 0 || w_src <= 0 || h <= 0 || w <= 0) return mot_fail(MOT_ERR_ARG, "mot_resize_gray_host: bad argument");
    if ((long)h_src * w_src > (1L << 28) || (long)h * w > (1L << 28)) return mot_fail(MOT_ERR_SHAPE, "mot_resize_gray_host: image too large");
    CU(cudaSetDevice(c->device));
    const size_t ns = (size_t)h_src * w_src, nd = (size_t)h * w;
    CU(c->d_scratch.ensure(sizeof(float) * ns)); CU(c->d_gray.ensure(nd));
    CU(cudaMemcpyAsync(c->d_scratch.p, src_host, sizeof(float) * ns, cudaMemcpyHostToDevice, c->stream));
    const unsigned blocks = (unsigned)std::min<size_t>((nd + 255) / 256, 2048);
    bilinear_gray_kernel<<<blocks, 256, 0, c->stream>>>(c->d_gray.p, reinterpret_cast<const float *>(c->d_scratch.p), h_src, w_src, h, w);
    CU(cudaGetLastError());
    c->launches += 1;
    CU(cudaMemcpyAsync(dst_host, c->d_gray.p, sizeof(float) * nd, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return 0;
}

}  // extern "C"
