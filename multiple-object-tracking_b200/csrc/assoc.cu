// assoc.cu -- batched association launches: cost matrices and the reference's Munkres solver, one CTA per problem (device code: assoc.cuh).
#include "assoc.cuh"

namespace mot {

constexpr int MUNKRES_THREADS = 1024;

__global__ void cost_kernel(const AssocLaunch p)
{
    const int m = blockIdx.y;
    const int T = p.T[m], D = p.D[m];
    const mot_bbox_t *trk = p.trk + (long)m * p.trk_stride, *det = p.det + (long)m * p.det_stride;
    double *dist = p.dist + (long)m * p.dist_stride;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < T * D; idx += gridDim.x * blockDim.x) {
        int i, j;
        if (T < D) { j = idx / T; i = idx - j * T; }      // rows = trackers  (top/td.cpp:388-421)
        else       { i = idx / D; j = idx - i * D; }      // rows = detections (top/td.cpp:424-456)
        dist[idx] = cost_cell(trk[i], det[j], p.cost_mode, p.screen_dis);
    }
}

__global__ void __launch_bounds__(MUNKRES_THREADS, 1) munkres_kernel(const AssocLaunch p, const int *rows_cols, const int smem_mat_doubles)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // persistent CTAs over the problems (the launch may hold fewer CTAs than problems: see assoc_solve)
    for (int m = blockIdx.x; m < p.n_mat; m += gridDim.x) {
        int nR, nC;
        if (rows_cols) { nR = rows_cols[2 * m]; nC = rows_cols[2 * m + 1]; }
        else { const int T = p.T[m], D = p.D[m]; if (T < D) { nR = T; nC = D; } else { nR = D; nC = T; } }
        munkres_cta<MUNKRES_THREADS>(p.dist + (long)m * p.dist_stride, p.work + (long)m * p.work_stride, nR, nC, p.max_dim, p.assign + (long)m * p.assign_stride,
                                     p.cost + m, smem_raw, smem_mat_doubles);
        __syncthreads();
    }
}

int assoc_cost(const AssocLaunch &p, cudaStream_t s)
{
    if (p.n_mat <= 0) return 0;
    const int cells = p.max_dim * p.max_dim;
    dim3 grid((unsigned)((cells + 255) / 256 > 64 ? 64 : (cells + 255) / 256), (unsigned)p.n_mat);
    if (grid.x == 0) grid.x = 1;
    cost_kernel<<<grid, 256, 0, s>>>(p);
    return (int)cudaGetLastError();
}

int assoc_solve(const AssocLaunch &p, const int *rows_cols, cudaStream_t s)
{
    if (p.n_mat <= 0) return 0;
    if (p.max_dim > 1024 || p.max_dim < 1) return -1001;
    const int mat_doubles = munkres_mat_doubles(p.max_dim);
    const size_t bytes = munkres_smem_bytes(p.max_dim, mat_doubles);
    cudaError_t e = cudaFuncSetAttribute((const void *)munkres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    // one CTA per problem (measured on config 5: capping the resident CTAs so that the working copies stay in L2 does not pay --
    // a problem's rate is set by its own dependent passes, not by HBM)
    const int grid = p.n_mat;
    munkres_kernel<<<grid, MUNKRES_THREADS, bytes, s>>>(p, rows_cols, mat_doubles);
    return (int)cudaGetLastError();
}

}  // namespace mot
