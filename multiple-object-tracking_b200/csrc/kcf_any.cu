// kcf_any.cu -- host side of the fused any-size KCF kernel (kernel: kcf_any_kernel.cuh, instantiated in kcf_any_inst.cu): sizing and launch.
#include "kcf_any.cuh"
#include <cstdlib>

namespace mot {

const void *kcf_any_fn_0_0(int strips, int v); const void *kcf_any_fn_1_0(int strips, int v);
const void *kcf_any_fn_0_1(int strips, int v); const void *kcf_any_fn_1_1(int strips, int v);

size_t kcf_any_smem_bytes(int hr, int wc, int lut_floats)
{
    const AnyGeo g = any_geo(hr, wc, lut_floats);
    return g.ok ? (size_t)g.total * sizeof(float) : 0;
}

size_t kcf_any_scratch_bytes(int hr, int wc, int lut_floats)
{
    const AnyGeo g = any_geo(hr, wc, lut_floats);
    return (g.ok && g.strips) ? (size_t)18 * g.os * sizeof(float) : 0;
}

int kcf_launch_any(int mode, const KcfLaunch &p, const AnyTablesDev &at, size_t smem_bytes, int threads, int ctas_per_sm, int *err_flag,
                   float *scratch, long scratch_stride_floats, int max_ctas, cudaStream_t s)
{
    const FhogTablesDev &t = p.tab;
    int lut_floats = 2 * (2 << t.rsqrt_bits) + ((2 * t.bin_nseg + 3) & ~3);
    if (lut_floats > 8192) lut_floats = 0;
    const bool dump = p.dump.gray != nullptr, strips = scratch != nullptr;
    // two register budgets: 128 registers per thread (at most 512 threads per SM in total) or 64 (1024 threads per SM)
    bool r128 = threads * (ctas_per_sm < 1 ? 1 : ctas_per_sm) <= 512;
    if (getenv("MOT_ANY_THREADS")) threads = atoi(getenv("MOT_ANY_THREADS"));
    if (getenv("MOT_ANY_CTAS")) ctas_per_sm = atoi(getenv("MOT_ANY_CTAS"));
    if (getenv("MOT_ANY_R128")) r128 = atoi(getenv("MOT_ANY_R128")) != 0;
    if (r128 && threads > 512) threads = 512;
    const int pm = mode == KCF_MODE_PREDICT ? 0 : 1, v = dump ? 0 : r128 ? 1 : 2;
    const bool ext = p.ext.gaussian || p.ext.subpixel || p.ext.padding > 1.0f || p.ext.osf > 0.0f;
    const void *fn = ext ? (pm ? kcf_any_fn_1_1(strips, v) : kcf_any_fn_0_1(strips, v)) : (pm ? kcf_any_fn_1_0(strips, v) : kcf_any_fn_0_0(strips, v));
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int cap = sms * (ctas_per_sm < 1 ? 1 : ctas_per_sm);
    if (max_ctas > 0 && cap > max_ctas) cap = max_ctas;
    const int grid = p.n_jobs < cap ? p.n_jobs : cap;
    int smem_floats = (int)(smem_bytes / sizeof(float));
    int *err = err_flag;
    void *args[7] = { (void *)&p, (void *)&at, (void *)&lut_floats, (void *)&smem_floats, (void *)&err, (void *)&scratch, (void *)&scratch_stride_floats };
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3((unsigned)threads), args, smem_bytes, s);
    return (int)e;
}

}  // namespace mot
