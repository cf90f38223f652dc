// kcf_dispatch.cu -- size dispatch of the fused KCF kernels (one instantiation per supported window size, kcf_inst.cu).
#include "mot_internal.h"

#define MOT_KCF_SIZES(X) X(8, 8) X(8, 16) X(16, 8) X(16, 16) X(16, 32) X(32, 16) X(32, 32) X(8, 32) X(32, 8)

namespace mot {
#define X(H, W) int kcf_launch_##H##_##W(int mode, const KcfLaunch &p, cudaStream_t s); size_t kcf_smem_##H##_##W(int lut_floats);
MOT_KCF_SIZES(X)
#undef X

int kcf_launch_fast(int mode, int hr, int wc, const KcfLaunch &p, cudaStream_t s)
{
#define X(H, W) if (hr == H && wc == W) return kcf_launch_##H##_##W(mode, p, s);
    MOT_KCF_SIZES(X)
#undef X
    return -1000;
}

size_t kcf_fast_smem_bytes(int hr, int wc)
{
#define X(H, W) if (hr == H && wc == W) return kcf_smem_##H##_##W(4256);
    MOT_KCF_SIZES(X)
#undef X
    return 0;
}
}  // namespace mot
