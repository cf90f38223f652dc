// kcf_inst.cu -- one translation unit per window size (compiled with -DKCF_HR=.. -DKCF_WC=..) so the sizes build in parallel.
#include "kcf_fused.cuh"

#define CAT3_(a, b, c) a##b##_##c
#define CAT3(a, b, c) CAT3_(a, b, c)

namespace mot {
int CAT3(kcf_launch_, KCF_HR, KCF_WC)(int mode, const KcfLaunch &p, cudaStream_t s) { return kcf_launch_size<KCF_HR, KCF_WC>(mode, p, s); }
size_t CAT3(kcf_smem_, KCF_HR, KCF_WC)(int lut_floats) { return smem_bytes<KCF_HR, KCF_WC>(lut_floats); }
}  // namespace mot
