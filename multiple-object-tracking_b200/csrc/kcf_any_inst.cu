// kcf_any_inst.cu -- the instantiations of the any-size kernel, one translation unit per (MODE, EXT) pair (compiled with
// -DANY_MODE=0|1 -DANY_EXT=0|1) so that they build in parallel: stage dumps / 128 registers / 64 registers, with and without strips.
#include "kcf_any_kernel.cuh"

#define CAT4_(a, b, c) a##b##_##c
#define CAT4(a, b, c) CAT4_(a, b, c)

namespace mot {
// v: 0 = stage dumps (1024-thread budget), 1 = 128 registers (CTAs of up to 512 threads), 2 = 64 registers
const void *CAT4(kcf_any_fn_, ANY_MODE, ANY_EXT)(int strips, int v)
{
#define ANY_ROW(S) { (const void *)kcf_any_kernel<ANY_MODE, true, 1024, S, ANY_EXT != 0>, (const void *)kcf_any_kernel<ANY_MODE, false, 512, S, ANY_EXT != 0>, \
                     (const void *)kcf_any_kernel<ANY_MODE, false, 1024, S, ANY_EXT != 0> }
    static const void *const fns[2][3] = { ANY_ROW(false), ANY_ROW(true) };
#undef ANY_ROW
    return fns[strips ? 1 : 0][v];
}
}  // namespace mot
