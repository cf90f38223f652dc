/*
 * oracle/port/port_fft.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Double-precision mixed-radix DFT used (a) by the FFTW shim that completes the
 * compiled reference (oracle/capi/fftw_shim.cpp) and (b) by the C restatement of
 * trackers/kcf.cpp (oracle/port/port_kcf.c).
 *
 * The reference calls FFTW 3.3.5 single precision (libfftw3f-3), whose source is NOT
 * under /root/reference (only include/fftw3.h and a Windows DLL are vendored), at
 *   trackers/kcf.cpp:134  fftwf_plan_dft_r2c_2d (labels)
 *   trackers/kcf.cpp:180  fftwf_plan_dft_r2c_2d (31 feature channels)
 *   trackers/kcf.cpp:189  fftwf_plan_dft_c2r_2d (response)
 *   trackers/kcf.cpp:142,265,399  fftwf_execute
 * The DFT is defined mathematically (FFTW manual, "What FFTW Really Computes"):
 *   r2c: Y[j][k] = sum_a sum_b X[a][b] exp(-2 pi i (j a/n0 + k b/n1)),  k = 0..n1/2
 *   c2r: un-normalised inverse of the Hermitian-extended half spectrum
 * so the oracle evaluates that definition in double precision and rounds to float.
 * Parity at this boundary is UNPINNED by the reference (it has no tests or vectors).
 */
#ifndef PORT_FFT_H
#define PORT_FFT_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double re, im; } pf_cpx;

/* out[k] = sum_a in[a*istride] * exp(sign * 2 pi i * a k / n), out contiguous. */
void pf_dft(int n, const pf_cpx *in, int istride, pf_cpx *out, int sign);

/* 2-D real -> half complex.  in: n0 x n1 floats (n1 contiguous); out: n0 x (n1/2+1) float pairs. */
void pf_r2c_2d(int n0, int n1, const float *in, float *out);

/* 2-D half complex -> real, un-normalised.  in: n0 x (n1/2+1) float pairs; out: n0 x n1 floats. */
void pf_c2r_2d(int n0, int n1, const float *in, float *out);

#ifdef __cplusplus
}
#endif
#endif
