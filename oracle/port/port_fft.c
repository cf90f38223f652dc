/*
 * oracle/port/port_fft.c -- TEST INFRASTRUCTURE ONLY.  See port_fft.h for what it restates
 * (the FFTW 3.3.5 r2c/c2r calls at trackers/kcf.cpp:134,180,189,142,265,399).
 *
 * Plain recursive mixed-radix Cooley-Tukey in double precision: n = p*m with p the
 * smallest prime factor; p interleaved sub-transforms of size m, then an O(n*p)
 * combine.  Works for any n (prime n degenerates to the O(n^2) definition).
 */
#include "port_fft.h"
#include <math.h>
#include <stdlib.h>

static int smallest_factor(int n)
{
    for (int p = 2; (long)p * p <= n; ++p)
        if (n % p == 0) return p;
    return n;
}

/* tw[q] = exp(sign*2*pi*i*q/N) for the top-level N; tws = N / n at this level. */
static void rec(int n, const pf_cpx *in, int is, pf_cpx *out, const pf_cpx *tw, int N, int tws, pf_cpx *scratch)
{
    if (n == 1) { out[0] = in[0]; return; }
    int p = smallest_factor(n), m = n / p;
    for (int r = 0; r < p; ++r)
        rec(m, in + (long)r * is, is * p, out + (long)r * m, tw, N, tws * p, scratch + p);
    for (int k = 0; k < m; ++k) {
        for (int r = 0; r < p; ++r) scratch[r] = out[(long)r * m + k];
        for (int q = 0; q < p; ++q) {
            long kk = (long)k + (long)q * m;   /* output index */
            double sr = 0.0, si = 0.0;
            for (int r = 0; r < p; ++r) {
                long idx = ((long)r * kk % n) * tws % N;
                double wr = tw[idx].re, wi = tw[idx].im;
                sr += scratch[r].re * wr - scratch[r].im * wi;
                si += scratch[r].re * wi + scratch[r].im * wr;
            }
            out[kk].re = sr; out[kk].im = si;
        }
    }
}

__attribute__((visibility("default")))
void pf_dft(int n, const pf_cpx *in, int istride, pf_cpx *out, int sign)
{
    pf_cpx *tw = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n);
    pf_cpx *scratch = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)(n + 64) * 2);
    const double two_pi = 6.283185307179586476925286766559;
    for (int q = 0; q < n; ++q) {
        double a = two_pi * (double)q / (double)n;
        tw[q].re = cos(a);
        tw[q].im = (sign < 0) ? -sin(a) : sin(a);
    }
    rec(n, in, istride, out, tw, n, 1, scratch);
    free(tw); free(scratch);
}

__attribute__((visibility("default")))
void pf_r2c_2d(int n0, int n1, const float *in, float *out)
{
    int nh = n1 / 2 + 1;
    pf_cpx *row = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n1);
    pf_cpx *rout = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n1);
    pf_cpx *half = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n0 * nh);
    pf_cpx *col = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n0);
    for (int j = 0; j < n0; ++j) {
        for (int b = 0; b < n1; ++b) { row[b].re = in[(long)j * n1 + b]; row[b].im = 0.0; }
        pf_dft(n1, row, 1, rout, -1);
        for (int k = 0; k < nh; ++k) half[(long)j * nh + k] = rout[k];
    }
    for (int k = 0; k < nh; ++k) {
        pf_dft(n0, half + k, nh, col, -1);
        for (int j = 0; j < n0; ++j) {
            out[2 * ((long)j * nh + k) + 0] = (float)col[j].re;
            out[2 * ((long)j * nh + k) + 1] = (float)col[j].im;
        }
    }
    free(row); free(rout); free(half); free(col);
}

__attribute__((visibility("default")))
void pf_c2r_2d(int n0, int n1, const float *in, float *out)
{
    int nh = n1 / 2 + 1;
    pf_cpx *half = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n0 * nh);
    pf_cpx *col = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n0);
    pf_cpx *row = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n1);
    pf_cpx *rout = (pf_cpx *)malloc(sizeof(pf_cpx) * (size_t)n1);
    /* inverse along n0 on each stored column k */
    for (int k = 0; k < nh; ++k) {
        for (int j = 0; j < n0; ++j) {
            half[(long)j * nh + k].re = in[2 * ((long)j * nh + k) + 0];
            half[(long)j * nh + k].im = in[2 * ((long)j * nh + k) + 1];
        }
    }
    for (int k = 0; k < nh; ++k) {
        pf_dft(n0, half + k, nh, col, +1);
        for (int j = 0; j < n0; ++j) half[(long)j * nh + k] = col[j];
    }
    /* Hermitian-extend each row along n1 and inverse; keep the real part */
    for (int j = 0; j < n0; ++j) {
        for (int k = 0; k < nh; ++k) row[k] = half[(long)j * nh + k];
        for (int k = nh; k < n1; ++k) {
            row[k].re = half[(long)j * nh + (n1 - k)].re;
            row[k].im = -half[(long)j * nh + (n1 - k)].im;
        }
        /* c2r semantics: imaginary parts of DC (and Nyquist for even n1) do not contribute */
        row[0].im = 0.0;
        if ((n1 & 1) == 0) row[n1 / 2].im = 0.0;
        pf_dft(n1, row, 1, rout, +1);
        for (int b = 0; b < n1; ++b) out[(long)j * n1 + b] = (float)rout[b].re;
    }
    free(half); free(col); free(row); free(rout);
}
