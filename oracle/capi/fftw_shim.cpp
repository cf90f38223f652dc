/*
 * oracle/capi/fftw_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Supplies the four FFTW 3.3.5 (single precision) entry points that the compiled reference
 * trackers/kcf.cpp needs (call sites kcf.cpp:134,180,189 plan; :142,265,399 execute;
 * :143,230,232 destroy).  FFTW's source is not under /root/reference and libfftw3f is not in
 * this image, so the mathematically defined DFT is supplied by one of two providers:
 *
 *   "dft64"  (default)  double-precision mixed-radix DFT rounded to float (oracle/port/port_fft.c).
 *                       Used for every parity check.
 *   "mkl"               oneMKL DFTI single precision, resolved at run time with dlopen/dlsym from
 *                       torch's libtorch_cpu.so (it exports the Dfti* entry points), thread limit 1.
 *                       Used ONLY to time the reference CPU path fairly (an optimised FFT like the
 *                       FFTW the reference ships with); selected with REF_FFT_PROVIDER=mkl and a
 *                       path in REF_FFT_MKL_LIB.  Falls back to dft64 if the library cannot be loaded.
 *
 * ref_fft_provider() reports which one is live so every CPU number can say what it used.
 */
#include <fftw3.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../port/port_fft.h"

struct fftwf_plan_s {
    int kind;          /* 0 = r2c, 1 = c2r */
    int n0, n1;
    void *in, *out;
    void *mkl;         /* DFTI descriptor or NULL */
};

/* ---- optional MKL DFTI provider ------------------------------------------------------- */
typedef long (*dfti_create_md_t)(void **, int, long, long *);
typedef long (*dfti_set_t)(void *, int, ...);
typedef long (*dfti_commit_t)(void *);
typedef long (*dfti_compute_t)(void *, void *, ...);
typedef long (*dfti_free_t)(void **);

static struct {
    int tried, ok;
    dfti_create_md_t create; dfti_set_t set; dfti_commit_t commit;
    dfti_compute_t fwd, bwd; dfti_free_t free_;
} g_mkl;

enum { DFTI_CONJUGATE_EVEN_STORAGE = 10, DFTI_PLACEMENT = 11, DFTI_INPUT_STRIDES = 12,
       DFTI_OUTPUT_STRIDES = 13, DFTI_THREAD_LIMIT = 27, DFTI_REAL = 33,
       DFTI_COMPLEX_COMPLEX = 39, DFTI_NOT_INPLACE = 44 };

static int mkl_available(void)
{
    if (g_mkl.tried) return g_mkl.ok;
    g_mkl.tried = 1;
    const char *want = getenv("REF_FFT_PROVIDER");
    if (!want || strcmp(want, "mkl") != 0) return 0;
    const char *lib = getenv("REF_FFT_MKL_LIB");
    void *h = dlopen(lib ? lib : "libtorch_cpu.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "[fftw_shim] mkl requested but dlopen failed: %s\n", dlerror()); return 0; }
    g_mkl.create = (dfti_create_md_t)dlsym(h, "DftiCreateDescriptor_s_md");
    g_mkl.set    = (dfti_set_t)dlsym(h, "DftiSetValue");
    g_mkl.commit = (dfti_commit_t)dlsym(h, "DftiCommitDescriptor");
    g_mkl.fwd    = (dfti_compute_t)dlsym(h, "DftiComputeForward");
    g_mkl.bwd    = (dfti_compute_t)dlsym(h, "DftiComputeBackward");
    g_mkl.free_  = (dfti_free_t)dlsym(h, "DftiFreeDescriptor");
    g_mkl.ok = g_mkl.create && g_mkl.set && g_mkl.commit && g_mkl.fwd && g_mkl.bwd && g_mkl.free_;
    if (!g_mkl.ok) fprintf(stderr, "[fftw_shim] mkl requested but Dfti* symbols missing\n");
    return g_mkl.ok;
}

static void *mkl_make(int n0, int n1, int backward)
{
    void *d = NULL;
    long len[2] = { n0, n1 };
    long sr[3] = { 0, n1, 1 }, sc[3] = { 0, n1 / 2 + 1, 1 };
    if (g_mkl.create(&d, DFTI_REAL, 2L, len) != 0) return NULL;
    g_mkl.set(d, DFTI_PLACEMENT, DFTI_NOT_INPLACE);
    g_mkl.set(d, DFTI_CONJUGATE_EVEN_STORAGE, DFTI_COMPLEX_COMPLEX);
    g_mkl.set(d, DFTI_INPUT_STRIDES, backward ? sc : sr);
    g_mkl.set(d, DFTI_OUTPUT_STRIDES, backward ? sr : sc);
    g_mkl.set(d, DFTI_THREAD_LIMIT, 1L);
    if (g_mkl.commit(d) != 0) { g_mkl.free_(&d); return NULL; }
    return d;
}

extern "C" __attribute__((visibility("default"))) const char *ref_fft_provider(void)
{
    return mkl_available() ? "mkl-dfti(libtorch_cpu, 1 thread)" : "dft64(port_fft, double precision)";
}

/* ---- the four FFTW symbols -------------------------------------------------------------- */
extern "C" {

fftwf_plan fftwf_plan_dft_r2c_2d(int n0, int n1, float *in, fftwf_complex *out, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->kind = 0; p->n0 = n0; p->n1 = n1; p->in = in; p->out = out;
    if (mkl_available()) p->mkl = mkl_make(n0, n1, 0);
    return p;
}

fftwf_plan fftwf_plan_dft_c2r_2d(int n0, int n1, fftwf_complex *in, float *out, unsigned flags)
{
    (void)flags;
    fftwf_plan p = (fftwf_plan)calloc(1, sizeof(*p));
    p->kind = 1; p->n0 = n0; p->n1 = n1; p->in = in; p->out = out;
    if (mkl_available()) p->mkl = mkl_make(n0, n1, 1);
    return p;
}

void fftwf_execute(const fftwf_plan p)
{
    if (p->mkl) {
        if (p->kind == 0) g_mkl.fwd(p->mkl, p->in, p->out);
        else              g_mkl.bwd(p->mkl, p->in, p->out);
        return;
    }
    if (p->kind == 0) pf_r2c_2d(p->n0, p->n1, (const float *)p->in, (float *)p->out);
    else              pf_c2r_2d(p->n0, p->n1, (const float *)p->in, (float *)p->out);
}

void fftwf_destroy_plan(fftwf_plan p)
{
    if (!p) return;
    if (p->mkl) g_mkl.free_(&p->mkl);
    free(p);
}

} /* extern "C" */
