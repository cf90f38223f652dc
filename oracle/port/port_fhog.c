/*
 * oracle/port/port_fhog.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the fHOG feature extractor as the reference drives it
 * (libhog/fhog.h:16-38: gradMag(I,M,O,h,w,d=1,full=true) then
 *  fhog(M,O,H,h,w,bin=4,nOrients=9,softBin=-1,clip=0.2f)), written as plain scalar loops:
 *   port_gradmag    <- grad1 / gradMag / acosTable   libhog/gradientMex.cpp:15-37, 59-100, 47-56
 *   port_gradhist18 <- gradQuantize / gradHist       libhog/gradientMex.cpp:112-145, 148-231
 *   port_hognorm    <- hogNormMatrix                 libhog/gradientMex.cpp:236-253
 *   port_fhog       <- hogChannels / fhog            libhog/gradientMex.cpp:256-280, 298-317
 * All arrays are column-major (row index y fastest), channels outermost.
 *
 * The two SSE approximations the reference relies on (RCPSQRT/RCP = _mm_rsqrt_ps/_mm_rcp_ps,
 * gradientMex.cpp:83-84, sse.hpp:40-41) are 12-bit table instructions whose error (3e-4) exceeds
 * the 1e-4 parity tolerance, so they are issued here as the very same instructions
 * (scalar forms rsqrtss/rcpss) rather than replaced by exact math.
 * No multiply-add may be contracted: build with -ffp-contract=off (oracle/Makefile).
 */
#include <emmintrin.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PORT_PI 3.14159265f                  /* gradientMex.cpp:12 */

static float sse_rsqrt(float x) { return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x))); }
static float sse_rcp(float x)   { return _mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x))); }

/* the two instructions on arrays, for the tests that pin the product's harvested tables against the CPU they run on */
__attribute__((visibility("default"))) void port_sse_rsqrt(const float *x, float *y, long n) { for (long i = 0; i < n; ++i) y[i] = sse_rsqrt(x[i]); }
__attribute__((visibility("default"))) void port_sse_rcp(const float *x, float *y, long n) { for (long i = 0; i < n; ++i) y[i] = sse_rcp(x[i]); }

/* gradientMex.cpp:47-56.  a1[i] ~ acos(i/10000) for i in [-10010, 10010); the argument is a float and the
 * reference is C++, so the float overload of acos is the one called. */
static const float *acos_table(void)
{
    enum { n = 10000, b = 10 };
    static float a[n * 2 + b * 2];
    static int init = 0;
    float *a1 = a + n + b;
    if (init) return a1;
    for (int i = -n - b; i < -n; ++i) a1[i] = PORT_PI;
    for (int i = -n; i < n; ++i) a1[i] = acosf((float)i / (float)n);
    for (int i = n; i < n + b; ++i) a1[i] = 0.0f;
    for (int i = -n - b; i < n / 10; ++i) if (a1[i] > PORT_PI - 1e-6f) a1[i] = PORT_PI - 1e-6f;
    init = 1;
    return a1;
}

/* expose the table so the tests can pin the product's host-harvested copy against it */
__attribute__((visibility("default")))
void port_acos_table(float *out /* 20020 */) { memcpy(out, acos_table() - 10010, sizeof(float) * 20020); }

/* gradientMex.cpp:59-100 with d=1, full=true.  I, M, O: h x w column-major. */
__attribute__((visibility("default")))
void port_gradmag(const float *I, float *M, float *O, int h, int w)
{
    const float *acost = acos_table();
    for (int x = 0; x < w; ++x) {
        const float *Ic = I + (long)x * h;
        /* grad1 (:15-37): central differences, one-sided (x1) at the borders */
        const float *Ip = Ic - h, *In = Ic + h; float r = .5f;
        if (x == 0) { r = 1.f; Ip += h; } else if (x == w - 1) { r = 1.f; In -= h; }
        for (int y = 0; y < h; ++y) {
            float gx = (In[y] - Ip[y]) * r;
            float gy;
            if (y == 0)          gy = (Ic[1] - Ic[0]) * 1.f;
            else if (y == h - 1) gy = (Ic[h - 1] - Ic[h - 2]) * 1.f;
            else                 gy = (Ic[y + 1] - Ic[y - 1]) * .5f;
            float a = gx * gx, b = gy * gy;
            float m2 = a + b;                                         /* :74 */
            float m = sse_rsqrt(m2); if (!(m < 1e10f)) m = 1e10f;      /* :83 MIN(RCPSQRT(M2),1e10f): minps returns the 2nd operand unless 1st < 2nd */
            M[(long)x * h + y] = sse_rcp(m);                          /* :84 */
            float g = (gx * m) * 10000.0f;                            /* :85 */
            union { float f; unsigned u; } ug, uy; ug.f = g; uy.f = gy;
            ug.u ^= (uy.u & 0x80000000u);                             /* :86 flip by the sign BIT of Gy (incl. -0.0) */
            float o = acost[(int)ug.f];                               /* :90 */
            if (gy < 0) o += PORT_PI;                                 /* :91-97 */
            O[(long)x * h + y] = o;
        }
    }
}

/* gradientMex.cpp:148-231 on the branch fhog takes: nOrients=18, full=true, softBin=-1, bin=4:
 * nearest orientation bin (gradQuantize non-interpolating, :130-131 / :143-144), bilinear spatial
 * interpolation (the "trilinear" branch with softBin<0, :183-221), boundary cells x 8/7 (:225-230).
 * R1: 18 x wb x hb (hb fastest). */
__attribute__((visibility("default")))
void port_gradhist18(const float *M, const float *O, float *R1, int h, int w)
{
    const int bin = 4, nOrients = 18;
    const int hb = h / bin, wb = w / bin, h0 = hb * bin, w0 = wb * bin, nb = wb * hb;
    const float s = (float)bin, sInv = 1 / s, sInv2 = 1 / s / s;
    const float oMult = (float)nOrients / (2 * PORT_PI);
    const int oMax = nOrients * nb;
    memset(R1, 0, sizeof(float) * (size_t)nb * nOrients);
    float xb = 0, init = 0;
    for (int x = 0; x < w0; ++x) {
        const float *Oc = O + (long)x * h, *Mc = M + (long)x * h;
        if (x == 0) { init = (0 + .5f) * sInv - 0.5f; xb = init; }
        int hasLf = xb >= 0, xb0 = hasLf ? (int)xb : -1, hasRt = xb0 < wb - 1;
        float xd = xb - xb0; xb += sInv;
        float yb = init;
        for (int y = 0; y < h0; ++y) {
            float o = Oc[y] * oMult; int o0 = (int)(o + .5f);
            o0 *= nb; if (o0 >= oMax) o0 = 0;
            float m0 = Mc[y] * sInv2;
            int yb0 = (y < bin / 2) ? -1 : (int)yb;
            float yd = yb - yb0; yb += sInv;
            float xyd = xd * yd;
            float ms0 = 1 - xd - yd + xyd, ms1 = yd - xyd, ms2 = xd - xyd, ms3 = xyd;
            float *H0 = R1 + o0 + (long)xb0 * hb + yb0;
            int hasTop = yb0 >= 0, hasBot = yb0 < hb - 1;
            if (hasLf) { if (hasTop) H0[0] += ms0 * m0;  if (hasBot) H0[1] += ms1 * m0; }
            if (hasRt) { if (hasTop) H0[hb] += ms2 * m0; if (hasBot) H0[hb + 1] += ms3 * m0; }
        }
    }
    for (int o = 0; o < nOrients; ++o) {
        int x, y;
        x = 0;      for (y = 0; y < hb; ++y) R1[o * nb + x * hb + y] *= 8.f / 7.f;
        y = 0;      for (x = 0; x < wb; ++x) R1[o * nb + x * hb + y] *= 8.f / 7.f;
        x = wb - 1; for (y = 0; y < hb; ++y) R1[o * nb + x * hb + y] *= 8.f / 7.f;
        y = hb - 1; for (x = 0; x < wb; ++x) R1[o * nb + x * hb + y] *= 8.f / 7.f;
    }
}

/* gradientMex.cpp:236-253.  R2: 9 x wb x hb; N: (wb+1) x (hb+1), (hb+1) fastest. */
__attribute__((visibility("default")))
void port_hognorm(const float *R2, float *N, int hb, int wb)
{
    const int hb1 = hb + 1, wb1 = wb + 1, bin = 4, nOrients = 9;
    const float eps = 1e-4f / 4 / bin / bin / bin / bin;
    memset(N, 0, sizeof(float) * (size_t)hb1 * wb1);
    float *N1 = N + hb1 + 1;
    for (int o = 0; o < nOrients; ++o) for (int x = 0; x < wb; ++x) for (int y = 0; y < hb; ++y) {
        float v = R2[(long)o * wb * hb + x * hb + y];
        float sq = v * v;
        N1[x * hb1 + y] += sq;
    }
    for (int x = 0; x < wb - 1; ++x) for (int y = 0; y < hb - 1; ++y) {
        float *n = N1 + x * hb1 + y;
        float e = n[0] + n[1]; e = e + n[hb1]; e = e + n[hb1 + 1]; e = e + eps;
        *n = 1 / sqrtf(e);
    }
    int x, y, dx, dy;
    x = 0;       dx = 1;  dy = 1;  y = 0;                    N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    x = 0;       dx = 1;  dy = 0;  for (y = 0; y < hb1; ++y) N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    x = 0;       dx = 1;  dy = -1; y = hb1 - 1;              N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    x = wb1 - 1; dx = -1; dy = 1;  y = 0;                    N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    x = wb1 - 1; dx = -1; dy = 0;  for (y = 0; y < hb1; ++y) N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    x = wb1 - 1; dx = -1; dy = -1; y = hb1 - 1;              N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    y = 0;       dx = 0;  dy = 1;  for (x = 0; x < wb1; ++x) N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
    y = hb1 - 1; dx = 0;  dy = -1; for (x = 0; x < wb1; ++x) N[x * hb1 + y] = N[(x + dx) * hb1 + y + dy];
}

/* gradientMex.cpp:256-280 (types 1 and 2) + :298-317.  H: 32 x wb x hb, channel 31 left at zero
 * (the driver memsets n_chns = 32 channels, libhog/fhog.h:27-31). */
__attribute__((visibility("default")))
void port_fhog_from_mo(const float *M, const float *O, float *H, int h, int w)
{
    const int bin = 4, hb = h / bin, wb = w / bin, nb = hb * wb, hb1 = hb + 1;
    const float clip = 0.2f, rtex = .2357f;
    float *R1 = (float *)malloc(sizeof(float) * (size_t)nb * 18);
    float *R2 = (float *)malloc(sizeof(float) * (size_t)nb * 9);
    float *N = (float *)malloc(sizeof(float) * (size_t)(hb + 1) * (wb + 1));
    memset(H, 0, sizeof(float) * (size_t)nb * 32);
    port_gradhist18(M, O, R1, h, w);
    for (int o = 0; o < 9; ++o) for (int i = 0; i < nb; ++i) R2[o * nb + i] = R1[o * nb + i] + R1[(o + 9) * nb + i];
    port_hognorm(R2, N, hb, wb);
    const int blk[4] = { 0, 1, hb1, hb1 + 1 };
    for (int pass = 0; pass < 3; ++pass) {
        const float *R = (pass == 1) ? R2 : R1;
        int nOr = (pass == 1) ? 9 : 18;
        float *Hb = H + (long)nb * 9 * (pass == 0 ? 0 : (pass == 1 ? 2 : 3));
        for (int o = 0; o < nOr; ++o) for (int x = 0; x < wb; ++x) {
            const float *Rc = R + (long)o * nb + x * hb, *N1 = N + x * hb1 + hb1 + 1;
            for (int y = 0; y < hb; ++y) for (int c = 0; c < 4; ++c) {
                float t = Rc[y] * N1[y - blk[c]]; if (t > clip) t = clip;
                if (pass < 2) Hb[(long)o * nb + x * hb + y] += t * .5f;
                else          Hb[(long)c * nb + x * hb + y] += t * rtex;
            }
        }
    }
    free(R1); free(R2); free(N);
}

/* libhog/fhog.h:16-38 */
__attribute__((visibility("default")))
void port_fhog_extract(const float *I, int h, int w, float *H)
{
    float *M = (float *)malloc(sizeof(float) * (size_t)h * w * 2), *O = M + (long)h * w;
    port_gradmag(I, M, O, h, w);
    port_fhog_from_mo(M, O, H, h, w);
    free(M);
}
