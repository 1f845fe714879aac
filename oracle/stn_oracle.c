/*
 * CPU oracle for the LoANs STN crop path -- plain C restatement.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may build, load or call this
 * file.  The product library (loans_b200/csrc) never links it and has no CPU fallback.
 *
 * It follows oracle/stn_numpy.py operation by operation (same float32 roundings, same float64
 * promotion of the bilinear weights, same scatter order), so that full-size parity runs finish in
 * seconds.  tests/test_oracle.py checks it bit-for-bit against the numpy restatement.
 *
 * Reference being restated (paths relative to /root/reference):
 *   rotation dropout  functions/rotation_droput.py:26-48                     (in tree; pinned by golden vectors)
 *   affine grid       chainer.functions.spatial_transformer_grid, called at sheep/sheep_localizer.py:62,170
 *   bilinear sampler  chainer.functions.spatial_transformer_sampler, called at sheep/sheep_localizer.py:63,171
 * The last two live in chainer==4.1.0 (requirements.txt:1), absent from /root/reference and not
 * installable here: PARITY UNPINNED for them (no reference test/golden vector exists); see the header
 * of oracle/stn_numpy.py for what they are cross-checked against instead.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, every product and sum
 * below rounds separately exactly where numpy's does).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* numpy.linspace(-1, 1, n, dtype=float32): evaluated in float64 as arange(n)*step + start with
 * step = 2/(n-1), last element forced to stop, then cast. */
static void linspace_pm1(int n, float *out)
{
    if (n == 1) { out[0] = -1.0f; return; }
    double step = 2.0 / (double)(n - 1);
    for (int k = 0; k < n; ++k) out[k] = (float)((double)k * step + -1.0);
    out[n - 1] = 1.0f;
}

/* One grid element.  The K=3 contraction is evaluated as  rn(t2 + fma(t0, xs, rn(t1*ys))) : this is
 * the order OpenBLAS 0.3.30's sgemm produces for numpy's theta.dot(coords) (verified bit-for-bit in
 * tests/test_oracle.py); any other order is within one float32 ulp of it.  With t1 == 0 (LoANs always
 * zeroes the rotation terms, sheep/sheep_localizer.py:61) it is order-independent. */
static inline float grid_elem(float t0, float t1, float t2, float xs, float ys)
{
    float e = t1 * ys;
    e = fmaf(t0, xs, e);
    return t2 + e;
}

void oracle_rotation_dropout(const float *theta_in, float mask_value, float *theta_out, int n)
{
    /* functions/rotation_droput.py:33-36 / :39-45 and :48 -- y = x * mask, mask = 1 except [.,0,1],[.,1,0] */
    for (int b = 0; b < n; ++b) {
        const float *s = theta_in + 6 * b;
        float *d = theta_out + 6 * b;
        d[0] = s[0] * 1.0f; d[1] = s[1] * mask_value; d[2] = s[2] * 1.0f;
        d[3] = s[3] * mask_value; d[4] = s[4] * 1.0f; d[5] = s[5] * 1.0f;
    }
}

void oracle_grid_forward(const float *theta, float *grid, int n, int oh, int ow)
{
    float *xs = (float *)malloc(sizeof(float) * (size_t)(ow + oh));
    float *ys = xs + ow;
    linspace_pm1(ow, xs);
    linspace_pm1(oh, ys);
    for (int b = 0; b < n; ++b) {
        const float *t = theta + 6 * b;
        float *g0 = grid + (size_t)b * 2 * oh * ow;
        float *g1 = g0 + (size_t)oh * ow;
        for (int i = 0; i < oh; ++i)
            for (int j = 0; j < ow; ++j) {
                g0[i * ow + j] = grid_elem(t[0], t[1], t[2], xs[j], ys[i]);
                g1[i * ow + j] = grid_elem(t[3], t[4], t[5], xs[j], ys[i]);
            }
    }
    free(xs);
}

/* gtheta = ggrid . coords^T.  numpy runs this through sgemm with K = oH*oW, whose summation order is
 * BLAS-internal; the oracle accumulates in float64 and rounds once (the centre of the tolerance band). */
void oracle_grid_backward(const float *ggrid, float *gtheta, int n, int oh, int ow)
{
    float *xs = (float *)malloc(sizeof(float) * (size_t)(ow + oh));
    float *ys = xs + ow;
    linspace_pm1(ow, xs);
    linspace_pm1(oh, ys);
    for (int b = 0; b < n; ++b)
        for (int r = 0; r < 2; ++r) {
            const float *g = ggrid + ((size_t)b * 2 + r) * oh * ow;
            double sx = 0, sy = 0, s1 = 0;
            for (int i = 0; i < oh; ++i)
                for (int j = 0; j < ow; ++j) {
                    double v = g[i * ow + j];
                    sx += v * xs[j]; sy += v * ys[i]; s1 += v;
                }
            gtheta[b * 6 + r * 3 + 0] = (float)sx;
            gtheta[b * 6 + r * 3 + 1] = (float)sy;
            gtheta[b * 6 + r * 3 + 2] = (float)s1;
        }
    free(xs);
}

typedef struct {
    float u, v;       /* unclipped padded coordinates */
    float uc, vc;     /* clipped */
    int u0, v0;       /* top-left tap, padded index space */
} tap_t;

static inline tap_t make_tap(float gu, float gv, int h, int w)
{
    tap_t t;
    /* u = (u + 1) * (W - 1) / 2 + 1  -- four float32 roundings */
    float a = gu + 1.0f; a = a * (float)(w - 1); a = a / 2.0f; t.u = a + 1.0f;
    float b = gv + 1.0f; b = b * (float)(h - 1); b = b / 2.0f; t.v = b + 1.0f;
    t.uc = fminf(fmaxf(t.u, 0.0f), (float)(w + 1));
    t.vc = fminf(fmaxf(t.v, 0.0f), (float)(h + 1));
    int u0 = (int)floorf(t.uc); if (u0 < 0) u0 = 0; if (u0 > w) u0 = w;
    int v0 = (int)floorf(t.vc); if (v0 < 0) v0 = 0; if (v0 > h) v0 = h;
    t.u0 = u0; t.v0 = v0;
    return t;
}

/* x_pad[., v, u] with the one-pixel zero frame, without materialising the padded copy */
static inline float xpad(const float *xc, int h, int w, int v, int u)
{
    return (v >= 1 && v <= h && u >= 1 && u <= w) ? xc[(size_t)(v - 1) * w + (u - 1)] : 0.0f;
}

/* crop n samples frame n / k  (k = crops per frame; Chainer itself only has k == 1) */
void oracle_sampler_forward(const float *x, const float *grid, float *y,
                            int n, int k, int c, int h, int w, int oh, int ow)
{
    const size_t np_ = (size_t)oh * ow;
    for (int b = 0; b < n; ++b) {
        const float *xb = x + (size_t)(b / k) * c * h * w;
        const float *g0 = grid + (size_t)b * 2 * np_;
        const float *g1 = g0 + np_;
        for (size_t p = 0; p < np_; ++p) {
            tap_t t = make_tap(g0[p], g1[p], h, w);
            int u1 = t.u0 + 1, v1 = t.v0 + 1;
            double du1 = (double)u1 - (double)t.uc, du0 = (double)t.uc - (double)t.u0;
            double dv1 = (double)v1 - (double)t.vc, dv0 = (double)t.vc - (double)t.v0;
            float w1 = (float)(du1 * dv1), w2 = (float)(du0 * dv1);
            float w3 = (float)(du1 * dv0), w4 = (float)(du0 * dv0);
            for (int ch = 0; ch < c; ++ch) {
                const float *xc = xb + (size_t)ch * h * w;
                float acc = w1 * xpad(xc, h, w, t.v0, t.u0);
                acc += w2 * xpad(xc, h, w, t.v0, u1);
                acc += w3 * xpad(xc, h, w, v1, t.u0);
                acc += w4 * xpad(xc, h, w, v1, u1);
                y[((size_t)b * c + ch) * np_ + p] = acc;
            }
        }
    }
}

/* gx (n/k, c, h, w) and ggrid (n, 2, oh, ow); either may be NULL. */
void oracle_sampler_backward(const float *x, const float *grid, const float *gy,
                             float *gx, float *ggrid,
                             int n, int k, int c, int h, int w, int oh, int ow)
{
    const size_t np_ = (size_t)oh * ow;
    const size_t pad_sz = (size_t)c * (h + 2) * (w + 2);
    float *gpad = gx ? (float *)malloc(sizeof(float) * pad_sz) : NULL;
    if (gx) memset(gx, 0, sizeof(float) * (size_t)(n / k) * c * h * w);
    for (int b = 0; b < n; ++b) {
        const float *xb = x + (size_t)(b / k) * c * h * w;
        const float *g0 = grid + (size_t)b * 2 * np_;
        const float *g1 = g0 + np_;
        const float *gyb = gy + (size_t)b * c * np_;
        if (ggrid) {
            for (size_t p = 0; p < np_; ++p) {
                tap_t t = make_tap(g0[p], g1[p], h, w);
                int u1 = t.u0 + 1, v1 = t.v0 + 1;
                float wu0 = (float)((double)t.uc - (double)t.u0), wu1 = (float)((double)u1 - (double)t.uc);
                float wv0 = (float)((double)t.vc - (double)t.v0), wv1 = (float)((double)v1 - (double)t.vc);
                float su = 0.0f, sv = 0.0f;
                for (int ch = 0; ch < c; ++ch) {
                    const float *xc = xb + (size_t)ch * h * w;
                    float x1 = xpad(xc, h, w, t.v0, t.u0), x2 = xpad(xc, h, w, t.v0, u1);
                    float x3 = xpad(xc, h, w, v1, t.u0), x4 = xpad(xc, h, w, v1, u1);
                    float gu = -wv1 * x1; gu += wv1 * x2; gu -= wv0 * x3; gu += wv0 * x4;
                    float gv = -wu1 * x1; gv -= wu0 * x2; gv += wu1 * x3; gv += wu0 * x4;
                    float g = gyb[(size_t)ch * np_ + p];
                    gu = gu * g; gv = gv * g;
                    if (ch == 0) { su = gu; sv = gv; } else { su += gu; sv += gv; }
                }
                su = su / 2.0f * (float)(w - 1);
                sv = sv / 2.0f * (float)(h - 1);
                su = su * (float)(t.u > 0.0f) * (float)(t.u < (float)(w + 1));
                sv = sv * (float)(t.v > 0.0f) * (float)(t.v < (float)(h + 1));
                ggrid[(size_t)b * 2 * np_ + p] = su;
                ggrid[(size_t)b * 2 * np_ + np_ + p] = sv;
            }
        }
        if (gx) {
            /* numpy.add.at, four passes in the reference's order, into a zeroed padded buffer */
            memset(gpad, 0, sizeof(float) * pad_sz);
            const int ws = w + 2;
            for (int pass = 0; pass < 4; ++pass)
                for (int ch = 0; ch < c; ++ch) {
                    float *gc = gpad + (size_t)ch * (h + 2) * ws;
                    for (size_t p = 0; p < np_; ++p) {
                        tap_t t = make_tap(g0[p], g1[p], h, w);
                        int u1 = t.u0 + 1, v1 = t.v0 + 1;
                        float wu0 = (float)((double)t.uc - (double)t.u0), wu1 = (float)((double)u1 - (double)t.uc);
                        float wv0 = (float)((double)t.vc - (double)t.v0), wv1 = (float)((double)v1 - (double)t.vc);
                        float g = gyb[(size_t)ch * np_ + p];
                        float val; int vv, uu;
                        switch (pass) {
                        case 0:  val = g * wu1 * wv1; vv = t.v0; uu = t.u0; break;
                        case 1:  val = g * wu0 * wv1; vv = t.v0; uu = u1;   break;
                        case 2:  val = g * wu1 * wv0; vv = v1;   uu = t.u0; break;
                        default: val = g * wu0 * wv0; vv = v1;   uu = u1;   break;
                        }
                        gc[(size_t)vv * ws + uu] += val;
                    }
                }
            float *gxb = gx + (size_t)(b / k) * c * h * w;
            for (int ch = 0; ch < c; ++ch)
                for (int r = 0; r < h; ++r)
                    for (int s = 0; s < w; ++s)
                        gxb[((size_t)ch * h + r) * w + s] += gpad[((size_t)ch * (h + 2) + r + 1) * ws + s + 1];
        }
    }
    free(gpad);
}

/* The composite of sheep/sheep_localizer.py:61-63: rotation_dropout -> grid -> sampler. */
void oracle_crop_forward(const float *x, const float *theta, float mask_value, float *y, float *grid,
                         int n, int k, int c, int h, int w, int oh, int ow)
{
    float *th = (float *)malloc(sizeof(float) * 6 * (size_t)n);
    float *g = grid ? grid : (float *)malloc(sizeof(float) * (size_t)n * 2 * oh * ow);
    oracle_rotation_dropout(theta, mask_value, th, n);
    oracle_grid_forward(th, g, n, oh, ow);
    oracle_sampler_forward(x, g, y, n, k, c, h, w, oh, ow);
    if (!grid) free(g);
    free(th);
}

/* gtheta (n,2,3) always; gx (n/k,c,h,w), ggrid_out (n,2,oh,ow), ggrid_upstream optional (NULL). */
void oracle_crop_backward(const float *x, const float *theta, float mask_value, const float *gy,
                          const float *ggrid_upstream, float *gtheta, float *gx, float *ggrid_out,
                          int n, int k, int c, int h, int w, int oh, int ow)
{
    const size_t gsz = (size_t)n * 2 * oh * ow;
    float *th = (float *)malloc(sizeof(float) * 6 * (size_t)n);
    float *g = (float *)malloc(sizeof(float) * gsz);
    float *gg = (float *)malloc(sizeof(float) * gsz);
    oracle_rotation_dropout(theta, mask_value, th, n);
    oracle_grid_forward(th, g, n, oh, ow);
    oracle_sampler_backward(x, g, gy, gx, gg, n, k, c, h, w, oh, ow);
    if (ggrid_out) memcpy(ggrid_out, gg, sizeof(float) * gsz);
    if (ggrid_upstream)
        for (size_t i = 0; i < gsz; ++i) gg[i] = gg[i] + ggrid_upstream[i];
    oracle_grid_backward(gg, gtheta, n, oh, ow);
    oracle_rotation_dropout(gtheta, mask_value, gtheta, n);
    free(gg); free(g); free(th);
}
