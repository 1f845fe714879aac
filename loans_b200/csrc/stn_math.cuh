// stn_math.cuh -- per-pixel arithmetic of the STN crop path, shared by every kernel in this library.
//
// Everything here is __host__ __device__ so that tests/hostemu can run the very same per-pixel code on
// the CPU against the oracle (a test harness, never a product path).  On the device the float32
// operations whose rounding the reference's numpy path fixes are written with the _rn intrinsics so
// that ptxas cannot contract them into FMAs; on the host the harness is compiled with
// -ffp-contract=off for the same effect.
//
// Arithmetic being reproduced (see oracle/stn_numpy.py for the statement-by-statement restatement):
//   grid    : chainer 4.1.0 SpatialTransformerGrid._forward      (call site sheep/sheep_localizer.py:62)
//   sampler : chainer 4.1.0 SpatialTransformerSampler._forward/_backward (call site sheep/sheep_localizer.py:63)
//   mask    : functions/rotation_droput.py:33-36,39-45
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define STN_HD __host__ __device__ __forceinline__
#else
#define STN_HD inline
#endif

namespace stn {

#if defined(__CUDA_ARCH__)
STN_HD float f_mul(float a, float b) { return __fmul_rn(a, b); }
STN_HD float f_add(float a, float b) { return __fadd_rn(a, b); }
STN_HD float f_sub(float a, float b) { return __fsub_rn(a, b); }
STN_HD float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
STN_HD int f_floor_i(float a) { return __float2int_rd(a); }
STN_HD int f_ceil_i(float a) { return __float2int_ru(a); }
STN_HD double d_mul(double a, double b) { return __dmul_rn(a, b); }
STN_HD double d_add(double a, double b) { return __dadd_rn(a, b); }
#define STN_LDG(p) __ldg(p)
#else
STN_HD float f_mul(float a, float b) { return a * b; }
STN_HD float f_add(float a, float b) { return a + b; }
STN_HD float f_sub(float a, float b) { return a - b; }
STN_HD float f_fma(float a, float b, float c) { return fmaf(a, b, c); }
STN_HD int f_floor_i(float a) { return (int)floorf(a); }
STN_HD int f_ceil_i(float a) { return (int)ceilf(a); }
STN_HD double d_mul(double a, double b) { return a * b; }
STN_HD double d_add(double a, double b) { return a + b; }
#define STN_LDG(p) (*(p))
#endif

// numpy.linspace(-1, 1, n, dtype=float32)[k]: float64 arange(n)*step + start, endpoint forced, then cast.
// step = 2/(n-1) is computed once on the host (IEEE double division) and passed in.
STN_HD float linspace_pm1(int k, int n, double step)
{
    if (n == 1) return -1.0f;
    if (k == n - 1) return 1.0f;
    return (float)d_add(d_mul((double)k, step), -1.0);
}

struct Theta {
    float t00, t01, t02, t10, t11, t12;
};

// rotation dropout fused in front of the grid: theta * mask, mask = 1 except [0,1] and [1,0]
STN_HD Theta load_theta_masked(const float *th, float mask01)
{
    Theta t;
    t.t00 = STN_LDG(th + 0);
    t.t01 = f_mul(STN_LDG(th + 1), mask01);
    t.t02 = STN_LDG(th + 2);
    t.t10 = f_mul(STN_LDG(th + 3), mask01);
    t.t11 = STN_LDG(th + 4);
    t.t12 = STN_LDG(th + 5);
    return t;
}

// One element of theta . [xs; ys; 1].  Evaluation order rn(t2 + fma(t0, xs, rn(t1*ys))) -- the order the
// oracle fixes (oracle/stn_oracle.c:grid_elem); order-independent when t1 == 0 (LoANs' ratio=0.0 case).
STN_HD float grid_elem(float t0, float t1, float t2, float xs, float ys)
{
    float e = f_mul(t1, ys);
    e = f_fma(t0, xs, e);
    return f_add(t2, e);
}

// [-1,1] -> padded pixel coordinate:  (g + 1) * (size - 1) / 2 + 1, four float32 roundings (/2 is exact).
STN_HD float to_padded_px(float g, float size_m1)
{
    float a = f_add(g, 1.0f);
    a = f_mul(a, size_m1);
    a = f_mul(a, 0.5f);
    return f_add(a, 1.0f);
}

struct Tap {
    float u, v;      // unclipped padded coordinates (gradient mask uses these)
    float wu0, wu1;  // uc - u0, u1 - uc   (exact in float32 wherever they meet a non-padding tap)
    float wv0, wv1;
    int u0, v0;      // top-left tap in padded index space, u0 in [0,W], v0 in [0,H]
};

STN_HD Tap make_tap(float g0, float g1, int H, int W)
{
    Tap t;
    t.u = to_padded_px(g0, (float)(W - 1));
    t.v = to_padded_px(g1, (float)(H - 1));
    float uc = fminf(fmaxf(t.u, 0.0f), (float)(W + 1));
    float vc = fminf(fmaxf(t.v, 0.0f), (float)(H + 1));
    int u0 = f_floor_i(uc);
    u0 = u0 < 0 ? 0 : (u0 > W ? W : u0);
    int v0 = f_floor_i(vc);
    v0 = v0 < 0 ? 0 : (v0 > H ? H : v0);
    t.u0 = u0;
    t.v0 = v0;
    t.wu0 = f_sub(uc, (float)u0);
    t.wu1 = f_sub((float)(u0 + 1), uc);
    t.wv0 = f_sub(vc, (float)v0);
    t.wv1 = f_sub((float)(v0 + 1), vc);
    return t;
}

// Offsets of the four taps inside one H x W plane and whether each is a real pixel (not the zero frame).
struct TapAddr {
    int o00;               // offset of (v0,u0) in unpadded plane; others are +1, +W, +W+1
    bool c0, c1, r0, r1;   // column u0 / u1 valid, row v0 / v1 valid
};

STN_HD TapAddr make_tap_addr(const Tap &t, int H, int W)
{
    TapAddr a;
    a.c0 = t.u0 >= 1;
    a.c1 = t.u0 <= W - 1;
    a.r0 = t.v0 >= 1;
    a.r1 = t.v0 <= H - 1;
    a.o00 = (t.v0 - 1) * W + (t.u0 - 1);
    return a;
}

STN_HD void load_taps(const float *plane, const TapAddr &a, int W, float &x1, float &x2, float &x3, float &x4)
{
    // two address registers (top row, bottom row) shared by the tap pairs; the empty asm keeps the compiler from
    // re-deriving a 64-bit address under every load's own predicate
    const float *pt = plane + a.o00;
    const float *pb = pt + W;
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(pt), "+l"(pb));
#endif
    x1 = (a.r0 && a.c0) ? STN_LDG(pt) : 0.0f;
    x2 = (a.r0 && a.c1) ? STN_LDG(pt + 1) : 0.0f;
    x3 = (a.r1 && a.c0) ? STN_LDG(pb) : 0.0f;
    x4 = (a.r1 && a.c1) ? STN_LDG(pb + 1) : 0.0f;
}

struct Weights4 {
    float w1, w2, w3, w4;
};

STN_HD Weights4 make_weights(const Tap &t)
{
    Weights4 w;
    w.w1 = f_mul(t.wu1, t.wv1);
    w.w2 = f_mul(t.wu0, t.wv1);
    w.w3 = f_mul(t.wu1, t.wv0);
    w.w4 = f_mul(t.wu0, t.wv0);
    return w;
}

// y = w1*x1; y += w2*x2; y += w3*x3; y += w4*x4   (products and sums round separately)
STN_HD float interp(const Weights4 &w, float x1, float x2, float x3, float x4)
{
    float y = f_mul(w.w1, x1);
    y = f_add(y, f_mul(w.w2, x2));
    y = f_add(y, f_mul(w.w3, x3));
    y = f_add(y, f_mul(w.w4, x4));
    return y;
}

// per-channel coordinate gradients before the gy product (sampler _backward)
STN_HD void grad_uv(const Tap &t, float x1, float x2, float x3, float x4, float &gu, float &gv)
{
    gu = f_mul(-t.wv1, x1);
    gu = f_add(gu, f_mul(t.wv1, x2));
    gu = f_sub(gu, f_mul(t.wv0, x3));
    gu = f_add(gu, f_mul(t.wv0, x4));
    gv = f_mul(-t.wu1, x1);
    gv = f_sub(gv, f_mul(t.wu0, x2));
    gv = f_add(gv, f_mul(t.wu1, x3));
    gv = f_add(gv, f_mul(t.wu0, x4));
}

// gu/2*(W-1) masked where the UNclipped coordinate left the padded image (strict inequalities)
STN_HD void finish_grad_uv(const Tap &t, int H, int W, float &su, float &sv)
{
    su = f_mul(f_mul(su, 0.5f), (float)(W - 1));
    sv = f_mul(f_mul(sv, 0.5f), (float)(H - 1));
    if (!(t.u > 0.0f && t.u < (float)(W + 1))) su = f_mul(su, 0.0f);
    if (!(t.v > 0.0f && t.v < (float)(H + 1))) sv = f_mul(sv, 0.0f);
}

// ---------------------------------------------------------------------------------------------------
// Inverse mapping for the gather formulation of gx.
//
// For an affine theta the padded coordinates are affine in the crop indices:
//     u(i,j) ~ muj*j + mui*i + cu ,   v(i,j) ~ mvj*j + mvi*i + cv .
// Source pixel (pr, pc) (padded) receives gradient from exactly those crop pixels whose 2x2 tap window
// covers it, i.e. |u - pc| <= 1 and |v - pr| <= 1.  The candidates are enumerated scanline by scanline in
// (slightly widened) real arithmetic; each candidate is then re-evaluated with the exact forward chain
// and only counted if its integer taps really hit the pixel, so the widening costs time, never accuracy.
struct InvCrop {
    Theta th;
    float muj, mui, cu, mvj, mvi, cv;
    float i00, i01, i10, i11;   // (du,dv) -> (dj,di)
    float ext_j, ext_i;         // half extents of the candidate bounding box (1e30 when singular)
    float tol_u, tol_v;         // 1 + slack
    float ru, rv;               // 1/muj, 1/mvj or 0 when that direction does not restrict j
    int pad_;
};

STN_HD InvCrop make_inv_crop(const Theta &th, int H, int W, int oH, int oW)
{
    InvCrop c;
    c.th = th;
    const double sx = oW > 1 ? 2.0 / (double)(oW - 1) : 0.0;
    const double sy = oH > 1 ? 2.0 / (double)(oH - 1) : 0.0;
    const double hw = 0.5 * (double)(W - 1), hh = 0.5 * (double)(H - 1);
    const double muj = (double)th.t00 * sx * hw, mui = (double)th.t01 * sy * hw;
    const double cu = ((double)th.t02 - (double)th.t00 - (double)th.t01 + 1.0) * hw + 1.0;
    const double mvj = (double)th.t10 * sx * hh, mvi = (double)th.t11 * sy * hh;
    const double cv = ((double)th.t12 - (double)th.t10 - (double)th.t11 + 1.0) * hh + 1.0;
    const double nj = (double)(oW > 1 ? oW - 1 : 1), ni = (double)(oH > 1 ? oH - 1 : 1);
    const double tol_u = 1.0 + 1e-3 + 4e-6 * (fabs(muj) * nj + fabs(mui) * ni + fabs(cu) + W + 2.0);
    const double tol_v = 1.0 + 1e-3 + 4e-6 * (fabs(mvj) * nj + fabs(mvi) * ni + fabs(cv) + H + 2.0);
    const double det = muj * mvi - mui * mvj;
    const double scale = fabs(muj * mvi) + fabs(mui * mvj);
    double i00 = 0, i01 = 0, i10 = 0, i11 = 0, ext_j = 1e30, ext_i = 1e30;
    if (fabs(det) > 1e-9 * scale + 1e-30) {
        i00 = mvi / det; i01 = -mui / det; i10 = -mvj / det; i11 = muj / det;
        ext_j = fabs(i00) * tol_u + fabs(i01) * tol_v + 1e-2
              + 4e-6 * (fabs(i00) * (W + 2.0 + fabs(cu)) + fabs(i01) * (H + 2.0 + fabs(cv)));
        ext_i = fabs(i10) * tol_u + fabs(i11) * tol_v + 1e-2
              + 4e-6 * (fabs(i10) * (W + 2.0 + fabs(cu)) + fabs(i11) * (H + 2.0 + fabs(cv)));
        if (!(ext_j < 1e30)) ext_j = 1e30;
        if (!(ext_i < 1e30)) ext_i = 1e30;
    }
    c.muj = (float)muj; c.mui = (float)mui; c.cu = (float)cu;
    c.mvj = (float)mvj; c.mvi = (float)mvi; c.cv = (float)cv;
    c.i00 = (float)i00; c.i01 = (float)i01; c.i10 = (float)i10; c.i11 = (float)i11;
    c.ext_j = (float)ext_j; c.ext_i = (float)ext_i;
    c.tol_u = (float)tol_u; c.tol_v = (float)tol_v;
    // a direction restricts j only if the coordinate moves by more than 1e-3 px across the whole row
    c.ru = (fabs(muj) * nj > 1e-3) ? (float)(1.0 / muj) : 0.0f;
    c.rv = (fabs(mvj) * nj > 1e-3) ? (float)(1.0 / mvj) : 0.0f;
    c.pad_ = 0;
    return c;
}

// j-interval allowed by |a*j + b| <= tol; r = 1/a (0: no restriction beyond |b| <= tol + 1e-3)
STN_HD bool j_interval(float r, float b, float tol, float &lo, float &hi)
{
    if (r == 0.0f) {
        lo = -1.0f; hi = 3.0e9f;
        return fabsf(b) <= tol + 1e-3f;
    }
    float a0 = (-b - tol) * r, a1 = (-b + tol) * r;
    lo = fminf(a0, a1);
    hi = fmaxf(a0, a1);
    const float s = 2e-6f * (fabsf(lo) + fabsf(hi)) + 1e-4f;
    lo -= s;
    hi += s;
    return true;
}

// Gradient gathered by padded source pixel (pr, pc) from ONE crop, CG channels at a time.
//   gy_crop : this crop's gy, channel c0 at offset 0, channel stride = oH*oW elements
//   load_gy : functor (const GY*, index) -> float
template <int CG, typename GY, typename LoadGy>
STN_HD void gather_from_crop(const InvCrop &c, const float *xs, const float *ys, int H, int W, int oH, int oW,
                             int pr, int pc, const GY *gy_crop, int nc, LoadGy load_gy, float (&acc)[CG])
{
    const float du = (float)pc - c.cu, dv = (float)pr - c.cv;
    int i_lo = 0, i_hi = oH - 1;
    if (c.ext_i < 1e29f) {
        const float cj = c.i00 * du + c.i01 * dv;
        const float ci = c.i10 * du + c.i11 * dv;
        if (cj + c.ext_j < 0.0f || cj - c.ext_j > (float)(oW - 1)) return;
        const float fl = fmaxf(ci - c.ext_i, 0.0f), fh = fminf(ci + c.ext_i, (float)(oH - 1));
        if (fl > fh) return;
        i_lo = f_ceil_i(fl);
        i_hi = f_floor_i(fh);
    }
    const size_t plane = (size_t)oH * oW;
    for (int i = i_lo; i <= i_hi; ++i) {
        float lo_u, hi_u, lo_v, hi_v;
        const float fi = (float)i;
        if (!j_interval(c.ru, c.mui * fi - du, c.tol_u, lo_u, hi_u)) continue;
        if (!j_interval(c.rv, c.mvi * fi - dv, c.tol_v, lo_v, hi_v)) continue;
        const float fl = fmaxf(fmaxf(lo_u, lo_v), 0.0f), fh = fminf(fminf(hi_u, hi_v), (float)(oW - 1));
        if (fl > fh) continue;
        const int j_lo = f_ceil_i(fl), j_hi = f_floor_i(fh);
        const float ysi = ys[i];
        for (int j = j_lo; j <= j_hi; ++j) {
            const float xsj = xs[j];
            const Tap t = make_tap(grid_elem(c.th.t00, c.th.t01, c.th.t02, xsj, ysi),
                                   grid_elem(c.th.t10, c.th.t11, c.th.t12, xsj, ysi), H, W);
            float wu, wv;
            if (t.u0 == pc) wu = t.wu1; else if (t.u0 + 1 == pc) wu = t.wu0; else continue;
            if (t.v0 == pr) wv = t.wv1; else if (t.v0 + 1 == pr) wv = t.wv0; else continue;
            const size_t o = (size_t)i * oW + j;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (ch < nc) {
                    const float g = load_gy(gy_crop, (size_t)ch * plane + o);
                    acc[ch] = f_add(acc[ch], f_mul(f_mul(g, wu), wv));     // gy * wu * wv, reference order
                }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Geometry for the tile-scatter formulation of gx (the fast path; the gather above is its fallback).
//
// A CTA owns a tile of frame pixels in shared memory and adds into it the contributions of every crop pixel
// whose 2x2 tap window touches the tile.  Two crop pixels can only touch the same frame pixel if their
// coordinates differ by less than 2 in BOTH u and v; for an affine map that bounds their index distance by
//   |dj| < Ej = (|i00|+|i01|)*T ,  |di| < Ei = (|i10|+|i11|)*T ,  T = 2 + margin.
// Crop pixels are therefore processed in P*Q phases, (i mod P, j mod Q) = const per phase with P >= Ei,
// Q >= Ej: inside one phase no two pixels share a frame pixel, so plain read-modify-writes are race free and
// the summation order is fixed (deterministic).  Down-sampling by >= 2 gives P = Q = 1: a single phase.
// Everything here is float32 and only needs to be CONSERVATIVE (it selects candidates and phases); every
// candidate is re-evaluated with the exact forward chain before anything is added.
struct ScatterGeom {
    Theta th;
    float muj, mui, cu, mvj, mvi, cv;   // approximate affine map (padded px) for the cheap pre-test
    float i00, i01, i10, i11;           // its inverse
    float slack;                        // pre-test slack in px
    int P, Q;                           // phase periods; P == 0 -> use the gather fallback for this crop
    int r_min, r_max, s_min, s_max;     // unpadded frame rows / columns any tap of this crop can touch (conservative)
};

constexpr int kMaxScatterPhases = 64;

STN_HD ScatterGeom make_scatter_geom(const Theta &th, int H, int W, int oH, int oW)
{
    ScatterGeom g;
    g.th = th;
    const float sx = oW > 1 ? 2.0f / (float)(oW - 1) : 0.0f;
    const float sy = oH > 1 ? 2.0f / (float)(oH - 1) : 0.0f;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    g.muj = th.t00 * sx * hw; g.mui = th.t01 * sy * hw; g.cu = (th.t02 - th.t00 - th.t01 + 1.0f) * hw + 1.0f;
    g.mvj = th.t10 * sx * hh; g.mvi = th.t11 * sy * hh; g.cv = (th.t12 - th.t10 - th.t11 + 1.0f) * hh + 1.0f;
    const float nj = (float)(oW > 1 ? oW - 1 : 1), ni = (float)(oH > 1 ? oH - 1 : 1);
    const float mag = fabsf(g.muj) * nj + fabsf(g.mui) * ni + fabsf(g.cu) + fabsf(g.mvj) * nj + fabsf(g.mvi) * ni
                    + fabsf(g.cv) + (float)(W + H);
    g.slack = 0.05f + 1e-5f * mag;
    const float det = g.muj * g.mvi - g.mui * g.mvj;
    const float scale = fabsf(g.muj * g.mvi) + fabsf(g.mui * g.mvj);
    {   // bounding box of the four crop corners in padded pixel coordinates, widened by 2 px + slack
        const float au = fabsf(g.muj) * nj, bu = fabsf(g.mui) * ni, av = fabsf(g.mvj) * nj, bv = fabsf(g.mvi) * ni;
        const float u_lo = g.cu + fminf(g.muj * nj, 0.0f) + fminf(g.mui * ni, 0.0f), v_lo = g.cv + fminf(g.mvj * nj, 0.0f) + fminf(g.mvi * ni, 0.0f);
        const float m = 2.0f + g.slack;
        const float fs0 = fmaxf(u_lo - m - 1.0f, -1.0f), fs1 = fminf(u_lo + au + bu + m, 2.0e9f);
        const float fr0 = fmaxf(v_lo - m - 1.0f, -1.0f), fr1 = fminf(v_lo + av + bv + m, 2.0e9f);
        const bool ok = mag < 1e8f;                      // otherwise: no restriction
        g.s_min = ok ? f_floor_i(fs0) : -1; g.s_max = ok ? f_ceil_i(fs1) : 0x7fffffff;
        g.r_min = ok ? f_floor_i(fr0) : -1; g.r_max = ok ? f_ceil_i(fr1) : 0x7fffffff;
    }
    g.P = 0; g.Q = 0;
    g.i00 = g.i01 = g.i10 = g.i11 = 0.0f;
    if (fabsf(det) > 1e-5f * scale + 1e-20f && mag < 1e6f) {
        const float r = 1.0f / det;
        g.i00 = g.mvi * r; g.i01 = -g.mui * r; g.i10 = -g.mvj * r; g.i11 = g.muj * r;
        const float T = 2.0f + 2.0f * g.slack;
        const float ej = (fabsf(g.i00) + fabsf(g.i01)) * T * 1.001f, ei = (fabsf(g.i10) + fabsf(g.i11)) * T * 1.001f;
        if (ej < 1e4f && ei < 1e4f) {
            const int q = ej <= 1.0f ? 1 : f_ceil_i(ej), pp = ei <= 1.0f ? 1 : f_ceil_i(ei);
            if ((long long)pp * q <= kMaxScatterPhases) { g.P = pp; g.Q = q; }
        }
    }
    return g;
}

// Index box [i_lo,i_hi] x [j_lo,j_hi] of the crop pixels that can touch frame tile rows [r0,r0+tr) x cols [s0,s0+tw)
// (unpadded).  Returns false when empty.  A crop pixel touches the tile iff v0 in [r0, r0+tr] and u0 in [s0, s0+tw]
// (padded tap indices), i.e. v in [r0, r0+tr+1), u in [s0, s0+tw+1).
STN_HD bool scatter_box(const ScatterGeom &g, int r0, int tr, int s0, int tw, int oH, int oW,
                        int &i_lo, int &i_hi, int &j_lo, int &j_hi)
{
    const float hu = 0.5f * (float)(tw + 1) + g.slack, hv = 0.5f * (float)(tr + 1) + g.slack;
    const float du = (float)s0 + 0.5f * (float)(tw + 1) - g.cu, dv = (float)r0 + 0.5f * (float)(tr + 1) - g.cv;
    const float cj = g.i00 * du + g.i01 * dv, ci = g.i10 * du + g.i11 * dv;
    const float ej = fabsf(g.i00) * hu + fabsf(g.i01) * hv, ei = fabsf(g.i10) * hu + fabsf(g.i11) * hv;
    const float sj = 0.01f + 2e-6f * (fabsf(cj) + ej), si = 0.01f + 2e-6f * (fabsf(ci) + ei);
    const float jl = fmaxf(cj - ej - sj, 0.0f), jh = fminf(cj + ej + sj, (float)(oW - 1));
    const float il = fmaxf(ci - ei - si, 0.0f), ih = fminf(ci + ei + si, (float)(oH - 1));
    if (!(jl <= jh) || !(il <= ih)) return false;
    j_lo = f_ceil_i(jl); j_hi = f_floor_i(jh);
    i_lo = f_ceil_i(il); i_hi = f_floor_i(ih);
    return j_lo <= j_hi && i_lo <= i_hi;
}

// cheap conservative test: can crop pixel (i,j) touch the tile at all?
STN_HD bool scatter_pretest(const ScatterGeom &g, int i, int j, int r0, int tr, int s0, int tw)
{
    const float fi = (float)i, fj = (float)j;
    const float u = g.muj * fj + g.mui * fi + g.cu, v = g.mvj * fj + g.mvi * fi + g.cv;
    return u >= (float)s0 - g.slack && u <= (float)(s0 + tw + 1) + g.slack &&
           v >= (float)r0 - g.slack && v <= (float)(r0 + tr + 1) + g.slack;
}

// The exact part of the scatter: forward chain of crop pixel (xsj, ysi) and which of its four taps land inside
// the tile rows [r0,r0+tr) x cols [s0,s0+tw) (unpadded) with a non-zero weight.  Zero-weight taps are dropped:
// they add nothing, and clipped (out-of-image) samples -- which the phase argument does not cover -- only ever
// have zero-weight or zero-frame taps.
struct ScatterTaps {
    Tap t;
    int row0, col0;          // tile-relative position of tap (v0,u0)
    bool rv0, rv1, cv0, cv1;
};

STN_HD bool scatter_taps(const Theta &th, float xsj, float ysi, int H, int W, int r0, int tr, int s0, int tw,
                         ScatterTaps &o)
{
    o.t = make_tap(grid_elem(th.t00, th.t01, th.t02, xsj, ysi), grid_elem(th.t10, th.t11, th.t12, xsj, ysi), H, W);
    o.row0 = o.t.v0 - 1 - r0;
    o.col0 = o.t.u0 - 1 - s0;
    o.rv0 = o.row0 >= 0 && o.row0 < tr && o.t.wv1 != 0.0f;
    o.rv1 = o.row0 + 1 >= 0 && o.row0 + 1 < tr && o.t.wv0 != 0.0f;
    o.cv0 = o.col0 >= 0 && o.col0 < tw && o.t.wu1 != 0.0f;
    o.cv1 = o.col0 + 1 >= 0 && o.col0 + 1 < tw && o.t.wu0 != 0.0f;
    return (o.rv0 || o.rv1) && (o.cv0 || o.cv1);
}

// first index >= lo that is congruent to c modulo m
STN_HD int first_congruent(int lo, int c, int m) { return lo + (((c - lo) % m) + m) % m; }

// ---------------------------------------------------------------------------------------------------
// Separable (axis-aligned) transforms: theta01 * mask == theta10 * mask == 0, which is what LoANs always runs
// (rotation_dropout(..., ratio=0.0), sheep/sheep_localizer.py:61).  Then u depends on j only and v on i only, so
// the whole coordinate chain is evaluated oW + oH times per crop instead of oH * oW times: one AxisTap per crop
// column and one per crop row, bit-identical to what make_tap() yields for every pixel of that column / row.
struct AxisTap {
    float g;        // grid value (what spatial_transformer_grid outputs for this column / row)
    float coord;    // unclipped padded coordinate (gradient mask)
    float w0, w1;   // weight of tap idx0 + 1, weight of tap idx0
    int idx0;       // first tap, padded index space, in [0, size]
    float lin;      // the linspace value xs[j] / ys[i] itself (theta-gradient sums)
};

// t_lin * lin + (t_rot_masked * other) + t_shift with t_rot_masked == +-0: the same operations as grid_elem
STN_HD AxisTap make_axis_tap(float t_lin, float t_rot_masked, float t_shift, float lin, bool lin_is_x, int size)
{
    AxisTap a;
    a.lin = lin;
    // grid_elem(t0, t1, t2, xs, ys): row 0 has t0 = t_lin (times xs), t1 = t_rot (times ys);
    //                                row 1 has t0 = t_rot (times xs), t1 = t_lin (times ys)
    a.g = lin_is_x ? grid_elem(t_lin, t_rot_masked, t_shift, lin, 0.0f) : grid_elem(t_rot_masked, t_lin, t_shift, 0.0f, lin);
    a.coord = to_padded_px(a.g, (float)(size - 1));
    const float c = fminf(fmaxf(a.coord, 0.0f), (float)(size + 1));
    int i0 = f_floor_i(c);
    i0 = i0 < 0 ? 0 : (i0 > size ? size : i0);
    a.idx0 = i0;
    a.w0 = f_sub(c, (float)i0);
    a.w1 = f_sub((float)(i0 + 1), c);
    return a;
}

STN_HD Tap tap_from_axes(const AxisTap &col, const AxisTap &row)
{
    Tap t;
    t.u = col.coord; t.v = row.coord;
    t.wu0 = col.w0; t.wu1 = col.w1; t.wv0 = row.w0; t.wv1 = row.w1;
    t.u0 = col.idx0; t.v0 = row.idx0;
    return t;
}

}  // namespace stn
