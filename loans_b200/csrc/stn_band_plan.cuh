// stn_band_plan.cuh -- planning arithmetic of the band backward (stn_band.cu), __host__ __device__ so that
// tests/hostemu can run the very same code on the CPU (a test harness, never a product path).
//
// The band backward handles what LoANs always runs: one crop per frame, rotation terms masked to zero
// (rotation_dropout(..., ratio=0.0), reference sheep/sheep_localizer.py:61), crop box not mirrored.  Then the
// padded sample coordinate u depends on the crop column only and v on the crop row only, both monotone, and the
// gradient of the frame splits by crop rows:
//
//   * a BAND is a run of crop rows [a, b).  It OWNS the frame rows [Fa, Fb), Fa = first frame row row a touches,
//     Fb = first frame row row b touches (after the last crop row: the end E of the rows the crop touches).  The
//     bands of a crop partition the frame rows [L, E) the crop spans; the rows above L and below E are all zero
//     and are dealt evenly to the CTAs of the crop (band_edge_rows).  So every gx element is written exactly
//     once, zeros included, by the CTA that owns it -- no memset pass, no atomics;
//   * the band evaluates each of its crop pixels ONCE: taps + gy -> d/du, d/dv for gtheta, and gy * wu * wv
//     added into a shared-memory tile of the frame rows it owns.  When the crop steps by less than ~2 frame
//     pixels per crop pixel, neighbouring crop pixels can share a frame pixel: pixels are then processed in
//     P * Q phases ((i mod P, j mod Q) constant per phase), inside which no two pixels share a frame pixel, and
//     the band re-evaluates the P - 1 crop rows above it (its halo) for the taps that land in its rows;
//   * tile rows and the all-zero rows in between leave shared memory as TMA bulk copies.
//
// Two tile layouts: COMPACT (P == 1: crop rows touch disjoint frame row pairs; slot 2*(i-a)+t holds tap row t of
// crop row i, the untouched frame rows in between are stored from a zero plane) and DENSE (P > 1: slot r - base
// holds frame row r; the touched rows are contiguous up to single-row gaps).
#pragma once
#include "stn_math.cuh"

namespace stn {

constexpr int kBandMaxPhases = 16;        // P * Q beyond this: the crop goes to the general roles
constexpr int kBandMaxHalo = kBandMaxPhases - 1;

// One crop column or crop row, 16 bytes.  code: bits 0..23 idx0 (first tap, padded index space [0,size]),
// kAxActive: the unclipped coordinate is strictly inside the padded image (the reference's gradient mask),
// kAxTap0 / kAxTap1: tap idx0 / idx0+1 is a real pixel AND carries a non-zero weight (it scatters).
struct alignas(16) BandAxis {
    float w0, w1;   // weight of tap idx0 + 1, weight of tap idx0 (exact differences, as make_tap)
    float lin;      // the linspace value xs[j] / ys[i]
    int code;
};
constexpr int kAxIdxMask = 0xffffff, kAxActive = 1 << 24, kAxTap0 = 1 << 25, kAxTap1 = 1 << 26;

STN_HD BandAxis make_band_axis(float t_lin, float t_rot_masked, float t_shift, float lin, bool lin_is_x, int size)
{
    const AxisTap a = make_axis_tap(t_lin, t_rot_masked, t_shift, lin, lin_is_x, size);
    BandAxis b;
    b.w0 = a.w0;
    b.w1 = a.w1;
    b.lin = lin;
    int code = a.idx0;
    if (a.coord > 0.0f && a.coord < (float)(size + 1)) code |= kAxActive;
    if (a.idx0 >= 1 && a.w1 != 0.0f) code |= kAxTap0;
    if (a.idx0 <= size - 1 && a.w0 != 0.0f) code |= kAxTap1;
    b.code = code;
    return b;
}

STN_HD Tap tap_from_band_axes(const BandAxis &col, const BandAxis &row, int H, int W)
{
    // u / v are only used for the strict in-image test of finish_grad_uv: any value with the same verdict does
    Tap t;
    t.u = (col.code & kAxActive) ? 1.0f : 0.0f;
    t.v = (row.code & kAxActive) ? 1.0f : 0.0f;
    t.wu0 = col.w0; t.wu1 = col.w1; t.wv0 = row.w0; t.wv1 = row.w1;
    t.u0 = col.code & kAxIdxMask; t.v0 = row.code & kAxIdxMask;
    (void)H; (void)W;
    return t;
}

// Per-crop verdict: can the band path take this crop, and with which phase periods.
struct BandCrop {
    int ok, P, Q;
};

STN_HD BandCrop make_band_crop(const Theta &th, int H, int W, int oH, int oW)
{
    BandCrop c;
    c.ok = 0; c.P = 1; c.Q = 1;
    if (!(th.t01 == 0.0f && th.t10 == 0.0f)) return c;        // rotation terms must be masked to (+-)0
    if (!(th.t00 > 0.0f && th.t11 > 0.0f)) return c;          // mirrored or degenerate boxes (and NaN): general roles
    if (H > kAxIdxMask - 2 || W > kAxIdxMask - 2) return c;
    const float sx = oW > 1 ? 2.0f / (float)(oW - 1) : 0.0f;
    const float sy = oH > 1 ? 2.0f / (float)(oH - 1) : 0.0f;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    const float muj = th.t00 * sx * hw, mvi = th.t11 * sy * hh;          // frame pixels per crop pixel
    const float cu = (th.t02 - th.t00 + 1.0f) * hw + 1.0f, cv = (th.t12 - th.t11 + 1.0f) * hh + 1.0f;
    const float nj = (float)(oW > 1 ? oW - 1 : 1), ni = (float)(oH > 1 ? oH - 1 : 1);
    const float mag = muj * nj + fabsf(cu) + mvi * ni + fabsf(cv) + (float)(W + H);
    if (!(mag < 1e6f)) return c;
    // two crop pixels share a frame pixel only if their coordinates differ by less than 2; T adds the float32 slack
    const float T = (2.0f + 2.0f * (0.05f + 1e-5f * mag)) * 1.001f;
    int P = 1, Q = 1;
    if (oW > 1 && muj < T) {
        if (!(muj * (float)kBandMaxPhases >= T)) return c;
        Q = f_ceil_i(T / muj);
    }
    if (oH > 1 && mvi < T) {
        if (!(mvi * (float)kBandMaxPhases >= T)) return c;
        P = f_ceil_i(T / mvi);
    }
    if (P * Q > kBandMaxPhases) return c;
    c.ok = 1; c.P = P; c.Q = Q;
    return c;
}

// Frame rows [lo, hi) a crop row can touch (clipped to the frame).  Rows whose coordinate is clipped touch
// nothing: they collapse to [0,0) above the frame and [H,H) below it, which keeps lo / hi monotone in the crop row.
STN_HD void band_row_range(const BandAxis &a, int H, int &lo, int &hi)
{
    const int idx0 = a.code & kAxIdxMask;
    if (a.code & kAxActive) {
        lo = idx0 - 1 < 0 ? 0 : idx0 - 1;
        hi = idx0 + 1 > H ? H : idx0 + 1;
    } else if (idx0 == 0) {
        lo = hi = 0;
    } else {
        lo = hi = H;
    }
}

struct BandPlan {
    int a, b;        // crop rows of the band
    int h;           // first halo row (h <= a): rows [h, b) are evaluated, rows [a, b) count for gtheta
    int Fa, Fb;      // frame rows owned
    int base, tend;  // DENSE: tile slot s holds frame row base + s, rows [base, tend) are stored from the tile
    int compact;
    int nslots;      // tile rows in use (to be zeroed)
};

// Plans the band that starts at crop row a inside a CTA's row range [.., i1).  rowtab[i - t0] describes crop row i
// and must cover rows [max(a - (P-1), 0), min(i1 + 1, oH)).  cap = tile rows available (>= 2), max_rows = crop rows
// per band at most, E = end of the frame rows the crop touches (band_row_range of the last crop row: its hi).
STN_HD BandPlan plan_band(const BandAxis *rowtab, int t0, int a, int i1, int oH, int H, int P, int cap, int max_rows, int E)
{
    BandPlan pl;
    pl.a = a;
    pl.compact = P == 1;
    pl.h = a - (P - 1) < 0 ? 0 : a - (P - 1);
    int lo_a, hi_a;
    band_row_range(rowtab[a - t0], H, lo_a, hi_a);
    pl.Fa = lo_a;
    pl.base = lo_a;
    int b, tend = lo_a;
    if (pl.compact) {
        const int nr = cap / 2 < max_rows ? cap / 2 : max_rows;
        b = a + nr < i1 ? a + nr : i1;
        pl.nslots = 2 * (b - a);
    } else {
        b = a;
        while (b < i1) {
            int lo, hi;
            band_row_range(rowtab[b - t0], H, lo, hi);
            int e = (rowtab[b - t0].code & kAxActive) && hi > tend ? hi : tend;
            if ((e - pl.base > cap || b - a >= max_rows) && b > a) break;
            tend = e;
            ++b;
        }
        pl.nslots = tend - pl.base < cap ? tend - pl.base : cap;    // a single row always fits (cap >= 2)
    }
    pl.b = b;
    if (b >= oH) pl.Fb = E;
    else {
        int lo, hi;
        band_row_range(rowtab[b - t0], H, lo, hi);
        pl.Fb = lo;
    }
    if (pl.Fb < pl.Fa) pl.Fb = pl.Fa;                                   // cannot happen (monotone); keeps spans sane
    pl.tend = tend < pl.Fb ? tend : pl.Fb;
    if (pl.tend < pl.base) pl.tend = pl.base;
    return pl;
}

// Tile slots of the two tap rows of crop row i (halo rows included); -1: that tap adds nothing to this band's tile
// (zero frame, zero weight, or a frame row another band owns).
STN_HD void band_row_slots(const BandPlan &pl, const BandAxis &row, int i, int &s0, int &s1)
{
    const int r0 = (row.code & kAxIdxMask) - 1;
    s0 = s1 = -1;
    if (pl.compact) {
        if (row.code & kAxTap0) s0 = 2 * (i - pl.a);
        if (row.code & kAxTap1) s1 = 2 * (i - pl.a) + 1;
    } else {
        if ((row.code & kAxTap0) && r0 >= pl.base && r0 < pl.tend) s0 = r0 - pl.base;
        if ((row.code & kAxTap1) && r0 + 1 >= pl.base && r0 + 1 < pl.tend) s1 = r0 + 1 - pl.base;
    }
}

// The all-zero frame rows above and below the crop, [0, L) and [E, H), dealt evenly to the cs CTAs of the crop:
// CTA `rank` zero-fills rows [r0a, r0a + na) and [r0b, r0b + nb).  L = lo of crop row 0, E = hi of the last crop row.
STN_HD void band_edge_rows(int L, int E, int H, int rank, int cs, int &r0a, int &na, int &r0b, int &nb)
{
    const int Z = L + (H - E);                       // <= H < 2^24 and cs <= 16: the products below fit 32 bits
    const int z0 = Z * rank / cs, z1 = Z * (rank + 1) / cs;
    const int a1 = z1 < L ? z1 : L;
    r0a = z0; na = a1 - z0 > 0 ? a1 - z0 : 0;
    const int b0 = z0 > L ? z0 : L;
    r0b = E + (b0 - L); nb = z1 - b0 > 0 ? z1 - b0 : 0;
}

// The band's write-out as spans of whole frame rows: rows [row, row + nrows) of every channel come from tile slot
// `slot` onwards (slot >= 0) or are zero (slot < 0).  Span index t runs over [0, band_span_count(pl)); even t are
// zero spans, odd t tile spans; empty spans have nrows <= 0.  Together they cover [Fa, Fb) exactly once.
struct BandSpan {
    int row, nrows, slot;
};

STN_HD int band_span_count(const BandPlan &pl) { return pl.compact ? 2 * (pl.b - pl.a) + 1 : 3; }

STN_HD BandSpan band_span(const BandPlan &pl, const BandAxis *rowtab, int t0, int H, int t)
{
    BandSpan s;
    s.slot = -1;
    int lo, hi;
    if (!pl.compact) {
        if (t == 0) { lo = pl.Fa; hi = pl.base; }
        else if (t == 1) { lo = pl.base; hi = pl.tend; s.slot = 0; }
        else { lo = pl.tend; hi = pl.Fb; }
    } else {
        const int k = t >> 1;                            // crop row a + k
        if (t & 1) {
            band_row_range(rowtab[pl.a + k - t0], H, lo, hi);
            const int r0 = (rowtab[pl.a + k - t0].code & kAxIdxMask) - 1;
            if (lo < pl.Fa) lo = pl.Fa;
            s.slot = 2 * k + (lo - r0);
        } else {
            int l2, h2;
            if (k == 0) lo = pl.Fa;
            else { band_row_range(rowtab[pl.a + k - 1 - t0], H, l2, lo); }
            if (k == pl.b - pl.a) hi = pl.Fb;
            else { band_row_range(rowtab[pl.a + k - t0], H, hi, h2); }
        }
    }
    if (lo < pl.Fa) lo = pl.Fa;
    if (hi > pl.Fb) hi = pl.Fb;
    s.row = lo;
    s.nrows = hi - lo;
    return s;
}

// ---- several crops per frame (stn_kframe.cu): the per-crop verdict of the row-owner gx kernel.  Here (host + device) so that
// tests/hostemu can check its one assumption -- a crop it takes puts at most one crop row on a frame row and at most one crop
// column on a frame column -- with the very same arithmetic.
struct KfCrop {                       // one crop of the frame
    float t00, t01, t02, t10, t11, t12;   // masked theta
    int ok;                           // the crop is taken by this path
    int rmin, rmax;                   // unpadded frame rows any tap of the crop can touch (conservative)
};

// (idx0 + 1) - c and 1 - (c - idx0) are the same real number and both exactly representable (c - idx0 is exact and a
// multiple of ulp(c)), so the second tap weight follows from the first without a rounding of its own
STN_HD float kf_w1(float w0) { return f_sub(1.0f, w0); }

STN_HD KfCrop make_kf_crop(const Theta &th, int H, int W, int oH, int oW)
{
    KfCrop c;
    c.t00 = th.t00; c.t01 = th.t01; c.t02 = th.t02; c.t10 = th.t10; c.t11 = th.t11; c.t12 = th.t12;
    c.ok = 0;
    c.rmin = 0; c.rmax = -1;
    if (!(th.t01 == 0.0f && th.t10 == 0.0f)) return c;                 // rotation terms must be masked to (+-)0
    if (!(th.t00 > 0.0f && th.t11 > 0.0f)) return c;                   // mirrored or degenerate boxes (and NaN): general role
    const float sx = oW > 1 ? 2.0f / (float)(oW - 1) : 0.0f, sy = oH > 1 ? 2.0f / (float)(oH - 1) : 0.0f;
    const float hw = 0.5f * (float)(W - 1), hh = 0.5f * (float)(H - 1);
    const float muj = th.t00 * sx * hw, mvi = th.t11 * sy * hh;         // frame pixels per crop pixel
    const float cu = (th.t02 - th.t00 + 1.0f) * hw + 1.0f, cv = (th.t12 - th.t11 + 1.0f) * hh + 1.0f;
    const float nj = (float)(oW > 1 ? oW - 1 : 1), ni = (float)(oH > 1 ? oH - 1 : 1);
    const float mag = muj * nj + fabsf(cu) + mvi * ni + fabsf(cv) + (float)(W + H);
    if (!(mag < 1e6f)) return c;                                       // also NaN / inf
    // two crop pixels share a frame pixel only if their coordinates differ by less than 2; T adds the float32 slack
    const float slack = 0.05f + 1e-5f * mag;
    const float T = (2.0f + 2.0f * slack) * 1.001f;
    if (oW > 1 && muj < T) return c;
    if (oH > 1 && mvi < T) return c;
    c.ok = 1;
    const float v_a = cv, v_b = cv + mvi * ni, m = 2.0f + slack;
    // padded coordinate p touches unpadded rows floor(p) - 1 and floor(p)
    c.rmin = f_floor_i(fmaxf(fminf(v_a, v_b) - m - 1.0f, -4.0f));
    c.rmax = f_ceil_i(fminf(fmaxf(v_a, v_b) + m, 2.0e9f));
    return c;
}

}  // namespace stn
