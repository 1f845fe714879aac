// Host emulation harness -- TEST INFRASTRUCTURE ONLY, never loaded by the product.
// Runs the per-pixel device code of loans_b200/csrc/stn_math.cuh (the same header the CUDA kernels include)
// on the CPU, pixel by pixel, so that the arithmetic and above all the inverse-mapping gather for gx can be
// checked against the oracle in the CPU-only container.  Built with g++ -ffp-contract=off by
// tests/test_hostemu.py.
#include <cstddef>
#include <vector>

#include "../../loans_b200/csrc/stn_math.cuh"
#include "../../loans_b200/csrc/stn_band_plan.cuh"

using namespace stn;

struct LoadF {
    float operator()(const float *p, size_t i) const { return p[i]; }
};

extern "C" {

void emu_crop_fwd(const float *x, const float *theta, float mask01, float *y, float *grid,
                  int n, int k, int c, int h, int w, int oh, int ow)
{
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    const int npx = oh * ow;
    const size_t plane = (size_t)h * w;
    for (int b = 0; b < n; ++b) {
        const Theta th = load_theta_masked(theta + 6 * b, mask01);
        const float *xb = x + (size_t)(b / k) * c * plane;
        for (int q = 0; q < npx; ++q) {
            const int i = q / ow, j = q - i * ow;
            const float xs = linspace_pm1(j, ow, xstep), ys = linspace_pm1(i, oh, ystep);
            const float g0 = grid_elem(th.t00, th.t01, th.t02, xs, ys), g1 = grid_elem(th.t10, th.t11, th.t12, xs, ys);
            if (grid) { grid[(size_t)b * 2 * npx + q] = g0; grid[(size_t)b * 2 * npx + npx + q] = g1; }
            const Tap t = make_tap(g0, g1, h, w);
            const TapAddr a = make_tap_addr(t, h, w);
            const Weights4 wt = make_weights(t);
            for (int ch = 0; ch < c; ++ch) {
                float x1, x2, x3, x4;
                load_taps(xb + ch * plane, a, w, x1, x2, x3, x4);
                y[((size_t)b * c + ch) * npx + q] = interp(wt, x1, x2, x3, x4);
            }
        }
    }
}

void emu_crop_bwd(const float *x, const float *theta, float mask01, const float *gy, const float *ggrid_up,
                  float *gtheta, float *gx, float *ggrid_out, int n, int k, int c, int h, int w, int oh, int ow)
{
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    const int npx = oh * ow;
    const size_t plane = (size_t)h * w;
    std::vector<float> xs(ow), ys(oh);
    for (int j = 0; j < ow; ++j) xs[j] = linspace_pm1(j, ow, xstep);
    for (int i = 0; i < oh; ++i) ys[i] = linspace_pm1(i, oh, ystep);
    for (int b = 0; b < n; ++b) {
        const Theta th = load_theta_masked(theta + 6 * b, mask01);
        const float *xb = x + (size_t)(b / k) * c * plane;
        const float *gyb = gy + (size_t)b * c * npx;
        float s[6] = {0, 0, 0, 0, 0, 0};
        for (int q = 0; q < npx; ++q) {
            const int i = q / ow, j = q - i * ow;
            const Tap t = make_tap(grid_elem(th.t00, th.t01, th.t02, xs[j], ys[i]),
                                   grid_elem(th.t10, th.t11, th.t12, xs[j], ys[i]), h, w);
            const TapAddr a = make_tap_addr(t, h, w);
            float su = 0, sv = 0;
            for (int ch = 0; ch < c; ++ch) {
                float x1, x2, x3, x4, gu, gv;
                load_taps(xb + ch * plane, a, w, x1, x2, x3, x4);
                grad_uv(t, x1, x2, x3, x4, gu, gv);
                const float g = gyb[(size_t)ch * npx + q];
                gu = f_mul(gu, g); gv = f_mul(gv, g);
                if (ch == 0) { su = gu; sv = gv; } else { su = f_add(su, gu); sv = f_add(sv, gv); }
            }
            finish_grad_uv(t, h, w, su, sv);
            if (ggrid_out) { ggrid_out[(size_t)b * 2 * npx + q] = su; ggrid_out[(size_t)b * 2 * npx + npx + q] = sv; }
            if (ggrid_up) {
                su = f_add(su, ggrid_up[(size_t)b * 2 * npx + q]);
                sv = f_add(sv, ggrid_up[(size_t)b * 2 * npx + npx + q]);
            }
            s[0] += su * xs[j]; s[1] += su * ys[i]; s[2] += su;
            s[3] += sv * xs[j]; s[4] += sv * ys[i]; s[5] += sv;
        }
        s[1] = f_mul(s[1], mask01); s[3] = f_mul(s[3], mask01);
        for (int e = 0; e < 6; ++e) gtheta[6 * b + e] = s[e];
    }
    if (!gx) return;
    const int frames = n / k;
    std::vector<InvCrop> inv(k);
    for (int f = 0; f < frames; ++f) {
        for (int kk = 0; kk < k; ++kk)
            inv[kk] = make_inv_crop(load_theta_masked(theta + 6 * (f * k + kk), mask01), h, w, oh, ow);
        for (int r = 0; r < h; ++r)
            for (int sx = 0; sx < w; ++sx)
                for (int c0 = 0; c0 < c; c0 += 3) {
                    const int nc = c - c0 < 3 ? c - c0 : 3;
                    float acc[3] = {0, 0, 0};
                    for (int kk = 0; kk < k; ++kk)
                        gather_from_crop<3>(inv[kk], xs.data(), ys.data(), h, w, oh, ow, r + 1, sx + 1,
                                            gy + ((size_t)(f * k + kk) * c + c0) * npx, nc, LoadF(), acc);
                    for (int ch = 0; ch < nc; ++ch) gx[((size_t)f * c + c0 + ch) * plane + (size_t)r * w + sx] = acc[ch];
                }
    }
}

// The tile-scatter formulation of gx exactly as stn_bwd_kernel's gx role runs it (same geometry, box, pre-test,
// phases and exact tap code from stn_math.cuh), executed serially, with a per-phase write set: returns the number
// of shared-memory addresses that two crop pixels of the SAME phase wrote -- which on the GPU would be a race.
// stats[0] = crop pixels fully evaluated, stats[1] = crop pixels pre-tested, stats[2] = crops sent to the gather
// fallback, stats[3] = largest phase count.
long long emu_gx_scatter(const float *theta, float mask01, const float *gy, float *gx,
                         int n, int k, int c, int h, int w, int oh, int ow, int tile_rows, int tile_cols, long long *stats)
{
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    const int npx = oh * ow;
    const size_t plane = (size_t)h * w;
    std::vector<float> xs(ow), ys(oh);
    for (int j = 0; j < ow; ++j) xs[j] = linspace_pm1(j, ow, xstep);
    for (int i = 0; i < oh; ++i) ys[i] = linspace_pm1(i, oh, ystep);
    const int frames = n / k;
    const int twp = (tile_cols + 3) & ~3;
    long long conflicts = 0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    std::vector<float> tile((size_t)c * tile_rows * twp);
    std::vector<int> stamp((size_t)c * tile_rows * twp);
    std::vector<ScatterGeom> geom(k);
    std::vector<InvCrop> inv(k);
    int phase_id = 0;
    for (int f = 0; f < frames; ++f) {
        bool any_fb = false;
        for (int kk = 0; kk < k; ++kk) {
            const Theta th = load_theta_masked(theta + 6 * (f * k + kk), mask01);
            geom[kk] = make_scatter_geom(th, h, w, oh, ow);
            if (geom[kk].P == 0) { inv[kk] = make_inv_crop(th, h, w, oh, ow); any_fb = true; stats[2]++; }
            else if ((long long)geom[kk].P * geom[kk].Q > stats[3]) stats[3] = (long long)geom[kk].P * geom[kk].Q;
        }
        for (int r0 = 0; r0 < h; r0 += tile_rows)
            for (int s0 = 0; s0 < w; s0 += tile_cols) {
                const int tr = h - r0 < tile_rows ? h - r0 : tile_rows, tw = w - s0 < tile_cols ? w - s0 : tile_cols;
                std::fill(tile.begin(), tile.end(), 0.f);
                std::fill(stamp.begin(), stamp.end(), -1);
                const int tile_plane = tile_rows * twp;
                for (int kk = 0; kk < k; ++kk) {
                    const ScatterGeom &g = geom[kk];
                    if (g.P == 0) continue;
                    if (g.r_max < r0 || g.r_min >= r0 + tr || g.s_max < s0 || g.s_min >= s0 + tw) continue;   // as the kernel
                    int i_lo, i_hi, j_lo, j_hi;
                    if (!scatter_box(g, r0, tr, s0, tw, oh, ow, i_lo, i_hi, j_lo, j_hi)) continue;
                    const float *gyc = gy + (size_t)(f * k + kk) * c * npx;
                    for (int cp = 0; cp < g.P; ++cp)
                        for (int cq = 0; cq < g.Q; ++cq) {
                            ++phase_id;
                            for (int i = first_congruent(i_lo, cp, g.P); i <= i_hi; i += g.P)
                                for (int j = first_congruent(j_lo, cq, g.Q); j <= j_hi; j += g.Q) {
                                    stats[1]++;
                                    if (!scatter_pretest(g, i, j, r0, tr, s0, tw)) continue;
                                    stats[0]++;
                                    ScatterTaps st;
                                    if (!scatter_taps(g.th, xs[j], ys[i], h, w, r0, tr, s0, tw, st)) continue;
                                    for (int ch = 0; ch < c; ++ch) {
                                        const float gv = gyc[(size_t)ch * npx + (size_t)i * ow + j];
                                        const float a1 = f_mul(gv, st.t.wu1), a0 = f_mul(gv, st.t.wu0);
                                        const int base = ch * tile_plane + st.row0 * twp + st.col0;
                                        const int offs[4] = {0, 1, twp, twp + 1};
                                        const bool ok[4] = {st.rv0 && st.cv0, st.rv0 && st.cv1, st.rv1 && st.cv0, st.rv1 && st.cv1};
                                        const float val[4] = {f_mul(a1, st.t.wv1), f_mul(a0, st.t.wv1), f_mul(a1, st.t.wv0), f_mul(a0, st.t.wv0)};
                                        for (int t4 = 0; t4 < 4; ++t4)
                                            if (ok[t4]) {
                                                const int ad = base + offs[t4];
                                                if (stamp[ad] == phase_id) ++conflicts;
                                                stamp[ad] = phase_id;
                                                tile[ad] = f_add(tile[ad], val[t4]);
                                            }
                                    }
                                }
                        }
                }
                // everything the exact taps say touches the tile must have been found by box + pre-test: verified by
                // comparing the result with the oracle in the test
                for (int row = 0; row < tr; ++row)
                    for (int col = 0; col < tw; ++col)
                        for (int c0 = 0; c0 < c; c0 += 3) {
                            const int nc = c - c0 < 3 ? c - c0 : 3;
                            float acc[3] = {0, 0, 0};
                            for (int ch = 0; ch < nc; ++ch) acc[ch] = tile[(c0 + ch) * tile_plane + row * twp + col];
                            if (any_fb)
                                for (int kk = 0; kk < k; ++kk)
                                    if (geom[kk].P == 0)
                                        gather_from_crop<3>(inv[kk], xs.data(), ys.data(), h, w, oh, ow, r0 + row + 1, s0 + col + 1,
                                                            gy + ((size_t)(f * k + kk) * c + c0) * npx, nc, LoadF(), acc);
                            for (int ch = 0; ch < nc; ++ch)
                                gx[((size_t)f * c + c0 + ch) * plane + (size_t)(r0 + row) * w + s0 + col] = acc[ch];
                        }
            }
    }
    return conflicts;
}

// The band backward (stn_band.cu) exactly as its CTAs run it -- same planning code (stn_band_plan.cuh), same
// per-pixel arithmetic -- executed serially: `cs` CTAs per crop, `cap` tile rows.  Checks what cannot be seen from
// the result alone: returns the number of tile addresses two crop pixels of the SAME phase wrote (a race on the
// GPU); stats[0] = gx elements of band-path frames not written exactly once, stats[1] = largest phase count,
// stats[2] = bands, stats[3] = tile slots out of range.  ok[n] = 1 where the band path takes the crop; the other
// frames / crops are left untouched (the kernel sends them to the general roles).
long long emu_band_bwd(const float *x, const float *theta, float mask01, const float *gy, const float *ggrid_up,
                       float *gtheta, float *gx, float *ggrid_out, int *ok,
                       int n, int c, int h, int w, int oh, int ow, int cs, int cap, int max_rows, long long *stats)
{
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    const int npx = oh * ow;
    const size_t plane = (size_t)h * w;
    long long conflicts = 0;
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    std::vector<BandAxis> coltab(ow), rowtab(oh + 1);
    std::vector<float> tile((size_t)c * cap * w);
    std::vector<int> stamp((size_t)c * cap * w);
    std::vector<int> written(plane * c);
    int phase_id = 0;
    for (int b = 0; b < n; ++b) {
        const Theta th = load_theta_masked(theta + 6 * b, mask01);
        const BandCrop bc = make_band_crop(th, h, w, oh, ow);
        ok[b] = bc.ok;
        if (!bc.ok) continue;
        if (bc.P * bc.Q > stats[1]) stats[1] = bc.P * bc.Q;
        const float *xb = x + (size_t)b * c * plane;
        const float *gyb = gy + (size_t)b * c * npx;
        float *gxb = gx + (size_t)b * c * plane;
        std::fill(written.begin(), written.end(), 0);
        for (int j = 0; j < ow; ++j) coltab[j] = make_band_axis(th.t00, th.t01, th.t02, linspace_pm1(j, ow, xstep), true, w);
        float s[6] = {0, 0, 0, 0, 0, 0};
        const int rows_cta = (oh + cs - 1) / cs;
        int L, E, tmp;
        band_row_range(make_band_axis(th.t11, th.t10, th.t12, linspace_pm1(0, oh, ystep), false, h), h, L, tmp);
        band_row_range(make_band_axis(th.t11, th.t10, th.t12, linspace_pm1(oh - 1, oh, ystep), false, h), h, tmp, E);
        for (int rank = 0; rank < cs; ++rank) {
            const int i0 = rank * rows_cta, i1 = i0 + rows_cta < oh ? i0 + rows_cta : oh;
            {   // this CTA's share of the all-zero rows above and below the crop
                int ra, na, rb, nb;
                band_edge_rows(L, E, h, rank, cs, ra, na, rb, nb);
                for (int ch = 0; ch < c; ++ch)
                    for (int part = 0; part < 2; ++part)
                        for (int r = 0; r < (part ? nb : na); ++r)
                            for (int col = 0; col < w; ++col) {
                                const size_t o = (size_t)ch * plane + (size_t)((part ? rb : ra) + r) * w + col;
                                gxb[o] = 0.0f;
                                written[o]++;
                            }
            }
            if (i0 >= i1) continue;
            const int t0 = i0 - (bc.P - 1) < 0 ? 0 : i0 - (bc.P - 1), t1 = i1 + 1 < oh ? i1 + 1 : oh;
            for (int i = t0; i < t1; ++i)
                rowtab[i - t0] = make_band_axis(th.t11, th.t10, th.t12, linspace_pm1(i, oh, ystep), false, h);
            for (int a = i0; a < i1;) {
                const BandPlan pl = plan_band(rowtab.data(), t0, a, i1, oh, h, bc.P, cap, max_rows, E);
                stats[2]++;
                if (pl.nslots > cap || pl.b <= a) { stats[3]++; return -1; }
                std::fill(tile.begin(), tile.end(), 0.f);
                std::fill(stamp.begin(), stamp.end(), -1);
                const int tile_plane = cap * w;
                for (int ph = 0; ph < bc.P * bc.Q; ++ph) {
                    ++phase_id;
                    for (int i = pl.h; i < pl.b; ++i)
                        for (int j = 0; j < ow; ++j) {
                            if ((i % bc.P) * bc.Q + (j % bc.Q) != ph) continue;
                            const BandAxis &col = coltab[j], &row = rowtab[i - t0];
                            const Tap t = tap_from_band_axes(col, row, h, w);
                            const TapAddr ad = make_tap_addr(t, h, w);
                            int s0, s1;
                            band_row_slots(pl, row, i, s0, s1);
                            if (s0 >= cap || s1 >= cap) { stats[3]++; return -1; }
                            float su = 0, sv = 0;
                            for (int ch = 0; ch < c; ++ch) {
                                float x1, x2, x3, x4, gu, gv;
                                load_taps(xb + ch * plane, ad, w, x1, x2, x3, x4);
                                const float g = gyb[(size_t)ch * npx + (size_t)i * ow + j];
                                grad_uv(t, x1, x2, x3, x4, gu, gv);
                                gu = f_mul(gu, g); gv = f_mul(gv, g);
                                if (ch == 0) { su = gu; sv = gv; } else { su = f_add(su, gu); sv = f_add(sv, gv); }
                                const float a1 = f_mul(g, t.wu1), a0 = f_mul(g, t.wu0);
                                const int cix = t.u0 - 1;
                                const int adr[4] = {s0 * w + cix, s0 * w + cix + 1, s1 * w + cix, s1 * w + cix + 1};
                                const bool okt[4] = {s0 >= 0 && (col.code & kAxTap0), s0 >= 0 && (col.code & kAxTap1),
                                                     s1 >= 0 && (col.code & kAxTap0), s1 >= 0 && (col.code & kAxTap1)};
                                const float val[4] = {f_mul(a1, t.wv1), f_mul(a0, t.wv1), f_mul(a1, t.wv0), f_mul(a0, t.wv0)};
                                for (int t4 = 0; t4 < 4; ++t4)
                                    if (okt[t4]) {
                                        const int o = ch * tile_plane + adr[t4];
                                        if (stamp[o] == phase_id) ++conflicts;
                                        stamp[o] = phase_id;
                                        tile[o] = f_add(tile[o], val[t4]);
                                    }
                            }
                            if (i < pl.a) continue;                 // halo row: scatter only
                            finish_grad_uv(t, h, w, su, sv);
                            const size_t q = (size_t)i * ow + j;
                            if (ggrid_out) { ggrid_out[(size_t)b * 2 * npx + q] = su; ggrid_out[(size_t)b * 2 * npx + npx + q] = sv; }
                            if (ggrid_up) {
                                su = f_add(su, ggrid_up[(size_t)b * 2 * npx + q]);
                                sv = f_add(sv, ggrid_up[(size_t)b * 2 * npx + npx + q]);
                            }
                            s[0] += su * col.lin; s[1] += su * row.lin; s[2] += su;
                            s[3] += sv * col.lin; s[4] += sv * row.lin; s[5] += sv;
                        }
                }
                const int ns = band_span_count(pl);
                for (int t = 0; t < ns; ++t) {
                    const BandSpan sp = band_span(pl, rowtab.data(), t0, h, t);
                    if (sp.nrows <= 0) continue;
                    if (sp.slot >= 0 && sp.slot + sp.nrows > cap) { stats[3]++; return -1; }
                    for (int ch = 0; ch < c; ++ch)
                        for (int r = 0; r < sp.nrows; ++r)
                            for (int col = 0; col < w; ++col) {
                                const size_t o = (size_t)ch * plane + (size_t)(sp.row + r) * w + col;
                                gxb[o] = sp.slot >= 0 ? tile[(size_t)ch * tile_plane + (size_t)(sp.slot + r) * w + col] : 0.0f;
                                written[o]++;
                            }
                }
                a = pl.b;
            }
        }
        for (size_t o = 0; o < plane * c; ++o) stats[0] += written[o] != 1;
        s[1] = f_mul(s[1], mask01); s[3] = f_mul(s[3], mask01);
        for (int e = 0; e < 6; ++e) gtheta[6 * b + e] = s[e];
    }
    return conflicts;
}

// The row-owner gx kernel of stn_kframe.cu (several crops per frame), executed serially with the kernel's own planning code
// (make_kf_crop, make_band_axis): per frame the verdict of every crop, the inverse row map, then frame row by frame row the crop
// rows on it in ascending crop order, gy * wu * wv added into a row buffer in the kernel's order.  ok[b] = 1 where the frame is
// taken (every crop passes); other frames are left untouched.  Returns the number of buffer addresses two columns of ONE crop
// row wrote (a race between the lanes of a warp on the GPU); stats[0] = inverse-row-map slots written twice (two crop rows of
// one crop on one frame row: the assumption the kernel rests on), stats[1] = frames taken.
long long emu_kframe_gx(const float *theta, float mask01, const float *gy, float *gx, int *ok,
                        int n, int k, int c, int h, int w, int oh, int ow, long long *stats)
{
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    const int npx = oh * ow, frames = n / k;
    const size_t plane = (size_t)h * w;
    long long conflicts = 0;
    stats[0] = stats[1] = 0;
    std::vector<KfCrop> crops(k);
    std::vector<BandAxis> coltab((size_t)k * ow), rowtab((size_t)k * oh);
    std::vector<int> rowinv((size_t)k * h);
    std::vector<float> buf((size_t)c * w);
    std::vector<int> stamp((size_t)c * w);
    int stamp_id = 0;
    for (int b = 0; b < frames; ++b) {
        bool all = true;
        for (int kk = 0; kk < k; ++kk) {
            crops[kk] = make_kf_crop(load_theta_masked(theta + 6 * ((size_t)b * k + kk), mask01), h, w, oh, ow);
            all = all && crops[kk].ok;
        }
        ok[b] = all;
        if (!all) continue;
        stats[1]++;
        std::fill(rowinv.begin(), rowinv.end(), -1);
        for (int kk = 0; kk < k; ++kk) {
            const KfCrop &cr = crops[kk];
            for (int j = 0; j < ow; ++j) coltab[(size_t)kk * ow + j] = make_band_axis(cr.t00, cr.t01, cr.t02, linspace_pm1(j, ow, xstep), true, w);
            for (int i = 0; i < oh; ++i) {
                const BandAxis a = make_band_axis(cr.t11, cr.t10, cr.t12, linspace_pm1(i, oh, ystep), false, h);
                rowtab[(size_t)kk * oh + i] = a;
                const int t0 = (a.code & kAxIdxMask) - 1;
                if ((a.code & kAxTap0) && t0 >= 0 && t0 < h) { if (rowinv[(size_t)kk * h + t0] >= 0) stats[0]++; rowinv[(size_t)kk * h + t0] = 2 * i; }
                if ((a.code & kAxTap1) && t0 + 1 >= 0 && t0 + 1 < h) { if (rowinv[(size_t)kk * h + t0 + 1] >= 0) stats[0]++; rowinv[(size_t)kk * h + t0 + 1] = 2 * i + 1; }
            }
        }
        float *gxb = gx + (size_t)b * c * plane;
        for (int r = 0; r < h; ++r) {
            std::fill(buf.begin(), buf.end(), 0.f);
            for (int kk = 0; kk < k; ++kk) {
                const int cur = rowinv[(size_t)kk * h + r];
                if (cur < 0) continue;
                const int i = cur >> 1;
                const BandAxis &rw = rowtab[(size_t)kk * oh + i];
                const float wv = (cur & 1) ? rw.w0 : kf_w1(rw.w0);
                const float *gyc = gy + ((size_t)b * k + kk) * c * npx + (size_t)i * ow;
                ++stamp_id;
                for (int j = 0; j < ow; ++j) {
                    const BandAxis &col = coltab[(size_t)kk * ow + j];
                    const int u = (col.code & kAxIdxMask) - 1;
                    for (int ch = 0; ch < c; ++ch) {
                        const float g = gyc[(size_t)ch * npx + j];
                        if (col.code & kAxTap0) {
                            const int o = ch * w + u;
                            if (stamp[o] == stamp_id) ++conflicts;
                            stamp[o] = stamp_id;
                            buf[o] = f_add(buf[o], f_mul(f_mul(g, kf_w1(col.w0)), wv));
                        }
                        if (col.code & kAxTap1) {
                            const int o = ch * w + u + 1;
                            if (stamp[o] == stamp_id) ++conflicts;
                            stamp[o] = stamp_id;
                            buf[o] = f_add(buf[o], f_mul(f_mul(g, col.w0), wv));
                        }
                    }
                }
            }
            for (int ch = 0; ch < c; ++ch)
                for (int col = 0; col < w; ++col) gxb[(size_t)ch * plane + (size_t)r * w + col] = buf[(size_t)ch * w + col];
        }
    }
    return conflicts;
}

}  // extern "C"
