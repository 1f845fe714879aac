// stn_band.cu -- the backward LoANs actually runs, in one pass over the crop pixels.
//
// LoANs calls rotation_dropout(..., ratio=0.0) in front of the grid (reference sheep/sheep_localizer.py:61,169), so
// mask01 == 0 and every crop is an upright box: u depends on the crop column only, v on the crop row only.  For that
// case (one crop per frame, the only thing Chainer's sampler has) the gradient of the frame splits by crop rows, and
// a thread-block CLUSTER per crop does the whole backward with every crop pixel evaluated exactly ONCE:
//
//   * a CTA takes a run of crop rows and works through it in BANDS (stn_band_plan.cuh).  A band owns the frame rows
//     between the first row it touches and the first row the next band touches: the bands of a crop partition the
//     frame, so gx is written exactly once, zeros included -- no memset pass, no atomics, bit-reproducible;
//   * per crop pixel: 4 taps x C channels + gy -> d/du, d/dv (bit-exact with the oracle) accumulated into the six
//     gtheta sums, and gy * wu * wv added into the band's shared-memory tile of frame rows.  Column / row
//     coordinate chains come from BandAxis tables built once per CTA (oW + rows entries instead of one chain per
//     pixel).  When the crop steps by less than ~2 frame pixels the adds run in P * Q conflict-free phases;
//   * the tile rows leave shared memory as TMA bulk copies (cp.async.bulk.global.shared::cta, SASS UBLKCP), whole
//     frame rows at a time; the untouched frame rows between, above and below them are bulk-copied from a zero
//     plane, issued BEFORE the pixel work so the TMA unit streams them out while the SM computes;
//   * gtheta: warp shuffles + shared memory, then cluster rank 0 adds the per-CTA partials over distributed shared
//     memory in rank order (reduce_gtheta, stn_theta_role.cuh).
//
// Crops the plan declines (mirrored boxes, up-sampling by more than kBandMaxPhases, non-finite theta) are handled in
// the same launch by the same cluster running the general roles (stn_gx_role.cuh, stn_theta_role.cuh).
#include <cooperative_groups.h>

#include "stn_band_plan.cuh"
#include "stn_common.cuh"
#include "stn_gx_role.cuh"
#include "stn_theta_role.cuh"

namespace stn {


#ifdef STN_BAND_TRACE
// debug build only: per-CTA timestamps (globaltimer ns) at the stages of the band kernel, 16 slots per CTA
__device__ long long *g_band_trace = nullptr;
__device__ __forceinline__ void trace(int slot)
{
    if (threadIdx.x == 0 && g_band_trace) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_band_trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}
#define TRACE(k) trace(k)
#else
#define TRACE(k)
#endif

__device__ __forceinline__ void bulk_s2g(float *dst, const float *src_smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst),
                 "r"((uint32_t)__cvta_generic_to_shared(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// A crop the band plan declines: the cluster runs the general roles on it (tiles over the whole frame, then gtheta).
template <typename GT, int CG, bool GRAY>
__device__ __noinline__ void band_declined(const CropParams &p, const Theta &th, unsigned char *smem_raw, BwdSmem &sm,
                                           float *xs, float *ys, ScatterGeom *geom, int n, int rank)
{
    fill_axis_tables(p, xs, ys);
    int fb = 0;
    if (threadIdx.x == 0) {
        geom[0] = make_scatter_geom(th, p.H, p.W, p.oH, p.oW);
        fb = geom[0].P == 0;
    }
    const int any_fb = __syncthreads_or(fb);
    gx_role<GT, CG, true, GRAY>(p, nullptr, xs, ys, any_fb != 0, geom, reinterpret_cast<float *>(smem_raw), nullptr, n, rank,
                          p.band_fb_tiles_per_warp);
    theta_role<GT, CG, true, GRAY>(p, xs, ys, sm, (int)blockIdx.x);
}

// ILP = crop pixels in flight per thread, MINB = CTAs per SM the register budget is set for
// ROWBAND: every WARP works through bands of ONE crop row on a private two-row tile, synchronising with __syncwarp() only
// (no CTA barrier between the prologue and the gtheta reduction); otherwise the CTA works through multi-row bands together.
template <typename GT, int CG, int ILP, int MINB, bool GRAY, bool ROWBAND = false>
__global__ void __launch_bounds__(kThreads, MINB) stn_bwd_band_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [tile | zero plane] (the general roles' warp tiles alias it) [BwdSmem] [coltab] [rowtab] [rowslot] [xs|ys] [geom]
    float *tile = reinterpret_cast<float *>(smem_raw);
    float *zero_plane = reinterpret_cast<float *>(smem_raw + p.band_tile_bytes);
    unsigned char *q = smem_raw + p.band_region_bytes;
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(q);
    q += sizeof(BwdSmem);
    BandAxis *coltab = reinterpret_cast<BandAxis *>(q);
    q += sizeof(BandAxis) * p.oW;
    BandAxis *rowtab = reinterpret_cast<BandAxis *>(q);
    q += sizeof(BandAxis) * p.band_tab_rows;
    int2 *rowslot = reinterpret_cast<int2 *>(q);
    q += sizeof(int2) * p.band_tab_rows;
    float *xs = reinterpret_cast<float *>(q);
    float *ys = xs + p.oW;
    q += sizeof(float) * ((p.oW + p.oH + 3) & ~3);
    ScatterGeom *geom = reinterpret_cast<ScatterGeom *>(q);

    TRACE(0);
    pdl_launch_dependents();
    {   // the zero plane needs no input: filled while the previous kernel of the stream drains
        float4 *z4 = reinterpret_cast<float4 *>(zero_plane);
        for (int e = threadIdx.x; e < p.band_zero_bytes / 16; e += kThreads) z4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_async_smem();
    }
    const int cs = p.ctas_per_crop;
    const int n = blockIdx.x / cs, rank = blockIdx.x - n * cs;
    const int tid = threadIdx.x;
#ifndef STN_BAND_PULL_REDUCE
    push_reduce_init(sm, rank, cs);                                   // shared memory only: before the wait
#endif
    pdl_prefetch_theta(p, n);
#ifdef STN_NO_PDL_PREFETCH
    pdl_wait();
#endif
    {   // this CTA's gy rows do not depend on theta: started towards L2 before the wait (prefetches only, see pdl_prefetch_theta)
        const int r0 = rank * p.band_rows_cta, r1 = min(p.oH, r0 + p.band_rows_cta);
        const int row_elems = max(r1 - r0, 0) * p.oW;
        const int lines = (row_elems * (int)sizeof(GT) + 127) / 128;
        constexpr int gplanes = crop_planes<GT, GRAY>(CG);
        for (int e = tid; e < lines * gplanes; e += kThreads) {
            const int ch = e / lines, l = e - ch * lines;
            const char *a = reinterpret_cast<const char *>(reinterpret_cast<const GT *>(p.gy) + ((size_t)n * gplanes + ch) * p.oH * p.oW +
                                                           (size_t)r0 * p.oW) + (size_t)l * 128;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
        }
    }
#ifndef STN_NO_PDL_PREFETCH
    pdl_wait();
#endif
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    TRACE(1);
    const int H = p.H, W = p.W, oH = p.oH, oW = p.oW;
    const int i0 = rank * p.band_rows_cta, i1 = min(oH, i0 + p.band_rows_cta);
    const int t0 = max(i0 - kBandMaxHalo, 0), t1 = min(i1 + 1, oH);   // rows tabulated: the longest halo the plan can ask for
    // ---- prologue: the crop's verdict (one warp), coordinate chains once per crop column / row (the others)
    if (tid >= kThreads - 32) {
        const BandCrop v = make_band_crop(th, H, W, oH, oW);              // the same in every CTA of the cluster
        if (tid == kThreads - 32) { sm.bc[0] = v.ok; sm.bc[1] = v.P; sm.bc[2] = v.Q; }
        if (tid >= kThreads - 2) {                                        // first and last frame row the crop touches: L, E
            const int last = tid == kThreads - 1;
            int lo, hi;
            band_row_range(make_band_axis(th.t11, th.t10, th.t12, lin_y_at(p, last ? oH - 1 : 0), false, H), H, lo, hi);
            sm.flags[last] = last ? hi : lo;
        }
    } else {
        for (int k = tid; k < oW + (t1 - t0); k += kThreads - 32) {
            if (k < oW) coltab[k] = make_band_axis(th.t00, th.t01, th.t02, lin_x_at(p, k), true, W);
            else rowtab[k - oW] = make_band_axis(th.t11, th.t10, th.t12, lin_y_at(p, t0 + k - oW), false, H);
        }
    }
    __syncthreads();
    TRACE(2);
    if (!sm.bc[0]) {
        band_declined<GT, CG, GRAY>(p, th, smem_raw, sm, xs, ys, geom, n, rank);
        return;
    }
    const int P = sm.bc[1], Q = sm.bc[2];
    if (!ROWBAND && !(p.band_flags & 2) && i0 < i1) {
        // every frame row segment this CTA will gather from, requested into L2 now, in one go, before any CTA has started to
        // store: a load that queues behind a burst of bulk stores waits for the whole burst (per-CTA timestamps, profiles/README.md)
        const int ua = (coltab[0].code & kAxIdxMask) - 1, ub = (coltab[oW - 1].code & kAxIdxMask);
        const int c_lo = max(ua, 0) & ~31, c_hi = min(ub, W - 1);                  // columns, 128-byte lines
        const int lines = c_hi >= c_lo ? (c_hi - c_lo) / 32 + 1 : 0;
        const int h0 = max(i0 - (P - 1), 0);                                        // first row evaluated (halo included)
        const int nrow2 = 2 * (i1 - h0);                                            // two tap rows per crop row
        const float *xn = p.x + (size_t)n * CG * ((size_t)H * W);
        for (int e = tid; e < nrow2 * lines * CG; e += kThreads) {
            const int l = e % lines, rc = e / lines;
            const int ch = rc % CG, r2 = rc / CG;
            const int fr = (rowtab[h0 - t0 + (r2 >> 1)].code & kAxIdxMask) - 1 + (r2 & 1);
            if (fr >= 0 && fr < H) {
                const float *a = xn + (size_t)ch * H * W + (size_t)fr * W + c_lo + 32 * l;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a));
            }
        }
    }

    const int npx = oH * oW;
    const int plane = H * W;
    const size_t fpx = (size_t)plane;
    const float *xb = p.x + (size_t)n * CG * fpx;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * crop_planes<GT, GRAY>(CG) * npx;
    float *gxb = p.gx + (size_t)n * CG * fpx;
    float *ggo = p.ggrid_out ? p.ggrid_out + (size_t)n * 2 * npx : nullptr;
    const float *ggu = p.ggrid_up ? p.ggrid_up + (size_t)n * 2 * npx : nullptr;
    const int tile_plane = p.band_cap * W;                            // floats per channel of the tile
    const int zrows = p.band_zero_bytes / (W * 4);
    const float inv_ow = 1.0f / (float)oW;
    const int nphases = P * Q;

    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (ROWBAND) {
        // ---- one crop row per band, one warp per band.  The band of crop row a owns the frame rows from the first one row a
        // touches to the first one row a + 1 touches; at most two of them are touched (the rest is zero), so the warp's tile is
        // two frame rows per channel.  With P > 1 the P - 1 crop rows above are re-evaluated for the taps that land in those rows.
        const int wrp = tid >> 5, ln = tid & 31;
        float *wtile = tile + wrp * (CG * 2 * W);
        const int wplane = 2 * W;
        bool wpending = false;
        for (int a = i0 + wrp; a < i1; a += kWarps) {
            const BandPlan pl = plan_band(rowtab, t0, a, i1, oH, H, P, 2, 1, sm.flags[1]);
            if (wpending) {                                          // this warp's previous tile has been read by the TMA unit
                bulk_wait_read();
                __syncwarp();
            }
            bool need_zero = true;                                   // the tile is zeroed while the first loads are in flight
            TRACE(4);
            for (int row = pl.h; row < pl.b; ++row) {
                const BandAxis rw = rowtab[row - t0];
                int s0, s1;
                band_row_slots(pl, rw, row, s0, s1);
                const bool own = row >= pl.a;
                if (!own && s0 < 0 && s1 < 0) continue;                // a halo row that reaches none of this band's frame rows
                const int v0 = rw.code & kAxIdxMask;
                // ILP crop pixels per lane in flight: the taps and gy of all of them are requested before any is reduced
                struct RPx {
                    float v[CG][4], g[CG];
                    int j;
                    bool live;
                };
                auto rprepare = [&](RPx &px, int j) {
                    px.j = j;
                    px.live = j < oW;
                    if (!px.live) return;
                    const int u0 = coltab[j].code & kAxIdxMask;
                    TapAddr ta;
                    ta.c0 = u0 >= 1; ta.c1 = u0 <= W - 1; ta.r0 = v0 >= 1; ta.r1 = v0 <= H - 1;
                    ta.o00 = (v0 - 1) * W + (u0 - 1);
                    const GT *gp = gyb + row * oW + j;
#pragma unroll
                    for (int ch = 0; ch < CG; ++ch) {
#ifndef STN_BAND_GXONLY                                  // A/B build (profiles/overlap_probe.py): gx alone, no taps, no gtheta
                        if (own) load_taps(xb + ch * plane, ta, W, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3]);
#endif
                        px.g[ch] = load_gy<GT, GRAY>(gp, ch, npx);
                    }
                };
                auto rreduce = [&](const RPx &px) {                    // gtheta sums and ggrid: the band's own row only
                    if (!px.live || !own) return;
                    const BandAxis col = coltab[px.j];
                    const Tap t = tap_from_band_axes(col, rw, H, W);
                    float su = 0.f, sv = 0.f;
#pragma unroll
                    for (int ch = 0; ch < CG; ++ch) {
                        float gu, gv;
                        grad_uv(t, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3], gu, gv);
                        gu = f_mul(gu, px.g[ch]);
                        gv = f_mul(gv, px.g[ch]);
                        if (ch == 0) { su = gu; sv = gv; }
                        else { su = f_add(su, gu); sv = f_add(sv, gv); }
                    }
                    finish_grad_uv(t, H, W, su, sv);
                    const int qq = row * oW + px.j;
                    if (ggo) {
                        ggo[qq] = su;
                        ggo[npx + qq] = sv;
                    }
                    if (ggu) {
                        su = f_add(su, __ldg(ggu + qq));
                        sv = f_add(sv, __ldg(ggu + npx + qq));
                    }
                    s[0] = fmaf(su, col.lin, s[0]); s[1] = fmaf(su, rw.lin, s[1]); s[2] += su;
                    s[3] = fmaf(sv, col.lin, s[3]); s[4] = fmaf(sv, rw.lin, s[4]); s[5] += sv;
                };
                auto rscatter = [&](const RPx &px) {                   // gy * wu * wv into the warp's tile
                    const BandAxis col = coltab[px.j];
                    const bool c0 = (col.code & kAxTap0) != 0, c1 = (col.code & kAxTap1) != 0;
                    float *tp = wtile + ((col.code & kAxIdxMask) - 1);
                    float *r0p = tp + s0 * W, *r1p = tp + s1 * W;
#pragma unroll
                    for (int ch = 0; ch < CG; ++ch) {
                        const float a1 = f_mul(px.g[ch], col.w1), a0 = f_mul(px.g[ch], col.w0);
                        if (s0 >= 0) {
                            if (c0) r0p[ch * wplane] = f_add(r0p[ch * wplane], f_mul(a1, rw.w1));
                            if (c1) r0p[ch * wplane + 1] = f_add(r0p[ch * wplane + 1], f_mul(a0, rw.w1));
                        }
                        if (s1 >= 0) {
                            if (c0) r1p[ch * wplane] = f_add(r1p[ch * wplane], f_mul(a1, rw.w0));
                            if (c1) r1p[ch * wplane + 1] = f_add(r1p[ch * wplane + 1], f_mul(a0, rw.w0));
                        }
                    }
                };
                const bool touches = s0 >= 0 || s1 >= 0;
                for (int j0 = 0; j0 < oW; j0 += 32 * ILP) {
                    RPx px[ILP];
#pragma unroll
                    for (int u = 0; u < ILP; ++u) rprepare(px[u], j0 + 32 * u + ln);
                    if (need_zero) {
                        float4 *t4 = reinterpret_cast<float4 *>(wtile);
                        for (int e = ln; e < CG * wplane / 4; e += 32) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                        __syncwarp();
                        need_zero = false;
                    }
                    TRACE(5);
#pragma unroll
#ifndef STN_BAND_GXONLY
                    for (int u = 0; u < ILP; ++u) rreduce(px[u]);
#endif
                    TRACE(6);
                    // crop pixels of one column phase never share a frame pixel (lanes of one chunk are 32 columns apart from the next)
#pragma unroll
                    for (int u = 0; u < ILP; ++u) {
                        for (int cq = 0; cq < Q; ++cq) {
                            if (px[u].live && touches && (Q == 1 || px[u].j % Q == cq)) rscatter(px[u]);
                            if (Q > 1) __syncwarp();
                        }
                        if (ILP > 1) __syncwarp();                      // chunk boundary: neighbouring columns meet at lanes 31 | 0
                    }
                }
                __syncwarp();                                          // rows are worked through one after the other
            }
            // tile rows and the zero rows around them -> gx, one bulk copy per lane
            TRACE(7);
            fence_async_smem();
            __syncwarp();
            {
                const int nspans = band_span_count(pl);
                if (ln < nspans * CG) {
                    const int t = ln / CG, ch = ln - t * CG;
                    const BandSpan sp = band_span(pl, rowtab, t0, H, t);
                    float *dst = gxb + (size_t)ch * fpx + (size_t)sp.row * W;
                    if (sp.slot >= 0) {
                        if (sp.nrows > 0) bulk_s2g(dst, wtile + ch * wplane + sp.slot * W, (uint32_t)(sp.nrows * W * 4));
                    } else {
                        for (int r = 0; r < sp.nrows; r += zrows)
                            bulk_s2g(dst + (size_t)r * W, zero_plane, (uint32_t)(min(zrows, sp.nrows - r) * W * 4));
                    }
                }
                bulk_commit();
            }
            wpending = true;
            TRACE(9);
        }
    }
    const int warp = tid >> 5, lane = tid & 31;
    bool pending = false;                                             // tile stores of the previous band in flight
    struct Px {
        float v[CG][4], g[CG];
        int i, j;
        bool live;
    };
    for (int a = i0; !ROWBAND && a < i1;) {
        const BandPlan pl = plan_band(rowtab, t0, a, i1, oH, H, P, p.band_cap, p.band_rows, sm.flags[1]);
        const int nspans = band_span_count(pl);
        const int npxb = (pl.b - pl.h) * oW;
        auto prepare = [&](Px &px, int e) {
            px.live = e < npxb;
            if (!px.live) return;
            const int ri = __float2int_rz(((float)e + 0.5f) * inv_ow);
            px.i = pl.h + ri;
            px.j = e - ri * oW;
            const int u0 = coltab[px.j].code & kAxIdxMask, v0 = rowtab[px.i - t0].code & kAxIdxMask;
            TapAddr ta;
            ta.c0 = u0 >= 1; ta.c1 = u0 <= W - 1; ta.r0 = v0 >= 1; ta.r1 = v0 <= H - 1;
            ta.o00 = (v0 - 1) * W + (u0 - 1);
            const GT *gp = gyb + px.i * oW + px.j;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                load_taps(xb + ch * plane, ta, W, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3]);
                px.g[ch] = load_gy<GT, GRAY>(gp, ch, npx);
            }
        };
        auto reduce = [&](const Px &px) {                             // gtheta sums and ggrid: own rows only
            if (!px.live || px.i < pl.a) return;
            const BandAxis col = coltab[px.j], row = rowtab[px.i - t0];
            const Tap t = tap_from_band_axes(col, row, H, W);
            float su = 0.f, sv = 0.f;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                float gu, gv;
                grad_uv(t, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3], gu, gv);
                gu = f_mul(gu, px.g[ch]);
                gv = f_mul(gv, px.g[ch]);
                if (ch == 0) { su = gu; sv = gv; }
                else { su = f_add(su, gu); sv = f_add(sv, gv); }      // numpy.sum over the channel axis
            }
            finish_grad_uv(t, H, W, su, sv);
            const int qq = px.i * oW + px.j;
            if (ggo) {
                ggo[qq] = su;
                ggo[npx + qq] = sv;
            }
            if (ggu) {
                su = f_add(su, __ldg(ggu + qq));
                sv = f_add(sv, __ldg(ggu + npx + qq));
            }
            s[0] = fmaf(su, col.lin, s[0]); s[1] = fmaf(su, row.lin, s[1]); s[2] += su;
            s[3] = fmaf(sv, col.lin, s[3]); s[4] = fmaf(sv, row.lin, s[4]); s[5] += sv;
        };
        auto scatter = [&](const Px &px) {                            // gy * wu * wv into the tile, reference order
            const int2 sl = rowslot[px.i - pl.h];
            if (sl.x < 0 && sl.y < 0) return;
            const BandAxis col = coltab[px.j], row = rowtab[px.i - t0];
            const bool c0 = (col.code & kAxTap0) != 0, c1 = (col.code & kAxTap1) != 0;
            float *tp = tile + ((col.code & kAxIdxMask) - 1);
            float *r0p = tp + sl.x * W, *r1p = tp + sl.y * W;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                const float a1 = f_mul(px.g[ch], col.w1), a0 = f_mul(px.g[ch], col.w0);
                if (sl.x >= 0) {
                    if (c0) r0p[ch * tile_plane] = f_add(r0p[ch * tile_plane], f_mul(a1, row.w1));
                    if (c1) r0p[ch * tile_plane + 1] = f_add(r0p[ch * tile_plane + 1], f_mul(a0, row.w1));
                }
                if (sl.y >= 0) {
                    if (c0) r1p[ch * tile_plane] = f_add(r1p[ch * tile_plane], f_mul(a1, row.w0));
                    if (c1) r1p[ch * tile_plane + 1] = f_add(r1p[ch * tile_plane + 1], f_mul(a0, row.w0));
                }
            }
        };
        auto zero_spans = [&]() {
            for (int e = warp + kWarps * lane; e < ((nspans + 1) >> 1) * CG; e += kThreads) {
                const int k = e / CG, ch = e - k * CG;
                const BandSpan sp = band_span(pl, rowtab, t0, H, 2 * k);
                float *dst = gxb + (size_t)ch * fpx + (size_t)sp.row * W;
                for (int r = 0; r < sp.nrows; r += zrows)
                    bulk_s2g(dst + (size_t)r * W, zero_plane, (uint32_t)(min(zrows, sp.nrows - r) * W * 4));
            }
        };
        // (1) the first ILP crop pixels per thread: taps and gy requested before anything else (the band's latency)
        Px px[ILP];
#pragma unroll
        for (int u = 0; u < ILP; ++u) prepare(px[u], u * kThreads + tid);
        TRACE(3);
        if (p.band_flags & 1) zero_spans();
        if (pending) {                                                // the tile is free once the TMA unit has read it
            bulk_wait_read();
            __syncthreads();
        }
        // (3) zero the tile rows in use, tile slots of the crop rows
        {
            const int row4 = W >> 2, n4 = pl.nslots * row4;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                float4 *t4 = reinterpret_cast<float4 *>(tile + ch * tile_plane);
                for (int e = tid; e < n4; e += kThreads) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int e = tid; e < pl.b - pl.h; e += kThreads) {
                int s0, s1;
                band_row_slots(pl, rowtab[pl.h + e - t0], pl.h + e, s0, s1);
                rowslot[e] = make_int2(s0, s1);
            }
        }
        __syncthreads();
        TRACE(4);
        // (4) the crop pixels of rows [h, b), ILP per thread in flight
        for (int e0 = 0; e0 < npxb; e0 += ILP * kThreads) {
            if (e0 > 0) {
#pragma unroll
                for (int u = 0; u < ILP; ++u) prepare(px[u], e0 + u * kThreads + tid);
            }
            TRACE(5);
#pragma unroll
            for (int u = 0; u < ILP; ++u) reduce(px[u]);
            TRACE(6);
            if (nphases == 1) {
#pragma unroll
                for (int u = 0; u < ILP; ++u)
                    if (px[u].live) scatter(px[u]);
            } else {
                // crop pixels of one phase never share a frame pixel; phases are separated by CTA barriers
                for (int ph = 0; ph < nphases; ++ph) {
#pragma unroll
                    for (int u = 0; u < ILP; ++u)
                        if (px[u].live && (px[u].i % P) * Q + (px[u].j % Q) == ph) scatter(px[u]);
                    __syncthreads();
                }
            }
        }
        // (5) tile rows -> gx.  A bulk copy is issued per lane (the warp serialises them): the ops are dealt to the warps first
        TRACE(7);
        fence_async_smem();
        __syncthreads();
        TRACE(8);
        for (int e = warp + kWarps * lane; e < (nspans >> 1) * CG; e += kThreads) {
            const int k = e / CG, ch = e - k * CG;
            const BandSpan sp = band_span(pl, rowtab, t0, H, 2 * k + 1);
            if (sp.nrows > 0)
                bulk_s2g(gxb + (size_t)ch * fpx + (size_t)sp.row * W, tile + ch * tile_plane + sp.slot * W,
                         (uint32_t)(sp.nrows * W * 4));
        }
        // (6) the all-zero frame rows this band owns, from the zero plane.  After the tile rows by default: taps requested
        //     behind a burst of bulk stores wait for the whole burst to drain
        if (!(p.band_flags & 1)) zero_spans();
        bulk_commit();
        pending = true;
        a = pl.b;
        TRACE(9);
    }
    {   // this CTA's share of the all-zero rows above and below the crop
        int ra, na, rb, nb;
        band_edge_rows(sm.flags[0], sm.flags[1], H, rank, cs, ra, na, rb, nb);
        const int ca = (na + zrows - 1) / zrows, cb = (nb + zrows - 1) / zrows;
        for (int e = warp + kWarps * lane; e < (ca + cb) * CG; e += kThreads) {
            const int k = e / CG, ch = e - k * CG;
            const int row = k < ca ? ra + k * zrows : rb + (k - ca) * zrows;
            const int nr = k < ca ? min(zrows, na - k * zrows) : min(zrows, nb - (k - ca) * zrows);
            bulk_s2g(gxb + (size_t)ch * fpx + (size_t)row * W, zero_plane, (uint32_t)(nr * W * 4));
        }
        bulk_commit();
    }
#if defined(STN_BAND_GXONLY)
    (void)s;
#elif !defined(STN_BAND_PULL_REDUCE)
    reduce_gtheta_push(p, s, sm, n, rank, cs);
#else
    reduce_gtheta(p, s, sm, n, rank, cs);
#endif
    TRACE(10);
    bulk_wait_read();
    TRACE(11);
#ifdef STN_BAND_TRACE
    if (threadIdx.x == 0 && g_band_trace) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        g_band_trace[(size_t)blockIdx.x * 16 + 15] = smid;
    }
#endif                                                 // shared memory stays valid until the TMA unit has read it
}

// ------------------------------------------------------------------------------------------ frames without grad
// gx == NULL (every LoANs call: the frames are a raw array) and rotation terms masked: the theta gradient alone, with the
// coordinate chains taken from per-crop BandAxis tables (oW + rows entries per CTA) instead of being re-derived per pixel --
// the per-pixel work is two 16-byte table reads, 4 taps x C + gy, and the exact d/du, d/dv arithmetic.  Any number of crops
// per frame, any sign of the scales (nothing is scattered, so nothing can collide).  Crops whose rotation terms are not
// (+-)0 after masking run the general theta role in the same launch.
template <typename GT, int CG, bool GRAY>
__global__ void __launch_bounds__(kThreads, 4) stn_bwd_theta_tab_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    BandAxis *coltab = reinterpret_cast<BandAxis *>(smem_raw + sizeof(BwdSmem));
    BandAxis *rowtab = coltab + p.oW;
    float *xs = reinterpret_cast<float *>(rowtab + p.band_rows_cta);
    float *ys = xs + p.oW;
    const int cs = p.ctas_per_crop;
    const int n = blockIdx.x / cs, rank = blockIdx.x - n * cs;
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    push_reduce_init(sm, rank, cs);
    pdl_prefetch_theta(p, n);
    pdl_wait();
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    const float mag = fabsf(th.t00) + fabsf(th.t11) + fabsf(th.t02) + fabsf(th.t12);
    if (!(th.t01 == 0.0f && th.t10 == 0.0f && mag < 1e30f)) {           // rotated (or non-finite) crop: per-pixel chains
        fill_axis_tables(p, xs, ys);
        __syncthreads();
        theta_role<GT, CG, true, GRAY>(p, xs, ys, sm, (int)blockIdx.x);
        return;
    }
    const int H = p.H, W = p.W, oH = p.oH, oW = p.oW;
    const int i0 = rank * p.band_rows_cta, i1 = min(oH, i0 + p.band_rows_cta);
    for (int k = tid; k < oW + max(i1 - i0, 0); k += kThreads) {
        if (k < oW) coltab[k] = make_band_axis(th.t00, th.t01, th.t02, lin_x_at(p, k), true, W);
        else rowtab[k - oW] = make_band_axis(th.t11, th.t10, th.t12, lin_y_at(p, i0 + k - oW), false, H);
    }
    __syncthreads();
    const int npx = oH * oW, plane = H * W;
    const float *xb = p.x + (size_t)(n / p.K) * CG * plane;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * crop_planes<GT, GRAY>(CG) * npx;
    float *ggo = p.ggrid_out ? p.ggrid_out + (size_t)n * 2 * npx : nullptr;
    const float *ggu = p.ggrid_up ? p.ggrid_up + (size_t)n * 2 * npx : nullptr;
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int q_end = i1 * oW;
    for (PxWalk w(i0 * oW + tid, oW); w.q < q_end; w.next()) {
        const BandAxis col = coltab[w.j], row = rowtab[w.i - i0];
        const Tap t = tap_from_band_axes(col, row, H, W);
        const TapAddr a = make_tap_addr(t, H, W);
        float v[CG][4], g[CG];
        const GT *gp = gyb + w.q;
#pragma unroll
        for (int ch = 0; ch < CG; ++ch) {
            load_taps(xb + ch * plane, a, W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
            g[ch] = load_gy<GT, GRAY>(gp, ch, npx);
        }
        float su = 0.f, sv = 0.f;
#pragma unroll
        for (int ch = 0; ch < CG; ++ch) {
            float gu, gv;
            grad_uv(t, v[ch][0], v[ch][1], v[ch][2], v[ch][3], gu, gv);
            gu = f_mul(gu, g[ch]);
            gv = f_mul(gv, g[ch]);
            if (ch == 0) { su = gu; sv = gv; }
            else { su = f_add(su, gu); sv = f_add(sv, gv); }                  // numpy.sum over the channel axis
        }
        finish_grad_uv(t, H, W, su, sv);
        if (ggo) {
            ggo[w.q] = su;
            ggo[npx + w.q] = sv;
        }
        if (ggu) {
            su = f_add(su, __ldg(ggu + w.q));
            sv = f_add(sv, __ldg(ggu + npx + w.q));
        }
        s[0] = fmaf(su, col.lin, s[0]); s[1] = fmaf(su, row.lin, s[1]); s[2] += su;
        s[3] = fmaf(sv, col.lin, s[3]); s[4] = fmaf(sv, row.lin, s[4]); s[5] += sv;
    }
    reduce_gtheta_push(p, s, sm, n, rank, cs);
}

template <typename GT, int CG, bool GRAY>
static cudaError_t launch_theta_tab_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, cs);
    return cudaLaunchKernelEx(&cfg, stn_bwd_theta_tab_kernel<GT, CG, GRAY>, p);
}

// Returns -1 when the call is not one this kernel takes (gx wanted, rotation not masked, channel count not 1 / 3 / 4).
int launch_crop_bwd_theta_tab(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.gx) return -1;
    if (p.C != 1 && p.C != 3 && p.C != 4) return -1;
    if (p.H > kAxIdxMask - 2 || p.W > kAxIdxMask - 2) return -1;
    if ((long long)p.H * p.W * p.C > 0x7fffffffLL) return -1;
    // CTAs per crop: enough to fill the machine (as the general theta role), at most 8, at least 64 pixels per thread-block row
    const long long npx = (long long)p.oH * p.oW;
    unsigned cs = 1;
    while (cs < 8 && (long long)p.N * cs < 2LL * num_sms() && (int)(2 * cs) <= p.oH && npx / (2 * cs) >= kThreads / 2) cs *= 2;
    p.ctas_per_crop = (int)cs;
    p.px_per_cta = (int)((npx + cs - 1) / cs);                         // rotated crops: the general theta role's share
    p.band_rows_cta = (p.oH + (int)cs - 1) / (int)cs;
    const size_t smem = sizeof(BwdSmem) + sizeof(BandAxis) * (size_t)(p.oW + p.band_rows_cta) + sizeof(float) * (size_t)((p.oW + p.oH + 3) & ~3);
    if (smem > 48 * 1024) return -1;
    const long long ctas = (long long)p.N * cs;
    if (ctas > 0x7fffffffLL) return -1;
    cudaError_t e;
    if (p.nhwc)
        e = launch_theta_tab_tt<Nhwc4, 3, false>(p, (unsigned)ctas, cs, smem, stream);
    else if (gy_dtype == 0)
        e = p.C == 1 ? launch_theta_tab_tt<float, 1, false>(p, (unsigned)ctas, cs, smem, stream)
          : p.C == 4 ? launch_theta_tab_tt<float, 4, false>(p, (unsigned)ctas, cs, smem, stream)
          : p.gray   ? launch_theta_tab_tt<float, 3, true>(p, (unsigned)ctas, cs, smem, stream)
                     : launch_theta_tab_tt<float, 3, false>(p, (unsigned)ctas, cs, smem, stream);
    else
        e = p.C == 1 ? launch_theta_tab_tt<__nv_bfloat16, 1, false>(p, (unsigned)ctas, cs, smem, stream)
          : p.C == 4 ? launch_theta_tab_tt<__nv_bfloat16, 4, false>(p, (unsigned)ctas, cs, smem, stream)
          : p.gray   ? launch_theta_tab_tt<__nv_bfloat16, 3, true>(p, (unsigned)ctas, cs, smem, stream)
                     : launch_theta_tab_tt<__nv_bfloat16, 3, false>(p, (unsigned)ctas, cs, smem, stream);
    count_launch();
    note_kernel("stn_bwd_theta_tab_kernel");
    if (e != cudaSuccess) return set_error("crop_bwd (theta, tables) launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ------------------------------------------------------------------------------------------ host launcher
// tuning knobs (loans_stn_configure): 0 = automatic
static int g_band_cs = 0, g_band_rows = 0, g_band_tile_kb = 0, g_band_variant = 0;
void band_tuning(int which, int value)
{
    if (which == 0) g_band_cs = value;
    else if (which == 1) g_band_rows = value;
    else if (which == 2) g_band_tile_kb = value;
    else g_band_variant = value;
}

template <typename GT, int CG, int ILP, int MINB, bool GRAY = false, bool ROWBAND = false>
static cudaError_t launch_band_ttt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    {
        const cudaError_t e = grant_dynamic_smem(reinterpret_cast<const void *>(&stn_bwd_band_kernel<GT, CG, ILP, MINB, GRAY, ROWBAND>), smem);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, cs);
    return cudaLaunchKernelEx(&cfg, stn_bwd_band_kernel<GT, CG, ILP, MINB, GRAY, ROWBAND>, p);
}

// kind 1: CTA bands, one crop pixel in flight per thread, 64 registers, four CTAs per SM; kind 2: CTA bands, two pixels in
// flight, 80 registers, three CTAs per SM (kept for A/B runs: never faster); kind 3: row bands (one crop row per warp)
template <typename GT, int CG>
static cudaError_t launch_band_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s, int kind)
{
    if constexpr (CG == 3) {
        if (p.gray)                                                    // grayscale epilogue: its own kernel instances
            return kind == 3 ? launch_band_ttt<GT, 3, 1, 4, true, true>(p, ctas, cs, smem, s)
                             : launch_band_ttt<GT, 3, 1, 4, true, false>(p, ctas, cs, smem, s);
    }
    switch (kind) {
    case 3: return launch_band_ttt<GT, CG, 1, 4, false, true>(p, ctas, cs, smem, s);
#ifdef STN_DEVEL
    case 2: return launch_band_ttt<GT, CG, 2, 3>(p, ctas, cs, smem, s);
#endif
    default: return launch_band_ttt<GT, CG, 1, 4>(p, ctas, cs, smem, s);
    }
}

// Returns -1 when the shape is not one the band kernels take -- or, with by_measurement, not one where they measured faster
// than the general kernel (profiles/README.md) -- and the caller then launches the general kernel.
//   row bands: narrow frames (eight two-row tiles fit the shared-memory budget) and enough crops to fill the machine --
//     they win from 64 crops of 64x64 (16.1 vs 20.9 us; 30 vs 36 us at 128, 56 vs 67 us at 256) and from 256 crops of 75x75
//     (74 vs 83 us; 241 vs 262 us at 1024), lose below (12.4 vs 13.7 us at 32 crops of 64x64, 44 vs 48 us at 128 crops of
//     75x75): a warp works through its rows and 32-pixel chunks one memory round trip after the other;
//   CTA bands: frame rows of at least 4 KiB (512-px RGB frames: 192 vs 209-214 us at BASELINE config 3).
int launch_crop_bwd_band(CropParams p, int gy_dtype, cudaStream_t stream, bool by_measurement)
{
    if (!p.gx || p.K != 1) return -1;
    if (p.C != 1 && p.C != 3 && p.C != 4) return -1;
    if (p.W % 4 != 0 || (reinterpret_cast<uintptr_t>(p.gx) & 15) != 0) return -1;
    if (p.H > kAxIdxMask - 2 || p.W > kAxIdxMask - 2) return -1;
    if ((long long)p.H * p.W * p.C > 0x7fffffffLL) return -1;
    const size_t slot_bytes = sizeof(float) * (size_t)p.W * p.C;
    // shared memory: tile + zero plane within ~53 KB, so that four CTAs fit an SM next to their tables
    const size_t budget = (size_t)(g_band_tile_kb > 0 ? g_band_tile_kb : 53) * 1024;
    int zrows = (int)(2048 / (sizeof(float) * p.W));
    if (zrows < 1) zrows = 1;
    p.band_zero_bytes = (int)(sizeof(float) * p.W * zrows);
    if (budget < (size_t)p.band_zero_bytes + 4 * slot_bytes) return -1;   // frame rows too wide for the tile
    const int cap_max = (int)((budget - p.band_zero_bytes) / slot_bytes);
    // CTAs per crop (= cluster size): 8 while that leaves every CTA at least one crop row
    unsigned cs = g_band_cs > 0 ? (unsigned)g_band_cs : 8;
    while (cs > 1 && (int)cs > p.oH) cs >>= 1;
    // row bands: a CTA's eight warps take eight crop rows at a time -- the cluster size (any size up to 8) that leaves the
    // fewest warps idle in the last pass, the largest one among equals
    unsigned cs_row = cs;
    if (g_band_cs == 0) {
        double best = 0.0;
        for (unsigned c = 1; c <= 8 && (int)c <= p.oH; ++c) {
            const int rows = (p.oH + (int)c - 1) / (int)c;
            const double util = (double)p.oH / ((double)c * kWarps * ((rows + kWarps - 1) / kWarps));
            if (util >= best * 0.999) { if (util > best) best = util; cs_row = c; }
        }
    }
    const bool fits_row = budget >= (size_t)p.band_zero_bytes + 2 * kWarps * slot_bytes;
    const int knob = g_band_variant & 15;                               // 0: automatic; 1, 2: CTA bands; 3: row bands
    bool rowband;
    if (knob == 3) {
        if (!fits_row) return -1;
        rowband = true;
    } else if (knob == 1 || knob == 2) {
        rowband = false;
    } else {
        rowband = fits_row;
        if (by_measurement) {
            if (p.gray) rowband = false;             // behind the grayscale epilogue the general kernel measured faster (295 vs 336 us at cfg5)
            if (rowband) {
                const int passes = (((p.oH + (int)cs_row - 1) / (int)cs_row) + kWarps - 1) / kWarps, chunks = (p.oW + 31) / 32;
                if ((long long)p.N * cs_row < 200LL * passes * chunks) rowband = false;
            }
            if (!rowband && slot_bytes < 4096) return -1;
        }
    }
    if (rowband && g_band_cs == 0) {
        // among the cluster sizes that keep the warps equally busy, the smallest one (>= 2) that still launches ~1.7 waves of
        // CTAs (4 x 148 resident): fewer CTAs per crop amortise the per-CTA prologue and reduction.  Measured, 75x75 crops:
        // 125.5 vs 127.3 us at 512 crops and 228.0 vs 238.7 us at 1024 (2 vs 5 CTAs per crop; a tie at 256); 64x64 crops:
        // 49.1 vs 53.0 us at 256 (4 vs 8), 91.9 vs 99.6 us at 512 (2 vs 8); one CTA per crop loses everywhere (267 us at 1024)
        const int rows_big = (p.oH + (int)cs_row - 1) / (int)cs_row;
        const double util_big = (double)p.oH / ((double)cs_row * kWarps * ((rows_big + kWarps - 1) / kWarps));
        for (unsigned c = 2; c < cs_row; ++c) {
            const int rows = (p.oH + (int)c - 1) / (int)c;
            const double util = (double)p.oH / ((double)c * kWarps * ((rows + kWarps - 1) / kWarps));
            if (util >= util_big * 0.999 && (long long)p.N * c >= 1024) { cs_row = c; break; }
        }
    }
    if (rowband) cs = cs_row;
    p.ctas_per_crop = (int)cs;
    p.px_per_cta = (int)(((long long)p.oH * p.oW + cs - 1) / cs);     // declined crops: theta role share
    p.band_rows_cta = (p.oH + (int)cs - 1) / (int)cs;
    // crop rows per band: a band of r rows needs 2r tile rows when the crop steps by >= ~2 frame rows (compact tile), and up
    // to ceil(2.2 (r - 1)) + 3 when it steps by less (dense tile, stn_band_plan.cuh); bands of a CTA evenly sized
    int max_r = 1;
    while (max_r < p.band_rows_cta && (22 * max_r + 9) / 10 + 3 <= cap_max && 2 * (max_r + 1) <= cap_max) ++max_r;
    const int nb = (p.band_rows_cta + max_r - 1) / max_r;
    p.band_rows = (p.band_rows_cta + nb - 1) / nb;
    if (g_band_rows > 0 && g_band_rows < p.band_rows) p.band_rows = g_band_rows;
    {
        const int need = (22 * (p.band_rows - 1) + 9) / 10 + 3 > 2 * p.band_rows ? (22 * (p.band_rows - 1) + 9) / 10 + 3 : 2 * p.band_rows;
        p.band_cap = need < cap_max ? need : cap_max;
    }
    if (rowband) {
        // row bands: eight warp-private tiles of two frame rows each
        p.band_rows = 1;
        p.band_cap = 2 * kWarps;                                       // rows of the whole tile region (two per warp)
    }
    p.band_flags = g_band_variant >> 4;
    p.band_tab_rows = p.band_rows_cta + kBandMaxHalo + 1;
    p.band_tile_bytes = (int)(slot_bytes * p.band_cap);
    // declined crops run the general gx role with vector stores: same tile geometry as launch_crop_bwd
    {
        const int nx = (p.W + 63) / 64;
        const int tw = (((p.W + nx - 1) / nx) + 3) & ~3;
        const int tr = p.H < 8 ? p.H : 8;
        p.gx_tile_rows = tr; p.gx_tile_cols = tw; p.gx_tile_pitch = tw;
        p.gx_tiles_x = (p.W + tw - 1) / tw;
        p.gx_tiles_per_frame = p.gx_tiles_x * ((p.H + tr - 1) / tr);
        p.gx_tile_bytes = (int)(sizeof(float) * (size_t)p.C * tr * tw * kWarps);
        p.gx_vec4 = 1; p.gx_tma_store = 0; p.gx_zero_bytes = 0;
        p.band_fb_tiles_per_warp = (p.gx_tiles_per_frame + kWarps * (int)cs - 1) / (kWarps * (int)cs);
    }
    const int region = p.band_tile_bytes + p.band_zero_bytes;
    p.band_region_bytes = ((region > p.gx_tile_bytes ? region : p.gx_tile_bytes) + 127) & ~127;
    const size_t smem = (size_t)p.band_region_bytes + sizeof(BwdSmem) + sizeof(BandAxis) * (size_t)(p.oW + p.band_tab_rows) +
                        sizeof(int2) * (size_t)p.band_tab_rows + sizeof(float) * (size_t)((p.oW + p.oH + 3) & ~3) + sizeof(ScatterGeom);
    if (smem > 200 * 1024) return -1;
    const long long ctas = (long long)p.N * cs;
    if (ctas > 0x7fffffffLL) return -1;
    const int kind = rowband ? 3 : (knob == 2 ? 2 : 1);
    cudaError_t e;
    if (p.nhwc)
        e = kind == 3 ? launch_band_ttt<Nhwc4, 3, 1, 4, false, true>(p, (unsigned)ctas, cs, smem, stream)
                      : launch_band_ttt<Nhwc4, 3, 1, 4, false, false>(p, (unsigned)ctas, cs, smem, stream);
    else if (gy_dtype == 0)
        e = p.C == 1 ? launch_band_tt<float, 1>(p, (unsigned)ctas, cs, smem, stream, kind)
          : p.C == 3 ? launch_band_tt<float, 3>(p, (unsigned)ctas, cs, smem, stream, kind)
                     : launch_band_tt<float, 4>(p, (unsigned)ctas, cs, smem, stream, kind);
    else
        e = p.C == 1 ? launch_band_tt<__nv_bfloat16, 1>(p, (unsigned)ctas, cs, smem, stream, kind)
          : p.C == 3 ? launch_band_tt<__nv_bfloat16, 3>(p, (unsigned)ctas, cs, smem, stream, kind)
                     : launch_band_tt<__nv_bfloat16, 4>(p, (unsigned)ctas, cs, smem, stream, kind);
    count_launch();
    note_kernel(rowband ? "stn_bwd_band_kernel/row" : "stn_bwd_band_kernel/cta");
    if (e != cudaSuccess) return set_error("crop_bwd (band) launch failed: %s", cudaGetErrorString(e));
    return 0;
}

#ifdef STN_BAND_TRACE
extern "C" int loans_stn_debug_band_trace(void *buf)
{
    long long *q = reinterpret_cast<long long *>(buf);
    return cudaMemcpyToSymbol(g_band_trace, &q, sizeof(q)) == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace stn
