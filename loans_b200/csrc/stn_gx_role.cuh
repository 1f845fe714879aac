// stn_gx_role.cuh -- the general gx role of the backward (any theta, any number of crops per frame): warp-owned
// shared-memory tiles of frame pixels, phased scatter of the crop pixels that touch them, each gx element written
// exactly once.  Shared by stn_bwd_kernel (stn_crop.cu) and, for the crops it declines, the band backward
// (stn_band.cu).
#pragma once
#include <cuda.h>

#include "stn_common.cuh"

namespace stn {

// gy element i of a crop for the per-frame-pixel gather (i = channel * plane + pixel); behind the grayscale epilogue
// every channel reads the single gray plane, scaled by its coefficient
template <typename GT, bool GRAY = false>
struct GyLoader {
    size_t plane;
    __device__ __forceinline__ float operator()(const GT *p, size_t i) const
    {
        if (!GRAY) return Elem<GT>::load(p, i);
        const int ch = (int)(i / plane);
        return f_mul(gray_coef(ch), Elem<GT>::load(p, i - (size_t)ch * plane));
    }
};

template <>
struct GyLoader<Nhwc4, false> {                // channels-last: element (channel, pixel) of a crop
    size_t plane;
    __device__ __forceinline__ float operator()(const Nhwc4 *p, size_t i) const
    {
        const int ch = (int)(i / plane);
        return __bfloat162float(p[i - (size_t)ch * plane].c[ch]);
    }
};

// this crop's gy for channel group c0 (pixel 0): planar (N, C, oH, oW), one gray plane, or channels-last pixels
template <typename GT, bool GRAY>
__device__ __forceinline__ const GT *crop_gy(const GT *gy, size_t crop, int C, int c0, int npx)
{
    return crop_planes<GT, GRAY>(C) == 1 ? gy + crop * npx : gy + (crop * C + c0) * npx;
}

#ifndef STN_BWD_MIN_CTAS
#define STN_BWD_MIN_CTAS 4
#endif

// e / d for 0 <= e < 2^20 via one multiply (exact there: |error| <= 1.2e-7 * (e/d) < 0.5/d), integer division beyond
__device__ __forceinline__ int div_small(int e, int d, float inv_d, bool small)
{
    return small ? __float2int_rz(((float)e + 0.5f) * inv_d) : e / d;
}

template <typename GT, int CG, bool GRAY>
__device__ __noinline__ void gx_writeout_slow(const CropParams &p, const float *xs, const float *ys, const float *tile,
                                              float *gxb, const GT *gy, int b, int c0, int nc,
                                              int r0, int tr, int s0, int tw, bool any_fallback);

// gx role.  Every WARP owns one tile of frame pixels (gx_tile_rows x gx_tile_cols x CG channels) in shared memory and
// works through it on its own: zero, phased scatter of the crop pixels that touch it, write-out -- synchronising
// with __syncwarp() only.  The eight warps of a CTA take eight consecutive tiles of the same frame and share one
// prologue (axis tables + per-crop geometry) behind the CTA's single barrier.  The CTA is number cta_in_frame of the
// CTAs working on frame b; each of its warps takes tiles_per_warp tiles.
template <typename GT, int CG, bool EXACT, bool GRAY = false>
__device__ __forceinline__ void gx_role(const CropParams &p, const CUtensorMap *gx_map, const float *xs, const float *ys,
                                        const bool any_fallback, const ScatterGeom *geom, float *tiles, const float *zero_plane,
                                        const int b, const int cta_in_frame, const int tiles_per_warp)
{
    const int C = EXACT ? CG : p.C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this warp's tiles: gx_tiles_per_warp of them, interleaved over the warps of the CTA (warp-uniform loop; no CTA
    // barrier below)
  for (int tt = 0; tt < tiles_per_warp; ++tt) {
    const int tix = (cta_in_frame * tiles_per_warp + tt) * kWarps + warp;
    if (tix >= p.gx_tiles_per_frame) break;
    const int ty = tix / p.gx_tiles_x, tx = tix - ty * p.gx_tiles_x;
    const int r0 = ty * p.gx_tile_rows, s0 = tx * p.gx_tile_cols;
    const int tr = min(p.gx_tile_rows, p.H - r0), tw = min(p.gx_tile_cols, p.W - s0);
    const int twp = p.gx_tile_pitch;
    const int tile_plane = p.gx_tile_rows * twp;
    float *tile = tiles + warp * (CG * tile_plane);
    const int npx = p.oH * p.oW, fpx = p.H * p.W;
    const GT *gy = reinterpret_cast<const GT *>(p.gy);

    for (int c0 = 0; c0 < C; c0 += CG) {
        const int nc = EXACT ? CG : min(CG, C - c0);
        bool touched = false;                                                  // warp-uniform
        for (int kk = 0; kk < p.K; ++kk) {
            const ScatterGeom &g = geom[kk];
            if (g.P == 0) continue;                                            // gather fallback crop
            // cheap reject first: the crop's frame bounding box against the tile
            if (g.r_max < r0 || g.r_min >= r0 + tr || g.s_max < s0 || g.s_min >= s0 + tw) continue;
            int i_lo, i_hi, j_lo, j_hi;
            if (!scatter_box(g, r0, tr, s0, tw, p.oH, p.oW, i_lo, i_hi, j_lo, j_hi)) continue;
            if (!touched) {                                                    // zero the tile on first use only
                float4 *t4 = reinterpret_cast<float4 *>(tile);
                const int n4 = CG * tile_plane / 4;
                for (int e = lane; e < n4; e += 32) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
                touched = true;
            }
            const GT *gyc = crop_gy<GT, GRAY>(gy, (size_t)(b * p.K + kk), C, c0, npx);
            const Theta th = g.th;
            const int P = g.P, Q = g.Q;
            for (int cp = 0; cp < P; ++cp)
                for (int cq = 0; cq < Q; ++cq) {
                    // phase (cp, cq): crop pixels with i == cp (mod P), j == cq (mod Q), dealt to the 32 lanes
                    const bool one = (P | Q) == 1;                              // down-sampling by >= 2: the usual case
                    const int ia = one ? i_lo : first_congruent(i_lo, cp, P), ja = one ? j_lo : first_congruent(j_lo, cq, Q);
                    const int nrows = one ? i_hi - i_lo + 1 : (ia <= i_hi ? (i_hi - ia) / P + 1 : 0);
                    const int ncols = one ? j_hi - j_lo + 1 : (ja <= j_hi ? (j_hi - ja) / Q + 1 : 0);
                    const int total = nrows * ncols;
                    const bool small = total < (1 << 20);
                    const float inv_nc = 1.0f / (float)max(ncols, 1);
                    for (int e = lane; e < total; e += 32) {
                        const int rr = div_small(e, ncols, inv_nc, small);
                        const int i = ia + rr * P, j = ja + (e - rr * ncols) * Q;
                        if (!scatter_pretest(g, i, j, r0, tr, s0, tw)) continue;
                        float gv[CG];
                        const GT *gp = gyc + i * p.oW + j;
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch) gv[ch] = ch < nc ? load_gy<GT, GRAY>(gp, ch, npx) : 0.f;
                        ScatterTaps st;
                        if (!scatter_taps(th, xs[j], ys[i], p.H, p.W, r0, tr, s0, tw, st)) continue;
                        const Tap &t = st.t;
                        float *t00 = tile + st.row0 * twp + st.col0;
                        const bool b00 = st.rv0 && st.cv0, b01 = st.rv0 && st.cv1, b10 = st.rv1 && st.cv0, b11 = st.rv1 && st.cv1;
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch)
                            if (ch < nc) {
                                float *tc = t00 + ch * tile_plane;
                                const float a1 = f_mul(gv[ch], t.wu1), a0 = f_mul(gv[ch], t.wu0);   // gy * wu * wv, reference order
                                if (b00) tc[0] = f_add(tc[0], f_mul(a1, t.wv1));
                                if (b01) tc[1] = f_add(tc[1], f_mul(a0, t.wv1));
                                if (b10) tc[twp] = f_add(tc[twp], f_mul(a1, t.wv0));
                                if (b11) tc[twp + 1] = f_add(tc[twp + 1], f_mul(a0, t.wv0));
                            }
                    }
                    __syncwarp();                                              // phase (and crop) boundary
                }
        }
        // write the tile out: each gx element exactly once, zeros included
        float *gxb = p.gx + ((size_t)b * C + c0) * fpx;
        if (p.gx_tma_store && !any_fallback && !touched) {
            // no crop reaches this tile: its gx is zero -- one tensor store per channel from the CTA's zero plane
            if (lane < nc) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(zero_plane);
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(gx_map), "r"(s0), "r"(r0), "r"(b * C + c0 + lane), "r"(src) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (p.gx_vec4 && !any_fallback && !touched) {
            // no crop reaches this tile: its gx is zero, written straight from registers
            const int tw4 = tw >> 2;
            const int total = tr * tw4;
            int row = lane / tw4, c4 = lane - row * tw4;
            const int drow = 32 / tw4, dc4 = 32 - drow * tw4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int e = lane; e < total; e += 32) {
                float *gp = gxb + (size_t)(r0 + row) * p.W + s0 + 4 * c4;
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    if (ch < nc) *reinterpret_cast<float4 *>(gp + (size_t)ch * fpx) = z;
                row += drow; c4 += dc4;
                if (c4 >= tw4) { c4 -= tw4; ++row; }
            }
        } else if (p.gx_tma_store && !any_fallback) {
            // TMA tensor store: gx is described to the TMA unit as a (W, H, B*C) tensor with a (tile_cols, tile_rows, 1)
            // box; one instruction per channel moves the whole tile from shared memory, clipped at the frame edges by
            // the hardware.  The warp only waits until the tile has been READ before reusing the shared memory.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy tile writes -> async proxy
            __syncwarp();
            if (lane < nc) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + lane * tile_plane);
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                             ::"l"(gx_map), "r"(s0), "r"(r0), "r"(b * C + c0 + lane), "r"(src) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
        } else if (p.gx_vec4 && !any_fallback) {
            const int tw4 = tw >> 2;                                           // tw % 4 == 0 guaranteed by the host
            const int total = tr * tw4;
            int row = lane / tw4, c4 = lane - row * tw4;
            const int drow = 32 / tw4, dc4 = 32 - drow * tw4;
            for (int e = lane; e < total; e += 32) {
                const float *tp = tile + row * twp + 4 * c4;
                float *gp = gxb + (size_t)(r0 + row) * p.W + s0 + 4 * c4;
                float4 v[CG];
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    if (ch < nc) v[ch] = *reinterpret_cast<const float4 *>(tp + ch * tile_plane);
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    if (ch < nc) *reinterpret_cast<float4 *>(gp + (size_t)ch * fpx) = v[ch];
                row += drow; c4 += dc4;
                if (c4 >= tw4) { c4 -= tw4; ++row; }
            }
        } else {
            if (!touched) {
                float4 *t4 = reinterpret_cast<float4 *>(tile);
                const int n4 = CG * tile_plane / 4;
                for (int e = lane; e < n4; e += 32) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                __syncwarp();
            }
            gx_writeout_slow<GT, CG, GRAY>(p, xs, ys, tile, gxb, gy, b, c0, nc, r0, tr, s0, tw, any_fallback);
        }
        __syncwarp();
    }
  }
    if (p.gx_tma_store) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // this thread's tensor stores have landed
}

// Scalar write-out of a warp's tile, plus -- for crops whose transform is too degenerate for the phased scatter --
// their contribution gathered per frame pixel (exact, slow, rare).  Out of line: keeps the fast path's registers low.
template <typename GT, int CG, bool GRAY>
__device__ __noinline__ void gx_writeout_slow(const CropParams &p, const float *xs, const float *ys, const float *tile,
                                              float *gxb, const GT *gy, int b, int c0, int nc,
                                              int r0, int tr, int s0, int tw, bool any_fallback)
{
    const int lane = threadIdx.x & 31;
    const int twp = p.gx_tile_pitch, tile_plane = p.gx_tile_rows * twp;
    const int npx = p.oH * p.oW, fpx = p.H * p.W;
    for (int e = lane; e < tr * tw; e += 32) {
        const int row = e / tw, col = e - row * tw;
        float acc[CG];
#pragma unroll
        for (int ch = 0; ch < CG; ++ch) acc[ch] = ch < nc ? tile[ch * tile_plane + row * twp + col] : 0.f;
        if (any_fallback)
            for (int kk = 0; kk < p.K; ++kk) {
                const Theta th = load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01);
                if (make_scatter_geom(th, p.H, p.W, p.oH, p.oW).P != 0) continue;
                const InvCrop inv = make_inv_crop(th, p.H, p.W, p.oH, p.oW);
                const GyLoader<GT, GRAY> ld = {(size_t)npx};
                gather_from_crop<CG>(inv, xs, ys, p.H, p.W, p.oH, p.oW, r0 + row + 1, s0 + col + 1,
                                     crop_gy<GT, GRAY>(gy, (size_t)(b * p.K + kk), p.C, c0, npx),
                                     nc, ld, acc);
            }
#pragma unroll
        for (int ch = 0; ch < CG; ++ch)
            if (ch < nc) gxb[(size_t)ch * fpx + (size_t)(r0 + row) * p.W + s0 + col] = acc[ch];
    }
}

}  // namespace stn
