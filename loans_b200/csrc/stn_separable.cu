// stn_separable.cu -- the production path: axis-aligned crops (theta01 * mask == theta10 * mask == 0), which is
// what LoANs always runs because it calls rotation_dropout(..., ratio=0.0) in front of the grid
// (reference sheep/sheep_localizer.py:61,169).  Selected by the host whenever mask01 == 0; results are bit-identical
// to the general kernels of stn_crop.cu (same float32 operations, evaluated once per crop column / row instead
// of once per pixel).
//
// What the structure buys:
//   * the coordinate chain runs oW + oH times per crop instead of oH * oW times (AxisTap tables in shared memory);
//   * the frame rows a CTA needs are exactly two per crop row, over one contiguous column range: they are staged
//     into shared memory with TMA bulk copies (cp.async.bulk, completion on an mbarrier), i.e. few large coalesced
//     requests in flight instead of 12 scattered 4-byte loads per pixel, and the taps are then read from shared
//     memory;
//   * the gx scatter needs no candidate search: a crop row / column touches a tile iff its table entry says so.
#include <cooperative_groups.h>

#include "stn_common.cuh"

namespace cg = cooperative_groups;

namespace stn {

// ------------------------------------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init()
{
    // make the initialised barrier visible to the async proxy that will complete transactions on it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "STN_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra STN_DONE_%=;\n\t"
        "bra STN_WAIT_%=;\n\t"
        "STN_DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA, 1-D): 16-byte aligned addresses, size a multiple of 16
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------------------ staging
// Shared-memory plan of a CTA that samples crop rows [ia, ia + nrows) of one crop:
//   col[oW], row[rows_cap] AxisTap tables, then rows_cap * 2 * C staged frame-row segments of `pitch` floats.
struct SepStage {
    int s_lo;       // first staged frame column (unpadded, multiple of 4)
    int len;        // staged columns (multiple of 4); 0: nothing to stage (box entirely outside the frame)
};

// Column range all taps of this crop need, widened to 16-byte boundaries.  col[] is monotone in j.
__device__ __forceinline__ SepStage stage_plan(const AxisTap *col, int oW, int W)
{
    const int a = col[0].idx0, b = col[oW - 1].idx0;
    int lo = min(a, b), hi = max(a, b) + 1;          // padded tap columns lo .. hi
    lo = max(lo, 1);
    hi = min(hi, W);
    SepStage st;
    if (lo > hi) { st.s_lo = 0; st.len = 0; return st; }
    st.s_lo = (lo - 1) & ~3;
    st.len = min((((hi - 1) - st.s_lo + 1) + 3) & ~3, W - st.s_lo);     // W % 4 == 0 on this path
    return st;
}

// Warp 0 issues the bulk copies of rows [0, nrows) x {v0, v0+1} x C channels; everybody then waits on `bar`.
// Rows in the zero frame are not copied (their taps are read as 0 by the consumer).
__device__ __forceinline__ void stage_issue(const CropParams &p, const float *frame, const AxisTap *row, int nrows,
                                            const SepStage st, float *buf, int pitch, uint64_t *bar)
{
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    const int plane = p.H * p.W;
    const uint32_t bytes = (uint32_t)st.len * 4u;
    int valid = 0;
    if (st.len > 0)
        for (int r = 0; r < nrows; ++r) {
            const int v0 = row[r].idx0;
            valid += (v0 >= 1 && v0 <= p.H) + (v0 + 1 >= 1 && v0 + 1 <= p.H);
        }
    if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)valid * (uint32_t)p.C * bytes);
    __syncwarp();
    if (st.len == 0) return;
    const int total = nrows * 2 * p.C;
    for (int e = lane; e < total; e += 32) {
        const int c = e % p.C, rt = e / p.C;
        const int r = rt >> 1, t = rt & 1;
        const int v = row[r].idx0 + t;                       // padded frame row
        if (v < 1 || v > p.H) continue;
        tma_bulk_g2s(buf + (size_t)e * pitch, frame + (size_t)c * plane + (size_t)(v - 1) * p.W + st.s_lo, bytes, bar);
    }
}

// the four taps of crop pixel (r local row, column tap ct) for channel c, read from the staged rows
__device__ __forceinline__ void staged_taps(const float *buf, int pitch, int C, int r, int c, const AxisTap &ct,
                                            const AxisTap &rt, const SepStage st, int H, int W,
                                            float &x1, float &x2, float &x3, float &x4)
{
    const bool c0 = ct.idx0 >= 1, c1 = ct.idx0 <= W - 1;
    const bool r0 = rt.idx0 >= 1, r1 = rt.idx0 <= H - 1;
    const float *top = buf + (size_t)((r * 2 + 0) * C + c) * pitch + (ct.idx0 - 1 - st.s_lo);
    const float *bot = buf + (size_t)((r * 2 + 1) * C + c) * pitch + (ct.idx0 - 1 - st.s_lo);
    x1 = (r0 && c0) ? top[0] : 0.0f;
    x2 = (r0 && c1) ? top[1] : 0.0f;
    x3 = (r1 && c0) ? bot[0] : 0.0f;
    x4 = (r1 && c1) ? bot[1] : 0.0f;
}

struct SepSmem {
    uint64_t bar;
    float red[kWarps][6];
    float part[6];
    int pad_[2];
};

// ------------------------------------------------------------------------------------------ forward
template <typename YT>
__global__ void __launch_bounds__(kThreads) stn_sep_fwd_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SepSmem &sm = *reinterpret_cast<SepSmem *>(smem_raw);
    AxisTap *col = reinterpret_cast<AxisTap *>(smem_raw + sizeof(SepSmem));
    AxisTap *row = col + p.oW;
    float *buf = reinterpret_cast<float *>(smem_raw + p.sep_buf_offset);
    const int pitch = p.sep_pitch;

    const int n = blockIdx.x / p.ctas_per_crop;
    const int tile = blockIdx.x - n * p.ctas_per_crop;
    const int ia = tile * p.sep_rows, nrows = min(p.sep_rows, p.oH - ia);
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar, 1);
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k < p.oW + nrows; k += kThreads) {
        if (k < p.oW) col[k] = make_axis_tap(th.t00, th.t01, th.t02, linspace_pm1(k, p.oW, p.xstep), true, p.W);
        else row[k - p.oW] = make_axis_tap(th.t11, th.t10, th.t12, linspace_pm1(ia + k - p.oW, p.oH, p.ystep), false, p.H);
    }
    __syncthreads();
    const SepStage st = stage_plan(col, p.oW, p.W);
    const float *frame = p.x + (size_t)(n / p.K) * p.C * p.H * p.W;
    stage_issue(p, frame, row, nrows, st, buf, pitch, &sm.bar);

    const int npx = p.oH * p.oW;
    YT *yb = reinterpret_cast<YT *>(p.y) + (size_t)n * p.C * npx + (size_t)ia * p.oW;
    // the grid is a pure broadcast of the two tables: write it while the copies are in flight
    if (p.grid_out) {
        float *g0 = p.grid_out + (size_t)n * 2 * npx + (size_t)ia * p.oW, *g1 = g0 + npx;
        for (int e = threadIdx.x; e < nrows * p.oW; e += kThreads) {
            const int r = e / p.oW, j = e - r * p.oW;
            g0[e] = col[j].g;
            g1[e] = row[r].g;
        }
    }
    mbar_wait(&sm.bar, 0);
    for (int e = threadIdx.x; e < nrows * p.oW; e += kThreads) {
        const int r = e / p.oW, j = e - r * p.oW;
        const AxisTap ct = col[j], rt = row[r];
        const Weights4 wt = make_weights(tap_from_axes(ct, rt));
        for (int c = 0; c < p.C; ++c) {
            float x1, x2, x3, x4;
            staged_taps(buf, pitch, p.C, r, c, ct, rt, st, p.H, p.W, x1, x2, x3, x4);
            Elem<YT>::store(yb, (size_t)c * npx + e, interp(wt, x1, x2, x3, x4));
        }
    }
}

int launch_sep_fwd(CropParams p, int y_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    // rows per CTA: as many as fit ~44 KB of staging (two frame rows per crop row, full width worst case)
    const int pitch = (p.W + 3) & ~3;
    const size_t per_row = (size_t)2 * p.C * pitch * sizeof(float);
    int rows = (int)((44 * 1024) / per_row);
    if (rows < 1) return -1;                                     // caller falls back to the general kernel
    if (rows > 16) rows = 16;
    if (rows > p.oH) rows = p.oH;
    // do not starve the machine: at least ~2 CTAs per SM when the batch is small
    while (rows > 1 && (long long)p.N * ((p.oH + rows - 1) / rows) < 2LL * kNumSMs) rows = (rows + 1) / 2;
    p.sep_rows = rows;
    p.sep_pitch = pitch;
    p.ctas_per_crop = (p.oH + rows - 1) / rows;
    size_t off = sizeof(SepSmem) + sizeof(AxisTap) * (size_t)(p.oW + rows);
    off = (off + 127) & ~(size_t)127;
    p.sep_buf_offset = (int)off;
    const size_t smem = off + per_row * rows;
    const long long ctas = (long long)p.N * p.ctas_per_crop;
    if (ctas > 0x7fffffffLL) return set_error("sep_fwd: too many CTAs (%lld)", ctas);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) {
        static size_t granted[2] = {0, 0};
        if (smem > granted[y_dtype]) {
            e = y_dtype == 0 ? cudaFuncSetAttribute(stn_sep_fwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                             : cudaFuncSetAttribute(stn_sep_fwd_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return set_error("sep_fwd: cannot get %zu B of shared memory: %s", smem, cudaGetErrorString(e));
            granted[y_dtype] = smem;
        }
    }
    if (y_dtype == 0) stn_sep_fwd_kernel<float><<<(unsigned)ctas, kThreads, smem, stream>>>(p);
    else stn_sep_fwd_kernel<__nv_bfloat16><<<(unsigned)ctas, kThreads, smem, stream>>>(p);
    count_launch();
    return check_launch("sep_fwd");
}


// ------------------------------------------------------------------------------------------ backward
#ifndef STN_SEP_BWD_MIN_CTAS
#define STN_SEP_BWD_MIN_CTAS 4
#endif

struct AxisTap16 {          // what the gx scatter needs of an AxisTap, one 128-bit shared-memory load
    int idx0;
    float w0, w1;
    int pad_;
};

__device__ __forceinline__ float warp_sum_sep(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// theta role: a cluster of CTAs per crop, each CTA takes a contiguous band of crop rows, stages the frame rows of
// `sep_rows` crop rows at a time with TMA and reduces its six partial sums; cluster rank 0 adds the partials.
template <typename GT, int CG>
__device__ __forceinline__ void sep_theta_role(const CropParams &p, unsigned char *smem_raw, int cta)
{
    SepSmem &sm = *reinterpret_cast<SepSmem *>(smem_raw);
    AxisTap *col = reinterpret_cast<AxisTap *>(smem_raw + sizeof(SepSmem));
    AxisTap *row = col + p.oW;
    float *buf = reinterpret_cast<float *>(smem_raw + p.sep_buf_offset);
    const int pitch = p.sep_pitch;
    const int cs = p.ctas_per_crop;
    const int n = cta / cs, rank = cta - n * cs;
    const int rows_per_cta = (p.oH + cs - 1) / cs;
    const int ia = rank * rows_per_cta, ib = min(p.oH, ia + rows_per_cta);
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar, 1);
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k < p.oW; k += kThreads)
        col[k] = make_axis_tap(th.t00, th.t01, th.t02, linspace_pm1(k, p.oW, p.xstep), true, p.W);
    const int npx = p.oH * p.oW;
    const float *frame = p.x + (size_t)(n / p.K) * p.C * p.H * p.W;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * p.C * npx;
    float *ggo = p.ggrid_out ? p.ggrid_out + (size_t)n * 2 * npx : nullptr;
    const float *ggu = p.ggrid_up ? p.ggrid_up + (size_t)n * 2 * npx : nullptr;
    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t parity = 0;
    for (int i0 = ia; i0 < ib; i0 += p.sep_rows) {
        const int nrows = min(p.sep_rows, ib - i0);
        __syncthreads();                                   // previous chunk fully consumed (and col[] written)
        for (int k = threadIdx.x; k < nrows; k += kThreads)
            row[k] = make_axis_tap(th.t11, th.t10, th.t12, linspace_pm1(i0 + k, p.oH, p.ystep), false, p.H);
        __syncthreads();
        const SepStage st = stage_plan(col, p.oW, p.W);
        stage_issue(p, frame, row, nrows, st, buf, pitch, &sm.bar);
        // gy does not depend on the staged rows: the first pixel's loads go out before the wait, every later
        // pixel's one iteration ahead of its use
        const int cnt = nrows * p.oW;
        float gnext[CG];
#pragma unroll
        for (int ch = 0; ch < CG; ++ch)
            gnext[ch] = ((int)threadIdx.x < cnt && ch < p.C) ? Elem<GT>::load(gyb, (size_t)ch * npx + (size_t)i0 * p.oW + threadIdx.x) : 0.f;
        mbar_wait(&sm.bar, parity);
        parity ^= 1u;
        for (int e = threadIdx.x; e < cnt; e += kThreads) {
            const int r = e / p.oW, j = e - r * p.oW;
            const int q = i0 * p.oW + e;
            const AxisTap ct = col[j], rt = row[r];
            const Tap t = tap_from_axes(ct, rt);
            float su = 0.f, sv = 0.f;
            for (int c0 = 0; c0 < p.C; c0 += CG) {
                float g[CG];
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    g[ch] = c0 == 0 ? gnext[ch] : (c0 + ch < p.C ? Elem<GT>::load(gyb, (size_t)(c0 + ch) * npx + q) : 0.f);
                if (c0 == 0) {
                    const int en = e + kThreads;
#pragma unroll
                    for (int ch = 0; ch < CG; ++ch)
                        gnext[ch] = (en < cnt && ch < p.C) ? Elem<GT>::load(gyb, (size_t)ch * npx + (size_t)i0 * p.oW + en) : 0.f;
                }
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    if (c0 + ch < p.C) {
                        float x1, x2, x3, x4, gu, gv;
                        staged_taps(buf, pitch, p.C, r, c0 + ch, ct, rt, st, p.H, p.W, x1, x2, x3, x4);
                        grad_uv(t, x1, x2, x3, x4, gu, gv);
                        gu = f_mul(gu, g[ch]);
                        gv = f_mul(gv, g[ch]);
                        if (c0 + ch == 0) { su = gu; sv = gv; }
                        else { su = f_add(su, gu); sv = f_add(sv, gv); }
                    }
            }
            finish_grad_uv(t, p.H, p.W, su, sv);
            if (ggo) { ggo[q] = su; ggo[npx + q] = sv; }
            if (ggu) { su = f_add(su, __ldg(ggu + q)); sv = f_add(sv, __ldg(ggu + npx + q)); }
            const float xsj = ct.lin, ysi = rt.lin;
            s[0] = fmaf(su, xsj, s[0]); s[1] = fmaf(su, ysi, s[1]); s[2] += su;
            s[3] = fmaf(sv, xsj, s[3]); s[4] = fmaf(sv, ysi, s[4]); s[5] += sv;
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float r = warp_sum_sep(s[k]);
        if (lane == 0) sm.red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float tot = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi) tot += sm.red[wi][threadIdx.x];
        sm.part[threadIdx.x] = tot;
    }
    float *out = p.gtheta + 6 * (size_t)n;
    if (cs > 1) {
        cg::cluster_group cl = cg::this_cluster();
        cl.sync();
        if (rank == 0 && threadIdx.x < 6) {
            float tot = 0.f;
            for (int r = 0; r < cs; ++r) tot += *cl.map_shared_rank(&sm.part[threadIdx.x], r);
            if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
            out[threadIdx.x] = tot;
        }
        cl.sync();
    } else if (threadIdx.x < 6) {
        float tot = sm.part[threadIdx.x];
        if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
        out[threadIdx.x] = tot;
    }
}

// contiguous index range {k : tab[k].idx0 in [lo, hi]} of a monotone table, found by the whole warp with ballots
__device__ __forceinline__ bool warp_index_range(const AxisTap16 *tab, int n, int lo, int hi, int &ka, int &kb)
{
    const int lane = threadIdx.x & 31;
    ka = n; kb = -1;
    for (int base = 0; base < n; base += 32) {
        const int k = base + lane;
        const bool in = k < n && tab[k].idx0 >= lo && tab[k].idx0 <= hi;
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m) {
            ka = min(ka, base + __ffs(m) - 1);
            kb = max(kb, base + 31 - __clz(m));
        }
    }
    return ka <= kb;
}

// gx role: warp-owned tiles as in stn_crop.cu, but which crop pixels touch a tile is read off the per-crop column
// / row tables (no candidate search, no pre-test) and the taps come from the tables (no coordinate chain).
template <typename GT, int CG>
__device__ __forceinline__ void sep_gx_role(const CropParams &p, unsigned char *smem_raw)
{
    float *tiles = reinterpret_cast<float *>(smem_raw);
    AxisTap16 *tabs = reinterpret_cast<AxisTap16 *>(smem_raw + p.gx_tile_bytes);       // per crop: col[oW] then row[oH]
    int *pq = reinterpret_cast<int *>(tabs + (size_t)p.K * (p.oW + p.oH));              // per crop: P, Q
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / p.gx_ctas_per_frame;
    if (b >= p.N / p.K) return;                                                // padding CTA (cluster rounding)
    const int per = p.oW + p.oH;
    for (int e = threadIdx.x; e < p.K * per; e += kThreads) {
        const int kk = e / per, k = e - kk * per;
        const Theta th = load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01);
        const AxisTap a = k < p.oW ? make_axis_tap(th.t00, th.t01, th.t02, linspace_pm1(k, p.oW, p.xstep), true, p.W)
                                   : make_axis_tap(th.t11, th.t10, th.t12, linspace_pm1(k - p.oW, p.oH, p.ystep), false, p.H);
        AxisTap16 t16; t16.idx0 = a.idx0; t16.w0 = a.w0; t16.w1 = a.w1; t16.pad_ = 0;
        tabs[e] = t16;
    }
    for (int kk = threadIdx.x; kk < p.K; kk += kThreads) {
        const ScatterGeom g = make_scatter_geom(load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01),
                                                p.H, p.W, p.oH, p.oW);
        pq[2 * kk] = g.P; pq[2 * kk + 1] = g.Q;
    }
    __syncthreads();
    const int twp = p.gx_tile_pitch, tile_plane = p.gx_tile_rows * twp;
    float *tile = tiles + warp * (CG * tile_plane);
    const int npx = p.oH * p.oW, fpx = p.H * p.W;
    const GT *gy = reinterpret_cast<const GT *>(p.gy);
    const int cta_in_frame = blockIdx.x - b * p.gx_ctas_per_frame;

  // the warp works through gx_tiles_per_warp tiles of this frame, one after the other, on its own
  for (int tt = 0; tt < p.gx_tiles_per_warp; ++tt) {
    const int tix = (cta_in_frame * p.gx_tiles_per_warp + tt) * kWarps + warp;
    if (tix >= p.gx_tiles_per_frame) break;
    const int ty = tix / p.gx_tiles_x, tx = tix - ty * p.gx_tiles_x;
    const int r0 = ty * p.gx_tile_rows, s0 = tx * p.gx_tile_cols;
    const int tr = min(p.gx_tile_rows, p.H - r0), tw = min(p.gx_tile_cols, p.W - s0);

    for (int c0 = 0; c0 < p.C; c0 += CG) {
        const int nc = min(CG, p.C - c0);
        {
            float4 *t4 = reinterpret_cast<float4 *>(tile);
            const int n4 = CG * tile_plane / 4;
            for (int e = lane; e < n4; e += 32) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
        for (int kk = 0; kk < p.K; ++kk) {
            const int P = pq[2 * kk], Q = pq[2 * kk + 1];
            const AxisTap16 *col = tabs + (size_t)kk * per, *row = col + p.oW;
            int ja, jb, ia, ib;
            // taps of column j are padded columns idx0, idx0+1 = unpadded idx0-1, idx0: inside [s0, s0+tw) iff idx0 in [s0, s0+tw]
            if (!warp_index_range(col, p.oW, s0, s0 + tw, ja, jb)) continue;
            if (!warp_index_range(row, p.oH, r0, r0 + tr, ia, ib)) continue;
            const GT *gyc = gy + ((size_t)(b * p.K + kk) * p.C + c0) * npx;
            // P == 0 (degenerate scale: many crop rows/columns on one frame pixel): every pixel its own phase
            const int PP = P > 0 ? P : (ib - ia + 1), QQ = P > 0 ? Q : (jb - ja + 1);
            for (int cp = 0; cp < PP; ++cp)
                for (int cq = 0; cq < QQ; ++cq) {
                    const int i1 = first_congruent(ia, cp, PP), j1 = first_congruent(ja, cq, QQ);
                    const int nrows = i1 <= ib ? (ib - i1) / PP + 1 : 0;
                    const int ncols = j1 <= jb ? (jb - j1) / QQ + 1 : 0;
                    const int total = nrows * ncols;
                    const float inv_nc = 1.0f / (float)max(ncols, 1);
                    // candidate e of this lane -> (i, j); gy of the NEXT candidate is requested before the current
                    // one is added, so the load latency of a batch hides behind the previous batch
                    int e = lane, ci = 0, cj = 0;
                    float gnext[CG];
#pragma unroll
                    for (int ch = 0; ch < CG; ++ch) gnext[ch] = 0.f;
                    if (e < total) {
                        const int rr = __float2int_rz(((float)e + 0.5f) * inv_nc);
                        ci = i1 + rr * PP; cj = j1 + (e - rr * ncols) * QQ;
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch) if (ch < nc) gnext[ch] = Elem<GT>::load(gyc, (size_t)ch * npx + ci * p.oW + cj);
                    }
                    for (; e < total; e += 32) {
                        const int i = ci, j = cj;
                        float gv[CG];
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch) gv[ch] = gnext[ch];
                        const int en = e + 32;
                        if (en < total) {
                            const int rr = __float2int_rz(((float)en + 0.5f) * inv_nc);
                            ci = i1 + rr * PP; cj = j1 + (en - rr * ncols) * QQ;
#pragma unroll
                            for (int ch = 0; ch < CG; ++ch) if (ch < nc) gnext[ch] = Elem<GT>::load(gyc, (size_t)ch * npx + ci * p.oW + cj);
                        }
                        const AxisTap16 ct = col[j], rt = row[i];
                        const int row0 = rt.idx0 - 1 - r0, col0 = ct.idx0 - 1 - s0;
                        const bool rv0 = row0 >= 0 && row0 < tr && rt.w1 != 0.0f, rv1 = row0 + 1 >= 0 && row0 + 1 < tr && rt.w0 != 0.0f;
                        const bool cv0 = col0 >= 0 && col0 < tw && ct.w1 != 0.0f, cv1 = col0 + 1 >= 0 && col0 + 1 < tw && ct.w0 != 0.0f;
                        const bool b00 = rv0 && cv0, b01 = rv0 && cv1, b10 = rv1 && cv0, b11 = rv1 && cv1;
                        float *t00 = tile + row0 * twp + col0;
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch)
                            if (ch < nc) {
                                float *tc = t00 + ch * tile_plane;
                                const float a1 = f_mul(gv[ch], ct.w1), a0 = f_mul(gv[ch], ct.w0);   // gy * wu * wv
                                if (b00) tc[0] = f_add(tc[0], f_mul(a1, rt.w1));
                                if (b01) tc[1] = f_add(tc[1], f_mul(a0, rt.w1));
                                if (b10) tc[twp] = f_add(tc[twp], f_mul(a1, rt.w0));
                                if (b11) tc[twp + 1] = f_add(tc[twp + 1], f_mul(a0, rt.w0));
                            }
                    }
                    __syncwarp();
                }
        }
        float *gxb = p.gx + ((size_t)b * p.C + c0) * fpx;
        if (p.gx_vec4) {
            const int tw4 = tw >> 2;
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
            for (int e = lane; e < tr * tw; e += 32) {
                const int row = e / tw, cc = e - row * tw;
                for (int ch = 0; ch < nc; ++ch)
                    gxb[(size_t)ch * fpx + (size_t)(r0 + row) * p.W + s0 + cc] = tile[ch * tile_plane + row * twp + cc];
            }
        }
        __syncwarp();
    }
  }
}

template <typename GT, int CG>
__global__ void __launch_bounds__(kThreads, STN_SEP_BWD_MIN_CTAS) stn_sep_bwd_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if ((int)blockIdx.x < p.gx_ctas) sep_gx_role<GT, CG>(p, smem_raw);
    else sep_theta_role<GT, CG>(p, smem_raw, (int)blockIdx.x - p.gx_ctas);
}

template <typename GT, int CG>
static cudaError_t launch_sep_bwd_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    if (smem > 48 * 1024) {
        static size_t granted = 0;
        if (smem > granted) {
            cudaError_t e = cudaFuncSetAttribute(stn_sep_bwd_kernel<GT, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            granted = smem;
        }
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, stn_sep_bwd_kernel<GT, CG>, p);
}

int launch_sep_bwd(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const int cgsel = p.C == 1 ? 1 : (p.C % 3 == 0 ? 3 : 4);
    // theta role: cluster of cs CTAs per crop (rows split evenly), staging sep_rows crop rows at a time
    const int pitch = (p.W + 3) & ~3;
    const size_t per_row = (size_t)2 * p.C * pitch * sizeof(float);
    int rows = (int)((44 * 1024) / per_row);
    if (rows < 1) return -1;
    if (rows > 16) rows = 16;
#ifndef STN_SEP_THETA_CS_MAX
#define STN_SEP_THETA_CS_MAX 8
#endif
    unsigned cs = 1;
    while (cs < STN_SEP_THETA_CS_MAX && (long long)p.N * cs < 2LL * kNumSMs && (p.oH + 2 * cs - 1) / (2 * cs) >= 2) cs *= 2;
    const int rows_per_cta = (p.oH + (int)cs - 1) / (int)cs;
    if (rows > rows_per_cta) rows = rows_per_cta;
    p.sep_rows = rows;
    p.sep_pitch = pitch;
    p.ctas_per_crop = (int)cs;
    size_t off = sizeof(SepSmem) + sizeof(AxisTap) * (size_t)(p.oW + rows);
    off = (off + 127) & ~(size_t)127;
    p.sep_buf_offset = (int)off;
    size_t smem = off + per_row * rows;
    const long long theta_ctas = (long long)p.N * cs;
    long long gx_ctas = 0;
    if (p.gx) {
#ifndef STN_GX_TILE_ROWS
#define STN_GX_TILE_ROWS 8
#endif
#ifndef STN_GX_TILE_COLS
#define STN_GX_TILE_COLS 64
#endif
        const int nx = (p.W + STN_GX_TILE_COLS - 1) / STN_GX_TILE_COLS;
        const int tw = (((p.W + nx - 1) / nx) + 3) & ~3;
        int tr = STN_GX_TILE_ROWS;
        if (tr > p.H) tr = p.H;
        p.gx_tile_rows = tr; p.gx_tile_cols = tw; p.gx_tile_pitch = tw;
        p.gx_tiles_x = (p.W + tw - 1) / tw;
        p.gx_tiles_per_frame = p.gx_tiles_x * ((p.H + tr - 1) / tr);
        // tiles per warp: the fewest that keep gx CTAs + theta CTAs within about one wave of resident CTAs
#ifndef STN_SEP_WAVE_CTAS
#define STN_SEP_WAVE_CTAS (8 * kNumSMs)
#endif
        long long gx_budget = (long long)STN_SEP_WAVE_CTAS - theta_ctas;
        if (gx_budget < STN_SEP_WAVE_CTAS / 2) gx_budget = STN_SEP_WAVE_CTAS / 2;
        int tpw = 1;
        while (tpw < 64 && (long long)(p.N / p.K) * ((p.gx_tiles_per_frame + kWarps * tpw - 1) / (kWarps * tpw)) > gx_budget) ++tpw;
        p.gx_tiles_per_warp = tpw;
        p.gx_ctas_per_frame = (p.gx_tiles_per_frame + kWarps * tpw - 1) / (kWarps * tpw);
        p.gx_tile_bytes = (int)(sizeof(float) * (size_t)cgsel * tr * tw * kWarps);
        p.gx_vec4 = (p.W % 4 == 0 && (reinterpret_cast<uintptr_t>(p.gx) & 15) == 0) ? 1 : 0;
        const size_t gx_smem = (size_t)p.gx_tile_bytes + (sizeof(AxisTap16) * (size_t)(p.oW + p.oH) + 2 * sizeof(int)) * (size_t)p.K;
        if (gx_smem > 160 * 1024) return -1;                    // too many crops per frame for the tables: general kernel
        if (gx_smem > smem) smem = gx_smem;
        const long long n_gx = (long long)(p.N / p.K) * p.gx_ctas_per_frame;
        if (n_gx > 0x3fffffffLL) return set_error("sep_bwd: too many gx CTAs (%lld)", n_gx);
        gx_ctas = ((n_gx + cs - 1) / cs) * cs;
    }
    p.gx_ctas = (int)gx_ctas;
    const long long ctas = theta_ctas + gx_ctas;
    if (ctas > 0x7fffffffLL) return set_error("sep_bwd: too many CTAs (%lld)", ctas);
    cudaError_t e;
    if (gy_dtype == 0) {
        e = cgsel == 1 ? launch_sep_bwd_tt<float, 1>(p, (unsigned)ctas, cs, smem, stream)
          : cgsel == 3 ? launch_sep_bwd_tt<float, 3>(p, (unsigned)ctas, cs, smem, stream)
                       : launch_sep_bwd_tt<float, 4>(p, (unsigned)ctas, cs, smem, stream);
    } else {
        e = cgsel == 1 ? launch_sep_bwd_tt<__nv_bfloat16, 1>(p, (unsigned)ctas, cs, smem, stream)
          : cgsel == 3 ? launch_sep_bwd_tt<__nv_bfloat16, 3>(p, (unsigned)ctas, cs, smem, stream)
                       : launch_sep_bwd_tt<__nv_bfloat16, 4>(p, (unsigned)ctas, cs, smem, stream);
    }
    count_launch();
    if (e != cudaSuccess) return set_error("sep_bwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace stn
