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
#include "stn_theta_role.cuh"

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
        if (k < p.oW) col[k] = make_axis_tap(th.t00, th.t01, th.t02, lin_x_at(p, k), true, p.W);
        else row[k - p.oW] = make_axis_tap(th.t11, th.t10, th.t12, lin_y_at(p, ia + k - p.oW), false, p.H);
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

struct AxisTap16 {          // what the gx role needs of an AxisTap, one 128-bit shared-memory load
    int idx0;
    float w0, w1;
    int pad_;
};

struct RowMatch {           // crop rows whose tap lands on one frame row: [a0, a0+ca) via tap v0 (weight w1),
    int a0, ca, b0, cb;     //                                             [b0, b0+cb) via tap v1 (weight w0)
};

constexpr int kSepOutPerThread = 4;      // float4 outputs per thread and channel in the gx role

// contiguous index range {k : tab[k].idx0 in [lo, hi]} of a monotone table, found by one warp with ballots
__device__ __forceinline__ void warp_index_range(const AxisTap16 *tab, int n, int lo, int hi, int &ka, int &kb)
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
}

// gx role for axis-aligned crops, as two separable passes -- gx = Wv^T (gy Wu):
//   pass 1  T[i][s] = sum_j gy[i][j] * wu(j -> s)        every crop row that reaches this tile is spread along the
//                                                        frame columns into a shared-memory row (1-D scatter, in Q
//                                                        conflict-free phases, see ScatterGeom);
//   pass 2  gx[r][s] = sum_i wv(i -> r) * T[i][s]        every frame row of the tile is a weighted sum of the (at most
//                                                        two, when down-sampling) T rows that reach it: float4 row
//                                                        operations, accumulated in registers over the crops of the
//                                                        frame and written to gx exactly once, zeros included.
// Products are formed as (gy * wu) * wv like the reference's scatter_add; no atomics, fixed order, deterministic.
template <typename GT, int CG>
__device__ __forceinline__ void sep_gx_role(const CropParams &p, unsigned char *smem_raw)
{
    // layout: [BwdSmem][xs|ys (theta role)] | tabs[K][oW+oH] | q[K] | rm[TR] | T[nt_cap][CG][twp]
    unsigned char *base = smem_raw + p.sep_buf_offset;
    AxisTap16 *tabs = reinterpret_cast<AxisTap16 *>(base);
    const int per = p.oW + p.oH;
    int *qtab = reinterpret_cast<int *>(tabs + (size_t)p.K * per);
    int *range = reinterpret_cast<int *>(qtab + ((p.K + 3) & ~3));        // [ia, ib, ja, jb]
    float *T = reinterpret_cast<float *>(range + 4);
    const int b = blockIdx.x / p.gx_tiles_per_frame;
    if (b >= p.N / p.K) return;                                                // padding CTA (cluster rounding)
    const int tix = blockIdx.x - b * p.gx_tiles_per_frame;
    const int ty = tix / p.gx_tiles_x, tx = tix - ty * p.gx_tiles_x;
    const int r0 = ty * p.gx_tile_rows, s0 = tx * p.gx_tile_cols;
    const int tr = min(p.gx_tile_rows, p.H - r0), tw = min(p.gx_tile_cols, p.W - s0);
    const int twp = p.gx_tile_pitch;
    const int npx = p.oH * p.oW, fpx = p.H * p.W;
    const GT *gy = reinterpret_cast<const GT *>(p.gy);

#ifdef STN_DEBUG_GX_ZERO_ONLY
    {   // experiment: the gx role only writes zeros (how much of the kernel is the dense write itself?)
        const int tw4z = (tw + 3) >> 2;
        for (int c = 0; c < p.C; ++c)
            for (int e = threadIdx.x; e < tr * tw4z; e += kThreads) {
                const int rr = e / tw4z, c4 = e - rr * tw4z;
                *reinterpret_cast<float4 *>(p.gx + ((size_t)b * p.C + c) * fpx + (size_t)(r0 + rr) * p.W + s0 + 4 * c4) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        return;
    }
#endif
    // tables of the K crops of this frame
    for (int e = threadIdx.x; e < p.K * per; e += kThreads) {
        const int kk = e / per, k = e - kk * per;
        const Theta th = load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01);
        const AxisTap a = k < p.oW ? make_axis_tap(th.t00, th.t01, th.t02, lin_x_at(p, k), true, p.W)
                                   : make_axis_tap(th.t11, th.t10, th.t12, lin_y_at(p, k - p.oW), false, p.H);
        AxisTap16 t16; t16.idx0 = a.idx0; t16.w0 = a.w0; t16.w1 = a.w1; t16.pad_ = 0;
        tabs[e] = t16;
    }
    for (int kk = threadIdx.x; kk < p.K; kk += kThreads) {
        // column phase period: two crop columns can share a frame column only if |du| < 2, i.e. |dj| < 2 / |du/dj|
        const Theta th = load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01);
        const float muj = fabsf(th.t00) * (p.oW > 1 ? 2.0f / (float)(p.oW - 1) : 0.0f) * 0.5f * (float)(p.W - 1);
        const float ej = muj > 1e-6f ? 2.2f / muj : 3.0e9f;
        qtab[kk] = ej <= 1.0f ? 1 : (ej < (float)p.oW ? (int)ceilf(ej) : p.oW);
    }

    // this thread's outputs: float4 column groups of tile rows
    const int tw4 = (tw + 3) >> 2;
    int o_row[kSepOutPerThread], o_c4[kSepOutPerThread];
    {
        int row = threadIdx.x / tw4, c4 = threadIdx.x - row * tw4;
        const int drow = kThreads / tw4, dc4 = kThreads - drow * tw4;
#pragma unroll
        for (int m = 0; m < kSepOutPerThread; ++m) {
            o_row[m] = row < tr ? row : -1;
            o_c4[m] = c4;
            row += drow; c4 += dc4;
            if (c4 >= tw4) { c4 -= tw4; ++row; }
        }
    }

    for (int c0 = 0; c0 < p.C; c0 += CG) {
        const int nc = min(CG, p.C - c0);
        float4 acc[kSepOutPerThread][CG];
#pragma unroll
        for (int m = 0; m < kSepOutPerThread; ++m)
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) acc[m][ch] = make_float4(0.f, 0.f, 0.f, 0.f);

        for (int kk = 0; kk < p.K; ++kk) {
            const AxisTap16 *col = tabs + (size_t)kk * per, *row = col + p.oW;
            __syncthreads();                                   // tables ready / previous crop's T and rm consumed
            // which crop rows / columns reach the tile at all (tables are monotone: contiguous ranges)
            if (threadIdx.x < 32) {
                int ka, kb;
                warp_index_range(row, p.oH, r0, r0 + tr, ka, kb);
                if (threadIdx.x == 0) { range[0] = ka; range[1] = kb; }
            } else if (threadIdx.x < 64) {
                int ka, kb;
                warp_index_range(col, p.oW, s0, s0 + tw, ka, kb);
                if (threadIdx.x == 32) { range[2] = ka; range[3] = kb; }
            }
            __syncthreads();
            const int ia = range[0], ib = range[1], ja = range[2], jb = range[3];
            if (ia > ib || ja > jb) continue;                  // CTA-uniform: this crop does not reach the tile
            const GT *gyc = gy + ((size_t)(b * p.K + kk) * p.C + c0) * npx;
            const int Q = qtab[kk];
            for (int i0 = ia; i0 <= ib; i0 += p.sep_rows) {    // T rows in chunks of sep_rows crop rows
                const int nt = min(p.sep_rows, ib - i0 + 1);
                if (i0 != ia) __syncthreads();                 // previous chunk consumed
                {
                    float4 *t4 = reinterpret_cast<float4 *>(T);
                    const int n4 = nt * CG * twp / 4;
                    for (int e = threadIdx.x; e < n4; e += kThreads) t4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                __syncthreads();
                // pass 1: spread gy rows along the frame columns
                for (int cq = 0; cq < Q; ++cq) {
                    const int j1 = first_congruent(ja, cq, Q);
                    const int ncols = j1 <= jb ? (jb - j1) / Q + 1 : 0;
                    const int total = nt * ncols;
                    const float inv_nc = 1.0f / (float)max(ncols, 1);
                    for (int e = threadIdx.x; e < total; e += kThreads) {
                        const int ii = total < (1 << 20) ? __float2int_rz(((float)e + 0.5f) * inv_nc) : e / ncols;
                        const int j = j1 + (e - ii * ncols) * Q;
                        const AxisTap16 ct = col[j];
                        const GT *gp = gyc + (i0 + ii) * p.oW + j;
                        const int cc = ct.idx0 - 1 - s0;
                        const bool v0 = cc >= 0 && cc < tw && ct.w1 != 0.0f, v1 = cc + 1 >= 0 && cc + 1 < tw && ct.w0 != 0.0f;
                        float *tc = T + (size_t)ii * CG * twp + cc;
#pragma unroll
                        for (int ch = 0; ch < CG; ++ch)
                            if (ch < nc) {
                                const float g = Elem<GT>::load(gp, (size_t)ch * npx);
                                if (v0) tc[ch * twp] = f_add(tc[ch * twp], f_mul(g, ct.w1));
                                if (v1) tc[ch * twp + 1] = f_add(tc[ch * twp + 1], f_mul(g, ct.w0));
                            }
                    }
                    __syncthreads();
                }
                // pass 2: weighted sums of T rows into this thread's output registers.  Crop row i reaches padded
                // frame rows idx0 (weight w1) and idx0 + 1 (weight w0).
                for (int ii = 0; ii < nt; ++ii) {
                    const AxisTap16 rt = row[i0 + ii];
                    const int ta = rt.idx0 - 1 - r0;               // tile row of tap v0; tap v1 is ta + 1
                    const float *trow = T + (size_t)ii * CG * twp;
#pragma unroll
                    for (int m = 0; m < kSepOutPerThread; ++m) {
                        const float wv = o_row[m] == ta ? rt.w1 : rt.w0;
                        if ((o_row[m] == ta || o_row[m] == ta + 1) && o_row[m] >= 0) {
                            const float *tp = trow + 4 * o_c4[m];
#pragma unroll
                            for (int ch = 0; ch < CG; ++ch)
                                if (ch < nc) {
                                    const float4 t = *reinterpret_cast<const float4 *>(tp + ch * twp);
                                    acc[m][ch].x = f_add(acc[m][ch].x, f_mul(t.x, wv)); acc[m][ch].y = f_add(acc[m][ch].y, f_mul(t.y, wv));
                                    acc[m][ch].z = f_add(acc[m][ch].z, f_mul(t.z, wv)); acc[m][ch].w = f_add(acc[m][ch].w, f_mul(t.w, wv));
                                }
                        }
                    }
                }
            }
        }
        // every gx element of the tile exactly once
        float *gxb = p.gx + ((size_t)b * p.C + c0) * fpx;
#pragma unroll
        for (int m = 0; m < kSepOutPerThread; ++m) {
            if (o_row[m] < 0) continue;
            float *gp = gxb + (size_t)(r0 + o_row[m]) * p.W + s0 + 4 * o_c4[m];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (ch < nc) {
                    if (p.gx_vec4) {
                        *reinterpret_cast<float4 *>(gp + (size_t)ch * fpx) = acc[m][ch];
                    } else {
                        const float v[4] = {acc[m][ch].x, acc[m][ch].y, acc[m][ch].z, acc[m][ch].w};
                        for (int k = 0; k < 4; ++k)
                            if (4 * o_c4[m] + k < tw) gp[(size_t)ch * fpx + k] = v[k];
                    }
                }
        }
    }
}

template <typename GT, int CG, bool EXACT>
__global__ void __launch_bounds__(kThreads, STN_SEP_BWD_MIN_CTAS) stn_sep_bwd_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    if ((int)blockIdx.x < p.gx_ctas) {
        sep_gx_role<GT, CG>(p, smem_raw);
    } else {
        BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
        float *xs = reinterpret_cast<float *>(smem_raw + sizeof(BwdSmem));
        float *ys = xs + p.oW;
        fill_axis_tables(p, xs, ys);
        __syncthreads();
        theta_role<GT, CG, EXACT>(p, xs, ys, sm, (int)blockIdx.x - p.gx_ctas);
    }
}

template <typename GT, int CG, bool EXACT>
static cudaError_t launch_sep_bwd_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    if (smem > 48 * 1024) {
        static size_t granted = 0;
        if (smem > granted) {
            cudaError_t e = cudaFuncSetAttribute(stn_sep_bwd_kernel<GT, CG, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
    return cudaLaunchKernelEx(&cfg, stn_sep_bwd_kernel<GT, CG, EXACT>, p);
}

template <typename GT>
static cudaError_t launch_sep_bwd_t(const CropParams &p, int cgsel, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    const bool exact = p.C == cgsel;
    switch (cgsel) {
    case 1: return launch_sep_bwd_tt<GT, 1, true>(p, ctas, cs, smem, s);
    case 3: return exact ? launch_sep_bwd_tt<GT, 3, true>(p, ctas, cs, smem, s) : launch_sep_bwd_tt<GT, 3, false>(p, ctas, cs, smem, s);
    default: return exact ? launch_sep_bwd_tt<GT, 4, true>(p, ctas, cs, smem, s) : launch_sep_bwd_tt<GT, 4, false>(p, ctas, cs, smem, s);
    }
}

int launch_sep_bwd(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const int cgsel = p.C == 1 ? 1 : (p.C % 3 == 0 ? 3 : 4);
    const long long npx = (long long)p.oH * p.oW;
    // theta role exactly as in the general backward (stn_crop.cu: launch_crop_bwd)
    unsigned cs = 1;
    while (cs < 8 && (long long)p.N * cs < 2LL * kNumSMs && npx / (2 * cs) >= kThreads / 2) cs *= 2;
    p.ctas_per_crop = (int)cs;
    p.px_per_cta = (int)((npx + cs - 1) / cs);
    const long long theta_ctas = (long long)p.N * cs;
    size_t smem = sizeof(BwdSmem) + sizeof(float) * (size_t)((p.oW + p.oH + 1) & ~1);
    smem = (smem + 127) & ~(size_t)127;
    p.sep_buf_offset = (int)smem;
    long long gx_ctas = 0;
    if (p.gx) {
        // tile: W cut evenly in column pieces of <= 256 (multiples of 4), rows so that a thread owns <= 4 float4 outputs
        const int nx = (p.W + 255) / 256;
        const int tw = (((p.W + nx - 1) / nx) + 3) & ~3;
        const int tw4 = tw / 4;
        int tr = (kSepOutPerThread * kThreads) / tw4;
        if (tr > 16) tr = 16;
        if (tr > p.H) tr = p.H;
        if (tr < 1) return -1;
        p.gx_tile_rows = tr; p.gx_tile_cols = tw; p.gx_tile_pitch = tw;
        p.gx_tiles_x = (p.W + tw - 1) / tw;
        p.gx_tiles_per_frame = p.gx_tiles_x * ((p.H + tr - 1) / tr);
        p.gx_vec4 = (p.W % 4 == 0 && (reinterpret_cast<uintptr_t>(p.gx) & 15) == 0) ? 1 : 0;
        // T rows per chunk: the crop rows that reach a tile when down-sampling by >= 2 (tr / 2 + 2), more if memory allows
        int nt = tr / 2 + 2;
        const size_t row_bytes = sizeof(float) * (size_t)cgsel * tw;
        while (nt < 16 && (size_t)(nt + 1) * row_bytes <= 32 * 1024) ++nt;
        if ((size_t)nt * row_bytes > 96 * 1024) return -1;
        p.sep_rows = nt;
        const size_t tabs = sizeof(AxisTap16) * (size_t)p.K * (p.oW + p.oH) + sizeof(int) * (size_t)((p.K + 3) & ~3)
                          + 4 * sizeof(int);
        const size_t gx_smem = (size_t)p.sep_buf_offset + ((tabs + 15) & ~(size_t)15) + (size_t)nt * row_bytes;
        if (gx_smem > 160 * 1024) return -1;                    // too many crops per frame for the tables: general kernel
        if (gx_smem > smem) smem = gx_smem;
        p.gx_tile_bytes = 0;
        const long long n_gx = (long long)(p.N / p.K) * p.gx_tiles_per_frame;
        if (n_gx > 0x3fffffffLL) return set_error("sep_bwd: too many gx CTAs (%lld)", n_gx);
        gx_ctas = ((n_gx + cs - 1) / cs) * cs;
    }
    p.gx_ctas = (int)gx_ctas;
    const long long ctas = theta_ctas + gx_ctas;
    if (ctas > 0x7fffffffLL) return set_error("sep_bwd: too many CTAs (%lld)", ctas);
    cudaError_t e = gy_dtype == 0 ? launch_sep_bwd_t<float>(p, cgsel, (unsigned)ctas, cs, smem, stream)
                                  : launch_sep_bwd_t<__nv_bfloat16>(p, cgsel, (unsigned)ctas, cs, smem, stream);
    count_launch();
    if (e != cudaSuccess) return set_error("sep_bwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace stn
