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
// The backward of axis-aligned crops runs the general kernel (stn_crop.cu): two table-driven variants of the gx role
// and a TMA-staged theta role were built and measured this round and were not faster (profiles/README.md).
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
    while (rows > 1 && (long long)p.N * ((p.oH + rows - 1) / rows) < 2LL * num_sms()) rows = (rows + 1) / 2;
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
    e = y_dtype == 0 ? grant_dynamic_smem(reinterpret_cast<const void *>(&stn_sep_fwd_kernel<float>), smem)
                     : grant_dynamic_smem(reinterpret_cast<const void *>(&stn_sep_fwd_kernel<__nv_bfloat16>), smem);
    if (e != cudaSuccess) return set_error("sep_fwd: cannot get %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    if (y_dtype == 0) stn_sep_fwd_kernel<float><<<(unsigned)ctas, kThreads, smem, stream>>>(p);
    else stn_sep_fwd_kernel<__nv_bfloat16><<<(unsigned)ctas, kThreads, smem, stream>>>(p);
    count_launch();
    note_kernel("stn_sep_fwd_kernel");
    return check_launch("sep_fwd");
}


}  // namespace stn
