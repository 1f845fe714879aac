// stn_crop.cu -- the fused hot path: rotation_dropout -> affine grid -> bilinear sampler, forward and
// backward, one kernel launch per direction (reference sheep/sheep_localizer.py:61-63).
//
// Forward  (stn_fwd_kernel): one thread per crop pixel computes its own sample coordinates from the masked
//   theta (no grid round trip through HBM), gathers the 4 taps of every channel through the read-only
//   path with all loads of a channel group in flight together, and stores the crop (fp32 or bf16) and,
//   if asked, the grid -- a required output of the reference API -- from the same registers.
//
// Backward (stn_bwd_kernel): one launch, two CTA roles.
//   * gx role, CTAs [0, gx_ctas) (scheduled first: the long pole).  A CTA owns a tile of frame pixels in shared
//     memory.  For every crop of that frame it walks the crop pixels whose 2x2 tap window can touch the tile
//     (found by inverting the affine map), re-evaluates each with the exact forward chain and adds
//     gy * wu * wv into the tile.  Crop pixels are processed in P*Q conflict-free phases (stn_math.cuh:
//     ScatterGeom) so the adds are plain shared-memory read-modify-writes: no atomics, fixed summation order,
//     bit-reproducible.  The tile -- zeros included -- is then written to gx exactly once with coalesced
//     128-bit stores: no memset pass.  Work is proportional to the number of CROP pixels, not frame pixels.
//     Degenerate transforms (singular / extreme up-sampling) fall back, per crop, to an exact per-frame-pixel
//     gather at write-out time.
//   * theta role, the remaining CTAs: a thread-block CLUSTER per crop.  Each CTA walks its share of the crop
//     pixels (taps + gy -> d/du, d/dv, plus the upstream grid gradient), reduces the six sums of
//     gtheta = ggrid . [xs; ys; 1]^T with warp shuffles and shared memory, and rank 0 of the cluster adds the
//     per-CTA partials through distributed shared memory in a fixed order: deterministic, no atomics, no
//     zero-initialised output, no workspace.
#include <cooperative_groups.h>
#include <cuda.h>             // CUtensorMap (the encoder is fetched through cudaGetDriverEntryPoint: no libcuda link)

#include <cstring>
#include <type_traits>

#include "stn_common.cuh"
#include "stn_gx_role.cuh"
#include "stn_theta_role.cuh"

namespace cg = cooperative_groups;

namespace stn {

// ------------------------------------------------------------------------------------------ forward
// register budget: 64 (four CTAs per SM: cfg2's 512 CTAs are one wave) for float32 crops -- 6.2 vs 6.5 us at cfg2, no spills;
// with bf16 crops the same bound measured 3.7 % slower at cfg3, so those keep three CTAs per SM (76 registers)
template <typename YT> struct FwdMinCtas { static constexpr int value = 4; };
template <> struct FwdMinCtas<__nv_bfloat16> { static constexpr int value = 3; };
template <> struct FwdMinCtas<Nhwc4> { static constexpr int value = 3; };

template <typename YT, int CG, bool FROM_GRID, bool EXACT>
__global__ void __launch_bounds__(kThreads, FwdMinCtas<YT>::value) stn_fwd_kernel(const __grid_constant__ CropParams p)
{
    const int C = EXACT ? CG : p.C;          // EXACT: one channel group covers all channels, loops fold away
    extern __shared__ float smem[];
    float *xs = smem, *ys = smem + p.oW;
    pdl_launch_dependents();
    const int n = blockIdx.x / p.ctas_per_crop;
    if (!FROM_GRID) {
        pdl_prefetch_theta(p, n);
        fill_axis_tables(p, xs, ys);                 // pure arithmetic: overlaps the tail of the previous kernel
        __syncthreads();
    }
    pdl_wait();
    const int tile = blockIdx.x - n * p.ctas_per_crop;
    const int npx = p.oH * p.oW;
    const int q_end = min(npx, (tile + 1) * p.px_per_cta);
    Theta th = {};
    if (!FROM_GRID) th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    const int plane = p.H * p.W;
    const float *xb = p.x + (size_t)(n / p.K) * C * plane;
    YT *yb = reinterpret_cast<YT *>(p.y) + (size_t)n * (p.gray ? 1 : crop_planes<YT, false>(C)) * npx;
    float *gout = p.grid_out ? p.grid_out + (size_t)n * 2 * npx : nullptr;
    const float *gin = FROM_GRID ? p.grid_in + (size_t)n * 2 * npx : nullptr;
    if (!FROM_GRID && p.corners_out && tile == 0 && threadIdx.x < 4) {
        // the four corner points of the grid, the only ones LoANs' regularisers and box extraction read (reference
        // common/utils.py:141-159, sheep/sheep_localizer.py:84-91): same operations as the dense grid, so bit-identical
        const int ci = threadIdx.x >> 1, cj = threadIdx.x & 1;
        const float xv = xs[cj ? p.oW - 1 : 0], yv = ys[ci ? p.oH - 1 : 0];
        float *co = p.corners_out + 8 * (size_t)n + 2 * ci + cj;
        co[0] = grid_elem(th.t00, th.t01, th.t02, xv, yv);
        co[4] = grid_elem(th.t10, th.t11, th.t12, xv, yv);
    }

    if (EXACT) {
        // one channel group covers all channels: two pixels per thread in flight -- both pixels' 8*CG tap loads are
        // issued before either is interpolated, so a thread pays one memory round trip per PAIR of pixels
        struct Px {
            Weights4 wt;
            float v[CG][4];
            int q;
            bool live;
        };
        auto prepare = [&](Px &px, const PxWalk &w) {
            px.live = w.q < q_end;
            px.q = w.q;
            if (!px.live) return;
            float g0, g1;
            if (FROM_GRID) {
                g0 = __ldg(gin + w.q);
                g1 = __ldg(gin + npx + w.q);
            } else {
                const float xsj = xs[w.j], ysi = ys[w.i];
                g0 = grid_elem(th.t00, th.t01, th.t02, xsj, ysi);
                g1 = grid_elem(th.t10, th.t11, th.t12, xsj, ysi);
                if (gout) {
                    gout[w.q] = g0;
                    gout[npx + w.q] = g1;
                }
            }
            const Tap t = make_tap(g0, g1, p.H, p.W);
            const TapAddr a = make_tap_addr(t, p.H, p.W);
            px.wt = make_weights(t);
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) load_taps(xb + ch * plane, a, p.W, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3]);
        };
        auto finish = [&](const Px &px) {
            if (!px.live) return;
            if (CG == 3 && p.gray) {                   // grayscale epilogue: one output channel
                float c[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) c[ch] = interp(px.wt, px.v[ch % CG][0], px.v[ch % CG][1], px.v[ch % CG][2], px.v[ch % CG][3]);
                Elem<YT>::store(yb + px.q, 0, gray_mix(c[0], c[1], c[2]));
                return;
            }
            if constexpr (std::is_same<YT, Nhwc4>::value) {      // channels-last: the pixel's three channels in one 8-byte store
                float c[3];
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) c[ch] = interp(px.wt, px.v[ch % CG][0], px.v[ch % CG][1], px.v[ch % CG][2], px.v[ch % CG][3]);
                Elem<Nhwc4>::store_px(yb + px.q, c[0], c[1], c[2]);
                return;
            }
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                Elem<YT>::store(yb + px.q, ch * npx, interp(px.wt, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3]));
        };
        PxWalk w(tile * p.px_per_cta + threadIdx.x, p.oW);
        while (w.q < q_end) {
            Px A, B;
            prepare(A, w);
            w.next();
            prepare(B, w);
            w.next();
            finish(A);
            finish(B);
        }
    } else {
    for (PxWalk w(tile * p.px_per_cta + threadIdx.x, p.oW); w.q < q_end; w.next()) {
        float g0, g1;
        if (FROM_GRID) {
            g0 = __ldg(gin + w.q);
            g1 = __ldg(gin + npx + w.q);
        } else {
            const float xsj = xs[w.j], ysi = ys[w.i];
            g0 = grid_elem(th.t00, th.t01, th.t02, xsj, ysi);
            g1 = grid_elem(th.t10, th.t11, th.t12, xsj, ysi);
            if (gout) {
                gout[w.q] = g0;
                gout[npx + w.q] = g1;
            }
        }
        const Tap t = make_tap(g0, g1, p.H, p.W);
        const TapAddr a = make_tap_addr(t, p.H, p.W);
        const Weights4 wt = make_weights(t);
        const float *xc = xb;
        YT *yc = yb + w.q;
        for (int c0 = 0; c0 < C; c0 += CG) {
            float v[CG][4];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < C) load_taps(xc + ch * plane, a, p.W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < C) Elem<YT>::store(yc, ch * npx, interp(wt, v[ch][0], v[ch][1], v[ch][2], v[ch][3]));
            xc += CG * plane;
            yc += CG * npx;
        }
    }
}
}

// ------------------------------------------------------------------------------------------ backward
#ifdef STN_BAND_TRACE
// debug build only: per-CTA timestamps (globaltimer ns) of the general backward, 8 slots per CTA
__device__ long long *g_bwd_trace = nullptr;
__device__ __forceinline__ void bwd_trace(int slot, long long v = -1)
{
    if (threadIdx.x == 0 && g_bwd_trace) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_bwd_trace[(size_t)blockIdx.x * 8 + slot] = v >= 0 ? v : t;
    }
}
#define BTRACE(...) bwd_trace(__VA_ARGS__)
#else
#define BTRACE(...)
#endif

template <typename GT, int CG, bool EXACT, bool GRAY>
__global__ void __launch_bounds__(kThreads, STN_BWD_MIN_CTAS) stn_bwd_kernel(const __grid_constant__ CropParams p, const __grid_constant__ CUtensorMap gx_map)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];       // TMA tensor stores read 128-byte aligned tiles
    // layout: [8 warp tiles (gx role only, 16 B aligned)] [BwdSmem] [xs | ys] [ScatterGeom[K]]
    float *tiles = reinterpret_cast<float *>(smem_raw);
    float *zero_plane = reinterpret_cast<float *>(smem_raw + p.gx_tile_bytes);           // gx role, TMA path only
    unsigned char *q = smem_raw + p.gx_tile_bytes + p.gx_zero_bytes;
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(q);
    q += sizeof(BwdSmem);
    float *xs = reinterpret_cast<float *>(q);
    float *ys = xs + p.oW;
    q += sizeof(float) * ((p.oW + p.oH + 1) & ~1);
    ScatterGeom *geom = reinterpret_cast<ScatterGeom *>(q);
    // role of this CTA and its index inside the role; which role is scheduled first is a launch parameter
    const bool gx_cta = p.theta_first ? (int)blockIdx.x >= p.theta_ctas : (int)blockIdx.x < p.gx_ctas;
    const int role_idx = (int)blockIdx.x - (p.theta_first ? (gx_cta ? p.theta_ctas : 0) : (gx_cta ? 0 : p.gx_ctas));
    BTRACE(0);
    pdl_launch_dependents();
    fill_axis_tables(p, xs, ys);
    if (gx_cta && p.gx_zero_bytes) {
        float4 *z4 = reinterpret_cast<float4 *>(zero_plane);
        for (int e = threadIdx.x; e < p.gx_zero_bytes / 16; e += kThreads) z4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    pdl_wait();                                    // everything above touches shared memory only
    if (gx_cta) {
        // per-crop geometry of this CTA's frame (float32, conservative); P == 0 marks a gather-fallback crop.  The
        // CTA's single barrier also tells everybody whether any crop of the frame needs the fallback.
        const int b = role_idx / p.gx_ctas_per_frame;
        int fallback = 0;
        if (b < p.N / p.K)
            for (int kk = threadIdx.x; kk < p.K; kk += kThreads) {
                geom[kk] = make_scatter_geom(load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01),
                                             p.H, p.W, p.oH, p.oW);
                fallback |= geom[kk].P == 0;
            }
        const int any_fallback = __syncthreads_or(fallback);
        if (b >= p.N / p.K) return;                                            // padding CTA (cluster rounding)
        BTRACE(1);
        gx_role<GT, CG, EXACT, GRAY>(p, &gx_map, xs, ys, any_fallback != 0, geom, tiles, zero_plane, b,
                               role_idx - b * p.gx_ctas_per_frame, p.gx_tiles_per_warp);
        BTRACE(2);
        BTRACE(4, 1);
    } else {
        __syncthreads();
        BTRACE(1);
        theta_role<GT, CG, EXACT, GRAY>(p, xs, ys, sm, role_idx);
        BTRACE(2);
        BTRACE(4, 2);
    }
#ifdef STN_BAND_TRACE
    {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        BTRACE(5, (long long)smid);
    }
#endif
}

// Backward of frames that take no gradient (gx == NULL: every LoANs call, the frames are a raw array): the theta role
// alone, as its own lean kernel (no gx code, no tile shared memory).  Measured against the two-role kernel launched without
// gx CTAs: 7.3 vs 8.9 us at cfg2, 71 vs 76 us at cfg5, 4.0 vs 5.6 us at cfg1; two pixels in flight per thread (80
// registers, 3 CTAs/SM) measured slower here too (11.2 us at cfg2).
template <typename GT, int CG, bool EXACT, bool GRAY>
__global__ void __launch_bounds__(kThreads, 4) stn_bwd_theta_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    float *xs = reinterpret_cast<float *>(smem_raw + sizeof(BwdSmem));
    float *ys = xs + p.oW;
    pdl_launch_dependents();
    pdl_prefetch_theta(p, (int)blockIdx.x / p.ctas_per_crop);
    fill_axis_tables(p, xs, ys);
    pdl_wait();
    __syncthreads();
    theta_role<GT, CG, EXACT, GRAY, false>(p, xs, ys, sm, (int)blockIdx.x);
}

template <typename GT, int CG, bool EXACT, bool GRAY = false>
static cudaError_t launch_theta_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, cs);
    return cudaLaunchKernelEx(&cfg, stn_bwd_theta_kernel<GT, CG, EXACT, GRAY>, p);
}

template <typename GT>
static cudaError_t launch_theta_t(const CropParams &p, int cgsel, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    const bool exact = p.C == cgsel;
    switch (cgsel) {
    case 1: return launch_theta_tt<GT, 1, true>(p, ctas, cs, smem, s);
    case 3:
        if (exact && p.gray) return launch_theta_tt<GT, 3, true, true>(p, ctas, cs, smem, s);
        return exact ? launch_theta_tt<GT, 3, true>(p, ctas, cs, smem, s) : launch_theta_tt<GT, 3, false>(p, ctas, cs, smem, s);
    default: return exact ? launch_theta_tt<GT, 4, true>(p, ctas, cs, smem, s) : launch_theta_tt<GT, 4, false>(p, ctas, cs, smem, s);
    }
}

// ------------------------------------------------------------------------------------------ host launchers
static int pick_channel_group(int C) { return C == 1 ? 1 : (C % 3 == 0 ? 3 : 4); }

template <typename YT, int CG, bool FROM_GRID, bool EXACT>
static cudaError_t launch_fwd_tt(const CropParams &p, dim3 grid, size_t smem, cudaStream_t s)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, 0);
    return cudaLaunchKernelEx(&cfg, stn_fwd_kernel<YT, CG, FROM_GRID, EXACT>, p);
}

template <typename YT, bool FROM_GRID>
static cudaError_t launch_fwd_t(const CropParams &p, int cgsel, dim3 grid, size_t smem, cudaStream_t s)
{
    const bool exact = p.C == cgsel;
    switch (cgsel) {
    case 1: return launch_fwd_tt<YT, 1, FROM_GRID, true>(p, grid, smem, s);          // C == 1
    case 3: return exact ? launch_fwd_tt<YT, 3, FROM_GRID, true>(p, grid, smem, s) : launch_fwd_tt<YT, 3, FROM_GRID, false>(p, grid, smem, s);
    default: return exact ? launch_fwd_tt<YT, 4, FROM_GRID, true>(p, grid, smem, s) : launch_fwd_tt<YT, 4, FROM_GRID, false>(p, grid, smem, s);
    }
}

int launch_crop_fwd(CropParams p, bool from_grid, int y_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const long long npx = (long long)p.oH * p.oW;
    // enough CTAs to give every SM a few, at most 8 pixels per thread
    long long per = (npx * p.N + 4LL * num_sms() - 1) / (4LL * num_sms());
    per = ((per + kThreads - 1) / kThreads) * kThreads;
    if (per < kThreads) per = kThreads;
    if (per > 8 * kThreads) per = 8 * kThreads;
    if (fwd_px_per_cta_override() > 0) per = fwd_px_per_cta_override();
    p.px_per_cta = (int)per;
    p.ctas_per_crop = (int)((npx + per - 1) / per);
    const long long ctas = (long long)p.N * p.ctas_per_crop;
    if (ctas > 0x7fffffffLL) return set_error("crop_fwd: too many CTAs (%lld)", ctas);
    const dim3 grid((unsigned)ctas);
    const size_t smem = sizeof(float) * (size_t)(p.oW + p.oH);
    const int cgsel = pick_channel_group(p.C);
    cudaError_t e;
    if (p.nhwc)                                            // c == 3, bf16, fused path only (checked by the entry point)
        e = launch_fwd_tt<Nhwc4, 3, false, true>(p, grid, smem, stream);
    else if (y_dtype == 0)
        e = from_grid ? launch_fwd_t<float, true>(p, cgsel, grid, smem, stream)
                      : launch_fwd_t<float, false>(p, cgsel, grid, smem, stream);
    else
        e = from_grid ? launch_fwd_t<__nv_bfloat16, true>(p, cgsel, grid, smem, stream)
                      : launch_fwd_t<__nv_bfloat16, false>(p, cgsel, grid, smem, stream);
    count_launch();
    note_kernel(from_grid ? "stn_fwd_kernel/grid" : "stn_fwd_kernel");
    if (e != cudaSuccess) return set_error("crop_fwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

template <typename GT, int CG, bool EXACT, bool GRAY = false>
static cudaError_t launch_bwd_tt(const CropParams &p, const CUtensorMap &gx_map, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    {
        const cudaError_t e = grant_dynamic_smem(reinterpret_cast<const void *>(&stn_bwd_kernel<GT, CG, EXACT, GRAY>), smem);
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
    return cudaLaunchKernelEx(&cfg, stn_bwd_kernel<GT, CG, EXACT, GRAY>, p, gx_map);
}

template <typename GT>
static cudaError_t launch_bwd_t(const CropParams &p, const CUtensorMap &m, int cgsel, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    const bool exact = p.C == cgsel;
    switch (cgsel) {
    case 1: return launch_bwd_tt<GT, 1, true>(p, m, ctas, cs, smem, s);
    case 3:
        if (exact && p.gray) return launch_bwd_tt<GT, 3, true, true>(p, m, ctas, cs, smem, s);      // grayscale epilogue: its own kernel
        return exact ? launch_bwd_tt<GT, 3, true>(p, m, ctas, cs, smem, s) : launch_bwd_tt<GT, 3, false>(p, m, ctas, cs, smem, s);
    default: return exact ? launch_bwd_tt<GT, 4, true>(p, m, ctas, cs, smem, s) : launch_bwd_tt<GT, 4, false>(p, m, ctas, cs, smem, s);
    }
}

// gx as a 3-D tensor (W, H, B*C) with a (tile_cols, tile_rows, 1) box for the TMA tensor stores of the gx role.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool encode_gx_map(CUtensorMap *map, float *gx, int planes, int H, int W, int tile_rows, int tile_cols)
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
        else
            cudaGetLastError();
    }
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tile_cols, (cuuint32_t)tile_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, gx, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int launch_crop_bwd(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const long long npx = (long long)p.oH * p.oW;
    const int cgsel = pick_channel_group(p.C);
    // cluster size: enough theta-role CTAs to occupy the machine twice over, 8 (portable maximum) at most
#ifndef STN_THETA_CS_MAX
#define STN_THETA_CS_MAX 4
#endif
    unsigned cs = 1;
    // (the theta-only kernel of gx == NULL has the machine to itself: up to 8 CTAs per crop)
    const bool theta_only = !p.gx && theta_only_kernel_enabled();
    const unsigned cs_max = theta_only ? 8 : STN_THETA_CS_MAX;
    while (cs < cs_max && (long long)p.N * cs < 2LL * num_sms() && npx / (2 * cs) >= kThreads / 2) cs *= 2;
    p.ctas_per_crop = (int)cs;
    p.px_per_cta = (int)((npx + cs - 1) / cs);
    const long long theta_ctas = (long long)p.N * cs;
    long long gx_ctas = 0;
    alignas(64) CUtensorMap gx_map;
    memset(&gx_map, 0, sizeof(gx_map));
    size_t smem = sizeof(BwdSmem) + sizeof(float) * (size_t)((p.oW + p.oH + 1) & ~1);
    if (p.gx) {
        // one tile per warp: STN_GX_TILE_ROWS rows x (W cut evenly in pieces of <= STN_GX_TILE_COLS, multiple of 4)
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
        const int tiles_y = (p.H + tr - 1) / tr;
        p.gx_tiles_per_frame = p.gx_tiles_x * tiles_y;
        // tiles per warp: one while the batch is small (every CTA is resident at once and latency rules), more as the
        // tile count grows, so that the per-CTA prologue (tables, geometry, barrier) is amortised over several tiles
#ifndef STN_GX_MAX_TILES_PER_WARP
#define STN_GX_MAX_TILES_PER_WARP 8
#endif
        {
            const long long all_tiles = (long long)(p.N / p.K) * p.gx_tiles_per_frame;
            long long tpw = all_tiles / ((long long)kWarps * num_sms() * 6 * p.K);     // K crops per frame: K times the work per tile
            if (tpw < 1) tpw = 1;
            // several crops per frame: the per-CTA prologue derives K geometries -- three tiles per warp amortise it
            // (cfg4: 541 vs 561 us with one, 545 with two, 601 with eight)
            if (p.K > 1 && tpw < 3 && all_tiles >= 3LL * kWarps * 4 * num_sms()) tpw = 3;
            if (tpw > STN_GX_MAX_TILES_PER_WARP) tpw = STN_GX_MAX_TILES_PER_WARP;
            if (gx_tiles_per_warp_override() > 0) tpw = gx_tiles_per_warp_override();
            p.gx_tiles_per_warp = (int)tpw;
        }
        p.gx_ctas_per_frame = (p.gx_tiles_per_frame + kWarps * p.gx_tiles_per_warp - 1) / (kWarps * p.gx_tiles_per_warp);
        p.gx_tile_bytes = (int)(sizeof(float) * (size_t)cgsel * tr * tw * kWarps);
        p.gx_vec4 = (p.W % 4 == 0 && (reinterpret_cast<uintptr_t>(p.gx) & 15) == 0) ? 1 : 0;
        // TMA tensor store of the tiles: 16-byte aligned rows, 128-byte aligned channel planes in shared memory
#ifndef STN_GX_TMA_STORE
#define STN_GX_TMA_STORE 1
#endif
        p.gx_tma_store = 0;
        if (STN_GX_TMA_STORE && p.gx_vec4 && tw <= 256 && tr <= 256 && ((size_t)tr * tw * sizeof(float)) % 128 == 0 &&
            (long long)p.H * p.W * 4 < (1LL << 40))
            p.gx_tma_store = encode_gx_map(&gx_map, p.gx, (p.N / p.K) * p.C, p.H, p.W, tr, tw) ? 1 : 0;
        const long long n_gx = (long long)(p.N / p.K) * p.gx_ctas_per_frame;
        if (n_gx > 0x3fffffffLL) return set_error("crop_bwd: too many gx CTAs (%lld)", n_gx);
        gx_ctas = ((n_gx + cs - 1) / cs) * cs;                                  // cluster boundaries stay on role boundaries
        p.gx_zero_bytes = p.gx_tma_store ? (int)(sizeof(float) * (size_t)tr * tw) : 0;
        smem += (size_t)p.gx_tile_bytes + p.gx_zero_bytes + sizeof(ScatterGeom) * (size_t)p.K;
    }
    p.gx_ctas = (int)gx_ctas;
    p.theta_ctas = (int)theta_ctas;
    p.theta_first = theta_first_enabled() ? 1 : 0;
    const long long ctas = theta_ctas + gx_ctas;
    if (ctas > 0x7fffffffLL) return set_error("crop_bwd: too many CTAs (%lld)", ctas);
    if (smem > 200 * 1024) return set_error("crop_bwd: %d crops per frame need %zu B of shared memory (max 200 KiB)", p.K, smem);
    cudaError_t e;
    if (!p.gx && theta_only_kernel_enabled())
        e = p.nhwc ? launch_theta_tt<Nhwc4, 3, true>(p, (unsigned)ctas, cs, smem, stream)
          : gy_dtype == 0 ? launch_theta_t<float>(p, cgsel, (unsigned)ctas, cs, smem, stream)
                          : launch_theta_t<__nv_bfloat16>(p, cgsel, (unsigned)ctas, cs, smem, stream);
    else
        e = p.nhwc ? launch_bwd_tt<Nhwc4, 3, true>(p, gx_map, (unsigned)ctas, cs, smem, stream)
          : gy_dtype == 0 ? launch_bwd_t<float>(p, gx_map, cgsel, (unsigned)ctas, cs, smem, stream)
                          : launch_bwd_t<__nv_bfloat16>(p, gx_map, cgsel, (unsigned)ctas, cs, smem, stream);
    count_launch();
    note_kernel(!p.gx && theta_only_kernel_enabled() ? "stn_bwd_theta_kernel" : "stn_bwd_kernel");
    if (e != cudaSuccess) return set_error("crop_bwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

#ifdef STN_BAND_TRACE
extern "C" int loans_stn_debug_bwd_trace(void *buf)
{
    long long *q = reinterpret_cast<long long *>(buf);
    return cudaMemcpyToSymbol(g_bwd_trace, &q, sizeof(q)) == cudaSuccess ? 0 : 1;
}
#endif

}  // namespace stn
