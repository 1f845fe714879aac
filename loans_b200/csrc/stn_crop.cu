// stn_crop.cu -- the fused hot path: rotation_dropout -> affine grid -> bilinear sampler, forward and
// backward, one kernel launch per direction (reference sheep/sheep_localizer.py:61-63).
//
// Forward  (stn_fwd_kernel): one thread per crop pixel computes its own sample coordinates from the masked
//   theta (no grid round trip through HBM), gathers the 4 taps of every channel through the read-only
//   path with all loads of a channel group in flight together, and stores the crop (fp32 or bf16) and,
//   if asked, the grid -- a required output of the reference API -- from the same registers.
//
// Backward (stn_bwd_kernel): one launch, two CTA roles.
//   * theta role, CTAs [0, theta_ctas): a thread-block CLUSTER per crop.  Each CTA walks its share of the
//     crop pixels (taps + gy -> d/du, d/dv, plus the upstream grid gradient), reduces the six sums of
//     gtheta = ggrid . [xs; ys; 1]^T with warp shuffles and shared memory, and rank 0 of the cluster adds
//     the per-CTA partials through distributed shared memory in a fixed order: deterministic, no atomics,
//     no zero-initialised output, no workspace.
//   * gx role, the remaining CTAs: the image gradient as a GATHER.  Each thread owns one frame pixel,
//     inverts the affine map to find the few crop pixels whose 2x2 taps cover it (stn_math.cuh), and
//     writes the dense gx exactly once with coalesced stores, zeros included -- no memset pass, no float
//     atomics, bit-reproducible.  With K crops per frame the K contributions are summed in registers.
#include <cooperative_groups.h>

#include "stn_common.cuh"

namespace cg = cooperative_groups;

namespace stn {

// ------------------------------------------------------------------------------------------ forward
template <typename YT, int CG, bool FROM_GRID>
__global__ void __launch_bounds__(kThreads) stn_fwd_kernel(const CropParams p)
{
    extern __shared__ float smem[];
    float *xs = smem, *ys = smem + p.oW;
    if (!FROM_GRID) {
        fill_axis_tables(xs, ys, p.oW, p.oH, p.xstep, p.ystep);
        __syncthreads();
    }
    const int n = blockIdx.x / p.ctas_per_crop;
    const int tile = blockIdx.x - n * p.ctas_per_crop;
    const int npx = p.oH * p.oW;
    const int q_end = min(npx, (tile + 1) * p.px_per_cta);
    Theta th = {};
    if (!FROM_GRID) th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    const size_t plane = (size_t)p.H * p.W;
    const float *xb = p.x + (size_t)(n / p.K) * p.C * plane;
    YT *yb = reinterpret_cast<YT *>(p.y) + (size_t)n * p.C * npx;
    const size_t gbase = (size_t)n * 2 * npx;

    for (int q = tile * p.px_per_cta + threadIdx.x; q < q_end; q += kThreads) {
        float g0, g1;
        if (FROM_GRID) {
            g0 = __ldg(p.grid_in + gbase + q);
            g1 = __ldg(p.grid_in + gbase + npx + q);
        } else {
            const int i = q / p.oW, j = q - i * p.oW;
            const float xsj = xs[j], ysi = ys[i];
            g0 = grid_elem(th.t00, th.t01, th.t02, xsj, ysi);
            g1 = grid_elem(th.t10, th.t11, th.t12, xsj, ysi);
            if (p.grid_out) {
                p.grid_out[gbase + q] = g0;
                p.grid_out[gbase + npx + q] = g1;
            }
        }
        const Tap t = make_tap(g0, g1, p.H, p.W);
        const TapAddr a = make_tap_addr(t, p.H, p.W);
        const Weights4 w = make_weights(t);
        for (int c0 = 0; c0 < p.C; c0 += CG) {
            float v[CG][4];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C)
                    load_taps(xb + (size_t)(c0 + ch) * plane, a, p.W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C)
                    Elem<YT>::store(yb, (size_t)(c0 + ch) * npx + q, interp(w, v[ch][0], v[ch][1], v[ch][2], v[ch][3]));
        }
    }
}

// ------------------------------------------------------------------------------------------ backward
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct BwdSmem {
    float red[kWarps][6];
    float part[6];        // this CTA's partial gtheta sums, read by cluster rank 0 through DSMEM
};

template <typename GT, int CG>
__device__ __forceinline__ void theta_role(const CropParams &p, const float *xs, const float *ys, BwdSmem &sm)
{
    const int cs = p.ctas_per_crop;
    const int n = blockIdx.x / cs;
    const int rank = blockIdx.x - n * cs;
    const int npx = p.oH * p.oW;
    const int q_end = min(npx, (rank + 1) * p.px_per_cta);
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    const size_t plane = (size_t)p.H * p.W;
    const float *xb = p.x + (size_t)(n / p.K) * p.C * plane;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * p.C * npx;
    const size_t gbase = (size_t)n * 2 * npx;

    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int q = rank * p.px_per_cta + threadIdx.x; q < q_end; q += kThreads) {
        const int i = q / p.oW, j = q - i * p.oW;
        const float xsj = xs[j], ysi = ys[i];
        const Tap t = make_tap(grid_elem(th.t00, th.t01, th.t02, xsj, ysi),
                               grid_elem(th.t10, th.t11, th.t12, xsj, ysi), p.H, p.W);
        const TapAddr a = make_tap_addr(t, p.H, p.W);
        float su = 0.f, sv = 0.f;
        for (int c0 = 0; c0 < p.C; c0 += CG) {
            float v[CG][4], g[CG];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C) {
                    load_taps(xb + (size_t)(c0 + ch) * plane, a, p.W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
                    g[ch] = Elem<GT>::load(gyb, (size_t)(c0 + ch) * npx + q);
                }
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C) {
                    float gu, gv;
                    grad_uv(t, v[ch][0], v[ch][1], v[ch][2], v[ch][3], gu, gv);
                    gu = f_mul(gu, g[ch]);
                    gv = f_mul(gv, g[ch]);
                    if (c0 + ch == 0) { su = gu; sv = gv; }
                    else { su = f_add(su, gu); sv = f_add(sv, gv); }          // numpy.sum over the channel axis
                }
        }
        finish_grad_uv(t, p.H, p.W, su, sv);
        if (p.ggrid_out) {
            p.ggrid_out[gbase + q] = su;
            p.ggrid_out[gbase + npx + q] = sv;
        }
        if (p.ggrid_up) {
            su = f_add(su, __ldg(p.ggrid_up + gbase + q));
            sv = f_add(sv, __ldg(p.ggrid_up + gbase + npx + q));
        }
        s[0] = fmaf(su, xsj, s[0]); s[1] = fmaf(su, ysi, s[1]); s[2] += su;
        s[3] = fmaf(sv, xsj, s[3]); s[4] = fmaf(sv, ysi, s[4]); s[5] += sv;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float r = warp_sum(s[k]);
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
        cl.sync();                                         // every CTA's part[] is written and visible
        if (rank == 0 && threadIdx.x < 6) {
            float tot = 0.f;
            for (int r = 0; r < cs; ++r) tot += *cl.map_shared_rank(&sm.part[threadIdx.x], r);
            // backward of the rotation mask: [0,1] and [1,0] are scaled (functions/rotation_droput.py:48)
            if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
            out[threadIdx.x] = tot;
        }
        cl.sync();                                         // peers keep their shared memory until rank 0 has read it
    } else if (threadIdx.x < 6) {
        float tot = sm.part[threadIdx.x];
        if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
        out[threadIdx.x] = tot;
    }
}

template <typename GT>
struct GyLoader {
    __device__ __forceinline__ float operator()(const GT *p, size_t i) const { return Elem<GT>::load(p, i); }
};

template <typename GT, int CG>
__device__ __forceinline__ void gx_role(const CropParams &p, const float *xs, const float *ys, InvCrop *inv)
{
    const int gx_ctas = (int)gridDim.x - p.theta_ctas;
    const int me = (int)blockIdx.x - p.theta_ctas;
    const int B = p.N / p.K;
    const int total = B * p.gx_tiles_per_frame;
    const int per = (total + gx_ctas - 1) / gx_ctas;       // contiguous tile range: few frame switches per CTA
    const int t_begin = me * per, t_end = min(total, t_begin + per);
    const int npx = p.oH * p.oW, fpx = p.H * p.W;
    const GT *gy = reinterpret_cast<const GT *>(p.gy);
    int cur_b = -1;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int b = tile / p.gx_tiles_per_frame;
        const int chunk = tile - b * p.gx_tiles_per_frame;
        if (b != cur_b) {
            __syncthreads();
            for (int kk = threadIdx.x; kk < p.K; kk += kThreads)
                inv[kk] = make_inv_crop(load_theta_masked(p.theta + 6 * ((size_t)b * p.K + kk), p.mask01),
                                        p.H, p.W, p.oH, p.oW);
            __syncthreads();
            cur_b = b;
        }
        const int q_end = min(fpx, (chunk + 1) * p.gx_tile_px);
        for (int q = chunk * p.gx_tile_px + threadIdx.x; q < q_end; q += kThreads) {
            const int r = q / p.W, s = q - r * p.W;
            for (int c0 = 0; c0 < p.C; c0 += CG) {
                const int nc = min(CG, p.C - c0);
                float acc[CG];
#pragma unroll
                for (int ch = 0; ch < CG; ++ch) acc[ch] = 0.f;
                for (int kk = 0; kk < p.K; ++kk)
                    gather_from_crop<CG>(inv[kk], xs, ys, p.H, p.W, p.oH, p.oW, r + 1, s + 1,
                                         gy + ((size_t)(b * p.K + kk) * p.C + c0) * npx, nc, GyLoader<GT>(), acc);
#pragma unroll
                for (int ch = 0; ch < CG; ++ch)
                    if (ch < nc) p.gx[((size_t)b * p.C + c0 + ch) * fpx + q] = acc[ch];
            }
        }
    }
}

template <typename GT, int CG>
__global__ void __launch_bounds__(kThreads) stn_bwd_kernel(const CropParams p)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    BwdSmem &sm = *reinterpret_cast<BwdSmem *>(smem_raw);
    float *xs = reinterpret_cast<float *>(smem_raw + sizeof(BwdSmem));
    float *ys = xs + p.oW;
    InvCrop *inv = reinterpret_cast<InvCrop *>(ys + p.oH + ((p.oW + p.oH) & 1));    // 8-byte aligned
    fill_axis_tables(xs, ys, p.oW, p.oH, p.xstep, p.ystep);
    __syncthreads();
    if ((int)blockIdx.x < p.theta_ctas) theta_role<GT, CG>(p, xs, ys, sm);
    else gx_role<GT, CG>(p, xs, ys, inv);
}

// ------------------------------------------------------------------------------------------ host launchers
static int pick_channel_group(int C) { return C == 1 ? 1 : (C % 3 == 0 ? 3 : 4); }

template <typename YT, bool FROM_GRID>
static cudaError_t launch_fwd_t(const CropParams &p, int cgsel, dim3 grid, size_t smem, cudaStream_t s)
{
    switch (cgsel) {
    case 1: stn_fwd_kernel<YT, 1, FROM_GRID><<<grid, kThreads, smem, s>>>(p); break;
    case 3: stn_fwd_kernel<YT, 3, FROM_GRID><<<grid, kThreads, smem, s>>>(p); break;
    default: stn_fwd_kernel<YT, 4, FROM_GRID><<<grid, kThreads, smem, s>>>(p); break;
    }
    return cudaGetLastError();
}

int launch_crop_fwd(CropParams p, bool from_grid, int y_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const long long npx = (long long)p.oH * p.oW;
    // enough CTAs to give every SM a few, at most 8 pixels per thread
    long long per = (npx * p.N + 4LL * kNumSMs - 1) / (4LL * kNumSMs);
    per = ((per + kThreads - 1) / kThreads) * kThreads;
    if (per < kThreads) per = kThreads;
    if (per > 8 * kThreads) per = 8 * kThreads;
    p.px_per_cta = (int)per;
    p.ctas_per_crop = (int)((npx + per - 1) / per);
    const long long ctas = (long long)p.N * p.ctas_per_crop;
    if (ctas > 0x7fffffffLL) return set_error("crop_fwd: too many CTAs (%lld)", ctas);
    const dim3 grid((unsigned)ctas);
    const size_t smem = sizeof(float) * (size_t)(p.oW + p.oH);
    const int cgsel = pick_channel_group(p.C);
    cudaError_t e;
    if (y_dtype == 0)
        e = from_grid ? launch_fwd_t<float, true>(p, cgsel, grid, smem, stream)
                      : launch_fwd_t<float, false>(p, cgsel, grid, smem, stream);
    else
        e = from_grid ? launch_fwd_t<__nv_bfloat16, true>(p, cgsel, grid, smem, stream)
                      : launch_fwd_t<__nv_bfloat16, false>(p, cgsel, grid, smem, stream);
    count_launch();
    if (e != cudaSuccess) return set_error("crop_fwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

template <typename GT, int CG>
static cudaError_t launch_bwd_tt(const CropParams &p, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    if (smem > 48 * 1024) {      // only very large crops-per-frame counts leave the default dynamic limit
        cudaError_t e = cudaFuncSetAttribute(stn_bwd_kernel<GT, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
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
    return cudaLaunchKernelEx(&cfg, stn_bwd_kernel<GT, CG>, p);
}

template <typename GT>
static cudaError_t launch_bwd_t(const CropParams &p, int cgsel, unsigned ctas, unsigned cs, size_t smem, cudaStream_t s)
{
    switch (cgsel) {
    case 1: return launch_bwd_tt<GT, 1>(p, ctas, cs, smem, s);
    case 3: return launch_bwd_tt<GT, 3>(p, ctas, cs, smem, s);
    default: return launch_bwd_tt<GT, 4>(p, ctas, cs, smem, s);
    }
}

int launch_crop_bwd(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    const long long npx = (long long)p.oH * p.oW;
    // cluster size: enough theta-role CTAs to occupy the machine twice over, 8 (portable maximum) at most
    unsigned cs = 1;
    while (cs < 8 && (long long)p.N * cs < 2LL * kNumSMs && npx / (2 * cs) >= kThreads / 2) cs *= 2;
    p.ctas_per_crop = (int)cs;
    p.px_per_cta = (int)((npx + cs - 1) / cs);
    const long long theta_ctas = (long long)p.N * cs;
    long long gx_ctas = 0;
    if (p.gx) {
        const long long fpx = (long long)p.H * p.W;
        p.gx_tile_px = 4 * kThreads;
        p.gx_tiles_per_frame = (int)((fpx + p.gx_tile_px - 1) / p.gx_tile_px);
        const long long tiles = (long long)(p.N / p.K) * p.gx_tiles_per_frame;
        gx_ctas = tiles < 6LL * kNumSMs ? tiles : 6LL * kNumSMs;
        gx_ctas = ((gx_ctas + cs - 1) / cs) * cs;
    }
    const long long ctas = theta_ctas + gx_ctas;
    if (ctas > 0x7fffffffLL) return set_error("crop_bwd: too many CTAs (%lld)", ctas);
    p.theta_ctas = (int)theta_ctas;
    size_t smem = sizeof(BwdSmem) + sizeof(float) * (size_t)(p.oW + p.oH + 1);
    if (p.gx) smem += sizeof(InvCrop) * (size_t)p.K;
    if (smem > 200 * 1024) return set_error("crop_bwd: %d crops per frame need %zu B of shared memory (max 200 KiB)", p.K, smem);
    const int cgsel = pick_channel_group(p.C);
    cudaError_t e = gy_dtype == 0 ? launch_bwd_t<float>(p, cgsel, (unsigned)ctas, cs, smem, stream)
                                  : launch_bwd_t<__nv_bfloat16>(p, cgsel, (unsigned)ctas, cs, smem, stream);
    count_launch();
    if (e != cudaSuccess) return set_error("crop_bwd launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace stn
