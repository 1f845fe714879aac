// stn_unfused.cu -- the three reference operators as separate nodes, for strict drop-in use where the
// fused kernel cannot be used (a grid that did not come from our grid node, or only one of the calls
// replaced):
//   rotation_dropout forward/backward      functions/rotation_droput.py:26-48
//   spatial_transformer_grid fwd/bwd       call site sheep/sheep_localizer.py:62
//   spatial_transformer_sampler bwd on an arbitrary grid (forward reuses stn_fwd_kernel<.,.,true>)
//                                          call site sheep/sheep_localizer.py:63
#include "stn_common.cuh"

namespace stn {

__global__ void __launch_bounds__(kThreads) rotation_dropout_kernel(const float *in, float mask01, float *out, int n6)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i < n6) {
        const int e = i % 6;
        const float v = __ldg(in + i);
        out[i] = (e == 1 || e == 3) ? f_mul(v, mask01) : v;
    }
}

int launch_rotation_dropout(const float *in, float mask01, float *out, int n, cudaStream_t stream)
{
    if (n == 0) return 0;
    const int n6 = n * 6;
    rotation_dropout_kernel<<<(n6 + kThreads - 1) / kThreads, kThreads, 0, stream>>>(in, mask01, out, n6);
    count_launch();
    note_kernel("rotation_dropout_kernel");
    return check_launch("rotation_dropout");
}

__global__ void __launch_bounds__(kThreads) grid_fwd_kernel(const float *theta, float *grid, int N, int oH, int oW,
                                                            double xstep, double ystep, int ctas_per_crop, int px_per_cta)
{
    extern __shared__ float smem[];
    float *xs = smem, *ys = smem + oW;
    fill_axis_tables(xs, ys, oW, oH, xstep, ystep);
    __syncthreads();
    const int n = blockIdx.x / ctas_per_crop, tile = blockIdx.x - n * ctas_per_crop;
    const int npx = oH * oW;
    const Theta th = load_theta_masked(theta + 6 * (size_t)n, 1.0f);
    const int q_end = min(npx, (tile + 1) * px_per_cta);
    float *g = grid + (size_t)n * 2 * npx;
    for (int q = tile * px_per_cta + threadIdx.x; q < q_end; q += kThreads) {
        const int i = q / oW, j = q - i * oW;
        g[q] = grid_elem(th.t00, th.t01, th.t02, xs[j], ys[i]);
        g[npx + q] = grid_elem(th.t10, th.t11, th.t12, xs[j], ys[i]);
    }
}

int launch_grid_fwd(const float *theta, float *grid, int n, int oh, int ow, cudaStream_t stream)
{
    if (n == 0) return 0;
    const int npx = oh * ow;
    const int px_per_cta = 4 * kThreads;
    const int cpc = (npx + px_per_cta - 1) / px_per_cta;
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    grid_fwd_kernel<<<(unsigned)((long long)n * cpc), kThreads, sizeof(float) * (ow + oh), stream>>>(
        theta, grid, n, oh, ow, xstep, ystep, cpc, px_per_cta);
    count_launch();
    note_kernel("grid_fwd_kernel");
    return check_launch("grid_fwd");
}

// gtheta[n][r][:] = sum_ij ggrid[n][r][i][j] * {xs[j], ys[i], 1};  one CTA per (crop, row of theta)
__global__ void __launch_bounds__(kThreads) grid_bwd_kernel(const float *ggrid, float *gtheta, int oH, int oW,
                                                            double xstep, double ystep)
{
    extern __shared__ float smem[];
    float *xs = smem, *ys = smem + oW;
    __shared__ float red[kWarps][3];
    fill_axis_tables(xs, ys, oW, oH, xstep, ystep);
    __syncthreads();
    const int npx = oH * oW;
    const float *g = ggrid + (size_t)blockIdx.x * npx;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    for (int q = threadIdx.x; q < npx; q += kThreads) {
        const int i = q / oW, j = q - i * oW;
        const float v = __ldg(g + q);
        s0 = fmaf(v, xs[j], s0);
        s1 = fmaf(v, ys[i], s1);
        s2 += v;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane == 0) { red[warp][0] = s0; red[warp][1] = s1; red[warp][2] = s2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float tot = 0.f;
#pragma unroll
        for (int wi = 0; wi < kWarps; ++wi) tot += red[wi][threadIdx.x];
        gtheta[(size_t)blockIdx.x * 3 + threadIdx.x] = tot;
    }
}

int launch_grid_bwd(const float *ggrid, float *gtheta, int n, int oh, int ow, cudaStream_t stream)
{
    if (n == 0) return 0;
    const double xstep = ow > 1 ? 2.0 / (ow - 1) : 0.0, ystep = oh > 1 ? 2.0 / (oh - 1) : 0.0;
    grid_bwd_kernel<<<(unsigned)(2 * n), kThreads, sizeof(float) * (ow + oh), stream>>>(ggrid, gtheta, oh, ow, xstep, ystep);
    count_launch();
    note_kernel("grid_bwd_kernel");
    return check_launch("grid_bwd");
}

// ---- sampler backward on an arbitrary grid: ggrid per pixel + gx by scatter-add.
// The four tap offsets of a pixel are shared by all its channels, so duplicates inside a warp (neighbouring
// crop pixels whose windows overlap: up-sampling) are found once per tap with __match_any_sync and their
// values pre-added with shuffles; one lane per distinct address issues the red.global.add.f32.
__device__ __forceinline__ void warp_aggregated_add(float *base, int off, bool valid, float val, unsigned peers, int rounds)
{
    if (rounds <= 1) {                    // warp-uniform: the common down-sampling case, no duplicates at all
        if (valid) atomicAdd(base + off, val);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    unsigned m = peers;
    float sum = 0.f;
    // `rounds` = largest peer group in the warp; in round r every lane fetches the value of its r-th peer
    for (int r = 0; r < rounds; ++r) {
        const int src = m ? __ffs(m) - 1 : lane;
        const float v = __shfl_sync(0xffffffffu, val, src);
        if (m) sum += v;
        m &= m - 1;
    }
    if (valid && lane == leader) atomicAdd(base + off, sum);
}

template <typename GT, int CG>
__global__ void __launch_bounds__(kThreads) sampler_bwd_kernel(const CropParams p)
{
    const int n = blockIdx.x / p.ctas_per_crop;
    const int tile = blockIdx.x - n * p.ctas_per_crop;
    const int npx = p.oH * p.oW;
    const int q_begin = tile * p.px_per_cta;
    const int q_end = min(npx, q_begin + p.px_per_cta);
    const size_t plane = (size_t)p.H * p.W;
    const size_t frame = (size_t)(n / p.K) * p.C * plane;
    const float *xb = p.x + frame;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * p.C * npx;
    const size_t gbase = (size_t)n * 2 * npx;
    // uniform trip count so that the warp-wide primitives below always see all 32 lanes
    for (int q0 = q_begin; q0 < q_end; q0 += kThreads) {
        const int q = q0 + threadIdx.x;
        const bool live = q < q_end;
        const int qq = live ? q : q_begin;
        const Tap t = make_tap(__ldg(p.grid_in + gbase + qq), __ldg(p.grid_in + gbase + npx + qq), p.H, p.W);
        const TapAddr a = make_tap_addr(t, p.H, p.W);
        const bool ok[4] = {live && a.r0 && a.c0, live && a.r0 && a.c1, live && a.r1 && a.c0, live && a.r1 && a.c1};
        const int off[4] = {a.o00, a.o00 + 1, a.o00 + p.W, a.o00 + p.W + 1};
        unsigned peers[4] = {0, 0, 0, 0};
        int rounds[4] = {1, 1, 1, 1};
        if (p.gx) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                // dead taps get a unique negative key so they never pair up with a live one
                const int key = ok[k] ? off[k] : -1 - (int)(threadIdx.x & 31);
                peers[k] = __match_any_sync(0xffffffffu, key);
                rounds[k] = __reduce_max_sync(0xffffffffu, __popc(peers[k]));
            }
        }
        float su = 0.f, sv = 0.f;
        for (int c0 = 0; c0 < p.C; c0 += CG) {
            float v[CG][4], g[CG];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C) {
                    load_taps(xb + (size_t)(c0 + ch) * plane, a, p.W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
                    g[ch] = Elem<GT>::load(gyb, (size_t)(c0 + ch) * npx + qq);
                }
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < p.C) {
                    float gu, gv;
                    grad_uv(t, v[ch][0], v[ch][1], v[ch][2], v[ch][3], gu, gv);
                    gu = f_mul(gu, g[ch]);
                    gv = f_mul(gv, g[ch]);
                    if (c0 + ch == 0) { su = gu; sv = gv; }
                    else { su = f_add(su, gu); sv = f_add(sv, gv); }
                    if (p.gx) {
                        float *gxc = p.gx + frame + (size_t)(c0 + ch) * plane;
                        // gy * wu * wv, reference order (sampler _backward scatter_add)
                        warp_aggregated_add(gxc, off[0], ok[0], f_mul(f_mul(g[ch], t.wu1), t.wv1), peers[0], rounds[0]);
                        warp_aggregated_add(gxc, off[1], ok[1], f_mul(f_mul(g[ch], t.wu0), t.wv1), peers[1], rounds[1]);
                        warp_aggregated_add(gxc, off[2], ok[2], f_mul(f_mul(g[ch], t.wu1), t.wv0), peers[2], rounds[2]);
                        warp_aggregated_add(gxc, off[3], ok[3], f_mul(f_mul(g[ch], t.wu0), t.wv0), peers[3], rounds[3]);
                    }
                }
        }
        if (p.ggrid_out && live) {
            finish_grad_uv(t, p.H, p.W, su, sv);
            p.ggrid_out[gbase + q] = su;
            p.ggrid_out[gbase + npx + q] = sv;
        }
    }
}

int launch_sampler_bwd(CropParams p, int gy_dtype, cudaStream_t stream)
{
    if (p.N == 0) return 0;
    if (p.gx) {
        cudaError_t e = cudaMemsetAsync(p.gx, 0, sizeof(float) * (size_t)(p.N / p.K) * p.C * p.H * p.W, stream);
        if (e != cudaSuccess) return set_error("sampler_bwd: zero-fill of gx failed: %s", cudaGetErrorString(e));
    }
    const long long npx = (long long)p.oH * p.oW;
    p.px_per_cta = 2 * kThreads;
    p.ctas_per_crop = (int)((npx + p.px_per_cta - 1) / p.px_per_cta);
    const long long ctas = (long long)p.N * p.ctas_per_crop;
    if (ctas > 0x7fffffffLL) return set_error("sampler_bwd: too many CTAs (%lld)", ctas);
    const int cgsel = p.C == 1 ? 1 : (p.C % 3 == 0 ? 3 : 4);
    const dim3 grid((unsigned)ctas);
    if (gy_dtype == 0) {
        if (cgsel == 1) sampler_bwd_kernel<float, 1><<<grid, kThreads, 0, stream>>>(p);
        else if (cgsel == 3) sampler_bwd_kernel<float, 3><<<grid, kThreads, 0, stream>>>(p);
        else sampler_bwd_kernel<float, 4><<<grid, kThreads, 0, stream>>>(p);
    } else {
        if (cgsel == 1) sampler_bwd_kernel<__nv_bfloat16, 1><<<grid, kThreads, 0, stream>>>(p);
        else if (cgsel == 3) sampler_bwd_kernel<__nv_bfloat16, 3><<<grid, kThreads, 0, stream>>>(p);
        else sampler_bwd_kernel<__nv_bfloat16, 4><<<grid, kThreads, 0, stream>>>(p);
    }
    count_launch();
    note_kernel("sampler_bwd_kernel");
    return check_launch("sampler_bwd");
}

}  // namespace stn
