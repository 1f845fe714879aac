// stn_theta_role.cuh -- the theta-gradient role of the backward kernels (shared by the general and the axis-aligned
// backward): a thread-block CLUSTER per crop; each CTA walks its share of the crop pixels (taps gathered straight
// from global memory through the read-only path, gy, upstream grid gradient), reduces the six sums of
// gtheta = ggrid . [xs; ys; 1]^T with warp shuffles and shared memory, and rank 0 of the cluster adds the per-CTA
// partials through distributed shared memory in a fixed order: deterministic, no atomics, no workspace.
#pragma once
#include <cooperative_groups.h>

#include "stn_common.cuh"

namespace stn {

// per-thread walk over a CTA's share of the crop pixels without divisions in the loop
struct PxWalk {
    int q, i, j;
    int di, dj, ow;
    __device__ __forceinline__ PxWalk(int q0, int ow_) : q(q0), ow(ow_)
    {
        i = q0 / ow_;
        j = q0 - i * ow_;
        di = kThreads / ow_;
        dj = kThreads - di * ow_;
    }
    __device__ __forceinline__ void next()
    {
        q += kThreads; i += di; j += dj;
        if (j >= ow) { j -= ow; ++i; }
    }
};

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

struct alignas(16) BwdSmem {
    float red[kWarps][6];
    float part[6];        // this CTA's partial gtheta sums, read by cluster rank 0 through DSMEM
    int flags[2];
    int bc[4];            // band kernels: the crop's verdict (ok, P, Q), computed once per CTA
    // push-model reduction (band kernels): the peers of a cluster store their partials into rank 0's peer[] over DSMEM
    // and arrive on rank 0's mbarrier; only rank 0 waits
    unsigned long long bar;
    float peer[8][6];
};

// The six per-thread partial sums of gtheta -> gtheta[n]: warp shuffles, shared memory, then cluster rank 0 adds the
// per-CTA partials through distributed shared memory in rank order (deterministic, no atomics, no workspace).
// Every CTA of the cluster must call it.
__device__ __forceinline__ void reduce_gtheta(const CropParams &p, const float (&s)[6], BwdSmem &sm, int n, int rank, int cs)
{
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
    // gradient arriving on the four corner points of the grid (loans_stn_crop_bwd_corners): gtheta += gc . [xs; ys; 1]^T
    auto corner_term = [&](int k) {
        if (!p.gcorners) return 0.f;
        const float *gc = p.gcorners + 8 * (size_t)n + 4 * (k / 3);
        const int m = k % 3;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ci = q >> 1, cj = q & 1;
            const float w = m == 0 ? lin_x_at(p, cj ? p.oW - 1 : 0) : (m == 1 ? lin_y_at(p, ci ? p.oH - 1 : 0) : 1.0f);
            acc = fmaf(__ldg(gc + q), w, acc);
        }
        return acc;
    };
    if (cs > 1) {
        cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
        cl.sync();                                         // every CTA's part[] is written and visible
        if (rank == 0 && threadIdx.x < 6) {
            float tot = 0.f;
            for (int r = 0; r < cs; ++r) tot += *cl.map_shared_rank(&sm.part[threadIdx.x], r);
            tot += corner_term(threadIdx.x);
            // backward of the rotation mask: [0,1] and [1,0] are scaled (functions/rotation_droput.py:48)
            if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
            out[threadIdx.x] = tot;
        }
        cl.sync();                                         // peers keep their shared memory until rank 0 has read it
    } else if (threadIdx.x < 6) {
        float tot = sm.part[threadIdx.x] + corner_term(threadIdx.x);
        if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
        out[threadIdx.x] = tot;
    }

}

// ILP2: two crop pixels in flight per thread (needs the channel count at compile time and ~80 registers: the theta-only
// kernel of frames that take no gradient; inside the 64-register two-role kernel it measured slower)
// ---- push-model reduction.  reduce_gtheta() makes every CTA of the cluster wait twice for all its peers (rank 0 PULLS the
// partials, so the peers' shared memory has to stay alive); in a kernel of many waves that wait holds SM slots.  Here the
// peers PUSH: six threads each store one partial sum into rank 0's shared memory (st.shared::cluster) and arrive on rank 0's
// mbarrier with release semantics at cluster scope, then the CTA is done; rank 0 alone waits (acquire).  Nobody touches a
// peer's shared memory, so peers may exit at once.  push_reduce_init() must run at kernel entry in every CTA of the cluster.
__device__ __forceinline__ void push_reduce_init(BwdSmem &sm, int rank, int cs)
{
    if (cs <= 1) return;
    if (rank == 0 && threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&sm.bar)), "r"(6 * (cs - 1)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cooperative_groups::this_cluster().sync();            // the barrier exists before any peer can arrive on it
}

__device__ __forceinline__ void reduce_gtheta_push(const CropParams &p, const float (&s)[6], BwdSmem &sm, int n, int rank, int cs)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const float r = warp_sum(s[k]);
        if (lane == 0) sm.red[warp][k] = r;
    }
    __syncthreads();
    if (threadIdx.x >= 6) return;
    float tot = 0.f;
#pragma unroll
    for (int wi = 0; wi < kWarps; ++wi) tot += sm.red[wi][threadIdx.x];
    if (cs > 1 && rank != 0) {
        unsigned dst, bar;
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(dst) : "r"((unsigned)__cvta_generic_to_shared(&sm.peer[rank][threadIdx.x])));
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(bar) : "r"((unsigned)__cvta_generic_to_shared(&sm.bar)));
        asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(dst), "f"(tot) : "memory");
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar) : "memory");
        return;
    }
    if (cs > 1) {
        const unsigned bar = (unsigned)__cvta_generic_to_shared(&sm.bar);
        unsigned done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
        for (int r = 1; r < cs; ++r) tot += sm.peer[r][threadIdx.x];                  // fixed order: deterministic
    }
    if (p.gcorners) {                                                                   // see reduce_gtheta
        const float *gc = p.gcorners + 8 * (size_t)n + 4 * (threadIdx.x / 3);
        const int m = threadIdx.x % 3;
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int ci = q >> 1, cj = q & 1;
            const float w = m == 0 ? lin_x_at(p, cj ? p.oW - 1 : 0) : (m == 1 ? lin_y_at(p, ci ? p.oH - 1 : 0) : 1.0f);
            acc = fmaf(__ldg(gc + q), w, acc);
        }
        tot += acc;
    }
    if (threadIdx.x == 1 || threadIdx.x == 3) tot = f_mul(tot, p.mask01);
    p.gtheta[6 * (size_t)n + threadIdx.x] = tot;
}

template <typename GT, int CG, bool EXACT, bool GRAY = false, bool ILP2 = false>
__device__ __forceinline__ void theta_role(const CropParams &p, const float *xs, const float *ys, BwdSmem &sm, int cta)
{
    const int C = EXACT ? CG : p.C;
    const int cs = p.ctas_per_crop;
    const int n = cta / cs;
    const int rank = cta - n * cs;
    const int npx = p.oH * p.oW;
    const int q_end = min(npx, (rank + 1) * p.px_per_cta);
    const Theta th = load_theta_masked(p.theta + 6 * (size_t)n, p.mask01);
    const int plane = p.H * p.W;
    const float *xb = p.x + (size_t)(n / p.K) * C * plane;
    const GT *gyb = reinterpret_cast<const GT *>(p.gy) + (size_t)n * crop_planes<GT, GRAY>(C) * npx;
    float *ggo = p.ggrid_out ? p.ggrid_out + (size_t)n * 2 * npx : nullptr;
    const float *ggu = p.ggrid_up ? p.ggrid_up + (size_t)n * 2 * npx : nullptr;

    float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#ifndef STN_THETA_ILP
#define STN_THETA_ILP 1      // 2 = two pixels in flight per thread: measured slower here (register pressure)
#endif
    if (EXACT && (ILP2 || STN_THETA_ILP == 2)) {
        // all channels in one group: two crop pixels per thread in flight (taps and gy of both requested before
        // either is reduced), one memory round trip per pair of pixels
        struct Px {
            Tap t;
            float v[CG][4], g[CG];
            float xsj, ysi;
            int q;
            bool live;
        };
        auto prepare = [&](Px &px, const PxWalk &w) {
            px.live = w.q < q_end;
            px.q = w.q;
            if (!px.live) return;
            px.xsj = xs[w.j]; px.ysi = ys[w.i];
            px.t = make_tap(grid_elem(th.t00, th.t01, th.t02, px.xsj, px.ysi),
                            grid_elem(th.t10, th.t11, th.t12, px.xsj, px.ysi), p.H, p.W);
            const TapAddr a = make_tap_addr(px.t, p.H, p.W);
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                load_taps(xb + ch * plane, a, p.W, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3]);
                px.g[ch] = load_gy<GT, GRAY>(gyb + w.q, ch, npx);
            }
        };
        auto finish = [&](const Px &px) {
            if (!px.live) return;
            float su = 0.f, sv = 0.f;
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                float gu, gv;
                grad_uv(px.t, px.v[ch][0], px.v[ch][1], px.v[ch][2], px.v[ch][3], gu, gv);
                gu = f_mul(gu, px.g[ch]);
                gv = f_mul(gv, px.g[ch]);
                if (ch == 0) { su = gu; sv = gv; }
                else { su = f_add(su, gu); sv = f_add(sv, gv); }              // numpy.sum over the channel axis
            }
            finish_grad_uv(px.t, p.H, p.W, su, sv);
            if (ggo) {
                ggo[px.q] = su;
                ggo[npx + px.q] = sv;
            }
            if (ggu) {
                su = f_add(su, __ldg(ggu + px.q));
                sv = f_add(sv, __ldg(ggu + npx + px.q));
            }
            s[0] = fmaf(su, px.xsj, s[0]); s[1] = fmaf(su, px.ysi, s[1]); s[2] += su;
            s[3] = fmaf(sv, px.xsj, s[3]); s[4] = fmaf(sv, px.ysi, s[4]); s[5] += sv;
        };
        PxWalk w(rank * p.px_per_cta + threadIdx.x, p.oW);
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
    for (PxWalk w(rank * p.px_per_cta + threadIdx.x, p.oW); w.q < q_end; w.next()) {
        const float xsj = xs[w.j], ysi = ys[w.i];
        const Tap t = make_tap(grid_elem(th.t00, th.t01, th.t02, xsj, ysi),
                               grid_elem(th.t10, th.t11, th.t12, xsj, ysi), p.H, p.W);
        const TapAddr a = make_tap_addr(t, p.H, p.W);
        float su = 0.f, sv = 0.f;
        const float *xc = xb;
        const GT *gc = gyb + w.q;
        for (int c0 = 0; c0 < C; c0 += CG) {
            float v[CG][4], g[CG];
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < C) {
                    load_taps(xc + ch * plane, a, p.W, v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
                    g[ch] = load_gy<GT, GRAY>(gc, ch, npx);
                }
#pragma unroll
            for (int ch = 0; ch < CG; ++ch)
                if (c0 + ch < C) {
                    float gu, gv;
                    grad_uv(t, v[ch][0], v[ch][1], v[ch][2], v[ch][3], gu, gv);
                    gu = f_mul(gu, g[ch]);
                    gv = f_mul(gv, g[ch]);
                    if (c0 + ch == 0) { su = gu; sv = gv; }
                    else { su = f_add(su, gu); sv = f_add(sv, gv); }          // numpy.sum over the channel axis
                }
            xc += CG * plane;
            gc += CG * npx;
        }
        finish_grad_uv(t, p.H, p.W, su, sv);
        if (ggo) {
            ggo[w.q] = su;
            ggo[npx + w.q] = sv;
        }
        if (ggu) {
            su = f_add(su, __ldg(ggu + w.q));
            sv = f_add(sv, __ldg(ggu + npx + w.q));
        }
        s[0] = fmaf(su, xsj, s[0]); s[1] = fmaf(su, ysi, s[1]); s[2] += su;
        s[3] = fmaf(sv, xsj, s[3]); s[4] = fmaf(sv, ysi, s[4]); s[5] += sv;
    }
    }
    reduce_gtheta(p, s, sm, n, rank, cs);
}


}  // namespace stn
