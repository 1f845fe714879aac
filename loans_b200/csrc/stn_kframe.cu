// stn_kframe.cu -- gx of axis-aligned crops with SEVERAL crops per frame (BASELINE config 4: 16 jittered boxes per frame,
// the assessor feed; no Chainer equivalent -- Chainer's sampler has one grid per frame -- semantics in include/loans_stn.h:
// crop i samples frame i / k, so gx of a frame is the sum of its k crops' scatters).
//
// The general gx role (stn_gx_role.cuh) inverts the affine map per (tile, crop) pair and re-evaluates every candidate crop
// pixel with the full per-pixel coordinate chain (379 M warp instructions at config 4).  With the rotation terms masked (LoANs'
// ratio = 0.0) u depends on the crop column only and v on the crop row only, the gradient of the frame needs no taps at all --
// only gy and the weights -- and, when every crop of the frame steps by at least ~2 frame pixels per crop pixel (down-sampling,
// LoANs' regime), a frame row receives AT MOST ONE crop row of each crop, and that crop row touches each frame column at most
// once.  So:
//
//   * a CTA takes a band of frame rows of one frame and builds, once, the column and row tables of EVERY crop of the frame
//     (K x (oW + oH) entries of 8 bytes: weight and tap index / flags, bit-identical to the per-pixel chain) and the INVERSE row
//     map of its band: rowinv[crop][frame row] = (crop row, which of its two tap rows) or nothing;
//   * a WARP owns one frame row at a time (all columns, all channels) in a private shared-memory row buffer.  For each crop
//     that has a crop row on it (a table look-up, no candidate search) the warp walks that crop row in 32-column chunks:
//     gy coalesced, the column entry from the table, gy * wu * wv added into the row buffer with plain read-modify-writes --
//     lanes of one chunk never share a frame pixel, chunks and crops follow each other in program order (`__syncwarp()` only,
//     no CTA barrier after the prologue).  The gy of the next chunk is requested before the current one is added;
//   * the finished row leaves with 16-byte stores and the buffer is zeroed on the way out: gx is written exactly once, zeros
//     included, in a fixed order (crops ascending): bit-reproducible, no memset pass, no atomics.
//
// gtheta needs the taps and every crop pixel exactly once: it stays with the table-driven theta kernel
// (stn_bwd_theta_tab_kernel, stn_band.cu), launched in front of this one by launch_crop_bwd_kframe.  A frame with a crop this
// path does not take (rotation not masked, a step below ~2, mirrored or degenerate scale, non-finite theta) runs the general
// gx role here, in the same launch.
#include "stn_band_plan.cuh"
#include "stn_common.cuh"
#include "stn_gx_role.cuh"

namespace stn {

int launch_crop_bwd_theta_tab(CropParams p, int gy_dtype, cudaStream_t stream);

struct alignas(8) KfCol {             // one crop column / row: weight of tap idx0 + 1 (w1 = 1 - w0 exactly, see kf_w1), BandAxis code
    float w0;
    int code;
};

// A frame with a crop this path does not take: the whole frame through the general gx role (same result, any theta).
// xs / ys / geom alias the (not yet built) tables, not the crops.  Out of line: the general role's registers are its own.
template <typename GT, int CG, bool GRAY>
__device__ __noinline__ void kf_declined(const CropParams &p, const KfCrop *crops, float *xs, float *ys, ScatterGeom *geom,
                                         float *region, int b, int part)
{
    const int K = p.K, tid = threadIdx.x;
    fill_axis_tables(p, xs, ys);
    int fb = 0;
    for (int kk = tid; kk < K; kk += kThreads) {
        const KfCrop &c = crops[kk];
        Theta th;
        th.t00 = c.t00; th.t01 = c.t01; th.t02 = c.t02; th.t10 = c.t10; th.t11 = c.t11; th.t12 = c.t12;
        geom[kk] = make_scatter_geom(th, p.H, p.W, p.oH, p.oW);
        fb |= geom[kk].P == 0;
    }
    const int any_fb = __syncthreads_or(fb);
    gx_role<GT, CG, true, GRAY>(p, nullptr, xs, ys, any_fb != 0, geom, region, nullptr, b, part, p.band_fb_tiles_per_warp);
}

template <typename GT, int CG, bool GRAY, int NCH>
__global__ void __launch_bounds__(kThreads, 3) stn_bwd_kframe_kernel(const __grid_constant__ CropParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [row buffers: kWarps x CG x W floats] (the general role's warp tiles alias them) [crops[K]]
    //         [coltab[K * oW] | rowtab[K * oH]] [rowinv[K * rows]]   (declined frames: [xs | ys] [geom[K]] alias the tables)
    const int K = p.K, H = p.H, W = p.W, oH = p.oH, oW = p.oW;
    float *region = reinterpret_cast<float *>(smem_raw);
    unsigned char *q = smem_raw + p.kf_region_bytes;
    KfCrop *crops = reinterpret_cast<KfCrop *>(q);
    q += (sizeof(KfCrop) * K + 15) & ~(size_t)15;
    KfCol *coltab = reinterpret_cast<KfCol *>(q);
    KfCol *rowtab = coltab + (size_t)K * oW;
    short *rowinv = reinterpret_cast<short *>(rowtab + (size_t)K * oH);          // 2 * crop row + tap row, 16 bits (oH < 16384: launcher)
    float *xs = reinterpret_cast<float *>(q);
    float *ys = xs + oW;
    ScatterGeom *geom = reinterpret_cast<ScatterGeom *>(q + sizeof(float) * ((oW + oH + 3) & ~3));

    const int tid = threadIdx.x, warp = tid >> 5, ln = tid & 31;
    pdl_launch_dependents();
    // bands in the middle of a frame carry most crop rows, the ones at its top and bottom few or none: the grid walks the bands
    // centre-out (all frames' middle bands first), so that the light bands fill the tail of the last wave
    const int frames = p.N / p.K;
    const int order = blockIdx.x / frames, b = blockIdx.x - order * frames;
    const int part = (p.kf_ctas_per_frame - 1) / 2 + ((order & 1) ? (order + 1) / 2 : -(order / 2));
    const int rows = p.kf_rows_cta;
    const int r0 = part * rows, nr = min(rows, H - r0);                 // the band: frame rows [r0, r0 + nr)
    {   // zero the row buffers: needs no input, overlaps the tail of the previous kernel
        float4 *b4 = reinterpret_cast<float4 *>(region);
        for (int e = tid; e < kWarps * CG * W / 4; e += kThreads) b4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    pdl_wait();
    // ---- the frame's crops: verdict, one thread per crop
    int declined = 0;
    for (int kk = tid; kk < K; kk += kThreads) {
        const Theta th = load_theta_masked(p.theta + 6 * ((size_t)b * K + kk), p.mask01);
        crops[kk] = make_kf_crop(th, H, W, oH, oW);
        declined |= !crops[kk].ok;
    }
    if (__syncthreads_or(declined)) {
        kf_declined<GT, CG, GRAY>(p, crops, xs, ys, geom, region, b, part);
        return;
    }
    // ---- column tables of all K crops, inverse row map cleared (same operations as the per-pixel chain: make_band_axis)
    for (int e = tid; e < K * rows; e += kThreads) rowinv[e] = -1;
    for (int e = tid; e < K * oW; e += kThreads) {
        const int kk = e / oW, t = e - kk * oW;
        const KfCrop &c = crops[kk];
        const BandAxis a = make_band_axis(c.t00, c.t01, c.t02, lin_x_at(p, t), true, W);
        KfCol kc;
        kc.w0 = a.w0;
        kc.code = a.code;
        coltab[e] = kc;
    }
    __syncthreads();
    // ---- row tables, and for the rows of the band: which crop row lands on them.  A crop that steps by >= ~2 frame rows puts
    //      at most one of its tap rows on a frame row (make_kf_crop), so every slot is written at most once
    for (int e = tid; e < K * oH; e += kThreads) {
        const int kk = e / oH, i = e - kk * oH;
        const KfCrop &c = crops[kk];
        KfCol kc;
        kc.w0 = 0.f;
        kc.code = 0;
        // rows far from the band are not needed (the band's rows come from rowinv only)
        const BandAxis a = make_band_axis(c.t11, c.t10, c.t12, lin_y_at(p, i), false, H);
        kc.w0 = a.w0;
        kc.code = a.code;
        rowtab[e] = kc;
        const int t0 = (a.code & kAxIdxMask) - 1 - r0;                  // band row of tap row 0 (tap row 1: + 1)
        if ((a.code & kAxTap0) && t0 >= 0 && t0 < nr) rowinv[kk * rows + t0] = (short)(2 * i);
        if ((a.code & kAxTap1) && t0 + 1 >= 0 && t0 + 1 < nr) rowinv[kk * rows + t0 + 1] = (short)(2 * i + 1);
    }
    __syncthreads();

    const int npx = oH * oW;
    constexpr int gplanes = crop_planes<GT, GRAY>(CG);
    const GT *gy = reinterpret_cast<const GT *>(p.gy) + (size_t)b * K * gplanes * npx;
    float *rowbuf = region + warp * (CG * W);
    float *gxf = p.gx + (size_t)b * CG * ((size_t)H * W);
    const size_t fpx = (size_t)H * W;

    // one crop row that lands on the frame row at hand: gy and the column entries of all its chunks, the row weight
    struct Hit {
        float g[NCH][CG];
        float cw0[NCH];
        int code[NCH];
        float wv;
    };
    const int kgroups = (K + 31) >> 5;
    for (int rr = warp; rr < nr; rr += kWarps) {
        // the crop rows on this frame row, crops in ascending order: lane l of group gk looks up crop 32 gk + l
        auto fetch = [&](Hit &h, int kk, int cur) {
            const int i = cur >> 1;
            const KfCol rw = rowtab[kk * oH + i];
            h.wv = (cur & 1) ? rw.w0 : kf_w1(rw.w0);                    // tap row 1 carries w0, tap row 0 carries w1
            const KfCol *ct = coltab + kk * oW + ln;
            const GT *gp = gy + (size_t)kk * gplanes * npx + i * oW + ln;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                h.code[c] = 0;
                h.cw0[c] = 0.f;
                if (32 * c + ln < oW) {
                    const KfCol col = ct[32 * c];
                    h.code[c] = col.code;
                    h.cw0[c] = col.w0;
                }
                const bool any = (h.code[c] & (kAxTap0 | kAxTap1)) != 0;
#pragma unroll
                for (int ch = 0; ch < CG; ++ch) h.g[c][ch] = any ? load_gy<GT, GRAY>(gp + 32 * c, ch, npx) : 0.f;
            }
        };
        auto add = [&](const Hit &h) {
            // the chunks of one crop row never share a frame pixel (the crop steps by >= ~2): all reads, then all writes
            float v[NCH][CG][2];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                const float *p0 = rowbuf + (h.code[c] & kAxIdxMask) - 1;  // tap column u0 - 1 of channel 0
                const bool c0ok = (h.code[c] & kAxTap0) != 0, c1ok = (h.code[c] & kAxTap1) != 0;
#pragma unroll
                for (int ch = 0; ch < CG; ++ch) {
                    v[c][ch][0] = c0ok ? p0[ch * W] : 0.f;
                    v[c][ch][1] = c1ok ? p0[ch * W + 1] : 0.f;
                }
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                float *p0 = rowbuf + (h.code[c] & kAxIdxMask) - 1;
                const bool c0ok = (h.code[c] & kAxTap0) != 0, c1ok = (h.code[c] & kAxTap1) != 0;
                const float cw1 = kf_w1(h.cw0[c]);
#pragma unroll
                for (int ch = 0; ch < CG; ++ch) {
                    const float a1 = f_mul(h.g[c][ch], cw1), a0 = f_mul(h.g[c][ch], h.cw0[c]);   // gy * wu * wv, reference order
                    if (c0ok) p0[ch * W] = f_add(v[c][ch][0], f_mul(a1, h.wv));
                    if (c1ok) p0[ch * W + 1] = f_add(v[c][ch][1], f_mul(a0, h.wv));
                }
            }
            __syncwarp();                                               // the next crop's lanes may meet these frame pixels
        };
        bool touched = false;
        for (int gk = 0; gk < kgroups; ++gk) {
            const int kl = 32 * gk + ln;
            const int mine = kl < K ? rowinv[kl * rows + rr] : -1;
            unsigned hits = __ballot_sync(0xffffffffu, mine >= 0);
            if (!hits) continue;
            touched = true;
            Hit A, B;
            int la = __ffs(hits) - 1;
            hits &= hits - 1;
            fetch(A, 32 * gk + la, __shfl_sync(0xffffffffu, mine, la));
            while (true) {
                const bool haveB = hits != 0;
                if (haveB) {
                    const int lb = __ffs(hits) - 1;
                    hits &= hits - 1;
                    fetch(B, 32 * gk + lb, __shfl_sync(0xffffffffu, mine, lb));
                }
                add(A);
                if (!haveB) break;
                const bool haveA = hits != 0;
                if (haveA) {
                    la = __ffs(hits) - 1;
                    hits &= hits - 1;
                    fetch(A, 32 * gk + la, __shfl_sync(0xffffffffu, mine, la));
                }
                add(B);
                if (!haveA) break;
            }
        }
        // the finished row: 16-byte stores, the buffer zeroed on the way out
        const int r = r0 + rr;
        if (touched) {
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                float4 *s4 = reinterpret_cast<float4 *>(rowbuf + ch * W);
                float4 *g4 = reinterpret_cast<float4 *>(gxf + (size_t)ch * fpx + (size_t)r * W);
                for (int e = ln; e < W / 4; e += 32) {
                    const float4 v = s4[e];
                    s4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
                    g4[e] = v;
                }
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int ch = 0; ch < CG; ++ch) {
                float4 *g4 = reinterpret_cast<float4 *>(gxf + (size_t)ch * fpx + (size_t)r * W);
                for (int e = ln; e < W / 4; e += 32) g4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
}

template <typename GT, int CG, bool GRAY, int NCH>
static cudaError_t launch_kframe_ttt(const CropParams &p, unsigned ctas, size_t smem, cudaStream_t s)
{
    const cudaError_t g = grant_dynamic_smem(reinterpret_cast<const void *>(&stn_bwd_kframe_kernel<GT, CG, GRAY, NCH>), smem);
    if (g != cudaSuccess) return g;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, 0);
    return cudaLaunchKernelEx(&cfg, stn_bwd_kframe_kernel<GT, CG, GRAY, NCH>, p);
}

// NCH: 32-column chunks of a crop row (crops up to 128 pixels wide)
template <typename GT, int CG, bool GRAY>
static cudaError_t launch_kframe_tt(const CropParams &p, unsigned ctas, size_t smem, cudaStream_t s)
{
    switch ((p.oW + 31) >> 5) {
    case 1: return launch_kframe_ttt<GT, CG, GRAY, 1>(p, ctas, smem, s);
    case 2: return launch_kframe_ttt<GT, CG, GRAY, 2>(p, ctas, smem, s);
    case 3: return launch_kframe_ttt<GT, CG, GRAY, 3>(p, ctas, smem, s);
    default: return launch_kframe_ttt<GT, CG, GRAY, 4>(p, ctas, smem, s);
    }
}

static int g_kf_rows = 0, g_kf_pad_kb = 0;                              // tuning knobs (loans_stn_configure): 0 = automatic
// bits 0-15: frame rows per CTA; bits 16...: KiB of shared-memory padding instead of the automatic rule (A/B of the CTAs per SM)
void kframe_tuning(int rows) { g_kf_rows = rows & 0xffff; g_kf_pad_kb = rows >> 16; }

// Returns -1 when the call is not one this path takes (one crop per frame, frames without grad, channel count not 1 / 3 / 4,
// unaligned rows): the caller then launches the general kernel.  Two launches: gtheta (table-driven theta kernel), then gx.
int launch_crop_bwd_kframe(CropParams p, int gy_dtype, cudaStream_t stream, bool by_measurement)
{
    if (!p.gx || p.K < 1) return -1;
    // measured against the general kernel at BASELINE config 4's shapes: 128 us vs 155 us with 32 frames, 311-349 vs 541 us with
    // 128; 50 vs 45 us with 8 frames and 32 vs 29 us with 2 (two launches, tables of every crop per CTA): taken from 16 frames
    if (by_measurement && p.N / p.K < 16) return -1;
    // the gx kernel takes a frame only if every crop on it steps by >= ~2 frame pixels per crop pixel (other frames run the
    // general role inside the launch, at the general kernel's speed): by default only where a box of half the frame still does
    if (by_measurement && ((p.oW > 1 && (p.W - 1) < 4 * (p.oW - 1)) || (p.oH > 1 && (p.H - 1) < 4 * (p.oH - 1)))) return -1;
    if (p.C != 1 && p.C != 3 && p.C != 4) return -1;
    if (p.oW > 128 || p.oH > 16383) return -1;                                            // a crop row is held in registers, 32 columns per lane slot
    if (p.W % 4 != 0 || (reinterpret_cast<uintptr_t>(p.gx) & 15) != 0) return -1;
    if (p.H > kAxIdxMask - 2 || p.W > kAxIdxMask - 2) return -1;
    if ((long long)p.H * p.W * p.C > 0x7fffffffLL) return -1;
    if ((long long)p.K * p.C * p.oH * p.oW > 0x7fffffffLL) return -1;    // gy of one frame's crops is indexed with 32-bit integers
    const int frames = p.N / p.K;
    // a frame with a crop this path declines runs the general gx role: same tile geometry as launch_crop_bwd, vector stores
    {
        const int nx = (p.W + 63) / 64;
        const int tw = (((p.W + nx - 1) / nx) + 3) & ~3;
        const int tr = p.H < 8 ? p.H : 8;
        p.gx_tile_rows = tr; p.gx_tile_cols = tw; p.gx_tile_pitch = tw;
        p.gx_tiles_x = (p.W + tw - 1) / tw;
        p.gx_tiles_per_frame = p.gx_tiles_x * ((p.H + tr - 1) / tr);
        p.gx_tile_bytes = (int)(sizeof(float) * (size_t)p.C * tr * tw * kWarps);
        p.gx_vec4 = 1; p.gx_tma_store = 0; p.gx_zero_bytes = 0;
    }
    // frame rows per CTA: a multiple of the warps per CTA, a little over ONE wave of CTAs (three per SM) over the machine.  The
    // tables of all the frame's crops are built once per CTA, so long bands are cheaper; the grid walks the bands centre-out, so
    // the few CTAs behind the first wave are the light top / bottom bands.  Measured at BASELINE config 4 (128 frames, backward
    // incl. the theta kernel): 312 / 302 us with 128 / 136 rows (4 CTAs per frame), 360 us with 120 (5), 333-350 us with 16 ... 104
    int rows = g_kf_rows;
    if (rows <= 0) {
        long long per_frame = (long long)(1.15 * 3.0 * num_sms() / (double)frames + 0.5);        // CTAs per frame wanted
        if (per_frame < 1) per_frame = 1;
        rows = (int)((p.H + per_frame - 1) / per_frame);
        // ... but no longer than the tables ask for: with few crops per frame they are cheap and short bands balance better
        // (one crop per frame, config 3: 172 us with 16 / 32 rows, 177 with 64, 184 with 112, 208 with 256)
        if (rows > 8 * p.K) rows = 8 * p.K;
        if (rows < 4 * kWarps) rows = 4 * kWarps;
        rows = (rows + kWarps - 1) / kWarps * kWarps;
    }
    if (rows > p.H) rows = p.H;
    p.kf_rows_cta = rows;
    p.kf_ctas_per_frame = (p.H + rows - 1) / rows;
    p.band_fb_tiles_per_warp = (p.gx_tiles_per_frame + kWarps * p.kf_ctas_per_frame - 1) / (kWarps * p.kf_ctas_per_frame);
    const size_t tables = sizeof(KfCol) * (size_t)p.K * (p.oW + p.oH) + ((sizeof(short) * (size_t)p.K * rows + 15) & ~(size_t)15);
    const size_t general = sizeof(float) * (size_t)((p.oW + p.oH + 3) & ~3) + sizeof(ScatterGeom) * (size_t)p.K;   // aliases the tables
    const size_t fixed = ((sizeof(KfCrop) * (size_t)p.K + 15) & ~(size_t)15) + (tables > general ? tables : general) + 16;
    const size_t region = sizeof(float) * (size_t)p.C * p.W * kWarps;
    p.kf_region_bytes = (int)(((region > (size_t)p.gx_tile_bytes ? region : (size_t)p.gx_tile_bytes) + 127) & ~(size_t)127);
    size_t smem = (size_t)p.kf_region_bytes + fixed + (size_t)g_kf_pad_kb * 1024;
    // several crops per frame: TWO CTAs per SM, not three.  Every gy row is read twice, by the warps that own the two frame rows
    // it lands on; with two CTAs the SM keeps ~90 KB of L1 and the second read hits it, with three (216 KB of shared memory) 28 KB
    // are left and it does not: 310-313 vs 367-370 us at BASELINE config 4 (no difference with one crop per frame: 172-177 us)
    if (p.K > 1 && g_kf_pad_kb == 0 && smem < 77 * 1024) smem = 77 * 1024;
    if (smem > 200 * 1024) return -1;                                    // frame rows too wide or too many crops per frame
    const long long ctas = (long long)frames * p.kf_ctas_per_frame;
    if (ctas > 0x7fffffffLL) return -1;
    // gtheta first (every crop pixel once, with its taps): the table-driven theta kernel, as for frames without grad
    {
        CropParams pt = p;
        pt.gx = nullptr;
        const int rc = launch_crop_bwd_theta_tab(pt, gy_dtype, stream);
        if (rc != 0) return rc;                         // -1: not taken (nothing launched), > 0: error
    }
    p.ggrid_out = nullptr;                              // written by the theta kernel
    cudaError_t e;
    if (p.nhwc)
        e = launch_kframe_tt<Nhwc4, 3, false>(p, (unsigned)ctas, smem, stream);
    else if (gy_dtype == 0)
        e = p.C == 1 ? launch_kframe_tt<float, 1, false>(p, (unsigned)ctas, smem, stream)
          : p.C == 4 ? launch_kframe_tt<float, 4, false>(p, (unsigned)ctas, smem, stream)
          : p.gray   ? launch_kframe_tt<float, 3, true>(p, (unsigned)ctas, smem, stream)
                     : launch_kframe_tt<float, 3, false>(p, (unsigned)ctas, smem, stream);
    else
        e = p.C == 1 ? launch_kframe_tt<__nv_bfloat16, 1, false>(p, (unsigned)ctas, smem, stream)
          : p.C == 4 ? launch_kframe_tt<__nv_bfloat16, 4, false>(p, (unsigned)ctas, smem, stream)
          : p.gray   ? launch_kframe_tt<__nv_bfloat16, 3, true>(p, (unsigned)ctas, smem, stream)
                     : launch_kframe_tt<__nv_bfloat16, 3, false>(p, (unsigned)ctas, smem, stream);
    count_launch();
    note_kernel("stn_bwd_kframe_kernel");                              // appended to the theta kernel's name
    if (e != cudaSuccess) return set_error("crop_bwd (crops per frame > 1) launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace stn
