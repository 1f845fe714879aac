// stn_common.cuh -- launch parameter block, dtype helpers and error plumbing shared by the .cu files.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "stn_math.cuh"

namespace stn {

constexpr int kThreads = 256;       // threads per CTA, every kernel
constexpr int kWarps = kThreads / 32;

// multiprocessors of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on B200): the grid-size
// rules are written in multiples of it
int num_sms();
// opt a kernel into more than 48 KiB of dynamic shared memory on the CURRENT device (the attribute is per device and per
// function: cached per (function, device), safe from several threads)
cudaError_t grant_dynamic_smem(const void *func, size_t bytes);
// name of the kernel a launcher picked, for loans_stn_last_kernel() (thread-local, last compute call)
void note_kernel(const char *name);

// One parameter block for the whole family; unused pointers are null.
struct CropParams {
    const float *x;          // (B,C,H,W)
    const float *theta;      // (N,2,3)        fused path
    const float *grid_in;    // (N,2,oH,oW)    explicit-grid path
    float mask01;
    void *y;                 // (N,C,oH,oW) f32|bf16
    float *grid_out;         // (N,2,oH,oW) or null
    const void *gy;          // (N,C,oH,oW) f32|bf16
    const float *ggrid_up;   // (N,2,oH,oW) or null
    float *gtheta;           // (N,2,3)
    float *gx;               // (B,C,H,W) or null
    float *ggrid_out;        // (N,2,oH,oW) or null
    float *corners_out;      // (N,2,2,2) or null: the grid at its four corners [., ., {0,oH-1}, {0,oW-1}] (forward)
    const float *gcorners;   // (N,2,2,2) or null: gradient arriving on those four grid points (backward)
    int nhwc;                // C == 3, bf16 only: y / gy are (N,oH,oW,4) channels-last, channel count padded to four
    int gray;                // C == 3 only: y / gy are (N,1,oH,oW), y = 0.299*ch2 + 0.587*ch1 + 0.114*ch0 (see gray_coef)
    int N, K, C, H, W, oH, oW;
    double xstep, ystep;     // 2/(oW-1), 2/(oH-1)
    int px_per_cta;          // crop pixels handled by one CTA of the per-crop roles
    int ctas_per_crop;       // == cluster size in the backward theta role
    // backward gx role: CTAs [0, gx_ctas), eight warp-owned tiles of one frame each; the rest reduce gtheta
    int gx_ctas, gx_ctas_per_frame, gx_tiles_per_frame, gx_tiles_x;
    int theta_ctas, theta_first;   // theta_first: the theta-role CTAs take the low block indices (scheduled first)
    int gx_tile_rows, gx_tile_cols, gx_tile_pitch, gx_tile_bytes;
    int gx_vec4;             // 128-bit stores of gx are legal (W % 4 == 0, aligned base)
    int gx_tma_store;        // tiles leave shared memory through TMA tensor stores (map passed next to this struct)
    int gx_zero_bytes;       // one zero plane per CTA behind the tiles: source of the stores of untouched tiles
    // separable path (stn_separable.cu)
    int sep_rows, sep_pitch, sep_buf_offset, gx_tiles_per_warp;
    // band backward (stn_band.cu): ctas_per_crop CTAs (one cluster) per crop, band_rows_cta crop rows each, worked
    // through in bands of at most band_rows crop rows; the tile holds band_cap frame rows of W floats per channel
    int band_rows_cta, band_rows, band_cap, band_tab_rows;
    int band_tile_bytes, band_zero_bytes, band_region_bytes;   // region = tile + zero plane, or the general roles' warp tiles
    int band_flags;                                             // A/B switches (bit 0: zero rows stored early, bit 1: no L2 prefetch of the taps)
    int band_fb_tiles_per_warp;                                 // declined crops: tiles per warp in the general gx role
    // several crops per frame, gx of axis-aligned crops (stn_kframe.cu): kf_ctas_per_frame CTAs per frame, kf_rows_cta frame
    // rows each; the region holds the warps' row buffers (or the general role's warp tiles)
    int kf_rows_cta, kf_ctas_per_frame, kf_region_bytes;
};

// Channels-last crop pixel with the channel count padded to four (LOANS_STN_FLAG_NHWC4: c == 3, bf16): y / gy are
// (N, oH, oW, 4) bf16, one 8-byte store / load per crop pixel -- the layout a tensor-core first convolution of the assessor
// consumes (reference common/net.py:15-25 is the consumer).  Used as the crop element type of the kernel templates.
struct alignas(8) Nhwc4 {
    __nv_bfloat16 c[4];
};
// planes of y / gy per crop, in elements of the crop type: C for planar crops, 1 behind the grayscale epilogue and for Nhwc4
template <typename T, bool GRAY> __host__ __device__ constexpr int crop_planes(int C) { return GRAY ? 1 : C; }
template <> __host__ __device__ constexpr int crop_planes<Nhwc4, false>(int) { return 1; }
template <> __host__ __device__ constexpr int crop_planes<Nhwc4, true>(int) { return 1; }

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float load(const float *p, size_t i) { return __ldg(p + i); }
    static __device__ __forceinline__ void store(float *p, size_t i, float v) { p[i] = v; }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float load(const __nv_bfloat16 *p, size_t i)
    {
        return __bfloat162float(__ldg(p + i));
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, size_t i, float v)
    {
        p[i] = __float2bfloat16_rn(v);
    }
};

template <> struct Elem<Nhwc4> {               // planar accessors are never reached with Nhwc4 crops (C == 3: the EXACT paths)
    static __device__ __forceinline__ float load(const Nhwc4 *p, size_t i) { return __bfloat162float(p[i].c[0]); }
    static __device__ __forceinline__ void store(Nhwc4 *, size_t, float) {}
    // one crop pixel: three channels + a zero, one 8-byte store
    static __device__ __forceinline__ void store_px(Nhwc4 *p, float c0, float c1, float c2)
    {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(c0, c1), hi = __floats2bfloat162_rn(c2, 0.0f);
        uint2 v;
        v.x = *reinterpret_cast<const unsigned *>(&lo);
        v.y = *reinterpret_cast<const unsigned *>(&hi);
        *reinterpret_cast<uint2 *>(p) = v;
    }
};

// Grayscale epilogue of the localizer (reference sheep/sheep_localizer.py:65-68, transform_rois_to_grayscale):
//     b, g, r = F.split_axis(rois, 3, axis=1);  rois = 0.299 * r + 0.587 * g + 0.114 * b
// i.e. channel 0 is taken as b and channel 2 as r, float32 products, summed left to right; the backward hands
// coef[ch] * ggray to channel ch (MulConstant's backward).
__device__ __forceinline__ float gray_coef(int ch) { return ch == 0 ? 0.114f : (ch == 1 ? 0.587f : 0.299f); }
__device__ __forceinline__ float gray_mix(float c0, float c1, float c2)
{
    return f_add(f_add(f_mul(0.299f, c2), f_mul(0.587f, c1)), f_mul(0.114f, c0));
}
// upstream gradient of channel ch at crop pixel q: gy[ch][q], or coef[ch] * ggray[q] behind the grayscale epilogue.
// base = this crop's gy at pixel q (channel 0 of the channel group at hand), npx = oH*oW.  GRAY is a compile-time switch:
// the backward kernels are register-bound, a run-time flag in their inner loops cost 6 % at every size.
template <typename GT, bool GRAY>
__device__ __forceinline__ float load_gy(const GT *base, int ch, int npx)
{
    if (GRAY) return f_mul(gray_coef(ch), Elem<GT>::load(base, 0));
    return Elem<GT>::load(base, ch * npx);
}
// channels-last: base points at the pixel; the 8-byte load is the same for every channel (one LDG.64 after CSE), bf16 ->
// float32 is a shift
template <>
__device__ __forceinline__ float load_gy<Nhwc4, false>(const Nhwc4 *base, int ch, int)
{
    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(base));
    const unsigned w = ch < 2 ? v.x : v.y;
    return __uint_as_float((ch & 1) ? (w & 0xffff0000u) : (w << 16));
}

// xs[0..oW) and ys[0..oH) into shared memory (numpy.linspace(-1,1,n,dtype=float32), see stn_math.cuh)
__device__ __forceinline__ void fill_axis_tables(float *xs, float *ys, int oW, int oH, double xstep, double ystep)
{
    for (int k = threadIdx.x; k < oW + oH; k += blockDim.x) {
        if (k < oW) xs[k] = linspace_pm1(k, oW, xstep);
        else ys[k - oW] = linspace_pm1(k - oW, oH, ystep);
    }
}

__device__ __forceinline__ float lin_x_at(const CropParams &p, int j)
{
    return linspace_pm1(j, p.oW, p.xstep);
}
__device__ __forceinline__ float lin_y_at(const CropParams &p, int i)
{
    return linspace_pm1(i, p.oH, p.ystep);
}
__device__ __forceinline__ void fill_axis_tables(const CropParams &p, float *xs, float *ys)
{
    for (int k = threadIdx.x; k < p.oW + p.oH; k += blockDim.x) {
        if (k < p.oW) xs[k] = lin_x_at(p, k);
        else ys[k - p.oW] = lin_y_at(p, k - p.oW);
    }
}

// ---- programmatic dependent launch (PDL).  The fused kernels are launched with cudaLaunchAttributeProgrammaticStream-
// Serialization: their CTAs may become resident, and run the part of their prologue that touches no global memory, while
// the previous kernel of the stream is still draining.  pdl_launch_dependents() at entry lets the NEXT kernel do the
// same; pdl_wait() returns once every prerequisite grid has completed and its memory is visible -- it comes before the
// first global access, reads and writes alike.  Both are no-ops for a launch without programmatic edges.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// L2 prefetches are the one kind of global access that may come BEFORE pdl_wait(): nothing is read into registers and L2 is
// the point of coherence, so whatever the previous kernel still writes is what a later load sees.  The first warp of a CTA
// asks for the 128-byte lines of theta from its own crop onwards (32 lines = 170 crops): the CTAs that become resident
// while the previous kernel drains warm theta for the ones behind them.
__device__ __forceinline__ void pdl_prefetch_theta(const CropParams &p, int n)
{
#ifndef STN_NO_PDL_PREFETCH
    if (threadIdx.x < 32) {
        const char *base = reinterpret_cast<const char *>(p.theta);
        const size_t bytes = 24 * (size_t)p.N, off = 24 * (size_t)n + 128 * (size_t)threadIdx.x;
        if (off < bytes) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
#endif
}

// ---- host-side error plumbing (definitions in stn_abi.cu)
int set_error(const char *fmt, ...);
void count_launch(int n = 1);
int check_launch(const char *what);
bool pdl_enabled();
bool theta_first_enabled();
int gx_tiles_per_warp_override();
int fwd_px_per_cta_override();
bool theta_only_kernel_enabled();
// fills the launch attributes shared by the fused kernels: [cluster dimension,] programmatic stream serialisation
inline unsigned fill_launch_attrs(cudaLaunchAttribute *attr, unsigned cluster)
{
    unsigned na = 0;
    if (cluster > 0) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cluster;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    return na;
}

}  // namespace stn
