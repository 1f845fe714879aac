// stn_abi.cu -- extern "C" entry points declared in include/loans_stn.h: argument validation, error
// reporting, launch accounting.  No torch types, no allocation, no synchronisation.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/loans_stn.h"
#include "../../include/loans_stn_devel.h"
#include "stn_common.cuh"

namespace stn {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

void count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

// ---- loans_stn_last_kernel(): the kernels the process's last compute entry point launched, '+'-separated.  Process-wide
// (a backward usually runs on the autograd engine's thread, the question is asked from the main thread); a thread-local
// being-built note is published when the entry point returns through a launcher
static std::mutex g_kernel_mu;
static char g_kernel[192] = "";
static thread_local char g_kernel_ret[192] = "";
static void reset_kernel_note()
{
    std::lock_guard<std::mutex> lock(g_kernel_mu);
    g_kernel[0] = 0;
}
void note_kernel(const char *name)
{
    std::lock_guard<std::mutex> lock(g_kernel_mu);
    const size_t have = strlen(g_kernel);
    snprintf(g_kernel + have, sizeof(g_kernel) - have, "%s%s", have ? "+" : "", name);
}

// ---- per-device facts, cached
static constexpr int kMaxDevices = 64;
int num_sms()
{
    static std::atomic<int> cache[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return 148; }
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
    cache[dev].store(v, std::memory_order_relaxed);
    return v;
}

cudaError_t grant_dynamic_smem(const void *func, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::unordered_map<unsigned long long, size_t> granted;       // (function, device) -> bytes granted so far
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    // kernels are 16-byte aligned objects: the low bits of the address carry the device ordinal
    const unsigned long long key = ((unsigned long long)reinterpret_cast<uintptr_t>(func) << 8) ^ (unsigned long long)(dev & 0xff);
    std::lock_guard<std::mutex> lock(mu);
    size_t &g = granted[key];
    if (bytes <= g) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) g = bytes;
    return e;
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

int launch_rotation_dropout(const float *in, float mask01, float *out, int n, cudaStream_t stream);
int launch_grid_fwd(const float *theta, float *grid, int n, int oh, int ow, cudaStream_t stream);
int launch_grid_bwd(const float *ggrid, float *gtheta, int n, int oh, int ow, cudaStream_t stream);
int launch_sampler_bwd(CropParams p, int gy_dtype, cudaStream_t stream);
int launch_crop_fwd(CropParams p, bool from_grid, int y_dtype, cudaStream_t stream);
int launch_crop_bwd(CropParams p, int gy_dtype, cudaStream_t stream);
#ifdef STN_DEVEL
int launch_sep_fwd(CropParams p, int y_dtype, cudaStream_t stream);
#endif
int launch_crop_bwd_band(CropParams p, int gy_dtype, cudaStream_t stream, bool by_measurement);   // -1: use the general kernel
int launch_crop_bwd_kframe(CropParams p, int gy_dtype, cudaStream_t stream, bool by_measurement); // -1: use the general kernel
void kframe_tuning(int rows);
void band_tuning(int which, int value);
int launch_crop_bwd_theta_tab(CropParams p, int gy_dtype, cudaStream_t stream);   // -1: not taken
int launch_prepare_images(const float *x, float *out, float scale, int b, int h, int w, cudaStream_t stream);      // -1: shape not supported, use the general kernel

static std::atomic<int> g_force_general{0};
static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load() != 0; }
static std::atomic<int> g_band_backward{-1};    // -1: by shape (wide frame rows), 0: never, 1: whenever it applies
#ifdef STN_DEVEL
// A/B switches of a -DSTN_DEVEL build (include/loans_stn_devel.h); the product build has the defaults compiled in
static std::atomic<int> g_tma_forward{0};
static std::atomic<int> g_theta_first{0};
static std::atomic<int> g_gx_tpw{0};
static std::atomic<int> g_fwd_px{0};
static std::atomic<int> g_theta_only{1};
bool theta_first_enabled() { return g_theta_first.load() != 0; }
int gx_tiles_per_warp_override() { return g_gx_tpw.load(); }
int fwd_px_per_cta_override() { return g_fwd_px.load(); }
bool theta_only_kernel_enabled() { return g_theta_only.load() != 0; }
#else
bool theta_first_enabled() { return false; }
int gx_tiles_per_warp_override() { return 0; }
int fwd_px_per_cta_override() { return 0; }
bool theta_only_kernel_enabled() { return true; }
#endif

static int need_device(const char *what)
{
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error("%s: no usable CUDA device (%s); this library has no CPU fallback", what, cudaGetErrorString(e));
    }
    return 0;
}

static int check_dims(const char *what, int n, int k, int c, int h, int w, int oh, int ow)
{
    if (n < 0 || k < 1 || c < 1 || h < 1 || w < 1 || oh < 1 || ow < 1)
        return set_error("%s: bad dimensions n=%d k=%d c=%d h=%d w=%d oh=%d ow=%d", what, n, k, c, h, w, oh, ow);
    if (n % k != 0) return set_error("%s: n=%d crops is not a multiple of k=%d crops per frame", what, n, k);
    if ((long long)h * w > 0x3fffffffLL || (long long)oh * ow > 0x3fffffffLL)
        return set_error("%s: a single plane exceeds 2^30 elements", what);
    return 0;
}

static int check_dtype(const char *what, int dt)
{
    if (dt != LOANS_STN_F32 && dt != LOANS_STN_BF16) return set_error("%s: unknown crop dtype %d", what, dt);
    return 0;
}

static CropParams base_params(int n, int k, int c, int h, int w, int oh, int ow)
{
    CropParams p = {};
    p.N = n; p.K = k; p.C = c; p.H = h; p.W = w; p.oH = oh; p.oW = ow;
    p.mask01 = 1.0f;
    p.xstep = ow > 1 ? 2.0 / (double)(ow - 1) : 0.0;
    p.ystep = oh > 1 ? 2.0 / (double)(oh - 1) : 0.0;
    return p;
}

// upright: the caller knows (mask01 == 0) or asserts (LOANS_STN_FLAG_UPRIGHT: theta came out of a rotation dropout that
// zeroed the rotation terms) that the crops are axis-aligned boxes.  It only SELECTS the kernels written for that case;
// every one of them checks each crop's own rotation terms on the device and runs the general roles for a crop that is
// rotated after all, inside the same launch -- a wrong hint costs time, never correctness.
static std::atomic<bool> g_kf_single{false};
static int crop_bwd_dispatch(const CropParams &p, bool upright, int gy_dtype, cudaStream_t stream)
{
    const int band = g_band_backward.load();
    if (upright && band != 0 && !g_force_general.load()) {
        // frames without grad (every LoANs call): the table-driven theta kernel, any number of crops per frame
        if (p.gx == nullptr && theta_only_kernel_enabled()) {
            const int rc = launch_crop_bwd_theta_tab(p, gy_dtype, stream);
            if (rc >= 0) return rc;
        }
        // one crop per frame, gx wanted: the band backward -- every crop pixel evaluated once, gx written once by the band
        // that owns the frame rows (stn_band.cu).  By default it is taken where it measured faster than the general kernel
        // (rule in launch_crop_bwd_band); LOANS_STN_CFG_BAND_BACKWARD = 1 / 0 forces it on (wherever it applies) / off
        // one crop per frame on WIDE frame rows (>= 4 KiB per row group, e.g. 512-px RGB frames): the two-kernel path of
        // stn_kframe.cu -- gtheta by the table-driven theta kernel, gx by warp-owned frame rows from gy and the table weights
        // alone -- measured faster than the CTA bands (172 vs 190 us at BASELINE config 3); narrow frames stay with the row
        // bands (cfg2: 28 vs 17 us, cfg5: 319 vs 229 us).  LOANS_STN_CFG_KFRAME_SINGLE forces it for any width (A/B arm)
        if (p.gx != nullptr && p.K == 1 && (g_kf_single.load() || (band < 0 && sizeof(float) * (size_t)p.W * p.C >= 4096))) {
            const int rc = launch_crop_bwd_kframe(p, gy_dtype, stream, band < 0 && !g_kf_single.load());
            if (rc >= 0) return rc;
        }
        if (p.gx != nullptr && p.K == 1) {
            const int rc = launch_crop_bwd_band(p, gy_dtype, stream, band < 0);
            if (rc >= 0) return rc;
        }
        // several crops per frame, gx wanted: warp-owned frame rows, each crop row found through an inverse row map and added
        // from gy and the table weights alone; gtheta by the table-driven theta kernel in front of it (stn_kframe.cu)
        if (p.gx != nullptr && p.K > 1) {
            const int rc = launch_crop_bwd_kframe(p, gy_dtype, stream, band < 0);
            if (rc >= 0) return rc;
        }
    }
    return launch_crop_bwd(p, gy_dtype, stream);
}

}  // namespace stn

using namespace stn;

#define REQUIRE_PTR(what, ptr)                                                    \
    do {                                                                          \
        if ((ptr) == nullptr) return set_error("%s: %s is NULL", what, #ptr);     \
    } while (0)

extern "C" {

int loans_stn_abi_version(void) { return LOANS_STN_ABI_VERSION; }

const char *loans_stn_last_error(void) { return g_err; }

unsigned long long loans_stn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char *loans_stn_last_kernel(void)
{
    std::lock_guard<std::mutex> lock(g_kernel_mu);
    memcpy(g_kernel_ret, g_kernel, sizeof(g_kernel_ret));
    return g_kernel_ret;
}

int loans_stn_configure(int key, int value)
{
    if (key == LOANS_STN_CFG_FORCE_GENERAL) { g_force_general.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_PDL) { g_pdl.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_BAND_BACKWARD) { g_band_backward.store(value < 0 ? -1 : (value != 0)); return 0; }
    if (key >= LOANS_STN_CFG_BAND_CS && key <= LOANS_STN_CFG_BAND_VARIANT) {          // test hooks (loans_stn_devel.h)
        if (value < 0) return set_error("loans_stn_configure: key %d needs a value >= 0", key);
        band_tuning(key - LOANS_STN_CFG_BAND_CS, value);
        return 0;
    }
    if (key == LOANS_STN_CFG_KFRAME_SINGLE) { g_kf_single.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_KFRAME_ROWS) {
        if (value < 0) return set_error("loans_stn_configure: key %d needs a value >= 0", key);
        kframe_tuning(value);
        return 0;
    }
#ifdef STN_DEVEL
    if (key == LOANS_STN_CFG_TMA_FORWARD) { g_tma_forward.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_GX_TILES_PER_WARP) { g_gx_tpw.store(value < 0 ? 0 : (value > 64 ? 64 : value)); return 0; }
    if (key == LOANS_STN_CFG_THETA_ONLY_KERNEL) { g_theta_only.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_FWD_PX_PER_CTA) { g_fwd_px.store(value > 0 ? ((value + 255) / 256) * 256 : 0); return 0; }
    if (key == LOANS_STN_CFG_THETA_FIRST) { g_theta_first.store(value != 0); return 0; }
#else
    if (key == LOANS_STN_CFG_TMA_FORWARD || (key >= LOANS_STN_CFG_THETA_FIRST && key <= LOANS_STN_CFG_FWD_PX_PER_CTA))
        return set_error("loans_stn_configure: key %d is an A/B switch of a -DSTN_DEVEL build (include/loans_stn_devel.h)", key);
#endif
    return set_error("loans_stn_configure: unknown key %d", key);
}

int loans_stn_rotation_dropout(const float *theta_in, float mask01, float *theta_out, int n, void *stream)
{
    const char *what = "loans_stn_rotation_dropout";
    reset_kernel_note();
    if (n < 0) return set_error("%s: n=%d", what, n);
    if (n == 0) return 0;
    REQUIRE_PTR(what, theta_in);
    REQUIRE_PTR(what, theta_out);
    if (need_device(what)) return 1;
    return launch_rotation_dropout(theta_in, mask01, theta_out, n, (cudaStream_t)stream);
}

int loans_stn_prepare_images(const float *x, float scale, float *out, int b, int c, int h, int w, void *stream)
{
    const char *what = "loans_stn_prepare_images";
    reset_kernel_note();
    if (check_dims(what, b, 1, c, h, w, 1, 1)) return 1;
    if (c != 3) return set_error("%s: frames must have 3 channels (RGB -> BGR), got %d", what, c);
    if (b == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, out);
    if (x == out) return set_error("%s: in-place is not possible (the channel order is reversed)", what);
    if (need_device(what)) return 1;
    return launch_prepare_images(x, out, scale, b, h, w, (cudaStream_t)stream);
}

int loans_stn_grid_fwd(const float *theta, float *grid, int n, int oh, int ow, void *stream)
{
    const char *what = "loans_stn_grid_fwd";
    reset_kernel_note();
    if (check_dims(what, n, 1, 1, 1, 1, oh, ow)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, grid);
    if (need_device(what)) return 1;
    return launch_grid_fwd(theta, grid, n, oh, ow, (cudaStream_t)stream);
}

int loans_stn_grid_bwd(const float *ggrid, float *gtheta, int n, int oh, int ow, void *stream)
{
    const char *what = "loans_stn_grid_bwd";
    reset_kernel_note();
    if (check_dims(what, n, 1, 1, 1, 1, oh, ow)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, ggrid);
    REQUIRE_PTR(what, gtheta);
    if (need_device(what)) return 1;
    return launch_grid_bwd(ggrid, gtheta, n, oh, ow, (cudaStream_t)stream);
}

int loans_stn_sampler_fwd(const float *x, const float *grid, void *y,
                          int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    const char *what = "loans_stn_sampler_fwd";
    reset_kernel_note();
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, y_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, grid);
    REQUIRE_PTR(what, y);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.grid_in = grid; p.y = y;
    return launch_crop_fwd(p, true, y_dtype, (cudaStream_t)stream);
}

int loans_stn_sampler_bwd(const float *x, const float *grid, const void *gy, float *gx, float *ggrid,
                          int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    const char *what = "loans_stn_sampler_bwd";
    reset_kernel_note();
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, gy_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, grid);
    REQUIRE_PTR(what, gy);
    if (gx == nullptr && ggrid == nullptr) return 0;
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.grid_in = grid; p.gy = gy; p.gx = gx; p.ggrid_out = ggrid;
    return launch_sampler_bwd(p, gy_dtype, (cudaStream_t)stream);
}

// ---- the composite: every public spelling funnels into these two
static int crop_fwd_impl(const char *what, const float *x, const float *theta, float mask01, void *y, float *grid, float *corners,
                         int flags, int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    reset_kernel_note();
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, y_dtype)) return 1;
    if (flags & ~LOANS_STN_FLAGS_ALL) return set_error("%s: unknown flags 0x%x", what, flags);
    if ((flags & LOANS_STN_FLAG_GRAY) && c != 3) return set_error("%s: the grayscale epilogue needs 3 channels, got %d", what, c);
    if (flags & LOANS_STN_FLAG_NHWC4) {
        if (c != 3 || y_dtype != LOANS_STN_BF16 || (flags & LOANS_STN_FLAG_GRAY))
            return set_error("%s: LOANS_STN_FLAG_NHWC4 needs 3 channels, bf16 crops and no grayscale epilogue", what);
        if (reinterpret_cast<uintptr_t>(y) & 7) return set_error("%s: NHWC4 crops must be 8-byte aligned", what);
    }
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, y);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.y = y; p.grid_out = grid; p.corners_out = corners;
    p.gray = (flags & LOANS_STN_FLAG_GRAY) ? 1 : 0;
    p.nhwc = (flags & LOANS_STN_FLAG_NHWC4) ? 1 : 0;
#ifdef STN_DEVEL
    // axis-aligned crops through the AxisTap-table + TMA-bulk-copy-staged kernel (stn_separable.cu): an A/B switch of the
    // devel build -- measured on B200 it is never faster than the direct gather (profiles/README.md)
    if (mask01 == 0.0f && flags == 0 && corners == nullptr && g_tma_forward.load() && !g_force_general.load() && w % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int rc = launch_sep_fwd(p, y_dtype, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
#endif
    return launch_crop_fwd(p, false, y_dtype, (cudaStream_t)stream);
}

static int crop_bwd_impl(const char *what, const float *x, const float *theta, float mask01, const void *gy,
                         const float *ggrid_upstream, const float *gcorners, float *gtheta, float *gx, float *ggrid_out,
                         int flags, int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    reset_kernel_note();
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, gy_dtype)) return 1;
    if (flags & ~LOANS_STN_FLAGS_ALL) return set_error("%s: unknown flags 0x%x", what, flags);
    if ((flags & LOANS_STN_FLAG_GRAY) && c != 3) return set_error("%s: the grayscale epilogue needs 3 channels, got %d", what, c);
    if (flags & LOANS_STN_FLAG_NHWC4) {
        if (c != 3 || gy_dtype != LOANS_STN_BF16 || (flags & LOANS_STN_FLAG_GRAY))
            return set_error("%s: LOANS_STN_FLAG_NHWC4 needs 3 channels, bf16 crops and no grayscale epilogue", what);
        if (reinterpret_cast<uintptr_t>(gy) & 7) return set_error("%s: NHWC4 crops must be 8-byte aligned", what);
    }
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, gy);
    REQUIRE_PTR(what, gtheta);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.gy = gy; p.ggrid_up = ggrid_upstream; p.gcorners = gcorners;
    p.gtheta = gtheta; p.gx = gx; p.ggrid_out = ggrid_out;
    p.gray = (flags & LOANS_STN_FLAG_GRAY) ? 1 : 0;
    p.nhwc = (flags & LOANS_STN_FLAG_NHWC4) ? 1 : 0;
    const bool upright = mask01 == 0.0f || (flags & LOANS_STN_FLAG_UPRIGHT) != 0;
    return crop_bwd_dispatch(p, upright, gy_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_fwd(const float *x, const float *theta, float mask01, void *y, float *grid,
                       int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    return crop_fwd_impl("loans_stn_crop_fwd", x, theta, mask01, y, grid, nullptr, 0, n, k, c, h, w, oh, ow, y_dtype, stream);
}

int loans_stn_crop_bwd(const float *x, const float *theta, float mask01, const void *gy,
                       const float *ggrid_upstream, float *gtheta, float *gx, float *ggrid_out,
                       int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    return crop_bwd_impl("loans_stn_crop_bwd", x, theta, mask01, gy, ggrid_upstream, nullptr, gtheta, gx, ggrid_out, 0,
                         n, k, c, h, w, oh, ow, gy_dtype, stream);
}

int loans_stn_crop_fwd_corners(const float *x, const float *theta, float mask01, void *y, float *corners,
                               int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    const char *what = "loans_stn_crop_fwd_corners";
    if (n > 0) REQUIRE_PTR(what, corners);
    return crop_fwd_impl(what, x, theta, mask01, y, nullptr, corners, 0, n, k, c, h, w, oh, ow, y_dtype, stream);
}

int loans_stn_crop_bwd_corners(const float *x, const float *theta, float mask01, const void *gy,
                               const float *gcorners, float *gtheta, float *gx,
                               int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    return crop_bwd_impl("loans_stn_crop_bwd_corners", x, theta, mask01, gy, nullptr, gcorners, gtheta, gx, nullptr, 0,
                         n, k, c, h, w, oh, ow, gy_dtype, stream);
}

int loans_stn_crop_fwd_ex(const float *x, const float *theta, float mask01, void *y, float *grid, float *corners,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    return crop_fwd_impl("loans_stn_crop_fwd_ex", x, theta, mask01, y, grid, corners, flags, n, k, c, h, w, oh, ow, y_dtype, stream);
}

int loans_stn_crop_bwd_ex(const float *x, const float *theta, float mask01, const void *gy, const float *ggrid_upstream,
                          const float *gcorners, float *gtheta, float *gx, float *ggrid_out,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    return crop_bwd_impl("loans_stn_crop_bwd_ex", x, theta, mask01, gy, ggrid_upstream, gcorners, gtheta, gx, ggrid_out, flags,
                         n, k, c, h, w, oh, ow, gy_dtype, stream);
}

}  // extern "C"
