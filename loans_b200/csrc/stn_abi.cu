// stn_abi.cu -- extern "C" entry points declared in include/loans_stn.h: argument validation, error
// reporting, launch accounting.  No torch types, no allocation, no synchronisation.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/loans_stn.h"
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
int launch_sep_fwd(CropParams p, int y_dtype, cudaStream_t stream);
int launch_crop_bwd_band(CropParams p, int gy_dtype, cudaStream_t stream, bool by_measurement);   // -1: use the general kernel
void band_tuning(int which, int value);
int launch_crop_bwd_theta_tab(CropParams p, int gy_dtype, cudaStream_t stream);   // -1: not taken
int launch_prepare_images(const float *x, float *out, float scale, int b, int h, int w, cudaStream_t stream);      // -1: shape not supported, use the general kernel

static std::atomic<int> g_force_general{0};
static std::atomic<int> g_tma_forward{0};
static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load() != 0; }
static std::atomic<int> g_theta_first{0};
bool theta_first_enabled() { return g_theta_first.load() != 0; }
static std::atomic<int> g_gx_tpw{0};
int gx_tiles_per_warp_override() { return g_gx_tpw.load(); }
static std::atomic<int> g_fwd_px{0};
int fwd_px_per_cta_override() { return g_fwd_px.load(); }
static std::atomic<int> g_theta_only{1};
bool theta_only_kernel_enabled() { return g_theta_only.load() != 0; }
static std::atomic<int> g_band_backward{-1};    // -1: by shape (wide frame rows), 0: never, 1: whenever it applies

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

static int crop_bwd_dispatch(const CropParams &p, float mask01, int k, int c, int w, const float *gx, int gy_dtype, cudaStream_t stream)
{
    // mask01 == 0 (LoANs' ratio = 0.0), one crop per frame, gx wanted: the band backward -- every crop pixel evaluated
    // once, gx written once by the band that owns the frame rows (stn_band.cu); crops it declines run the general roles
    // inside the same launch.  By default it is taken where it measured faster than the general kernel on B200 (row bands
    // for narrow frames and enough crops, CTA bands for wide frame rows; the rule is in launch_crop_bwd_band);
    // LOANS_STN_CFG_BAND_BACKWARD = 1 / 0 forces it on (wherever it applies) / off.
    const int band = g_band_backward.load();
    // frames without grad, rotation masked: the table-driven theta kernel (any number of crops per frame)
    if (mask01 == 0.0f && gx == nullptr && band != 0 && theta_only_kernel_enabled() && !g_force_general.load()) {
        const int rc = launch_crop_bwd_theta_tab(p, gy_dtype, stream);
        if (rc >= 0) return rc;
    }
    if (mask01 == 0.0f && k == 1 && gx != nullptr && band != 0 && !g_force_general.load()) {
        const int rc = launch_crop_bwd_band(p, gy_dtype, stream, band < 0);
        if (rc >= 0) return rc;
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

int loans_stn_configure(int key, int value)
{
    if (key == LOANS_STN_CFG_FORCE_GENERAL) { g_force_general.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_TMA_FORWARD) { g_tma_forward.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_GX_TILES_PER_WARP) { g_gx_tpw.store(value < 0 ? 0 : (value > 64 ? 64 : value)); return 0; }
    if (key == LOANS_STN_CFG_THETA_ONLY_KERNEL) { g_theta_only.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_FWD_PX_PER_CTA) { g_fwd_px.store(value > 0 ? ((value + 255) / 256) * 256 : 0); return 0; }
    if (key == LOANS_STN_CFG_THETA_FIRST) { g_theta_first.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_PDL) { g_pdl.store(value != 0); return 0; }
    if (key == LOANS_STN_CFG_BAND_BACKWARD) { g_band_backward.store(value < 0 ? -1 : (value != 0)); return 0; }
    if (key >= LOANS_STN_CFG_BAND_CS && key <= LOANS_STN_CFG_BAND_VARIANT) {
        if (value < 0) return set_error("loans_stn_configure: key %d needs a value >= 0", key);
        band_tuning(key - LOANS_STN_CFG_BAND_CS, value);
        return 0;
    }
    return set_error("loans_stn_configure: unknown key %d", key);
}

int loans_stn_rotation_dropout(const float *theta_in, float mask01, float *theta_out, int n, void *stream)
{
    const char *what = "loans_stn_rotation_dropout";
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

int loans_stn_crop_fwd(const float *x, const float *theta, float mask01, void *y, float *grid,
                       int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    const char *what = "loans_stn_crop_fwd";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, y_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, y);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.y = y; p.grid_out = grid;
    // mask01 == 0 (LoANs' ratio = 0.0): every crop is axis-aligned -> the table + TMA-staged kernel applies.  It is
    // opt-in: measured on B200 it is never faster than the direct gather (profiles/README.md), so the default is off.
    if (mask01 == 0.0f && g_tma_forward.load() && !g_force_general.load() && w % 4 == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int rc = launch_sep_fwd(p, y_dtype, (cudaStream_t)stream);
        if (rc >= 0) return rc;
    }
    return launch_crop_fwd(p, false, y_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_fwd_ex(const float *x, const float *theta, float mask01, void *y, float *grid, float *corners,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    const char *what = "loans_stn_crop_fwd_ex";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, y_dtype)) return 1;
    if (flags & ~LOANS_STN_FLAG_GRAY) return set_error("%s: unknown flags 0x%x", what, flags);
    if ((flags & LOANS_STN_FLAG_GRAY) && c != 3) return set_error("%s: the grayscale epilogue needs 3 channels, got %d", what, c);
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, y);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.y = y; p.grid_out = grid; p.corners_out = corners;
    p.gray = (flags & LOANS_STN_FLAG_GRAY) ? 1 : 0;
    return launch_crop_fwd(p, false, y_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_bwd_ex(const float *x, const float *theta, float mask01, const void *gy, const float *ggrid_upstream,
                          const float *gcorners, float *gtheta, float *gx, float *ggrid_out,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    const char *what = "loans_stn_crop_bwd_ex";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, gy_dtype)) return 1;
    if (flags & ~LOANS_STN_FLAG_GRAY) return set_error("%s: unknown flags 0x%x", what, flags);
    if ((flags & LOANS_STN_FLAG_GRAY) && c != 3) return set_error("%s: the grayscale epilogue needs 3 channels, got %d", what, c);
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
    return crop_bwd_dispatch(p, mask01, k, c, w, gx, gy_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_fwd_corners(const float *x, const float *theta, float mask01, void *y, float *corners,
                               int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream)
{
    const char *what = "loans_stn_crop_fwd_corners";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, y_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, y);
    REQUIRE_PTR(what, corners);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.y = y; p.corners_out = corners;
    return launch_crop_fwd(p, false, y_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_bwd_corners(const float *x, const float *theta, float mask01, const void *gy,
                               const float *gcorners, float *gtheta, float *gx,
                               int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    const char *what = "loans_stn_crop_bwd_corners";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, gy_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, gy);
    REQUIRE_PTR(what, gtheta);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.gy = gy; p.gcorners = gcorners;
    p.gtheta = gtheta; p.gx = gx;
    return crop_bwd_dispatch(p, mask01, k, c, w, gx, gy_dtype, (cudaStream_t)stream);
}

int loans_stn_crop_bwd(const float *x, const float *theta, float mask01, const void *gy,
                       const float *ggrid_upstream, float *gtheta, float *gx, float *ggrid_out,
                       int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream)
{
    const char *what = "loans_stn_crop_bwd";
    if (check_dims(what, n, k, c, h, w, oh, ow) || check_dtype(what, gy_dtype)) return 1;
    if (n == 0) return 0;
    REQUIRE_PTR(what, x);
    REQUIRE_PTR(what, theta);
    REQUIRE_PTR(what, gy);
    REQUIRE_PTR(what, gtheta);
    if (need_device(what)) return 1;
    CropParams p = base_params(n, k, c, h, w, oh, ow);
    p.x = x; p.theta = theta; p.mask01 = mask01; p.gy = gy; p.ggrid_up = ggrid_upstream;
    p.gtheta = gtheta; p.gx = gx; p.ggrid_out = ggrid_out;
    return crop_bwd_dispatch(p, mask01, k, c, w, gx, gy_dtype, (cudaStream_t)stream);
}

}  // extern "C"
