// stn_ingest.cu -- the loader's frame path on the device: what the reference does on the host,
// per sample, in loader threads (common/datasets/image_dataset.py:16-28 resize_image, :98 / :181 `image / 255`):
//
//     PIL.Image.fromarray(uint8 HWC).convert('RGB').resize((W', H'), Image.LANCZOS) -> float32 CHW -> / 255
//
// as two integer kernels over a batch of decoded uint8 HWC frames.  Pillow's 8-bit resampling is reproduced BIT FOR BIT:
// the coefficient tables are computed on the host exactly as Pillow's precompute_coeffs / normalize_coeffs_8bpc do (float64
// Lanczos-3 windowed sinc from libm's sin, support 3 * max(scale, 1), normalised, 22-bit fixed point rounded half away from
// zero); the kernels are Pillow's two passes -- horizontal, then vertical -- each an int32 dot product started at 1 << 21,
// shifted and clipped to uint8; the last pass also converts: float32(v) / 255 (IEEE division, as numpy) into NCHW.
// Oracle: oracle/ingest_numpy.py, pinned against the real PIL.
#include <cmath>
#include <map>
#include <mutex>
#include <vector>

#include "../../include/loans_stn.h"
#include "stn_common.cuh"

namespace stn {

constexpr int kPrecisionBits = 32 - 8 - 2;

static double lanczos3(double x)
{
    if (-3.0 <= x && x < 3.0) {
        if (x == 0.0) return 1.0;
        const double a = x * M_PI, b = x / 3.0 * M_PI;
        return (sin(a) / a) * (b != 0.0 ? sin(b) / b : 1.0);
    }
    return 0.0;
}

static int axis_ksize(int in_size, int out_size)
{
    const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
    return (int)ceil(3.0 * filterscale) * 2 + 1;
}

// table of one axis: out_size records of (first tap, tap count, ksize coefficients), Pillow's precompute_coeffs +
// normalize_coeffs_8bpc with box = (0, in_size)
static std::vector<int> axis_table(int in_size, int out_size)
{
    const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale, ss = 1.0 / filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    std::vector<int> tab((size_t)out_size * (2 + ksize), 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = (xx + 0.5) * scale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        double ww = 0.0;
        for (int x = 0; x < xmax; ++x) {
            k[x] = lanczos3((x + xmin - center + 0.5) * ss);
            ww += k[x];
        }
        int *rec = tab.data() + (size_t)xx * (2 + ksize);
        rec[0] = xmin;
        rec[1] = xmax;
        for (int x = 0; x < xmax; ++x) {
            const double w = ww != 0.0 ? k[x] / ww : k[x];
            rec[2 + x] = w < 0 ? (int)(-0.5 + w * (1 << kPrecisionBits)) : (int)(0.5 + w * (1 << kPrecisionBits));
        }
    }
    return tab;
}

__device__ __forceinline__ int clip8(int acc)
{
    const int v = acc >> kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// horizontal pass: rows x w x 3 bytes -> rows x ow x 3 bytes (or, TO_FLOAT, straight to the float32 NCHW output when the
// height does not change)
template <bool TO_FLOAT>
__global__ void __launch_bounds__(kThreads) ingest_h_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ tmp,
                                                            float *__restrict__ out, const int *__restrict__ tab, int ksize,
                                                            long long npix, int h, int w, int ow)
{
    pdl_launch_dependents();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= npix) return;
    const long long row = idx / ow;
    const int xx = (int)(idx - row * ow);
    const int *rec = tab + (size_t)xx * (2 + ksize);
    const int xmin = rec[0], n = rec[1];
    const unsigned char *p = src + (row * w + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < n; ++x) {
        const int k = rec[2 + x];
        s0 += (int)p[3 * x] * k;
        s1 += (int)p[3 * x + 1] * k;
        s2 += (int)p[3 * x + 2] * k;
    }
    if (TO_FLOAT) {
        const long long b = row / h;
        const int y = (int)(row - b * h);
        float *o = out + ((size_t)b * 3 * h + y) * ow + xx;
        const size_t plane = (size_t)h * ow;
        o[0] = __fdiv_rn((float)clip8(s0), 255.0f);
        o[plane] = __fdiv_rn((float)clip8(s1), 255.0f);
        o[2 * plane] = __fdiv_rn((float)clip8(s2), 255.0f);
    } else {
        unsigned char *o = tmp + idx * 3;
        o[0] = (unsigned char)clip8(s0);
        o[1] = (unsigned char)clip8(s1);
        o[2] = (unsigned char)clip8(s2);
    }
}

// vertical pass + conversion: b x h x ow x 3 bytes -> b x 3 x oh x ow float32 / 255.  ksize == 0: no resampling at all
// (frames already have the target size): conversion only.
__global__ void __launch_bounds__(kThreads) ingest_v_kernel(const unsigned char *__restrict__ src, float *__restrict__ out,
                                                            const int *__restrict__ tab, int ksize, long long npix, int h, int oh, int ow)
{
    pdl_launch_dependents();
    pdl_wait();
    const long long idx = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= npix) return;
    const long long bo = idx / ow;                      // b * oh + yy
    const int xx = (int)(idx - bo * ow);
    const long long b = bo / oh;
    const int yy = (int)(bo - b * oh);
    int v0, v1, v2;
    if (ksize == 0) {
        const unsigned char *p = src + ((b * h + yy) * ow + xx) * 3;
        v0 = p[0]; v1 = p[1]; v2 = p[2];
    } else {
        const int *rec = tab + (size_t)yy * (2 + ksize);
        const int ymin = rec[0], n = rec[1];
        const unsigned char *p = src + ((b * h + ymin) * ow + xx) * 3;
        const size_t stride = (size_t)ow * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int y = 0; y < n; ++y) {
            const int k = rec[2 + y];
            s0 += (int)p[0] * k;
            s1 += (int)p[1] * k;
            s2 += (int)p[2] * k;
            p += stride;
        }
        v0 = clip8(s0); v1 = clip8(s1); v2 = clip8(s2);
    }
    float *o = out + ((size_t)b * 3 * oh + yy) * ow + xx;
    const size_t plane = (size_t)oh * ow;
    o[0] = __fdiv_rn((float)v0, 255.0f);
    o[plane] = __fdiv_rn((float)v1, 255.0f);
    o[2 * plane] = __fdiv_rn((float)v2, 255.0f);
}

struct IngestLayout {
    int kx, ky;                       // coefficients per output column / row (0: that axis is not resampled)
    size_t off_x, off_y, off_tmp, total;
};

static IngestLayout ingest_layout(long long b, int h, int w, int oh, int ow)
{
    IngestLayout L = {};
    L.kx = w != ow ? axis_ksize(w, ow) : 0;
    L.ky = h != oh ? axis_ksize(h, oh) : 0;
    size_t o = 0;
    L.off_x = o; o += L.kx ? sizeof(int) * (size_t)ow * (2 + L.kx) : 0; o = (o + 255) & ~(size_t)255;
    L.off_y = o; o += L.ky ? sizeof(int) * (size_t)oh * (2 + L.ky) : 0; o = (o + 255) & ~(size_t)255;
    L.off_tmp = o; o += (L.kx && L.ky) ? (size_t)b * h * ow * 3 : 0; o = (o + 255) & ~(size_t)255;
    L.total = o ? o : 256;
    return L;
}

}  // namespace stn

using namespace stn;

extern "C" {

long long loans_stn_ingest_workspace_bytes(int b, int h, int w, int oh, int ow)
{
    if (b < 0 || h < 1 || w < 1 || oh < 1 || ow < 1) return -1;
    return (long long)ingest_layout(b, h, w, oh, ow).total;
}

int loans_stn_ingest_prepare(void *workspace, int h, int w, int oh, int ow, void *stream)
{
    const char *what = "loans_stn_ingest_prepare";
    if (h < 1 || w < 1 || oh < 1 || ow < 1) return set_error("%s: bad dimensions h=%d w=%d oh=%d ow=%d", what, h, w, oh, ow);
    if (!workspace) return set_error("%s: workspace is NULL", what);
    // the tables live in a process-wide cache so that the host memory outlives the asynchronous upload
    static std::mutex mu;
    static std::map<std::pair<int, int>, std::vector<int>> cache;
    const IngestLayout L = ingest_layout(0, h, w, oh, ow);
    const int in_size[2] = {w, h}, out_size[2] = {ow, oh};
    const size_t off[2] = {L.off_x, L.off_y};
    for (int a = 0; a < 2; ++a) {
        if (in_size[a] == out_size[a]) continue;
        const std::vector<int> *tab;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto key = std::make_pair(in_size[a], out_size[a]);
            auto it = cache.find(key);
            if (it == cache.end()) it = cache.emplace(key, axis_table(in_size[a], out_size[a])).first;
            tab = &it->second;
        }
        const cudaError_t e = cudaMemcpyAsync(static_cast<char *>(workspace) + off[a], tab->data(), sizeof(int) * tab->size(),
                                              cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) return set_error("%s: uploading the coefficient table failed: %s", what, cudaGetErrorString(e));
    }
    return 0;
}

int loans_stn_ingest_u8(const unsigned char *frames_hwc, float *out_nchw, const void *workspace,
                        int b, int h, int w, int oh, int ow, void *stream)
{
    const char *what = "loans_stn_ingest_u8";
    if (b < 0 || h < 1 || w < 1 || oh < 1 || ow < 1) return set_error("%s: bad dimensions b=%d h=%d w=%d oh=%d ow=%d", what, b, h, w, oh, ow);
    if (b == 0) return 0;
    if (!frames_hwc || !out_nchw || !workspace) return set_error("%s: NULL pointer", what);
    const IngestLayout L = ingest_layout(b, h, w, oh, ow);
    const char *ws = static_cast<const char *>(workspace);
    const int *tx = reinterpret_cast<const int *>(ws + L.off_x), *ty = reinterpret_cast<const int *>(ws + L.off_y);
    unsigned char *tmp = reinterpret_cast<unsigned char *>(const_cast<char *>(ws + L.off_tmp));
    cudaStream_t s = (cudaStream_t)stream;
    auto launch = [&](auto kernel, long long npix, auto... args) -> cudaError_t {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((npix + kThreads - 1) / kThreads));
        cfg.blockDim = dim3(kThreads);
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 0);
        count_launch();
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    cudaError_t e = cudaSuccess;
    const long long rows = (long long)b * h;
    if ((rows * (long long)(w > ow ? w : ow) + kThreads) / kThreads > 0x7fffffffLL) return set_error("%s: batch too large", what);
    if (L.kx && L.ky) {
        e = launch(ingest_h_kernel<false>, rows * ow, frames_hwc, tmp, (float *)nullptr, tx, L.kx, rows * ow, h, w, ow);
        if (e == cudaSuccess)
            e = launch(ingest_v_kernel, (long long)b * oh * ow, (const unsigned char *)tmp, out_nchw, ty, L.ky, (long long)b * oh * ow, h, oh, ow);
    } else if (L.kx) {
        e = launch(ingest_h_kernel<true>, rows * ow, frames_hwc, (unsigned char *)nullptr, out_nchw, tx, L.kx, rows * ow, h, w, ow);
    } else {
        e = launch(ingest_v_kernel, (long long)b * oh * ow, frames_hwc, out_nchw, ty, L.ky, (long long)b * oh * ow, h, oh, ow);
    }
    note_kernel("ingest");
    if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
