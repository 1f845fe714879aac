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
// height does not change).  A CTA stages kIngestRows source rows in shared memory with coalesced 16-byte loads (the taps of
// neighbouring output pixels overlap: every source byte is read from DRAM once) and its threads walk the rows' output pixels.
constexpr int kIngestRows = 8;

template <bool TO_FLOAT>
__global__ void __launch_bounds__(kThreads) ingest_h_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ tmp,
                                                            float *__restrict__ out, const int *__restrict__ tab, int ksize,
                                                            long long rows, int h, int w, int ow)
{
    extern __shared__ __align__(16) unsigned char srow[];             // kIngestRows x pitch source bytes | kIngestRows x ow * 3 output bytes
    pdl_launch_dependents();
    pdl_wait();
    const long long r0 = (long long)blockIdx.x * kIngestRows;
    const int nr = (int)min((long long)kIngestRows, rows - r0);
    const int wb = w * 3, pitch = (wb + 15) & ~15;
    {   // stage: the nr rows are contiguous in the source (HWC, rows back to back)
        const unsigned char *g = src + r0 * wb;
        const long long total = (long long)nr * wb;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && (wb & 15) == 0) {
            const uint4 *g4 = reinterpret_cast<const uint4 *>(g);
            uint4 *s4 = reinterpret_cast<uint4 *>(srow);
            for (int e = threadIdx.x; e < (int)(total >> 4); e += kThreads) s4[e] = __ldg(g4 + e);       // pitch == wb here
        } else {
            for (int e = threadIdx.x; e < (int)total; e += kThreads) {
                const int r = e / wb, c = e - r * wb;
                srow[r * pitch + c] = __ldg(g + e);
            }
        }
    }
    __syncthreads();
    unsigned char *sout = srow + kIngestRows * pitch;                  // the rows' output bytes, back to back as in tmp
    for (int it = threadIdx.x; it < nr * ow; it += kThreads) {
        const int r = it / ow, xx = it - r * ow;
        const int *rec = tab + (size_t)xx * (2 + ksize);
        const int xmin = __ldg(rec), n = __ldg(rec + 1);
        const unsigned char *p = srow + r * pitch + xmin * 3;
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
        for (int x = 0; x < n; ++x) {
            const int k = __ldg(rec + 2 + x);
            s0 += (int)p[3 * x] * k;
            s1 += (int)p[3 * x + 1] * k;
            s2 += (int)p[3 * x + 2] * k;
        }
        const long long row = r0 + r;
        if (TO_FLOAT) {
            const long long b = row / h;
            const int y = (int)(row - b * h);
            float *o = out + ((size_t)b * 3 * h + y) * ow + xx;
            const size_t plane = (size_t)h * ow;
            o[0] = __fdiv_rn((float)clip8(s0), 255.0f);
            o[plane] = __fdiv_rn((float)clip8(s1), 255.0f);
            o[2 * plane] = __fdiv_rn((float)clip8(s2), 255.0f);
        } else {
            unsigned char *o = sout + it * 3;
            o[0] = (unsigned char)clip8(s0);
            o[1] = (unsigned char)clip8(s1);
            o[2] = (unsigned char)clip8(s2);
        }
    }
    if (!TO_FLOAT) {
        __syncthreads();
        unsigned char *g = tmp + r0 * ow * 3;
        const int total = nr * ow * 3;
        if ((reinterpret_cast<uintptr_t>(g) & 15) == 0 && ((kIngestRows * pitch) & 15) == 0) {
            uint4 *g4 = reinterpret_cast<uint4 *>(g);
            const uint4 *s4 = reinterpret_cast<const uint4 *>(sout);
            for (int e = threadIdx.x; e < (total >> 4); e += kThreads) g4[e] = s4[e];
            for (int e = (total & ~15) + threadIdx.x; e < total; e += kThreads) g[e] = sout[e];
        } else {
            for (int e = threadIdx.x; e < total; e += kThreads) g[e] = sout[e];
        }
    }
}

// vertical pass + conversion: b x h x ow x 3 bytes -> b x 3 x oh x ow float32 / 255.  One CTA per output row: the row's
// coefficients sit in shared memory, a thread takes four consecutive BYTE columns (pixel x channel, channel fastest, as the
// source is laid out) with one 32-bit load per tap row, and the finished row goes through shared memory so that each channel
// plane is written with coalesced float stores.  ksize == 0: no resampling at all (frames already have the target size).
__global__ void __launch_bounds__(kThreads) ingest_v_kernel(const unsigned char *__restrict__ src, float *__restrict__ out,
                                                            const int *__restrict__ tab, int ksize, int h, int oh, int ow)
{
    extern __shared__ __align__(16) unsigned char vsm[];               // [ksize + 2 ints] [ow * 3 floats]
    int *rec = reinterpret_cast<int *>(vsm);
    float *res = reinterpret_cast<float *>(vsm + (((size_t)(ksize + 2) * sizeof(int) + 15) & ~(size_t)15));
    pdl_launch_dependents();
    pdl_wait();
    const long long bo = blockIdx.x;                    // b * oh + yy
    const long long b = bo / oh;
    const int yy = (int)(bo - b * oh);
    const int owb = ow * 3;
    if (ksize) {
        for (int e = threadIdx.x; e < ksize + 2; e += kThreads) rec[e] = __ldg(tab + (size_t)yy * (2 + ksize) + e);
        __syncthreads();
    }
    const int ymin = ksize ? rec[0] : yy, n = ksize ? rec[1] : 1;
    const unsigned char *base = src + (b * h + ymin) * owb;
    const bool words = (owb & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0;
    for (int q = 4 * threadIdx.x; q < owb; q += 4 * kThreads) {
        int acc[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) acc[t] = ksize ? 1 << (kPrecisionBits - 1) : 0;
        const unsigned char *p = base + q;
        const int nb = min(4, owb - q);
        for (int y = 0; y < n; ++y) {
            const int k = ksize ? rec[2 + y] : 1;
            unsigned wv = 0;
            if (words) wv = __ldg(reinterpret_cast<const unsigned *>(p));
            else
                for (int t = 0; t < nb; ++t) wv |= (unsigned)__ldg(p + t) << (8 * t);
#pragma unroll
            for (int t = 0; t < 4; ++t) acc[t] += (int)((wv >> (8 * t)) & 0xffu) * k;
            p += owb;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (t < nb) res[q + t] = __fdiv_rn((float)(ksize ? clip8(acc[t]) : acc[t]), 255.0f);
    }
    __syncthreads();
    float *o = out + ((size_t)b * 3 * oh + yy) * ow;
    const size_t plane = (size_t)oh * ow;
    for (int e = threadIdx.x; e < owb; e += kThreads) {
        const int ch = e / ow, xx = e - ch * ow;
        o[ch * plane + xx] = res[xx * 3 + ch];
    }
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
    auto launch = [&](auto kernel, long long ctas, size_t smem, auto... args) -> cudaError_t {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)ctas);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 0);
        count_launch();
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    cudaError_t e = cudaSuccess;
    const long long rows = (long long)b * h;
    const long long h_ctas = (rows + kIngestRows - 1) / kIngestRows;
    const size_t h_smem = (size_t)kIngestRows * ((((size_t)w * 3 + 15) & ~(size_t)15) + (L.ky ? (size_t)ow * 3 : 0)) + 16;
    const long long v_ctas = (long long)b * oh;
    const size_t v_smem = (((size_t)(L.ky + 2) * sizeof(int) + 15) & ~(size_t)15) + sizeof(float) * (size_t)ow * 3;
    if (h_ctas > 0x7fffffffLL || v_ctas > 0x7fffffffLL) return set_error("%s: batch too large", what);
    if ((L.kx && h_smem > 48 * 1024) || v_smem > 48 * 1024)
        return set_error("%s: frame rows of %d -> %d pixels are too wide for the staging buffers", what, w, ow);
    if (L.kx && L.ky) {
        e = launch(ingest_h_kernel<false>, h_ctas, h_smem, frames_hwc, tmp, (float *)nullptr, tx, L.kx, rows, h, w, ow);
        if (e == cudaSuccess)
            e = launch(ingest_v_kernel, v_ctas, v_smem, (const unsigned char *)tmp, out_nchw, ty, L.ky, h, oh, ow);
    } else if (L.kx) {
        e = launch(ingest_h_kernel<true>, h_ctas, h_smem, frames_hwc, (unsigned char *)nullptr, out_nchw, tx, L.kx, rows, h, w, ow);
    } else {
        e = launch(ingest_v_kernel, v_ctas, v_smem, frames_hwc, out_nchw, ty, L.ky, h, oh, ow);
    }
    note_kernel("ingest");
    if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
