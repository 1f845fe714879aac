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
//
// How the dot products are evaluated (round 2, second version).  A tap is uint8 x int32 coefficient (|k| < 2^23).  Every
// coefficient is split on the host into three bytes, k = k0 + 2^8 k1 + 2^16 k2 (k0, k1 unsigned, k2 signed), so that four
// taps are ONE 32-bit word of pixels against one word of coefficient bytes per plane: three `dp4a` per four taps instead of
// four byte extractions and four multiply-adds, and sum(p k) = A0 + 2^8 A1 + 2^16 A2 exactly (integer arithmetic modulo 2^32;
// the true sum fits int32 as it does in Pillow).  For that the four taps must be neighbours in memory:
//   * horizontal pass: a CTA stages eight source rows in shared memory DE-INTERLEAVED (HWC -> one byte plane per channel,
//     three 32-bit loads and six PRMT per four pixels); a thread owns an output column -- its window starts at a word boundary
//     (the coefficient bytes are shifted by xmin & 3 on the host, zeros in front), its coefficient words live in registers
//     for all rows and channels -- and the intermediate image is written ROW-PLANAR (row, channel, column) so that
//   * vertical pass: a thread takes four neighbouring columns of one channel with one 32-bit load per tap row, transposes
//     four rows x four columns with eight PRMT, and accumulates twelve dp4a per sixteen taps; `/ 255` is a 256-entry table of
//     IEEE quotients in shared memory; the four results leave as one 16-byte store into the channel plane.
#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <type_traits>
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
// 32-bit words of four taps per output element, counted from its first tap
static int axis_words(int ksize) { return (ksize + 3) / 4; }

// Pillow's precompute_coeffs + normalize_coeffs_8bpc with box = (0, in_size): first tap, tap count and the fixed-point
// coefficients of every output element
struct AxisCoeffs {
    int ksize;
    std::vector<int> first, count, k;          // k: out_size x ksize
};
static AxisCoeffs axis_coeffs(int in_size, int out_size)
{
    const double scale = (double)in_size / out_size, filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = 3.0 * filterscale, ss = 1.0 / filterscale;
    AxisCoeffs A;
    A.ksize = (int)ceil(support) * 2 + 1;
    A.first.assign(out_size, 0);
    A.count.assign(out_size, 0);
    A.k.assign((size_t)out_size * A.ksize, 0);
    std::vector<double> k(A.ksize);
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
        A.first[xx] = xmin;
        A.count[xx] = xmax;
        for (int x = 0; x < xmax; ++x) {
            const double w = ww != 0.0 ? k[x] / ww : k[x];
            A.k[(size_t)xx * A.ksize + x] = w < 0 ? (int)(-0.5 + w * (1 << kPrecisionBits)) : (int)(0.5 + w * (1 << kPrecisionBits));
        }
    }
    return A;
}

// byte `plane` of coefficient k: k = b0 + 2^8 b1 + 2^16 b2 with b0, b1 in [0, 255] and b2 the signed rest (|k| < 2^23)
static unsigned coeff_byte(int k, int plane)
{
    const int k2 = k >> 16;                                 // floor (arithmetic shift)
    const int rem = k - k2 * 65536;                         // [0, 65535]
    return plane == 0 ? (unsigned)(rem & 255) : (plane == 1 ? (unsigned)(rem >> 8) : (unsigned)(k2 & 255));
}

// device table of the HORIZONTAL axis, ints: [ow] first tap xmin | [nw * 3][ow] coefficient words, word j plane p of column
// xx at (j * 3 + p) * ow + xx (neighbouring columns side by side: coalesced), byte t of word j = the coefficient byte of
// source column xmin + 4 j + t (the device shifts the pixel words into place: the window starts at any byte)
static std::vector<int> table_x(int in_size, int out_size)
{
    const AxisCoeffs A = axis_coeffs(in_size, out_size);
    const int nw = axis_words(A.ksize);
    std::vector<int> tab((size_t)out_size * (1 + 3 * nw), 0);
    for (int xx = 0; xx < out_size; ++xx) {
        tab[xx] = A.first[xx];
        for (int x = 0; x < A.count[xx]; ++x)
            for (int p = 0; p < 3; ++p)
                tab[(size_t)out_size + (size_t)((x >> 2) * 3 + p) * out_size + xx] |= (int)(coeff_byte(A.k[(size_t)xx * A.ksize + x], p) << (8 * (x & 3)));
    }
    return tab;
}
// device table of the VERTICAL axis, ints: per output row a record [first tap row, groups of four taps, ng_max x 3 words]
static std::vector<int> table_y(int in_size, int out_size)
{
    const AxisCoeffs A = axis_coeffs(in_size, out_size);
    const int ng = axis_words(A.ksize), rec = 2 + 3 * ng;
    std::vector<int> tab((size_t)out_size * rec, 0);
    for (int yy = 0; yy < out_size; ++yy) {
        int *r = tab.data() + (size_t)yy * rec;
        r[0] = A.first[yy];
        r[1] = (A.count[yy] + 3) / 4;
        for (int y = 0; y < A.count[yy]; ++y)
            for (int p = 0; p < 3; ++p) r[2 + (y >> 2) * 3 + p] |= (int)(coeff_byte(A.k[(size_t)yy * A.ksize + y], p) << (8 * (y & 3)));
    }
    return tab;
}

__device__ __forceinline__ unsigned dp4a_uu(unsigned a, unsigned b, unsigned c)
{
    unsigned d;
    asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b, int c)           // unsigned pixels x signed coefficient bytes
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// (1 << 21) + sum of the taps -> Pillow's clip8
__device__ __forceinline__ int clip8(unsigned a0, unsigned a1, int a2)
{
    const int acc = (int)(a0 + (a1 << 8) + ((unsigned)a2 << 16));
    const int v = acc >> kPrecisionBits;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}
constexpr unsigned kAccStart = 1u << (kPrecisionBits - 1);

// float32(v) / 255 for v = 0...255: IEEE quotients, computed once per CTA (kThreads == 256)
__device__ __forceinline__ void fill_quotients(float *lut) { lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f); }
static_assert(kThreads == 256, "the quotient table is filled by one thread per byte value");

constexpr int kIngestRows = 8;            // source rows per CTA of the horizontal pass

// horizontal pass: rows x w x 3 bytes -> rows x 3 x owp bytes, row-planar (or, TO_FLOAT, straight to the float32 NCHW output when
// the height does not change).  NW: coefficient words per output column held in registers (0: read from the table per use).
template <int NW, bool TO_FLOAT>
__global__ void __launch_bounds__(kThreads) ingest_h_kernel(const unsigned char *__restrict__ src, unsigned char *__restrict__ tmp,
                                                            float *__restrict__ out, const int *__restrict__ tab, int nw_rt,
                                                            long long rows, int h, int w, int ow, int owp, int wp, int tcols)
{
    extern __shared__ __align__(16) unsigned char hsm[];               // [kIngestRows][3][wp] source planes | [kIngestRows][3][owp] results
    __shared__ float lut[256];
    pdl_launch_dependents();
    if (TO_FLOAT) fill_quotients(lut);
    pdl_wait();
    const long long r0 = (long long)blockIdx.x * kIngestRows;
    const int nr = (int)min((long long)kIngestRows, rows - r0);
    const int wb = w * 3, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpw = wp >> 2;
    unsigned char *sout = hsm + (size_t)kIngestRows * 3 * wp;
    {   // stage and de-interleave: the nr rows are contiguous in the source (HWC, rows back to back); a warp takes a row
        const unsigned char *g = src + r0 * wb;
        if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 3) == 0) {
            const int gw = w >> 2;                                     // groups of four pixels = three words per row
            for (int r = warp; r < nr; r += kWarps) {
                const unsigned *g4 = reinterpret_cast<const unsigned *>(g) + r * 3 * gw;
                unsigned *d = reinterpret_cast<unsigned *>(hsm) + r * 3 * wpw;
                for (int q = lane; q < gw; q += 32) {
                    const unsigned w0 = __ldg(g4 + 3 * q), w1 = __ldg(g4 + 3 * q + 1), w2 = __ldg(g4 + 3 * q + 2);
                    d[q] = __byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210);
                    d[q + wpw] = __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210);
                    d[q + 2 * wpw] = __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410);
                }
            }
        } else {
            for (int r = warp; r < nr; r += kWarps)
                for (int x = lane; x < w; x += 32) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) hsm[(r * 3 + c) * wp + x] = __ldg(g + r * wb + 3 * x + c);
                }
        }
        // the slack behind each plane row is read (against zero coefficients): keep it defined; so are the pad bytes of the results
        {
            const int ws = (w + 3) >> 2, sw = wpw - ws;                 // the partly filled last word is rewritten whole: bytes >= w are slack
            unsigned *p32 = reinterpret_cast<unsigned *>(hsm);
            for (int e = threadIdx.x; e < nr * 3 * sw; e += kThreads) {
                const int pr = e / sw;
                p32[pr * wpw + ws + (e - pr * sw)] = 0;
            }
            if (w & 3)
                for (int pr = threadIdx.x; pr < nr * 3; pr += kThreads)
                    for (int x = w; x < 4 * ws; ++x) hsm[pr * wp + x] = 0;
            if (!TO_FLOAT && owp != ow)
                for (int pr = threadIdx.x; pr < nr * 3; pr += kThreads)
                    for (int x = ow; x < owp; ++x) sout[pr * owp + x] = 0;
        }
    }
    __syncthreads();
    const int nw = NW ? NW : nw_rt;
    // tcols is a multiple of 32 (or kThreads): whole warps per row lane; small integers, exact in float
    const int tcw = tcols >> 5, tr = (int)(((float)warp + 0.5f) * __frcp_rn((float)tcw)), tc = threadIdx.x - tr * tcols, trows = kWarps / tcw;
    if (tr < trows)
        for (int xx = tc; xx < ow; xx += tcols) {
            const int xmin = __ldg(tab + xx), sh = 8 * (xmin & 3);
            const int *kt = tab + ow + xx;
            unsigned kreg[NW ? 3 * NW : 1];
            if (NW) {
#pragma unroll
                for (int j = 0; j < 3 * NW; ++j) kreg[j] = (unsigned)__ldg(kt + j * ow);
            }
            const unsigned *d0 = reinterpret_cast<const unsigned *>(hsm) + (xmin >> 2);
            for (int r = tr; r < nr; r += trows) {
                unsigned a0[3], a1[3];
                int a2[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    a0[c] = kAccStart;
                    a1[c] = 0;
                    a2[c] = 0;
                }
                const unsigned *d = d0 + r * 3 * wpw;
                if (NW) {
                    unsigned lo[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) lo[c] = d[c * wpw];
#pragma unroll
                    for (int j = 0; j < NW; ++j) {
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const unsigned hi = d[c * wpw + j + 1];
                            const unsigned wv = __funnelshift_r(lo[c], hi, sh);          // pixels xmin + 4 j ... + 3
                            lo[c] = hi;
                            a0[c] = dp4a_uu(wv, kreg[3 * j], a0[c]);
                            a1[c] = dp4a_uu(wv, kreg[3 * j + 1], a1[c]);
                            a2[c] = dp4a_us(wv, kreg[3 * j + 2], a2[c]);
                        }
                    }
                } else {
                    unsigned lo[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c) lo[c] = d[c * wpw];
                    for (int j = 0; j < nw; ++j) {
                        const unsigned k0 = (unsigned)__ldg(kt + (3 * j) * ow), k1 = (unsigned)__ldg(kt + (3 * j + 1) * ow),
                                       k2 = (unsigned)__ldg(kt + (3 * j + 2) * ow);
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const unsigned hi = d[c * wpw + j + 1];
                            const unsigned wv = __funnelshift_r(lo[c], hi, sh);
                            lo[c] = hi;
                            a0[c] = dp4a_uu(wv, k0, a0[c]);
                            a1[c] = dp4a_uu(wv, k1, a1[c]);
                            a2[c] = dp4a_us(wv, k2, a2[c]);
                        }
                    }
                }
                if (TO_FLOAT) {
                    const long long row = r0 + r, b = row / h;
                    const int y = (int)(row - b * h);
                    float *o = out + ((size_t)b * 3 * h + y) * ow + xx;
                    const size_t plane = (size_t)h * ow;
#pragma unroll
                    for (int c = 0; c < 3; ++c) o[c * plane] = lut[clip8(a0[c], a1[c], a2[c])];
                } else {
                    unsigned char *so = sout + r * 3 * owp + xx;
#pragma unroll
                    for (int c = 0; c < 3; ++c) so[c * owp] = (unsigned char)clip8(a0[c], a1[c], a2[c]);
                }
            }
        }
    if (!TO_FLOAT) {
        __syncthreads();
        // the CTA's rows are contiguous in tmp as they are in shared memory; both 16-byte aligned (owp % 8 == 0, 8 rows x 3 planes)
        uint4 *g4 = reinterpret_cast<uint4 *>(tmp + r0 * 3 * owp);
        const uint4 *s4 = reinterpret_cast<const uint4 *>(sout);
        const int total = nr * 3 * owp;
        for (int e = threadIdx.x; e < (total >> 4); e += kThreads) g4[e] = s4[e];
        for (int e = (total & ~15) + threadIdx.x; e < total; e += kThreads) tmp[r0 * 3 * owp + e] = sout[e];
    }
}

// four rows x four byte columns -> four words of one column each (byte r = row r)
__device__ __forceinline__ void transpose4(unsigned ra, unsigned rb, unsigned rc, unsigned rd, unsigned *col)
{
    const unsigned t0 = __byte_perm(ra, rb, 0x5140), t1 = __byte_perm(rc, rd, 0x5140);
    const unsigned t2 = __byte_perm(ra, rb, 0x7362), t3 = __byte_perm(rc, rd, 0x7362);
    col[0] = __byte_perm(t0, t1, 0x5410);
    col[1] = __byte_perm(t0, t1, 0x7632);
    col[2] = __byte_perm(t2, t3, 0x5410);
    col[3] = __byte_perm(t2, t3, 0x7632);
}

constexpr int kIngestVItems = 2;          // items (groups of byte columns) a thread of the vertical pass works through

// vertical pass + conversion.  PLANAR: src is the row-planar intermediate image (b x h rows of 3 x owp bytes, owp % 8 == 0) and a
// thread's item is EIGHT neighbouring pixels of one channel (one 8-byte load per tap row); otherwise src is the HWC frame itself
// (the width does not change), an item is four byte columns = interleaved channels, loaded as a word where the rows are
// word-aligned (ALIGNED) and byte by byte where not.  A CTA takes kThreads * kIngestVItems consecutive items of one frame
// (blockIdx.y); item -> (output row, column group) by one division per thread, then incrementally.  Output (b,3,oh,ow) float32 =
// quotient table of the clipped sums.
template <bool PLANAR, bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 4) ingest_v_kernel(const unsigned char *__restrict__ src, float *__restrict__ out,
                                                            const int *__restrict__ tab, int ng_max, int b_total, int h, int oh,
                                                            int ow, int pitch, int ipr)
{
    constexpr int CW = PLANAR ? 2 : 1;                                 // words per item
    __shared__ float lut[256];
    pdl_launch_dependents();
    fill_quotients(lut);
    pdl_wait();
    __syncthreads();
    const int items = oh * ipr;                                        // per frame (ipr: items per output row)
    const int step_r = kThreads / ipr, step_q = kThreads - step_r * ipr;
    for (int b = blockIdx.y; b < b_total; b += gridDim.y) {
        const unsigned char *fb = src + (size_t)b * h * pitch;
        float *ob = out + (size_t)b * 3 * oh * ow;
        int id = blockIdx.x * (kThreads * kIngestVItems) + threadIdx.x;
        int yy = id / ipr, q = id - yy * ipr;
        for (int it = 0; it < kIngestVItems && id < items; ++it, id += kThreads) {
            const int *rec = tab + yy * (2 + 3 * ng_max);
            const int ymin = __ldg(rec), ng = __ldg(rec + 1);
            const unsigned char *base = fb + 4 * CW * q;
            unsigned a0[4 * CW], a1[4 * CW];
            int a2[4 * CW];
#pragma unroll
            for (int t = 0; t < 4 * CW; ++t) {
                a0[t] = kAccStart;
                a1[t] = 0;
                a2[t] = 0;
            }
#pragma unroll 1
            for (int g = 0; g < ng; ++g) {
                const int y0 = ymin + 4 * g;
                unsigned rw[4][CW];
                // rows behind the last tap carry zero coefficients: any row of the frame will do
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // the intermediate image has three rows of slack behind its last frame; the caller's frames have not
                    const unsigned char *p = base + (PLANAR ? y0 + k : min(y0 + k, h - 1)) * pitch;
                    if (PLANAR) {
                        const uint2 v = __ldg(reinterpret_cast<const uint2 *>(p));
                        rw[k][0] = v.x;
                        rw[k][CW - 1] = v.y;
                    } else if (ALIGNED) {
                        rw[k][0] = __ldg(reinterpret_cast<const unsigned *>(p));
                    } else {
                        const int nb = min(4, pitch - 4 * q);
                        unsigned v = 0;
                        for (int t = 0; t < nb; ++t) v |= (unsigned)__ldg(p + t) << (8 * t);
                        rw[k][0] = v;
                    }
                }
                const unsigned k0 = (unsigned)__ldg(rec + 2 + 3 * g), k1 = (unsigned)__ldg(rec + 3 + 3 * g), k2 = (unsigned)__ldg(rec + 4 + 3 * g);
#pragma unroll
                for (int cw = 0; cw < CW; ++cw) {
                    unsigned col[4];
                    transpose4(rw[0][cw], rw[1][cw], rw[2][cw], rw[3][cw], col);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        a0[4 * cw + t] = dp4a_uu(col[t], k0, a0[4 * cw + t]);
                        a1[4 * cw + t] = dp4a_uu(col[t], k1, a1[4 * cw + t]);
                        a2[4 * cw + t] = dp4a_us(col[t], k2, a2[4 * cw + t]);
                    }
                }
            }
            float v[4 * CW];
#pragma unroll
            for (int t = 0; t < 4 * CW; ++t) v[t] = lut[clip8(a0[t], a1[t], a2[t])];
            if (PLANAR) {
                const int ipc = ipr / 3;                               // items per channel row (owp / 8)
                const int c = (q >= ipc) + (q >= 2 * ipc), xx = 8 * (q - c * ipc);
                float *o = ob + (c * oh + yy) * ow + xx;
                if ((ow & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
                    *reinterpret_cast<float4 *>(o) = make_float4(v[0], v[1], v[2], v[3]);
                    if (xx + 4 < ow) *reinterpret_cast<float4 *>(o + 4) = make_float4(v[4 * CW - 4], v[4 * CW - 3], v[4 * CW - 2], v[4 * CW - 1]);
                } else {
#pragma unroll
                    for (int t = 0; t < 4 * CW; ++t)
                        if (xx + t < ow) o[t] = v[t];
                }
            } else {
                const int plane = oh * ow;
                float *o = ob + yy * ow;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int e = 4 * q + t, xx = e / 3;
                    if (e < 3 * ow) o[(e - 3 * xx) * plane + xx] = v[t];
                }
            }
            yy += step_r;
            q += step_q;
            if (q >= ipr) {
                q -= ipr;
                ++yy;
            }
        }
    }
}

// no resampling at all (frames already have the target size): uint8 HWC -> float32 NCHW through the quotient table, four pixels
// (three words in, one 16-byte store per channel plane) per thread
__global__ void __launch_bounds__(kThreads) ingest_convert_kernel(const unsigned char *__restrict__ src, float *__restrict__ out,
                                                                  long long groups, int hw, int vec)
{
    __shared__ float lut[256];
    pdl_launch_dependents();
    fill_quotients(lut);
    pdl_wait();
    __syncthreads();
    const long long id = (long long)blockIdx.x * kThreads + threadIdx.x;
    if (id >= groups) return;
    if (vec) {                                                         // hw % 4 == 0, aligned pointers: a group never straddles frames
        const long long px = 4 * id, b = px / hw;
        const int o0 = (int)(px - b * hw);
        const unsigned *g = reinterpret_cast<const unsigned *>(src) + 3 * id;
        const unsigned w0 = __ldg(g), w1 = __ldg(g + 1), w2 = __ldg(g + 2);
        const unsigned pl[3] = {__byte_perm(__byte_perm(w0, w1, 0x0630), w2, 0x5210), __byte_perm(__byte_perm(w0, w1, 0x0741), w2, 0x6210),
                                __byte_perm(__byte_perm(w0, w1, 0x0052), w2, 0x7410)};
#pragma unroll
        for (int c = 0; c < 3; ++c)
            *reinterpret_cast<float4 *>(out + ((size_t)b * 3 + c) * hw + o0) =
                make_float4(lut[pl[c] & 255u], lut[(pl[c] >> 8) & 255u], lut[(pl[c] >> 16) & 255u], lut[pl[c] >> 24]);
    } else {                                                           // one pixel per "group"
        const long long b = id / hw;
        const int o0 = (int)(id - b * hw);
#pragma unroll
        for (int c = 0; c < 3; ++c) out[((size_t)b * 3 + c) * hw + o0] = lut[__ldg(src + 3 * id + c)];
    }
}

struct IngestLayout {
    int kx, ky;                       // coefficients per output column / row (0: that axis is not resampled)
    int nwx, ngy;                     // coefficient words per output column / row
    int owp;                          // pitch of a channel row of the intermediate image (ow rounded up to eight bytes)
    size_t off_x, off_y, off_tmp, total;
};

static IngestLayout ingest_layout(long long b, int h, int w, int oh, int ow)
{
    IngestLayout L = {};
    L.kx = w != ow ? axis_ksize(w, ow) : 0;
    L.ky = h != oh ? axis_ksize(h, oh) : 0;
    L.nwx = L.kx ? axis_words(L.kx) : 0;
    L.ngy = L.ky ? axis_words(L.ky) : 0;
    L.owp = (ow + 7) & ~7;
    size_t o = 0;
    L.off_x = o; o += L.kx ? sizeof(int) * (size_t)ow * (1 + 3 * L.nwx) : 0; o = (o + 255) & ~(size_t)255;
    L.off_y = o; o += L.ky ? sizeof(int) * (size_t)oh * (2 + 3 * L.ngy) : 0; o = (o + 255) & ~(size_t)255;
    L.off_tmp = o; o += (L.kx && L.ky) ? ((size_t)b * h + 3) * 3 * L.owp : 0; o = (o + 255) & ~(size_t)255;   // + 3 rows the vertical pass may read
   
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
    static std::map<std::pair<int, std::pair<int, int>>, std::vector<int>> cache;
    const IngestLayout L = ingest_layout(0, h, w, oh, ow);
    const int in_size[2] = {w, h}, out_size[2] = {ow, oh};
    const size_t off[2] = {L.off_x, L.off_y};
    for (int a = 0; a < 2; ++a) {
        if (in_size[a] == out_size[a]) continue;
        const std::vector<int> *tab;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto key = std::make_pair(a, std::make_pair(in_size[a], out_size[a]));
            auto it = cache.find(key);
            if (it == cache.end()) it = cache.emplace(key, a == 0 ? table_x(in_size[a], out_size[a]) : table_y(in_size[a], out_size[a])).first;
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
    auto launch = [&](auto kernel, long long ctas, int ctas_y, size_t smem, auto... args) -> cudaError_t {
        if (ctas > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)ctas, (unsigned)ctas_y);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = s;
        cudaLaunchAttribute attr[2];
        cfg.attrs = attr;
        cfg.numAttrs = fill_launch_attrs(attr, 0);
        if (smem > 48 * 1024) {
            const cudaError_t g = grant_dynamic_smem((const void *)kernel, smem);
            if (g != cudaSuccess) return g;
        }
        count_launch();
        return cudaLaunchKernelEx(&cfg, kernel, args...);
    };
    cudaError_t e = cudaSuccess;
    const long long rows = (long long)b * h;
    const long long h_ctas = (rows + kIngestRows - 1) / kIngestRows;
    // plane pitch of the staged source rows: the window of the last column may start in the last word and reads nwx + 1 words
    const int wp = (((w + 3) & ~3) + 4 * (L.nwx + 1) + 15) & ~15;
    const size_t h_smem = (size_t)kIngestRows * 3 * wp + (L.ky ? (size_t)kIngestRows * 3 * L.owp : 0);
    const int tcols = ow >= kThreads ? kThreads : (ow <= 32 ? 32 : ((ow + 31) & ~31));
    if (L.kx && h_smem > 200 * 1024)
        return set_error("%s: frame rows of %d pixels are too wide for the staging buffers", what, w);
    // the kernels index inside a frame with 32-bit integers
    if ((long long)h * 3 * std::max(L.owp, w) > 0x7fffffffLL || (long long)oh * 3 * L.owp > 0x7fffffffLL)
        return set_error("%s: frame too large", what);
    auto v_ctas = [&](int ipr) { return ((long long)oh * ipr + kThreads * kIngestVItems - 1) / (kThreads * kIngestVItems); };
    const int v_ctas_y = b < 65535 ? b : 65535;
    auto launch_h = [&](auto to_float) -> cudaError_t {
        constexpr bool TF = decltype(to_float)::value;
        unsigned char *t = TF ? (unsigned char *)nullptr : tmp;
        float *o = TF ? out_nchw : (float *)nullptr;
#define STN_INGEST_H(NW)                                                                                                          \
    case NW:                                                                                                                      \
        return launch(ingest_h_kernel<NW, TF>, h_ctas, 1, h_smem, frames_hwc, t, o, tx, L.nwx, rows, h, w, ow, L.owp, wp, tcols)
        switch (L.nwx <= 8 ? L.nwx : 0) {
            STN_INGEST_H(2);
            STN_INGEST_H(3);
            STN_INGEST_H(4);
            STN_INGEST_H(5);
            STN_INGEST_H(6);
            STN_INGEST_H(7);
            STN_INGEST_H(8);
        default:
            return launch(ingest_h_kernel<0, TF>, h_ctas, 1, h_smem, frames_hwc, t, o, tx, L.nwx, rows, h, w, ow, L.owp, wp, tcols);
        }
#undef STN_INGEST_H
    };
    if (L.kx && L.ky) {
        e = launch_h(std::false_type());
        if (e == cudaSuccess)
            e = launch(ingest_v_kernel<true, true>, v_ctas(3 * (L.owp >> 3)), v_ctas_y, 0, (const unsigned char *)tmp, out_nchw, ty, L.ngy, b, h, oh, ow,
                       3 * L.owp, 3 * (L.owp >> 3));
    } else if (L.kx) {
        e = launch_h(std::true_type());
    } else if (L.ky) {
        const int ipr = (3 * ow + 3) >> 2;
        if (((3 * ow) & 3) == 0 && (reinterpret_cast<uintptr_t>(frames_hwc) & 3) == 0)
            e = launch(ingest_v_kernel<false, true>, v_ctas(ipr), v_ctas_y, 0, frames_hwc, out_nchw, ty, L.ngy, b, h, oh, ow, 3 * ow, ipr);
        else
            e = launch(ingest_v_kernel<false, false>, v_ctas(ipr), v_ctas_y, 0, frames_hwc, out_nchw, ty, L.ngy, b, h, oh, ow, 3 * ow, ipr);
    } else {
        const long long hw = (long long)h * w;
        if (hw > 0x7fffffffLL) return set_error("%s: frame too large", what);
        const int vec = (hw & 3) == 0 && (reinterpret_cast<uintptr_t>(frames_hwc) & 3) == 0 && (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0;
        const long long groups = vec ? (long long)b * hw / 4 : (long long)b * hw;
        e = launch(ingest_convert_kernel, (groups + kThreads - 1) / kThreads, 1, 0, frames_hwc, out_nchw, groups, (int)hw, vec);
    }
    note_kernel("ingest");
    if (e != cudaSuccess) return set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
