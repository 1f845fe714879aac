// stn_prepare.cu -- SURVEY.md section 8(f) rank 1: SheepLocalizer.prepare_images on the device.
//
// The reference turns the [0,1] RGB frame batch into the ResNet trunk's input with a device -> host -> device round
// trip and a per-image Python loop (sheep/sheep_localizer.py:45,72-82):
//     input_images = self.prepare_images(images.copy() * 255)
//     prepare_images: [chainer.links.model.vision.resnet.prepare(image.data, size=None) for image in images]
// and resnet.prepare (chainer 4.1.0, third party) does, per image:  image.astype(numpy.uint8) -> PIL RGB ->
// numpy float32 -> [:, :, ::-1] (RGB -> BGR) -> minus the mean pixel [103.063, 115.903, 123.152] -> (3,H,W).
// So  out[b, c, i, j] = float(uint8(scaled[b, 2 - c, i, j])) - mean[c]  with C-cast truncation toward zero.
// One streaming pass: 128-bit loads and stores, grid sized to the machine, HBM-bound (4 B read + 4 B written / element).
// `scale` folds the caller's `images * 255` in (scale = 255, raw [0,1] frames in) or leaves it to the caller
// (scale = 1, the drop-in replacement of the method, which receives the scaled frames).
#include "stn_common.cuh"

namespace stn {

__device__ __forceinline__ float quantise_u8(float v)
{
    // numpy float32 -> uint8 on the host is (unsigned char)(int)v: truncation toward zero, low byte of the int32.
    // Defined input range is [0, 256) (frames are [0,1] * 255); outside it this keeps the low byte like x86 does.
    return (float)(__float2int_rz(v) & 0xff);
}

__global__ void __launch_bounds__(kThreads) prepare_images_kernel(const float *__restrict__ x, float *__restrict__ out, float scale,
                                                                   int plane4, int planes, int vec)
{
    // blockIdx.y walks the output planes (b, c), which read input plane (b, 2 - c); blockIdx.x strides through the plane
    // in float4 (vec) or float items
    pdl_launch_dependents();
    pdl_wait();
    for (int pl = blockIdx.y; pl < planes; pl += gridDim.y) {
        const int b = pl / 3, c = pl - 3 * b;
        const float mean = c == 0 ? 103.063f : (c == 1 ? 115.903f : 123.152f);
        const size_t src = ((size_t)b * 3 + (2 - c)) * plane4, dst = (size_t)pl * plane4;
        for (int o = blockIdx.x * kThreads + threadIdx.x; o < plane4; o += gridDim.x * kThreads) {
            if (vec) {
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(x) + src + o);
                float4 r;
                r.x = f_sub(quantise_u8(f_mul(v.x, scale)), mean);
                r.y = f_sub(quantise_u8(f_mul(v.y, scale)), mean);
                r.z = f_sub(quantise_u8(f_mul(v.z, scale)), mean);
                r.w = f_sub(quantise_u8(f_mul(v.w, scale)), mean);
                reinterpret_cast<float4 *>(out)[dst + o] = r;
            } else {
                out[dst + o] = f_sub(quantise_u8(f_mul(__ldcs(x + src + o), scale)), mean);
            }
        }
    }
}

int launch_prepare_images(const float *x, float *out, float scale, int b, int h, int w, cudaStream_t stream)
{
    const long long plane = (long long)h * w;                  // <= 2^30 (checked by the caller)
    const int vec = (plane % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 1 : 0;
    const int plane4 = (int)(vec ? plane / 4 : plane);
    const long long planes = 3LL * b;
    if (planes > 0x7fffffffLL) return set_error("prepare_images: too many planes");
    // four items per thread where the plane is big enough; about 16 CTAs per SM in total, grid-stride beyond
    int gx = (plane4 + 4 * kThreads - 1) / (4 * kThreads);
    int gy = (int)(planes < 65535 ? planes : 65535);
    const long long want = 16LL * num_sms();
    if ((long long)gx * gy > want) {
        gx = (int)((want + gy - 1) / gy);
        if (gx < 1) gx = 1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)gx, (unsigned)gy);
    cfg.blockDim = dim3(kThreads);
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    cfg.numAttrs = fill_launch_attrs(attr, 0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, prepare_images_kernel, x, out, scale, plane4, (int)planes, vec);
    count_launch();
    note_kernel("prepare_images_kernel");
    if (e != cudaSuccess) return set_error("prepare_images launch failed: %s", cudaGetErrorString(e));
    return 0;
}

}  // namespace stn
