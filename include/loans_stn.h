/*
 * loans_stn.h -- C ABI of the B200-native STN crop stage of LoANs (libloans_stn.so).
 *
 * This is the drop-in boundary for ONE path of the reference (Bartzi/loans): the three operator calls at
 * the end of the localizer forward,
 *
 *     transform_params = rotation_dropout(F.reshape(transform_params, (-1, 2, 3)), ratio=0.0)
 *     points = F.spatial_transformer_grid(transform_params, self.out_size)
 *     rois   = F.spatial_transformer_sampler(images, points)
 *                                    (reference sheep/sheep_localizer.py:61-63 and :169-171)
 *
 * and their backward passes.  The reference is pure Python on Chainer 4.1; it has no FFI of its own, so
 * every entry point below names the reference operator (file:line) it stands in for.  The reference-side
 * binding (Chainer FunctionNodes over ctypes + cupy pointers) is shown in INTEGRATION.md and shipped in
 * loans_b200/chainer_compat.py.
 *
 * Conventions (all entry points)
 *   - plain pointers and ints only; every pointer is a DEVICE pointer on the current CUDA device unless the
 *     function name ends in _host; tensors are C-contiguous, frames NCHW;
 *   - the library never allocates, frees or retains caller memory; outputs are fully overwritten
 *     (gx includes its zeros); work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream) and the call returns without synchronising;
 *   - return value 0 = success; non-zero = error, message via loans_stn_last_error() (thread-local).
 *     Shape/dtype violations are reported before anything is launched;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns an error.
 *
 * Shapes: n = number of crops = b * k, b = frames, k = crops per frame (crop i samples frame i / k;
 * Chainer's sampler only has k == 1), c channels, frame h x w, crop oh x ow.
 *   theta (n,2,3) f32 | grid (n,2,oh,ow) f32, channel 0 = x, 1 = y in [-1,1], (-1,-1) = centre of the top-left
 *   pixel | x (b,c,h,w) f32 | y, gy (n,c,oh,ow) f32 or bf16 | gx (b,c,h,w) f32 | gtheta (n,2,3) f32.
 */
#ifndef LOANS_STN_H_
#define LOANS_STN_H_

#ifdef __cplusplus
extern "C" {
#endif

#define LOANS_STN_ABI_VERSION 2

/* element type of the crops y / gy */
#define LOANS_STN_F32  0
#define LOANS_STN_BF16 1     /* bf16 = round-to-nearest-even of the fp32 result; no reference equivalent
                                (Chainer's sampler type-checks float32), defined by BASELINE.json config 3 */

int loans_stn_abi_version(void);
const char *loans_stn_last_error(void);
/* number of kernels this library has launched from the calling process so far (bench.py's gpu_launches) */
unsigned long long loans_stn_launch_count(void);
/* names of the kernels the LAST compute entry point called in this process launched, '+'-separated (e.g.
 * "stn_bwd_band_kernel/row"; process-wide, since a backward usually runs on the framework's autograd thread): lets tests
 * and bench.py state which kernel a dispatch rule picked */
const char *loans_stn_last_kernel(void);

/* process-wide switches (results are identical either way; further test hooks in loans_stn_devel.h).
 * LOANS_STN_CFG_FORCE_GENERAL != 0: never take a kernel written for axis-aligned crops.
 * LOANS_STN_CFG_BAND_BACKWARD: backward of axis-aligned crops (one crop per frame, gx wanted, w % 4 == 0) through the
 *   band kernels (stn_band.cu).  -1 (default): where they measured faster (row bands: narrow frames and enough crops to
 *   fill the machine, e.g. 64 crops of 64 rows from 224-px frames; CTA bands: frame rows of >= 4 KiB, e.g. 512-px
 *   frames), 1: whenever they apply, 0: never.  gx, ggrid: same values; gtheta: same sums in a different order (both
 *   within the 1e-4 bar).
 * LOANS_STN_CFG_PDL (default 1): the fused kernels are launched with programmatic stream serialisation: their CTAs may
 *   become resident, and fill their shared-memory tables, while the previous kernel of the stream is still draining; they
 *   touch global memory only after griddepcontrol.wait, so stream order is preserved whatever the neighbouring kernels
 *   are.  0: plain launches. */
#define LOANS_STN_CFG_FORCE_GENERAL 1
#define LOANS_STN_CFG_BAND_BACKWARD 3
#define LOANS_STN_CFG_PDL 8
int loans_stn_configure(int key, int value);

/* ---- a1  rotation_dropout forward AND backward: out = in * mask, mask = 1 except [.,0,1] = [.,1,0] = mask01.
 *      Replaces RotationDropout.forward / .backward, reference functions/rotation_droput.py:26-45 / :47-48.
 *      mask01 is drawn ON THE HOST by the caller (train: float(rand(1) < ratio), one draw per call, :41;
 *      test: ratio, :33-35) so the RNG stream stays the host framework's.  In-place (out == in) allowed. */
int loans_stn_rotation_dropout(const float *theta_in, float mask01, float *theta_out, int n, void *stream);

/* ---- f1  SheepLocalizer.prepare_images (reference sheep/sheep_localizer.py:45,72-82): the localizer's per-step
 *      device -> host -> device round trip with a per-image Python loop around chainer's resnet.prepare(image,
 *      size=None), as one streaming kernel:  out[b,ch,i,j] = float(uint8(x[b,2-ch,i,j] * scale)) - mean[ch],
 *      uint8() = C-cast truncation, mean = (103.063, 115.903, 123.152) for output channels B, G, R.
 *      scale = 1: x is what the reference hands the method (`images.copy() * 255`); scale = 255: x is the raw [0,1]
 *      frame batch and the caller's multiply is folded in.  x, out (b,3,h,w) f32, out != x.  No backward: the
 *      reference takes no gradient through it (it builds new arrays on the host). */
int loans_stn_prepare_images(const float *x, float scale, float *out, int b, int c, int h, int w, void *stream);

/* ---- f4  the loader's frame path (reference common/datasets/image_dataset.py:16-28 resize_image, :98 and :181 `image / 255`):
 *          Image.fromarray(uint8 HWC).convert('RGB').resize((ow, oh), Image.LANCZOS) -> float32 CHW -> / 255
 *      for a batch of decoded frames on the device.  frames_hwc (b,h,w,3) uint8 -> out_nchw (b,3,oh,ow) f32 in [0,1], bit for
 *      bit what Pillow's 8-bit Lanczos resampling gives (two fixed-point passes, coefficient tables computed on the host in
 *      float64 exactly as Pillow does).  h == oh and w == ow: the conversion alone.
 *      workspace: caller-owned device memory of loans_stn_ingest_workspace_bytes(b,h,w,oh,ow) bytes (coefficient tables +
 *      the intermediate image); loans_stn_ingest_prepare fills the tables for (h,w,oh,ow) once (a host-to-device copy ordered
 *      on `stream`: not capturable into a CUDA graph); loans_stn_ingest_u8 then only launches kernels, any number of times. */
long long loans_stn_ingest_workspace_bytes(int b, int h, int w, int oh, int ow);
int loans_stn_ingest_prepare(void *workspace, int h, int w, int oh, int ow, void *stream);
int loans_stn_ingest_u8(const unsigned char *frames_hwc, float *out_nchw, const void *workspace,
                        int b, int h, int w, int oh, int ow, void *stream);

/* ---- a2  F.spatial_transformer_grid forward / backward (call site sheep/sheep_localizer.py:62,170;
 *      arithmetic in chainer 4.1.0 chainer/functions/array/spatial_transformer_grid.py, restated in
 *      oracle/stn_numpy.py:grid_forward/grid_backward). */
int loans_stn_grid_fwd(const float *theta, float *grid, int n, int oh, int ow, void *stream);
int loans_stn_grid_bwd(const float *ggrid, float *gtheta, int n, int oh, int ow, void *stream);

/* ---- a3/a4  F.spatial_transformer_sampler forward / backward on an EXPLICIT (arbitrary) grid
 *      (call site sheep/sheep_localizer.py:63,171; chainer 4.1.0 .../spatial_transformer_sampler.py,
 *      restated in oracle/stn_numpy.py:sampler_forward/sampler_backward).
 *      bwd: gx and/or ggrid may be NULL (skipped).  gx is zero-filled and scatter-added with
 *      warp-aggregated float atomics (order-dependent in the last bits). */
int loans_stn_sampler_fwd(const float *x, const float *grid, void *y,
                          int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream);
int loans_stn_sampler_bwd(const float *x, const float *grid, const void *gy, float *gx, float *ggrid,
                          int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream);

/* ---- a5  the fused composite: rotation_dropout -> grid -> sampler in ONE kernel per direction
 *      (reference sheep/sheep_localizer.py:61-63 as a whole).
 *      fwd: y always; grid may be NULL (not materialised).
 *      bwd: gtheta always (already multiplied by the rotation mask, i.e. the gradient w.r.t. the
 *           un-masked theta the localizer predicted); ggrid_upstream (n,2,oh,ow) or NULL is the gradient
 *           arriving on the grid output from other consumers (the corner regularisers, reference
 *           common/utils.py:142-178,301-316) and is folded into gtheta; gx may be NULL (LoANs never needs
 *           it: frames do not require grad); ggrid_out may be NULL.
 *           gx is produced by a deterministic gather (each element written once, no atomics). */
int loans_stn_crop_fwd(const float *x, const float *theta, float mask01, void *y, float *grid,
                       int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream);
int loans_stn_crop_bwd(const float *x, const float *theta, float mask01, const void *gy,
                       const float *ggrid_upstream, float *gtheta, float *gx, float *ggrid_out,
                       int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream);

/* ---- f2  the composite with the grid reduced to its four corner points.
 *      Everything LoANs does with `points` besides sampling reads only grid[:, :, {0,-1}, {0,-1}]: the direction and
 *      out-of-image regularisers (reference common/utils.py:141-178,301-316), extract_corners (sheep/sheep_localizer.py:
 *      84-91), the evaluator (sheep/sheep_evaluator.py:17-30).  corners (n,2,2,2) f32 = grid[:, :, {0,oh-1}, {0,ow-1}],
 *      bit-identical to those elements of the dense grid and itself a valid (n,2,2,2) `points` array for every one of those
 *      consumers (they take height and width from its shape).  The dense grid (8*oh*ow bytes per crop, written in the
 *      forward and read back as ggrid_upstream in the backward) is then never materialised.
 *      bwd: gcorners (n,2,2,2) or NULL = gradient arriving on those four points, folded into gtheta. */
int loans_stn_crop_fwd_corners(const float *x, const float *theta, float mask01, void *y, float *corners,
                               int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream);
int loans_stn_crop_bwd_corners(const float *x, const float *theta, float mask01, const void *gy,
                               const float *gcorners, float *gtheta, float *gx,
                               int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream);

/* ---- f3  the composite with every optional input / output and the localizer's grayscale epilogue (SURVEY.md section
 *      8f rank 3; reference sheep/sheep_localizer.py:65-68, `transform_rois_to_grayscale`):
 *          b, g, r = F.split_axis(rois, 3, axis=1);  rois = 0.299 * r + 0.587 * g + 0.114 * b
 *      flags & LOANS_STN_FLAG_GRAY (c == 3): y / gy are (n,1,oh,ow); the three channels are mixed in registers (float32
 *      products summed left to right, channel 0 taken as b, channel 2 as r, exactly as above) and the backward hands
 *      coef[ch] * gy to channel ch, so the 3-channel crops are never written or read.  grid, corners (fwd) and
 *      ggrid_upstream, gcorners, gx, ggrid_out (bwd) may each be NULL.  With flags == 0 these are the entry points above. */
#define LOANS_STN_FLAG_GRAY 1
/* LOANS_STN_FLAG_UPRIGHT (bwd): the caller asserts that theta's rotation terms are already zero although mask01 != 0 --
 *      theta is the OUTPUT of a rotation dropout that zeroed them (the strict three-node use of the reference operators:
 *      rotation_dropout -> grid -> sampler with the masked theta materialised in between).  It selects the same kernels
 *      as mask01 == 0.  A hint only: those kernels test every crop's own rotation terms on the device and run the general
 *      roles for a crop that is rotated after all, in the same launch -- a wrong hint costs time, never correctness. */
#define LOANS_STN_FLAG_UPRIGHT 2
/* LOANS_STN_FLAG_NHWC4 (c == 3, bf16 crops, no GRAY): y / gy are channels-last with the channel count padded to four,
 *      (n, oh, ow, 4) bf16 -- one 8-byte store / load per crop pixel, the layout a tensor-core first convolution of the
 *      assessor consumes (reference common/net.py:15-25,77-90 is the consumer; Chainer itself only has float32 NCHW).
 *      fwd writes the padding channel as +0; bwd ignores it.  Values are those of the bf16 NCHW crops. */
#define LOANS_STN_FLAG_NHWC4 4
#define LOANS_STN_FLAGS_ALL 7
int loans_stn_crop_fwd_ex(const float *x, const float *theta, float mask01, void *y, float *grid, float *corners,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int y_dtype, void *stream);
int loans_stn_crop_bwd_ex(const float *x, const float *theta, float mask01, const void *gy, const float *ggrid_upstream,
                          const float *gcorners, float *gtheta, float *gx, float *ggrid_out,
                          int flags, int n, int k, int c, int h, int w, int oh, int ow, int gy_dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LOANS_STN_H_ */
