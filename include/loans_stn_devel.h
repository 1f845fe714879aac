/*
 * loans_stn_devel.h -- test hooks and A/B switches of libloans_stn.so.  NOT part of the operator ABI (loans_stn.h): nothing a
 * drop-in caller needs is here, and results are bit-identical whatever these are set to.
 *
 * Always available (tests use them to run, on small inputs against the oracle, the kernel variants the automatic rules
 * pick at other sizes):
 *   LOANS_STN_CFG_BAND_CS / _ROWS / _TILE_KB / _VARIANT: CTAs per crop, crop rows per band, shared-memory tile budget in
 *   KiB, kernel variant (1: CTA bands, 3: row bands) of the band backward; 0 = automatic.
 *   LOANS_STN_CFG_KFRAME_ROWS: frame rows per CTA of the several-crops-per-frame gx kernel (stn_kframe.cu); 0 = automatic
 *     (bits 16 and up: KiB of shared-memory padding per CTA instead of the automatic residency rule).
 *   LOANS_STN_CFG_KFRAME_SINGLE != 0: that kernel (theta kernel + row-owner gx) also for ONE crop per frame, ahead of the band
 *     backward (A/B arm).
 *
 * Only in a -DSTN_DEVEL build (make -C loans_b200/csrc EXTRA=-DSTN_DEVEL; the product build answers them with an error):
 *   LOANS_STN_CFG_TMA_FORWARD != 0: forward of axis-aligned crops (mask01 == 0, w % 4 == 0) through the AxisTap-table +
 *     TMA-bulk-copy-staged kernel (stn_separable.cu) -- measured slower than the direct gather at every BASELINE size;
 *   LOANS_STN_CFG_THETA_FIRST: general backward, theta-role CTAs scheduled before the gx-role CTAs;
 *   LOANS_STN_CFG_GX_TILES_PER_WARP: general backward, frame tiles per warp of the gx role (0 = automatic);
 *   LOANS_STN_CFG_THETA_ONLY_KERNEL (default 1): gx == NULL through the theta-only kernels; 0: the two-role kernel;
 *   LOANS_STN_CFG_FWD_PX_PER_CTA: forward, crop pixels per CTA (rounded up to 256; 0 = automatic);
 *   band variant 2 (CTA bands with two crop pixels in flight per thread) and the per-CTA stage timestamps
 *   (-DSTN_BAND_TRACE: loans_stn_debug_band_trace / loans_stn_debug_bwd_trace).
 */
#ifndef LOANS_STN_DEVEL_H_
#define LOANS_STN_DEVEL_H_

#define LOANS_STN_CFG_TMA_FORWARD 2
#define LOANS_STN_CFG_BAND_CS 4
#define LOANS_STN_CFG_BAND_ROWS 5
#define LOANS_STN_CFG_BAND_TILE_KB 6
#define LOANS_STN_CFG_BAND_VARIANT 7
#define LOANS_STN_CFG_THETA_FIRST 9
#define LOANS_STN_CFG_GX_TILES_PER_WARP 10
#define LOANS_STN_CFG_THETA_ONLY_KERNEL 11
#define LOANS_STN_CFG_FWD_PX_PER_CTA 12
#define LOANS_STN_CFG_KFRAME_ROWS 13
#define LOANS_STN_CFG_KFRAME_SINGLE 14

/* measurement aid of bench.py's "floor" block (stn_probe.cu): what a kernel of our launch shape costs before any STN
 * arithmetic -- mode 0 an empty kernel with our launch attributes, mode 1 two dependent DRAM round trips and a store,
 * mode 2 in_bytes read + out_bytes written as plain coalesced 16-byte accesses, mode 3 the chain of mode 1 and then the bytes of
 * mode 2 (an ideal fused kernel: same bytes, same two dependent round trips, no arithmetic).  ctas x 256 threads on `stream`. */
#ifdef __cplusplus
extern "C"
#endif
int loans_stn_probe(int mode, const void *in, void *out, long long in_bytes, long long out_bytes, int ctas, void *stream);

#endif /* LOANS_STN_DEVEL_H_ */
