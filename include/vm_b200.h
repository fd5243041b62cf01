/*
 * vm_b200.h - C ABI of libvm_sm100a.so: the B200 (sm_100a) replacement for the per-frame
 * data path of tangih/video-matting (flow.py, tps.py, augmentation.py, reader.py).
 *
 * The reference has no FFI: its hot path is a set of Python module-level functions that
 * bottom out in OpenCV / SciPy / NumPy CPU kernels.  Each entry point below replaces one of
 * those call sites (reference file:line given per function) and is what a Python-side
 * binding (ctypes) loads.  Conventions:
 *
 *   - plain C: pointers + sizes, no C++/torch types.  Unless the name ends in `_host`, every
 *     data pointer is a DEVICE pointer on the current CUDA device, dense row-major,
 *     channel-last (OpenCV layout), no padding.
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream).  Calls
 *     only enqueue work; nothing allocates, frees or synchronises inside the library.
 *   - return value: 0 = VM_OK, otherwise a VM_ERR_* code; `vm_last_error_string()` returns
 *     a thread-local description.  Nothing throws across the boundary.
 *   - `status` arguments point to a DEVICE `int32[VM_STATUS_WORDS]` block the kernels OR /
 *     add into; the host binding reads it to reproduce the reference's exceptions
 *     (IndexError / ValueError in flow.correct_alpha) and to count knife-edge events.
 *   - flow fields are float32 (H,W,2) = (dx, dy) in pixels on the OUTPUT grid pointing into
 *     the SOURCE (reader.py:29, flow.py:13-17).
 */
#ifndef VM_B200_H
#define VM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VM_OK            0
#define VM_ERR_ARG       1   /* bad size / null pointer / unsupported combination        */
#define VM_ERR_CUDA      2   /* a CUDA runtime call or launch failed                     */

/* element types of generic (drop-in) entry points */
#define VM_U8   0
#define VM_F32  1
#define VM_F64  2

/* layout of a status block (device int32[VM_STATUS_WORDS]) */
#define VM_STATUS_WORDS        8
#define VM_STATUS_INDEX_ERR    0   /* # pixels where flow.py:46 would raise IndexError     */
#define VM_STATUS_NAN_ERR      1   /* # pixels where int(nan/inf) would raise ValueError   */
#define VM_STATUS_MASKED       2   /* # pixels zeroed by the consistency test              */
#define VM_STATUS_TPS_OUTSIDE  3   /* # TPS samples outside [0,n-1] (map_coordinates -> 0) */
#define VM_STATUS_SLOW_TILES   4   /* # tiles of a fused kernel that took the gather path  */
#define VM_STATUS_BAD_TABLE    5   /* # tiles skipped: axis tables inconsistent with a /2 grid */
#define VM_STATUS_NEAR_KNIFE   6   /* # resampled colour samples within 1e-9 of a rounding boundary (SURVEY 8a-6:
                                      these are the ones a ~1e-11 px transform difference could flip)      */

int         vm_version(void);
const char *vm_last_error_string(void);
/* Upload the log table used by the TPS kernels to the current device (idempotent, per
 * device; called lazily by the TPS entry points, exposed so that it can be done up front). */
int         vm_init(void);

/* ---- flow.warp_img / flow.warp_bgr : cv2.remap(INTER_LINEAR, BORDER_CONSTANT 0) ---------
 * reference: flow.py:9-18 (1 channel), flow.py:21-33 (3 channels, per-channel remap).
 * src is (sh, sw, channels) of `dtype`; flow and dst are on the (h, w) output grid.
 * uint8: bit-exact fixed-point bilinear; float32/float64: float32 table weights, sum in the
 * source precision, left to right (bit-equal to OpenCV 4.13).                              */
int vm_flow_warp(const void *src, int dtype, int channels, int sh, int sw,
                 const float *flow, int h, int w, void *dst, void *stream);

/* ---- flow.correct_alpha : forward/backward consistency mask ------------------------------
 * reference: flow.py:36-65.  `vm_occlusion_mask` evaluates flow.py:41-48 for every pixel and
 * writes mask[i,j] = (err > 15) as uint8; IndexError / NaN conditions are counted in
 * `status`.  `vm_apply_mask` is flow.py:49-50 (alpha[mask] = 0, in place).                 */
int vm_occlusion_mask(const float *backward, const float *forward, int h, int w,
                      uint8_t *mask, int32_t *status, void *stream);
int vm_apply_mask(void *alpha, int dtype, const uint8_t *mask, int64_t n, void *stream);

/* ---- reader.create_composite_image : I = alpha*F + (1-alpha)*B ---------------------------
 * reference: reader.py:72-79.  fg/bg (h,w,3) of fg_dtype/bg_dtype, alpha (h,w) float64,
 * out (h,w,3) float64, evaluated as tri*fg + (1-tri)*bg in float64.                        */
int vm_composite(const void *fg, int fg_dtype, const void *bg, int bg_dtype,
                 const double *alpha, int h, int w, double *out, void *stream);

/* ---- tps._make_warp / _calculate_f : coarse-grid TPS evaluation ---------------------------
 * reference: tps.py:101-110, 120-121 as called from tps.py:47-51.  For each of `n` frames:
 * ctrl (N,2) float64 control points (row, col), coef (N+3,2) float64 from the host solve
 * (tps.py:119), coarse point (k,l) = (k*step_x + x0, l*step_y + y0); writes coarse (n, 2, nx, ny)
 * float64.  U(r) = r^2 log r is evaluated as 0.5*r2*log(r2) in float64.                    */
int vm_tps_coarse(const double *ctrl, const double *coef, int n, int N,
                  int nx, int ny, double step_x, double step_y, double x0, double y0,
                  double *coarse, void *stream);

/* Axis tables of the bilinear up-sampling of the coarse transform (tps.py:55-63): for output
 * index i in [0, len]: frac = modf((steps-1)*i/len), i0, i1 = min(i0+1, steps-1).  They depend
 * on the frame size only and are computed once on the host by the binding.                 */
typedef struct { double frac; int32_t i0; int32_t i1; } vm_axis_entry;

/* ---- tps._make_inverse_warp : materialise the (h+1, w+1) transform ------------------------
 * reference: tps.py:55-74, same operation order.  out (2, h+1, w+1) float64.               */
int vm_tps_upsample(const double *coarse, int nx, int ny,
                    const vm_axis_entry *rows, const vm_axis_entry *cols, int h, int w,
                    double *out, void *stream);

/* ---- tps.warp_images : up-sample + scipy.ndimage.map_coordinates(order=1) ------------------
 * reference: tps.py:34 (+ 55-74).  src (sh, sw, channels) of dtype (uint8 or float64), dst
 * (oh, ow, channels), oh <= h+1, ow <= w+1 rows/cols of the transform.  uint8 output is
 * floor(v + 0.5) clamped, float64 is v (same operation order as scipy).                    */
int vm_tps_warp(const void *src, int dtype, int channels, int sh, int sw,
                const double *coarse, int nx, int ny,
                const vm_axis_entry *rows, const vm_axis_entry *cols,
                int oh, int ow, void *dst, int32_t *status, void *stream);
/* Same with the `interpolation_order` argument of tps.warp_images (tps.py:14,34): 1 = bilinear (above),
 * 0 = nearest neighbour as scipy does it: cval 0 outside [0, n-1], else the sample at floor(t + 1/2). */
int vm_tps_warp_order(const void *src, int dtype, int channels, int sh, int sw,
                      const double *coarse, int nx, int ny,
                      const vm_axis_entry *rows, const vm_axis_entry *cols,
                      int oh, int ow, void *dst, int32_t *status, int order, void *stream);

/* scipy.ndimage.map_coordinates(order=1, mode='constant', cval=0) with an explicit transform
 * (tps.py:34 when approximate_grid is None/1): t0/t1 (oh, ow) float64 row / column coords.  */
int vm_map_coordinates(const void *src, int dtype, int channels, int sh, int sw,
                       const double *t0, const double *t1, int oh, int ow, void *dst,
                       int32_t *status, void *stream);
int vm_map_coordinates_order(const void *src, int dtype, int channels, int sh, int sw,
                             const double *t0, const double *t1, int oh, int ow, void *dst,
                             int32_t *status, int order, void *stream);

/* ---- cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0), legacy fixed-point path -----------
 * reference: augmentation.py:59-62.  M is the forward 2x3 matrix (row-major, 6 doubles,
 * HOST pointer); inversion and the AB_BITS=10 coordinate generation follow OpenCV.         */
int vm_warp_affine(const void *src, int dtype, int channels, int sh, int sw,
                   const double *M_host, int dh, int dw, void *dst, void *stream);

/* ---- augmentation.change_illumination ----------------------------------------------------
 * reference: augmentation.py:88-99.  BGR2HSV integer model (exact), S/V gamma in float64 with truncation,
 * HSV2BGR bit-exact for OpenCV 4.13 (row f2): float32 sector formula with the contracted 1 - s*f, result
 * TRUNCATED in the SIMD body of every image row and rounded half-to-even in the row's scalar tail - cv2 converts
 * `hsv_vec` pixels per SIMD step (32 with AVX2, 16 with SSE, 64 with AVX-512; the binding probes its cv2), so
 * pixel x of a row of w pixels is "body" iff x < w - w % hsv_vec.                                          */
int vm_illumination_lut_rows(const uint8_t *bgr, int64_t rows, int w, const uint8_t *lut_host, int hsv_vec,
                             uint8_t *out, void *stream);
/* The buffer as ONE row of npx pixels, hsv_vec = 32 (kept for callers without a row structure).         */
int vm_change_illumination(const uint8_t *bgr, int64_t npx, double a, double b, double c,
                           uint8_t *out, void *stream);
/* Same with the 256-entry S/V transfer table supplied by the caller (HOST pointer), so that a
 * NumPy host can build it with the reference's own expression (augmentation.py:91-98).     */
int vm_illumination_lut(const uint8_t *bgr, int64_t npx, const uint8_t *lut_host,
                        uint8_t *out, void *stream);

/* ---- cv2.resize(uint8, dsize, INTER_LINEAR) (row f2) ----------------------------------------
 * reference: reader.py:41,53 (backgrounds to the frame size), augmentation.py:160.  n images (sh, sw, channels)
 * -> (dh, dw, channels), channels 1 / 3 / 4; bit-exact for OpenCV 4.13: 11-bit coefficients from float32
 * fractions, horizontal fraction forced to 0 at the row ends, rows clipped vertically, exact 2x reductions as
 * the 2x2 block mean (OpenCV switches INTER_LINEAR to INTER_AREA there).                              */
int vm_resize_u8(const uint8_t *src, int n, int sh, int sw, int channels, uint8_t *dst, int dh, int dw, void *stream);

/* ---- augmentation.object_size / fg_center -------------------------------------------------
 * reference: augmentation.py:10-21.  out (device uint64[3]) += {count(alpha != 0),
 * sum of row indices, sum of column indices}; caller zeroes `out` first.                   */
int vm_alpha_stats(const void *alpha, int dtype, int h, int w, unsigned long long *out,
                   void *stream);

/* ======================= fused clip-level kernels (canonical layouts) =====================
 * fg     : (n, h, w, 4) uint8  BGRA, alpha = A/255  (reader.py:16-17)
 * flows  : (n, h, w, 2) float32
 * bg     : (n_bg, h, w, 3) uint8 BGR, frame f uses bg[f % n_bg]
 */

/* warp_bgr + warp_img + correct_alpha in one pass (flow.py:9-65).
 * out_bgr (n,h,w,3) uint8 bit-exact; out_alpha (n,h,w) float32 (<= 1e-6 rel of the float64
 * reference).  forward may be NULL (no consistency test).  27 B/px of HBM traffic.         */
int vm_flow_warp_mask_bgra(const uint8_t *fg, const float *backward, const float *forward,
                           int n, int h, int w, uint8_t *out_bgr, float *out_alpha,
                           int32_t *status, void *stream);

/* TPS warp of a BGRA frame + composite onto bg (tps.py:14-34, reader.py:72-79), i.e.
 * warp_image(.., identity affine, thin) for fg and alpha followed by create_composite_image.
 * The coarse-grid spline evaluation (tps.py:101-121) is fused in: ctrl (n,N,2) / coef (n,N+3,2)
 * float64 as for vm_tps_coarse, coarse point (k,l) = (k*step_x, l*step_y).
 * out (n,h,w,4) float32 = {B, G, R composite (0..255), warped alpha}.  23 B/px.
 * `scratch`: device workspace of vm_fused_scratch_bytes(n,h,w) bytes (see below; may be NULL
 * for this entry point unless the per-pixel gather variant is selected).                    */
int64_t vm_fused_scratch_bytes(int n, int h, int w);
int vm_tps_composite_bgra(const uint8_t *fg, const uint8_t *bg, int n_bg,
                          const double *ctrl, const double *coef, int N,
                          int nx, int ny, double step_x, double step_y,
                          const vm_axis_entry *rows, const vm_axis_entry *cols,
                          int n, int h, int w, float *out, void *scratch,
                          int32_t *status, void *stream);

/* flow warp + consistency mask + TPS + composite (SURVEY 8d "C4 pipeline"):
 * warp_bgr/warp_img (flow.py:9-33), correct_alpha (flow.py:36-65), warp_image(identity affine,
 * thin) (augmentation.py:44-63 -> tps.py:14-123), create_composite_image (reader.py:72-79).
 * 39 algorithmic B/px.  forward may be NULL (no consistency test).  The default variant (lean split
 * pipeline) runs four kernels per round of frames - flow stage (packed {bgr, TA} pixels, 8 B/px, into
 * `scratch`), float64 spline on the coarse grid, tile boxes, resampling + composite; fused_variant 5
 * runs the whole pipeline in one kernel and does not touch `scratch`.                         */
int vm_flow_tps_composite_bgra(const uint8_t *fg, const float *backward, const float *forward,
                               const uint8_t *bg, int n_bg,
                               const double *ctrl, const double *coef, int N,
                               int nx, int ny, double step_x, double step_y,
                               const vm_axis_entry *rows, const vm_axis_entry *cols,
                               int n, int h, int w, float *out, void *scratch,
                               int32_t *status, void *stream);

/* Tuning / test switches (process-wide configuration: set before use, not concurrently with calls; every
 * setting gives bit-identical output for variants 4 and 5):
 * "fused_variant" 4 = lean split pipeline (default: flow stage, float64 spline stage, TMA-tiled resampling
 * stage), 5 = single-pass warp-specialised kernel for the C4 entry point (no intermediate in HBM), 1 =
 * per-pixel gather kernels (generic fallback); "lean_chunk" frames per stage round (0 = automatic: about 64 frames of 1080p worth of pixels), "lean_sub", "lean_rb",
 * "lean_b1_warps", "lean_b1_dyr", "lean_b1_ctas", "lean_minb", "lean_fine_rows", "lean_box_cap", "lean_aug_minb",
 * "lean_floors", "flow_stage_layout", "lean_tmap" (1 = tensor-map staging of the resampling stage, the default when available; 0 = one bulk copy per row): schedule parameters of variant 4; "fuse_ctas": persistent CTAs of
 * variant 5; "lean_timing" 1 = record CUDA events around the stages of variant 4 (vm_lean_stage_ms).       */
int vm_set_option(const char *key, int value);

/* Batched augmentation (reference augmentation.py:102-135, `augment` for a whole clip; BASELINE config 5).
 * The host draws the reference's random parameters frame by frame and solves the TPS systems; the device
 * stages are:
 *   vm_alpha_stats_bgra   object_size / fg_center sums of every BGRA frame (augmentation.py:10-21):
 *                         out (n,3) uint64 {count(A != 0), sum(rows), sum(cols)}, zeroed by the caller;
 *   vm_tps_coarse_packed  spline on the coarse grid (tps.py:101-123) as (n,nx,ny) double2 {row, col}
 *                         (`counter`: one zeroed device word of workspace);
 *   vm_aug_tps            tps.warp_images on B,G,R (uint8, half-up) and alpha = A/255 of every frame on the
 *                         (h+1) x (w+1) grid (tps.py:14-75) -> (n,h+1,w+1) uint2 {B|G<<8|R<<16, float32 alpha bits};
 *   vm_aug_affine         the two cv2.warpAffine passes of augmentation.warp_image (augmentation.py:59-62:
 *                         integer translation, then rotation/scale, both (w,h)) fused with
 *                         change_illumination (augmentation.py:88-99).  mode 1: src = vm_aug_tps output ->
 *                         out_bgr (n,h,w,3) uint8 + out_alpha (n,h,w) float32; mode 0: src = (n,h,w,3) uint8
 *                         background -> out_bgr.  params: device array of n {double M[6]; int32 tu, tv}
 *                         (M = 2x3 matrix of the second pass); luts: device (n,256) uint8 S/V tables.
 *   float64 alpha (optional, both NULL otherwise): vm_aug_tps also writes alpha64 (n,h+1,w+1) in scipy's float64
 *                         operation order and vm_aug_affine (mode 1) reads it and writes out_alpha64 (n,h,w) float64
 *                         instead of out_alpha - what the reference's augment() returns (augmentation.py:125).      */
int vm_alpha_stats_bgra(const uint8_t *bgra, int n, int h, int w, unsigned long long *out, void *stream);
int vm_tps_coarse_packed(const double *ctrl, const double *coef, int n, int N, int nx, int ny,
                         double step_x, double step_y, void *T, unsigned int *counter, void *stream);
int vm_aug_tps(const uint8_t *fg_bgra, const void *T, int nx, int ny, const vm_axis_entry *rows,
               const vm_axis_entry *cols, int n, int h, int w, void *inter, double *alpha64, int32_t *status, void *stream);
int vm_aug_affine(int mode, const void *src, const double *alpha64, const void *params, const uint8_t *luts, int n, int h, int w,
                  uint8_t *out_bgr, float *out_alpha, double *out_alpha64, int hsv_vec, void *stream);

/* ---- batch loader: loader.load_and_crop / simple_load_crop / video_load_crop ----
 * reference: loader.py:39-85, 119-157, 285-330 (one sample), loader.py:93-116, 160-171, 333-345 (batch).
 * The host decodes the files, draws the reference's np.random.randint sequence (crop type, padding
 * offsets of get_padded_img loader.py:10-36, crop origins) and describes each sample by two views;
 * the device does everything after the decode for the whole batch in one launch: flow.warp_img of the
 * previous alpha (flow.py:9-18), padding, crop, cv2.resize(INTER_LINEAR) of the float64 planes
 * (loader.py:316-319), create_composite_image (reader.py:72-79), VGG-mean subtraction (loader.py:322-323)
 * and the optional mirror of get_batch (loader.py:107-110).
 *
 * A view maps the window that is resized back to image pixels: window (r, c) is canvas
 * (wi + r, wj + c); canvas cells inside [vi0,vi1) x [vj0,vj1) hold image pixel
 * (si + row - vi0, sj + col - vj0), all others hold 0 (the zero canvas of get_padded_img).
 * mode 0: separable linear interpolation with scale = 1 / (dst / src) per axis (double coefficients,
 * as OpenCV 4.13 does for CV_64F); mode 1: 2x2 block mean (OpenCV's INTER_AREA switch for exact 2x).  */
typedef struct {
    int32_t win_h, win_w;            /* size of the resized window                                */
    int32_t wi, wj;                  /* window origin in the canvas                               */
    int32_t vi0, vi1, vj0, vj1;      /* canvas rectangle that holds image data                    */
    int32_t si, sj;                  /* image pixel at canvas (vi0, vj0)                          */
    int32_t mode, reserved;
    double  scale_y, scale_x;
} vm_loader_view;

typedef struct {
    const uint8_t *fg;               /* (fh, fw, 4) uint8 BGRA: foreground B,G,R and A = 255 * alpha; may be a
                                        sub-rectangle of the decoded image whose origin is (oy, ox)       */
    const uint8_t *prev;             /* alpha byte of the previous frame's pixel (0, 0), pixels prev_stride
                                        bytes apart, rows pw pixels long, (ph, pw) pixels; or NULL        */
    const float   *flow;             /* (fh, fw, 2) float32 on the same rectangle as fg; required with prev */
    const uint8_t *bg;               /* (bh, bw, 3) uint8 BGR                                              */
    int32_t fh, fw, bh, bw;
    int32_t oy, ox;                  /* full-image coordinates of fg / flow element (0, 0)                 */
    int32_t ph, pw, prev_stride;
    int32_t flip;                    /* flip != 0: outputs mirrored along the columns                      */
    vm_loader_view fgv, bgv;
} vm_loader_sample;

/* samples: DEVICE array of n descriptors.  Outputs are (n, out_h, out_w, C) of out_dtype (VM_F64 as the
 * reference, or VM_F32): cmp C=3 (composite - mean), bg C=3 (resized background - mean), label C=1
 * (alpha), warped C=3 (resized flow-warped previous alpha, repeated), fg C=3 (resized foreground); any of
 * them may be NULL.  mean_host: 3 doubles on the HOST (params.VGG_MEAN).                                */
int vm_loader_batch(const vm_loader_sample *samples, int n, int out_h, int out_w, const double *mean_host,
                    int out_dtype, void *cmp, void *bg, void *label, void *warped, void *fg, void *stream);

/* loader.psnr (loader.py:214-227): adds sum((a - b)^2) over n elements into the device double *out
 * (zeroed by the caller); dtype VM_F32 or VM_F64.                                                       */
int vm_sq_err_sum(const void *a, const void *b, int dtype, int64_t n, double *out, void *stream);

/* ---- data.trimap_from_matte -------------------------------------------------------
 * reference: data.py:37-67, a pure-Python raster scan.  matte: (n, h, w) float64 in [0,1] (VM_F64, the
 * reference's argument) or the uint8 alpha bytes (VM_U8: 255 = 1., 0 = 0.); out (n, h, w) uint8 in
 * {0, 128, 255}.  Reproduces the scan-order dependence of the reference: an alpha==1 (alpha==0) pixel
 * becomes 128 iff a fractional pixel within Chebyshev distance 3 (1) follows it in raster order.      */
int vm_trimap_from_matte(const void *matte, int dtype, int n, int h, int w, uint8_t *out, void *stream);

/* ---- reader.read_fg_img, uint16 branch (clip ingest) ------------------------
 * reference: reader.py:13-15, (((img + 1) / 256.) - 1).astype(uint8) with the uint16 wrap of img + 1 and
 * the x86 wrap of -1.0 -> 255.  src: n uint16 elements as decoded (16-byte aligned), dst: n uint8.       */
int vm_fg_from_u16(const uint16_t *src, int64_t n, uint8_t *dst, void *stream);

/* Measurement hooks of the fused paths (no reference counterpart; used by bench.py).
 * vm_lean_stage_ms: durations in ms of {spline, tile boxes, flow stage, resampling+composite}
 * of the first chunk of the last call of the lean split pipeline (fused_variant 4, and C3) made with
 * "lean_timing" = 1 on the current device, after the stream was synchronised.
 * vm_lean_launch_count / vm_fuse_launch_count: kernels launched so far by the lean pipeline / by the
 * single-pass C4 kernel (fused_variant 5, the default).                                       */
int       vm_lean_stage_ms(float *out4);
long long vm_lean_launch_count(void);
long long vm_fuse_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VM_B200_H */
