// Thin-plate-spline kernels: coarse-grid radial-basis evaluation (float64), bilinear
// up-sampling of the transform, scipy-style resampling, and the fused TPS(+flow)+composite
// kernels.  Reference behaviour: tps.py:14-123, augmentation.py:44-63, reader.py:72-79.
#include "vm_common.cuh"
#include <math.h>
#include <string.h>
#include <mutex>

int vm_lean_set_option(const char *key, int value);                                // vm_lean.cu
int vm_fuse_set_option(const char *key, int value);                                // vm_fuse.cu
extern int g_vm_flow_stage_layout;                                                 // vm_flow.cu

// ---------------------------------------------------------------------------------------
// float64 log for the radial basis U(r) = r^2 log r = 0.5 * r2 * log(r2).
//
// x = 2^e * m, m in [1,2).  The top VM_LOG_BITS mantissa bits pick c_k = 1 + (k + 0.5)/2^B;
// table holds {1/c_k, log(c_k)} (from long double), r = fma(m, 1/c_k, -1) has |r| <= 2^-(B+1)
// and log(m) = log(c_k) + log1p(r) with a degree-5 Taylor polynomial (|r|^6/6 < 3e-21).
// Absolute error ~1 ulp of the result, i.e. the same class as numpy's log; the TPS transform
// then agrees with the reference to ~1e-11 px (tests/test_gpu_parity.py asserts 5e-10).
// ---------------------------------------------------------------------------------------
#define VM_LOG_BITS 10
#define VM_LOG_N (1 << VM_LOG_BITS)

__device__ double2 g_vm_log_tab[VM_LOG_N];

static std::mutex g_init_mu;
static bool g_init_done[64];

extern "C" int vm_init(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        vm_set_error("vm_init: cudaGetDevice failed");
        return VM_ERR_CUDA;
    }
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (g_init_done[dev]) return VM_OK;
    static double2 tab[VM_LOG_N];
    for (int k = 0; k < VM_LOG_N; ++k) {
        const long double c = 1.0L + ((long double)k + 0.5L) / (long double)VM_LOG_N;
        const double inv = (double)(1.0L / c);
        tab[k].x = inv;
        tab[k].y = (double)(-logl((long double)inv));   // log(1/inv): consistent with the rounded 1/c
    }
    cudaError_t e = cudaMemcpyToSymbol(g_vm_log_tab, tab, sizeof(tab));
    if (e != cudaSuccess) {
        vm_set_error("vm_init: cudaMemcpyToSymbol: %s", cudaGetErrorString(e));
        return VM_ERR_CUDA;
    }
    g_init_done[dev] = true;
    return VM_OK;
}

// log(x) for finite x >= 2^-1000 (callers special-case tiny / zero r2).
__device__ __forceinline__ double vm_log_pos(double x, const double2 *__restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const int k = (hi >> (20 - VM_LOG_BITS)) & (VM_LOG_N - 1);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double2 t = tab[k];
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, 0.2, -0.25);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    const double r2 = r * r;
    const double l1p = fma(r2, q, r);                    // log1p(r)
    return fma((double)e, 0.6931471805599453094, t.y + l1p);
}

// a1 + ax*x + ay*y + sum_i w_i U(|(x,y) - P_i|) for both output coordinates.
// sp: N * {Px, Py, w0, w1}; aff: {a1_0, ax_0, ay_0, a1_1, ax_1, ay_1}
__device__ __forceinline__ void vm_tps_point(const double4 *__restrict__ sp, const double *__restrict__ aff,
                                             int N, double x, double y, const double2 *__restrict__ tab,
                                             double &o0, double &o1) {
    double s0 = 0.0, s1 = 0.0;
    for (int a = 0; a < N; ++a) {
        const double4 p = sp[a];
        const double dx = x - p.x, dy = y - p.y;
        const double r2 = fma(dx, dx, dy * dy);
        double U = 0.0;
        if (r2 >= 1e-200) U = r2 * (0.5 * vm_log_pos(r2, tab));
        s0 = fma(p.z, U, s0);
        s1 = fma(p.w, U, s1);
    }
    o0 = ((aff[0] + aff[1] * x) + aff[2] * y) + s0;
    o1 = ((aff[3] + aff[4] * x) + aff[5] * y) + s1;
}

#define VM_TPS_MAX_N 256

__global__ void __launch_bounds__(256)
k_tps_coarse(const double *__restrict__ ctrl, const double *__restrict__ coef, int N, int nx, int ny,
             double step_x, double step_y, double x0, double y0, double *__restrict__ coarse) {
    __shared__ double4 sp[VM_TPS_MAX_N];
    __shared__ double aff[6];
    const int frame = blockIdx.z;
    const double *P = ctrl + (int64_t)frame * N * 2;
    const double *C = coef + (int64_t)frame * (N + 3) * 2;
    for (int a = threadIdx.y * blockDim.x + threadIdx.x; a < N; a += blockDim.x * blockDim.y)
        sp[a] = make_double4(P[2 * a], P[2 * a + 1], C[2 * a], C[2 * a + 1]);
    if (threadIdx.y == 0 && threadIdx.x < 6) {
        const int c = threadIdx.x / 3, r = threadIdx.x % 3;
        aff[threadIdx.x] = C[(N + r) * 2 + c];
    }
    __syncthreads();
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y * blockDim.y + threadIdx.y;
    if (k >= nx || l >= ny) return;
    double o0, o1;
    vm_tps_point(sp, aff, N, (double)k * step_x + x0, (double)l * step_y + y0, g_vm_log_tab, o0, o1);
    double *out = coarse + (int64_t)frame * 2 * nx * ny;
    out[(int64_t)k * ny + l] = o0;
    out[(int64_t)nx * ny + (int64_t)k * ny + l] = o1;
}

extern "C" int vm_tps_coarse(const double *ctrl, const double *coef, int n, int N, int nx, int ny,
                             double step_x, double step_y, double x0, double y0, double *coarse,
                             void *stream) {
    VM_REQUIRE(ctrl && coef && coarse, "null pointer");
    VM_REQUIRE(n >= 0 && N >= 1 && N <= VM_TPS_MAX_N, "control point count out of range");
    VM_REQUIRE(nx >= 1 && ny >= 1 && n < 65536, "bad size");
    if (n == 0) return VM_OK;
    int rc = vm_init();
    if (rc != VM_OK) return rc;
    dim3 block(32, 8), grid((ny + 31) / 32, (nx + 7) / 8, n);
    k_tps_coarse<<<grid, block, 0, (cudaStream_t)stream>>>(ctrl, coef, N, nx, ny, step_x, step_y, x0, y0, coarse);
    return vm_check_launch("vm_tps_coarse");
}

// ---------------------------------------------------------------------------------------
// up-sampled transform at fine position (i, j) of one frame (exact reference order)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void vm_tps_fine(const double *__restrict__ coarse, int nx, int ny,
                                            const vm_axis_entry &re, const vm_axis_entry &ce,
                                            double &t0, double &t1) {
    const double *T0 = coarse, *T1 = coarse + (int64_t)nx * ny;
    const int64_t o00 = (int64_t)re.i0 * ny + ce.i0, o01 = (int64_t)re.i0 * ny + ce.i1;
    const int64_t o10 = (int64_t)re.i1 * ny + ce.i0, o11 = (int64_t)re.i1 * ny + ce.i1;
    t0 = vm_upsample_exact(__ldg(T0 + o00), __ldg(T0 + o01), __ldg(T0 + o10), __ldg(T0 + o11), re.frac, ce.frac);
    t1 = vm_upsample_exact(__ldg(T1 + o00), __ldg(T1 + o01), __ldg(T1 + o10), __ldg(T1 + o11), re.frac, ce.frac);
}

__global__ void __launch_bounds__(256)
k_tps_upsample(const double *__restrict__ coarse, int nx, int ny, const vm_axis_entry *__restrict__ rows,
               const vm_axis_entry *__restrict__ cols, int h, int w, double *__restrict__ out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j > w || i > h) return;
    double t0, t1;
    vm_tps_fine(coarse, nx, ny, vm_ld_axis(rows + i), vm_ld_axis(cols + j), t0, t1);
    const int64_t plane = (int64_t)(h + 1) * (w + 1), o = (int64_t)i * (w + 1) + j;
    out[o] = t0;
    out[plane + o] = t1;
}

extern "C" int vm_tps_upsample(const double *coarse, int nx, int ny, const vm_axis_entry *rows,
                               const vm_axis_entry *cols, int h, int w, double *out, void *stream) {
    VM_REQUIRE(coarse && rows && cols && out, "null pointer");
    VM_REQUIRE(nx >= 1 && ny >= 1 && h >= 1 && w >= 1 && h < 65535, "bad size");
    dim3 grid((w + 1 + 255) / 256, h + 1);
    k_tps_upsample<<<grid, 256, 0, (cudaStream_t)stream>>>(coarse, nx, ny, rows, cols, h, w, out);
    return vm_check_launch("vm_tps_upsample");
}

// ---------------------------------------------------------------------------------------
// generic tps.warp_images: up-sample + map_coordinates(order=1) for uint8 / float64 images
// ---------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256)
k_tps_warp(const T *__restrict__ src, int sh, int sw, const double *__restrict__ coarse, int nx, int ny,
           const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols,
           int oh, int ow, T *__restrict__ dst, int32_t *__restrict__ status, int order) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ow || i >= oh) return;
    double t0, t1;
    vm_tps_fine(coarse, nx, ny, vm_ld_axis(rows + i), vm_ld_axis(cols + j), t0, t1);
    const VmBilin64 s = vm_mapcoord_setup(t0, t1, sh, sw);
    T *o = dst + ((int64_t)i * ow + j) * C;
    if (!s.inside) {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = T(0);
        if (status) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, 1);
        return;
    }
    if (order == 0) {                       // scipy order 0: the sample at floor(t + 1/2), no arithmetic on the value
        const T *pn = src + ((int64_t)vm_nearest(t0) * sw + vm_nearest(t1)) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = __ldg(pn + c);
        return;
    }
    const T *p00 = src + ((int64_t)s.i0 * sw + s.j0) * C, *p01 = src + ((int64_t)s.i0 * sw + s.j1) * C;
    const T *p10 = src + ((int64_t)s.i1 * sw + s.j0) * C, *p11 = src + ((int64_t)s.i1 * sw + s.j1) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double v = vm_mapcoord_blend(s, (double)__ldg(p00 + c), (double)__ldg(p01 + c),
                                           (double)__ldg(p10 + c), (double)__ldg(p11 + c));
        if (sizeof(T) == 1) o[c] = (T)vm_round_half_up_u8(v);
        else o[c] = (T)v;
    }
}

extern "C" int vm_tps_warp(const void *src, int dtype, int channels, int sh, int sw,
                           const double *coarse, int nx, int ny, const vm_axis_entry *rows,
                           const vm_axis_entry *cols, int oh, int ow, void *dst, int32_t *status,
                           void *stream) {
    return vm_tps_warp_order(src, dtype, channels, sh, sw, coarse, nx, ny, rows, cols, oh, ow, dst, status, 1, stream);
}

extern "C" int vm_tps_warp_order(const void *src, int dtype, int channels, int sh, int sw,
                                 const double *coarse, int nx, int ny, const vm_axis_entry *rows,
                                 const vm_axis_entry *cols, int oh, int ow, void *dst, int32_t *status,
                                 int order, void *stream) {
    VM_REQUIRE(src && coarse && rows && cols && dst, "null pointer");
    VM_REQUIRE(order == 0 || order == 1, "interpolation order must be 0 or 1");
    VM_REQUIRE(sh >= 1 && sw >= 1 && oh >= 1 && ow >= 1 && oh < 65536, "bad size");
    dim3 grid((ow + 255) / 256, oh);
    cudaStream_t st = (cudaStream_t)stream;
#define VM_TW(T, C) k_tps_warp<T, C><<<grid, 256, 0, st>>>((const T *)src, sh, sw, coarse, nx, ny, rows, cols, oh, ow, (T *)dst, status, order)
    if (dtype == VM_U8 && channels == 1) VM_TW(uint8_t, 1);
    else if (dtype == VM_U8 && channels == 3) VM_TW(uint8_t, 3);
    else if (dtype == VM_U8 && channels == 4) VM_TW(uint8_t, 4);
    else if (dtype == VM_F64 && channels == 1) VM_TW(double, 1);
    else if (dtype == VM_F64 && channels == 3) VM_TW(double, 3);
    else if (dtype == VM_F32 && channels == 1) VM_TW(float, 1);
    else { vm_set_error("vm_tps_warp: unsupported dtype/channels %d/%d", dtype, channels); return VM_ERR_ARG; }
#undef VM_TW
    return vm_check_launch("vm_tps_warp");
}

template <typename T, int C>
__global__ void __launch_bounds__(256)
k_map_coordinates(const T *__restrict__ src, int sh, int sw, const double *__restrict__ t0p,
                  const double *__restrict__ t1p, int oh, int ow, T *__restrict__ dst,
                  int32_t *__restrict__ status, int order) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= ow || i >= oh) return;
    const int64_t q = (int64_t)i * ow + j;
    const VmBilin64 s = vm_mapcoord_setup(t0p[q], t1p[q], sh, sw);
    T *o = dst + q * C;
    if (!s.inside) {
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = T(0);
        if (status) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, 1);
        return;
    }
    if (order == 0) {
        const T *pn = src + ((int64_t)vm_nearest(t0p[q]) * sw + vm_nearest(t1p[q])) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = __ldg(pn + c);
        return;
    }
    const T *p00 = src + ((int64_t)s.i0 * sw + s.j0) * C, *p01 = src + ((int64_t)s.i0 * sw + s.j1) * C;
    const T *p10 = src + ((int64_t)s.i1 * sw + s.j0) * C, *p11 = src + ((int64_t)s.i1 * sw + s.j1) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double v = vm_mapcoord_blend(s, (double)__ldg(p00 + c), (double)__ldg(p01 + c),
                                           (double)__ldg(p10 + c), (double)__ldg(p11 + c));
        if (sizeof(T) == 1) o[c] = (T)vm_round_half_up_u8(v);
        else o[c] = (T)v;
    }
}

extern "C" int vm_map_coordinates(const void *src, int dtype, int channels, int sh, int sw,
                                  const double *t0, const double *t1, int oh, int ow, void *dst,
                                  int32_t *status, void *stream) {
    return vm_map_coordinates_order(src, dtype, channels, sh, sw, t0, t1, oh, ow, dst, status, 1, stream);
}

extern "C" int vm_map_coordinates_order(const void *src, int dtype, int channels, int sh, int sw,
                                        const double *t0, const double *t1, int oh, int ow, void *dst,
                                        int32_t *status, int order, void *stream) {
    VM_REQUIRE(src && t0 && t1 && dst, "null pointer");
    VM_REQUIRE(order == 0 || order == 1, "interpolation order must be 0 or 1");
    VM_REQUIRE(sh >= 1 && sw >= 1 && oh >= 1 && ow >= 1 && oh < 65536, "bad size");
    dim3 grid((ow + 255) / 256, oh);
    cudaStream_t st = (cudaStream_t)stream;
#define VM_MC(T, C) k_map_coordinates<T, C><<<grid, 256, 0, st>>>((const T *)src, sh, sw, t0, t1, oh, ow, (T *)dst, status, order)
    if (dtype == VM_U8 && channels == 1) VM_MC(uint8_t, 1);
    else if (dtype == VM_U8 && channels == 3) VM_MC(uint8_t, 3);
    else if (dtype == VM_F64 && channels == 1) VM_MC(double, 1);
    else if (dtype == VM_F32 && channels == 1) VM_MC(float, 1);
    else { vm_set_error("vm_map_coordinates: unsupported dtype/channels %d/%d", dtype, channels); return VM_ERR_ARG; }
#undef VM_MC
    return vm_check_launch("vm_map_coordinates");
}

// ---------------------------------------------------------------------------------------
// fused TPS (+ flow warp + mask) + composite on BGRA frames, one output pixel per thread.
//
// Source of the TPS resampling is either the BGRA frame itself (C3) or the flow-warped,
// consistency-masked frame evaluated on the fly at the 4 integer neighbours (C4): each
// neighbour is {B,G,R uint8 bit-exact, alpha = TA/261120}.  Output float4 {B,G,R,alpha'}.
// ---------------------------------------------------------------------------------------
template <bool FLOW>
__global__ void __launch_bounds__(256)
k_tps_composite(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd, const float2 *__restrict__ fwd,
                const uint8_t *__restrict__ bg, int n_bg, const double *__restrict__ coarse, int nx, int ny,
                const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols,
                int h, int w, float4 *__restrict__ out, int32_t *__restrict__ status) {
    const int j = blockIdx.x * 32 + threadIdx.x, i = blockIdx.y * 8 + threadIdx.y, frame = blockIdx.z;
    if (j >= w || i >= h) return;
    const int64_t fbase = (int64_t)frame * h * w;
    const uint8_t *fgf = fg + fbase * 4;
    const float2 *bf = FLOW ? bwd + fbase : nullptr;
    const float2 *ff = (FLOW && fwd) ? fwd + fbase : nullptr;
    double t0, t1;
    vm_tps_fine(coarse + (int64_t)frame * 2 * nx * ny, nx, ny, vm_ld_axis(rows + i), vm_ld_axis(cols + j), t0, t1);
    const VmBilin64 s = vm_mapcoord_setup(t0, t1, h, w);
    const int64_t p = (int64_t)i * w + j;
    const uint8_t *bgp = bg + ((int64_t)(frame % n_bg) * h * w + p) * 3;
    const double bb = (double)__ldg(bgp), bgc = (double)__ldg(bgp + 1), br = (double)__ldg(bgp + 2);
    double fb_ = 0.0, fg_ = 0.0, fr_ = 0.0, a2 = 0.0;
    int flags = 0;
    if (s.inside) {
        const VmSrcPx s00 = vm_src_px<FLOW>(fgf, bf, ff, h, w, s.i0, s.j0, flags);
        const VmSrcPx s01 = vm_src_px<FLOW>(fgf, bf, ff, h, w, s.i0, s.j1, flags);
        const VmSrcPx s10 = vm_src_px<FLOW>(fgf, bf, ff, h, w, s.i1, s.j0, flags);
        const VmSrcPx s11 = vm_src_px<FLOW>(fgf, bf, ff, h, w, s.i1, s.j1, flags);
        fb_ = (double)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.b, s01.b, s10.b, s11.b));
        fg_ = (double)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.g, s01.g, s10.g, s11.g));
        fr_ = (double)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.r, s01.r, s10.r, s11.r));
        a2 = vm_mapcoord_blend(s, s00.a, s01.a, s10.a, s11.a);
    } else if (status) {
        atomicAdd(status + VM_STATUS_TPS_OUTSIDE, 1);
    }
    const double na = 1.0 - a2;
    float4 o;
    o.x = (float)(a2 * fb_ + na * bb);
    o.y = (float)(a2 * fg_ + na * bgc);
    o.z = (float)(a2 * fr_ + na * br);
    o.w = (float)a2;
    out[fbase + p] = o;
    if (flags && status) {
        if (flags & 1) atomicAdd(status + VM_STATUS_INDEX_ERR, 1);
        if (flags & 2) atomicAdd(status + VM_STATUS_NAN_ERR, 1);
    }
}

// option block (vm_set_option): fused_variant 4 = lean split pipeline (default, vm_lean.cu), 5 = single-pass kernel
// for C4 (vm_fuse.cu), 1 = per-pixel gather kernels (the generic fallback for anything the others do not take).
static int g_opt_variant = 4;
#define VM_FUSED_MAX_N 64                    /* control points the fused fast paths take (shared-memory tables) */

extern "C" int vm_set_option(const char *key, int value) {
    if (!key) return VM_ERR_ARG;
    if (!strcmp(key, "fused_variant") && (value == 1 || value == 4 || value == 5)) { g_opt_variant = value; return VM_OK; }
    if (!strncmp(key, "lean_", 5) && vm_lean_set_option(key, value) == VM_OK) return VM_OK;
    if (!strncmp(key, "fuse_", 5) && vm_fuse_set_option(key, value) == VM_OK) return VM_OK;
    if (!strcmp(key, "flow_stage_layout") && (value == 0 || value == 1)) { g_vm_flow_stage_layout = value; return VM_OK; }
    vm_set_error("vm_set_option: unknown option %s=%d", key, value);
    return VM_ERR_ARG;
}

int64_t vm_lean_scratch_bytes(int n, int h, int w);                                // vm_lean.cu
int vm_lean_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, void *scratch, int32_t *status, cudaStream_t st, const char *what);
int vm_fuse_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, int32_t *status, cudaStream_t st, const char *what);  // vm_fuse.cu

static int launch_fused(bool flow, const uint8_t *fg, const float *backward, const float *forward,
                        const uint8_t *bg, int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny,
                        double step_x, double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n,
                        int h, int w, float *out, void *scratch, int32_t *status, void *stream, const char *what) {
    if (n == 0) return VM_OK;
    int rc = vm_init();
    if (rc != VM_OK) return rc;
    VM_REQUIRE(h <= 32767 && w <= 32767 && (int64_t)h * w < (1ll << 28), "frame too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool small_n = N <= VM_FUSED_MAX_N && h >= 2 && w >= 2;
    // variant 5: C4 in ONE kernel (vm_fuse.cu: every byte touches HBM once; measured slower than the lean split pipeline,
    // DESIGN.md 5f); C3 and anything it does not take goes to the lean pipeline
    if (g_opt_variant == 5 && flow && small_n && (int64_t)n * ((h + 59) / 60) * ((w + 59) / 60) < (1ll << 31))
        return vm_fuse_launch(forward ? 2 : 1, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y,
                              rows, cols, n, h, w, out, status, st, what);
    if ((g_opt_variant == 4 || g_opt_variant == 5) && small_n && nx <= h / 2 + 1 && ny <= w / 2 + 1)
        return vm_lean_launch(flow ? (forward ? 2 : 1) : 0, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny,
                              step_x, step_y, rows, cols, n, h, w, out, scratch, status, st, what);
    // generic fallback (any control-point count, any coarse grid): coarse transform through global scratch, then one
    // pixel per thread with every source pixel evaluated from the inputs
    VM_REQUIRE(scratch, "scratch workspace (vm_fused_scratch_bytes) required");
    double *coarse = (double *)scratch;
    rc = vm_tps_coarse(ctrl, coef, n, N, nx, ny, step_x, step_y, 0.0, 0.0, coarse, stream);
    if (rc != VM_OK) return rc;
    dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, n);
    if (flow)
        k_tps_composite<true><<<grid, block, 0, st>>>(fg, (const float2 *)backward, (const float2 *)forward, bg,
                                                      n_bg, coarse, nx, ny, rows, cols, h, w, (float4 *)out, status);
    else
        k_tps_composite<false><<<grid, block, 0, st>>>(fg, nullptr, nullptr, bg, n_bg, coarse, nx, ny, rows,
                                                       cols, h, w, (float4 *)out, status);
    return vm_check_launch(what);
}

extern "C" int64_t vm_fused_scratch_bytes(int n, int h, int w) {
    // enough for every variant selectable with vm_set_option for this clip shape, any coarse grid up to (h/2+1) x (w/2+1)
    const int64_t gather = (int64_t)n * 2 * (h / 2 + 1) * (w / 2 + 1) * (int64_t)sizeof(double);
    const int64_t lean = vm_lean_scratch_bytes(n, h, w);
    return gather > lean ? gather : lean;
}

extern "C" int vm_tps_composite_bgra(const uint8_t *fg, const uint8_t *bg, int n_bg, const double *ctrl,
                                     const double *coef, int N, int nx, int ny, double step_x, double step_y,
                                     const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                                     float *out, void *scratch, int32_t *status, void *stream) {
    if (n == 0) return VM_OK;                                          // empty clip: nothing to read, nothing to launch
    VM_REQUIRE(fg && bg && ctrl && coef && rows && cols && out, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && h > 0 && w > 0 && n_bg >= 1 && nx >= 1 && ny >= 1 && N >= 1 && N <= VM_TPS_MAX_N, "bad size");
    return launch_fused(false, fg, nullptr, nullptr, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols,
                        n, h, w, out, scratch, status, stream, "vm_tps_composite_bgra");
}

extern "C" int vm_flow_tps_composite_bgra(const uint8_t *fg, const float *backward, const float *forward,
                                          const uint8_t *bg, int n_bg, const double *ctrl, const double *coef,
                                          int N, int nx, int ny, double step_x, double step_y,
                                          const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h,
                                          int w, float *out, void *scratch, int32_t *status, void *stream) {
    if (n == 0) return VM_OK;
    VM_REQUIRE(fg && backward && bg && ctrl && coef && rows && cols && out, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && h > 0 && w > 0 && n_bg >= 1 && nx >= 1 && ny >= 1 && N >= 1 && N <= VM_TPS_MAX_N, "bad size");
    return launch_fused(true, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols,
                        n, h, w, out, scratch, status, stream, "vm_flow_tps_composite_bgra");
}
