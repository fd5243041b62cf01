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

// ---------------------------------------------------------------------------------------
// Tiled fused kernel (the fast path): one CTA per TT_H x TT_W output tile.
//
//   P0  control points / affine part / log table / axis entries of the tile -> shared memory
//   P1  radial-basis sum on the tile's coarse sub-grid (float64, table log) -> shared memory;
//       bounding box of the coarse values = bounding box of every fine coordinate of the tile
//       (the fine transform is a convex combination of coarse values)
//   P3  the source pixels of that box are produced ONCE into shared memory: for C4 the
//       flow-warped, consistency-masked pixel {B,G,R uint8 exact, alpha code}; for C3 the
//       BGRA pixel itself.  Flow, forward flow and BGRA taps come through L1/L2.
//   P4  per output pixel: float64 up-sampling of the transform, map_coordinates geometry in
//       float64, colour blend in float32 with an exact float64 re-evaluation whenever the
//       float32 value is within 5e-4 of a rounding boundary, alpha and (1 - alpha) blended
//       separately (relative accuracy for the composite), one 16-byte store.
//
// HBM traffic is the algorithmic 39 (C4) / 23 (C3) bytes per pixel plus tile halos served by
// L2.  A tile whose source box does not fit the shared-memory budget (degenerate grids) takes
// the per-pixel gather path of k_tps_composite and is counted in VM_STATUS_SLOW_TILES.
// ---------------------------------------------------------------------------------------
#define TT_W 64
#define TT_THREADS 256
#define TT_MAX_N 64

template <int TH> struct TileCfg {
    static constexpr int CR = TH / 2 + 3;                  // max coarse rows of a tile
    static constexpr int CC = TT_W / 2 + 3;                // max coarse cols of a tile (35)
    static constexpr int IMAX = (TH == 64) ? 8448 : 4608;  // source-box entries (8 B each)
    static constexpr int SEG = (TH == 64) ? 5 : 3;         // coarse points per thread (one row run)
    static constexpr int NSEG = (CC + SEG - 1) / SEG;      // threads per coarse row
    static_assert(CR * NSEG <= TT_THREADS, "coarse tile does not fit one pass");
};

template <int TH> struct __align__(16) CoarseSmem {
    double2 logtab[VM_LOG_N];                              // P1 only; reused as the bg tile afterwards
    double4 ctrl[TT_MAX_N];                                // {Px, Py, w0/2, w1/2}
    vm_axis_entry rows[TH];
    vm_axis_entry cols[TT_W];
    double2 T[TileCfg<TH>::CR * TileCfg<TH>::CC];          // {row coord, col coord} per coarse point
    double aff[6];
    int box[4];                                            // rmin, rmax, cmin, cmax (floors)
    int bad;
    int pad;
};

template <int TH> struct __align__(16) TileSmem : CoarseSmem<TH> {
    uint2 inter[TileCfg<TH>::IMAX];                        // {B | G<<8 | R<<16, alpha code}
};

// log(x), x > 0 finite (x = 0 gives a finite value, so that 0 * log(0) = 0 as in tps.py:81)
__device__ __forceinline__ double vm_log_tab_smem(double x, const double2 *__restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double2 t = *reinterpret_cast<const double2 *>(
        reinterpret_cast<const char *>(tab) + ((hi >> (16 - VM_LOG_BITS)) & ((VM_LOG_N - 1) << 4)));
    // exponent as a double without a conversion instruction: 2^52 + biased_e - (2^52 + 1023)
    const double ed = __hiloint2double(0x43300000, (int)((unsigned)hi >> 20)) - 4503599627371519.0;
    const double r = fma(m, t.x, -1.0);                    // |r| <= 2^-11
    double q = fma(r, -0.25, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    q = fma(r, q, 1.0);                                    // log1p(r)/r to 6e-18
    return fma(ed, 0.6931471805599453094, fma(r, q, t.y));
}

// P0 + P1 of a tile: stage control points / log table / axis entries, evaluate the spline on the
// tile's coarse sub-grid into S.T and reduce the bounding box of the coarse values.  Returns
// false (after counting it) when the axis tables are not those of a /2 grid.
template <int TH>
__device__ __forceinline__ bool vm_tile_coarse(CoarseSmem<TH> &S, const double *__restrict__ ctrl,
                                               const double *__restrict__ coef, int N, double step_x, double step_y,
                                               const vm_axis_entry *__restrict__ rows,
                                               const vm_axis_entry *__restrict__ cols, int frame, int I0, int J0,
                                               int th, int tw, int32_t *__restrict__ status, int &kr0, int &kc0,
                                               int &nkr_out, int &nkc_out) {
    using Cfg = TileCfg<TH>;
    const int tid = threadIdx.x;
    // ---- P0 ----------------------------------------------------------------------------
    {
        const double *P = ctrl + (int64_t)frame * N * 2;
        const double *C = coef + (int64_t)frame * (N + 3) * 2;
        for (int k = tid; k < VM_LOG_N; k += TT_THREADS) S.logtab[k] = g_vm_log_tab[k];
        if (tid < N) S.ctrl[tid] = make_double4(P[2 * tid], P[2 * tid + 1], 0.5 * C[2 * tid], 0.5 * C[2 * tid + 1]);
        if (tid < 6) S.aff[tid] = C[(N + tid % 3) * 2 + tid / 3];
        if (tid < th) S.rows[tid] = vm_ld_axis(rows + I0 + tid);
        if (tid >= 64 && tid < 64 + tw) S.cols[tid - 64] = vm_ld_axis(cols + J0 + tid - 64);
        if (tid == 0) { S.box[0] = INT_MAX; S.box[1] = INT_MIN; S.box[2] = INT_MAX; S.box[3] = INT_MIN; S.bad = 0; }
    }
    __syncthreads();
    kr0 = S.rows[0].i0; kc0 = S.cols[0].i0;
    const int nkr = S.rows[th - 1].i1 - kr0 + 1, nkc = S.cols[tw - 1].i1 - kc0 + 1;
    nkr_out = nkr; nkc_out = nkc;
    if (nkr < 1 || nkc < 1 || nkr > Cfg::CR || nkc > Cfg::CC) {      // axis tables are not those of a /2 grid
        if (status && tid == 0) atomicAdd(status + VM_STATUS_BAD_TABLE, 1);
        return false;
    }

    // ---- P1: coarse radial-basis evaluation: thread = run of SEG points in one coarse row ---
    {
        const int k = tid / Cfg::NSEG, l0 = (tid - k * Cfg::NSEG) * Cfg::SEG;
        const bool active = k < nkr && l0 < nkc;
        int rmin = INT_MAX, rmax = INT_MIN, cmin = INT_MAX, cmax = INT_MIN, bad = 0;
        if (active) {
            const double x = (double)(kr0 + k) * step_x;
            double py[Cfg::SEG], s0[Cfg::SEG], s1[Cfg::SEG];
#pragma unroll
            for (int m = 0; m < Cfg::SEG; ++m) {
                py[m] = (double)(kc0 + min(l0 + m, nkc - 1)) * step_y;
                s0[m] = 0.0; s1[m] = 0.0;
            }
            for (int a = 0; a < N; ++a) {
                const double4 c = S.ctrl[a];
                const double dx = x - c.x;
                const double dx2 = dx * dx;
#pragma unroll
                for (int m = 0; m < Cfg::SEG; ++m) {
                    const double dy = py[m] - c.y;
                    const double r2 = fma(dy, dy, dx2);
                    const double U = r2 * vm_log_tab_smem(r2, S.logtab);
                    s0[m] = fma(c.z, U, s0[m]);
                    s1[m] = fma(c.w, U, s1[m]);
                }
            }
#pragma unroll
            for (int m = 0; m < Cfg::SEG; ++m) {
                if (l0 + m < nkc) {
                    const double v0 = ((S.aff[0] + S.aff[1] * x) + S.aff[2] * py[m]) + s0[m];
                    const double v1 = ((S.aff[3] + S.aff[4] * x) + S.aff[5] * py[m]) + s1[m];
                    S.T[k * nkc + l0 + m] = make_double2(v0, v1);
                    if (!(fabs(v0) < 1.0e9) || !(fabs(v1) < 1.0e9)) bad = 1;
                    else {
                        const int f0 = __double2int_rd(v0), f1 = __double2int_rd(v1);
                        rmin = min(rmin, f0); rmax = max(rmax, f0); cmin = min(cmin, f1); cmax = max(cmax, f1);
                    }
                }
            }
        }
        rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
        cmin = __reduce_min_sync(0xffffffffu, cmin); cmax = __reduce_max_sync(0xffffffffu, cmax);
        bad = __reduce_max_sync(0xffffffffu, bad);
        if ((tid & 31) == 0) {
            atomicMin(&S.box[0], rmin); atomicMax(&S.box[1], rmax);
            atomicMin(&S.box[2], cmin); atomicMax(&S.box[3], cmax);
            if (bad) atomicOr(&S.bad, 1);
        }
    }
    __syncthreads();
    return true;
}

template <bool FLOW, int TH>
__global__ void __launch_bounds__(TT_THREADS, (TH == 64) ? 2 : 3)
k_tps_tiled(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd, const float2 *__restrict__ fwd,
            const uint8_t *__restrict__ bg, int n_bg, const double *__restrict__ ctrl,
            const double *__restrict__ coef, int N, int nx, int ny, double step_x, double step_y,
            const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols,
            int h, int w, int tiles_x, int tiles_y, float4 *__restrict__ out, int32_t *__restrict__ status) {
    using Cfg = TileCfg<TH>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem<TH> &S = *reinterpret_cast<TileSmem<TH> *>(smem_raw);
    const int tid = threadIdx.x;
    const int per = tiles_x * tiles_y;
    const int frame = blockIdx.x / per;
    const int tl = blockIdx.x - frame * per;
    const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
    const int I0 = ty * TH, J0 = tx * TT_W;
    const int th = min(TH, h - I0), tw = min(TT_W, w - J0);
    const int64_t fbase = (int64_t)frame * h * w;
    const uint32_t *fg32 = reinterpret_cast<const uint32_t *>(fg) + fbase;
    const float2 *bf = FLOW ? bwd + fbase : nullptr;
    const float2 *ff = (FLOW && fwd) ? fwd + fbase : nullptr;

    int kr0, kc0, nkr, nkc;
    if (!vm_tile_coarse<TH>(S, ctrl, coef, N, step_x, step_y, rows, cols, frame, I0, J0, th, tw, status, kr0, kc0, nkr, nkc))
        return;
    // source box actually addressed by in-range pixels: rows [rmin, rmax+1], cols [cmin, cmax+1]
    const int rmin = max(S.box[0], 0), rmax = min(S.box[1] + 1, h - 1);
    const int cmin = max(S.box[2], 0), cmax = min(S.box[3] + 1, w - 1);
    const int RH = max(rmax - rmin + 1, 0), RW = max(cmax - cmin + 1, 0);
    const bool tiled = !S.bad && RH >= 2 && RW >= 2 && RH <= 4096 && RW <= 4096 && RH * RW <= Cfg::IMAX;
    int flags = 0;

    // ---- P3: source box -> shared memory (warp per row, 3 x 32 columns per step; the flow
    // vectors of the next step are prefetched while the current one is blended).  The background
    // tile is fetched asynchronously into the (now dead) log-table space meanwhile.
    uint8_t *bgt = reinterpret_cast<uint8_t *>(S.logtab);           // [TH][TT_W * 3]
    const uint8_t *bgf = bg + (int64_t)(frame % n_bg) * h * w * 3;
    const bool bg_async = (w & 15) == 0 && tw == TT_W && (reinterpret_cast<uintptr_t>(bg) & 15) == 0;
    if (bg_async) {
        for (int c = tid; c < th * (TT_W * 3 / 16); c += TT_THREADS) {
            const int r = c / (TT_W * 3 / 16), k = c - r * (TT_W * 3 / 16);
            const uint8_t *src = bgf + ((int64_t)(I0 + r) * w + J0) * 3 + k * 16;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(bgt + r * (TT_W * 3) + k * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    } else {
        for (int c = tid; c < th * tw * 3; c += TT_THREADS) {
            const int r = c / (tw * 3), k = c - r * (tw * 3);
            bgt[r * (TT_W * 3) + k] = __ldg(bgf + ((int64_t)(I0 + r) * w + J0) * 3 + k);
        }
    }
    if (tiled) {
        const int warp = tid >> 5, lane = tid & 31;
        const int ngrp = (RW + 95) / 96;
        const int nrow = (RH - warp + TT_THREADS / 32 - 1) / (TT_THREADS / 32);   // rows of this warp
        const int nit = max(nrow, 0) * ngrp;
        float2 fnext[3];
        if (FLOW && nit > 0) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int cc = u * 32 + lane;
                fnext[u] = (cc < RW) ? __ldg(bf + ((rmin + warp) * w + cmin + cc)) : make_float2(0.f, 0.f);
            }
        }
        for (int it = 0; it < nit; ++it) {
            const int rr = it / ngrp, g = it - rr * ngrp;
            const int r = warp + rr * (TT_THREADS / 32), c0 = g * 96;
            const int qi = rmin + r;
            uint2 e[3];
            if (FLOW) {
                float2 fb[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) fb[u] = fnext[u];
                if (it + 1 < nit) {
                    const int rr2 = (it + 1) / ngrp, g2 = (it + 1) - rr2 * ngrp;
                    const int r2 = warp + rr2 * (TT_THREADS / 32);
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int cc = g2 * 96 + u * 32 + lane;
                        fnext[u] = (cc < RW) ? __ldg(bf + ((rmin + r2) * w + cmin + cc)) : make_float2(0.f, 0.f);
                    }
                }
                const float fi = (float)qi;
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int qj = cmin + min(c0 + u * 32 + lane, RW - 1);
                    VmFlowPx px;
                    if (ff) px = vm_flow_px<true>(fg32, ff, h, w, qi, qj, fi, (float)qj, fb[u], flags);
                    else px = vm_flow_px<false>(fg32, ff, h, w, qi, qj, fi, (float)qj, fb[u], flags);
                    e[u].x = px.bgr;
                    e[u].y = px.masked ? 0u : vm_alpha_code(px.ta);
                }
            } else {
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int cc = c0 + u * 32 + lane;
                    const uint32_t sfg = (cc < RW) ? __ldg(fg32 + (qi * w + cmin + cc)) : 0u;
                    e[u].x = sfg & 0x00FFFFFFu;
                    e[u].y = vm_alpha_code((sfg >> 24) * 1024u);       // A/255 = 1024 A / 261120
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u)
                if (c0 + u * 32 + lane < RW) S.inter[r * RW + c0 + u * 32 + lane] = e[u];
        }
    }
    if (bg_async) asm volatile("cp.async.wait_group 0;");
    __syncthreads();

    // ---- P4: per-pixel resampling + composite (two rows per step for ILP) --------------------
    const int jc = tid & (TT_W - 1);
    int outside = 0;
    if (jc < tw) {
        const vm_axis_entry ce = S.cols[jc];
        const int c0 = ce.i0 - kc0, c1 = ce.i1 - kc0;
        const double yf = ce.frac, y1 = 1.0 - yf;
        const int j = J0 + jc;
        constexpr int RSTEP = TT_THREADS / TT_W;                       // 4 rows between a thread's pixels
        for (int ir0 = tid / TT_W; ir0 < th; ir0 += 2 * RSTEP) {
            float4 o[2];
            double t0v[2], t1v[2];
            uint2 ev[2][4];
            float wv[2][4];
            bool fast[2], live[2];
            unsigned unc = 0;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int ir = ir0 + z * RSTEP;
                live[z] = ir < th;
                const vm_axis_entry re = S.rows[min(ir, th - 1)];
                const int o0 = (re.i0 - kr0) * nkc, o1 = (re.i1 - kr0) * nkc;
                const double xf = re.frac, x1 = 1.0 - xf;
                const double2 T00 = S.T[o0 + c0], T01 = S.T[o0 + c1], T10 = S.T[o1 + c0], T11 = S.T[o1 + c1];
                // bilinear up-sampling of the transform (tps.py:68,73); weights formed once for
                // both coordinates (differs from the reference's operation order by < 1e-12 px)
                const double u00 = x1 * y1, u01 = x1 * yf, u10 = xf * y1, u11 = xf * yf;
                const double t0 = fma(T11.x, u11, fma(T10.x, u10, fma(T01.x, u01, T00.x * u00)));
                const double t1 = fma(T11.y, u11, fma(T10.y, u10, fma(T01.y, u01, T00.y * u00)));
                t0v[z] = t0; t1v[z] = t1;
                // map_coordinates geometry: floor via round-to-nearest magic + fix-up
                const double m0 = t0 + 6755399441055744.0, m1 = t1 + 6755399441055744.0;
                const double d0 = t0 - (m0 - 6755399441055744.0), d1 = t1 - (m1 - 6755399441055744.0);
                const int n0 = __double2loint(m0) - (d0 < 0.0 ? 1 : 0);
                const int n1 = __double2loint(m1) - (d1 < 0.0 ? 1 : 0);
                // inside the staged source box (which is clamped to the frame, so this also implies
                // 0 <= n0 < h-1 and 0 <= n1 < w-1); anything else goes the exact per-pixel way
                fast[z] = tiled && (unsigned)(n0 - rmin) <= (unsigned)(RH - 2) && (unsigned)(n1 - cmin) <= (unsigned)(RW - 2);
                const int q = fast[z] ? (n0 - rmin) * RW + (n1 - cmin) : 0, qs = fast[z] ? RW : 0;
                ev[z][0] = S.inter[q]; ev[z][1] = S.inter[q + 1];
                ev[z][2] = S.inter[q + qs]; ev[z][3] = S.inter[q + qs + 1];
                const float af = (float)d0 + (d0 < 0.0 ? 1.f : 0.f), bfr = (float)d1 + (d1 < 0.0 ? 1.f : 0.f);
                const float a0f = 1.f - af, b0f = 1.f - bfr;
                wv[z][0] = a0f * b0f; wv[z][1] = a0f * bfr; wv[z][2] = af * b0f; wv[z][3] = af * bfr;
            }
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int ir = min(ir0 + z * RSTEP, th - 1);
                const uint8_t *bp = bgt + ir * (TT_W * 3) + jc * 3;
                const float bb = vm_u2f(bp[0]), bgc = vm_u2f(bp[1]), br = vm_u2f(bp[2]);
                float col[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float v = __fmaf_rn(vm_byte2f(ev[z][3].x, c), wv[z][3], __fmaf_rn(vm_byte2f(ev[z][2].x, c), wv[z][2],
                                    __fmaf_rn(vm_byte2f(ev[z][1].x, c), wv[z][1], vm_byte2f(ev[z][0].x, c) * wv[z][0])));
                    col[c] = (v + 12582912.f) - 12582912.f;              // nearest integer
                    if (fabsf(v - col[c]) > 0.4995f) unc |= 1u << (z * 4 + c);   // float32 cannot decide
                }
                float al[4], nl[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) vm_alpha_decode(ev[z][k].y, al[k], nl[k]);
                const float a2 = __fmaf_rn(al[3], wv[z][3], __fmaf_rn(al[2], wv[z][2], __fmaf_rn(al[1], wv[z][1], al[0] * wv[z][0])));
                const float na = __fmaf_rn(nl[3], wv[z][3], __fmaf_rn(nl[2], wv[z][2], __fmaf_rn(nl[1], wv[z][1], nl[0] * wv[z][0])));
                o[z].x = __fmaf_rn(a2, col[0], na * bb);
                o[z].y = __fmaf_rn(a2, col[1], na * bgc);
                o[z].z = __fmaf_rn(a2, col[2], na * br);
                o[z].w = a2;
                if (!fast[z]) unc |= 8u << (z * 4);
            }
            if (unc) {
                // rare: exact float64 re-evaluation (knife-edge samples, frame borders, gather tiles)
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    if (!((unc >> (z * 4)) & 15u) || !live[z]) continue;
                    const int ir = ir0 + z * RSTEP;
                    const uint8_t *bp = bgt + ir * (TT_W * 3) + jc * 3;
                    const float bb = vm_u2f(bp[0]), bgc = vm_u2f(bp[1]), br = vm_u2f(bp[2]);
                    const VmBilin64 s = vm_mapcoord_setup(t0v[z], t1v[z], h, w);
                    if (fast[z]) {
                        float col[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c)
                            col[c] = (float)vm_round_half_up_u8(vm_mapcoord_blend(
                                s, (double)((ev[z][0].x >> (8 * c)) & 255u), (double)((ev[z][1].x >> (8 * c)) & 255u),
                                (double)((ev[z][2].x >> (8 * c)) & 255u), (double)((ev[z][3].x >> (8 * c)) & 255u)));
                        float al[4], nl[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) vm_alpha_decode(ev[z][k].y, al[k], nl[k]);
                        const float a2 = o[z].w;
                        const float na = __fmaf_rn(nl[3], wv[z][3], __fmaf_rn(nl[2], wv[z][2], __fmaf_rn(nl[1], wv[z][1], nl[0] * wv[z][0])));
                        o[z].x = __fmaf_rn(a2, col[0], na * bb);
                        o[z].y = __fmaf_rn(a2, col[1], na * bgc);
                        o[z].z = __fmaf_rn(a2, col[2], na * br);
                    } else if (s.inside) {
                        // last row / column of the frame, or a tile on the gather path
                        const uint8_t *fg8 = reinterpret_cast<const uint8_t *>(fg32);
                        const VmSrcPx s00 = vm_src_px<FLOW>(fg8, bf, ff, h, w, s.i0, s.j0, flags);
                        const VmSrcPx s01 = vm_src_px<FLOW>(fg8, bf, ff, h, w, s.i0, s.j1, flags);
                        const VmSrcPx s10 = vm_src_px<FLOW>(fg8, bf, ff, h, w, s.i1, s.j0, flags);
                        const VmSrcPx s11 = vm_src_px<FLOW>(fg8, bf, ff, h, w, s.i1, s.j1, flags);
                        const float cb = (float)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.b, s01.b, s10.b, s11.b));
                        const float cg = (float)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.g, s01.g, s10.g, s11.g));
                        const float cr = (float)vm_round_half_up_u8(vm_mapcoord_blend(s, s00.r, s01.r, s10.r, s11.r));
                        const double a64 = vm_mapcoord_blend(s, s00.a, s01.a, s10.a, s11.a);
                        const float a2 = (float)a64, na = (float)(1.0 - a64);
                        o[z] = make_float4(__fmaf_rn(a2, cb, na * bb), __fmaf_rn(a2, cg, na * bgc), __fmaf_rn(a2, cr, na * br), a2);
                    } else {
                        o[z] = make_float4(bb, bgc, br, 0.f);
                        outside++;
                    }
                }
            }
#pragma unroll
            for (int z = 0; z < 2; ++z)
                if (live[z]) out[fbase + (I0 + ir0 + z * RSTEP) * w + j] = o[z];
        }
    }
    if (status) {
        outside = __reduce_add_sync(0xffffffffu, outside);
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((tid & 31) == 0) {
            if (outside) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside);
            if (flags & 1) atomicAdd(status + VM_STATUS_INDEX_ERR, 1);
            if (flags & 2) atomicAdd(status + VM_STATUS_NAN_ERR, 1);
        }
        if (tid == 0 && !tiled) atomicAdd(status + VM_STATUS_SLOW_TILES, 1);
    }
}

// ---------------------------------------------------------------------------------------
// Split pipeline, stage B: TPS resampling + composite with the source pixels gathered from
// global memory through L1/L2 (no shared-memory source box, so 4 CTAs per SM).
//
//   PACKED = true : src is the (n,h,w) uint2 {bgr, alpha code} intermediate written by stage A
//                   (k_flow_warp_mask_bgra<.., PACKED>) for the same frames - produced a few
//                   microseconds earlier, so it is served from the 126 MB L2;
//   PACKED = false: src is the BGRA clip itself (C3: no flow stage at all).
// Phases P0/P1 (spline on the tile's coarse sub-grid) are those of the tiled kernel.
// ---------------------------------------------------------------------------------------
template <bool PACKED>
__device__ __forceinline__ uint2 vm_ld_src(const void *__restrict__ src, int idx) {
    if (PACKED) return __ldg(reinterpret_cast<const uint2 *>(src) + idx);
    const uint32_t s = __ldg(reinterpret_cast<const uint32_t *>(src) + idx);
    return make_uint2(s & 0x00FFFFFFu, vm_alpha_code((s >> 24) * 1024u));
}

template <bool PACKED, int TH>
__global__ void __launch_bounds__(TT_THREADS, 4)
k_tps_gather2(const void *__restrict__ src_all, const uint8_t *__restrict__ bg, int n_bg, int frame0,
              const double *__restrict__ ctrl, const double *__restrict__ coef, int N, double step_x,
              double step_y, const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols,
              int h, int w, float4 *__restrict__ out, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoarseSmem<TH> &S = *reinterpret_cast<CoarseSmem<TH> *>(smem_raw);
    const int tid = threadIdx.x;
    const int frame = blockIdx.z, I0 = blockIdx.y * TH, J0 = blockIdx.x * TT_W;   // frame: index in this launch
    const int th = min(TH, h - I0), tw = min(TT_W, w - J0);
    const int64_t fbase = (int64_t)frame * h * w;
    const void *src = PACKED ? (const void *)(reinterpret_cast<const uint2 *>(src_all) + fbase)
                             : (const void *)(reinterpret_cast<const uint32_t *>(src_all) + fbase);
    int kr0, kc0, nkr, nkc;
    if (!vm_tile_coarse<TH>(S, ctrl, coef, N, step_x, step_y, rows, cols, frame, I0, J0, th, tw, status, kr0, kc0, nkr, nkc))
        return;
    const bool finite = !S.bad;

    // background tile -> (dead) log-table space, asynchronously
    uint8_t *bgt = reinterpret_cast<uint8_t *>(S.logtab);           // [TH][TT_W * 3]
    const uint8_t *bgf = bg + (int64_t)((frame0 + frame) % n_bg) * h * w * 3;
    const bool bg_async = (w & 15) == 0 && tw == TT_W && (reinterpret_cast<uintptr_t>(bg) & 15) == 0;
    if (bg_async) {
        for (int c = tid; c < th * (TT_W * 3 / 16); c += TT_THREADS) {
            const int r = c / (TT_W * 3 / 16), k = c - r * (TT_W * 3 / 16);
            const uint8_t *g = bgf + ((int64_t)(I0 + r) * w + J0) * 3 + k * 16;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(bgt + r * (TT_W * 3) + k * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g));
        }
        asm volatile("cp.async.commit_group;");
        asm volatile("cp.async.wait_group 0;");
    } else {
        for (int c = tid; c < th * tw * 3; c += TT_THREADS) {
            const int r = c / (tw * 3), k = c - r * (tw * 3);
            bgt[r * (TT_W * 3) + k] = __ldg(bgf + ((int64_t)(I0 + r) * w + J0) * 3 + k);
        }
    }
    __syncthreads();

    const int jc = tid & (TT_W - 1);
    int outside = 0;
    if (jc < tw) {
        const vm_axis_entry ce = S.cols[jc];
        const int c0 = ce.i0 - kc0, c1 = ce.i1 - kc0;
        const double yf = ce.frac, y1 = 1.0 - yf;
        const int j = J0 + jc;
        constexpr int RSTEP = TT_THREADS / TT_W;
        for (int ir0 = tid / TT_W; ir0 < th; ir0 += 2 * RSTEP) {
            float4 o[2];
            double t0v[2], t1v[2];
            uint2 ev[2][4];
            float wv[2][4];
            bool fast[2], live[2];
            unsigned unc = 0;
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int ir = ir0 + z * RSTEP;
                live[z] = ir < th;
                const vm_axis_entry re = S.rows[min(ir, th - 1)];
                const int o0 = (re.i0 - kr0) * nkc, o1 = (re.i1 - kr0) * nkc;
                const double xf = re.frac, x1 = 1.0 - xf;
                const double2 T00 = S.T[o0 + c0], T01 = S.T[o0 + c1], T10 = S.T[o1 + c0], T11 = S.T[o1 + c1];
                const double u00 = x1 * y1, u01 = x1 * yf, u10 = xf * y1, u11 = xf * yf;
                const double t0 = fma(T11.x, u11, fma(T10.x, u10, fma(T01.x, u01, T00.x * u00)));
                const double t1 = fma(T11.y, u11, fma(T10.y, u10, fma(T01.y, u01, T00.y * u00)));
                t0v[z] = t0; t1v[z] = t1;
                const double m0 = t0 + 6755399441055744.0, m1 = t1 + 6755399441055744.0;
                const double d0 = t0 - (m0 - 6755399441055744.0), d1 = t1 - (m1 - 6755399441055744.0);
                const int n0 = __double2loint(m0) - (d0 < 0.0 ? 1 : 0);
                const int n1 = __double2loint(m1) - (d1 < 0.0 ? 1 : 0);
                // strictly inside the frame with both neighbours (|t| < 2^31 is implied by `finite`)
                fast[z] = finite && (unsigned)n0 < (unsigned)(h - 1) && (unsigned)n1 < (unsigned)(w - 1);
                const int q = fast[z] ? n0 * w + n1 : 0, qs = fast[z] ? w : 0;
                ev[z][0] = vm_ld_src<PACKED>(src, q); ev[z][1] = vm_ld_src<PACKED>(src, q + 1);
                ev[z][2] = vm_ld_src<PACKED>(src, q + qs); ev[z][3] = vm_ld_src<PACKED>(src, q + qs + 1);
                const float af = (float)d0 + (d0 < 0.0 ? 1.f : 0.f), bfr = (float)d1 + (d1 < 0.0 ? 1.f : 0.f);
                const float a0f = 1.f - af, b0f = 1.f - bfr;
                wv[z][0] = a0f * b0f; wv[z][1] = a0f * bfr; wv[z][2] = af * b0f; wv[z][3] = af * bfr;
            }
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const int ir = min(ir0 + z * RSTEP, th - 1);
                const uint8_t *bp = bgt + ir * (TT_W * 3) + jc * 3;
                const float bb = vm_u2f(bp[0]), bgc = vm_u2f(bp[1]), br = vm_u2f(bp[2]);
                float col[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float v = __fmaf_rn(vm_byte2f(ev[z][3].x, c), wv[z][3], __fmaf_rn(vm_byte2f(ev[z][2].x, c), wv[z][2],
                                    __fmaf_rn(vm_byte2f(ev[z][1].x, c), wv[z][1], vm_byte2f(ev[z][0].x, c) * wv[z][0])));
                    col[c] = (v + 12582912.f) - 12582912.f;
                    if (fabsf(v - col[c]) > 0.4995f) unc |= 1u << (z * 4 + c);
                }
                float al[4], nl[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) vm_alpha_decode(ev[z][k].y, al[k], nl[k]);
                const float a2 = __fmaf_rn(al[3], wv[z][3], __fmaf_rn(al[2], wv[z][2], __fmaf_rn(al[1], wv[z][1], al[0] * wv[z][0])));
                const float na = __fmaf_rn(nl[3], wv[z][3], __fmaf_rn(nl[2], wv[z][2], __fmaf_rn(nl[1], wv[z][1], nl[0] * wv[z][0])));
                o[z].x = __fmaf_rn(a2, col[0], na * bb);
                o[z].y = __fmaf_rn(a2, col[1], na * bgc);
                o[z].z = __fmaf_rn(a2, col[2], na * br);
                o[z].w = a2;
                if (!fast[z]) unc |= 8u << (z * 4);
            }
            if (unc) {
                // rare: exact float64 re-evaluation (knife-edge samples, last row/column, outside)
#pragma unroll
                for (int z = 0; z < 2; ++z) {
                    if (!((unc >> (z * 4)) & 15u) || !live[z]) continue;
                    const int ir = ir0 + z * RSTEP;
                    const uint8_t *bp = bgt + ir * (TT_W * 3) + jc * 3;
                    const float bb = vm_u2f(bp[0]), bgc = vm_u2f(bp[1]), br = vm_u2f(bp[2]);
                    const VmBilin64 s = vm_mapcoord_setup(t0v[z], t1v[z], h, w);
                    if (!s.inside) {
                        o[z] = make_float4(bb, bgc, br, 0.f);
                        outside++;
                        continue;
                    }
                    if (!fast[z]) {                           // t on the last row / column: clamped neighbours
                        ev[z][0] = vm_ld_src<PACKED>(src, s.i0 * w + s.j0); ev[z][1] = vm_ld_src<PACKED>(src, s.i0 * w + s.j1);
                        ev[z][2] = vm_ld_src<PACKED>(src, s.i1 * w + s.j0); ev[z][3] = vm_ld_src<PACKED>(src, s.i1 * w + s.j1);
                    }
                    float col[3];
#pragma unroll
                    for (int c = 0; c < 3; ++c)
                        col[c] = (float)vm_round_half_up_u8(vm_mapcoord_blend(
                            s, (double)((ev[z][0].x >> (8 * c)) & 255u), (double)((ev[z][1].x >> (8 * c)) & 255u),
                            (double)((ev[z][2].x >> (8 * c)) & 255u), (double)((ev[z][3].x >> (8 * c)) & 255u)));
                    float al[4], nl[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) vm_alpha_decode(ev[z][k].y, al[k], nl[k]);
                    const float w0 = (float)(s.a0 * s.b0), w1 = (float)(s.a0 * s.b1), w2 = (float)(s.a1 * s.b0), w3 = (float)(s.a1 * s.b1);
                    const float a2 = __fmaf_rn(al[3], w3, __fmaf_rn(al[2], w2, __fmaf_rn(al[1], w1, al[0] * w0)));
                    const float na = __fmaf_rn(nl[3], w3, __fmaf_rn(nl[2], w2, __fmaf_rn(nl[1], w1, nl[0] * w0)));
                    o[z] = make_float4(__fmaf_rn(a2, col[0], na * bb), __fmaf_rn(a2, col[1], na * bgc),
                                       __fmaf_rn(a2, col[2], na * br), a2);
                }
            }
#pragma unroll
            for (int z = 0; z < 2; ++z)
                if (live[z]) out[fbase + (I0 + ir0 + z * RSTEP) * w + j] = o[z];
        }
    }
    if (status) {
        outside = __reduce_add_sync(0xffffffffu, outside);
        if ((tid & 31) == 0 && outside) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside);
    }
}

// option block (vm_set_option): variant 4 = lean split pipeline (default, vm_lean.cu), 0 = first split
// pipeline, 1 = per-pixel gather kernels, 2 = single shared-memory tiled kernel, 3 = persistent
// role-specialised pipeline (vm_pipe.cu).  The non-default ones are kept as independent
// implementations for differential tests.
static int g_opt_variant = 4;
static int g_opt_chunk = 4;                  // frames per stage-A/stage-B launch pair
static int g_opt_tile_h = 32;

extern "C" int vm_set_option(const char *key, int value) {
    if (!key) return VM_ERR_ARG;
    if (!strcmp(key, "fused_variant")) { g_opt_variant = value; return VM_OK; }
    if (!strcmp(key, "tile_h") && (value == 32 || value == 64)) { g_opt_tile_h = value; return VM_OK; }
    if (!strcmp(key, "chunk_frames") && value >= 1 && value <= 4096) { g_opt_chunk = value; return VM_OK; }
    if (!strcmp(key, "pipe_lead") && value >= 1 && value <= 4096) { g_vp_lead = value; return VM_OK; }
    if (!strcmp(key, "pipe_ring_rows") && value >= 16) { g_vp_ring_rows = value; return VM_OK; }
    if (!strcmp(key, "pipe_cring_rows") && value >= 16) { g_vp_cring_rows = value; return VM_OK; }
    if (!strcmp(key, "pipe_blocks") && value >= 0) { g_vp_blocks = value; return VM_OK; }
    if (!strcmp(key, "pipe_roles") && value >= 1 && value <= 7) { g_vp_roles = value; return VM_OK; }
    if (!strncmp(key, "lean_", 5) && vm_lean_set_option(key, value) == VM_OK) return VM_OK;
    if (!strncmp(key, "fuse_", 5) && vm_fuse_set_option(key, value) == VM_OK) return VM_OK;
    if (!strcmp(key, "flow_stage_layout") && (value == 0 || value == 1)) { g_vm_flow_stage_layout = value; return VM_OK; }
    vm_set_error("vm_set_option: unknown option %s=%d", key, value);
    return VM_ERR_ARG;
}

template <bool FLOW, int TH>
static int launch_tiled(const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                        int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                        double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                        float *out, int32_t *status, cudaStream_t st, const char *what) {
    const int tiles_x = (w + TT_W - 1) / TT_W, tiles_y = (h + TH - 1) / TH;
    const int64_t tiles = (int64_t)n * tiles_x * tiles_y;
    VM_REQUIRE(tiles < (1ll << 31), "too many tiles for one launch");
    const size_t smem = sizeof(TileSmem<TH>);
    cudaError_t e = cudaFuncSetAttribute(k_tps_tiled<FLOW, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { vm_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return VM_ERR_CUDA; }
    k_tps_tiled<FLOW, TH><<<(unsigned)tiles, TT_THREADS, smem, st>>>(
        fg, (const float2 *)backward, (const float2 *)forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y,
        rows, cols, h, w, tiles_x, tiles_y, (float4 *)out, status);
    return vm_check_launch(what);
}

int vm_launch_flow_stage(const uint8_t *fg, const float *backward, const float *forward, int n, int h, int w,
                         void *packed, int32_t *status, cudaStream_t st, bool raw_ta); // vm_flow.cu
int64_t vm_lean_scratch_bytes(int n, int h, int w);                                // vm_lean.cu
int vm_lean_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, void *scratch, int32_t *status, cudaStream_t st, const char *what);
int vm_fuse_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, int32_t *status, cudaStream_t st, const char *what);  // vm_fuse.cu
int64_t vm_pipe_scratch_bytes(int n, int h, int w);                                // vm_pipe.cu
int vm_pipe_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, void *scratch, int32_t *status, cudaStream_t st, const char *what);

template <bool PACKED, int TH>
static int launch_gather2(const void *src, const uint8_t *bg, int n_bg, int frame0, const double *ctrl,
                          const double *coef, int N, double step_x, double step_y, const vm_axis_entry *rows,
                          const vm_axis_entry *cols, int n, int h, int w, float *out, int32_t *status,
                          cudaStream_t st, const char *what) {
    const dim3 grid((w + TT_W - 1) / TT_W, (h + TH - 1) / TH, n);
    VM_REQUIRE(grid.y <= 65535 && n <= 65535, "too many tiles for one launch");
    const size_t smem = sizeof(CoarseSmem<TH>);
    k_tps_gather2<PACKED, TH><<<grid, TT_THREADS, smem, st>>>(src, bg, n_bg, frame0, ctrl, coef, N, step_x, step_y,
                                                              rows, cols, h, w, (float4 *)out, status);
    return vm_check_launch(what);
}

static int64_t fused_scratch_bytes(int n, int h, int w) {
    if (g_opt_variant == 0) return (int64_t)(n < g_opt_chunk ? n : g_opt_chunk) * h * w * 8;
    if (g_opt_variant == 1) return (int64_t)n * 2 * (h / 2) * (w / 2) * (int64_t)sizeof(double);
    return 0;
}

static int launch_fused(bool flow, const uint8_t *fg, const float *backward, const float *forward,
                        const uint8_t *bg, int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny,
                        double step_x, double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n,
                        int h, int w, float *out, void *scratch, int32_t *status, void *stream, const char *what) {
    if (n == 0) return VM_OK;
    int rc = vm_init();
    if (rc != VM_OK) return rc;
    VM_REQUIRE(h <= 32767 && w <= 32767 && (int64_t)h * w < (1ll << 28), "frame too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool small_n = N <= TT_MAX_N && h >= 2 && w >= 2;
    // variant 5: C4 in ONE kernel (vm_fuse.cu: every byte touches HBM once; measured slower than the lean split pipeline,
    // DESIGN.md 5f); C3 and anything it does not take goes to the lean pipeline
    if (g_opt_variant == 5 && flow && small_n && (int64_t)n * ((h + 59) / 60) * ((w + 59) / 60) < (1ll << 31))
        return vm_fuse_launch(forward ? 2 : 1, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y,
                              rows, cols, n, h, w, out, status, st, what);
    if ((g_opt_variant == 4 || g_opt_variant == 5) && small_n && nx <= h / 2 + 1 && ny <= w / 2 + 1)
        return vm_lean_launch(flow ? (forward ? 2 : 1) : 0, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny,
                              step_x, step_y, rows, cols, n, h, w, out, scratch, status, st, what);
    if (g_opt_variant == 3 && small_n && nx == h / 2 && ny == w / 2)
        return vm_pipe_launch(flow ? (forward ? 2 : 1) : 0, fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny,
                              step_x, step_y, rows, cols, n, h, w, out, scratch, status, st, what);
    if ((g_opt_variant == 0 || g_opt_variant == 3 || g_opt_variant == 4 || g_opt_variant == 5) && small_n) {
        // split pipeline: stage A (flow warp + mask -> packed intermediate, L2 resident) and stage B
        // (TPS + composite) per chunk of frames; C3 has no stage A.
        if (!flow)
            return launch_gather2<false, 32>(fg, bg, n_bg, 0, ctrl, coef, N, step_x, step_y, rows, cols, n, h, w, out,
                                             status, st, what);
        VM_REQUIRE(scratch, "scratch workspace (vm_fused_scratch_bytes) required");
        const int64_t px = (int64_t)h * w;
        for (int f0 = 0; f0 < n; f0 += g_opt_chunk) {
            const int m = (n - f0 < g_opt_chunk) ? n - f0 : g_opt_chunk;
            rc = vm_launch_flow_stage(fg + f0 * px * 4, backward + f0 * px * 2, forward ? forward + f0 * px * 2 : nullptr,
                                      m, h, w, scratch, status, st, false);
            if (rc != VM_OK) return rc;
            rc = launch_gather2<true, 32>(scratch, bg, n_bg, f0, ctrl + (int64_t)f0 * N * 2, coef + (int64_t)f0 * (N + 3) * 2,
                                          N, step_x, step_y, rows, cols, m, h, w, out + f0 * px * 4, status, st, what);
            if (rc != VM_OK) return rc;
        }
        return VM_OK;
    }
    if (g_opt_variant == 2 && small_n) {
        if (g_opt_tile_h == 64)
            return flow ? launch_tiled<true, 64>(fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols, n, h, w, out, status, st, what)
                        : launch_tiled<false, 64>(fg, nullptr, nullptr, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols, n, h, w, out, status, st, what);
        return flow ? launch_tiled<true, 32>(fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols, n, h, w, out, status, st, what)
                    : launch_tiled<false, 32>(fg, nullptr, nullptr, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols, n, h, w, out, status, st, what);
    }
    // per-pixel gather variant: coarse transform through global scratch, then one pixel per thread
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
    // enough for every variant selectable with vm_set_option for this clip shape
    const int64_t a = (int64_t)(n < g_opt_chunk ? n : g_opt_chunk) * h * w * 8;
    const int64_t b = (int64_t)n * 2 * (h / 2) * (w / 2) * (int64_t)sizeof(double);
    (void)fused_scratch_bytes;
    const int64_t c = vm_pipe_scratch_bytes(n, h, w), d = vm_lean_scratch_bytes(n, h, w);
    int64_t m = a > b ? a : b;
    m = m > c ? m : c;
    return m > d ? m : d;
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
