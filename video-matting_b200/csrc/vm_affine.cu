// cv2.warpAffine (legacy fixed-point path), HSV illumination jitter and alpha statistics.
// Reference behaviour: augmentation.py:10-21, 44-63, 88-99 (see include/vm_b200.h).
#include "vm_common.cuh"
#include <math.h>
#include <mutex>

struct VmAffineInv { double i00, i01, b0, i10, i11, b1; };

// OpenCV's closed-form inverse (imgwarp.cpp, cv::warpAffine without WARP_INVERSE_MAP).
static VmAffineInv vm_affine_invert(const double *M) {
    double D = M[0] * M[4] - M[1] * M[3];
    D = D != 0 ? 1. / D : 0;
    const double A11 = M[4] * D, A22 = M[0] * D;
    VmAffineInv r;
    r.i00 = A11; r.i01 = M[1] * (-D);
    r.i10 = M[3] * (-D); r.i11 = A22;
    r.b0 = -r.i00 * M[2] - r.i01 * M[5];
    r.b1 = -r.i10 * M[2] - r.i11 * M[5];
    return r;
}

// AB_BITS = 10: adelta/bdelta per column, X0/Y0 per row (+16 rounding), >> 5 -> 1/32 px.
__device__ __forceinline__ void vm_affine_coords(const VmAffineInv &A, int x, int y, int &SX, int &SY) {
    const int adelta = vm_cvround_f64(__dmul_rn(__dmul_rn(A.i00, (double)x), 1024.0));
    const int bdelta = vm_cvround_f64(__dmul_rn(__dmul_rn(A.i10, (double)x), 1024.0));
    const int X0 = (int)((unsigned)vm_cvround_f64(__dmul_rn(__dadd_rn(__dmul_rn(A.i01, (double)y), A.b0), 1024.0)) + 16u);
    const int Y0 = (int)((unsigned)vm_cvround_f64(__dmul_rn(__dadd_rn(__dmul_rn(A.i11, (double)y), A.b1), 1024.0)) + 16u);
    SX = (int)((unsigned)X0 + (unsigned)adelta) >> 5;
    SY = (int)((unsigned)Y0 + (unsigned)bdelta) >> 5;
}

template <typename T, int C>
__global__ void __launch_bounds__(256)
k_warp_affine(const T *__restrict__ src, int sh, int sw, VmAffineInv A, int dh, int dw, T *__restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    int SX, SY;
    vm_affine_coords(A, x, y, SX, SY);
    T out[C];
    vm_sample_fixed<T, C>(src, sh, sw, SX, SY, out);
    T *o = dst + ((int64_t)y * dw + x) * C;
#pragma unroll
    for (int c = 0; c < C; ++c) o[c] = out[c];
}

extern "C" int vm_warp_affine(const void *src, int dtype, int channels, int sh, int sw,
                              const double *M_host, int dh, int dw, void *dst, void *stream) {
    VM_REQUIRE(src && M_host && dst, "null pointer");
    VM_REQUIRE(sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && dh < 65536, "bad size");
    const VmAffineInv A = vm_affine_invert(M_host);
    dim3 grid((dw + 255) / 256, dh);
    cudaStream_t st = (cudaStream_t)stream;
#define VM_WA(T, C) k_warp_affine<T, C><<<grid, 256, 0, st>>>((const T *)src, sh, sw, A, dh, dw, (T *)dst)
    if (dtype == VM_U8 && channels == 1) VM_WA(uint8_t, 1);
    else if (dtype == VM_U8 && channels == 3) VM_WA(uint8_t, 3);
    else if (dtype == VM_U8 && channels == 4) VM_WA(uint8_t, 4);
    else if (dtype == VM_F32 && channels == 1) VM_WA(float, 1);
    else if (dtype == VM_F32 && channels == 3) VM_WA(float, 3);
    else if (dtype == VM_F64 && channels == 1) VM_WA(double, 1);
    else if (dtype == VM_F64 && channels == 3) VM_WA(double, 3);
    else { vm_set_error("vm_warp_affine: unsupported dtype/channels %d/%d", dtype, channels); return VM_ERR_ARG; }
#undef VM_WA
    return vm_check_launch("vm_warp_affine");
}

// ---------------------------------------------------------------------------------------
// change_illumination: BGR2HSV (integer tables, hsv_shift = 12) -> S/V LUT -> HSV2BGR (float)
// ---------------------------------------------------------------------------------------
struct VmLut256 { uint8_t v[256]; };

// one pixel of change_illumination: BGR2HSV (integer tables) -> S/V through `slut` -> HSV2BGR, bit-exact for
// OpenCV 4.13's 8-bit HSV2BGR (probed over all 180 x 256 x 256 inputs against cv2, see the CPU test suite):
//   s = S*(1/255.f), v = V*(1/255.f), hh = H*(6.f/180.f), f = hh - floor(hh),
//   tab = {v, v*(1-s), v*fma(-s,f,1), v*fma(-s,1-f,1)}, out = tab*255.f,
//   TRUNCATED in the SIMD body of a row (`trunc`), rounded half-to-even in the row's scalar tail.
// Returns B | G<<8 | R<<16.
__device__ __forceinline__ uint32_t vm_illum_px(int b, int g, int r, const int *__restrict__ sdiv, const int *__restrict__ hdiv,
                                                const uint8_t *__restrict__ slut, bool trunc) {
    const int v = max(max(b, g), r), vmin = min(min(b, g), r), d = v - vmin;
    int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
    const int s = (d * sdiv[v] + (1 << 11)) >> 12;
    hh = (hh * hdiv[d] + (1 << 11)) >> 12;
    if (hh < 0) hh += 180;
    const int s2 = slut[s], v2 = slut[v];
    const float fv = __fmul_rn((float)v2, 1.f / 255.f), fs = __fmul_rn((float)s2, 1.f / 255.f);
    float ob = fv, og = fv, orr = fv;
    if (s2 != 0) {
        float hf = __fmul_rn((float)hh, 6.f / 180.f);
        const float fl = floorf(hf);
        int sec = (int)fl;
        hf = __fsub_rn(hf, fl);
        if ((unsigned)sec >= 6u) { sec = 0; hf = 0.f; }
        const float p_ = __fmul_rn(fv, __fsub_rn(1.f, fs));
        const float q_ = __fmul_rn(fv, __fmaf_rn(-fs, hf, 1.f));
        const float t_ = __fmul_rn(fv, __fmaf_rn(-fs, __fsub_rn(1.f, hf), 1.f));
        // OpenCV's sector table {v, p, q, t}[{1,1,3,0,0,2} / {3,0,0,2,1,1} / {0,2,1,1,3,0}][sec] as selects (an indexed
        // local array would live in local memory)
        ob = sec < 2 ? p_ : (sec == 2 ? t_ : (sec == 5 ? q_ : fv));
        og = sec == 0 ? t_ : (sec < 3 ? fv : (sec == 3 ? q_ : p_));
        orr = (sec == 0 || sec == 5) ? fv : (sec == 1 ? q_ : (sec == 4 ? t_ : p_));
    }
    const float xb = __fmul_rn(ob, 255.f), xg = __fmul_rn(og, 255.f), xr = __fmul_rn(orr, 255.f);
    const uint32_t B = (uint32_t)max(0, min(255, trunc ? __float2int_rz(xb) : __float2int_rn(xb)));
    const uint32_t G = (uint32_t)max(0, min(255, trunc ? __float2int_rz(xg) : __float2int_rn(xg)));
    const uint32_t R = (uint32_t)max(0, min(255, trunc ? __float2int_rz(xr) : __float2int_rn(xr)));
    return B | (G << 8) | (R << 16);
}

// cv2 converts `vec` pixels per SIMD step and rows independently: pixel x of a row of w pixels is in the truncating
// SIMD body iff x < w - w % vec
__device__ __forceinline__ bool vm_hsv_body(int x, int w, int vec) { return x < w - w % vec; }

__global__ void __launch_bounds__(256)
k_illumination(const uint8_t *__restrict__ bgr, int64_t rows, int w, int vec, VmLut256 lut, uint8_t *__restrict__ out) {
    __shared__ int sdiv[256], hdiv[256];
    __shared__ uint8_t slut[256];
    {
        const int k = threadIdx.x;
        sdiv[k] = k ? __double2int_rn(1044480.0 / (double)k) : 0;            // 255 << 12
        hdiv[k] = k ? __double2int_rn(737280.0 / (6.0 * (double)k)) : 0;     // 180 << 12
        slut[k] = lut.v[k];
    }
    __syncthreads();
    const int64_t npx = rows * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int x = (int)(p % w);
        const uint32_t o = vm_illum_px(bgr[p * 3], bgr[p * 3 + 1], bgr[p * 3 + 2], sdiv, hdiv, slut, vm_hsv_body(x, w, vec));
        out[p * 3] = (uint8_t)o; out[p * 3 + 1] = (uint8_t)(o >> 8); out[p * 3 + 2] = (uint8_t)(o >> 16);
    }
}

extern "C" int vm_illumination_lut_rows(const uint8_t *bgr, int64_t rows, int w, const uint8_t *lut_host, int vec,
                                        uint8_t *out, void *stream) {
    VM_REQUIRE(bgr && lut_host && out, "null pointer");
    VM_REQUIRE(rows >= 0 && w >= 0 && vec >= 1 && vec <= 1024, "bad size");
    if (rows * w <= 0) return VM_OK;
    VmLut256 lut;
    for (int k = 0; k < 256; ++k) lut.v[k] = lut_host[k];
    k_illumination<<<min(vm_blocks(rows * w, 256), 148u * 16u), 256, 0, (cudaStream_t)stream>>>(bgr, rows, w, vec, lut, out);
    return vm_check_launch("vm_change_illumination");
}

// the buffer as ONE row of npx pixels on an AVX2 host (32 pixels per SIMD step); images go through vm_illumination_lut_rows
extern "C" int vm_illumination_lut(const uint8_t *bgr, int64_t npx, const uint8_t *lut_host,
                                   uint8_t *out, void *stream) {
    VM_REQUIRE(npx < (1ll << 31), "too many pixels for one row");
    return vm_illumination_lut_rows(bgr, 1, (int)npx, lut_host, 32, out, stream);
}

extern "C" int vm_change_illumination(const uint8_t *bgr, int64_t npx, double a, double b, double c,
                                      uint8_t *out, void *stream) {
    // augmentation.py:91-98: clip(a*(x/255.)**b + c, 0, 1), (255.*new).astype(uint8) (truncation)
    uint8_t lut[256];
    for (int k = 0; k < 256; ++k) {
        double nv = a * pow((double)k / 255., b) + c;
        nv = nv < 0. ? 0. : (nv > 1. ? 1. : nv);
        lut[k] = (uint8_t)(255. * nv);
    }
    return vm_illumination_lut(bgr, npx, lut, out, stream);
}

// ---------------------------------------------------------------------------------------
// cv2.resize(uint8, INTER_LINEAR) (reader.py:41,53, augmentation.py:160), bit-exact for OpenCV 4.13:
// 11-bit coefficients from float32 fractions, int32 horizontal pass, vertical pass
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; the horizontal fraction is forced to 0 at the row
// ends, vertically rows are clipped instead; exact 2x reductions are the 2x2 block mean (OpenCV switches to INTER_AREA).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void vm_resize_coef(int d, double scale, int sn, bool clamp_frac, int &i, int &w0, int &w1) {
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    const float fl = floorf(f);
    i = (int)fl;
    f = __fsub_rn(f, fl);
    if (clamp_frac) {
        if (i < 0) { i = 0; f = 0.f; }
        if (i >= sn - 1) { i = sn - 1; f = 0.f; }
    }
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

template <int C>
__global__ void __launch_bounds__(256)
k_resize_u8(const uint8_t *__restrict__ src, int sh, int sw, uint8_t *__restrict__ dst, int dh, int dw, double scale_x,
            double scale_y, int area2) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, img = blockIdx.z;
    if (x >= dw) return;
    const uint8_t *s = src + (int64_t)img * sh * sw * C;
    uint8_t *o = dst + (((int64_t)img * dh + y) * dw + x) * C;
    if (area2) {
        const uint8_t *p0 = s + ((int64_t)(2 * y) * sw + 2 * x) * C, *p1 = p0 + (int64_t)sw * C;
#pragma unroll
        for (int c = 0; c < C; ++c) o[c] = (uint8_t)((p0[c] + p0[C + c] + p1[c] + p1[C + c] + 2) >> 2);
        return;
    }
    int xi, a0, a1, yi, b0, b1;
    vm_resize_coef(x, scale_x, sw, true, xi, a0, a1);
    vm_resize_coef(y, scale_y, sh, false, yi, b0, b1);
    const int x1 = min(xi + 1, sw - 1), r0 = max(0, min(yi, sh - 1)), r1 = max(0, min(yi + 1, sh - 1));
    const uint8_t *q0 = s + (int64_t)r0 * sw * C, *q1 = s + (int64_t)r1 * sw * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int S0 = __ldg(q0 + xi * C + c) * a0 + __ldg(q0 + x1 * C + c) * a1;
        const int S1 = __ldg(q1 + xi * C + c) * a0 + __ldg(q1 + x1 * C + c) * a1;
        const int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)max(0, min(255, v));
    }
}

extern "C" int vm_resize_u8(const uint8_t *src, int n, int sh, int sw, int channels, uint8_t *dst, int dh, int dw, void *stream) {
    VM_REQUIRE(src && dst, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && dh < 65536, "bad size");
    if (n == 0) return VM_OK;
    const double scale_x = 1.0 / ((double)dw / (double)sw), scale_y = 1.0 / ((double)dh / (double)sh);   // OpenCV: 1. / inv_scale
    const int area2 = (sw == 2 * dw && sh == 2 * dh) ? 1 : 0;
    const dim3 grid((dw + 255) / 256, dh, n);
    cudaStream_t st = (cudaStream_t)stream;
    if (channels == 1) k_resize_u8<1><<<grid, 256, 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, area2);
    else if (channels == 3) k_resize_u8<3><<<grid, 256, 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, area2);
    else if (channels == 4) k_resize_u8<4><<<grid, 256, 0, st>>>(src, sh, sw, dst, dh, dw, scale_x, scale_y, area2);
    else { vm_set_error("vm_resize_u8: channels must be 1, 3 or 4"); return VM_ERR_ARG; }
    return vm_check_launch("vm_resize_u8");
}

// ---------------------------------------------------------------------------------------
// object_size / fg_center reductions: count(alpha != 0), sum(row), sum(col)
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_alpha_stats(const T *__restrict__ alpha, int h, int w, unsigned long long *__restrict__ out) {
    unsigned long long cnt = 0, si = 0, sj = 0;
    const int64_t n = (int64_t)h * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        if (alpha[p] != T(0)) {
            const int i = (int)(p / w);
            cnt += 1; si += (unsigned long long)i; sj += (unsigned long long)(p - (int64_t)i * w);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        si += __shfl_down_sync(0xffffffffu, si, o);
        sj += __shfl_down_sync(0xffffffffu, sj, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(out, cnt); atomicAdd(out + 1, si); atomicAdd(out + 2, sj);
    }
}

extern "C" int vm_alpha_stats(const void *alpha, int dtype, int h, int w, unsigned long long *out, void *stream) {
    VM_REQUIRE(alpha && out, "null pointer");
    const int64_t n = (int64_t)h * w;
    if (n <= 0) return VM_OK;
    const unsigned grid = min(vm_blocks(n, 256), 148u * 8u);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case VM_U8:  k_alpha_stats<uint8_t><<<grid, 256, 0, st>>>((const uint8_t *)alpha, h, w, out); break;
    case VM_F32: k_alpha_stats<float><<<grid, 256, 0, st>>>((const float *)alpha, h, w, out); break;
    case VM_F64: k_alpha_stats<double><<<grid, 256, 0, st>>>((const double *)alpha, h, w, out); break;
    default: vm_set_error("vm_alpha_stats: unsupported dtype %d", dtype); return VM_ERR_ARG;
    }
    return vm_check_launch("vm_alpha_stats");
}

// ---------------------------------------------------------------------------------------
// Batched augmentation (augmentation.py:102-135 for a whole clip): alpha statistics of BGRA frames, and
// the fused tail of warp_image + change_illumination:
//   warpAffine([[1,0,tu],[0,1,tv]], (w,h))  - an exact integer shift of the (sh, sw) source, 0 outside -
//   warpAffine(getRotationMatrix2D(...), (w,h)) - cv2 fixed-point bilinear of that shifted (h, w) image -
//   change_illumination (uint8 colours only).
// Each tap (Y, X) of the second pass is the shifted image's pixel: source[Y - tv][X - tu] when 0 <= Y < h,
// 0 <= X < w and the source index is inside (sh, sw); 0 otherwise - so the two passes need no intermediate.
// ---------------------------------------------------------------------------------------
struct VmAugParams { double M[6]; int tu, tv; };         // M: 2x3 matrix of the second pass; (tu, tv) of the first
#define VA_AFF_ROWS 16                                   // rows per CTA of k_aug_affine (amortises the table set-up)

__global__ void __launch_bounds__(256)
k_alpha_stats_bgra(const uint32_t *__restrict__ bgra, int h, int w, unsigned long long *__restrict__ out) {
    const int frame = blockIdx.y;
    const uint32_t *a = bgra + (int64_t)frame * h * w;
    unsigned long long cnt = 0, si = 0, sj = 0;
    const int64_t n = (int64_t)h * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        if (__ldg(a + p) >> 24) {
            const int i = (int)(p / w);
            cnt += 1; si += (unsigned long long)i; sj += (unsigned long long)(p - (int64_t)i * w);
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        si += __shfl_down_sync(0xffffffffu, si, o);
        sj += __shfl_down_sync(0xffffffffu, sj, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt) {
        atomicAdd(out + frame * 3, cnt); atomicAdd(out + frame * 3 + 1, si); atomicAdd(out + frame * 3 + 2, sj);
    }
}

// out (n, 3) uint64 {count(A != 0), sum(rows), sum(cols)} per frame; `out` must be zeroed by the caller
extern "C" int vm_alpha_stats_bgra(const uint8_t *bgra, int n, int h, int w, unsigned long long *out, void *stream) {
    VM_REQUIRE(bgra && out, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && h >= 1 && w >= 1, "bad size");
    if (n == 0) return VM_OK;
    const dim3 grid(min(vm_blocks((int64_t)h * w, 1024), 592u), n);
    k_alpha_stats_bgra<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(bgra), h, w, out);
    return vm_check_launch("vm_alpha_stats_bgra");
}

// BGR2HSV division tables of OpenCV (hsv_shift = 12): the same for every launch, computed once per device
__device__ int g_va_sdiv[256], g_va_hdiv[256];
__global__ void k_va_tables() {
    const int k = threadIdx.x;
    g_va_sdiv[k] = k ? __double2int_rn(1044480.0 / (double)k) : 0;            // 255 << 12
    g_va_hdiv[k] = k ? __double2int_rn(737280.0 / (6.0 * (double)k)) : 0;     // 180 << 12
}

// three bytes at byte offset `o` of a uint8 buffer through aligned 32-bit loads (two words and a funnel shift instead
// of three byte loads); `last` = index of the last whole word of the buffer, so nothing is read behind it
__device__ __forceinline__ uint32_t vm_ld3(const uint32_t *__restrict__ words, int64_t o, int64_t last) {
    const int64_t wi = o >> 2;
    const uint32_t lo = __ldg(words + wi), hi = __ldg(words + min(wi + 1, last));
    return __funnelshift_r(lo, hi, ((unsigned)o & 3u) * 8u) & 0x00FFFFFFu;
}

// FG: source = packed (sh, sw) = (h+1, w+1) intermediate {bgr, alpha float bits} of vm_aug_tps -> new_fg uint8 x3
// (after the illumination change) + new_alpha float32.  !FG: source = (h, w, 3) uint8 background -> new_bg.
// CTA = 256 columns x VA_AFF_ROWS rows, one column per thread.  Per-row terms of the coordinate transform come from
// shared memory (they are the same for the whole row), the division tables from global memory.  W4 (w % 4 == 0 and
// 4-byte aligned planes): *interior* pixels - all four taps inside the shifted image and inside the source, the bulk of
// a frame - take a branch-free path: one 32-bit offset from the frame's base pointer, the background's two taps of a
// row as three aligned words + funnel shifts, the lane-packed blend of vm_blend_lanes (same integers as VmTap<uint8_t>),
// float32 alpha taps blended in float32 (the float64 plane, when carried, in cv2's float64 order); the three colour
// bytes of four neighbouring lanes leave as three 32-bit stores (one shuffle per lane).  Everything else goes through the
// tap-by-tap code with OpenCV's border rule.  ncu before this path: 366 executed instructions per pixel, issue slots
// 82 % busy - the kernel was bound by its predicates, branches and 64-bit index arithmetic, not by memory.
template <bool FG, bool W4>
__global__ void __launch_bounds__(256)
k_aug_affine(const void *__restrict__ src_all, const double *__restrict__ alpha64, const VmAugParams *__restrict__ params,
             const uint8_t *__restrict__ luts, int h, int w, uint8_t *__restrict__ out_bgr, float *__restrict__ out_alpha,
             double *__restrict__ out_alpha64, int vec, int64_t src_last_word) {
    __shared__ int sdiv[256], hdiv[256];
    __shared__ uint8_t slut[256];
    __shared__ VmAffineInv Ash;
    __shared__ int tsh[2];
    __shared__ int rowX0[VA_AFF_ROWS], rowY0[VA_AFF_ROWS];
    const int frame = blockIdx.z;
    const int y0 = blockIdx.y * VA_AFF_ROWS;
    {
        const int k = threadIdx.x;
        sdiv[k] = g_va_sdiv[k];
        hdiv[k] = g_va_hdiv[k];
        slut[k] = luts[frame * 256 + k];
        if (k < 32) {
            const VmAugParams P = params[frame];
            double D = P.M[0] * P.M[4] - P.M[1] * P.M[3];                      // OpenCV's closed-form inverse
            D = D != 0 ? 1. / D : 0;
            VmAffineInv r;
            r.i00 = P.M[4] * D; r.i01 = P.M[1] * (-D);
            r.i10 = P.M[3] * (-D); r.i11 = P.M[0] * D;
            r.b0 = -r.i00 * P.M[2] - r.i01 * P.M[5];
            r.b1 = -r.i10 * P.M[2] - r.i11 * P.M[5];
            if (k == 0) { Ash = r; tsh[0] = P.tu; tsh[1] = P.tv; }
            if (k < VA_AFF_ROWS) {                                             // the row terms of vm_affine_coords
                const double y = (double)(y0 + k);
                rowX0[k] = (int)((unsigned)vm_cvround_f64(__dmul_rn(__dadd_rn(__dmul_rn(r.i01, y), r.b0), 1024.0)) + 16u);
                rowY0[k] = (int)((unsigned)vm_cvround_f64(__dmul_rn(__dadd_rn(__dmul_rn(r.i11, y), r.b1), 1024.0)) + 16u);
            }
        }
    }
    __syncthreads();
    const int sh = FG ? h + 1 : h, sw = FG ? w + 1 : w;
    const int tu = tsh[0], tv = tsh[1];
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = x < w;
    if (!W4 && !live) return;                                              // W4: whole groups of four lanes stay for the shuffle
    const int adelta = vm_cvround_f64(__dmul_rn(__dmul_rn(Ash.i00, (double)x), 1024.0));
    const int bdelta = vm_cvround_f64(__dmul_rn(__dmul_rn(Ash.i10, (double)x), 1024.0));
    const bool body = vm_hsv_body(x, w, vec);
    const int yend = min(y0 + VA_AFF_ROWS, h);
    const int64_t fsrc = (int64_t)frame * sh * sw;
    // interior: iy in [ylo, yhi] and ix in [xlo, xhi] <=> 0 <= Y, Y + 1 < h, 0 <= Y - tv, Y + 1 - tv < sh (same for X)
    const int ylo = max(0, tv), yhi = min(h - 2, sh - 2 + tv), xlo = max(0, tu), xhi = min(w - 2, sw - 2 + tu);
    const bool any_fast = W4 && yhi >= ylo && xhi >= xlo && (int64_t)sh * sw < (1ll << 28);
    const unsigned yspan = (unsigned)(yhi - ylo), xspan = (unsigned)(xhi - xlo);
    const int qoff = tv * sw + tu;
    const uint2 *fsrc2 = reinterpret_cast<const uint2 *>(src_all) + (FG ? fsrc : 0);               // FG: packed {bgr, alpha} frame
    const uint32_t *fsrcw = reinterpret_cast<const uint32_t *>(src_all) + (FG ? 0 : (fsrc * 3 >> 2));   // !FG, W4: rows are 4-byte aligned
    asm volatile("" : "+l"(fsrc2));
    asm volatile("" : "+l"(fsrcw));
    for (int y = y0; y < yend; ++y) {
        const int SX = (int)((unsigned)rowX0[y - y0] + (unsigned)adelta) >> 5;
        const int SY = (int)((unsigned)rowY0[y - y0] + (unsigned)bdelta) >> 5;
        const int fx = SX & 31, fy = SY & 31;
        uint32_t bgr;
        float af = 0.f;
        double ad = 0.0;
        if (any_fast && live && (unsigned)((SY >> 5) - ylo) <= yspan && (unsigned)((SX >> 5) - xlo) <= xspan) {
            const unsigned q = (unsigned)((SY >> 5) * sw + (SX >> 5) - qoff);
            uint32_t c0, c1, c2, c3;
            if (FG) {
                const uint2 *g0 = fsrc2 + q, *g1 = g0 + (unsigned)sw;
                const uint2 e0 = __ldg(g0), e1 = __ldg(g0 + 1), e2 = __ldg(g1), e3 = __ldg(g1 + 1);
                c0 = e0.x & 0x00FFFFFFu; c1 = e1.x & 0x00FFFFFFu; c2 = e2.x & 0x00FFFFFFu; c3 = e3.x & 0x00FFFFFFu;
                if (alpha64) {
                    const double *ap = alpha64 + fsrc + q;
                    ad = VmTap<double>::blend(__ldg(ap), __ldg(ap + 1), __ldg(ap + sw), __ldg(ap + sw + 1), fx, fy);
                } else {
                    af = VmTap<float>::blend(__uint_as_float(e0.y), __uint_as_float(e1.y), __uint_as_float(e2.y), __uint_as_float(e3.y), fx, fy);
                }
            } else {
                // the two taps of a row are 6 consecutive bytes at byte offset 3 q: three aligned words and funnel shifts;
                // the row pitch 3 sw is a multiple of 4, so both rows share the byte phase
                const unsigned o = 3u * q, sh8 = (o & 3u) * 8u;
                const uint32_t *r0 = fsrcw + (o >> 2), *r1 = r0 + (unsigned)(3 * sw >> 2);
                const uint32_t u0 = __ldg(r0), u1 = __ldg(r0 + 1), u2 = __ldg(r0 + 2);
                const uint32_t v0 = __ldg(r1), v1 = __ldg(r1 + 1), v2 = __ldg(r1 + 2);
                const uint32_t ul = __funnelshift_r(u0, u1, sh8), uh = __funnelshift_r(u1, u2, sh8);
                const uint32_t vl = __funnelshift_r(v0, v1, sh8), vh = __funnelshift_r(v1, v2, sh8);
                c0 = ul & 0x00FFFFFFu; c1 = __funnelshift_r(ul, uh, 24) & 0x00FFFFFFu;
                c2 = vl & 0x00FFFFFFu; c3 = __funnelshift_r(vl, vh, 24) & 0x00FFFFFFu;
            }
            uint32_t ta_unused;
            vm_blend_lanes(c0, c1, c2, c3, (uint32_t)fx, (uint32_t)fy, bgr, ta_unused);
        } else {
            const int ix = vm_sat_s16(SX >> 5), iy = vm_sat_s16(SY >> 5);
            uint32_t c[4];
            double al[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int Y = iy + (t >> 1), X = ix + (t & 1);
                const int sy = Y - tv, sx = X - tu;
                const bool ok = live && (unsigned)Y < (unsigned)h && (unsigned)X < (unsigned)w && (unsigned)sy < (unsigned)sh && (unsigned)sx < (unsigned)sw;
                c[t] = 0u; al[t] = 0.0;
                if (ok) {
                    const int64_t q = fsrc + (int64_t)sy * sw + sx;
                    if (FG) {
                        const uint2 e = __ldg(reinterpret_cast<const uint2 *>(src_all) + q);
                        c[t] = e.x; al[t] = alpha64 ? __ldg(alpha64 + q) : (double)__uint_as_float(e.y);
                    } else if (W4) {
                        c[t] = vm_ld3(reinterpret_cast<const uint32_t *>(src_all), q * 3, src_last_word);
                    } else {
                        const uint8_t *p = reinterpret_cast<const uint8_t *>(src_all) + q * 3;
                        c[t] = (uint32_t)__ldg(p) | ((uint32_t)__ldg(p + 1) << 8) | ((uint32_t)__ldg(p + 2) << 16);
                    }
                }
            }
            bgr = 0u;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                bgr |= (uint32_t)VmTap<uint8_t>::blend((c[0] >> (8 * k)) & 255, (c[1] >> (8 * k)) & 255, (c[2] >> (8 * k)) & 255, (c[3] >> (8 * k)) & 255, fx, fy) << (8 * k);
            if (FG) {
                // float32 alpha taps: float32 products of exact weights, summed in tap order - the same expression as the
                // interior path, so a pixel's value does not depend on which path took it
                if (alpha64) ad = VmTap<double>::blend(al[0], al[1], al[2], al[3], fx, fy);
                else af = VmTap<float>::blend((float)al[0], (float)al[1], (float)al[2], (float)al[3], fx, fy);
            }
        }
        const uint32_t o = vm_illum_px((int)(bgr & 255u), (int)((bgr >> 8) & 255u), (int)(bgr >> 16), sdiv, hdiv, slut, body);
        const int64_t p = ((int64_t)frame * h + y) * w + x;
        if (W4) {
            // lanes 4g .. 4g+3 hold pixels o0 .. o3 = 12 bytes = words {o0 | o1 << 24, o1 >> 8 | o2 << 16, o2 >> 16 | o3 << 8}
            const uint32_t nxt = __shfl_down_sync(0xffffffffu, o, 1);
            const int k = threadIdx.x & 3;
            if (live && k < 3) {
                const uint32_t word = (o >> (8 * k)) | (nxt << (24 - 8 * k));
                reinterpret_cast<uint32_t *>(out_bgr)[((p - k) * 3 >> 2) + k] = word;
            }
        } else {
            out_bgr[p * 3] = (uint8_t)o; out_bgr[p * 3 + 1] = (uint8_t)(o >> 8); out_bgr[p * 3 + 2] = (uint8_t)(o >> 16);
        }
        if (FG && live) {
            if (out_alpha64) out_alpha64[p] = ad; else out_alpha[p] = af;
        }
    }
}

// mode 0: background (src = (n,h,w,3) uint8, out_alpha unused); mode 1: foreground (src = vm_aug_tps intermediate).
// params: device array of n {double M[6]; int tu, tv}; luts: device (n, 256) uint8 S/V tables.
extern "C" int vm_aug_affine(int mode, const void *src, const double *alpha64, const void *params, const uint8_t *luts, int n, int h,
                             int w, uint8_t *out_bgr, float *out_alpha, double *out_alpha64, int hsv_vec, void *stream) {
    VM_REQUIRE(src && params && luts && out_bgr && (mode == 0 || out_alpha || out_alpha64), "null pointer");
    VM_REQUIRE(hsv_vec >= 1 && hsv_vec <= 1024, "hsv_vec out of range");
    VM_REQUIRE(mode == 0 || (alpha64 != nullptr) == (out_alpha64 != nullptr), "float64 alpha needs both the float64 source plane and the float64 output");
    VM_REQUIRE(n >= 0 && n < 65536 && h >= 1 && h < 65536 && w >= 1, "bad size");
    if (n == 0) return VM_OK;
    const dim3 grid((w + 255) / 256, (h + VA_AFF_ROWS - 1) / VA_AFF_ROWS, n);
    cudaStream_t st = (cudaStream_t)stream;
    {
        static std::mutex mu;
        static bool done[64];
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { vm_set_error("vm_aug_affine: cudaGetDevice failed"); return VM_ERR_CUDA; }
        std::lock_guard<std::mutex> lk(mu);
        if (!done[dev]) {                                                  // division tables: once per device (stream-ordered before the first use)
            k_va_tables<<<1, 256, 0, st>>>();
            if (cudaStreamSynchronize(st) != cudaSuccess) { vm_set_error("vm_aug_affine: table init failed"); return VM_ERR_CUDA; }
            done[dev] = true;
        }
    }
    // 32-bit colour accesses need 4-byte aligned planes and rows: w % 4 == 0 makes every row start a multiple of 12 bytes
    const bool w4 = (w & 3) == 0 && vm_aligned(out_bgr, 4) && (mode == 1 || vm_aligned(src, 4));
    const int64_t last_word = mode == 0 ? ((int64_t)n * h * w * 3) / 4 - 1 : 0;
    if (mode == 1) {
        if (w4) k_aug_affine<true, true><<<grid, 256, 0, st>>>(src, alpha64, (const VmAugParams *)params, luts, h, w, out_bgr, out_alpha, out_alpha64, hsv_vec, last_word);
        else k_aug_affine<true, false><<<grid, 256, 0, st>>>(src, alpha64, (const VmAugParams *)params, luts, h, w, out_bgr, out_alpha, out_alpha64, hsv_vec, last_word);
    } else {
        if (w4) k_aug_affine<false, true><<<grid, 256, 0, st>>>(src, nullptr, (const VmAugParams *)params, luts, h, w, out_bgr, nullptr, nullptr, hsv_vec, last_word);
        else k_aug_affine<false, false><<<grid, 256, 0, st>>>(src, nullptr, (const VmAugParams *)params, luts, h, w, out_bgr, nullptr, nullptr, hsv_vec, last_word);
    }
    return vm_check_launch("vm_aug_affine");
}
