// Shared device helpers for libvm_sm100a.so (compiled with -fmad=false: every fused
// multiply-add in these kernels is an explicit fma()/__fmaf_rn so that the float64 paths
// keep the reference's operation order bit for bit).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include "../../include/vm_b200.h"

void vm_set_error(const char *fmt, ...);
int  vm_check_launch(const char *what);

#define VM_REQUIRE(cond, msg)                                   \
    do { if (!(cond)) { vm_set_error("%s: %s", __func__, msg); return VM_ERR_ARG; } } while (0)

static inline bool vm_aligned(const void *p, uintptr_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }   // nullptr counts as aligned

static inline unsigned vm_blocks(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    return (unsigned)(b < 1 ? 1 : b);
}

// ---------------------------------------------------------------------------------------
// cv2.remap / warpAffine fixed-point tap rule
// ---------------------------------------------------------------------------------------

// cvRound(m * 32) with the x86 cvtps2dq "integer indefinite" for NaN / out-of-range.
__device__ __forceinline__ int vm_cvround_x32(float m) {
    float v = __fmul_rn(m, 32.0f);
    if (!(fabsf(v) < 2147483648.0f)) return INT_MIN;
    return __float2int_rn(v);
}

// cvRound(double) -> int32 (cvtsd2si semantics)
__device__ __forceinline__ int vm_cvround_f64(double v) {
    if (!(fabs(v) < 2147483648.0)) return INT_MIN;
    return __double2int_rn(v);
}

__device__ __forceinline__ int vm_sat_s16(int v) { return max(-32768, min(32767, v)); }

// flow.py:13-17 : map = float32(int64 grid + float32 flow)  ==  fp32 rn(float(j) + dx)
__device__ __forceinline__ float vm_map_coord(int j, float d) { return __fadd_rn((float)j, d); }

template <typename T> struct VmTap;   // per-dtype bilinear accumulation

template <> struct VmTap<uint8_t> {
    // (sum S*w + 16384) >> 15 with w = wx*wy*32  ==  (sum S*wx*wy + 512) >> 10
    static __device__ __forceinline__ uint8_t blend(int s00, int s01, int s10, int s11, int fx, int fy) {
        int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
        return (uint8_t)((s00 * w00 + s01 * w01 + s10 * w10 + s11 * w11 + 512) >> 10);
    }
};
template <> struct VmTap<float> {
    static __device__ __forceinline__ float blend(float s00, float s01, float s10, float s11, int fx, int fy) {
        float ax = (float)fx * 0.03125f, ay = (float)fy * 0.03125f;           // exact
        float w00 = __fmul_rn(1.f - ax, 1.f - ay), w01 = __fmul_rn(ax, 1.f - ay);
        float w10 = __fmul_rn(1.f - ax, ay), w11 = __fmul_rn(ax, ay);          // exact products
        float acc = __fmul_rn(s00, w00);
        acc = __fadd_rn(acc, __fmul_rn(s01, w01));
        acc = __fadd_rn(acc, __fmul_rn(s10, w10));
        acc = __fadd_rn(acc, __fmul_rn(s11, w11));
        return acc;
    }
};
template <> struct VmTap<double> {
    static __device__ __forceinline__ double blend(double s00, double s01, double s10, double s11, int fx, int fy) {
        float ax = (float)fx * 0.03125f, ay = (float)fy * 0.03125f;
        double w00 = (double)__fmul_rn(1.f - ax, 1.f - ay), w01 = (double)__fmul_rn(ax, 1.f - ay);
        double w10 = (double)__fmul_rn(1.f - ax, ay), w11 = (double)__fmul_rn(ax, ay);
        double acc = __dmul_rn(s00, w00);
        acc = __dadd_rn(acc, __dmul_rn(s01, w01));
        acc = __dadd_rn(acc, __dmul_rn(s10, w10));
        acc = __dadd_rn(acc, __dmul_rn(s11, w11));
        return acc;
    }
};

// Bilinear sample of a (H, W, C) image at fixed-point position (SX, SY) in 1/32 px; taps
// outside the image read 0 (BORDER_CONSTANT).  Writes C values to dst.
template <typename T, int C>
__device__ __forceinline__ void vm_sample_fixed(const T *__restrict__ src, int H, int W,
                                                int SX, int SY, T *__restrict__ dst) {
    const int ix = vm_sat_s16(SX >> 5), iy = vm_sat_s16(SY >> 5);
    const int fx = SX & 31, fy = SY & 31;
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const T *r0 = src + ((int64_t)iy * W + ix) * C;
    const T *r1 = r0 + (int64_t)W * C;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        T s00 = (y0 && x0) ? __ldg(r0 + c) : T(0);
        T s01 = (y0 && x1) ? __ldg(r0 + C + c) : T(0);
        T s10 = (y1 && x0) ? __ldg(r1 + c) : T(0);
        T s11 = (y1 && x1) ? __ldg(r1 + C + c) : T(0);
        dst[c] = VmTap<T>::blend(s00, s01, s10, s11, fx, fy);
    }
}

// ---------------------------------------------------------------------------------------
// BGRA foreground: flow-warped pixel = {B, G, R bit-exact uint8, alpha numerator TA}
// where warp_img(A/255.)[q] = TA / (1024 * 255) exactly (integer weights wx*wy sum to 1024).
// ---------------------------------------------------------------------------------------
#define VM_ALPHA_DEN 261120.0   /* 1024 * 255 */

struct VmWarped { uint32_t bgr; uint32_t ta; };   // bgr: B | G<<8 | R<<16

__device__ __forceinline__ uint32_t vm_ldg_bgra(const uint8_t *__restrict__ fg, int64_t px) {
    return __ldg(reinterpret_cast<const uint32_t *>(fg) + px);
}

// 4 BGRA taps blended with integer weights.  s?? are packed BGRA words.
__device__ __forceinline__ VmWarped vm_blend_bgra(uint32_t s00, uint32_t s01, uint32_t s10,
                                                  uint32_t s11, int fx, int fy) {
    const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
    VmWarped o;
    uint32_t b = ((s00 & 255u) * w00 + (s01 & 255u) * w01 + (s10 & 255u) * w10 + (s11 & 255u) * w11 + 512u) >> 10;
    uint32_t g = (((s00 >> 8) & 255u) * w00 + ((s01 >> 8) & 255u) * w01 + ((s10 >> 8) & 255u) * w10 +
                  ((s11 >> 8) & 255u) * w11 + 512u) >> 10;
    uint32_t r = (((s00 >> 16) & 255u) * w00 + ((s01 >> 16) & 255u) * w01 + ((s10 >> 16) & 255u) * w10 +
                  ((s11 >> 16) & 255u) * w11 + 512u) >> 10;
    o.ta = (s00 >> 24) * w00 + (s01 >> 24) * w01 + (s10 >> 24) * w10 + (s11 >> 24) * w11;
    o.bgr = b | (g << 8) | (r << 16);
    return o;
}

// warp_bgr / warp_img at integer output position (i, j) given its backward flow vector.
__device__ __forceinline__ VmWarped vm_flow_warp_bgra(const uint8_t *__restrict__ fg, int H, int W,
                                                      int i, int j, float2 fb) {
    const int SX = vm_cvround_x32(vm_map_coord(j, fb.x));
    const int SY = vm_cvround_x32(vm_map_coord(i, fb.y));
    const int ix = vm_sat_s16(SX >> 5), iy = vm_sat_s16(SY >> 5);
    const int fx = SX & 31, fy = SY & 31;
    const int64_t base = (int64_t)iy * W + ix;
    uint32_t s00, s01, s10, s11;
    if ((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1)) {
        s00 = vm_ldg_bgra(fg, base);     s01 = vm_ldg_bgra(fg, base + 1);
        s10 = vm_ldg_bgra(fg, base + W); s11 = vm_ldg_bgra(fg, base + W + 1);
    } else {
        const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
        const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
        s00 = (y0 && x0) ? vm_ldg_bgra(fg, base) : 0u;
        s01 = (y0 && x1) ? vm_ldg_bgra(fg, base + 1) : 0u;
        s10 = (y1 && x0) ? vm_ldg_bgra(fg, base + W) : 0u;
        s11 = (y1 && x1) ? vm_ldg_bgra(fg, base + W + 1) : 0u;
    }
    return vm_blend_bgra(s00, s01, s10, s11, fx, fy);
}

// flow.py:41-48 at pixel (i, j).  Returns 1 if err > 15 (alpha must be zeroed); flags:
// bit0 = IndexError condition, bit1 = NaN/inf condition (reference would raise).
__device__ __forceinline__ int vm_consistency(const float2 *__restrict__ forward, int H, int W,
                                              int i, int j, float2 fb, int &flags) {
    const float a = __fadd_rn(fb.x, (float)j), b = __fadd_rn(fb.y, (float)i);
    if (!isfinite(a) || !isfinite(b)) { flags |= 2; return 1; }      // int(nan/inf) raises
    // Python ints do not overflow: min() clamps the top; anything below -W / -H cannot be
    // wrapped and raises IndexError.
    const int j0 = (a >= (float)W) ? W - 1 : ((a <= -1.0e9f) ? -W - 1 : min(__float2int_rz(a), W - 1));
    const int i0 = (b >= (float)H) ? H - 1 : ((b <= -1.0e9f) ? -H - 1 : min(__float2int_rz(b), H - 1));
    const int jw = j0 < 0 ? j0 + W : j0, iw = i0 < 0 ? i0 + H : i0;
    if (jw < 0 || iw < 0) { flags |= 1; return 1; }
    const float2 ff = __ldg(forward + (int64_t)iw * W + jw);
    const float c = __fadd_rn(ff.x, (float)j0), d = __fadd_rn(ff.y, (float)i0);
    if (!isfinite(c) || !isfinite(d)) { flags |= 2; return 1; }
    // |c| may exceed int32: Python ints do not overflow, min() clamps the top, the bottom
    // only grows the error -> masked either way.
    const float cc = fmaxf(c, -1.0e9f), dd = fmaxf(d, -1.0e9f);
    const int j1 = (cc >= (float)W) ? W - 1 : min(__float2int_rz(cc), W - 1);
    const int i1 = (dd >= (float)H) ? H - 1 : min(__float2int_rz(dd), H - 1);
    const long long di = (long long)i1 - i, dj = (long long)j1 - j;
    return (di * di + dj * dj > 225) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------
// scipy.ndimage.map_coordinates(order=1, mode='constant', cval=0)
// ---------------------------------------------------------------------------------------
struct VmBilin64 { double a0, a1, b0, b1; int i0, i1, j0, j1; bool inside; };

__device__ __forceinline__ VmBilin64 vm_mapcoord_setup(double t0, double t1, int H, int W) {
    VmBilin64 s;
    s.inside = (t0 >= 0.0) && (t0 <= (double)(H - 1)) && (t1 >= 0.0) && (t1 <= (double)(W - 1));
    const double f0 = floor(t0), f1 = floor(t1);
    s.i0 = s.inside ? (int)f0 : 0;
    s.j0 = s.inside ? (int)f1 : 0;
    const double a = t0 - f0, b = t1 - f1;
    s.a0 = 1.0 - a; s.a1 = 1.0 - s.a0;          // scipy: w1 = 1 - w0
    s.b0 = 1.0 - b; s.b1 = 1.0 - s.b0;
    s.i1 = min(s.i0 + 1, H - 1);
    s.j1 = min(s.j0 + 1, W - 1);
    return s;
}

__device__ __forceinline__ double vm_mapcoord_blend(const VmBilin64 &s, double s00, double s01,
                                                    double s10, double s11) {
    double v = __dmul_rn(__dmul_rn(s00, s.a0), s.b0);
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(s01, s.a0), s.b1));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(s10, s.a1), s.b0));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(s11, s.a1), s.b1));
    return v;
}

// scipy order 0 (nearest): index floor(t + 1/2) of an in-range coordinate (0 <= t <= n - 1)
__device__ __forceinline__ int vm_nearest(double t) { return (int)floor(t + 0.5); }

__device__ __forceinline__ uint8_t vm_round_half_up_u8(double v) {
    const double r = floor(v + 0.5);
    return (uint8_t)(r < 0.0 ? 0 : (r > 255.0 ? 255 : (int)r));
}

// ---------------------------------------------------------------------------------------
// bilinear up-sampling of the coarse TPS transform (tps.py:55-74)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double vm_upsample_exact(double t00, double t01, double t10, double t11,
                                                    double xf, double yf) {
    const double x1 = 1.0 - xf, y1 = 1.0 - yf;
    double v = __dmul_rn(__dmul_rn(t00, x1), y1);
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(t01, x1), yf));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(t10, xf), y1));
    v = __dadd_rn(v, __dmul_rn(__dmul_rn(t11, xf), yf));
    return v;
}


// ---------------------------------------------------------------------------------------
// Instruction-lean fast paths used by the fused kernels.  Each one is exact (same integers
// as the generic functions above) inside a guarded range and reports when it cannot be used.
// ---------------------------------------------------------------------------------------

// rne(32*m) for |m| < 65536 without a conversion instruction: (32m + 1.5*2^23) keeps the
// integer in the mantissa; 32m is exact so this is a single rounding = cvRound(m*32).
__device__ __forceinline__ int vm_fix5_fast(float m) {
    return __float_as_int(__fmaf_rn(m, 32.f, 12582912.f)) - 0x4B400000;
}

// uint8 (already isolated in the low byte of `u`, value < 2^23) -> float, no conversion pipe.
__device__ __forceinline__ float vm_u2f(uint32_t u) { return __uint_as_float(u | 0x4B000000u) - 8388608.f; }
__device__ __forceinline__ float vm_byte2f(uint32_t word, int byte) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650 + byte)) - 8388608.f;
}

struct VmFlowPx { uint32_t bgr; uint32_t ta; int masked; };

// The four channel blends of vm_blend_bgra as two 64-bit multiply-add chains: the tap words are split
// into {B, R} and {G, A} with one channel in byte 0 and the other in byte 3, the weights are scaled by 64
// (sum 2^16), so a channel's sum (+ 2^15 for the rounding) stays below 2^24 and never reaches its
// neighbour: lane 0 = (sum S w + 512) << 6, i.e. the blended byte sits in bits 16..23; the alpha lane
// (no rounding term) holds TA << 6 from bit 24 on.  Same integers as vm_blend_bgra, 19 instead of 40
// instructions.
__device__ __forceinline__ void vm_blend_lanes(uint32_t s00, uint32_t s01, uint32_t s10, uint32_t s11, uint32_t fx,
                                               uint32_t fy, uint32_t &bgr, uint32_t &ta) {
    const uint32_t gx = 32u - fx, fy6 = fy << 6, gy6 = 2048u - fy6;
    const uint32_t w00 = gx * gy6, w01 = fx * gy6, w10 = gx * fy6, w11 = fx * fy6;
    unsigned long long br = 0x0000008000008000ull, ga = 0x0000000000008000ull;
    br += (unsigned long long)__byte_perm(s00, 0, 0x2440) * w00; ga += (unsigned long long)__byte_perm(s00, 0, 0x3441) * w00;
    br += (unsigned long long)__byte_perm(s01, 0, 0x2440) * w01; ga += (unsigned long long)__byte_perm(s01, 0, 0x3441) * w01;
    br += (unsigned long long)__byte_perm(s10, 0, 0x2440) * w10; ga += (unsigned long long)__byte_perm(s10, 0, 0x3441) * w10;
    br += (unsigned long long)__byte_perm(s11, 0, 0x2440) * w11; ga += (unsigned long long)__byte_perm(s11, 0, 0x3441) * w11;
    const uint32_t brl = (uint32_t)br, brh = (uint32_t)(br >> 32), gal = (uint32_t)ga, gah = (uint32_t)(ga >> 32);
    // B = bits 16..23 of br, R = bits 40..47 (byte 1 of the high word, whose byte 3 is zero), G = bits 16..23 of ga
    bgr = __byte_perm(__byte_perm(brl, brh, 0x7572), gal, 0x3260);
    ta = __funnelshift_r(gal, gah, 30);
}

// warp_bgr / warp_img / correct_alpha for the pixel (i, j) of a BGRA frame: fast path for
// in-range coordinates, otherwise the generic routines.  fi/fj = (float)i / (float)j.
template <bool MASK>
__device__ __forceinline__ VmFlowPx vm_flow_px(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd,
                                               int H, int W, int i, int j, float fi, float fj, float2 fb,
                                               int &flags) {
    VmFlowPx o;
    const float mx = __fadd_rn(fj, fb.x), my = __fadd_rn(fi, fb.y);
    // fmaxf() returns the non-NaN operand, so each component is tested on its own: false for NaN
    if ((fabsf(mx) < 60000.f) & (fabsf(my) < 60000.f)) {
        const int SX = vm_fix5_fast(mx), SY = vm_fix5_fast(my);
        const int ix = SX >> 5, iy = SY >> 5, fx = SX & 31, fy = SY & 31;
        uint32_t s00, s01, s10, s11;
        const int base = (int)((unsigned)iy * (unsigned)W + (unsigned)ix);
        if ((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1)) {
            const uint32_t *p = fg32 + (unsigned)base, *q = p + (unsigned)W;
            s00 = __ldg(p); s01 = __ldg(p + 1); s10 = __ldg(q); s11 = __ldg(q + 1);
        } else {
            const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
            const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
            s00 = (y0 && x0) ? __ldg(fg32 + base) : 0u;
            s01 = (y0 && x1) ? __ldg(fg32 + base + 1) : 0u;
            s10 = (y1 && x0) ? __ldg(fg32 + base + W) : 0u;
            s11 = (y1 && x1) ? __ldg(fg32 + base + W + 1) : 0u;
        }
        vm_blend_lanes(s00, s01, s10, s11, (uint32_t)fx, (uint32_t)fy, o.bgr, o.ta);
        o.masked = 0;
        if (MASK) {
            // flow.py:44: a = bx + j (same float as mx), trunc toward zero; fast when it lands
            // inside the frame (no clamp, no wrap)
            const int j0 = __float2int_rz(mx), i0 = __float2int_rz(my);
            if ((unsigned)j0 < (unsigned)W && (unsigned)i0 < (unsigned)H) {
                const float2 ff = __ldg(fwd + (unsigned)(i0 * W + j0));
                const float c = __fadd_rn(ff.x, (float)j0), d = __fadd_rn(ff.y, (float)i0);
                if ((fabsf(c) < 3.0e38f) & (fabsf(d) < 3.0e38f)) {              // finite, and false for NaN
                    // min(trunc(c), W-1) - j, in float: exact for |.| < 2^24, monotone beyond
                    const float dj = __fadd_rn(fminf(truncf(c), (float)(W - 1)), -fj);
                    const float di = __fadd_rn(fminf(truncf(d), (float)(H - 1)), -fi);
                    o.masked = __fmaf_rn(dj, dj, __fmul_rn(di, di)) > 225.f;
                } else {
                    flags |= 2; o.masked = 1;
                }
            } else {
                o.masked = vm_consistency(fwd, H, W, i, j, fb, flags);
            }
        }
    } else {
        const VmWarped wv = vm_flow_warp_bgra(reinterpret_cast<const uint8_t *>(fg32), H, W, i, j, fb);
        o.bgr = wv.bgr; o.ta = wv.ta;
        o.masked = MASK ? vm_consistency(fwd, H, W, i, j, fb, flags) : 0;
    }
    return o;
}

// G horizontally adjacent pixels at once, in two phases: every tap load and every forward-flow load
// of the group is issued (from clamped, always valid addresses) before the first value is used, so the
// 20 gathers of a thread are in flight together instead of as eight dependent round trips.  Pixels that
// leave the fast path (frame border, out-of-range or NaN coordinates) are redone by the same code as
// vm_flow_px afterwards; results are identical to four vm_flow_px calls.
// gather addresses of one pixel (always valid): the four taps p, p + dx, q, q + dx and the forward-flow vector pf.
// Pixels off the fast path (frame border, out-of-range or NaN coordinates) get pixel 0; vm_flow_px_finish redoes them.
// (dx as an immediate 1 saves four instructions and costs a spill at the 48-register cap: 0.877 ms against 0.813 ms.)
template <bool MASK>
__device__ __forceinline__ void vm_flow_px_addr(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd, int H, int W,
                                                float fi, float fj, float2 fb, const uint32_t *&p, const uint32_t *&q,
                                                int &dx, const float2 *&pf) {
    const float mx = __fadd_rn(fj, fb.x), my = __fadd_rn(fi, fb.y);
    const bool inr = (fabsf(mx) < 60000.f) & (fabsf(my) < 60000.f);
    const int ix = vm_fix5_fast(mx) >> 5, iy = vm_fix5_fast(my) >> 5;
    const bool inner = inr && (unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1);
    p = fg32 + (inner ? (unsigned)iy * (unsigned)W + (unsigned)ix : 0u);
    q = p + (inner ? (unsigned)W : 0u);
    dx = inner ? 1 : 0;
    if (MASK) {
        const int j0 = __float2int_rz(mx), i0 = __float2int_rz(my);
        const bool fwd_ok = inr && (unsigned)j0 < (unsigned)W && (unsigned)i0 < (unsigned)H;
        pf = fwd + (fwd_ok ? (unsigned)(i0 * W + j0) : 0u);
    }
}

// the arithmetic of vm_flow_px on values gathered from the addresses of vm_flow_px_addr
template <bool MASK>
__device__ __forceinline__ VmFlowPx vm_flow_px_finish(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd,
                                                      int H, int W, int i, int j, float fi, float fj, float2 fb,
                                                      uint32_t s00, uint32_t s01, uint32_t s10, uint32_t s11, float2 fv,
                                                      int &flags) {
    VmFlowPx o;
    const float mx = __fadd_rn(fj, fb.x), my = __fadd_rn(fi, fb.y);
    if ((fabsf(mx) < 60000.f) & (fabsf(my) < 60000.f)) {
        const int SX = vm_fix5_fast(mx), SY = vm_fix5_fast(my);
        const int ix = SX >> 5, iy = SY >> 5, fx = SX & 31, fy = SY & 31;
        if (!((unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1))) {     // frame border: taps outside read 0
            const int base = (int)((unsigned)iy * (unsigned)W + (unsigned)ix);
            const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
            const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
            s00 = (y0 && x0) ? __ldg(fg32 + base) : 0u;
            s01 = (y0 && x1) ? __ldg(fg32 + base + 1) : 0u;
            s10 = (y1 && x0) ? __ldg(fg32 + base + W) : 0u;
            s11 = (y1 && x1) ? __ldg(fg32 + base + W + 1) : 0u;
        }
        vm_blend_lanes(s00, s01, s10, s11, (uint32_t)fx, (uint32_t)fy, o.bgr, o.ta);
        o.masked = 0;
        if (MASK) {
            const int j0 = __float2int_rz(mx), i0 = __float2int_rz(my);
            if ((unsigned)j0 < (unsigned)W && (unsigned)i0 < (unsigned)H) {
                const float c = __fadd_rn(fv.x, (float)j0), d = __fadd_rn(fv.y, (float)i0);
                if ((fabsf(c) < 3.0e38f) & (fabsf(d) < 3.0e38f)) {
                    const float dj = __fadd_rn(fminf(truncf(c), (float)(W - 1)), -fj);
                    const float di = __fadd_rn(fminf(truncf(d), (float)(H - 1)), -fi);
                    o.masked = __fmaf_rn(dj, dj, __fmul_rn(di, di)) > 225.f;
                } else {
                    flags |= 2; o.masked = 1;
                }
            } else {
                o.masked = vm_consistency(fwd, H, W, i, j, fb, flags);
            }
        }
    } else {
        const VmWarped wv = vm_flow_warp_bgra(reinterpret_cast<const uint8_t *>(fg32), H, W, i, j, fb);
        o.bgr = wv.bgr; o.ta = wv.ta;
        o.masked = MASK ? vm_consistency(fwd, H, W, i, j, fb, flags) : 0;
    }
    return o;
}

template <bool MASK, int G>
__device__ __forceinline__ void vm_flow_pxn(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd,
                                            int H, int W, int i, int j, float fi, float fj, const float2 *__restrict__ fb,
                                            VmFlowPx *__restrict__ o, int &flags) {
    uint32_t s[G][4];
    float2 fv[G];
    // phase 1: addresses and loads only (the geometry is recomputed in phase 2: a few ALU instructions
    // are cheaper than carrying it in registers across the outstanding loads)
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const uint32_t *p, *q;
        const float2 *pf = nullptr;
        int dx;
        vm_flow_px_addr<MASK>(fg32, fwd, H, W, fi, fj + (float)k, fb[k], p, q, dx, pf);
        s[k][0] = __ldg(p); s[k][1] = __ldg(p + dx); s[k][2] = __ldg(q); s[k][3] = __ldg(q + dx);
        fv[k] = MASK ? __ldg(pf) : make_float2(0.f, 0.f);
    }
    // phase 2: the arithmetic of vm_flow_px on the loaded values
#pragma unroll
    for (int k = 0; k < G; ++k)
        o[k] = vm_flow_px_finish<MASK>(fg32, fwd, H, W, i, j + k, fi, fj + (float)k, fb[k], s[k][0], s[k][1], s[k][2], s[k][3], fv[k], flags);
}

// alpha numerator ta (0..261120) -> signed-complement float: +alpha when alpha <= 1/2,
// -(1 - alpha) otherwise.  Both alpha and 1 - alpha are then recoverable with ~1e-7 RELATIVE
// error, which the composite needs (|out - ref| <= 1e-5 rel even where (1-alpha)*B dominates).
#define VM_ALPHA_INV 3.82965686274509803e-06f      /* 1 / 261120 */
__device__ __forceinline__ uint32_t vm_alpha_code(uint32_t ta) {
    const uint32_t nta = 261120u - ta;
    const float s = (float)min(ta, nta) * VM_ALPHA_INV;
    return __float_as_uint(s) | ((ta > nta) ? 0x80000000u : 0u);
}
__device__ __forceinline__ void vm_alpha_decode(uint32_t code, float &a, float &na) {
    const float s = fabsf(__uint_as_float(code)), t = 1.f - s;
    const bool big = (int)code < 0;
    a = big ? t : s;
    na = big ? s : t;
}

// source pixel of the TPS resampling evaluated from the original inputs (exact per-pixel path):
// FLOW: the flow-warped, consistency-masked pixel; otherwise the BGRA pixel itself.
struct VmSrcPx { double b, g, r, a; };

template <bool FLOW>
__device__ __forceinline__ VmSrcPx vm_src_px(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                             const float2 *__restrict__ fwd, int H, int W, int qi, int qj,
                                             int &flags) {
    VmSrcPx o;
    if (FLOW) {
        const float2 fb = __ldg(bwd + (int64_t)qi * W + qj);
        const VmWarped wv = vm_flow_warp_bgra(fg, H, W, qi, qj, fb);
        const int m = fwd ? vm_consistency(fwd, H, W, qi, qj, fb, flags) : 0;
        o.b = (double)(wv.bgr & 255u); o.g = (double)((wv.bgr >> 8) & 255u); o.r = (double)((wv.bgr >> 16) & 255u);
        o.a = m ? 0.0 : (double)wv.ta / VM_ALPHA_DEN;
    } else {
        const uint32_t s = vm_ldg_bgra(fg, (int64_t)qi * W + qj);
        o.b = (double)(s & 255u); o.g = (double)((s >> 8) & 255u); o.r = (double)((s >> 16) & 255u);
        o.a = (double)(s >> 24) / 255.0;                       // reader.py:16
    }
    return o;
}

__device__ __forceinline__ vm_axis_entry vm_ld_axis(const vm_axis_entry *p) {
    const int4 v = __ldg(reinterpret_cast<const int4 *>(p));
    vm_axis_entry e;
    e.frac = __hiloint2double(v.y, v.x);
    e.i0 = v.z; e.i1 = v.w;
    return e;
}


// ---------------------------------------------------------------------------------------
// fused flow stage (C2 / stage A of C4) - shared by vm_flow.cu and the dependency-driven kernel of vm_lean.cu
// ---------------------------------------------------------------------------------------
// Gathers of a pixel: vm_flow_pxn<., 1> issues the four taps AND the forward-flow vector (whose address depends on the
// coordinate only) before the first use, so a pixel costs one memory round trip instead of two.  Measured on the
// stage-A launch (1080p x 64, round 2): 0.813 ms against 0.878 ms pixel by pixel (vm_flow_px), 0.87 ms with the gathers
// of two pixels grouped (spills at 48 registers), 0.90-1.06 ms at 4 CTAs per SM; the lane-packed blend and the opaque frame
// pointers are worth 1 % each on their own, 4 % together with the hoisted forward-flow load.
#define C2_TW 128
#define C2_TH 8                    /* rows in flight per CTA (one per warp) */
#define C2_ROWS 40                 /* rows per CTA: each warp walks C2_ROWS / C2_TH of them */

// alpha = ta / 261120 as float32 with <= 1.2e-7 relative error and exact 0 / 1 end points
__device__ __forceinline__ float vm_alpha_f32(uint32_t ta) {
    const uint32_t nta = 261120u - ta;
    const float s = (float)min(ta, nta) * VM_ALPHA_INV;
    return (ta > nta) ? 1.f - s : s;
}

// PACKED = 0: out_bgr (n,h,w,3) uint8 + out_alpha (n,h,w) float32 (the C2 result);
// PACKED = 1: out_bgr is a (n,h,w) uint2 array {B|G<<8|R<<16, alpha code} - the stage-A
//             intermediate of the split C4 pipeline (see vm_tps.cu), out_alpha unused;
// PACKED = 2: same, with the raw alpha numerator {B|G<<8|R<<16, TA} (TA = 0 where masked) -
//             the intermediate of the lean pipeline (vm_lean.cu), alpha = TA / 261120.
template <int PACKED> __device__ __forceinline__ uint32_t vm_pack_alpha(const VmFlowPx &px) {
    return px.masked ? 0u : (PACKED == 2 ? px.ta : vm_alpha_code(px.ta));
}

// One flow-stage unit = 128 columns x C2_ROWS rows of frame `frame`, executed by 256 threads: warp k walks
// rows k, k + 8, ...; the two 16-byte flow loads of the next row are in flight while the current row's
// taps are gathered and blended.
template <bool HAS_FWD, int PACKED>
__device__ __forceinline__ void vm_flow_unit_t(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                             const float2 *__restrict__ fwd, int h, int w, int frame, int ty, int tx, int tid,
                                             uint8_t *__restrict__ out_bgr, float *__restrict__ out_alpha,
                                             int32_t *__restrict__ status) {
    const int ibeg = ty * C2_ROWS + (tid >> 5), iend = min((ty + 1) * C2_ROWS, h);
    const int j = tx * C2_TW + (tid & 31) * 4;
    if (ibeg >= iend || j >= w) return;
    const int64_t fbase = (int64_t)frame * h * w;
    const uint32_t *fg32 = reinterpret_cast<const uint32_t *>(fg) + fbase;
    const float2 *bf = bwd + fbase;
    const float2 *ff = HAS_FWD ? fwd + fbase : nullptr;
    // the frame pointers as opaque per-thread values: a gather address is then one 32 x 32 + 64 multiply-add
    // instead of (index + frame offset) -> 64-bit scale -> add, five instructions per address
    asm volatile("" : "+l"(fg32));
    if (HAS_FWD) asm volatile("" : "+l"(ff));
    const float fj = (float)j;
    int flags = 0;
    if (j + 3 < w && (w & 3) == 0) {
        float4 n01 = __ldcs(reinterpret_cast<const float4 *>(bf + ibeg * w + j));       // streamed once: evict first
        float4 n23 = __ldcs(reinterpret_cast<const float4 *>(bf + ibeg * w + j + 2));
        for (int i = ibeg; i < iend; i += C2_TH) {
            const float4 f01 = n01, f23 = n23;
            if (i + C2_TH < iend) {
                n01 = __ldcs(reinterpret_cast<const float4 *>(bf + (i + C2_TH) * w + j));
                n23 = __ldcs(reinterpret_cast<const float4 *>(bf + (i + C2_TH) * w + j + 2));
            }
            const int p = i * w + j;
            const float fi = (float)i;
            const float2 fl[4] = {{f01.x, f01.y}, {f01.z, f01.w}, {f23.x, f23.y}, {f23.z, f23.w}};
            VmFlowPx px[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                vm_flow_pxn<HAS_FWD, 1>(fg32, ff, h, w, i, j + k, fi, fj + (float)k, fl + k, px + k, flags);
            if (PACKED) {
                uint4 *op = reinterpret_cast<uint4 *>(reinterpret_cast<uint2 *>(out_bgr) + fbase + p);
                op[0] = make_uint4(px[0].bgr, vm_pack_alpha<PACKED>(px[0]), px[1].bgr, vm_pack_alpha<PACKED>(px[1]));
                op[1] = make_uint4(px[2].bgr, vm_pack_alpha<PACKED>(px[2]), px[3].bgr, vm_pack_alpha<PACKED>(px[3]));
            } else {
                float al[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) al[k] = px[k].masked ? 0.f : vm_alpha_f32(px[k].ta);
                // 12 bytes of BGR: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
                const uint32_t c0 = px[0].bgr, c1 = px[1].bgr, c2 = px[2].bgr, c3 = px[3].bgr;
                uint32_t *ob = reinterpret_cast<uint32_t *>(out_bgr + (fbase + p) * 3);
                __stcs(ob, c0 | (c1 << 24));
                __stcs(ob + 1, (c1 >> 8) | (c2 << 16));
                __stcs(ob + 2, (c2 >> 16) | (c3 << 8));
                __stcs(reinterpret_cast<float4 *>(out_alpha + fbase + p), make_float4(al[0], al[1], al[2], al[3]));
            }
        }
    } else {
        for (int i = ibeg; i < iend; i += C2_TH) {
            const int p = i * w + j;
            const float fi = (float)i;
            for (int k = 0; k < 4 && j + k < w; ++k) {
                const float2 f = __ldg(bf + p + k);
                const VmFlowPx px = vm_flow_px<HAS_FWD>(fg32, ff, h, w, i, j + k, fi, fj + (float)k, f, flags);
                if (PACKED) {
                    reinterpret_cast<uint2 *>(out_bgr)[fbase + p + k] = make_uint2(px.bgr, vm_pack_alpha<PACKED>(px));
                    continue;
                }
                uint8_t *ob = out_bgr + (fbase + p + k) * 3;
                ob[0] = (uint8_t)px.bgr; ob[1] = (uint8_t)(px.bgr >> 8); ob[2] = (uint8_t)(px.bgr >> 16);
                out_alpha[fbase + p + k] = px.masked ? 0.f : vm_alpha_f32(px.ta);
            }
        }
    }
    if (HAS_FWD && flags && status) {
        if (flags & 1) atomicAdd(status + VM_STATUS_INDEX_ERR, 1);
        if (flags & 2) atomicAdd(status + VM_STATUS_NAN_ERR, 1);
    }
}


template <bool HAS_FWD, int PACKED>
__device__ __forceinline__ void vm_flow_unit(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                             const float2 *__restrict__ fwd, int h, int w, int frame, int ty, int tx,
                                             uint8_t *__restrict__ out_bgr, float *__restrict__ out_alpha,
                                             int32_t *__restrict__ status) {
    vm_flow_unit_t<HAS_FWD, PACKED>(fg, bwd, fwd, h, w, frame, ty, tx, (int)threadIdx.x, out_bgr, out_alpha, status);
}

// packed {bgr, TA} output, explicit flat thread index (for CTAs that are not 256 x 1 shaped)
template <bool HAS_FWD>
__device__ __forceinline__ void vm_flow_unit_flat(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                                  const float2 *__restrict__ fwd, int h, int w, int frame, int ty, int tx,
                                                  int tid, uint2 *__restrict__ packed, int32_t *__restrict__ status) {
    vm_flow_unit_t<HAS_FWD, 2>(fg, bwd, fwd, h, w, frame, ty, tx, tid, reinterpret_cast<uint8_t *>(packed), nullptr, status);
}
