// Device helpers shared by the lean split pipeline (vm_lean.cu) and the single-pass fused kernel (vm_fuse.cu):
// the table-driven r^2 log r^2, the fixed-point resampling / composite of one pixel, bulk-copy wrappers.
#pragma once
#include "vm_common.cuh"
#include <math.h>

// ---------------------------------------------------------------------------------------
// log table: value x = 2^e * m, e in [VL_EMIN, VL_EMAX), m in [1 + k/2^B, 1 + (k+1)/2^B), B = VL_BITS;
// entry (e,k) = {RN(1/c_k) * 2^-e, e ln2 - log(RN(1/c_k))}, c_k = 1 + (k + 1/2)/2^B.
// t = fma(x, entry.x, -1) has |t| <= 2^-(VL_BITS+1);  log x = entry.y + t * q(t).
// ---------------------------------------------------------------------------------------
#define VL_EMIN (-6)
#define VL_EMAX 25
#define VL_BITS 6
#define VL_TAB_N ((VL_EMAX - VL_EMIN) << VL_BITS)
#define VL_TAB_BYTES (VL_TAB_N * 16)
#define VL_HI_MIN ((1023 + VL_EMIN) << 20)
#define VL_HI_MAX ((1023 + VL_EMAX) << 20)

// minimax fit of log1p(t)/t on |t| <= 2^-7 (max error of t*q(t): 2.4e-15 absolute).  64 entries per
// octave keep the table index of neighbouring coarse columns within one entry of each other for every
// control point further than ~370 px away, so a quarter-warp reads consecutive entries: few bank
// conflicts.  Measured alternatives: 256 entries + degree 3 (9 DP, bound by shared-memory bank conflicts),
// 32 entries + degree 5 (11 DP, 5 % slower than this one).
#define VL_Q4 0.2000108996732607315039
#define VL_Q3 (-0.2500127162639107632768)
#define VL_Q2 0.3333333331670152271109
#define VL_Q1 (-0.4999999998059626206361)
#define VL_Q0 1.0


#define VL_MAX_N 64

__device__ __forceinline__ uint32_t vl_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// r2 * log(r2) through the shared-memory table; r2 must lie in [2^VL_EMIN, 2^VL_EMAX)
__device__ __forceinline__ double vl_u_fast(double r2, uint32_t tab_adj) {
    const uint32_t addr = tab_adj + (((uint32_t)__double2hiint(r2) >> (20 - VL_BITS)) << 4);
    double ex, ey;
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(ex), "=d"(ey) : "r"(addr));
    const double t = fma(r2, ex, -1.0);
    double q = fma(t, VL_Q4, VL_Q3);
    q = fma(t, q, VL_Q2);
    q = fma(t, q, VL_Q1);
    q = fma(t, q, VL_Q0);
    return r2 * fma(t, q, ey);
}

// any r2 >= 0 (tps.py:78-82: U = 0 for r < 1e-100)
static __device__ __noinline__ double vl_u_any(double r2, uint32_t tab_adj) {
    const int hi = __double2hiint(r2);
    if (hi >= VL_HI_MIN && hi < VL_HI_MAX) return vl_u_fast(r2, tab_adj);
    if (r2 >= 1e-200) return r2 * log(r2);
    return (r2 == r2) ? 0.0 : r2;
}


#define VL_MAGIC 1572864.0                         // 1.5 * 2^20: ulp = 2^-32, exponent field 0x413
#define VL_MAGIC_HI 0x41380000

struct VlC { double x, y; };                       // column-interpolated coarse row {row coord, column coord}

template <int SRC> struct VlSrc;
template <> struct VlSrc<0> {                      // BGRA frame: alpha = A / 255
    typedef uint32_t elem;
    static constexpr double DEN = 255.0;
    static __device__ __forceinline__ uint2 ld(const elem *p) { const uint32_t s = __ldg(p); return make_uint2(s, s >> 24); }
    static __device__ __forceinline__ uint2 lds(const elem *p) { const uint32_t s = *p; return make_uint2(s, s >> 24); }
};
template <> struct VlSrc<1> {                      // packed {bgr, TA}: alpha = TA / 261120
    typedef uint2 elem;
    static constexpr double DEN = VM_ALPHA_DEN;
    static __device__ __forceinline__ uint2 ld(const elem *p) { return __ldg(p); }
    static __device__ __forceinline__ uint2 lds(const elem *p) { return *p; }
};

#define VL_NEAR_KNIFE_UNIT 65536                    // per-thread counter: low 16 bits = samples outside the source, high = near-knife samples

// exact per-pixel evaluation (frame borders, samples outside the source, undecidable roundings):
// scipy's float64 arithmetic on the fast path's coordinates.  `mask` bit c set =
// recompute colour c; bit 3 = the fast geometry did not apply: recompute everything.
// o.w / na = the fast path's alpha', 1 - alpha'.
template <int SRC>
__device__ __noinline__ float4 vl_exact_px(const typename VlSrc<SRC>::elem *__restrict__ src, double t0, double t1,
                                           int h, int w, const uint8_t *__restrict__ bp, float4 o, float na,
                                           unsigned mask, int *outside) {
    const float bgv[3] = {(float)__ldg(bp), (float)__ldg(bp + 1), (float)__ldg(bp + 2)};
    const VmBilin64 s = vm_mapcoord_setup(t0, t1, h, w);
    if (!s.inside) {
        (*outside)++;
        return make_float4(bgv[0], bgv[1], bgv[2], 0.f);
    }
    const uint2 e0 = VlSrc<SRC>::ld(src + ((int64_t)s.i0 * w + s.j0)), e1 = VlSrc<SRC>::ld(src + ((int64_t)s.i0 * w + s.j1));
    const uint2 e2 = VlSrc<SRC>::ld(src + ((int64_t)s.i1 * w + s.j0)), e3 = VlSrc<SRC>::ld(src + ((int64_t)s.i1 * w + s.j1));
    float a2 = o.w;
    if (mask & 8u) {
        const double a64 = vm_mapcoord_blend(s, (double)e0.y / VlSrc<SRC>::DEN, (double)e1.y / VlSrc<SRC>::DEN,
                                             (double)e2.y / VlSrc<SRC>::DEN, (double)e3.y / VlSrc<SRC>::DEN);
        a2 = (float)a64;
        na = (float)(1.0 - a64);
        mask = 15u;
    }
    float res[3] = {o.x, o.y, o.z};
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (mask & (1u << c)) {
            const double v = vm_mapcoord_blend(s, (double)((e0.x >> (8 * c)) & 255u), (double)((e1.x >> (8 * c)) & 255u),
                                               (double)((e2.x >> (8 * c)) & 255u), (double)((e3.x >> (8 * c)) & 255u));
            // samples whose half-up rounding a 1e-9 level perturbation of the value could flip (knife-edge
            // samples are counted, never masked): high half of the per-thread counter -> VM_STATUS_NEAR_KNIFE
            const double fr = (v + 0.5) - floor(v + 0.5);
            if (fr < 1e-9 || fr > 1.0 - 1e-9) *outside += VL_NEAR_KNIFE_UNIT;
            res[c] = __fmaf_rn(a2, (float)vm_round_half_up_u8(v), na * bgv[c]);
        }
    }
    return make_float4(res[0], res[1], res[2], a2);
}

// column interpolation of one coarse row (tps.py:68,73 with the column weights applied first)
__device__ __forceinline__ VlC vl_col_lerp(double2 a, double2 b, double y1, double yf) {
    VlC c;
    c.x = fma(b.x, yf, a.x * y1);
    c.y = fma(b.y, yf, a.y * y1);
    return c;
}

// fixed-point blend + composite of one pixel; returns the mask of what the exact path must redo
template <int SRC>
__device__ __forceinline__ unsigned vl_blend(const uint2 (&e)[4], uint32_t fa, uint32_t fb, bool fast, float b0, float b1,
                                             float b2, float4 &o, float &na_out) {
    constexpr float SCALE = (float)(1.0 / (VlSrc<SRC>::DEN * 1073741824.0));
    constexpr uint32_t DEN_U = (uint32_t)VlSrc<SRC>::DEN;
    // 2^-30 fixed-point weights (floor: each is at most one unit below the true product)
    const uint32_t A1 = fa >> 1, A0 = 0x80000000u - A1, B1 = fb >> 1, B0 = 0x80000000u - B1;
    const uint32_t W00 = __umulhi(A0, B0), W01 = __umulhi(A0, B1), W10 = __umulhi(A1, B0), W11 = __umulhi(A1, B1);
    const uint32_t w0 = W00 >> 6, w1 = W01 >> 6, w2 = W10 >> 6, w3 = W11 >> 6;     // 2^-24
    bool knife = false;
    uint32_t vv[3];
    float col[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t s0 = c == 0 ? (e[0].x & 255u) : __byte_perm(e[0].x, 0, 0x4440 + c);
        const uint32_t s1 = c == 0 ? (e[1].x & 255u) : __byte_perm(e[1].x, 0, 0x4440 + c);
        const uint32_t s2 = c == 0 ? (e[2].x & 255u) : __byte_perm(e[2].x, 0, 0x4440 + c);
        const uint32_t s3 = c == 0 ? (e[3].x & 255u) : __byte_perm(e[3].x, 0, 0x4440 + c);
        // value * 2^24 + 1/2 + window: the true value lies in [v, v + 1040] units (floored weights),
        // +-1 unit for the rounding of the fraction itself
        const uint32_t v = s0 * w0 + s1 * w1 + s2 * w2 + s3 * w3 + (8388608u + 1100u);
        col[c] = (float)(v >> 24);
        vv[c] = v;
        knife |= (v & 0x00FFFFFFu) < 1108u;
    }
    const unsigned long long a2f = (unsigned long long)e[0].y * W00 + (unsigned long long)e[1].y * W01 +
                                   (unsigned long long)e[2].y * W10 + (unsigned long long)e[3].y * W11;
    const uint32_t ws = W00 + W01 + W10 + W11;
    const unsigned long long naf = (unsigned long long)DEN_U * ws - a2f;
    const float a2 = (float)a2f * SCALE, na = (float)naf * SCALE;
    o.x = __fmaf_rn(a2, col[0], na * b0);
    o.y = __fmaf_rn(a2, col[1], na * b1);
    o.z = __fmaf_rn(a2, col[2], na * b2);
    o.w = a2;
    na_out = na;
    if (fast && !knife) return 0u;
    unsigned unc = fast ? 0u : 8u;                                     // rare: which colours need the exact evaluation
#pragma unroll
    for (int c = 0; c < 3; ++c)
        if ((vv[c] & 0x00FFFFFFu) < 1108u) unc |= 1u << c;
    return unc;
}

// floor + 2^-32 fraction of both coordinates; true when all four taps are strictly inside the
// frame and away from the t = 0 edge (the exact path decides there)
__device__ __forceinline__ bool vl_geometry(double t0, double t1, int h, int w, int &n0, int &n1, uint32_t &fa, uint32_t &fb) {
    const double m0 = t0 + VL_MAGIC, m1 = t1 + VL_MAGIC;
    n0 = __double2hiint(m0) - VL_MAGIC_HI; n1 = __double2hiint(m1) - VL_MAGIC_HI;
    fa = (uint32_t)__double2loint(m0); fb = (uint32_t)__double2loint(m1);
    return (unsigned)(n0 - 1) < (unsigned)(h - 2) && (unsigned)(n1 - 1) < (unsigned)(w - 2);
}

__device__ __forceinline__ void vl_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void vl_mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
