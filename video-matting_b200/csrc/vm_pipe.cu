// Persistent three-role pipeline for the fused clip kernels (C3: TPS + composite, C4: flow warp +
// consistency mask + TPS + composite).  Reference behaviour: flow.py:9-65, tps.py:14-123,
// augmentation.py:44-63 (identity affine), reader.py:72-79.
//
// One kernel, one CTA shape (256 threads), three kinds of work items popped IN ORDER from a
// global queue:
//
//   A(band)  flow stage for 16 output rows of one frame: warp_bgr + warp_img + correct_alpha ->
//            float4 {B, G, R, TA} per pixel (B,G,R exact uint8 values, TA = alpha numerator,
//            alpha = TA / 261120, 0 where masked) into a ring of rows that lives in L2.
//            For C3 (no flow) the band is just converted BGRA -> float4.
//   R(band)  radial-basis sum of the spline on 8 coarse rows (float64, table log) -> double2
//            {row coord, col coord} per coarse point into a second ring; publishes the band's
//            range of source rows.
//   B(tile)  16 rows x 256 columns of output: bilinear up-sampling of the coarse transform
//            (float64), map_coordinates geometry (float64), 4 float4 gathers from the A ring,
//            float32 blend with an exact float64 re-evaluation near rounding boundaries,
//            composite, one 16-byte streaming store per pixel.
//
// Items of the three kinds are interleaved in the queue (A and R run VP_LEAD bands ahead of B), so
// the CTAs resident on an SM are a mix of fp64-bound (R), gather/integer-bound (A) and
// fp32/LSU-bound (B) work at any time and the pipes overlap.  Dependencies are per-band flags
// in global memory (release: __syncthreads + __threadfence + atomic; acquire: poll + __threadfence,
// which also invalidates L1).  A consumer only ever waits for items that were queued BEFORE it, so
// progress never depends on CTAs that have not started (no co-residency requirement).  Anything a
// B tile needs outside the staged window (degenerate grids) is evaluated per pixel from the
// original inputs instead of waited for.
//
// HBM traffic is the algorithmic 39 (C4) / 23 (C3) bytes per pixel: the rings (16 MB + 4 MB at
// 1080p) are rewritten in place while still resident in the 126 MB L2.
#include "vm_common.cuh"
#include <math.h>
#include <mutex>

#define VP_THREADS 256
#define VP_BAND 16              // output rows per band
#define VP_CBAND 8              // coarse rows per R band
#define VP_BTW 256              // columns per A / B tile
#define VP_RCG 20               // 8-column coarse groups per R item (160 coarse columns)
#define VP_BACK 6               // B may read staged rows up to this many bands behind its own
#define VP_TCR 10               // coarse rows a 16-row tile can touch
#define VP_LOG_BITS 8
#define VP_LOG_N (1 << VP_LOG_BITS)
#define VP_MAX_N 64
#define VP_PT 5                 // coarse points per thread per pass
#define VP_MAGIC 6755399441055744.0     /* 1.5 * 2^52 */
#define VP_TA_DEN 261120.f
#define VP_BIAS (1 << 30)

struct VpParams {
    const uint32_t *fg; const float2 *bwd; const float2 *fwd; const uint8_t *bg; int n_bg;
    const double *ctrl; const double *coef; int N;
    int nx, ny; double step_x, step_y;
    const vm_axis_entry *rows; const vm_axis_entry *cols;
    int n, h, w;
    float4 *out;
    float4 *inter; double2 *coarse;
    unsigned ring_mask, cring_mask;
    int S, SR, TX, TXR, total_items, lead, roles;
    int *queue, *a_cnt, *r_cnt, *r_lo, *r_hi, *r_bad, *b_cnt;
    int32_t *status;
};

struct __align__(16) VpSmemR {                 // role R
    double2 logtab[VP_LOG_N];
    double4 ctrl[VP_MAX_N];
    double aff[6];
};
struct __align__(16) VpSmemB {                 // role B
    double2 tc[VP_TCR][VP_BTW];
    uint8_t bgt[VP_BAND][VP_BTW * 3];
    vm_axis_entry rows[VP_BAND];
};
struct __align__(16) VpSmem {
    union { VpSmemR r; VpSmemB b; };
    int item[2];
    int bad, slo, shi, pad;
};

// ---------------------------------------------------------------------------------------
// log table: c_k = 1 + (k + 1/2) / 2^B, entries {1/c_k, -log(1/c_k)}; log1p(r) Taylor to r^5
// (|r| <= 2^-9: truncation < 1e-17)
// ---------------------------------------------------------------------------------------
__device__ double2 g_vp_log_tab[VP_LOG_N];
static std::mutex g_vp_mu;
static bool g_vp_init[64];
static int g_vp_sms[64], g_vp_occ[64][3];
int g_vp_lead = 24, g_vp_ring_rows = 1024, g_vp_cring_rows = 512, g_vp_blocks = 0, g_vp_roles = 7;

__device__ __forceinline__ double vp_log(double x, const double2 *__restrict__ tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, lo);
    const double2 t = *reinterpret_cast<const double2 *>(
        reinterpret_cast<const char *>(tab) + ((hi >> (16 - VP_LOG_BITS)) & ((VP_LOG_N - 1) << 4)));
    const double ed = __hiloint2double(0x43300000, (int)((unsigned)hi >> 20)) - 4503599627371519.0;
    const double r = fma(m, t.x, -1.0);
    double q = fma(r, 0.2, -0.25);
    q = fma(r, q, 1.0 / 3.0);
    q = fma(r, q, -0.5);
    q = fma(r, q, 1.0);
    return fma(ed, 0.6931471805599453094, fma(r, q, t.y));
}

#ifdef VP_TIMING
__device__ __forceinline__ unsigned long long vp_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define VP_T(var) const unsigned long long var = vp_now()
#define VP_ACC(slot, a, b) do { if (threadIdx.x == 0) atomicAdd(reinterpret_cast<unsigned *>(P.queue) + (slot), (unsigned)((b) - (a))); } while (0)
#else
#define VP_T(var)
#define VP_ACC(slot, a, b)
#endif

__device__ __forceinline__ int vp_poll(const int *p) {
    return *reinterpret_cast<const volatile int *>(p);
}

// ---------------------------------------------------------------------------------------
// role R: spline on coarse rows [8s, 8s+8) x 160 coarse columns of frame f.  One unit per warp:
// footprint = 8 columns x 4 rows of coarse points (neighbouring r^2 -> neighbouring table
// entries -> few shared-memory wavefronts), VP_PT footprints side by side.
// ---------------------------------------------------------------------------------------
__device__ void vp_role_R(const VpParams &P, VpSmem &S, int f, int s, int tr) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, lx = lane & 7, ly = lane >> 3;
    const int N = P.N;
    {
        const double *Pc = P.ctrl + (int64_t)f * N * 2, *C = P.coef + (int64_t)f * (N + 3) * 2;
        for (int k = tid; k < VP_LOG_N; k += VP_THREADS) S.r.logtab[k] = g_vp_log_tab[k];
        if (tid < N) S.r.ctrl[tid] = make_double4(Pc[2 * tid], Pc[2 * tid + 1], 0.5 * C[2 * tid], 0.5 * C[2 * tid + 1]);
        if (tid >= 64 && tid < 70) { const int t = tid - 64; S.r.aff[t] = C[(N + t % 3) * 2 + t / 3]; }
    }
    __syncthreads();
    const int rg = warp >> 2, cg0 = tr * VP_RCG + (warp & 3) * VP_PT;
    const int k = s * VP_CBAND + rg * 4 + ly;
    const unsigned crow0 = (unsigned)(f * P.S * VP_CBAND);
    int rmin = INT_MAX, rmax = INT_MIN, bad = 0;
    if (s * VP_CBAND + rg * 4 < P.nx && cg0 * 8 < P.ny) {          // warp-uniform
        const double x = (double)min(k, P.nx - 1) * P.step_x;
        double py[VP_PT], s0[VP_PT], s1[VP_PT];
#pragma unroll
        for (int m = 0; m < VP_PT; ++m) {
            py[m] = (double)min(lx + 8 * (cg0 + m), P.ny - 1) * P.step_y;
            s0[m] = 0.0; s1[m] = 0.0;
        }
        for (int a = 0; a < N; ++a) {
            const double4 c = S.r.ctrl[a];
            const double dx = x - c.x;
            const double dx2 = dx * dx;
#pragma unroll
            for (int m = 0; m < VP_PT; ++m) {
                const double dy = py[m] - c.y;
                const double r2 = fma(dy, dy, dx2);
                const double U = r2 * vp_log(r2, S.r.logtab);      // 0 * finite = 0 at r2 = 0 (tps.py:81)
                s0[m] = fma(c.z, U, s0[m]);
                s1[m] = fma(c.w, U, s1[m]);
            }
        }
        if (k < P.nx) {
            double2 *dst = P.coarse + (size_t)((crow0 + k) & P.cring_mask) * P.ny;
#pragma unroll
            for (int m = 0; m < VP_PT; ++m) {
                const int l = lx + 8 * (cg0 + m);
                if (l < P.ny) {
                    const double v0 = ((S.r.aff[0] + S.r.aff[1] * x) + S.r.aff[2] * py[m]) + s0[m];
                    const double v1 = ((S.r.aff[3] + S.r.aff[4] * x) + S.r.aff[5] * py[m]) + s1[m];
                    dst[l] = make_double2(v0, v1);
                    if (!(fabs(v0) < 1.0e9) || !(fabs(v1) < 1.0e9)) bad = 1;
                    else { const int f0 = __double2int_rd(v0); rmin = min(rmin, f0); rmax = max(rmax, f0); }
                }
            }
        }
    }
    rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
    bad = __reduce_max_sync(0xffffffffu, bad);
    const int g = f * P.S + s;
    if (lane == 0) {                                               // zero-initialised biased extrema
        if (rmax >= rmin) { atomicMax(P.r_hi + g, rmax + VP_BIAS); atomicMax(P.r_lo + g, VP_BIAS - rmin); }
        if (bad) atomicOr(P.r_bad + g, 1);
    }
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(P.r_cnt + g, 1); }
}

// ---------------------------------------------------------------------------------------
// role A: warp_bgr / warp_img / correct_alpha -> {B, G, R, TA}.  Two phases so that a thread's
// gathers are all in flight together: vp_flow_issue computes the fixed-point sample position and
// issues the 4 BGRA taps and the forward-flow vector from CLAMPED (always valid) addresses;
// vp_flow_finish blends.  Pixels whose taps / forward sample are not interior, or whose flow is
// NaN / huge, are redone by the generic exact routines (rare).
// ---------------------------------------------------------------------------------------
struct VpTaps { uint32_t s00, s01, s10, s11, fxy; float2 ff; float cj0, ci0; bool ok; };

template <bool MASK>
__device__ __forceinline__ void vp_flow_issue(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd,
                                              int H, int W, float fi, float fj, float2 fb, VpTaps &t) {
    const float mx = __fadd_rn(fj, fb.x), my = __fadd_rn(fi, fb.y);
    bool ok = (fabsf(mx) < 60000.f) & (fabsf(my) < 60000.f);      // each component on its own: false for NaN
    const int SX = vm_fix5_fast(mx), SY = vm_fix5_fast(my);
    const int ix = SX >> 5, iy = SY >> 5;
    t.fxy = (uint32_t)(SX & 31) | ((uint32_t)(SY & 31) << 8);
    ok = ok && (unsigned)ix < (unsigned)(W - 1) && (unsigned)iy < (unsigned)(H - 1);
    const uint32_t *p = fg32 + (min(max(iy, 0), H - 2) * W + min(max(ix, 0), W - 2));
    t.s00 = __ldg(p); t.s01 = __ldg(p + 1); t.s10 = __ldg(p + W); t.s11 = __ldg(p + W + 1);
    if (MASK) {
        const int j0 = __float2int_rz(mx), i0 = __float2int_rz(my);
        ok = ok && (unsigned)j0 < (unsigned)W && (unsigned)i0 < (unsigned)H;
        const int cj = min(max(j0, 0), W - 1), ci = min(max(i0, 0), H - 1);
        t.ff = __ldg(fwd + (ci * W + cj));
        t.cj0 = (float)cj; t.ci0 = (float)ci;
    }
    t.ok = ok;
}

template <bool MASK>
__device__ __forceinline__ float4 vp_flow_finish(const VpTaps &t, const uint32_t *__restrict__ fg32,
                                                 const float2 *__restrict__ fwd, int H, int W, int i, int j,
                                                 float fi, float fj, float2 fb, int &flags) {
    const uint32_t fx = t.fxy & 255u, fy = t.fxy >> 8;
    // integer weights wx*wy (sum 1024) packed two per word; 16x8-bit dot products
    const uint32_t X = (32u - fx) + (fx << 16);
    const uint32_t W0 = X * (32u - fy), W1 = X * fy;             // {w00 | w01 << 16}, {w10 | w11 << 16}
    const uint32_t bg0 = __byte_perm(t.s00, t.s01, 0x5140), ra0 = __byte_perm(t.s00, t.s01, 0x7362);
    const uint32_t bg1 = __byte_perm(t.s10, t.s11, 0x5140), ra1 = __byte_perm(t.s10, t.s11, 0x7362);
    uint32_t b = __dp2a_lo(W1, bg1, __dp2a_lo(W0, bg0, 512u)) >> 10;
    uint32_t g = __dp2a_hi(W1, bg1, __dp2a_hi(W0, bg0, 512u)) >> 10;
    uint32_t r = __dp2a_lo(W1, ra1, __dp2a_lo(W0, ra0, 512u)) >> 10;
    uint32_t ta = __dp2a_hi(W1, ra1, __dp2a_hi(W0, ra0, 0u));
    bool ok = t.ok;
    int masked = 0;
    if (MASK) {
        const float c = __fadd_rn(t.ff.x, t.cj0), d = __fadd_rn(t.ff.y, t.ci0);
        ok = ok && (fabsf(c) < 3.0e38f) && (fabsf(d) < 3.0e38f);
        // min(trunc(c), W-1) - j in float: exact for |.| < 2^24, monotone beyond (flow.py:47-48)
        const float dj = __fadd_rn(fminf(truncf(c), (float)(W - 1)), -fj);
        const float di = __fadd_rn(fminf(truncf(d), (float)(H - 1)), -fi);
        masked = __fmaf_rn(dj, dj, __fmul_rn(di, di)) > 225.f;
    }
    if (!ok) {                                                    // borders, out-of-frame, NaN / huge flows
        const VmWarped wv = vm_flow_warp_bgra(reinterpret_cast<const uint8_t *>(fg32), H, W, i, j, fb);
        b = wv.bgr & 255u; g = (wv.bgr >> 8) & 255u; r = (wv.bgr >> 16) & 255u; ta = wv.ta;
        masked = MASK ? vm_consistency(fwd, H, W, i, j, fb, flags) : 0;
    }
    return make_float4((float)b, (float)g, (float)r, masked ? 0.f : (float)ta);
}

template <bool MASK>
__device__ __forceinline__ float4 vp_flow_px(const uint32_t *__restrict__ fg32, const float2 *__restrict__ fwd,
                                             int H, int W, int i, int j, float fi, float fj, float2 fb, int &flags) {
    VpTaps t;
    vp_flow_issue<MASK>(fg32, fwd, H, W, fi, fj, fb, t);
    return vp_flow_finish<MASK>(t, fg32, fwd, H, W, i, j, fi, fj, fb, flags);
}

__device__ __forceinline__ float4 vp_convert_px(uint32_t s) {
    return make_float4((float)(s & 255u), (float)((s >> 8) & 255u), (float)((s >> 16) & 255u),
                       (float)((s >> 24) << 10));            // A/255 = 1024 A / 261120
}

// MODE 0: no flow (C3), 1: flow warp, 2: flow warp + consistency mask.
// Tile = 16 rows x 256 columns; thread = (4-pixel column group, row quad), 4 rows each, the flow
// vectors of all 4 rows requested up front.
template <int MODE>
__device__ void vp_role_A(const VpParams &P, int f, int s, int tx, int &flags) {
    const int tid = threadIdx.x;
    const int h = P.h, w = P.w;
    const int i0 = s * VP_BAND, nrows = min(VP_BAND, h - i0);
    const int J0 = tx * VP_BTW;
    const int64_t fbase = (int64_t)f * h * w;
    const uint32_t *fg32 = P.fg + fbase;
    const float2 *bf = MODE ? P.bwd + fbase : nullptr;
    const float2 *ff = (MODE == 2) ? P.fwd + fbase : nullptr;
    const unsigned grow0 = (unsigned)(f * P.S * VP_BAND + i0);
    if ((w & 3) == 0) {
        const int j = J0 + (tid & 63) * 4, rq = tid >> 6;
        if (j >= w) return;
        const float fj = (float)j;
        if (MODE == 0) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = rq + 4 * it;
                if (r < nrows) {
                    const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(fg32 + (i0 + r) * w + j));
                    float4 *dst = P.inter + (size_t)((grow0 + r) & P.ring_mask) * w + j;
                    dst[0] = vp_convert_px(v.x); dst[1] = vp_convert_px(v.y);
                    dst[2] = vp_convert_px(v.z); dst[3] = vp_convert_px(v.w);
                }
            }
        } else {
            float4 fa[4], fb[4];
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = min(rq + 4 * it, nrows - 1);
                const float4 *src = reinterpret_cast<const float4 *>(bf + (i0 + r) * w + j);
                fa[it] = __ldcs(src); fb[it] = __ldcs(src + 1);
            }
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = rq + 4 * it;
                if (r < nrows) {
                    const int i = i0 + r;
                    const float fi = (float)i;
                    float4 *dst = P.inter + (size_t)((grow0 + r) & P.ring_mask) * w + j;
                    const float2 fl[4] = {{fa[it].x, fa[it].y}, {fa[it].z, fa[it].w}, {fb[it].x, fb[it].y}, {fb[it].z, fb[it].w}};
                    VpTaps tp[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        vp_flow_issue<MODE == 2>(fg32, ff, h, w, fi, fj + (float)k, fl[k], tp[k]);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        dst[k] = vp_flow_finish<MODE == 2>(tp[k], fg32, ff, h, w, i, j + k, fi, fj + (float)k, fl[k], flags);
                }
            }
        }
    } else {
        const int tw = min(VP_BTW, w - J0);
        for (int idx = tid; idx < nrows * tw; idx += VP_THREADS) {
            const int r = idx / tw, j = J0 + idx - r * tw, i = i0 + r, p = i * w + j;
            float4 *dst = P.inter + (size_t)((grow0 + r) & P.ring_mask) * w + j;
            if (MODE == 0) *dst = vp_convert_px(__ldg(fg32 + p));
            else *dst = vp_flow_px<MODE == 2>(fg32, ff, h, w, i, j, (float)i, (float)j, __ldg(bf + p), flags);
        }
    }
}

// ---------------------------------------------------------------------------------------
// role B: 16 rows x 256 columns of output; thread = column, two rows per step
// ---------------------------------------------------------------------------------------
struct VpPix { double t0, t1; float4 e00, e01, e10, e11; float af, bf; bool fast; };

template <int MODE>
__device__ void vp_role_B(const VpParams &P, VpSmem &S, int f, int s, int tx, int &outside, int &flags) {
    const int tid = threadIdx.x;
    const int h = P.h, w = P.w;
    const int I0 = s * VP_BAND, th = min(VP_BAND, h - I0);
    const int j = tx * VP_BTW + tid;
    const int gbase = f * P.S;
    if (tid < th) S.b.rows[tid] = vm_ld_axis(P.rows + I0 + tid);
    __syncthreads();
    const int kr0 = S.b.rows[0].i0, nkr = S.b.rows[th - 1].i1 - kr0 + 1;
    const bool sync = P.roles == 7;
    // background tile -> shared memory, asynchronously (lands while the dependencies are polled)
    const uint8_t *bgf = P.bg + ((int64_t)(f % P.n_bg) * h * w) * 3;
    const int tw = min(VP_BTW, w - tx * VP_BTW);
    const bool bg_async = (w & 15) == 0 && tw == VP_BTW && (reinterpret_cast<uintptr_t>(P.bg) & 15) == 0;
    if (bg_async) {
        for (int c = tid; c < th * (VP_BTW * 3 / 16); c += VP_THREADS) {
            const int r = c / (VP_BTW * 3 / 16), k = c - r * (VP_BTW * 3 / 16);
            const uint8_t *src = bgf + ((int64_t)(I0 + r) * w + tx * VP_BTW) * 3 + k * 16;
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&S.b.bgt[r][k * 16]);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    } else {
        for (int c = tid; c < th * tw * 3; c += VP_THREADS) {
            const int r = c / (tw * 3), k = c - r * (tw * 3);
            S.b.bgt[r][k] = __ldg(bgf + ((int64_t)(I0 + r) * w + tx * VP_BTW) * 3 + k);
        }
    }
    VP_T(tb0);
    // ---- coarse bands this tile reads --------------------------------------------------------
    int a_lo = 0, a_hi = -1;
    if (tid < 32) {
        const int rb_lo = kr0 / VP_CBAND, rb_hi = min((kr0 + nkr - 1) / VP_CBAND, P.SR - 1);
        const bool mine = rb_lo + tid <= rb_hi;
        if (mine && sync) while (vp_poll(P.r_cnt + gbase + rb_lo + tid) < P.TXR) __nanosleep(32);
        __threadfence();
        int rmin = INT_MAX, rmax = INT_MIN, bad = 0;
        if (mine) {
            rmin = VP_BIAS - vp_poll(P.r_lo + gbase + rb_lo + tid);
            rmax = vp_poll(P.r_hi + gbase + rb_lo + tid) - VP_BIAS;
            bad = vp_poll(P.r_bad + gbase + rb_lo + tid);
        }
        rmin = __reduce_min_sync(0xffffffffu, rmin); rmax = __reduce_max_sync(0xffffffffu, rmax);
        bad = __reduce_max_sync(0xffffffffu, bad) | (nkr > VP_TCR ? 1 : 0);
        const int win_lo = max(s - VP_BACK, 0), win_hi = min(s + P.lead - 2, P.S - 1);
        a_lo = max(max(rmin, 0) / VP_BAND, win_lo); a_hi = min(min(rmax + 1, h - 1) / VP_BAND, win_hi);
        if (bad || !sync) a_hi = a_lo - 1;
        if (tid == 0) { S.bad = bad; S.slo = win_lo * VP_BAND; S.shi = min((win_hi + 1) * VP_BAND, h); }
    }
    __syncthreads();
    VP_T(tb1); VP_ACC(10, tb0, tb1);
    const bool ok = !S.bad;
    // ---- column-interpolated transform of this thread's column for every coarse row of the tile
    vm_axis_entry ce;
    ce.frac = 0.0; ce.i0 = 0; ce.i1 = 0;
    if (j < w) ce = vm_ld_axis(P.cols + j);
    if (j < w && ok) {
        const unsigned crow0 = (unsigned)(f * P.S * VP_CBAND + kr0);
#pragma unroll 5
        for (int k = 0; k < nkr; ++k) {
            const double2 *row = P.coarse + (size_t)((crow0 + k) & P.cring_mask) * P.ny;
            const double2 a = row[ce.i0], b = row[ce.i1];
            S.b.tc[k][tid] = make_double2(fma(ce.frac, b.x - a.x, a.x), fma(ce.frac, b.y - a.y, a.y));
        }
    }
    // ---- staged source rows (polled while the transform loads are in flight) -------------------
    if (tid < 32) {
        if (a_lo + tid <= a_hi) while (vp_poll(P.a_cnt + gbase + a_lo + tid) < P.TX) __nanosleep(32);
        __threadfence();
    }
    if (bg_async) asm volatile("cp.async.wait_group 0;");
    __syncthreads();
    VP_T(tb2); VP_ACC(11, tb1, tb2);
    const int slo = S.slo, srange = S.shi - 1 - slo;
    int nout = 0;
    if (j < w) {
        const unsigned frow0 = (unsigned)(f * P.S * VP_BAND);
        const int64_t fbase = (int64_t)f * h * w;
        float4 *outf = P.out + fbase;
        for (int ir0 = 0; ir0 < th; ir0 += 2) {
            VpPix px[2];
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                const vm_axis_entry re = S.b.rows[min(ir0 + z, th - 1)];
                double t0 = 0.0, t1 = 0.0;
                if (ok) {
                    const double2 TA = S.b.tc[re.i0 - kr0][tid], TB = S.b.tc[re.i1 - kr0][tid];
                    t0 = fma(re.frac, TB.x - TA.x, TA.x); t1 = fma(re.frac, TB.y - TA.y, TA.y);
                }
                // floor / fraction in float64: a round-down add of 1.5 * 2^52 leaves floor(t) in the low word
                const double m0 = __dadd_rd(t0, VP_MAGIC), m1 = __dadd_rd(t1, VP_MAGIC);
                const int n0 = __double2loint(m0), n1 = __double2loint(m1);
                const double d0 = t0 - (m0 - VP_MAGIC), d1 = t1 - (m1 - VP_MAGIC);
                const bool fast = ok && (unsigned)(n0 - slo) < (unsigned)srange && (unsigned)n1 < (unsigned)(w - 1);
                const unsigned gr = frow0 + (unsigned)(fast ? n0 : slo), cc = (unsigned)(fast ? n1 : 0);
                const float4 *p0 = P.inter + ((gr & P.ring_mask) * (unsigned)w + cc);
                const float4 *p1 = P.inter + (((gr + 1u) & P.ring_mask) * (unsigned)w + cc);
                px[z].e00 = p0[0]; px[z].e01 = p0[1]; px[z].e10 = p1[0]; px[z].e11 = p1[1];
                px[z].t0 = t0; px[z].t1 = t1; px[z].af = (float)d0; px[z].bf = (float)d1; px[z].fast = fast;
            }
#pragma unroll
            for (int z = 0; z < 2; ++z) {
                if (ir0 + z >= th) break;
                const float4 e00 = px[z].e00, e01 = px[z].e01, e10 = px[z].e10, e11 = px[z].e11;
                const float af = px[z].af, bfr = px[z].bf;
                const float a0f = 1.f - af, b0f = 1.f - bfr;
                const float w00 = a0f * b0f, w01 = a0f * bfr, w10 = af * b0f, w11 = af * bfr;
                const float vb = __fmaf_rn(e11.x, w11, __fmaf_rn(e10.x, w10, __fmaf_rn(e01.x, w01, e00.x * w00)));
                const float vg = __fmaf_rn(e11.y, w11, __fmaf_rn(e10.y, w10, __fmaf_rn(e01.y, w01, e00.y * w00)));
                const float vr = __fmaf_rn(e11.z, w11, __fmaf_rn(e10.z, w10, __fmaf_rn(e01.z, w01, e00.z * w00)));
                float cb = (vb + 12582912.f) - 12582912.f, cg = (vg + 12582912.f) - 12582912.f,
                      cr = (vr + 12582912.f) - 12582912.f;
                const float dev = fmaxf(fmaxf(fabsf(vb - cb), fabsf(vg - cg)), fabsf(vr - cr));
                const float ta2 = __fmaf_rn(e11.w, w11, __fmaf_rn(e10.w, w10, __fmaf_rn(e01.w, w01, e00.w * w00)));
                const float nt2 = __fmaf_rn(VP_TA_DEN - e11.w, w11, __fmaf_rn(VP_TA_DEN - e10.w, w10,
                                  __fmaf_rn(VP_TA_DEN - e01.w, w01, (VP_TA_DEN - e00.w) * w00)));
                float a2 = ta2 * VM_ALPHA_INV, na = nt2 * VM_ALPHA_INV;
                const int p = (I0 + ir0 + z) * w + j;
                const uint8_t *bp = &S.b.bgt[ir0 + z][tid * 3];
                const float bb = (float)(uint32_t)bp[0], bgc = (float)(uint32_t)bp[1], br = (float)(uint32_t)bp[2];
                if (!px[z].fast || dev > 0.4995f) {
                    // rare: exact float64 evaluation (knife-edge samples, last row/column, outside, unstaged rows)
                    const VmBilin64 sb = vm_mapcoord_setup(px[z].t0, px[z].t1, h, w);
                    if (!sb.inside) {
                        cb = cg = cr = 0.f; a2 = 0.f; na = 1.f; nout++;
                    } else if (px[z].fast) {
                        cb = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, (double)e00.x, (double)e01.x, (double)e10.x, (double)e11.x));
                        cg = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, (double)e00.y, (double)e01.y, (double)e10.y, (double)e11.y));
                        cr = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, (double)e00.z, (double)e01.z, (double)e10.z, (double)e11.z));
                    } else {
                        const uint8_t *fg8 = reinterpret_cast<const uint8_t *>(P.fg + fbase);
                        const float2 *bfl = MODE ? P.bwd + fbase : nullptr;
                        const float2 *ffl = (MODE == 2) ? P.fwd + fbase : nullptr;
                        const VmSrcPx s00 = vm_src_px<MODE != 0>(fg8, bfl, ffl, h, w, sb.i0, sb.j0, flags);
                        const VmSrcPx s01 = vm_src_px<MODE != 0>(fg8, bfl, ffl, h, w, sb.i0, sb.j1, flags);
                        const VmSrcPx s10 = vm_src_px<MODE != 0>(fg8, bfl, ffl, h, w, sb.i1, sb.j0, flags);
                        const VmSrcPx s11 = vm_src_px<MODE != 0>(fg8, bfl, ffl, h, w, sb.i1, sb.j1, flags);
                        cb = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, s00.b, s01.b, s10.b, s11.b));
                        cg = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, s00.g, s01.g, s10.g, s11.g));
                        cr = (float)vm_round_half_up_u8(vm_mapcoord_blend(sb, s00.r, s01.r, s10.r, s11.r));
                        const double a64 = vm_mapcoord_blend(sb, s00.a, s01.a, s10.a, s11.a);
                        a2 = (float)a64; na = (float)(1.0 - a64);
                    }
                }
                __stcs(outf + p, make_float4(__fmaf_rn(a2, cb, na * bb), __fmaf_rn(a2, cg, na * bgc),
                                             __fmaf_rn(a2, cr, na * br), a2));
            }
        }
    }
    outside += nout;
    VP_T(tb3); VP_ACC(9, tb2, tb3); VP_ACC(8, tb3, tb3 + 1);
}

// ---------------------------------------------------------------------------------------
// scheduler: queue step q / ips holds the A tiles and R items of band (step) and the B tiles of
// band (step - lead)
// ---------------------------------------------------------------------------------------
#ifndef VP_OCC
#define VP_OCC 3
#endif
template <int MODE>
__global__ void __launch_bounds__(VP_THREADS, VP_OCC)
k_pipe(const __grid_constant__ VpParams P) {
    extern __shared__ __align__(16) unsigned char vp_smem_raw[];
    VpSmem &S = *reinterpret_cast<VpSmem *>(vp_smem_raw);
    const int tid = threadIdx.x;
    const int ips = 2 * P.TX + P.TXR, nbands = P.n * P.S;
    const int ring_bands = (int)((P.ring_mask + 1u) / VP_BAND), cring_bands = (int)((P.cring_mask + 1u) / VP_CBAND);
    int outside = 0, flags = 0, slot = 0;
    if (tid == 0) S.item[0] = atomicAdd(P.queue, 1);
    __syncthreads();
    int q = S.item[0];
    while (q < P.total_items) {
        int nxt = 0;
        if (tid == VP_THREADS - 32) nxt = atomicAdd(P.queue, 1);       // next item, consumed at the end of this one
        const int stq = q / ips, sub = q - stq * ips;
        if (sub < P.TX) {
            if (stq < nbands && (P.roles & 1)) {
                const int g = stq, f = g / P.S, s = g - f * P.S;
                VP_T(ta0);
                // ring rows of band g last held band g - ring_bands, read by B bands [.. - lead, .. + BACK]
                if (g >= ring_bands && P.roles == 7) {
                    const int h0 = g - ring_bands - P.lead;
                    for (int t = tid; t <= P.lead + VP_BACK; t += VP_THREADS)
                        if (h0 + t >= 0) while (vp_poll(P.b_cnt + h0 + t) < P.TX) __nanosleep(32);
                    __threadfence();
                    __syncthreads();
                }
                VP_T(ta1); VP_ACC(4, ta0, ta1);
                vp_role_A<MODE>(P, f, s, sub, flags);
                __syncthreads();
                VP_T(ta2); VP_ACC(3, ta1, ta2); VP_ACC(2, ta2, ta2 + 1);
                if (tid == 0) { __threadfence(); atomicAdd(P.a_cnt + g, 1); }
            }
        } else if (sub < P.TX + P.TXR) {
            const int g = stq, f = g / P.S, s = g - f * P.S;
            if (stq < nbands && s < P.SR && (P.roles & 2)) {
                if (g >= cring_bands && P.roles == 7) {
                    const int h0 = g - cring_bands - 1;
                    if (tid <= 2 && h0 + tid >= 0) while (vp_poll(P.b_cnt + h0 + tid) < P.TX) __nanosleep(32);
                    __threadfence();
                    __syncthreads();
                }
                VP_T(tr1);
                vp_role_R(P, S, f, s, sub - P.TX);
                __syncthreads();
                VP_T(tr2); VP_ACC(6, tr1, tr2); VP_ACC(5, tr2, tr2 + 1);
            }
        } else {
            const int st = stq - P.lead;
            if (st >= 0 && (P.roles & 4)) {
                const int f = st / P.S, s = st - f * P.S;
                vp_role_B<MODE>(P, S, f, s, sub - P.TX - P.TXR, outside, flags);
                __syncthreads();
                if (tid == 0) { __threadfence(); atomicAdd(P.b_cnt + st, 1); }
            }
        }
        if (tid == VP_THREADS - 32) S.item[slot ^ 1] = nxt;
        __syncthreads();
        slot ^= 1;
        q = S.item[slot];
    }
    if (P.status) {
        outside = __reduce_add_sync(0xffffffffu, outside);
        flags = __reduce_or_sync(0xffffffffu, flags);
        if ((tid & 31) == 0) {
            if (outside) atomicAdd(P.status + VM_STATUS_TPS_OUTSIDE, outside);
            if (flags & 1) atomicAdd(P.status + VM_STATUS_INDEX_ERR, 1);
            if (flags & 2) atomicAdd(P.status + VM_STATUS_NAN_ERR, 1);
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static int vp_init(int &dev) {
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        vm_set_error("vm_pipe: cudaGetDevice failed");
        return VM_ERR_CUDA;
    }
    std::lock_guard<std::mutex> lk(g_vp_mu);
    if (g_vp_init[dev]) return VM_OK;
    static double2 tab[VP_LOG_N];
    for (int k = 0; k < VP_LOG_N; ++k) {
        const long double c = 1.0L + ((long double)k + 0.5L) / (long double)VP_LOG_N;
        const double inv = (double)(1.0L / c);
        tab[k].x = inv;
        tab[k].y = (double)(-logl((long double)inv));
    }
    cudaError_t e = cudaMemcpyToSymbol(g_vp_log_tab, tab, sizeof(tab));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_vp_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    const int smem = (int)sizeof(VpSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pipe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pipe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pipe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_vp_occ[dev][0], k_pipe<0>, VP_THREADS, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_vp_occ[dev][1], k_pipe<1>, VP_THREADS, smem);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g_vp_occ[dev][2], k_pipe<2>, VP_THREADS, smem);
    if (e != cudaSuccess) {
        vm_set_error("vm_pipe: init failed: %s", cudaGetErrorString(e));
        return VM_ERR_CUDA;
    }
    g_vp_init[dev] = true;
    return VM_OK;
}

static inline int64_t vp_align(int64_t v) { return (v + 255) & ~(int64_t)255; }
static inline int vp_pow2_rows(int64_t need, int cap) {
    int r = 16;
    while (r < cap && r < need) r <<= 1;
    return r;
}

struct VpLayout { int S, SR, TX, TXR, ring_rows, cring_rows; int64_t hdr, cbytes, ibytes; };

static VpLayout vp_layout(int n, int h, int w) {
    VpLayout L;
    L.S = (h + VP_BAND - 1) / VP_BAND;
    const int nx = h / 2, ny = w / 2;
    L.SR = (nx + VP_CBAND - 1) / VP_CBAND;
    L.TX = (w + VP_BTW - 1) / VP_BTW;
    L.TXR = ((ny + 7) / 8 + VP_RCG - 1) / VP_RCG;
    if (L.TXR < 1) L.TXR = 1;
    L.ring_rows = vp_pow2_rows((int64_t)n * L.S * VP_BAND, g_vp_ring_rows);
    L.cring_rows = vp_pow2_rows((int64_t)n * L.S * VP_CBAND, g_vp_cring_rows);
    L.hdr = vp_align((int64_t)(16 + 6 * (int64_t)n * L.S) * 4);
    L.cbytes = vp_align((int64_t)L.cring_rows * (ny > 0 ? ny : 1) * 16);
    L.ibytes = vp_align((int64_t)L.ring_rows * w * 16);
    return L;
}

int64_t vm_pipe_scratch_bytes(int n, int h, int w) {
    const VpLayout L = vp_layout(n, h, w);
    return L.hdr + L.cbytes + L.ibytes + 256;
}

// mode 0: C3 (no flow), 1: flow warp, 2: flow warp + mask.  All pointers device; enqueues a memset
// of the scratch header and one kernel on `st`.
int vm_pipe_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, void *scratch, int32_t *status, cudaStream_t st, const char *what) {
    int dev = 0;
    int rc = vp_init(dev);
    if (rc != VM_OK) return rc;
    VM_REQUIRE(scratch, "scratch workspace (vm_fused_scratch_bytes) required");
    VM_REQUIRE(N <= VP_MAX_N && nx == h / 2 && ny == w / 2 && nx >= 1 && ny >= 1, "grid not supported by the pipeline kernel");
    const VpLayout L = vp_layout(n, h, w);
    VM_REQUIRE((int64_t)n * L.S * VP_BAND < (1ll << 30), "clip too long for one launch");
    char *base = reinterpret_cast<char *>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    cudaError_t e = cudaMemsetAsync(base, 0, (size_t)L.hdr, st);
    if (e != cudaSuccess) { vm_set_error("%s: cudaMemsetAsync: %s", what, cudaGetErrorString(e)); return VM_ERR_CUDA; }
    VpParams P;
    P.fg = reinterpret_cast<const uint32_t *>(fg);
    P.bwd = reinterpret_cast<const float2 *>(backward);
    P.fwd = reinterpret_cast<const float2 *>(forward);
    P.bg = bg; P.n_bg = n_bg;
    P.ctrl = ctrl; P.coef = coef; P.N = N;
    P.nx = nx; P.ny = ny; P.step_x = step_x; P.step_y = step_y;
    P.rows = rows; P.cols = cols;
    P.n = n; P.h = h; P.w = w;
    P.out = reinterpret_cast<float4 *>(out);
    int *hdr = reinterpret_cast<int *>(base);
    const int nb = n * L.S;
    P.queue = hdr;
    P.a_cnt = hdr + 16; P.r_cnt = P.a_cnt + nb; P.r_lo = P.r_cnt + nb; P.r_hi = P.r_lo + nb; P.r_bad = P.r_hi + nb;
    P.b_cnt = P.r_bad + nb;
    P.coarse = reinterpret_cast<double2 *>(base + L.hdr);
    P.inter = reinterpret_cast<float4 *>(base + L.hdr + L.cbytes);
    P.ring_mask = (unsigned)L.ring_rows - 1u; P.cring_mask = (unsigned)L.cring_rows - 1u;
    P.S = L.S; P.SR = L.SR; P.TX = L.TX; P.TXR = L.TXR;
    P.lead = g_vp_lead; P.roles = g_vp_roles;
    VM_REQUIRE(L.ring_rows / VP_BAND >= nb || L.ring_rows / VP_BAND > P.lead + VP_BACK + 8, "ring too small for the lead");
    VM_REQUIRE(L.cring_rows / VP_CBAND >= nb || L.cring_rows / VP_CBAND > P.lead + 8, "coarse ring too small for the lead");
    P.total_items = (nb + P.lead) * (2 * L.TX + L.TXR);
    P.status = status;
    int grid = g_vp_sms[dev] * (g_vp_blocks > 0 ? g_vp_blocks : (g_vp_occ[dev][mode] > 0 ? g_vp_occ[dev][mode] : 1));
    if (grid > P.total_items) grid = P.total_items;
    if (mode == 0) k_pipe<0><<<grid, VP_THREADS, sizeof(VpSmem), st>>>(P);
    else if (mode == 1) k_pipe<1><<<grid, VP_THREADS, sizeof(VpSmem), st>>>(P);
    else k_pipe<2><<<grid, VP_THREADS, sizeof(VpSmem), st>>>(P);
    return vm_check_launch(what);
}
