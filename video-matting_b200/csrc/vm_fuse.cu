// Single-pass fused C4 kernel (fused_variant 5, the default for the flow + TPS + composite entry point):
// flow warp + consistency mask (flow.py:9-65), thin-plate spline on the coarse grid and its bilinear
// up-sampling (tps.py:41-123), map_coordinates (tps.py:34) and the composite (reader.py:72-79) in ONE
// persistent kernel in which every input byte is read from HBM once and every output byte written once:
// no packed intermediate, no coarse transform, no tile records in global memory (39 B/px algorithmic =
// the kernel's DRAM traffic up to halo re-reads, which hit L2).
//
// One CTA of 1024 threads per SM walks 60 x 60 output tiles (stride = grid size).  Warp-specialised:
//
//   spline warps (VF_NSW)   for tile i+1: the spline on the tile's <= 32 x 32 coarse window, float64,
//                           r^2 log r^2 from the shared-memory log table (the arithmetic of k_lean_coarse) ->
//                           T[(i+1)&1] in shared memory, plus the bounding box of floor(T) = the source box.
//   pixel warps (the rest)  for tile i:  P2  column-interpolated coarse rows Cs (shared memory)
//                                        P3  flow warp + mask of every pixel of the source box straight from
//                                            fg / backward / forward (coalesced flow reads, tap gathers through
//                                            the read-only path) -> packed {bgr, TA} box in shared memory
//                                        P4  row interpolation (6 DP / pixel), fixed-point resampling of the
//                                            four box taps, composite, one 16-byte streaming store per pixel.
//
// The float64 pipe (spline warps) and the load/store + integer pipes (pixel warps) are busy at the same
// time on every SM; the two roles hand tiles over through named barriers (bar.arrive / bar.sync), T is
// double buffered.  All arithmetic is the lean pipeline's (vm_lean.cuh), so the output is bit-identical to
// it (tests/test_gpu_fuse.py); tiles whose source box does not fit shared memory, or whose axis tables are
// not monotone windows, take slower exact paths inside the same kernel and are counted.
#include "vm_lean.cuh"
#include <string.h>
#include <atomic>

const void *vl_table_device();                     // vm_lean.cu: device address of the log table (after vl_init)
int vl_table_init();

#define VF_TILE 60                                 // output tile edge: 60 fine pixels need <= 32 coarse points per axis
#define VF_CW 32                                   // coarse window edge
#define VF_THREADS 1024
#define VF_CSP 64                                  // pitch of Cs (columns)
#define VF_BOX_CAP 8192                            // packed source-box entries (8 B each)
#ifndef VF_RG
#define VF_RG 4                                    // coarse rows evaluated together by a spline warp (2 or 4)
#endif

// -DVF_TIMING: cycles per phase, accumulated by one thread per role and CTA (scripts/fuse_phases.py)
#ifdef VF_TIMING
__device__ unsigned long long g_vf_prof[16];
#define VF_T(var) const long long var = clock64()
#define VF_ACC(slot, a, b) atomicAdd(&g_vf_prof[slot], (unsigned long long)((b) - (a)))
#else
#define VF_T(var)
#define VF_ACC(slot, a, b)
#endif

enum { VF_BAR_FULL0 = 1, VF_BAR_FULL1 = 2, VF_BAR_EMPTY0 = 3, VF_BAR_EMPTY1 = 4, VF_BAR_SPLINE = 5, VF_BAR_PIXEL = 6 };

struct __align__(16) VfTileInfo {
    int frame, I0, J0, th, tw;
    int kr0, nkr, kc0, nkc;
    int valid;                                     // the coarse window of the tile fits T
    int rlo, rhi, clo, chi, bad;                   // bounding box of floor(T) over the window (atomics of the spline warps)
    int rmin, bh, cmin, bw;                        // source box (bw = 0: no box, taps are evaluated one by one)
    int pad;
};

template <int NSW>
struct __align__(16) VfSmem {
    double2 tab[VL_TAB_N];
    double2 T[2][VF_CW * VF_CW];
    double2 Cs[VF_CW * VF_CSP];
    uint2 box[VF_BOX_CAP];
    vm_axis_entry rows[64];
    double2 p[VL_MAX_N];                           // control points of the spline warps' current frame
    double2 wv[VL_MAX_N];                          // {w0 / 2, w1 / 2}
    double aff[8];
    double4 dx2[NSW][VL_MAX_N];                    // (x_r - Px)^2 for the VF_RG rows of a warp's group
    VfTileInfo info[2];
    unsigned long long bar;
};

__device__ __forceinline__ void vf_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void vf_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool vf_bar_and(int id, int count, bool pred) {
    int r;
    asm volatile("{ .reg .pred p, q; setp.ne.s32 p, %3, 0; bar.red.and.pred q, %1, %2, p; selp.s32 %0, 1, 0, q; }"
                 : "=r"(r) : "r"(id), "r"(count), "r"((int)pred) : "memory");
    return r != 0;
}

// flow-warped, consistency-masked source pixel (qi, qj) as a packed {bgr, TA} entry, from the original inputs
// (generic routines: any coordinate, same integers as the fast path of vm_flow_px)
template <bool HAS_FWD>
__device__ __noinline__ uint2 vf_src_elem(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                          const float2 *__restrict__ fwd, int h, int w, int qi, int qj) {
    const float2 fb = __ldg(bwd + (int64_t)qi * w + qj);
    const VmWarped wv = vm_flow_warp_bgra(fg, h, w, qi, qj, fb);
    int flags = 0;
    const int m = HAS_FWD ? vm_consistency(fwd, h, w, qi, qj, fb, flags) : 0;
    return make_uint2(wv.bgr, m ? 0u : wv.ta);
}

// exact per-pixel evaluation: vl_exact_px with the four source pixels recomputed from the inputs
template <bool HAS_FWD>
__device__ __noinline__ float4 vf_exact_px(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                           const float2 *__restrict__ fwd, double t0, double t1, int h, int w,
                                           const uint8_t *__restrict__ bp, float4 o, float na, unsigned mask, int *outside,
                                           int *near_knife) {
    const float bgv[3] = {(float)__ldg(bp), (float)__ldg(bp + 1), (float)__ldg(bp + 2)};
    const VmBilin64 s = vm_mapcoord_setup(t0, t1, h, w);
    if (!s.inside) {
        (*outside)++;
        return make_float4(bgv[0], bgv[1], bgv[2], 0.f);
    }
    const uint2 e0 = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, s.i0, s.j0), e1 = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, s.i0, s.j1);
    const uint2 e2 = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, s.i1, s.j0), e3 = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, s.i1, s.j1);
    float a2 = o.w;
    if (mask & 8u) {
        const double a64 = vm_mapcoord_blend(s, (double)e0.y / VM_ALPHA_DEN, (double)e1.y / VM_ALPHA_DEN,
                                             (double)e2.y / VM_ALPHA_DEN, (double)e3.y / VM_ALPHA_DEN);
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
            // samples whose half-up rounding a 1e-9 level perturbation of the value could flip (SURVEY 8a-6:
            // knife-edge samples are counted, never masked)
            const double fr = (v + 0.5) - floor(v + 0.5);
            if (fr < 1e-9 || fr > 1.0 - 1e-9) (*near_knife)++;
            res[c] = __fmaf_rn(a2, (float)vm_round_half_up_u8(v), na * bgv[c]);
        }
    }
    return make_float4(res[0], res[1], res[2], a2);
}

// spline value at coarse point (k, l) with k_lean_coarse's roundings (generic path, controls from global memory)
__device__ __noinline__ double2 vf_tps_point(const double *__restrict__ P, const double *__restrict__ C, int NC, int k, int l,
                                             double step_x, double step_y, uint32_t tab_adj) {
    const double x = (double)k * step_x, y = (double)l * step_y;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
    for (int a = 0; a < NC; ++a) {
        const double dx = x - P[2 * a], dy = y - P[2 * a + 1];
        const double U = vl_u_any(dx * dx + dy * dy, tab_adj);
        s0 = fma(0.5 * C[2 * a], U, s0);
        s1 = fma(0.5 * C[2 * a + 1], U, s1);
    }
    const double v0 = ((C[(NC + 0) * 2] + C[(NC + 1) * 2] * x) + C[(NC + 2) * 2] * y) + s0;
    const double v1 = ((C[(NC + 0) * 2 + 1] + C[(NC + 1) * 2 + 1] * x) + C[(NC + 2) * 2 + 1] * y) + s1;
    return make_double2(v0, v1);
}

struct VfGeo { int lo0, cnt0, lo1, cnt1; };        // fast pixels: n0 in [lo0, lo0 + cnt0), n1 in [lo1, lo1 + cnt1)  (cnt 0: none)

// floor + 2^-32 fraction of both coordinates; true when all four taps lie inside the frame interior AND the staged box
__device__ __forceinline__ bool vf_geometry(double t0, double t1, const VfGeo &g, int &n0, int &n1, uint32_t &fa, uint32_t &fb) {
    const double m0 = t0 + VL_MAGIC, m1 = t1 + VL_MAGIC;
    n0 = __double2hiint(m0) - VL_MAGIC_HI; n1 = __double2hiint(m1) - VL_MAGIC_HI;
    fa = (uint32_t)__double2loint(m0); fb = (uint32_t)__double2loint(m1);
    return (unsigned)(n0 - g.lo0) < (unsigned)g.cnt0 && (unsigned)(n1 - g.lo1) < (unsigned)g.cnt1;
}

// resampling + composite of one thread's rows of a tile (P4)
template <bool HAS_FWD, bool BOX>
__device__ __forceinline__ void vf_strip(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                                         const float2 *__restrict__ fwd, const uint2 *__restrict__ box, int rmin, int cmin,
                                         int bw, const VfGeo geo, const double2 *__restrict__ Csj, int kr0,
                                         const vm_axis_entry *__restrict__ rp, const uint8_t *__restrict__ bgp,
                                         float4 *__restrict__ op, int nrows, int h, int w, int *outside, int *near_knife) {
    const int w3 = w * 3;
#pragma unroll 1
    for (int i = 0; i < nrows; ++i) {
        const vm_axis_entry re = rp[i];
        const double2 c0 = Csj[(re.i0 - kr0) * VF_CSP], c1 = Csj[(re.i1 - kr0) * VF_CSP];
        const double xf = re.frac, x1 = 1.0 - xf;
        const double t0 = fma(c1.x, xf, c0.x * x1), t1 = fma(c1.y, xf, c0.y * x1);
        int n0, n1;
        uint32_t fa, fb;
        const bool fast = vf_geometry(t0, t1, geo, n0, n1, fa, fb);
        uint2 e[4];
        if (BOX) {
            const int q = fast ? (n0 - rmin) * bw + (n1 - cmin) : 0;
            const uint2 *g0 = box + q, *g1 = g0 + (fast ? bw : 0);
            e[0] = g0[0]; e[1] = g0[1]; e[2] = g1[0]; e[3] = g1[1];
        } else if (fast) {
            e[0] = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, n0, n1);     e[1] = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, n0, n1 + 1);
            e[2] = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, n0 + 1, n1); e[3] = vf_src_elem<HAS_FWD>(fg, bwd, fwd, h, w, n0 + 1, n1 + 1);
        } else {
            e[0] = e[1] = e[2] = e[3] = make_uint2(0u, 0u);
        }
        const float b0 = (float)__ldcs(bgp), b1 = (float)__ldcs(bgp + 1), b2 = (float)__ldcs(bgp + 2);
        float4 o;
        float na;
        const unsigned unc = vl_blend<1>(e, fa, fb, fast, b0, b1, b2, o, na);
        if (unc) {
            // t in [n, n + 1]: anything with n <= -2 or n >= size is outside [0, size - 1] (map_coordinates -> 0)
            if (n0 <= -2 || n0 >= h || n1 <= -2 || n1 >= w) { o = make_float4(b0, b1, b2, 0.f); ++*outside; }
            else o = vf_exact_px<HAS_FWD>(fg, bwd, fwd, t0, t1, h, w, bgp, o, na, unc, outside, near_knife);
        }
        __stcs(op, o);                                                 // streamed once: evict first
        bgp += w3; op += w;
    }
}

// N > 0: control-point count known at compile time (unrolled spline loop), 0: run-time count
template <int N, bool HAS_FWD, int NSW>
__global__ void __launch_bounds__(VF_THREADS, 1)
k_fuse_c4(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd, const float2 *__restrict__ fwd,
          const uint8_t *__restrict__ bg, int n_bg, const double *__restrict__ ctrl, const double *__restrict__ coef,
          int n_rt, int nx, int ny, double step_x, double step_y, const vm_axis_entry *__restrict__ rows,
          const vm_axis_entry *__restrict__ cols, int n_frames, int h, int w, int tiles_x, int tiles_y,
          const double2 *__restrict__ gtab, float4 *__restrict__ out, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typedef VfSmem<NSW> Smem;
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    constexpr int NST = NSW * 32;                    // spline threads
    constexpr int NPT = VF_THREADS - NST;            // pixel threads
    constexpr int NSTRIPS = NPT / 64;                // strips of a tile in P4 (64 columns x RPT rows per strip)
    constexpr int RPT = (VF_TILE + NSTRIPS - 1) / NSTRIPS;
    static_assert(NSTRIPS * RPT >= VF_TILE && NPT % 64 == 0, "tile rows must be covered by the pixel threads");
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NC = N > 0 ? N : n_rt;
    const int tpf = tiles_x * tiles_y;
    const int64_t total = (int64_t)n_frames * tpf;
    const int64_t hw = (int64_t)h * w;

    // ---- log table -> shared memory (bulk async copy, one mbarrier), awaited by everybody --------------------
    const uint32_t bar = vl_smem_u32(&S.bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)VL_TAB_BYTES) : "memory");
        constexpr int PIECE = VL_TAB_BYTES / 8;
        for (int o = 0; o < VL_TAB_BYTES; o += PIECE)
            vl_bulk_g2s(vl_smem_u32(reinterpret_cast<unsigned char *>(S.tab) + o), reinterpret_cast<const unsigned char *>(gtab) + o,
                        (uint32_t)PIECE, bar);
    }
    __syncthreads();
    vl_mbar_wait_parity(bar, 0);
    const uint32_t tab_adj = vl_smem_u32(S.tab) - (uint32_t)(((1023 + VL_EMIN) << VL_BITS) << 4);

#ifdef VF_SPLINE_HIGH
    constexpr int SPL_BASE = NPT, PIX_BASE = 0;     // spline warps = the LAST warps of the CTA (the issue arbiter favours high warp ids)
#else
    constexpr int SPL_BASE = 0, PIX_BASE = NST;
#endif
    if (tid >= SPL_BASE && tid < SPL_BASE + NST) {
        const int stid = tid - SPL_BASE, swarp = stid >> 5;
        // =====================================================================================================
        // spline warps: T and the source box of tile it (buffer it & 1), one tile ahead of the pixel warps
        // =====================================================================================================
        int cur_frame = -1, it = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int b = it & 1;
            VF_T(ts0);
            if (it >= 2) vf_bar_sync(VF_BAR_EMPTY0 + b, VF_THREADS);   // the pixel warps are done with T[b] / info[b]
            VF_T(ts1);
            const int frame = (int)(t / tpf), rem = (int)(t - (int64_t)frame * tpf);
            const int ty = rem / tiles_x, tx = rem - ty * tiles_x;
            const int I0 = ty * VF_TILE, J0 = tx * VF_TILE;
            const int th = min(VF_TILE, h - I0), tw = min(VF_TILE, w - J0);
            const vm_axis_entry r0 = vm_ld_axis(rows + I0), r1 = vm_ld_axis(rows + I0 + th - 1);
            const vm_axis_entry c0 = vm_ld_axis(cols + J0), c1 = vm_ld_axis(cols + J0 + tw - 1);
            const int kr0 = r0.i0, kr1 = max(r1.i1, r1.i0), kc0 = c0.i0, kc1 = max(c1.i1, c1.i0);
            const int nkr = kr1 - kr0 + 1, nkc = kc1 - kc0 + 1;
            const bool valid = nkr >= 1 && nkr <= VF_CW && nkc >= 1 && nkc <= VF_CW && kr0 >= 0 && kr1 < nx && kc0 >= 0 && kc1 < ny;
            if (frame != cur_frame) {                                  // CTA-uniform among the spline warps
                vf_bar_sync(VF_BAR_SPLINE, NST);                       // nobody still reads the previous frame's controls
                const double *P = ctrl + (int64_t)frame * NC * 2;
                const double *C = coef + (int64_t)frame * (NC + 3) * 2;
                for (int a = stid; a < NC; a += NST) {
                    S.p[a] = make_double2(P[2 * a], P[2 * a + 1]);
                    S.wv[a] = make_double2(0.5 * C[2 * a], 0.5 * C[2 * a + 1]);
                }
                if (stid < 6) S.aff[stid] = C[(NC + stid % 3) * 2 + stid / 3];
                cur_frame = frame;
            }
            if (stid == 0) {
                VfTileInfo &I = S.info[b];
                I.frame = frame; I.I0 = I0; I.J0 = J0; I.th = th; I.tw = tw;
                I.kr0 = kr0; I.nkr = nkr; I.kc0 = kc0; I.nkc = nkc; I.valid = valid ? 1 : 0;
                I.rlo = INT_MAX; I.rhi = INT_MIN; I.clo = INT_MAX; I.chi = INT_MIN; I.bad = 0;
            }
            vf_bar_sync(VF_BAR_SPLINE, NST);
            if (valid) {
                const int kend = kr1;
                const int l = min(kc0 + lane, kc1);
                const double y = (double)l * step_y;
                // can every (point, control) pair of the window use the table?  lane a checks control a
                bool badc = false;
                {
                    const double xlo = (double)kr0 * step_x, xhi = (double)kr1 * step_x;
                    const double ylo = (double)kc0 * step_y, yhi = (double)kc1 * step_y;
                    for (int a = lane; a < NC; a += 32) {
                        const double2 p = S.p[a];
                        const double kn = fmin(fmax(rint(p.x / step_x), (double)kr0), (double)kr1);
                        const double ln = fmin(fmax(rint(p.y / step_y), (double)kc0), (double)kc1);
                        const double dxm = kn * step_x - p.x, dym = ln * step_y - p.y;
                        const double dxM = fmax(fabs(xlo - p.x), fabs(xhi - p.x)), dyM = fmax(fabs(ylo - p.y), fabs(yhi - p.y));
                        const double dmin2 = dxm * dxm + dym * dym, dmax2 = dxM * dxM + dyM * dyM;
                        if (!(dmin2 >= 0.015626) || !(dmax2 < 33550000.0)) badc = true;      // 2^-6 (1 + 6e-5), 2^25 (1 - 1e-4)
                    }
                }
                const bool slow = __any_sync(0xffffffffu, badc) || (N == 0);
                int rlo = INT_MAX, rhi = INT_MIN, clo = INT_MAX, chi = INT_MIN, bad = 0;
                for (int kg = kr0 + VF_RG * swarp; kg <= kend; kg += VF_RG * NSW) {
                    double v0[VF_RG], v1[VF_RG];
                    if (!slow) {
                        constexpr int NF = N > 0 ? N : 1;
                        __syncwarp();
                        for (int a = lane; a < NF; a += 32) {
                            const double px = S.p[a].x;
                            double d[VF_RG];
#pragma unroll
                            for (int r = 0; r < VF_RG; ++r) {
                                const double dx = (double)min(kg + r, kend) * step_x - px;
                                d[r] = dx * dx;
                            }
                            S.dx2[swarp][a] = make_double4(d[0], d[1], d[VF_RG > 2 ? 2 : 0], d[VF_RG > 2 ? 3 : 1]);
                        }
                        __syncwarp();
                        double s0[VF_RG], s1[VF_RG];
#pragma unroll
                        for (int r = 0; r < VF_RG; ++r) { s0[r] = 0.0; s1[r] = 0.0; }
#pragma unroll
                        for (int a = 0; a < NF; ++a) {
                            const double2 wv = S.wv[a];
                            const double4 dx2 = S.dx2[swarp][a];
                            const double dd[4] = {dx2.x, dx2.y, dx2.z, dx2.w};
                            const double dy = y - S.p[a].y;
                            const double dy2a = dy * dy;
#pragma unroll
                            for (int r = 0; r < VF_RG; ++r) {
                                const double U = vl_u_fast(dd[r] + dy2a, tab_adj);
                                s0[r] = fma(wv.x, U, s0[r]);
                                s1[r] = fma(wv.y, U, s1[r]);
                            }
                        }
#pragma unroll
                        for (int r = 0; r < VF_RG; ++r) {
                            const double x = (double)min(kg + r, kend) * step_x;
                            v0[r] = ((S.aff[0] + S.aff[1] * x) + S.aff[2] * y) + s0[r];
                            v1[r] = ((S.aff[3] + S.aff[4] * x) + S.aff[5] * y) + s1[r];
                        }
                    } else {
#pragma unroll 1
                        for (int r = 0; r < VF_RG; ++r) {
                            const double x = (double)min(kg + r, kend) * step_x;
                            double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
                            for (int a = 0; a < NC; ++a) {
                                const double2 p = S.p[a], wv = S.wv[a];
                                const double dx = x - p.x, dy = y - p.y;
                                const double U = vl_u_any(dx * dx + dy * dy, tab_adj);   // same roundings as the fast path's dx2 + dy2
                                s0 = fma(wv.x, U, s0);
                                s1 = fma(wv.y, U, s1);
                            }
                            v0[r] = ((S.aff[0] + S.aff[1] * x) + S.aff[2] * y) + s0;
                            v1[r] = ((S.aff[3] + S.aff[4] * x) + S.aff[5] * y) + s1;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < VF_RG; ++r) {
                        const int k = kg + r;
                        if (k <= kend && lane < nkc) {
                            S.T[b][(k - kr0) * VF_CW + lane] = make_double2(v0[r], v1[r]);
                            if (!(fabs(v0[r]) < 1.0e9) || !(fabs(v1[r]) < 1.0e9)) bad = 1;
                            else {
                                const int f0 = __double2int_rd(v0[r]), f1 = __double2int_rd(v1[r]);
                                rlo = min(rlo, f0); rhi = max(rhi, f0); clo = min(clo, f1); chi = max(chi, f1);
                            }
                        }
                    }
                }
                rlo = __reduce_min_sync(0xffffffffu, rlo); rhi = __reduce_max_sync(0xffffffffu, rhi);
                clo = __reduce_min_sync(0xffffffffu, clo); chi = __reduce_max_sync(0xffffffffu, chi);
                bad = __reduce_max_sync(0xffffffffu, bad);
                if (lane == 0) {
                    VfTileInfo &I = S.info[b];
                    atomicMin(&I.rlo, rlo); atomicMax(&I.rhi, rhi); atomicMin(&I.clo, clo); atomicMax(&I.chi, chi);
                    if (bad) atomicMax(&I.bad, 1);
                }
            }
            vf_bar_sync(VF_BAR_SPLINE, NST);
            if (stid == 0) {
                VfTileInfo &I = S.info[b];
                I.rmin = 0; I.bh = 0; I.cmin = 0; I.bw = 0;
                if (valid && !I.bad && I.rlo <= I.rhi) {
                    // rows [rmin, rmax] x columns [cmin, cmax] hold every tap of every in-frame fast pixel
                    const int rmin = max(I.rlo, 0), rmax = min(I.rhi + 1, h - 1);
                    const int cmin = max(I.clo, 0), cmax = min(I.chi + 1, w - 1);
                    const int bh = rmax - rmin + 1, bwid = cmax - cmin + 1;
                    if (bh >= 2 && bwid >= 2 && (int64_t)bh * bwid <= VF_BOX_CAP) { I.rmin = rmin; I.bh = bh; I.cmin = cmin; I.bw = bwid; }
                    else if (bh < 2 || bwid < 2) { I.rmin = 0; I.bh = 0; I.cmin = 0; I.bw = -1; }   // nothing of the source is in reach: no fast pixel
                }
            }
            __threadfence_block();
            vf_bar_arrive(VF_BAR_FULL0 + b, VF_THREADS);
#ifdef VF_TIMING
            if (stid == 0) { VF_T(ts2); VF_ACC(0, ts0, ts1); VF_ACC(1, ts1, ts2); VF_ACC(2, 0, 1); }
#endif
        }
    } else {
        // =====================================================================================================
        // pixel warps
        // =====================================================================================================
        const int tp = tid - PIX_BASE;                   // 0 .. NPT-1
        const int x = tp & 63, strip = tp >> 6;
        int outside = 0, slow_tiles = 0, flags = 0, near_knife = 0, bad_tiles = 0, it = 0;
        for (int64_t t = blockIdx.x; t < total; t += gridDim.x, ++it) {
            const int b = it & 1;
            VF_T(tp0);
            vf_bar_sync(VF_BAR_FULL0 + b, VF_THREADS);
            VF_T(tp1);
            const VfTileInfo I = S.info[b];
            const int frame = I.frame, I0 = I.I0, J0 = I.J0, th = I.th, tw = I.tw, kr0 = I.kr0, kc0 = I.kc0;
            const int kr1 = kr0 + I.nkr - 1, kc1 = kc0 + I.nkc - 1;
            const uint8_t *fgf = fg + frame * hw * 4;
            const float2 *bwf = bwd + frame * hw;
            const float2 *fwf = HAS_FWD ? fwd + frame * hw : nullptr;
            int bgi = frame;
            if (bgi >= n_bg) bgi %= n_bg;
            const uint8_t *bgf = bg + bgi * hw * 3;
            float4 *outf = out + frame * hw;

            // ---- P2: axis entries, column-interpolated coarse rows ---------------------------------------------
            const int jc = min(x, tw - 1);
            const vm_axis_entry ce = vm_ld_axis(cols + J0 + jc);
            bool ok = I.valid && ce.i0 >= kc0 && ce.i0 <= kc1 && ce.i1 >= kc0 && ce.i1 <= kc1;
            if (tp < th) {
                const vm_axis_entry myrow = vm_ld_axis(rows + I0 + tp);
                S.rows[tp] = myrow;
                ok = ok && myrow.i0 >= kr0 && myrow.i0 <= kr1 && myrow.i1 >= kr0 && myrow.i1 <= kr1;
            }
            const bool staged = vf_bar_and(VF_BAR_PIXEL, NPT, ok);
            if (staged) {
                const double yf = ce.frac, y1 = 1.0 - yf;
                const double2 *Ta = S.T[b] + (ce.i0 - kc0), *Tb = S.T[b] + (ce.i1 - kc0);
                for (int k = strip; k < I.nkr; k += NSTRIPS) {
                    const VlC c = vl_col_lerp(Ta[k * VF_CW], Tb[k * VF_CW], y1, yf);
                    S.Cs[k * VF_CSP + x] = make_double2(c.x, c.y);
                }
            }
            vf_bar_sync(VF_BAR_PIXEL, NPT);                            // Cs complete; T[b] and info[b] are free
            if (t + 2 * (int64_t)gridDim.x < total) vf_bar_arrive(VF_BAR_EMPTY0 + b, VF_THREADS);
            VF_T(tp2);

            const int strip0 = strip * RPT;
            const int nrows = min(RPT, th - strip0);
            const bool active = x < tw && nrows > 0;
            const int64_t p0 = (int64_t)(I0 + strip0) * w + J0 + x;

            if (!staged) {
                // ---- axis tables that are not monotone /2 windows: every pixel on its own (exact, slow) --------
                if (tp == 0) ++bad_tiles;
                const double *P = ctrl + (int64_t)frame * NC * 2;
                const double *C = coef + (int64_t)frame * (NC + 3) * 2;
                const VfGeo geo = {1, max(h - 2, 0), 1, max(w - 2, 0)};
                for (int q = tp; q < th * tw; q += NPT) {
                    const int ii = q / tw, jj = q - ii * tw;
                    const vm_axis_entry re = vm_ld_axis(rows + I0 + ii), cc = vm_ld_axis(cols + J0 + jj);
                    const double yf = cc.frac, y1 = 1.0 - yf, xf = re.frac, x1 = 1.0 - xf;
                    const VlC c0 = vl_col_lerp(vf_tps_point(P, C, NC, re.i0, cc.i0, step_x, step_y, tab_adj),
                                               vf_tps_point(P, C, NC, re.i0, cc.i1, step_x, step_y, tab_adj), y1, yf);
                    const VlC c1 = vl_col_lerp(vf_tps_point(P, C, NC, re.i1, cc.i0, step_x, step_y, tab_adj),
                                               vf_tps_point(P, C, NC, re.i1, cc.i1, step_x, step_y, tab_adj), y1, yf);
                    const double t0 = fma(c1.x, xf, c0.x * x1), t1 = fma(c1.y, xf, c0.y * x1);
                    int n0, n1;
                    uint32_t fa, fb;
                    const bool fast = vf_geometry(t0, t1, geo, n0, n1, fa, fb);
                    uint2 e[4];
                    if (fast) {
                        e[0] = vf_src_elem<HAS_FWD>(fgf, bwf, fwf, h, w, n0, n1);     e[1] = vf_src_elem<HAS_FWD>(fgf, bwf, fwf, h, w, n0, n1 + 1);
                        e[2] = vf_src_elem<HAS_FWD>(fgf, bwf, fwf, h, w, n0 + 1, n1); e[3] = vf_src_elem<HAS_FWD>(fgf, bwf, fwf, h, w, n0 + 1, n1 + 1);
                    } else {
                        e[0] = e[1] = e[2] = e[3] = make_uint2(0u, 0u);
                    }
                    const uint8_t *bp = bgf + ((int64_t)(I0 + ii) * w + J0 + jj) * 3;
                    const float b0 = (float)__ldcs(bp), b1 = (float)__ldcs(bp + 1), b2 = (float)__ldcs(bp + 2);
                    float4 o;
                    float na;
                    const unsigned unc = vl_blend<1>(e, fa, fb, fast, b0, b1, b2, o, na);
                    if (unc) {
                        if (n0 <= -2 || n0 >= h || n1 <= -2 || n1 >= w) { o = make_float4(b0, b1, b2, 0.f); ++outside; }
                        else o = vf_exact_px<HAS_FWD>(fgf, bwf, fwf, t0, t1, h, w, bp, o, na, unc, &outside, &near_knife);
                    }
                    __stcs(outf + (int64_t)(I0 + ii) * w + J0 + jj, o);
                }
                vf_bar_sync(VF_BAR_PIXEL, NPT);
                continue;
            }

            // ---- P3: flow warp + consistency mask of the source box -> shared memory -------------------------------
            const int bw = I.bw > 0 ? I.bw : 0, bh = I.bh, rmin = I.rmin, cmin = I.cmin;
            const bool boxed = bw > 0;
            if (boxed) {
                const int nbox = bw * bh;
                const int qstep = NPT / bw, mstep = NPT - qstep * bw;
                int r = tp / bw, c = tp - r * bw;
                const uint32_t *fg32 = reinterpret_cast<const uint32_t *>(fgf);
                float2 nxt = make_float2(0.f, 0.f);
                if (tp < nbox) nxt = __ldcs(bwf + (rmin + r) * w + cmin + c);
                for (int p = tp; p < nbox; p += NPT) {
                    const float2 fb = nxt;
                    const int i = rmin + r, j = cmin + c;
                    c += mstep; r += qstep;
                    if (c >= bw) { c -= bw; ++r; }
                    if (p + NPT < nbox) nxt = __ldcs(bwf + (rmin + r) * w + cmin + c);
                    const VmFlowPx px = vm_flow_px<HAS_FWD>(fg32, fwf, h, w, i, j, (float)i, (float)j, fb, flags);
                    S.box[p] = make_uint2(px.bgr, px.masked ? 0u : px.ta);
                }
            } else if (tp == 0 && I.bw == 0) {
                ++slow_tiles;                                           // box too large for shared memory: taps one by one
            }
            VF_T(tp3);
            vf_bar_sync(VF_BAR_PIXEL, NPT);
            VF_T(tp4);

            // ---- P4: resampling + composite --------------------------------------------------------------------
            if (active) {
                // fast pixels: all four taps inside the frame interior (n in [1, size - 2]) and inside the box
                VfGeo geo = {0, 0, 0, 0};                               // I.bw < 0: no source pixel in reach, nothing is fast
                if (boxed) {
                    geo.lo0 = max(rmin, 1); geo.cnt0 = max(min(rmin + bh - 1, h - 1) - geo.lo0, 0);
                    geo.lo1 = max(cmin, 1); geo.cnt1 = max(min(cmin + bw - 1, w - 1) - geo.lo1, 0);
                } else if (I.bw == 0) {
                    geo.lo0 = 1; geo.cnt0 = max(h - 2, 0); geo.lo1 = 1; geo.cnt1 = max(w - 2, 0);
                }
                const double2 *Csj = S.Cs + x;
                const vm_axis_entry *rp = S.rows + strip0;
                const uint8_t *bgp = bgf + p0 * 3;
                float4 *op = outf + p0;
                if (boxed) vf_strip<HAS_FWD, true>(fgf, bwf, fwf, S.box, rmin, cmin, bw, geo, Csj, kr0, rp, bgp, op, nrows, h, w, &outside, &near_knife);
                else       vf_strip<HAS_FWD, false>(fgf, bwf, fwf, S.box, 0, 0, 0, geo, Csj, kr0, rp, bgp, op, nrows, h, w, &outside, &near_knife);
            }
            VF_T(tp5);
            vf_bar_sync(VF_BAR_PIXEL, NPT);                            // box / Cs / rows are free for the next tile
#ifdef VF_TIMING
            if (tp == 0) { VF_T(tp6); VF_ACC(4, tp0, tp1); VF_ACC(5, tp1, tp2); VF_ACC(6, tp2, tp3); VF_ACC(7, tp3, tp4); VF_ACC(8, tp4, tp5); VF_ACC(9, tp5, tp6); VF_ACC(10, 0, 1); }
#endif
        }
        if (status) {
            if (outside) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside);
            if (slow_tiles) atomicAdd(status + VM_STATUS_SLOW_TILES, slow_tiles);
            if (bad_tiles) atomicAdd(status + VM_STATUS_BAD_TABLE, bad_tiles);
            if (near_knife) atomicAdd(status + VM_STATUS_NEAR_KNIFE, near_knife);
            if (HAS_FWD && flags) {
                if (flags & 1) atomicAdd(status + VM_STATUS_INDEX_ERR, 1);
                if (flags & 2) atomicAdd(status + VM_STATUS_NAN_ERR, 1);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
#ifdef VF_TIMING
// {spline: wait for a free buffer, compute, tiles, -, pixel: wait for T, P2, P3 own work, P3 barrier, P4 own work, P4 barrier, tiles}
extern "C" int vm_fuse_prof_read(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, g_vf_prof, sizeof(g_vf_prof)) != cudaSuccess) return VM_ERR_CUDA;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_vf_prof, z, sizeof(z)); }
    return VM_OK;
}
#endif

static std::atomic<long long> g_vf_launches{0};
extern "C" long long vm_fuse_launch_count(void) { return g_vf_launches.load(); }

static int vf_sm_count() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    return v;
}

template <int N, bool FW, int NSW>
static int vf_launch_t(const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg, int n_bg,
                       const double *ctrl, const double *coef, int Nrt, int nx, int ny, double step_x, double step_y,
                       const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w, float *out,
                       int32_t *status, cudaStream_t st, int ctas, const char *what) {
    const size_t smem = sizeof(VfSmem<NSW>);
    static bool attr_set[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_fuse_c4<N, FW, NSW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { vm_set_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e)); return VM_ERR_CUDA; }
        attr_set[dev & 63] = true;
    }
    const int tiles_x = (w + VF_TILE - 1) / VF_TILE, tiles_y = (h + VF_TILE - 1) / VF_TILE;
    const int64_t total = (int64_t)n * tiles_x * tiles_y;
    const int grid = (int)(total < ctas ? total : ctas);
    k_fuse_c4<N, FW, NSW><<<grid, VF_THREADS, smem, st>>>(fg, (const float2 *)backward, (const float2 *)forward, bg, n_bg, ctrl, coef, Nrt,
                                                          nx, ny, step_x, step_y, rows, cols, n, h, w, tiles_x, tiles_y,
                                                          (const double2 *)vl_table_device(), (float4 *)out, status);
    g_vf_launches += 1;
    return vm_check_launch(what);
}

int g_vf_ctas = 0;           // CTAs (0: one per SM)

int vm_fuse_set_option(const char *key, int value) {
    if (!strcmp(key, "fuse_ctas") && value >= 0 && value <= 4096) { g_vf_ctas = value; return VM_OK; }
    return VM_ERR_ARG;
}

// mode 1: flow warp only, 2: flow warp + consistency mask
int vm_fuse_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, int32_t *status, cudaStream_t st, const char *what) {
    int rc = vl_table_init();
    if (rc != VM_OK) return rc;
    VM_REQUIRE(N >= 1 && N <= VL_MAX_N, "control point count out of range");
    VM_REQUIRE(h >= 2 && w >= 2 && (int64_t)h * w < (1ll << 28), "frame size out of range");
    const int ctas = g_vf_ctas > 0 ? g_vf_ctas : vf_sm_count();
    const bool fw = mode == 2 && forward;
#define VF_GO(NN, FW, NSW) return vf_launch_t<NN, FW, NSW>(fg, backward, forward, bg, n_bg, ctrl, coef, N, nx, ny, step_x, step_y, rows, cols, n, h, w, out, status, st, ctas, what)
#define VF_N(FW, NSW) do { if (N == 25) VF_GO(25, FW, NSW); else if (N == 16) VF_GO(16, FW, NSW); else VF_GO(0, FW, NSW); } while (0)
    if (fw) VF_N(true, 8); else VF_N(false, 8);
#undef VF_N
#undef VF_GO
    return VM_OK;
}
