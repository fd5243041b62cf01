// Lean split pipeline for the fused C3 / C4 entry points (fused_variant 4, the default).
//
// Per chunk of frames (sized so that the two intermediates stay in the 126 MB L2):
//
//   stage A   k_flow_warp_mask_bgra<.., 2> (vm_flow.cu)   flow warp + consistency mask
//             -> packed (n,h,w) uint2 {B|G<<8|R<<16, TA}, alpha = TA / 261120     [C4 only]
//   stage B1  k_lean_coarse   thin-plate spline on the coarse grid (tps.py:101-123), float64,
//             -> T (n,nx,ny) double2 {row coordinate, column coordinate}
//   stage B2  k_lean_fine     bilinear up-sampling of T (tps.py:55-74), map_coordinates
//             (tps.py:34) on the packed intermediate / the BGRA frame, composite (reader.py:72-79)
//             -> out (n,h,w) float4 {B, G, R, alpha'}
//
// B1 is bound by the float64 pipe: 10 DP instructions per (coarse point, control point) - the
// log comes from a 31-octave x 64-entry {1/c, log c} table staged in shared memory by a bulk
// async copy (TMA, 32 KB per persistent CTA) plus a degree-4 minimax polynomial.  B2 is bound
// by instruction issue: coordinates are the only float64 work (6 DP per pixel thanks to
// rolling column-interpolated coarse rows); weights, colour blending and alpha are integer
// fixed point (2^-30 weights), with an exact float64 re-evaluation of the rare samples whose
// rounding the fixed-point value cannot decide.
#include "vm_lean.cuh"
#include <cuda.h>
#include <string.h>
#include <atomic>
#include <mutex>

int vm_launch_flow_stage(const uint8_t *fg, const float *backward, const float *forward, int n, int h, int w,
                         void *packed, int32_t *status, cudaStream_t st, bool raw_ta);     // vm_flow.cu

int g_vl_chunk = 0;          // frames per A/B1/B2 round; 0: as many as make ~64 frames of 1080p (long launches: no ramp/tail losses)
static int vl_chunk_for(int h, int w) {
    if (g_vl_chunk > 0) return g_vl_chunk;
    const int64_t px = (int64_t)h * w;
    int64_t c = (64ll * 1080 * 1920 + px / 2) / (px > 0 ? px : 1);
    return (int)(c < 8 ? 8 : (c > 1024 ? 1024 : c));
}
int g_vl_rb = 0;             // coarse rows per B1 unit (0 = pick on the host)
int g_vl_fine_rows = 8;      // fine rows per B2 thread
int g_vl_timing = 0;         // 1: bracket the stages of the first chunk of every call with CUDA events (vm_lean_stage_ms)
static cudaEvent_t g_vl_tev[64][5];
static int g_vl_tpair[64][4][2];                  // event indices bracketing {spline, tile boxes, flow stage, resampling}
static bool g_vl_tev_ok[64];
static std::atomic<long long> g_vl_launches{0};   // kernels launched by this library's lean path (bench.py "gpu_launches")
int g_vl_sub = 0;            // frames per flow-stage / resampling sub-round inside a round (0: the whole round)
int g_vl_box_cap = 0;        // source-box entries per B2 tile (0: as many as the occupancy target allows)
int g_vl_floors = 1;         // 1: the spline stage also writes packed int16 floors of T for the tile-box stage
int g_vl_minb = 4;           // B2 occupancy target (CTAs of 256 threads per SM: 2, 3 or 4)
int g_vl_tmap = 1;           // B2 staging: 1 = two tensor bulk copies per tile (when available), 0 = one bulk copy per row
int g_va_minb = 12;          // k_aug_tps: CTAs of 128 threads per SM: 12 (40 registers, 60 bytes of spills, 48 warps per SM) measured 4.22 ms per
                             // augment_clip call of 64 x 1080p against 4.32-4.47 at 8 (64 registers), 4.33 at 10, 4.76 at 16: the gathers want warps

__device__ double2 g_vl_tab[VL_TAB_N];
// -DVL_TIMING: cycles per phase of the resampling CTAs, accumulated by thread 0 and by the last thread of every CTA
// (scripts/fine_phases.py); not part of the product build
#ifdef VL_TIMING
__constant__ int c_vl_abl;        // timing-only ablation (wrong output): 1 no source-box copies, 2 no background copies, 4 no transform loads, 8 no store
__device__ unsigned long long g_vl_prof[16];
#define VL_T(i) do { if (vl_tim) { const long long t_ = clock64(); vl_acc[i] += (unsigned long long)(t_ - vl_t0); vl_t0 = t_; } } while (0)
extern "C" int vm_lean_prof_read(unsigned long long *out16, int reset) {
    if (cudaMemcpyFromSymbol(out16, g_vl_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return VM_ERR_CUDA;
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_vl_prof, z, sizeof(z)); }
    return VM_OK;
}
#define VL_ABL(b) (c_vl_abl & (b))
#else
#define VL_T(i) do { } while (0)
#define VL_ABL(b) false
#endif

static std::mutex g_vl_mu;
static bool g_vl_done[64];

static int vl_init(int *dev_out) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
        vm_set_error("vm_lean: cudaGetDevice failed");
        return VM_ERR_CUDA;
    }
    *dev_out = dev;
    std::lock_guard<std::mutex> lk(g_vl_mu);
    if (g_vl_done[dev]) return VM_OK;
    static double2 tab[VL_TAB_N];
    for (int e = VL_EMIN; e < VL_EMAX; ++e)
        for (int k = 0; k < (1 << VL_BITS); ++k) {
            const long double c = 1.0L + ((long double)k + 0.5L) / (long double)(1 << VL_BITS);
            const double inv = (double)(1.0L / c);
            double2 &t = tab[((e - VL_EMIN) << VL_BITS) + k];
            t.x = ldexp(inv, -e);
            t.y = (double)((long double)e * 0.693147180559945309417232121458L - logl((long double)inv));
        }
    const cudaError_t err = cudaMemcpyToSymbol(g_vl_tab, tab, sizeof(tab));
    if (err != cudaSuccess) {
        vm_set_error("vm_lean: init: %s", cudaGetErrorString(err));
        return VM_ERR_CUDA;
    }
    g_vl_done[dev] = true;
    return VM_OK;
}

// the log table for other translation units (vm_fuse.cu): make sure it is uploaded / return its device address
int vl_table_init() { int dev = 0; return vl_init(&dev); }
const void *vl_table_device() {
    static void *addr[64];
    int dev = 0;
    cudaGetDevice(&dev);
    void *&p = addr[dev & 63];
    if (!p) cudaGetSymbolAddress(&p, g_vl_tab);
    return p;
}

// ---------------------------------------------------------------------------------------
// stage B1
// ---------------------------------------------------------------------------------------
#define VL_B1_WARPS 16
#define VL_B1_THREADS (VL_B1_WARPS * 32)
#define VL_B1_WARPS_HI 24
#define VL_B1_THREADS_HI (VL_B1_WARPS_HI * 32)
#ifndef VL_RG
#define VL_RG 4                                    // coarse rows evaluated together (4 or 8)
#endif

struct __align__(16) VlWarpSmem {
    double2 p[VL_MAX_N];                           // control point {x (row axis), y (column axis)}
    double2 wv[VL_MAX_N];                          // {w0 / 2, w1 / 2}
    double4 dx2[VL_RG / 4][VL_MAX_N];              // (x_r - Px)^2 for the VL_RG rows of the group
    double aff[8];
};

struct __align__(16) VlCoarseSmem {
    double2 tab[VL_TAB_N];
    unsigned long long bar;
    unsigned long long pad;
    VlWarpSmem w[1];                               // one slice per launched warp (dynamic shared memory)
};
static inline size_t vl_coarse_smem_bytes(int warps) { return sizeof(VlCoarseSmem) + (size_t)(warps - 1) * sizeof(VlWarpSmem); }

// floor of both coordinates of a coarse point as two int16 (what the tile-box stage needs: a quarter of the
// bytes of T); 0x8000 in both halves marks a non-finite / absurd value (the tile then takes the gather path)
#define VL_FLOOR_BAD ((int)0x80008000)
__device__ __forceinline__ int vl_pack_floor(double v0, double v1) {
    if (!(fabs(v0) < 1.0e9) || !(fabs(v1) < 1.0e9)) return VL_FLOOR_BAD;
    // clamped to [-32767, 32766]: 0x7FFF stays free for the box stage's "no point" filler, 0x8000 for VL_FLOOR_BAD
    const int f0 = max(-32767, min(32766, __double2int_rd(v0))), f1 = max(-32767, min(32766, __double2int_rd(v1)));
    return (f0 & 0xFFFF) | (f1 << 16);
}

// N > 0: number of control points known at compile time; 0: run-time count.
// DYR: (y - Py)^2 of every control point held in registers (16 warps per SM at 128 registers);
// otherwise it is recomputed per row group and the kernel runs 24 warps per SM.
template <int N, bool DYR>
__global__ void __launch_bounds__(DYR ? VL_B1_THREADS : VL_B1_THREADS_HI, 1)
k_lean_coarse(const double *__restrict__ ctrl, const double *__restrict__ coef, int n_rt, int n_frames, int nx, int ny,
              double step_x, double step_y, int rb, int nbands, int ncb, double2 *__restrict__ T,
              int *__restrict__ F, unsigned int *__restrict__ next_unit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VlCoarseSmem &S = *reinterpret_cast<VlCoarseSmem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NC = N > 0 ? N : n_rt;

    // log table -> shared memory with one bulk async copy per 16 KB piece, all on one mbarrier
    const uint32_t bar = vl_smem_u32(&S.bar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)VL_TAB_BYTES) : "memory");
        constexpr int PIECE = VL_TAB_BYTES / 8;
        static_assert(VL_TAB_BYTES % PIECE == 0 && PIECE % 16 == 0, "bulk copy pieces");
        for (int o = 0; o < VL_TAB_BYTES; o += PIECE)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(vl_smem_u32(reinterpret_cast<unsigned char *>(S.tab) + o)),
                           "l"(reinterpret_cast<const unsigned char *>(g_vl_tab) + o), "r"((uint32_t)PIECE), "r"(bar)
                         : "memory");
    }

    VlWarpSmem &Ws = S.w[warp];
    const uint32_t tab_adj = vl_smem_u32(S.tab) - (uint32_t)(((1023 + VL_EMIN) << VL_BITS) << 4);
    const int units_per_frame = nbands * ncb;
    const int units = n_frames * units_per_frame;
    int cur_frame = -1;
    bool tab_ready = false;

    // dynamic unit queue: unit costs differ (table conflicts near control points, generic-path units)
    for (;;) {
        int u = 0;
        if (lane == 0) u = (int)atomicAdd(next_unit, 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= units) break;
        const int frame = u / units_per_frame;
        const int rem = u - frame * units_per_frame;
        const int band = rem / ncb, cb = rem - band * ncb;
        const int k0 = band * rb, kend = min(k0 + rb, nx) - 1;        // rows of the unit
        const int l0 = cb * 32, lend = min(l0 + 32, ny) - 1;
        const int l = min(l0 + lane, ny - 1);
        const double y = (double)l * step_y;

        if (frame != cur_frame) {                                     // warp-uniform
            __syncwarp();
            const double *P = ctrl + (int64_t)frame * NC * 2;
            const double *C = coef + (int64_t)frame * (NC + 3) * 2;
            for (int a = lane; a < NC; a += 32) {
                Ws.p[a] = make_double2(P[2 * a], P[2 * a + 1]);
                Ws.wv[a] = make_double2(0.5 * C[2 * a], 0.5 * C[2 * a + 1]);
            }
            if (lane < 6) Ws.aff[lane] = C[(NC + lane % 3) * 2 + lane / 3];
            cur_frame = frame;
            __syncwarp();
        }

        // can every (point, control) pair of the unit use the table?  lane a checks control a
        bool bad = false;
        {
            const double xlo = (double)k0 * step_x, xhi = (double)kend * step_x;
            const double ylo = (double)l0 * step_y, yhi = (double)lend * step_y;
            for (int a = lane; a < NC; a += 32) {
                const double2 p = Ws.p[a];
                const double kn = fmin(fmax(rint(p.x / step_x), (double)k0), (double)kend);
                const double ln = fmin(fmax(rint(p.y / step_y), (double)l0), (double)lend);
                const double dxm = kn * step_x - p.x, dym = ln * step_y - p.y;
                const double dxM = fmax(fabs(xlo - p.x), fabs(xhi - p.x)), dyM = fmax(fabs(ylo - p.y), fabs(yhi - p.y));
                const double dmin2 = dxm * dxm + dym * dym, dmax2 = dxM * dxM + dyM * dyM;
                if (!(dmin2 >= 0.015626) || !(dmax2 < 33550000.0)) bad = true;      // 2^-6 (1 + 6e-5), 2^25 (1 - 1e-4)
            }
        }
        const bool slow = __any_sync(0xffffffffu, bad) || (N == 0);

        if (!tab_ready) {                                             // first unit: wait for the table
            uint32_t ok = 0;
            while (!ok)
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                             : "=r"(ok) : "r"(bar) : "memory");
            tab_ready = true;
        }

        if (!slow) {
            // ---- fast path: dy^2 per control in registers, VL_RG rows per step ----------------
            constexpr int NF = N > 0 ? N : 1;
            double dy2[DYR ? NF : 1];
            if (DYR) {
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    const double dy = y - Ws.p[a].y;
                    dy2[a] = dy * dy;
                }
            }
            for (int kg = k0; kg <= kend; kg += VL_RG) {
                __syncwarp();
                for (int a = lane; a < NF; a += 32) {
                    const double px = Ws.p[a].x;
                    double d[VL_RG];
#pragma unroll
                    for (int r = 0; r < VL_RG; ++r) {
                        const double dx = (double)min(kg + r, kend) * step_x - px;
                        d[r] = dx * dx;
                    }
#pragma unroll
                    for (int g = 0; g < VL_RG / 4; ++g) Ws.dx2[g][a] = make_double4(d[4 * g], d[4 * g + 1], d[4 * g + 2], d[4 * g + 3]);
                }
                __syncwarp();
                double s0[VL_RG], s1[VL_RG];
#pragma unroll
                for (int r = 0; r < VL_RG; ++r) { s0[r] = 0.0; s1[r] = 0.0; }
#pragma unroll
                for (int a = 0; a < NF; ++a) {
                    const double2 wv = Ws.wv[a];
                    double dd[VL_RG];
#pragma unroll
                    for (int g = 0; g < VL_RG / 4; ++g) {
                        const double4 dx2 = Ws.dx2[g][a];
                        dd[4 * g] = dx2.x; dd[4 * g + 1] = dx2.y; dd[4 * g + 2] = dx2.z; dd[4 * g + 3] = dx2.w;
                    }
                    double dy2a;
                    if (DYR) dy2a = dy2[a];
                    else { const double dy = y - Ws.p[a].y; dy2a = dy * dy; }
#pragma unroll
                    for (int r = 0; r < VL_RG; ++r) {
                        const double U = vl_u_fast(dd[r] + dy2a, tab_adj);
                        s0[r] = fma(wv.x, U, s0[r]);
                        s1[r] = fma(wv.y, U, s1[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < VL_RG; ++r) {
                    const int k = kg + r;
                    if (k <= kend && l0 + lane < ny) {
                        const double x = (double)k * step_x;
                        const double v0 = ((Ws.aff[0] + Ws.aff[1] * x) + Ws.aff[2] * y) + s0[r];
                        const double v1 = ((Ws.aff[3] + Ws.aff[4] * x) + Ws.aff[5] * y) + s1[r];
                        T[((int64_t)frame * nx + k) * ny + l] = make_double2(v0, v1);
                        if (F) F[((int64_t)frame * nx + k) * ny + l] = vl_pack_floor(v0, v1);
                    }
                }
            }
        } else {
            // ---- generic path: any control count, any distance -------------------------------
            for (int k = k0; k <= kend; ++k) {
                const double x = (double)k * step_x;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll 1
                for (int a = 0; a < NC; ++a) {
                    const double2 p = Ws.p[a], wv = Ws.wv[a];
                    const double dx = x - p.x, dy = y - p.y;
                    const double U = vl_u_any(dx * dx + dy * dy, tab_adj);   // same roundings as the fast path's dx2 + dy2
                    s0 = fma(wv.x, U, s0);
                    s1 = fma(wv.y, U, s1);
                }
                if (l0 + lane < ny) {
                    const double v0 = ((Ws.aff[0] + Ws.aff[1] * x) + Ws.aff[2] * y) + s0;
                    const double v1 = ((Ws.aff[3] + Ws.aff[4] * x) + Ws.aff[5] * y) + s1;
                    T[((int64_t)frame * nx + k) * ny + l] = make_double2(v0, v1);
                    if (F) F[((int64_t)frame * nx + k) * ny + l] = vl_pack_floor(v0, v1);
                }
            }
        }
    }
    if (!tab_ready && tid == 0) {                                      // never leave a bulk copy in flight
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar) : "memory");
    }
    __syncthreads();
}

// rows per unit: minimise (waves of units over the persistent warps) x (rows + set-up cost)
int g_vl_b1_warps = VL_B1_WARPS;
int g_vl_b1_ctas = 0;        // persistent CTAs of the spline stage (0: one per SM)
int g_vl_b1_dyr = 1;         // 1: dy^2 in registers / <= 16 warps per SM; 0: recomputed / <= 24 warps per SM

static int vl_pick_rb(int n, int nx, int ny, int ctas) {
    if (g_vl_rb > 0) return (g_vl_rb + VL_RG - 1) / VL_RG * VL_RG;
    const int ncb = (ny + 31) / 32;
    const int64_t warps = (int64_t)ctas * g_vl_b1_warps;
    int best = VL_RG;
    double best_cost = 1e300;
    for (int rb = VL_RG; rb <= 32; rb += VL_RG) {
        const int64_t units = (int64_t)n * ((nx + rb - 1) / rb) * ncb;
        const double cost = (double)((units + warps - 1) / warps) * (rb + 1.0);
        if (cost < best_cost - 1e-9) { best_cost = cost; best = rb; }
    }
    return best;
}

static int vl_sm_count() {
    static int sms[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!sms[dev]) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

static int vl_launch_coarse(const double *ctrl, const double *coef, int N, int n, int nx, int ny, double step_x,
                            double step_y, double2 *T, int *F, unsigned int *counter, cudaStream_t st) {
    if (cudaMemsetAsync(counter, 0, sizeof(unsigned int), st) != cudaSuccess) {
        vm_set_error("vm_lean: cudaMemsetAsync failed");
        return VM_ERR_CUDA;
    }
    const int ctas = g_vl_b1_ctas > 0 ? g_vl_b1_ctas : vl_sm_count();
    const int warps_cta = g_vl_b1_dyr ? (g_vl_b1_warps < VL_B1_WARPS ? g_vl_b1_warps : VL_B1_WARPS) : g_vl_b1_warps;
    const int rb = vl_pick_rb(n, nx, ny, ctas);
    const int nbands = (nx + rb - 1) / rb, ncb = (ny + 31) / 32;
    const size_t smem = vl_coarse_smem_bytes(VL_B1_WARPS_HI);             // attribute: the largest launch
#define VL_LAUNCH(NN, DY)                                                                                           \
    do {                                                                                                          \
        static bool attr_set[64];                                                                                 \
        int dev = 0;                                                                                              \
        cudaGetDevice(&dev);                                                                                      \
        if (!attr_set[dev & 63]) {                                                                                \
            cudaError_t e = cudaFuncSetAttribute(k_lean_coarse<NN, DY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) { vm_set_error("vm_lean: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VM_ERR_CUDA; } \
            attr_set[dev & 63] = true;                                                                            \
        }                                                                                                         \
        k_lean_coarse<NN, DY><<<ctas, warps_cta * 32, vl_coarse_smem_bytes(warps_cta), st>>>(ctrl, coef, N, n, nx, ny, step_x, step_y, rb, nbands, ncb, T, F, counter); \
    } while (0)
    if (g_vl_b1_dyr) {
        switch (N) {
        case 16: VL_LAUNCH(16, true); break;
        case 25: VL_LAUNCH(25, true); break;
        default: VL_LAUNCH(0, true); break;
        }
    } else {
        switch (N) {
        case 16: VL_LAUNCH(16, false); break;
        case 25: VL_LAUNCH(25, false); break;
        default: VL_LAUNCH(0, false); break;
        }
    }
#undef VL_LAUNCH
    return vm_check_launch("vm_lean coarse stage");
}

// ---------------------------------------------------------------------------------------
// stage B2
//
// CTA = VL_FW columns x (VL_FS * rpt) rows, thread = one column of one strip of rpt rows.
//   P0  axis entries of the tile -> registers / shared memory
//   P1  bulk async copies (TMA): the coarse transform window of the tile and its background rows
//   P2  bounding box of the window (= bounding box of every fine coordinate: the fine transform
//       is a convex combination of coarse values) and the column-interpolated coarse rows Cs
//   P3  bulk async copies: the source rows of the box (packed stage-A pixels or BGRA)
//   P4  per pixel: row interpolation of Cs (float64), fixed-point weights, integer blend of the
//       four taps read from shared memory, composite, one 16-byte streaming store
// Tiles whose box does not fit (strongly stretched grids) gather the taps from global memory;
// plans whose axis tables are not monotone windows take everything from global memory.
// ---------------------------------------------------------------------------------------
#ifndef VL_FW
#define VL_FW 64                                   // columns per CTA
#endif
#define VL_BG ((VL_FW / 2 + 4 + 31) / 32)          // 32-lane groups covering the coarse columns of a tile
#ifndef VL_FS
#define VL_FS 4                                    // row strips per CTA (256 threads; -DVL_FS=2: 128-thread CTAs, 64 x 16 tiles)
#endif
#define VL_FROWS_MAX (VL_FS * 8)                   // fine rows per CTA (VL_FS * rpt)
#define VL_TC (VL_FW / 2 + 4)                      // coarse columns staged per CTA (36)
#define VL_TR (VL_FROWS_MAX / 2 + 4)               // coarse rows staged per CTA (20)
#define VL_BOX_MAX 5632                            // source-box entries staged per CTA (8 B each)

struct __align__(128) VlFineSmem {                 // (128: destination alignment of tensor bulk copies - Cs, bgt and box qualify)
    double2 Cs[VL_TR * VL_FW];                     // coarse rows interpolated at the tile's fine columns
    unsigned char bgt[VL_FROWS_MAX * VL_FW * 3];
    vm_axis_entry rows[VL_FROWS_MAX];
    unsigned long long bar[2];
    alignas(128) uint2 box[1];                     // box_cap entries (dynamic): source rows [rmin, rmin + bh) x [cmin, cmin + bw)
};

// Tensor-map staging (default when the driver provides cuTensorMapEncodeTiled and the planes qualify): the source box of a
// tile and its background rows arrive through TWO tensor bulk copies issued by one thread instead of one 1-D bulk copy per
// row (~72 per tile).  A bulk-copy instruction is executed lane by lane (operands through uniform registers: 8 issue slots
// per copy, 10 % of the kernel's instructions in the ncu source view) and the TMA unit takes ~35 cycles per small copy.  A
// tensor copy moves a box of FIXED size, so the tile-box stage picks, per tile, the first of VL_NSHAPE shapes that holds the
// tile's bounding box (bw x bh entries <= the shared-memory budget); taps outside the frame are zero-filled by the hardware
// and never read.
#define VL_NSHAPE 12
struct alignas(64) VlTmaps {
    CUtensorMap box[VL_NSHAPE];                    // (frames, h, w) elements of the source, box {bw[s], bh[s], 1}
    CUtensorMap bg;                                // (n_bg, h, 3 w) bytes, box {3 VL_FW, VL_FROWS_MAX, 1}
    CUtensorMap T;                                 // (frames, nx, 2 ny) doubles, box {2 VL_TCP, VL_TR, 1}
    int bw[VL_NSHAPE], bh[VL_NSHAPE];
    int on;                                        // 0: one bulk copy per row (fallback)
};
struct VlShapes { int bw[VL_NSHAPE], bh[VL_NSHAPE], on; };
#define VL_TCP 40                                  // columns of the raw coarse window in shared memory (>= VL_TC, 16-byte elements)
struct __align__(128) VlFineSmemT {                // tensor-map kernel: the raw window instead of Cs leaves 8 KB more for the box
    double2 Traw[VL_TR * VL_TCP];
    unsigned char bgt[VL_FROWS_MAX * VL_FW * 3];
    vm_axis_entry rows[VL_FROWS_MAX];
    unsigned long long bar[2];
    alignas(128) uint2 box[1];
};
static inline size_t vl_fine_smem_bytes_t(int box_cap) { return sizeof(VlFineSmemT) + (size_t)(box_cap - 1) * sizeof(uint2); }
static inline size_t vl_fine_smem_bytes(int box_cap) { return sizeof(VlFineSmem) + (size_t)(box_cap - 1) * sizeof(uint2); }

// per-tile record written by k_lean_boxes: source box of the tile (bw = 0: does not fit / not usable)
struct __align__(16) VlTileBox { int rmin, bh, cmin, bw; };

// ---- P4, tile path: Cs in shared memory; taps from the staged box (BOX) or from global memory ----
// RAWT: Csj points at the thread's first column of the RAW coarse window (rows VL_TCP apart); the column interpolation
// (weights y1, yf between the columns Csj[.] and Csj[. + dq]) is done here, per pixel, with the arithmetic of vl_col_lerp -
// the same bits as the pre-interpolated Cs, which that variant does not have to build.
template <int SRC, bool BOX, bool BGSM, bool RAWT = false>
__device__ __forceinline__ void vl_strip_tile(const typename VlSrc<SRC>::elem *__restrict__ src,
                                              const typename VlSrc<SRC>::elem *__restrict__ box, int rmin, int cmin, int bw,
                                              const double2 *__restrict__ Csj, int kr0, const vm_axis_entry *__restrict__ rp,
                                              const unsigned char *__restrict__ bgl, const uint8_t *__restrict__ bgp,
                                              float4 *__restrict__ op, int nrows, int h, int w, int *outside,
                                              int dq = 0, double y1 = 0.0, double yf = 0.0) {
    typedef typename VlSrc<SRC>::elem elem;
    const int w3 = w * 3;
    constexpr int pitch = VL_FW * 3;
#pragma unroll 2
    for (int i = 0; i < nrows; ++i) {
        const vm_axis_entry re = rp[i];
        double2 c0, c1;
        if (RAWT) {
            const double2 *r0 = Csj + (re.i0 - kr0) * VL_TCP, *r1 = Csj + (re.i1 - kr0) * VL_TCP;
            const VlC u0 = vl_col_lerp(r0[0], r0[dq], y1, yf), u1 = vl_col_lerp(r1[0], r1[dq], y1, yf);
            c0 = make_double2(u0.x, u0.y); c1 = make_double2(u1.x, u1.y);
        } else {
            c0 = Csj[(re.i0 - kr0) * VL_FW]; c1 = Csj[(re.i1 - kr0) * VL_FW];
        }
        const double xf = re.frac, x1 = 1.0 - xf;
        const double t0 = fma(c1.x, xf, c0.x * x1), t1 = fma(c1.y, xf, c0.y * x1);
        int n0, n1;
        uint32_t fa, fb;
        const bool fast = vl_geometry(t0, t1, h, w, n0, n1, fa, fb);
        uint2 e[4];
        if (BOX) {
            const int q = fast ? (n0 - rmin) * bw + (n1 - cmin) : 0;
            const elem *g0 = box + q, *g1 = g0 + (fast ? bw : 0);
            e[0] = VlSrc<SRC>::lds(g0); e[1] = VlSrc<SRC>::lds(g0 + 1); e[2] = VlSrc<SRC>::lds(g1); e[3] = VlSrc<SRC>::lds(g1 + 1);
        } else {
            const int q = fast ? n0 * w + n1 : 0;
            const elem *g0 = src + q, *g1 = g0 + (fast ? w : 0);
            e[0] = VlSrc<SRC>::ld(g0); e[1] = VlSrc<SRC>::ld(g0 + 1); e[2] = VlSrc<SRC>::ld(g1); e[3] = VlSrc<SRC>::ld(g1 + 1);
        }
        float b0, b1, b2;
        if (BGSM) { b0 = (float)bgl[0]; b1 = (float)bgl[1]; b2 = (float)bgl[2]; }
        else { b0 = (float)__ldcs(bgp); b1 = (float)__ldcs(bgp + 1); b2 = (float)__ldcs(bgp + 2); }
        float4 o;
        float na;
        const unsigned unc = vl_blend<SRC>(e, fa, fb, fast, b0, b1, b2, o, na);
        if (unc) {
            // t in [n, n + 1]: anything with n <= -2 or n >= size is outside [0, size - 1] (map_coordinates -> 0)
            if (n0 <= -2 || n0 >= h || n1 <= -2 || n1 >= w) { o = make_float4(b0, b1, b2, 0.f); ++*outside; }
            else o = vl_exact_px<SRC>(src, t0, t1, h, w, bgp, o, na, unc, outside);
        }
        if (!VL_ABL(8) || o.x == 1.2345e33f) __stcs(op, o);             // streamed once: evict first
        bgl += pitch; bgp += w3; op += w;
    }
}

// ---- P4, generic path: rolling column-interpolated rows straight from global memory ----------
template <int SRC>
__device__ __noinline__ void vl_strip_generic(const typename VlSrc<SRC>::elem *__restrict__ src,
                                              const double2 *__restrict__ Ta, const double2 *__restrict__ Tb, int ny,
                                              const vm_axis_entry *__restrict__ rp, double yf,
                                              const uint8_t *__restrict__ bgp, float4 *__restrict__ op, int nrows,
                                              int h, int w, int *outside) {
    typedef typename VlSrc<SRC>::elem elem;
    const double y1 = 1.0 - yf;
    const int w3 = w * 3;
    int k0 = -1, k1 = -1;
    VlC C0 = {0.0, 0.0}, C1 = {0.0, 0.0};
    for (int i = 0; i < nrows; ++i) {
        const vm_axis_entry re = vm_ld_axis(rp + i);
        if (re.i0 != k0) {
            if (re.i0 == k1) C0 = C1; else C0 = vl_col_lerp(__ldg(Ta + re.i0 * ny), __ldg(Tb + re.i0 * ny), y1, yf);
            k0 = re.i0;
        }
        if (re.i1 != k1) {
            if (re.i1 == k0) C1 = C0; else C1 = vl_col_lerp(__ldg(Ta + re.i1 * ny), __ldg(Tb + re.i1 * ny), y1, yf);
            k1 = re.i1;
        }
        const double xf = re.frac, x1 = 1.0 - xf;
        const double t0 = fma(C1.x, xf, C0.x * x1), t1 = fma(C1.y, xf, C0.y * x1);
        int n0, n1;
        uint32_t fa, fb;
        const bool fast = vl_geometry(t0, t1, h, w, n0, n1, fa, fb);
        const int q = fast ? n0 * w + n1 : 0;
        const elem *g0 = src + q, *g1 = g0 + (fast ? w : 0);
        uint2 e[4];
        e[0] = VlSrc<SRC>::ld(g0); e[1] = VlSrc<SRC>::ld(g0 + 1); e[2] = VlSrc<SRC>::ld(g1); e[3] = VlSrc<SRC>::ld(g1 + 1);
        const float b0 = (float)__ldcs(bgp), b1 = (float)__ldcs(bgp + 1), b2 = (float)__ldcs(bgp + 2);
        float4 o;
        float na;
        const unsigned unc = vl_blend<SRC>(e, fa, fb, fast, b0, b1, b2, o, na);
        if (unc) o = vl_exact_px<SRC>(src, t0, t1, h, w, bgp, o, na, unc, outside);
        __stcs(op, o);
        bgp += w3; op += w;
    }
}

// one warp polls the mbarrier (phase 0), the CTA barrier releases everybody else without spinning;
// every thread then observes the completed phase itself (one try_wait that succeeds at once), which is
// what orders the async-proxy writes of the bulk copies before its own shared-memory reads
// `vote`: the CTA barrier doubles as an AND-vote (returned)
__device__ __forceinline__ bool vl_cta_wait(uint32_t bar, bool poll, bool used, uint32_t parity, bool vote = true) {
    uint32_t ok = 0;
    if (poll) {
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
    const bool all = __syncthreads_and(vote);
    if (used && !poll) {
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
    return all;
}


// Source box of every B2 tile: bounding box of the coarse transform window the tile interpolates
// in (the fine transform is a convex combination of those values), aligned for 16-byte bulk copies.
// One warp per tile; runs right behind B1 on the same stream.
template <int SRC>
__global__ void __launch_bounds__(128)
k_lean_boxes(const double2 *__restrict__ T, const int *__restrict__ F, int nx, int ny, const vm_axis_entry *__restrict__ rows,
             const vm_axis_entry *__restrict__ cols, int h, int w, int rpt, int tiles_x, int tiles_y, int n_tiles,
             int box_cap, VlTileBox *__restrict__ boxes, const VlShapes shapes) {
    constexpr int EPV = 16 / (int)sizeof(typename VlSrc<SRC>::elem);
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (t >= n_tiles) return;
    const int per = tiles_x * tiles_y;
    const int frame = t / per, tl = t - frame * per;
    const int ty = tl / tiles_x, tx = tl - ty * tiles_x;
    const int I0 = ty * (VL_FS * rpt), J0 = tx * VL_FW;
    const int th = min(VL_FS * rpt, h - I0), tw = min(VL_FW, w - J0);
    const vm_axis_entry r0 = vm_ld_axis(rows + I0), r1 = vm_ld_axis(rows + I0 + th - 1);
    const vm_axis_entry c0 = vm_ld_axis(cols + J0), c1 = vm_ld_axis(cols + J0 + tw - 1);
    const int kr0 = r0.i0, kr1 = max(r1.i1, r1.i0), kc0 = c0.i0, kc1 = max(c1.i1, c1.i0);
    const int nkr = kr1 - kr0 + 1, nkc = kc1 - kc0 + 1;
    VlTileBox rec = {0, 0, 0, 0};
    if (nkr >= 1 && nkc >= 1 && nkr <= 4096 && nkc <= 4096 && kr0 >= 0 && kr1 < nx && kc0 >= 0 && kc1 < ny) {
        const int64_t woff = (int64_t)frame * nx * ny + (int64_t)kr0 * ny + kc0;
        const double2 *Tf = T + woff;
        int rlo = INT_MAX, rhi = INT_MIN, clo = INT_MAX, chi = INT_MIN, bad = 0;
        if (F && nkr <= VL_TR && nkc <= 32 * VL_BG) {                            // packed floors written by the spline stage:
            const int *Ff = F + woff;                                    // every load of the window is issued before the
            int v[VL_TR][VL_BG];                                             // first one is used (the stage is latency bound)
#pragma unroll
            for (int r = 0; r < VL_TR; ++r)
#pragma unroll
                for (int g = 0; g < VL_BG; ++g) {
                    const int c = lane + 32 * g;
                    v[r][g] = (r < nkr && c < nkc) ? __ldg(Ff + (int64_t)r * ny + c) : 0x7FFF7FFF;     // 0x7FFF7FFF: no point
                }
#pragma unroll
            for (int r = 0; r < VL_TR; ++r)
#pragma unroll
                for (int g = 0; g < VL_BG; ++g) {
                    const int x = v[r][g];
                    if (x == VL_FLOOR_BAD) bad = 1;
                    else if (x != 0x7FFF7FFF) {
                        const int f0 = (int)(short)(x & 0xFFFF), f1 = x >> 16;
                        rlo = min(rlo, f0); rhi = max(rhi, f0); clo = min(clo, f1); chi = max(chi, f1);
                    }
                }
        } else {
            for (int r = 0; r < nkr; ++r)
                for (int c = lane; c < nkc; c += 32) {
                    const double2 v = __ldg(Tf + (int64_t)r * ny + c);
                    if (!(fabs(v.x) < 1.0e9) || !(fabs(v.y) < 1.0e9)) bad = 1;
                    else {
                        const int f0 = __double2int_rd(v.x), f1 = __double2int_rd(v.y);
                        rlo = min(rlo, f0); rhi = max(rhi, f0); clo = min(clo, f1); chi = max(chi, f1);
                    }
                }
        }
        rlo = __reduce_min_sync(0xffffffffu, rlo); rhi = __reduce_max_sync(0xffffffffu, rhi);
        clo = __reduce_min_sync(0xffffffffu, clo); chi = __reduce_max_sync(0xffffffffu, chi);
        bad = __reduce_max_sync(0xffffffffu, bad);
        // rows [rmin, rmax] x columns [cmin, cmin + bw) hold every tap of every in-frame fast pixel
        const int rmin = max(rlo, 0), rmax = min(rhi + 1, h - 1);
        int cmin = max(clo, 0) & ~(EPV - 1);
        const int cmax = min(chi + 1, w - 1);
        const int bh = rmax - rmin + 1;
        const int bw = (cmax - cmin + 1 + EPV - 1) & ~(EPV - 1);
        if (cmin + bw > w) cmin = w - bw;
        if (shapes.on) {                                                 // tensor copies: the first fixed shape that holds the box
            if (!bad && bh >= 1 && bw >= 1 && cmin >= 0) {
#pragma unroll
                for (int s = VL_NSHAPE - 1; s >= 0; --s)
                    if (bw <= shapes.bw[s] && bh <= shapes.bh[s]) { rec.rmin = rmin; rec.bh = shapes.bh[s]; rec.cmin = cmin; rec.bw = shapes.bw[s]; }
            }
        } else if (!bad && bh >= 1 && bw >= EPV && cmin >= 0 && (w & (EPV - 1)) == 0 && bh * bw <= box_cap) {
            rec.rmin = rmin; rec.bh = bh; rec.cmin = cmin; rec.bw = bw;
        }
    }
    if (lane == 0) boxes[t] = rec;
}

// One B2 tile (VL_FW columns x VL_FS * rpt rows at tile (bx, by) of frame `frame`) by the 256 threads of a
// CTA.  S.bar[0] must have been initialised (count 1); `phase` is its current parity and is toggled when
// the tile used it, so the function can be called for tile after tile by a persistent CTA.
template <int SRC, bool XCTA>      // XCTA: the source was written by other CTAs of the same kernel (full proxy fence)
__device__ __forceinline__ void vl_fine_tile(VlFineSmem &S, uint32_t &phase, const void *__restrict__ src_all,
                                             const uint8_t *__restrict__ bg, int n_bg, int frame0,
                                             const double2 *__restrict__ T, int nx, int ny,
                                             const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols,
                                             int h, int w, int rpt, VlTileBox rec, int frame, int by, int bx,
                                             float4 *__restrict__ out, int &outside, int &slow) {
    typedef typename VlSrc<SRC>::elem elem;
    const int tid = threadIdx.y * VL_FW + threadIdx.x;
#ifdef VL_TIMING
    const bool vl_tim = tid == 0 || tid == VL_FW * VL_FS - 1;          // warp 0 (issues / polls the bulk copies) and the last warp
    long long vl_t0 = clock64();
    unsigned long long vl_acc[6] = {0, 0, 0, 0, 0, 0};
#endif
    const int J0 = bx * VL_FW, I0 = by * (VL_FS * rpt);
    const int tw = min(VL_FW, w - J0), th = min(VL_FS * rpt, h - I0);
    const int jc = min((int)threadIdx.x, tw - 1), j = J0 + jc;
    const uint32_t bar0 = vl_smem_u32(&S.bar[0]);
    // frame offsets as ONE 32 x 32 -> 64 multiply each (h * w < 2^28 and frame < 2^16 are checked on the host)
    const uint32_t hw = (uint32_t)h * (uint32_t)w;
    const int64_t fbase = (int64_t)((uint64_t)(uint32_t)frame * hw);
    const elem *src = reinterpret_cast<const elem *>(src_all) + fbase;
    // background index (frame0 + frame) mod n_bg without an integer division: float estimate of the quotient + one correction
    uint32_t bgi = (uint32_t)(frame0 + frame);
    if (bgi >= (uint32_t)n_bg) {
        const uint32_t q = (uint32_t)__fmul_rz(__uint2float_rz(bgi), __frcp_rz(__uint2float_ru((uint32_t)n_bg)));   // <= true quotient, at most 1 short (operands < 2^24)
        bgi -= q * (uint32_t)n_bg;
        if (bgi >= (uint32_t)n_bg) bgi -= (uint32_t)n_bg;
    }
    const uint8_t *bgf = bg + (int64_t)((uint64_t)bgi * (hw * 3u));
    const double2 *Tf = T + (int64_t)((uint64_t)(uint32_t)frame * ((uint32_t)nx * (uint32_t)ny));

    // ---- P0: axis entries.  Every load of the prologue that does not depend on another one is issued here, before the
    // first use: the column entry of the thread, the first / last column and row entries of the tile (they give the coarse
    // window without a round trip through shared memory) and the thread's row entry ------------------------------------
    const vm_axis_entry ce = vm_ld_axis(cols + j);
    const vm_axis_entry cf = vm_ld_axis(cols + J0), cl = vm_ld_axis(cols + J0 + tw - 1);
    const vm_axis_entry rf = vm_ld_axis(rows + I0), rl = vm_ld_axis(rows + I0 + th - 1);
    vm_axis_entry myrow = {0.0, 0, 0};
    if (tid < th) myrow = vm_ld_axis(rows + I0 + tid);
    // ---- P1: bulk async copies (TMA): background rows + source box, all on one mbarrier.  They need the tile record only
    // and are issued while the axis entries requested above are still on their way
    const bool bg_sm = tw == VL_FW && (w & 15) == 0 && (reinterpret_cast<uintptr_t>(bg) & 15) == 0;
    const bool boxed = rec.bw > 0 && (reinterpret_cast<uintptr_t>(src_all) & 15) == 0;
    const bool copy_box = boxed && !VL_ABL(1), copy_bg = bg_sm && !VL_ABL(2);
    const bool used = copy_bg || copy_box;
    elem *boxp = reinterpret_cast<elem *>(S.box);
    if (used) {
        // the copies are dealt out over all warps - copy c goes to lane c / W of warp c % W (W warps per CTA): a warp's
        // bulk-copy instruction is executed lane by lane (~65 cycles per copy, -DVL_TIMING), so 72 copies by the lanes of ONE
        // warp kept that warp busy for a third of the tile's time while the other seven waited at the barrier
        constexpr int NW = VL_FW * VL_FS / 32;
        if (tid == 0) {
            const uint32_t bytes = (copy_bg ? (uint32_t)(th * VL_FW * 3) : 0u) + (copy_box ? (uint32_t)(rec.bh * rec.bw * (int)sizeof(elem)) : 0u);
            // earlier generic accesses (this CTA's reads of the buffers; with XCTA also the acquired global
            // writes of other CTAs) are ordered before the async-proxy copies
            if (XCTA) asm volatile("fence.proxy.async;" ::: "memory");
            else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(bytes) : "memory");
        }
        const int c0 = (tid & 31) * NW + (tid >> 5);                   // this thread's first copy; further ones 32 * NW apart
        const int nbox = copy_box ? rec.bh : 0, nbg = copy_bg ? th : 0;
        const uint32_t row_bytes = (uint32_t)(rec.bw * (int)sizeof(elem));
        const elem *gb = src + ((uint32_t)rec.rmin * (uint32_t)w + (uint32_t)rec.cmin);
        const uint8_t *gg = bgf + ((uint32_t)I0 * (uint32_t)w + (uint32_t)J0) * 3u;
        for (int c = c0; c < nbox + nbg; c += 32 * NW) {
            if (c < nbox) vl_bulk_g2s(vl_smem_u32(boxp + c * rec.bw), gb + (uint32_t)c * (uint32_t)w, row_bytes, bar0);
            else { const int r = c - nbox; vl_bulk_g2s(vl_smem_u32(S.bgt + r * (VL_FW * 3)), gg + (uint32_t)r * (uint32_t)(w * 3), VL_FW * 3, bar0); }
        }
    }

    const int kr0 = rf.i0, kr1 = max(rl.i1, rl.i0);
    const int nkr = kr1 - kr0 + 1;
    const int kc0 = cf.i0, kc1 = max(cl.i1, cl.i0);
    // every (i0, i1) of the tile's rows must lie inside the window that Cs covers, and every (i0, i1) of its columns
    // inside the window k_lean_boxes scanned
    const bool row_ok = tid >= th || (myrow.i0 >= kr0 && myrow.i0 <= kr1 && myrow.i1 >= kr0 && myrow.i1 <= kr1);
    const bool win_ok = nkr >= 1 && nkr <= VL_TR && kr0 >= 0 && kr1 < nx && kc0 >= 0 && kc1 < ny &&
                        ce.i0 >= kc0 && ce.i0 <= kc1 && ce.i1 >= kc0 && ce.i1 <= kc1;
    // the thread's share of the coarse window (P2) is requested now: its latency runs under the vote and the bulk-copy issue
    constexpr int KMAX = (VL_TR + VL_FS - 1) / VL_FS;
    double2 ta[KMAX], tb[KMAX];
    if (win_ok) {
        const double2 *Ta = Tf + (int64_t)kr0 * ny + ce.i0, *Tb = Tf + (int64_t)kr0 * ny + ce.i1;
#pragma unroll
        for (int u = 0; u < KMAX; ++u) {
            const int k = min((int)threadIdx.y + u * VL_FS, nkr - 1);
            const uint32_t o = (uint32_t)k * (uint32_t)ny;
            ta[u] = __ldg(Ta + o); tb[u] = __ldg(Tb + o);
        }
    }
    if (tid < th) S.rows[tid] = myrow;
    VL_T(0);                                                            // P0: set-up, axis entries, bulk-copy issue, transform requests

    const int strip0 = threadIdx.y * rpt;                               // first row of the strip inside the CTA
    const int nrows = min(rpt, th - strip0);
    const int64_t p0 = (int64_t)(I0 + strip0) * w + j;
    const uint8_t *bgp = bgf + p0 * 3;
    float4 *op = out + fbase + p0;
    const bool active = (int)threadIdx.x < tw && nrows > 0;

    // ---- P2: column-interpolated coarse rows; thread (jc, strip) takes rows strip, strip + VL_FS, ...
    if (win_ok) {
        const double yf = ce.frac, y1 = 1.0 - yf;
#pragma unroll
        for (int u = 0; u < KMAX; ++u) {
            const int k = threadIdx.y + u * VL_FS;
            if (k < nkr) {
                const VlC c = vl_col_lerp(ta[u], tb[u], y1, yf);
                S.Cs[k * VL_FW + threadIdx.x] = make_double2(c.x, c.y);
            }
        }
    }
    VL_T(2);                                                            // P2: transform window -> Cs
    // one CTA barrier for three things: the bulk copies have landed, Cs / S.rows are published, and the vote whether every
    // thread found its axis entries inside the staged windows
    const bool all_staged = vl_cta_wait(bar0, tid < 32 && used, used, phase, row_ok && win_ok);
    if (used) phase ^= 1u;
    VL_T(3);                                                            // wait for the bulk copies + CTA barrier

    if (!all_staged) {                                                  // generic axis tables: everything from global memory
        if (active) vl_strip_generic<SRC>(src, Tf + ce.i0, Tf + ce.i1, ny, rows + I0 + strip0, ce.frac, bgp, op, nrows, h, w, &outside);
        if (tid == 0) ++slow;
        return;
    }

    const unsigned char *bgl = S.bgt + (strip0 * VL_FW + (int)threadIdx.x) * 3;
    const double2 *Csj = S.Cs + threadIdx.x;

    // ---- P3: per-pixel resampling + composite ------------------------------------------------
    if (!boxed && tid == 0) ++slow;
    if (active) {
        const vm_axis_entry *rp = S.rows + strip0;
        if (boxed && bg_sm) vl_strip_tile<SRC, true, true>(src, boxp, rec.rmin, rec.cmin, rec.bw, Csj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside);
        else if (boxed)     vl_strip_tile<SRC, true, false>(src, boxp, rec.rmin, rec.cmin, rec.bw, Csj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside);
        else                vl_strip_tile<SRC, false, false>(src, boxp, 0, 0, 0, Csj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside);
    }
    VL_T(4);                                                            // P3: resampling + composite
#ifdef VL_TIMING
    if (vl_tim) {
        const int o = tid == 0 ? 0 : 8;
        for (int k = 0; k < 5; ++k) atomicAdd(&g_vl_prof[o + k], vl_acc[k]);
        atomicAdd(&g_vl_prof[o + 7], 1ull);
    }
#endif
}

__device__ __forceinline__ void vl_bar_init(VlFineSmem &S) {
    if (threadIdx.y * VL_FW + threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(vl_smem_u32(&S.bar[0])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
}

template <int SRC, int MINB>
__global__ void __launch_bounds__(VL_FW * VL_FS, MINB)
k_lean_fine(const void *__restrict__ src_all, const uint8_t *__restrict__ bg, int n_bg, int frame0,
            const double2 *__restrict__ T, int nx, int ny, const vm_axis_entry *__restrict__ rows,
            const vm_axis_entry *__restrict__ cols, int h, int w, int rpt, const VlTileBox *__restrict__ boxes,
            float4 *__restrict__ out, int32_t *__restrict__ status) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    VlFineSmem &S = *reinterpret_cast<VlFineSmem *>(smem_raw);
    const VlTileBox rec = boxes[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x];     // < 2^31 tiles per launch
    vl_bar_init(S);
    uint32_t phase = 0;
    int outside = 0, slow = 0;
    vl_fine_tile<SRC, false>(S, phase, src_all, bg, n_bg, frame0, T, nx, ny, rows, cols, h, w, rpt, rec, blockIdx.z, blockIdx.y, blockIdx.x,
                      out, outside, slow);
    if (status) {
        if (outside & (VL_NEAR_KNIFE_UNIT - 1)) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside & (VL_NEAR_KNIFE_UNIT - 1));
        if (outside >= VL_NEAR_KNIFE_UNIT) atomicAdd(status + VM_STATUS_NEAR_KNIFE, outside / VL_NEAR_KNIFE_UNIT);
        if (slow) atomicAdd(status + VM_STATUS_SLOW_TILES, slow);
    }
}

// The resampling kernel with tensor-map staging.  Per tile: thread 0 reads the tile record and the first row / column
// entries (which give the origin of the coarse window), then issues THREE tensor bulk copies on one mbarrier - the source
// box, the background rows, the raw coarse window of the transform - and the CTA has ONE barrier before the pixel loop:
// the copies have landed, the row entries are published, and the vote whether every thread's axis entries lie inside the
// staged windows.  No Cs pass, no per-thread transform loads, no address arithmetic on 64-bit pointers in the prologue.
template <int SRC, int MINB>
__global__ void __launch_bounds__(VL_FW * VL_FS, MINB)
k_lean_fine_tm(const void *__restrict__ src_all, const uint8_t *__restrict__ bg, int n_bg, int frame0,
               const double2 *__restrict__ T, int nx, int ny, const vm_axis_entry *__restrict__ rows,
               const vm_axis_entry *__restrict__ cols, int h, int w, int rpt, const VlTileBox *__restrict__ boxes,
               float4 *__restrict__ out, int32_t *__restrict__ status, const __grid_constant__ VlTmaps tm) {
    typedef typename VlSrc<SRC>::elem elem;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    VlFineSmemT &S = *reinterpret_cast<VlFineSmemT *>(smem_raw);
    const int tid = threadIdx.y * VL_FW + threadIdx.x;
    const int frame = blockIdx.z, by = blockIdx.y, bx = blockIdx.x;
    const int J0 = bx * VL_FW, I0 = by * (VL_FS * rpt);
    const int tw = min(VL_FW, w - J0), th = min(VL_FS * rpt, h - I0);
    const int jc = min((int)threadIdx.x, tw - 1), j = J0 + jc;
    const uint32_t bar0 = vl_smem_u32(&S.bar[0]);
    // every load of the prologue is independent of the others and issued here
    const VlTileBox rec = boxes[(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x];
    const vm_axis_entry rf = vm_ld_axis(rows + I0), rl = vm_ld_axis(rows + I0 + th - 1);
    const vm_axis_entry cf = vm_ld_axis(cols + J0), cl = vm_ld_axis(cols + J0 + tw - 1);
    const vm_axis_entry ce = vm_ld_axis(cols + j);
    vm_axis_entry myrow = {0.0, 0, 0};
    if (tid < th) myrow = vm_ld_axis(rows + I0 + tid);
    const uint32_t hw = (uint32_t)h * (uint32_t)w;
    const int64_t fbase = (int64_t)((uint64_t)(uint32_t)frame * hw);
    const elem *src = reinterpret_cast<const elem *>(src_all) + fbase;
    uint32_t bgi = (uint32_t)(frame0 + frame);
    if (bgi >= (uint32_t)n_bg) {                                        // (frame0 + frame) mod n_bg without an integer division
        const uint32_t q = (uint32_t)__fmul_rz(__uint2float_rz(bgi), __frcp_rz(__uint2float_ru((uint32_t)n_bg)));
        bgi -= q * (uint32_t)n_bg;
        if (bgi >= (uint32_t)n_bg) bgi -= (uint32_t)n_bg;
    }
    const uint8_t *bgf = bg + (int64_t)((uint64_t)bgi * (hw * 3u));
    const double2 *Tf = T + (int64_t)((uint64_t)(uint32_t)frame * ((uint32_t)nx * (uint32_t)ny));
    const int kr0 = rf.i0, kr1 = max(rl.i1, rl.i0), kc0 = cf.i0, kc1 = max(cl.i1, cl.i0);
    const int nkr = kr1 - kr0 + 1, nkc = kc1 - kc0 + 1;
    const bool bg_sm = tw == VL_FW;
    const bool boxed = rec.bw > 0;
    elem *boxp = reinterpret_cast<elem *>(S.box);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        const uint32_t bytes = (uint32_t)(VL_TR * VL_TCP * 16) + (bg_sm ? (uint32_t)(VL_FROWS_MAX * VL_FW * 3) : 0u) +
                               (boxed ? (uint32_t)(rec.bh * rec.bw * (int)sizeof(elem)) : 0u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(vl_smem_u32(S.Traw)), "l"(reinterpret_cast<uint64_t>(&tm.T)), "r"(2 * kc0), "r"(kr0), "r"(frame), "r"(bar0) : "memory");
        if (boxed) {
            int sidx = 0;
#pragma unroll
            for (int q = 1; q < VL_NSHAPE; ++q) if (tm.bw[q] == rec.bw && tm.bh[q] == rec.bh) sidx = q;
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(vl_smem_u32(boxp)), "l"(reinterpret_cast<uint64_t>(&tm.box[sidx])), "r"(rec.cmin), "r"(rec.rmin), "r"(frame), "r"(bar0) : "memory");
        }
        if (bg_sm)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(vl_smem_u32(S.bgt)), "l"(reinterpret_cast<uint64_t>(&tm.bg)), "r"(J0 * 3), "r"(I0), "r"((int)bgi), "r"(bar0) : "memory");
    }
    // every (i0, i1) of the tile's rows / columns must lie inside the staged window
    const bool row_ok = tid >= th || (myrow.i0 >= kr0 && myrow.i0 <= kr1 && myrow.i1 >= kr0 && myrow.i1 <= kr1);
    const bool win_ok = nkr >= 1 && nkr <= VL_TR && nkc >= 1 && nkc <= VL_TCP && kr0 >= 0 && kr1 < nx && kc0 >= 0 && kc1 < ny &&
                        ce.i0 >= kc0 && ce.i0 <= kc1 && ce.i1 >= kc0 && ce.i1 <= kc1;
    if (tid < th) S.rows[tid] = myrow;
    const int strip0 = threadIdx.y * rpt;
    const int nrows = min(rpt, th - strip0);
    const uint32_t p0 = (uint32_t)(I0 + strip0) * (uint32_t)w + (uint32_t)j;
    const uint8_t *bgp = bgf + p0 * 3u;
    float4 *op = out + fbase + p0;
    const bool active = (int)threadIdx.x < tw && nrows > 0;
    // one CTA barrier - S.rows is published, the mbarrier's initialisation is visible, and the vote - then every warp waits
    // for the copies on the mbarrier itself (try_wait suspends the warp in hardware): no polling warp, and no second barrier
    // between the arrival of the data and the pixel loop
    const bool all_staged = __syncthreads_and(row_ok && win_ok);
    vl_mbar_wait_parity(bar0, 0);
    int outside = 0, slow = 0;
    if (!all_staged) {                                                  // generic axis tables: everything from global memory
        if (active) vl_strip_generic<SRC>(src, Tf + ce.i0, Tf + ce.i1, ny, rows + I0 + strip0, ce.frac, bgp, op, nrows, h, w, &outside);
        if (tid == 0) ++slow;
    } else {
        if (!boxed && tid == 0) ++slow;
        if (active) {
            const unsigned char *bgl = S.bgt + (strip0 * VL_FW + (int)threadIdx.x) * 3;
            const double2 *Tj = S.Traw + (ce.i0 - kc0);
            const int dq = ce.i1 - ce.i0;
            const double yf = ce.frac, y1 = 1.0 - yf;
            const vm_axis_entry *rp = S.rows + strip0;
            if (boxed && bg_sm) vl_strip_tile<SRC, true, true, true>(src, boxp, rec.rmin, rec.cmin, rec.bw, Tj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside, dq, y1, yf);
            else if (boxed)     vl_strip_tile<SRC, true, false, true>(src, boxp, rec.rmin, rec.cmin, rec.bw, Tj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside, dq, y1, yf);
            else                vl_strip_tile<SRC, false, false, true>(src, boxp, 0, 0, 0, Tj, kr0, rp, bgl, bgp, op, nrows, h, w, &outside, dq, y1, yf);
        }
    }
    if (status) {
        if (outside & (VL_NEAR_KNIFE_UNIT - 1)) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside & (VL_NEAR_KNIFE_UNIT - 1));
        if (outside >= VL_NEAR_KNIFE_UNIT) atomicAdd(status + VM_STATUS_NEAR_KNIFE, outside / VL_NEAR_KNIFE_UNIT);
        if (slow) atomicAdd(status + VM_STATUS_SLOW_TILES, slow);
    }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
static inline int64_t vl_align(int64_t v) { return (v + 255) & ~(int64_t)255; }

// scratch of one slot of `m` frames: packed flow-warped frames, coarse transform, packed floors, tile boxes, counter
struct VlSlot { void *packed; double2 *T; int *F; VlTileBox *boxes; unsigned int *counter; };
static int64_t vl_slot_bytes(int m, int h, int w) {
    const int64_t tiles = (int64_t)m * ((w + VL_FW - 1) / VL_FW) * ((h + 3) / 4);     // >= tiles for any rows-per-thread setting
    const int64_t cp = (int64_t)m * (h / 2 + 1) * (w / 2 + 1);
    return 256 + vl_align((int64_t)m * h * w * 8) + vl_align(cp * 16) + vl_align(cp * 4) + vl_align(tiles * 16);
}
static VlSlot vl_slot_at(unsigned char *base, int m, int h, int w) {
    const int64_t tiles = (int64_t)m * ((w + VL_FW - 1) / VL_FW) * ((h + 3) / 4);
    const int64_t cp = (int64_t)m * (h / 2 + 1) * (w / 2 + 1);
    VlSlot s;
    s.counter = reinterpret_cast<unsigned int *>(base); base += 256;
    s.packed = base; base += vl_align((int64_t)m * h * w * 8);
    s.T = reinterpret_cast<double2 *>(base); base += vl_align(cp * 16);
    s.F = reinterpret_cast<int *>(base); base += vl_align(cp * 4);
    s.boxes = reinterpret_cast<VlTileBox *>(base);
    (void)tiles;
    return s;
}

int64_t vm_lean_scratch_bytes(int n, int h, int w) {
    const int chunk = vl_chunk_for(h, w), m = n < chunk ? n : chunk;
    return 512 + vl_slot_bytes(m, h, w);
}

int vm_lean_set_option(const char *key, int value) {
    if (!strcmp(key, "lean_chunk") && value >= 0 && value <= 4096) { g_vl_chunk = value; return VM_OK; }
    if (!strcmp(key, "lean_rb") && value >= 0 && value <= 64) { g_vl_rb = value; return VM_OK; }
    if (!strcmp(key, "lean_b1_warps") && value >= 1 && value <= VL_B1_WARPS_HI) { g_vl_b1_warps = value; return VM_OK; }
    if (!strcmp(key, "lean_b1_dyr") && value >= 0 && value <= 1) { g_vl_b1_dyr = value; return VM_OK; }
    if (!strcmp(key, "lean_timing") && value >= 0 && value <= 1) { g_vl_timing = value; return VM_OK; }
    if (!strcmp(key, "lean_sub") && value >= 0 && value <= 4096) { g_vl_sub = value; return VM_OK; }
    if (!strcmp(key, "lean_box_cap") && value >= 0 && value <= 16384) { g_vl_box_cap = value; return VM_OK; }
    if (!strcmp(key, "lean_minb") && (value >= 2 && value <= 6 || value == 8)) { g_vl_minb = value; return VM_OK; }
    if (!strcmp(key, "lean_fine_rows") && value >= 1 && value <= 256) { g_vl_fine_rows = value; return VM_OK; }
    if (!strcmp(key, "lean_floors") && (value == 0 || value == 1)) { g_vl_floors = value; return VM_OK; }
    if (!strcmp(key, "lean_b1_ctas") && value >= 0 && value <= 4096) { g_vl_b1_ctas = value; return VM_OK; }
    if (!strcmp(key, "lean_tmap") && (value == 0 || value == 1)) { g_vl_tmap = value; return VM_OK; }
#ifdef VL_TIMING
    if (!strcmp(key, "lean_abl") && value >= 0 && value < 16) return cudaMemcpyToSymbol(c_vl_abl, &value, sizeof(int)) == cudaSuccess ? VM_OK : VM_ERR_CUDA;
#endif
    if (!strcmp(key, "lean_aug_minb") && (value == 8 || value == 12)) { g_va_minb = value; return VM_OK; }
    return VM_ERR_ARG;
}

// geometry shared by the launches of one call
struct VlCall {
    int mode, N, nx, ny, n_bg, h, w, rpt, box_cap;
    double step_x, step_y;
    const uint8_t *fg, *bg;
    const float *backward, *forward;
    const double *ctrl, *coef;
    const vm_axis_entry *rows, *cols;
    float *out;
    int32_t *status;
    size_t fine_smem;
    VlShapes shapes;                                 // fixed box shapes of the tensor-map staging (on = 0: one bulk copy per row)
    const char *what;
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*VlEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static VlEncodeTiled vl_encode_tiled() {
    static std::mutex mu;
    static bool tried = false;
    static VlEncodeTiled fn = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!tried) {
        tried = true;
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<VlEncodeTiled>(f);
        else
            cudaGetLastError();
    }
    return fn;
}
// (d2, d1, d0) tensor of `esize`-byte elements, innermost dimension d0, box {b0, b1, 1}
static bool vl_encode3(CUtensorMap *tm, const void *base, int esize, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
    VlEncodeTiled enc = vl_encode_tiled();
    if (!enc) return false;
    const CUtensorMapDataType dt = esize == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : (esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8);
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * (uint64_t)esize, d0 * d1 * (uint64_t)esize};
    const cuuint32_t box[3] = {b0, b1, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, dt, 3, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// spline on the coarse grid of frames [f0, f0 + m) into slot `sl`
static int *vl_floors_ptr(const VlCall &c, const VlSlot &sl, bool floors_in_packed) {
    return g_vl_floors && c.h <= 32766 && c.w <= 32766 ? (floors_in_packed ? reinterpret_cast<int *>(sl.packed) : sl.F) : nullptr;
}
static int vl_enqueue_coarse(const VlCall &c, const VlSlot &sl, int f0, int m, bool floors_in_packed, cudaStream_t st) {
    g_vl_launches += 1;
    return vl_launch_coarse(c.ctrl + (int64_t)f0 * c.N * 2, c.coef + (int64_t)f0 * (c.N + 3) * 2, c.N, m, c.nx, c.ny, c.step_x, c.step_y,
                            sl.T, vl_floors_ptr(c, sl, floors_in_packed), sl.counter, st);
}
// source box of every resampling tile of the slot's m frames
static int vl_enqueue_boxes(const VlCall &c, const VlSlot &sl, int m, bool floors_in_packed, cudaStream_t st) {
    int *F = vl_floors_ptr(c, sl, floors_in_packed);
    const dim3 grid((c.w + VL_FW - 1) / VL_FW, (c.h + VL_FS * c.rpt - 1) / (VL_FS * c.rpt), m);
    const int n_tiles = (int)(grid.x * grid.y * m);
    if (c.mode != 0) k_lean_boxes<1><<<(n_tiles + 3) / 4, 128, 0, st>>>(sl.T, F, c.nx, c.ny, c.rows, c.cols, c.h, c.w, c.rpt, grid.x, grid.y, n_tiles, c.box_cap, sl.boxes, c.shapes);
    else             k_lean_boxes<0><<<(n_tiles + 3) / 4, 128, 0, st>>>(sl.T, F, c.nx, c.ny, c.rows, c.cols, c.h, c.w, c.rpt, grid.x, grid.y, n_tiles, c.box_cap, sl.boxes, c.shapes);
    g_vl_launches += 1;
    return vm_check_launch("vm_lean box stage");
}

static int vl_enqueue_flow(const VlCall &c, const VlSlot &sl, int f0, int m, cudaStream_t st) {
    const int64_t px = (int64_t)c.h * c.w;
    g_vl_launches += 1;
    return vm_launch_flow_stage(c.fg + f0 * px * 4, c.backward + f0 * px * 2, (c.mode == 2 && c.forward) ? c.forward + f0 * px * 2 : nullptr,
                                m, c.h, c.w, sl.packed, c.status, st, true);
}

// resampling + composite of frames [f0, f0 + m); the slot's T / boxes start at frame `fs` of the slot
static int vl_enqueue_fine(const VlCall &c, const VlSlot &sl, int f0, int fs, int m, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    const int64_t px = (int64_t)c.h * c.w;
    const dim3 block(VL_FW, VL_FS), sgrid((c.w + VL_FW - 1) / VL_FW, (c.h + VL_FS * c.rpt - 1) / (VL_FS * c.rpt), m);
    const int64_t tiles_per_frame = (int64_t)sgrid.x * sgrid.y;
    const double2 *Ts = sl.T + (int64_t)fs * c.nx * c.ny;
    const VlTileBox *bs = sl.boxes + fs * tiles_per_frame;
    float4 *o4 = reinterpret_cast<float4 *>(c.out) + f0 * px;
    const void *src = c.mode != 0 ? (const void *)(reinterpret_cast<const uint2 *>(sl.packed) + fs * px) : (const void *)(c.fg + f0 * px * 4);
    const size_t fine_smem = c.fine_smem;
    VlTmaps tm;
    memset(&tm, 0, sizeof(tm));
    if (c.shapes.on) {
        const int esize = c.mode != 0 ? 8 : 4;
        bool ok = vl_encode3(&tm.bg, c.bg, 1, (uint64_t)c.w * 3, (uint64_t)c.h, (uint64_t)c.n_bg, VL_FW * 3, VL_FROWS_MAX);
        for (int q = 0; q < VL_NSHAPE && ok; ++q) {
            tm.bw[q] = c.shapes.bw[q]; tm.bh[q] = c.shapes.bh[q];
            ok = vl_encode3(&tm.box[q], src, esize, (uint64_t)c.w, (uint64_t)c.h, (uint64_t)m, (uint32_t)tm.bw[q], (uint32_t)tm.bh[q]);
        }
        ok = ok && vl_encode3(&tm.T, Ts, 8, (uint64_t)c.ny * 2, (uint64_t)c.nx, (uint64_t)m, 2 * VL_TCP, VL_TR);
        if (!ok) { vm_set_error("vm_lean: cuTensorMapEncodeTiled failed"); return VM_ERR_CUDA; }
        tm.on = 1;
    }
#define VL_FINE_TM(S, MB)                                                                                           \
    do {                                                                                                            \
        static size_t attr_set[64];                                                                                 \
        if (attr_set[dev & 63] < fine_smem) {                                                                       \
            cudaError_t e = cudaFuncSetAttribute(k_lean_fine_tm<S, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_smem); \
            if (e != cudaSuccess) { vm_set_error("vm_lean: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VM_ERR_CUDA; } \
            attr_set[dev & 63] = fine_smem;                                                                         \
        }                                                                                                           \
        k_lean_fine_tm<S, MB><<<sgrid, block, fine_smem, st>>>(src, c.bg, c.n_bg, f0, Ts, c.nx, c.ny, c.rows, c.cols, c.h, c.w, c.rpt, bs, o4, c.status, tm); \
    } while (0)
    if (tm.on) {
        if (c.mode != 0) { if (g_vl_minb == 3) VL_FINE_TM(1, 3); else VL_FINE_TM(1, 4); }
        else             { if (g_vl_minb == 3) VL_FINE_TM(0, 3); else VL_FINE_TM(0, 4); }
        g_vl_launches += 1;
        return vm_check_launch(c.what);
    }
#undef VL_FINE_TM
#define VL_FINE(S, MB)                                                                                              \
    do {                                                                                                            \
        static size_t attr_set[64];                                                                                 \
        if (attr_set[dev & 63] < fine_smem) {                                                                       \
            cudaError_t e = cudaFuncSetAttribute(k_lean_fine<S, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fine_smem); \
            if (e != cudaSuccess) { vm_set_error("vm_lean: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return VM_ERR_CUDA; } \
            attr_set[dev & 63] = fine_smem;                                                                         \
        }                                                                                                           \
        k_lean_fine<S, MB><<<sgrid, block, fine_smem, st>>>(src, c.bg, c.n_bg, f0, Ts, c.nx, c.ny, c.rows, c.cols, c.h, c.w, c.rpt, bs, o4, c.status); \
    } while (0)
    if (c.mode != 0) { if (g_vl_minb == 8) VL_FINE(1, 8); else if (g_vl_minb == 6) VL_FINE(1, 6); else if (g_vl_minb == 5) VL_FINE(1, 5); else if (g_vl_minb == 4) VL_FINE(1, 4); else if (g_vl_minb == 3) VL_FINE(1, 3); else VL_FINE(1, 2); }
    else             { if (g_vl_minb == 8) VL_FINE(0, 8); else if (g_vl_minb == 6) VL_FINE(0, 6); else if (g_vl_minb == 5) VL_FINE(0, 5); else if (g_vl_minb == 4) VL_FINE(0, 4); else if (g_vl_minb == 3) VL_FINE(0, 3); else VL_FINE(0, 2); }
#undef VL_FINE
    g_vl_launches += 1;
    return vm_check_launch(c.what);
}

// mode 0: C3 (no flow stage), 1: flow warp only, 2: flow warp + consistency mask
int vm_lean_launch(int mode, const uint8_t *fg, const float *backward, const float *forward, const uint8_t *bg,
                   int n_bg, const double *ctrl, const double *coef, int N, int nx, int ny, double step_x,
                   double step_y, const vm_axis_entry *rows, const vm_axis_entry *cols, int n, int h, int w,
                   float *out, void *scratch, int32_t *status, cudaStream_t st, const char *what) {
    int dev = 0;
    int rc = vl_init(&dev);
    if (rc != VM_OK) return rc;
    VM_REQUIRE(scratch, "scratch workspace (vm_fused_scratch_bytes) required");
    VM_REQUIRE(N >= 1 && N <= VL_MAX_N, "control point count out of range");
    VM_REQUIRE(nx <= h / 2 + 1 && ny <= w / 2 + 1, "coarse grid larger than the scratch layout");
    VlCall c;
    c.mode = mode; c.N = N; c.nx = nx; c.ny = ny; c.n_bg = n_bg; c.h = h; c.w = w;
    c.step_x = step_x; c.step_y = step_y; c.fg = fg; c.bg = bg; c.backward = backward; c.forward = forward;
    c.ctrl = ctrl; c.coef = coef; c.rows = rows; c.cols = cols; c.out = out; c.status = status; c.what = what;
    c.rpt = g_vl_fine_rows < VL_FROWS_MAX / VL_FS ? g_vl_fine_rows : VL_FROWS_MAX / VL_FS;
    {
        const dim3 grid((w + VL_FW - 1) / VL_FW, (h + VL_FS * c.rpt - 1) / (VL_FS * c.rpt), 1);
        VM_REQUIRE(grid.y <= 65535 && n <= 65535, "too many tiles for one launch");
    }
    // shared memory per resampling CTA: 227 KB per SM (1 KB reserved per CTA) split over the occupancy target
    c.box_cap = g_vl_box_cap;
    if (c.box_cap <= 0) {
        const int64_t per_cta = (227 * 1024) / g_vl_minb - 1024 - (int64_t)sizeof(VlFineSmem);
        c.box_cap = (int)(per_cta / 8) & ~63;
        if (c.box_cap > 8192) c.box_cap = 8192;
    }
    c.fine_smem = vl_fine_smem_bytes(c.box_cap);
    // tensor-map staging of the resampling stage: needs the driver entry point, 16-byte aligned planes / rows and boxes of <= 256
    memset(&c.shapes, 0, sizeof(c.shapes));
    if (g_vl_tmap && (g_vl_minb == 3 || g_vl_minb == 4) && vl_encode_tiled() && (w & 15) == 0 && vm_aligned(bg, 16) && vm_aligned(fg, 16) &&
        vm_aligned(scratch, 16)) {
        if (g_vl_box_cap <= 0) {                                        // the raw transform window replaces Cs: more room for the box
            const int64_t per_cta = (227 * 1024) / g_vl_minb - 1024 - (int64_t)sizeof(VlFineSmemT);
            c.box_cap = (int)(per_cta / 8) & ~63;
            if (c.box_cap > 8192) c.box_cap = 8192;
        }
        static const int kBw[VL_NSHAPE] = {80, 76, 84, 72, 88, 68, 92, 64, 96, 104, 112, 124};   // widths (multiples of 4), height = budget / width
        bool ok = true;
        for (int q = 0; q < VL_NSHAPE; ++q) {
            c.shapes.bw[q] = kBw[q];
            c.shapes.bh[q] = c.box_cap / kBw[q] < 256 ? c.box_cap / kBw[q] : 256;
            ok = ok && c.shapes.bh[q] >= 4;
        }
        c.shapes.on = ok ? 1 : 0;
        if (ok) c.fine_smem = vl_fine_smem_bytes_t(c.box_cap);
        else if (g_vl_box_cap <= 0) {
            const int64_t per_cta = (227 * 1024) / g_vl_minb - 1024 - (int64_t)sizeof(VlFineSmem);
            c.box_cap = (int)(per_cta / 8) & ~63;
            if (c.box_cap > 8192) c.box_cap = 8192;
        }
    }
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cap);
    // ---- one round per chunk of `lean_chunk` frames on the caller's stream: spline, tile boxes, flow stage, resampling
    const int chunk = vl_chunk_for(h, w);
    const int mc = n < chunk ? n : chunk;
    const VlSlot sl = vl_slot_at(base, mc, h, w);
    const bool timing = g_vl_timing && cap == cudaStreamCaptureStatusNone;
    if (timing && !g_vl_tev_ok[dev]) {
        for (int k = 0; k < 5; ++k)
            if (cudaEventCreate(&g_vl_tev[dev][k]) != cudaSuccess) { vm_set_error("vm_lean: cudaEventCreate failed"); return VM_ERR_CUDA; }
        g_vl_tev_ok[dev] = true;
    }
    for (int f0 = 0, ci = 0; f0 < n; f0 += chunk, ++ci) {
        const int m = (n - f0 < chunk) ? n - f0 : chunk;
        const bool tev = timing && ci == 0;
        auto mark = [&](int k) { if (tev) cudaEventRecord(g_vl_tev[dev][k], st); };
        if (tev) { const int pr[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 4}}; memcpy(g_vl_tpair[dev], pr, sizeof(pr)); }
        mark(0);
        // the packed floors of T live at the start of the flow stage's output buffer, which nobody touches until
        // the box stage is done (same stream)
        rc = vl_enqueue_coarse(c, sl, f0, m, true, st);
        if (rc != VM_OK) return rc;
        mark(1);
        rc = vl_enqueue_boxes(c, sl, m, true, st);
        if (rc != VM_OK) return rc;
        mark(2);
        // flow stage, then resampling + composite, in sub-rounds of `lean_sub` frames (0 = the whole round)
        const int sub = (g_vl_sub > 0 && g_vl_sub < m && !tev) ? g_vl_sub : m;
        for (int fs = 0; fs < m; fs += sub) {
            const int ms = (m - fs < sub) ? m - fs : sub;
            if (mode != 0) {
                rc = vl_enqueue_flow(c, sl, f0 + fs, ms, st);           // the sub-round's packed frames start the buffer
                if (rc != VM_OK) return rc;
            }
            mark(3);
            VlSlot part = sl;
            part.T = sl.T + (int64_t)fs * nx * ny;
            part.boxes = sl.boxes + (int64_t)fs * ((w + VL_FW - 1) / VL_FW) * ((h + VL_FS * c.rpt - 1) / (VL_FS * c.rpt));
            rc = vl_enqueue_fine(c, part, f0 + fs, 0, ms, st);
            if (rc != VM_OK) return rc;
        }
        mark(4);
    }
    return VM_OK;
}

// Durations (ms) of the four stages of the first chunk of the last timed call on the current device:
// {spline (k_lean_coarse), tile boxes (k_lean_boxes), flow stage (k_flow_warp_mask_bgra), resampling +
// composite (k_lean_fine)}.  Requires vm_set_option("lean_timing", 1) before the call and a synchronised
// stream.  Returns VM_OK, or VM_ERR_ARG when no timed call has completed.
extern "C" int vm_lean_stage_ms(float *out4) {
    int dev = 0;
    if (!out4 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_vl_tev_ok[dev]) {
        vm_set_error("vm_lean_stage_ms: no timed call on this device");
        return VM_ERR_ARG;
    }
    for (int k = 0; k < 4; ++k)
        if (g_vl_tpair[dev][k][0] == g_vl_tpair[dev][k][1]) out4[k] = 0.f;
        else if (cudaEventElapsedTime(out4 + k, g_vl_tev[dev][g_vl_tpair[dev][k][0]], g_vl_tev[dev][g_vl_tpair[dev][k][1]]) != cudaSuccess) {
            cudaGetLastError();
            vm_set_error("vm_lean_stage_ms: events not complete");
            return VM_ERR_ARG;
        }
    return VM_OK;
}

// number of kernels the lean path has launched in this process (all devices)
extern "C" long long vm_lean_launch_count(void) { return g_vl_launches.load(); }

// ---------------------------------------------------------------------------------------
// Batched augmentation (SURVEY 8a-11 / BASELINE config 5): TPS stage of augmentation.warp_image for a
// whole clip.  BGRA frames (alpha = A/255, reader.py:16) are resampled through the per-frame spline on
// the (h+1) x (w+1) grid of tps.warp_images (tps.py:55) into a packed intermediate {B|G<<8|R<<16 with the
// uint8 half-up rounding of map_coordinates, alpha' as float32 bits}; the two warpAffine passes and the
// illumination change follow in vm_affine.cu (k_aug_affine).  Taps come straight from global memory:
// this path is bound by its host orchestration (RNG draws and pinv per frame), not by this kernel.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint2 vl_exact_cols(const uint32_t *__restrict__ src, double t0, double t1, int h, int w, int *outside,
                                            double *a64_out) {
    const VmBilin64 s = vm_mapcoord_setup(t0, t1, h, w);
    if (!s.inside) {
        (*outside)++;
        *a64_out = 0.0;
        return make_uint2(0u, 0u);                                      // map_coordinates: cval = 0
    }
    const uint32_t e0 = __ldg(src + ((int64_t)s.i0 * w + s.j0)), e1 = __ldg(src + ((int64_t)s.i0 * w + s.j1));
    const uint32_t e2 = __ldg(src + ((int64_t)s.i1 * w + s.j0)), e3 = __ldg(src + ((int64_t)s.i1 * w + s.j1));
    uint32_t bgr = 0;
#pragma unroll
    for (int c = 0; c < 3; ++c)
        bgr |= (uint32_t)vm_round_half_up_u8(vm_mapcoord_blend(s, (double)((e0 >> (8 * c)) & 255u), (double)((e1 >> (8 * c)) & 255u),
                                                               (double)((e2 >> (8 * c)) & 255u), (double)((e3 >> (8 * c)) & 255u))) << (8 * c);
    const double a64 = vm_mapcoord_blend(s, (double)(e0 >> 24) / 255.0, (double)(e1 >> 24) / 255.0,
                                         (double)(e2 >> 24) / 255.0, (double)(e3 >> 24) / 255.0);
    *a64_out = a64;
    return make_uint2(bgr, __float_as_uint((float)a64));
}

#define VA_ROWS 16
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_aug_tps(const uint32_t *__restrict__ fg, const double2 *__restrict__ T, int nx, int ny,
          const vm_axis_entry *__restrict__ rows, const vm_axis_entry *__restrict__ cols, int h, int w,
          uint2 *__restrict__ inter, double *__restrict__ alpha64, int32_t *__restrict__ status) {
    const int oh = h + 1, ow = w + 1;
    const int j = blockIdx.x * 128 + threadIdx.x, frame = blockIdx.z;
    const int ibeg = blockIdx.y * VA_ROWS, iend = min(ibeg + VA_ROWS, oh);
    if (j >= ow) return;
    const uint32_t *src = fg + (int64_t)frame * h * w;
    uint2 *op = inter + (int64_t)frame * oh * ow + j;
    double *ap = alpha64 ? alpha64 + (int64_t)frame * oh * ow + j : nullptr;
    const vm_axis_entry ce = vm_ld_axis(cols + j);
    const double yf = ce.frac, y1 = 1.0 - yf;
    const double2 *Ta = T + (int64_t)frame * nx * ny + ce.i0, *Tb = T + (int64_t)frame * nx * ny + ce.i1;
    int k0 = -1, k1 = -1, outside = 0;
    VlC C0 = {0.0, 0.0}, C1 = {0.0, 0.0};
    for (int i = ibeg; i < iend; ++i) {
        const vm_axis_entry re = vm_ld_axis(rows + i);
        if (re.i0 != k0) {
            if (re.i0 == k1) C0 = C1; else C0 = vl_col_lerp(__ldg(Ta + re.i0 * ny), __ldg(Tb + re.i0 * ny), y1, yf);
            k0 = re.i0;
        }
        if (re.i1 != k1) {
            if (re.i1 == k0) C1 = C0; else C1 = vl_col_lerp(__ldg(Ta + re.i1 * ny), __ldg(Tb + re.i1 * ny), y1, yf);
            k1 = re.i1;
        }
        const double xf = re.frac, x1 = 1.0 - xf;
        const double t0 = fma(C1.x, xf, C0.x * x1), t1 = fma(C1.y, xf, C0.y * x1);
        int n0, n1;
        uint32_t fa, fb;
        const bool fast = vl_geometry(t0, t1, h, w, n0, n1, fa, fb);
        uint2 o;
        double a64 = 0.0;
        if (fast) {
            const uint32_t *g0 = src + (n0 * w + n1), *g1 = g0 + w;
            const uint32_t e0 = __ldg(g0), e1 = __ldg(g0 + 1), e2 = __ldg(g1), e3 = __ldg(g1 + 1);
            const uint32_t A1 = fa >> 1, A0 = 0x80000000u - A1, B1 = fb >> 1, B0 = 0x80000000u - B1;
            const uint32_t W00 = __umulhi(A0, B0), W01 = __umulhi(A0, B1), W10 = __umulhi(A1, B0), W11 = __umulhi(A1, B1);
            const uint32_t w0 = W00 >> 6, w1 = W01 >> 6, w2 = W10 >> 6, w3 = W11 >> 6;
            uint32_t bgr = 0;
            bool knife = false;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const uint32_t v = ((e0 >> (8 * c)) & 255u) * w0 + ((e1 >> (8 * c)) & 255u) * w1 + ((e2 >> (8 * c)) & 255u) * w2 +
                                   ((e3 >> (8 * c)) & 255u) * w3 + (8388608u + 1100u);
                bgr |= (v >> 24) << (8 * c);
                knife |= (v & 0x00FFFFFFu) < 1108u;
            }
            const unsigned long long a2f = (unsigned long long)(e0 >> 24) * W00 + (unsigned long long)(e1 >> 24) * W01 +
                                           (unsigned long long)(e2 >> 24) * W10 + (unsigned long long)(e3 >> 24) * W11;
            o = make_uint2(bgr, __float_as_uint((float)a2f * (float)(1.0 / (255.0 * 1073741824.0))));
            if (knife) o = vl_exact_cols(src, t0, t1, h, w, &outside, &a64);
            else if (ap) {                                              // scipy's float64 operation order (tps.py:34)
                const VmBilin64 sb = vm_mapcoord_setup(t0, t1, h, w);
                a64 = vm_mapcoord_blend(sb, (double)(e0 >> 24) / 255.0, (double)(e1 >> 24) / 255.0,
                                        (double)(e2 >> 24) / 255.0, (double)(e3 >> 24) / 255.0);
            }
        } else {
            o = vl_exact_cols(src, t0, t1, h, w, &outside, &a64);
        }
        op[(int64_t)i * ow] = o;
        if (ap) ap[(int64_t)i * ow] = a64;
    }
    if (status && outside) atomicAdd(status + VM_STATUS_TPS_OUTSIDE, outside);
}

// coarse transform of n frames as (n, nx, ny) double2 {row coordinate, column coordinate} - the spline stage
// of the fused paths as a stand-alone entry point.  `counter`: one zero-initialisable device word.
extern "C" int vm_tps_coarse_packed(const double *ctrl, const double *coef, int n, int N, int nx, int ny, double step_x,
                                    double step_y, void *T, unsigned int *counter, void *stream) {
    VM_REQUIRE(ctrl && coef && T && counter, "null pointer");
    VM_REQUIRE(n >= 0 && N >= 1 && N <= VL_MAX_N && nx >= 1 && ny >= 1, "bad size");
    if (n == 0) return VM_OK;
    int dev = 0;
    int rc = vl_init(&dev);
    if (rc != VM_OK) return rc;
    return vl_launch_coarse(ctrl, coef, N, n, nx, ny, step_x, step_y, reinterpret_cast<double2 *>(T), nullptr, counter, (cudaStream_t)stream);
}

extern "C" int vm_aug_tps(const uint8_t *fg_bgra, const void *T, int nx, int ny, const vm_axis_entry *rows,
                          const vm_axis_entry *cols, int n, int h, int w, void *inter, double *alpha64, int32_t *status, void *stream) {
    VM_REQUIRE(fg_bgra && T && rows && cols && inter, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && h >= 2 && w >= 2 && (int64_t)(h + 1) * (w + 1) < (1ll << 28), "bad size");
    if (n == 0) return VM_OK;
    const dim3 grid((w + 1 + 127) / 128, (h + 1 + VA_ROWS - 1) / VA_ROWS, n);
    VM_REQUIRE(grid.y <= 65535, "frame too tall");
#define VA_TPS(U) k_aug_tps<U><<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint32_t *>(fg_bgra), reinterpret_cast<const double2 *>(T), nx, ny, \
                                                       rows, cols, h, w, reinterpret_cast<uint2 *>(inter), alpha64, status)
    if (g_va_minb == 8) VA_TPS(8); else VA_TPS(12);
#undef VA_TPS
    return vm_check_launch("vm_aug_tps");
}
