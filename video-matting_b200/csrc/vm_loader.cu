// Training-sample assembly of the reference batch loader (loader.py:39-85, 119-157, 285-330) for a
// whole batch in one launch: flow warp of the previous alpha (flow.py:9-18), zero-padded canvas +
// random crop (loader.py:10-36, 294-309), cv2.resize(INTER_LINEAR) of the float64 crop
// (loader.py:316-319), composite (reader.py:72-79) and VGG-mean subtraction (loader.py:322-323).
//
// Nothing is materialised between those steps: a thread owns one output pixel of one sample,
// walks back through the resize taps -> crop window -> canvas -> image, and evaluates the flow
// warp only at the (at most four) image pixels its alpha taps land on.  The kernel is bound by
// its float64 output stores (13 values per pixel); inputs are the decoded uint8 files.
#include "vm_common.cuh"
#include <math.h>

static_assert(sizeof(vm_loader_view) == 64 && sizeof(vm_loader_sample) == 200, "descriptor layout (host: loader.py SAMPLE_DTYPE)");

// one tap position of a resize: two source indices per axis and the weight pair
struct VmResizeTaps { int y0, y1, x0, x1; double fy, fx; bool area; };

// OpenCV resize.cpp, INTER_LINEAR on a 64-bit image: double coefficients, pixel centres aligned;
// columns beyond the first / last sample centre replicate with the fraction forced to 0, rows are
// clamped with the fraction kept.  Exactly-2x reductions use the 2x2 block mean (INTER_AREA).
__device__ __forceinline__ VmResizeTaps vm_resize_taps(const vm_loader_view &v, int y, int x) {
    VmResizeTaps t;
    t.area = v.mode == 1;
    if (t.area) {
        t.y0 = 2 * y; t.y1 = 2 * y + 1; t.x0 = 2 * x; t.x1 = 2 * x + 1;
        t.fy = t.fx = 0.0;
        return t;
    }
    const double fxd = __dsub_rn(__dmul_rn(__dadd_rn((double)x, 0.5), v.scale_x), 0.5);
    const double fyd = __dsub_rn(__dmul_rn(__dadd_rn((double)y, 0.5), v.scale_y), 0.5);
    const double sxf = floor(fxd), syf = floor(fyd);
    int sx = (int)sxf, sy = (int)syf;
    t.fx = __dsub_rn(fxd, sxf);
    t.fy = __dsub_rn(fyd, syf);
    if (sx < 0) { sx = 0; t.fx = 0.0; }
    if (sx >= v.win_w - 1) { sx = v.win_w - 1; t.fx = 0.0; }
    t.x0 = sx; t.x1 = min(sx + 1, v.win_w - 1);
    t.y0 = max(0, min(sy, v.win_h - 1));
    t.y1 = max(0, min(sy + 1, v.win_h - 1));
    return t;
}

__device__ __forceinline__ double vm_resize_mix(const VmResizeTaps &t, double s00, double s01, double s10, double s11) {
    if (t.area) return __dmul_rn(__dadd_rn(__dadd_rn(__dadd_rn(s00, s01), s10), s11), 0.25);
    const double a0 = __dsub_rn(1.0, t.fx), b0 = __dsub_rn(1.0, t.fy);
    const double h0 = __dadd_rn(__dmul_rn(s00, a0), __dmul_rn(s01, t.fx));
    const double h1 = __dadd_rn(__dmul_rn(s10, a0), __dmul_rn(s11, t.fx));
    return __dadd_rn(__dmul_rn(h0, b0), __dmul_rn(h1, t.fy));
}

// window coordinates -> image pixel index, or -1 where the canvas holds padding zeros
__device__ __forceinline__ int64_t vm_view_pixel(const vm_loader_view &v, int r, int c, int img_w) {
    const int cr = v.wi + r, cc = v.wj + c;
    if (cr < v.vi0 || cr >= v.vi1 || cc < v.vj0 || cc >= v.vj1) return -1;
    return (int64_t)(cr - v.vi0 + v.si) * img_w + (cc - v.vj0 + v.sj);
}

// flow.warp_img(prev_alpha, flow) at element px of the staged rectangle, prev_alpha = A / 255. (float64),
// flow.py:9-18; map coordinates and taps are in full-image coordinates
__device__ __forceinline__ double vm_warped_alpha_at(const vm_loader_sample &S, int64_t px) {
    const int i = (int)(px / S.fw), j = (int)(px - (int64_t)i * S.fw);
    const float2 d = __ldg(reinterpret_cast<const float2 *>(S.flow) + px);
    const int SX = vm_cvround_x32(vm_map_coord(j + S.ox, d.x)), SY = vm_cvround_x32(vm_map_coord(i + S.oy, d.y));
    const int ix = vm_sat_s16(SX >> 5), iy = vm_sat_s16(SY >> 5);
    const int H = S.ph, W = S.pw;
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const int64_t ps = S.prev_stride;
    const uint8_t *r0 = S.prev + ((int64_t)iy * W + ix) * ps, *r1 = r0 + (int64_t)W * ps;
    const double s00 = (y0 && x0) ? __ddiv_rn((double)__ldg(r0), 255.0) : 0.0;
    const double s01 = (y0 && x1) ? __ddiv_rn((double)__ldg(r0 + ps), 255.0) : 0.0;
    const double s10 = (y1 && x0) ? __ddiv_rn((double)__ldg(r1), 255.0) : 0.0;
    const double s11 = (y1 && x1) ? __ddiv_rn((double)__ldg(r1 + ps), 255.0) : 0.0;
    return VmTap<double>::blend(s00, s01, s10, s11, SX & 31, SY & 31);
}

template <typename OT>
__global__ void __launch_bounds__(256)
k_loader_batch(const vm_loader_sample *__restrict__ samples, int out_h, int out_w, double m0, double m1, double m2,
               OT *__restrict__ o_cmp, OT *__restrict__ o_bg, OT *__restrict__ o_label, OT *__restrict__ o_warped,
               OT *__restrict__ o_fg) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= out_w || y >= out_h) return;
    const vm_loader_sample S = samples[blockIdx.z];

    // foreground window: B, G, R, alpha and (video loader) the warped previous alpha
    const VmResizeTaps t = vm_resize_taps(S.fgv, y, x);
    const int rr[4] = {t.y0, t.y0, t.y1, t.y1}, cc[4] = {t.x0, t.x1, t.x0, t.x1};
    double sb[4], sg[4], sr[4], sa[4], sw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t px = vm_view_pixel(S.fgv, rr[k], cc[k], S.fw);
        sb[k] = sg[k] = sr[k] = sa[k] = sw[k] = 0.0;
        if (px >= 0) {
            const uint32_t p = __ldg(reinterpret_cast<const uint32_t *>(S.fg) + px);
            sb[k] = (double)(p & 255u);
            sg[k] = (double)((p >> 8) & 255u);
            sr[k] = (double)((p >> 16) & 255u);
            sa[k] = __ddiv_rn((double)(p >> 24), 255.0);                  // reader.py:16
            if (S.prev) sw[k] = vm_warped_alpha_at(S, px);
        }
    }
    const double fb = vm_resize_mix(t, sb[0], sb[1], sb[2], sb[3]);
    const double fg = vm_resize_mix(t, sg[0], sg[1], sg[2], sg[3]);
    const double fr = vm_resize_mix(t, sr[0], sr[1], sr[2], sr[3]);
    const double al = vm_resize_mix(t, sa[0], sa[1], sa[2], sa[3]);

    // background window
    const VmResizeTaps u = vm_resize_taps(S.bgv, y, x);
    const int br[4] = {u.y0, u.y0, u.y1, u.y1}, bc[4] = {u.x0, u.x1, u.x0, u.x1};
    double gb[4], gg[4], gr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t px = vm_view_pixel(S.bgv, br[k], bc[k], S.bw);
        gb[k] = gg[k] = gr[k] = 0.0;
        if (px >= 0) {
            const uint8_t *p = S.bg + px * 3;
            gb[k] = (double)__ldg(p); gg[k] = (double)__ldg(p + 1); gr[k] = (double)__ldg(p + 2);
        }
    }
    const double bb = vm_resize_mix(u, gb[0], gb[1], gb[2], gb[3]);
    const double bg = vm_resize_mix(u, gg[0], gg[1], gg[2], gg[3]);
    const double br_ = vm_resize_mix(u, gr[0], gr[1], gr[2], gr[3]);

    // reader.py:72-79: tri * fg + (1 - tri) * bg, then the mean subtraction of loader.py:322-323
    const double na = __dsub_rn(1.0, al);
    const double cb = __dadd_rn(__dmul_rn(al, fb), __dmul_rn(na, bb));
    const double cg = __dadd_rn(__dmul_rn(al, fg), __dmul_rn(na, bg));
    const double cr = __dadd_rn(__dmul_rn(al, fr), __dmul_rn(na, br_));

    const int xo = S.flip ? out_w - 1 - x : x;                           // loader.py:107-110 (rd_mirror)
    const int64_t o = ((int64_t)blockIdx.z * out_h + y) * out_w + xo;
    if (o_cmp) { o_cmp[o * 3] = (OT)__dsub_rn(cb, m0); o_cmp[o * 3 + 1] = (OT)__dsub_rn(cg, m1); o_cmp[o * 3 + 2] = (OT)__dsub_rn(cr, m2); }
    if (o_bg) { o_bg[o * 3] = (OT)__dsub_rn(bb, m0); o_bg[o * 3 + 1] = (OT)__dsub_rn(bg, m1); o_bg[o * 3 + 2] = (OT)__dsub_rn(br_, m2); }
    if (o_label) o_label[o] = (OT)al;
    if (o_fg) { o_fg[o * 3] = (OT)fb; o_fg[o * 3 + 1] = (OT)fg; o_fg[o * 3 + 2] = (OT)fr; }
    if (o_warped) {
        const OT wv = S.prev ? (OT)vm_resize_mix(t, sw[0], sw[1], sw[2], sw[3]) : (OT)0;
        o_warped[o * 3] = wv; o_warped[o * 3 + 1] = wv; o_warped[o * 3 + 2] = wv;
    }
}

extern "C" int vm_loader_batch(const vm_loader_sample *samples, int n, int out_h, int out_w, const double *mean_host,
                               int out_dtype, void *cmp, void *bg, void *label, void *warped, void *fg, void *stream) {
    VM_REQUIRE(samples && mean_host, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && out_h >= 1 && out_w >= 1, "bad size");
    if (n == 0) return VM_OK;
    dim3 grid((out_w + 31) / 32, (out_h + 7) / 8, n), block(32, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == VM_F64)
        k_loader_batch<double><<<grid, block, 0, st>>>(samples, out_h, out_w, mean_host[0], mean_host[1], mean_host[2],
                                                       (double *)cmp, (double *)bg, (double *)label, (double *)warped, (double *)fg);
    else if (out_dtype == VM_F32)
        k_loader_batch<float><<<grid, block, 0, st>>>(samples, out_h, out_w, mean_host[0], mean_host[1], mean_host[2],
                                                      (float *)cmp, (float *)bg, (float *)label, (float *)warped, (float *)fg);
    else { vm_set_error("vm_loader_batch: out_dtype must be VM_F32 or VM_F64"); return VM_ERR_ARG; }
    return vm_check_launch("vm_loader_batch");
}

// loader.psnr (loader.py:214-227): sum over every element of (a - b)^2 into one double
template <typename T>
__global__ void __launch_bounds__(256) k_sq_err(const T *__restrict__ a, const T *__restrict__ b, int64_t n, double *__restrict__ out) {
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d = (double)a[i] - (double)b[i];
        acc += d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < 8; ++k) s += part[k];
        atomicAdd(out, s);
    }
}

extern "C" int vm_sq_err_sum(const void *a, const void *b, int dtype, int64_t n, double *out, void *stream) {
    VM_REQUIRE(a && b && out && n >= 0, "bad argument");
    if (n == 0) return VM_OK;
    const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);   // 148 SMs x 8
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VM_F64) k_sq_err<double><<<blocks, 256, 0, st>>>((const double *)a, (const double *)b, n, out);
    else if (dtype == VM_F32) k_sq_err<float><<<blocks, 256, 0, st>>>((const float *)a, (const float *)b, n, out);
    else { vm_set_error("vm_sq_err_sum: dtype must be VM_F32 or VM_F64"); return VM_ERR_ARG; }
    return vm_check_launch("vm_sq_err_sum");
}

// ---------------------------------------------------------------------------------------
// data.trimap_from_matte (data.py:37-67).  The reference scans in raster order: every pixel first
// receives its own class (255 / 0 / 128), and a fractional pixel then paints 128 over the alpha==1
// pixels within +-3 and the alpha==0 pixels within +-1.  Paint that lands on a pixel before the scan
// reaches it is overwritten, so: a 1-pixel (0-pixel) ends up 128 iff a fractional pixel within
// Chebyshev distance 3 (1) comes LATER in raster order.  One thread per pixel; the 8 x 32 tile and
// its halo (3 rows below, 3 columns either side) are classified once into shared memory.
// ---------------------------------------------------------------------------------------
#define TM_W 32
#define TM_H 8
#define TM_R 3

template <typename T> __device__ __forceinline__ int vm_matte_class(T v);
template <> __device__ __forceinline__ int vm_matte_class<double>(double v) { return v == 1.0 ? 1 : (v == 0.0 ? 0 : 2); }
template <> __device__ __forceinline__ int vm_matte_class<uint8_t>(uint8_t v) { return v == 255 ? 1 : (v == 0 ? 0 : 2); }

template <typename T>
__global__ void __launch_bounds__(TM_W * TM_H)
k_trimap(const T *__restrict__ matte, int h, int w, uint8_t *__restrict__ out) {
    __shared__ uint8_t frac[TM_H + TM_R][TM_W + 2 * TM_R];      // 1 where the pixel is fractional (in frame)
    const int x0 = blockIdx.x * TM_W, y0 = blockIdx.y * TM_H;
    const int64_t plane = (int64_t)blockIdx.z * h * w;
    for (int k = threadIdx.y * TM_W + threadIdx.x; k < (TM_H + TM_R) * (TM_W + 2 * TM_R); k += TM_W * TM_H) {
        const int r = k / (TM_W + 2 * TM_R), c = k - r * (TM_W + 2 * TM_R);
        const int y = y0 + r, x = x0 + c - TM_R;
        int cls = 0;
        if (y < h && x >= 0 && x < w) cls = vm_matte_class<T>(__ldg(matte + plane + (int64_t)y * w + x));
        frac[r][c] = cls == 2;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int own = vm_matte_class<T>(__ldg(matte + plane + (int64_t)y * w + x));
    uint8_t v = own == 2 ? 128 : (own == 1 ? 255 : 0);
    if (own != 2) {
        const int R = own == 1 ? 3 : 1;                         // crop = 3 for alpha == 1, dilate = 1 for alpha == 0
        bool later = false;
        for (int dk = 0; dk <= R; ++dk)
            for (int dl = -R; dl <= R; ++dl) {
                if (dk == 0 && dl <= 0) continue;               // same row: only pixels to the right come later
                later |= frac[threadIdx.y + dk][threadIdx.x + TM_R + dl] != 0;
            }
        if (later) v = 128;
    }
    out[plane + (int64_t)y * w + x] = v;
}

extern "C" int vm_trimap_from_matte(const void *matte, int dtype, int n, int h, int w, uint8_t *out, void *stream) {
    VM_REQUIRE(matte && out, "null pointer");
    VM_REQUIRE(n >= 0 && n < 65536 && h >= 1 && w >= 1 && (h + TM_H - 1) / TM_H < 65536, "bad size");
    if (n == 0) return VM_OK;
    dim3 grid((w + TM_W - 1) / TM_W, (h + TM_H - 1) / TM_H, n), block(TM_W, TM_H);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == VM_F64) k_trimap<double><<<grid, block, 0, st>>>((const double *)matte, h, w, out);
    else if (dtype == VM_U8) k_trimap<uint8_t><<<grid, block, 0, st>>>((const uint8_t *)matte, h, w, out);
    else { vm_set_error("vm_trimap_from_matte: dtype must be VM_U8 or VM_F64"); return VM_ERR_ARG; }
    return vm_check_launch("vm_trimap_from_matte");
}

// ---------------------------------------------------------------------------------------
// reader.read_fg_img, uint16 branch (reader.py:13-15): (((img + 1) / 256.) - 1).astype(uint8) where
// img + 1 wraps in uint16 and the cast of -1.0 wraps to 255.  With v' = (v + 1) mod 2^16:
// v' = 0 -> 255, 1..255 -> 0 (a value in (-1, 0) truncates to 0), otherwise (v' >> 8) - 1.
// Eight elements per thread (16-byte loads, 8-byte stores); the tail is done element-wise.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t vm_u16_quirk(uint32_t v) {
    const uint32_t q = (v + 1u) & 0xFFFFu;
    return q == 0u ? 255u : (q < 256u ? 0u : (q >> 8) - 1u);
}

// vec: src is 16-byte and dst 8-byte aligned (vector accesses); otherwise element by element
__global__ void __launch_bounds__(256) k_fg_from_u16(const uint16_t *__restrict__ src, int64_t n, uint8_t *__restrict__ dst, int vec) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
    if (i >= n) return;
    if (vec && i + 8 <= n) {
        const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(src + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        uint32_t o[2] = {0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[k >> 1] |= vm_u16_quirk(w[k] & 0xFFFFu) << (16 * (k & 1));
            o[k >> 1] |= vm_u16_quirk(w[k] >> 16) << (16 * (k & 1) + 8);
        }
        __stcs(reinterpret_cast<uint2 *>(dst + i), make_uint2(o[0], o[1]));
    } else {
        for (int64_t k = i; k < n && k < i + 8; ++k) dst[k] = (uint8_t)vm_u16_quirk(src[k]);
    }
}

extern "C" int vm_fg_from_u16(const uint16_t *src, int64_t n, uint8_t *dst, void *stream) {
    VM_REQUIRE(src && dst && n >= 0, "bad argument");
    VM_REQUIRE((reinterpret_cast<uintptr_t>(src) & 1) == 0, "source is not 2-byte aligned");
    if (n == 0) return VM_OK;
    // any destination offset works (reader.load_clip converts into slices of a clip: odd frame sizes give 4-byte
    // aligned slices); the vector path needs 16 / 8 byte alignment
    const int vec = (reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0;
    const int64_t threads = (n + 7) / 8;
    k_fg_from_u16<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, dst, vec);
    return vm_check_launch("vm_fg_from_u16");
}
