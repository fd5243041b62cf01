// cv2.warpAffine (legacy fixed-point path), HSV illumination jitter and alpha statistics.
// Reference behaviour: augmentation.py:10-21, 44-63, 88-99 (see include/vm_b200.h).
#include "vm_common.cuh"
#include <math.h>

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

__global__ void __launch_bounds__(256)
k_illumination(const uint8_t *__restrict__ bgr, int64_t npx, VmLut256 lut, uint8_t *__restrict__ out) {
    __shared__ int sdiv[256], hdiv[256];
    __shared__ uint8_t slut[256];
    {
        const int k = threadIdx.x;
        sdiv[k] = k ? __double2int_rn(1044480.0 / (double)k) : 0;            // 255 << 12
        hdiv[k] = k ? __double2int_rn(737280.0 / (6.0 * (double)k)) : 0;     // 180 << 12
        slut[k] = lut.v[k];
    }
    __syncthreads();
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int b = bgr[p * 3], g = bgr[p * 3 + 1], r = bgr[p * 3 + 2];
        const int v = max(max(b, g), r), vmin = min(min(b, g), r), d = v - vmin;
        int hh = (v == r) ? (g - b) : ((v == g) ? (b - r + 2 * d) : (r - g + 4 * d));
        const int s = (d * sdiv[v] + (1 << 11)) >> 12;
        hh = (hh * hdiv[d] + (1 << 11)) >> 12;
        if (hh < 0) hh += 180;
        const int s2 = slut[s], v2 = slut[v];
        // HSV2BGR, float formulation (OpenCV HSV2RGB_f with hscale = 6/180)
        const float fv = (float)v2 * (1.f / 255.f), fs = (float)s2 * (1.f / 255.f);
        float ob = fv, og = fv, orr = fv;
        if (s2 != 0) {
            float hf = (float)hh * (6.f / 180.f);
            const float fl = floorf(hf);
            int sec = (int)fl;
            hf -= fl;
            sec %= 6; if (sec < 0) sec += 6;
            float tab[4];
            tab[0] = fv;
            tab[1] = __fmul_rn(fv, 1.f - fs);
            tab[2] = __fmul_rn(fv, 1.f - __fmul_rn(fs, hf));
            tab[3] = __fmul_rn(fv, 1.f - __fmul_rn(fs, 1.f - hf));
            const int ib[6] = {1, 1, 3, 0, 0, 2}, ig[6] = {3, 0, 0, 2, 1, 1}, ir[6] = {0, 2, 1, 1, 3, 0};
            ob = tab[ib[sec]]; og = tab[ig[sec]]; orr = tab[ir[sec]];
        }
        out[p * 3] = (uint8_t)max(0, min(255, __float2int_rn(ob * 255.f)));
        out[p * 3 + 1] = (uint8_t)max(0, min(255, __float2int_rn(og * 255.f)));
        out[p * 3 + 2] = (uint8_t)max(0, min(255, __float2int_rn(orr * 255.f)));
    }
}

extern "C" int vm_illumination_lut(const uint8_t *bgr, int64_t npx, const uint8_t *lut_host,
                                   uint8_t *out, void *stream) {
    VM_REQUIRE(bgr && lut_host && out, "null pointer");
    if (npx <= 0) return VM_OK;
    VmLut256 lut;
    for (int k = 0; k < 256; ++k) lut.v[k] = lut_host[k];
    k_illumination<<<min(vm_blocks(npx, 256), 148u * 16u), 256, 0, (cudaStream_t)stream>>>(bgr, npx, lut, out);
    return vm_check_launch("vm_change_illumination");
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
