// Flow warp, forward/backward consistency mask and compositing kernels.
// Reference behaviour: flow.py:9-65, reader.py:72-79 (see include/vm_b200.h).
#include "vm_common.cuh"

// ---------------------------------------------------------------------------------------
// generic drop-in kernels (any size, any of uint8/float32/float64, 1/3/4 channels)
// ---------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256)
k_flow_warp(const T *__restrict__ src, int sh, int sw, const float2 *__restrict__ flow,
            int h, int w, T *__restrict__ dst) {
    const int64_t n = (int64_t)h * w;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
        const float2 f = __ldg(flow + p);
        const int SX = vm_cvround_x32(vm_map_coord(j, f.x));
        const int SY = vm_cvround_x32(vm_map_coord(i, f.y));
        T out[C];
        vm_sample_fixed<T, C>(src, sh, sw, SX, SY, out);
#pragma unroll
        for (int c = 0; c < C; ++c) dst[p * C + c] = out[c];
    }
}

template <typename T>
static int launch_flow_warp(const void *src, int channels, int sh, int sw, const float *flow,
                            int h, int w, void *dst, cudaStream_t st) {
    const int64_t n = (int64_t)h * w;
    const unsigned grid = min(vm_blocks(n, 256), 148u * 32u);
    const float2 *f2 = reinterpret_cast<const float2 *>(flow);
    switch (channels) {
    case 1: k_flow_warp<T, 1><<<grid, 256, 0, st>>>((const T *)src, sh, sw, f2, h, w, (T *)dst); break;
    case 3: k_flow_warp<T, 3><<<grid, 256, 0, st>>>((const T *)src, sh, sw, f2, h, w, (T *)dst); break;
    case 4: k_flow_warp<T, 4><<<grid, 256, 0, st>>>((const T *)src, sh, sw, f2, h, w, (T *)dst); break;
    default: vm_set_error("vm_flow_warp: channels must be 1, 3 or 4"); return VM_ERR_ARG;
    }
    return vm_check_launch("vm_flow_warp");
}

extern "C" int vm_flow_warp(const void *src, int dtype, int channels, int sh, int sw,
                            const float *flow, int h, int w, void *dst, void *stream) {
    VM_REQUIRE(src && flow && dst, "null pointer");
    VM_REQUIRE(sh > 0 && sw > 0 && h >= 0 && w >= 0, "bad size");
    if ((int64_t)h * w == 0) return VM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case VM_U8:  return launch_flow_warp<uint8_t>(src, channels, sh, sw, flow, h, w, dst, st);
    case VM_F32: return launch_flow_warp<float>(src, channels, sh, sw, flow, h, w, dst, st);
    case VM_F64: return launch_flow_warp<double>(src, channels, sh, sw, flow, h, w, dst, st);
    }
    vm_set_error("vm_flow_warp: unsupported dtype %d", dtype);
    return VM_ERR_ARG;
}

__global__ void __launch_bounds__(256)
k_occlusion_mask(const float2 *__restrict__ bwd, const float2 *__restrict__ fwd, int h, int w,
                 uint8_t *__restrict__ mask, int32_t *__restrict__ status) {
    const int64_t n = (int64_t)h * w;
    int idx_err = 0, nan_err = 0, masked = 0;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / w), j = (int)(p - (int64_t)i * w);
        int flags = 0;
        const int m = vm_consistency(fwd, h, w, i, j, __ldg(bwd + p), flags);
        mask[p] = (uint8_t)(m && !flags);
        idx_err += flags & 1; nan_err += (flags >> 1) & 1; masked += (m && !flags);
    }
    if (status) {
        if (idx_err) atomicAdd(status + VM_STATUS_INDEX_ERR, idx_err);
        if (nan_err) atomicAdd(status + VM_STATUS_NAN_ERR, nan_err);
        if (masked) atomicAdd(status + VM_STATUS_MASKED, masked);
    }
}

extern "C" int vm_occlusion_mask(const float *backward, const float *forward, int h, int w,
                                 uint8_t *mask, int32_t *status, void *stream) {
    VM_REQUIRE(backward && forward && mask, "null pointer");
    VM_REQUIRE(h >= 0 && w >= 0, "bad size");
    const int64_t n = (int64_t)h * w;
    if (n == 0) return VM_OK;
    k_occlusion_mask<<<min(vm_blocks(n, 256), 148u * 32u), 256, 0, (cudaStream_t)stream>>>(
        (const float2 *)backward, (const float2 *)forward, h, w, mask, status);
    return vm_check_launch("vm_occlusion_mask");
}

template <typename T>
__global__ void k_apply_mask(T *__restrict__ alpha, const uint8_t *__restrict__ mask, int64_t n) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x)
        if (mask[p]) alpha[p] = T(0);
}

extern "C" int vm_apply_mask(void *alpha, int dtype, const uint8_t *mask, int64_t n, void *stream) {
    VM_REQUIRE(alpha && mask, "null pointer");
    if (n <= 0) return VM_OK;
    const unsigned grid = min(vm_blocks(n, 256), 148u * 32u);
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
    case VM_U8:  k_apply_mask<uint8_t><<<grid, 256, 0, st>>>((uint8_t *)alpha, mask, n); break;
    case VM_F32: k_apply_mask<float><<<grid, 256, 0, st>>>((float *)alpha, mask, n); break;
    case VM_F64: k_apply_mask<double><<<grid, 256, 0, st>>>((double *)alpha, mask, n); break;
    default: vm_set_error("vm_apply_mask: unsupported dtype %d", dtype); return VM_ERR_ARG;
    }
    return vm_check_launch("vm_apply_mask");
}

// reader.py:72-79: tri*fg + (1-tri)*bg in float64, no FMA contraction.
template <typename TF, typename TB>
__global__ void k_composite(const TF *__restrict__ fg, const TB *__restrict__ bg,
                            const double *__restrict__ alpha, int64_t npx, double *__restrict__ out) {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npx;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double a = alpha[p], na = __dadd_rn(1.0, -a);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            out[p * 3 + c] = __dadd_rn(__dmul_rn(a, (double)fg[p * 3 + c]), __dmul_rn(na, (double)bg[p * 3 + c]));
    }
}

extern "C" int vm_composite(const void *fg, int fg_dtype, const void *bg, int bg_dtype,
                            const double *alpha, int h, int w, double *out, void *stream) {
    VM_REQUIRE(fg && bg && alpha && out, "null pointer");
    const int64_t n = (int64_t)h * w;
    if (n <= 0) return VM_OK;
    const unsigned grid = min(vm_blocks(n, 256), 148u * 32u);
    cudaStream_t st = (cudaStream_t)stream;
#define VM_CMP(TF, TB) k_composite<TF, TB><<<grid, 256, 0, st>>>((const TF *)fg, (const TB *)bg, alpha, n, out)
    if (fg_dtype == VM_U8 && bg_dtype == VM_U8) VM_CMP(uint8_t, uint8_t);
    else if (fg_dtype == VM_U8 && bg_dtype == VM_F64) VM_CMP(uint8_t, double);
    else if (fg_dtype == VM_F64 && bg_dtype == VM_U8) VM_CMP(double, uint8_t);
    else if (fg_dtype == VM_F64 && bg_dtype == VM_F64) VM_CMP(double, double);
    else if (fg_dtype == VM_F32 && bg_dtype == VM_F32) VM_CMP(float, float);
    else if (fg_dtype == VM_U8 && bg_dtype == VM_F32) VM_CMP(uint8_t, float);
    else if (fg_dtype == VM_F32 && bg_dtype == VM_U8) VM_CMP(float, uint8_t);
    else { vm_set_error("vm_composite: unsupported dtype pair %d/%d", fg_dtype, bg_dtype); return VM_ERR_ARG; }
#undef VM_CMP
    return vm_check_launch("vm_composite");
}

// ---------------------------------------------------------------------------------------
// fused C2 kernel: warp_bgr + warp_img + correct_alpha on BGRA frames, 27 B/px.
//
// CTA = 256 threads = 8 warps; tile = 128 px wide x 8 rows; each thread owns 4 consecutive
// pixels of one row: two 16-byte flow loads, 16 BGRA taps through the read-only path (tap rows
// of vertically adjacent warps overlap in L1), 4 nearest forward-flow gathers, then one
// 16-byte alpha store and 12 bytes of BGR (three 4-byte stores; a thread's 4 px * 3 B are
// 4-byte aligned because the thread's first pixel index is a multiple of 4).
// ---------------------------------------------------------------------------------------
// (C2_* tile constants, vm_alpha_f32, vm_pack_alpha and the unit body vm_flow_unit live in vm_common.cuh)

#ifndef C2_MINB
#define C2_MINB 5                  /* CTAs of 256 threads per SM (48 registers) */
#endif
template <bool HAS_FWD, int PACKED>
__global__ void __launch_bounds__(256, C2_MINB)
k_flow_warp_mask_bgra(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd,
                      const float2 *__restrict__ fwd, int h, int w, int tiles_x, int tiles_y,
                      uint8_t *__restrict__ out_bgr, float *__restrict__ out_alpha,
                      int32_t *__restrict__ status) {
    vm_flow_unit<HAS_FWD, PACKED>(fg, bwd, fwd, h, w, blockIdx.z, blockIdx.y, blockIdx.x, out_bgr, out_alpha, status);   // 3-D grid: no integer division
}

// Packed-output variant used by the C4 pipelines: thread = one column x 4 consecutive rows, so
// every tap load of a warp covers 32 neighbouring pixels (one or two cache lines instead of four) and
// vertically adjacent taps of the same thread reuse L1 lines.  CTA = 32 columns x C2P_ROWS rows;
// warp k owns rows 4k .. 4k+3 of every 32-row group.
#define C2P_ROWS 96
template <bool HAS_FWD, int PACKED>
__global__ void __launch_bounds__(256, 5)
k_flow_stage_packed(const uint8_t *__restrict__ fg, const float2 *__restrict__ bwd, const float2 *__restrict__ fwd,
                    int h, int w, uint2 *__restrict__ out, int32_t *__restrict__ status) {
    const int frame = blockIdx.z;
    const int j = blockIdx.x * 32 + (threadIdx.x & 31);
    const int rbeg = blockIdx.y * C2P_ROWS + (threadIdx.x >> 5) * 4, rend = min((blockIdx.y + 1) * C2P_ROWS, h);
    if (j >= w || rbeg >= rend) return;
    const int64_t fbase = (int64_t)frame * h * w;
    const uint32_t *fg32 = reinterpret_cast<const uint32_t *>(fg) + fbase;
    const float2 *bf = bwd + fbase + j;
    const float2 *ff = HAS_FWD ? fwd + fbase : nullptr;
    uint2 *op = out + fbase + j;
    asm volatile("" : "+l"(fg32));
    if (HAS_FWD) asm volatile("" : "+l"(ff));
    const float fj = (float)j;
    int flags = 0;
    float2 nx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) nx[k] = __ldcs(bf + (unsigned)(min(rbeg + k, h - 1) * w));      // streamed once: evict first
    for (int i = rbeg; i < rend; i += 32) {
        float2 fl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) fl[k] = nx[k];
        if (i + 32 < rend) {
#pragma unroll
            for (int k = 0; k < 4; ++k) nx[k] = __ldcs(bf + (unsigned)(min(i + 32 + k, h - 1) * w));
        }
        VmFlowPx px[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ii = min(i + k, h - 1);
            vm_flow_pxn<HAS_FWD, 1>(fg32, ff, h, w, ii, j, (float)ii, fj, &fl[k], &px[k], flags);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i + k < rend) op[(unsigned)((i + k) * w)] = make_uint2(px[k].bgr, vm_pack_alpha<PACKED>(px[k]));
    }
    if (HAS_FWD && flags && status) {
        if (flags & 1) atomicAdd(status + VM_STATUS_INDEX_ERR, 1);
        if (flags & 2) atomicAdd(status + VM_STATUS_NAN_ERR, 1);
    }
}

// stage A of the split C4 pipelines: (n,h,w) uint2 {bgr, alpha code} or, raw_ta, {bgr, TA}
int g_vm_flow_stage_layout = 0;      // 1: column x 4 rows per thread (k_flow_stage_packed); 0: 4 columns per thread

int vm_launch_flow_stage(const uint8_t *fg, const float *backward, const float *forward, int n, int h, int w,
                         void *packed, int32_t *status, cudaStream_t st, bool raw_ta) {
    const float2 *b2 = (const float2 *)backward, *f2 = (const float2 *)forward;
    VM_REQUIRE(vm_aligned(fg, 4) && vm_aligned(backward, (w & 3) ? 8 : 16) && vm_aligned(forward, 8) &&
               vm_aligned(packed, (w & 3) ? 8 : 16), "unaligned buffer");
    if (g_vm_flow_stage_layout == 1) {
        const dim3 grid((w + 31) / 32, (h + C2P_ROWS - 1) / C2P_ROWS, n);
        uint2 *o = (uint2 *)packed;
        if (forward && raw_ta) k_flow_stage_packed<true, 2><<<grid, 256, 0, st>>>(fg, b2, f2, h, w, o, status);
        else if (forward)      k_flow_stage_packed<true, 1><<<grid, 256, 0, st>>>(fg, b2, f2, h, w, o, status);
        else if (raw_ta)       k_flow_stage_packed<false, 2><<<grid, 256, 0, st>>>(fg, b2, nullptr, h, w, o, status);
        else                   k_flow_stage_packed<false, 1><<<grid, 256, 0, st>>>(fg, b2, nullptr, h, w, o, status);
        return vm_check_launch("vm_flow_stage");
    }
    const int tiles_x = (w + C2_TW - 1) / C2_TW, tiles_y = (h + C2_ROWS - 1) / C2_ROWS;
    const dim3 tiles(tiles_x, tiles_y, n);
    uint8_t *o = (uint8_t *)packed;
    if (forward && raw_ta) k_flow_warp_mask_bgra<true, 2><<<tiles, 256, 0, st>>>(fg, b2, f2, h, w, tiles_x, tiles_y, o, nullptr, status);
    else if (forward)      k_flow_warp_mask_bgra<true, 1><<<tiles, 256, 0, st>>>(fg, b2, f2, h, w, tiles_x, tiles_y, o, nullptr, status);
    else if (raw_ta)       k_flow_warp_mask_bgra<false, 2><<<tiles, 256, 0, st>>>(fg, b2, nullptr, h, w, tiles_x, tiles_y, o, nullptr, status);
    else                   k_flow_warp_mask_bgra<false, 1><<<tiles, 256, 0, st>>>(fg, b2, nullptr, h, w, tiles_x, tiles_y, o, nullptr, status);
    return vm_check_launch("vm_flow_stage");
}

extern "C" int vm_flow_warp_mask_bgra(const uint8_t *fg, const float *backward, const float *forward,
                                      int n, int h, int w, uint8_t *out_bgr, float *out_alpha,
                                      int32_t *status, void *stream) {
    if (n == 0) return VM_OK;                                          // empty clip
    VM_REQUIRE(fg && backward && out_bgr && out_alpha, "null pointer");
    VM_REQUIRE(n >= 0 && h > 0 && w > 0 && h <= 32767 && w <= 32767, "bad size");
    if (n == 0) return VM_OK;
    // the 4-pixel path (w % 4 == 0) uses 16-byte flow loads / alpha stores and 4-byte BGRA / BGR accesses
    VM_REQUIRE(vm_aligned(fg, 4) && vm_aligned(backward, (w & 3) ? 8 : 16) && vm_aligned(forward, 8) &&
               vm_aligned(out_alpha, (w & 3) ? 4 : 16) && ((w & 3) || vm_aligned(out_bgr, 4)), "unaligned buffer");
    const int tiles_x = (w + C2_TW - 1) / C2_TW, tiles_y = (h + C2_ROWS - 1) / C2_ROWS;
    VM_REQUIRE(tiles_y <= 65535 && n <= 65535, "too many tiles for one launch");
    const dim3 tiles(tiles_x, tiles_y, n);
    cudaStream_t st = (cudaStream_t)stream;
    if (forward)
        k_flow_warp_mask_bgra<true, 0><<<tiles, 256, 0, st>>>(
            fg, (const float2 *)backward, (const float2 *)forward, h, w, tiles_x, tiles_y, out_bgr, out_alpha, status);
    else
        k_flow_warp_mask_bgra<false, 0><<<tiles, 256, 0, st>>>(
            fg, (const float2 *)backward, nullptr, h, w, tiles_x, tiles_y, out_bgr, out_alpha, status);
    return vm_check_launch("vm_flow_warp_mask_bgra");
}
