// Shared helpers for libmpn_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../../include/mpn_b200.h"

void mpn_set_error(const char* fmt, ...);

#define MPN_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      mpn_set_error(__VA_ARGS__);           \
      return MPN_ERR_ARG;                   \
    }                                       \
  } while (0)

#define MPN_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      mpn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MPN_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define MPN_LAUNCH_OK() MPN_CUDA_OK(cudaGetLastError())

static inline int mpn_divup(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- activation element access shared by the CUDA-core kernels ---------------------------------
__device__ __forceinline__ float mpn_bf16_bits_to_float(unsigned short b) { return __uint_as_float(((unsigned)b) << 16); }

// ---- MPN_FMT_F16F8: x = fp16 hi + 2^-12 * e5m2 lo (+ an e5m2 copy of x for the tensor core's cross term).  The `lo`
// pointer of that format addresses TWO consecutive byte planes: [lo8 : plane elements][h8 : plane elements].
#define MPN_F8_LO_SCALE 4096.f
#define MPN_F8_LO_INV (1.f / 4096.f)
__device__ __forceinline__ float mpn_e5m2_to_float(unsigned char b) { return __half2float(__ushort_as_half((unsigned short)((unsigned short)b << 8))); }
__device__ __forceinline__ unsigned char mpn_float_to_e5m2(float v) { return (unsigned char)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E5M2); }
__device__ __forceinline__ unsigned char mpn_float_to_e4m3(float v) { return (unsigned char)__nv_cvt_float_to_fp8(v, __NV_SATFINITE, __NV_E4M3); }

// fp32 -> fp16 with saturation to +-65504 (one F2FP.SATFINITE instruction): an activation beyond the fp16 range is clamped
// instead of becoming inf (and NaN once the e5m2 residual plane is added back) -- the overflow guard of MPN_FMT_F16F8.
__device__ __forceinline__ uint32_t mpn_pack_f16x2_sat(float lo_elem, float hi_elem) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
  return r;
}
__device__ __forceinline__ __half mpn_f16_sat(float v) {
  return __ushort_as_half((unsigned short)(mpn_pack_f16x2_sat(v, 0.f) & 0xFFFFu));
}

__device__ __forceinline__ float mpn_load_act(const void* hi, const void* lo, long long idx, int fmt) {
  if (fmt == MPN_FMT_F32) return ((const float*)hi)[idx];
  if (fmt == MPN_FMT_F16F8)
    return __half2float(((const __half*)hi)[idx]) + mpn_e5m2_to_float(((const unsigned char*)lo)[idx]) * MPN_F8_LO_INV;
  float v = __bfloat162float(((const __nv_bfloat16*)hi)[idx]);
  if (fmt == MPN_FMT_BF16X2) v += __bfloat162float(((const __nv_bfloat16*)lo)[idx]);
  return v;
}

// plane = number of elements of one plane of the destination tensor (needed by MPN_FMT_F16F8 to find the h8 plane; < 0: no h8 plane)
__device__ __forceinline__ void mpn_store_act(void* hi, void* lo, long long idx, int fmt, float v, long long plane = 0) {
  if (fmt == MPN_FMT_F32) {
    ((float*)hi)[idx] = v;
  } else if (fmt == MPN_FMT_F16F8) {
    const __half h = mpn_f16_sat(v);
    ((__half*)hi)[idx] = h;
    ((unsigned char*)lo)[idx] = mpn_float_to_e5m2((v - __half2float(h)) * MPN_F8_LO_SCALE);
    if (plane >= 0) ((unsigned char*)lo)[plane + idx] = mpn_float_to_e5m2(v);   // plane < 0: tensor stored without its h8 plane
  } else {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    ((__nv_bfloat16*)hi)[idx] = h;
    if (fmt == MPN_FMT_BF16X2) ((__nv_bfloat16*)lo)[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// PyTorch legacy 'nearest' source index (fpn.py:95 F.upsample(size=...)): floor(dst * in/out) clamped
__device__ __forceinline__ int mpn_nearest_src(int dst, int in_size, int out_size) {
  if (out_size == 2 * in_size) return dst >> 1;
  float scale = (float)in_size / (float)out_size;
  int s = (int)floorf((float)dst * scale);
  return s < in_size - 1 ? s : in_size - 1;
}
