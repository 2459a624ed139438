// Layout conversion, filter packing, BN folding, max-pool, ReLU: the bandwidth-bound glue of the path.
// All kernels are grid-stride, vector-friendly and launched on the caller's stream.
#include <stdarg.h>
#include <string.h>

#include "mpn_common.cuh"

static thread_local char g_err[512] = "";

void mpn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* mpn_last_error(void) { return g_err; }
extern "C" int mpn_version(void) { return 100; }
// sizeof of the two structs that cross the boundary by pointer: a binding checks its own layout against these at load time
extern "C" int mpn_sizeof_conv_desc(void) { return (int)sizeof(mpn_conv_desc); }
extern "C" int mpn_sizeof_conv_ptrs(void) { return (int)sizeof(mpn_conv_ptrs); }

extern "C" int mpn_device_supports_tcgen05(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10;
}

static inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 16;  // 16 resident CTAs of 256 threads per SM, grid-stride beyond
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// ---------------------------------------------------------------------------------------------
__global__ void pack_filter_f32_kernel(const float* __restrict__ w, float* __restrict__ dst, int Cout, int Cin, int R,
                                       int S, int CoutPad) {
  long long total = (long long)R * S * Cin * CoutPad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % CoutPad);
    long long t = i / CoutPad;
    int ci = (int)(t % Cin);
    t /= Cin;
    int s = (int)(t % S);
    int r = (int)(t / S);
    dst[i] = co < Cout ? w[(((long long)co * Cin + ci) * R + r) * S + s] : 0.f;
  }
}

extern "C" int mpn_pack_filter_f32(const float* w, float* dst, int Cout, int Cin, int R, int S, int CoutPad, void* stream) {
  MPN_CHECK_ARG(w && dst && Cout > 0 && Cin > 0 && R > 0 && S > 0 && CoutPad >= Cout, "mpn_pack_filter_f32: bad argument");
  long long total = (long long)R * S * Cin * CoutPad;
  pack_filter_f32_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, dst, Cout, Cin, R, S, CoutPad);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

__global__ void pack_filter_bf16_kernel(const float* __restrict__ w, const float* __restrict__ scale, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, int Cout, int Cin, int R, int S) {
  long long total = (long long)Cout * R * S * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(i % Cin);
    long long t = i / Cin;
    int s = (int)(t % S);
    t /= S;
    int r = (int)(t % R);
    int co = (int)(t / R);
    float v = w[(((long long)co * Cin + ci) * R + r) * S + s];
    if (scale) v = __fmul_rn(v, scale[co]);
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

extern "C" int mpn_pack_filter_bf16(const float* w, void* hi, void* lo, int Cout, int Cin, int R, int S, void* stream) {
  return mpn_pack_filter_bf16_scaled(w, nullptr, hi, lo, Cout, Cin, R, S, stream);
}

extern "C" int mpn_pack_filter_bf16_scaled(const float* w, const float* scale, void* hi, void* lo, int Cout, int Cin, int R, int S,
                                           void* stream) {
  MPN_CHECK_ARG(w && hi && Cout > 0 && Cin > 0 && R > 0 && S > 0, "mpn_pack_filter_bf16: bad argument");
  long long total = (long long)Cout * R * S * Cin;
  pack_filter_bf16_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, scale, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                                 Cout, Cin, R, S);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// MPN_FMT_F16F8 filters (see mpn_b200.h).  amax: max |w * scale[co]| over the filter (non-negative floats order like their bits).
__global__ void filter_absmax_kernel(const float* __restrict__ w, const float* __restrict__ scale, long long total, int per_cout,
                                     float* amax) {
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = w[i];
    if (scale) v = __fmul_rn(v, scale[i / per_cout]);
    m = fmaxf(m, fabsf(v));
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(reinterpret_cast<int*>(amax), __float_as_int(m));
}

extern "C" int mpn_filter_absmax(const float* w, const float* scale, int Cout, int per_cout, float* amax, void* stream) {
  MPN_CHECK_ARG(w && amax && Cout > 0 && per_cout > 0, "mpn_filter_absmax: bad argument");
  const long long total = (long long)Cout * per_cout;
  filter_absmax_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(w, scale, total, per_cout, amax);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

__global__ void pack_filter_f16f8_kernel(const float* __restrict__ w, const float* __restrict__ scale, float wscale, __half* __restrict__ hi,
                                         unsigned char* __restrict__ lo8, unsigned char* __restrict__ h8, __half* __restrict__ lo16,
                                         int Cout, int Cin, int R, int S, int stem, int merged) {
  const long long total = stem ? (long long)Cout * 256 : (long long)Cout * R * S * Cin;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float v = 0.f;
    int co;
    if (stem) {  // [co][r2][k], k = s2*16 + cc (stem_pack_filter_kernel)
      const int k = (int)(i & 63), r2 = (int)((i >> 6) & 3);
      co = (int)(i >> 8);
      const int s2 = k >> 4, cc = k & 15;
      if (cc < 12) {
        const int c = cc % 3, ph = (cc / 3) >> 1, pw = (cc / 3) & 1;
        const int r = 2 * r2 + ph - 1, s = 2 * s2 + pw - 1;
        if (r >= 0 && r < 7 && s >= 0 && s < 7) v = w[((co * 3 + c) * 7 + r) * 7 + s];
      }
    } else {
      const int ci = (int)(i % Cin);
      long long t = i / Cin;
      const int s = (int)(t % S);
      t /= S;
      const int r = (int)(t % R);
      co = (int)(t / R);
      v = w[(((long long)co * Cin + ci) * R + r) * S + s];
    }
    if (scale) v = __fmul_rn(v, scale[co]);
    v = __fmul_rn(v, wscale);  // a power of two: exact
    const __half h = __float2half_rn(v);
    hi[i] = h;
    if (merged) {   // layout bit 2: one 128-byte group per 64 K elements, [64 lo8 | 64 h8] (h8 == lo8 + 64 on entry)
      const long long j = (i >> 6) * 128 + (i & 63);
      lo8[j] = mpn_float_to_e4m3(v - __half2float(h));
      h8[j] = mpn_float_to_e4m3(v * MPN_F8_LO_INV);
      continue;
    }
    if (lo16) lo16[i] = __float2half_rn(v - __half2float(h));   // layout 1: pairs with the activations' fp16 hi plane (MPN_IN_NO_H8)
    else lo8[i] = mpn_float_to_e4m3(v - __half2float(h));       // layout 0: pairs with the activations' e5m2 copy
    h8[i] = mpn_float_to_e4m3(v * MPN_F8_LO_INV);               // pairs with the activations' lo8 = e5m2((x - hi) * 2^12)
  }
}

extern "C" int mpn_pack_filter_f16f8(const float* w, const float* scale, float wscale, void* hi, void* lo8h8, int Cout, int Cin, int R,
                                     int S, int layout, void* stream) {
  // layout bit 0: the stem's space-to-depth filter; bit 1: [lo16 fp16 plane][h8 plane] for inputs without an h8 plane; bit 2: the
  // lo8 and h8 planes interleaved per 64 K elements ([Cout][K/64][64 lo8 | 64 h8]): one 128-byte TMA row instead of two 64-byte
  // ones (MPN_W_MERGED) -- the TMA unit is paced per row, not per byte
  const int stem = layout & 1, lay16 = (layout >> 1) & 1, merged = (layout >> 2) & 1;
  MPN_CHECK_ARG(!(merged && lay16), "mpn_pack_filter_f16f8: layouts 2 and 4 exclude each other");
  MPN_CHECK_ARG(!merged || stem || ((long long)R * S * Cin) % 64 == 0, "mpn_pack_filter_f16f8: the merged layout needs K %% 64 == 0");
  MPN_CHECK_ARG(w && hi && lo8h8 && Cout > 0 && Cin > 0 && R > 0 && S > 0 && wscale > 0.f, "mpn_pack_filter_f16f8: bad argument");
  MPN_CHECK_ARG(!stem || (Cin == 3 && R == 7 && S == 7), "mpn_pack_filter_f16f8: the stem layout is for a [Cout,3,7,7] filter");
  const long long total = stem ? (long long)Cout * 256 : (long long)Cout * R * S * Cin;
  unsigned char* lo = (unsigned char*)lo8h8;
  pack_filter_f16f8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w, scale, wscale, (__half*)hi, lo, merged ? lo + 64 : (lay16 ? lo + 2 * total : lo + total), lay16 ? (__half*)lo : nullptr, Cout, Cin, R,
      S, stem, merged);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

__global__ void fold_bn_kernel(const float* g, const float* b, const float* m, const float* v, float eps, float* scale,
                               float* bias, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C) {
    // (x - mean) / sqrt(var + eps) * gamma + beta  ==  x*scale + bias
    float s = __fdiv_rn(g[i], __fsqrt_rn(__fadd_rn(v[i], eps)));
    scale[i] = s;
    bias[i] = __fsub_rn(b[i], __fmul_rn(m[i], s));
  }
}

extern "C" int mpn_fold_bn(const float* g, const float* b, const float* m, const float* v, float eps, float* scale,
                           float* bias, int C, void* stream) {
  MPN_CHECK_ARG(g && b && m && v && scale && bias && C > 0, "mpn_fold_bn: bad argument");
  fold_bn_kernel<<<mpn_divup(C, 256), 256, 0, (cudaStream_t)stream>>>(g, b, m, v, eps, scale, bias, C);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// NCHW fp32 -> NHWC: one thread per (n, h, w, c) destination element, c fastest (coalesced writes);
// reads are strided by H*W but the tensors this is used on are the 3-channel image and small maps.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, void* hi, void* lo, int N, int C, int H, int W,
                                    int cstride, int fmt) {
  long long total = (long long)N * H * W * cstride;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % cstride);
    long long p = i / cstride;
    int w = (int)(p % W);
    long long q = p / W;
    int h = (int)(q % H);
    int n = (int)(q / H);
    float v = c < C ? src[(((long long)n * C + c) * H + h) * W + w] : 0.f;
    mpn_store_act(hi, lo, i, fmt, v, total);
  }
}

extern "C" int mpn_nchw_to_nhwc(const float* src, void* hi, void* lo, int N, int C, int H, int W, int cstride, int fmt,
                                void* stream) {
  MPN_CHECK_ARG(src && hi && N > 0 && C > 0 && H > 0 && W > 0 && cstride >= C, "mpn_nchw_to_nhwc: bad argument");
  MPN_CHECK_ARG((fmt != MPN_FMT_BF16X2 && fmt != MPN_FMT_F16F8) || lo, "mpn_nchw_to_nhwc: split formats need a lo plane");
  long long total = (long long)N * H * W * cstride;
  nchw_to_nhwc_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(src, hi, lo, N, C, H, W, cstride, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// NHWC -> NCHW fp32 through a 32x32 shared-memory transpose of the (pixel, channel) plane.
__global__ void nhwc_to_nchw_kernel(const void* __restrict__ hi, const void* __restrict__ lo, float* __restrict__ dst,
                                    int C, long long HW, int cstride, int fmt) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    long long p = p0 + j;
    int c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (p < HW && c < C) ? mpn_load_act(hi, lo, ((long long)n * HW + p) * cstride + c, fmt) : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j;
    long long p = p0 + threadIdx.x;
    if (p < HW && c < C) dst[((long long)n * C + c) * HW + p] = tile[threadIdx.x][j];
  }
}

extern "C" int mpn_nhwc_to_nchw(const void* hi, const void* lo, float* dst, int N, int C, int H, int W, int cstride, int fmt,
                                void* stream) {
  MPN_CHECK_ARG(hi && dst && N > 0 && C > 0 && H > 0 && W > 0 && cstride >= C, "mpn_nhwc_to_nchw: bad argument");
  long long HW = (long long)H * W;
  dim3 grid(mpn_divup(HW, 32), mpn_divup(C, 32), N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(hi, lo, dst, C, HW, cstride, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// max_pool2d(3, stride 2, pad 1), NHWC.  One thread per (pixel, 4-channel group); -inf padding.
__global__ void maxpool3x3s2_kernel(const void* __restrict__ xhi, const void* __restrict__ xlo, void* yhi, void* ylo, int N,
                                    int H, int W, int C, int OH, int OW, int fmt, int no_h8) {
  long long total = (long long)N * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int ow = (int)(p % OW);
    long long q = p / OW;
    int oh = (int)(q % OH);
    int n = (int)(q / OH);
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        m = fmaxf(m, mpn_load_act(xhi, xlo, (((long long)n * H + ih) * W + iw) * C + c, fmt));
      }
    }
    mpn_store_act(yhi, ylo, i, fmt, m, no_h8 ? -1 : total);
  }
}

// bf16 planes, 8 channels (one 16-byte vector) per thread; hi/lo pairs are compared as hi+lo values.
template <bool SPLIT>
__global__ void maxpool3x3s2_bf16_kernel(const uint4* __restrict__ xhi, const uint4* __restrict__ xlo, uint4* __restrict__ yhi,
                                         uint4* __restrict__ ylo, int N, int H, int W, int C8, int OH, int OW) {
  long long total = (long long)N * OH * OW * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long p = i / C8;
    int ow = (int)(p % OW);
    long long q = p / OW;
    int oh = (int)(q % OH);
    int n = (int)(q / OH);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        long long o = (((long long)n * H + ih) * W + iw) * C8 + c;
        uint4 h = __ldg(xhi + o);
        const unsigned hv[4] = {h.x, h.y, h.z, h.w};
        float v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[2 * j] = __uint_as_float(hv[j] << 16);
          v[2 * j + 1] = __uint_as_float(hv[j] & 0xFFFF0000u);
        }
        if (SPLIT) {
          uint4 l = __ldg(xlo + o);
          const unsigned lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            v[2 * j] += __uint_as_float(lv[j] << 16);
            v[2 * j + 1] += __uint_as_float(lv[j] & 0xFFFF0000u);
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
      }
    }
    unsigned oh4[4], ol4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 t = __floats2bfloat162_rn(m[2 * j], m[2 * j + 1]);
      oh4[j] = *reinterpret_cast<unsigned*>(&t);
      if (SPLIT) {
        __nv_bfloat162 u = __floats2bfloat162_rn(m[2 * j] - __uint_as_float(oh4[j] << 16),
                                                 m[2 * j + 1] - __uint_as_float(oh4[j] & 0xFFFF0000u));
        ol4[j] = *reinterpret_cast<unsigned*>(&u);
      }
    }
    yhi[i] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
    if (SPLIT) ylo[i] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
  }
}

// MPN_FMT_F16F8 planes, 8 channels per thread (16 B of fp16 hi + 8 B of e5m2 lo8): same arithmetic as the generic kernel
// (value = hi + 2^-12 * lo8, maximum, re-split into hi / lo8 / h8), vector loads and stores.
__global__ void maxpool3x3s2_f16f8_kernel(const uint4* __restrict__ xhi, const uint2* __restrict__ xlo, uint4* __restrict__ yhi,
                                          uint2* __restrict__ ylo, uint2* __restrict__ yh8, int N, int H, int W, int C8, int OH, int OW) {
  long long total = (long long)N * OH * OW * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C8);
    long long p = i / C8;
    int ow = (int)(p % OW);
    long long q = p / OW;
    int oh = (int)(q % OH);
    int n = (int)(q / OH);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      int ih = oh * 2 - 1 + r;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        int iw = ow * 2 - 1 + s;
        if (iw < 0 || iw >= W) continue;
        long long o = (((long long)n * H + ih) * W + iw) * C8 + c;
        const uint4 h = __ldg(xhi + o);
        const uint2 l = __ldg(xlo + o);
        const unsigned hv[4] = {h.x, h.y, h.z, h.w};
        const unsigned lv[2] = {l.x, l.y};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hv[j]));
          const unsigned b0 = (lv[j >> 1] >> (16 * (j & 1))) & 0xFFu, b1 = (lv[j >> 1] >> (16 * (j & 1) + 8)) & 0xFFu;
          m[2 * j] = fmaxf(m[2 * j], f.x + mpn_e5m2_to_float((unsigned char)b0) * MPN_F8_LO_INV);
          m[2 * j + 1] = fmaxf(m[2 * j + 1], f.y + mpn_e5m2_to_float((unsigned char)b1) * MPN_F8_LO_INV);
        }
      }
    }
    unsigned oh4[4], ol2[2] = {0u, 0u}, og2[2] = {0u, 0u};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      oh4[j] = mpn_pack_f16x2_sat(m[2 * j], m[2 * j + 1]);
      const __half2 t = *reinterpret_cast<const __half2*>(&oh4[j]);
      const float2 f = __half22float2(t);
      const unsigned l0 = mpn_float_to_e5m2((m[2 * j] - f.x) * MPN_F8_LO_SCALE), l1 = mpn_float_to_e5m2((m[2 * j + 1] - f.y) * MPN_F8_LO_SCALE);
      const unsigned g0 = mpn_float_to_e5m2(m[2 * j]), g1 = mpn_float_to_e5m2(m[2 * j + 1]);
      ol2[j >> 1] |= (l0 | (l1 << 8)) << (16 * (j & 1));
      og2[j >> 1] |= (g0 | (g1 << 8)) << (16 * (j & 1));
    }
    yhi[i] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
    ylo[i] = make_uint2(ol2[0], ol2[1]);
    if (yh8) yh8[i] = make_uint2(og2[0], og2[1]);
  }
}

extern "C" int mpn_maxpool3x3s2(const void* xhi, const void* xlo, void* yhi, void* ylo, int N, int H, int W, int C, int fmt,
                                int flags, void* stream) {
  const int no_h8 = (fmt == MPN_FMT_F16F8 && (flags & MPN_EPI_NO_H8)) ? 1 : 0;
  MPN_CHECK_ARG(xhi && yhi && N > 0 && H > 0 && W > 0 && C > 0, "mpn_maxpool3x3s2: bad argument");
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * OH * OW * C;
  if ((fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && C % 8 == 0) {
    long long t8 = total / 8;
    if (fmt == MPN_FMT_BF16X2)
      maxpool3x3s2_bf16_kernel<true><<<grid_for(t8, 256), 256, 0, (cudaStream_t)stream>>>(
          (const uint4*)xhi, (const uint4*)xlo, (uint4*)yhi, (uint4*)ylo, N, H, W, C / 8, OH, OW);
    else
      maxpool3x3s2_bf16_kernel<false><<<grid_for(t8, 256), 256, 0, (cudaStream_t)stream>>>(
          (const uint4*)xhi, nullptr, (uint4*)yhi, nullptr, N, H, W, C / 8, OH, OW);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  if (fmt == MPN_FMT_F16F8 && C % 8 == 0) {
    maxpool3x3s2_f16f8_kernel<<<grid_for(total / 8, 256), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)xhi, (const uint2*)xlo, (uint4*)yhi, (uint2*)ylo, no_h8 ? nullptr : (uint2*)((unsigned char*)ylo + total), N, H, W,
        C / 8, OH, OW);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  maxpool3x3s2_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(xhi, xlo, yhi, ylo, N, H, W, C, OH, OW, fmt, no_h8);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

__global__ void relu_kernel(const void* xhi, const void* xlo, void* yhi, void* ylo, long long n, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mpn_store_act(yhi, ylo, i, fmt, fmaxf(mpn_load_act(xhi, xlo, i, fmt), 0.f), n);
}

extern "C" int mpn_relu(const void* xhi, const void* xlo, void* yhi, void* ylo, long long n, int fmt, void* stream) {
  MPN_CHECK_ARG(xhi && yhi && n > 0, "mpn_relu: bad argument");
  relu_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(xhi, xlo, yhi, ylo, n, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor-core stem helpers (see mpn_b200.h).  X2[n][hp][wp][cc], cc = (ph*2+pw)*3 + c, hp = h/2 + 2.
// One thread per (image, hp, wp) position of the space-to-depth tensor: 16 channels = the 2 x 2 pixel block (ph, pw) x 3 colours
// + 4 zero pads, written with vector stores (32 B hi + the format's second / third plane) -- r02: the element-per-thread version
// ran at 0.8 TB/s and its uint8 twin was slower than the fp32 one although it reads 4x fewer bytes.
__device__ __forceinline__ void stem_store16(void* hi, void* lo, long long pos, int fmt, const float* v, long long plane) {
  if (fmt == MPN_FMT_F16F8) {
    unsigned h[8], l[4], g[4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = mpn_pack_f16x2_sat(v[2 * j], v[2 * j + 1]);
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
      const unsigned l0 = mpn_float_to_e5m2((v[2 * j] - f.x) * MPN_F8_LO_SCALE), l1 = mpn_float_to_e5m2((v[2 * j + 1] - f.y) * MPN_F8_LO_SCALE);
      const unsigned g0 = mpn_float_to_e5m2(v[2 * j]), g1 = mpn_float_to_e5m2(v[2 * j + 1]);
      if ((j & 1) == 0) { l[j >> 1] = l0 | (l1 << 8); g[j >> 1] = g0 | (g1 << 8); }
      else { l[j >> 1] |= (l0 | (l1 << 8)) << 16; g[j >> 1] |= (g0 | (g1 << 8)) << 16; }
    }
    uint4* ph = reinterpret_cast<uint4*>((__half*)hi + pos * 16);
    ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
    ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>((unsigned char*)lo + pos * 16) = make_uint4(l[0], l[1], l[2], l[3]);
    if (plane >= 0) *reinterpret_cast<uint4*>((unsigned char*)lo + plane + pos * 16) = make_uint4(g[0], g[1], g[2], g[3]);   // < 0: no h8 plane
  } else {
    unsigned h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      h[j] = *reinterpret_cast<const unsigned*>(&t);
      const __nv_bfloat162 r = __floats2bfloat162_rn(v[2 * j] - __uint_as_float(h[j] << 16), v[2 * j + 1] - __uint_as_float(h[j] & 0xFFFF0000u));
      l[j] = *reinterpret_cast<const unsigned*>(&r);
    }
    uint4* ph = reinterpret_cast<uint4*>((__nv_bfloat16*)hi + pos * 16);
    ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
    ph[1] = make_uint4(h[4], h[5], h[6], h[7]);
    if (fmt == MPN_FMT_BF16X2) {
      uint4* pl = reinterpret_cast<uint4*>((__nv_bfloat16*)lo + pos * 16);
      pl[0] = make_uint4(l[0], l[1], l[2], l[3]);
      pl[1] = make_uint4(l[4], l[5], l[6], l[7]);
    }
  }
}

__global__ void stem_pack_input_kernel(const float* __restrict__ img, void* hi, void* lo, int N, int H, int W, int H2p, int W2p,
                                       int fmt, int no_h8) {
  const long long npos = (long long)N * H2p * W2p;
  for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < npos; pos += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(pos % W2p);
    const long long q = pos / W2p;
    const int hp = (int)(q % H2p), n = (int)(q / H2p);
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const int ih = 2 * (hp - 2) + ph;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const int iw = 2 * (wp - 2) + pw;
        if (iw < 0 || iw >= W) continue;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ph * 2 + pw) * 3 + c] = __ldg(img + (((long long)n * 3 + c) * H + ih) * W + iw);
      }
    }
    stem_store16(hi, lo, pos, fmt, v, no_h8 ? -1 : npos * 16);
  }
}

extern "C" int mpn_stem_pack_input(const float* img, void* hi, void* lo, int N, int H, int W, int fmt, int flags, void* stream) {
  const int no_h8 = (fmt == MPN_FMT_F16F8 && (flags & MPN_EPI_NO_H8)) ? 1 : 0;
  MPN_CHECK_ARG(img && hi && N > 0 && H > 0 && W > 0, "mpn_stem_pack_input: bad argument");
  MPN_CHECK_ARG(fmt == MPN_FMT_BF16 || ((fmt == MPN_FMT_BF16X2 || fmt == MPN_FMT_F16F8) && lo), "mpn_stem_pack_input: fmt must be BF16, BF16X2 or F16F8 (with lo)");
  int H2p = (H + 1) / 2 + 3, W2p = (W + 1) / 2 + 3;
  long long total = (long long)N * H2p * W2p;
  stem_pack_input_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img, hi, lo, N, H, W, H2p, W2p, fmt, no_h8);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

__global__ void stem_pack_filter_kernel(const float* __restrict__ w, __nv_bfloat16* hi, __nv_bfloat16* lo, int Cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // [co][r2][k], k = s2*16 + cc
  if (i >= Cout * 4 * 64) return;
  int k = i & 63, r2 = (i >> 6) & 3, co = i >> 8;
  int s2 = k >> 4, cc = k & 15;
  float v = 0.f;
  if (cc < 12) {
    int c = cc % 3, ph = (cc / 3) >> 1, pw = (cc / 3) & 1;
    int r = 2 * r2 + ph - 1, s = 2 * s2 + pw - 1;  // tap of the 7x7 filter this (row, pixel, parity) slot holds
    if (r >= 0 && r < 7 && s >= 0 && s < 7) v = w[((co * 3 + c) * 7 + r) * 7 + s];
  }
  __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

extern "C" int mpn_stem_pack_filter(const float* w, void* hi, void* lo, int Cout, void* stream) {
  MPN_CHECK_ARG(w && hi && Cout > 0, "mpn_stem_pack_filter: bad argument");
  stem_pack_filter_kernel<<<mpn_divup(Cout * 256, 256), 256, 0, (cudaStream_t)stream>>>(w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Cout);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// PRN tail: out = softmax(a + res) per row (posenet.py:147-148).  One 512-thread CTA per row; the row
// (34272 floats for the reference's 56x36x17 grid) is streamed three times from L2: max, sum, normalise.
__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float r = red[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

__global__ void __launch_bounds__(512) add_softmax_rows_kernel(const void* __restrict__ ahi, const void* __restrict__ alo,
                                                               const float* __restrict__ res, float* __restrict__ out, int D,
                                                               int a_stride, int fmt) {
  __shared__ float red[16];
  const long long r = blockIdx.x;
  const float* rr = res + r * D;
  float* oo = out + r * D;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float v = mpn_load_act(ahi, alo, r * a_stride + j, fmt) + rr[j];
    oo[j] = v;  // staged in the output row
    m = fmaxf(m, v);
  }
  m = block_reduce(m, red, true);
  float s = 0.f;
  for (int j = threadIdx.x; j < D; j += blockDim.x) {
    float e = expf(oo[j] - m);
    oo[j] = e;
    s += e;
  }
  s = block_reduce(s, red, false);
  const float inv = 1.f / s;
  for (int j = threadIdx.x; j < D; j += blockDim.x) oo[j] = oo[j] * inv;
}

extern "C" int mpn_add_softmax_rows(const void* ahi, const void* alo, const float* res, float* out, int P, int D, int a_stride,
                                    int fmt, void* stream) {
  MPN_CHECK_ARG(ahi && res && out && P > 0 && D > 0 && a_stride >= D, "mpn_add_softmax_rows: bad argument");
  MPN_CHECK_ARG((fmt != MPN_FMT_BF16X2 && fmt != MPN_FMT_F16F8) || alo, "mpn_add_softmax_rows: split formats need a lo plane");
  add_softmax_rows_kernel<<<P, 512, 0, (cudaStream_t)stream>>>(ahi, alo, res, out, D, a_stride, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// resnet_preprocess (datasets/coco_data/preprocessing.py:15-26) on the device, for uint8 HWC BGR images as cv2.imread
// returns them: out[c] = ((float(u8[2-c]) / 255) - mean[c]) / std[c] in fp32, operation for operation (no FMA, IEEE
// division), so the result is bit-identical to the numpy code.
__device__ __forceinline__ float resnet_preprocess_px(const unsigned char* __restrict__ px, int c) {
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const float v = __fdiv_rn((float)px[2 - c], 255.f);  // BGR -> RGB
  return __fdiv_rn(__fsub_rn(v, mean[c]), stdv[c]);
}

__global__ void preprocess_u8_nchw_kernel(const unsigned char* __restrict__ img, float* __restrict__ out, int N, int H, int W) {
  long long total = (long long)N * 3 * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    long long t = i / W;
    int h = (int)(t % H);
    t /= H;
    int c = (int)(t % 3);
    int n = (int)(t / 3);
    out[i] = resnet_preprocess_px(img + (((long long)n * H + h) * W + w) * 3, c);
  }
}

extern "C" int mpn_preprocess_u8_nchw(const unsigned char* img_nhwc_bgr, float* out_nchw, int N, int H, int W, void* stream) {
  MPN_CHECK_ARG(img_nhwc_bgr && out_nchw && N > 0 && H > 0 && W > 0, "mpn_preprocess_u8_nchw: bad argument");
  long long total = (long long)N * 3 * H * W;
  preprocess_u8_nchw_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img_nhwc_bgr, out_nchw, N, H, W);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// the same, fused into the tensor-core stem's space-to-depth packing (see mpn_stem_pack_input)
__global__ void stem_pack_input_u8_kernel(const unsigned char* __restrict__ img, void* hi, void* lo, int N, int H, int W, int H2p,
                                          int W2p, int fmt, int no_h8) {
  // a byte has 256 values: tabulate the exact (IEEE-division) result per channel once per CTA instead of dividing per pixel
  __shared__ float lut[3][256];
  for (int t = threadIdx.x; t < 768; t += blockDim.x) {
    const unsigned char fake[3] = {(unsigned char)(t & 255), (unsigned char)(t & 255), (unsigned char)(t & 255)};
    lut[t >> 8][t & 255] = resnet_preprocess_px(fake, t >> 8);
  }
  __syncthreads();
  const long long npos = (long long)N * H2p * W2p;
  for (long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x; pos < npos; pos += (long long)gridDim.x * blockDim.x) {
    const int wp = (int)(pos % W2p);
    const long long q = pos / W2p;
    const int hp = (int)(q % H2p), n = (int)(q / H2p);
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0.f;
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {
      const int ih = 2 * (hp - 2) + ph;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const int iw = 2 * (wp - 2) + pw;
        if (iw < 0 || iw >= W) continue;
        const unsigned char* px = img + (((long long)n * H + ih) * W + iw) * 3;   // BGR
#pragma unroll
        for (int c = 0; c < 3; ++c) v[(ph * 2 + pw) * 3 + c] = lut[c][px[2 - c]];
      }
    }
    stem_store16(hi, lo, pos, fmt, v, no_h8 ? -1 : npos * 16);
  }
}

extern "C" int mpn_stem_pack_input_u8(const unsigned char* img_nhwc_bgr, void* hi, void* lo, int N, int H, int W, int fmt,
                                      int flags, void* stream) {
  const int no_h8 = (fmt == MPN_FMT_F16F8 && (flags & MPN_EPI_NO_H8)) ? 1 : 0;
  MPN_CHECK_ARG(img_nhwc_bgr && hi && N > 0 && H > 0 && W > 0, "mpn_stem_pack_input_u8: bad argument");
  MPN_CHECK_ARG(fmt == MPN_FMT_BF16 || ((fmt == MPN_FMT_BF16X2 || fmt == MPN_FMT_F16F8) && lo), "mpn_stem_pack_input_u8: fmt must be BF16, BF16X2 or F16F8 (with lo)");
  int H2p = (H + 1) / 2 + 3, W2p = (W + 1) / 2 + 3;
  long long total = (long long)N * H2p * W2p;
  stem_pack_input_u8_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(img_nhwc_bgr, hi, lo, N, H, W, H2p, W2p, fmt, no_h8);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
