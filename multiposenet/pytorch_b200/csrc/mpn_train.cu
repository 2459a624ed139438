// Training-step kernels of the keypoint subnet: BatchNorm2d in train mode (batch statistics,
// running-stat update, backward), ReLU / max-pool / nearest-upsample / replication backward, zero-insertion for
// stride-2 data gradients, per-channel reductions (bias gradients), the weighted-MSE heat-map loss of
// network/posenet.py:367-403 with its gradient, and filter re-layouts for dgrad / wgrad.
// All tensors are NHWC activations in one of the MPN_FMT_* formats with dense channels (cstride == C) unless a
// stride is passed explicitly.  These are bandwidth-bound passes; the convolutions themselves (forward, data
// gradient = the same tcgen05 kernel on a flipped/transposed filter, weight gradient = mpn_wgrad_tc.cu) carry the FLOPs.
#include "mpn_common.cuh"

namespace {

inline int grid_for(long long n, int block) {
  long long g = (n + block - 1) / block;
  const long long cap = 148LL * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

constexpr int PIX_PER_BLOCK = 64;

// sum[c] += sum_p x[p,c] ; sumsq[c] += sum_p x[p,c]^2   (fp32 partials per 64 pixels, fp64 atomics across blocks)
__global__ void channel_stats_kernel(const void* __restrict__ hi, const void* __restrict__ lo, long long pixels, int C,
                                     int cstride, int coffset, int fmt, double* __restrict__ sum, double* __restrict__ sumsq) {
  const long long p0 = (long long)blockIdx.x * PIX_PER_BLOCK;
  const long long p1 = p0 + PIX_PER_BLOCK < pixels ? p0 + PIX_PER_BLOCK : pixels;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (long long p = p0; p < p1; ++p) {
      float v = mpn_load_act(hi, lo, p * cstride + coffset + c, fmt);
      s += v;
      q = fmaf(v, v, q);
    }
    atomicAdd(sum + c, (double)s);
    if (sumsq) atomicAdd(sumsq + c, (double)q);
  }
}

__global__ void bn_finalize_kernel(const double* sum, const double* sumsq, double n, float* mean, float* var, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    double m = sum[c] / n;
    double v = sumsq[c] / n - m * m;
    mean[c] = (float)m;
    var[c] = (float)(v > 0.0 ? v : 0.0);  // biased variance, as used for normalisation
  }
}

__global__ void bn_update_running_kernel(const float* mean, const float* var, float* rm, float* rv, double n, float momentum, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    rm[c] = (1.f - momentum) * rm[c] + momentum * mean[c];
    double unbiased = n > 1.0 ? (double)var[c] * n / (n - 1.0) : (double)var[c];
    rv[c] = (1.f - momentum) * rv[c] + momentum * (float)unbiased;
  }
}

__global__ void bn_apply_kernel(const void* yhi, const void* ylo, const float* __restrict__ mean, const float* __restrict__ var,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float eps, const void* rhi,
                                const void* rlo, int relu, void* zhi, void* zlo, long long total, int C, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float invstd = rsqrtf(var[c] + eps);
    float v = (mpn_load_act(yhi, ylo, i, fmt) - mean[c]) * invstd * gamma[c] + beta[c];
    if (rhi) v += mpn_load_act(rhi, rlo, i, fmt);
    if (relu) v = fmaxf(v, 0.f);
    mpn_store_act(zhi, zlo, i, fmt, v);
  }
}

// s1[c] = sum g, s2[c] = sum g * yhat, with g = dz * (z > 0 if relu), yhat = (y - mean) * invstd
__global__ void bn_bwd_reduce_kernel(const void* dzhi, const void* dzlo, const void* zhi, const void* zlo, const void* yhi,
                                     const void* ylo, const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                     int relu, long long pixels, int C, int fmt, double* __restrict__ s1, double* __restrict__ s2) {
  const long long p0 = (long long)blockIdx.x * PIX_PER_BLOCK;
  const long long p1 = p0 + PIX_PER_BLOCK < pixels ? p0 + PIX_PER_BLOCK : pixels;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m = mean[c], invstd = rsqrtf(var[c] + eps);
    float a = 0.f, b = 0.f;
    for (long long p = p0; p < p1; ++p) {
      const long long i = p * C + c;
      float g = mpn_load_act(dzhi, dzlo, i, fmt);
      if (relu && !(mpn_load_act(zhi, zlo, i, fmt) > 0.f)) g = 0.f;
      a += g;
      b = fmaf(g, (mpn_load_act(yhi, ylo, i, fmt) - m) * invstd, b);
    }
    atomicAdd(s1 + c, (double)a);
    atomicAdd(s2 + c, (double)b);
  }
}

// dy = gamma * invstd * (g - s1/n - yhat * s2/n);  gmask (optional) = g (gradient handed to the shortcut branch)
__global__ void bn_bwd_apply_kernel(const void* dzhi, const void* dzlo, const void* zhi, const void* zlo, const void* yhi,
                                    const void* ylo, const float* __restrict__ mean, const float* __restrict__ var,
                                    const float* __restrict__ gamma, float eps, int relu, const double* __restrict__ s1,
                                    const double* __restrict__ s2, double n, void* dyhi, void* dylo, void* ghi, void* glo,
                                    long long total, int C, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    const float invstd = rsqrtf(var[c] + eps);
    float g = mpn_load_act(dzhi, dzlo, i, fmt);
    if (relu && !(mpn_load_act(zhi, zlo, i, fmt) > 0.f)) g = 0.f;
    const float yhat = (mpn_load_act(yhi, ylo, i, fmt) - mean[c]) * invstd;
    const float m1 = (float)(s1[c] / n), m2 = (float)(s2[c] / n);
    mpn_store_act(dyhi, dylo, i, fmt, gamma[c] * invstd * (g - m1 - yhat * m2));
    if (ghi) mpn_store_act(ghi, glo, i, fmt, g);
  }
}

__global__ void double_to_float_kernel(const double* a, float* o, int n, float scale) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) o[i] = (float)(a[i] * scale);
}

__global__ void relu_bwd_kernel(const void* dzhi, const void* dzlo, const void* zhi, const void* zlo, void* ohi, void* olo,
                                long long n, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float g = mpn_load_act(dzhi, dzlo, i, fmt);
    if (!(mpn_load_act(zhi, zlo, i, fmt) > 0.f)) g = 0.f;
    mpn_store_act(ohi, olo, i, fmt, g);
  }
}

__global__ void add_act_kernel(const void* ahi, const void* alo, const void* bhi, const void* blo, void* ohi, void* olo,
                               long long n, int fmt) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mpn_store_act(ohi, olo, i, fmt, mpn_load_act(ahi, alo, i, fmt) + mpn_load_act(bhi, blo, i, fmt));
}

// max_pool2d(3, 2, 1) backward: every input pixel gathers dy from the (<= 4) windows whose arg-max it is
// (first maximum in row-major window order, like ATen's max_pool2d_with_indices).
__global__ void maxpool_bwd_kernel(const void* xhi, const void* xlo, const void* dyhi, const void* dylo, void* dxhi, void* dxlo,
                                   int N, int H, int W, int C, int OH, int OW, int fmt) {
  long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int iw = (int)(p % W);
    long long q = p / W;
    int ih = (int)(q % H);
    int n = (int)(q / H);
    float acc = 0.f;
    for (int oh = (ih + 1 - 2 + 1) / 2 > 0 ? (ih + 1 - 2 + 1) / 2 : 0; oh <= (ih + 1) / 2 && oh < OH; ++oh)
      for (int ow = (iw + 1 - 2 + 1) / 2 > 0 ? (iw + 1 - 2 + 1) / 2 : 0; ow <= (iw + 1) / 2 && ow < OW; ++ow) {
        // window rows 2*oh-1 .. 2*oh+1
        float best = -INFINITY;
        int bh = -1, bw = -1;
        for (int r = 0; r < 3; ++r) {
          int y = 2 * oh - 1 + r;
          if (y < 0 || y >= H) continue;
          for (int s = 0; s < 3; ++s) {
            int x = 2 * ow - 1 + s;
            if (x < 0 || x >= W) continue;
            float v = mpn_load_act(xhi, xlo, (((long long)n * H + y) * W + x) * C + c, fmt);
            if (v > best || bh < 0) { best = v; bh = y; bw = x; }
          }
        }
        if (bh == ih && bw == iw) acc += mpn_load_act(dyhi, dylo, (((long long)n * OH + oh) * OW + ow) * C + c, fmt);
      }
    mpn_store_act(dxhi, dxlo, i, fmt, acc);
  }
}

// coarse[n,h,w,c] = sum over the r x r block of fine[n, h*r+dy, w*r+dx, coffset + c]  (backward of nearest upsample
// by an integer factor: fpn.py:84-95 with exact 2x, posenet.py:180-182 with 8/4/2)
__global__ void block_sum_kernel(const void* fhi, const void* flo, int f_cstride, int f_coffset, void* chi, void* clo, int N,
                                 int h, int w, int C, int r, int fmt) {
  long long total = (long long)N * h * w * C;
  const int H = h * r, W = w * r;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int x = (int)(p % w);
    long long q = p / w;
    int y = (int)(q % h);
    int n = (int)(q / h);
    float acc = 0.f;
    for (int dy = 0; dy < r; ++dy)
      for (int dx = 0; dx < r; ++dx)
        acc += mpn_load_act(fhi, flo, (((long long)n * H + y * r + dy) * W + x * r + dx) * f_cstride + f_coffset + c, fmt);
    mpn_store_act(chi, clo, i, fmt, acc);
  }
}

// out[n, 2*oh, 2*ow, :] = dy[n, oh, ow, :], zero elsewhere; out is [N,H,W,C]
__global__ void zero_insert2_kernel(const void* dyhi, const void* dylo, void* ohi, void* olo, int N, int OH, int OW, int C, int H,
                                    int W, int fmt) {
  long long total = (long long)N * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long p = i / C;
    int x = (int)(p % W);
    long long q = p / W;
    int y = (int)(q % H);
    int n = (int)(q / H);
    float v = 0.f;
    if (!(y & 1) && !(x & 1) && (y >> 1) < OH && (x >> 1) < OW)
      v = mpn_load_act(dyhi, dylo, (((long long)n * OH + (y >> 1)) * OW + (x >> 1)) * C + c, fmt);
    mpn_store_act(ohi, olo, i, fmt, v);
  }
}

// Weighted MSE of posenet.py:380-387 for one of the 5 supervised maps:
//   loss += mean_{b,c<18,h,w} (pred*w - gt*w)^2 ;  dpred = grad_scale * 2*w*(pred*w - gt*w)/numel  (channels >= 18: 0)
// pred is fp32 NCHW [B,Cp,H,W] (Cp = 18 or 19), gt/w fp32 NCHW [B,18,H,W]; dpred is an NHWC activation with Cd channels.
__global__ void mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ wt, int B,
                                int Cp, int H, int W, double* loss, void* dhi, void* dlo, int Cd, int fmt, float grad_scale) {
  const long long HW = (long long)H * W;
  const long long total = (long long)B * HW * Cd;
  const double numel = (double)B * 18.0 * (double)HW;
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Cd);
    long long p = i / Cd;  // b*HW + hw
    long long hw = p % HW;
    int b = (int)(p / HW);
    float g = 0.f;
    if (c < 18) {
      float w = wt[((long long)b * 18 + c) * HW + hw];
      float d = pred[((long long)b * Cp + c) * HW + hw] * w - gt[((long long)b * 18 + c) * HW + hw] * w;
      local = fmaf(d, d, local);
      g = grad_scale * 2.f * w * d / (float)numel;
    }
    mpn_store_act(dhi, dlo, i, fmt, g);
  }
  // block reduction of the loss
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    atomicAdd(loss, (double)s / numel);
  }
}

// OIHW fp32 -> data-gradient filter [Cin][R][S][CoutPad] bf16 hi(+lo): W'[ci][r'][s'][co] = W[co][ci][R-1-r'][S-1-s']
__global__ void pack_filter_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* hi, __nv_bfloat16* lo, int Cout, int Cin, int R,
                                         int S, int CoutPad) {
  long long total = (long long)Cin * R * S * CoutPad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % CoutPad);
    long long t = i / CoutPad;
    int s = (int)(t % S);
    t /= S;
    int r = (int)(t % R);
    int ci = (int)(t / R);
    float v = co < Cout ? w[(((long long)co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s)] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// [Cout][R][S][Cin] fp32 (wgrad layout) -> OIHW fp32 (torch parameter layout)
__global__ void unpack_filter_grad_kernel(const float* __restrict__ g, float* __restrict__ out, int Cout, int Cin, int R, int S) {
  long long total = (long long)Cout * Cin * R * S;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int s = (int)(i % S);
    long long t = i / S;
    int r = (int)(t % R);
    t /= R;
    int ci = (int)(t % Cin);
    int co = (int)(t / Cin);
    out[i] = g[(((long long)co * R + r) * S + s) * Cin + ci];
  }
}

// stem: gradient of the [64][4][64] space-to-depth filter (see mpn_stem_pack_filter) -> OIHW [64,3,7,7]
__global__ void stem_unpack_filter_grad_kernel(const float* __restrict__ g, float* __restrict__ out, int Cout) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;  // OIHW index
  if (i >= Cout * 147) return;
  int s = i % 7, r = (i / 7) % 7, c = (i / 49) % 3, co = i / 147;
  int r2 = (r + 1) >> 1, ph = (r + 1) & 1, s2 = (s + 1) >> 1, pw = (s + 1) & 1;  // r = 2*r2 + ph - 1
  int cc = (ph * 2 + pw) * 3 + c;
  out[i] = g[(co * 4 + r2) * 64 + s2 * 16 + cc];
}


// ---------------------------------------------------------------------------------------------
// Vectorised bf16 / bf16x2 variants (8 channels = one 16-byte vector per thread).  The scalar kernels above stay as
// the generic (fp32 / odd channel count) path.
template <bool SPLIT>
__device__ __forceinline__ void load8(const uint4* __restrict__ hi, const uint4* __restrict__ lo, long long i, float* v) {
  const uint4 h = __ldg(hi + i);
  const unsigned hv[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __uint_as_float(hv[j] << 16);
    v[2 * j + 1] = __uint_as_float(hv[j] & 0xFFFF0000u);
  }
  if (SPLIT) {
    const uint4 l = __ldg(lo + i);
    const unsigned lv[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] += __uint_as_float(lv[j] << 16);
      v[2 * j + 1] += __uint_as_float(lv[j] & 0xFFFF0000u);
    }
  }
}

template <bool SPLIT>
__device__ __forceinline__ void store8(uint4* hi, uint4* lo, long long i, const float* v) {
  unsigned oh[4], ol[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    oh[j] = *reinterpret_cast<unsigned*>(&t);
    if (SPLIT) {
      __nv_bfloat162 u = __floats2bfloat162_rn(v[2 * j] - __uint_as_float(oh[j] << 16), v[2 * j + 1] - __uint_as_float(oh[j] & 0xFFFF0000u));
      ol[j] = *reinterpret_cast<unsigned*>(&u);
    }
  }
  hi[i] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
  if (SPLIT) lo[i] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
}


// per-channel sum / sum of squares; dense channels (C8 = C/8 vectors per pixel), 256 threads.
// grid = (pixel chunks, channel slices of <= 32 vectors): small-spatial / wide layers still fill the machine.
template <bool SPLIT>
__global__ void __launch_bounds__(256) channel_stats_v8_kernel(const uint4* __restrict__ hi, const uint4* __restrict__ lo,
                                                               long long pixels, int C8, int vpix, double* __restrict__ sum,
                                                               double* __restrict__ sumsq) {
  // fp64 accumulators: the cross-thread sums must not depend on the (non-deterministic) order of the atomics, otherwise
  // batch statistics differ in the last fp32 bit from run to run and the bf16 hi/lo re-split turns that into ~1e-5 jitter
  // per-thread fp32 partials -> shared memory [lane][channel] -> one fp64 column sum per channel -> ONE fp64 global atomic per
  // channel and CTA (r02: the 4096 contended fp64 shared-memory atomics per CTA and the un-unrolled load loop held these
  // reductions at 1.2-2 TB/s)
  __shared__ float part[2][2048];
  const int groups = C8 < 32 ? C8 : 32;          // channel vectors of this CTA's slice (power of two)
  const int cg = blockIdx.y * 32 + (threadIdx.x % groups);
  const int lanes = 256 / groups, lane = threadIdx.x / groups;
  const long long p0 = (long long)blockIdx.x * vpix;
  const long long p1 = p0 + vpix < pixels ? p0 + vpix : pixels;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
#pragma unroll 4
  for (long long p = p0 + lane; p < p1; p += lanes) {
    float v[8];
    load8<SPLIT>(hi, lo, p * C8 + cg, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
  }
  {
    const int g = threadIdx.x % groups;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part[0][lane * (groups * 8) + g * 8 + j] = s[j];
      part[1][lane * (groups * 8) + g * 8 + j] = q[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < groups * 8) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l) {
      a += (double)part[0][l * (groups * 8) + threadIdx.x];
      b += (double)part[1][l * (groups * 8) + threadIdx.x];
    }
    const int c = blockIdx.y * 256 + threadIdx.x;
    atomicAdd(sum + c, a);
    if (sumsq) atomicAdd(sumsq + c, b);
  }
}

// coef[0..C) = gamma*invstd, coef[C..2C) = beta - mean*gamma*invstd
__global__ void bn_fwd_coef_kernel(const float* mean, const float* var, const float* gamma, const float* beta, float eps, float* coef, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float a = gamma[c] * rsqrtf(var[c] + eps);
    coef[c] = a;
    coef[C + c] = beta[c] - mean[c] * a;
  }
}

template <bool SPLIT>
__global__ void bn_apply_v8_kernel(const uint4* __restrict__ yhi, const uint4* __restrict__ ylo, const float* __restrict__ coef, int C,
                                   const uint4* __restrict__ rhi, const uint4* __restrict__ rlo, int relu, uint4* zhi,
                                   uint4* zlo, long long total8, int C8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % C8) * 8;
    float v[8], r[8];
    load8<SPLIT>(yhi, ylo, i, v);
    if (rhi) load8<SPLIT>(rhi, rlo, i, r);
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(coef + c0)), a1 = __ldg(reinterpret_cast<const float4*>(coef + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(coef + C + c0)), b1 = __ldg(reinterpret_cast<const float4*>(coef + C + c0 + 4));
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float o = fmaf(v[j], a[j], b[j]);
      if (rhi) o += r[j];
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
    store8<SPLIT>(zhi, zlo, i, v);
  }
}

template <bool SPLIT>
__global__ void __launch_bounds__(256) bn_bwd_reduce_v8_kernel(const uint4* __restrict__ dzhi, const uint4* __restrict__ dzlo,
                                                               const uint4* __restrict__ zhi, const uint4* __restrict__ zlo,
                                                               const uint4* __restrict__ yhi, const uint4* __restrict__ ylo,
                                                               const float* __restrict__ mean, const float* __restrict__ var,
                                                               float eps, int relu, long long pixels, int C8, int vpix,
                                                               double* __restrict__ s1, double* __restrict__ s2) {
  // fp64 accumulators: the cross-thread sums must not depend on the (non-deterministic) order of the atomics, otherwise
  // batch statistics differ in the last fp32 bit from run to run and the bf16 hi/lo re-split turns that into ~1e-5 jitter
  __shared__ float part[2][2048];
  const int groups = C8 < 32 ? C8 : 32;
  const int cg = blockIdx.y * 32 + (threadIdx.x % groups);
  const int lanes = 256 / groups, lane = threadIdx.x / groups;
  const long long p0 = (long long)blockIdx.x * vpix;
  const long long p1 = p0 + vpix < pixels ? p0 + vpix : pixels;
  float a[8], b[8], m[8], is[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = b[j] = 0.f;
    m[j] = mean[cg * 8 + j];
    is[j] = rsqrtf(var[cg * 8 + j] + eps);
  }
#pragma unroll 2
  for (long long p = p0 + lane; p < p1; p += lanes) {
    float g[8], z[8], y[8];
    load8<SPLIT>(dzhi, dzlo, p * C8 + cg, g);
    if (relu) load8<SPLIT>(zhi, zlo, p * C8 + cg, z);
    load8<SPLIT>(yhi, ylo, p * C8 + cg, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gg = (relu && !(z[j] > 0.f)) ? 0.f : g[j];
      a[j] += gg;
      b[j] = fmaf(gg, (y[j] - m[j]) * is[j], b[j]);
    }
  }
  {
    const int g2 = threadIdx.x % groups;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      part[0][lane * (groups * 8) + g2 * 8 + j] = a[j];
      part[1][lane * (groups * 8) + g2 * 8 + j] = b[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < groups * 8) {
    double sa = 0.0, sb = 0.0;
    for (int l = 0; l < lanes; ++l) {
      sa += (double)part[0][l * (groups * 8) + threadIdx.x];
      sb += (double)part[1][l * (groups * 8) + threadIdx.x];
    }
    const int c = blockIdx.y * 256 + threadIdx.x;
    atomicAdd(s1 + c, sa);
    atomicAdd(s2 + c, sb);
  }
}

// dy = A*g + Bc*y + D per channel:  A = gamma*invstd, Bc = -A*invstd*m2, D = -A*m1 + A*invstd*m2*mean   (m1 = s1/n, m2 = s2/n)
__global__ void bn_bwd_coef_kernel(const float* mean, const float* var, const float* gamma, float eps, const double* s1, const double* s2,
                                   double n, float* coef, int C) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float invstd = rsqrtf(var[c] + eps);
    const float A = gamma[c] * invstd;
    const float m1 = (float)(s1[c] / n), m2 = (float)(s2[c] / n);
    coef[c] = A;
    coef[C + c] = -A * invstd * m2;
    coef[2 * C + c] = -A * m1 + A * invstd * m2 * mean[c];
  }
}

template <bool SPLIT>
__global__ void bn_bwd_apply_v8_kernel(const uint4* __restrict__ dzhi, const uint4* __restrict__ dzlo, const uint4* __restrict__ zhi,
                                       const uint4* __restrict__ zlo, const uint4* __restrict__ yhi, const uint4* __restrict__ ylo,
                                       const float* __restrict__ coef, int C, int relu, uint4* dyhi, uint4* dylo, uint4* ghi, uint4* glo,
                                       long long total8, int C8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % C8) * 8;
    float g[8], z[8], y[8], o[8];
    load8<SPLIT>(dzhi, dzlo, i, g);
    if (relu) load8<SPLIT>(zhi, zlo, i, z);
    load8<SPLIT>(yhi, ylo, i, y);
    float A[8], Bc[8], D[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 t0 = __ldg(reinterpret_cast<const float4*>(coef + c0 + 4 * h));
      const float4 t1 = __ldg(reinterpret_cast<const float4*>(coef + C + c0 + 4 * h));
      const float4 t2 = __ldg(reinterpret_cast<const float4*>(coef + 2 * C + c0 + 4 * h));
      A[4 * h] = t0.x; A[4 * h + 1] = t0.y; A[4 * h + 2] = t0.z; A[4 * h + 3] = t0.w;
      Bc[4 * h] = t1.x; Bc[4 * h + 1] = t1.y; Bc[4 * h + 2] = t1.z; Bc[4 * h + 3] = t1.w;
      D[4 * h] = t2.x; D[4 * h + 1] = t2.y; D[4 * h + 2] = t2.z; D[4 * h + 3] = t2.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (relu && !(z[j] > 0.f)) g[j] = 0.f;
      o[j] = fmaf(A[j], g[j], fmaf(Bc[j], y[j], D[j]));
    }
    store8<SPLIT>(dyhi, dylo, i, o);
    if (ghi) store8<SPLIT>(ghi, glo, i, g);
  }
}

template <bool SPLIT>
__global__ void relu_bwd_v8_kernel(const uint4* __restrict__ dzhi, const uint4* __restrict__ dzlo, const uint4* __restrict__ zhi,
                                   const uint4* __restrict__ zlo, uint4* ohi, uint4* olo, long long total8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], z[8];
    load8<SPLIT>(dzhi, dzlo, i, g);
    load8<SPLIT>(zhi, zlo, i, z);
#pragma unroll
    for (int j = 0; j < 8; ++j) if (!(z[j] > 0.f)) g[j] = 0.f;
    store8<SPLIT>(ohi, olo, i, g);
  }
}

template <bool SPLIT>
__global__ void add_act_v8_kernel(const uint4* __restrict__ ahi, const uint4* __restrict__ alo, const uint4* __restrict__ bhi,
                                  const uint4* __restrict__ blo, uint4* ohi, uint4* olo, long long total8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8<SPLIT>(ahi, alo, i, a);
    load8<SPLIT>(bhi, blo, i, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    store8<SPLIT>(ohi, olo, i, a);
  }
}


// max_pool2d(3,2,1) backward, 8 channels per thread.  A window's arg-max is the FIRST maximum in row-major order
// (ATen semantics); it is recognised by comparing against the pooled value and checking the earlier window positions.
template <bool SPLIT>
__global__ void maxpool_bwd_v8_kernel(const uint4* __restrict__ xhi, const uint4* __restrict__ xlo, const uint4* __restrict__ dyhi,
                                      const uint4* __restrict__ dylo, uint4* dxhi, uint4* dxlo, int N, int H, int W, int C8, int OH,
                                      int OW) {
  long long total = (long long)N * H * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long p = i / C8;
    const int iw = (int)(p % W);
    long long q = p / W;
    const int ih = (int)(q % H);
    const int n = (int)(q / H);
    float xv[8], acc[8];
    load8<SPLIT>(xhi, xlo, i, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oh = ih / 2; oh <= (ih + 1) / 2 && oh < OH; ++oh)
      for (int ow = iw / 2; ow <= (iw + 1) / 2 && ow < OW; ++ow) {
        // is (ih, iw) the first maximum of window (oh, ow)?  per channel: not beaten by any element, not tied by an earlier one
        bool win[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) win[j] = true;
        for (int r = 0; r < 3; ++r) {
          const int y = 2 * oh - 1 + r;
          if (y < 0 || y >= H) continue;
          for (int s2 = 0; s2 < 3; ++s2) {
            const int x = 2 * ow - 1 + s2;
            if (x < 0 || x >= W || (y == ih && x == iw)) continue;
            float v[8];
            load8<SPLIT>(xhi, xlo, (((long long)n * H + y) * W + x) * C8 + c, v);
            const bool earlier = (y < ih) || (y == ih && x < iw);
#pragma unroll
            for (int j = 0; j < 8; ++j) win[j] = win[j] && (earlier ? (v[j] < xv[j]) : (v[j] <= xv[j]));
          }
        }
        float g[8];
        load8<SPLIT>(dyhi, dylo, (((long long)n * OH + oh) * OW + ow) * C8 + c, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += win[j] ? g[j] : 0.f;
      }
    store8<SPLIT>(dxhi, dxlo, i, acc);
  }
}

template <bool SPLIT>
__global__ void zero_insert2_v8_kernel(const uint4* __restrict__ dyhi, const uint4* __restrict__ dylo, uint4* ohi, uint4* olo, int N,
                                       int OH, int OW, int C8, int H, int W) {
  long long total = (long long)N * H * W * C8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C8);
    long long p = i / C8;
    const int x = (int)(p % W);
    long long q = p / W;
    const int y = (int)(q % H);
    const int n = (int)(q / H);
    uint4 h = make_uint4(0, 0, 0, 0), l = make_uint4(0, 0, 0, 0);
    if (!(y & 1) && !(x & 1) && (y >> 1) < OH && (x >> 1) < OW) {
      const long long o = (((long long)n * OH + (y >> 1)) * OW + (x >> 1)) * C8 + c;
      h = __ldg(dyhi + o);
      if (SPLIT) l = __ldg(dylo + o);
    }
    ohi[i] = h;
    if (SPLIT) olo[i] = l;
  }
}

// pixels per CTA of the per-channel reductions: aim at >= 4 CTAs per SM over the (pixel chunk, channel slice) grid
inline int reduce_vpix(long long pixels, int C8) {
  const long long slices = C8 < 32 ? 1 : C8 / 32;
  long long v = pixels * slices / (148 * 4);
  if (v > 512) v = 512;
  if (v < 32) v = 32;
  return (int)v;
}

inline bool vec_ok(int fmt, int C) {
  const int c8 = C / 8;
  return (fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && C % 8 == 0 && C <= 4096 && (c8 & (c8 - 1)) == 0;  // power-of-two groups
}

}  // namespace

#define ST ((cudaStream_t)stream)

extern "C" int mpn_channel_sums(const void* hi, const void* lo, long long pixels, int C, int cstride, int coffset, int fmt,
                                double* sum, double* sumsq, void* stream) {
  MPN_CHECK_ARG(hi && sum && pixels > 0 && C > 0 && cstride >= coffset + C, "mpn_channel_sums: bad argument");
  MPN_CUDA_OK(cudaMemsetAsync(sum, 0, sizeof(double) * C, ST));
  if (sumsq) MPN_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(double) * C, ST));
  if (vec_ok(fmt, C) && cstride == C && coffset == 0) {
    const int vpix = reduce_vpix(pixels, C / 8);
    dim3 grid(mpn_divup(pixels, vpix), C / 8 < 32 ? 1 : C / 8 / 32);
    if (fmt == MPN_FMT_BF16X2)
      channel_stats_v8_kernel<true><<<grid, 256, 0, ST>>>((const uint4*)hi, (const uint4*)lo, pixels, C / 8, vpix, sum, sumsq);
    else
      channel_stats_v8_kernel<false><<<grid, 256, 0, ST>>>((const uint4*)hi, nullptr, pixels, C / 8, vpix, sum, sumsq);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  channel_stats_kernel<<<mpn_divup(pixels, PIX_PER_BLOCK), 256, 0, ST>>>(hi, lo, pixels, C, cstride, coffset, fmt, sum, sumsq);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_bn_stats(const void* yhi, const void* ylo, long long pixels, int C, int fmt, float* mean, float* var,
                            double* workspace, void* stream) {
  MPN_CHECK_ARG(mean && var && workspace, "mpn_bn_stats: bad argument");
  int rc = mpn_channel_sums(yhi, ylo, pixels, C, C, 0, fmt, workspace, workspace + C, stream);
  if (rc) return rc;
  bn_finalize_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(workspace, workspace + C, (double)pixels, mean, var, C);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_bn_update_running(const float* mean, const float* var, float* running_mean, float* running_var, long long n,
                                     float momentum, int C, void* stream) {
  MPN_CHECK_ARG(mean && var && running_mean && running_var && C > 0 && n > 0, "mpn_bn_update_running: bad argument");
  bn_update_running_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(mean, var, running_mean, running_var, (double)n, momentum, C);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_bn_apply(const void* yhi, const void* ylo, const float* mean, const float* var, const float* gamma,
                            const float* beta, float eps, const void* rhi, const void* rlo, int relu, void* zhi, void* zlo,
                            long long pixels, int C, int fmt, float* coef, void* stream) {
  MPN_CHECK_ARG(yhi && mean && var && gamma && beta && zhi && pixels > 0 && C > 0, "mpn_bn_apply: bad argument");
  long long total = pixels * C;
  if (vec_ok(fmt, C)) {
    const long long t8 = total / 8;
    MPN_CHECK_ARG(coef, "mpn_bn_apply: coefficient workspace (3*C floats) missing");
    bn_fwd_coef_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(mean, var, gamma, beta, eps, coef, C);
    if (fmt == MPN_FMT_BF16X2)
      bn_apply_v8_kernel<true><<<grid_for(t8, 256), 256, 0, ST>>>((const uint4*)yhi, (const uint4*)ylo, coef, C, (const uint4*)rhi,
                                                                  (const uint4*)rlo, relu, (uint4*)zhi, (uint4*)zlo, t8, C / 8);
    else
      bn_apply_v8_kernel<false><<<grid_for(t8, 256), 256, 0, ST>>>((const uint4*)yhi, nullptr, coef, C, (const uint4*)rhi, nullptr, relu,
                                                                   (uint4*)zhi, nullptr, t8, C / 8);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  bn_apply_kernel<<<grid_for(total, 256), 256, 0, ST>>>(yhi, ylo, mean, var, gamma, beta, eps, rhi, rlo, relu, zhi, zlo, total, C, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_bn_backward(const void* dzhi, const void* dzlo, const void* zhi, const void* zlo, const void* yhi, const void* ylo,
                               const float* mean, const float* var, const float* gamma, float eps, int relu, long long pixels, int C,
                               int fmt, void* dyhi, void* dylo, void* ghi, void* glo, float* dgamma, float* dbeta,
                               double* workspace, float* coef, void* stream) {
  MPN_CHECK_ARG(dzhi && yhi && mean && var && gamma && dyhi && workspace && coef && pixels > 0 && C > 0, "mpn_bn_backward: bad argument");
  MPN_CHECK_ARG(!relu || zhi, "mpn_bn_backward: relu mask needs z");
  double* s1 = workspace;
  double* s2 = workspace + C;
  MPN_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(double) * 2 * C, ST));
  long long total = pixels * C;
  if (vec_ok(fmt, C)) {
    const int vpix = reduce_vpix(pixels, C / 8);
    dim3 grid(mpn_divup(pixels, vpix), C / 8 < 32 ? 1 : C / 8 / 32);
    const long long t8 = total / 8;
    if (fmt == MPN_FMT_BF16X2) {
      bn_bwd_reduce_v8_kernel<true><<<grid, 256, 0, ST>>>((const uint4*)dzhi, (const uint4*)dzlo, (const uint4*)zhi, (const uint4*)zlo,
                                                          (const uint4*)yhi, (const uint4*)ylo, mean, var, eps, relu, pixels, C / 8, vpix, s1, s2);
      bn_bwd_coef_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(mean, var, gamma, eps, s1, s2, (double)pixels, coef, C);
      bn_bwd_apply_v8_kernel<true><<<grid_for(t8, 256), 256, 0, ST>>>((const uint4*)dzhi, (const uint4*)dzlo, (const uint4*)zhi,
                                                                      (const uint4*)zlo, (const uint4*)yhi, (const uint4*)ylo, coef, C, relu,
                                                                      (uint4*)dyhi, (uint4*)dylo, (uint4*)ghi, (uint4*)glo, t8, C / 8);
    } else {
      bn_bwd_reduce_v8_kernel<false><<<grid, 256, 0, ST>>>((const uint4*)dzhi, nullptr, (const uint4*)zhi, nullptr, (const uint4*)yhi,
                                                           nullptr, mean, var, eps, relu, pixels, C / 8, vpix, s1, s2);
      bn_bwd_coef_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(mean, var, gamma, eps, s1, s2, (double)pixels, coef, C);
      bn_bwd_apply_v8_kernel<false><<<grid_for(t8, 256), 256, 0, ST>>>((const uint4*)dzhi, nullptr, (const uint4*)zhi, nullptr,
                                                                       (const uint4*)yhi, nullptr, coef, C, relu, (uint4*)dyhi, nullptr,
                                                                       (uint4*)ghi, nullptr, t8, C / 8);
    }
    MPN_LAUNCH_OK();
  } else {
    bn_bwd_reduce_kernel<<<mpn_divup(pixels, PIX_PER_BLOCK), 256, 0, ST>>>(dzhi, dzlo, zhi, zlo, yhi, ylo, mean, var, eps, relu, pixels,
                                                                            C, fmt, s1, s2);
    MPN_LAUNCH_OK();
    bn_bwd_apply_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dzhi, dzlo, zhi, zlo, yhi, ylo, mean, var, gamma, eps, relu, s1, s2,
                                                             (double)pixels, dyhi, dylo, ghi, glo, total, C, fmt);
    MPN_LAUNCH_OK();
  }
  if (dgamma) double_to_float_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(s2, dgamma, C, 1.f);
  if (dbeta) double_to_float_kernel<<<mpn_divup(C, 256), 256, 0, ST>>>(s1, dbeta, C, 1.f);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_double_to_float(const double* a, float* out, int n, float scale, void* stream) {
  MPN_CHECK_ARG(a && out && n > 0, "mpn_double_to_float: bad argument");
  double_to_float_kernel<<<mpn_divup(n, 256), 256, 0, ST>>>(a, out, n, scale);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_relu_backward(const void* dzhi, const void* dzlo, const void* zhi, const void* zlo, void* ohi, void* olo,
                                 long long n, int fmt, void* stream) {
  MPN_CHECK_ARG(dzhi && zhi && ohi && n > 0, "mpn_relu_backward: bad argument");
  if ((fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && n % 8 == 0) {
    if (fmt == MPN_FMT_BF16X2)
      relu_bwd_v8_kernel<true><<<grid_for(n / 8, 256), 256, 0, ST>>>((const uint4*)dzhi, (const uint4*)dzlo, (const uint4*)zhi,
                                                                      (const uint4*)zlo, (uint4*)ohi, (uint4*)olo, n / 8);
    else
      relu_bwd_v8_kernel<false><<<grid_for(n / 8, 256), 256, 0, ST>>>((const uint4*)dzhi, nullptr, (const uint4*)zhi, nullptr, (uint4*)ohi,
                                                                       nullptr, n / 8);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  relu_bwd_kernel<<<grid_for(n, 256), 256, 0, ST>>>(dzhi, dzlo, zhi, zlo, ohi, olo, n, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_add_act(const void* ahi, const void* alo, const void* bhi, const void* blo, void* ohi, void* olo, long long n,
                           int fmt, void* stream) {
  MPN_CHECK_ARG(ahi && bhi && ohi && n > 0, "mpn_add_act: bad argument");
  if ((fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && n % 8 == 0) {
    if (fmt == MPN_FMT_BF16X2)
      add_act_v8_kernel<true><<<grid_for(n / 8, 256), 256, 0, ST>>>((const uint4*)ahi, (const uint4*)alo, (const uint4*)bhi, (const uint4*)blo,
                                                                     (uint4*)ohi, (uint4*)olo, n / 8);
    else
      add_act_v8_kernel<false><<<grid_for(n / 8, 256), 256, 0, ST>>>((const uint4*)ahi, nullptr, (const uint4*)bhi, nullptr, (uint4*)ohi,
                                                                      nullptr, n / 8);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  add_act_kernel<<<grid_for(n, 256), 256, 0, ST>>>(ahi, alo, bhi, blo, ohi, olo, n, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_maxpool3x3s2_backward(const void* xhi, const void* xlo, const void* dyhi, const void* dylo, void* dxhi, void* dxlo,
                                         int N, int H, int W, int C, int fmt, void* stream) {
  MPN_CHECK_ARG(xhi && dyhi && dxhi && N > 0 && H > 0 && W > 0 && C > 0, "mpn_maxpool3x3s2_backward: bad argument");
  int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  long long total = (long long)N * H * W * C;
  if ((fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && C % 8 == 0) {
    if (fmt == MPN_FMT_BF16X2)
      maxpool_bwd_v8_kernel<true><<<grid_for(total / 8, 256), 256, 0, ST>>>((const uint4*)xhi, (const uint4*)xlo, (const uint4*)dyhi,
                                                                             (const uint4*)dylo, (uint4*)dxhi, (uint4*)dxlo, N, H, W, C / 8, OH, OW);
    else
      maxpool_bwd_v8_kernel<false><<<grid_for(total / 8, 256), 256, 0, ST>>>((const uint4*)xhi, nullptr, (const uint4*)dyhi, nullptr,
                                                                              (uint4*)dxhi, nullptr, N, H, W, C / 8, OH, OW);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  maxpool_bwd_kernel<<<grid_for(total, 256), 256, 0, ST>>>(xhi, xlo, dyhi, dylo, dxhi, dxlo, N, H, W, C, OH, OW, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_block_sum(const void* fhi, const void* flo, int f_cstride, int f_coffset, void* chi, void* clo, int N, int h,
                             int w, int C, int r, int fmt, void* stream) {
  MPN_CHECK_ARG(fhi && chi && N > 0 && h > 0 && w > 0 && C > 0 && r >= 1 && f_cstride >= f_coffset + C, "mpn_block_sum: bad argument");
  long long total = (long long)N * h * w * C;
  block_sum_kernel<<<grid_for(total, 256), 256, 0, ST>>>(fhi, flo, f_cstride, f_coffset, chi, clo, N, h, w, C, r, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_zero_insert2(const void* dyhi, const void* dylo, void* ohi, void* olo, int N, int OH, int OW, int C, int H, int W,
                                int fmt, void* stream) {
  MPN_CHECK_ARG(dyhi && ohi && N > 0 && OH > 0 && OW > 0 && C > 0 && H >= 2 * OH - 1 && W >= 2 * OW - 1, "mpn_zero_insert2: bad argument");
  long long total = (long long)N * H * W * C;
  if ((fmt == MPN_FMT_BF16 || fmt == MPN_FMT_BF16X2) && C % 8 == 0) {
    if (fmt == MPN_FMT_BF16X2)
      zero_insert2_v8_kernel<true><<<grid_for(total / 8, 256), 256, 0, ST>>>((const uint4*)dyhi, (const uint4*)dylo, (uint4*)ohi, (uint4*)olo,
                                                                              N, OH, OW, C / 8, H, W);
    else
      zero_insert2_v8_kernel<false><<<grid_for(total / 8, 256), 256, 0, ST>>>((const uint4*)dyhi, nullptr, (uint4*)ohi, nullptr, N, OH, OW,
                                                                               C / 8, H, W);
    MPN_LAUNCH_OK();
    return MPN_OK;
  }
  zero_insert2_kernel<<<grid_for(total, 256), 256, 0, ST>>>(dyhi, dylo, ohi, olo, N, OH, OW, C, H, W, fmt);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_mse_heatmap_loss(const float* pred, const float* gt, const float* weight, int B, int Cp, int H, int W, double* loss,
                                    void* dhi, void* dlo, int Cd, int fmt, float grad_scale, void* stream) {
  MPN_CHECK_ARG(pred && gt && weight && loss && dhi && B > 0 && Cp >= 18 && H > 0 && W > 0 && Cd >= 18, "mpn_mse_heatmap_loss: bad argument");
  long long total = (long long)B * H * W * Cd;
  mse_loss_kernel<<<grid_for(total, 256), 256, 0, ST>>>(pred, gt, weight, B, Cp, H, W, loss, dhi, dlo, Cd, fmt, grad_scale);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_pack_filter_dgrad_bf16(const float* w, void* hi, void* lo, int Cout, int Cin, int R, int S, int CoutPad,
                                          void* stream) {
  MPN_CHECK_ARG(w && hi && Cout > 0 && Cin > 0 && R > 0 && S > 0 && CoutPad >= Cout, "mpn_pack_filter_dgrad_bf16: bad argument");
  long long total = (long long)Cin * R * S * CoutPad;
  pack_filter_dgrad_kernel<<<grid_for(total, 256), 256, 0, ST>>>(w, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, Cout, Cin, R, S, CoutPad);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_unpack_filter_grad(const float* g, float* out_oihw, int Cout, int Cin, int R, int S, void* stream) {
  MPN_CHECK_ARG(g && out_oihw && Cout > 0 && Cin > 0 && R > 0 && S > 0, "mpn_unpack_filter_grad: bad argument");
  long long total = (long long)Cout * Cin * R * S;
  unpack_filter_grad_kernel<<<grid_for(total, 256), 256, 0, ST>>>(g, out_oihw, Cout, Cin, R, S);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_stem_unpack_filter_grad(const float* g, float* out_oihw, int Cout, void* stream) {
  MPN_CHECK_ARG(g && out_oihw && Cout > 0, "mpn_stem_unpack_filter_grad: bad argument");
  stem_unpack_filter_grad_kernel<<<mpn_divup(Cout * 147, 256), 256, 0, ST>>>(g, out_oihw, Cout);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
