// CUDA-core fp32 implicit-GEMM convolution (NHWC), the MPN_FMT_F32 path and the stem of every mode.
//
// It is the "exact" mode of the library (plain fp32 FMA accumulation like the reference's fp32 conv,
// fpn.py:14-25 / posenet.py:165-187) and the on-device reference the tcgen05 kernel is validated against.
// Tile: 64 output pixels x 64 output channels per 256-thread CTA, 4x4 outputs per thread, K chunks of 16
// staged in shared memory; filters are pre-packed [R][S][Cin][CoutPad] so both operands load coalesced.
#include "mpn_common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct SimtParams {
  mpn_conv_desc d;
  mpn_conv_ptrs p;
  int in_fmt;  // format of x (F32 for the stem variant)
  int M;       // N*OH*OW
  int K;       // R*S*Cin
};

// One output element through the whole epilogue (order documented in mpn_b200.h).
__device__ __forceinline__ void epilogue_store(const SimtParams& P, int n, int oh, int ow, int co, float acc) {
  const mpn_conv_desc& d = P.d;
  float v = acc;
  if (P.p.scale) v = __fmul_rn(v, P.p.scale[co]);
  if (P.p.bias) v = __fadd_rn(v, P.p.bias[co]);
  if (d.res_cstride > 0) v += mpn_load_act(P.p.res_hi, P.p.res_lo, (((long long)n * d.OH + oh) * d.OW + ow) * d.res_cstride + co, d.fmt);
  if (d.up_cstride > 0) {
    int sh = mpn_nearest_src(oh, d.up_h, d.OH), sw = mpn_nearest_src(ow, d.up_w, d.OW);
    float u = mpn_load_act(P.p.up_hi, P.p.up_lo, (((long long)n * d.up_h + sh) * d.up_w + sw) * d.up_cstride + co, d.fmt);
    v = u + v;  // F.upsample(x) + y  (fpn.py:95)
  }
  if (d.flags & MPN_EPI_RELU) v = fmaxf(v, 0.f);
  if (d.flags & MPN_EPI_SIGMOID) v = 1.f / (1.f + expf(-v));
  const int rep = d.out_rep;
  const int OHr = d.OH * rep, OWr = d.OW * rep;
  const long long nstride = d.out_nstride > 0 ? d.out_nstride
                          : (d.out_mode == MPN_OUT_F32_NCHW ? (long long)d.Cout * OHr * OWr : (long long)OHr * OWr * d.out_cstride);
  for (int ry = 0; ry < rep; ++ry)
    for (int rx = 0; rx < rep; ++rx) {
      int y = oh * rep + ry, x = ow * rep + rx;
      if (d.out_mode == MPN_OUT_F32_NCHW) {
        ((float*)P.p.y_hi)[(long long)n * nstride + ((long long)(d.out_coffset + co) * OHr + y) * OWr + x] = v;
      } else {
        long long idx = (long long)n * nstride + ((long long)y * OWr + x) * d.out_cstride + d.out_coffset + co;
        if (d.out_mode == MPN_OUT_F32_NHWC) ((float*)P.p.y_hi)[idx] = v;
        else mpn_store_act(P.p.y_hi, P.p.y_lo, idx, d.fmt, v);
      }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(NT) conv_simt_kernel(const SimtParams P) {
  const mpn_conv_desc& d = P.d;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid % 16, ty = tid / 16;  // tx -> 4 channels, ty -> 4 pixels

  // A-load assignment: thread -> (pixel a_m, k-quad a_kq)
  const int a_m = tid / 4, a_kq = tid % 4;
  int a_n = 0, a_oh = 0, a_ow = 0;
  const bool a_valid = (m0 + a_m) < P.M;
  if (a_valid) {
    int m = m0 + a_m;
    a_ow = m % d.OW;
    int t = m / d.OW;
    a_oh = t % d.OH;
    a_n = t / d.OH;
  }
  // B-load assignment: thread -> (k row b_k, 4 channels b_c4)
  const int b_k = tid / 16, b_c4 = (tid % 16) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const float* wf = (const float*)P.p.w_hi;
  for (int k0 = 0; k0 < P.K; k0 += BK) {
    // ---- A tile
    float av[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {  // Cin % 16 == 0: the BK chunk sits inside one filter tap
      int tap = k0 / d.Cin, c0 = k0 - tap * d.Cin;
      int r = tap / d.S, s = tap - r * d.S;
      int ih = a_oh * d.stride - d.pad + r, iw = a_ow * d.stride - d.pad + s;
      if (a_valid && ih >= 0 && ih < d.H && iw >= 0 && iw < d.W) {
        long long base = (((long long)a_n * d.H + ih) * d.W + iw) * d.in_cstride + c0 + a_kq * 4;
        if (P.in_fmt == MPN_FMT_F32) {
          float4 t = *reinterpret_cast<const float4*>((const float*)P.p.x_hi + base);
          av[0] = t.x; av[1] = t.y; av[2] = t.z; av[3] = t.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) av[i] = mpn_load_act(P.p.x_hi, P.p.x_lo, base + i, P.in_fmt);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int k = k0 + a_kq * 4 + i;
        if (a_valid && k < P.K) {
          int tap = k / d.Cin, c = k - tap * d.Cin;
          int r = tap / d.S, s = tap - r * d.S;
          int ih = a_oh * d.stride - d.pad + r, iw = a_ow * d.stride - d.pad + s;
          if (ih >= 0 && ih < d.H && iw >= 0 && iw < d.W)
            av[i] = mpn_load_act(P.p.x_hi, P.p.x_lo, (((long long)a_n * d.H + ih) * d.W + iw) * d.in_cstride + c, P.in_fmt);
        }
      }
    }
    // ---- B tile
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      int k = k0 + b_k;
      if (k < P.K && (n0 + b_c4) < d.w_cout_pad) bv = *reinterpret_cast<const float4*>(wf + (long long)k * d.w_cout_pad + n0 + b_c4);
    }
    __syncthreads();  // previous iteration's reads are done
#pragma unroll
    for (int i = 0; i < 4; ++i) As[a_kq * 4 + i][a_m] = av[i];
    *reinterpret_cast<float4*>(&Bs[b_k][b_c4]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float aa[4] = {a.x, a.y, a.z, a.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
    }
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
    int ow = m % d.OW;
    int t = m / d.OW;
    int oh = t % d.OH;
    int n = t / d.OH;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co < d.Cout) epilogue_store(P, n, oh, ow, co, acc[i][j]);
    }
  }
}

int check_desc(const mpn_conv_desc* d, const mpn_conv_ptrs* p) {
  MPN_CHECK_ARG(d && p, "conv: null descriptor");
  MPN_CHECK_ARG(d->N > 0 && d->H > 0 && d->W > 0 && d->Cin > 0 && d->Cout > 0 && d->R > 0 && d->S > 0 && d->stride > 0 && d->pad >= 0,
                "conv: bad problem size");
  MPN_CHECK_ARG(d->OH == (d->H + 2 * d->pad - d->R) / d->stride + 1 && d->OW == (d->W + 2 * d->pad - d->S) / d->stride + 1,
                "conv: OH/OW do not match (H + 2*pad - R)/stride + 1");
  MPN_CHECK_ARG(d->in_cstride >= d->Cin || d->k_overlap, "conv: in_cstride < Cin");
  MPN_CHECK_ARG(d->out_rep == 1 || d->out_rep == 2 || d->out_rep == 4 || d->out_rep == 8, "conv: out_rep must be 1,2,4,8");
  MPN_CHECK_ARG(d->out_mode >= 0 && d->out_mode <= 2, "conv: bad out_mode");
  MPN_CHECK_ARG(d->out_mode == MPN_OUT_F32_NCHW || d->out_cstride >= d->out_coffset + d->Cout, "conv: out_cstride too small");
  MPN_CHECK_ARG(p->x_hi && p->w_hi && p->y_hi, "conv: null x/w/y");
  MPN_CHECK_ARG(d->res_cstride == 0 || p->res_hi, "conv: residual pointer missing");
  MPN_CHECK_ARG(d->up_cstride == 0 || (p->up_hi && d->up_h > 0 && d->up_w > 0), "conv: upsample source missing");
  if (d->fmt == MPN_FMT_BF16X2 || d->fmt == MPN_FMT_F16F8) {
    MPN_CHECK_ARG(d->out_mode != MPN_OUT_ACT || p->y_lo, "conv: BF16X2 output needs y_lo");
    MPN_CHECK_ARG(d->res_cstride == 0 || p->res_lo, "conv: BF16X2 residual needs res_lo");
    MPN_CHECK_ARG(d->up_cstride == 0 || p->up_lo, "conv: BF16X2 upsample source needs up_lo");
  }
  return MPN_OK;
}

int launch_simt(const mpn_conv_desc* d, const mpn_conv_ptrs* p, int in_fmt, void* stream) {
  int rc = check_desc(d, p);
  if (rc) return rc;
  MPN_CHECK_ARG(d->w_cout_pad >= d->Cout && d->w_cout_pad % 4 == 0, "conv(fp32): w_cout_pad must be >= Cout and a multiple of 4");
  MPN_CHECK_ARG(!d->k_overlap && !d->in_wpitch && !d->in_hpitch, "conv(fp32): pitched / overlapped inputs are a tcgen05-path feature");
  MPN_CHECK_ARG(d->gat_n == 0, "conv(fp32): phase-class addends are a tcgen05-path feature");
  MPN_CHECK_ARG(in_fmt != MPN_FMT_BF16X2 || p->x_lo, "conv: BF16X2 input needs x_lo");
  SimtParams P;
  P.d = *d;
  P.p = *p;
  P.in_fmt = in_fmt;
  long long M = (long long)d->N * d->OH * d->OW;
  MPN_CHECK_ARG(M < (1LL << 31), "conv: too many output pixels");
  P.M = (int)M;
  P.K = d->R * d->S * d->Cin;
  dim3 grid(mpn_divup(M, BM), mpn_divup(d->Cout, BN));
  bool vec = (d->Cin % BK == 0) && (d->in_cstride % 4 == 0);
  if (vec) conv_simt_kernel<true><<<grid, NT, 0, (cudaStream_t)stream>>>(P);
  else conv_simt_kernel<false><<<grid, NT, 0, (cudaStream_t)stream>>>(P);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

}  // namespace

int mpn_conv_tc_launch(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream);  // mpn_conv_tc.cu

extern "C" int mpn_conv2d_fwd(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream) {
  MPN_CHECK_ARG(d, "conv: null descriptor");
  if (d->fmt == MPN_FMT_F32) return launch_simt(d, p, MPN_FMT_F32, stream);
  if (d->fmt == MPN_FMT_BF16 || d->fmt == MPN_FMT_BF16X2 || d->fmt == MPN_FMT_F16F8) {
    int rc = check_desc(d, p);
    if (rc) return rc;
    return mpn_conv_tc_launch(d, p, stream);
  }
  mpn_set_error("conv: unknown fmt %d", d->fmt);
  return MPN_ERR_ARG;
}

int mpn_conv_tc_launch_multi(const mpn_conv_desc* ds, const mpn_conv_ptrs* ps, int nseg, void* stream);  // mpn_conv_tc.cu

extern "C" int mpn_conv2d_fwd_multi(const mpn_conv_desc* ds, const mpn_conv_ptrs* ps, int nseg, void* stream) {
  MPN_CHECK_ARG(ds && ps && nseg >= 1, "conv (multi-level): null descriptor array");
  MPN_CHECK_ARG(ds->fmt == MPN_FMT_BF16 || ds->fmt == MPN_FMT_BF16X2 || ds->fmt == MPN_FMT_F16F8,
                "conv (multi-level): tensor-core formats only (got fmt %d)", ds->fmt);
  for (int i = 0; i < nseg; ++i) {
    int rc = check_desc(ds + i, ps + i);
    if (rc) return rc;
  }
  return mpn_conv_tc_launch_multi(ds, ps, nseg, stream);
}

extern "C" int mpn_conv2d_fwd_f32in(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream) {
  MPN_CHECK_ARG(d, "conv: null descriptor");
  MPN_CHECK_ARG(d->fmt != MPN_FMT_F16F8, "conv(fp32 input): F16F8 outputs come from the tensor-core stem only (MPN_TC_STEM=1)");
  return launch_simt(d, p, MPN_FMT_F32, stream);
}
