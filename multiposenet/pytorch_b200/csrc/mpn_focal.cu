// RetinaNet focal + smooth-L1 loss of the detection subnet: network/losses.py:5-137 (calc_iou, FocalLoss.forward)
// for a whole batch in two launches, with the gradients w.r.t. the class scores and box regressions.
//
// Reference, per image j (a Python loop over the batch): annotations with class != -1; IoU(anchor, annotation) WITHOUT the +1
// pixel convention and with the union clamped at 1e-8 (losses.py:5-24); per anchor the best annotation (first maximum);
// IoU < 0.4 -> negative, IoU >= 0.5 -> positive (one-hot target of the annotation's class), in between -> ignored;
// classification clamped to [1e-4, 1 - 1e-4]; focal term alpha_t * (1 - p_t)^2 * BCE (alpha .25, gamma 2) summed and divided by
// max(#positives, 1); smooth-L1 (beta 1/9) of (targets - regression) over the positives, targets = ((gcx - acx)/aw, (gcy - acy)/ah,
// log(gw/aw), log(gh/ah)) / (.1, .1, .2, .2) with gw, gh clamped at >= 1 AFTER the centre was taken, averaged over #positives x 4.
// An image without annotations contributes 0 to both.  The caller averages the per-image values over the batch
// (losses.py:137, posenet.py:413-416).
//
// Per-anchor arithmetic is fp32 in the reference's operation order; the sums run in fp64 (the reference: torch.sum in fp32).
#include "mpn_common.cuh"

namespace {

constexpr int FL_THREADS = 256;
constexpr int FL_MAX_ANN = 256;   // annotations per image staged in shared memory

struct FocalAcc {   // per image
  double cls_sum, reg_sum;
  int num_pos, num_ann;
};

__device__ __forceinline__ double block_sum_d(double v, double* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < FL_THREADS / 32; ++i) t += red[i];
  return t;   // valid in thread 0
}

// smooth-L1 element (losses.py:124-130) and its derivative w.r.t. the regression value
__device__ __forceinline__ float smooth_l1(float target, float r, float* grad) {
  const float d = target - r, ad = fabsf(d);
  const bool quad = ad <= 1.0f / 9.0f;
  *grad = quad ? -9.0f * d : (d > 0.f ? -1.f : (d < 0.f ? 1.f : 0.f));
  return quad ? 0.5f * 9.0f * ad * ad : ad - 0.5f / 9.0f;
}

// pass 1: assignment, per-anchor loss terms, per-image sums.  state[b][a] = argmax annotation (>= 0) for positives, -1 negative,
// -2 ignored.
__global__ void __launch_bounds__(FL_THREADS) focal_assign_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                                  const float* __restrict__ anchors, const float* __restrict__ ann,
                                                                  int A, int C, int M, int* __restrict__ state,
                                                                  FocalAcc* __restrict__ acc) {
  __shared__ float sann[FL_MAX_ANN][5];
  __shared__ int nann;
  __shared__ double red[FL_THREADS / 32];
  const int b = blockIdx.y, a = blockIdx.x * FL_THREADS + threadIdx.x;
  if (threadIdx.x == 0) {   // compaction in order (bbox_annotation[bbox_annotation[:, 4] != -1], losses.py:50-51)
    int n = 0;
    for (int m = 0; m < M && n < FL_MAX_ANN; ++m) {
      const float* r = ann + ((long long)b * M + m) * 5;
      if (r[4] != -1.f) {
        for (int k = 0; k < 5; ++k) sann[n][k] = r[k];
        ++n;
      }
    }
    nann = n;
    if (blockIdx.x == 0) acc[b].num_ann = n;
  }
  __syncthreads();
  const int n = nann;
  double lc = 0.0, lr = 0.0;
  int pos = 0;
  if (n > 0 && a < A) {
    const float4 an = reinterpret_cast<const float4*>(anchors)[a];
    const float aw = an.z - an.x, ah = an.w - an.y;
    float best = -INFINITY;
    int arg = 0;
    for (int m = 0; m < n; ++m) {   // calc_iou (losses.py:5-24)
      const float bx0 = sann[m][0], by0 = sann[m][1], bx1 = sann[m][2], by1 = sann[m][3];
      const float area = __fmul_rn(bx1 - bx0, by1 - by0);
      const float iw = fmaxf(fminf(an.z, bx1) - fmaxf(an.x, bx0), 0.f);
      const float ih = fmaxf(fminf(an.w, by1) - fmaxf(an.y, by0), 0.f);
      const float inter = __fmul_rn(iw, ih);
      const float ua = fmaxf(__fsub_rn(__fadd_rn(__fmul_rn(aw, ah), area), inter), 1e-8f);
      const float iou = __fdiv_rn(inter, ua);
      if (iou > best) { best = iou; arg = m; }   // first maximum
    }
    const bool positive = best >= 0.5f, negative = best < 0.4f;
    const int tclass = positive ? (int)sann[arg][4] : -1;
    state[(long long)b * A + a] = positive ? arg : (negative ? -1 : -2);
    if (positive || negative) {
      for (int c = 0; c < C; ++c) {   // losses.py:78-92
        const float p = fminf(fmaxf(cls[((long long)b * A + a) * C + c], 1e-4f), 1.0f - 1e-4f);
        const bool one = positive && c == tclass;
        const float alpha = one ? 0.25f : 0.75f;
        const float fw0 = one ? 1.f - p : p;
        const float fw = alpha * (fw0 * fw0);
        const float bce = one ? -logf(p) : -logf(1.0f - p);
        lc += (double)(fw * bce);
      }
    }
    if (positive) {
      pos = 1;
      const float acx = __fadd_rn(an.x, __fmul_rn(0.5f, aw)), acy = __fadd_rn(an.y, __fmul_rn(0.5f, ah));
      float gw = sann[arg][2] - sann[arg][0], gh = sann[arg][3] - sann[arg][1];
      const float gcx = __fadd_rn(sann[arg][0], __fmul_rn(0.5f, gw)), gcy = __fadd_rn(sann[arg][1], __fmul_rn(0.5f, gh));
      gw = fmaxf(gw, 1.f);
      gh = fmaxf(gh, 1.f);
      const float t[4] = {__fdiv_rn(__fdiv_rn(gcx - acx, aw), 0.1f), __fdiv_rn(__fdiv_rn(gcy - acy, ah), 0.1f),
                          __fdiv_rn(logf(__fdiv_rn(gw, aw)), 0.2f), __fdiv_rn(logf(__fdiv_rn(gh, ah)), 0.2f)};
      const float4 r = reinterpret_cast<const float4*>(reg)[(long long)b * A + a];
      const float rv[4] = {r.x, r.y, r.z, r.w};
      float g;
      for (int k = 0; k < 4; ++k) lr += (double)smooth_l1(t[k], rv[k], &g);
    }
  }
  const double sc = block_sum_d(lc, red);
  const double sr = block_sum_d(lr, red);
  const int np = __syncthreads_count(pos);
  if (threadIdx.x == 0 && n > 0) {
    if (sc != 0.0) atomicAdd(&acc[b].cls_sum, sc);
    if (sr != 0.0) atomicAdd(&acc[b].reg_sum, sr);
    if (np) atomicAdd(&acc[b].num_pos, np);
  }
}

// pass 2: per-image losses and the gradients of gscale_cls * mean_b(cls_loss_b) + gscale_reg * mean_b(reg_loss_b)
__global__ void __launch_bounds__(FL_THREADS) focal_grad_kernel(const float* __restrict__ cls, const float* __restrict__ reg,
                                                                const float* __restrict__ anchors, const float* __restrict__ ann,
                                                                int B, int A, int C, int M, const int* __restrict__ state,
                                                                const FocalAcc* __restrict__ acc, float* __restrict__ cls_loss,
                                                                float* __restrict__ reg_loss, float* __restrict__ dcls,
                                                                float* __restrict__ dreg, float gscale_cls, float gscale_reg) {
  __shared__ float sann[FL_MAX_ANN][5];
  const int b = blockIdx.y, a = blockIdx.x * FL_THREADS + threadIdx.x;
  const FocalAcc ac = acc[b];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    cls_loss[b] = ac.num_ann > 0 ? (float)(ac.cls_sum / (double)max(ac.num_pos, 1)) : 0.f;            // losses.py:94
    reg_loss[b] = (ac.num_ann > 0 && ac.num_pos > 0) ? (float)(ac.reg_sum / (4.0 * ac.num_pos)) : 0.f;  // losses.py:131
  }
  if (!dcls && !dreg) return;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int m = 0; m < M && n < FL_MAX_ANN; ++m) {
      const float* r = ann + ((long long)b * M + m) * 5;
      if (r[4] != -1.f) {
        for (int k = 0; k < 5; ++k) sann[n][k] = r[k];
        ++n;
      }
    }
  }
  __syncthreads();
  if (a >= A) return;
  const long long ia = (long long)b * A + a;
  const int st = ac.num_ann > 0 ? state[ia] : -2;
  const float wc = gscale_cls / ((float)B * (float)max(ac.num_pos, 1));
  if (dcls) {
    for (int c = 0; c < C; ++c) {
      float g = 0.f;
      const float praw = cls[ia * C + c];
      if (st != -2 && praw >= 1e-4f && praw <= 1.0f - 1e-4f) {   // torch.clamp passes the gradient inside [min, max]
        const float p = praw;
        const bool one = st >= 0 && c == (int)sann[st][4];
        if (one) g = 0.25f * (2.f * (1.f - p) * logf(p) - (1.f - p) * (1.f - p) / p);
        else g = 0.75f * (-2.f * p * logf(1.f - p) + p * p / (1.f - p));
        g *= wc;
      }
      dcls[ia * C + c] = g;
    }
  }
  if (dreg) {
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (st >= 0) {
      const float4 an = reinterpret_cast<const float4*>(anchors)[a];
      const float aw = an.z - an.x, ah = an.w - an.y;
      const float acx = __fadd_rn(an.x, __fmul_rn(0.5f, aw)), acy = __fadd_rn(an.y, __fmul_rn(0.5f, ah));
      float gw = sann[st][2] - sann[st][0], gh = sann[st][3] - sann[st][1];
      const float gcx = __fadd_rn(sann[st][0], __fmul_rn(0.5f, gw)), gcy = __fadd_rn(sann[st][1], __fmul_rn(0.5f, gh));
      gw = fmaxf(gw, 1.f);
      gh = fmaxf(gh, 1.f);
      const float t[4] = {__fdiv_rn(__fdiv_rn(gcx - acx, aw), 0.1f), __fdiv_rn(__fdiv_rn(gcy - acy, ah), 0.1f),
                          __fdiv_rn(logf(__fdiv_rn(gw, aw)), 0.2f), __fdiv_rn(logf(__fdiv_rn(gh, ah)), 0.2f)};
      const float4 r = reinterpret_cast<const float4*>(reg)[ia];
      const float rv[4] = {r.x, r.y, r.z, r.w};
      const float wr = gscale_reg / ((float)B * 4.f * (float)ac.num_pos);
      float g[4];
      for (int k = 0; k < 4; ++k) { smooth_l1(t[k], rv[k], &g[k]); g[k] *= wr; }
      g4 = make_float4(g[0], g[1], g[2], g[3]);
    }
    reinterpret_cast<float4*>(dreg)[ia] = g4;
  }
}

}  // namespace

extern "C" size_t mpn_focal_loss_workspace_bytes(int B, int A) {
  if (B <= 0 || A <= 0) return 0;
  return (size_t)B * sizeof(FocalAcc) + (size_t)B * A * sizeof(int) + 256;
}

extern "C" int mpn_focal_loss(const float* cls, const float* reg, const float* anchors, const float* annotations, int B, int A, int C,
                              int M, float* cls_loss, float* reg_loss, float* dcls, float* dreg, float gscale_cls, float gscale_reg,
                              void* workspace, size_t workspace_bytes, void* stream) {
  MPN_CHECK_ARG(cls && reg && anchors && annotations && cls_loss && reg_loss && workspace, "mpn_focal_loss: null pointer");
  MPN_CHECK_ARG(B > 0 && A > 0 && C > 0 && M >= 0 && B <= 65535, "mpn_focal_loss: bad sizes");
  MPN_CHECK_ARG(M <= FL_MAX_ANN, "mpn_focal_loss: at most %d annotation rows per image (got %d)", FL_MAX_ANN, M);
  MPN_CHECK_ARG(workspace_bytes >= mpn_focal_loss_workspace_bytes(B, A), "mpn_focal_loss: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  FocalAcc* acc = (FocalAcc*)workspace;
  int* state = (int*)((char*)workspace + (((size_t)B * sizeof(FocalAcc) + 255) & ~(size_t)255));
  MPN_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)B * sizeof(FocalAcc), st));
  dim3 grid(mpn_divup(A, FL_THREADS), B);
  focal_assign_kernel<<<grid, FL_THREADS, 0, st>>>(cls, reg, anchors, annotations, A, C, M, state, acc);
  MPN_LAUNCH_OK();
  focal_grad_kernel<<<grid, FL_THREADS, 0, st>>>(cls, reg, anchors, annotations, B, A, C, M, state, acc, cls_loss, reg_loss, dcls, dreg,
                                                 gscale_cls, gscale_reg);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
