// Pose-residual-network assignment on the device: evaluate/tester.py:333-513 (Tester.prn_process)
// for all person boxes of a batch of images at once.  The reference does this per image in Python: a triple loop over
// peaks x boxes, one skimage gaussian per (box, joint) plane, one batch-1 PRN call per box, np.argwhere/np.sum per peak
// and list comprehensions for the assignment table.
//
//   prn_scatter_kernel  : one CTA per box; every peak of the box's image is tested against the enlarged box (:366-369)
//                         and dropped into its grid cell (:371-392) -- last writer in peak-id order wins (atomicMax)
//   prn_blur_kernel     : one CTA per (box, joint) plane: one-hot plane -> separable sigma=1 Gaussian in float64 with
//                         scipy's operation order (skimage.filters.gaussian, :396-398) -> fp32 PRN input [P,h,w,17]
//   prn_score_kernel    : one thread per grid cell that holds a peak: sum of the <=15x15 window of the PRN output
//                         around it (:412-430) in numpy's pairwise float32 order -> S[box, peak]
//   prn_assign_kernel   : one CTA per (image, joint type), one thread per box: the greedy table walk (:432-470)
//   prn_fallback_kernel : one CTA per (box, joint) plane: arg-max fallback (:471-483)
// Every float64 / float32 operation that decides an index is issued with explicit round-to-nearest intrinsics so that
// no fused multiply-add changes a comparison (oracle/prn_oracle.py is the numpy statement of the same arithmetic).
#include <stdint.h>

#include "mpn_common.cuh"

namespace {

constexpr int NJ = 17;
constexpr int PRN_THREADS = 256;

struct GaussW {
  double w[5];
};

__global__ void __launch_bounds__(PRN_THREADS) prn_scatter_kernel(const double* __restrict__ peak_xy, const int32_t* __restrict__ peak_type,
                                                                  const int32_t* __restrict__ peak_img_start,
                                                                  const double* __restrict__ boxes, const int32_t* __restrict__ box_img,
                                                                  int gh, int gw, double in_thres, int32_t* __restrict__ owner) {
  const int p = blockIdx.x;
  const int img = box_img[p];
  const double b0 = boxes[p * 4 + 0], b1 = boxes[p * 4 + 1], b2 = boxes[p * 4 + 2], b3 = boxes[p * 4 + 3];
  const double lo_x = __dsub_rn(b0, __dmul_rn(b2, in_thres)), lo_y = __dsub_rn(b1, __dmul_rn(b3, in_thres));
  const double grow = __dadd_rn(1.0, in_thres);
  const double hi_x = __dadd_rn(b0, __dmul_rn(b2, grow)), hi_y = __dadd_rn(b1, __dmul_rn(b3, grow));
  const double x_scale = __ddiv_rn((double)gw, ceil(b2)), y_scale = __ddiv_rn((double)gh, ceil(b3));
  for (int k = peak_img_start[img] + threadIdx.x; k < peak_img_start[img + 1]; k += PRN_THREADS) {
    const double px = peak_xy[2 * k], py = peak_xy[2 * k + 1];
    if (!(px > lo_x && py > lo_y && px < hi_x && py < hi_y)) continue;
    int x0 = (int)__dmul_rn(__dsub_rn(px, b0), x_scale);  // int(): truncation toward zero
    int y0 = (int)__dmul_rn(__dsub_rn(py, b1), y_scale);
    // tester.py:377-390 is an elif chain: exactly one correction is applied
    if (x0 >= gw && y0 >= gh) { x0 = gw - 1; y0 = gh - 1; }
    else if (x0 >= gw) x0 = gw - 1;
    else if (y0 >= gh) y0 = gh - 1;
    else if (x0 < 0 && y0 < 0) { x0 = 0; y0 = 0; }
    else if (x0 < 0) x0 = 0;
    else if (y0 < 0) y0 = 0;
    if (x0 < 0) x0 += gw;  // an index the chain left negative wraps (numpy indexing, :392)
    if (y0 < 0) y0 += gh;
    if (x0 < 0 || y0 < 0) continue;  // beyond one wrap the reference raises IndexError; in_thres < 1 rules it out
    atomicMax(&owner[(((long long)p * NJ + peak_type[k]) * gh + y0) * gw + x0], k);
  }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// dynamic shared memory: 2 * gh * gw doubles
__global__ void __launch_bounds__(PRN_THREADS) prn_blur_kernel(const int32_t* __restrict__ owner, int gh, int gw, GaussW g,
                                                               float* __restrict__ inp, int32_t* __restrict__ plane_has) {
  extern __shared__ double sm[];
  double* a = sm;
  double* v = sm + gh * gw;
  const int plane = blockIdx.x, p = plane / NJ, t = plane - p * NJ;
  const int32_t* o = owner + (long long)plane * gh * gw;
  int any = 0;
  for (int i = threadIdx.x; i < gh * gw; i += PRN_THREADS) {
    const int h = o[i] >= 0;
    a[i] = h ? 1.0 : 0.0;
    any |= h;
  }
  any = __syncthreads_or(any);
  if (threadIdx.x == 0) plane_has[plane] = any;
  if (!any) return;  // inp was zero-filled: the Gaussian of a zero plane is zero
  // scipy NI_Correlate1D, symmetric weights, mode 'nearest': c*w0 + sum_{j=4..1} (a[-j] + a[+j])*w[j]; axis 0 first
  for (int i = threadIdx.x; i < gh * gw; i += PRN_THREADS) {
    const int y = i / gw, x = i - y * gw;
    double s = __dmul_rn(a[i], g.w[0]);
#pragma unroll
    for (int j = 4; j >= 1; --j)
      s = __dadd_rn(s, __dmul_rn(__dadd_rn(a[clampi(y - j, 0, gh - 1) * gw + x], a[clampi(y + j, 0, gh - 1) * gw + x]), g.w[j]));
    v[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < gh * gw; i += PRN_THREADS) {
    const int y = i / gw, x = i - y * gw;
    double s = __dmul_rn(v[i], g.w[0]);
#pragma unroll
    for (int j = 4; j >= 1; --j)
      s = __dadd_rn(s, __dmul_rn(__dadd_rn(v[y * gw + clampi(x - j, 0, gw - 1)], v[y * gw + clampi(x + j, 0, gw - 1)]), g.w[j]));
    inp[((long long)p * gh * gw + i) * NJ + t] = (float)s;
  }
}

// numpy pairwise_sum over <= 128 float32 values of a row-major window (element i -> row i / cw, column i % cw)
struct Window {
  const float* base;  // plane origin (+ joint offset); pixel stride NJ floats
  int r0, c0, cw, gw;
  __device__ __forceinline__ float at(int i) const {
    const int r = i / cw, c = i - r * cw;
    return base[((r0 + r) * gw + c0 + c) * NJ];
  }
};

__device__ float pairwise_block(const Window& wn, int start, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, wn.at(start + i));
    return res;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = wn.at(start + j);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], wn.at(start + i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])), __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __fadd_rn(res, wn.at(start + i));
  return res;
}

__global__ void __launch_bounds__(PRN_THREADS) prn_score_kernel(const int32_t* __restrict__ owner, const float* __restrict__ output,
                                                                const int32_t* __restrict__ box_img,
                                                                const int32_t* __restrict__ peak_img_start, int P, int gh, int gw,
                                                                int kmax, double* __restrict__ S, int32_t* __restrict__ colhit) {
  const long long cell = (long long)blockIdx.x * PRN_THREADS + threadIdx.x;
  if (cell >= (long long)P * NJ * gh * gw) return;
  const int k = owner[cell];
  if (k < 0) return;
  const int x = (int)(cell % gw), y = (int)((cell / gw) % gh), t = (int)((cell / ((long long)gw * gh)) % NJ), p = (int)(cell / ((long long)gw * gh * NJ));
  // prn_gaussian.py:122-146 crop(img, (y, x), N=15)
  int r0 = y - 7, c0 = x - 7, r1 = y + 8, c1 = x + 8;
  if (r0 < 0) r0 = 0;
  if (c0 < 0) c0 = 0;
  if (r1 > gh - 1) r1 = gh;
  if (c1 > gw - 1) c1 = gw;
  Window wn{output + (long long)p * gh * gw * NJ + t, r0, c0, c1 - c0, gw};
  const int n = (r1 - r0) * (c1 - c0);
  float sum;
  if (n <= 128) {
    sum = pairwise_block(wn, 0, n);
  } else {  // one level of numpy's recursion covers n <= 256
    int n2 = n / 2;
    n2 -= n2 % 8;
    sum = __fadd_rn(pairwise_block(wn, 0, n2), pairwise_block(wn, n2, n - n2));
  }
  const int img = box_img[p];
  S[(long long)p * kmax + (k - peak_img_start[img])] = (double)sum;  // kp_score (== 1) * score, :425
  colhit[k] = 1;
}

// one CTA per (image, joint type); thread = box of the image
__global__ void __launch_bounds__(128) prn_assign_kernel(const double* __restrict__ peak_xy, const int32_t* __restrict__ peak_img_start,
                                                         const int32_t* __restrict__ joint_start /* [B][18] */,
                                                         const int32_t* __restrict__ box_img_start, int kmax,
                                                         const double* __restrict__ S, const int32_t* __restrict__ colhit,
                                                         double* __restrict__ bbox_keypoints) {
  const int img = blockIdx.x / NJ, jt = blockIdx.x - img * NJ;
  const int bs = box_img_start[img], be = box_img_start[img + 1];
  const int k0 = peak_img_start[img];
  const int ks = joint_start[img * (NJ + 1) + jt] - k0, ke = joint_start[img * (NJ + 1) + jt + 1] - k0;  // local peak range of this joint
  const int32_t* ch = colhit + k0;
  for (int bbox = bs + threadIdx.x; bbox < be; bbox += blockDim.x) {
    const double* row = S + (long long)bbox * kmax;
    // walk the row in descending score, ties by ascending index (stable argsort of -row, :455)
    double prev_s = 0.0;
    int prev_i = -1;
    bool first = true;
    int chosen = -1;
    while (true) {
      double best = 0.0;
      int bi = -1;
      for (int k = ks; k < ke; ++k) {
        if (!ch[k]) continue;  // not a table column (:438)
        const double s = row[k];
        if (!(s > 0.0)) continue;  // :456/:458 -- non-positive scores end the walk
        if (!first && !(s < prev_s || (s == prev_s && k > prev_i))) continue;
        if (bi < 0 || s > best) { best = s; bi = k; }
      }
      if (bi < 0) break;
      first = false;
      prev_s = best;
      prev_i = bi;
      // column0 = box with the largest score for this peak (first among ties), :460
      int col0 = bs;
      double cbest = S[(long long)bs * kmax + bi];
      for (int b = bs + 1; b < be; ++b) {
        const double s = S[(long long)b * kmax + bi];
        if (s > cbest) { cbest = s; col0 = b; }
      }
      if (col0 == bbox) { chosen = bi; break; }
      // else: is this peak the lowest-scoring column of that box (first among ties)?  :468-469
      const double* row0 = S + (long long)col0 * kmax;
      int wi = -1;
      double wv = 0.0;
      for (int k = ks; k < ke; ++k) {
        if (!ch[k]) continue;
        if (wi < 0 || row0[k] < wv) { wv = row0[k]; wi = k; }
      }
      if (wi == bi) { chosen = bi; break; }
    }
    if (chosen >= 0) {
      double* o = bbox_keypoints + ((long long)bbox * NJ + jt) * 3;
      o[0] = peak_xy[2 * (k0 + chosen)];
      o[1] = peak_xy[2 * (k0 + chosen) + 1];
      o[2] = 1.0;
    }
  }
}

__global__ void __launch_bounds__(PRN_THREADS) prn_fallback_kernel(const float* __restrict__ output, const double* __restrict__ boxes,
                                                                   const int32_t* __restrict__ box_img,
                                                                   const int32_t* __restrict__ joint_start, const int32_t* __restrict__ colhit,
                                                                   const int32_t* __restrict__ plane_has, int gh, int gw,
                                                                   double* __restrict__ bbox_keypoints) {
  const int plane = blockIdx.x, p = plane / NJ, t = plane - p * NJ;
  if (plane_has[plane]) return;  // :478-479: only (box, joint) planes without any peak
  const int img = box_img[p];
  // the branch runs iff some joint type of the image has no scored peak at all (:471)
  __shared__ int joint_any[NJ];
  __shared__ float wbest[PRN_THREADS / 32];
  __shared__ int wbesti[PRN_THREADS / 32];
  if (threadIdx.x < NJ) joint_any[threadIdx.x] = 0;
  __syncthreads();
  const int* js = joint_start + img * (NJ + 1);
  for (int k = js[0] + threadIdx.x; k < js[NJ]; k += PRN_THREADS) {
    if (colhit[k]) {
      int j = 0;
      while (k >= js[j + 1]) ++j;
      joint_any[j] = 1;
    }
  }
  __syncthreads();
  bool empty_joint = false;
  for (int j = 0; j < NJ; ++j) empty_joint |= (joint_any[j] == 0);
  if (!empty_joint) return;
  // first maximum of the plane in row-major order (:480 np.argwhere(out == out.max())[0])
  const float* o = output + (long long)p * gh * gw * NJ + t;
  float best = -INFINITY;
  int besti = 0x7fffffff;
  for (int i = threadIdx.x; i < gh * gw; i += PRN_THREADS) {
    const float v = o[(long long)i * NJ];
    if (v > best) { best = v; besti = i; }
  }
  for (int s = 16; s; s >>= 1) {
    const float ov = __shfl_down_sync(0xffffffffu, best, s);
    const int oi = __shfl_down_sync(0xffffffffu, besti, s);
    if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
  }
  if ((threadIdx.x & 31) == 0) { wbest[threadIdx.x >> 5] = best; wbesti[threadIdx.x >> 5] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < PRN_THREADS / 32; ++w)
      if (wbest[w] > best || (wbest[w] == best && wbesti[w] < besti)) { best = wbest[w]; besti = wbesti[w]; }
    const double b0 = boxes[p * 4 + 0], b1 = boxes[p * 4 + 1], b2 = boxes[p * 4 + 2], b3 = boxes[p * 4 + 3];
    const double x_scale = __ddiv_rn((double)gw, ceil(b2)), y_scale = __ddiv_rn((double)gh, ceil(b3));
    const int my = besti / gw, mx = besti - my * gw;
    double* r = bbox_keypoints + (long long)plane * 3;
    r[0] = __dadd_rn(__ddiv_rn((double)mx, x_scale), b0);  // :481-482
    r[1] = __dadd_rn(__ddiv_rn((double)my, y_scale), b1);
    r[2] = 0.0;
  }
}

}  // namespace

extern "C" size_t mpn_prn_workspace_bytes(int P, int n_peaks, int kmax) {
  if (P < 0 || n_peaks < 0 || kmax < 0) return 0;
  // S [P, kmax] f64 | colhit [n_peaks] i32 | plane_has [P*17] i32
  return (size_t)P * (size_t)(kmax > 0 ? kmax : 1) * sizeof(double) + (size_t)(n_peaks + 1) * sizeof(int32_t) + (size_t)P * NJ * sizeof(int32_t) + 64;
}

static inline size_t prn_align8(size_t v) { return (v + 7) & ~(size_t)7; }

extern "C" int mpn_prn_build_inputs(const double* peak_xy, const int32_t* peak_type, const int32_t* peak_img_start, int n_peaks,
                                    const double* boxes_xywh, const int32_t* box_img, int P, int gh, int gw, double in_thres,
                                    const double* gauss_w_host, int32_t* owner, float* inp, void* workspace, size_t workspace_bytes,
                                    int kmax, void* stream) {
  MPN_CHECK_ARG(P > 0 && n_peaks >= 0 && gh > 0 && gw > 0 && gh * gw <= 8192, "mpn_prn_build_inputs: bad sizes");
  MPN_CHECK_ARG(peak_img_start && boxes_xywh && box_img && gauss_w_host && owner && inp && workspace, "mpn_prn_build_inputs: null pointer");
  MPN_CHECK_ARG(n_peaks == 0 || (peak_xy && peak_type), "mpn_prn_build_inputs: null peaks");
  MPN_CHECK_ARG(in_thres >= 0.0 && in_thres < 1.0, "mpn_prn_build_inputs: in_thres must be in [0, 1)");
  MPN_CHECK_ARG(workspace_bytes >= mpn_prn_workspace_bytes(P, n_peaks, kmax), "mpn_prn_build_inputs: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cells = (size_t)P * NJ * gh * gw;
  MPN_CUDA_OK(cudaMemsetAsync(owner, 0xFF, cells * sizeof(int32_t), st));
  MPN_CUDA_OK(cudaMemsetAsync(inp, 0, cells * sizeof(float), st));
  if (n_peaks > 0) {
    prn_scatter_kernel<<<P, PRN_THREADS, 0, st>>>(peak_xy, peak_type, peak_img_start, boxes_xywh, box_img, gh, gw, in_thres, owner);
    MPN_LAUNCH_OK();
  }
  GaussW g;
  for (int i = 0; i < 5; ++i) g.w[i] = gauss_w_host[i];
  int32_t* plane_has = (int32_t*)((char*)workspace + prn_align8((size_t)P * (size_t)(kmax > 0 ? kmax : 1) * sizeof(double)) +
                                  prn_align8((size_t)(n_peaks + 1) * sizeof(int32_t)));
  const size_t blur_smem = 2 * (size_t)gh * gw * sizeof(double);
  if (blur_smem > 48 * 1024) MPN_CUDA_OK(cudaFuncSetAttribute(prn_blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blur_smem));
  prn_blur_kernel<<<P * NJ, PRN_THREADS, blur_smem, st>>>(owner, gh, gw, g, inp, plane_has);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_prn_assign(const double* peak_xy, const int32_t* peak_img_start, const int32_t* joint_start, int n_peaks,
                              const double* boxes_xywh, const int32_t* box_img, const int32_t* box_img_start, int P, int B, int gh, int gw,
                              const int32_t* owner, const float* output, double* bbox_keypoints, void* workspace,
                              size_t workspace_bytes, int kmax, void* stream) {
  MPN_CHECK_ARG(P > 0 && B > 0 && n_peaks >= 0 && gh > 0 && gw > 0 && kmax >= 0, "mpn_prn_assign: bad sizes");
  MPN_CHECK_ARG(peak_img_start && joint_start && boxes_xywh && box_img && box_img_start && owner && output && bbox_keypoints && workspace,
                "mpn_prn_assign: null pointer");
  MPN_CHECK_ARG(workspace_bytes >= mpn_prn_workspace_bytes(P, n_peaks, kmax), "mpn_prn_assign: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t s_bytes = prn_align8((size_t)P * (size_t)(kmax > 0 ? kmax : 1) * sizeof(double));
  double* S = (double*)workspace;
  int32_t* colhit = (int32_t*)((char*)workspace + s_bytes);
  int32_t* plane_has = (int32_t*)((char*)workspace + s_bytes + prn_align8((size_t)(n_peaks + 1) * sizeof(int32_t)));  // written by mpn_prn_build_inputs
  MPN_CUDA_OK(cudaMemsetAsync(S, 0, s_bytes + prn_align8((size_t)(n_peaks + 1) * sizeof(int32_t)), st));
  MPN_CUDA_OK(cudaMemsetAsync(bbox_keypoints, 0, (size_t)P * NJ * 3 * sizeof(double), st));
  const long long cells = (long long)P * NJ * gh * gw;
  if (n_peaks > 0) {
    prn_score_kernel<<<mpn_divup(cells, PRN_THREADS), PRN_THREADS, 0, st>>>(owner, output, box_img, peak_img_start, P, gh, gw, kmax, S, colhit);
    MPN_LAUNCH_OK();
    prn_assign_kernel<<<B * NJ, 128, 0, st>>>(peak_xy, peak_img_start, joint_start, box_img_start, kmax, S, colhit, bbox_keypoints);
    MPN_LAUNCH_OK();
  }
  prn_fallback_kernel<<<P * NJ, PRN_THREADS, 0, st>>>(output, boxes_xywh, box_img, joint_start, colhit, plane_has, gh, gw, bbox_keypoints);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
