// Heat-map peak extraction on the device (the step right after the path: network/joint_utils.py:19-32, 61-152 as called
// by evaluate/tester.py:215-221).  Replaces: 44 MB/step heat-map D2H + scipy maximum_filter + a Python loop with one
// cv2.resize per peak.
//
//   peak_count_kernel : one CTA per (image, joint) plane: number of pixels with v > thre1 and v >= its 4 neighbours
//   peak_emit_kernel  : same scan in row-major chunks with an ordered block compaction (= np.nonzero order); one warp
//                       per peak refines it: <=5x5 patch -> x factor bicubic (A = -0.75, half-pixel centres, clamped
//                       taps; horizontal then vertical pass, unfused float32 ops in the order of oracle/peaks_oracle.py)
//                       -> first arg-max -> row (x, y, score, id, joint) at the peak's global rank.
#include <stdint.h>

#include "mpn_common.cuh"

namespace {

constexpr int PK_THREADS = 256;
constexpr int PK_WARPS = PK_THREADS / 32;
constexpr int MAX_FACTOR = 8;
constexpr int MAX_UP = 5 * MAX_FACTOR;  // upsampled patch side

__device__ __forceinline__ bool is_peak(const float* __restrict__ m, int H, int W, int y, int x, float thre1) {
  const float v = m[y * W + x];
  if (!(v > thre1)) return false;
  // maximum_filter(mode='reflect') repeats the border pixel: only in-range neighbours can beat v
  if (y > 0 && m[(y - 1) * W + x] > v) return false;
  if (y + 1 < H && m[(y + 1) * W + x] > v) return false;
  if (x > 0 && m[y * W + x - 1] > v) return false;
  if (x + 1 < W && m[y * W + x + 1] > v) return false;
  return true;
}

__global__ void __launch_bounds__(PK_THREADS) peak_count_kernel(const float* __restrict__ heat, int C, int H, int W,
                                                               long long image_stride, float thre1, int32_t* __restrict__ counts) {
  const int plane = blockIdx.x, b = plane / C, c = plane - b * C;
  const float* m = heat + (long long)b * image_stride + (long long)c * H * W;
  int n = 0;
  for (int i = threadIdx.x; i < H * W; i += PK_THREADS) n += is_peak(m, H, W, i / W, i - (i / W) * W, thre1) ? 1 : 0;
  __shared__ int wsum[PK_WARPS];
  for (int o = 16; o; o >>= 1) n += __shfl_down_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < PK_WARPS; ++w) t += wsum[w];
    counts[plane] = t;
  }
}

// OpenCV interpolateCubic in unfused float32 arithmetic
__device__ __forceinline__ void cubic_coeffs(float x, float* c) {
  const float A = -0.75f;
  const float xp1 = __fadd_rn(x, 1.f), omx = __fsub_rn(1.f, x);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), __fmul_rn(5.f, A)), xp1), __fmul_rn(8.f, A)), xp1),
                   __fmul_rn(4.f, A));
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f)), x), x), 1.f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), omx), __fadd_rn(A, 3.f)), omx), omx), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

__device__ __forceinline__ float dot4(const float* s, const float* c) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s[0], c[0]), __fmul_rn(s[1], c[1])), __fmul_rn(s[2], c[2])), __fmul_rn(s[3], c[3]));
}

__global__ void __launch_bounds__(PK_THREADS) peak_emit_kernel(const float* __restrict__ heat, int C, int H, int W,
                                                              long long image_stride, float thre1, int factor,
                                                              const int32_t* __restrict__ counts, float* __restrict__ peaks,
                                                              int max_peaks, int32_t* __restrict__ count_out) {
  const int plane = blockIdx.x, b = plane / C, c = plane - b * C;
  const float* m = heat + (long long)b * image_stride + (long long)c * H * W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ int wsum[PK_WARPS];
  __shared__ int list[PK_THREADS];            // pixel index of the chunk's peaks, in order
  __shared__ float coef[MAX_UP][4];           // per destination index of an upsampled patch side
  __shared__ int tap0[MAX_UP];                // first source tap (before clamping)
  __shared__ float hr[PK_WARPS][5][MAX_UP];   // per-warp horizontal pass
  int base = 0;                               // rows of this image written by earlier joint types (the reference's id counter)
  for (int j = 0; j < c; ++j) base += counts[b * C + j];
  if (c == C - 1 && threadIdx.x == 0) count_out[b] = base + counts[plane];
  if (threadIdx.x < 5 * factor) {
    // destination index d samples the source at fx = (d + 0.5)/factor - 0.5 (double expression, then float: resize.cpp)
    const float fx = (float)__dsub_rn(__dmul_rn((double)threadIdx.x + 0.5, 1.0 / (double)factor), 0.5);  // no fma contraction
    const float fl = floorf(fx);
    tap0[threadIdx.x] = (int)fl - 1;
    cubic_coeffs(__fsub_rn(fx, fl), coef[threadIdx.x]);
  }
  __syncthreads();
  int running = 0;
  for (int start = 0; start < H * W; start += PK_THREADS) {
    const int i = start + threadIdx.x;
    const bool pk = i < H * W && is_peak(m, H, W, i / W, i - (i / W) * W, thre1);
    const unsigned bal = __ballot_sync(0xffffffffu, pk);
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < PK_WARPS; ++w) {
      if (w < warp) before += wsum[w];
      total += wsum[w];
    }
    if (pk) list[before + __popc(bal & ((1u << lane) - 1u))] = i;
    __syncthreads();
    for (int e = warp; e < total; e += PK_WARPS) {
      const int pi = list[e], py = pi / W, px = pi - py * W;
      const int x_min = max(0, px - 2), y_min = max(0, py - 2);
      const int x_max = min(W - 1, px + 2), y_max = min(H - 1, py + 2);
      const int pw = x_max - x_min + 1, ph = y_max - y_min + 1, uw = pw * factor, uh = ph * factor;
      // horizontal pass: hr[y][d] for the patch rows
      for (int t = lane; t < ph * uw; t += 32) {
        const int y = t / uw, d = t - y * uw;
        float s[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = m[(y_min + y) * W + x_min + min(max(tap0[d] + j, 0), pw - 1)];
        hr[warp][y][d] = dot4(s, coef[d]);
      }
      __syncwarp();
      // vertical pass + first arg-max in row-major order
      float best = -INFINITY;
      int besti = 0x7fffffff;
      for (int t = lane; t < uh * uw; t += 32) {
        const int ey = t / uw, d = t - ey * uw;
        float s[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] = hr[warp][min(max(tap0[ey] + k, 0), ph - 1)][d];
        const float v = dot4(s, coef[ey]);
        if (v > best) { best = v; besti = t; }  // increasing t per lane: keeps the first maximum
      }
      for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_down_sync(0xffffffffu, best, o);
        const int oi = __shfl_down_sync(0xffffffffu, besti, o);
        if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
      }
      if (lane == 0) {
        const int rank = base + running + e;
        if (rank < max_peaks) {
          float* row = peaks + ((long long)b * max_peaks + rank) * 5;
          row[0] = (float)(x_min * factor + besti % uw);
          row[1] = (float)(y_min * factor + besti / uw);
          row[2] = best;
          row[3] = (float)rank;
          row[4] = (float)c;
        }
      }
      __syncwarp();
    }
    running += total;
    __syncthreads();
  }
}

}  // namespace

extern "C" size_t mpn_heatmap_peaks_workspace_bytes(int B, int C) { return (size_t)(B > 0 ? B : 0) * (size_t)(C > 0 ? C : 0) * sizeof(int32_t); }

extern "C" int mpn_heatmap_peaks(const float* heat, int B, int C, int H, int W, long long image_stride, float thre1, int factor,
                                 float* peaks, int max_peaks, int32_t* count, void* workspace, size_t workspace_bytes, void* stream) {
  MPN_CHECK_ARG(heat && peaks && count && workspace && B > 0 && C > 0 && H > 0 && W > 0 && max_peaks > 0, "mpn_heatmap_peaks: bad argument");
  MPN_CHECK_ARG(factor >= 1 && factor <= MAX_FACTOR, "mpn_heatmap_peaks: upsampling factor must be an integer in [1, %d]", MAX_FACTOR);
  MPN_CHECK_ARG(image_stride >= (long long)C * H * W, "mpn_heatmap_peaks: image stride smaller than C*H*W");
  MPN_CHECK_ARG(workspace_bytes >= mpn_heatmap_peaks_workspace_bytes(B, C), "mpn_heatmap_peaks: workspace too small");
  MPN_CHECK_ARG((long long)H * W < (1LL << 30), "mpn_heatmap_peaks: plane too large");
  cudaStream_t st = (cudaStream_t)stream;
  int32_t* counts = (int32_t*)workspace;
  peak_count_kernel<<<B * C, PK_THREADS, 0, st>>>(heat, C, H, W, image_stride, thre1, counts);
  MPN_LAUNCH_OK();
  peak_emit_kernel<<<B * C, PK_THREADS, 0, st>>>(heat, C, H, W, image_stride, thre1, factor, counts, peaks, max_peaks, count);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
