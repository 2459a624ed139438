// Multi-scale test-time augmentation, device side (evaluate/tester.py:264-331):
//   heatmap = cv2.resize(heatmaps[:h/4, :w/4], None, fx=4, fy=4, INTER_CUBIC)          tester.py:298-299
//   heatmap = cv2.resize(heatmap[:real_h, :real_w], (W0, H0), INTER_CUBIC)              tester.py:300-302
//   heatmap_avg = heatmap_avg + heatmap / len(multiplier)                               tester.py:304
//   averaged = (normal + flipped[:, ::-1, :][:, :, swap_heat]) / 2.                     tester.py:316-331
// cv2.resize INTER_CUBIC on float32 planes (imgproc/resize.cpp): per destination index d the source coordinate is
// fx = float((d + 0.5) * scale - 0.5) (double expression), taps floor(fx) - 1 .. + 2 clamped to the plane (border
// replicate by index clamping), Keys weights A = -0.75 in float32; a horizontal pass rounds to float32 rows, then the vertical
// pass combines four of them.  OpenCV's own vector / scalar code paths differ in FMA use, so the contract here is a tolerance
// (1e-5 of the plane's max, tests/test_gpu_tta.py; measured 3.9e-6), not bit-exactness.
#include "mpn_common.cuh"

namespace {

__device__ __forceinline__ void cubic_w(float x, float* c) {  // interpolateCubic (resize.cpp), unfused float32
  const float A = -0.75f;
  const float xp1 = __fadd_rn(x, 1.f), omx = __fsub_rn(1.f, x);
  c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), __fmul_rn(5.f, A)), xp1), __fmul_rn(8.f, A)), xp1),
                   __fmul_rn(4.f, A));
  c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), x), __fadd_rn(A, 3.f)), x), x), 1.f);
  c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(A, 2.f), omx), __fadd_rn(A, 3.f)), omx), omx), 1.f);
  c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
}

// tap table of one axis: for each destination index the first source tap (floor(fx) - 1) and the four weights
__global__ void cubic_table_kernel(int dn, double scale, int* __restrict__ tap0, float* __restrict__ coef) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dn) return;
  const float fx = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
  const float fl = floorf(fx);
  tap0[d] = (int)fl - 1;
  cubic_w(__fsub_rn(fx, fl), coef + 4 * d);
}

// One thread per destination pixel of one plane.  out_mode 0: dst fp32 = value; 1: dst fp64 += (double)(value / div) with the
// destination plane chosen by plane_map (channel swap of the flipped pass) and the column mirrored when mirror != 0.
__global__ void __launch_bounds__(256) resize_cubic_kernel(const float* __restrict__ src, long long src_plane, int src_pitch, int sh,
                                                           int sw, void* __restrict__ dst, long long dst_plane, int dst_pitch, int dh,
                                                           int dw, const int* __restrict__ xt, const float* __restrict__ xc,
                                                           const int* __restrict__ yt, const float* __restrict__ yc, int out_mode,
                                                           float div, int mirror, const int* __restrict__ plane_map) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, p = blockIdx.z;
  if (x >= dw) return;
  const float* s = src + (long long)p * src_plane;
  const int x0 = xt[x], y0 = yt[y];
  float a[4], b[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { a[j] = xc[4 * x + j]; b[j] = yc[4 * y + j]; }
  int xs[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) xs[j] = min(max(x0 + j, 0), sw - 1);
  float rows[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float* r = s + (long long)min(max(y0 + k, 0), sh - 1) * src_pitch;
    // HResizeCubic: S[sx-1]*a0 + S[sx]*a1 + S[sx+1]*a2 + S[sx+2]*a3, left to right
    rows[k] = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r[xs[0]], a[0]), __fmul_rn(r[xs[1]], a[1])), __fmul_rn(r[xs[2]], a[2])), __fmul_rn(r[xs[3]], a[3]));
  }
  // VResizeCubic (vector path): fma(S0, b0, fma(S1, b1, fma(S2, b2, S3 * b3)))
  const float v = fmaf(rows[0], b[0], fmaf(rows[1], b[1], fmaf(rows[2], b[2], __fmul_rn(rows[3], b[3]))));
  if (out_mode == 0) {
    ((float*)dst)[(long long)p * dst_plane + (long long)y * dst_pitch + x] = v;
  } else {
    const int dp = plane_map ? plane_map[p] : p;
    const int dx = mirror ? dw - 1 - x : x;
    double* o = (double*)dst + (long long)dp * dst_plane + (long long)y * dst_pitch + dx;
    *o = __dadd_rn(*o, (double)__fdiv_rn(v, div));   // heatmap_avg + heatmap / len(multiplier): float32 quotient, float64 sum
  }
}

// out[c][y][x] = (a[c][y][x] + b[c][y][x]) / 2 in float64 (b already mirrored / channel-swapped), also as float32
__global__ void tta_combine_kernel(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, float* __restrict__ out32,
                                   long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double v = b ? __ddiv_rn(__dadd_rn(a[i], b[i]), 2.0) : a[i];
    if (out) out[i] = v;
    if (out32) out32[i] = (float)v;
  }
}

}  // namespace

extern "C" size_t mpn_resize_cubic_workspace_bytes(int dh, int dw) {
  if (dh <= 0 || dw <= 0) return 0;
  return (size_t)(dh + dw) * (sizeof(int) + 4 * sizeof(float)) + 64;
}

extern "C" int mpn_resize_cubic(const float* src, long long src_plane, int src_pitch, int sh, int sw, void* dst, long long dst_plane,
                                int dst_pitch, int dh, int dw, int planes, double scale_x, double scale_y, int out_mode, float div,
                                int mirror, const int* plane_map, void* workspace, size_t workspace_bytes, void* stream) {
  MPN_CHECK_ARG(src && dst && workspace && planes > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, "mpn_resize_cubic: bad argument");
  MPN_CHECK_ARG(src_pitch >= sw && dst_pitch >= dw && scale_x > 0 && scale_y > 0, "mpn_resize_cubic: bad pitch / scale");
  MPN_CHECK_ARG(out_mode == 0 || (out_mode == 1 && div != 0.f), "mpn_resize_cubic: out_mode 0 (fp32 store) or 1 (fp64 accumulate, div != 0)");
  MPN_CHECK_ARG(workspace_bytes >= mpn_resize_cubic_workspace_bytes(dh, dw), "mpn_resize_cubic: workspace too small");
  MPN_CHECK_ARG(planes <= 65535 && dh <= 65535, "mpn_resize_cubic: too many planes / rows for one launch");
  cudaStream_t st = (cudaStream_t)stream;
  int* xt = (int*)workspace;
  int* yt = xt + dw;
  float* xc = (float*)(yt + dh);
  float* yc = xc + 4 * (size_t)dw;
  cubic_table_kernel<<<mpn_divup(dw, 128), 128, 0, st>>>(dw, scale_x, xt, xc);
  MPN_LAUNCH_OK();
  cubic_table_kernel<<<mpn_divup(dh, 128), 128, 0, st>>>(dh, scale_y, yt, yc);
  MPN_LAUNCH_OK();
  dim3 grid(mpn_divup(dw, 256), dh, planes);
  resize_cubic_kernel<<<grid, 256, 0, st>>>(src, src_plane, src_pitch, sh, sw, dst, dst_plane, dst_pitch, dh, dw, xt, xc, yt, yc, out_mode,
                                            div, mirror, plane_map);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" int mpn_tta_combine(const double* normal, const double* flipped, double* out, float* out32, long long n, void* stream) {
  MPN_CHECK_ARG(normal && (out || out32) && n > 0, "mpn_tta_combine: bad argument");
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  tta_combine_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(normal, flipped, out, out32, n);
  MPN_LAUNCH_OK();
  return MPN_OK;
}
