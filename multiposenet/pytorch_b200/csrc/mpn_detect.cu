// Detection post-process on the device: anchors, box decode + clip, score filter, stable descending
// radix sort, bit-mask IoU and the greedy reduction -- the reference's
//   Anchors.forward (network/anchors.py:21-37)  [float64 host math, cached by the caller]
//   BBoxTransform / ClipBoxes (network/utils.py:19-61)
//   score > 0.05 boolean indexing (network/posenet.py:269-279)
//   pth_nms -> gpu_nms -> nms_kernel + the serial host loop (lib/nms/pth_nms.py:25-44,
//     src/nms_cuda.c:17-67, src/cuda/nms_kernel.cu:16-70)
// without the reference's device->host mask copy and host loop.  Integer results are bit-exact with
// oracle/nms_oracle.c: the IoU arithmetic uses explicit round-to-nearest intrinsics in the reference's
// operation order (no FMA contraction is possible in devIoU either).
#include <cub/cub.cuh>
#include <math.h>
#include <stdlib.h>

#include "mpn_common.cuh"

namespace {

constexpr int kLevels[5] = {3, 4, 5, 6, 7};

// ---------------------------------------------------------------------------------------------
__global__ void decode_clip_kernel(const float* __restrict__ anchors, const float* __restrict__ reg, float* __restrict__ boxes,
                                   int B, int A, float Hf, float Wf, int clip) {
  long long total = (long long)B * A;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int a = (int)(i % A);
    float4 an = reinterpret_cast<const float4*>(anchors)[a];
    float4 dl = reinterpret_cast<const float4*>(reg)[i];
    // utils.py:21-29 -- every product and sum is a separate fp32 op in torch, so no FMA here
    float w = __fsub_rn(an.z, an.x), h = __fsub_rn(an.w, an.y);
    float cx = __fadd_rn(an.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(an.y, __fmul_rn(0.5f, h));
    float dx = __fadd_rn(__fmul_rn(dl.x, 0.1f), 0.f), dy = __fadd_rn(__fmul_rn(dl.y, 0.1f), 0.f);
    float dw = __fadd_rn(__fmul_rn(dl.z, 0.2f), 0.f), dh = __fadd_rn(__fmul_rn(dl.w, 0.2f), 0.f);
    float pcx = __fadd_rn(cx, __fmul_rn(dx, w)), pcy = __fadd_rn(cy, __fmul_rn(dy, h));
    float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
    float4 o;
    o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
    o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
    o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
    o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
    if (clip) {
      o.x = fmaxf(o.x, 0.f);  // utils.py:56-57
      o.y = fmaxf(o.y, 0.f);
      o.z = fminf(o.z, Wf);   // utils.py:59-60
      o.w = fminf(o.w, Hf);
    }
    reinterpret_cast<float4*>(boxes)[i] = o;
  }
}

// One CTA per image walks the A scores in order; survivors keep their relative order
// (boolean indexing semantics).  rank_in = position in the filtered set, key = score.
__global__ void __launch_bounds__(1024) filter_compact_kernel(const float* __restrict__ cls, int A, float thr, int max_cand,
                                                             int32_t* __restrict__ cand_idx, int32_t* __restrict__ cand_cnt,
                                                             float* __restrict__ keys, int32_t* __restrict__ ranks) {
  __shared__ int warp_sums[32];
  __shared__ int running;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) running = 0;
  __syncthreads();
  const float* s = cls + (long long)b * A;
  for (int base = 0; base < A; base += 1024) {
    int a = base + tid;
    float v = a < A ? s[a] : 0.f;
    bool keep = a < A && v > thr;
    unsigned bal = __ballot_sync(0xffffffffu, keep);
    int within = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) warp_sums[wid] = __popc(bal);
    __syncthreads();
    int before = 0;
    int tot = 0;
#pragma unroll
    for (int w2 = 0; w2 < 32; ++w2) {
      int c = warp_sums[w2];
      before += (w2 < wid) ? c : 0;
      tot += c;
    }
    int pos = running + before + within;
    if (keep && pos < max_cand) {
      cand_idx[(long long)b * max_cand + pos] = a;
      keys[(long long)b * max_cand + pos] = v;
      ranks[(long long)b * max_cand + pos] = pos;
    }
    __syncthreads();
    if (tid == 0) running += tot;
    __syncthreads();
  }
  if (tid == 0) cand_cnt[b] = running;
}

__global__ void iota_dets_kernel(const float* __restrict__ dets, int n, float* keys, int32_t* ranks, int32_t* cand_idx,
                                 int32_t* cand_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    keys[i] = dets[(long long)i * 5 + 4];
    ranks[i] = i;
    cand_idx[i] = i;
  }
  if (i == 0) cand_cnt[0] = n;
}

__global__ void segments_kernel(const int32_t* cand_cnt, int B, int max_cand, int* seg_begin, int* seg_end) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    int c = cand_cnt[b];
    seg_begin[b] = b * max_cand;
    seg_end[b] = b * max_cand + (c < max_cand ? c : max_cand);
  }
}

// sdets[b][j] = (box of the j-th best candidate, score)
__global__ void gather_sorted_kernel(const float* __restrict__ boxes, int box_stride, long long box_image_stride,
                                     const int32_t* __restrict__ cand_idx, const int32_t* __restrict__ cand_cnt,
                                     const float* __restrict__ keys_sorted, const int32_t* __restrict__ ranks_sorted,
                                     int max_cand, float* __restrict__ sdets) {
  const int b = blockIdx.y;
  int n = cand_cnt[b];
  n = n < max_cand ? n : max_cand;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    long long o = (long long)b * max_cand + j;
    int a = cand_idx[(long long)b * max_cand + ranks_sorted[o]];
    const float* bx = boxes + (long long)b * box_image_stride + (long long)a * box_stride;
    float* d = sdets + o * 5;
    d[0] = bx[0]; d[1] = bx[1]; d[2] = bx[2]; d[3] = bx[3];
    d[4] = keys_sorted[o];
  }
}

// grid (upper-triangle block pair, 1, image): only col_block >= row_block exists (the reduction never
// reads the lower triangle, nms_cuda.c:52).  64 threads: thread t owns row box row_block*64+t.
__global__ void __launch_bounds__(64) nms_mask_kernel(const float* __restrict__ sdets, const int32_t* __restrict__ cand_cnt,
                                                     int n_fixed, int max_cand, int cb_stride, int cb_grid, float thr,
                                                     int ge, unsigned long long* __restrict__ mask) {
  const int b = blockIdx.z;
  int n = cand_cnt ? cand_cnt[b] : n_fixed;
  n = n < max_cand ? n : max_cand;
  // blockIdx.x enumerates the upper triangle of the cb_grid x cb_grid block matrix row by row
  int row_start = 0, t = blockIdx.x;
  while (t >= cb_grid - row_start) {
    t -= cb_grid - row_start;
    ++row_start;
  }
  const int col_start = row_start + t;
  if (row_start * 64 >= n || col_start * 64 >= n) return;
  const int row_size = min(n - row_start * 64, 64), col_size = min(n - col_start * 64, 64);
  const float* dets = sdets + (long long)b * max_cand * 5;
  __shared__ float block_boxes[64 * 5];  // x1 y1 x2 y2 area  (area = (x2-x1+1)*(y2-y1+1), nms_kernel.cu:22-23)
  if (threadIdx.x < col_size) {
    float bx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bx[k] = dets[(long long)(64 * col_start + threadIdx.x) * 5 + k];
#pragma unroll
    for (int k = 0; k < 4; ++k) block_boxes[threadIdx.x * 5 + k] = bx[k];
    block_boxes[threadIdx.x * 5 + 4] = __fmul_rn(__fadd_rn(__fsub_rn(bx[2], bx[0]), 1.f), __fadd_rn(__fsub_rn(bx[3], bx[1]), 1.f));
  }
  __syncthreads();
  if (threadIdx.x < row_size) {
    const int cur = 64 * row_start + threadIdx.x;
    float cb[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cb[k] = dets[(long long)cur * 5 + k];
    const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(cb[2], cb[0]), 1.f), __fadd_rn(__fsub_rn(cb[3], cb[1]), 1.f));
    unsigned long long t = 0;
    int start = (row_start == col_start) ? threadIdx.x + 1 : 0;
    // The shortcuts assume a positive finite threshold.  (1) disjoint boxes: interS == 0 exactly -> IoU is 0 (or NaN for a
    // 0/0 union), never above the threshold.  (2) interS vs thr*union with a 1e-6 relative guard band (8 ulps, the
    // products carry < 1 ulp of rounding): outside the band the correctly rounded quotient of nms_kernel.cu:24 is
    // certainly above / below thr, so the division is only evaluated inside the band.
    const bool fast_ok = thr > 1e-3f && thr < 1e3f;
    for (int i = start; i < col_size; ++i) {
      const float* bb = block_boxes + i * 5;
      const float w = __fadd_rn(__fsub_rn(fminf(cb[2], bb[2]), fmaxf(cb[0], bb[0])), 1.f);
      const float h = __fadd_rn(__fsub_rn(fminf(cb[3], bb[3]), fmaxf(cb[1], bb[1])), 1.f);
      if (fast_ok && (!(w > 0.f) || !(h > 0.f))) continue;
      const float interS = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
      const float uni = __fsub_rn(__fadd_rn(Sa, bb[4]), interS);
      bool over;
      const float pth = __fmul_rn(thr, uni);
      if (fast_ok && uni > 0.f && interS > __fmul_rn(pth, 1.000001f)) {
        over = true;
      } else if (fast_ok && uni > 0.f && interS < __fmul_rn(pth, 0.999999f)) {
        over = false;
      } else {
        const float v = __fdiv_rn(interS, uni);
        over = ge ? (v >= thr) : (v > thr);
      }
      if (over) t |= 1ULL << i;
    }
    mask[((long long)b * max_cand + cur) * cb_stride + col_start] = t;
  }
}

// The IoU test of one (row box, column box) pair, in the reference's operation order (nms_kernel.cu:16-24) with explicit
// round-to-nearest intrinsics.  Shortcuts (positive finite threshold only): (1) disjoint boxes: interS == 0 exactly -> IoU is
// 0 (or NaN for a 0/0 union), never above the threshold; (2) interS vs thr*union with a 1e-6 relative guard band (8 ulps, the
// products carry < 1 ulp of rounding): outside the band the correctly rounded quotient is certainly above / below thr, so the
// division is only evaluated inside the band.
__device__ __forceinline__ bool nms_pair_over(const float4& a, float Sa, const float4& bb, float Sb, float thr, int ge, bool fast_ok) {
  const float w = __fadd_rn(__fsub_rn(fminf(a.z, bb.z), fmaxf(a.x, bb.x)), 1.f);
  const float h = __fadd_rn(__fsub_rn(fminf(a.w, bb.w), fmaxf(a.y, bb.y)), 1.f);
  if (fast_ok && (!(w > 0.f) || !(h > 0.f))) return false;
  const float interS = __fmul_rn(fmaxf(w, 0.f), fmaxf(h, 0.f));
  const float uni = __fsub_rn(__fadd_rn(Sa, Sb), interS);
  const float pth = __fmul_rn(thr, uni);
  if (fast_ok && uni > 0.f && interS > __fmul_rn(pth, 1.000001f)) return true;
  if (fast_ok && uni > 0.f && interS < __fmul_rn(pth, 0.999999f)) return false;
  const float v = __fdiv_rn(interS, uni);
  return ge ? (v >= thr) : (v > thr);
}

// Second-generation mask kernel.  grid (row block, column group, image); a CTA of 128 threads owns one 64-box row block and
// up to NMS_CG consecutive column blocks of the upper triangle (first group starts at the diagonal).  Thread t: row box
// t & 63, column half t >> 6 (warp-uniform) -> one 32-bit half of every mask word.  The column boxes of block j + 1 are
// fetched from global memory into registers while block j is evaluated from shared memory (double buffer, one barrier
// per block), so a CTA pays the global-load latency once instead of once per 64 x 64 pairs, and the row box is loaded once
// per NMS_CG blocks.  Same bits as nms_mask_kernel (same pair function).
constexpr int NMS_CG = 8;
__global__ void __launch_bounds__(128) nms_mask_kernel_v2(const float* __restrict__ sdets, const int32_t* __restrict__ cand_cnt,
                                                         int n_fixed, int max_cand, int cb_stride, float thr, int ge,
                                                         unsigned long long* __restrict__ mask) {
  const int b = blockIdx.z;
  int n = cand_cnt ? cand_cnt[b] : n_fixed;
  n = n < max_cand ? n : max_cand;
  const int row_start = blockIdx.x;
  const int col_first = row_start + blockIdx.y * NMS_CG;
  if (row_start * 64 >= n || col_first * 64 >= n) return;
  const int col_blocks = (n + 63) >> 6;
  const int nblk = min(NMS_CG, col_blocks - col_first);
  const float* dets = sdets + (long long)b * max_cand * 5;
  __shared__ float4 cbox[2][64];
  __shared__ float carea[2][64];
  const int tid = threadIdx.x, r = tid & 63, half = tid >> 6;
  const int cur = 64 * row_start + r;
  const bool row_ok = cur < n;
  float4 rbx = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row_ok) rbx = make_float4(dets[(long long)cur * 5], dets[(long long)cur * 5 + 1], dets[(long long)cur * 5 + 2], dets[(long long)cur * 5 + 3]);
  const float Sa = __fmul_rn(__fadd_rn(__fsub_rn(rbx.z, rbx.x), 1.f), __fadd_rn(__fsub_rn(rbx.w, rbx.y), 1.f));
  const bool fast_ok = thr > 1e-3f && thr < 1e3f;
  // threads 0..63 stage column box `tid` of a block (boxes beyond n: zeros, never tested)
  auto fetch = [&](int cblk, float4& v) {
    const int c = 64 * cblk + tid;
    v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 64 && c < n) v = make_float4(dets[(long long)c * 5], dets[(long long)c * 5 + 1], dets[(long long)c * 5 + 2], dets[(long long)c * 5 + 3]);
  };
  auto stage = [&](int buf, const float4& v) {
    if (tid < 64) {
      cbox[buf][tid] = v;
      carea[buf][tid] = __fmul_rn(__fadd_rn(__fsub_rn(v.z, v.x), 1.f), __fadd_rn(__fsub_rn(v.w, v.y), 1.f));
    }
  };
  float4 nxt;
  fetch(col_first, nxt);
  stage(0, nxt);
  __syncthreads();
  unsigned int* mask32 = reinterpret_cast<unsigned int*>(mask + ((long long)b * max_cand + cur) * cb_stride);
  for (int j = 0; j < nblk; ++j) {
    const int cblk = col_first + j, buf = j & 1;
    if (j + 1 < nblk) fetch(cblk + 1, nxt);
    if (row_ok) {
      const int col_size = min(n - cblk * 64, 64);
      const int base = 32 * half;
      const int lo = max(base, cblk == row_start ? r + 1 : 0), hi = min(base + 32, col_size);
      unsigned int t = 0u;
      if (!fast_ok) {
        for (int i = lo; i < hi; ++i)
          if (nms_pair_over(rbx, Sa, cbox[buf][i], carea[buf][i], thr, ge, false)) t |= 1u << (i & 31);
      } else if (lo < hi) {
        // Phase 1, branch-free over all 32 columns of the half: which column boxes overlap the row box at all?  w > 0 <=> d_w > -1
        // exactly (w = fl(d_w + 1) and the floats next to -1 are 2^-24 apart), so the test needs 2 min/max + 1 subtract per axis.
        // Columns outside [lo, hi) are evaluated on whatever the staging buffer holds (finite zeros) and masked off.
        unsigned int cand = 0u;
#pragma unroll 8
        for (int ii = 0; ii < 32; ++ii) {
          const float4 bb = cbox[buf][base + ii];
          const float dw_ = __fsub_rn(fminf(rbx.z, bb.z), fmaxf(rbx.x, bb.x));
          const float dh_ = __fsub_rn(fminf(rbx.w, bb.w), fmaxf(rbx.y, bb.y));
          cand |= ((dw_ > -1.f && dh_ > -1.f) ? 1u : 0u) << ii;
        }
        const unsigned int upto = (hi - base) >= 32 ? 0xffffffffu : ((1u << (hi - base)) - 1u);
        cand &= upto & ~((1u << (lo - base)) - 1u);
        // Phase 2: the IoU decision proper only for the overlapping pairs (a few per cent of all pairs)
        while (cand) {
          const int ii = __ffs((int)cand) - 1;
          cand &= cand - 1u;
          if (nms_pair_over(rbx, Sa, cbox[buf][base + ii], carea[buf][base + ii], thr, ge, true)) t |= 1u << ii;
        }
      }
      mask32[2 * cblk + half] = t;   // little endian: half 0 = columns 0..31 of the 64-bit word
    }
    if (j + 1 < nblk) stage(buf ^ 1, nxt);
    __syncthreads();
  }
}

// One CTA per image: the serial host loop of nms_cuda.c:46-58, 64 boxes at a time.  Per 64-box block the critical path is
// (a) resolving the block against its own diagonal mask word by word -- only over the still-alive boxes -- and (b) OR-ing
// the rows of the boxes it keeps into the later column blocks; the next block's diagonal words are prefetched meanwhile,
// and the kept boxes are written out in parallel after the scan.
__global__ void __launch_bounds__(128) nms_reduce_kernel(const unsigned long long* __restrict__ mask,
                                                        const int32_t* __restrict__ cand_cnt, int max_cand, int cb_stride,
                                                        const int32_t* __restrict__ ranks_sorted,
                                                        const float* __restrict__ sdets, int64_t* __restrict__ keep_idx,
                                                        int32_t* __restrict__ keep_cnt, float* __restrict__ out_scores,
                                                        float* __restrict__ out_boxes) {
  extern __shared__ unsigned long long dyn[];  // remv[cb_stride] | keepbits[cb_stride] | prefix[cb_stride + 1] (as int)
  unsigned long long* remv = dyn;
  unsigned long long* keepbits = dyn + cb_stride;
  int* prefix = reinterpret_cast<int*>(dyn + 2 * cb_stride);
  __shared__ unsigned long long diag[2][64];
  const int b = blockIdx.x, tid = threadIdx.x;
  int n = cand_cnt[b];
  n = n < max_cand ? n : max_cand;
  const int col_blocks = (n + 63) / 64;
  for (int j = tid; j < cb_stride; j += blockDim.x) { remv[j] = 0ULL; keepbits[j] = 0ULL; }
  const unsigned long long* m = mask + (long long)b * max_cand * cb_stride;
  if (tid < 64 && col_blocks > 0) diag[0][tid] = tid < min(n, 64) ? m[(long long)tid * cb_stride] : 0ULL;
  __syncthreads();
  for (int rb = 0; rb < col_blocks; ++rb) {
    const int rows = min(n - rb * 64, 64);
    const unsigned long long* dg = diag[rb & 1];
    // prefetch the next block's diagonal words (independent of the scan state)
    unsigned long long nd = 0ULL;
    if (tid >= 64 && rb + 1 < col_blocks) {
      const int r = tid - 64, nrows = min(n - (rb + 1) * 64, 64);
      if (r < nrows) nd = m[(long long)((rb + 1) * 64 + r) * cb_stride + rb + 1];
    }
    if (tid == 0) {
      const unsigned long long valid = rows == 64 ? ~0ULL : ((1ULL << rows) - 1ULL);
      unsigned long long cur = remv[rb], kb = 0ULL;
      unsigned long long alive = ~cur & valid;
      while (alive) {  // boxes are visited in score order; diag[t] only has bits above t
        const int t = __ffsll((long long)alive) - 1;
        kb |= 1ULL << t;
        cur |= dg[t];
        alive &= ~cur;
        alive &= ~((2ULL << t) - 1ULL);
      }
      keepbits[rb] = kb;
    }
    if (tid >= 64 && rb + 1 < col_blocks) diag[(rb + 1) & 1][tid - 64] = nd;
    __syncthreads();
    const unsigned long long kb = keepbits[rb];
    // later column blocks: OR in the rows of every kept box of this block
    for (int j = rb + 1 + tid; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = remv[j];
      unsigned long long bits = kb;
      while (bits) {
        int t = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= m[(long long)(rb * 64 + t) * cb_stride + j];
      }
      remv[j] = acc;
    }
    __syncthreads();
  }
  // exclusive prefix of the per-block keep counts (col_blocks <= a few hundred: one thread)
  if (tid == 0) {
    int run = 0;
    for (int j = 0; j < col_blocks; ++j) { prefix[j] = run; run += __popcll(keepbits[j]); }
    prefix[col_blocks] = run;
    keep_cnt[b] = run;
  }
  __syncthreads();
  const long long o = (long long)b * max_cand;
  for (int srt = tid; srt < n; srt += blockDim.x) {
    const int rb = srt >> 6, t = srt & 63;
    const unsigned long long kb = keepbits[rb];
    if (!((kb >> t) & 1ULL)) continue;
    const int pos = prefix[rb] + __popcll(kb & ((1ULL << t) - 1ULL));
    keep_idx[o + pos] = (int64_t)ranks_sorted[o + srt];
    if (out_scores) out_scores[o + pos] = sdets[(o + srt) * 5 + 4];
    if (out_boxes) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out_boxes[(o + pos) * 4 + k] = sdets[(o + srt) * 5 + k];
    }
  }
}

// Second-generation reduction: same serial semantics, but the mask rows of a 64-box block (words rb .. col_blocks) are staged
// in shared memory by cp.async two blocks ahead, so neither the resolve of the diagonal words nor the OR of the kept rows
// waits on an L2 round trip inside the serial chain.  Dynamic smem: remv | keepbits | prefix | 2 x [64][cbs] words.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__global__ void __launch_bounds__(128) nms_reduce_kernel_v2(const unsigned long long* __restrict__ mask,
                                                           const int32_t* __restrict__ cand_cnt, int max_cand, int cb_stride,
                                                           const int32_t* __restrict__ ranks_sorted,
                                                           const float* __restrict__ sdets, int64_t* __restrict__ keep_idx,
                                                           int32_t* __restrict__ keep_cnt, float* __restrict__ out_scores,
                                                           float* __restrict__ out_boxes) {
  extern __shared__ __align__(16) unsigned long long dyn2[];
  const int cbs = (cb_stride + 1) & ~1;   // row pitch of the staged blocks in words (16-byte chunks)
  unsigned long long* remv = dyn2;
  unsigned long long* keepbits = dyn2 + cbs;
  int* prefix = reinterpret_cast<int*>(dyn2 + 2 * cbs);
  unsigned long long* stg = dyn2 + 2 * cbs + (((cbs + 2) / 2 + 1) & ~1);   // 16-byte aligned
  const int b = blockIdx.x, tid = threadIdx.x;
  int n = cand_cnt[b];
  n = n < max_cand ? n : max_cand;
  const int col_blocks = (n + 63) / 64;
  for (int j = tid; j < cbs; j += blockDim.x) { remv[j] = 0ULL; keepbits[j] = 0ULL; }
  const unsigned long long* m = mask + (long long)b * max_cand * cb_stride;
  // stage block rb: rows rb*64 .. +63 (those < n), words (rb & ~1) .. col_blocks-1 rounded up to a pair; cb_stride is even or
  // the row start may be 8-byte aligned only -> fall back to 8-byte copies when the pitch is odd
  const bool pair_ok = (cb_stride & 1) == 0;
  auto prefetch = [&](int rb) {
    if (rb < col_blocks) {
      unsigned long long* dst = stg + (size_t)(rb & 1) * 64 * cbs;
      const int rows = min(n - rb * 64, 64);
      if (pair_ok) {
        const int w0 = rb & ~1, nch = (col_blocks - w0 + 1) >> 1;   // 16-byte chunks per row
        for (int i = tid; i < rows * nch; i += blockDim.x) {
          const int rr = i / nch, c = w0 + 2 * (i - rr * nch);
          cp_async16(dst + rr * cbs + c, m + (long long)(rb * 64 + rr) * cb_stride + c);
        }
      } else {
        const int nw = col_blocks - rb;
        for (int i = tid; i < rows * nw; i += blockDim.x) {
          const int rr = i / nw, c = rb + (i - rr * nw);
          dst[rr * cbs + c] = m[(long long)(rb * 64 + rr) * cb_stride + c];
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(0);
  prefetch(1);
  for (int rb = 0; rb < col_blocks; ++rb) {
    asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest group (block rb + 1) has landed
    __syncthreads();
    const unsigned long long* blk = stg + (size_t)(rb & 1) * 64 * cbs;
    const int rows = min(n - rb * 64, 64);
    if (tid == 0) {
      const unsigned long long valid = rows == 64 ? ~0ULL : ((1ULL << rows) - 1ULL);
      unsigned long long cur = remv[rb], kb = 0ULL;
      unsigned long long alive = ~cur & valid;
      while (alive) {  // boxes are visited in score order; the diagonal word of box t only has bits above t
        const int t = __ffsll((long long)alive) - 1;
        kb |= 1ULL << t;
        cur |= blk[t * cbs + rb];
        alive &= ~cur;
        alive &= ~((2ULL << t) - 1ULL);
      }
      keepbits[rb] = kb;
    }
    __syncthreads();
    const unsigned long long kb = keepbits[rb];
    for (int j = rb + 1 + tid; j < col_blocks; j += blockDim.x) {
      unsigned long long acc = remv[j];
      unsigned long long bits = kb;
      while (bits) {
        const int t = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        acc |= blk[t * cbs + j];
      }
      remv[j] = acc;
    }
    __syncthreads();   // every thread is done with this staging buffer: refill it with block rb + 2
    prefetch(rb + 2);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (tid == 0) {
    int run = 0;
    for (int j = 0; j < col_blocks; ++j) { prefix[j] = run; run += __popcll(keepbits[j]); }
    prefix[col_blocks] = run;
    keep_cnt[b] = run;
  }
  __syncthreads();
  const long long o = (long long)b * max_cand;
  for (int srt = tid; srt < n; srt += blockDim.x) {
    const int rb = srt >> 6, t = srt & 63;
    const unsigned long long kb = keepbits[rb];
    if (!((kb >> t) & 1ULL)) continue;
    const int pos = prefix[rb] + __popcll(kb & ((1ULL << t) - 1ULL));
    keep_idx[o + pos] = (int64_t)ranks_sorted[o + srt];
    if (out_scores) out_scores[o + pos] = sdets[(o + srt) * 5 + 4];
    if (out_boxes) {
#pragma unroll
      for (int k = 0; k < 4; ++k) out_boxes[(o + pos) * 4 + k] = sdets[(o + srt) * 5 + k];
    }
  }
}

struct Workspace {
  float* keys_in;
  float* keys_out;
  int32_t* ranks_in;
  int32_t* ranks_out;
  int* seg_begin;
  int* seg_end;
  float* sdets;
  unsigned long long* mask;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cub_temp_bytes(int B, int max_cand) {
  size_t bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairsDescending((void*)nullptr, bytes, (const float*)nullptr, (float*)nullptr,
                                                     (const int32_t*)nullptr, (int32_t*)nullptr, B * max_cand, B,
                                                     (const int*)nullptr, (const int*)nullptr);
  return bytes;
}

Workspace carve(void* base, int B, int max_cand) {
  Workspace w;
  size_t off = 0;
  char* p = (char*)base;
  size_t n = (size_t)B * max_cand;
  int cb = (max_cand + 63) / 64;
  auto take = [&](size_t bytes) {
    char* q = p ? p + off : nullptr;
    off += align256(bytes);
    return q;
  };
  w.keys_in = (float*)take(n * 4);
  w.keys_out = (float*)take(n * 4);
  w.ranks_in = (int32_t*)take(n * 4);
  w.ranks_out = (int32_t*)take(n * 4);
  w.seg_begin = (int*)take((size_t)B * 4);
  w.seg_end = (int*)take((size_t)B * 4);
  w.sdets = (float*)take(n * 5 * 4);
  w.mask = (unsigned long long*)take(n * cb * 8);
  w.cub_bytes = cub_temp_bytes(B, max_cand);
  w.cub_temp = take(w.cub_bytes);
  w.total = off;
  return w;
}

// MPN_NMS_V2=0 selects the first-generation mask / reduce kernels (kept as the A/B reference)
int nms_v2_enabled() {
  static const int on = getenv("MPN_NMS_V2") ? atoi(getenv("MPN_NMS_V2")) : 1;
  return on;
}

size_t reduce_v2_smem(int cb) {
  const int cbs = (cb + 1) & ~1;
  return (size_t)(2 * cbs + (((cbs + 2) / 2 + 1) & ~1) + 2 * 64 * cbs) * sizeof(unsigned long long);
}

int launch_mask(const float* sdets, const int32_t* cand_cnt, int n_fixed, int max_cand, int B, float iou_thr, int ge,
                unsigned long long* mask, cudaStream_t st) {
  const int cb = (max_cand + 63) / 64;
  if (nms_v2_enabled() && cb <= 65535) {
    dim3 mg(cb, mpn_divup(cb, NMS_CG), B);
    nms_mask_kernel_v2<<<mg, 128, 0, st>>>(sdets, cand_cnt, n_fixed, max_cand, cb, iou_thr, ge, mask);
  } else {
    dim3 mg(cb * (cb + 1) / 2, 1, B);
    nms_mask_kernel<<<mg, 64, 0, st>>>(sdets, cand_cnt, n_fixed, max_cand, cb, cb, iou_thr, ge, mask);
  }
  MPN_LAUNCH_OK();
  return MPN_OK;
}

// ev (optional): 5 events recorded after sort, gather, mask, reduce (ev[0] is recorded by the caller before the filter,
// ev[1] after it) -- mpn_filter_sort_nms_profile
int run_core(const float* boxes, int box_stride, long long box_image_stride, int B, int max_cand, float iou_thr, int ge,
             const int32_t* cand_idx, const int32_t* cand_cnt, int64_t* keep_idx, int32_t* keep_cnt, float* out_scores,
             float* out_boxes, Workspace& w, cudaStream_t st, cudaEvent_t* ev = nullptr) {
  segments_kernel<<<mpn_divup(B, 128), 128, 0, st>>>(cand_cnt, B, max_cand, w.seg_begin, w.seg_end);
  MPN_LAUNCH_OK();
  size_t tb = w.cub_bytes;
  MPN_CUDA_OK(cub::DeviceSegmentedRadixSort::SortPairsDescending(w.cub_temp, tb, w.keys_in, w.keys_out, w.ranks_in, w.ranks_out,
                                                                 B * max_cand, B, w.seg_begin, w.seg_end, 0, 32, st));
  if (ev) MPN_CUDA_OK(cudaEventRecord(ev[2], st));
  dim3 gg(mpn_divup(max_cand, 256), B);
  gather_sorted_kernel<<<gg, 256, 0, st>>>(boxes, box_stride, box_image_stride, cand_idx, cand_cnt, w.keys_out, w.ranks_out,
                                           max_cand, w.sdets);
  MPN_LAUNCH_OK();
  if (ev) MPN_CUDA_OK(cudaEventRecord(ev[3], st));
  const int cb = (max_cand + 63) / 64;
  int rc = launch_mask(w.sdets, cand_cnt, 0, max_cand, B, iou_thr, ge, w.mask, st);
  if (rc) return rc;
  if (ev) MPN_CUDA_OK(cudaEventRecord(ev[4], st));
  const size_t v2_smem = reduce_v2_smem(cb);
  if (nms_v2_enabled() && v2_smem <= 200 * 1024) {
    if (v2_smem > 48 * 1024) MPN_CUDA_OK(cudaFuncSetAttribute(nms_reduce_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v2_smem));
    nms_reduce_kernel_v2<<<B, 128, v2_smem, st>>>(w.mask, cand_cnt, max_cand, cb, w.ranks_out, w.sdets, keep_idx, keep_cnt, out_scores,
                                                   out_boxes);
  } else {
    const size_t red_smem = (size_t)(2 * cb + (cb + 2) / 2) * sizeof(unsigned long long);
    MPN_CHECK_ARG(red_smem <= 200 * 1024, "nms: too many candidates for the on-device reduction (%d)", max_cand);
    if (red_smem > 48 * 1024) MPN_CUDA_OK(cudaFuncSetAttribute(nms_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)red_smem));
    nms_reduce_kernel<<<B, 128, red_smem, st>>>(w.mask, cand_cnt, max_cand, cb, w.ranks_out, w.sdets, keep_idx, keep_cnt, out_scores,
                                                out_boxes);
  }
  MPN_LAUNCH_OK();
  if (ev) MPN_CUDA_OK(cudaEventRecord(ev[5], st));
  return MPN_OK;
}

}  // namespace

extern "C" int mpn_num_anchors(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  long long a = 0;
  for (int l : kLevels) {
    int s = 1 << l;
    a += (long long)((H + s - 1) / s) * ((W + s - 1) / s) * 9;
  }
  return (int)a;
}

extern "C" int mpn_generate_anchors(int H, int W, float* out) {
  MPN_CHECK_ARG(H > 0 && W > 0 && out, "mpn_generate_anchors: bad argument");
  const double ratios[3] = {0.5, 1.0, 2.0};
  const double scales[3] = {pow(2.0, 0.0), pow(2.0, 1.0 / 3.0), pow(2.0, 2.0 / 3.0)};  // anchors.py:19
  long long o = 0;
  for (int l : kLevels) {
    const int stride = 1 << l;
    const double base_size = (double)(1 << (l + 2));  // anchors.py:15
    double base[9][4];
    for (int ri = 0; ri < 3; ++ri)
      for (int si = 0; si < 3; ++si) {
        double side = base_size * scales[si];   // anchors.py:55
        double area = side * side;              // :58
        double w = sqrt(area / ratios[ri]);     // :61
        double h = w * ratios[ri];              // :62
        double* b = base[ri * 3 + si];
        b[0] = 0.0 - w * 0.5; b[1] = 0.0 - h * 0.5; b[2] = w - w * 0.5; b[3] = h - h * 0.5;  // :65-66
      }
    const int fh = (H + stride - 1) / stride, fw = (W + stride - 1) / stride;  // anchors.py:25
    for (int y = 0; y < fh; ++y) {
      const double sy = ((double)y + 0.5) * stride;  // anchors.py:108
      for (int x = 0; x < fw; ++x) {
        const double sx = ((double)x + 0.5) * stride;  // :107
        for (int a = 0; a < 9; ++a) {
          out[o++] = (float)(base[a][0] + sx);
          out[o++] = (float)(base[a][1] + sy);
          out[o++] = (float)(base[a][2] + sx);
          out[o++] = (float)(base[a][3] + sy);
        }
      }
    }
  }
  return MPN_OK;
}

extern "C" int mpn_decode_clip(const float* anchors, const float* reg, float* boxes, int B, int A, int H, int W, void* stream) {
  MPN_CHECK_ARG(anchors && reg && boxes && B > 0 && A > 0, "mpn_decode_clip: bad argument");
  const int clip = (H > 0 && W > 0) ? 1 : 0;  // H <= 0: decode only (BBoxTransform without ClipBoxes)
  long long total = (long long)B * A;
  int grid = (int)((total + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  decode_clip_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(anchors, reg, boxes, B, A, (float)H, (float)W, clip);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

extern "C" size_t mpn_detect_workspace_bytes(int B, int A, int max_cand) {
  (void)A;
  if (B <= 0 || max_cand <= 0) return 0;
  return carve(nullptr, B, max_cand).total;
}

extern "C" int mpn_filter_sort_nms(const float* cls, const float* boxes, int B, int A, float score_thresh, float iou_thresh, int ge,
                                   int max_cand, int32_t* cand_idx, int32_t* cand_cnt, int64_t* keep_idx, int32_t* keep_cnt,
                                   float* out_scores, float* out_boxes, void* workspace, size_t workspace_bytes, void* stream) {
  MPN_CHECK_ARG(cls && boxes && B > 0 && A > 0 && max_cand > 0, "mpn_filter_sort_nms: bad argument");
  MPN_CHECK_ARG(cand_idx && cand_cnt && keep_idx && keep_cnt && workspace, "mpn_filter_sort_nms: null output/workspace");
  MPN_CHECK_ARG((long long)B * max_cand < (1LL << 30), "mpn_filter_sort_nms: B*max_cand too large");
  Workspace w = carve(workspace, B, max_cand);
  MPN_CHECK_ARG(workspace_bytes >= w.total, "mpn_filter_sort_nms: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  cudaStream_t st = (cudaStream_t)stream;
  filter_compact_kernel<<<B, 1024, 0, st>>>(cls, A, score_thresh, max_cand, cand_idx, cand_cnt, w.keys_in, w.ranks_in);
  MPN_LAUNCH_OK();
  return run_core(boxes, 4, (long long)A * 4, B, max_cand, iou_thresh, ge, cand_idx, cand_cnt, keep_idx, keep_cnt, out_scores,
                  out_boxes, w, st);
}

extern "C" int mpn_filter_sort_nms_profile(const float* cls, const float* boxes, int B, int A, float score_thresh, float iou_thresh,
                                           int ge, int max_cand, int32_t* cand_idx, int32_t* cand_cnt, int64_t* keep_idx,
                                           int32_t* keep_cnt, float* out_scores, float* out_boxes, void* workspace,
                                           size_t workspace_bytes, void* stream, float* stage_ms) {
  MPN_CHECK_ARG(cls && boxes && B > 0 && A > 0 && max_cand > 0 && stage_ms, "mpn_filter_sort_nms_profile: bad argument");
  MPN_CHECK_ARG(cand_idx && cand_cnt && keep_idx && keep_cnt && workspace, "mpn_filter_sort_nms_profile: null output/workspace");
  Workspace w = carve(workspace, B, max_cand);
  MPN_CHECK_ARG(workspace_bytes >= w.total, "mpn_filter_sort_nms_profile: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t ev[6];
  for (int i = 0; i < 6; ++i) MPN_CUDA_OK(cudaEventCreate(&ev[i]));
  MPN_CUDA_OK(cudaEventRecord(ev[0], st));
  filter_compact_kernel<<<B, 1024, 0, st>>>(cls, A, score_thresh, max_cand, cand_idx, cand_cnt, w.keys_in, w.ranks_in);
  MPN_LAUNCH_OK();
  MPN_CUDA_OK(cudaEventRecord(ev[1], st));
  int rc = run_core(boxes, 4, (long long)A * 4, B, max_cand, iou_thresh, ge, cand_idx, cand_cnt, keep_idx, keep_cnt, out_scores,
                    out_boxes, w, st, ev);
  if (rc == MPN_OK) {
    MPN_CUDA_OK(cudaEventSynchronize(ev[5]));
    for (int i = 0; i < 5; ++i) MPN_CUDA_OK(cudaEventElapsedTime(&stage_ms[i], ev[i], ev[i + 1]));
  }
  for (int i = 0; i < 6; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

extern "C" size_t mpn_nms_workspace_bytes(int n) {
  if (n <= 0) return 256;
  // extra: cand_idx[n] + cand_cnt[1]
  return carve(nullptr, 1, n).total + align256((size_t)n * 4) + 256;
}

extern "C" int mpn_nms(const float* dets, int n, float iou_thresh, int ge, int64_t* keep, int32_t* num_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  MPN_CHECK_ARG(n >= 0 && num_out, "mpn_nms: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    MPN_CUDA_OK(cudaMemsetAsync(num_out, 0, sizeof(int32_t), st));
    return MPN_OK;
  }
  MPN_CHECK_ARG(dets && keep && workspace, "mpn_nms: null pointer");
  MPN_CHECK_ARG(workspace_bytes >= mpn_nms_workspace_bytes(n), "mpn_nms: workspace too small");
  Workspace w = carve(workspace, 1, n);
  int32_t* cand_idx = (int32_t*)((char*)workspace + w.total);
  int32_t* cand_cnt = (int32_t*)((char*)cand_idx + align256((size_t)n * 4));
  iota_dets_kernel<<<mpn_divup(n, 256), 256, 0, st>>>(dets, n, w.keys_in, w.ranks_in, cand_idx, cand_cnt);
  MPN_LAUNCH_OK();
  return run_core(dets, 5, 0, 1, n, iou_thresh, ge, cand_idx, cand_cnt, keep, num_out, nullptr, nullptr, w, st);
}

extern "C" int mpn_nms_mask(const float* sorted_dets, int n, float iou_thresh, int ge, uint64_t* mask, void* stream) {
  MPN_CHECK_ARG(sorted_dets && mask && n > 0, "mpn_nms_mask: bad argument");
  return launch_mask(sorted_dets, nullptr, n, n, 1, iou_thresh, ge, (unsigned long long*)mask, (cudaStream_t)stream);
}
