/* oracle/nms_oracle.c -- TEST INFRASTRUCTURE: scalar C restatement of the reference NMS.
 *
 * The reference's own lib/nms/src/nms.c and nms_cuda.c cannot be compiled (they are
 * written against the removed TH/THC C API), so their algorithms are restated here on
 * plain arrays.  Compile with -ffp-contract=off so no FMA is formed (the CUDA source has
 * no contractible expression either: nms_kernel.cu:16-24).
 *
 *   oracle_cpu_nms  <- lib/nms/src/nms.c:4-69       (greedy O(N^2), suppress at ovr >= thresh,
 *                      areas and order supplied by the caller as in pth_nms.py:16-17)
 *   oracle_iou      <- lib/nms/src/cuda/nms_kernel.cu:16-24   (devIoU, +1 pixel convention)
 *   oracle_nms_mask <- lib/nms/src/cuda/nms_kernel.cu:26-70   (64-wide bit mask of j>i, IoU > thresh)
 *   oracle_gpu_nms  <- lib/nms/src/nms_cuda.c:17-67           (mask + serial host reduction)
 * Return value 1 = success, like the reference entry points (nms.c:68, nms_cuda.c:66).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TPB 64 /* threadsPerBlock = sizeof(unsigned long long)*8, nms_kernel.h:11 */
#define DIVUP(m, n) ((m) / (n) + ((m) % (n) > 0))

int oracle_cpu_nms(int64_t* keep_out, int64_t* num_out, const float* boxes, int64_t boxes_num, int64_t boxes_dim,
                   const int64_t* order, const float* areas, float thresh) {
  unsigned char* suppressed = (unsigned char*)calloc(boxes_num > 0 ? boxes_num : 1, 1);
  int64_t num_to_keep = 0;
  for (int64_t _i = 0; _i < boxes_num; ++_i) {
    int64_t i = order[_i];
    if (suppressed[i] == 1) continue;
    keep_out[num_to_keep++] = i;
    float ix1 = boxes[i * boxes_dim], iy1 = boxes[i * boxes_dim + 1];
    float ix2 = boxes[i * boxes_dim + 2], iy2 = boxes[i * boxes_dim + 3];
    float iarea = areas[i];
    for (int64_t _j = _i + 1; _j < boxes_num; ++_j) {
      int64_t j = order[_j];
      if (suppressed[j] == 1) continue;
      float xx1 = fmaxf(ix1, boxes[j * boxes_dim]);
      float yy1 = fmaxf(iy1, boxes[j * boxes_dim + 1]);
      float xx2 = fminf(ix2, boxes[j * boxes_dim + 2]);
      float yy2 = fminf(iy2, boxes[j * boxes_dim + 3]);
      float w = fmaxf(0.0f, xx2 - xx1 + 1);
      float h = fmaxf(0.0f, yy2 - yy1 + 1);
      float inter = w * h;
      float ovr = inter / (iarea + areas[j] - inter);
      if (ovr >= thresh) suppressed[j] = 1; /* nms.c:59 */
    }
  }
  *num_out = num_to_keep;
  free(suppressed);
  return 1;
}

float oracle_iou(const float* a, const float* b) {
  float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
  float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
  float width = fmaxf(right - left + 1, 0.f), height = fmaxf(bottom - top + 1, 0.f);
  float interS = width * height;
  float Sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
  float Sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
  return interS / (Sa + Sb - interS);
}

/* mask[i*col_blocks + cb] bit k set <=> j = cb*64+k, j > i (within the diagonal block; every j of a
 * later block), IoU(i,j) > thresh.  Blocks with cb < row block are written too by the reference
 * (the early-out is commented, nms_kernel.cu:31) but never read by the reduction (nms_cuda.c:52). */
void oracle_nms_mask(const float* boxes, int n, float thresh, int ge, uint64_t* mask) {
  const int col_blocks = DIVUP(n, TPB);
  for (int i = 0; i < n; ++i) {
    const int rb = i / TPB;
    for (int cb = 0; cb < col_blocks; ++cb) {
      const int col_size = (n - cb * TPB) < TPB ? (n - cb * TPB) : TPB;
      uint64_t t = 0;
      int start = (rb == cb) ? (i % TPB) + 1 : 0;
      for (int k = start; k < col_size; ++k) {
        float v = oracle_iou(boxes + (size_t)i * 5, boxes + (size_t)(cb * TPB + k) * 5);
        if (ge ? (v >= thresh) : (v > thresh)) t |= 1ULL << k;
      }
      mask[(size_t)i * col_blocks + cb] = t;
    }
  }
}

/* boxes: [n,5] already sorted by descending score (pth_nms.py:33-35). keep: n int64. */
int oracle_gpu_nms(int64_t* keep, int64_t* num_out, const float* boxes, int n, float thresh, int ge) {
  const int col_blocks = DIVUP(n, TPB);
  uint64_t* mask = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(n > 0 ? n : 1) * (col_blocks > 0 ? col_blocks : 1));
  uint64_t* remv = (uint64_t*)calloc(col_blocks > 0 ? col_blocks : 1, sizeof(uint64_t));
  oracle_nms_mask(boxes, n, thresh, ge, mask);
  int64_t num_to_keep = 0;
  for (int i = 0; i < n; i++) { /* nms_cuda.c:46-58 */
    int nblock = i / TPB, inblock = i % TPB;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      const uint64_t* p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
    }
  }
  *num_out = num_to_keep;
  free(mask);
  free(remv);
  return 1;
}

/* Reduction alone, from a mask produced elsewhere (e.g. oracle/_ref's nms_kernel on the GPU). */
int oracle_reduce_mask(int64_t* keep, int64_t* num_out, const uint64_t* mask, int n) {
  const int col_blocks = DIVUP(n, TPB);
  uint64_t* remv = (uint64_t*)calloc(col_blocks > 0 ? col_blocks : 1, sizeof(uint64_t));
  int64_t num_to_keep = 0;
  for (int i = 0; i < n; i++) {
    int nblock = i / TPB, inblock = i % TPB;
    if (!(remv[nblock] & (1ULL << inblock))) {
      keep[num_to_keep++] = i;
      const uint64_t* p = mask + (size_t)i * col_blocks;
      for (int j = nblock; j < col_blocks; j++) remv[j] |= p[j];
    }
  }
  *num_out = num_to_keep;
  free(remv);
  return 1;
}
