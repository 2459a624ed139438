// Experiment (not product code): how fast does one SM's TMA unit move [128 rows x W bytes] boxes as a function of the box row width
// W (32 / 64 / 128 bytes; global row pitch 2048 B, L2-resident source), for loads, stores and a load+store mix?  The epilogue of
// conv_tc_kernel moves its shortcut and output tiles as 64-byte (fp16 plane) and 32-byte (e5m2 plane) rows; the clock64 timeline
// (scripts/exp/trace_conv.py) shows its TMA instructions waiting ~2500 cycles each.  If the unit is paced per row rather than per
// byte, 64-channel chunks (128 / 64-byte rows) halve the epilogue's share.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_row_rate_probe scripts/exp/tma_row_rate_probe.cu && gpurun_out/tma_row_rate_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int DEPTH = 8;        // boxes in flight
constexpr int ROWS = 128;
constexpr int PITCH = 2048;

// mode 0: loads, 1: stores, 2: one load + one store per iteration
__global__ void probe(const __grid_constant__ CUtensorMap map, int W, int iters, int mode, int rows_total, unsigned long long* cycles,
                      const uint8_t* flat) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[DEPTH];
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * i));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const int box_bytes = W * ROWS;
  const int row0 = (int)((blockIdx.x * ROWS) % rows_total);
  const int ncol = PITCH / W;
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const int slot = it % DEPTH;
    const uint32_t dst = sbase + (W == 256 ? (slot % 4) * 32768 : slot * 16384);
    const uint32_t bar = b0 + 8 * slot;
    const int c0 = (it % ncol) * W;
    if (mode == 0 || mode == 2) {
      if (it >= DEPTH) {   // wait for the load that used this slot
        uint32_t ok = 0;
        const uint32_t par = ((it / DEPTH) - 1) & 1;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                   "l"(&map), "r"(bar), "r"(c0), "r"(row0)
                   : "memory");
    }
    if (mode == 3) {   // non-tensor bulk copy: W * ROWS contiguous bytes
      if (it >= DEPTH) {
        uint32_t ok = 0;
        const uint32_t par = ((it / DEPTH) - 1) & 1;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(flat + (size_t)row0 * PITCH + (size_t)(it % 8) * box_bytes), "r"(box_bytes), "r"(bar)
                   : "memory");
    }
    if (mode == 1 || mode == 2) {
      const uint32_t src = sbase + (DEPTH + slot % 4) * 16384;
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map), "r"(src), "r"((c0 + PITCH / 2) % PITCH),
                   "r"(row0)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
  }
  if (mode == 0 || mode == 2 || mode == 3) {
    for (int it = iters; it < iters + DEPTH; ++it) {
      const int slot = it % DEPTH;
      if (it < DEPTH) continue;
      uint32_t ok = 0;
      const uint32_t par = ((it / DEPTH) - 1) & 1;
      while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b0 + 8 * slot), "r"(par) : "memory");
    }
  }
  if (mode == 1 || mode == 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  cycles[blockIdx.x] = clock64() - t0;
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int rows_total = 148 * ROWS;   // 148 x 128 rows x 2 KB = 38.8 MB: L2-resident
  uint8_t* d;
  unsigned long long* dc;
  cudaMalloc(&d, (size_t)rows_total * PITCH);
  cudaMemset(d, 1, (size_t)rows_total * PITCH);
  cudaMalloc(&dc, 148 * 8);
  const int SMEM = (DEPTH + 4) * 16384 + 2048;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  const int iters = 2048;
  const char* mname[4] = {"load", "store", "load+store", "bulk-1d"};
  printf("%-11s %5s %5s %8s | %12s %10s %10s %12s\n", "mode", "W", "grid", "swizzle", "cycles/box", "B/clk/SM", "rows/clk", "chip TB/s@1.9");
  for (int mode = 0; mode < 4; ++mode)
    for (int W = 32; W <= 256; W *= 2)
      for (int sw = 0; sw < 2; ++sw)
        for (int grid = 1; grid <= 148; grid += 147) {
          if (W == 256 && (sw == 1 || mode == 1 || mode == 2)) continue;   // 256-byte rows: no swizzle mode; 32 KB boxes: loads only
          if (mode == 3 && sw == 1) continue;
          CUtensorMap map;
          cuuint64_t dims[2] = {(cuuint64_t)PITCH, (cuuint64_t)rows_total};
          cuuint64_t strides[1] = {(cuuint64_t)PITCH};
          cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)ROWS};
          cuuint32_t es[2] = {1, 1};
          const CUtensorMapSwizzle swz = sw == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                                                 : (W == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : W == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
          CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<grid, 32, SMEM>>>(map, W, iters, mode, rows_total, dc, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 2; }
          }
          unsigned long long hc[148];
          cudaMemcpy(hc, dc, grid * 8, cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < grid; ++i) avg += (double)hc[i];
          avg /= grid;
          const double boxes = (mode == 2 ? 2.0 : 1.0) * iters;
          const double bpc = boxes * W * ROWS / avg;
          printf("%-11s %5d %5d %8s | %12.1f %10.2f %10.3f %12.2f\n", mname[mode], W, grid, sw ? "matched" : "none", avg / boxes, bpc, boxes * ROWS / avg,
                 bpc * grid * 1.9e9 / 1e12);
        }
  return 0;
}
