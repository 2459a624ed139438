// Experiment (not product code): can TMA deliver the HIGH bytes of an fp16 plane (every second byte, starting at byte 1) as a
// K-major byte tile, i.e. can the e5m2 "h8" activation plane of MPN_FMT_F16F8 be dropped from HBM and derived from the fp16
// plane by the loader?  Encodes a UINT8 tensor map over the fp16 rows with elementStrides = {2, 1} and loads a box that starts
// at inner coordinate 1, for swizzle NONE / 64B / 128B and both candidate transaction sizes; dumps shared memory and reports
// where source byte (row, col) landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/tma_stride_probe scripts/exp/tma_stride_probe.cu && gpurun_out/tma_stride_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, uint8_t* out, int out_bytes, int tx_bytes, int c0, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  const uint32_t sbase = ((uint32_t)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;
  uint8_t* sgen = smem + (sbase - (uint32_t)__cvta_generic_to_shared(smem));
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) sgen[i] = 0xEE;
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(tx_bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(sbase),
                 "l"(&map), "r"(b), "r"(c0), "r"(0)
                 : "memory");
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
    *status = (int)ok;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < out_bytes; i += blockDim.x) out[i] = sgen[i];
}

int main() {
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no encode fn\n"); return 1; }
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  const int ROWS = 128, ROWB = 128;  // 64 fp16 per row
  std::vector<uint8_t> src(ROWS * ROWB);
  uint8_t *dsrc, *dout;
  int* dstat;
  const int OUT = 16384;
  cudaMalloc(&dsrc, src.size()); cudaMalloc(&dout, OUT); cudaMalloc(&dstat, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, OUT + 2048);
  const CUtensorMapSwizzle swz[3] = {CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B};
  const char* swn[3] = {"NONE", "64B", "128B"};
  for (int estride = 1; estride <= 2; ++estride)
    for (int si = 0; si < 3; ++si)
      for (int boxi = 0; boxi < 2; ++boxi) {
        // estride 2: box 128 (traversal) or 64; estride 1 (control): box 64 bytes of a 128-byte row
        const int box0 = estride == 2 ? (boxi == 0 ? 128 : 64) : 64;
        if (estride == 1 && boxi == 1) continue;
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)ROWB, (cuuint64_t)ROWS};
        cuuint64_t strides[1] = {(cuuint64_t)ROWB};
        cuuint32_t box[2] = {(cuuint32_t)box0, (cuuint32_t)ROWS};
        cuuint32_t es[2] = {(cuuint32_t)estride, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, dsrc, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz[si],
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("\n== elementStride %d, box0 %d, swizzle %s: encode rc %d\n", estride, box0, swn[si], (int)r);
        if (r != CUDA_SUCCESS) continue;
        for (int c0 = 0; c0 <= 1; ++c0) {
          const int expect_elems = estride == 2 ? (box0 + 1) / 2 : box0;
          for (int txi = 0; txi < 2; ++txi) {
            const int tx = (txi == 0 ? expect_elems : box0) * ROWS;
            if (txi == 1 && tx == expect_elems * ROWS) continue;
            std::vector<uint8_t> outc(OUT), outr(OUT);
            int st[2] = {0, 0};
            for (int pass = 0; pass < 2; ++pass) {
              for (int rr = 0; rr < ROWS; ++rr)
                for (int cc = 0; cc < ROWB; ++cc) src[rr * ROWB + cc] = pass == 0 ? (uint8_t)cc : (uint8_t)rr;
              cudaMemcpy(dsrc, src.data(), src.size(), cudaMemcpyHostToDevice);
              cudaMemset(dstat, 0, 4);
              probe<<<1, 128, OUT + 2048>>>(map, dout, OUT, tx, c0, dstat);
              cudaError_t e = cudaDeviceSynchronize();
              if (e != cudaSuccess) { printf("   c0 %d tx %d: kernel error %s\n", c0, tx, cudaGetErrorString(e)); return 2; }
              cudaMemcpy(pass == 0 ? outc.data() : outr.data(), dout, OUT, cudaMemcpyDeviceToHost);
              cudaMemcpy(&st[pass], dstat, 4, cudaMemcpyDeviceToHost);
            }
            // analyse: which (row, col) sits at each smem offset; compare with candidate layouts
            int written = 0, dense64 = 0, sw64 = 0, dense128 = 0, sw128 = 0, odd = 0, even = 0;
            for (int o = 0; o < OUT; ++o) {
              if (outc[o] == 0xEE && outr[o] == 0xEE) continue;
              ++written;
              const int cc = outc[o], rr = outr[o];
              if (cc & 1) ++odd; else ++even;
              const int k = (cc - c0) / estride;  // element index within the row's box
              if (o == rr * 64 + k) ++dense64;
              if (o == rr * 64 + ((((k >> 4) ^ ((rr >> 1) & 3)) << 4) | (k & 15))) ++sw64;
              if (o == rr * 128 + k) ++dense128;
              if (o == rr * 128 + ((((k >> 4) ^ (rr & 7)) << 4) | (k & 15))) ++sw128;
            }
            printf("   c0 %d tx %5d: barrier %s | bytes written %5d (odd cols %d, even cols %d) | layout match: dense64 %d  swz64 %d  dense128 %d  swz128 %d\n",
                   c0, tx, (st[0] && st[1]) ? "completed" : "TIMEOUT", written, odd, even, dense64, sw64, dense128, sw128);
            printf("      row0 first 24 smem bytes -> source cols:");
            for (int o = 0; o < 24; ++o) printf(" %d", outc[o] == 0xEE ? -1 : outc[o]);
            printf("\n");
          }
        }
      }
  return 0;
}
