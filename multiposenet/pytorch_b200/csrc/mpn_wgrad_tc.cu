// tcgen05 + TMA weight-gradient kernel (training step).
//
//   dW[co, r, s, ci] = sum over output pixels (n, oh, ow) of  dY[n, oh, ow, co] * X[n, oh*stride + r - pad, ow*stride + s - pad, ci]
//
// GEMM view per filter tap: M = 128 output channels, N = BN input channels, K = output pixels.  Both operands are
// "MN-major" for the tensor core (the contiguous dimension of an NHWC tensor is the channel = M or N, the K index is the
// pixel), which tcgen05 supports directly through the instruction descriptor's major bits, so the activations and
// the output gradients are consumed in place: a K block is one 4-D TMA box [64 ch, TW, TH, TN] of <= 64 pixels of dY
// and the correspondingly shifted box of X (zero padding = TMA out-of-bounds fill; stride-2 convs read the four
// phase views, exactly like the forward kernel).  Each work item (Cout tile, Cin tile, tap, pixel chunk)
// accumulates its [128 x BN] fp32 tile in TMEM over its pixel chunk and adds it to dW with fp32 atomics
// (split-K over pixels keeps all 148 SMs busy even for 64x64 filters).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "mpn_common.cuh"
#include "mpn_tc_ptx.cuh"

namespace {

constexpr int KROWS = 64;                      // pixels per K block (box slot size)
constexpr int BOX_BYTES = KROWS * 128;         // one [64 pixels x 64 channels] bf16 box, 128-byte swizzled rows
constexpr int WG_THREADS = 256;
constexpr int WG_SMEM_LIMIT = 227 * 1024;
constexpr int WG_BAR_BYTES = 256;
constexpr int WG_EPI_STAGE = 4096;  // per epilogue warp: 32 x 32 fp32 tile, 16-byte chunks XOR-swizzled

struct WMaps {
  CUtensorMap dy[2];    // [plane]
  CUtensorMap x[2][4];  // [plane][phase]
};

struct WParams {
  int N, OH, OW, Cout, Cin;
  int TW, TH, TN, rows;
  int tiles_w, tiles_h, tiles_n, kblocks;
  int R, S, stride, pad;
  int co_tiles, ci_tiles, k_chunks, kb_per_chunk, total_items;
  int phase_empty;
  float* dw;
};

template <int BN, bool SPLIT, int STAGES>
__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WMaps maps, const WParams P) {
  constexpr int PLANES = SPLIT ? 2 : 1;
  constexpr int A_BOXES = 2, B_BOXES = BN / 64;
  constexpr int A_BYTES = A_BOXES * BOX_BYTES, B_BYTES = B_BOXES * BOX_BYTES;
  constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // f32 accumulate, bf16 x bf16, A and B both MN-major (bits 15, 16), N = BN, M = 128
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));
  uint8_t* epi_stage = smem_gen + STAGES * STAGE_BYTES + WG_BAR_BYTES;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // K rows a box never writes (rows < 64) must read as zero: clear the whole ring once
  for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += WG_THREADS) reinterpret_cast<uint4*>(smem_gen)[i] = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy zeros visible to the async proxy (TMA / MMA)

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const int taps = P.R * P.S;

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    const uint32_t tx_bytes = (uint32_t)PLANES * (uint32_t)(A_BOXES + B_BOXES) * (uint32_t)(P.rows * 128);
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x) {
      int t = item;
      const int kc = t % P.k_chunks; t /= P.k_chunks;
      const int tap = t % taps; t /= taps;
      const int ci_t = t % P.ci_tiles;
      const int co_t = t / P.ci_tiles;
      const int r = tap / P.S, s = tap - r * P.S;
      const int dh = r - P.pad, dw = s - P.pad;
      int ph = 0, hoff, woff;
      if (P.stride == 1) {
        hoff = dh;
        woff = dw;
      } else {
        const int hp = dh & 1, wp = dw & 1;
        hoff = (dh - hp) / 2;
        woff = (dw - wp) / 2;
        ph = hp * 2 + wp;
      }
      const bool empty_view = (P.phase_empty >> ph) & 1;
      const int kb0 = kc * P.kb_per_chunk, kb1 = min(kb0 + P.kb_per_chunk, P.kblocks);
      for (int kb = kb0; kb < kb1; ++kb) {
        int m = kb;
        const int tw_i = m % P.tiles_w; m /= P.tiles_w;
        const int th_i = m % P.tiles_h;
        const int tn_i = m / P.tiles_h;
        const int ow0 = tw_i * P.TW, oh0 = th_i * P.TH, n0 = tn_i * P.TN;
        mbar_wait(empty_bar(stage), phase ^ 1u);
        mbar_expect_tx(full_bar(stage), tx_bytes);
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + PLANES * A_BYTES;
#pragma unroll
        for (int p = 0; p < PLANES; ++p) {
#pragma unroll
          for (int j = 0; j < A_BOXES; ++j)
            tma_load_4d(sa + p * A_BYTES + j * BOX_BYTES, &maps.dy[p], full_bar(stage), co_t * 128 + j * 64, ow0, oh0, n0);
#pragma unroll
          for (int j = 0; j < B_BOXES; ++j)
            tma_load_4d(sb + p * B_BYTES + j * BOX_BYTES, &maps.x[p][empty_view ? 0 : ph], full_bar(stage), ci_t * BN + j * 64,
                        ow0 + woff, empty_view ? (1 << 24) : oh0 + hoff, n0);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // =============================== MMA issuer ===============================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++it) {
      const int kc = item % P.k_chunks;
      const int kb0 = kc * P.kb_per_chunk, kb1 = min(kb0 + P.kb_per_chunk, P.kblocks);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sa = smem_base + stage * STAGE_BYTES;
        const uint32_t sb = sa + PLANES * A_BYTES;
#pragma unroll
        for (int k = 0; k < KROWS / 16; ++k) {
          const uint64_t a_hi = make_sdesc_mn(sa + k * 2048, BOX_BYTES, 1024);
          const uint64_t b_hi = make_sdesc_mn(sb + k * 2048, BOX_BYTES, 1024);
          umma_bf16(d_tmem, a_hi, b_hi, IDESC, (kb > kb0 || k > 0) ? 1u : 0u);
          if (SPLIT) {
            const uint64_t a_lo = make_sdesc_mn(sa + A_BYTES + k * 2048, BOX_BYTES, 1024);
            const uint64_t b_lo = make_sdesc_mn(sb + B_BYTES + k * 2048, BOX_BYTES, 1024);
            umma_bf16(d_tmem, a_lo, b_hi, IDESC, 1u);
            umma_bf16(d_tmem, a_hi, b_lo, IDESC, 1u);
          }
        }
        umma_commit(empty_bar(stage));
        if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (kb1 <= kb0) umma_commit(tfull_bar(acc));  // empty chunk (cannot happen with the host's chunking): keep the protocol alive
    }
  } else if (warp >= 4) {
    // =============================== epilogue: TMEM -> fp32 atomics on dW ===============================
    const int q = warp & 3;
    int it = 0;
    for (int item = blockIdx.x; item < P.total_items; item += gridDim.x, ++it) {
      int t = item;
      const int kc = t % P.k_chunks; t /= P.k_chunks;
      const int tap = t % taps; t /= taps;
      const int ci_t = t % P.ci_tiles;
      const int co_t = t / P.ci_tiles;
      const int kb0 = kc * P.kb_per_chunk, kb1 = min(kb0 + P.kb_per_chunk, P.kblocks);
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int co_base = co_t * 128 + q * 32;  // first output channel (dW row) of this warp
      float* stg = reinterpret_cast<float*>(epi_stage + (warp - 4) * WG_EPI_STAGE);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        if (ci_t * BN + c0 >= P.Cin) break;
        uint32_t raw[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
        TMEM_LD_32x32b_X32(taddr, raw);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // thread-per-row accumulators -> swizzled staging -> lane-per-column, so each RED instruction of the warp
        // adds 32 consecutive floats of one dW row (one 128-byte L2 transaction instead of 32 scattered ones)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) =
              make_uint4(raw[4 * c], raw[4 * c + 1], raw[4 * c + 2], raw[4 * c + 3]);
        __syncwarp();
        const int ci = ci_t * BN + c0 + lane;
        if (ci < P.Cin && kb1 > kb0) {
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int co = co_base + rr;
            if (co >= P.Cout) break;
            const float v = stg[rr * 32 + ((((lane >> 2) ^ (rr & 7))) << 2) + (lane & 3)];
            atomicAdd(P.dw + ((long long)co * taps + tap) * P.Cin + ci, v);
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// pixel box of at most 64 rows with the fewest wasted K rows
void choose_kbox(int N, int OH, int OW, int* TW, int* TH, int* TN) {
  double best = -1.0;
  int bw = 1, bh = 1, bn = 1;
  for (int tw = 1; tw <= KROWS && tw <= OW; ++tw)
    for (int th = 1; th * tw <= KROWS && th <= OH; ++th) {
      int tn = 1;
      if (tw == OW && th == OH) {
        tn = KROWS / (tw * th);
        if (tn > N) tn = N;
        if (tn < 1) tn = 1;
      }
      long long tiles = (long long)((OW + tw - 1) / tw) * ((OH + th - 1) / th) * ((N + tn - 1) / tn);
      double score = (double)N * OH * OW / (double)(tiles * KROWS) + 1e-4 * tw;
      if (score > best) { best = score; bw = tw; bh = th; bn = tn; }
    }
  *TW = bw; *TH = bh; *TN = bn;
}

template <int BN, bool SPLIT>
int launch_wgrad(const WMaps& maps, const WParams& P, cudaStream_t st) {
  constexpr int PLANES = SPLIT ? 2 : 1;
  constexpr int STAGE_BYTES = PLANES * (2 + BN / 64) * BOX_BYTES;
  constexpr int MAXS = (WG_SMEM_LIMIT - 1024 - WG_BAR_BYTES - 4 * WG_EPI_STAGE) / STAGE_BYTES;
  constexpr int STAGES = MAXS > 6 ? 6 : MAXS;
  static_assert(STAGES >= 2, "wgrad: need a 2-stage ring");
  const int smem = STAGES * STAGE_BYTES + 1024 + WG_BAR_BYTES + 4 * WG_EPI_STAGE;
  auto kern = wgrad_tc_kernel<BN, SPLIT, STAGES>;
  MPN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int dev = 0, sms = 148;
  MPN_CUDA_OK(cudaGetDevice(&dev));
  MPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = P.total_items < sms ? P.total_items : sms;
  kern<<<grid, WG_THREADS, smem, st>>>(maps, P);
  MPN_LAUNCH_OK();
  return MPN_OK;
}

}  // namespace

// d describes the FORWARD conv (x [N,H,W,Cin] -> y [N,OH,OW,Cout]); p->x_* = forward input, p->res_* = dY (gradient of
// the conv output, channel stride d->res_cstride >= Cout), dw = [Cout][R][S][Cin] fp32 (overwritten).
extern "C" int mpn_conv2d_wgrad(const mpn_conv_desc* d, const mpn_conv_ptrs* p, float* dw, void* stream) {
  MPN_CHECK_ARG(d && p && dw, "wgrad: null argument");
  MPN_CHECK_ARG(d->fmt == MPN_FMT_BF16 || d->fmt == MPN_FMT_BF16X2, "wgrad: tcgen05 path only (fmt BF16 / BF16X2)");
  const bool split = d->fmt == MPN_FMT_BF16X2;
  MPN_CHECK_ARG(d->stride == 1 || d->stride == 2, "wgrad: stride must be 1 or 2");
  MPN_CHECK_ARG(d->Cin % 64 == 0, "wgrad: Cin must be a multiple of 64");
  MPN_CHECK_ARG(p->x_hi && p->res_hi && (!split || (p->x_lo && p->res_lo)), "wgrad: missing x / dY planes");
  MPN_CHECK_ARG(d->res_cstride >= d->Cout && d->res_cstride % 8 == 0 && d->in_cstride % 8 == 0, "wgrad: bad channel strides");
  MPN_CHECK_ARG(d->OH == (d->H + 2 * d->pad - d->R) / d->stride + 1 && d->OW == (d->W + 2 * d->pad - d->S) / d->stride + 1,
                "wgrad: OH/OW do not match");
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    mpn_set_error("wgrad: cuTensorMapEncodeTiled entry point not available");
    return MPN_ERR_CUDA;
  }
  cudaStream_t st = (cudaStream_t)stream;
  MPN_CUDA_OK(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)d->Cout * d->R * d->S * d->Cin, st));

  WParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.OH = d->OH; P.OW = d->OW; P.Cout = d->Cout; P.Cin = d->Cin;
  choose_kbox(d->N, d->OH, d->OW, &P.TW, &P.TH, &P.TN);
  P.rows = P.TW * P.TH * P.TN;
  P.tiles_w = mpn_divup(d->OW, P.TW);
  P.tiles_h = mpn_divup(d->OH, P.TH);
  P.tiles_n = mpn_divup(d->N, P.TN);
  P.kblocks = P.tiles_w * P.tiles_h * P.tiles_n;
  P.R = d->R; P.S = d->S; P.stride = d->stride; P.pad = d->pad;
  const int BN = d->Cin >= 256 ? 256 : d->Cin >= 128 ? 128 : 64;
  P.co_tiles = mpn_divup(d->Cout, 128);
  P.ci_tiles = mpn_divup(d->Cin, BN);
  const int base_items = P.co_tiles * P.ci_tiles * d->R * d->S;
  int kc = mpn_divup(148, base_items);  // ~one wave of work items: every extra pixel chunk costs a full tile of atomics
  if (kc > P.kblocks) kc = P.kblocks;
  if (kc < 1) kc = 1;
  P.kb_per_chunk = mpn_divup(P.kblocks, kc);
  P.k_chunks = mpn_divup(P.kblocks, P.kb_per_chunk);  // every chunk non-empty
  P.total_items = base_items * P.k_chunks;
  P.dw = dw;

  alignas(64) WMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int planes = split ? 2 : 1;
  const int s2 = d->stride;
  const int nphase = s2 == 1 ? 1 : 4;
  const long long wpitch = d->in_wpitch > 0 ? d->in_wpitch : d->W;
  const long long hpitch = d->in_hpitch > 0 ? d->in_hpitch : d->H;
  MPN_CHECK_ARG(d->in_cstride >= d->Cin || d->k_overlap, "wgrad: in_cstride < Cin");
  for (int pl = 0; pl < planes; ++pl) {
    const char* xb = (const char*)(pl == 0 ? p->x_hi : p->x_lo);
    for (int ph = 0; ph < nphase; ++ph) {
      const int hp = ph >> 1, wp = ph & 1;
      const int Hp = (d->H - hp + s2 - 1) / s2, Wp = (d->W - wp + s2 - 1) / s2;
      if (Hp <= 0 || Wp <= 0) {
        P.phase_empty |= 1 << ph;
        continue;
      }
      cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)d->N};
      cuuint64_t strides[3] = {(cuuint64_t)d->in_cstride * s2 * 2ULL, (cuuint64_t)wpitch * d->in_cstride * s2 * 2ULL,
                               (cuuint64_t)hpitch * wpitch * d->in_cstride * 2ULL};
      cuuint32_t box[4] = {64u, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
      const char* base = xb + ((long long)hp * wpitch + wp) * d->in_cstride * 2LL;
      int rc = encode(fn, &maps.x[pl][ph], base, 4, dims, strides, box);
      if (rc) return rc;
    }
    MPN_CHECK_ARG(!(P.phase_empty & 1), "wgrad: empty input");
    cuuint64_t ydims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->OW, (cuuint64_t)d->OH, (cuuint64_t)d->N};
    cuuint64_t ystr[3] = {(cuuint64_t)d->res_cstride * 2ULL, (cuuint64_t)d->OW * d->res_cstride * 2ULL,
                          (cuuint64_t)d->OH * d->OW * d->res_cstride * 2ULL};
    cuuint32_t ybox[4] = {64u, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
    int rc = encode(fn, &maps.dy[pl], pl == 0 ? p->res_hi : p->res_lo, 4, ydims, ystr, ybox);
    if (rc) return rc;
  }
  if (split) {
    switch (BN) {
      case 256: return launch_wgrad<256, true>(maps, P, st);
      case 128: return launch_wgrad<128, true>(maps, P, st);
      default: return launch_wgrad<64, true>(maps, P, st);
    }
  }
  switch (BN) {
    case 256: return launch_wgrad<256, false>(maps, P, st);
    case 128: return launch_wgrad<128, false>(maps, P, st);
    default: return launch_wgrad<64, false>(maps, P, st);
  }
}
