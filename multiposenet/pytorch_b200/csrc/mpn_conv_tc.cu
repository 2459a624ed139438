// tcgen05 + TMA implicit-GEMM convolution for sm_100a (NHWC bf16 planes, fp32 accumulation in TMEM).
//
//   D[pixel, cout] = sum_{tap (r,s)} sum_{c}  X[n, oh*stride + r - pad, ow*stride + s - pad, c] * W[cout, r, s, c]
//
// GEMM view: M = a TW x TH x TN box of output pixels (<= 128 rows), N = BN output channels, K walks
// the filter taps and 64-channel blocks.  For every (tap, channel block) ONE 4-D TMA box load brings
// the shifted input window [64ch, TW, TH, TN] into a 128B-swizzled K-major shared-memory tile; the
// zero padding of the convolution is the TMA out-of-bounds fill, and stride-2 convolutions read one of
// four "phase" views of the input (tensor maps with doubled pixel strides and an offset base), so no
// im2col buffer ever exists in HBM.  The filter is a plain 2-D [Cout, R*S*Cin] K-major tensor.
//
// Warp roles (384 threads, or 640 with the wide epilogue; 1 CTA/SM or one CTA pair per TPC, persistent over tiles):
//   warp 0 lane 0 : TMA producer           (mbarrier full/empty ring of STAGES stages)
//   warp 1 lane 0 : tcgen05.mma issuer     (accumulators double-buffered in TMEM, 2 x BN columns; pair kernels: the leader CTA)
//   warp 2        : TMEM allocate / free
//   warps 2, 3    : after that, "agents": they issue the epilogue's bulk-tensor instructions (EPI_TMA: the stores of epilogue
//                   half 0 / 1; wide epilogue: shortcut loads and stores of group pair 0 / 1) -- a thread that issues one stalls
//                   until the TMA unit accepts it, which the math threads must not pay; MODE_F16F8C: the e5m2 converters
//   warps 4..11   : epilogue (two warps per TMEM lane quarter; wide: warps 4..19, four per quarter), variants (template EPI):
//                   EPI_TMA  tcgen05.ld (thread = pixel row) -> BN-fold scale/bias, residual, ReLU/sigmoid -> bf16 hi/lo
//                            -> 64B-swizzled smem box -> cp.async.bulk.tensor store (no per-row address math, no LSU stores);
//                            optional phase-class addends gathered per pixel (keypoint head conv2)
//                   EPI_TMA_RES  the same with the shortcut (or the 2x-upsample source) brought in by TMA boxes; NG = 4: the
//                            wide variant (64-channel boxes, shortcut added in place)
//                   EPI_LSU  same math through a swizzled smem transpose and 64-byte row segments per 4 lanes; handles the
//                            ragged nearest-upsample add and x2/x4/x8 replicated outputs (FPN laterals, keypoint concat)
//                   EPI_F32  fp32 NHWC / NCHW heads (Cout <= 64), thread-per-row stores
// The TMA unit is paced per box ROW (~2.25 cycles for any row <= 128 bytes, scripts/exp/tma_row_rate_probe.cu): operand and
// epilogue boxes are laid out for 128-byte rows wherever the formats allow (merged filter byte planes, 64-channel wide boxes).
// Precision modes: MPN_FMT_BF16   one bf16 plane per operand, one MMA per K step;
//                  MPN_FMT_BF16X2 hi/lo bf16 planes (x = hi + lo to ~2^-17), three MMAs per K step
//                  (hi*hi + lo*hi + hi*lo) accumulated in fp32 -> fp32-grade parity with the reference.
//                  MPN_FMT_F16F8  fp16 hi plane + two fp8 byte planes (lo8 = e5m2((x - hi) * 2^12), h8 = e5m2(x)); filters
//                  prescaled by 2^k: fp16 hi, lo8 = e4m3(w' - hi), h8 = e4m3(w' * 2^-12).  Per 64-channel K block:
//                  4 kind::f16 MMAs (hi*hi) + 2 + 2 kind::f8f6f4 MMAs of K = 32 (xlo8*wh8, xh8*wlo8) into the SAME fp32
//                  accumulator = 8 MMA slots instead of 12, product error ~2^-15 (tests/tools/precision_study.py).
//                  Operand variants: MODE_F16F8 (copy plane stored), MODE_F16F8B (input without it: fp16 weight residual, 10
//                  slots), MODE_F16F8C (input without it: derived in shared memory by warps 2, 3; opt-in).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "mpn_common.cuh"
#include "mpn_tc_ptx.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;           // bf16 elements = one 128-byte swizzle row
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 384;

// Debug build only (-DMPN_CONV_TRACE, scripts/exp/trace_conv.py): clock64 stamps of the producer / MMA issuer / epilogue issuers of
// the first CTAs into a global buffer -- the timeline of a persistent CTA.  Compiled out of the product library.
#ifdef MPN_CONV_TRACE
__device__ unsigned long long* g_trace_buf = nullptr;
__device__ int g_trace_cap = 0;
constexpr int TRACE_CTAS = 4, TRACE_ROLES = 8;
struct Tracer {
  unsigned long long* p;
  int n, cap;
  __device__ __forceinline__ void init(int role) {
    n = 0;
    cap = (blockIdx.x < TRACE_CTAS && g_trace_buf) ? g_trace_cap : 0;
    p = g_trace_buf + ((size_t)(blockIdx.x * TRACE_ROLES + role) * (size_t)g_trace_cap) * 2;
  }
  __device__ __forceinline__ void rec(int ev, int a, int b) {
    if (n < cap) {
      p[2 * n] = (unsigned long long)clock64();
      p[2 * n + 1] = (unsigned long long)ev | ((unsigned long long)(unsigned)a << 8) | ((unsigned long long)(unsigned)b << 32);
      ++n;
    }
  }
};
#define TRACE_INIT(role) Tracer tr_; tr_.init(role)
#define TRACE(ev, a, b) tr_.rec(ev, a, b)
#else
#define TRACE_INIT(role)
#define TRACE(ev, a, b)
#endif
constexpr int NUM_EPI_WARPS = 8;   // NUM_THREADS = 4 control warps (TMA, MMA, TMEM alloc / h8 converters) + 8 epilogue warps
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int BIAS_BYTES = 1024;  // EPI_TMA: the tile's folded-BN bias, 2 halves x 4 chunks x 32 channels (behind the barriers)
constexpr int EPI_STAGE_BYTES = 4096;  // per epilogue warp: 32 rows x 32 fp32 accumulators, 16B chunks XOR-swizzled

struct Maps {
  CUtensorMap a[3][5];  // [plane hi/lo(/h8)][phase hp*2+wp, or the segment of a multi-level launch]; F16F8: planes 1, 2 are byte tensors, 64B swizzle
  CUtensorMap b[3];     // [plane]
  CUtensorMap bs[3];    // [plane] filter map with a tail_bn-row box (tail sub-tiles)
  CUtensorMap y[3][5];  // [plane][segment] output tensor (EPI_TMA): box {32 ch, TW, TH, TN}, 64B swizzle (F16F8 byte planes: none)
  CUtensorMap r[2];     // [plane] residual tensor (res_mma): box {64 ch, TW, TH, TN}, 128B swizzle = MN-major B operand
  CUtensorMap ident;    // 128 x 128 bf16 identity matrix (res_mma): box {64, 128}, K-major A operand
};

// Multi-level launches (mpn_conv2d_fwd_multi: one layer of a RetinaNet tower over the five pyramid levels, posenet.py:262-263):
// the persistent tile list is the concatenation of the levels' tiles; a level ("segment") has its own geometry, input / output
// tensor maps and, for fp32 outputs, its offset in the concatenated output.  Single launches are one segment.
constexpr int MAX_SEG = 5;
struct Seg {
  int N, OH, OW, TW, TH, TN, rows;
  int tiles_w, tiles_h, tiles_n, m_tiles;
  int unit0;           // first schedulable M unit (pair kernels: pair of M tiles) of this segment
  long long out_off;   // EPI_F32: element offset of this segment's output inside y_hi
};

struct TcParams {
  int nseg;
  Seg seg[MAX_SEG];
  int N, OH, OW, Cout;
  int TW, TH, TN, rows;
  int tiles_w, tiles_h, tiles_n, tiles_co, total_tiles;
  int m_tiles;  // tiles_w * tiles_h * tiles_n; CTA-pair kernels: a "tile" is a pair of consecutive M tiles x one Cout block
  // Tail balancing: tiles [0, main_tiles) are BN wide; every further BN-wide tile is cut into tail_split sub-tiles of
  // tail_bn columns so that the last partial round of the persistent schedule is spread over all SMs.
  int main_tiles, tail_split, tail_bn;
  // res_mma: the residual tile is added by the tensor core: after the filter taps, one extra ring iteration per plane
  // loads I (128x128) as the A operand and the residual box [BN ch x 128 pixels] as an MN-major B operand, D += I * R.
  int res_mma;
  int R, S, stride, pad, kb_per_tap, Cin;
  int phase_empty;  // bit p set: phase view p has no pixels (tiny maps) -> load an all-OOB box instead
  const float* scale;
  const float* bias;
  const __nv_bfloat16* res_hi;
  const __nv_bfloat16* res_lo;
  int res_cstride;
  const __nv_bfloat16* up_hi;
  const __nv_bfloat16* up_lo;
  int up_h, up_w, up_cstride;
  int up_tma;  // EPI_TMA_RES brings in the 2x-nearest-upsample source (box {32 ch, TW/2, TH/2, TN}) instead of a shortcut tensor
  int flags, out_mode, out_cstride, out_coffset, out_rep;
  long long out_nstride;
  // phase-class addends (mpn_conv_desc.gat_*): per output pixel a low-resolution pixel and one of nine Cout-wide channel slices
#ifdef MPN_CONV_TRACE
  int dbg_skip;   // timing experiment (results are garbage): bit 0 skips the B h8 plane load, bit 1 the A h8 plane load
#endif
  int w_merged; // F16F8 / F16F8C: filter byte planes interleaved per K block (MPN_W_MERGED): one 128B-swizzled tile [lo8 | h8]
  int epi_agent;  // EPI_TMA: the stores are issued by the agent warps (MPN_EPI_AGENT=0: by a thread of each epilogue half)
  int res_pf;   // EPI_TMA_RES: L2 prefetch of the next tile's shortcut boxes (MPN_RES_PF=0 disables)
  int gat_n, gat_shift[2], gat_h[2], gat_w[2], gat_cstride[2];
  const void* gat_hi[2];
  const void* gat_lo[2];
  void* y_hi;
  void* y_lo;
  long long y_plane;  // F16F8: elements of one output plane (the h8 plane starts y_plane bytes after y_lo)
  float acc_scale;    // F16F8: 2^-k of the filter prescale, applied to the accumulator first
};

struct TileCoord {
  int co0, ncols, tw_i, th_i, tn_i, seg;
};

template <int BN, bool PAIR = false>
__device__ __forceinline__ TileCoord decode_tile(const TcParams& P, int tile, int rank = 0) {
  TileCoord c;
  int t = tile, sub = 0;
  c.ncols = BN;
  if (tile >= P.main_tiles) {
    const int j = tile - P.main_tiles;
    t = P.main_tiles + j / P.tail_split;
    sub = j - (j / P.tail_split) * P.tail_split;
    c.ncols = P.tail_bn;
  }
  const int co_t = t % P.tiles_co;
  int mt = t / P.tiles_co;   // M unit over all segments
  c.co0 = co_t * BN + sub * c.ncols;
  int sg = 0;
  while (sg + 1 < P.nseg && mt >= P.seg[sg + 1].unit0) ++sg;
  const Seg& g = P.seg[sg];
  c.seg = sg;
  mt -= g.unit0;
  if (PAIR) {
    mt = 2 * mt + rank;
    if (mt >= g.m_tiles) {  // odd tile count: the pair's second half is a box beyond the last image (zero loads, clipped stores)
      c.tw_i = c.th_i = 0;
      c.tn_i = g.tiles_n;
      return c;
    }
  }
  c.tw_i = mt % g.tiles_w;
  mt /= g.tiles_w;
  c.th_i = mt % g.tiles_h;
  c.tn_i = mt / g.tiles_h;
  return c;
}

// ---------------------------------------------------------------------------------------------
// EPI_TMA_RES = EPI_TMA with the shortcut tensor brought in by TMA as well (32-channel boxes, double-buffered per epilogue half)
enum { EPI_LSU = 0, EPI_F32 = 1, EPI_TMA = 2, EPI_TMA_RES = 3 };
constexpr int RES_STAGE_BYTES = 16384;  // per (half, buffer): [hi: rows x 64 B, 64B swizzle][lo: rows x 64 B | lo8: rows x 32 B]
// Wide epilogue (NG = 4 column groups of four warps, F16F8 shortcut convolutions whose output is stored without h8): per
// group pair one 24 KB output staging box and one 24 KB shortcut box of 64 channels, each [hi: rows x 128 B, 128B swizzle][lo8:
// rows x 64 B, 64B swizzle, at +16384]
constexpr int WIDE_BOX_BYTES = 24576;

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// MODE_F16F8B: F16F8 for an activation tensor stored WITHOUT its h8 plane (3 bytes per element in HBM instead of 4): the
// weight-residual term x * w_lo runs as x_hi(fp16) * w_lo16(fp16) on kind::f16 (4 more K = 16 MMAs per K block: 10 slots
// instead of 8, and a more accurate term) -- used where the layer is bound by HBM / the epilogue, not by the tensor pipe
// (every 1x1 conv of the backbone, the FPN laterals).  Filter planes: hi fp16, lo16 fp16, h8 e4m3.
// MODE_F16F8C: F16F8 for an activation tensor stored without its h8 plane, the plane DERIVED in shared memory: TMA pays ~2.25
// cycles per box row whatever its width (32..128 B: scripts/exp/tma_row_rate_probe.cu), so the 64-byte h8 rows cost as much as
// the 128-byte fp16 rows they copy.  Warps 2 and 3 round the fp16 tile TMA delivered to e5m2 (top byte, round half up, clamped
// to the largest finite e5m2) into the stage's h8 tile; 8 MMA slots per K block like MODE_F16F8, filter planes hi | lo8 | h8.
enum { MODE_BF16 = 0, MODE_BF16X2 = 1, MODE_F16F8 = 2, MODE_F16F8B = 3, MODE_F16F8C = 4 };
__host__ __device__ constexpr int a_stage_bytes(int mode) { return mode == MODE_BF16 ? 16384 : mode == MODE_F16F8B ? 24576 : 32768; }
__host__ __device__ constexpr int b_row_bytes(int mode) { return mode == MODE_BF16 ? 128 : mode == MODE_F16F8B ? 320 : 256; }

__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major byte operand, 64-byte rows, 64-byte swizzle, 8-row groups 512 B apart
__device__ __forceinline__ uint64_t make_sdesc64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;  // SWIZZLE_64B
  return d;
}
__device__ __forceinline__ uint32_t pack_f16(float a, float b) { return mpn_pack_f16x2_sat(a, b); }  // saturating: no inf planes
__device__ __forceinline__ float f16_lo_f(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xFFFFu))); }
__device__ __forceinline__ float f16_hi_f(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }
__device__ __forceinline__ uint32_t e5m2x2(float a, float b) {  // a in the low byte
  return (uint32_t)__nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E5M2);
}
__device__ __forceinline__ uint32_t e5m2x4(float a, float b, float c, float d) { return e5m2x2(a, b) | (e5m2x2(c, d) << 16); }
// v[0..7] += 8 fp16 values (hi plane) ; v[0..3] += 2^-12 * 4 e5m2 bytes (lo8 plane)
__device__ __forceinline__ void add_f16x8_reg(float* v, const uint4& t) {
  v[0] += f16_lo_f(t.x); v[1] += f16_hi_f(t.x); v[2] += f16_lo_f(t.y); v[3] += f16_hi_f(t.y);
  v[4] += f16_lo_f(t.z); v[5] += f16_hi_f(t.z); v[6] += f16_lo_f(t.w); v[7] += f16_hi_f(t.w);
}
__device__ __forceinline__ void add_e5m2x4_reg(float* v, uint32_t w) {
  // an e5m2 byte is the upper byte of an fp16: one PRMT turns two bytes into a half2 (the products with 2^-12 are exact, so
  // the FMA equals multiply-then-add)
  const uint32_t p01 = __byte_perm(w, 0u, 0x1404), p23 = __byte_perm(w, 0u, 0x3424);
  v[0] = fmaf(f16_lo_f(p01), MPN_F8_LO_INV, v[0]);
  v[1] = fmaf(f16_hi_f(p01), MPN_F8_LO_INV, v[1]);
  v[2] = fmaf(f16_lo_f(p23), MPN_F8_LO_INV, v[2]);
  v[3] = fmaf(f16_hi_f(p23), MPN_F8_LO_INV, v[3]);
}
// four fp16 (two words) -> their e5m2 roundings (top byte after adding half an e5m2 ulp; the carry runs into the exponent as it
// should), magnitude clamped to 0x7B = 57344, the largest finite e5m2: a saturated activation must not become inf * w
__device__ __forceinline__ uint32_t h8_round4(uint32_t w0, uint32_t w1) {
  // magnitude + half an e5m2 ulp, clamped below the first inf pattern (one VIADDMNMX.U16x2 per word), top bytes, signs back in
  const uint32_t m0 = __viaddmin_u16x2(w0 & 0x7FFF7FFFu, 0x00800080u, 0x7BFF7BFFu);
  const uint32_t m1 = __viaddmin_u16x2(w1 & 0x7FFF7FFFu, 0x00800080u, 0x7BFF7BFFu);
  return __byte_perm(m0, m1, 0x7531) | (__byte_perm(w0, w1, 0x7531) & 0x80808080u);
}
// 8 floats -> fp16 hi (uint4), lo8 = e5m2((v - hi) * 2^12) (uint2), h8 = e5m2(v) (uint2; skipped for tensors stored without it)
__device__ __forceinline__ void split_f16f8x8(const float* w, uint4& hi, uint2& lo8, uint2& h8, bool want_h8 = true) {
  hi.x = pack_f16(w[0], w[1]); hi.y = pack_f16(w[2], w[3]); hi.z = pack_f16(w[4], w[5]); hi.w = pack_f16(w[6], w[7]);
  lo8.x = e5m2x4((w[0] - f16_lo_f(hi.x)) * MPN_F8_LO_SCALE, (w[1] - f16_hi_f(hi.x)) * MPN_F8_LO_SCALE,
                 (w[2] - f16_lo_f(hi.y)) * MPN_F8_LO_SCALE, (w[3] - f16_hi_f(hi.y)) * MPN_F8_LO_SCALE);
  lo8.y = e5m2x4((w[4] - f16_lo_f(hi.z)) * MPN_F8_LO_SCALE, (w[5] - f16_hi_f(hi.z)) * MPN_F8_LO_SCALE,
                 (w[6] - f16_lo_f(hi.w)) * MPN_F8_LO_SCALE, (w[7] - f16_hi_f(hi.w)) * MPN_F8_LO_SCALE);
  if (want_h8) {
    h8.x = e5m2x4(w[0], w[1], w[2], w[3]);
    h8.y = e5m2x4(w[4], w[5], w[6], w[7]);
  } else {
    h8.x = h8.y = 0u;
  }
}

// PAIR: the kernel runs as 2-CTA clusters (one TPC).  The two CTAs take two consecutive M tiles of the same Cout block; each
// loads its own activation box and HALF of the filter rows, the leader (cluster rank 0) issues one M = 256
// tcgen05.mma.cta_group::2 per K step for both, so every SM reads half the B operand from shared memory and fills half of
// it by TMA -- the single-CTA kernel is bound by shared-memory bandwidth (96 B/clk of MMA operand reads + the TMA fill
// against 128 B/clk), not by the tensor pipe.
template <int BN, int MODE, int STAGES, int EPI, bool PAIR, int NG = 2>
__global__ void __launch_bounds__(128 + 128 * NG, 1) conv_tc_kernel(const __grid_constant__ Maps maps, const TcParams P) {
  constexpr bool WIDE = NG == 4;             // 16 epilogue warps: four per TMEM lane quarter, each group owns every 4th 32-column chunk
  constexpr int NEPI = 4 * NG;               // epilogue warps
  constexpr int NTHREADS = 128 + 128 * NG;   // 4 control warps + the epilogue warps
  constexpr bool SPLIT = MODE == MODE_BF16X2;
  constexpr bool F8B = MODE == MODE_F16F8B;             // operand side: no h8 activation plane, fp16 weight residual
  constexpr bool F8C = MODE == MODE_F16F8C;             // operand side: no h8 activation plane, derived in shared memory
  constexpr bool F8 = MODE == MODE_F16F8 || F8B || F8C; // format side (epilogue, shortcut): fp16 hi + e5m2 lo8 (+ optional h8)
  // "plane units" of 128 bytes per row and K block: BF16X2 = hi + lo; F16F8 = hi (128 B) + lo8 (64 B) + h8 (64 B)
  constexpr int PLANES = MODE == MODE_BF16 ? 1 : 2;
  constexpr int NMAPS = F8 ? 3 : PLANES;
  constexpr int B_TILE_BYTES = (PAIR ? BN / 2 : BN) * BLOCK_K * 2;  // one 128-byte plane of the filter rows held by THIS CTA
  constexpr int A_STAGE = a_stage_bytes(MODE);                      // A planes of one K block
  constexpr int B_ROW = b_row_bytes(MODE);                          // B bytes per filter row and K block (all planes)
  constexpr int A_ROW_TX = (F8B || F8C) ? 192 : PLANES * 128;       // A bytes TMA delivers per box row
  constexpr int STAGE_BYTES = A_STAGE + (PAIR ? BN / 2 : BN) * B_ROW;
  constexpr int MMA_M = PAIR ? 2 * BLOCK_M : BLOCK_M;
  constexpr bool TMAEPI = EPI == EPI_TMA || EPI == EPI_TMA_RES;
  constexpr bool RESLD = EPI == EPI_TMA_RES;
  // TMA-store epilogue without a shortcut: the bulk-tensor stores of epilogue half h are issued by control warp 2 + h (a thread
  // that issues one stalls until the TMA unit accepts it -- ~500 cycles per chunk on the math threads' critical path otherwise)
  constexpr bool AGENT = EPI == EPI_TMA && MODE != MODE_F16F8C;
  constexpr int EPI_BYTES = WIDE ? 4 * WIDE_BOX_BYTES : NUM_EPI_WARPS * EPI_STAGE_BYTES + (RESLD ? 4 * RES_STAGE_BYTES : 0);
  static_assert(!WIDE || (RESLD && PAIR && F8 && !F8C), "the wide epilogue is the F16F8 shortcut epilogue (its agents are the F8C converter warps)");
  constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  // c = F32; a, b = BF16 (1) or, for F16F8, F16 (0);  | (N >> 3) << 17 per tile
  constexpr uint32_t IDESC_BASE = (1u << 4) | (F8 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(MMA_M >> 4) << 24);
  constexpr uint32_t IDESC_F8 = (1u << 4) | (1u << 7) | ((uint32_t)(MMA_M >> 4) << 24);  // a = E5M2 (1), b = E4M3 (0)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  auto res_full_bar = [&](int h, int b) { return bar_base + 8u * (2 * STAGES + 5 + 2 * h + b); };
  // F8C: afull = this CTA's activation planes of a stage have landed (the converter warps wait for it); conv = the h8 tiles of the
  // stage are complete in every CTA of the pair (on the leader; the MMA issuer waits for it in addition to `full`, which counts
  // the filter bytes there)
  auto afull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 9 + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (3 * STAGES + 9 + s); };
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + EPI_BYTES + 8 * (2 * STAGES + 4));
  uint8_t* epi_stage = smem_gen + STAGES * STAGE_BYTES;  // 1024-byte aligned (TMA-store swizzle pattern)
  float* bias_smem = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_k_iters = P.R * P.S * P.kb_per_tap;
  const int cta_rank = PAIR ? (int)cluster_ctarank() : 0;
  const bool leader = cta_rank == 0;
  const int tile0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;       // first (pair) tile of this CTA / cluster
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int p = 0; p < NMAPS; ++p) {
      prefetch_tmap(&maps.b[p]);
      for (int sg = 0; sg < P.nseg; ++sg) {
        if (!((F8B || F8C) && p == 2)) prefetch_tmap(&maps.a[p][sg]);
        if (TMAEPI && !(F8 && p == 2 && (P.flags & MPN_EPI_NO_H8))) prefetch_tmap(&maps.y[p][sg]);
      }
      if (P.tail_split > 1) prefetch_tmap(&maps.bs[p]);
      if (TMAEPI && (P.res_mma || RESLD) && p < 2) prefetch_tmap(&maps.r[p]);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), PAIR ? 2 * NEPI : NEPI);  // pair: the epilogue warps of both CTAs release the leader
    }
    if (RESLD) {
      for (int i = 0; i < 4; ++i) mbar_init(res_full_bar(i >> 1, i & 1), 1);
    }
    if (F8C) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(afull_bar(s), 1);
        mbar_init(conv_bar(s), PAIR ? 4 : 2);   // one arrival per converter warp of each CTA
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (TMAEPI && P.res_mma && P.rows < BLOCK_M) {
    // rows >= P.rows of a residual box are never written by TMA but are read (times a zero of I) by the MMA:
    // make sure no stale NaN/Inf bit pattern of an earlier kernel sits there
    uint4* z = reinterpret_cast<uint4*>(smem_gen);
    for (int i = threadIdx.x; i < STAGES * STAGE_BYTES / 16; i += NTHREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // the peer's barriers are initialised before anything is signalled across the pair
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  // Programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) overlapped the tail of the
  // previous kernel in the stream; its results are visible after the wait.  The next kernel may start its own prologue
  // as soon as every CTA of this grid got here.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0 && lane == 0) {
    // =============================== TMA producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    TRACE_INIT(0);
    const uint32_t lead_full0 = PAIR ? mapa_shared(full_bar(0), 0) : 0u;  // the leader's full barriers, as cluster addresses
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
      const bool tail = tc.ncols != BN;
      const int brows = PAIR ? tc.ncols / 2 : tc.ncols;                     // filter rows this CTA loads
      // bytes of one K block landing in this CTA; the leader's barrier of a pair expects both CTAs' bytes
      const Seg& g = P.seg[tc.seg];
#ifdef MPN_CONV_TRACE
      const uint32_t tx_bytes = (uint32_t)(g.rows * (A_ROW_TX - ((P.dbg_skip & 2) ? 64 : 0)) + brows * (B_ROW - ((P.dbg_skip & 1) ? 64 : 0))) * (PAIR ? 2u : 1u);
#else
      const uint32_t tx_bytes = (uint32_t)(g.rows * A_ROW_TX + brows * B_ROW) * (PAIR ? 2u : 1u);
#endif
      // narrow tail tiles are latency-bound on the ring: pack as many K blocks as fit into one stage (sub-blocks of
      // [A planes][B planes], the B planes ncols*128 bytes apart) so that twice the bytes are in flight
      const uint32_t bplane = tail ? (uint32_t)brows * 128u : (uint32_t)B_TILE_BYTES;
      const uint32_t sub_bytes = (uint32_t)A_STAGE + (bplane >> 7) * (uint32_t)B_ROW;
      const int kpack = (tail && !PAIR) ? (int)(STAGE_BYTES / sub_bytes) : 1;
      const int ow0 = tc.tw_i * g.TW, oh0 = tc.th_i * g.TH, n0 = tc.tn_i * g.TN, co0 = tc.co0 + cta_rank * brows;
      int u = 0, ki = 0;
      for (int tap = 0; tap < P.R * P.S; ++tap) {
        const int r = tap / P.S, s = tap - r * P.S;
        const int dh = r - P.pad, dw = s - P.pad;
        int ph = 0, hc, wc;
        if (P.stride == 1) {
          hc = oh0 + dh;
          wc = ow0 + dw;
        } else {  // stride 2: input row 2*oh + dh lives in phase view (dh & 1) at row oh + (dh - (dh&1))/2
          const int hp = dh & 1, wp = dw & 1;
          hc = oh0 + (dh - hp) / 2;
          wc = ow0 + (dw - wp) / 2;
          ph = hp * 2 + wp;
        }
        if ((P.phase_empty >> ph) & 1) {  // view without pixels: any fully out-of-range box reads zeros
          ph = 0;
          hc = 1 << 24;
        }
        if (P.nseg > 1) ph = tc.seg;   // multi-level launches are stride 1: the map slot is the segment
        for (int kb = 0; kb < P.kb_per_tap; ++kb) {
          if (u == 0) {
            const int group = min(kpack, num_k_iters - ki);
            mbar_wait(empty_bar(stage), phase ^ 1u);
            TRACE(1, tile, ki);
            if constexpr (F8C) {   // activation bytes -> this CTA's afull barrier, filter bytes -> the (leader's) full barrier
              mbar_expect_tx(afull_bar(stage), (uint32_t)(g.rows * A_ROW_TX) * (uint32_t)group);
              if (leader) mbar_expect_tx(full_bar(stage), (uint32_t)(brows * B_ROW) * (PAIR ? 2u : 1u) * (uint32_t)group);
            } else {
              if (leader) mbar_expect_tx(full_bar(stage), tx_bytes * (uint32_t)group);
            }
          }
          const uint32_t sa = smem_base + stage * STAGE_BYTES + u * sub_bytes;
          const uint32_t sb = sa + A_STAGE;
          const int kcol = tap * P.Cin + kb * BLOCK_K;
          if constexpr (PAIR) {
            const uint32_t lb = lead_full0 + 8u * stage;
            if constexpr (F8C) {
              // A: [hi 16 KB][lo8 8 KB][h8 8 KB, written by the converter warps]; B: [hi rows*128][lo8 rows*64][h8 rows*64]
              tma_load_4d(sa, &maps.a[0][ph], afull_bar(stage), kb * BLOCK_K, wc, hc, n0);
              tma_load_4d(sa + A_TILE_BYTES, &maps.a[1][ph], afull_bar(stage), kb * BLOCK_K, wc, hc, n0);
              tma_load_2d_pair(sb, tail ? &maps.bs[0] : &maps.b[0], lb, kcol, co0);
              if (P.w_merged) {
                tma_load_2d_pair(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], lb, 2 * kcol, co0);
              } else {
                tma_load_2d_pair(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], lb, kcol, co0);
                tma_load_2d_pair(sb + bplane + bplane / 2, tail ? &maps.bs[2] : &maps.b[2], lb, kcol, co0);
              }
            } else if constexpr (F8B) {
              // A: [hi 16 KB][lo8 8 KB]; B: [hi rows*128][lo16 rows*128][h8 rows*64]
              tma_load_4d_pair(sa, &maps.a[0][ph], lb, kb * BLOCK_K, wc, hc, n0);
              tma_load_4d_pair(sa + A_TILE_BYTES, &maps.a[1][ph], lb, kb * BLOCK_K, wc, hc, n0);
              tma_load_2d_pair(sb, tail ? &maps.bs[0] : &maps.b[0], lb, kcol, co0);
              tma_load_2d_pair(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], lb, kcol, co0);
              tma_load_2d_pair(sb + 2 * bplane, tail ? &maps.bs[2] : &maps.b[2], lb, kcol, co0);
            } else if constexpr (F8) {
              tma_load_4d_pair(sa, &maps.a[0][ph], lb, kb * BLOCK_K, wc, hc, n0);
              tma_load_4d_pair(sa + A_TILE_BYTES, &maps.a[1][ph], lb, kb * BLOCK_K, wc, hc, n0);
#ifdef MPN_CONV_TRACE
              if (!(P.dbg_skip & 2))
#endif
              tma_load_4d_pair(sa + A_TILE_BYTES + A_TILE_BYTES / 2, &maps.a[2][ph], lb, kb * BLOCK_K, wc, hc, n0);
              tma_load_2d_pair(sb, tail ? &maps.bs[0] : &maps.b[0], lb, kcol, co0);
              if (P.w_merged) {
                tma_load_2d_pair(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], lb, 2 * kcol, co0);
              } else {
              tma_load_2d_pair(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], lb, kcol, co0);
#ifdef MPN_CONV_TRACE
              if (!(P.dbg_skip & 1))
#endif
              tma_load_2d_pair(sb + bplane + bplane / 2, tail ? &maps.bs[2] : &maps.b[2], lb, kcol, co0);
              }
            } else {
#pragma unroll
              for (int p = 0; p < PLANES; ++p) {
                tma_load_4d_pair(sa + p * A_TILE_BYTES, &maps.a[p][ph], lb, kb * BLOCK_K, wc, hc, n0);
                tma_load_2d_pair(sb + p * bplane, tail ? &maps.bs[p] : &maps.b[p], lb, kcol, co0);
              }
            }
          } else if constexpr (F8C) {
            tma_load_4d(sa, &maps.a[0][ph], afull_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_4d(sa + A_TILE_BYTES, &maps.a[1][ph], afull_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_2d(sb, tail ? &maps.bs[0] : &maps.b[0], full_bar(stage), kcol, co0);
            if (P.w_merged) {
              tma_load_2d(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], full_bar(stage), 2 * kcol, co0);
            } else {
              tma_load_2d(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], full_bar(stage), kcol, co0);
              tma_load_2d(sb + bplane + bplane / 2, tail ? &maps.bs[2] : &maps.b[2], full_bar(stage), kcol, co0);
            }
          } else if constexpr (F8B) {
            tma_load_4d(sa, &maps.a[0][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_4d(sa + A_TILE_BYTES, &maps.a[1][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_2d(sb, tail ? &maps.bs[0] : &maps.b[0], full_bar(stage), kcol, co0);
            tma_load_2d(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], full_bar(stage), kcol, co0);
            tma_load_2d(sb + 2 * bplane, tail ? &maps.bs[2] : &maps.b[2], full_bar(stage), kcol, co0);
          } else if constexpr (F8) {
            // A: [hi 16 KB][lo8 8 KB][h8 8 KB]; B: [hi ncols*128][lo8 ncols*64][h8 ncols*64]
            tma_load_4d(sa, &maps.a[0][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_4d(sa + A_TILE_BYTES, &maps.a[1][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_4d(sa + A_TILE_BYTES + A_TILE_BYTES / 2, &maps.a[2][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
            tma_load_2d(sb, tail ? &maps.bs[0] : &maps.b[0], full_bar(stage), kcol, co0);
            if (P.w_merged) {
              tma_load_2d(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], full_bar(stage), 2 * kcol, co0);
            } else {
              tma_load_2d(sb + bplane, tail ? &maps.bs[1] : &maps.b[1], full_bar(stage), kcol, co0);
              tma_load_2d(sb + bplane + bplane / 2, tail ? &maps.bs[2] : &maps.b[2], full_bar(stage), kcol, co0);
            }
          } else {
#pragma unroll
            for (int p = 0; p < PLANES; ++p) {
              tma_load_4d(sa + p * A_TILE_BYTES, &maps.a[p][ph], full_bar(stage), kb * BLOCK_K, wc, hc, n0);
              tma_load_2d(sb + p * bplane, tail ? &maps.bs[p] : &maps.b[p], full_bar(stage), kcol, co0);
            }
          }
          ++ki;
          if (++u == kpack || ki == num_k_iters) {
            u = 0;
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
      if (TMAEPI && !PAIR && P.res_mma) {
#pragma unroll 1
        for (int p = 0; p < PLANES; ++p) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          mbar_expect_tx(full_bar(stage), (uint32_t)(2 * A_TILE_BYTES + (tc.ncols / 64) * P.rows * 128));
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE;
          tma_load_2d(sa, &maps.ident, full_bar(stage), 0, 0);
          tma_load_2d(sa + A_TILE_BYTES, &maps.ident, full_bar(stage), 64, 0);
          for (int j = 0; j < tc.ncols / 64; ++j)
            tma_load_4d(sb + j * (BLOCK_M * 128), &maps.r[p], full_bar(stage), co0 + 64 * j, ow0, oh0, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // =============================== MMA issuer (pair: the leader CTA only) ===============================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    TRACE_INIT(1);
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      TRACE(2, tile, 0);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
      const int ncols = decode_tile<BN, PAIR>(P, tile, 0).ncols;
      const uint32_t IDESC = IDESC_BASE | ((uint32_t)(ncols >> 3) << 17);
      const bool res_mma = TMAEPI && !PAIR && P.res_mma;
      const bool tail = ncols != BN;
      const uint32_t bplane = tail ? (uint32_t)(PAIR ? ncols / 2 : ncols) * 128u : (uint32_t)B_TILE_BYTES;
      const uint32_t sub_bytes = (uint32_t)A_STAGE + (bplane >> 7) * (uint32_t)B_ROW;
      const int kpack = (tail && !PAIR) ? (int)(STAGE_BYTES / sub_bytes) : 1;
      for (int ki = 0; ki < num_k_iters;) {
        mbar_wait(full_bar(stage), phase);
        TRACE(16, tile, ki);
        if constexpr (F8C) mbar_wait(conv_bar(stage), phase);
        TRACE(3, tile, ki);
        tc_fence_after();
        const int group = min(kpack, num_k_iters - ki);
        for (int u = 0; u < group; ++u, ++ki) {
          const uint32_t sa = smem_base + stage * STAGE_BYTES + u * sub_bytes;
          const uint32_t sb = sa + A_STAGE;
          auto mma16 = [&](uint64_t a, uint64_t b, uint32_t accum) {
            if (PAIR) umma_f16_pair(d_tmem, a, b, IDESC, accum);
            else umma_bf16(d_tmem, a, b, IDESC, accum);
          };
          if constexpr (F8) {
            const uint32_t idesc8 = IDESC_F8 | ((uint32_t)(ncols >> 3) << 17);
            auto mma8 = [&](uint64_t a, uint64_t b) {
              if (PAIR) umma_f8_pair(d_tmem, a, b, idesc8, 1u);
              else umma_f8(d_tmem, a, b, idesc8, 1u);
            };
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k)  // fp16 hi * fp16 hi, K = 16 per instruction
              mma16(make_sdesc(sa + k * 32), make_sdesc(sb + k * 32), (ki > 0 || k > 0) ? 1u : 0u);
            if constexpr (F8B) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 16; ++k)  // fp16 hi * fp16 weight residual
                mma16(make_sdesc(sa + k * 32), make_sdesc(sb + bplane + k * 32), 1u);
#pragma unroll
              for (int k = 0; k < BLOCK_K / 32; ++k)  // xlo8 * wh8, K = 32 per instruction
                mma8(make_sdesc64(sa + A_TILE_BYTES + k * 32), make_sdesc64(sb + 2 * bplane + k * 32));
            } else {
#pragma unroll
              for (int k = 0; k < BLOCK_K / 32; ++k) {  // fp8 cross terms, K = 32 per instruction
                // merged filter planes: one 128B-swizzled tile of 128-byte rows, lo8 in bytes 0..63 and h8 in bytes 64..127
                const uint64_t b_h8 = P.w_merged ? make_sdesc(sb + bplane + 64 + k * 32) : make_sdesc64(sb + bplane + bplane / 2 + k * 32);
                const uint64_t b_lo8 = P.w_merged ? make_sdesc(sb + bplane + k * 32) : make_sdesc64(sb + bplane + k * 32);
                mma8(make_sdesc64(sa + A_TILE_BYTES + k * 32), b_h8);                          // xlo8 * wh8
                mma8(make_sdesc64(sa + A_TILE_BYTES + A_TILE_BYTES / 2 + k * 32), b_lo8);      // xh8 * wlo8
              }
            }
          } else {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              const uint64_t a_hi = make_sdesc(sa + k * 32);
              const uint64_t b_hi = make_sdesc(sb + k * 32);
              mma16(a_hi, b_hi, (ki > 0 || k > 0) ? 1u : 0u);
              if (SPLIT) {
                const uint64_t a_lo = make_sdesc(sa + A_TILE_BYTES + k * 32);
                const uint64_t b_lo = make_sdesc(sb + bplane + k * 32);
                mma16(a_lo, b_hi, 1u);
                mma16(a_hi, b_lo, 1u);
              }
            }
          }
        }
        if (PAIR) {  // multicast: the barrier at the same offset in both CTAs of the pair
          umma_commit_pair(empty_bar(stage));
          if (ki == num_k_iters) umma_commit_pair(tfull_bar(acc));
        } else {
          umma_commit(empty_bar(stage));                                    // frees the smem stage when the MMAs retire
          if (ki == num_k_iters && !res_mma) umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (res_mma) {
        // D += I(128x128, K-major) * R(128 pixels x ncols channels, MN-major: 64-channel blocks 16 KB apart)
        const uint32_t idesc_mn = IDESC | (1u << 16);
#pragma unroll 1
        for (int p = 0; p < PLANES; ++p) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_STAGE;
#pragma unroll
          for (int k = 0; k < BLOCK_M / 16; ++k) {
            const uint64_t a_id = make_sdesc(sa + (k >> 2) * A_TILE_BYTES + (k & 3) * 32);
            const uint64_t b_rs = make_sdesc_mn(sb + k * 2048, BLOCK_M * 128, 1024);
            umma_bf16(d_tmem, a_id, b_rs, idesc_mn, 1u);
          }
          umma_commit(empty_bar(stage));
          if (p == PLANES - 1) umma_commit(tfull_bar(acc));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (AGENT && P.epi_agent && (warp == 2 || warp == 3)) {
    // =============================== TMA-store agents (EPI_TMA) ===============================
    // Warp 2 + h walks the (tile, chunk) sequence of epilogue half h: barrier 3 + h = "the staging box of the half is free" (this
    // warp arrives once its previous store has read it), barrier 5 + h = "the box is complete", then lane 0 stores the planes.
    const int half = warp - 2;
    const uint32_t stg_u32 = smem_base + STAGES * STAGE_BYTES + half * (4 * EPI_STAGE_BYTES);
    const bool want_h8 = !(P.flags & MPN_EPI_NO_H8);
    bool pending = false;
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
      const Seg& g = P.seg[tc.seg];
      const int ow0 = tc.tw_i * g.TW, oh0 = tc.th_i * g.TH, n0 = tc.tn_i * g.TN;
      for (int c0 = half * 32; c0 < tc.ncols; c0 += 64) {
        const int cbase = tc.co0 + c0;
        if (cbase >= P.Cout) break;
        named_bar_sync(3 + half, 160);
        named_bar_sync(5 + half, 160);
        if (lane == 0) {
          tma_store_4d(&maps.y[0][tc.seg], stg_u32, cbase, ow0, oh0, n0);
          if (SPLIT) tma_store_4d(&maps.y[1][tc.seg], stg_u32 + 8192, cbase, ow0, oh0, n0);
          if (F8) {
            tma_store_4d(&maps.y[1][tc.seg], stg_u32 + 8192, cbase, ow0, oh0, n0);
            if (want_h8) tma_store_4d(&maps.y[2][tc.seg], stg_u32 + 12288, cbase, ow0, oh0, n0);
          }
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          pending = true;
        }
      }
    }
    if (lane == 0 && pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (WIDE && (warp == 2 || warp == 3)) {
    // =============================== wide epilogue: TMA agents ===============================
    // Warp 2 + pg issues every bulk-tensor instruction of group pair pg (see the wide epilogue below): the shortcut box of chunk
    // n lands in buffer n & 1, the 256 math threads turn it into the output in place and meet this warp at named barrier 5 + pg,
    // lane 0 stores the buffer, waits until the store has read it and requests chunk n + 2 into it.
    const int pg = warp - 2;
    constexpr int LO_OFF = 16384;
    const uint32_t buf0 = smem_base + STAGES * STAGE_BYTES + (uint32_t)(pg * 2) * WIDE_BOX_BYTES;
    const uint32_t res_bytes = (uint32_t)P.rows * 192u;
    int pf_tile = tile0, pf_c0 = pg * 64;   // cursor: the next (tile, chunk) of this pair whose shortcut box is not requested yet
    uint32_t pf_n = 0;
    auto pf_issue = [&]() {
      while (pf_tile < P.total_tiles) {
        const TileCoord t = decode_tile<BN, PAIR>(P, pf_tile, cta_rank);
        if (pf_c0 < t.ncols && t.co0 + pf_c0 < P.Cout) {
          const uint32_t bar = res_full_bar(pg, (int)(pf_n & 1u)), dst = buf0 + (pf_n & 1u) * WIDE_BOX_BYTES;
          mbar_expect_tx(bar, res_bytes);
          tma_load_4d(dst, &maps.r[0], bar, t.co0 + pf_c0, t.tw_i * P.TW, t.th_i * P.TH, t.tn_i * P.TN);
          tma_load_4d(dst + LO_OFF, &maps.r[1], bar, t.co0 + pf_c0, t.tw_i * P.TW, t.th_i * P.TH, t.tn_i * P.TN);
          pf_c0 += 128;
          ++pf_n;
          return;
        }
        pf_tile += tile_step;   // no (further) chunk of this pair in that tile
        pf_c0 = pg * 64;
      }
    };
    if (lane == 0) { pf_issue(); pf_issue(); }
    uint32_t n = 0;
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
      const int ow0 = tc.tw_i * P.TW, oh0 = tc.th_i * P.TH, n0 = tc.tn_i * P.TN;
      for (int c0 = pg * 64; c0 < tc.ncols; c0 += 128) {
        const int cbase = tc.co0 + c0;
        if (cbase >= P.Cout) break;
        named_bar_sync(5 + pg, 288);
        if (lane == 0) {
          const uint32_t src = buf0 + (n & 1u) * WIDE_BOX_BYTES;
          tma_store_4d(&maps.y[0][0], src, cbase, ow0, oh0, n0);
          tma_store_4d(&maps.y[1][0], src + LO_OFF, cbase, ow0, oh0, n0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          pf_issue();   // chunk n + 2 into the buffer just read
        }
        ++n;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (F8C && (warp == 2 || warp == 3)) {
    // =============================== h8 converters (MODE_F16F8C) ===============================
    // Per K block: the 128 x 64 fp16 tile (128-byte rows, 128B swizzle) -> its e5m2 rounding as a 128 x 64 byte tile (64-byte rows,
    // 64B swizzle) behind the lo8 tile.  Lane = row: a quarter-warp's eight rows read / write eight different 16-byte columns.
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t conv0 = PAIR ? mapa_shared(conv_bar(0), 0) : conv_bar(0);
    TRACE_INIT(6 + (warp - 2));
#ifdef MPN_CONV_TRACE
    if (lane != 0) tr_.cap = 0;
#endif
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step) {
      const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
      const bool tail = tc.ncols != BN;
      const int brows = PAIR ? tc.ncols / 2 : tc.ncols;
      const uint32_t bplane = tail ? (uint32_t)brows * 128u : (uint32_t)B_TILE_BYTES;
      const uint32_t sub_bytes = (uint32_t)A_STAGE + (bplane >> 7) * (uint32_t)B_ROW;
      const int kpack = (tail && !PAIR) ? (int)(STAGE_BYTES / sub_bytes) : 1;
      for (int ki = 0; ki < num_k_iters;) {
        const int group = min(kpack, num_k_iters - ki);
        TRACE(12, tile, ki);
        mbar_wait(afull_bar(stage), phase);
        TRACE(13, tile, ki);
        for (int u = 0; u < group; ++u) {
          uint8_t* a_hi = smem_gen + stage * STAGE_BYTES + u * sub_bytes;
          uint8_t* a_h8 = a_hi + A_TILE_BYTES + A_TILE_BYTES / 2;
          // all sixteen loads of the thread's two rows first (the stores below may alias them for all the compiler knows)
          uint4 x[2][8];
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = (warp - 2) * 64 + i * 32 + lane;
#pragma unroll
            for (int c = 0; c < 8; ++c) x[i][c] = *reinterpret_cast<const uint4*>(a_hi + r * 128 + ((c ^ (r & 7)) << 4));
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int r = (warp - 2) * 64 + i * 32 + lane;
            uint8_t* dst = a_h8 + r * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 x0 = x[i][2 * j], x1 = x[i][2 * j + 1];
              uint4 o;
              o.x = h8_round4(x0.x, x0.y); o.y = h8_round4(x0.z, x0.w); o.z = h8_round4(x1.x, x1.y); o.w = h8_round4(x1.z, x1.w);
              *reinterpret_cast<uint4*>(dst + ((j ^ ((r >> 1) & 3)) << 4)) = o;
            }
          }
        }
        ki += group;
        TRACE(14, tile, ki);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) {
          if (PAIR) mbar_arrive_cluster(conv0 + 8u * stage);
          else mbar_arrive(conv_bar(stage));
        }
        TRACE(15, tile, ki);
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp >= 4) {
    // =============================== epilogue (8 warps) ===============================
    // Two warps per TMEM lane quarter (one per SM sub-partition pair): warp w owns quarter w&3 and the
    // 32-column chunks of parity (w-4)>>2, so each scheduler has two independent instruction streams.
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    if constexpr (WIDE) {
      // =============== wide epilogue: 4 column groups x 4 warps (one per TMEM lane quarter) ===============
      // The 1x1 convolutions with a shortcut (bottleneck conv3) are bound by their epilogue, and the epilogue by its TMA traffic and
      // the latency of the shortcut boxes (clock64 timelines, profiles/r02q and r02r): the unit is paced per box ROW (~2 cycles
      // whatever the row width) and a thread that issues a bulk-tensor instruction stalls until the unit accepts it.  Here
      //  * every sub-partition runs four epilogue warps working in 16-column register blocks (<= 96 registers at 640 threads);
      //  * two neighbouring groups share 64-channel boxes (128-byte fp16 rows, 64-byte e5m2 rows: half the rows of 32-channel boxes):
      //    group pair `pg` owns the 64-column chunks pg, pg + 2 of a tile, group 2 * pg + sub the 32-column half `sub` of each;
      //  * the shortcut box IS the output staging box: a thread adds its row's 32 values in place (same layout, no other thread
      //    touches them), so the pair's two buffers double-buffer the shortcut loads -- chunk n + 2 is requested as soon as the
      //    store of chunk n has read its buffer;
      //  * all bulk-tensor instructions of the pair are issued by an otherwise idle control warp (warp 2 + pg, below): the 256 math
      //    threads never wait on the TMA unit, they meet the agent at one named barrier per chunk.
      const int grp = half, pg = half >> 1, sub = half & 1;
      const int row = q * 32 + lane;
      const int sw7 = row & 7, sw3 = (row >> 1) & 3;   // 128B swizzle of the fp16 rows, 64B swizzle of the e5m2 rows
      constexpr int LO_OFF = 16384;                      // [hi: rows x 128 B][lo8: rows x 64 B]
      float* bias_g = bias_smem + grp * 64;
      const uint32_t lead_tempty0 = mapa_shared(tempty_bar(0), 0);
      const float as = P.acc_scale;
      TRACE_INIT(2 + grp);
#ifdef MPN_CONV_TRACE
      if (!(q == 0 && lane == 0)) tr_.cap = 0;
#endif
      uint32_t res_n = 0;   // chunks of this pair consumed so far: buffer = res_n & 1, mbarrier parity = (res_n >> 1) & 1
      int it = 0;
      for (int tile = tile0; tile < P.total_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
        const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
        const int co0 = tc.co0, ncols = tc.ncols;
        if (P.bias) {   // the group's (up to) two half-chunks x 32 bias values of this tile
          const int tq = q * 32 + lane;
          if (tq < 64) {
            const int col = grp * 32 + 128 * (tq >> 5) + (tq & 31);
            bias_g[tq] = (col < ncols && co0 + col < P.Cout) ? __ldg(P.bias + co0 + col) : 0.f;
          }
          named_bar_sync(1 + grp, 128);
        }
        TRACE(4, tile, 0);
        mbar_wait(tfull_bar(acc), acc_phase);
        TRACE(5, tile, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = pg * 64; c0 < ncols; c0 += 128) {
          if (co0 + c0 >= P.Cout) break;  // uniform over the pair
          const int b = (int)(res_n & 1u);
          uint8_t* buf = epi_stage + (pg * 2 + b) * WIDE_BOX_BYTES;
          mbar_wait(res_full_bar(pg, b), (res_n >> 1) & 1u);
          TRACE(6, tile, c0);
          ++res_n;
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            uint32_t raw[16];
            TMEM_LD_32x32b_X16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0 + sub * 32 + 16 * s2), raw);
            uint4* ph0 = reinterpret_cast<uint4*>(buf + row * 128 + (((sub * 4 + 2 * s2) ^ sw7) << 4));
            uint4* ph1 = reinterpret_cast<uint4*>(buf + row * 128 + (((sub * 4 + 2 * s2 + 1) ^ sw7) << 4));
            uint4* pl = reinterpret_cast<uint4*>(buf + LO_OFF + row * 64 + (((sub * 2 + s2) ^ sw3) << 4));
            const uint4 rh0 = *ph0, rh1 = *ph1, rl = *pl;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float v[16];
            if (P.bias) {
              const float4* b4 = reinterpret_cast<const float4*>(bias_g + (c0 >> 7) * 32 + 16 * s2);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 t = b4[j];
                v[4 * j] = fmaf(__uint_as_float(raw[4 * j]), as, t.x); v[4 * j + 1] = fmaf(__uint_as_float(raw[4 * j + 1]), as, t.y);
                v[4 * j + 2] = fmaf(__uint_as_float(raw[4 * j + 2]), as, t.z); v[4 * j + 3] = fmaf(__uint_as_float(raw[4 * j + 3]), as, t.w);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) * as;
            }
            add_f16x8_reg(v, rh0);
            add_f16x8_reg(v + 8, rh1);
            add_e5m2x4_reg(v, rl.x); add_e5m2x4_reg(v + 4, rl.y); add_e5m2x4_reg(v + 8, rl.z); add_e5m2x4_reg(v + 12, rl.w);
            if (P.flags & MPN_EPI_RELU) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (P.flags & MPN_EPI_SIGMOID) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
            }
            uint4 h0, h1;
            uint2 l0, l1, unused;
            split_f16f8x8(v, h0, l0, unused, false);
            split_f16f8x8(v + 8, h1, l1, unused, false);
            *ph0 = h0;   // in place: this thread's own 16 values of the shortcut box become the output
            *ph1 = h1;
            *pl = make_uint4(l0.x, l0.y, l1.x, l1.y);
          }
          TRACE(7, tile, c0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          named_bar_sync(5 + pg, 288);   // buffer complete: the pair's agent warp stores it and refills it with chunk n + 2
          TRACE(10, tile, c0);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(lead_tempty0 + 8u * acc);  // the leader's MMA issuer waits for both CTAs' epilogues
        TRACE(11, tile, 0);
      }
    } else {
    const bool want_h8 = !(P.flags & MPN_EPI_NO_H8);  // F16F8 outputs: the e5m2 copy plane is only stored for tensors a 3x3 conv reads
    const int rep = P.out_rep, OHr = P.OH * rep, OWr = P.OW * rep;
    const long long nstride = P.out_nstride > 0 ? P.out_nstride
                              : (P.out_mode == MPN_OUT_F32_NCHW ? (long long)P.Cout * OHr * OWr : (long long)OHr * OWr * P.out_cstride);
    int it = 0;
    bool store_pending = false;  // EPI_TMA: a bulk store of this half may still be reading its staging box
    TRACE_INIT(2 + half);
#ifdef MPN_CONV_TRACE
    if (!(q == 0 && lane == 0)) tr_.cap = 0;
#endif
    const uint32_t lead_tempty0 = PAIR ? mapa_shared(tempty_bar(0), 0) : 0u;
    // RESLD: the shortcut box of 32-channel chunk n + 1 of this half (next chunk of the tile, or the first one of the CTA's next
    // tile) is requested by TMA when chunk n starts, into the buffer chunk n - 1 was read from (every thread of the half has
    // passed that chunk's barriers), so a whole chunk period covers the load latency and no LSU load touches the shortcut.
    uint32_t res_n = 0;  // chunks of this half consumed so far: buffer = res_n & 1, mbarrier parity = (res_n >> 1) & 1
    // up_tma: the box holds the half-resolution source pixels of the tile (fpn.py:84-95, exact 2x nearest upsample): output
    // pixel (th, tw) of the tile adds source row (th/2) * (TW/2) + tw/2 of the box (tile origins and sizes are even)
    const int src_rows = P.up_tma ? (P.TW >> 1) * (P.TH >> 1) * P.TN : P.rows;
    const uint32_t res_bytes = (uint32_t)src_rows * (SPLIT ? 128u : F8 ? 96u : 64u);
    auto res_issue = [&](uint32_t n, int cb, int w0, int h0, int i0) {
      const uint32_t bar = res_full_bar(half, (int)(n & 1u));
      const uint32_t dst = smem_base + STAGES * STAGE_BYTES + NUM_EPI_WARPS * EPI_STAGE_BYTES + (uint32_t)(half * 2 + (int)(n & 1u)) * RES_STAGE_BYTES;
      if (P.up_tma) { w0 >>= 1; h0 >>= 1; }
      mbar_expect_tx(bar, res_bytes);
      tma_load_4d(dst, &maps.r[0], bar, cb, w0, h0, i0);
      if (SPLIT || F8) tma_load_4d(dst + 8192, &maps.r[1], bar, cb, w0, h0, i0);
    };
    if (RESLD && q == 0 && lane == 0 && tile0 < P.total_tiles) {
      const TileCoord t0 = decode_tile<BN, PAIR>(P, tile0, cta_rank);
      res_issue(0u, t0.co0 + half * 32, t0.tw_i * P.TW, t0.th_i * P.TH, t0.tn_i * P.TN);
    }
    for (int tile = tile0; tile < P.total_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
      const TileCoord tc = decode_tile<BN, PAIR>(P, tile, cta_rank);
      const int tw_i = tc.tw_i, th_i = tc.th_i, tn_i = tc.tn_i, co0 = tc.co0, ncols = tc.ncols;
      if constexpr (TMAEPI) {
        // ---- thread = pixel row (the TMEM lane): 32 channels per step -> bf16 hi/lo -> 64B-swizzled box in smem -> TMA store.
        // The two halves (4 warps each) own separate 16 KB staging boxes and alternate over the 32-channel blocks.
        const int row = q * 32 + lane;
        uint8_t* stg = epi_stage + half * (4 * EPI_STAGE_BYTES);       // [hi: rows x 64 B][lo: rows x 64 B] at +8192
        const uint32_t stg_u32 = smem_base + STAGES * STAGE_BYTES + half * (4 * EPI_STAGE_BYTES);
        const int sw = (row >> 1) & 3;                                 // CU_TENSOR_MAP_SWIZZLE_64B: chunk ^= (byte >> 7) & 3
        int rrow = row;                                                // row of the shortcut / upsample-source box this pixel adds
        if (RESLD && P.up_tma) {
          const int tw2 = row % P.TW, th2 = (row / P.TW) % P.TH, tn2 = row / (P.TW * P.TH);
          rrow = row < P.rows ? (tn2 * (P.TH >> 1) + (th2 >> 1)) * (P.TW >> 1) + (tw2 >> 1) : 0;
        }
        const int rsw = (rrow >> 1) & 3;
        const bool issuer = (q == 0 && lane == 0);
        const Seg& g = P.seg[tc.seg];
        const int ow0 = tw_i * g.TW, oh0 = th_i * g.TH, n0 = tn_i * g.TN;
        if (RESLD && issuer && P.res_pf && tile + tile_step < P.total_tiles) {
          // the shortcut boxes are requested one chunk ahead only (two buffers per half): L2 is asked for the next tile's boxes a
          // whole tile ahead, so that those requests find their lines there instead of in HBM
          const TileCoord t2 = decode_tile<BN, PAIR>(P, tile + tile_step, cta_rank);
          int w2 = t2.tw_i * P.TW, h2 = t2.th_i * P.TH;
          if (P.up_tma) { w2 >>= 1; h2 >>= 1; }
          for (int c2 = half * 32; c2 < t2.ncols && t2.co0 + c2 < P.Cout; c2 += 64) {
            tma_prefetch_l2_4d(&maps.r[0], t2.co0 + c2, w2, h2, t2.tn_i * P.TN);
            if (SPLIT || F8) tma_prefetch_l2_4d(&maps.r[1], t2.co0 + c2, w2, h2, t2.tn_i * P.TN);
          }
        }
        long long res_off = 0;
        const bool res_lsu = !RESLD && P.res_cstride > 0 && !P.res_mma;
        if (res_lsu) {
          const int tw2 = row % P.TW, th2 = (row / P.TW) % P.TH, tn2 = row / (P.TW * P.TH);
          const int ow2 = ow0 + tw2, oh2 = oh0 + th2, n2 = n0 + tn2;
          const bool ok = row < P.rows && ow2 < P.OW && oh2 < P.OH && n2 < P.N;
          res_off = ok ? ((long long)(n2 * P.OH + oh2) * P.OW + ow2) * P.res_cstride : 0;  // invalid rows read pixel 0 (never stored)
        }
        // phase-class addends: element offset of this row's (low-resolution pixel, class slice) in each gathered tensor
        long long gat_off[2] = {0, 0};
        if (!RESLD && P.gat_n > 0) {
          const int tw2 = row % P.TW, th2 = (row / P.TW) % P.TH, tn2 = row / (P.TW * P.TH);
          const int ow2 = ow0 + tw2, oh2 = oh0 + th2, n2 = n0 + tn2;
          const bool ok = row < P.rows && ow2 < P.OW && oh2 < P.OH && n2 < P.N;
#pragma unroll
          for (int gi = 0; gi < 2; ++gi) {
            if (gi < P.gat_n && ok) {
              const int sh = P.gat_shift[gi], m = (1 << sh) - 1;
              const int a = oh2 & m, b = ow2 & m;
              const int cls = (a == 0 ? 0 : a == m ? 2 : 1) * 3 + (b == 0 ? 0 : b == m ? 2 : 1);
              gat_off[gi] = ((long long)(n2 * P.gat_h[gi] + (oh2 >> sh)) * P.gat_w[gi] + (ow2 >> sh)) * P.gat_cstride[gi] + (long long)cls * P.Cout;
            }
          }
        }
        // The shortcut operand does not depend on the accumulator: the 64 (+64 / +32) bytes of a row's next 32-channel chunk are
        // requested one chunk ahead -- the first one before the accumulator is even complete -- so the global-load latency
        // hides behind the previous chunk's convert / stage / store instead of stalling every chunk (4 per tile and warp).
        uint4 nh[4], nl[4];
        auto load_res = [&](int cb) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            nh[i] = __ldg(reinterpret_cast<const uint4*>(P.res_hi + res_off + cb) + i);
            if (SPLIT) nl[i] = __ldg(reinterpret_cast<const uint4*>(P.res_lo + res_off + cb) + i);
            if (F8 && i < 2) nl[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(P.res_lo) + res_off + cb) + i);
          }
        };
        if (res_lsu && half * 32 < ncols && co0 + half * 32 < P.Cout) load_res(co0 + half * 32);
        // The bias of the tile's columns goes through shared memory once per tile: with ~30 KB of L1 left beside the 225 KB
        // carve-out the per-chunk __ldg of 128 bytes missed to L2 every time (ncu r01v: 4-8 % of the epilogue warps' samples
        // on the dependent FMA).  Each half keeps the 4 x 32 values of its own chunks, so the halves stay independent; every
        // thread of the half is past the barriers of the previous tile's last chunk, i.e. done reading the old values.
        float* bias_h = bias_smem + half * 128;
        const bool bias_staged = P.bias != nullptr && BN <= 256;
        if (bias_staged) {
            const int tq = q * 32 + lane;                       // 0..127 within the half
            const int col = half * 32 + 64 * (tq >> 5) + (tq & 31);
            bias_h[tq] = (col < ncols && co0 + col < P.Cout) ? __ldg(P.bias + co0 + col) : 0.f;
            named_bar_sync(1 + half, 128);
        }
        TRACE(4, tile, 0);
        mbar_wait(tfull_bar(acc), acc_phase);
        TRACE(5, tile, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = half * 32; c0 < ncols; c0 += 64) {
          const int cbase = co0 + c0;
          if (cbase >= P.Cout) break;  // uniform over the half
          uint32_t raw[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
          TMEM_LD_32x32b_X32(taddr, raw);
          uint4 rh[4], rl[4];
          if (res_lsu) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { rh[i] = nh[i]; rl[i] = nl[i]; }
            if (c0 + 64 < ncols && cbase + 64 < P.Cout) load_res(cbase + 64);
          }
          if constexpr (RESLD) {
            if (issuer) {
              if (c0 + 64 < ncols && cbase + 64 < P.Cout) {
                res_issue(res_n + 1u, cbase + 64, ow0, oh0, n0);
              } else if (tile + tile_step < P.total_tiles) {
                const TileCoord t2 = decode_tile<BN, PAIR>(P, tile + tile_step, cta_rank);
                res_issue(res_n + 1u, t2.co0 + half * 32, t2.tw_i * P.TW, t2.th_i * P.TH, t2.tn_i * P.TN);
              }
            }
            mbar_wait(res_full_bar(half, (int)(res_n & 1u)), (res_n >> 1) & 1u);
            const uint8_t* rs = epi_stage + NUM_EPI_WARPS * EPI_STAGE_BYTES + (half * 2 + (int)(res_n & 1u)) * RES_STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              rh[i] = *reinterpret_cast<const uint4*>(rs + rrow * 64 + ((i ^ rsw) << 4));
              if (SPLIT) rl[i] = *reinterpret_cast<const uint4*>(rs + 8192 + rrow * 64 + ((i ^ rsw) << 4));
              if (F8 && i < 2) rl[i] = *reinterpret_cast<const uint4*>(rs + 8192 + rrow * 32 + (i << 4));
            }
            ++res_n;
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          // F16F8: the accumulator carries the filter prescale 2^k; acc_scale = 2^-k is exact, so folding it into the bias add
          // (one FMA) gives the same bits as multiply-then-add
          const bool fused_bias = F8 && P.bias && !P.scale;
          if (F8 && !fused_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= P.acc_scale;
          }
          if (P.scale) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(P.scale + cbase) + j);
              v[4 * j] *= t.x; v[4 * j + 1] *= t.y; v[4 * j + 2] *= t.z; v[4 * j + 3] *= t.w;
            }
          }
          const float4* bias4 = reinterpret_cast<const float4*>(bias_h + (c0 >> 6) * 32);  // chunk (c0 - half*32) / 64 of this half
          if (fused_bias) {
            const float as = P.acc_scale;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = bias4[j];
              v[4 * j] = fmaf(v[4 * j], as, t.x); v[4 * j + 1] = fmaf(v[4 * j + 1], as, t.y);
              v[4 * j + 2] = fmaf(v[4 * j + 2], as, t.z); v[4 * j + 3] = fmaf(v[4 * j + 3], as, t.w);
            }
          } else if (P.bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 t = bias4[j];
              v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
          }
          if (!RESLD && P.gat_n > 0) {   // += G_g[low-res pixel][class * Cout + channel], straight from L2 (the MMA phase of these
                                         // 3x3 convolutions is long: the epilogue has slack)
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
              if (gi >= P.gat_n) break;
              if constexpr (F8) {
                const uint4* gh = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(P.gat_hi[gi]) + gat_off[gi] + cbase);
                const uint4* gl = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(P.gat_lo[gi]) + gat_off[gi] + cbase);
#pragma unroll
                for (int i = 0; i < 4; ++i) add_f16x8_reg(v + 8 * i, __ldg(gh + i));
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                  const uint4 t = __ldg(gl + i);
                  add_e5m2x4_reg(v + 16 * i, t.x); add_e5m2x4_reg(v + 16 * i + 4, t.y);
                  add_e5m2x4_reg(v + 16 * i + 8, t.z); add_e5m2x4_reg(v + 16 * i + 12, t.w);
                }
              } else {
                const uint4* gh = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(P.gat_hi[gi]) + gat_off[gi] + cbase);
#pragma unroll
                for (int i = 0; i < 4; ++i) add_bf16x8_reg(v + 8 * i, __ldg(gh + i));
                if (SPLIT) {
                  const uint4* gl = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(P.gat_lo[gi]) + gat_off[gi] + cbase);
#pragma unroll
                  for (int i = 0; i < 4; ++i) add_bf16x8_reg(v + 8 * i, __ldg(gl + i));
                }
              }
            }
          }
          if (res_lsu || RESLD) {
            if constexpr (F8) {
#pragma unroll
              for (int i = 0; i < 4; ++i) add_f16x8_reg(v + 8 * i, rh[i]);
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                add_e5m2x4_reg(v + 16 * i, rl[i].x); add_e5m2x4_reg(v + 16 * i + 4, rl[i].y);
                add_e5m2x4_reg(v + 16 * i + 8, rl[i].z); add_e5m2x4_reg(v + 16 * i + 12, rl[i].w);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                add_bf16x8_reg(v + 8 * i, rh[i]);
                if (SPLIT) add_bf16x8_reg(v + 8 * i, rl[i]);
              }
            }
          }
          if (P.flags & MPN_EPI_RELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (P.flags & MPN_EPI_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
          }
          uint4 hi4[4], lo4[4];
          uint2 l8[4], h8[4];
          if constexpr (F8) {
#pragma unroll
            for (int i = 0; i < 4; ++i) split_f16f8x8(v + 8 * i, hi4[i], l8[i], h8[i], want_h8);
          } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float* w = v + 8 * i;
            hi4[i].x = pack_bf16(w[0], w[1]); hi4[i].y = pack_bf16(w[2], w[3]);
            hi4[i].z = pack_bf16(w[4], w[5]); hi4[i].w = pack_bf16(w[6], w[7]);
            if (SPLIT) {
              lo4[i].x = pack_bf16(w[0] - bf16_lo_f(hi4[i].x), w[1] - bf16_hi_f(hi4[i].x));
              lo4[i].y = pack_bf16(w[2] - bf16_lo_f(hi4[i].y), w[3] - bf16_hi_f(hi4[i].y));
              lo4[i].z = pack_bf16(w[4] - bf16_lo_f(hi4[i].z), w[5] - bf16_hi_f(hi4[i].z));
              lo4[i].w = pack_bf16(w[6] - bf16_lo_f(hi4[i].w), w[7] - bf16_hi_f(hi4[i].w));
            }
          }
          }
          // the previous store of this half must have finished READING the staging box before it is overwritten
          TRACE(7, tile, c0);
          if (AGENT && P.epi_agent) {
            named_bar_sync(3 + half, 160);   // the agent warp joins once its previous store has read the box
          } else if (store_pending) {
            if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            TRACE(8, tile, c0);
            named_bar_sync(1 + half, 128);
          }
          TRACE(9, tile, c0);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            *reinterpret_cast<uint4*>(stg + row * 64 + ((i ^ sw) << 4)) = hi4[i];
            if (SPLIT) *reinterpret_cast<uint4*>(stg + 8192 + row * 64 + ((i ^ sw) << 4)) = lo4[i];
          }
          if constexpr (F8) {  // byte planes: [lo8: rows x 32 B] at +8192, [h8: rows x 32 B] at +12288, unswizzled boxes
            *reinterpret_cast<uint4*>(stg + 8192 + row * 32) = make_uint4(l8[0].x, l8[0].y, l8[1].x, l8[1].y);
            *reinterpret_cast<uint4*>(stg + 8192 + row * 32 + 16) = make_uint4(l8[2].x, l8[2].y, l8[3].x, l8[3].y);
            if (want_h8) {
              *reinterpret_cast<uint4*>(stg + 12288 + row * 32) = make_uint4(h8[0].x, h8[0].y, h8[1].x, h8[1].y);
              *reinterpret_cast<uint4*>(stg + 12288 + row * 32 + 16) = make_uint4(h8[2].x, h8[2].y, h8[3].x, h8[3].y);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          if (AGENT && P.epi_agent) {
            named_bar_sync(5 + half, 160);   // box complete: the agent stores it
            TRACE(10, tile, c0);
            continue;
          }
          named_bar_sync(1 + half, 128);
          TRACE(10, tile, c0);
          if (issuer) {
            tma_store_4d(&maps.y[0][tc.seg], stg_u32, cbase, ow0, oh0, n0);
            if (SPLIT) tma_store_4d(&maps.y[1][tc.seg], stg_u32 + 8192, cbase, ow0, oh0, n0);
            if (F8) {
              tma_store_4d(&maps.y[1][tc.seg], stg_u32 + 8192, cbase, ow0, oh0, n0);
              if (want_h8) tma_store_4d(&maps.y[2][tc.seg], stg_u32 + 12288, cbase, ow0, oh0, n0);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          TRACE(17, tile, c0);
          store_pending = true;
        }
      } else if constexpr (EPI == EPI_LSU) {
        // ---- per-tile row bookkeeping for the coalesced phase: lane serves rows (lane>>2) + 8*i, i = 0..3
        long long obase[4];  // destination element offset of (n, oh*rep, ow*rep, out_coffset)
        int pixv[4];         // flat output pixel index, or -1 if the row is outside the tensor
        int upv[4];          // flat pixel index in the upsample source
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int trow = q * 32 + (lane >> 2) + 8 * i;
          const int tw2 = trow % P.TW, th2 = (trow / P.TW) % P.TH, tn2 = trow / (P.TW * P.TH);
          const int ow2 = tw_i * P.TW + tw2, oh2 = th_i * P.TH + th2, n2 = tn_i * P.TN + tn2;
          const bool ok = trow < P.rows && ow2 < P.OW && oh2 < P.OH && n2 < P.N;
          pixv[i] = ok ? (n2 * P.OH + oh2) * P.OW + ow2 : -1;
          obase[i] = (long long)n2 * nstride + ((long long)(oh2 * rep) * OWr + ow2 * rep) * P.out_cstride + P.out_coffset;
          upv[i] = 0;
          if (ok && P.up_cstride > 0)
            upv[i] = (n2 * P.up_h + mpn_nearest_src(oh2, P.up_h, P.OH)) * P.up_w + mpn_nearest_src(ow2, P.up_w, P.OW);
        }
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        float* stg = reinterpret_cast<float*>(epi_stage + (warp - 4) * EPI_STAGE_BYTES);
        const int g = lane & 3;  // 8-channel group of this lane in the coalesced phase
#pragma unroll 1
        for (int c0 = half * 32; c0 < ncols; c0 += 64) {
          const int cbase = co0 + c0;
          if (cbase >= P.Cout) break;  // warp-uniform
          uint32_t raw[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
          TMEM_LD_32x32b_X32(taddr, raw);
          // prefetch the residual / upsample operands of this chunk while the TMEM load is in flight
          const int ch = cbase + 8 * g;
          uint4 rh[4], rl[4], uh[4], ul[4];
          if (P.res_cstride > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const long long o = (long long)(pixv[i] < 0 ? 0 : pixv[i]) * P.res_cstride + ch;
              rh[i] = __ldg(reinterpret_cast<const uint4*>(P.res_hi + o));
              if (SPLIT) rl[i] = __ldg(reinterpret_cast<const uint4*>(P.res_lo + o));
              if (F8) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(P.res_lo) + o));
                rl[i].x = t.x; rl[i].y = t.y;
              }
            }
          }
          if (P.up_cstride > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const long long o = (long long)upv[i] * P.up_cstride + ch;
              uh[i] = __ldg(reinterpret_cast<const uint4*>(P.up_hi + o));
              if (SPLIT) ul[i] = __ldg(reinterpret_cast<const uint4*>(P.up_lo + o));
              if (F8) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const unsigned char*>(P.up_lo) + o));
                ul[i].x = t.x; ul[i].y = t.y;
              }
            }
          }
          float4 sc0 = make_float4(1.f, 1.f, 1.f, 1.f), sc1 = sc0, bi0 = make_float4(0.f, 0.f, 0.f, 0.f), bi1 = bi0;
          if (P.scale) { sc0 = __ldg(reinterpret_cast<const float4*>(P.scale + ch)); sc1 = __ldg(reinterpret_cast<const float4*>(P.scale + ch + 4)); }
          if (P.bias) { bi0 = __ldg(reinterpret_cast<const float4*>(P.bias + ch)); bi1 = __ldg(reinterpret_cast<const float4*>(P.bias + ch + 4)); }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // thread-per-row accumulators -> swizzled staging -> 4 lanes per row (coalesced 64-byte row segments)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) =
                make_uint4(raw[4 * c], raw[4 * c + 1], raw[4 * c + 2], raw[4 * c + 3]);
          __syncwarp();
          const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
          const float biv[8] = {bi0.x, bi0.y, bi0.z, bi0.w, bi1.x, bi1.y, bi1.z, bi1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = (lane >> 2) + 8 * i;
            const uint4 f0 = *reinterpret_cast<const uint4*>(stg + rr * 32 + (((2 * g) ^ (rr & 7)) << 2));
            const uint4 f1 = *reinterpret_cast<const uint4*>(stg + rr * 32 + (((2 * g + 1) ^ (rr & 7)) << 2));
            float v[8] = {__uint_as_float(f0.x), __uint_as_float(f0.y), __uint_as_float(f0.z), __uint_as_float(f0.w),
                          __uint_as_float(f1.x), __uint_as_float(f1.y), __uint_as_float(f1.z), __uint_as_float(f1.w)};
            if (F8) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= P.acc_scale;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], scv[j], biv[j]);
            if constexpr (F8) {
              if (P.res_cstride > 0) { add_f16x8_reg(v, rh[i]); add_e5m2x4_reg(v, rl[i].x); add_e5m2x4_reg(v + 4, rl[i].y); }
              if (P.up_cstride > 0) { add_f16x8_reg(v, uh[i]); add_e5m2x4_reg(v, ul[i].x); add_e5m2x4_reg(v + 4, ul[i].y); }
            } else {
              if (P.res_cstride > 0) {
                add_bf16x8_reg(v, rh[i]);
                if (SPLIT) add_bf16x8_reg(v, rl[i]);
              }
              if (P.up_cstride > 0) {
                add_bf16x8_reg(v, uh[i]);
                if (SPLIT) add_bf16x8_reg(v, ul[i]);
              }
            }
            if (P.flags & MPN_EPI_RELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (P.flags & MPN_EPI_SIGMOID) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
            }
            uint4 hi4, lo4;
            uint2 l8, h8;
            if constexpr (F8) {
              split_f16f8x8(v, hi4, l8, h8, want_h8);
              if (pixv[i] >= 0) {
                unsigned char* ylo = reinterpret_cast<unsigned char*>(P.y_lo);
                for (int ry = 0; ry < rep; ++ry)
                  for (int rx = 0; rx < rep; ++rx) {
                    const long long o = obase[i] + ((long long)ry * OWr + rx) * P.out_cstride + ch;
                    *reinterpret_cast<uint4*>((__half*)P.y_hi + o) = hi4;
                    *reinterpret_cast<uint2*>(ylo + o) = l8;
                    if (want_h8) *reinterpret_cast<uint2*>(ylo + P.y_plane + o) = h8;
                  }
              }
              continue;
            }
            hi4.x = pack_bf16(v[0], v[1]); hi4.y = pack_bf16(v[2], v[3]); hi4.z = pack_bf16(v[4], v[5]); hi4.w = pack_bf16(v[6], v[7]);
            if (SPLIT) {
              lo4.x = pack_bf16(v[0] - bf16_lo_f(hi4.x), v[1] - bf16_hi_f(hi4.x));
              lo4.y = pack_bf16(v[2] - bf16_lo_f(hi4.y), v[3] - bf16_hi_f(hi4.y));
              lo4.z = pack_bf16(v[4] - bf16_lo_f(hi4.z), v[5] - bf16_hi_f(hi4.z));
              lo4.w = pack_bf16(v[6] - bf16_lo_f(hi4.w), v[7] - bf16_hi_f(hi4.w));
            }
            if (pixv[i] >= 0) {
              if (rep == 1) {
                *reinterpret_cast<uint4*>((__nv_bfloat16*)P.y_hi + obase[i] + ch) = hi4;
                if (SPLIT) *reinterpret_cast<uint4*>((__nv_bfloat16*)P.y_lo + obase[i] + ch) = lo4;
              } else {
                for (int ry = 0; ry < rep; ++ry)
                  for (int rx = 0; rx < rep; ++rx) {
                    const long long o = obase[i] + ((long long)ry * OWr + rx) * P.out_cstride + ch;
                    *reinterpret_cast<uint4*>((__nv_bfloat16*)P.y_hi + o) = hi4;
                    if (SPLIT) *reinterpret_cast<uint4*>((__nv_bfloat16*)P.y_lo + o) = lo4;
                  }
              }
            }
          }
          __syncwarp();  // the staging tile is rewritten by the next chunk
        }
      } else {
        // ---- fp32 outputs (heads: Cout <= 64): thread-per-row, 16 columns at a time
        const int row = q * 32 + lane;
        const Seg& g = P.seg[tc.seg];
        const int tw = row % g.TW, th = (row / g.TW) % g.TH, tn = row / (g.TW * g.TH);
        const int ow = tw_i * g.TW + tw, oh = th_i * g.TH + th, n = tn_i * g.TN + tn;
        const bool valid = row < g.rows && ow < g.OW && oh < g.OH && n < g.N;
        const int OHr = g.OH * rep, OWr = g.OW * rep;   // shadows the launch-wide values: per segment
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = half * 16; c0 < ncols; c0 += 32) {
          const int cbase = co0 + c0;
          if (cbase >= P.Cout) break;  // warp-uniform
          uint32_t raw[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c0);
          TMEM_LD_32x32b_X16(taddr, raw);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (!valid) continue;
          const int nc = min(16, P.Cout - cbase);
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            v[j] = __uint_as_float(raw[j]);
            if (F8) v[j] *= P.acc_scale;
            if (j < nc) {
              if (P.scale) v[j] *= __ldg(P.scale + cbase + j);
              if (P.bias) v[j] += __ldg(P.bias + cbase + j);
            }
            if (P.flags & MPN_EPI_RELU) v[j] = fmaxf(v[j], 0.f);
            if (P.flags & MPN_EPI_SIGMOID) v[j] = 1.f / (1.f + expf(-v[j]));
          }
          for (int ry = 0; ry < rep; ++ry)
            for (int rx = 0; rx < rep; ++rx) {
              if (P.out_mode == MPN_OUT_F32_NHWC) {
                float* dst = (float*)P.y_hi + g.out_off + (long long)n * nstride +
                             ((long long)(oh * rep + ry) * OWr + (ow * rep + rx)) * P.out_cstride + P.out_coffset + cbase;
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < nc) dst[j] = v[j];
              } else {  // NCHW: for a fixed channel the 32 lanes are neighbouring pixels -> coalesced rows
                float* dst = (float*)P.y_hi + (long long)n * nstride + (long long)(P.out_coffset + cbase) * OHr * OWr +
                             (long long)(oh * rep + ry) * OWr + (ow * rep + rx);
#pragma unroll
                for (int j = 0; j < 16; ++j) if (j < nc) dst[(long long)j * OHr * OWr] = v[j];
              }
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(lead_tempty0 + 8u * acc);  // the leader's MMA issuer waits for both CTAs' epilogues
        else mbar_arrive(tempty_bar(acc));
      }
    }
    if (TMAEPI && store_pending && (warp & 3) == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }  // !WIDE
  }
  // =============================== teardown ===============================
  tc_fence_before();
  if (PAIR) cluster_sync_all();  // no CTA of the pair frees TMEM / exits while the other may still signal or read it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// 128 x 128 bf16 identity matrix in device memory (one per device, written once): the A operand of the residual MMA.
__device__ __nv_bfloat16 g_identity[BLOCK_M * BLOCK_M];

__global__ void init_identity_kernel() {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < BLOCK_M * BLOCK_M) g_identity[i] = __float2bfloat16_rn((i / BLOCK_M) == (i % BLOCK_M) ? 1.f : 0.f);
}

int identity_matrix(const void** ptr, cudaStream_t st) {
  static std::mutex mu;
  static bool ready[64] = {};
  static const void* addr[64] = {};
  int dev = 0;
  MPN_CUDA_OK(cudaGetDevice(&dev));
  MPN_CHECK_ARG(dev >= 0 && dev < 64, "conv(tcgen05): device index out of range");
  std::lock_guard<std::mutex> lock(mu);
  if (!addr[dev]) {
    void* a = nullptr;
    MPN_CUDA_OK(cudaGetSymbolAddress(&a, g_identity));
    addr[dev] = a;
  }
  if (!ready[dev]) {
    // stream-ordered before the convolution that needs it; inside a stream capture the launch becomes a graph node and
    // the matrix is not marked ready (an eager call later writes it for good)
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    MPN_CUDA_OK(cudaStreamIsCapturing(st, &cs));
    init_identity_kernel<<<BLOCK_M * BLOCK_M / 256, 256, 0, st>>>();
    MPN_LAUNCH_OK();
    if (cs == cudaStreamCaptureStatusNone) ready[dev] = true;
  }
  *ptr = addr[dev];
  return MPN_OK;
}

// ---------------------------------------------------------------------------------------------
// Pick the pixel box (TW x TH x TN <= 128) that wastes the fewest MMA rows.
// even: only boxes with even width and height (the 2x-upsample source of a tile is then one half-resolution box)
void choose_tile(int N, int OH, int OW, int* TW, int* TH, int* TN, bool even = false) {
  double best = -1.0;
  int bw = 1, bh = 1, bn = 1;
  for (int tw = 1; tw <= 128 && tw <= OW; ++tw) {
    for (int th = 1; th * tw <= 128 && th <= OH; ++th) {
      if (even && ((tw | th) & 1)) continue;
      int tn = 1;
      if (tw == OW && th == OH) {
        tn = 128 / (tw * th);
        if (tn > N) tn = N;
        if (tn < 1) tn = 1;
      }
      long long tiles = (long long)((OW + tw - 1) / tw) * ((OH + th - 1) / th) * ((N + tn - 1) / tn);
      double util = (double)N * OH * OW / (double)(tiles * 128);
      // prefer wide rows (longer contiguous pixel runs per TMA box row and coalesced NCHW stores)
      double score = util + 1e-4 * tw;
      if (score > best) { best = score; bw = tw; bh = th; bn = tn; }
    }
  }
  *TW = bw; *TH = bh; *TN = bn;
}

template <int BN, int MODE, int EPI, bool PAIR = false, int NG = 2>
int launch(const Maps& maps, const TcParams& P, cudaStream_t st, int sms) {
  constexpr int STAGE_BYTES = a_stage_bytes(MODE) + (PAIR ? BN / 2 : BN) * b_row_bytes(MODE);
  constexpr int EPI_BYTES = NG == 4 ? 4 * WIDE_BOX_BYTES : NUM_EPI_WARPS * EPI_STAGE_BYTES + (EPI == EPI_TMA_RES ? 4 * RES_STAGE_BYTES : 0);
  constexpr int MAXS = (SMEM_LIMIT - 1024 - BAR_BYTES - BIAS_BYTES - EPI_BYTES) / STAGE_BYTES;
  constexpr int STAGES = MAXS > 8 ? 8 : (MAXS < 2 ? 2 : MAXS);
  if constexpr (MAXS < 2) {   // MODE_F16F8B, 256-wide single-CTA tiles: the host plan never selects them (BN is capped at 128)
    mpn_set_error("conv(tcgen05): tile configuration does not fit shared memory (BN %d, mode %d)", BN, MODE);
    return MPN_ERR_UNSUPPORTED;
  } else {
  static_assert(8 * ((MODE == MODE_F16F8C ? 4 : 2) * STAGES + 9) <= BAR_BYTES, "barrier area too small");
  const int smem = STAGES * STAGE_BYTES + 1024 + BAR_BYTES + BIAS_BYTES + EPI_BYTES;
  static_assert(STAGES * STAGE_BYTES + 1024 + BAR_BYTES + BIAS_BYTES + EPI_BYTES <= SMEM_LIMIT, "shared memory budget");
  auto kern = conv_tc_kernel<BN, MODE, STAGES, EPI, PAIR, NG>;
  MPN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int grid = P.total_tiles < sms ? P.total_tiles : sms;
  if (PAIR) grid = 2 * (P.total_tiles < sms / 2 ? P.total_tiles : sms / 2);  // one 2-CTA cluster per pair tile, <= SMs/2 clusters
  static const int pdl = getenv("MPN_PDL") ? atoi(getenv("MPN_PDL")) : 0;  // opt-in: no step-time gain inside a CUDA graph (r01l)
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128 + 128 * NG);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  MPN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, maps, P));
  MPN_LAUNCH_OK();
  return MPN_OK;
  }
}

}  // namespace

// ds / ps: nseg descriptors that differ only in N, H, W, OH, OW and the x / y pointers (one per pyramid level); nseg == 1 is the
// ordinary launch.
static int tc_launch(const mpn_conv_desc* ds, const mpn_conv_ptrs* ps, int nseg, void* stream) {
  const mpn_conv_desc* d = ds;
  const mpn_conv_ptrs* p = ps;
  MPN_CHECK_ARG(nseg >= 1 && nseg <= MAX_SEG, "conv(tcgen05): 1..%d segments per launch", MAX_SEG);
  for (int sg = 1; sg < nseg; ++sg) {
    const mpn_conv_desc* e = ds + sg;
    MPN_CHECK_ARG(e->Cin == d->Cin && e->Cout == d->Cout && e->R == d->R && e->S == d->S && e->stride == d->stride && e->pad == d->pad &&
                      e->fmt == d->fmt && e->flags == d->flags && e->out_mode == d->out_mode && e->in_cstride == d->in_cstride &&
                      e->out_cstride == d->out_cstride && e->out_coffset == d->out_coffset && e->out_nstride == d->out_nstride &&
                      e->acc_scale == d->acc_scale,
                  "conv(tcgen05) multi-level launch: the levels must share filter, format, flags and output strides");
    MPN_CHECK_ARG(ps[sg].w_hi == p->w_hi && ps[sg].w_lo == p->w_lo && ps[sg].bias == p->bias && ps[sg].scale == p->scale,
                  "conv(tcgen05) multi-level launch: the levels must share the packed filter and bias");
  }
  if (nseg > 1) {
    for (int sg = 0; sg < nseg; ++sg) {
      const mpn_conv_desc* e = ds + sg;
      MPN_CHECK_ARG(e->stride == 1 && e->res_cstride == 0 && e->up_cstride == 0 && e->out_rep == 1 && !e->k_overlap && !e->in_wpitch &&
                        !e->in_hpitch && e->OH == e->H + 2 * e->pad - e->R + 1 && e->OW == e->W + 2 * e->pad - e->S + 1 && e->N > 0,
                    "conv(tcgen05) multi-level launch: stride-1 convolutions without shortcut / upsample / replication only");
      MPN_CHECK_ARG(e->out_mode == MPN_OUT_ACT || e->out_nstride > 0, "conv(tcgen05) multi-level launch: fp32 outputs need out_nstride");
    }
  }
  const bool split = d->fmt == MPN_FMT_BF16X2;
  const bool f8 = d->fmt == MPN_FMT_F16F8;
  const bool f8c = f8 && (d->flags & MPN_IN_DERIVE_H8);              // input without its h8 plane, derived in shared memory (MODE_F16F8C)
  const bool f8b = f8 && !f8c && (d->flags & MPN_IN_NO_H8);          // input without its h8 plane: fp16 weight-residual term (MODE_F16F8B)
  MPN_CHECK_ARG(!f8 || (p->x_lo && p->w_lo), "conv(tcgen05): F16F8 needs the byte planes of x and w");
  MPN_CHECK_ARG(f8 || !(d->flags & (MPN_IN_NO_H8 | MPN_EPI_NO_H8)), "conv(tcgen05): MPN_IN_NO_H8 / MPN_EPI_NO_H8 are MPN_FMT_F16F8 flags");
  MPN_CHECK_ARG(!f8 || (d->in_cstride % 16 == 0 && d->res_cstride % 16 == 0 && d->up_cstride % 16 == 0 &&
                        (d->out_mode != MPN_OUT_ACT || (d->out_cstride % 16 == 0 && d->out_coffset % 16 == 0))),
                "conv(tcgen05): F16F8 byte planes need channel strides / offsets that are multiples of 16");
  MPN_CHECK_ARG(d->stride == 1 || d->stride == 2, "conv(tcgen05): stride must be 1 or 2");
  MPN_CHECK_ARG(d->Cin % BLOCK_K == 0, "conv(tcgen05): Cin must be a multiple of 64 (got %d)", d->Cin);
  MPN_CHECK_ARG(d->in_cstride % 8 == 0, "conv(tcgen05): in_cstride must be a multiple of 8");
  MPN_CHECK_ARG(!split || (p->x_lo && p->w_lo), "conv(tcgen05): BF16X2 needs lo planes of x and w");
  MPN_CHECK_ARG(d->res_cstride % 8 == 0 && d->up_cstride % 8 == 0, "conv(tcgen05): residual/upsample channel strides must be multiples of 8");
  if (d->out_mode == MPN_OUT_ACT)
    MPN_CHECK_ARG(d->Cout % 32 == 0 && d->out_cstride % 8 == 0 && d->out_coffset % 8 == 0,
                  "conv(tcgen05): activation outputs need Cout %% 32 == 0 and 16-byte aligned channel offsets");
  if (d->res_cstride > 0 || d->up_cstride > 0) MPN_CHECK_ARG(d->Cout % 32 == 0, "conv(tcgen05): residual/upsample add needs Cout %% 32 == 0");
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    mpn_set_error("conv(tcgen05): cuTensorMapEncodeTiled entry point not available");
    return MPN_ERR_CUDA;
  }

  TcParams P;
  memset(&P, 0, sizeof(P));
  P.N = d->N; P.OH = d->OH; P.OW = d->OW; P.Cout = d->Cout;
  {
    static const int up_even = getenv("MPN_UP_TMA") ? atoi(getenv("MPN_UP_TMA")) : 1;
    const bool up2x = up_even && d->up_cstride > 0 && d->OH == 2 * d->up_h && d->OW == 2 * d->up_w && d->OH >= 2 && d->OW >= 2;
    choose_tile(d->N, d->OH, d->OW, &P.TW, &P.TH, &P.TN, up2x);
  }
  P.rows = P.TW * P.TH * P.TN;
  P.tiles_w = mpn_divup(d->OW, P.TW);
  P.tiles_h = mpn_divup(d->OH, P.TH);
  P.tiles_n = mpn_divup(d->N, P.TN);
  // per-segment geometry (segment 0 = the values above); unit0 is filled in once pair / single-CTA scheduling is known
  P.nseg = nseg;
  long long m_tiles_all = 0;
  for (int sg = 0; sg < nseg; ++sg) {
    const mpn_conv_desc* e = ds + sg;
    Seg& g = P.seg[sg];
    g.N = e->N; g.OH = e->OH; g.OW = e->OW;
    if (sg == 0) { g.TW = P.TW; g.TH = P.TH; g.TN = P.TN; }
    else choose_tile(e->N, e->OH, e->OW, &g.TW, &g.TH, &g.TN, false);
    g.rows = g.TW * g.TH * g.TN;
    g.tiles_w = mpn_divup(e->OW, g.TW);
    g.tiles_h = mpn_divup(e->OH, g.TH);
    g.tiles_n = mpn_divup(e->N, g.TN);
    g.m_tiles = g.tiles_w * g.tiles_h * g.tiles_n;
    g.out_off = (long long)(((const char*)ps[sg].y_hi - (const char*)p->y_hi) / 4);   // fp32 outputs: element offset of the level
    m_tiles_all += g.m_tiles;
  }
  // ---- tile plan.  A persistent launch runs `rounds` full rounds of G = min(tiles, SMs) tiles plus a partial round of
  // `rem` tiles that would leave most SMs idle; the plan cuts those rem tiles into s sub-tiles of BN/s columns (>= 32)
  // so the tail costs one narrow tile instead of one full tile.  Relative tile cost ~ (columns + 64): the fixed part is
  // the 128-row activation window every tile loads regardless of its width (fitted on r01c/r01d: 256-wide tiles take
  // 1.66x a 128-wide tile).  BN in {256, 128} (when Cout > 128) and s are chosen to minimise the makespan.  MPN_SPLIT_BN256 = 0/1 forces the width, MPN_TAIL_SPLIT = 0 disables the tail cut.
  int sms = 148;
  {
    int dev = 0;
    MPN_CUDA_OK(cudaGetDevice(&dev));
    MPN_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  static const int split256 = getenv("MPN_SPLIT_BN256") ? atoi(getenv("MPN_SPLIT_BN256")) : -1;
  static const int tail_on = getenv("MPN_TAIL_SPLIT") ? atoi(getenv("MPN_TAIL_SPLIT")) : 1;
  const long long m_tiles = m_tiles_all;
  const bool f32out = d->out_mode != MPN_OUT_ACT;
  // Activation outputs without upsample-add / replication go through the TMA-store epilogue (MPN_EPI_TMA=0 forces the LSU
  // one); there a split-format residual without an epilogue scale is added by the tensor core (MPN_RES_MMA=0: by the LSU).
  static const int epi_tma_on = getenv("MPN_EPI_TMA") ? atoi(getenv("MPN_EPI_TMA")) : 1;
  static const int res_mma_on = getenv("MPN_RES_MMA") ? atoi(getenv("MPN_RES_MMA")) : 1;
  // ... and an exact-2x nearest-upsample add (FPN laterals, fpn.py:84-95) rides on the same TMA machinery as the shortcut
  // (MPN_UP_TMA=0: the LSU epilogue): even tile sizes put every tile's source pixels in one half-resolution box
  static const int up_tma_on = getenv("MPN_UP_TMA") ? atoi(getenv("MPN_UP_TMA")) : 1;
  static const int pair_on_ = getenv("MPN_PAIR") ? atoi(getenv("MPN_PAIR")) : 1;
  static const int res_tma_on_ = getenv("MPN_RES_TMA") ? atoi(getenv("MPN_RES_TMA")) : 1;
  const bool up_tma = up_tma_on && epi_tma_on && pair_on_ && res_tma_on_ && !f32out && d->out_rep == 1 && d->up_cstride > 0 &&
                      d->res_cstride == 0 && d->OH == 2 * d->up_h && d->OW == 2 * d->up_w && P.TW % 2 == 0 && P.TH % 2 == 0 &&
                      d->Cout >= 128 && d->Cout % 64 == 0 && m_tiles >= 2 && sms >= 2;
  const bool epi_tma = epi_tma_on && !f32out && d->out_rep == 1 && (d->up_cstride == 0 || up_tma);
  const bool res_mma = res_mma_on && epi_tma && split && d->res_cstride > 0 && !p->scale && d->Cout % 64 == 0;
  // CTA-pair kernel (MPN_PAIR=0 disables): activation outputs of >= 128 channels, no residual (the tensor-core residual add
  // stays on the single-CTA kernel), enough M tiles to form pairs.  A pair tile = two consecutive M tiles x one Cout block.
  static const int pair_on = getenv("MPN_PAIR") ? atoi(getenv("MPN_PAIR")) : 1;
  const bool pair = pair_on && !f32out && !res_mma && d->Cout >= 128 && m_tiles >= 2 && sms >= 2;
  // ... and there the shortcut of a residual conv comes in by TMA (MPN_RES_TMA=0: per-thread global loads in the epilogue)
  static const int res_tma_on = getenv("MPN_RES_TMA") ? atoi(getenv("MPN_RES_TMA")) : 1;
  const bool res_tma = res_tma_on && pair && epi_tma && (d->res_cstride > 0 || up_tma) && d->Cout % 64 == 0;
  const int min_tail_bn = pair ? 64 : res_mma ? 64 : 32;
  int BN = d->Cout > 128 ? 256 : d->Cout > 64 ? 128 : d->Cout > 32 ? 64 : 32;
  if (f8b && !pair && BN > 128) BN = 128;   // 256 filter rows x 320 B per K block do not leave room for a 2-stage ring
  int tail_s = 1;
  long long sched_m = 0;                                           // schedulable M units (pair kernels: pairs of M tiles of a level)
  for (int sg = 0; sg < nseg; ++sg) {
    P.seg[sg].unit0 = (int)sched_m;
    sched_m += pair ? (P.seg[sg].m_tiles + 1) / 2 : P.seg[sg].m_tiles;
  }
  MPN_CHECK_ARG(nseg == 1 || f32out || (epi_tma && !res_tma && !res_mma), "conv(tcgen05) multi-level launch: TMA-store or fp32 epilogue only");
  const int sched_sms = pair ? sms / 2 : sms;                      // ... and the units that run at once
  {
    int cand[3] = {BN, 0, 0};
    if (BN == 256) {
      if (split256 == 0) cand[0] = 128;
      else if (split256 < 0) cand[1] = 128;
    }
    double best = 1e30;
    for (int ci = 0; ci < 3 && cand[ci]; ++ci) {
      const int bn = cand[ci];
      const long long T = sched_m * mpn_divup(d->Cout, bn);
      const long long G = T < sched_sms ? T : sched_sms;
      const long long rounds = T / G, rem = T % G;
      const double c_full = bn + 64.0;
      double cost = (double)(rounds + (rem > 0)) * c_full;
      int s_best = 1;
      if (tail_on && !f32out && rem > 0) {
        for (int sdiv = 2; bn / sdiv >= min_tail_bn && rem * sdiv <= G; sdiv *= 2) {
          const double c = (double)rounds * c_full + (bn / sdiv + 64.0);
          if (c < cost) { cost = c; s_best = sdiv; }
        }
      }
      if (cost < best) { best = cost; BN = bn; tail_s = s_best; }
    }
  }
  P.tiles_co = mpn_divup(d->Cout, BN);
  {
    const long long T = sched_m * P.tiles_co;
    MPN_CHECK_ARG(T < (1LL << 27), "conv(tcgen05): too many tiles");
    const long long G = T < sched_sms ? T : sched_sms;
    P.m_tiles = (int)m_tiles;
    const long long rem = tail_s > 1 ? T % G : 0;
    P.main_tiles = (int)(T - rem);
    P.tail_split = tail_s;
    P.tail_bn = BN / tail_s;
    P.total_tiles = (int)(T - rem + rem * tail_s);
  }
  P.res_mma = res_mma ? 1 : 0;
  P.up_tma = up_tma ? 1 : 0;
  // wide epilogue (MPN_EPI_WIDE=0 disables): F16F8 shortcut convolutions whose output is stored without the e5m2 copy plane; its
  // output and shortcut boxes are 64 channels wide (half the TMA rows of the 32-channel boxes of the other epilogues)
  static const int wide_on = getenv("MPN_EPI_WIDE") ? atoi(getenv("MPN_EPI_WIDE")) : 1;
  static const int wide_maxk = getenv("MPN_EPI_WIDE_MAXK") ? atoi(getenv("MPN_EPI_WIDE_MAXK")) : 256;
  // ... and whose reduction is short (r02i, 32-channel boxes: with eight K blocks per tile the MMA phase covers the epilogue and
  // the single shortcut buffer cost more than the extra warps won)
  const bool wide = wide_on && res_tma && !up_tma && f8 && !f8c && (d->flags & MPN_EPI_NO_H8) && !p->scale && nseg == 1 && P.tail_split <= 4 &&
                    d->Cin * d->R * d->S <= wide_maxk;
  P.R = d->R; P.S = d->S; P.stride = d->stride; P.pad = d->pad; P.Cin = d->Cin; P.kb_per_tap = d->Cin / BLOCK_K;
  P.scale = p->scale; P.bias = p->bias;
  P.res_hi = (const __nv_bfloat16*)p->res_hi; P.res_lo = (const __nv_bfloat16*)p->res_lo; P.res_cstride = d->res_cstride;
  P.up_hi = (const __nv_bfloat16*)p->up_hi; P.up_lo = (const __nv_bfloat16*)p->up_lo;
  P.up_h = d->up_h; P.up_w = d->up_w; P.up_cstride = d->up_cstride;
  P.flags = d->flags; P.out_mode = d->out_mode; P.out_cstride = d->out_cstride; P.out_coffset = d->out_coffset;
  P.out_rep = d->out_rep; P.out_nstride = d->out_nstride;
  P.y_hi = p->y_hi; P.y_lo = p->y_lo;
  P.acc_scale = d->acc_scale != 0.f ? d->acc_scale : 1.f;
#ifdef MPN_CONV_TRACE
  P.dbg_skip = getenv("MPN_DEBUG_SKIP") ? atoi(getenv("MPN_DEBUG_SKIP")) : 0;
#endif
  P.w_merged = (f8 && !f8b && (d->flags & MPN_W_MERGED)) ? 1 : 0;
  MPN_CHECK_ARG(!(f8b && (d->flags & MPN_W_MERGED)), "conv(tcgen05): MPN_W_MERGED filters cannot serve MPN_IN_NO_H8 (fp16 residual plane)");
  static const int epi_agent_on = getenv("MPN_EPI_AGENT") ? atoi(getenv("MPN_EPI_AGENT")) : 1;
  P.epi_agent = epi_agent_on;
  static const int res_pf_on = getenv("MPN_RES_PF") ? atoi(getenv("MPN_RES_PF")) : 0;
  P.res_pf = res_pf_on;
  P.gat_n = d->gat_n;
  MPN_CHECK_ARG(d->gat_n >= 0 && d->gat_n <= 2, "conv(tcgen05): gat_n must be 0, 1 or 2");
  if (d->gat_n > 0) {
    MPN_CHECK_ARG(epi_tma && !res_tma && nseg == 1 && d->Cout % 32 == 0, "conv(tcgen05): phase-class addends need the TMA-store epilogue without a shortcut");
    for (int gi = 0; gi < d->gat_n; ++gi) {
      const int sh = d->gat_shift[gi];
      MPN_CHECK_ARG(sh >= 1 && sh <= 4 && (d->gat_h[gi] << sh) == d->OH && (d->gat_w[gi] << sh) == d->OW &&
                        d->gat_cstride[gi] >= 9 * d->Cout && d->gat_cstride[gi] % 16 == 0 && p->gat_hi[gi] &&
                        (p->gat_lo[gi] || d->fmt == MPN_FMT_BF16),
                    "conv(tcgen05): phase-class addend %d: needs OH = gat_h << shift, OW = gat_w << shift, cstride >= 9*Cout", gi);
      P.gat_shift[gi] = sh; P.gat_h[gi] = d->gat_h[gi]; P.gat_w[gi] = d->gat_w[gi]; P.gat_cstride[gi] = d->gat_cstride[gi];
      P.gat_hi[gi] = p->gat_hi[gi]; P.gat_lo[gi] = p->gat_lo[gi];
    }
  }
  {
    const long long rep2 = (long long)d->out_rep * d->out_rep;
    const long long ns = d->out_nstride > 0 ? d->out_nstride : (long long)d->OH * d->OW * rep2 * d->out_cstride;
    P.y_plane = (long long)d->N * ns;
  }

  alignas(64) Maps maps;
  memset(&maps, 0, sizeof(maps));
  const int planes = f8 ? 3 : split ? 2 : 1;
  const int st = d->stride;
  const int nphase = st == 1 ? 1 : 4;
  const long long wpitch = d->in_wpitch > 0 ? d->in_wpitch : d->W;   // pixels per row in memory
  const long long hpitch = d->in_hpitch > 0 ? d->in_hpitch : d->H;   // rows per image in memory
  MPN_CHECK_ARG(wpitch >= d->W && hpitch >= d->H, "conv(tcgen05): pitches smaller than the logical size");
  MPN_CHECK_ARG(!d->k_overlap || (st == 1 && d->Cin == BLOCK_K && d->in_cstride % 8 == 0 &&
                                  wpitch >= d->W + d->Cin / d->in_cstride - 1),
                "conv(tcgen05): k_overlap needs stride 1, Cin == 64 and a row pitch covering the window");
  const long long x_plane = (long long)d->N * hpitch * wpitch * d->in_cstride;   // F16F8: h8 plane = lo8 plane + x_plane bytes
  const long long w_plane = (long long)d->Cout * d->R * d->S * d->Cin;
  // ---- activation planes: hi (2-byte elements, 128B swizzle); BF16X2: lo; F16F8: lo8 and (unless MPN_IN_NO_H8) h8 byte planes
  const int a_planes = (f8b || f8c) ? 2 : planes;
  for (int pl = 0; pl < a_planes; ++pl) {
    const bool bytes = f8 && pl > 0;                 // byte planes: 1-byte elements, 64-byte swizzle
    const unsigned long long es = bytes ? 1ULL : 2ULL;
    const CUtensorMapSwizzle swz = bytes ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapDataType dt = bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const char* xb = (const char*)(pl == 0 ? p->x_hi : p->x_lo) + (f8 && pl == 2 ? x_plane : 0);
    for (int ph = 0; ph < nphase; ++ph) {
      const int hp = ph >> 1, wp = ph & 1;
      const int Hp = (d->H - hp + st - 1) / st, Wp = (d->W - wp + st - 1) / st;
      if (Hp <= 0 || Wp <= 0) {
        P.phase_empty |= 1 << ph;
        continue;
      }
      cuuint64_t dims[4] = {(cuuint64_t)d->Cin, (cuuint64_t)Wp, (cuuint64_t)Hp, (cuuint64_t)d->N};
      cuuint64_t strides[3] = {(cuuint64_t)d->in_cstride * st * es, (cuuint64_t)wpitch * d->in_cstride * st * es,
                               (cuuint64_t)hpitch * wpitch * d->in_cstride * es};
      cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
      const char* base = xb + ((long long)hp * wpitch + wp) * d->in_cstride * (long long)es;
      int rc = encode(fn, &maps.a[pl][ph], base, 4, dims, strides, box, swz, dt);
      if (rc) return rc;
    }
    MPN_CHECK_ARG(!(P.phase_empty & 1), "conv(tcgen05): empty input");
    for (int sg = 1; sg < nseg; ++sg) {   // multi-level launch (stride 1, dense tensors): map slot = segment
      const mpn_conv_desc* e = ds + sg;
      const Seg& g = P.seg[sg];
      const long long xpl = (long long)e->N * e->H * e->W * e->in_cstride;
      const char* xs = (const char*)(pl == 0 ? ps[sg].x_hi : ps[sg].x_lo) + (f8 && pl == 2 ? xpl : 0);
      MPN_CHECK_ARG(ps[sg].x_hi && (pl == 0 || ps[sg].x_lo), "conv(tcgen05) multi-level launch: missing input plane");
      cuuint64_t dims[4] = {(cuuint64_t)e->Cin, (cuuint64_t)e->W, (cuuint64_t)e->H, (cuuint64_t)e->N};
      cuuint64_t strides[3] = {(cuuint64_t)e->in_cstride * es, (cuuint64_t)e->W * e->in_cstride * es,
                               (cuuint64_t)e->H * e->W * e->in_cstride * es};
      cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TN};
      int rc = encode(fn, &maps.a[pl][sg], xs, 4, dims, strides, box, swz, dt);
      if (rc) return rc;
    }
  }
  // ---- filter planes: [Cout, R*S*Cin] K-major.  F16F8: hi fp16 | lo8 bytes | h8 bytes; MPN_IN_NO_H8: hi fp16 | lo16 fp16 | h8 bytes
  const bool wm = P.w_merged != 0;
  for (int pl = 0; pl < planes; ++pl) {
    const bool bytes = f8 && (f8b ? pl == 2 : pl > 0);
    const bool mrg = wm && pl > 0;                   // merged byte planes: [Cout][2K] bytes, boxes of 128-byte rows (slot 2 repeats slot 1)
    const unsigned long long es = bytes ? 1ULL : 2ULL;
    const CUtensorMapSwizzle swz = (bytes && !mrg) ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
    const CUtensorMapDataType dt = bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const long long off = !f8 || pl < 2 || mrg ? 0 : (f8b ? 2 * w_plane : w_plane);   // byte offset of the third plane inside w_lo
    const char* wb = (const char*)(pl == 0 ? p->w_hi : p->w_lo) + off;
    const cuuint64_t K = (cuuint64_t)d->R * d->S * d->Cin * (mrg ? 2 : 1);
    cuuint64_t wdims[2] = {K, (cuuint64_t)d->Cout};
    cuuint64_t wstrides[1] = {K * es};
    cuuint32_t wbox[2] = {(cuuint32_t)(mrg ? 2 * BLOCK_K : BLOCK_K), (cuuint32_t)(pair ? BN / 2 : BN)};   // pair: each CTA loads half of the filter rows
    int rc = encode(fn, &maps.b[pl], wb, 2, wdims, wstrides, wbox, swz, dt);
    if (rc) return rc;
    if (P.tail_split > 1) {
      cuuint32_t sbox[2] = {(cuuint32_t)(mrg ? 2 * BLOCK_K : BLOCK_K), (cuuint32_t)(pair ? P.tail_bn / 2 : P.tail_bn)};
      rc = encode(fn, &maps.bs[pl], wb, 2, wdims, wstrides, sbox, swz, dt);
      if (rc) return rc;
    }
  }
  if (epi_tma) {
    for (int sg = 0; sg < nseg; ++sg) {
      const mpn_conv_desc* e = ds + sg;
      const Seg& g = P.seg[sg];
      const long long nstride = e->out_nstride > 0 ? e->out_nstride : (long long)e->OH * e->OW * e->out_cstride;
      const long long ypl = (long long)e->N * nstride;   // elements of one output plane of this level (F16F8: h8 follows lo8)
      for (int pl = 0; pl < planes; ++pl) {
        if (f8 && pl == 2 && (d->flags & MPN_EPI_NO_H8)) continue;   // output stored without its h8 plane
        const bool bytes = f8 && pl > 0;
        const unsigned long long es = bytes ? 1ULL : 2ULL;
        cuuint64_t ydims[4] = {(cuuint64_t)e->Cout, (cuuint64_t)e->OW, (cuuint64_t)e->OH, (cuuint64_t)e->N};
        cuuint64_t ystr[3] = {(cuuint64_t)e->out_cstride * es, (cuuint64_t)e->OW * e->out_cstride * es, (cuuint64_t)nstride * es};
        cuuint32_t ybox[4] = {wide ? 64u : 32u, (cuuint32_t)g.TW, (cuuint32_t)g.TH, (cuuint32_t)g.TN};
        MPN_CHECK_ARG(ps[sg].y_hi && (pl == 0 || ps[sg].y_lo), "conv(tcgen05): missing output plane");
        const char* yb = (const char*)(pl == 0 ? ps[sg].y_hi : ps[sg].y_lo) + (f8 && pl == 2 ? ypl : 0) + (long long)e->out_coffset * (long long)es;
        int rc = encode(fn, &maps.y[pl][sg], yb, 4, ydims, ystr, ybox,
                        wide ? (bytes ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)
                             : (bytes ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B),
                        bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
        if (rc) return rc;
      }
    }
  }
  if (res_mma) {
    const void* ident = nullptr;
    int rc = identity_matrix(&ident, (cudaStream_t)stream);
    if (rc) return rc;
    cuuint64_t idims[2] = {128, 128};
    cuuint64_t istr[1] = {256};
    cuuint32_t ibox[2] = {64, 128};
    rc = encode(fn, &maps.ident, ident, 2, idims, istr, ibox);
    if (rc) return rc;
    for (int pl = 0; pl < 2; ++pl) {
      cuuint64_t rdims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)d->OW, (cuuint64_t)d->OH, (cuuint64_t)d->N};
      cuuint64_t rstr[3] = {(cuuint64_t)d->res_cstride * 2ULL, (cuuint64_t)d->OW * d->res_cstride * 2ULL,
                            (cuuint64_t)d->OH * d->OW * d->res_cstride * 2ULL};
      cuuint32_t rbox[4] = {64u, (cuuint32_t)P.TW, (cuuint32_t)P.TH, (cuuint32_t)P.TN};
      rc = encode(fn, &maps.r[pl], pl == 0 ? p->res_hi : p->res_lo, 4, rdims, rstr, rbox);
      if (rc) return rc;
    }
  }
  if (res_tma) {  // shortcut boxes in the geometry of the output boxes: 32 channels x the tile's pixels (up_tma: x its source pixels)
    const int rw = up_tma ? d->up_w : d->OW, rh_ = up_tma ? d->up_h : d->OH, rcs = up_tma ? d->up_cstride : d->res_cstride;
    for (int pl = 0; pl < (f8 || split ? 2 : 1); ++pl) {
      const bool bytes = f8 && pl > 0;
      const unsigned long long es = bytes ? 1ULL : 2ULL;
      cuuint64_t rdims[4] = {(cuuint64_t)d->Cout, (cuuint64_t)rw, (cuuint64_t)rh_, (cuuint64_t)d->N};
      cuuint64_t rstr[3] = {(cuuint64_t)rcs * es, (cuuint64_t)rw * rcs * es, (cuuint64_t)rh_ * rw * rcs * es};
      cuuint32_t rbox[4] = {wide ? 64u : 32u, (cuuint32_t)(up_tma ? P.TW / 2 : P.TW), (cuuint32_t)(up_tma ? P.TH / 2 : P.TH), (cuuint32_t)P.TN};
      const void* rbase = up_tma ? (pl == 0 ? p->up_hi : p->up_lo) : (pl == 0 ? p->res_hi : p->res_lo);
      int rc = encode(fn, &maps.r[pl], rbase, 4, rdims, rstr, rbox,
                      wide ? (bytes ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B)
                           : (bytes ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B),
                      bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
      if (rc) return rc;
    }
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (f32out) {  // head outputs (18/19/36/9 channels): the thread-per-row fp32 store path, small-N tiles only
    MPN_CHECK_ARG(BN <= 64, "conv(tcgen05): fp32 outputs are built for Cout <= 64 (got %d)", d->Cout);
    if (f8b) return BN == 64 ? launch<64, MODE_F16F8B, EPI_F32>(maps, P, s, sms) : launch<32, MODE_F16F8B, EPI_F32>(maps, P, s, sms);
    if (f8c) return BN == 64 ? launch<64, MODE_F16F8C, EPI_F32>(maps, P, s, sms) : launch<32, MODE_F16F8C, EPI_F32>(maps, P, s, sms);
    if (f8) return BN == 64 ? launch<64, MODE_F16F8, EPI_F32>(maps, P, s, sms) : launch<32, MODE_F16F8, EPI_F32>(maps, P, s, sms);
    if (split) return BN == 64 ? launch<64, MODE_BF16X2, EPI_F32>(maps, P, s, sms) : launch<32, MODE_BF16X2, EPI_F32>(maps, P, s, sms);
    return BN == 64 ? launch<64, MODE_BF16, EPI_F32>(maps, P, s, sms) : launch<32, MODE_BF16, EPI_F32>(maps, P, s, sms);
  }
#define MPN_TC_DISPATCH(SPLIT_, EPI_)                                     \
  if (pair) {                                                              \
    if (BN == 256) return launch<256, SPLIT_, EPI_, true>(maps, P, s, sms); \
    return launch<128, SPLIT_, EPI_, true>(maps, P, s, sms);               \
  }                                                                        \
  switch (BN) {                                                            \
    case 256: return launch<256, SPLIT_, EPI_>(maps, P, s, sms);           \
    case 128: return launch<128, SPLIT_, EPI_>(maps, P, s, sms);           \
    case 64: return launch<64, SPLIT_, EPI_>(maps, P, s, sms);             \
    default: return launch<32, SPLIT_, EPI_>(maps, P, s, sms);             \
  }
  if (wide) {
    if (f8b) return BN == 256 ? launch<256, MODE_F16F8B, EPI_TMA_RES, true, 4>(maps, P, s, sms) : launch<128, MODE_F16F8B, EPI_TMA_RES, true, 4>(maps, P, s, sms);
    return BN == 256 ? launch<256, MODE_F16F8, EPI_TMA_RES, true, 4>(maps, P, s, sms) : launch<128, MODE_F16F8, EPI_TMA_RES, true, 4>(maps, P, s, sms);
  }
#define MPN_TC_DISPATCH_RES(SPLIT_)                                                  \
  if (res_tma) {                                                                      \
    if (BN == 256) return launch<256, SPLIT_, EPI_TMA_RES, true>(maps, P, s, sms);    \
    return launch<128, SPLIT_, EPI_TMA_RES, true>(maps, P, s, sms);                   \
  }
  if (f8b) {
    MPN_TC_DISPATCH_RES(MODE_F16F8B)
    if (epi_tma) { MPN_TC_DISPATCH(MODE_F16F8B, EPI_TMA) }
    MPN_TC_DISPATCH(MODE_F16F8B, EPI_LSU)
  }
  if (f8c) {
    MPN_TC_DISPATCH_RES(MODE_F16F8C)
    if (epi_tma) { MPN_TC_DISPATCH(MODE_F16F8C, EPI_TMA) }
    MPN_TC_DISPATCH(MODE_F16F8C, EPI_LSU)
  }
  if (f8) {
    MPN_TC_DISPATCH_RES(MODE_F16F8)
    if (epi_tma) { MPN_TC_DISPATCH(MODE_F16F8, EPI_TMA) }
    MPN_TC_DISPATCH(MODE_F16F8, EPI_LSU)
  }
  if (split) {
    MPN_TC_DISPATCH_RES(MODE_BF16X2)
    if (epi_tma) { MPN_TC_DISPATCH(MODE_BF16X2, EPI_TMA) }
    MPN_TC_DISPATCH(MODE_BF16X2, EPI_LSU)
  }
  MPN_TC_DISPATCH_RES(MODE_BF16)
  if (epi_tma) { MPN_TC_DISPATCH(MODE_BF16, EPI_TMA) }
  MPN_TC_DISPATCH(MODE_BF16, EPI_LSU)
#undef MPN_TC_DISPATCH
#undef MPN_TC_DISPATCH_RES
}

int mpn_conv_tc_launch(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream) { return tc_launch(d, p, 1, stream); }

int mpn_conv_tc_launch_multi(const mpn_conv_desc* ds, const mpn_conv_ptrs* ps, int nseg, void* stream) {
  return tc_launch(ds, ps, nseg, stream);
}

#ifdef MPN_CONV_TRACE
extern "C" int mpn_debug_conv_trace(void* buf, int cap) {
  unsigned long long* b = (unsigned long long*)buf;
  MPN_CUDA_OK(cudaMemcpyToSymbol(g_trace_buf, &b, sizeof(b)));
  MPN_CUDA_OK(cudaMemcpyToSymbol(g_trace_cap, &cap, sizeof(cap)));
  return MPN_OK;
}
#endif
