/* mpn_b200.h -- C ABI of libmpn_b200.so, the B200 (sm_100a) hot path of MultiPoseNet.
 *
 * The reference's only native boundary is lib/nms (cffi):
 *     int cpu_nms(THLongTensor* keep_out, THLongTensor* num_out, THFloatTensor* boxes,
 *                 THLongTensor* order, THFloatTensor* areas, float thresh);        lib/nms/src/nms.h:1
 *     int gpu_nms(THLongTensor* keep_out, THLongTensor* num_out, THCudaTensor* boxes,
 *                 float thresh);                                                   lib/nms/src/nms_cuda.h:1
 * everything else on the path is a torch.nn call from network/posenet.py / network/fpn.py.  This
 * header therefore (1) mirrors gpu_nms/cpu_nms on plain pointers (mpn_nms*), and (2) exposes one entry
 * point per torch.nn call-site family of the path (conv2d+epilogue, max-pool, layout, decode/filter)
 * so the Python host mirror of network/posenet.py can drive them with raw device pointers.
 *
 * Conventions: plain `extern "C"`, raw DEVICE pointers unless a parameter says host, sizes as ints,
 * an explicit stream (void* = cudaStream_t; NULL = legacy default stream), caller-owned outputs and
 * workspaces, no global mutable state (error text is thread-local).  Return 0 = success, <0 = error
 * (text via mpn_last_error()).  The reference's "1 = success" convention (nms.c:68) is restored by the
 * Python shim lib/nms/pth_nms.py.
 */
#ifndef MPN_B200_H_
#define MPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPN_OK 0
#define MPN_ERR_ARG (-1)
#define MPN_ERR_CUDA (-2)
#define MPN_ERR_UNSUPPORTED (-3)

/* Activation storage formats (NHWC, channel stride given separately). */
#define MPN_FMT_F32 0    /* one fp32 plane                    -> CUDA-core fp32 path           */
#define MPN_FMT_BF16 1   /* one bf16 plane                    -> tcgen05 single-pass bf16      */
#define MPN_FMT_BF16X2 2 /* bf16 hi plane + bf16 lo plane     -> tcgen05 3-pass split (~fp32)  */
#define MPN_FMT_F16F8 3  /* fp16 hi plane + two fp8 byte planes -> tcgen05 1 f16 + 2 half-cost fp8 passes (~fp32, |x| < 65504):
                            x = hi + 2^-12 * lo8, lo8 = e5m2((x - hi) * 2^12), h8 = e5m2(x).  Every `*_lo` pointer of this
                            format addresses the two byte planes back to back: [lo8 : E elements][h8 : E elements], E = the
                            element count of the hi plane.  Filters: hi = fp16(w'), lo8 = e4m3(w' - hi), h8 = e4m3(w' * 2^-12)
                            with w' = w * 2^k (one power of two per filter, mpn_conv_desc.acc_scale = 2^-k), so that
                            x*w*2^k = xhi*whi (kind::f16) + xlo8*wh8 + xh8*wlo8 (kind::f8f6f4) lands in ONE accumulator.   */

/* Output modes of a conv epilogue. */
#define MPN_OUT_ACT 0      /* activation format `fmt`, NHWC                                   */
#define MPN_OUT_F32_NHWC 1 /* fp32 [N, OH*rep, OW*rep, Cout] (== permute(0,2,3,1), posenet.py:67,111) */
#define MPN_OUT_F32_NCHW 2 /* fp32 [N, Cout, OH*rep, OW*rep] (the boundary layout of the heat maps)   */

/* Epilogue flags. */
#define MPN_EPI_RELU 1
#define MPN_EPI_SIGMOID 2
/* MPN_FMT_F16F8 only.  MPN_EPI_NO_H8: the output tensor is stored without its e5m2 copy plane (3 bytes per element); the `lo`
 * buffer then holds the lo8 plane alone.  MPN_IN_NO_H8: the input tensor has no e5m2 copy plane; the convolution then evaluates
 * the weight-residual term as x_hi (fp16) * w_lo16 (fp16) and expects the filter packed by mpn_pack_filter_f16f8 with layout 1:
 * w_hi fp16 [Cout,R,S,Cin]; w_lo = [lo16 fp16 plane][h8 e4m3 plane]. */
#define MPN_EPI_NO_H8 4
#define MPN_IN_NO_H8 8
/* MPN_IN_DERIVE_H8: the input tensor has no e5m2 copy plane and the kernel derives the copy in shared memory from the fp16 tile
 * TMA delivered (two converter warps per CTA); the filter is packed in the ordinary layout 0 (hi | lo8 | h8) and the MMA schedule
 * is the 8-slot one.  With it no activation needs its copy plane in HBM at all: the plane exists because TMA cannot gather every
 * second byte, and its 64-byte box rows cost the TMA unit as much as 128-byte ones (~2.25 cycles per row). */
#define MPN_IN_DERIVE_H8 16
/* MPN_W_MERGED: the filter's two byte planes were packed interleaved (mpn_pack_filter_f16f8 layout bit 2: per filter row and 64 K
 * elements one 128-byte group [64 lo8 | 64 h8]), so a K block's filter tile arrives as two TMA boxes of 128-byte rows instead of
 * three.  Not combinable with MPN_IN_NO_H8 (whose residual plane is fp16). */
#define MPN_W_MERGED 32

typedef struct mpn_conv_desc {
  /* problem: y = epilogue(conv2d(x, w)); torch.nn.Conv2d semantics (cross-correlation, zero pad) */
  int N, H, W, Cin;          /* input  [N,H,W,Cin]  (NHWC)                                    */
  int Cout, R, S;            /* filter [Cout][R][S][Cin] for tcgen05, [R][S][Cin][CoutPad] for fp32 */
  int stride, pad;
  int OH, OW;                /* (H + 2*pad - R)/stride + 1                                    */
  int fmt;                   /* MPN_FMT_* of x, residual, upsample source and MPN_OUT_ACT output */
  int in_cstride;            /* channel stride (elements per pixel) of x                      */
  int flags;                 /* MPN_EPI_*                                                     */
  /* epilogue, applied in this order: v = acc*scale[c] + bias[c]; v += residual; v += up; relu/sigmoid */
  int res_cstride;           /* residual [N,OH,OW,res_cstride] (0 = none)                     */
  int up_h, up_w, up_cstride;/* nearest-upsampled addend [N,up_h,up_w,up_cstride] (0 = none), fpn.py:84-95 */
  int out_mode;              /* MPN_OUT_*                                                     */
  int out_cstride;           /* channels per pixel of the destination (>= out_coffset+Cout)   */
  int out_coffset;           /* first destination channel (concat writes, posenet.py:254)     */
  int out_rep;               /* nearest-upsample replication factor on store: 1,2,4,8 (posenet.py:180-182) */
  long long out_nstride;     /* elements between images in the destination (0 = dense)        */
  int w_cout_pad;            /* fp32 path: padded Cout of the [R][S][Cin][CoutPad] filter     */
  int in_wpitch;             /* pixels per input row in memory (0 = W); tcgen05 path only     */
  int in_hpitch;             /* rows per input image in memory (0 = H); tcgen05 path only     */
  int k_overlap;             /* 1: in_cstride < Cin is intended -- each "pixel" of Cin channels is a window of
                                Cin/in_cstride neighbouring pixels (stem as a space-to-depth conv) */
  float acc_scale;           /* MPN_FMT_F16F8: the accumulator is multiplied by this (2^-k of the filter packing) before
                                scale/bias; 0 is read as 1 */
  /* Phase-class addends (tcgen05 path, MPN_OUT_ACT through the TMA-store epilogue; posenet.py:243-257 without the concat): a
   * 3x3 convolution over a x2^shift nearest-upsampled map equals, per output pixel, ONE of nine 3x3 convolutions of the
   * low-resolution map -- which one depends only on (oh mod 2^shift, ow mod 2^shift): class = rc*3 + cc with rc / cc = 0 on the
   * first row / column of a block, 2 on the last, 1 inside.  The caller evaluates the nine class convolutions at low resolution
   * (one convolution with 9*Cout output channels, class-major) and this epilogue adds, for g < gat_n,
   *   v[c] += G_g[n, oh >> gat_shift[g], ow >> gat_shift[g], class * Cout + c]   before ReLU / sigmoid. */
  int gat_n;                 /* 0, 1 or 2 gathered addends                                                     */
  int gat_shift[2];          /* log2 of the upsampling factor (2 -> x4, 3 -> x8)                             */
  int gat_h[2], gat_w[2];    /* low-resolution map size                                                        */
  int gat_cstride[2];        /* channel stride of G_g (>= 9 * Cout)                                            */
} mpn_conv_desc;

typedef struct mpn_conv_ptrs {
  const void* x_hi;   const void* x_lo;     /* lo planes only for MPN_FMT_BF16X2 */
  const void* w_hi;   const void* w_lo;
  const float* scale; const float* bias;    /* per-Cout fp32, either may be NULL */
  const void* res_hi; const void* res_lo;
  const void* up_hi;  const void* up_lo;
  void* y_hi;         void* y_lo;
  const void* gat_hi[2]; const void* gat_lo[2];   /* phase-class addends (fmt planes: hi, lo / lo8) */
} mpn_conv_ptrs;

const char* mpn_last_error(void);
int mpn_version(void);
/* sizeof(mpn_conv_desc) / sizeof(mpn_conv_ptrs) as compiled into the library: a foreign-language binding (ctypes, cgo, JNI)
 * compares them with its own struct layout when it loads the library. */
int mpn_sizeof_conv_desc(void);
int mpn_sizeof_conv_ptrs(void);
/* 1 if the current device can run the tcgen05 kernels (compute capability 10.x). */
int mpn_device_supports_tcgen05(void);

/* ---- conv2d (+BN-fold/bias/residual/upsample-add/ReLU/sigmoid): fpn.py:14-25,42-74,99-124,
 *      posenet.py:37-49,79-92,165-187.  fmt F32 -> CUDA-core kernel, BF16/BF16X2 -> tcgen05+TMA. */
int mpn_conv2d_fwd(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream);
/* One layer applied to several tensors with SHARED weights in ONE persistent launch -- a RetinaNet tower layer over the five
 * pyramid levels (posenet.py:262-263: `torch.cat([self.regressionModel(feature) for feature in features])`).  descs / ptrs are
 * arrays of nseg (<= 5) entries that differ only in N, H, W, OH, OW and the x / y pointers; stride 1, no shortcut / upsample /
 * replication; tensor-core formats only.  The levels' tiles form one tile list, so the small levels fill the SMs the large
 * ones leave idle instead of running as latency-bound launches of their own. */
int mpn_conv2d_fwd_multi(const mpn_conv_desc* descs, const mpn_conv_ptrs* ptrs, int nseg, void* stream);
/* fp32-input variant used for the stem: x is fp32 NHWC (Cin may be 3), output in d->fmt. */
int mpn_conv2d_fwd_f32in(const mpn_conv_desc* d, const mpn_conv_ptrs* p, void* stream);

/* ---- filter packing (host-side weights are torch OIHW fp32 on the device) */
/* OIHW fp32 -> [R][S][Cin][CoutPad] fp32 (zero padded) */
int mpn_pack_filter_f32(const float* w_oihw, float* dst, int Cout, int Cin, int R, int S, int CoutPad, void* stream);
/* OIHW fp32 -> [Cout][R][S][Cin] bf16 hi (+ lo = bf16(w - hi) if dst_lo != NULL) */
int mpn_pack_filter_bf16(const float* w_oihw, void* dst_hi, void* dst_lo, int Cout, int Cin, int R, int S, void* stream);
/* same with the eval-mode BatchNorm scale folded into the filter (w * scale[cout], rounded once in fp32, before the
 * hi/lo split): the epilogue then only adds the folded bias, and a residual (fpn.py:45-46: out += shortcut(x)) can be
 * accumulated by the tensor core ahead of it.  scale == NULL: plain packing. */
int mpn_pack_filter_bf16_scaled(const float* w_oihw, const float* scale, void* dst_hi, void* dst_lo, int Cout, int Cin, int R, int S,
                                void* stream);
/* MPN_FMT_F16F8 filters: max|w * scale| (device scalar, *amax must be zeroed by the caller) -> the host picks k with
 * max|w * scale| * 2^k in [2^14, 2^15) -> [Cout][R][S][Cin] fp16 hi, e4m3 lo8 / h8 byte planes (dst_lo8h8 = the two planes back
 * to back).  layout bit 0: the [64][4][64] space-to-depth layout of mpn_stem_pack_filter (Cin = 3, R = S = 7); bit 1: the planes
 * of a convolution whose input has no h8 plane (MPN_IN_NO_H8): dst_lo8h8 = [lo16 = fp16(w' - hi) : 2 bytes per element][h8]. */
int mpn_filter_absmax(const float* w_oihw, const float* scale, int Cout, int per_cout, float* amax, void* stream);
int mpn_pack_filter_f16f8(const float* w_oihw, const float* scale, float wscale, void* dst_hi, void* dst_lo8h8, int Cout, int Cin,
                          int R, int S, int layout, void* stream);
/* BatchNorm (eval) fold: scale = gamma/sqrt(var+eps), bias = beta - mean*scale   (fpn.py:15-19,25,43) */
int mpn_fold_bn(const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                float* scale, float* bias, int C, void* stream);

/* ---- tensor-core stem (fpn.py:99: 7x7/2 conv, 3 -> 64).  The image is re-laid as a zero-padded
 * space-to-depth tensor X2[n, H/2+3, W/2+3, 16] (channel = (row parity, col parity, rgb), 12 used) in
 * which the 7x7/2 conv is a 4x4/1 conv; one filter row (4 pixels x 16 ch = 64 contiguous elements) is one
 * K block, so mpn_conv2d_fwd runs it with R=4, S=1, Cin=64, in_cstride=16, k_overlap=1. */
int mpn_stem_pack_input(const float* img_nchw, void* dst_hi, void* dst_lo, int N, int H, int W, int fmt, int flags /* MPN_EPI_NO_H8 */,
                        void* stream);
/* OIHW [64,3,7,7] fp32 -> [64][4][64] bf16 hi (+lo) matching that layout */
int mpn_stem_pack_filter(const float* w_oihw, void* dst_hi, void* dst_lo, int Cout, void* stream);

/* ---- input side: resnet_preprocess of datasets/coco_data/preprocessing.py:15-26 on the device.
 * img: uint8 [N,H,W,3] BGR (cv2.imread layout).  Bit-identical to the numpy code (fp32 divide / subtract / divide). */
int mpn_preprocess_u8_nchw(const unsigned char* img_nhwc_bgr, float* out_nchw, int N, int H, int W, void* stream);
/* ... fused into mpn_stem_pack_input: uint8 image -> normalised, zero-padded space-to-depth stem operand */
int mpn_stem_pack_input_u8(const unsigned char* img_nhwc_bgr, void* dst_hi, void* dst_lo, int N, int H, int W, int fmt, int flags,
                           void* stream);

/* ---- layout / elementwise */
/* fp32 NCHW -> NHWC in `fmt` (dst_lo for BF16X2); cstride >= C, padding channels are zeroed */
int mpn_nchw_to_nhwc(const float* src, void* dst_hi, void* dst_lo, int N, int C, int H, int W, int cstride, int fmt, void* stream);
/* NHWC `fmt` -> fp32 NCHW (first C channels) */
int mpn_nhwc_to_nchw(const void* src_hi, const void* src_lo, float* dst, int N, int C, int H, int W, int cstride, int fmt, void* stream);
/* F.max_pool2d(kernel 3, stride 2, pad 1) on NHWC (fpn.py:100).  flags: MPN_EPI_NO_H8 = the MPN_FMT_F16F8 output is stored
 * without its h8 plane (its consumers are 1x1 convolutions running with MPN_IN_NO_H8) */
int mpn_maxpool3x3s2(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, int N, int H, int W, int C, int fmt, int flags,
                     void* stream);
/* y = relu(x) on n elements (fpn.py:108: conv7(F.relu(p6))) */
int mpn_relu(const void* x_hi, const void* x_lo, void* y_hi, void* y_lo, long long n, int fmt, void* stream);

/* out[r,:] = softmax(a[r,:] + res[r,:]) over D columns (posenet.py:147-148 / 345-346: Add then Softmax(dim=1)).
 * a: activation planes `fmt` with row stride a_stride elements; res, out: fp32 [P, D] dense. */
int mpn_add_softmax_rows(const void* a_hi, const void* a_lo, const float* res, float* out, int P, int D, int a_stride, int fmt,
                         void* stream);

/* ---- training step of the keypoint subnet (posenet.py:367-403 loss, BatchNorm2d in train mode,
 *      the autograd backward of fpn.py / posenet.py).  Activations are dense NHWC (cstride == C) unless noted. */
/* dW[Cout][R][S][Cin] fp32 = sum_pixels dY (x) X  -- tcgen05, MN-major operands, split-K over pixels with fp32 atomics.
 * d = the FORWARD conv descriptor; p->x_* = forward input, p->res_* (+ d->res_cstride) = dY. */
int mpn_conv2d_wgrad(const mpn_conv_desc* d, const mpn_conv_ptrs* p, float* dw, void* stream);
/* data-gradient filter of a conv: OIHW fp32 -> [Cin][R][S][CoutPad] bf16 hi(+lo), taps flipped; dX = conv(dY, W') with
 * pad' = R-1-pad runs on mpn_conv2d_fwd (stride-2 convs: dY zero-inserted first, mpn_zero_insert2). */
int mpn_pack_filter_dgrad_bf16(const float* w_oihw, void* dst_hi, void* dst_lo, int Cout, int Cin, int R, int S, int CoutPad, void* stream);
int mpn_unpack_filter_grad(const float* g_corsci, float* out_oihw, int Cout, int Cin, int R, int S, void* stream);
int mpn_stem_unpack_filter_grad(const float* g_stem, float* out_oihw, int Cout, void* stream);
int mpn_zero_insert2(const void* dy_hi, const void* dy_lo, void* out_hi, void* out_lo, int N, int OH, int OW, int C, int H, int W, int fmt, void* stream);
/* per-channel sums (fp64 accumulators): bias gradients, BN statistics */
int mpn_channel_sums(const void* x_hi, const void* x_lo, long long pixels, int C, int cstride, int coffset, int fmt,
                     double* sum, double* sumsq_or_null, void* stream);
int mpn_double_to_float(const double* a, float* out, int n, float scale, void* stream);
/* BatchNorm2d, training mode (fpn.py:15-25 under model.train(), trainer.py:170-174): batch mean / biased variance,
 * running-stat update (momentum 0.1, unbiased variance), normalise (+residual, +ReLU), and the backward pass.
 * workspace: 2*C doubles. */
int mpn_bn_stats(const void* y_hi, const void* y_lo, long long pixels, int C, int fmt, float* mean, float* var, double* workspace, void* stream);
int mpn_bn_update_running(const float* mean, const float* var, float* running_mean, float* running_var, long long n, float momentum, int C, void* stream);
int mpn_bn_apply(const void* y_hi, const void* y_lo, const float* mean, const float* var, const float* gamma, const float* beta, float eps,
                 const void* res_hi, const void* res_lo, int relu, void* z_hi, void* z_lo, long long pixels, int C, int fmt,
                 float* coef_workspace /* 3*C floats */, void* stream);
/* g = dz * (z > 0 if relu); dy = gamma*invstd*(g - mean(g) - yhat*mean(g*yhat)); dgamma = sum g*yhat; dbeta = sum g;
 * g_out (optional) receives g (the gradient of the residual / shortcut input). */
int mpn_bn_backward(const void* dz_hi, const void* dz_lo, const void* z_hi, const void* z_lo, const void* y_hi, const void* y_lo,
                    const float* mean, const float* var, const float* gamma, float eps, int relu, long long pixels, int C, int fmt,
                    void* dy_hi, void* dy_lo, void* g_hi, void* g_lo, float* dgamma, float* dbeta, double* workspace,
                    float* coef_workspace /* 3*C floats */, void* stream);
int mpn_relu_backward(const void* dz_hi, const void* dz_lo, const void* z_hi, const void* z_lo, void* out_hi, void* out_lo, long long n, int fmt, void* stream);
int mpn_add_act(const void* a_hi, const void* a_lo, const void* b_hi, const void* b_lo, void* out_hi, void* out_lo, long long n, int fmt, void* stream);
int mpn_maxpool3x3s2_backward(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, void* dx_hi, void* dx_lo,
                              int N, int H, int W, int C, int fmt, void* stream);
/* backward of a nearest upsample by an integer factor r: coarse = sum of the r x r fine block (channel slice of the fine tensor) */
int mpn_block_sum(const void* fine_hi, const void* fine_lo, int f_cstride, int f_coffset, void* coarse_hi, void* coarse_lo,
                  int N, int h, int w, int C, int r, int fmt, void* stream);
/* weighted MSE of one supervised heat map (posenet.py:380-387): *loss += mean((pred*w - gt*w)^2) over [B,18,H,W];
 * dpred (NHWC, Cd >= 18 channels, extra channels zero) = grad_scale * d loss / d pred. pred fp32 NCHW [B,Cp,H,W]. */
int mpn_mse_heatmap_loss(const float* pred, const float* gt, const float* weight, int B, int Cp, int H, int W, double* loss,
                         void* dpred_hi, void* dpred_lo, int Cd, int fmt, float grad_scale, void* stream);

/* ---- detection post-process: anchors.py:21-37, utils.py:19-61, posenet.py:269-285, lib/nms */
/* number of anchors for an image (levels 3..7, 9 per cell) */
int mpn_num_anchors(int H, int W);
/* HOST output [A,4] fp32, bit-identical to Anchors.forward (float64 math, cast to fp32) */
int mpn_generate_anchors(int H, int W, float* anchors_host);
/* boxes[b,a,:] = clip(decode(anchors[a], reg[b,a,:])); device pointers; reg [B,A,4], boxes [B,A,4];
 * H <= 0 or W <= 0: decode only (no clipping) */
int mpn_decode_clip(const float* anchors, const float* reg, float* boxes, int B, int A, int H, int W, void* stream);
/* Per image: select a where cls[b,a] > score_thresh (order preserved, posenet.py:271-279), sort by
 * score descending (stable), NMS (IoU > iou_thresh, or >= if ge), and write
 *   cand_idx  [B,max_cand] int32 : anchor index of each candidate in filtered order
 *   cand_cnt  [B]          int32 : number of candidates (N_s)
 *   keep_idx  [B,max_cand] int64 : kept indices INTO THE FILTERED SET, descending score (== pth_nms)
 *   keep_cnt  [B]          int32
 *   out_scores[B,max_cand] fp32, out_boxes [B,max_cand,4] fp32 : gathered per kept index
 * Candidates beyond max_cand are dropped (cand_cnt still reports the true count so the host can detect it).
 * workspace: mpn_detect_workspace_bytes(B, A, max_cand) bytes of device memory. */
size_t mpn_detect_workspace_bytes(int B, int A, int max_cand);
int mpn_filter_sort_nms(const float* cls, const float* boxes, int B, int A, float score_thresh, float iou_thresh, int ge,
                        int max_cand, int32_t* cand_idx, int32_t* cand_cnt, int64_t* keep_idx, int32_t* keep_cnt,
                        float* out_scores, float* out_boxes, void* workspace, size_t workspace_bytes, void* stream);
/* Profiling twin of mpn_filter_sort_nms (bench.py roofline_aux): same work and outputs, plus CUDA events between the stages on
 * `stream`; BLOCKS until the last stage finished and writes stage_ms[5] (host) = filter/compact, segment sort, gather, IoU
 * bit-mask, greedy reduction. */
int mpn_filter_sort_nms_profile(const float* cls, const float* boxes, int B, int A, float score_thresh, float iou_thresh, int ge,
                                int max_cand, int32_t* cand_idx, int32_t* cand_cnt, int64_t* keep_idx, int32_t* keep_cnt,
                                float* out_scores, float* out_boxes, void* workspace, size_t workspace_bytes, void* stream,
                                float* stage_ms);
/* pth_nms equivalents on raw device memory.  dets [n,5] (x1,y1,x2,y2,score), any order.
 * keep [n] int64 (device), num_out [1] int32 (device). Mirrors gpu_nms (ge=0) / cpu_nms (ge=1) + the
 * sort/gather of pth_nms.py:25-44. workspace: mpn_nms_workspace_bytes(n). */
size_t mpn_nms_workspace_bytes(int n);
int mpn_nms(const float* dets, int n, float iou_thresh, int ge, int64_t* keep, int32_t* num_out,
            void* workspace, size_t workspace_bytes, void* stream);
/* mask stage alone on already-sorted dets (== _nms() of nms_kernel.cu:73): mask [n, ceil(n/64)] u64 */
int mpn_nms_mask(const float* sorted_dets, int n, float iou_thresh, int ge, uint64_t* mask, void* stream);

/* ---- heat-map peak extraction (the step after the path)
 * network/joint_utils.py:19-32 (find_peaks: 3x3-cross maximum == value and value > thre1), :61-138 (NMS: every peak's
 * <=5x5 patch upsampled x factor with cv2.INTER_CUBIC, first arg-max = refined location and score, ids counting up over
 * the joint types) and :141-152 (get_joint_list rows), as called by evaluate/tester.py:215-221 with factor = 480/120.
 *   heat  : fp32 [B][C][H][W] (image_stride elements between images, >= C*H*W: pass the 18 used channels of a 19-channel map)
 *   peaks : fp32 [B][max_peaks][5] rows (x, y, score, id, joint_type) in the reference's order (joint type, then row-major);
 *           x, y are integers in the factor-times-upsampled grid (the caller multiplies by its image scale, :146-147)
 *   count : int32 [B] number of peaks found (rows beyond max_peaks are counted, not written)
 * The bicubic arithmetic is OpenCV's (Keys A = -0.75, half-pixel centres, clamped taps) in unfused fp32. */
size_t mpn_heatmap_peaks_workspace_bytes(int B, int C);
int mpn_heatmap_peaks(const float* heat, int B, int C, int H, int W, long long image_stride, float thre1, int factor,
                      float* peaks, int max_peaks, int32_t* count, void* workspace, size_t workspace_bytes, void* stream);

/* ---- PRN assignment: evaluate/tester.py:333-513 (Tester.prn_process) for every person box of a
 * batch of images.  Peaks are the joint-list rows of the images regrouped by (image, joint type) in their original order
 * (:337-350, the neck row dropped by the caller, tester.py:224-229); a peak's id is its index in that order.
 *   peak_xy        f64 [n_peaks][2]   peak_type i32 [n_peaks] (0..16)   peak_img_start i32 [B+1]
 *   joint_start    i32 [B][18]  first peak of joint type t of image b (entry 17 = end of the image's peaks)
 *   boxes_xywh     f64 [P][4]   (x1, y1, x2-x1, y2-y1) (:356-358), boxes sorted by image
 *   box_img        i32 [P]      box_img_start i32 [B+1]
 *   gauss_w_host   f64 [5]      HOST: scipy's normalised sigma=1 weights, index = distance from the centre
 *   owner          i32 [P][17][gh][gw]  id of the peak in each grid cell (-1 = none), written by build_inputs
 *   inp / output   f32 [P][gh][gw][17]  PRN input (gaussian-blurred one-hots, :396-403) / PRN output (posenet.py:337-350)
 *   bbox_keypoints f64 [P][17][3]       (x, y, 1) of the assigned peak, (x, y, 0) of the arg-max fallback, else zeros (:451-483)
 * kmax >= the largest number of peaks of one image.  workspace is shared by the two calls (build_inputs leaves the
 * per-plane occupancy there for assign).  Index decisions use unfused float64 / float32 arithmetic in the reference's
 * operation order (scipy's correlate1d, numpy's pairwise float32 sum). */
size_t mpn_prn_workspace_bytes(int P, int n_peaks, int kmax);
int mpn_prn_build_inputs(const double* peak_xy, const int32_t* peak_type, const int32_t* peak_img_start, int n_peaks,
                         const double* boxes_xywh, const int32_t* box_img, int P, int gh, int gw, double in_thres,
                         const double* gauss_w_host, int32_t* owner, float* inp, void* workspace, size_t workspace_bytes,
                         int kmax, void* stream);
int mpn_prn_assign(const double* peak_xy, const int32_t* peak_img_start, const int32_t* joint_start, int n_peaks,
                   const double* boxes_xywh, const int32_t* box_img, const int32_t* box_img_start, int P, int B, int gh, int gw,
                   const int32_t* owner, const float* output, double* bbox_keypoints, void* workspace,
                   size_t workspace_bytes, int kmax, void* stream);

/* ---- Multi-scale / flip test-time augmentation, heat-map side: evaluate/tester.py:298-304, :316-331.
 * mpn_resize_cubic = cv2.resize(..., interpolation=cv2.INTER_CUBIC) on `planes` fp32 planes: the sh x sw valid region of every
 * source plane (row pitch src_pitch, plane stride src_plane elements) -> dh x dw.  scale_x / scale_y = OpenCV's source step per
 * destination pixel (1 / fx for cv2.resize(None, fx, fy); 1 / (dw / sw) for an explicit dsize).
 *   out_mode 0: dst fp32 [planes][dh][dst_pitch] = resized
 *   out_mode 1: dst fp64 += (double)(resized / div)   (tester.py:304 heatmap_avg + heatmap / len(multiplier)), written to
 *               plane plane_map[p] (device int32, may be NULL) and mirrored along x when mirror != 0 (the flipped pass of
 *               Tester._handle_heat, tester.py:316-331)
 * Same Keys A = -0.75 arithmetic as OpenCV (float32 horizontal pass, then vertical); agreement with cv2 is within 1e-5 of the plane
 * maximum (measured 3.9e-6), not bit-exact (OpenCV's own SIMD and scalar paths differ in FMA use). */
size_t mpn_resize_cubic_workspace_bytes(int dh, int dw);
int mpn_resize_cubic(const float* src, long long src_plane, int src_pitch, int sh, int sw, void* dst, long long dst_plane,
                     int dst_pitch, int dh, int dw, int planes, double scale_x, double scale_y, int out_mode, float div,
                     int mirror, const int* plane_map, void* workspace, size_t workspace_bytes, void* stream);
/* out = (normal + flipped) / 2 in float64 (tester.py:329; flipped may be NULL: out = normal), optionally also as fp32 */
int mpn_tta_combine(const double* normal, const double* flipped, double* out, float* out32, long long n, void* stream);

/* ---- Detection-subnet training loss: network/losses.py:5-137 (calc_iou + FocalLoss.forward) for the whole
 * batch, and its gradients.
 *   cls fp32 [B][A][C] class scores AFTER the sigmoid (posenet.py:109-117), reg fp32 [B][A][4], anchors fp32 [A][4],
 *   annotations fp32 [B][M][5] = (x1, y1, x2, y2, class), rows with class == -1 are padding (datasets bbox_collater), M <= 256
 *   cls_loss / reg_loss fp32 [B]: the per-image values the reference stacks (losses.py:94,131); the caller takes the batch mean
 *   dcls [B][A][C] / dreg [B][A][4] (may be NULL): gradient of gscale_cls * mean_b(cls_loss) + gscale_reg * mean_b(reg_loss)
 * Anchor <-> annotation IoU without the +1 convention, first maximum, < 0.4 negative / >= 0.5 positive / else ignored, scores
 * clamped to [1e-4, 1 - 1e-4] (gradient passes inside the range), alpha .25, gamma 2, smooth-L1 beta 1/9, box targets divided
 * by (.1, .1, .2, .2).  fp32 per-anchor arithmetic in the reference's order, fp64 sums. */
size_t mpn_focal_loss_workspace_bytes(int B, int A);
int mpn_focal_loss(const float* cls, const float* reg, const float* anchors, const float* annotations, int B, int A, int C, int M,
                   float* cls_loss, float* reg_loss, float* dcls, float* dreg, float gscale_cls, float gscale_reg, void* workspace,
                   size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MPN_B200_H_ */
