/* cunet_b200.h -- C ABI of libcunet_b200.so, the sm_100a kernel library behind the CU-Net hot path.
 *
 * The reference (zhiqiangdon/CU-Net) has no FFI: its "lower layer" is PyTorch library calls made from
 * models/cu_net.py, cu-net.py, utils/quantize.py and pylib/Evaluation.py.  Each entry point below
 * replaces one group of those calls (reference file:line cited per function).  Conventions:
 *   - plain pointers and sizes, no torch types; every pointer is a DEVICE pointer unless stated;
 *   - the caller allocates every buffer, including workspaces;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no internal synchronisation, no host
 *     callbacks => CUDA-graph capturable;
 *   - return 0 on success, <0 on error; cunet_last_error() returns a thread-local message;
 *   - activations are NHWC ("pixel rows"): tensor [N*H*W][ld] of dtype CUNET_F32 / CUNET_BF16;
 *   - weights / gradients visible to the caller keep the reference layout [Cout][Cin][kh][kw] fp32.
 */
#ifndef CUNET_B200_H_
#define CUNET_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUNET_F32 0
#define CUNET_BF16 1
#define CUNET_MAX_SEG 8

const char* cunet_last_error(void);
int cunet_abi_version(void);

/* One source tensor of a virtual channel concat (torch.cat(inputs, 1), models/cu_net.py:13). */
typedef struct {
  const void* ptr;     /* [rows][ld] activations                                                   */
  const double* stats; /* [2*C]: per-channel sum, sum of squares of this tensor (train-mode BN)    */
  double inv_count;    /* 1 / (rows of this tensor) -- elements per channel                        */
  int C;               /* channels of this segment (multiple of 32)                                */
  int ld;              /* row stride in elements                                                    */
  int up;              /* 1: tensor is at half resolution; nearest x2 upsample fused (cu_net.py:250,265) */
  int reserved;
} cunet_seg;

/* A virtual concat + the BatchNorm(+ReLU) that precedes a conv: the "A operand" of every fused conv
 * (torch.cat -> norm -> relu of _bn_function_factory, models/cu_net.py:11-17). */
typedef struct {
  cunet_seg seg[CUNET_MAX_SEG];
  int nseg;
  int bn_train;       /* 1: batch statistics from seg[].stats; 0: running statistics;
                         2: identity (no BatchNorm, no ReLU: the stem's im2col operand) */
  const float* gamma; /* [Cin] BN weight, concat order                 */
  const float* beta;  /* [Cin] BN bias                                 */
  const float* rmean; /* [Cin] running mean  (eval mode)               */
  const float* rvar;  /* [Cin] running var   (eval mode)               */
  float eps;
  int act_bits;       /* 0: none.  2..8: QuanInput2d between the ReLU and the conv (utils/quantize.py:47-73;
                         models/cu_net_prev_version_wig.py:96-98,277-279): the activated operand becomes
                         a <- round(clamp(a, +-(1 - 2^-(bits-1))) * 2^(bits-1)) / 2^(bits-1); backward straight-through,
                         zero where |a| >= 1 */
} cunet_concat;

/* Fused  cat -> BatchNorm -> ReLU -> conv(1x1 | 3x3 pad 1)  [-> 2x2 maxpool]  forward.
 * Replaces _bn_function_factory + nn.Conv2d (models/cu_net.py:11-17, 24, 43, 47-48, 197) and the
 * nn.MaxPool2d that follows a down-block's adapters_ahead (models/cu_net.py:249, 260).
 * Epilogue also accumulates the per-channel sum / sum-of-squares of the tensor it writes, which is the
 * batch statistic every consumer BatchNorm of that tensor needs (nn.BatchNorm2d train mode). */
typedef struct {
  cunet_concat in;
  int N, H, W;        /* output resolution (before the optional pool) */
  int taps;           /* 1 (1x1) or 9 (3x3, pad 1)                     */
  const void* wpack;  /* packed weights from cunet_pack_weights (fwd image) */
  int Cout, CoutPad;  /* CoutPad: multiple of 16, <= 128              */
  void* out;          /* [rows][out_ld]                                */
  int out_ld;
  int out_fp32;       /* 1: store fp32 regardless of dtype (heatmap heads) */
  double* out_stats;  /* [2*Cout] accumulated (+=) or NULL             */
  uint8_t* pool_idx;  /* [N*H/2*W/2][Cout] argmax position (dy*2+dx) or NULL */
  int pool;           /* 1: write maxpool2x2(out) [N*H/2*W/2][out_ld] + pool_idx */
  int dtype;          /* CUNET_F32 (tf32 MMA) or CUNET_BF16           */
} cunet_conv_fwd_params;

int cunet_conv_fwd(const cunet_conv_fwd_params* p, void* stream);

/* The gradient w.r.t. a conv's OUTPUT tensor T, as the backward kernels consume it.
 * Every tensor of the network is consumed only through BatchNorm'd convs, so its gradient is
 *   dT = istd * (G - mean(G) - xhat * mean(G*xhat)),   G = sum over consumers of gamma_j * dz_j
 * (autograd of nn.BatchNorm2d train mode, summed over the consumers that share T's batch statistics).
 * The dgrad kernels of the consumers accumulate G and (sum G, sum G*xhat); the producer's backward kernels
 * evaluate dT on the fly per channel.  mode 0: dT = g (plain, e.g. dLoss/dhead).
 * pooled = 1: T is the 2x2-maxpooled conv output (g, t, pool_idx at half resolution); the gradient is
 * routed to the argmax position (autograd of nn.MaxPool2d, models/cu_net.py:249,260). */
typedef struct {
  const void* g;
  const void* t;
  const double* stats;     /* [2*C] sum, sumsq of T (forward)         */
  const double* gstats;    /* [2*C] sum G, sum G*xhat                  */
  const uint8_t* pool_idx; /* [rows][C] or NULL                        */
  double inv_count;        /* 1 / rows of T                            */
  int C, ld;
  int mode;                /* 0 plain, 1 batch-norm backward form      */
  int pooled;
  float eps;
  int reserved;
} cunet_grad_src;

/* Per-source accumulator written by a consumer's dgrad epilogue. */
typedef struct {
  void* G;        /* [rows][ld] dtype; NULL: this source needs no gradient  */
  double* gstats; /* non-NULL: add this consumer's share of (sum G, sum G*xhat) = gamma*(dbeta, dgamma); every
                     consumer of the tensor passes it, the buffer is zeroed once per backward pass          */
  int ld;
  int accumulate; /* 0: first consumer in backward order (write), 1: read-modify-write */
} cunet_gacc;

/* Backward-data of the fused conv (autograd of nn.Conv2d + ReLU + BatchNorm2d + cat/upsample split,
 * SURVEY.md section 8 A15): dA = dY * W, then per source: dz = dA * [bn(x) > 0],
 * dgamma += sum dz*xhat, dbeta += sum dz, G_src += gamma*dz (summing the 4 children of an upsampled src). */
typedef struct {
  cunet_concat in;         /* the conv's input side, exactly as in the forward call */
  cunet_gacc gacc[CUNET_MAX_SEG];
  cunet_grad_src dy;       /* gradient of the conv's output */
  int N, H, W, taps;
  const void* wpack_dgrad; /* dgrad image from cunet_pack_weights */
  int Cout, CoutPad;
  float* dgamma;           /* [Cin] fp32, accumulated with atomics (caller zeroes once per step) */
  float* dbeta;            /* [Cin] */
  int dtype;
  int reserved;
} cunet_conv_dgrad_params;

int cunet_conv_dgrad(const cunet_conv_dgrad_params* p, void* stream);
/* debug aid: device buffer (>= 512 int64) that receives CTA 0's clock64() timeline of the persistent bf16 dgrad
 * kernel (tools/trace_dgrad.py); NULL disables. */
int cunet_debug_dgrad_trace(void* buf);

/* Backward-filter of the fused conv: dW[co][k][tap] += sum_px dY[px][co] * relu(bn(x))[px+tap][k]
 * (autograd of nn.Conv2d w.r.t. weight), accumulated into the reference-layout fp32 gradient. */
typedef struct {
  cunet_concat in;
  cunet_grad_src dy;
  int N, H, W, taps;
  int Cout;
  float* dw;               /* [Cout][Cin][taps] fp32, atomically accumulated */
  int nsplit;              /* CTAs along the pixel dimension (0: library default) */
  int dtype;
  int dw_cin;              /* row length (Cin) of dw; 0: the concat's channel count.  Channels >= dw_cin are padding */
  int reserved;
} cunet_conv_wgrad_params;

int cunet_conv_wgrad(const cunet_conv_wgrad_params* p, void* stream);

/* debug / experiment switch: route bf16 1x1 forward calls with >= min_tiles 128-pixel tiles (and no pooling) to the
 * persistent bulk-landing kernel (csrc/conv_fwd_v2.cu); < 0 (the default) keeps the round-1 kernel.  Returns the
 * previous setting.  Also settable through the environment (CUNET_FWD_V2_MIN_TILES). */
int cunet_debug_fwd_v2_min_tiles(int min_tiles);
/* same for the third-generation kernel (csrc/conv_fwd_v3.cu: resident weights, transposed GEMM with per-channel
 * statistics in registers, pooling fused): bf16 1x1 forward calls with at least min_rows pixel rows use it (default
 * 0 = every eligible call, CUNET_FWD_V3_MIN_ROWS); negative disables.  Returns the previous setting. */
long cunet_debug_fwd_v3_min_rows(long min_rows);

/* Fused backward of the dense-layer 3x3 conv (models/cu_net.py:47-48, conv2 128 -> 32): exactly
 * cunet_conv_dgrad(d) followed by cunet_conv_wgrad(w) for the SAME op (same `in`, same `dy`), in one launch.
 * bf16, one 128-channel source, Cout == 32, batch-norm-form dy, W <= 64 run in the fused persistent kernel
 * (the im2col of the output gradient is built once in shared memory and feeds both contractions); every
 * other configuration falls back to the two calls above, in that order, on `stream`. */
int cunet_conv_bwd3x3(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, void* stream);

/* Fused backward of a 1x1 fused conv -- adapters (models/cu_net.py:19-35), dense-layer conv1 (:41-44, 52-61),
 * intermedia adapters (:147-190), heat-map heads (:192-198): exactly cunet_conv_dgrad(d) followed by
 * cunet_conv_wgrad(w) for the SAME op (same `in`, same `dy`), in one launch that lands every operand once
 * (csrc/conv_bwd1x1.cu).  bf16, taps == 1, Cin <= 384 in 32/64/128-channel segments, power-of-two H, W in [4, 64]
 * run in the fused persistent kernel; every other configuration falls back to the two calls above, in that
 * order, on `stream`. */
int cunet_conv_bwd1x1(const cunet_conv_dgrad_params* d, const cunet_conv_wgrad_params* w, void* stream);

/* ---- stem: conv0 7x7 s2 p3 -> norm0 -> relu0 -> pool0 (models/cu_net.py:299-304) ------------------
 * conv0 runs on the tensor cores through cunet_conv_fwd / cunet_conv_wgrad with an identity input
 * (bn_train == 2) over an im2col matrix of 160 columns (147 = 3*7*7 in the reference's weight-flattening
 * order c*49 + kh*7 + kw, zero padded to 160), stored as TWO dense column blocks -- columns 0..127 as
 * [N*Ho*Wo][128], then columns 128..159 as [N*Ho*Wo][32] -- i.e. as two concat segments (C = ld = 128 and 32). */
int cunet_stem_im2col(const float* img /* [N][3][Hi][Wi] fp32, NCHW as the reference feeds it */,
                      void* cols /* [N*Ho*Wo][128] ++ [N*Ho*Wo][32] dtype */, int N, int Hi, int Wi, int dtype,
                      void* stream);

/* norm0 -> relu0 -> pool0 forward: y [N*H*W][128] -> x [N*H/2*W/2][128], accumulating the statistics of x.
 * bn_train: 1 batch statistics from y_stats, 0 running statistics. */
typedef struct {
  const void* y;
  const double* y_stats; /* [256] sum, sumsq of y (train) */
  const float* gamma; const float* beta; const float* rmean; const float* rvar; /* [128] norm0 */
  void* x;
  double* x_stats;       /* [256] += */
  int N, H, W;           /* resolution of y */
  int bn_train;
  float eps;
  int dtype;
} cunet_stem_pool_params;
int cunet_stem_pool_fwd(const cunet_stem_pool_params* p, void* stream);

/* Backward of pool0 / relu0 / norm0 (autograd of nn.MaxPool2d, nn.ReLU, nn.BatchNorm2d train mode).
 * phase 0: dgamma0 += sum dz*yhat, dbeta0 += sum dz      (dz = gradient w.r.t. norm0's output)
 * phase 1: dy = gamma*istd*(dz - dbeta0/n - yhat*dgamma0/n)  written as [N*H*W][128] dtype
 * The gradient of x arrives in the batch-norm backward form (cunet_grad_src mode 1). */
typedef struct {
  const void* y;
  const double* y_stats;
  const float* gamma; const float* beta;
  cunet_grad_src dx;     /* gradient of x (pooled resolution), mode 1 */
  float* dgamma; float* dbeta; /* [128] fp32 (phase 0: accumulated; phase 1: read) */
  void* dy;              /* phase 1 output */
  int N, H, W;
  float eps;
  int dtype;
  int phase;
} cunet_stem_bwd_params;
int cunet_stem_bwd(const cunet_stem_bwd_params* p, void* stream);

/* ---- loss + decode ------------------------------------------------------------------------------
 * Multi-loss MSE  sum_k mean((out_k - heatmap)^2)  (cu-net.py:175-178) with its gradient, fused with the
 * argmax landmark decode of the last head (pylib/Evaluation.py:6-23: first maximum, 1-based (x, y),
 * zero where max <= 0).  Heads are NHWC fp32 [N*H*W][ld]; the target is NCHW fp32 [N][C][H][W]. */
typedef struct {
  const float* heads[16]; /* loss_num head outputs */
  void* dheads[16];       /* gradients [N*H*W][ld] dtype, or all NULL (loss / decode only) */
  int nheads;
  const float* target;
  int N, C, H, W, ld;
  double* loss;           /* [1 + nheads]: total, then per head; accumulated (+=) */
  unsigned long long* keys; /* [N*C] zero-initialised scratch for the decode */
  float grad_scale;       /* multiplies the gradient (1 / world_size under data parallelism) */
  int dtype;
} cunet_mse_params;
int cunet_mse_decode(const cunet_mse_params* p, void* stream);
/* keys -> preds [N][C][2] fp32 (x, y) 1-based, masked by max > 0 */
int cunet_decode_finalize(const unsigned long long* keys, float* preds, int NC, int W, void* stream);

/* ---- BatchNorm running statistics (nn.BatchNorm2d momentum update) ---------- */
typedef struct {
  const double* stats[CUNET_MAX_SEG]; /* per segment [2*C] */
  double inv_count[CUNET_MAX_SEG];
  int C[CUNET_MAX_SEG];
  int nseg;
  int reps;          /* 1, or 2 for BatchNorms that the reference re-runs under checkpointing */
  float* rmean; float* rvar;  /* [Cin] */
  double n_elems;    /* elements per channel seen by this BatchNorm (for the unbiased variance) */
  float momentum;
  int reserved;
} cunet_bn_update_desc;
int cunet_bn_running_update(const cunet_bn_update_desc* descs_dev, int ndesc, void* stream);

/* ---- optimizer: torch.optim.RMSprop(lr, alpha, eps, momentum=0, weight_decay=0) (cu-net.py:60-61) ---- */
int cunet_rmsprop_step(float* params, const float* grads, float* square_avg, long n, const float* lr_dev,
                       float alpha, float eps, void* stream);

/* ---- weight / activation / gradient quantizers ---------------------------------------------------------
 * QuanOp (utils/quantize.py:77-175) and BinOp (models/cu_net_prev_version.py:17-92) as multi-tensor launches:
 * one thread block per filter (output channel) of every target conv, all targets in one launch.
 *   forward (mode 0, QuanOp.quantization, quantize.py:104-149):
 *       w -= mean over Cin;  w = clamp(w, +-(1 - 2^-(bits_g-1)));  saved = round(w*2^(bits_g-1))/2^(bits_g-1);
 *       bits_w == 1: w = sign(w) [the scale is dropped by the fall-through at quantize.py:135,148-149]
 *       bits_w == 2: w = +1 if w > d, -1 if w < -d, else 0,  d = 0.7 * mean_filter |w|
 *       else       : w = round(clamp(w)*2^(bits_w-1))/2^(bits_w-1)
 *   forward (mode 1, BinOp.binarization, cu_net_prev_version.py:43-72):
 *       w -= mean over Cin;  w = clamp(w, +-1);  saved = w;  w = sign(w) * mean_filter |w|
 *   restore: w = saved                                         (quantize.py:151-153, cu_net_prev_version.py:74-76)
 *   grad (quantize.py:156-175, cu_net_prev_version.py:78-92), w = restored weights:
 *       bits_w == 1 or BinOp: g = (m*g + sign(w)*mean_filter(sign(w)*g)) * (1 - 1/Cin) * n,
 *                            m = mean_filter|w| where |w| <= 1 else 0  [QuanOp: m rounded to the bits_g grid]
 *       QuanOp: g = round(clamp(g)*2^(bits_g-1))/2^(bits_g-1) afterwards (and only that when bits_w != 1) */
typedef struct {
  float* w;        /* [Cout][Cin][taps] fp32 parameter (modified in place) */
  float* saved;    /* same shape: full-precision copy */
  float* grad;     /* same shape, or NULL */
  int Cout, Cin, taps;
  int first_block; /* index of this tensor's first filter in the launch grid */
} cunet_quant_desc;
int cunet_quant_forward(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, int mode, int bits_w, int bits_g,
                        void* stream);
int cunet_quant_restore(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, void* stream);
int cunet_quant_grad(const cunet_quant_desc* descs_dev, int ndesc, int nblocks, int mode, int bits_w, int bits_g,
                     void* stream);
/* QuanInput (utils/quantize.py:47-63): y = Q(C(x, bits), bits);  backward: dx = dy where |x| < 1 else 0 */
int cunet_quant_input_fwd(const float* x, float* y, long n, int bits, void* stream);
int cunet_quant_input_bwd(const float* x, const float* dy, float* dx, long n, void* stream);

/* Weight packing: reference-layout fp32 master weights -> tensor-core operand images.
 * One descriptor per conv; all descriptors processed by one launch.
 *   fwd image  : [tap][kb][CoutPad rows][128 B]  K = input channel  (B operand of the forward GEMM)
 *   dgrad image: [chunk][kb][128 rows][128 B]    rows = input channel, K = (tap, output channel)
 */
typedef struct {
  const float* w;  /* [Cout][Cin][k][k] fp32                               */
  void* fwd;       /* fwd image or NULL                                     */
  void* dgrad;     /* dgrad image or NULL                                   */
  int Cout, Cin, taps, CoutPad;
} cunet_pack_desc;

int cunet_pack_weights(const cunet_pack_desc* descs_dev, int ndesc, int dtype, int max_chunks, void* stream);
/* bytes of the images for one conv */
long cunet_pack_fwd_bytes(int Cin, int taps, int CoutPad, int dtype);
long cunet_pack_dgrad_bytes(int Cin, int taps, int CoutPad, int dtype);

#ifdef __cplusplus
}
#endif
#endif /* CUNET_B200_H_ */
