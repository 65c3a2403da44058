/*
 * icf.h — C-ABI of the B200-native conditional-BiGAN hot path (libicf_b200.so).
 *
 * The reference (wtaylor17/ImageCFGen-Pytorch) has no FFI: its hot path is the stock torch.nn layers
 * called from image_scms/{mnist,audio_mnist,whalecalls,esrf_acoustic}.py.  Each entry point below
 * replaces the torch operator(s) named in its comment (reference file:line); the Python host side in
 * imagecfgen-pytorch_b200/ binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller (PyTorch's
 *    caching allocator); the library never allocates or frees tensor memory and never synchronises.
 *  - every launch goes to the `stream` argument (a cudaStream_t passed as void*).
 *  - return 0 on success; non-zero -> icf_last_error() (thread-local) describes the failure. Invalid
 *    shapes / alignments are reported, never abort()ed. Asynchronous CUDA faults surface at the caller's
 *    next synchronisation, as with torch.
 *  - activations are NHWC ("pixel-major"): tensor[n][y][x][pitch], `pitch` >= channels, both multiples
 *    of 8 for bf16 tensor-core operands. At the module boundary C==1 or H==W==1, so NCHW == NHWC.
 *  - re-entrant: may be called from the Python thread and from torch's autograd worker thread.
 */
#ifndef ICF_H_
#define ICF_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ICF_VERSION 1

enum { ICF_F32 = 0, ICF_BF16 = 1 };
enum { ICF_ACT_NONE = 0, ICF_ACT_LRELU = 1, ICF_ACT_TANH = 2 };
enum { ICF_FORM_GATHER = 0,    /* src = dst*stride - pad + tap           (Conv2d fprop, ConvT dgrad) */
       ICF_FORM_TRANSPOSED = 1 /* src = (dst + pad - tap)/stride, exact  (ConvT fprop, Conv2d dgrad) */ };

const char* icf_last_error(void);
int icf_version(void);
/* 1 when the tcgen05/TMEM implicit-GEMM path is compiled in and allowed (env ICF_DISABLE_TC unset). */
int icf_tc_enabled(void);
void icf_set_tc_enabled(int on);
/* Kernel family the calling thread's last icf_conv_forward used: 0 SIMT fp32 FMA, 1 per-tap tcgen05 implicit GEMM,
 * 2 weight-stationary row-streaming tcgen05 kernel, 3 scatter-form tcgen05 kernel (<= 8 output channels), 4 channel-major
 * tcgen05 kernel (folded unit-stride layers on 16-byte pixels), -1 none yet (tests and profiles assert the intended path). */
enum { ICF_PATH_SIMT = 0, ICF_PATH_TC = 1, ICF_PATH_WS = 2, ICF_PATH_SC = 3, ICF_PATH_CM = 4 };
int icf_last_conv_path(void);

/* ------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution, forward and data-gradient.
 * Replaces aten::convolution / transposed convolution of nn.Conv2d, nn.ConvTranspose2d, nn.Linear
 * (image_scms/mnist.py:31-39,64-72,100-135; audio_mnist.py:187-197,226-242,273-302) and the dgrad half
 * of aten::convolution_backward, fused with bias + LeakyReLU/Tanh (+ Dropout2d mask, + BatchNorm
 * statistics of the result).
 *   dst[n,p,q,k] = mask[n,k] * act( bias[k] + sum_{r,s,c} src[n, y(p,r), x(q,s), c] * w[k][r*S+s][c] )
 * bf16: tcgen05.mma with TMEM accumulators, operands staged by TMA; f32: SIMT FMA.
 * When K is not a multiple of 8 and out_pitch leaves room, the bf16 kernels also store zeros into the pitch
 * padding dst[..., K .. roundup8(K)-1] (16-byte stores); that range must not alias live data.
 * ------------------------------------------------------------------------------------------------ */
typedef struct icf_conv_args {
  int32_t dtype;                 /* ICF_F32 | ICF_BF16: element type of src, w and (unless out_f32) dst */
  int32_t form;                  /* ICF_FORM_* */
  int32_t N;
  int32_t H, W, C, in_pitch;     /* src [N][H][W][in_pitch], C channels reduced over */
  int32_t P, Q, K, out_pitch;    /* dst [N][P][Q][out_pitch], K channels produced */
  int32_t R, S, stride, pad;
  int32_t w_rows;                /* rows physically present in w (>= K; extra rows must be zero) */
  int32_t w_pitch;               /* channels physically present per (row, tap) of w (>= C, zero padded) */
  int32_t act;                   /* ICF_ACT_* */
  float slope;                   /* LeakyReLU negative slope */
  int32_t out_f32;               /* 1: dst elements are float even when dtype == ICF_BF16 */
  int32_t mask_pitch;            /* row pitch of out_mask */
  int32_t accumulate;            /* 1: dst += result (f32 dst only; used by split dgrad) */
  int32_t win;                   /* >1 (tensor-core gather form only): the S filter columns are folded into the
                                    channel dimension — the operand row of (pixel, filter row r) is the contiguous
                                    run of win*in_pitch elements starting at that pixel; pass S = 1, C = win*in_pitch
                                    and w packed [rows][R][w_pitch].  Used for the small-channel first layers. */
  const void* src;
  const void* w;                 /* packed [w_rows][R*S][w_pitch], channels contiguous */
  const float* bias;             /* [K] or NULL */
  void* dst;
  const float* out_mask;         /* [N][mask_pitch] Dropout2d mask incl. 1/(1-p) scale, or NULL */
  float* stats;                  /* [2][K]: += sum, += sum of squares of dst values (BatchNorm), or NULL */
} icf_conv_args;
int icf_conv_forward(const icf_conv_args* a, void* stream);
/* Same operation with split-K allowed: `partial` is a ZEROED fp32 scratch of `partial_elems` floats supplied by the caller
 * (>= N*P*Q * K rounded up to 256).  When the tile grid of the layer alone would leave most SMs idle (a few hundred output pixels
 * against a 10-50 MB weight: the 1024-channel layers of the spectrogram families at their default batches) the (tap, channel
 * chunk) iterations of every tile are dealt to several CTAs that add raw fp32 sums into `partial`, and a finishing pass applies
 * bias / activation / Dropout2d mask.  `partial` is left dirty.  Otherwise (and for stats != NULL) this IS icf_conv_forward. */
int icf_conv_forward_splitk(const icf_conv_args* a, float* partial, int64_t partial_elems, void* stream);

/* Introspection / test hook, no kernel launch and no GPU needed: the plan the weight-stationary row-streaming kernel
 * would use for `a` (pointers in `a` only need to be 16-byte aligned, they are not dereferenced).  Returns 0 and fills
 * `out`, -1 when the geometry is outside that kernel's envelope, > 0 on error.  `out` needs 16 + 4*48 + 4096 int32 words:
 *   [0] classes (output-parity classes of a transposed conv, else 1)  [1] XG  [2] NG  (a column = XG output columns x NG
 *   images = 128 MMA rows)  [3] ring slots  [4] TMEM accumulators  [5] channel tile  [6] source step  [7] output step
 *   [8] issuer warps  [9] schedule words in total  [10] grid  [11] channel tiles  [12] image groups  [13] schedule bytes in smem
 *   per class c at [16 + 48*c]: Pi, Qj, py, px, ylo, yhi, dymax, groups, taps, x tiles, first CTA, CTAs,
 *                               schedule offsets of issuer 0..issuers (end), then per group g at [+16+4g]: dy, dprev, first tap, taps
 *   schedule words at [16 + 4*48]: bits 0-7 output row | 8-12 first tap | 13-17 tap count | 18-19 kind (0 MMA chain,
 *                               1 accumulator complete, 2 nothing) | 20 chain opens the accumulator | 21 last action of its source row;
 *                               every issuer's list ends with one spare word.
 * tests/test_host_cpu.py replays the schedules and checks the accumulator protocol for every layer of every family. */
int icf_ws_plan(const icf_conv_args* a, int32_t* out, int32_t out_words);

/* Weight gradient (the wgrad half of aten::convolution_backward), fp32 accumulation, split over pixels:
 *   dw[a][r*S+s][b] += sum_{n,p,q} small[n,p,q,a] * big[n, p*stride-pad+r, q*stride-pad+s, b]
 * Conv2d: small = dY, big = X.  ConvTranspose2d: small = X, big = dY. */
typedef struct icf_wgrad_args {
  int32_t dtype;
  int32_t N;
  int32_t P, Q, A, a_pitch;
  int32_t H, W, B, b_pitch;
  int32_t R, S, stride, pad;
  const void* small_t;
  const void* big_t;
  float* dw;                     /* [A][R*S][B] fp32, caller zeroes */
  int32_t win;                   /* >1: as in icf_conv_args — big's row is win*b_pitch contiguous elements; S = 1 */
} icf_wgrad_args;
int icf_conv_wgrad(const icf_wgrad_args* a, void* stream);

/* Introspection / test hook like icf_ws_plan, for the tensor-core weight-gradient kernel (no launch, no GPU needed; returns
 * -1 when that kernel declines the geometry).  `out` (>= 16 int32 words): [0] channel tile of the big operand  [1] taps per
 * CTA  [2] tap groups  [3] pipeline stages  [4] bytes per stage  [5] TMEM columns  [6] dynamic shared memory  [7] CTAs per SM
 * assumed  [8] grid.x = channel tiles x tap groups  [9] grid.y = split of the pixel blocks  [10..12] pixel-block box
 * (columns, rows, images)  [13] pixels per block  [14] pixel blocks  [15] taps */
int icf_wgrad_plan(const icf_wgrad_args* a, int32_t* out, int32_t out_words);

/* ------------------------------------------------------------------------------------------------
 * Weight (re)packing between the checkpoint layout (fp32 OIHW / IOHW / [out,in], App. A.5 of SURVEY.md)
 * and the K-major operand layout the conv kernels read.  Generic 3-index permutation:
 *   pack:    dst[i0][i1][i2] (dense, i2 padded to d2_pad, rows padded to rows_pad with zeros)
 *              = src[i0*s0 + i1*s1 + i2*s2]
 *   unpack:  dst[i0*s0 + i1*s1 + i2*s2] (=|+=) src[i0][i1][i2]        (fp32 -> fp32, for gradients)
 * ------------------------------------------------------------------------------------------------ */
typedef struct icf_perm {
  int64_t d0, d1, d2;            /* logical extents */
  int64_t s0, s1, s2;            /* element strides in the checkpoint-layout tensor */
  int64_t d2_pad;                /* physical extent of i2 in the packed tensor (>= d2) */
  int64_t d0_pad;                /* physical extent of i0 in the packed tensor (>= d0) */
} icf_perm;
int icf_pack(const float* src, void* dst, int32_t dst_dtype, const icf_perm* p, void* stream);
int icf_unpack(const float* src_packed, float* dst, const icf_perm* p, int32_t atomic_add, void* stream);
/* 4-index variant for the folded ("win") first-layer operands:
 *   packed[(i0*d1 + i1)*row_pitch + i2*d3_pad + i3]  <->  ref[i0*s0 + i1*s1 + i2*s2 + i3*s3]   (zero padding) */
typedef struct icf_perm4 {
  int64_t d0, d1, d2, d3;
  int64_t s0, s1, s2, s3;
  int64_t d3_pad;                /* >= d3 */
  int64_t row_pitch;             /* >= d2*d3_pad */
} icf_perm4;
int icf_pack4(const float* src, void* dst, int32_t dst_dtype, const icf_perm4* p, void* stream);
int icf_unpack4(const float* src_packed, float* dst, const icf_perm4* p, void* stream);

/* All re-packing jobs of a network in ONE launch (after an optimiser step every layer's operand copies are stale:
 * ~30 tiny launches per network otherwise).  `jobs` is an array in DEVICE memory, uploaded once by the caller;
 * kind 0 = icf_perm job (p), kind 1 = icf_perm4 job (p4); max_elems = largest padded element count of any job. */
typedef struct icf_pack_job {
  const float* src;
  void* dst;
  int32_t dst_dtype;
  int32_t kind;
  icf_perm p;
  icf_perm4 p4;
} icf_pack_job;
int icf_pack_multi(const icf_pack_job* jobs, int32_t n_jobs, int64_t max_elems, void* stream);
/* The reverse for gradients: every job copies its packed fp32 accumulator (src) into the checkpoint-layout gradient
 * (dst, assignment), icf_unpack / icf_unpack4 semantics; dst_dtype is ignored. */
int icf_unpack_multi(const icf_pack_job* jobs, int32_t n_jobs, int64_t max_elems, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attribute / latent feature assembly (mnist.py:47-55,77-85; audio_mnist.py:204-210,250-256).
 * ------------------------------------------------------------------------------------------------ */
#define ICF_MAX_PLANES 8
/* first-max argmax over rows of a one-hot / score matrix -> int32 (torch.argmax semantics) */
int icf_argmax_rows(const void* x, int32_t x_dtype /*0 f32,1 bf16,2 i32,3 i64*/, int32_t n, int32_t k,
                    int32_t* out, void* stream);

typedef struct icf_imgfeat_args {
  int32_t dtype;                 /* dtype of feat */
  int32_t N, H, W;
  int32_t feat_pitch;            /* channels physically present in feat (multiple of 8) */
  int32_t x_dtype;               /* dtype of x: ICF_F32 or ICF_BF16 */
  int32_t x_pitch;               /* elements between consecutive pixels of x (1 for a plain image) */
  int32_t n_emb, n_cont;
  int32_t mask_pitch;
  int32_t pad;                   /* fwd: feat is [N][H+2pad][W+2pad][feat_pitch] with a zero border (pre-padded conv input) */
  const void* x;                 /* [N][H][W][x_pitch], channel 0 is the image */
  const float* emb_table[ICF_MAX_PLANES];   /* [K_i][256] */
  const int32_t* emb_index[ICF_MAX_PLANES]; /* [N] */
  const float* cont[ICF_MAX_PLANES];        /* [N] constant-plane values */
  const float* mask;             /* [N][mask_pitch] Dropout2d mask on the feature stack, or NULL */
  void* feat;                    /* [N][H][W][feat_pitch]: ch0 image, then embedding planes, then constants */
  /* backward only */
  const void* dfeat;             /* [N][H][W][feat_pitch] gradient w.r.t. feat */
  float* demb_table[ICF_MAX_PLANES];        /* [K_i][256] += */
} icf_imgfeat_args;
int icf_image_features_fwd(const icf_imgfeat_args* a, void* stream);
int icf_image_features_bwd(const icf_imgfeat_args* a, void* stream);

typedef struct icf_latfeat_args {
  int32_t dtype;                 /* dtype of feat */
  int32_t N, latent;
  int32_t feat_pitch;            /* >= latent + 256*n_emb + n_cont, multiple of 8 */
  int32_t z_dtype, z_pitch;
  int32_t n_emb, n_cont;
  int32_t emb_k[ICF_MAX_PLANES];
  const void* z;                 /* [N][z_pitch] */
  const float* emb_table[ICF_MAX_PLANES];   /* [K_i][256] */
  const float* onehot[ICF_MAX_PLANES];      /* [N][K_i] dense (soft one-hots allowed) */
  const float* cont[ICF_MAX_PLANES];        /* [N] */
  void* feat;                    /* [N][feat_pitch] */
  /* backward only (any output may be NULL) */
  const void* dfeat;             /* [N][feat_pitch] */
  float* dz;                     /* [N][latent] = */
  float* demb_table[ICF_MAX_PLANES];        /* [K_i][256] += */
  float* donehot[ICF_MAX_PLANES];           /* [N][K_i] = */
  float* dcont[ICF_MAX_PLANES];             /* [N] = */
} icf_latfeat_args;
int icf_latent_features_fwd(const icf_latfeat_args* a, void* stream);
int icf_latent_features_bwd(const icf_latfeat_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BatchNorm2d in training mode (mnist.py:111,114,118,122) split around the convolutions:
 *  - the producing conv accumulates stats[2][C];
 *  - icf_bn_finalize turns them into scale/shift, saves mean/invstd, updates the running statistics
 *    (momentum 0.1, unbiased running_var, num_batches_tracked += 1), and re-zeroes `stats`;
 *  - icf_scale_shift_mask applies u = mask[n,c] * (scale[c]*y + shift[c]) (BatchNorm + Dropout2d),
 *    also used alone for a bare Dropout2d or a dtype cast.
 * ------------------------------------------------------------------------------------------------ */
int icf_bn_finalize(float* stats, int32_t C, double count, const float* gamma, const float* beta,
                    float eps, float momentum, float* running_mean, float* running_var,
                    int64_t* num_batches_tracked, float* scale, float* shift, float* save_mean,
                    float* save_invstd, void* stream);
int icf_scale_shift_mask(const void* y, int32_t y_dtype, int32_t y_pitch, void* u, int32_t u_dtype,
                         int32_t u_pitch, int64_t pixels, int32_t pixels_per_sample, int32_t C,
                         const float* scale, const float* shift, const float* mask, int32_t mask_pitch,
                         void* stream);
/* ------------------------------------------------------------------------------------------------
 * Spectrogram front end, the step in front of the hot path for the spectrogram families:
 *   icf_log_spectrogram  = torchaudio.transforms.Spectrogram(n_fft, win_length, pad) followed by (. + eps).log()
 *                          (audio_mnist.py:59-61,116: n_fft 255, win_length 128, pad 96, hop = win_length/2, eps 1e-6):
 *                          zero padding by `pad`, reflect centring by n_fft/2, periodic Hann window of win_length centred in
 *                          the n_fft frame, one-sided power spectrum.  wave fp32 [N][L] -> out fp32 [N][n_fft/2+1][frames],
 *                          frames = 1 + (L + 2*pad + 2*(n_fft/2) - n_fft)/hop.
 *   icf_spect_stats      sum[t] += sum_rows s[row][t], sumsq[t] += sum_rows s^2   (audio_mnist.py:347-358: statistics per time
 *                          frame over clips and frequency bins; the caller divides by the row count)
 *   icf_spect_to_img     clip((s - mean[t])/(std[t] + 1e-6), -k, k)/k             (audio_mnist.py:361-363), out f32 or bf16
 * ------------------------------------------------------------------------------------------------ */
int icf_log_spectrogram(const float* wave, int64_t N, int32_t L, int32_t n_fft, int32_t win_length, int32_t hop, int32_t pad,
                        float eps, float* out, int32_t frames, void* stream);
int icf_spect_stats(const float* s, int64_t rows, int32_t T, float* sum, float* sumsq, void* stream);
int icf_spect_to_img(const float* s, const float* mean, const float* std_, int64_t rows, int32_t T, float stds_kept, void* out,
                     int32_t out_dtype, void* stream);

/* "Taps as channels" for the layers with ONE (or very few) channels on one side and a stride-2 5x5 filter — the generator tail
 * ConvTranspose2d(C, 1, 5, 2, 2, 1) (audio_mnist.py:242, whalecalls.py:306, esrf_acoustic.py:197) and the data gradient of the
 * first conv towards the attribute-plane channels: the contraction over the C channels runs as a plain 1x1 icf_conv_forward /
 * icf_conv_wgrad GEMM whose other dimension are the filter taps, and these two kernels move between the tap-major matrix and
 * the image:
 *   icf_col2im_taps  out[n,oy,ox,k] = act(bias[k] + sum over taps (r,s) with (oy+pad-r, ox+pad-s) = stride*(iy,ix) of
 *                    T[n,iy,ix][k*TP + r*S+s])            T: bf16 [N*H*W][t_pitch], TP = tap count padded
 *   icf_im2col_taps  A[n,iy,ix][r*S+s] = src[n, iy*stride-pad+r, ix*stride-pad+s][0]  (zero outside / in the padding columns)
 *                    A: bf16 [N*H*W][a_pitch], a_pitch a multiple of 8 */
int icf_col2im_taps(const void* T, int32_t t_pitch, int32_t TP, int32_t N, int32_t H, int32_t W, int32_t P, int32_t Q, int32_t K,
                    int32_t R, int32_t S, int32_t stride, int32_t pad, const float* bias, int32_t act, float slope, void* out,
                    int32_t out_dtype, int32_t out_pitch, void* stream);
int icf_im2col_taps(const void* src, int32_t src_dtype, int32_t src_pitch, int32_t N, int32_t P, int32_t Q, int32_t H, int32_t W,
                    int32_t R, int32_t S, int32_t stride, int32_t pad, void* A, int32_t a_pitch, void* stream);
/* BatchNorm folded into the convolution that consumes its output when no Dropout2d sits in between and the convolution has no
 * padding (mnist.py:111-112: BatchNorm2d(32) -> Conv2d(32,64,4,2)): the normalised tensor is never materialised.
 *   forward   conv(scale*y + shift; w, b) = conv(y; w', b'),  w'[k][t][c] = w[k][t][c]*scale[c],  b'[k] = b[k] + sum_{t,c} w*shift[c]
 *             (w, w_out: packed operands [K][T][Cp]; scale/shift as written by icf_bn_finalize)
 *   backward  the data gradient uses the un-folded operand; the weight gradient G computed against y becomes
 *             dW[k][t][c] = G[k][t][c]*scale[c] + shift[c]*dbias[k]   (dw: packed fp32 accumulator [K][T][C], in place) */
int icf_bn_fold_weights(const void* w, int32_t dtype, int32_t K, int32_t T, int32_t Cp, int32_t C, const float* scale,
                        const float* shift, const float* bias, void* w_out, float* bias_out, void* stream);
int icf_bn_fold_wgrad(float* dw, int32_t K, int32_t T, int32_t C, const float* scale, const float* shift, const float* dbias,
                      void* stream);
/* sums[0][c] = sum du, sums[1][c] = sum du * xhat, du = dU*mask, xhat = (y-mean)*invstd  (+=) */
int icf_bn_bwd_reduce(const void* dU, int32_t d_dtype, int32_t d_pitch, const void* y, int32_t y_dtype,
                      int32_t y_pitch, int64_t pixels, int32_t pixels_per_sample, int32_t C,
                      const float* mask, int32_t mask_pitch, const float* save_mean,
                      const float* save_invstd, float* sums, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of the fused conv epilogue (aten::leaky_relu_backward / tanh_backward / dropout backward /
 * native_batch_norm_backward / bias gradient in one pass):
 *   g      = dOut[n,pix,c]                                   (gradient w.r.t. what the consumer read)
 *   if bn: g = gamma*invstd*( g*bn_mask - sums0/M - xhat*sums1/M ),  dgamma += sums1, dbeta += sums0 (once);
 *          M = pixels (or pixels*world under SyncBN, see bn_inv_world)
 *   dPre   = g * out_mask[n,c] * act'(y)                     (act' from the saved output y)
 *   dbias[c] += sum dPre
 * ------------------------------------------------------------------------------------------------ */
typedef struct icf_actbwd_args {
  int32_t d_dtype, d_pitch;      /* dOut */
  int32_t y_dtype, y_pitch;      /* saved layer output y */
  int32_t p_dtype, p_pitch;      /* dPre (written) */
  int64_t pixels;                /* N*P*Q */
  int32_t pixels_per_sample;
  int32_t C;
  int32_t act; float slope;
  int32_t mask_pitch, bn_mask_pitch;
  const void* dOut; const void* y; void* dPre;
  const float* out_mask;         /* Dropout2d mask applied in the forward epilogue, or NULL */
  float* dbias;                  /* [C] += , or NULL */
  int32_t bias_mod;              /* dbias index = c % bias_mod (0: = c) */
  /* BatchNorm that followed this layer's output (NULL bn_sums: none) */
  const float* bn_sums;          /* [2][C] from icf_bn_bwd_reduce */
  const float* bn_mask;          /* Dropout2d mask applied after the BatchNorm, or NULL */
  const float* bn_gamma; const float* bn_mean; const float* bn_invstd;
  float* bn_dgamma; float* bn_dbeta;   /* [C] += */
  float bn_inv_world;            /* SyncBN across `world` data-parallel ranks: bn_sums were summed over all ranks (the caller
                                    all-reduces them), so M = pixels*world and dgamma/dbeta += sums/world (the later gradient
                                    average re-multiplies by world/world).  0 = 1 (statistics of this tensor alone). */
} icf_actbwd_args;
int icf_act_backward(const icf_actbwd_args* a, void* stream);

/* ------------------------------------------------------------------------------------------------
 * nn.BCEWithLogitsLoss (mnist.py:181,228,234,239) forward + backward in one kernel:
 *   loss_out[0] += weight * mean_n( max(l,0) - l*t + log1p(exp(-|l|)) )
 *   dlogits[n]   = weight * (sigmoid(l) - t) / N
 * and the phase-D score (mnist.py:245-248): score_out[0] += mean_n sigmoid(l).
 * ------------------------------------------------------------------------------------------------ */
int icf_bce_logits(const void* logits, int32_t l_dtype, int32_t l_pitch, int32_t n, float target,
                   float weight, float* loss_out, void* dlogits, int32_t d_dtype, int32_t d_pitch,
                   void* stream);
int icf_sigmoid_mean(const void* logits, int32_t l_dtype, int32_t l_pitch, int32_t n, float* score_out,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fine-tune losses (finetune_mnist_bigan.py:68-86, finetune_audio_mnist_bigan.py:79-92, finetune_whale_bigan.py:58-73):
 *   rec    = torch.square(x - xr).mean()         -> icf_mse_loss: loss_out[0] += weight*(mean((xr - x)^2) + extra[0]),
 *                                                   dxr = weight * 2 (xr - x) / count  (gradient w.r.t. the reconstruction)
 *   latent = torch.square(codes).mean()          -> icf_latent_l2: loss_out[0] += weight*mean(z^2), dz (+)= weight*2z/count
 * x is fp32 [n_img][pixels_per_image] (target_stride = pixels_per_image) or ONE image broadcast over the batch
 * (target_stride = 0): finetune_whale_bigan.py:59-65 subtracts (N,256,256) from (N,1,256,256), i.e. the all-pairs mean,
 * which equals the MSE against the batch-mean image plus the mean pixel variance (`extra`, a device scalar or NULL);
 * icf_col_mean supplies both.
 * xr / dxr address channel 0 of [n_img*pixels][pitch] tensors.
 * ------------------------------------------------------------------------------------------------ */
int icf_mse_loss(const float* x, int64_t target_stride, const void* xr, int32_t xr_dtype, int32_t xr_pitch, int64_t n_img,
                 int64_t pixels_per_image, float weight, const float* extra, float* loss_out, void* dxr, int32_t d_dtype,
                 int32_t d_pitch, void* stream);
/* xbar[p] = mean_n x[n][p];  var_out[0] += mean_p(mean_n x[n][p]^2 - xbar[p]^2) */
int icf_col_mean(const float* x, int64_t n, int64_t p, float* xbar, float* var_out, void* stream);
int icf_latent_l2(const void* z, int32_t z_dtype, int32_t z_pitch, int64_t n, int32_t latent, float weight, float* loss_out,
                  float* dz /* fp32 [n][latent] or NULL */, int32_t accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attribute-SCM intervention on the device — the "intervene" step between E and G of a counterfactual
 * (attribute_scms/graph.py:144-184 sample_cf: abduct the noise of every observed variable, regenerate the
 * descendants of the intervened ones).  One conditional affine -> sigmoid -> affine mechanism per call, the shape
 * of MorphoMNIST's thickness -> intensity edge (attribute_scms/mnist.py:28-33,48):
 *   abduction      u = clamp((v - lo)/span, u_min, u_max);  s = log u - log1p(-u);  eps = (s - loc(p)) * exp(-ls(p))
 *   regeneration   s' = loc(p') + exp(ls(p')) * eps;         v' = lo + span * sigmoid(s')
 * (loc, ls)(p) = W2 relu(W1 p + b1) + b2 with `hidden` units (pyro ConditionalAutoRegressiveNN of a 1-D variable with a
 * 1-D context: the autoregressive mask removes the variable itself) or, hidden = 0, loc = closed[0] + closed[1]*p,
 * ls = closed[2] (the ground-truth SCM of create_train_dataset.py:42-46: loc = 2t - 5, scale 0.5); ls is clamped to
 * [clip_lo, clip_hi] (pyro: -5, 3).  p' = parent_cf[i], or parent[i] + parent_shift when parent_cf is NULL
 * (mnist_gan_counterfactuals.py:63 do(thickness + 2)).  *_scaled outputs = 2 (v - min)/(max - min) - 1 (:65-68).
 * ------------------------------------------------------------------------------------------------ */
typedef struct icf_scm_affine_args {
  int64_t n;
  int32_t hidden;                /* hyper-network width, 0 = closed form */
  float closed[3];               /* hidden == 0: loc = closed[0] + closed[1]*p, log scale = closed[2] */
  float clip_lo, clip_hi;        /* clamp of the log scale */
  float lo, span;                /* final affine transform of the mechanism */
  float u_min, u_max;            /* clamp of the sigmoid pre-image (torch SigmoidTransform: tiny, 1 - eps) */
  float parent_shift;
  float v_min, v_max, p_min, p_max;   /* min-max statistics for the scaled outputs */
  const float* w1; const float* b1;   /* [hidden] (context weight), [hidden] */
  const float* w2; const float* b2;   /* [2][hidden] (loc row, log-scale row), [2] */
  const float* value;            /* [n] observed child (intensity) */
  const float* parent;           /* [n] observed parent (thickness) */
  const float* parent_cf;        /* [n] counterfactual parent, or NULL */
  float* noise_out;              /* [n] abducted noise, or NULL */
  float* value_cf;               /* [n] raw counterfactual child, or NULL */
  float* parent_cf_out;          /* [n] raw counterfactual parent, or NULL */
  float* value_cf_scaled;        /* [n] or NULL */
  float* parent_cf_scaled;       /* [n] or NULL */
} icf_scm_affine_args;
int icf_scm_affine_cf(const icf_scm_affine_args* a, void* stream);
/* rows[n][:] = one_hot(idx[n]) for the rows with mask[n] != 0 (all rows when mask is NULL); other rows keep their content:
 * torch.eye(K)[idx] and the masked attribute swap of mnist_bigan_score.py:83-91 / audiomnist_cf_eval.py:82-83. */
int icf_onehot_swap(const void* idx, int32_t idx_is_int64, const uint8_t* mask, int64_t n, int32_t K, float* rows,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient-based counterfactual explainer (explain/cf_example.py:80-170 HingeLossCFExplainer.explain): per image the
 * optimised variables are RAW rows — one per attribute that is not ignored (:121-125) and, with train_z, the latent code
 * (:129-130); the generator is fed softmax(raw) for categorical attributes, tanh(raw) for the others and for z (:139-147).
 * All raw rows of a batch of images live in ONE flat fp32 buffer (group g = a [rows][width] block at `offset`), so that
 * the forward map, its backward and torch.optim.Adam (icf_adam_step over the flat buffer) are three launches per step.
 *   icf_explain_transform: out[g] = f_g(raw[g])
 *   icf_explain_backward : draw[g] = dout_g * f_g'(raw[g]) from the saved outputs (dout NULL -> 0)
 * ------------------------------------------------------------------------------------------------ */
#define ICF_EXPLAIN_MAX_GROUPS 12
#define ICF_EXPLAIN_COPY 0
#define ICF_EXPLAIN_TANH 1
#define ICF_EXPLAIN_SOFTMAX 2
typedef struct icf_explain_group {
  int32_t mode;                  /* ICF_EXPLAIN_* */
  int32_t width;                 /* elements per row */
  int64_t offset;                /* of the group's [rows][width] block in the flat buffers (elements) */
  const float* dout;             /* backward: gradient w.r.t. the transformed rows, [rows][width] fp32, or NULL */
} icf_explain_group;
int icf_explain_transform(const float* raw, float* out, const icf_explain_group* groups /* host */, int32_t n_groups,
                          int64_t rows, void* stream);
int icf_explain_backward(const float* out, float* draw, const icf_explain_group* groups /* host */, int32_t n_groups,
                         int64_t rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * torch.optim.Adam (mnist.py:176-179; eps 1e-8, no weight decay) over one flat fp32 buffer.
 * `state` = {step, lr, beta1, beta2, eps, grad_scale} in device memory so a captured CUDA graph can
 * be replayed: the kernel itself advances `step`.
 * ------------------------------------------------------------------------------------------------ */
int icf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float* state /*[8]*/, void* stream);

/* Workspace query of SURVEY.md §8(b): bytes of scratch memory the entry point named `entry_point` ("icf_conv_forward",
 * "icf_conv_wgrad", ...) would need for the given argument struct (may be NULL).  Every kernel of this library keeps its
 * staging buffers in shared / tensor memory and takes its accumulators (dw, stats, loss_out) as explicit arguments, so the
 * answer is 0 for every entry point; -1 for a name the library does not export.  Kept so that a binding written against
 * the survey's contract can size (and skip) its workspace allocation. */
int64_t icf_workspace_bytes(const char* entry_point, const void* args);

/* small utilities */
int icf_cast(const void* src, int32_t src_dtype, void* dst, int32_t dst_dtype, int64_t n, void* stream);
int icf_fill_f32(float* dst, float value, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ICF_H_ */
