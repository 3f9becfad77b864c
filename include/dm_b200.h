/*
 * dm_b200.h — C ABI of libdm_b200.so: the B200 (sm_100a) kernels behind the VAE / GAN / beta-VAE-GAN
 * training step of RicoFio/disentangle_mlp.
 *
 * The reference has no native layer: its hot path calls torch.nn library ops.  Each entry point below
 * replaces the ATen op(s) that a line of the reference invokes (cited as file:line, relative to the
 * reference checkout).  All pointers are DEVICE pointers borrowed for the duration of the call; every
 * function enqueues work on `stream` (a cudaStream_t passed as void*) and returns without synchronising.
 * Return value: 0 on success, non-zero on argument / launch error (see dm_last_error()).
 *
 * Data conventions
 *   - activations between layers: bf16, NHWC ("pixel-major": [batch, h, w, c]); matrices row-major
 *   - parameters / gradients / optimizer state: fp32 in the reference's own layouts
 *       Conv2d.weight [Co,Ci,5,5], ConvTranspose2d.weight [Ci,Co,5,5], Linear.weight [out,in]
 *   - "small"/"big": the low-/high-resolution side of a 5x5, pad-2 (transposed) convolution:
 *       Conv2d:          big = input,  small = output, weight [Cs=Co][Cb=Ci][5][5]
 *       ConvTranspose2d: small = input, big = output,  weight [Cs=Ci][Cb=Co][5][5]
 *     so one weight layout [Cs][Cb][25] serves both (models/model.py:449-457, 495-507, 388-398).
 */
#ifndef DM_B200_H_
#define DM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* dm_last_error(void);
int dm_version(void);
/* Number of kernels launched through this library since load (bench.py's gpu_launches). */
long long dm_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM-class ops (tcgen05 + TMEM + TMA).  bf16 operands, fp32 accumulation.
 * ---------------------------------------------------------------------------------------------- */

enum { DM_GEMM_NT = 0, /* D[m,n] = sum_k A[m,k] B[n,k]   (nn.Linear forward,  model.py:460-471,490,402-408) */
       DM_GEMM_NN = 1, /* D[m,n] = sum_k A[m,k] B[k,n]   (nn.Linear backward wrt input)                     */
       DM_GEMM_TN = 2  /* D[m,n] = sum_k A[k,m] B[k,n]   (nn.Linear backward wrt weight)                    */ };

/* BatchNorm statistics + finalize fused into the kernel that writes a pre-BatchNorm tensor (NT GEMM, dm_conv_down,
 * dm_conv_up, dm_conv_up_merged with bf16 output and no split-K; or dm_bn_stats for tensors no GEMM epilogue covers).
 * Producers add per-channel SHIFTED sums  sum (y - k), sum (y - k)^2  (k = running_mean; taken from the fp32 accumulator
 * + bias before the bf16 rounding in the GEMM case) to a slot scratch with fp32 atomic adds; the LAST CTA to finish
 * (ticket counter) sums the slots and writes scale_shift / mean_invstd for every stacked pass, updates the running
 * statistics in pass order (momentum, unbiased variance; num_batches_tracked += groups) and re-zeroes the scratch.
 * The consumer is then a plain dm_bn_apply_act(scale_shift).  scratch == NULL: off. */
typedef struct dm_bn_fuse {
  float* scratch;       /* dm_bn_scratch_floats(channels, groups) floats, ZERO on entry, left zero on exit */
  int groups;           /* passes stacked along the batch / rows; each must cover whole 128-row tiles in the GEMMs */
  long long rows;       /* rows (pixels) of ONE pass: the n of the statistics */
  const float* gamma;   /* [channels] */
  const float* beta;
  float* running_mean;  /* [channels] or NULL; also the shift k of the sums */
  float* running_var;
  long long* num_batches_tracked; /* or NULL */
  float momentum, eps;
  float* scale_shift;   /* out [groups][2][channels]: gamma*invstd, beta - mean*gamma*invstd */
  float* mean_invstd;   /* out [groups][2][channels]: saved for the backward pass */
} dm_bn_fuse;

typedef struct dm_gemm_desc {
  int layout;        /* DM_GEMM_* */
  int m, n, k;
  const void* a;     /* bf16 */
  long long lda;     /* leading dimension (elements) of A as stored */
  const void* b;     /* bf16 */
  long long ldb;
  void* d;           /* fp32 or bf16 */
  long long ldd_m;   /* element stride of D along m */
  long long ldd_n;   /* element stride of D along n (1 = row-major) */
  int d_f32;         /* 1: D is fp32, 0: bf16 */
  int accumulate;    /* 1: atomically add into D (fp32 only); required when splits > 1 */
  const float* bias; /* optional per-n bias (NT/NN only), added once */
  int m_store;       /* rows of D actually stored (0 = m) */
  int n_store;       /* columns of D actually stored (0 = n) */
  int splits;        /* split-K factor (>= 1) */
  int k_alg;         /* algorithmic K for FLOP accounting when k is zero-padded (0 = k) */
  dm_bn_fuse bn;     /* NT, bf16 D, splits == 1: fused BatchNorm statistics over the n_store columns (scratch NULL = off) */
} dm_gemm_desc;

int dm_gemm_bf16(const dm_gemm_desc* g, void* stream);

typedef struct dm_conv_geom {
  int batch;
  int hs, ws, cs; /* small side */
  int hb, wb, cb; /* big side; hb = hs*stride, wb = ws*stride */
  int stride;     /* 1 or 2 */
} dm_conv_geom;

/* small[b,hs,ws,cs] = sum_{kh,kw,cb} big[b, s*h+kh-2, s*w+kw-2, cb] * W[cs][cb][kh][kw]  (+ bias[cs])
 * = nn.Conv2d forward (model.py:449-457, 388-398) and nn.ConvTranspose2d input-gradient.
 * w_down: bf16 [25][cs][cb] (dm_pack_conv_weights).  Requires cb % 32 == 0, cs % 16 == 0. */
int dm_conv_down(const dm_conv_geom* g, const void* big, const void* w_down, const float* bias,
                 void* out_small, const dm_bn_fuse* bn, void* stream);

/* big[b, s*h+kh-2, s*w+kw-2, cb] += small[b,h,w,cs] * W[cs][cb][kh][kw]   (+ bias[cb])
 * = nn.ConvTranspose2d forward with output_padding = stride-1 (model.py:495-507, 555-563) and
 *   nn.Conv2d input-gradient.  w_up: bf16 [25][cb_pad][cs], cb_pad = max(cb,16) rounded up to 16.
 * out_big: bf16 NHWC [b,hb,wb,cb] or, when out_f32, fp32 NHWC (used for the 3-channel image side). */
int dm_conv_up(const dm_conv_geom* g, const void* small, const void* w_up, const float* bias,
               void* out_big, int out_f32, const dm_bn_fuse* bn, void* stream);

/* Phase-merged form of dm_conv_up for stride 2, cb == 32 (ConvTranspose2d(128, 32), model.py:500 / the input
 * gradient of Conv2d(32, 128), model.py:391): w_upm = dm_pack_up_merged(w_up) is [9][4*cb][cs]; one GEMM with N = 128
 * computes the four sub-pixel phases from 9 shared input taps.  Output bf16 NHWC like dm_conv_up. */
int dm_pack_up_merged(const void* w_up, int cs, int cb, void* w_upm, void* stream);

/* The stride-2 convolution with 32 input channels (Conv2d(32, 128), model.py:391; the input-gradient of
 * ConvTranspose2d(128, 32), model.py:500) with TWO filter columns per k-block: w_pair = dm_pack_down_pairs(w_down) is
 * [15][cs][64] (filter row kh, column pair j: columns 2j, 2j+1; the sixth column is zero); the A operand row is one
 * 128-byte run of the NHWC tensor (two adjacent pixels x 32 channels).  Same arguments as dm_conv_down otherwise. */
int dm_pack_down_pairs(const void* w_down, int cs, int cb, void* w_pair, void* stream);
int dm_conv_down_paired(const dm_conv_geom* g, const void* big, const void* w_pair, const float* bias, void* out_small,
                        const dm_bn_fuse* bn, void* stream);
int dm_conv_up_merged(const dm_conv_geom* g, const void* small, const void* w_upm, const float* bias, void* out_big,
                      const dm_bn_fuse* bn, void* stream);

/* Weight gradient of nn.Conv2d (small = grad_output, big = input) and of nn.ConvTranspose2d (small = input,
 * big = grad_output), fp32, accumulated atomically:
 *   direct_layout = 1: dw[cs][cb][kh][kw]          += sum_{b,h,w} small[b,h,w,cs] * big[b, s*h+kh-2, s*w+kw-2, cb]
 *   direct_layout = 0: dw_packed[kh*5+kw][cs][cb]  += (same sum)   tap-major layout, reduced with bulk tensor reductions
 *                      (cp.reduce.async.bulk.tensor, fp32 add) from a transposed smem staging tile.  The fused trainers
 *                      keep conv weights, gradients and Adam state in this layout (no unpack); dm_unpack_conv_grad
 *                      moves a packed scratch into the parameter layout for the drop-in modules.
 * Requires cs % 64 == 0 and cb % 64 == 0, or cb == 32 with stride 2. */
int dm_conv_wgrad(const dm_conv_geom* g, const void* small, const void* big, float* dw, int direct_layout,
                  void* stream);
/* dw[cs][cb][5][5] (+)= dw_packed[25][cs][cb]; dw_packed is zeroed for the next accumulation. */
int dm_unpack_conv_grad(float* dw_packed, int cs, int cb, int accumulate, float* dw, void* stream);

/* ---- TF32 precision mode (north-star: "bf16/TF32 with fp32 accumulation", tolerance 1e-3 per layer).
 * The same tcgen05 kernel instantiated for tcgen05.mma.kind::tf32: operands are FP32 in memory (activations fp32 NHWC,
 * weights fp32 in the same packed layouts: w_down [25][cs][cb], w_up [25][cb][cs], Linear [out][in]), read by the tensor
 * cores as TF32 (10-bit mantissa), accumulated in fp32 in TMEM, written as fp32.  The reference's layers are fp32
 * (model.py:388-408, 449-509); this is the mode that tracks them to 1e-3.  Requires channel counts % 32 == 0 (the three
 * 3-channel layers stay on the bf16 path); dw_packed of dm_conv_wgrad_tf32 is the tap-major [25][cs][cb] layout. */
int dm_gemm_tf32(const dm_gemm_desc* g, void* stream);
int dm_conv_down_tf32(const dm_conv_geom* g, const void* big, const void* w_down, const float* bias, void* out_small,
                      void* stream);
int dm_conv_up_tf32(const dm_conv_geom* g, const void* small, const void* w_up, const float* bias, void* out_big,
                    void* stream);
int dm_conv_wgrad_tf32(const dm_conv_geom* g, const void* small, const void* big, float* dw_packed, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound ops.  "rows x c" = channel-innermost matrix view of an NHWC activation (rows = b*h*w) or of a
 * Linear output (rows = batch).  act: 0 none, 1 ReLU, 2 LeakyReLU(slope).
 * ---------------------------------------------------------------------------------------------- */

/* nn.BatchNorm1d/2d, training mode (model.py:451,454,457,462,468,492,496,500,504,390,393,396,399), with the
 * ReLU / LeakyReLU(0.2) that follows every BatchNorm folded in.  Two launches per application and direction, no
 * finalize kernel:
 *   producer : per-channel partial sums into a small SLOT scratch [groups][dm_bn_slots()][2][c] (fp32 atomic adds), then
 *              the last block finalizes (see dm_bn_fuse above).  Forward: the GEMM that writes the tensor (dm_bn_fuse
 *              argument of the GEMM-class entry points) or dm_bn_stats; backward: the first kernel of dm_bn_backward
 *              (sum dz, sum dz*xhat, dz = dout*act'(z); dgamma / dbeta accumulated by its last block).
 *   consumer : dm_bn_apply_act (out = act(y*scale + shift)) / the second kernel of dm_bn_backward
 *              (dy = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat))).
 * `scratch`: dm_bn_scratch_floats(c, groups) floats per call site, ZERO on entry and left zero on exit (allocate and
 * clear once; never memset in the step).  scale_shift / mean_invstd: [groups][2][c].
 * groups > 1: y / out (dout / dy) hold `groups` independent batches of `rows` rows stacked along rows (several forward
 * passes of one network pushed through each GEMM together); statistics, normalisation and running-stat updates are per
 * group, in order; dgamma / dbeta accumulate over all groups.
 * Tensors with rows <= 256 (BatchNorm1d behind the Linear layers: rows = batch) take ONE launch with an exact two-pass
 * variance inside dm_bn_forward / dm_bn_backward (scratch unused). */
int dm_bn_slots(void);
long long dm_bn_scratch_floats(int c, int groups);
/* producer + finalize for a tensor already in memory; f->rows = rows of one pass */
int dm_bn_stats(const void* y, int y_f32, int c, const dm_bn_fuse* f, void* stream);
/* out = act(y*scale + shift) with given constants ([groups][2][c]); `rows` per group */
int dm_bn_apply_act(const void* y, int y_f32, long long rows, int c, const float* scale_shift, int act,
                    float slope, void* out_bf16, int groups, void* stream);
/* = dm_bn_stats + dm_bn_apply_act, or the single-launch small-row kernel */
int dm_bn_forward(const void* y, int y_f32, long long rows, int c, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, long long* num_batches_tracked, float momentum, float eps,
                  int act, float slope, float* scratch, float* scale_shift, float* mean_invstd, void* out_bf16,
                  int groups, void* stream);
/* nn.BatchNorm1d (training) + activation on a COLUMN BLOCK of a wider fp32 matrix (row stride ld_y), rows <= 256: the
 * two heads x_to_mu / x_to_logvar (model.py:460-471) computed by ONE GEMM over their stacked weights [4096][16384].
 * pre_bias: the head's Linear bias, which that GEMM does not add (it only moves running_mean).  dy of the backward
 * form goes to a column block of a [rows][ld_dy] matrix (the A operand of the fused input-gradient GEMM). */
int dm_bn1d_forward(const float* y, long long ld_y, int rows, int c, const float* pre_bias, const float* gamma,
                    const float* beta, float* running_mean, float* running_var, long long* num_batches_tracked,
                    float momentum, float eps, int act, float slope, float* scale_shift, float* mean_invstd,
                    void* out_bf16, void* stream);
int dm_bn1d_backward(const void* dout_bf16, const float* y, long long ld_y, int rows, int c, const float* scale_shift,
                     const float* mean_invstd, int act, float slope, void* dy_bf16, long long ld_dy, float* dgamma,
                     float* dbeta, void* stream);
int dm_bn_backward(const void* dout_bf16, const void* y, int y_f32, long long rows, int c,
                   const float* scale_shift, const float* mean_invstd, int act, float slope, float* scratch,
                   void* dy_bf16, float* dgamma, float* dbeta, int groups, void* stream);
/* rows of the [parts][c] partial-sum scratch of dm_act_backward / dm_colsum */
int dm_bn_parts(long long rows, int c);

/* out = act(acc + bias) after a split-K Linear (model.py:402-404); fp32 and/or bf16 outputs (NULL = skip). */
int dm_bias_act(const float* acc, long long rows, int c, const float* bias, int act, float slope,
                float* out_f32, void* out_bf16, void* stream);
/* dpre = dout * act'(out) as bf16; colsum[c] += sum_r dpre (the Linear bias gradient; NULL = skip). */
int dm_act_backward(const float* dout, const float* out, long long rows, int c, int act, float slope,
                    void* dpre_bf16, float* partials, float* colsum, void* stream);
/* colsum[c] += sum_r x[r][c] */
int dm_colsum(const void* x, int x_f32, long long rows, int c, float* partials, float* colsum, void* stream);

/* fp32 NCHW [b,3,h,w] image -> bf16 im2col matrix [b*(h/s)*(w/s), 80], column c*25+kh*5+kw (75 valid):
 * the A operand of the two 3-channel convolutions (model.py:449, 389) and of deconv4's input-gradient. */
int dm_im2col3(const float* x_nchw, int batch, int h, int w, int stride, void* col_bf16, void* stream);
/* fp32 NHWC(3) -> fp32 NCHW, optionally through nn.Tanh (model.py:509,565); pim_bf16 (may be NULL; 64x64 only) also
 * receives the result as a PADDED IMAGE (below): the discriminator's input operand. */
int dm_nhwc3_to_nchw(const float* src, long long batch, int hw, int apply_tanh, float* dst, void* pim_bf16, void* stream);
/* dy = dout*(1-out^2) (fp32 NCHW in); bias_grad[3] += per-channel sums of dy (deconv4.bias gradient).  dy is written as
 * fp32 NCHW (dy, may be NULL) and / or as a padded image (pim_bf16, may be NULL): the operand of deconv4's gradient GEMMs. */
int dm_tanh_backward(const float* dout, const float* out, long long batch, int hw, float* dy,
                     float* bias_grad, void* pim_bf16, void* stream);

/* ---- the 3-channel image side as TMA-fed implicit GEMMs over a PADDED IMAGE ("pim"):
 * bf16 [batch][68][72][4], pixel (h, w) of a 64x64 image at [h+2][w+2]; the 2-pixel border, 4 spare pixels per row and
 * the 4th channel are zero (dm_pim_elems(batch) elements incl. 64 of slack, 128-byte aligned).  A 5x5 filter ROW of an
 * output pixel is then 20 contiguous elements; the GEMMs read 8-pixel (16-pixel: weight gradient) windows starting at
 * even pixels over a tensor map with overlapping positions (TMA strides are multiples of 16 B = 2 pixels): one window
 * per output pixel for stride 2, one per output-pixel PAIR for stride 1 (the pair = 2*cs GEMM columns).  No im2col matrix.
 *   dm_pad_image3         image -> pim.  src fp32 NCHW [b,3,64,64] (src_u8 = 0), or uint8 NHWC [b,64,64,3] (src_u8 = 1):
 *                         the reference's input transform ToTensor() + Normalize(.5,.5) = (u/255 - .5)/.5
 *                         (dataloader/dataset.py:37-43) fused in; dst_nchw (may be NULL) then also gets the normalised
 *                         fp32 NCHW image (what the losses read).
 *   dm_pack_conv3_weights fp32 W[cs][3][5][5] -> bf16 w_win: stride 2 [5][cs][32] (element kw*4+c of filter row kh),
 *                         stride 1 [5][2*cs][32] (row pw*cs+n holds the filter shifted by pw pixels)
 *   dm_conv3_fwd          out[b,hs,ws,cs] = conv5x5(image, W) + bias, stride 1 / 2: nn.Conv2d(3, cs) forward
 *                         (model.py:449, 389) and the input-gradient of nn.ConvTranspose2d(cs, 3) (model.py:507);
 *                         optional fused BatchNorm statistics (dm_bn_fuse)
 *   dm_conv3_wgrad        dw_win[5][cs or 2*cs][64] (fp32) += sum_positions small x window: weight gradient of those
 *                         layers (small = grad_output of the Conv2d / input of deconv4); split-K, bulk tensor reductions
 *   dm_unpack_conv3_grad  dw[cs][3][5][5] += dw_win; dw_win is re-zeroed */
long long dm_pim_elems(int batch);
int dm_pad_image3(const void* src, int src_u8, int batch, void* pim_bf16, float* dst_nchw, void* stream);
int dm_pack_conv3_weights(const float* w, int cs, int stride, void* w_win, void* stream);
int dm_conv3_fwd(const dm_conv_geom* g, const void* pim, const void* w_win, const float* bias, void* out_small,
                 const dm_bn_fuse* bn, void* stream);
int dm_conv3_wgrad(const dm_conv_geom* g, const void* pim, const void* small, float* dw_win, void* stream);
int dm_unpack_conv3_grad(float* dw_win, int cs, int stride, float* dw, void* stream);
/* bf16 [batch][rows][cols] -> [batch][cols][rows]: NHWC <-> the NCHW flatten order that the 16384-wide
 * Linear layers are defined on (model.py:516-517, 540-543, 412-413). */
int dm_transpose_bf16(const void* src, int batch, int rows, int cols, void* dst, void* stream);
/* fp32 weight [cs][cb][5][5] -> bf16 operand packs: w_down [25][cs][cb], w_up [25][cb_pad][cs],
 * w_col [cs][128] (only when cb*25 <= 128).  Any output may be NULL.  For cb == 3 the first 5 planes of w_up hold the
 * kw-folded layout [kh][kw*3+cb][cs] that dm_conv_up uses for the 3-channel image side. */
int dm_pack_conv_weights(const float* w, int cs, int cb, void* w_down, void* w_up, void* w_col, void* stream);
int dm_cast_bf16(const float* src, long long n, void* dst, void* stream);

/* z = mu + eps*exp(0.5*logvar) (model.py:532-535) and its backward (+ externally supplied dmu/dlogvar). */
int dm_reparam_forward(const float* mu, const float* logvar, const float* eps, long long n, float* z_f32,
                       void* z_bf16, void* stream);
int dm_reparam_backward(const float* dz, const float* logvar, const float* eps, const float* dmu_ext,
                        const float* dlogvar_ext, long long n, void* dmu_bf16, void* dlogvar_bf16,
                        float* dmu_f32, float* dlogvar_f32, void* stream);

/* The second Linear of the encoder's two heads (Linear(2048, 128) of x_to_mu and x_to_logvar, model.py:464,470) for BOTH
 * heads in one launch (SIMT: 17 MFLOP each).  x bf16 [rows][k], w bf16 [n][k], out / d fp32 [rows][n].
 * backward: dx bf16 [rows][k] = d w;  dw fp32 [n][k] += d^T x and db fp32 [n] += column sums of d (skipped when dw0 or
 * dw1 is NULL).  k % 256 == 0; backward: n <= 128, n % 16 == 0, rows <= 256. */
int dm_linear_pair_forward(const void* x0, const void* x1, const void* w0, const void* w1, const float* b0, const float* b1,
                           int rows, int n, int k, float* out0, float* out1, void* stream);
int dm_linear_pair_backward(const float* d0, const float* d1, const void* x0, const void* x1, const void* w0,
                            const void* w1, int rows, int n, int k, void* dx0, void* dx1, float* dw0, float* dw1,
                            float* db0, float* db1, void* stream);

/* Discriminator head Linear(k,1)+Sigmoid (model.py:406-408) and its backward. */
int dm_head_forward(const float* feat, int rows, int k, const float* w, const float* b, float* prob,
                    void* stream);
int dm_head_backward(const float* dprob, const float* prob, const float* feat, const float* dfeat_ext,
                     int rows, int k, const float* w, float* dfeat, float* dw, float* db, void* stream);

/* Loss reductions (experiments/new_betavaegan.py:53,64-75; new_vae.py:39-48).  `loss` is a device scalar
 * that is atomically accumulated; gradients are written (or accumulated) when the pointer is non-NULL. */
int dm_mse_sum(const float* a, const float* b, long long n, float wloss, float* loss, float wgrad,
               int grad_accumulate, float* grad, void* stream);
int dm_kl(const float* mu, const float* logvar, long long n, float w, float* loss, int grad_accumulate,
          float* dmu, float* dlogvar, void* stream);
/* target_dev (may be NULL): read the label from device memory instead of `target` (CUDA-graph replay). */
int dm_bce_const(const float* p, int n, float n_total, float target, const float* target_dev, float w, float* loss,
                 int grad_accumulate, float* dprob, float* stat, void* stream);

/* torch.optim.Adam step on a flat fp32 buffer (new_betavaegan.py:49-50,123,164,193); `step` is the
 * 1-based step count; g is multiplied by grad_scale first; shadow_bf16 (may be NULL) receives bf16(p).
 * step_dev (may be NULL): device-side step counter, incremented and then used instead of `step` (CUDA-graph
 * replay: the bias corrections change every step). */
int dm_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1,
                 double beta2, double eps, int step, int* step_dev, float grad_scale, void* shadow_bf16,
                 void* stream);
/* Same on one SEGMENT of a flat buffer: g may be bf16 (g_bf16 != 0: the three 16384x2048 Linear weight gradients are
 * stored, all-reduced and read in bf16); count_step == 0 reuses the device step counter that an earlier segment of
 * the same optimizer step has already incremented. */
int dm_adam_step_ex(float* p, const void* g, int g_bf16, float* m, float* v, long long n, double lr, double beta1,
                    double beta2, double eps, int step, int* step_dev, int count_step, float grad_scale,
                    void* shadow_bf16, void* stream);

/* Workspace query: bytes of caller-owned scratch an op needs; all tensors -- inputs, outputs, saved-for-
 * backward, scratch -- are allocated by the caller (PyTorch's caching allocator) and borrowed for the call; the library
 * allocates nothing except cached tensor-map descriptors.  dims per op:
 *   DM_WS_GEMM {m,n,k,splits} = 0 (split-K reduces into D); DM_WS_CONV_FWD / _DGRAD {..} = 0 (no im2col buffer);
 *   DM_WS_CONV_WGRAD {cs,cb} packed-gradient scratch when dw is in the parameter layout; DM_WS_CONV3_WGRAD {cs,stride};
 *   DM_WS_BATCHNORM {c,groups} slot scratch (= 4 * dm_bn_scratch_floats); DM_WS_PADDED_IMAGE {batch} (= 2 * dm_pim_elems);
 *   DM_WS_COLSUM {rows,c} partial sums of dm_act_backward / dm_colsum.  Returns -1 for an unknown op. */
enum { DM_WS_GEMM = 0, DM_WS_CONV_FWD = 1, DM_WS_CONV_DGRAD = 2, DM_WS_CONV_WGRAD = 3, DM_WS_CONV3_WGRAD = 4,
       DM_WS_BATCHNORM = 5, DM_WS_PADDED_IMAGE = 6, DM_WS_COLSUM = 7 };
long long dm_workspace_bytes(int op, const long long* dims, int ndims);

/* Same on a segment whose update runs OUT OF ORDER with the rest of its optimizer step, off the critical path:
 *   EARLY    (step_offset = 1, enable_dev NULL): the two 33.5 M-element encoder Linear weight gradients are final as soon
 *            as the heads have been back-propagated; their Adam update (2/3 of the step's Adam traffic) runs on a side
 *            stream under the encoder's convolution backward, before the rest of the step has incremented the device
 *            step counter -- hence the offset;
 *   DEFERRED (step_offset = 0, enable_dev = device int "the gradient buffer holds an unapplied gradient"): applied at
 *            the start of the next step; a no-op when *enable_dev == 0.
 * Never increments the step counter. */
int dm_adam_step_gated(float* p, const void* g, int g_bf16, float* m, float* v, long long n, double lr, double beta1,
                       double beta2, double eps, int* step_dev, int step_offset, float grad_scale, void* shadow_bf16,
                       const int* enable_dev, void* stream);

/* Per-launch CUDA-event timing of the GEMM-class kernel (bench.py roofline). dm_profile_read synchronises the
 * device and returns the summed launch durations, algorithmic FLOPs (2*M*N*K; convolutions: 2*25*b*hs*ws*cs*cb)
 * and launch count since the previous read. */
int dm_profile_enable(int on);
int dm_profile_read(double* total_ms, double* total_flops, long long* launches);
/* One CSV row per profiled launch (tile grid, tile shape, us, GFLOP); clears the records; synchronises. */
int dm_profile_dump(const char* path);

/* Test hook: direct access to the tile plan of a GEMM-class call (tile counts, smem bytes). */
int dm_debug_last_plan(int* grid_xyz, int* smem_bytes, int* stages);

#ifdef __cplusplus
}
#endif
#endif /* DM_B200_H_ */
