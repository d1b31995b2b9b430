/*
 * jck_b200.h -- C ABI of libjck_b200.so: the sm_100a kernels under the DCGAN / CGAN train step.
 *
 * The reference (hy-vision-learning/jck-generation) has no native code and no FFI: its hot path is
 * Python calling torch operators.  Each entry point below therefore names the reference call site
 * (file:line, relative to the reference root) whose torch operator it replaces.  The boundary is
 * plain C: raw device pointers, int / float scalars, a cudaStream_t passed as void*.  No torch
 * types.  The Python host (jck_generation_b200/_lib.py) binds it with ctypes; INTEGRATION.md shows
 * the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative JCK_E_* code; the message is available from
 *     jck_last_error_string() (thread-local).  Nothing throws, exits, or falls back to a CPU path.
 *   - all work is asynchronous on `stream`; no entry point synchronises, allocates device memory,
 *     or reads device data on the host, so a whole train step is CUDA-graph capturable.
 *   - caller owns every buffer, workspaces included (sizes from the *_workspace_bytes queries).
 *   - activations are NHWC ("pixels x channels"), dtype JCK_F32 or JCK_BF16; images at the API edge
 *     are NCHW fp32 as in the reference; parameters / gradients / optimizer state are fp32 in the
 *     reference's own layouts ([Cout,Cin,4,4] for Conv2d, [Cin,Cout,4,4] for ConvTranspose2d).
 *   - every 4x4 stride-2 pad-1 layer is described by (Ca, Cb, Hs, Ws): Ca = channels of its
 *     low-resolution side ("small", Hs x Ws), Cb = channels of its high-resolution side ("large",
 *     2Hs x 2Ws).  Conv2d: Ca = Cout, Cb = Cin.  ConvTranspose2d: Ca = Cin, Cb = Cout.  Both store
 *     their weight as w4[Ca][Cb][4][4], so one packing serves both.
 *         down : large -> small   Conv2d forward            / ConvTranspose2d input-gradient
 *         up   : small -> large   ConvTranspose2d forward   / Conv2d input-gradient
 *         wgrad: (small, large) -> dw4                        weight gradient of either
 */
#ifndef JCK_B200_H_
#define JCK_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JCK_OK 0
#define JCK_E_BADARG (-1)
#define JCK_E_UNSUPPORTED_SHAPE (-2)
#define JCK_E_CUDA (-3)
#define JCK_E_DRIVER (-4)

/* activation dtypes */
#define JCK_F32 0
#define JCK_BF16 1

/* activation-side layout of IMAGE tensors (the nc-channel 64x64 side of D.conv1 / G.conv5):
 *   JCK_IMG_NHWC  dense [B][H][W][C]
 *   JCK_IMG_P4    [B][H+2][W+2][4], zero border and zero pad channels (C <= 4): every 4x4 stride-2
 *                 patch is then 4 runs of 32 contiguous bytes, which one TMA box fetches as a 128-byte
 *                 GEMM row -- the layout of the tcgen05 image-edge kernels (jck_edge_*). */
#define JCK_IMG_NHWC 0
#define JCK_IMG_P4 1

/* peer-memory communicator (jck_comm_*) */
#define JCK_COMM_HANDLE_BYTES 64 /* sizeof(cudaIpcMemHandle_t) */
#define JCK_COMM_MAX_N 3072      /* floats per small all-reduce: 3 groups x 2C, C <= 512 */

/* conv algorithm selector */
#define JCK_ALGO_AUTO 0  /* tcgen05 where dtype/shape allow, else SIMT */
#define JCK_ALGO_SIMT 1  /* CUDA-core fp32-FMA implicit GEMM (exact-fp32 parity mode; edge layers) */
#define JCK_ALGO_TC 2    /* tcgen05 / TMEM / TMA implicit GEMM; error if unsupported */

/* Traversal order of the streaming BatchNorm passes (a scheduling hint, never a change of the result).  A tensor that
 * does not fit the 126 MB L2 is walked by all thread blocks together as ONE front, ascending or descending, so that a
 * pass which starts where its producer stopped finds the producer's last ~40 MB still in L2 and leaves its own tail
 * where a consumer walking the other way begins.  Convolutions write ascending; the step runs bn_act_fwd and
 * bn_act_bwd_reduce descending, bn_act_bwd_apply ascending after a reduce pass / descending after a fused convolution.
 * JCK_ORDER_SLAB keeps the per-block contiguous slabs used for small tensors. */
#define JCK_ORDER_ASC 0
#define JCK_ORDER_DESC 1
#define JCK_ORDER_SLAB 2

int jck_version(void);
const char* jck_last_error_string(void);
/* number of kernels this library has launched in the calling process (bench.py "gpu_launches") */
unsigned long long jck_launch_count(void);

/* ---- image edge ------------------------------------------------------------------------------
 * out_nhwc[n,h,w,c] = v,  v = a1*x1 + b1*m1            (m1 may be NULL)
 *                       v = alpha[n]*v + (1-alpha[n])*x2 (when alpha != NULL)
 * x1, m1, x2: NCHW fp32.  Optional out_nchw_f32 receives v as NCHW fp32.
 * Replaces: instance noise `0.9*real + 0.1*randn` train/dcgan_trainer.py:160,171 and the
 * interpolation `alpha*real + (1-alpha)*fake` train/dcgan_trainer.py:112. */
int jck_prep_image(const float* x1, const float* m1, float a1, float b1, const float* x2,
                   const float* alpha, void* out_nhwc, float* out_nchw_f32, int B, int C, int H, int W,
                   int layout, int dtype, void* stream);

/* The same with the Gaussian m1 drawn in registers: m1[i] = what jck_randn(seed, stream_id, counter_base) writes at
 * flat NCHW index i, never stored (needs W % 4 == 0, C <= 4).  torch.randn_like train/dcgan_trainer.py:160. */
int jck_prep_image_rng(const float* x1, unsigned long long seed, unsigned long long stream_id,
                       const unsigned long long* counter_base, float a1, float b1, void* out_nhwc,
                       float* out_nchw_f32, int B, int C, int H, int W, int layout, int dtype, void* stream);

/* NHWC activation-dtype tensor -> NCHW fp32 (e.g. the GP input-gradient handed back to the caller). */
int jck_nhwc_to_nchw_f32(const void* in_nhwc, float* out_nchw, int B, int C, int H, int W, int layout,
                         int dtype, void* stream);

/* ---- weights ---------------------------------------------------------------------------------
 * w4[Ca][Cb][4][4] fp32 -> w_down[Ca][16][Cb] and w_up[4 phases][Cb][4 taps][Ca] (dtype).
 * Either output may be NULL. */
int jck_pack_weights(const float* w4, void* w_down, void* w_up, int Ca, int Cb, int dtype, void* stream);

/* ---- 4x4 stride-2 convolutions ---------------------------------------------------------------
 * stats (nullable): fp32 [groups][2*Cout] receiving per-channel sum and sum of squares of the
 * fp32 accumulators, ADDED atomically (caller zeroes); image n belongs to group n / imgs_per_group.
 * Replaces: nn.Conv2d forward model/DCGAN.py:30-33, nn.ConvTranspose2d forward model/DCGAN.py:63-66,
 * and aten::convolution_backward (autograd of the above; train/dcgan_trainer.py:164,175,187,116). */
int jck_conv_down(const void* in_large, const void* w_down, void* out_small, float* stats, int B, int Hs,
                  int Ws, int Ca, int Cb, int imgs_per_group, int dtype, int algo, void* stream);
int jck_conv_up(const void* in_small, const void* w_up, void* out_large, float* stats, int B, int Hs,
                int Ws, int Ca, int Cb, int imgs_per_group, int dtype, int algo, void* stream);
/* Input-gradient convolution with the BatchNorm-backward REDUCTION of the layer below fused into its epilogue
 * (bf16 / tcgen05 only; returns JCK_E_UNSUPPORTED_SHAPE otherwise -- callers then use jck_conv_up / jck_conv_down
 * followed by jck_bn_act_bwd_reduce).  With y_saved = that layer's raw conv output (same shape as the result) and
 * its statistics, the epilogue forms, from the fp32 accumulators,
 *     g = acc * act'(y*scale + shift),   sums[group][0:C] += sum g,   sums[group][C:2C] += sum g * xhat
 * and stores g (out_g), i.e. what jck_bn_act_bwd_apply consumes with slope = 1.
 * Replaces: autograd of nn.Conv2d / nn.ConvTranspose2d followed by native_batch_norm_backward's reduction
 * (model/DCGAN.py:30-33 and :62-65 under loss.backward(), train/dcgan_trainer.py:164,175,187). */
int jck_conv_up_bnbwd(const void* in_small, const void* w_up, const void* y_saved, const float* scale_shift,
                      const float* mean_rstd, float slope, void* out_g, float* sums, int B, int Hs, int Ws, int Ca,
                      int Cb, int imgs_per_group, int dtype, void* stream);
int jck_conv_down_bnbwd(const void* in_large, const void* w_down, const void* y_saved, const float* scale_shift,
                        const float* mean_rstd, float slope, void* out_g, float* sums, int B, int Hs, int Ws, int Ca,
                        int Cb, int imgs_per_group, int dtype, void* stream);

/* dw4[Ca][Cb][4][4] (+)= sum over pixels small (x) large.  workspace: jck_conv_wgrad_workspace_bytes. */
size_t jck_conv_wgrad_workspace_bytes(int B, int Hs, int Ws, int Ca, int Cb, int dtype, int algo);
int jck_conv_wgrad(const void* small, const void* large, float* dw4, void* workspace, size_t workspace_bytes,
                   int B, int Hs, int Ws, int Ca, int Cb, int accumulate, int dtype, int algo, void* stream);

/* ---- image-edge layers on tensor cores (bf16, nc <= 4 image channels, JCK_IMG_P4 image layout) -----
 * D.conv1 (model/DCGAN.py:10,30) and G.conv5 (model/DCGAN.py:58,66) have 3 channels on their large side:
 * K = 48 or N = 3 in the GEMM view.  With the image stored as JCK_IMG_P4 the whole 4x4x4 patch of an output
 * pixel is one 64-wide K step (down / wgrad): the kernels bulk-copy the 2*rows+2 image rows a tile needs into
 * shared memory and re-pack them there into the swizzled MMA operand -- no im2col / patch matrix ever exists in
 * HBM.  The transposed direction is a 3x3-shift GEMM with N = 16 (4 output parities x 4 channels).
 * w_down_e[Ca][64], w_up9[16][9*Ca] from jck_pack_weights_edge. */
int jck_pack_weights_edge(const float* w4, void* w_down_e, void* w_up9, int Ca, int nc, void* stream);
/* D.conv1 forward (nn.Conv2d(nc,64,4,2,1) model/DCGAN.py:10,30) / G.conv5 input-gradient (:58 under backward) */
int jck_edge_down_img(const void* img_p4, const void* w_down_e, void* out_small, float* stats, int B, int Hs, int Ws, int Ca,
                      int imgs_per_group, void* stream);
/* G.conv5 forward (nn.ConvTranspose2d(64,nc,4,2,1) model/DCGAN.py:58,66) / D.conv1 input-gradient (GP sweep, pass D) */
int jck_edge_up(const void* in_small, const void* w_up9, void* img_p4, int B, int Hs, int Ws, int Ca, void* stream);
/* The same product in scatter form (the step's path for 32-pixel rows): every 4-row activation tile is read once and
 * multiplied by the whole filter bank (w_down_e of jck_pack_weights_edge read as an MN-major operand), the 16 tap images
 * are folded in the epilogue; tiles overlap by one input row so every output pixel is written exactly once. */
int jck_edge_up_scatter(const void* in_small, const void* w_down_e, void* img_p4, int B, int Hs, int Ws, int Ca,
                        void* stream);
/* weight gradient of either (small = the 64-channel side, img_p4 = the image side) */
size_t jck_edge_wgrad_workspace_bytes(int B, int Hs, int Ws, int Ca);
int jck_edge_wgrad_img(const void* small, const void* img_p4, float* dw4, void* workspace, size_t workspace_bytes, int B,
                       int Hs, int Ws, int Ca, int nc, int accumulate, void* stream);

/* ---- dense layers (G.conv1: a 1x1 -> 4x4 transposed conv is a matrix product) -----------------
 * out[m][n] = sum_k x[m][k] * w[n][k];  x fp32 [M][K]; w, out activation dtype; stats per channel
 * (n % C).  Replaces nn.ConvTranspose2d(100,512,4,1,0) model/DCGAN.py:42,62. */
int jck_fc_fwd(const float* x, const void* w, void* out, float* stats, int M, int N, int K, int C,
               int dtype, void* stream);
/* dw[n][k] (+)= sum_m dy[m][n] * x[m][k]  (fp32 out) */
int jck_fc_wgrad(const void* dy, const float* x, float* dw, int M, int N, int K, int accumulate, int dtype,
                 void* stream);
/* G.conv1 weight w4[K][C][16] fp32 <-> fc layout w[n = tap*C + c][k] */
int jck_pack_fc(const float* w4, void* w_fc, int K, int C, int dtype, void* stream);
int jck_unpack_fc_grad(const float* dw_fc, float* dw4, int K, int C, int accumulate, void* stream);
/* G.conv1 on tcgen05 (jck_gemm_tc): weight as the MN-major operand w_t[k][n = tap*C + c] (bf16), its gradient back from
 * dw_t[k][n] (fp32), and fp32 rows x[M][K] -> bf16 rows of pitch ldo >= K (zero padded; TMA needs ldo % 8 == 0).
 * nn.ConvTranspose2d(100,512,4,1,0) model/DCGAN.py:42,62; CGAN.py:132 (200 inputs). */
int jck_pack_fc_t(const float* w4, void* w_t_bf16, int K, int C, void* stream);
int jck_unpack_fc_grad_t(const float* dw_t, float* dw4, int K, int C, int accumulate, void* stream);
int jck_cast_rows_bf16(const float* x, void* out_bf16, int M, int K, int ldo, void* stream);
/* out[m] = [a[m][0:K1] | b[m][0:K2] | 0 ...] (row pitch ldo), b fp32 or int64 (b_is_i64), out fp32 or bf16: the CGAN generator
 * input cat([z, labels.reshape(-1, n_classes, 1, 1)], 1) (model/CGAN.py:154-155) written straight into the operand of the
 * conv1 matrix product -- no torch.cat, no separate int64 -> float or fp32 -> bf16 pass. */
int jck_concat_rows(const float* a, const void* b, int b_is_i64, void* out, int M, int K1, int K2, int ldo, int dtype,
                    void* stream);
/* stream-ordered memset to zero (one graph memset node): the per-step accumulator arena (BatchNorm statistics / backward sums /
 * logged scalars), gradient buffers that accumulate.  Replaces torch.zeros / .zero_() inside the captured step. */
int jck_zero(void* p, size_t bytes, void* stream);
int jck_copy_f32(const float* src, float* dst, long long n, void* stream);

/* ---- BatchNorm2d (train mode) + activation ---------------------------------------------------
 * Replaces nn.BatchNorm2d + nn.LeakyReLU(0.2)/nn.ReLU model/DCGAN.py:11-12,43-44 (forward) and
 * aten::native_batch_norm_backward / leaky_relu_backward / threshold_backward (autograd).
 * finalize: for each group g in order: mean, biased var from stats[g] over `count` samples;
 *   running_mean/var momentum update with UNBIASED var; num_batches_tracked += 1;
 *   scale_shift[g] = {gamma*rstd, beta - mean*gamma*rstd};  mean_rstd[g] = {mean, rstd}. */
int jck_bn_finalize(const float* stats, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, long long* num_batches_tracked, float* scale_shift,
                    float* mean_rstd, int C, int groups, float count, float eps, float momentum,
                    void* stream);
/* a = act(scale*y + shift); slope 0 -> ReLU, 0.2 -> LeakyReLU.  pix_per_group pixels per group. */
int jck_bn_act_fwd(const void* y, const float* scale_shift, void* a, long long npix, int C,
                   long long pix_per_group, float slope, int dtype, int order, void* stream);
/* g = da * act'(scale*y+shift); sums[g] += {sum g, sum g*xhat} (caller zeroes sums [groups][2C]) */
int jck_bn_act_bwd_reduce(const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                          float* sums, long long npix, int C, long long pix_per_group, float slope,
                          int dtype, int order, void* stream);
/* dy = gamma*rstd*(g - sum_g/count - xhat*sum_gx/count) */
int jck_bn_act_bwd_apply(const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                         const float* gamma, const float* sums, void* dy, long long npix, int C,
                         long long pix_per_group, float count, float slope, int dtype, int order,
                         void* stream);
/* dgamma (+)= sum over groups of sum_gx ; dbeta (+)= sum over groups of sum_g */
int jck_bn_param_grad(const float* sums, float* dgamma, float* dbeta, int C, int groups, int accumulate,
                      void* stream);

/* ---- DCGAN discriminator head: conv5 (K = 16*C4 dot product) + sigmoid + BCE ------------------
 * Replaces nn.Conv2d(512,1,4,1,0)+nn.Sigmoid model/DCGAN.py:26-27,34 and nn.BCELoss
 * train/dcgan_trainer.py:64,163,174,186 (log clamped at -100 as torch does).
 * scalars (fp32, ADDED): scalars[0] += BCE mean, scalars[1] += mean(prob). w5 is NHWC-ordered. */
int jck_head_fwd(const void* a4, const void* w5, float* prob, float target, float* scalars, int B, int K,
                 int dtype, void* stream);
/* mode 0: dlogit = (p-target)/max(p(1-p),1e-12) * p(1-p) / mean_count   (BCE mean backward through sigmoid;
 *         mean_count = the batch the mean runs over: 0 -> B; the GLOBAL batch under data parallelism, so that the
 *         ranks' gradients SUM to the global-batch mean the reference computes)
 * mode 1: dlogit = p(1-p)                                   (grad_outputs = ones, the GP sweep)
 * mode 2: dlogit = dprob[b] * p(1-p)                        (caller-supplied gradient of the sigmoid output)
 * da4[b][k] = dlogit[b]*w5[k];  dw5[k] (+)= sum_b dlogit[b]*a4[b][k] when dw5 != NULL (fp32, NHWC order) */
int jck_head_bwd(const float* prob, const float* dprob, float target, const void* w5, const void* a4, void* da4,
                 float* dw5, int B, int mean_count, int K, int mode, int accumulate, int dtype, void* stream);
/* conv5 weight [1][C4][4][4] fp32 <-> NHWC-ordered [16*C4] */
int jck_pack_head(const float* w4, void* w5, int C4, int dtype, void* stream);
int jck_unpack_head_grad(const float* dw5, float* dw4, int C4, int accumulate, void* stream);

/* ---- generator output edge ---------------------------------------------------------------------
 * fake_raw = tanh(y5) (NCHW fp32); fake_mix = a*fake_raw + b*noise (NCHW fp32 and NHWC dtype).
 * Replaces nn.Tanh model/DCGAN.py:59,66 + train/dcgan_trainer.py:171.  Outputs nullable. */
int jck_g_out_fwd(const void* y5_nhwc, const float* noise, float a, float b, float* fake_raw_nchw,
                  float* fake_mix_nchw, void* fake_mix_nhwc, int B, int C, int H, int W, int layout, int dtype,
                  void* stream);
/* jck_g_out_fwd with the noise drawn in registers (see jck_prep_image_rng).  train/dcgan_trainer.py:171. */
int jck_g_out_fwd_rng(const void* y5_nhwc, unsigned long long seed, unsigned long long stream_id,
                      const unsigned long long* counter_base, float a, float b, float* fake_raw_nchw,
                      float* fake_mix_nchw, void* fake_mix_nhwc, int B, int C, int H, int W, int layout,
                      int dtype, void* stream);
/* dy5 = a * dmix * (1 - fake_raw^2)  (dmix NHWC dtype, fake_raw NCHW fp32, dy5 NHWC dtype) */
int jck_g_out_bwd(const void* dmix_nhwc, const float* fake_raw_nchw, float a, void* dy5_nhwc, int B, int C,
                  int H, int W, int layout, int dtype, void* stream);

/* ---- CGAN discriminator head and the second-order (gradient-penalty) sweep ---------------------------
 * model/CGAN.py:83-84,103-107,112-122 (label_embedding + LeakyReLU, flatten, cat, linear1, drop1, linear2,
 * sigmoid) and train/cgan_trainer.py:200-204 (error_d.backward() through the gradient penalty).
 * jck_dense: C[m][n] (+)= sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]; dtypes JCK_F32 / JCK_BF16 per operand. */
int jck_dense(const void* A, int a_dt, long long sam, long long sak, const void* B, int b_dt, long long sbn,
              long long sbk, void* C, int c_dt, long long ldc, int M, int N, int K, int accumulate, void* stream);
/* fp32 [M][N] helpers.  op 0: out = act(x + y[n]) (act slope s, y nullable); 1: out = x*y*s; 2: out = x*(y>0 ? 1 : s);
 * 3: out[m] = x[m] + y[m % rows_y]; 4: out[r] = sum_g x[g*rows_y + r]; 5: out[m][n] = x[m]*y[n]; 6: out[m][n] = x[m][n]*y[m]; 7: out = (x >= s) ? 1 : 0 */
int jck_rowop(int op, const float* x, const float* y, float* out, int M, int N, int rows_y, float s, void* stream);
/* prob = sigmoid(logit); scalars (nullable): [0] += BCE mean vs target, [1] += mean(prob) */
int jck_sigmoid_bce(const float* logit, float* prob, float target, float* scalars, int B, void* stream);
/* d/d(logit): mode 0 BCE-mean, 1 ones on prob, 2 up*p(1-p), 3 second order up*p(1-p)(1-2p); all times `scale` */
int jck_logit_grad(const float* prob, const float* up, float target, float* out, int B, int mode, float scale, void* stream);
int jck_i64_to_f32(const long long* in, float* out, long long n, void* stream);
int jck_f32_to_bf16(const float* in, void* out_bf16, long long n, void* stream);
/* FID feature moments on the tensor cores (metrics.py:118-124 np.mean / np.cov): hi + lo = x - mean as two bf16 matrices
 * of row pitch ldo (zero padded); the covariance is (hi^T hi + hi^T lo + lo^T hi) / (rows - 1) through jck_gemm_tc with
 * both operands MN-major. */
int jck_center_split_bf16(const float* x, const float* mean, void* hi_bf16, void* lo_bf16, long long rows, int d, int ldo,
                          void* stream);
/* The head's large products (Linear(16*C4 + E -> 256), CGAN.py:105,120: forward x.W^T, input gradient g.W, weight
 * gradient g^T.x and their second-order twins) on tcgen05:  C[m][n] (+)= sum_k A(m,k) * B(n,k), bf16 operands, fp32
 * accumulation.  Operand X is K-major (x_mn_major = 0: X[row*ldx + k]) or MN-major (1: X[k*ldx + row]); ldx % 8 == 0.
 * C: row-major, JCK_F32 (accumulate allowed) or JCK_BF16.  Split-K partials live in `workspace`
 * (jck_gemm_tc_workspace_bytes; 0 when the plan has one split).  `stats` (nullable, single-split plans): BatchNorm
 * sums of the product, stats[n % stats_channels] += sum_m C[m][n], stats[stats_channels + ...] += sum_m C[m][n]^2. */
size_t jck_gemm_tc_workspace_bytes(int M, int N, int K);
int jck_gemm_tc(const void* A, int a_mn_major, long long lda, const void* B, int b_mn_major, long long ldb, void* C,
                int c_dtype, long long ldc, int M, int N, int K, int accumulate, float* stats, int stats_channels,
                void* workspace, size_t workspace_bytes, void* stream);
int jck_axpy(const void* x, void* y, float a, long long n, int dtype, void* stream);   /* y += a*x */
/* per sample b: norm = ||v_b||; scalars[0] += (norm-1)^2/B; u_b = scale*(1 - 1/norm)*v_b (u nullable) */
int jck_gp_seed(const void* v, void* u, float* scalars, int B, long long per_sample, float scale, int dtype, void* stream);
/* linear1.weight [O][C*HW + E] (NCHW flatten order) <-> w_a [O][HW*C] (NHWC order, dtype), w_b [O][E] fp32 */
int jck_pack_linear(const float* w, void* w_a, float* w_b, int O, int C, int HW, int E, int dtype, void* stream);
int jck_unpack_linear_grad(const float* dwa, const float* dwb, float* dw, int O, int C, int HW, int E, int accumulate,
                           void* stream);
/* adjoint of the BatchNorm backward pass (formulas in csrc/cgan.cu): asums [3C] += {S1,S2,S3} (caller zeroes) */
int jck_bn_adj_reduce(const void* dbar, const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                      const float* sums1, float* asums, long long npix, int C, float count, float slope, int dtype, void* stream);
int jck_bn_adj_apply(const void* dbar, const void* da, const void* y, const float* scale_shift, const float* mean_rstd,
                     const float* gamma, const float* sums1, const float* asums, void* gbar_a, void* ybar, long long npix,
                     int C, float count, float slope, int dtype, void* stream);
int jck_bn_adj_param(const float* asums, const float* mean_rstd, float* dgamma, int C, float scale, void* stream);

/* ---- gradient penalty --------------------------------------------------------------------------
 * scalars[0] += mean_n (||dx[n,:]||_2 - 1)^2.  Replaces train/dcgan_trainer.py:125-126. */
int jck_gp_penalty(const void* dx, float* scalars, int B, long long per_sample, int dtype, void* stream);

/* ---- Adam (lr, betas, eps as torch.optim.Adam; no weight decay, no amsgrad) --------------------
 * One launch over a flat fp32 buffer.  step_count: device int32, the step number BEFORE this update
 * (kernel uses step_count+1; jck_adam_advance increments it) so a captured graph replays correctly.
 * Replaces optim.Adam.step train/dcgan_trainer.py:61-62,180,189. */
int jck_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
             float beta1, float beta2, float eps, const int* step_count, void* stream);
int jck_adam_advance(int* step_count, void* stream);

/* ---- random numbers (Philox4x32-10; performance mode only -- parity tests inject host tensors) --
 * counter_base: device uint64 advanced by jck_rng_advance.  Replaces torch.randn / torch.rand
 * train/dcgan_trainer.py:111,160,168,171. */
int jck_randn(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
              const unsigned long long* counter_base, void* stream);
int jck_rand(float* out, long long n, unsigned long long seed, unsigned long long stream_id,
             const unsigned long long* counter_base, void* stream);
int jck_rng_advance(unsigned long long* counter_base, unsigned long long by, void* stream);

/* ---- FID / IS evaluation: the Inception-v3 feature extractor of metrics.py on our kernels ----------------------
 * Replaces `self.inception_model(image)` metrics.py:87 (torchvision models.inception_v3, fc = Linear(2048, 100),
 * metrics.py:46-52), the pre-processing of the eval branch (train/dcgan_trainer.py:203-207, cgan_trainer.py:227-231)
 * and the score reduction metrics.py:96-110.  Activations: NHWC bf16 buffers described by
 *     geom = {Hb, Wb, by, bx, c_off}  + channel pitch ld:   pixel (b, y, x), channel c  at
 *     ((b*Hb + y + by)*Wb + x + bx)*ld + c_off + c        (a zero border of by / bx pixels, a slice of a wider concat).
 *
 * jck_conv_gemm: out = act(scale[n] * sum_{tap,c} A[m + shift[tap]][c] * W[n][tap*Cp + c] + bias[n]) on tcgen05.
 *   A: bf16 rows [rows_a][C], pitch lda (% 8 == 0); rows outside [0, rows_a) read as zero.  W: bf16 [N][ntaps*Cp],
 *   Cp = C rounded up to 64, or 32 when C <= 32 (zero filled).  scale / bias: fp32 [N], nullable (1 / 0).  Row m of the M = B*Hq*Wq computed
 *   rows is grid position (b, Y, X); it is stored iff oy = Y - oy0 in [0, Ho) and ox = X - ox0 in [0, Wo), at
 *   out[((b*Hob + oy + opy)*Wob + ox + opx)*ldc + c_off + n], dtype JCK_BF16 or JCK_F32.
 *   geom = {M, N, C, ntaps, Hq, Wq, oy0, ox0, Ho, Wo, Hob, Wob, opy, opx, c_off, relu, out_dtype, rows_a, shift[ntaps]}
 *   (host array).  A stride-1 kh x kw convolution with padding (py, px) over a buffer whose border is >= the padding:
 *   tap (ky, kx) has shift (ky - py)*Wb + (kx - px), oy0 = by, ox0 = bx; a 1x1 convolution, a Linear layer or a patch
 *   matrix from jck_im2col is the one-tap case. */
int jck_conv_gemm(const void* act, long long lda, const void* w, const float* scale, const float* bias, void* out,
                  long long ldc, const int* geom, int ngeom, void* stream);
/* patches[(b, oy, ox)][(ky*kw + kx)*C + c] (bf16, row pitch Kp % 8 == 0, zero beyond kh*kw*C and outside the image):
 * the stride-2 convolutions and the 3-channel stem.  in_geom as above (host array of 5 ints). */
int jck_im2col(const void* x, const int* in_geom, long long ldx, void* patches, int B, int H, int W, int C, int kh, int kw,
               int sy, int sx, int py, int px, int Ho, int Wo, int Kp, void* stream);
/* 3x3 pooling: mode 0 = max (F.max_pool2d(x, 3, stride), no padding), 1 = average over the zero-padded window / 9
 * (F.avg_pool2d(x, 3, 1, 1), count_include_pad).  Channels and pitches multiples of 8. */
int jck_pool3(const void* x, const int* in_geom, long long ldx, void* out, const int* out_geom, long long ldo, int B, int H,
              int W, int C, int stride, int pad, int Ho, int Wo, int mode, void* stream);
/* adaptive_avg_pool2d(x, 1) of a dense [B][HW][C] bf16 tensor -> fp32 and / or bf16 [B][C] */
int jck_global_avgpool(const void* x, float* out_f32, void* out_bf16, int B, int HW, int C, void* stream);
/* out[b][y][x][c] (NHWC bf16, channel pitch ldo, pad channels zero) = (a*bilinear(in_nchw)[b][c][y][x] + b - mean[c]) / std[c];
 * bilinear = F.resize / F.interpolate(align_corners=False) for an enlargement.  mean3 / std3: HOST pointers, 3 floats. */
int jck_resize_norm(const float* in_nchw, void* out_nhwc, int B, int C, int Hi, int Wi, int Ho, int Wo, int ldo, float a,
                    float b, const float* mean3, const float* std3, void* stream);
/* jck_resize_norm + jck_im2col of the 3x3 stride-2 stem (Conv2d_1a_3x3) in one pass: patches[(b, oy, ox)][(ky*3 + kx)*3 + c]
 * (bf16, row pitch 32, columns 27..31 zero) of the Hr x Wr resized, normalised 3-channel image, never materialised. */
int jck_stem_patches(const float* in_nchw, void* patches, int B, int Hi, int Wi, int Hr, int Wr, float a, float b,
                     const float* mean3, const float* std3, void* stream);
/* ---- split-precision mode of the extractor (fp32-grade features on the bf16 tensor cores) --------------------------
 * Every activation is kept as TWO bf16 planes, value = hi + lo with hi = bf16(v), lo = bf16(v - hi) (16 mantissa bits),
 * the low plane a fixed element offset behind the high one; weights are split the same way on the host.  A product is
 * hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM (lo*lo, 2^-18 relative, is dropped).  No new GEMM kernel is needed:
 * with the two planes stacked along the row dimension of A, "the low plane at tap shift s" is simply the shift s + R
 * (R = rows of one plane), so the caller passes 3*ntaps taps {s, s, s + R} against the weight blocks {W_hi, W_lo, W_hi}.
 * jck_conv_gemm writes a split OUTPUT when geom carries one more entry after the shifts:
 *     geom[18 + ntaps] = rows of one output plane   (low plane at out + that * ldc; bf16 output only).
 * The streaming kernels take the plane offsets (in elements, multiples of 8; 0 = plain bf16 tensor): */
int jck_pool3_split(const void* x, const int* in_geom, long long ldx, long long x_lo, void* out, const int* out_geom,
                    long long ldo, long long out_lo, int B, int H, int W, int C, int stride, int pad, int Ho, int Wo, int mode,
                    void* stream);
int jck_global_avgpool_split(const void* x, long long x_lo, float* out_f32, void* out_bf16, long long out_lo, int B, int HW,
                             int C, void* stream);
int jck_stem_patches_split(const float* in_nchw, void* patches, long long patches_lo, int B, int Hi, int Wi, int Hr, int Wr,
                           float a, float b, const float* mean3, const float* std3, void* stream);
/* metrics.py:96-110: scores[k] = exp(mean_i KL(softmax(logits_i) || mean_i softmax(logits_i))) over rows
 * [k*(n/splits), (k+1)*(n/splits)) of fp32 logits [n][d] */
int jck_inception_score(const float* logits, int n, int d, int splits, float* scores, void* stream);

/* ---- input pipeline on the device (preprocess/dcgan_data_preprocessor.py:38-49, cgan_data_preprocessor.py:11-16,51-62) ----
 * The uint8 dataset [N][Hi][Wi][C] lives in HBM.  out[b] = Normalize(ToTensor(Resize(data[index[b]]))) as NCHW fp32, bit-exact
 * with Pillow's bilinear resample (horizontal pass, then vertical pass on the uint8-rounded result; 22-bit fixed-point
 * coefficients: *_bounds [out][2] = (first source index, count), *_coef [out][ksize], device int32 tables built by the host as
 * Pillow's precompute_coeffs / normalize_coeffs_8bpc do) and with torchvision's fp32 (u8 / 255 - mean) / std.  index: device
 * int64 [B] (nullable = identity); mean / std: HOST pointers to C floats; a pass whose size does not change is skipped. */
int jck_u8_resize_norm(const void* data_u8, const long long* index, float* out_nchw, int B, int Hi, int Wi, int C, int Ho,
                       int Wo, const int* h_bounds, const int* h_coef, int h_ksize, const int* v_bounds, const int* v_coef,
                       int v_ksize, const float* mean, const float* std, void* stream);
/* out[b][j] = (labels[index[b]] == j), int64 (OneHotEncoder + default collate) */
int jck_one_hot_i64(const long long* labels, const long long* index, long long* out, int B, int n_classes, void* stream);

/* ---- data-parallel exchange over NVLink peer memory -------------------------------------------
 * The reference is single-GPU: these entry points have no reference counterpart.  They make
 * an N-GPU run equal the reference at the global batch: nn.BatchNorm2d's batch statistics (model/DCGAN.py:
 * 11,15,19,23,43,47,51,55) are taken over ALL ranks' rows.  One process per GPU; each rank creates a
 * communicator, the 64-byte IPC handles are exchanged by the host (torch.distributed all_gather) and
 * connected.  A small all-reduce is then one single-CTA kernel per rank that stores (value, sequence) words
 * into every peer's mailbox and spins on its own -- no host call, no NCCL launch, CUDA-graph capturable.
 * All ranks must issue the same sequence of jck_comm_* / *_sync calls. */
int jck_comm_create(int rank, int world, void** comm_out, void* ipc_handle_out /* JCK_COMM_HANDLE_BYTES */);
int jck_comm_connect(void* comm, const void* all_handles /* world x JCK_COMM_HANDLE_BYTES, rank order */);
int jck_comm_destroy(void* comm);
/* early_dependents = 0 (the default): this communicator's kernels never execute griddepcontrol.launch_dependents (their
 * dependents start when the exchange has completed).  At most ONE communicator of a process may have it on when exchanges
 * are issued from more than one stream -- see csrc/comm.cu for the cross-rank deadlock this prevents. */
int jck_comm_configure(void* comm, int early_dependents);
/* *flag_out <- 1 if any exchange on this communicator ever timed out waiting for a peer (JCK_COMM_TIMEOUT_S seconds,
 * default 120; the timed-out exchange returned NaN instead of a sum), else 0.  Synchronises with the device. */
int jck_comm_error(void* comm, int* flag_out);
/* data[0..n) <- sum over ranks, in place, n <= JCK_COMM_MAX_N; identical bits on every rank */
int jck_comm_allreduce_small(void* comm, float* data, int n, void* stream);
/* SyncBN forward: all-reduce stats[groups][2C] in place, then exactly jck_bn_finalize (count = GLOBAL samples) */
int jck_bn_finalize_sync(void* comm, float* stats, const float* gamma, const float* beta, float* running_mean,
                         float* running_var, long long* num_batches_tracked, float* scale_shift, float* mean_rstd, int C,
                         int groups, float count, float eps, float momentum, void* stream);
/* SyncBN backward: jck_bn_param_grad on the LOCAL sums (skipped when dgamma == dbeta == NULL), then all-reduce
 * sums[groups][2C] in place for jck_bn_act_bwd_apply */
int jck_bn_bwd_sums_sync(void* comm, float* sums, float* dgamma, float* dbeta, int C, int groups, int accumulate,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* JCK_B200_H_ */
