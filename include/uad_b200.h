/*
 * libuad_b200 - C ABI of the B200 (sm_100a) hot path for unsupervised brain-MRI anomaly detection.
 *
 * The reference (StefanDenn3r/Unsupervised_Anomaly_Detection_Brain_MRI) has no FFI of its own: its hot path runs inside
 * tf.Session.run (trainers/AE.py:83, trainers/VAE.py:96, trainers/ceVAE.py:107) on graphs built from
 * models/customlayers.py:16-38 and lowered to TensorFlow library ops.  Each entry point below replaces one of those
 * library ops (or a fused group of them); the reference interface it replaces is cited per function
 * (paths relative to the reference repo root).
 *
 * Conventions
 *  - every tensor pointer is a caller-owned DEVICE pointer, NHWC / row-major contiguous fp32 unless stated
 *  - functions enqueue on `stream` (a cudaStream_t passed as void*) and return without synchronising
 *  - no allocation inside hot calls: workspace is caller-provided (query with uad_conv_workspace_bytes)
 *  - return 0 on success; non-zero on error with a message in uad_last_error() (thread-local)
 *  - `accumulate` != 0 means "+=" into the gradient output (shared-weight branches, e.g. ceVAE), else overwrite
 *  - all reductions are deterministic (two-stage, no floating-point atomics)
 */
#ifndef UAD_B200_H_
#define UAD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UAD_ABI_VERSION 1

#define UAD_ACT_NONE 0
#define UAD_ACT_LEAKY 1   /* tf.keras.layers.LeakyReLU(alpha)  models/customlayers.py:23,36 */
#define UAD_ACT_RELU 2    /* tf.keras.layers.ReLU              models/customlayers.py:31    */
#define UAD_ACT_SIGMOID 3 /* models/fanogan.py:41,46 */
#define UAD_ACT_TANH 4    /* models/fanogan.py:29    */
/* OR-ed into `act` of uad_act_bn_bwd / uad_final1x1_l1_bwd_fused: the `z` argument holds the block's OUTPUT
 * a = act(gamma*bn_c*z+beta) instead of z (piecewise-linear activations only), so a training forward need not write z
 * at all (z_out = NULL): halves the forward's HBM write volume.  Requires gamma != 0 (dgamma divides by gamma). */
#define UAD_ACT_FROM_OUTPUT 0x100

#define UAD_OP_CONV_FWD 0
#define UAD_OP_CONV_DGRAD 1
#define UAD_OP_CONV_WGRAD 2
#define UAD_OP_CONVT_FWD 3
#define UAD_OP_CONVT_DGRAD 4
#define UAD_OP_CONVT_WGRAD 5

/* math mode of the conv kernels: 0 = fp32 SIMT FFMA; 1 = tcgen05 3xTF32 (fp32-accurate split); 2 = tcgen05 1xTF32 */
#define UAD_MATH_FP32_SIMT 0
#define UAD_MATH_TC_3XTF32 1
#define UAD_MATH_TC_1XTF32 2 /* ONE tf32 MMA per K-step on operands rounded to nearest tf32, fp32 storage / accumulation: ~1e-3 relative, the
                                 single-pass tensor-core arithmetic of config C4 - not the 1e-4 parity mode.  Shapes only the first-generation
                                 kernels cover run 3xTF32 in this mode too.  Any other value is rejected (bf16 storage is not built). */

const char* uad_last_error(void);
int uad_abi_version(void);
/* number of CUDA kernels this library has launched in this process (diagnostic; bench.py reports it) */
long long uad_launch_count(void);
/* 1 if the tcgen05 (tensor-core) conv path is compiled in and usable for the given op/shape, else 0 */
int uad_conv_tc_supported(int op, int B, int H, int W, int Cin, int Cout, int ksize);

/* Bytes of scratch an op needs. (H,W) is the spatial size of the op's LOW-resolution... see each op: it is always the
 * size of the tensor called `x` in the forward direction of the layer (conv: input; convT: input). */
size_t uad_conv_workspace_bytes(int op, int B, int H, int W, int Cin, int Cout, int ksize, int math_mode);

/* ---- strided conv block: Conv2D(k, s=2, 'same') + bias -> frozen BatchNormalization -> activation
 * replaces tf Conv2D + BatchNormalization(inference affine) + LeakyReLU  (models/customlayers.py:21-23)
 * x [B,H,W,Cin], w HWIO [k,k,Cin,Cout], z_out/a_out [B,H/2,W/2,Cout].
 * z = conv(x)+bias; a = act(gamma*bn_c*z + beta)  (gamma/beta NULL => identity affine).  z_out or a_out may be NULL. */
int uad_conv2d_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                   float* z_out, float* a_out, int B, int H, int W, int Cin, int Cout, int ksize, int act, float alpha,
                   float bn_c, int math_mode, void* ws, size_t ws_bytes, void* stream);
/* input gradient of the conv (tf.gradients -> Conv2DBackpropInput): dz [B,H/2,W/2,Cout] -> dx [B,H,W,Cin] */
int uad_conv2d_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cin, int Cout, int ksize,
                     int math_mode, void* ws, size_t ws_bytes, void* stream);
/* filter gradient (Conv2DBackpropFilter): dw HWIO */
int uad_conv2d_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                     int accumulate, int math_mode, void* ws, size_t ws_bytes, void* stream);

/* ---- transposed conv block: Conv2DTranspose(k, s=2, 'same') + bias -> frozen BN -> activation
 * replaces tf Conv2DTranspose + BatchNormalization + LeakyReLU (models/customlayers.py:34-36)
 * x [B,H,W,Cin], w [k,k,Cout,Cin] (TF layout), outputs [B,2H,2W,Cout]. */
int uad_convT2d_fwd(const float* x, const float* w, const float* bias, const float* gamma, const float* beta,
                    float* z_out, float* a_out, int B, int H, int W, int Cin, int Cout, int ksize, int act, float alpha,
                    float bn_c, int math_mode, void* ws, size_t ws_bytes, void* stream);
/* the last decoder block fused with the 1x1 conv that follows it (models/customlayers.py:34-37: Conv2DTranspose + BN + LeakyReLU,
 * then Conv2D(C, 1)): a_out as uad_convT2d_fwd, head_out[b, y, x] = sum_c a_out[b, y, x, c] * head_w[c] + head_b[0].
 * Tensor-core path only (Cout = 32); ask uad_convT2d_fwd_head_supported first and fall back to uad_convT2d_fwd +
 * uad_final1x1_l1_fwd when it returns 0.  The L1 residual / per-sample sums then come from uad_l1_map. */
int uad_convT2d_fwd_head_supported(int B, int H, int W, int Cin, int Cout, int ksize, int math_mode);
int uad_convT2d_fwd_head(const float* x, const float* w, const float* bias, const float* gamma, const float* beta, float* a_out,
                         const float* head_w, const float* head_b, float* head_out, int B, int H, int W, int Cin, int Cout,
                         int ksize, int act, float alpha, float bn_c, int math_mode, void* ws, size_t ws_bytes, void* stream);
int uad_convT2d_dgrad(const float* dz, const float* w, float* dx, int B, int H, int W, int Cin, int Cout, int ksize,
                      int math_mode, void* ws, size_t ws_bytes, void* stream);
int uad_convT2d_wgrad(const float* x, const float* dz, float* dw, int B, int H, int W, int Cin, int Cout, int ksize,
                      int accumulate, int math_mode, void* ws, size_t ws_bytes, void* stream);

/* ---- backward of "z -> frozen BN -> activation" plus the bias gradient (tf.gradients through
 * BatchNormalization/LeakyReLU/BiasAdd).  rows = B*H*W, C channels.
 * du = da*act'(gamma*bn_c*z+beta); dgamma = bn_c*sum(du*z); dbeta = sum(du); dz = gamma*bn_c*du; dbias = sum(dz).
 * dz may alias da.  gamma/beta NULL => identity affine (dgamma/dbeta ignored).  ws: >= uad_rowreduce_workspace_bytes. */
size_t uad_rowreduce_workspace_bytes(long long rows, int C);
int uad_act_bn_bwd(const float* da, const float* z, const float* gamma, const float* beta, float* dz, float* dgamma,
                   float* dbeta, float* dbias, long long rows, int C, int act, float alpha, float bn_c, int accumulate,
                   void* ws, size_t ws_bytes, void* stream);

/* ---- dense / 1x1-conv block: y = dropout(x.W + b) -> optional frozen BN -> activation
 * replaces tf Dense / Keras Conv2D(1x1) / Dropout (models/autoencoder.py:20-30, variational_autoencoder.py:20-36)
 * x [M,K], w [K,N], mask [M,N] of {0,1} or NULL, z = (x.W+b)*mask*mask_scale, a = act(gamma*bn_c*z+beta). */
int uad_dense_fwd(const float* x, const float* w, const float* bias, const float* mask, float mask_scale,
                  const float* gamma, const float* beta, float* z_out, float* a_out, int M, int K, int N, int act,
                  float alpha, float bn_c, void* ws, size_t ws_bytes, void* stream);
/* scratch for uad_dense_fwd / uad_dense_bwd (deterministic split-K partials); smaller buffers only reduce parallelism */
size_t uad_dense_workspace_bytes(int M, int K, int N);
/* dz [M,N] is the gradient w.r.t. z (post-dropout); dx may be NULL */
int uad_dense_bwd(const float* x, const float* w, const float* dz, const float* mask, float mask_scale, float* dx,
                  float* dw, float* dbias, int M, int K, int N, int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ---- reparameterise + KL (models/variational_autoencoder.py:33-34, trainers/VAE.py:38)
 * sigma=exp(ls); z=mu+eps*sigma; kl[b]=0.5*sum_j(mu^2+sigma^2-log(sigma^2)-1).  eps NULL => z=mu (ceVAE ce-branch) */
int uad_reparam_kl_fwd(const float* mu, const float* log_sigma, const float* eps, float* sigma, float* z, float* kl,
                       int B, int Z, void* stream);
/* dmu = dz + kl_scale*mu ; dls = dz*eps*sigma + kl_scale*(sigma^2-1)   (kl_scale = 1/B for loss=mean_b(rec+kl)) */
int uad_reparam_kl_bwd(const float* mu, const float* log_sigma, const float* eps, const float* dz, float kl_scale,
                       float* dmu, float* dls, int B, int Z, void* stream);

/* ---- fused final 1x1 conv (Cin->1) + L1 residual + per-sample sums
 * replaces dec_Conv2D_final (models/customlayers.py:37) + tf.losses.absolute_difference + reduce_sum
 * (trainers/AE.py:28-29).  a [B*HW, Cin], w [Cin], b [1], x [B*HW] -> xhat [B*HW], l1 [B*HW] (nullable), rec[B] (nullable) */
int uad_final1x1_l1_fwd(const float* a, const float* w, const float* bias, const float* x, float* xhat, float* l1,
                        float* rec, int B, int HW, int Cin, void* ws, size_t ws_bytes, void* stream);
/* dxhat = sign(xhat-x)*scale (sign(0)=0); da[p,c] = dxhat*w[c]; dw[c] = sum dxhat*a[p,c]; db = sum dxhat. */
int uad_final1x1_l1_bwd(const float* a, const float* w, const float* x, const float* xhat, float scale, float* da,
                        float* dw, float* dbias, int B, int HW, int Cin, int accumulate, void* ws, size_t ws_bytes,
                        void* stream);

/* fused variant for training: backward of the final 1x1+L1 AND of the preceding "z -> frozen BN -> act" block in one pass
 * (reads z, x, xhat; writes dz; produces dw/dbias of the 1x1 and dgamma/dbeta/dbias_prev of the block) - equals
 * uad_final1x1_l1_bwd followed by uad_act_bn_bwd without materialising da (saves ~3 activation-sized HBM passes). */
int uad_final1x1_l1_bwd_fused(const float* z, const float* gamma, const float* beta, const float* w, const float* x,
                              const float* xhat, float scale, float* dz, float* dgamma, float* dbeta, float* dbias_prev,
                              float* dw, float* dbias, int B, int HW, int Cin, int act, float alpha, float bn_c,
                              int accumulate, void* ws, size_t ws_bytes, void* stream);

/* ---- loss scalars: out[0]=mean(rec), out[1]=mean(kl) (0 if kl NULL), out[2]=mean(rec+kl) (trainers/VAE.py:40-42) */
int uad_loss_scalars(const float* rec, const float* kl, float* out3, int B, void* stream);

/* ---- TF-form Adam on a flat buffer (trainers/DLMODEL.py:112-131; tf.train.AdamOptimizer):
 * g = grad*grad_scale; m=b1*m+(1-b1)g; v=b2*v+(1-b2)g^2; p -= lr_t*m/(sqrt(v)+eps), lr_t precomputed by the host */
int uad_adam_tf_step(float* params, const float* grads, float* m, float* v, size_t n, float lr_t, float b1, float b2,
                     float eps, float grad_scale, const int64_t* step_dev, void* stream);
/* step_dev (nullable): device step counter t (1-based).  When given, `lr_t` is the BASE learning rate and the kernel
 * derives lr_t = lr*sqrt(1-b2^t)/(1-b1^t) itself (float64), so a captured CUDA graph needs no per-step host value. */

/* ---- data-parallel optimiser step as ONE kernel over NVLink peer memory (csrc/uad_peer.cu): reduce-scatter of the ranks' flat
 * gradient buffers, TF-form Adam (the arithmetic of uad_adam_tf_step; reference trainers/DLMODEL.py:112-131) on the rank's own
 * shard, all-gather of the updated parameters.  Replaces "all-reduce (NCCL) + uad_adam_tf_step".  Each rank allocates one region
 * [params | grads | flags] of uad_peer_region_bytes(numel) bytes with uad_peer_alloc, publishes its 64-byte IPC handle
 * (uad_peer_ipc_handle) through the host-side process group and maps the other ranks' regions with uad_peer_ipc_open;
 * regions[j] is rank j's region as mapped in the calling process.  m / v are local (only the rank's shard is used).
 * Every rank must issue the call the same number of times; a peer that never arrives traps after ~2 min instead of hanging for good. */
size_t uad_peer_region_bytes(size_t numel);
int uad_peer_alloc(size_t bytes, void** region_out);
int uad_peer_free(void* region);
int uad_peer_ipc_handle(void* region, void* handle64);
int uad_peer_ipc_open(const void* handle64, void** region_out);
int uad_peer_ipc_close(void* region);
int uad_peer_adam_step(void* const* regions, int rank, int world, size_t numel, size_t offset, size_t count, float* m, float* v,
                       float lr, float b1, float b2, float eps, float grad_scale, const int64_t* step_dev, void* stream);
/* [offset, offset + count): the slice of the flat index space this optimiser owns (everything, or one f-AnoGAN scope); m / v
 * point at the slice's first element. */

/* ---- Philox-4x32-10 streams for the live graph RNG nodes (tf.random_normal variational_autoencoder.py:34; Dropout) */
int uad_randn(float* out, size_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
int uad_dropout_mask(float* mask, size_t n, float rate, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     void* stream);
/* offset_dev (nullable): device counter added to `offset`, advanced with uad_counter_add (CUDA-graph friendly) */
int uad_counter_add(uint64_t* counter_dev, uint64_t inc, void* stream);

/* ---- LayerNormalization(axis=[1,2]) + activation (models/customlayers.py:22,30,35 with use_batchnorm=False; f-AnoGAN G / D):
 * mean / variance over (H,W) per (sample, channel), gamma / beta [H*W], y = act((x-mean)/sqrt(var+eps)*gamma[hw]+beta[hw]) */
size_t uad_layernorm_hw_workspace_bytes(int B, int HW, int C);
int uad_layernorm_hw_fwd(const float* x, const float* gamma_hw, const float* beta_hw, float* y, int B, int HW, int C,
                         float eps, int act, float alpha, void* ws, size_t ws_bytes, void* stream);
/* y = act(x) (tanh / sigmoid heads: models/fanogan.py:29,41,46) */
int uad_activation(const float* x, float* y, size_t n, int act, float alpha, void* stream);

/* ---- residual-map scoring (utils/Evaluation.py:282-291):
 * d = keep_positive ? max(x-xhat,0) : |x-xhat| (fp32); d *= mask (uint8, nullable); if apply_prior and (double)x < prior: d=0 */
int uad_residual_score(const float* x, const float* xhat, const uint8_t* mask, double prior_quantile, int keep_positive,
                       int apply_prior, float* diff, size_t n, void* stream);
/* ---- threshold + Dice counts (utils/Evaluation.py:453-457, trainers/Metrics.py:67-72):
 * for each threshold t_i (float64): P = (double)diff > t_i; counts[i] = {sum P*G, sum P, sum G} as int64.
 * mask_out (nullable, uint8 [n]) receives P for thresholds[0].  label uint8 (nullable => G=0). n_thr <= 32. */
int uad_threshold_counts(const float* diff, const uint8_t* label, size_t n, const double* thresholds_host, int n_thr,
                         int64_t* counts_dev, uint8_t* mask_out, void* stream);

/* ---- ceVAE anomaly map (trainers/ceVAE.py:51): anomaly = l1 * |gx| */
int uad_mul_abs(const float* l1, const float* gx, float* out, size_t n, void* stream);
/* gx[i] += -sign(xhat[i]-x[i])*scale : the direct dependence of |xhat-x| on x in d loss_vae/dx (trainers/ceVAE.py:51) */
int uad_l1_direct_term(const float* x, const float* xhat, float scale, float* gx, size_t n, void* stream);
/* developer aid: clock64 trace (64 int64, host buffer) of one CTA of the last N=32 tcgen05 launch run with UAD_TC_DEBUG=16 */
int uad_debug_trace(long long* out64_host);
/* y = a*x + b*y elementwise (gradient bookkeeping on flat buffers) */
int uad_axpby(float a, const float* x, float b, float* y, size_t n, void* stream);


/* ================= f-AnoGAN training (trainers/fAnoGAN.py:50-77; models/fanogan.py:50-82) =================
 * LayerNormalization([1,2]) with saved statistics, its backward, forward-mode derivative and joint backward: the pieces
 * of the WGAN-GP critic step (tf.gradients(d_hat, x_hat) differentiated again w.r.t. the critic weights, fAnoGAN.py:55-58).
 * stats = {mean[B*C], rstd[B*C]}; jstats = {mean(xdot)[B*C], mean(xdot*xhat)[B*C]}.  C in {8..128} power of two. */
size_t uad_layernorm_hw_train_workspace_bytes(int B, int HW, int C);
int uad_layernorm_hw_fwd_train(const float* x, const float* gamma_hw, const float* beta_hw, float* y, float* stats, int B,
                               int HW, int C, float eps, int act, float alpha, void* ws, size_t ws_bytes, void* stream);
/* dy = gradient w.r.t. the activated output; dx may alias dy; dgamma/dbeta [HW] nullable (both or neither) */
int uad_layernorm_hw_bwd(const float* dy, const float* x, const float* stats, const float* gamma_hw, const float* beta_hw,
                         float* dx, float* dgamma, float* dbeta, int B, int HW, int C, int act, float alpha, int accumulate,
                         void* ws, size_t ws_bytes, void* stream);
/* ydot = d/de act(LN(x + e*xdot)) at e = 0 */
int uad_layernorm_hw_jvp(const float* xdot, const float* x, const float* stats, const float* gamma_hw, const float* beta_hw,
                         float* ydot, float* jstats, int B, int HW, int C, int act, float alpha, void* ws, size_t ws_bytes,
                         void* stream);
/* reverse of the pair (y, ydot): adjoints dydot and dy (nullable) -> dxdot, dx, dgamma/dbeta */
int uad_layernorm_hw_bwd2(const float* dydot, const float* dy, const float* x, const float* xdot, const float* stats,
                          const float* jstats, const float* gamma_hw, const float* beta_hw, float* dxdot, float* dx,
                          float* dgamma, float* dbeta, int B, int HW, int C, int act, float alpha, int accumulate, void* ws,
                          size_t ws_bytes, void* stream);
/* backward of the final 1x1 conv (Cin -> 1) for an arbitrary incoming gradient dxhat [B*HW] (generator head, fanogan.py:41,46) */
int uad_final1x1_bwd(const float* a, const float* w, const float* dxhat, float* da, float* dw, float* dbias, int B, int HW,
                     int Cin, int accumulate, void* ws, size_t ws_bytes, void* stream);
/* dx = dy * act'(u), u the pre-activation (sigmoid / tanh heads) */
int uad_activation_bwd(const float* dy, const float* u, float* dx, size_t n, int act, float alpha, void* stream);
int uad_fill(float* y, float v, size_t n, void* stream);
/* uniform [0,1) Philox stream (tf.random_uniform, fanogan.py:67) */
int uad_uniform(float* out, size_t n, uint64_t seed, uint64_t offset, const uint64_t* offset_dev, void* stream);
/* x_hat = x + alpha[b]*(x_gen - x)  (fanogan.py:67-69) */
int uad_interpolate(const float* x, const float* x_gen, const float* alpha, float* out, int B, size_t per_sample,
                    void* stream);
/* deterministic reductions: out = scale*sum(x);  loss = loss_scale*sum((a-b)^2) and grad_a = grad_scale*(a-b) (nullable) */
size_t uad_reduce_workspace_bytes(void);
int uad_sum_scaled(const float* x, size_t n, double scale, float* out_dev, void* ws, size_t ws_bytes, void* stream);
int uad_mse(const float* a, const float* b, size_t n, float grad_scale, float* grad_a, double loss_scale, float* loss_out_dev,
            void* ws, size_t ws_bytes, void* stream);
/* gradient penalty (fAnoGAN.py:56-57): slope[b,j] = sqrt(sum_h ddx[b,h,j]^2) (axis 1 ONLY, as the reference),
 * gp = scale*mean((slope-1)^2), u = d gp / d ddx.  ddx [B,H,WC]; ws >= B*WC floats. */
int uad_gradient_penalty(const float* ddx, int B, int H, int WC, float scale, float* u_out, float* gp_out_dev, void* ws,
                         size_t ws_bytes, void* stream);
/* l1 = |xhat - x| (nullable), rec[b] = sum_hw l1 (nullable)  (fAnoGAN.py:65-66) */
int uad_l1_map(const float* x, const float* xhat, float* l1, float* rec, int B, int HW, void* stream);

/* ================= evaluation post-processing stencils (utils/Evaluation.py:84-89, 108-110, 311-312) =================
 * brain-mask erosion: scipy.ndimage.binary_erosion(mask, generate_binary_structure(2,1), iterations) per [H,W] slice of
 * mask [N,H,W] uint8, border_value 0.  iterations in [1,24] (the reference uses 12). */
int uad_binary_erosion_cross(const uint8_t* mask, uint8_t* out, int N, int H, int W, int iterations, void* stream);
/* scipy.ndimage.median_filter(volume, (5,5,5)) on a float32 volume [Z,H,W], boundary mode 'reflect'; out != vol */
int uad_median_filter3d_5(const float* vol, float* out, int Z, int H, int W, void* stream);

/* ================= spatial-bottleneck models (models/autoencoder_spatial.py:16-23) =================
 * z = x*(mask ? mask*keep : 1) (tf.keras Dropout on the encoder output), a = act(gamma*bn_c*z + beta) (the decoder's leading
 * BatchNormalization + ReLU, models/customlayers.py:30-31).  x, mask, z_out, a_out: [rows, C]; z_out / a_out nullable. */
int uad_mask_bn_act_fwd(const float* x, const float* mask, float keep, const float* gamma, const float* beta, float bn_c, int act,
                        float alpha, float* z_out, float* a_out, long long rows, int C, void* stream);
/* y = x*(mask ? mask : 1)*scale  (Dropout backward; y may alias x) */
int uad_mask_scale(const float* x, const float* mask, float scale, float* y, size_t n, void* stream);

/* ================= iterative MAP restoration (trainers/VAE_You.py:53-54,125-147; GMVAE.py:166-197) =================
 * One iteration = forward, g = dL/dxhat seed, dgrad chain to the input (uad_final1x1_bwd, uad_act_bn_bwd with NULL
 * parameter-gradient outputs, uad_conv*_dgrad, uad_dense_bwd with dw = NULL), update - all resident on the device.
 * seed: g = sign(xhat - x) - tv_lambda * dTV(d)/dd at d = x - xhat (tf.image.total_variation, sign(0) = 0);
 *       tv[b] (nullable) = TV(x - xhat) per image.  x, xhat, g: [B,H,W] single channel.  ws >= uad_tv_restore_workspace_bytes. */
size_t uad_tv_restore_workspace_bytes(int B, int H, int W);
int uad_tv_restore_seed(const float* x, const float* xhat, float tv_lambda, float* g, float* tv, int B, int H, int W, void* ws,
                        size_t ws_bytes, void* stream);
/* x <- x - lr*(gx - g)   (gx: gradient through the network; -g: direct dependence of the L1 and TV terms on x);
 * grads_out (nullable) receives gx - g, the tensor the reference fetches as losses['grads'] (VAE_You.py:54). */
int uad_restore_update(float* x, const float* gx, const float* g, float lr, float* grads_out, size_t n, void* stream);

/* ================= GMVAE latent block (models/gaussian_mixture_variational_autoencoder.py:64-71; trainers/GMVAE.py:66-88) =========
 * per sample: logit_c = sum_j(-0.5 (z_s_j - M_jc)^2 e^{S_jc} - S_jc + log pi), pc = softmax(logit) [B,dc] (nullable output);
 * con[b] = sum_c pc_c sum_j kl_jc with kl_jc = 0.5((e^{z_ls_j} + (z_mu_j - M_jc)^2)(e^{S_jc} + 1e-6) - S_jc - z_ls_j - 1)
 * (conditional_prior_loss); closs[b] = max(sum_c pc_c log(pc_c dc + 1e-8), c_lambda) (c_prior_loss).
 * z_mu, z_ls, z_s: [B,dz]; M = z_wc_mus, S = z_wc_log_sigma_invs: [B,dz,dc]; dc <= 32. */
int uad_gmvae_latent_fwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float* pc,
                         float* con, float* closs, int B, int dz, int dc, float c_lambda, void* stream);
/* gradients of scale*(con[b] + closs[b]) w.r.t. the five inputs (z_s as an independent input) */
int uad_gmvae_latent_bwd(const float* z_mu, const float* z_ls, const float* z_s, const float* M, const float* S, float scale,
                         float* dz_mu, float* dz_ls, float* dz_s, float* dM, float* dS, int B, int dz, int dc, float c_lambda,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UAD_B200_H_ */
