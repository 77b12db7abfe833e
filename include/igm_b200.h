/*
 * igm_b200.h — C ABI of libigm_b200.so: the B200 (sm_100a) implementation of the
 * DDPM U-Net hot path of Victarry/Image-Generation-models.
 *
 * The reference has no FFI (it is 100 % Python on torch ATen); the seam this ABI
 * attaches to is the Python class surface listed in SURVEY.md section 8(b).
 * Each entry point below names the reference method (file:line under
 * /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - tensors at the boundary are fp32, NCHW, contiguous; timesteps are int64
 *     (what torch hands the reference);
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *     all work is enqueued asynchronously on it;
 *   - every call returns 0 on success or a negative IGM_ERR_* code; the message
 *     is retrievable with igm_last_error(ctx) (ctx may be NULL for create errors);
 *   - the caller (PyTorch) owns parameters, gradients, Adam moments and all
 *     input/output tensors; the context owns packed weights, activations,
 *     workspaces and CUDA graphs.  No allocation happens after igm_unet_create.
 *   - one context per (process, device); not thread-safe.
 */
#ifndef IGM_B200_H
#define IGM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IGM_OK 0
#define IGM_ERR_INVALID (-1)   /* bad argument / unsupported shape            */
#define IGM_ERR_CUDA (-2)      /* a CUDA runtime / driver call failed         */
#define IGM_ERR_STATE (-3)     /* call order violated (e.g. backward w/o fwd) */
#define IGM_ERR_NOMEM (-4)

#define IGM_MAX_MULTS 8

typedef struct igm_ctx igm_ctx;

/* Shape of reference Unet(dim, channels, dim_mults) — src/models/ddpm.py:170-236 —
 * plus the image size BaseModel reads from the datamodule config
 * (src/models/base.py:20-22) and the diffusion length (ddpm.py:302). */
typedef struct igm_unet_cfg {
  int32_t dim;                       /* hidden_dim, multiple of 32            */
  int32_t channels;                  /* image channels (1 or 3)               */
  int32_t n_mults;
  int32_t dim_mults[IGM_MAX_MULTS];
  int32_t height, width;
  int32_t max_batch;                 /* activations are sized for this batch  */
  int32_t timesteps;                 /* T of GaussianDiffusion                */
  int32_t loss_type;                 /* 1 = l1 (ddpm.py:453), 2 = l2 (:455)   */
  int32_t training;                  /* 1: keep activations + grad workspaces */
} igm_unet_cfg;

/* Names the 12 schedule buffers of GaussianDiffusion.__init__ (ddpm.py:325-350);
 * each points at T fp32 values (device memory, caller-owned, must outlive ctx). */
typedef struct igm_schedule {
  const float* sqrt_alphas_cumprod;            /* ddpm.py:330 */
  const float* sqrt_one_minus_alphas_cumprod;  /* ddpm.py:331 */
  const float* sqrt_recip_alphas_cumprod;      /* ddpm.py:333 */
  const float* sqrt_recipm1_alphas_cumprod;    /* ddpm.py:334 */
  const float* posterior_log_variance_clipped; /* ddpm.py:344 */
  const float* posterior_mean_coef1;           /* ddpm.py:345 */
  const float* posterior_mean_coef2;           /* ddpm.py:346 */
} igm_schedule;

/* ---- lifecycle ---------------------------------------------------------- */
int igm_version(void);
const char* igm_last_error(const igm_ctx* ctx);

/* Builds the layer plan of Unet.__init__ (ddpm.py:170-236) for `cfg` on CUDA
 * device `device`, allocating every activation / workspace buffer once. */
int igm_unet_create(igm_ctx** out, const igm_unet_cfg* cfg, int device);
void igm_unet_destroy(igm_ctx* ctx);

/* ---- parameters ---------------------------------------------------------- */
/* Number of parameter tensors and total fp32 elements, in the exact order and
 * shapes of reference Unet(...).state_dict() (the flat arena layout). */
int igm_unet_num_params(const igm_ctx* ctx);
int64_t igm_unet_param_elems(const igm_ctx* ctx);
/* name_buf receives e.g. "downs.0.0.block1.block.0.weight"; shape gets <=4 dims. */
int igm_unet_param_info(const igm_ctx* ctx, int index, char* name_buf, int name_cap,
                        int64_t* offset, int32_t* ndim, int64_t shape[4]);
/* Bind the caller-owned flat fp32 arenas (params, grads: igm_unet_param_elems
 * elements each; grads may be NULL for inference).  nn.Parameter tensors of the
 * host-side mirror are views into these arenas, so state_dict() keeps the
 * reference's keys and NCHW shapes. */
int igm_unet_bind_params(igm_ctx* ctx, float* params, float* grads);
/* Re-pack bound parameters into the kernels' tap-major layouts.  Must be called
 * after the parameters change (load_state_dict, optimizer step). */
int igm_unet_pack_weights(igm_ctx* ctx, void* stream);

/* ---- U-Net ---------------------------------------------------------------- */
/* Unet.forward(x, time) — ddpm.py:238-261.  x,out: [B,C,H,W] fp32; t: [B] int64. */
int igm_unet_forward(igm_ctx* ctx, const float* x, const int64_t* t, float* out, int B,
                     void* stream);
/* Autograd of the last igm_unet_forward: d_out [B,C,H,W] -> ACCUMULATES parameter
 * gradients into the bound grad arena (like torch autograd into .grad) and, if
 * d_x is not NULL, writes dL/dx. */
int igm_unet_backward(igm_ctx* ctx, const float* d_out, float* d_x, void* stream);

/* Data-parallel gradient exchange ("one ncclAllReduce over the flat fp32 gradient arena per
 * step, chunked and overlapped with backward"; the reference itself delegates this to Lightning's DDP wrapper around
 * DDPM, src/train.py:27 + trainer strategy).  igm_unet_grad_buckets returns the number of buckets and fills up to `cap`
 * element ranges [lo, hi) of the gradient arena, in the order the backward pass completes them: bucket 0 = ups.*,
 * mid_*, final_conv.*; then downs.(n-1) ... downs.1; last = time_mlp.* + downs.0.*.  igm_unet_bucket_wait makes `stream`
 * wait (cudaStreamWaitEvent, no host sync) until every kernel of the MOST RECENT backward pass that writes bucket k has
 * finished, so the caller can start that bucket's all-reduce on `stream` while the backward pass is still running. */
int igm_unet_grad_buckets(const igm_ctx* ctx, int64_t* lo, int64_t* hi, int cap);
int igm_unet_bucket_wait(igm_ctx* ctx, int k, void* stream);

/* ---- diffusion ------------------------------------------------------------ */
int igm_ddpm_set_schedule(igm_ctx* ctx, const igm_schedule* sched);
/* GaussianDiffusion.q_sample — ddpm.py:433-444. */
int igm_ddpm_q_sample(igm_ctx* ctx, const float* x_start, const int64_t* t, const float* noise,
                      float* out, int B, void* stream);
/* GaussianDiffusion.p_losses — ddpm.py:446-460: q_sample + Unet.forward + l1/l2
 * mean loss, fused.  loss_out: 1 fp32 on the device. */
int igm_ddpm_p_losses(igm_ctx* ctx, const float* x_start, const int64_t* t, const float* noise,
                      float* loss_out, int B, void* stream);
/* d(loss)/d(params) of the last igm_ddpm_p_losses, accumulated into the grad
 * arena, seeded with  scale * (d_loss ? *d_loss : 1)  where d_loss is the
 * upstream gradient of the scalar loss as a DEVICE scalar (what autograd hands
 * loss.backward(); reading it on the device avoids a host sync) and `scale` a
 * host-side factor (1/world_size for a data-parallel mean). */
int igm_ddpm_p_losses_backward(igm_ctx* ctx, const float* d_loss, float scale, void* stream);
/* GaussianDiffusion.p_sample — ddpm.py:378-397 — for n_steps consecutive steps
 * t = t_start, t_start-1, ... (p_sample_loop, ddpm.py:406-407, when
 * t_start = T-1 and n_steps = T).  img [B,C,H,W] is updated in place.
 *   noise != NULL: [n_steps,B,C,H,W] injected draws (parity mode);
 *   noise == NULL: Philox4x32-10 + Box-Muller draws keyed by (seed, step, element).
 * The step is captured once into a CUDA graph and replayed n_steps times with
 * no host work in between. */
int igm_ddpm_sample_loop(igm_ctx* ctx, float* img, const float* noise, uint64_t seed, int B,
                         int t_start, int n_steps, int clip_denoised, void* stream);

/* ---- optimiser ------------------------------------------------------------ */
/* torch.optim.Adam step as configured at ddpm.py:502-512 (no weight decay, no
 * amsgrad) over n contiguous fp32 elements; grads are multiplied by grad_scale
 * first (1/world_size after an all-reduce SUM).  `step` is 1-based. */
int igm_adam_step(igm_ctx* ctx, float* params, const float* grads, float* exp_avg,
                  float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                  int step, float grad_scale, void* stream);

/* dst[i] += src[i] * scale * (alpha ? *alpha : 1) over n fp32 elements (both 16-byte aligned; alpha is a DEVICE scalar).
 * Host use: DDPM.training_step (ddpm.py:495-500) enqueues forward AND backward before it reads the loss back
 * (loss.item(), :499), with the gradients going to a pending arena; autograd's later loss.backward() — the point where
 * the reference computes them — is then this one kernel: .grad += d_loss * pending. */
int igm_grad_axpy(igm_ctx* ctx, float* dst, const float* src, const float* alpha, float scale, int64_t n,
                  void* stream);

/* ---- VQ-VAE quantiser ------------------------------------------------------- */
/* VectorQuantizer.forward — src/models/vqvae.py:24-43.  z, quant: [N, D, h, w] fp32 NCHW (HW = h*w);
 * codebook: [K, D]; idx: [N*h*w] int64 = argmin_k ||z - e_k|| with the FIRST index on ties;
 * losses[0] = vq_loss = mse(z, q), losses[1] = commit_loss = beta * mse(z, q)  (device floats).
 * ws: igm_vq_workspace_floats(N, HW) floats of scratch.  Context-free; errors via igm_last_error(NULL). */
int igm_vq_workspace_floats(int N, int HW);
int igm_vq_forward(const float* z, const float* codebook, int64_t* idx, float* quant, float* losses,
                   int N, int D, int HW, int K, float beta, float* ws, void* stream);
/* Autograd of the above: dz = d_commit*beta*2(z-q)/n (may be NULL); d_codebook[idx] +=
 * d_vq*2(q-z)/n + d_quant (ACCUMULATED; d_quant / d_vq / d_commit may be NULL = zero). */
int igm_vq_backward(const float* z, const float* codebook, const int64_t* idx, const float* d_quant,
                    const float* d_vq, const float* d_commit, float beta, float* dz, float* d_codebook,
                    int N, int D, int HW, int K, void* stream);

/* ---- PixelCNN ---------------------------------------------------------------- */
/* Incremental raster walk of the gated PixelCNN (src/models/pixelcnn.py:85-195), one CTA per image.
 *   mode 0: PixelCNN.sample (:167-195) with inverse-CDF draws k = #{j : cdf_j <= u}; u from
 *           uniforms[H*W][N*C] or, when NULL, Philox4x32-10 keyed by (seed, pixel, image*C+channel);
 *   mode 1: greedy argmax decode (first index on ties);
 *   mode 2: teacher forced, nothing drawn: with logits != NULL this is PixelCNN.forward (:128-154).
 * img [N,C,H,W] in/out (pixels k/255, or 2k/255-1 when normalize); skip[H*W] = 1 keeps a given pixel
 * (:185); logits [N,256,C,H,W] optional; weights = the live taps of every conv as [K][N] matrices in
 * the order documented in csrc/pixelcnn.cu (igm_pixelcnn_weight_floats elements);
 * ws = igm_pixelcnn_workspace_floats(N,C,H,W,Hd) floats.  Context-free; errors via igm_last_error(NULL). */
int64_t igm_pixelcnn_weight_floats(int C, int Hd);
int64_t igm_pixelcnn_workspace_floats(int N, int C, int H, int W, int Hd);
/* cond (nullable): [11][2][N][2*Hd] class-conditioning pre-gate addends, per layer the vertical then the
 * horizontal gate: cat(cond_proj_*1(y), cond_proj_*2(y)) of pixelcnn.py:71,:79. */
int igm_pixelcnn_run(const float* weights, float* img, const float* uniforms, const uint8_t* skip,
                     const float* cond, float* logits, float* ws, uint64_t seed, int N, int C, int H, int W, int Hd,
                     int mode, int normalize, void* stream);

/* ---- generic operators (secondary models: VQ-VAE encoder/decoder, PixelCNN training) --------- */
/* NHWC fp32 tensors (= torch channels_last); context-free, errors via igm_last_error(NULL).  Stride-1 1x1 / 3x3 "same"
 * convolutions with 64-multiple channel counts run on the tcgen05 bf16x3 engine (operands split and weights re-packed
 * per call into ws; IGM_OPS_TC=0 disables), everything else on the fp32 CUDA-core engine.
 * ws: igm_conv2d_workspace_floats(<the call's geometry>) floats of scratch.
 * transposed = 0: nn.Conv2d (weight OIHW, dilation allowed); 1: nn.ConvTranspose2d (weight IOHW).
 * Used behind src/networks/vqvae.py:5-136 and src/models/pixelcnn.py:12-82,:128-165. */
int64_t igm_conv2d_workspace_floats(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h,
                                    int pad_w, int dil, int OH, int OW);
int igm_conv2d_forward(const float* x, const float* w, const float* bias, const float* residual, float* y,
                       int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h,
                       int pad_w, int dil, int transposed, int OH, int OW, float* ws, void* stream);
/* dx (nullable) = dL/dx; dw += dL/dw in the weight's own layout; db (nullable) += dL/dbias. */
int igm_conv2d_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                        int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h,
                        int pad_w, int dil, int transposed, int OH, int OW, float* ws, void* stream);
/* kind 0 ReLU, 1 ELU, 2 tanh(a)*sigmoid(b), 3 tanh(a)*tanh(b) (gates: x [M,2C] -> y [M,C]).
 * backward: ref = ReLU OUTPUT for kind 0, the pre-activation x otherwise. */
int igm_act_forward(int kind, const float* x, const float* cond, int64_t hw, float* y, int64_t M, int C,
                    void* stream);
int igm_act_backward(int kind, const float* ref, const float* cond, int64_t hw, const float* dy, float* dx,
                     float* dcond, int64_t M, int C, void* stream);
/* kind 0: y = a + b; 1: y = a + (b - a), the straight-through value of src/models/vqvae.py:103. */
int igm_ewise(int kind, const float* a, const float* b, float* y, int64_t n, void* stream);
/* F.mse_loss (src/models/vqvae.py:106): loss[1] (nullable); da (nullable) = d_loss[0] * 2 (a-b) / n. */
int igm_mse(const float* a, const float* b, int64_t n, float* loss, const float* d_loss, float* da,
            void* stream);
/* PixelCNN.calc_likelihood's F.cross_entropy (pixelcnn.py:163) on NHWC logits [M, 256*C] with
 * channel = cls*C + ch and int64 targets [M, C]: nll[M*C] and/or d_logits = d_nll * (softmax - onehot). */
int igm_ce256(const float* logits, const int64_t* target, float* nll, float* d_logits, const float* d_nll,
              int64_t M, int C, void* stream);

/* ---- introspection (tests / profiling) ------------------------------------ */
/* Copies the named intermediate of the last forward (same names as
 * oracle/ddpm_oracle.py taps, e.g. "downs.0.0.block1.conv") to dst as NCHW fp32.
 * Returns the element count, or a negative error.  dst may be NULL to query. */
int64_t igm_debug_read_tap(igm_ctx* ctx, const char* name, float* dst_nchw, int64_t cap,
                           void* stream);
/* CUDA-event profiler (bench.py's roofline pass): between start and stop every launch of
 * this context is bracketed by an event pair; stop synchronises the device and returns,
 * per kernel class, the launch count, summed device time and the ALGORITHMIC flops /
 * bytes of those launches.  Adds per-launch overhead: never enabled in a timed region. */
typedef struct igm_profile_entry {
  char name[32];
  int64_t launches;
  double ms;
  double flops;
  double bytes;
} igm_profile_entry;
int igm_profile_start(igm_ctx* ctx);
int igm_profile_stop(igm_ctx* ctx, igm_profile_entry* out, int cap);
/* Number of kernels the library launched on behalf of this context so far. */
int64_t igm_launch_count(const igm_ctx* ctx);
/* Number of kernels launched so far by the context-free entry points (igm_conv2d_*, igm_act_*, igm_ce256, igm_ewise,
 * igm_mse, igm_vq_*, igm_pixelcnn_run) of this process. */
int64_t igm_ops_launch_count(void);
/* Diagnosis of the PixelCNN engine (IGM_PCNN_PROF=1 in the environment): SM cycles summed over all CTAs since the last call,
 * out[0] = row pass (vertical stack), out[1] = per-pixel chain, out[2] = head + draw, out[3] = pixels, out[4..7] = the chain
 * split into input fill / horiz_conv GEMV / gate / conv1x1_2 GEMV + tail (eight values); resets the counters. */
int igm_debug_pixelcnn_prof(unsigned long long* out);
/* One stride-1 KxK (K = 1 or 3, pad (K-1)/2) convolution on NHWC fp32 tensors, for kernel-level
 * parity tests: mode 0 = forward  x[B,H,W,Cin] -> out[B,H,W,Cout] (+bias, +add);
 *               mode 1 = data gradient  x = d_out[B,H,W,Cout] -> out = d_in[B,H,W,Cin].
 * w is the PyTorch OIHW weight.  engine as in igm_set_conv_engine; engine 3 = the CTA-pair (cta_group::2) tcgen05
 * kernel (conv_tc2.cu; N % 128 == 0; comparison only: parity-green but slower than engine 1).  Synchronous. */
int igm_debug_conv(int engine, int mode, const float* x, const float* w_oihw, const float* bias,
                   const float* add, float* out, int B, int H, int W, int Cin, int Cout, int K,
                   void* stream);
/* Kernel-level timing of the same convolution on the tcgen05 engines (1 = per-tap conv_tc.cu, 3 = CTA pair
 * conv_tc2.cu): synthetic operands staged once, `warm` untimed then `iters` timed launches between CUDA
 * events on `stream`; gn != 0 adds the fused GroupNorm partial statistics.  Synchronous. */
int igm_debug_conv_bench(int engine, int mode, int B, int H, int W, int Cin, int Cout, int K, int gn, int warm,
                         int iters, float* ms_per_launch, void* stream);
/* Weight gradient of the same convolution: gw[Cout,Cin,K,K] (fp32, ACCUMULATED) += dY (x) X with
 * x[B,H,W,Cin], dy[B,H,W,Cout] NHWC.  variant is a bring-up switch of the tcgen05 path (use 0). */
int igm_debug_wgrad(int engine, int variant, const float* x, const float* dy, float* gw, int B, int H,
                    int W, int Cin, int Cout, int K, void* stream);
/* The stride-2 resampling convs (kind 0: Conv2d(C,C2,3,2,1) weight OIHW; kind 1:
 * ConvTranspose2d(C,C2,4,2,1) weight IOHW) on either engine, NHWC fp32:
 *   mode 0 forward x[B,H,W,C] -> out[B,OH,OW,C2] (+bias); mode 1 data gradient x = dY -> out = dX (+add);
 *   mode 2 weight gradient (x, aux = dY) -> out = dW in the weight's own layout, ACCUMULATED. */
int igm_debug_resample(int engine, int kind, int mode, const float* x, const float* aux, const float* w,
                       const float* bias, const float* add, float* out, int B, int H, int W, int C, int C2,
                       void* stream);
/* Which conv engine is active: 0 = SIMT fp32 implicit GEMM, 1 = tcgen05 bf16x3 (default when
 * the layer shapes allow it; environment IGM_CONV_ENGINE=0 forces the SIMT engine). */
int igm_set_conv_engine(igm_ctx* ctx, int engine);
int igm_get_conv_engine(const igm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* IGM_B200_H */
