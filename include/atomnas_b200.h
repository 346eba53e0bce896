/*
 * atomnas_b200 C ABI -- the drop-in boundary of the B200-native AtomNAS hot path.
 *
 * The reference (meijieru/AtomNAS) has no FFI: its operators are PyTorch library calls made
 * from nn.Module.forward / autograd.  Each entry point below replaces one such call site on the
 * supernet training step; the reference file:line it stands in for is cited beside it.  The
 * Python host layer (atomnas_b200/*.py) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named host_*; no ownership is transferred;
 *   - activations are NHWC bf16 (a [rows = B*H*W, channels] row-major matrix, row pitch `ld*`
 *     in elements, a multiple of 8); statistics, BN vectors, master weights, grads are fp32;
 *   - `stream` is a cudaStream_t passed as void*; all calls are stream-ordered and asynchronous;
 *   - return value 0 = ok, non-zero = error (an_last_error() gives the message); no exceptions
 *     cross the boundary; there is no CPU fallback -- without an sm_100 device calls fail.
 */
#ifndef ATOMNAS_B200_H_
#define ATOMNAS_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

/* activation ids == get_active_fn names, models/mobilenet_base.py:407-415 */
#define AN_ACT_NONE 0
#define AN_ACT_RELU 1
#define AN_ACT_RELU6 2
#define AN_ACT_SWISH 3

const char* an_last_error(void);
int an_version(void);
/* number of SMs of the current device (grid sizing of the persistent kernels) */
int an_sm_count(void);
/* kernels launched by this library so far in this process (bench.py "gpu_launches") */
long long an_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Pointwise (1x1) convolution as a tcgen05 GEMM:   D[M,N] = Aeff[M,K] * W[N,K]^T
 * Replaces nn.Conv2d(k=1) forward and its autograd dgrad:
 *   expand  conv  models/mobilenet_base.py:315-320, project conv :338 (+ python sum :378),
 *   head conv models/mobilenet_supernet.py:148-153, classifier nn.Linear :160-163.
 * A operand prologue (applied to the smem tile before the MMA, so BN/activation of the
 * PRODUCER layer never makes an HBM round trip -- nn.BatchNorm2d/ReLU at :142):
 *   a_mode 0: Aeff = x1
 *   a_mode 1: Aeff = act(a1[k]*x1 + a3[k])                 (BN-apply + activation)
 *   a_mode 2: Aeff = a1[k]*x1 + a2[k]*x2 + a3[k]           (BN-backward apply, two inputs)
 * Epilogue:
 *   epi_mode 0: y = D (+ bias[n]); optional per-column sum / sum-of-squares partials of the
 *               ROUNDED output (training-mode BN statistics, F.batch_norm)
 *   epi_mode 1: g = D * act'(r_scale*r + r_shift); y = g; partials of sum(g), sum(g*r)
 *               (activation backward + BN-backward reductions; an_bn_bwd_finalize(raw_second=1)
 *               centres the second moment)
 *   epi_mode 2: y = D + r                             (residual gradient add; r may be NULL)
 * stats_part, when not NULL, receives an_pw_gemm_stats_rows(M,N) rows of [2][N] floats.
 * bf16 outputs without bias leave through a staged epilogue (TMEM -> swizzled smem -> TMA store)
 * and their partials are computed by the tensor core from the staged bf16 tile (diag(Y^T Y), Y^T 1);
 * fp32 output / bias (the classifier) takes the direct register -> global epilogue.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x1; long long ldx1;
  const void* x2; long long ldx2;
  const void* w;  long long ldw;        /* bf16 [N, ldw], K contiguous */
  int M, N, K;
  int a_mode, a_act;
  const float* a1; const float* a2; const float* a3;   /* [K] */
  int epi_mode;
  void* y; long long ldy; int y_is_f32;
  const float* bias;                    /* [N] or NULL (epi_mode 0) */
  float* stats_part;                    /* [rows][2][N] or NULL */
  const void* r; long long ldr;         /* bf16 [M, ldr] (epi_mode 1, 2) */
  const float* r_scale; const float* r_shift; const float* r_mean; const float* r_rstd;
  int r_act;
} an_pw_gemm_args;

int an_pw_gemm(const an_pw_gemm_args* args, void* stream);
int an_pw_gemm_stats_rows(int M, int N);

/* ------------------------------------------------------------------------------------------
 * Pointwise weight gradient:  dW[cb, cs] = sum_m Peff[m, cb] * q[m, cs]
 * Replaces the autograd wgrad of the same nn.Conv2d(k=1) / nn.Linear call sites.
 * P is the "big" channel side (prologue modes as above, coefficients [Cbig]); q is plain.
 * Deterministic two-stage reduction: the kernel writes `splits` fp32 partial matrices
 * [splits][Cbig][Csmall]; an_wgrad_reduce sums them in fixed order into dst with strides
 * (dst[cb*dst_sb + cs*dst_ss]).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* p1; long long ldp1;
  const void* p2; long long ldp2;
  const void* q;  long long ldq;
  int Mrows, Cbig, Csmall;
  int p_mode, p_act;
  const float* a1; const float* a2; const float* a3;   /* [Cbig] */
  float* partials;                     /* [splits][Cbig][Csmall] */
  int splits;                          /* from an_pw_wgrad_splits */
} an_pw_wgrad_args;

int an_pw_wgrad_splits(int Mrows, int Cbig, int Csmall);
int an_pw_wgrad(const an_pw_wgrad_args* args, void* stream);
/* dst[(cb-row0)*dst_sb + cs*dst_ss] = sum_s partials[s][cb][cs], cb in [row0,row0+nrows), cs < ncols */
int an_wgrad_reduce(const float* partials, int splits, int Cbig, int Csmall, int row0, int nrows,
                    int ncols, float* dst, long long dst_sb, long long dst_ss, void* stream);
/* up to 4 row ranges [row0, row0+nrows) of one partial buffer, each to its own destination, in ONE launch */
typedef struct { int row0; int nrows; float* dst; long long dst_sb; long long dst_ss; } an_wg_seg;
int an_wgrad_reduce_multi(const float* partials, int splits, int Cbig, int Csmall, int ncols,
                          const an_wg_seg* segs, int nseg, void* stream);

/* ------------------------------------------------------------------------------------------
 * Depthwise k x k convolution (k = 3, 5, 7; stride 1, 2; pad (k-1)/2), NHWC bf16: CUDA cores
 * (FFMA2 row streaming) or, for the shapes where it measured faster, the warp-level tensor cores.
 * Replaces ConvBNReLU(h, h, k, stride, groups=h) models/mobilenet_base.py:330-336 (pad :133)
 * and its autograd backward.  Pointers are pre-offset to the branch's first channel; C is the
 * branch width, ld* the full row pitch of the (branch-fused) tensor.
 *   forward : y = dwconv(act(scale*x + shift)) (scale == NULL: x used as is), BN partials of y
 *   bwd_data: du = dwconv^T(g);   g = dd, the BN-backward-applied gradient (an_affine_act)
 *             ge = du * act'(e_scale*eraw + e_shift); partials sum(ge), sum(ge*xhat_e)
 *   bwd_weight: dW[c][kr][kc] = sum g * act(e_scale*eraw + e_shift) (master layout [C,1,k,k];
 *             e_scale == NULL: eraw is the already activated input)
 * w is tap-major fp32 [k*k][C] (written by an_refresh_weights).
 * The producer's BatchNorm + activation is applied ON LOAD by the forward and weight-gradient
 * kernels (scale / e_scale != NULL), so the activated operand is never materialised in HBM.
 * Row-streaming kernels: csrc/dw2.cu (forward, stride-1 data gradient, weight gradient) and
 * csrc/dw_stream.cu (stride-2 data gradient).  Tensor-core kernels: csrc/dw_mma.cu -- stride 1, at
 * least 8 images, k = 7 forward / data gradient on maps >= 14 wide, k = 7 weight gradient, k = 5 weight
 * gradient on maps 14 / 56 wide; there the activated operand and the taps are multiplied in bf16
 * (fp32 accumulation), everywhere else in fp32.  AN_DW_MMA=<digits> forces those kernel sizes onto
 * the tensor cores for every width ("0": never); the *_rows helpers follow the same rule.
 * ca / cb / cc must be NULL (the on-load BN-backward
 * transform of g was removed together with the tile kernels); draw is unused.  The `fused`
 * argument of the *_rows helpers is ignored (one kernel family serves both operand forms).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* x; long long ldx;
  void* y; long long ldy;
  const float* w;
  const float* scale; const float* shift;
  float* stats_part;
  int act;
  int B, H, W, C, k, stride;
} an_dw_fwd_args;

typedef struct {
  const void* g; long long ldg;         /* bf16 [B,Ho,Wo,*] grad wrt this layer's BN output */
  const void* draw; long long ldd;      /* bf16 raw conv output saved by forward */
  const void* eraw; long long lde;      /* bf16 [B,H,W,*] raw input (producer conv output) */
  void* ge; long long ldge;             /* out (bwd_data) */
  const float* w;
  const float* ca; const float* cb; const float* cc;
  const float* e_scale; const float* e_shift; const float* e_mean; const float* e_rstd;
  float* stats_part;
  int act;
  int B, H, W, C, k, stride;
} an_dw_bwd_args;

int an_dw_fwd(const an_dw_fwd_args* args, void* stream);
int an_dw_fwd_stats_rows(int B, int C, int Ho, int Wo, int k, int stride, int fused);
int an_dw_bwd_data(const an_dw_bwd_args* args, void* stream);
int an_dw_bwd_data_stats_rows(int B, int C, int H, int W, int k, int stride, int fused);
int an_dw_bwd_weight(const an_dw_bwd_args* args, float* partials, float* dw, void* stream);
int an_dw_bwd_weight_rows(int B, int C, int H, int W, int k, int stride, int fused);
/* 1 if the (k, stride, batch, map width W of the stride-1 layer) shape runs on the tensor cores:
 * which = 0 forward / data gradient (bf16 operand + taps -- the oracle's bf16 points follow this rule),
 * which = 1 weight gradient. */
int an_dw_on_tensor_cores(int which, int k, int stride, int B, int W);

/* ------------------------------------------------------------------------------------------
 * BatchNorm2d (training mode) pieces -- nn.BatchNorm2d at models/mobilenet_base.py:142, :342.
 * The conv kernels produce per-CTA partials [rows][2][C]; these combine them.
 *   an_bn_finalize     : mean/rstd (saved for backward), scale = gamma*rstd, shift = beta-mean*scale,
 *                        running stats update (momentum m; calibration passes m = 1/nbt)
 *   an_bn_bwd_finalize : dgamma = sum g*xhat, dbeta = sum g, and (ca,cb,cc) with dx = ca*g+cb*x+cc
 *                        (raw_second: the partials hold sum g*x -- an_pw_gemm epi_mode 1)
 *   an_affine_act      : out = act(a1*x1 + a2*x2 + a3) (NULL coefficient = 1, 1, 0; x2 optional):
 *                        pw_bn apply + residual (:379-381), BN-backward apply
 *   an_col_reduce      : partials of sum g, sum g*xhat over a [M,C] matrix (pw_bn backward)
 * ------------------------------------------------------------------------------------------ */
int an_bn_finalize(const float* stats_part, int rows, int C, double count, const float* gamma,
                   const float* beta, float* running_mean, float* running_var, float momentum,
                   float eps, float* mean, float* rstd, float* scale, float* shift, void* stream);
int an_bn_bwd_finalize(const float* stats_part, int rows, int C, double count, const float* gamma,
                       const float* mean, const float* rstd, float* dgamma, float* dbeta, float* ca,
                       float* cb, float* cc, int raw_second, void* stream);

/* The same two finalizes for up to 4 independent partial buffers in ONE launch (the three depthwise
 * branches of a block reduce their statistics with different partial-row counts): segment s holds
 * [rows][2][C] partials of channels [c0, c0 + C) of the vectors passed (gamma ... shift are the
 * branch-fused vectors' base pointers). */
typedef struct { const float* part; int rows; int c0; int C; } an_bn_seg;
int an_bn_finalize_multi(const an_bn_seg* segs, int nseg, double count, const float* gamma,
                         const float* beta, float* running_mean, float* running_var, float momentum,
                         float eps, float* mean, float* rstd, float* scale, float* shift, void* stream);
int an_bn_bwd_finalize_multi(const an_bn_seg* segs, int nseg, double count, const float* gamma,
                             const float* mean, const float* rstd, float* dgamma, float* dbeta, float* ca,
                             float* cb, float* cc, int raw_second, void* stream);
int an_affine_act(const void* x1, long long ld1, const void* x2, long long ld2, const float* a1,
                  const float* a2, const float* a3, int act, void* out, long long ldo, long long M,
                  int C, void* stream);
int an_bn_eval_coeffs(const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, float eps, int C, float* scale, float* shift,
                      void* stream);
/* g = dy*act'(scale*x+shift) + partials of sum g, sum g*xhat (same row count as an_col_reduce) */
int an_act_bwd_reduce(const void* dy, long long ldy, const void* x, long long ldx, const float* scale,
                      const float* shift, const float* mean, const float* rstd, int act, void* g,
                      long long ldg, long long M, int C, float* stats_part, void* stream);
int an_sum_rows(const float* part, int rows, long long row_stride, long long n, float* dst,
                void* stream);
int an_col_reduce_rows(long long M, int C);
int an_col_reduce(const void* g, long long ldg, const void* x, long long ldx, const float* mean,
                  const float* rstd, long long M, int C, float* stats_part, void* stream);

/* ------------------------------------------------------------------------------------------
 * Network ends: stem conv 3x3/s2 on the fp32 NCHW input (models/mobilenet_supernet.py:126-132),
 * head BN+act fused into the global average pool (+ dropout) (:154-161), label-smoothed cross
 * entropy (utils/optim.py:199-207; loss per sample, dlogits scaled by grad_scale).
 * ------------------------------------------------------------------------------------------ */
int an_stem_fwd(const float* x, const float* w, void* y, long long ldy, float* stats_part, int B,
                int H, int W, int C, void* stream);
int an_stem_fwd_stats_rows(int B, int Ho, int Wo);
int an_stem_wgrad(const float* x, const void* g, long long ldg, const void* sraw, long long lds,
                  const float* ca, const float* cb, const float* cc, float* partials, float* dw,
                  int B, int H, int W, int C, void* stream);
int an_stem_wgrad_rows(int B, int Ho, int Wo);
int an_pool_fwd(const void* h, long long ldh, const float* scale, const float* shift, int act,
                void* pooled, long long ldp, int B, int HW, int C, float drop_p,
                unsigned long long seed, const unsigned long long* seed_dev, void* stream);
int an_pool_bwd(const void* dpooled, long long ldd, const void* h, long long ldh, const float* scale,
                const float* shift, const float* mean, const float* rstd, int act, void* gh,
                long long ldg, float* stats_part, int B, int HW, int C, float drop_p,
                unsigned long long seed, const unsigned long long* seed_dev, void* stream);
int an_pool_bwd_stats_rows(int B);
int an_ce_loss(const float* logits, long long ldl, const long long* labels, int B, int NC,
               float label_smoothing, float grad_scale, float* loss, void* dlogits, long long ldd,
               int nd_pad, void* stream);
int an_f32_to_bf16(const float* src, long long lds, void* dst, long long ldd, long long rows,
                   int cols, int cols_pad, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused optimizer over flat arenas (wd_flag / l1_pen are per-element arrays parallel to the arena;
 * wd_flag < 0 marks frozen padding elements):
 *   g' = g*grad_scale + weight_decay*wd_flag*p + rho*l1_pen*sign(p)
 *   RMSprop (utils/rmsprop.py:100-130), then ema = d*ema + (1-d)*p (utils/optim.py:64-65).
 * hparams_dev (optional, device) = {lr, rho, ema_decay}: per-step values read on the device so a
 * captured CUDA graph can be replayed while the schedules advance; seed_dev likewise offsets the
 * dropout seed (effective seed = seed + *seed_dev).
 * an_refresh_weights re-materialises the kernel-layout weight copies (bf16 [N,K], bf16 [K,N]
 * transposes for dgrad, fp32 tap-major depthwise) from the fp32 master arena via a table of
 * 9 x int64 entries {src_ld, src_off, dst_off, rows, cols, dst_ld, transpose, dst_f32, first_tile}.
 * an_gather_rows is the bit-exact `rhs[mask]` of models/compress_utils.py:31-37.
 * ------------------------------------------------------------------------------------------ */
int an_rmsprop_ema_step(float* p, const float* g, float* sq, float* buf, float* ema,
                        const float* wd_flag, const float* l1_pen, long long n, float lr,
                        float alpha, float momentum, float eps, int eps_inside_sqrt,
                        float grad_scale, float weight_decay, float rho, float ema_decay,
                        const float* hparams_dev, void* stream);
int an_ema_update(float* ema, const float* x, long long n, float decay, const float* hparams_dev,
                  void* stream);
int an_reg_losses(const float* p, const float* wd_flag, const float* l1_pen, long long n,
                  float* part, void* stream);
int an_refresh_weights(const void* table, int n_entries, long long total_tiles, const float* src,
                       void* dst_bf16, float* dst_f32, void* stream);
int an_gather_rows(const float* src, const long long* idx, long long n_out, long long inner,
                   float* dst, void* stream);

/* ------------------------------------------------------------------------------------------
 * Squeeze-and-Excitation of the fused atomic block (AtomNAS-A+/B+/C+):
 * SqueezeAndExcitation models/mobilenet_base.py:93-117 as used by
 * InvertedResidualChannelsFused.forward :256-267, and its autograd backward.
 *   v = act(scale*x + shift) (x = raw depthwise output, branch-fused NHWC bf16 [B*HW, C]);
 *   an_se_pool   : out[b,c] = mean_hw v           (D == NULL, mean = 1)
 *                  out[b,c] = sum_hw D*v          (D != NULL, mean = 0: d loss / d gate)
 *   an_se_gate   : h = W1 s + b1, a = act(h), z = W2 a + b2, gate = sigmoid(z); hpre = h is kept
 *   an_se_scale  : out = v * gate[sample]         (the projection GEMM's A operand)
 *   an_se_gate_bwd: dz, dh, ds (= d loss / d pooled) per sample, and the FC weight / bias gradients
 *                  summed over the batch in fixed order
 *   an_se_bwd_apply: g = (D*gate + ds/HW) * act'(scale*x + shift) -> bf16, partials sum(g), sum(g*xhat)
 * W1 [mid, total], W2 [total, mid], b1 [mid], b2 [total]: fp32 master layout (nn.Conv2d 1x1 with
 * bias); chmap[C] = master channel of a padded channel or -1, inv_map[total] its inverse.
 * ------------------------------------------------------------------------------------------ */
int an_se_pool(const void* x, long long ldx, const void* D, long long ldD, const float* scale,
               const float* shift, int act, int B, int HW, int C, int mean, float* out, void* stream);
int an_se_gate(const float* pooled, const float* W1, const float* b1, const float* W2, const float* b2,
               const int* chmap, int act, int B, int C, int total, int mid, float* hpre, float* gate,
               void* stream);
int an_se_scale(const void* x, long long ldx, const float* scale, const float* shift, int act,
                const float* gate, int B, int HW, int C, void* out, long long ldo, void* stream);
int an_se_gate_bwd(const float* dgate, const float* gate, const float* hpre, const float* pooled,
                   const float* W1, const float* W2, const int* chmap, const int* inv_map, int act, int B,
                   int C, int total, int mid, float* dz, float* dh, float* ds, float* dW1, float* db1,
                   float* dW2, float* db2, void* stream);
int an_se_bwd_apply_rows(long long M, int C);
int an_se_bwd_apply(const void* D, long long ldD, const void* x, long long ldx, const float* scale,
                    const float* shift, const float* mean, const float* rstd, int act, const float* gate,
                    const float* ds, int B, int HW, int C, void* g, long long ldg, float* stats_part,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATOMNAS_B200_H_ */
