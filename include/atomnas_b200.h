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
 *   epi_mode 1: g = D * act'(r_scale*r + r_shift); y = g; partials of sum(g), sum(g*xhat),
 *               xhat = (r - r_mean)*r_rstd            (ReLU backward + BN-backward reductions)
 *   epi_mode 2: y = D + r                             (residual gradient add; r may be NULL)
 * stats_part, when not NULL, receives an_pw_gemm_stats_rows(M,N) rows of [2][N] floats.
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
int an_wgrad_reduce(const float* partials, int splits, int Cbig, int Csmall, float* dst,
                    long long dst_sb, long long dst_ss, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATOMNAS_B200_H_ */
