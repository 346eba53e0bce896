// Fused multi-tensor optimizer step over flat arenas, and the refresh of the GEMM-ready weight
// copies.
//
// Reference: utils/rmsprop.py:70-132 (TF-style RMSprop, eps inside the sqrt), the EMA update
// utils/optim.py:54-65 driven by train.py:231-236, the analytic gradients of cal_l2_loss
// (utils/optim.py:226-249: wd*w for conv/fc weights and the classifier bias) and cal_bn_l1_loss
// (utils/prune.py:161-167: rho*penalty*sign(gamma)), and the 1/world_size of
// utils/distributed.py:136.  The reference runs ~6 tiny ATen kernels per parameter tensor
// (x497) plus 799 EMA launches; here it is one pass over the arena.
#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

// Arena convention: wd_flag / l1_pen are per-element arrays parallel to the arena (wd_flag: 1 =
// L2-regularised, 0 = not, < 0 = frozen padding element that must never move); n % 4 == 0.
__global__ void rmsprop_ema_kernel(float* __restrict__ p, const float* __restrict__ g,
                                   float* __restrict__ sq, float* __restrict__ buf,
                                   float* __restrict__ ema, const float* __restrict__ wd_flag,
                                   const float* __restrict__ l1_pen, long long n, float lr,
                                   float alpha, float momentum, float eps, int eps_inside_sqrt,
                                   float grad_scale, float weight_decay, float rho,
                                   float ema_decay, const float* __restrict__ hp) {
  if (hp != nullptr) {          // per-step values from device memory (CUDA-graph replay)
    lr = hp[0];
    rho = hp[1];
    ema_decay = hp[2];
  }
  const long long i4 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) * 4;
  if (i4 >= n) return;
  const float4 wf = *reinterpret_cast<const float4*>(wd_flag + i4);
  const float4 lp = *reinterpret_cast<const float4*>(l1_pen + i4);
  const float wfl[4] = {wf.x, wf.y, wf.z, wf.w}, lpe[4] = {lp.x, lp.y, lp.z, lp.w};
  float4 pv = *reinterpret_cast<float4*>(p + i4);
  const float4 gv = *reinterpret_cast<const float4*>(g + i4);
  float4 sv = *reinterpret_cast<float4*>(sq + i4);
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (momentum > 0.f) bv = *reinterpret_cast<float4*>(buf + i4);
  float pp[4] = {pv.x, pv.y, pv.z, pv.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
  float ss[4] = {sv.x, sv.y, sv.z, sv.w}, bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (wfl[k] < 0.f) continue;
    const float sgn = static_cast<float>(pp[k] > 0.f) - static_cast<float>(pp[k] < 0.f);
    const float gr = gg[k] * grad_scale + weight_decay * wfl[k] * pp[k] + rho * lpe[k] * sgn;
    ss[k] = ss[k] * alpha + (1.f - alpha) * gr * gr;
    const float avg = eps_inside_sqrt ? sqrtf(ss[k] + eps) : sqrtf(ss[k]) + eps;
    if (momentum > 0.f) {
      bb[k] = bb[k] * momentum + gr / avg;
      pp[k] -= lr * bb[k];
    } else {
      pp[k] -= lr * (gr / avg);
    }
  }
  *reinterpret_cast<float4*>(p + i4) = make_float4(pp[0], pp[1], pp[2], pp[3]);
  *reinterpret_cast<float4*>(sq + i4) = make_float4(ss[0], ss[1], ss[2], ss[3]);
  if (momentum > 0.f) *reinterpret_cast<float4*>(buf + i4) = make_float4(bb[0], bb[1], bb[2], bb[3]);
  if (ema != nullptr) {
    float4 ev = *reinterpret_cast<float4*>(ema + i4);
    ev.x = ev.x * ema_decay + (1.f - ema_decay) * pp[0];
    ev.y = ev.y * ema_decay + (1.f - ema_decay) * pp[1];
    ev.z = ev.z * ema_decay + (1.f - ema_decay) * pp[2];
    ev.w = ev.w * ema_decay + (1.f - ema_decay) * pp[3];
    *reinterpret_cast<float4*>(ema + i4) = ev;
  }
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ x, long long n, float decay,
                           const float* __restrict__ hp) {
  if (hp != nullptr) decay = hp[2];
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i < n) ema[i] = ema[i] * decay + (1.f - decay) * x[i];
}

// Regularisation loss values (logging only): out[0] = 0.5*wd*sum(flag*p^2), out[1] = rho*sum(pen*|p|)
__global__ void reg_loss_kernel(const float* __restrict__ p, const float* __restrict__ wd_flag,
                                const float* __restrict__ l1_pen, long long n, float* part) {
  __shared__ double r0[256], r1[256];
  double a = 0.0, b = 0.0;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    const float v = p[i];
    a += static_cast<double>(fmaxf(wd_flag[i], 0.f)) * v * v;
    b += static_cast<double>(l1_pen[i]) * fabsf(v);
  }
  r0[threadIdx.x] = a;
  r1[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { r0[threadIdx.x] += r0[threadIdx.x + s]; r1[threadIdx.x] += r1[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[blockIdx.x * 2] = static_cast<float>(r0[0]); part[blockIdx.x * 2 + 1] = static_cast<float>(r1[0]); }
}

// ------------------------------------------------------------------------------------------------
// weight refresh: table-driven strided 2-D copy/cast/transposes, master fp32 -> kernel layouts
// ------------------------------------------------------------------------------------------------
struct RefreshEntry {     // mirrored by atomnas_b200/engine.py (9 x int64)
  long long src_ld;       // source pitch (elements)
  long long src_off;      // elements into the fp32 source arena; matrix is [rows, cols], pitch src_ld
  long long dst_off;      // elements into the destination arena (bf16 or fp32)
  long long rows, cols;
  long long dst_ld;       // destination pitch (elements)
  long long transpose;    // 0: dst[r*ld + c], 1: dst[c*ld + r]
  long long dst_f32;      // 0: bf16 arena, 1: fp32 arena
  long long blk0;         // first 32x32 tile index of this entry
};

__global__ void refresh_kernel(const RefreshEntry* __restrict__ table, int n_entries,
                               const float* __restrict__ src, __nv_bfloat16* dst16, float* dst32) {
  __shared__ float tile[32][33];
  int lo = 0, hi = n_entries - 1;
  const long long blk = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (table[mid].blk0 <= blk) lo = mid; else hi = mid - 1;
  }
  const RefreshEntry e = table[lo];
  const long long tiles_c = (e.cols + 31) / 32;
  const long long t = blk - e.blk0;
  const long long r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int j = ty; j < 32; j += 8) {
    const long long r = r0 + j, c = c0 + tx;
    tile[j][tx] = (r < e.rows && c < e.cols) ? src[e.src_off + r * e.src_ld + c] : 0.f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    long long r, c;
    float v;
    if (e.transpose) { c = c0 + j; r = r0 + tx; v = tile[tx][j]; }
    else { r = r0 + j; c = c0 + tx; v = tile[j][tx]; }
    if (r < e.rows && c < e.cols) {
      const long long o = e.dst_off + (e.transpose ? c * e.dst_ld + r : r * e.dst_ld + c);
      if (e.dst_f32) dst32[o] = v; else dst16[o] = __float2bfloat16_rn(v);
    }
  }
}

// dst[j] = src[idx[j]] along dim 0 (rows of `inner` elements) -- the `rhs[mask]` of
// models/compress_utils.py:31-37 with idx = nonzero(mask); dim-1 masking is the same call on a
// transposed view done by the host.  Bit-exact copy.
__global__ void gather_rows_kernel(const float* __restrict__ src, const long long* __restrict__ idx,
                                   long long n_out, long long inner, float* dst) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n_out * inner) return;
  const long long r = i / inner, c = i % inner;
  dst[i] = src[idx[r] * inner + c];
}

}  // namespace an

using namespace an;

extern "C" int an_rmsprop_ema_step(float* p, const float* g, float* sq, float* buf, float* ema,
                                   const float* wd_flag, const float* l1_pen, long long n, float lr,
                                   float alpha, float momentum, float eps, int eps_inside_sqrt,
                                   float grad_scale, float weight_decay, float rho, float ema_decay,
                                   const float* hparams_dev, void* stream) {
  AN_REQUIRE(p && g && sq && wd_flag && l1_pen && n > 0 && n % 4 == 0, "an_rmsprop_ema_step: arena must be a multiple of 4 elements");
  AN_REQUIRE(momentum <= 0.f || buf, "an_rmsprop_ema_step: momentum needs a buffer");
  const long long threads = n / 4;
  rmsprop_ema_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, sq, buf, ema, wd_flag, l1_pen, n, lr, alpha, momentum, eps, eps_inside_sqrt, grad_scale, weight_decay,
      rho, ema_decay, hparams_dev);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_ema_update(float* ema, const float* x, long long n, float decay, const float* hparams_dev,
                             void* stream) {
  AN_REQUIRE(ema && x && n > 0, "an_ema_update: bad arguments");
  ema_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(ema, x, n, decay, hparams_dev);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_reg_losses(const float* p, const float* wd_flag, const float* l1_pen, long long n,
                             float* part /* [128][2] */, void* stream) {
  AN_REQUIRE(p && wd_flag && l1_pen && part && n > 0, "an_reg_losses: bad arguments");
  reg_loss_kernel<<<128, 256, 0, static_cast<cudaStream_t>(stream)>>>(p, wd_flag, l1_pen, n, part);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_refresh_weights(const void* table, int n_entries, long long total_tiles,
                                  const float* src, void* dst_bf16, float* dst_f32, void* stream) {
  AN_REQUIRE(table && n_entries > 0 && total_tiles > 0 && src, "an_refresh_weights: bad arguments");
  refresh_kernel<<<static_cast<unsigned>(total_tiles), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const RefreshEntry*>(table), n_entries, src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), dst_f32);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_gather_rows(const float* src, const long long* idx, long long n_out, long long inner,
                              float* dst, void* stream) {
  AN_REQUIRE(src && idx && dst && n_out >= 0 && inner > 0, "an_gather_rows: bad arguments");
  if (n_out == 0) return 0;
  const long long n = n_out * inner;
  gather_rows_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, idx, n_out, inner, dst);
  AN_CHECK_LAUNCH();
  return 0;
}
