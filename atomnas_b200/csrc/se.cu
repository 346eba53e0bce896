// Squeeze-and-Excitation for the fused atomic block (AtomNAS-A+/B+/C+, BASELINE config 4).
//
// Reference: SqueezeAndExcitation models/mobilenet_base.py:93-117, used by
// InvertedResidualChannelsFused.forward :256-267 on the concatenated, ACTIVATED depthwise output:
//     v = act(bn(d_raw));  s = mean_{H,W} v;  h = W1 s + b1;  a = act(h);  z = W2 a + b2;
//     gate = sigmoid(z);   out = v * gate            (out feeds the 1x1 projection)
// and its autograd backward (D = d loss / d out):
//     dgate[b,c] = sum_{hw} D*v;  dz = dgate*gate*(1-gate);  da = W2^T dz;  dh = da*act'(h);
//     ds = W1^T dh;   dv = D*gate + ds/HW;   dW2 = sum_b dz a^T, db2 = sum_b dz, dW1 = sum_b dh s^T, ...
//
// Layout: activations NHWC bf16 in the engine's branch-fused, 8-padded channel slots ([rows, C],
// pitch ld).  The two FC layers keep the reference's master layout (fp32, [mid, total] and
// [total, mid], `total` = sum of the true branch widths); `chmap[C]` maps a padded channel to its
// master index (-1 for padding, whose gate is 0).  Everything here is small (B x C vectors) or one
// streaming pass over the [rows, C] tensor: CUDA cores, fp32, deterministic reduction order.
#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

constexpr int kSeThreads = 256;

// pooled[b, c] = mean over HW of act(scale*x + shift); with D != NULL: sum over HW of D * act(...)
// grid (C/64 blocks of 64 channels, B); thread = 8 channels (one 16-byte load) x row lane (32 lanes),
// 4 rows in flight per thread; fixed-order cross-lane sum in shared memory.
__global__ void __launch_bounds__(256) se_pool_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                      const __nv_bfloat16* __restrict__ D, long long ldD,
                                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                                      int act, int HW, int C, float inv, float* __restrict__ out) {
  __shared__ float red[32][65];
  const int cgl = threadIdx.x & 7, rl = threadIdx.x >> 3;          // 8 channel groups x 32 row lanes
  const int ch = blockIdx.x * 64 + cgl * 8;
  const int b = blockIdx.y;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (ch < C) {
    float sc[8], sh[8];
    load_coef8(scale, ch, C, sc, 0.f);
    load_coef8(shift, ch, C, sh, 0.f);
    const __nv_bfloat16* xp = x + (static_cast<long long>(b) * HW) * ldx + ch;
    const __nv_bfloat16* dp = D ? D + (static_cast<long long>(b) * HW) * ldD + ch : nullptr;
    for (int r0 = rl; r0 < HW; r0 += 128) {
      uint4 v[4], dv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + 32 * k;
        v[k] = r < HW ? ldg_nc_v4(xp + static_cast<long long>(r) * ldx) : make_uint4(0, 0, 0, 0);
        dv[k] = (dp != nullptr && r < HW) ? ldg_nc_v4(dp + static_cast<long long>(r) * ldD) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (r0 + 32 * k < HW) {
          const uint32_t vi[4] = {v[k].x, v[k].y, v[k].z, v[k].w}, di[4] = {dv[k].x, dv[k].y, dv[k].z, dv[k].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float lo = act_fwd(fmaf(sc[2 * e], bf16_lo(vi[e]), sh[2 * e]), act);
            float hi = act_fwd(fmaf(sc[2 * e + 1], bf16_hi(vi[e]), sh[2 * e + 1]), act);
            if (dp != nullptr) {
              lo *= bf16_lo(di[e]);
              hi *= bf16_hi(di[e]);
            }
            acc[2 * e] += lo;
            acc[2 * e + 1] += hi;
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cgl * 8 + e] = acc[e];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int l = 0; l < 32; ++l) t += red[l][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < C) out[static_cast<long long>(b) * C + c] = t * inv;
  }
}

// One CTA per sample: h = W1 s + b1, a = act(h), z = W2 a + b2, gate = sigmoid(z).
__global__ void se_gate_kernel(const float* __restrict__ pooled, const float* __restrict__ W1,
                               const float* __restrict__ b1, const float* __restrict__ W2,
                               const float* __restrict__ b2, const int* __restrict__ chmap, int act, int C,
                               int total, int mid, float* __restrict__ hpre, float* __restrict__ gate) {
  extern __shared__ float sm[];
  float* s = sm;                 // [total]  (master order)
  float* a = sm + total;         // [mid]
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int m = chmap[c];
    if (m >= 0) s[m] = pooled[static_cast<long long>(b) * C + c];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int j = warp; j < mid; j += nw) {        // one warp per hidden unit: coalesced W1 row
    float t = 0.f;
    for (int m = lane; m < total; m += 32) t = fmaf(W1[static_cast<long long>(j) * total + m], s[m], t);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) {
      t += b1[j];
      hpre[static_cast<long long>(b) * mid + j] = t;
      a[j] = act_fwd(t, act);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int m = chmap[c];
    float g = 0.f;
    if (m >= 0) {
      float t = b2[m];
      for (int j = 0; j < mid; ++j) t = fmaf(W2[static_cast<long long>(m) * mid + j], a[j], t);
      g = 1.f / (1.f + __expf(-t));
    }
    gate[static_cast<long long>(b) * C + c] = g;
  }
}

// vg = act(scale*x + shift) * gate[sample]   (bf16, the projection GEMM's A operand)
__global__ void se_scale_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                const float* __restrict__ gate, int HW, long long M, int C,
                                __nv_bfloat16* __restrict__ out, long long ldo) {
  const int ncg = C / 8;
  const long long total = M * ncg;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = i / ncg;
    const int ch = static_cast<int>(i % ncg) * 8;
    const long long b = row / HW;
    float sc[8], sh[8], gt[8];
    load_coef8(scale, ch, C, sc, 1.f);
    load_coef8(shift, ch, C, sh, 0.f);
    load_coef8(gate + b * C, ch, C, gt, 0.f);
    const uint4 v = ldg_nc_v4(x + row * ldx + ch);
    const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float lo = act_fwd(fmaf(sc[2 * e], bf16_lo(vi[e]), sh[2 * e]), act) * gt[2 * e];
      const float hi = act_fwd(fmaf(sc[2 * e + 1], bf16_hi(vi[e]), sh[2 * e + 1]), act) * gt[2 * e + 1];
      o[e] = pack_bf16(lo, hi);
    }
    stg_v4(out + row * ldo + ch, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// One CTA per sample: dz = dgate*g*(1-g); da = W2^T dz; dh = da*act'(h); ds = W1^T dh.
__global__ void se_gate_bwd_kernel(const float* __restrict__ dgate, const float* __restrict__ gate,
                                   const float* __restrict__ hpre, const float* __restrict__ W1,
                                   const float* __restrict__ W2, const int* __restrict__ chmap, int act, int C,
                                   int total, int mid, float* __restrict__ dz, float* __restrict__ dh,
                                   float* __restrict__ ds) {
  extern __shared__ float sm[];
  float* z = sm;                 // [total] dz in master order
  float* h = sm + total;         // [mid]   dh
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int m = chmap[c];
    if (m >= 0) {
      const float g = gate[static_cast<long long>(b) * C + c];
      const float t = dgate[static_cast<long long>(b) * C + c] * g * (1.f - g);
      z[m] = t;
      dz[static_cast<long long>(b) * total + m] = t;
    }
  }
  __syncthreads();
  // da[j] = sum_m W2[m, j] * dz[m]: warps split m (fixed partition), lanes run along j (coalesced
  // rows of W2); per-warp partials are combined in warp order -> deterministic
  {
    float* part = sm + total + mid;                // [nwarps][mid]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j0 = 0; j0 < mid; j0 += 32) {
      const int j = j0 + lane;
      float t = 0.f;
      if (j < mid)
        for (int m = warp; m < total; m += nw) t = fmaf(W2[static_cast<long long>(m) * mid + j], z[m], t);
      if (j < mid) part[warp * mid + j] = t;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < mid; j += blockDim.x) {
      float t = 0.f;
      for (int wv = 0; wv < nw; ++wv) t += part[wv * mid + j];
      t *= act_bwd(hpre[static_cast<long long>(b) * mid + j], act);
      h[j] = t;
      dh[static_cast<long long>(b) * mid + j] = t;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int m = chmap[c];
    float t = 0.f;
    if (m >= 0)
      for (int j = 0; j < mid; ++j) t = fmaf(W1[static_cast<long long>(j) * total + m], h[j], t);
    ds[static_cast<long long>(b) * C + c] = t;
  }
}

// Weight gradients of the two FC layers, summed over the batch in fixed order (deterministic):
//   dW2[m, j] = sum_b dz[b, m] * act(h[b, j]);  db2[m] = sum_b dz[b, m]
//   dW1[j, m] = sum_b dh[b, j] * s[b, m];       db1[j] = sum_b dh[b, j]     (s in padded order via chmap^-1)
__global__ void se_wgrad_kernel(const float* __restrict__ dz, const float* __restrict__ dh,
                                const float* __restrict__ hpre, const float* __restrict__ pooled,
                                const int* __restrict__ inv_map, int act, int B, int C, int total, int mid,
                                float* __restrict__ dW1, float* __restrict__ db1, float* __restrict__ dW2,
                                float* __restrict__ db2) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long n = static_cast<long long>(total) * mid;
  if (i < n) {
    const int m = static_cast<int>(i / mid), j = static_cast<int>(i % mid);
    const int c = inv_map[m];
    float w2 = 0.f, w1 = 0.f;
    for (int b = 0; b < B; ++b) {
      w2 = fmaf(dz[static_cast<long long>(b) * total + m], act_fwd(hpre[static_cast<long long>(b) * mid + j], act), w2);
      w1 = fmaf(dh[static_cast<long long>(b) * mid + j], pooled[static_cast<long long>(b) * C + c], w1);
    }
    dW2[static_cast<long long>(m) * mid + j] = w2;
    dW1[static_cast<long long>(j) * total + m] = w1;
  }
  if (i < total) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dz[static_cast<long long>(b) * total + i];
    db2[i] = t;
  }
  if (i < mid) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dh[static_cast<long long>(b) * mid + i];
    db1[i] = t;
  }
}

// g = (D*gate + ds/HW) * act'(scale*x + shift), rounded to bf16 and stored; partials of sum(g) and
// sum(g*xhat): the activation / BN backward of the depthwise output under the SE product.
__global__ void se_bwd_apply_kernel(const __nv_bfloat16* __restrict__ D, long long ldD,
                                    const __nv_bfloat16* __restrict__ x, long long ldx,
                                    const float* __restrict__ scale, const float* __restrict__ shift,
                                    const float* __restrict__ mean, const float* __restrict__ rstd, int act,
                                    const float* __restrict__ gate, const float* __restrict__ ds, int HW,
                                    float inv_hw, __nv_bfloat16* gout, long long ldg, long long M, int C,
                                    int rows_per_cta, float* part) {
  __shared__ float red[kSeThreads * 16];
  const int ncg = C / 8;
  const int cgb = min(ncg, 32);
  const int lanes = kSeThreads / cgb;
  const int lcg = threadIdx.x % cgb, rl = threadIdx.x / cgb;
  const int cg = blockIdx.y * cgb + lcg;
  const int ch = cg * 8;
  float s1[8], s2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
  if (cg < ncg && rl < lanes) {
    float sc[8], sh[8], mu[8], rs[8];
    load_coef8(scale, ch, C, sc, 1.f);
    load_coef8(shift, ch, C, sh, 0.f);
    load_coef8(mean, ch, C, mu, 0.f);
    load_coef8(rstd, ch, C, rs, 1.f);
    const long long r0 = static_cast<long long>(blockIdx.x) * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    for (long long r = r0 + rl; r < r1; r += lanes) {
      const long long b = r / HW;
      float gt[8], dsv[8];
      load_coef8(gate + b * C, ch, C, gt, 0.f);
      load_coef8(ds + b * C, ch, C, dsv, 0.f);
      const uint4 dv = ldg_nc_v4(D + r * ldD + ch);
      const uint4 xv = ldg_nc_v4(x + r * ldx + ch);
      const uint32_t di[4] = {dv.x, dv.y, dv.z, dv.w}, xi[4] = {xv.x, xv.y, xv.z, xv.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = (e & 1) ? bf16_hi(di[e >> 1]) : bf16_lo(di[e >> 1]);
        const float xx = (e & 1) ? bf16_hi(xi[e >> 1]) : bf16_lo(xi[e >> 1]);
        const float dv_ = fmaf(d, gt[e], dsv[e] * inv_hw);
        const float gq = round_bf16(dv_ * act_bwd(fmaf(sc[e], xx, sh[e]), act));
        o[e] = gq;
        s1[e] += gq;
        s2[e] = fmaf(gq, (xx - mu[e]) * rs[e], s2[e]);
      }
      stg_v4(gout + r * ldg + ch,
             make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7])));
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    red[threadIdx.x * 16 + c] = s1[c];
    red[threadIdx.x * 16 + 8 + c] = s2[c];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < cgb * 16; o += kSeThreads) {
    const int l = o / 16, e = o % 16;
    float a = 0.f;
    for (int r = 0; r < lanes; ++r) a += red[(r * cgb + l) * 16 + e];
    const int c = (blockIdx.y * cgb + l) * 8 + (e & 7);
    if (c < C) part[(static_cast<long long>(blockIdx.x) * 2 + (e >> 3)) * C + c] = a;
  }
}

static int se_rows_per_cta(long long M, int C) {
  const int ncg = C / 8, cgb = ncg < 32 ? ncg : 32;
  const int ygrid = (ncg + cgb - 1) / cgb;
  long long want_ctas = (4LL * sm_count() + ygrid - 1) / ygrid;
  long long rpc = (M + want_ctas - 1) / want_ctas;
  if (rpc < 64) rpc = 64;
  return static_cast<int>(rpc);
}

}  // namespace an

using namespace an;

extern "C" int an_se_pool(const void* x, long long ldx, const void* D, long long ldD, const float* scale,
                          const float* shift, int act, int B, int HW, int C, int mean, float* out, void* stream) {
  AN_REQUIRE(x && scale && shift && out && B > 0 && HW > 0 && C > 0 && C % 8 == 0 && ldx % 8 == 0,
             "an_se_pool: bad arguments");
  dim3 grid((C + 63) / 64, B);
  se_pool_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(D), ldD, scale, shift,
      act, HW, C, mean ? 1.f / HW : 1.f, out);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_se_gate(const float* pooled, const float* W1, const float* b1, const float* W2, const float* b2,
                          const int* chmap, int act, int B, int C, int total, int mid, float* hpre, float* gate,
                          void* stream) {
  AN_REQUIRE(pooled && W1 && b1 && W2 && b2 && chmap && hpre && gate && B > 0 && total > 0 && mid > 0,
             "an_se_gate: bad arguments");
  const size_t smem = static_cast<size_t>(total + mid) * 4;
  AN_REQUIRE(smem <= 48 * 1024, "an_se_gate: %d + %d floats of shared memory", total, mid);
  se_gate_kernel<<<B, 256, smem, static_cast<cudaStream_t>(stream)>>>(pooled, W1, b1, W2, b2, chmap, act, C, total, mid,
                                                                      hpre, gate);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_se_scale(const void* x, long long ldx, const float* scale, const float* shift, int act,
                           const float* gate, int B, int HW, int C, void* out, long long ldo, void* stream) {
  AN_REQUIRE(x && scale && shift && gate && out && C % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0, "an_se_scale: bad arguments");
  const long long M = static_cast<long long>(B) * HW;
  const long long total = M * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 16LL * sm_count()) blocks = 16LL * sm_count();
  se_scale_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, scale, shift, act, gate, HW, M, C,
      reinterpret_cast<__nv_bfloat16*>(out), ldo);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_se_gate_bwd(const float* dgate, const float* gate, const float* hpre, const float* pooled,
                              const float* W1, const float* W2, const int* chmap, const int* inv_map, int act, int B,
                              int C, int total, int mid, float* dz, float* dh, float* ds, float* dW1, float* db1,
                              float* dW2, float* db2, void* stream) {
  AN_REQUIRE(dgate && gate && hpre && pooled && W1 && W2 && chmap && inv_map && dz && dh && ds && dW1 && db1 && dW2 &&
                 db2, "an_se_gate_bwd: bad arguments");
  const size_t smem = static_cast<size_t>(total + mid + 8 * mid) * 4;          // + per-warp partials of da
  AN_REQUIRE(smem <= 48 * 1024, "an_se_gate_bwd: %d + 9 x %d floats of shared memory", total, mid);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  se_gate_bwd_kernel<<<B, 256, smem, st>>>(dgate, gate, hpre, W1, W2, chmap, act, C, total, mid, dz, dh, ds);
  AN_CHECK_LAUNCH();
  const long long n = static_cast<long long>(total) * mid;
  se_wgrad_kernel<<<static_cast<int>((n + 127) / 128), 128, 0, st>>>(dz, dh, hpre, pooled, inv_map, act, B, C, total,
                                                                     mid, dW1, db1, dW2, db2);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_se_bwd_apply_rows(long long M, int C) {
  const int rpc = se_rows_per_cta(M, C);
  return static_cast<int>((M + rpc - 1) / rpc);
}

extern "C" int an_se_bwd_apply(const void* D, long long ldD, const void* x, long long ldx, const float* scale,
                               const float* shift, const float* mean, const float* rstd, int act, const float* gate,
                               const float* ds, int B, int HW, int C, void* g, long long ldg, float* stats_part,
                               void* stream) {
  AN_REQUIRE(D && x && g && stats_part && scale && shift && mean && rstd && gate && ds && C % 8 == 0 && ldD % 8 == 0 &&
                 ldx % 8 == 0 && ldg % 8 == 0, "an_se_bwd_apply: bad arguments");
  const long long M = static_cast<long long>(B) * HW;
  const int rpc = se_rows_per_cta(M, C);
  const int ncg = C / 8, cgb = ncg < 32 ? ncg : 32;
  dim3 grid(static_cast<unsigned>((M + rpc - 1) / rpc), (ncg + cgb - 1) / cgb);
  se_bwd_apply_kernel<<<grid, kSeThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(D), ldD, reinterpret_cast<const __nv_bfloat16*>(x), ldx, scale, shift,
      mean, rstd, act, gate, ds, HW, 1.f / HW, reinterpret_cast<__nv_bfloat16*>(g), ldg, M, C, rpc, stats_part);
  AN_CHECK_LAUNCH();
  return 0;
}
