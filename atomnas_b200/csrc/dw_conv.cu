// Depthwise k x k convolution (k in {3,5,7}, stride 1/2, pad (k-1)/2) on NHWC bf16 activations,
// CUDA cores only (it is not a dense contraction): forward, input gradient and weight gradient.
//
// Reference call site: ConvBNReLU(h, h, k, stride, groups=h) -- models/mobilenet_base.py:330-336,
// padding rule :133 -- executed there by nn.Conv2d / cuDNN, followed by nn.BatchNorm2d + act.
//
// Kernel pattern: the PRODUCER's BatchNorm-apply + activation is applied to the
// input while it is staged into shared memory (one pass, 16-byte vectors, 8 channels per
// thread), the k x k taps run out of registers/shared memory, and this layer's BN statistics
// (sum, sum of squares of the bf16-rounded output) are reduced in the epilogue into per-CTA
// partials (deterministic, no float atomics).
#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

struct DwFwdParams {
  const __nv_bfloat16* x;     // [B, H, W, ldx] (+ channel offset applied)
  __nv_bfloat16* y;           // [B, Ho, Wo, ldy]
  const float* w;             // [k*k][C] tap-major
  const float* scale;         // [C] producer BN scale (gamma * rstd)   -- nullptr: identity
  const float* shift;         // [C]
  float* stats_part;          // [parts][2][C]
  int act;
  int B, H, W, C, Ho, Wo;
  long long ldx, ldy;
  int ncg, tr, tc;            // channel groups (8 ch each), tile rows, tile col-strips (8 cols)
  int tiles_x, tiles_y, ipb;  // images per CTA
};

template <int K, int S>
__global__ void dw_fwd_kernel(const DwFwdParams p) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int P = (K - 1) / 2;
  const int ncg = p.ncg, tr = p.tr, tc = p.tc;
  const int rows_in = (tr - 1) * S + K;
  const int cols_in = (tc * 8 - 1) * S + K;
  uint4* patch = reinterpret_cast<uint4*>(dsm);                                  // [rows_in][cols_in][ncg]
  float* wsm = reinterpret_cast<float*>(dsm + (size_t)rows_in * cols_in * ncg * 16);   // [K*K][ncg*8]
  float* red = wsm + K * K * ncg * 8;                                            // [nthreads][16]

  const int nthreads = blockDim.x;
  const int tid = threadIdx.x;
  const int cg = tid % ncg;
  const int strip = (tid / ncg) % tc;
  const int trow = tid / (ncg * tc);
  const int cb0 = blockIdx.y * ncg * 8;
  const int c0 = cb0 + cg * 8;
  const bool cg_ok = c0 < p.C;
  const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x % p.tiles_x;
  const int oy0 = tile_y * tr, ox0 = tile_x * tc * 8;
  const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;

  // weights -> smem (zero for channels beyond C)
  for (int i = tid; i < K * K * ncg * 8; i += nthreads) {
    const int tap = i / (ncg * 8), c = cb0 + i % (ncg * 8);
    const int e8 = i % 8, lcg8 = (i % (ncg * 8)) / 8;      // [tap][half][cg][4]: 16-byte lane stride
    wsm[((tap * 2 + e8 / 4) * ncg + lcg8) * 4 + (e8 & 3)] = (c < p.C) ? p.w[(long long)tap * p.C + c] : 0.f;
  }
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc[i] = 1.f; sh[i] = 0.f; }

  float st1[8], st2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { st1[i] = 0.f; st2[i] = 0.f; }

  const int img0 = blockIdx.z * p.ipb;
  for (int ii = 0; ii < p.ipb; ++ii) {
    const int n = img0 + ii;
    if (n >= p.B) break;
    __syncthreads();     // previous image's compute done before the patch is overwritten
    // ---- stage the input patch with BN-apply + activation
    const int nchunks = rows_in * cols_in * ncg;
    for (int i = tid; i < nchunks; i += nthreads) {
      const int lcg = i % ncg;
      const int px = (i / ncg) % cols_in;
      const int py = i / (ncg * cols_in);
      const int iy = iy0 + py, ix = ix0 + px;
      const int ch = cb0 + lcg * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && ch < p.C) {
        v = ldg_nc_v4(p.x + (((long long)n * p.H + iy) * p.W + ix) * p.ldx + ch);
        if (p.scale != nullptr) {
          float a[8], b[8];
          load_coef8(p.scale, ch, p.C, a, 1.f);
          load_coef8(p.shift, ch, p.C, b, 0.f);
          uint32_t vi[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = act_fwd(fmaf(a[2 * e], bf16_lo(vi[e]), b[2 * e]), p.act);
            const float hi = act_fwd(fmaf(a[2 * e + 1], bf16_hi(vi[e]), b[2 * e + 1]), p.act);
            vi[e] = pack_bf16(lo, hi);
          }
          v = make_uint4(vi[0], vi[1], vi[2], vi[3]);
        }
      }
      patch[i] = v;
    }
    __syncthreads();

    // ---- compute 8 outputs x 8 channels per thread
    float acc[8][8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

    const int oy = oy0 + trow;
    if (cg_ok && oy < p.Ho) {
#pragma unroll 1
      for (int kr = 0; kr < K; ++kr) {
        float wr[K][8];
#pragma unroll
        for (int kc = 0; kc < K; ++kc) {
          const float4 lo = *reinterpret_cast<const float4*>(wsm + (((kr * K + kc) * 2 + 0) * ncg + cg) * 4);
          const float4 hi = *reinterpret_cast<const float4*>(wsm + (((kr * K + kc) * 2 + 1) * ncg + cg) * 4);
          wr[kc][0] = lo.x; wr[kc][1] = lo.y; wr[kc][2] = lo.z; wr[kc][3] = lo.w;
          wr[kc][4] = hi.x; wr[kc][5] = hi.y; wr[kc][6] = hi.z; wr[kc][7] = hi.w;
        }
        const uint4* prow = patch + ((size_t)(trow * S + kr) * cols_in + strip * 8 * S) * ncg + cg;
#pragma unroll
        for (int ic = 0; ic < 7 * S + K; ++ic) {
          const uint4 v = prow[(size_t)ic * ncg];
          const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                              bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kc = ic - j * S;
            if (kc >= 0 && kc < K) {
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(wr[kc][c], f[c], acc[j][c]);
            }
          }
        }
      }
      // ---- store + statistics of the rounded values
      __nv_bfloat16* yrow = p.y + (((long long)n * p.Ho + oy) * p.Wo) * p.ldy + c0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ox = ox0 + strip * 8 + j;
        if (ox < p.Wo) {
          float r[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            r[c] = round_bf16(acc[j][c]);
            st1[c] += r[c];
            st2[c] = fmaf(r[c], r[c], st2[c]);
          }
          stg_v4(yrow + (long long)ox * p.ldy,
                 make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]), pack_bf16(r[4], r[5]),
                            pack_bf16(r[6], r[7])));
        }
      }
    }
  }

  // ---- per-CTA partial statistics: fixed-order sum over the threads that share a channel group
  if (p.stats_part != nullptr) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      red[tid * 16 + c] = st1[c];
      red[tid * 16 + 8 + c] = st2[c];
    }
    __syncthreads();
    const int nsp = tr * tc;   // threads per channel group
    for (int o = tid; o < ncg * 16; o += nthreads) {
      const int lcg = o / 16, e = o % 16;
      float a = 0.f;
      for (int s = 0; s < nsp; ++s) a += red[(s * ncg + lcg) * 16 + e];
      const int c = cb0 + lcg * 8 + (e & 7);
      if (c < p.C) {
        const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
        p.stats_part[(part * 2 + (e >> 3)) * p.C + c] = a;
      }
    }
  }
}

struct DwCfg {
  int ncg, tr, tc, tiles_x, tiles_y, ipb, parts, threads;
  size_t smem;
};

// Channel groups (8 channels each) per CTA: the largest count <= 16 that tiles C/8 without a
// half-empty last block (branch widths are 6*Cin: 96, 144, 240, ... -- multiples of 48, not 64).
static int pick_ncg(int C) {
  const int g = (C + 7) / 8;
  if (g <= 16) return g;
  int best = 8;
  double best_util = 0.0;
  for (int n = 16; n >= 6; --n) {
    const int blocks = (g + n - 1) / n;
    const double util = static_cast<double>(g) / (blocks * n);
    if (util > best_util + 1e-9) { best_util = util; best = n; }
  }
  return best;
}

static void pick_tile(DwCfg* c, int C, int out_h, int out_w, bool wide_ok) {
  c->ncg = pick_ncg(C);
  c->tc = (out_w > 8 && wide_ok) ? 2 : 1;
  if (c->ncg <= 4) c->tc = (out_w > 16) ? 4 : c->tc;
  c->tr = 128 / (c->ncg * c->tc);
  if (c->tr < 4) c->tr = 4;
  if (c->tr > 16) c->tr = 16;
  if (c->tr > out_h) c->tr = out_h;
}

static DwCfg dw_fwd_cfg(int B, int C, int Ho, int Wo, int k, int s) {
  DwCfg c;
  pick_tile(&c, C, Ho, Wo, s == 1);
  c.threads = c.ncg * c.tr * c.tc;
  c.tiles_x = (Wo + c.tc * 8 - 1) / (c.tc * 8);
  c.tiles_y = (Ho + c.tr - 1) / c.tr;
  c.ipb = B >= 64 ? 8 : 1;
  c.parts = c.tiles_x * c.tiles_y * ((B + c.ipb - 1) / c.ipb);
  const int rows_in = (c.tr - 1) * s + k, cols_in = (c.tc * 8 - 1) * s + k;
  c.smem = (size_t)rows_in * cols_in * c.ncg * 16 + (size_t)k * k * c.ncg * 8 * 4 +
           (size_t)c.threads * 16 * 4;
  return c;
}

template <int K, int S>
static int launch_dw_fwd(const DwFwdParams& p, const DwCfg& c, cudaStream_t st) {
  static size_t configured = 0;
  if (c.smem > configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(dw_fwd_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = 160 * 1024;
  }
  dim3 grid(c.tiles_x * c.tiles_y, (p.C + c.ncg * 8 - 1) / (c.ncg * 8), (p.B + c.ipb - 1) / c.ipb);
  dw_fwd_kernel<K, S><<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

}  // namespace an

using namespace an;

namespace an {   // row-streaming kernels for materialised operands (dw_stream.cu)
int dw_stream_fwd_rows(int B, int C, int Ho, int Wo, int k, int stride);
int dw_stream_dgrad_rows(int B, int C, int H, int W, int k, int stride);
int dw_stream_wgrad_rows(int B, int C, int Ho, int Wo, int k, int stride);
int dw_stream_fwd(const an_dw_fwd_args* a, cudaStream_t st);
int dw_stream_bwd_data(const an_dw_bwd_args* a, cudaStream_t st);
int dw_stream_wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st);
}

extern "C" int an_dw_fwd_stats_rows(int B, int C, int Ho, int Wo, int k, int stride, int fused) {
  if (!fused) return dw_stream_fwd_rows(B, C, Ho, Wo, k, stride);
  return dw_fwd_cfg(B, C, Ho, Wo, k, stride).parts;
}

extern "C" int an_dw_fwd(const an_dw_fwd_args* a, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_fwd: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_fwd: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && a->ldx % 8 == 0 && a->ldy % 8 == 0, "an_dw_fwd: channels/pitches must be multiples of 8");
  if (a->scale == nullptr) return dw_stream_fwd(a, static_cast<cudaStream_t>(stream));
  const int P = (a->k - 1) / 2;
  const int Ho = (a->H + 2 * P - a->k) / a->stride + 1, Wo = (a->W + 2 * P - a->k) / a->stride + 1;
  const DwCfg c = dw_fwd_cfg(a->B, a->C, Ho, Wo, a->k, a->stride);
  AN_REQUIRE(c.smem <= 160 * 1024, "an_dw_fwd: tile too large (%zu B smem)", c.smem);
  DwFwdParams p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(a->x);
  p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
  p.w = a->w; p.scale = a->scale; p.shift = a->shift; p.stats_part = a->stats_part;
  p.act = a->act; p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.Ho = Ho; p.Wo = Wo;
  p.ldx = a->ldx; p.ldy = a->ldy;
  p.ncg = c.ncg; p.tr = c.tr; p.tc = c.tc; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->k * 10 + a->stride) {
    case 31: return launch_dw_fwd<3, 1>(p, c, st);
    case 32: return launch_dw_fwd<3, 2>(p, c, st);
    case 51: return launch_dw_fwd<5, 1>(p, c, st);
    case 52: return launch_dw_fwd<5, 2>(p, c, st);
    case 71: return launch_dw_fwd<7, 1>(p, c, st);
    case 72: return launch_dw_fwd<7, 2>(p, c, st);
  }
  return 2;
}

// =================================================================================================
// backward: input gradient
// =================================================================================================
// du[n,h,w,c] = sum_{kr,kc} dd[n,(h+P-kr)/S,(w+P-kc)/S,c] * wt[kr][kc][c]   (divisible terms only)
// with dd = ca*g + cb*d_raw + cc (BN-backward apply of THIS layer's BN, on load), then
// ge = du * act'(scale_e*e_raw + shift_e) (activation backward of the PRODUCER layer) and the
// partials sum(ge), sum(ge*xhat_e) for the producer's BN backward.
namespace an {

struct DwBwdDataParams {
  const __nv_bfloat16 *g, *draw, *eraw;
  __nv_bfloat16* ge;
  const float *w, *ca, *cb, *cc;
  const float *e_scale, *e_shift, *e_mean, *e_rstd;
  float* stats_part;
  int act;
  int B, H, W, C, Ho, Wo;
  long long ldg, ldd, lde, ldge;
  int ncg, tr, tc, tiles_x, tiles_y, ipb;
};

__host__ __device__ constexpr int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

template <int K, int S>
__global__ void dw_bwd_data_kernel(const DwBwdDataParams p) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int P = (K - 1) / 2;
  constexpr int FL = floordiv(-P, S);
  constexpr int NPC = floordiv(7 + P, S) - FL + 1;    // patch columns one 8-wide strip touches
  const int ncg = p.ncg, tr = p.tr, tc = p.tc;
  const int h0 = (blockIdx.x / p.tiles_x) * tr, w0t = (blockIdx.x % p.tiles_x) * tc * 8;
  const int ho_base = floordiv(h0 + P - K + 1, S);
  const int ho_last = floordiv(h0 + tr - 1 + P, S);
  const int rows_p = ho_last - ho_base + 1;
  const int rows_p_max = floordiv(tr - 1 + K - 1, S) + 2;
  const int cols_p = (tc - 1) * (8 / S) + NPC;
  const int wo_base = w0t / S + FL;
  uint4* patch = reinterpret_cast<uint4*>(dsm);                                          // [rows_p][cols_p][ncg]
  float* wsm = reinterpret_cast<float*>(dsm + (size_t)rows_p_max * cols_p * ncg * 16);   // [K*K][ncg*8]
  float* red = wsm + K * K * ncg * 8;

  const int nthreads = blockDim.x, tid = threadIdx.x;
  const int cg = tid % ncg, strip = (tid / ncg) % tc, trow = tid / (ncg * tc);
  const int cb0 = blockIdx.y * ncg * 8, c0 = cb0 + cg * 8;
  const bool cg_ok = c0 < p.C;

  for (int i = tid; i < K * K * ncg * 8; i += nthreads) {
    const int tap = i / (ncg * 8), c = cb0 + i % (ncg * 8);
    const int e8 = i % 8, lcg8 = (i % (ncg * 8)) / 8;      // [tap][half][cg][4]: 16-byte lane stride
    wsm[((tap * 2 + e8 / 4) * ncg + lcg8) * 4 + (e8 & 3)] = (c < p.C) ? p.w[(long long)tap * p.C + c] : 0.f;
  }
  float st1[8], st2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { st1[i] = 0.f; st2[i] = 0.f; }

  for (int ii = 0; ii < p.ipb; ++ii) {
    const int n = blockIdx.z * p.ipb + ii;
    if (n >= p.B) break;
    __syncthreads();
    const int nchunks = rows_p * cols_p * ncg;
    for (int i = tid; i < nchunks; i += nthreads) {
      const int lcg = i % ncg, pc = (i / ncg) % cols_p, pr = i / (ncg * cols_p);
      const int ho = ho_base + pr, wo = wo_base + pc, ch = cb0 + lcg * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (ho >= 0 && ho < p.Ho && wo >= 0 && wo < p.Wo && ch < p.C) {
        const long long pix = ((long long)n * p.Ho + ho) * p.Wo + wo;
        const uint4 gv = ldg_nc_v4(p.g + pix * p.ldg + ch);
        if (p.ca == nullptr) {
          v = gv;                                   // dd already materialised
        } else {
          const uint4 dv = ldg_nc_v4(p.draw + pix * p.ldd + ch);
          float a[8], b[8], c[8];
          load_coef8(p.ca, ch, p.C, a, 0.f);
          load_coef8(p.cb, ch, p.C, b, 0.f);
          load_coef8(p.cc, ch, p.C, c, 0.f);
          const uint32_t gi[4] = {gv.x, gv.y, gv.z, gv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float lo = fmaf(a[2 * e], bf16_lo(gi[e]), fmaf(b[2 * e], bf16_lo(di[e]), c[2 * e]));
            const float hi = fmaf(a[2 * e + 1], bf16_hi(gi[e]), fmaf(b[2 * e + 1], bf16_hi(di[e]), c[2 * e + 1]));
            o[e] = pack_bf16(lo, hi);
          }
          v = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      patch[i] = v;
    }
    __syncthreads();

    const int h = h0 + trow;
    if (cg_ok && h < p.H) {
      float acc[8][8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;
#pragma unroll 1
      for (int kr = 0; kr < K; ++kr) {
        const int t = h + P - kr;
        if (t < 0 || (t % S) != 0) continue;
        const int ho = t / S;
        if (ho >= p.Ho) continue;
        float wr[K][8];
#pragma unroll
        for (int kc = 0; kc < K; ++kc) {
          const float4 lo = *reinterpret_cast<const float4*>(wsm + (((kr * K + kc) * 2 + 0) * ncg + cg) * 4);
          const float4 hi = *reinterpret_cast<const float4*>(wsm + (((kr * K + kc) * 2 + 1) * ncg + cg) * 4);
          wr[kc][0] = lo.x; wr[kc][1] = lo.y; wr[kc][2] = lo.z; wr[kc][3] = lo.w;
          wr[kc][4] = hi.x; wr[kc][5] = hi.y; wr[kc][6] = hi.z; wr[kc][7] = hi.w;
        }
        const uint4* prow = patch + ((size_t)(ho - ho_base) * cols_p + strip * (8 / S)) * ncg + cg;
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
          const uint4 v = prow[(size_t)pc * ncg];
          const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                              bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int kc = j + P - S * FL - S * pc;
            if (kc >= 0 && kc < K) {
#pragma unroll
              for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(wr[kc][c], f[c], acc[j][c]);
            }
          }
        }
      }
      // ---- activation backward of the producer + statistics for its BN backward
      float sc[8], sh[8], mu[8], rs[8];
      load_coef8(p.e_scale, c0, p.C, sc, 1.f);
      load_coef8(p.e_shift, c0, p.C, sh, 0.f);
      load_coef8(p.e_mean, c0, p.C, mu, 0.f);
      load_coef8(p.e_rstd, c0, p.C, rs, 1.f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int w = w0t + strip * 8 + j;
        if (w < p.W) {
          const long long pix = ((long long)n * p.H + h) * p.W + w;
          const uint4 ev = ldg_nc_v4(p.eraw + pix * p.lde + c0);
          const uint32_t ei[4] = {ev.x, ev.y, ev.z, ev.w};
          float r[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float x = (c & 1) ? bf16_hi(ei[c >> 1]) : bf16_lo(ei[c >> 1]);
            const float gq = round_bf16(acc[j][c] * act_bwd(fmaf(sc[c], x, sh[c]), p.act));
            r[c] = gq;
            st1[c] += gq;
            st2[c] = fmaf(gq, (x - mu[c]) * rs[c], st2[c]);
          }
          stg_v4(p.ge + pix * p.ldge + c0,
                 make_uint4(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]), pack_bf16(r[4], r[5]),
                            pack_bf16(r[6], r[7])));
        }
      }
    }
  }

  if (p.stats_part != nullptr) {
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      red[tid * 16 + c] = st1[c];
      red[tid * 16 + 8 + c] = st2[c];
    }
    __syncthreads();
    const int nsp = tr * tc;
    for (int o = tid; o < ncg * 16; o += nthreads) {
      const int lcg = o / 16, e = o % 16;
      float a = 0.f;
      for (int s = 0; s < nsp; ++s) a += red[(s * ncg + lcg) * 16 + e];
      const int c = cb0 + lcg * 8 + (e & 7);
      if (c < p.C) {
        const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
        p.stats_part[(part * 2 + (e >> 3)) * p.C + c] = a;
      }
    }
  }
}

static DwCfg dw_bwd_data_cfg(int B, int C, int H, int W, int k, int s) {
  DwCfg c;
  pick_tile(&c, C, H, W, true);
  c.threads = c.ncg * c.tr * c.tc;
  c.tiles_x = (W + c.tc * 8 - 1) / (c.tc * 8);
  c.tiles_y = (H + c.tr - 1) / c.tr;
  c.ipb = B >= 64 ? 8 : 1;
  c.parts = c.tiles_x * c.tiles_y * ((B + c.ipb - 1) / c.ipb);
  const int P = (k - 1) / 2;
  const int FL = floordiv(-P, s), NPC = floordiv(7 + P, s) - FL + 1;
  const int rows_p_max = floordiv(c.tr - 1 + k - 1, s) + 2, cols_p = (c.tc - 1) * (8 / s) + NPC;
  c.smem = (size_t)rows_p_max * cols_p * c.ncg * 16 + (size_t)k * k * c.ncg * 8 * 4 + (size_t)c.threads * 16 * 4;
  return c;
}

template <int K, int S>
static int launch_dw_bwd_data(const DwBwdDataParams& p, const DwCfg& c, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(dw_bwd_data_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  dim3 grid(c.tiles_x * c.tiles_y, (p.C + c.ncg * 8 - 1) / (c.ncg * 8), (p.B + c.ipb - 1) / c.ipb);
  dw_bwd_data_kernel<K, S><<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

// =================================================================================================
// backward: weight gradient
// =================================================================================================
// dwt[kr][kc][c] = sum_{n,ho,wo} dd[n,ho,wo,c] * u[n, ho*S-P+kr, wo*S-P+kc, c],
// u = act(scale_e*e_raw + shift_e) (zero outside the image), dd as above.
// Thread = (channel group, tap row kr, row lane); the K tap columns of one tap row live in
// registers; per-CTA partials [cta][K*K][C], summed in fixed order by dw_wgrad_reduce_kernel.
struct DwBwdWeightParams {
  const __nv_bfloat16 *g, *draw, *eraw;
  const float *ca, *cb, *cc, *e_scale, *e_shift;
  float* part;            // [ctas][K*K][C]
  int act;
  int B, H, W, C, Ho, Wo;
  long long ldg, ldd, lde;
  int ncg, tro, two, tiles_x, tiles_y, ipb, rl;
};

template <int K, int S>
__global__ void dw_bwd_weight_kernel(const DwBwdWeightParams p) {
  extern __shared__ __align__(16) uint8_t dsm[];
  constexpr int P = (K - 1) / 2;
  const int ncg = p.ncg, tro = p.tro, two = p.two, rl = p.rl;
  const int rows_in = (tro - 1) * S + K, cols_in = (two - 1) * S + K;
  uint4* upatch = reinterpret_cast<uint4*>(dsm);                                   // [rows_in][cols_in][ncg]
  uint4* ddt = upatch + (size_t)rows_in * cols_in * ncg;                           // [tro][two][ncg]
  float* red = reinterpret_cast<float*>(ddt + (size_t)tro * two * ncg);            // [rl*K*ncg][K*8]
  const int nthreads = blockDim.x, tid = threadIdx.x;
  const int cg = tid % ncg, kr = (tid / ncg) % K, lane_r = tid / (ncg * K);
  const bool worker = lane_r < rl;
  const int cb0 = blockIdx.y * ncg * 8;

  float acc[K][8];
#pragma unroll
  for (int kc = 0; kc < K; ++kc)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[kc][c] = 0.f;

  for (int ii = 0; ii < p.ipb; ++ii) {
    const int n = blockIdx.x * p.ipb + ii;
    if (n >= p.B) break;
    for (int tile = 0; tile < p.tiles_x * p.tiles_y; ++tile) {
      const int oy0 = (tile / p.tiles_x) * tro, ox0 = (tile % p.tiles_x) * two;
      const int iy0 = oy0 * S - P, ix0 = ox0 * S - P;
      __syncthreads();
      for (int i = tid; i < rows_in * cols_in * ncg; i += nthreads) {
        const int lcg = i % ncg, px = (i / ncg) % cols_in, py = i / (ncg * cols_in);
        const int iy = iy0 + py, ix = ix0 + px, ch = cb0 + lcg * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W && ch < p.C) {
          v = ldg_nc_v4(p.eraw + (((long long)n * p.H + iy) * p.W + ix) * p.lde + ch);
          if (p.e_scale != nullptr) {
            float a[8], b[8];
            load_coef8(p.e_scale, ch, p.C, a, 1.f);
            load_coef8(p.e_shift, ch, p.C, b, 0.f);
            uint32_t vi[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
              vi[e] = pack_bf16(act_fwd(fmaf(a[2 * e], bf16_lo(vi[e]), b[2 * e]), p.act),
                                act_fwd(fmaf(a[2 * e + 1], bf16_hi(vi[e]), b[2 * e + 1]), p.act));
            v = make_uint4(vi[0], vi[1], vi[2], vi[3]);
          }
        }
        upatch[i] = v;
      }
      for (int i = tid; i < tro * two * ncg; i += nthreads) {
        const int lcg = i % ncg, ox = ox0 + (i / ncg) % two, oy = oy0 + i / (ncg * two), ch = cb0 + lcg * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (oy < p.Ho && ox < p.Wo && ch < p.C) {
          const long long pix = ((long long)n * p.Ho + oy) * p.Wo + ox;
          const uint4 gv = ldg_nc_v4(p.g + pix * p.ldg + ch);
          if (p.ca == nullptr) {
            v = gv;                                 // dd already materialised
          } else {
            const uint4 dv = ldg_nc_v4(p.draw + pix * p.ldd + ch);
            float a[8], b[8], c[8];
            load_coef8(p.ca, ch, p.C, a, 0.f);
            load_coef8(p.cb, ch, p.C, b, 0.f);
            load_coef8(p.cc, ch, p.C, c, 0.f);
            const uint32_t gi[4] = {gv.x, gv.y, gv.z, gv.w}, di[4] = {dv.x, dv.y, dv.z, dv.w};
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              o[e] = pack_bf16(fmaf(a[2 * e], bf16_lo(gi[e]), fmaf(b[2 * e], bf16_lo(di[e]), c[2 * e])),
                               fmaf(a[2 * e + 1], bf16_hi(gi[e]), fmaf(b[2 * e + 1], bf16_hi(di[e]), c[2 * e + 1])));
            v = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
        ddt[i] = v;
      }
      __syncthreads();
      if (worker) {
        for (int r = lane_r; r < tro; r += rl) {
          const uint4* urow = upatch + ((size_t)(r * S + kr) * cols_in) * ncg + cg;
          const uint4* drow = ddt + ((size_t)r * two) * ncg + cg;
          for (int wc = 0; wc < two; wc += 8) {
            float d[8][8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = drow[(size_t)(wc + j) * ncg];
              d[j][0] = bf16_lo(v.x); d[j][1] = bf16_hi(v.x); d[j][2] = bf16_lo(v.y); d[j][3] = bf16_hi(v.y);
              d[j][4] = bf16_lo(v.z); d[j][5] = bf16_hi(v.z); d[j][6] = bf16_lo(v.w); d[j][7] = bf16_hi(v.w);
            }
#pragma unroll
            for (int ic = 0; ic < 7 * S + K; ++ic) {
              const uint4 v = urow[(size_t)(wc * S + ic) * ncg];
              const float f[8] = {bf16_lo(v.x), bf16_hi(v.x), bf16_lo(v.y), bf16_hi(v.y),
                                  bf16_lo(v.z), bf16_hi(v.z), bf16_lo(v.w), bf16_hi(v.w)};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int kc = ic - j * S;
                if (kc >= 0 && kc < K) {
#pragma unroll
                  for (int c = 0; c < 8; ++c) acc[kc][c] = fmaf(d[j][c], f[c], acc[kc][c]);
                }
              }
            }
          }
        }
      }
    }
  }
  // ---- reduce over row lanes (fixed order), one partial per CTA
  __syncthreads();
  if (worker) {
#pragma unroll
    for (int kc = 0; kc < K; ++kc)
#pragma unroll
      for (int c = 0; c < 8; ++c) red[(size_t)tid * (K * 8) + kc * 8 + c] = acc[kc][c];
  }
  __syncthreads();
  for (int o = tid; o < ncg * K * K * 8; o += nthreads) {
    const int c8 = o % 8, kc = (o / 8) % K, lkr = (o / (8 * K)) % K, lcg = o / (8 * K * K);
    float a = 0.f;
    for (int l = 0; l < rl; ++l) a += red[(size_t)((l * K + lkr) * ncg + lcg) * (K * 8) + kc * 8 + c8];
    const int c = cb0 + lcg * 8 + c8;
    if (c < p.C) p.part[((long long)blockIdx.x * K * K + lkr * K + kc) * p.C + c] = a;
  }
}

struct DwWgCfg {
  int ncg, tro, two, tiles_x, tiles_y, ipb, rl, threads, ctas;
  size_t smem;
};

static DwWgCfg dw_bwd_weight_cfg(int B, int C, int Ho, int Wo, int k, int s) {
  DwWgCfg c;
  c.ncg = pick_ncg(C);
  if (c.ncg > 12) c.ncg = (c.ncg + 1) / 2;
  c.two = (Wo > 8 && s == 1) ? 16 : 8;
  c.tro = Ho < 8 ? Ho : 8;
  c.rl = 256 / (c.ncg * k);
  if (c.rl > c.tro) c.rl = c.tro;
  if (c.rl < 1) c.rl = 1;
  c.threads = c.ncg * k * c.rl;
  c.tiles_x = (Wo + c.two - 1) / c.two;
  c.tiles_y = (Ho + c.tro - 1) / c.tro;
  const int cblocks = (C + c.ncg * 8 - 1) / (c.ncg * 8);
  // enough CTAs for ~4 waves; each CTA walks every tile of its images
  int ipb = (B * cblocks) / (4 * sm_count());
  if (ipb < 1) ipb = 1;
  if (ipb > 8) ipb = 8;
  c.ipb = ipb;
  c.ctas = (B + ipb - 1) / ipb;
  const int rows_in = (c.tro - 1) * s + k, cols_in = (c.two - 1) * s + k;
  size_t tiles = ((size_t)rows_in * cols_in + (size_t)c.tro * c.two) * c.ncg * 16;
  size_t red = (size_t)c.threads * k * 8 * 4;
  c.smem = tiles + red;
  return c;
}

template <int K, int S>
static int launch_dw_bwd_weight(const DwBwdWeightParams& p, const DwWgCfg& c, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(dw_bwd_weight_kernel<K, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured = true;
  }
  dim3 grid(c.ctas, (p.C + c.ncg * 8 - 1) / (c.ncg * 8), 1);
  dw_bwd_weight_kernel<K, S><<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

// dst[c][tap] (master layout [C,1,k,k]) = sum_rows part[row][tap][c]
__global__ void dw_wgrad_reduce_kernel(const float* __restrict__ part, int rows, int taps, int C, float* dst) {
  // block = 32 consecutive (tap, channel) outputs x 8 row lanes (fixed order -> deterministic)
  __shared__ double red[8][33];
  const int ex = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ex;
  const int n = taps * C;
  double a = 0.0;
  if (i < n)
    for (int r = rl; r < rows; r += 8) a += static_cast<double>(part[(long long)r * n + i]);
  red[rl][ex] = a;
  __syncthreads();
  if (rl == 0 && i < n) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][ex];
    const int tap = i / C, c = i % C;
    dst[(long long)c * taps + tap] = static_cast<float>(t);
  }
}

}  // namespace an

extern "C" int an_dw_bwd_data_stats_rows(int B, int C, int H, int W, int k, int stride, int fused) {
  if (!fused) return dw_stream_dgrad_rows(B, C, H, W, k, stride);
  return dw_bwd_data_cfg(B, C, H, W, k, stride).parts;
}

extern "C" int an_dw_bwd_data(const an_dw_bwd_args* a, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_bwd_data: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_bwd_data: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && a->ldg % 8 == 0 && a->ldd % 8 == 0 && a->lde % 8 == 0 && a->ldge % 8 == 0, "an_dw_bwd_data: bad layout");
  AN_REQUIRE(a->stride == 1 || (a->H % 2 == 0 && a->W % 2 == 0), "an_dw_bwd_data: stride 2 needs even H, W");
  if (a->ca == nullptr) return dw_stream_bwd_data(a, static_cast<cudaStream_t>(stream));
  const int P = (a->k - 1) / 2;
  const int Ho = (a->H + 2 * P - a->k) / a->stride + 1, Wo = (a->W + 2 * P - a->k) / a->stride + 1;
  const DwCfg c = dw_bwd_data_cfg(a->B, a->C, a->H, a->W, a->k, a->stride);
  AN_REQUIRE(c.smem <= 160 * 1024, "an_dw_bwd_data: tile too large");
  DwBwdDataParams p;
  p.g = reinterpret_cast<const __nv_bfloat16*>(a->g); p.draw = reinterpret_cast<const __nv_bfloat16*>(a->draw);
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw); p.ge = reinterpret_cast<__nv_bfloat16*>(a->ge);
  p.w = a->w; p.ca = a->ca; p.cb = a->cb; p.cc = a->cc;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.Ho = Ho; p.Wo = Wo;
  p.ldg = a->ldg; p.ldd = a->ldd; p.lde = a->lde; p.ldge = a->ldge;
  p.ncg = c.ncg; p.tr = c.tr; p.tc = c.tc; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->k * 10 + a->stride) {
    case 31: return launch_dw_bwd_data<3, 1>(p, c, st);
    case 32: return launch_dw_bwd_data<3, 2>(p, c, st);
    case 51: return launch_dw_bwd_data<5, 1>(p, c, st);
    case 52: return launch_dw_bwd_data<5, 2>(p, c, st);
    case 71: return launch_dw_bwd_data<7, 1>(p, c, st);
    case 72: return launch_dw_bwd_data<7, 2>(p, c, st);
  }
  return 2;
}

extern "C" int an_dw_bwd_weight_rows(int B, int C, int H, int W, int k, int stride, int fused) {
  const int P = (k - 1) / 2;
  const int Ho = (H + 2 * P - k) / stride + 1, Wo = (W + 2 * P - k) / stride + 1;
  if (!fused) return dw_stream_wgrad_rows(B, C, Ho, Wo, k, stride);
  return dw_bwd_weight_cfg(B, C, Ho, Wo, k, stride).ctas;
}

extern "C" int an_dw_bwd_weight(const an_dw_bwd_args* a, float* partials, float* dw, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_bwd_weight: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_bwd_weight: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && partials && dw, "an_dw_bwd_weight: bad arguments");
  if (a->ca == nullptr && a->e_scale == nullptr) {
    cudaStream_t s1 = static_cast<cudaStream_t>(stream);
    if (int rc1 = dw_stream_wgrad(a, partials, s1)) return rc1;
    const int taps1 = a->k * a->k;
    dw_wgrad_reduce_kernel<<<(taps1 * a->C + 31) / 32, 256, 0, s1>>>(
        partials, an_dw_bwd_weight_rows(a->B, a->C, a->H, a->W, a->k, a->stride, 0), taps1, a->C, dw);
    AN_CHECK_LAUNCH();
    return 0;
  }
  const int P = (a->k - 1) / 2;
  const int Ho = (a->H + 2 * P - a->k) / a->stride + 1, Wo = (a->W + 2 * P - a->k) / a->stride + 1;
  const DwWgCfg c = dw_bwd_weight_cfg(a->B, a->C, Ho, Wo, a->k, a->stride);
  AN_REQUIRE(c.smem <= 160 * 1024, "an_dw_bwd_weight: tile too large");
  DwBwdWeightParams p;
  p.g = reinterpret_cast<const __nv_bfloat16*>(a->g); p.draw = reinterpret_cast<const __nv_bfloat16*>(a->draw);
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw);
  p.ca = a->ca; p.cb = a->cb; p.cc = a->cc; p.e_scale = a->e_scale; p.e_shift = a->e_shift;
  p.part = partials; p.act = a->act;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.Ho = Ho; p.Wo = Wo;
  p.ldg = a->ldg; p.ldd = a->ldd; p.lde = a->lde;
  p.ncg = c.ncg; p.tro = c.tro; p.two = c.two; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb; p.rl = c.rl;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = 2;
  switch (a->k * 10 + a->stride) {
    case 31: rc = launch_dw_bwd_weight<3, 1>(p, c, st); break;
    case 32: rc = launch_dw_bwd_weight<3, 2>(p, c, st); break;
    case 51: rc = launch_dw_bwd_weight<5, 1>(p, c, st); break;
    case 52: rc = launch_dw_bwd_weight<5, 2>(p, c, st); break;
    case 71: rc = launch_dw_bwd_weight<7, 1>(p, c, st); break;
    case 72: rc = launch_dw_bwd_weight<7, 2>(p, c, st); break;
  }
  if (rc) return rc;
  const int taps = a->k * a->k;
  dw_wgrad_reduce_kernel<<<(taps * a->C + 31) / 32, 256, 0, st>>>(partials, c.ctas, taps, a->C, dw);
  AN_CHECK_LAUNCH();
  return 0;
}
