// Stride-1 depthwise convolution kernels, "row-streaming" design with TMA-staged patches
// (85 % of the depthwise MACs of the AtomNAS supernets are stride 1; stride 2 stays on the tile
// kernels of dw_conv.cu).
//
//   staging = the (tile + halo) patch of the NHWC bf16 input is one 4-D TMA box
//             {64 channels, PW, PH, 1}; out-of-image coordinates are zero-filled by the TMA unit,
//             which IS the convolution padding.  Two patch buffers + mbarriers: the TMA engine
//             fetches image n+1 while all warps compute image n.  The input must already be the
//             operand (activated for forward, BN-backward-applied for the gradients) -- the
//             engine materialises it with one an_affine_act pass; depthwise is FMA-issue bound,
//             not HBM bound, so trading a little bandwidth for zero staging instructions pays.
//   thread  = one channel PAIR (a 32-bit bf16x2 word; the lanes of a warp are 32 consecutive
//             pairs, so every shared / global access of a warp is one contiguous 128-byte line)
//             x 4 output columns x every row of the tile, streamed top to bottom;
//   weights = all k*k taps of the pair stay in registers (float2 each);
//   math    = packed FFMA2 (fma.rn.f32x2), fp32 accumulation, a ring of k partially accumulated
//             output rows in registers; consecutive FFMA2 touch different accumulators.
//
// forward  (MODE 0): y = dwconv(u)                                        + BN partials of y
// backward (MODE 1): du = dwconv(dd, flipped taps);
//                    ge = du * act'(e_scale*eraw + e_shift)               + partials for BN_e
// weight gradient  : dW[c][kr][kc] = sum dd * u, all k*k accumulators of the pair in registers.
// Reference call site: ConvBNReLU(h, h, k, 1, groups=h), models/mobilenet_base.py:330-336.
#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  rc = *reinterpret_cast<const unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) { return make_float2(bf16_lo(w), bf16_hi(w)); }

struct DwS1Params {
  CUtensorMap tm_in;                  // MODE 0: activated input u;  MODE 1: dd
  const float* w;                     // [k*k][C] tap-major
  __nv_bfloat16* out;
  long long ldo;
  const __nv_bfloat16* eraw;          // MODE 1: producer raw output at the output position
  long long lde;
  const float *e_scale, *e_shift, *e_mean, *e_rstd;
  float* stats_part;
  int act;
  int B, H, W, C;                     // stride 1: output size == input size
  int nwc, th, tiles_x, tiles_y, ipb;
};

template <int K, int MODE>
__global__ void __launch_bounds__(128) dw_s1_kernel(const __grid_constant__ DwS1Params p) {
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 127) & ~static_cast<uintptr_t>(127));
  constexpr int P = (K - 1) / 2;
  const int nwc = p.nwc, th = p.th;
  const int PW = 4 * nwc + K - 1, PH = th + K - 1;
  const uint32_t patch_bytes = static_cast<uint32_t>(PH * PW * 128);
  uint8_t* patch0 = dsm;                                               // 2 x [PH][PW][32 words]
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + 2 * (size_t)patch_bytes);
  float* red = reinterpret_cast<float*>(full + 2);                     // [nwc][2][64]
  const int nthreads = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, g = tid >> 5;
  const int cb0 = blockIdx.y * 64;
  const int ch = cb0 + 2 * lane;
  const bool ch_ok = ch < p.C;
  const int oy0 = (blockIdx.x / p.tiles_x) * th, ox0 = (blockIdx.x % p.tiles_x) * 4 * nwc;
  const int n0 = blockIdx.z * p.ipb;
  const int nimg = min(p.ipb, p.B - n0);

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
  }
  __syncthreads();
  if (tid == 0 && nimg > 0) {
    mbar_expect_tx(&full[0], patch_bytes);
    tma_load_4d(patch0, &p.tm_in, &full[0], cb0, ox0 - P, oy0 - P, n0);
  }

  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    const int src = (MODE == 0) ? t : (K * K - 1 - t);              // data gradient: flipped taps
    w[t] = ch_ok ? make_float2(__ldg(p.w + (long long)src * p.C + ch), __ldg(p.w + (long long)src * p.C + ch + 1))
                 : make_float2(0.f, 0.f);
  }
  float2 es = make_float2(1.f, 1.f), eh = make_float2(0.f, 0.f), em = eh, er = es;
  const bool e_bn = (MODE == 1) && p.e_scale != nullptr;
  if (e_bn && ch_ok) {
    es = make_float2(p.e_scale[ch], p.e_scale[ch + 1]);
    eh = make_float2(p.e_shift[ch], p.e_shift[ch + 1]);
    em = make_float2(p.e_mean[ch], p.e_mean[ch + 1]);
    er = make_float2(p.e_rstd[ch], p.e_rstd[ch + 1]);
  }
  float2 st1 = make_float2(0.f, 0.f), st2 = st1;

  for (int ii = 0; ii < nimg; ++ii) {
    const int n = n0 + ii;
    const int b = ii & 1;
    if (tid == 0 && ii + 1 < nimg) {            // prefetch the next image's patch into the other buffer
      mbar_expect_tx(&full[b ^ 1], patch_bytes);
      tma_load_4d(patch0 + (size_t)(b ^ 1) * patch_bytes, &p.tm_in, &full[b ^ 1], cb0, ox0 - P, oy0 - P, n + 1);
    }
    mbar_wait(&full[b], (ii >> 1) & 1);
    const uint32_t* patch = reinterpret_cast<const uint32_t*>(patch0 + (size_t)b * patch_bytes);

    // ---- stream the patch rows; ring of K partially accumulated output rows in registers
    float2 acc[K][4];
#pragma unroll
    for (int s = 0; s < K; ++s)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[s][j] = make_float2(0.f, 0.f);
    const uint32_t* pcol = patch + (g * 4) * 32 + lane;
    for (int base = 0; base < PH; base += K) {
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const int ir = base + u;
        if (ir < PH) {
          float2 x[K + 3];
#pragma unroll
          for (int c = 0; c < K + 3; ++c) x[c] = unpack2(pcol[(ir * PW + c) * 32]);
          const int rd = ir - (K - 1);                    // output row completed by this input row
          uint32_t xraw[4] = {0u, 0u, 0u, 0u};            // producer output under the finishing row
          if (MODE == 1 && e_bn && ch_ok && rd >= 0 && oy0 + rd < p.H) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int ox = ox0 + g * 4 + j;
              if (ox < p.W)
                xraw[j] = __ldg(reinterpret_cast<const unsigned int*>(
                    p.eraw + (((long long)n * p.H + oy0 + rd) * p.W + ox) * p.lde + ch));
            }
          }
          bool vk[K];                                     // which of the K in-flight output rows exist
#pragma unroll
          for (int kr = 0; kr < K; ++kr) vk[kr] = (ir - kr >= 0) && (ir - kr < th);
          // kc outermost: consecutive FFMA2 hit 4*K different accumulators (no dependent chains)
#pragma unroll
          for (int kc = 0; kc < K; ++kc) {
#pragma unroll
            for (int kr = 0; kr < K; ++kr) {
              if (vk[kr]) {
                const int slot = (u - kr + K) % K;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[slot][j] = fma2(w[kr * K + kc], x[j + kc], acc[slot][j]);
              }
            }
          }
          const int sd = (u + 1) % K;
          if (rd >= 0) {
            const int oy = oy0 + rd;
            if (ch_ok && oy < p.H) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int ox = ox0 + g * 4 + j;
                if (ox < p.W) {
                  const long long pix = ((long long)n * p.H + oy) * p.W + ox;
                  float2 v = acc[sd][j];
                  if (MODE == 1 && e_bn) {
                    const float2 xr = unpack2(xraw[j]);
                    v.x = round_bf16(v.x * act_bwd(fmaf(es.x, xr.x, eh.x), p.act));
                    v.y = round_bf16(v.y * act_bwd(fmaf(es.y, xr.y, eh.y), p.act));
                    st1.x += v.x;
                    st1.y += v.y;
                    st2.x = fmaf(v.x, (xr.x - em.x) * er.x, st2.x);
                    st2.y = fmaf(v.y, (xr.y - em.y) * er.y, st2.y);
                  } else {
                    v.x = round_bf16(v.x);
                    v.y = round_bf16(v.y);
                    st1.x += v.x;
                    st1.y += v.y;
                    st2.x = fmaf(v.x, v.x, st2.x);
                    st2.y = fmaf(v.y, v.y, st2.y);
                  }
                  *reinterpret_cast<uint32_t*>(p.out + pix * p.ldo + ch) = pack_bf16(v.x, v.y);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[sd][j] = make_float2(0.f, 0.f);
          }
        }
      }
    }
    __syncthreads();          // every warp is done with buffer b before it is refilled
  }

  if (p.stats_part != nullptr) {
    red[(g * 2 + 0) * 64 + 2 * lane] = st1.x;
    red[(g * 2 + 0) * 64 + 2 * lane + 1] = st1.y;
    red[(g * 2 + 1) * 64 + 2 * lane] = st2.x;
    red[(g * 2 + 1) * 64 + 2 * lane + 1] = st2.y;
    __syncthreads();
    for (int o = tid; o < 128; o += nthreads) {
      const int s = o / 64, c = o % 64;
      float a = 0.f;
      for (int wv = 0; wv < nwc; ++wv) a += red[(wv * 2 + s) * 64 + c];
      if (cb0 + c < p.C) {
        const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
        p.stats_part[(part * 2 + s) * p.C + cb0 + c] = a;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient, stride 1
// ------------------------------------------------------------------------------------------------
struct DwS1WgParams {
  CUtensorMap tm_u, tm_dd;             // u: box {64, PW, PH, 1};  dd: box {64, TW, th, 1}
  float* part;                         // [ctas][K*K][C]
  int B, H, W, C;
  int nwc, th, tiles_x, tiles_y, ipb;
};

template <int K>
__global__ void __launch_bounds__(128) dw_s1_wgrad_kernel(const __grid_constant__ DwS1WgParams p) {
  extern __shared__ uint8_t dsm_raw[];
  uint8_t* dsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dsm_raw) + 127) & ~static_cast<uintptr_t>(127));
  constexpr int P = (K - 1) / 2;
  const int nwc = p.nwc, th = p.th;
  const int PW = 4 * nwc + K - 1, PH = th + K - 1, TW = 4 * nwc;
  const uint32_t u_bytes = static_cast<uint32_t>(PH * PW * 128), d_bytes = static_cast<uint32_t>(th * TW * 128);
  const uint32_t stage_bytes = u_bytes + d_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + 2 * (size_t)stage_bytes);
  float* red = reinterpret_cast<float*>(dsm);                                 // reused after the loops
  const int nthreads = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, g = tid >> 5;
  const int cb0 = blockIdx.y * 64;
  const int n0 = blockIdx.x * p.ipb;
  const int nimg = min(p.ipb, p.B - n0);
  const int ntiles = p.tiles_x * p.tiles_y;
  const int nwork = nimg * ntiles;

  auto issue = [&](int wi) {
    const int n = n0 + wi / ntiles, tile = wi % ntiles;
    const int oy0 = (tile / p.tiles_x) * th, ox0 = (tile % p.tiles_x) * TW;
    uint8_t* st = dsm + (size_t)(wi & 1) * stage_bytes;
    mbar_expect_tx(&full[wi & 1], stage_bytes);
    tma_load_4d(st, &p.tm_u, &full[wi & 1], cb0, ox0 - P, oy0 - P, n);
    tma_load_4d(st + u_bytes, &p.tm_dd, &full[wi & 1], cb0, ox0, oy0, n);
  };

  if (tid == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_u);
    tma_prefetch_desc(&p.tm_dd);
  }
  __syncthreads();
  if (tid == 0 && nwork > 0) issue(0);

  float2 acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);

  for (int wi = 0; wi < nwork; ++wi) {
    if (tid == 0 && wi + 1 < nwork) issue(wi + 1);
    mbar_wait(&full[wi & 1], (wi >> 1) & 1);
    const uint32_t* upatch = reinterpret_cast<const uint32_t*>(dsm + (size_t)(wi & 1) * stage_bytes);
    const uint32_t* ddt = reinterpret_cast<const uint32_t*>(dsm + (size_t)(wi & 1) * stage_bytes + u_bytes);
    const uint32_t* ucol = upatch + (g * 4) * 32 + lane;
    const uint32_t* dcol = ddt + (g * 4) * 32 + lane;
    // ring of the last K input rows (packed words) in registers: each row is read from smem once
    uint32_t ring[K][K + 3];
    for (int base = 0; base < PH; base += K) {
#pragma unroll
      for (int u = 0; u < K; ++u) {
        const int ir = base + u;
        if (ir < PH) {
#pragma unroll
          for (int c = 0; c < K + 3; ++c) ring[u][c] = ucol[(ir * PW + c) * 32];
          const int r = ir - (K - 1);            // output row whose K input rows are now resident
          if (r >= 0) {
            float2 d[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) d[j] = unpack2(dcol[(r * TW + j) * 32]);
#pragma unroll
            for (int kr = 0; kr < K; ++kr) {
              const int slot = (u + 1 + kr) % K;   // input row r + kr
              float2 x[K + 3];
#pragma unroll
              for (int c = 0; c < K + 3; ++c) x[c] = unpack2(ring[slot][c]);
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int kc = 0; kc < K; ++kc) acc[kr * K + kc] = fma2(d[j], x[j + kc], acc[kr * K + kc]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // ---- fixed-order reduction over the column-group warps, one partial per CTA
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    red[((size_t)g * K * K + t) * 64 + 2 * lane] = acc[t].x;
    red[((size_t)g * K * K + t) * 64 + 2 * lane + 1] = acc[t].y;
  }
  __syncthreads();
  for (int o = tid; o < K * K * 64; o += nthreads) {
    const int t = o / 64, c = o % 64;
    float a = 0.f;
    for (int wv = 0; wv < nwc; ++wv) a += red[((size_t)wv * K * K + t) * 64 + c];
    if (cb0 + c < p.C) p.part[((long long)blockIdx.x * K * K + t) * p.C + cb0 + c] = a;
  }
}

struct S1Cfg {
  int nwc, th, tiles_x, tiles_y, ipb, parts, threads;
  size_t smem;
};

static S1Cfg s1_cfg(int B, int C, int H, int W, int k, bool wgrad) {
  S1Cfg c;
  c.nwc = (W + 3) / 4;              // 4 warps x 4 columns: small CTAs, 2-3 resident per SM
  if (c.nwc > 4) c.nwc = 4;
  c.th = H < 14 ? H : 14;
  c.threads = 32 * c.nwc;
  c.tiles_x = (W + 4 * c.nwc - 1) / (4 * c.nwc);
  c.tiles_y = (H + c.th - 1) / c.th;
  const int PW = 4 * c.nwc + k - 1, PH = c.th + k - 1;
  const int cblocks = (C + 63) / 64;
  if (wgrad) {
    const size_t stage = (size_t)PH * PW * 128 + (size_t)c.th * 4 * c.nwc * 128;
    const size_t red = (size_t)c.nwc * k * k * 64 * 4;
    c.smem = (2 * stage > red ? 2 * stage : red) + 64 + 128;
    int ipb = (B * cblocks) / (4 * sm_count());     // ~4 waves of CTAs; each walks all tiles of its images
    if (ipb < 1) ipb = 1;
    if (ipb > 8) ipb = 8;
    c.ipb = ipb;
    c.parts = (B + ipb - 1) / ipb;
  } else {
    c.smem = 2 * (size_t)PH * PW * 128 + 64 + (size_t)c.nwc * 2 * 64 * 4 + 128;
    c.ipb = B >= 64 ? 8 : 1;
    c.parts = c.tiles_x * c.tiles_y * ((B + c.ipb - 1) / c.ipb);
  }
  return c;
}

template <int K, int MODE>
static int launch_s1(const DwS1Params& p, const S1Cfg& c, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(dw_s1_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(c.tiles_x * c.tiles_y, (p.C + 63) / 64, (p.B + c.ipb - 1) / c.ipb);
  dw_s1_kernel<K, MODE><<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

template <int K>
static int launch_s1_wgrad(const DwS1WgParams& p, const S1Cfg& c, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(dw_s1_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(c.parts, (p.C + 63) / 64, 1);
  dw_s1_wgrad_kernel<K><<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

// ---- entry points used by dw_conv.cu's C ABI functions when stride == 1 and the operands are
//      already materialised (scale == NULL resp. ca == NULL) ------------------------------------------
int dw_s1_fwd_rows(int B, int C, int H, int W, int k) { return s1_cfg(B, C, H, W, k, false).parts; }
int dw_s1_wgrad_rows(int B, int C, int H, int W, int k) { return s1_cfg(B, C, H, W, k, true).parts; }

int dw_s1_fwd(const an_dw_fwd_args* a, cudaStream_t st) {
  const S1Cfg c = s1_cfg(a->B, a->C, a->H, a->W, a->k, false);
  DwS1Params p;
  memset(&p, 0, sizeof(p));
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->x, a->C, a->W, a->H, a->B, a->ldx, 64, 4 * c.nwc + a->k - 1,
                                  c.th + a->k - 1)) return e;
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldo = a->ldy;
  p.stats_part = a->stats_part; p.act = a->act;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C;
  p.nwc = c.nwc; p.th = c.th; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb;
  switch (a->k) {
    case 3: return launch_s1<3, 0>(p, c, st);
    case 5: return launch_s1<5, 0>(p, c, st);
    case 7: return launch_s1<7, 0>(p, c, st);
  }
  return 2;
}

int dw_s1_bwd_data(const an_dw_bwd_args* a, cudaStream_t st) {
  const S1Cfg c = s1_cfg(a->B, a->C, a->H, a->W, a->k, false);
  DwS1Params p;
  memset(&p, 0, sizeof(p));
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->g, a->C, a->W, a->H, a->B, a->ldg, 64, 4 * c.nwc + a->k - 1,
                                  c.th + a->k - 1)) return e;
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->ge); p.ldo = a->ldge;
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw); p.lde = a->lde;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C;
  p.nwc = c.nwc; p.th = c.th; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb;
  switch (a->k) {
    case 3: return launch_s1<3, 1>(p, c, st);
    case 5: return launch_s1<5, 1>(p, c, st);
    case 7: return launch_s1<7, 1>(p, c, st);
  }
  return 2;
}

int dw_s1_wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st) {
  const S1Cfg c = s1_cfg(a->B, a->C, a->H, a->W, a->k, true);
  DwS1WgParams p;
  memset(&p, 0, sizeof(p));
  if (int e = make_tmap_bf16_nhwc(&p.tm_u, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, 4 * c.nwc + a->k - 1,
                                  c.th + a->k - 1)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->g, a->C, a->W, a->H, a->B, a->ldg, 64, 4 * c.nwc, c.th)) return e;
  p.part = partials;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C;
  p.nwc = c.nwc; p.th = c.th; p.tiles_x = c.tiles_x; p.tiles_y = c.tiles_y; p.ipb = c.ipb;
  switch (a->k) {
    case 3: return launch_s1_wgrad<3>(p, c, st);
    case 5: return launch_s1_wgrad<5>(p, c, st);
    case 7: return launch_s1_wgrad<7>(p, c, st);
  }
  return 2;
}

}  // namespace an
