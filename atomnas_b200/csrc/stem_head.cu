// Network ends: stem 3x3/s2 convolution on the fp32 NCHW input, global average pool fused with the
// head BatchNorm-apply + activation (+ dropout), and the label-smoothed cross entropy.
//
// Reference call sites: features[0] ConvBNReLU(3, c, stride=2) models/mobilenet_supernet.py:126-132,
// nn.AvgPool2d :154-155, nn.Dropout :160-161, CrossEntropyLabelSmooth utils/optim.py:199-207.
#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

// ------------------------------------------------------------------------------------------------
// stem forward: x fp32 [B,3,H,W] -> y bf16 [B,Ho,Wo,C] (C <= 32), raw conv output + BN partials
// ------------------------------------------------------------------------------------------------
constexpr int kStemThreads = 128;
constexpr int kStemChunks = 8;     // 128-pixel chunks per CTA

__global__ void __launch_bounds__(kStemThreads) stem_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, __nv_bfloat16* __restrict__ y,
    float* stats_part, int B, int H, int W, int C, int Ho, int Wo, long long ldy) {
  __shared__ float wsm[27 * 32];
  __shared__ float red[4 * 2 * 32];
  for (int i = threadIdx.x; i < 27 * 32; i += kStemThreads) {
    const int tap = i / 32, c = i % 32;        // w: [C][3][3][3] -> wsm[ci*9+kr*3+kc][c]
    wsm[i] = c < C ? w[c * 27 + tap] : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  float acc1[2] = {0.f, 0.f}, acc2[2] = {0.f, 0.f};   // per-lane column totals (2 groups of 16)
  for (int ch = 0; ch < kStemChunks; ++ch) {
    const long long pix = (static_cast<long long>(blockIdx.x) * kStemChunks + ch) * kStemThreads + threadIdx.x;
    const bool ok = pix < npix;
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = 0.f;
    if (ok) {
      const int ox = static_cast<int>(pix % Wo);
      const int oy = static_cast<int>((pix / Wo) % Ho);
      const int n = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
#pragma unroll
        for (int kr = 0; kr < 3; ++kr) {
          const int iy = oy * 2 - 1 + kr;
#pragma unroll
          for (int kc = 0; kc < 3; ++kc) {
            const int ix = ox * 2 - 1 + kc;
            float xv = 0.f;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W)
              xv = __ldg(x + ((static_cast<long long>(n) * 3 + ci) * H + iy) * W + ix);
            const float* wr = wsm + (ci * 9 + kr * 3 + kc) * 32;
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = fmaf(wr[c], xv, v[c]);
          }
        }
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = round_bf16(v[c]);
      __nv_bfloat16* yp = y + pix * ldy;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g * 8 < C)
          stg_v4(yp + g * 8, make_uint4(pack_bf16(v[8 * g], v[8 * g + 1]), pack_bf16(v[8 * g + 2], v[8 * g + 3]),
                                        pack_bf16(v[8 * g + 4], v[8 * g + 5]), pack_bf16(v[8 * g + 6], v[8 * g + 7])));
    }
    if (stats_part != nullptr) {
      float sq[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) sq[c] = v[c] * v[c];
      acc1[0] += warp_colsum16(v, lane);
      acc1[1] += warp_colsum16(v + 16, lane);
      acc2[0] += warp_colsum16(sq, lane);
      acc2[1] += warp_colsum16(sq + 16, lane);
    }
  }
  if (stats_part != nullptr) {
    if ((lane & 1) == 0) {
      red[(warp * 2 + 0) * 32 + (lane >> 1)] = acc1[0];
      red[(warp * 2 + 0) * 32 + 16 + (lane >> 1)] = acc1[1];
      red[(warp * 2 + 1) * 32 + (lane >> 1)] = acc2[0];
      red[(warp * 2 + 1) * 32 + 16 + (lane >> 1)] = acc2[1];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int s = threadIdx.x / 32, c = threadIdx.x % 32;
      float a = 0.f;
      for (int wv = 0; wv < 4; ++wv) a += red[(wv * 2 + s) * 32 + c];
      if (c < C) stats_part[(static_cast<long long>(blockIdx.x) * 2 + s) * C + c] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// stem weight gradient: dW[c][ci][kr][kc] = sum_pix dy[pix][c] * x[n, ci, 2oy-1+kr, 2ox-1+kc],
// dy = ca*g + cb*s_raw + cc (BN-backward apply of the stem BN on load).  Partials per CTA.
// A [32 channels x 27 taps] = dy^T [32 x pixels] * patches [pixels x 27] product on the warp-level tensor
// cores (mma.sync m16n8k16).  Both operands are split into bf16 hi + lo parts and three products are
// accumulated (hi*hi + hi*lo + lo*hi): ~16 bits of each operand survive, so the fp32 master gradient keeps
// its round-1 accuracy (the FFMA kernel this replaces spent ~7000 warp instructions per 64 pixels, 1.2 ms
// per step; this one ~800).  Tiles of 64 pixels are staged as [pixel][32 x bf16] rows with an 80-byte
// pitch (conflict-free ldmatrix); ldmatrix.trans pairs two adjacent PIXELS of one channel / tap in a
// register, which is exactly the k-pairing of both fragments.
// ------------------------------------------------------------------------------------------------
constexpr int kStemWgPix = 64;      // pixels per smem tile
constexpr int kStemWgPitch = 80;    // bytes per staged pixel row
constexpr int kStemWgThreads = 128;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma16816_bf16(float (&c)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                              const uint32_t a3, const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// 8 fp32 -> 8 bf16 hi parts and 8 bf16 lo parts (v - hi), one 16-byte store each
__device__ __forceinline__ void split_store8(const float (&v)[8], uint8_t* hi, uint8_t* lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
    l[i] = pack_bf16(v[2 * i] - bf16_lo(h[i]), v[2 * i + 1] - bf16_hi(h[i]));
  }
  *reinterpret_cast<uint4*>(hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(kStemWgThreads) stem_wgrad_kernel(
    const float* __restrict__ x, const __nv_bfloat16* __restrict__ g, long long ldg,
    const __nv_bfloat16* __restrict__ sraw, long long lds, const float* __restrict__ ca,
    const float* __restrict__ cb, const float* __restrict__ cc, float* part, int B, int H, int W,
    int C, int Ho, int Wo, int tiles_per_cta) {
  __shared__ __align__(16) uint8_t sm[4][kStemWgPix * kStemWgPitch];      // dy hi, dy lo, patch hi, patch lo
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pix = tid >> 1, half = tid & 1;                               // fill role: one pixel, 16 channels / taps
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  float a[16], b[16], d[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = half * 16 + i;
    a[i] = c < C ? ca[c] : 0.f; b[i] = c < C ? cb[c] : 0.f; d[i] = c < C ? cc[c] : 0.f;
  }
  float acc[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
  // ldmatrix row addresses of this lane (matrix m = lane / 8, row r = lane % 8)
  const int lm = lane >> 3, lr = lane & 7;
  const int a_off = (warp * 16 + (lm >> 1) * 8 + lr) * kStemWgPitch + (lm & 1) * 16;     // + mt * 32
  const int b_off = (warp * 16 + (lm & 1) * 8 + lr) * kStemWgPitch + (lm >> 1) * 16;     // + ntpair * 32
  uint8_t* my = reinterpret_cast<uint8_t*>(sm) + pix * kStemWgPitch + half * 32;
  constexpr int kPlane = kStemWgPix * kStemWgPitch;

  for (int t = 0; t < tiles_per_cta; ++t) {
    const long long p0 = (static_cast<long long>(blockIdx.x) * tiles_per_cta + t) * kStemWgPix;
    if (p0 >= npix) break;
    __syncthreads();                                       // previous tile's fragments are consumed
    const long long p = p0 + pix;
    float dy[16], xv[16];
    if (p < npix) {
      const uint4* gp = reinterpret_cast<const uint4*>(g + p * ldg + half * 16);
      const uint4* sp = reinterpret_cast<const uint4*>(sraw + p * lds + half * 16);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 gv = make_uint4(0u, 0u, 0u, 0u), sv = gv;
        if (half * 16 + q * 8 < C) { gv = __ldg(gp + q); sv = __ldg(sp + q); }
        const uint32_t gi[4] = {gv.x, gv.y, gv.z, gv.w}, si[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dy[q * 8 + 2 * i] = fmaf(a[q * 8 + 2 * i], bf16_lo(gi[i]), fmaf(b[q * 8 + 2 * i], bf16_lo(si[i]), d[q * 8 + 2 * i]));
          dy[q * 8 + 2 * i + 1] = fmaf(a[q * 8 + 2 * i + 1], bf16_hi(gi[i]), fmaf(b[q * 8 + 2 * i + 1], bf16_hi(si[i]), d[q * 8 + 2 * i + 1]));
        }
      }
      const int ox = static_cast<int>(p % Wo), oy = static_cast<int>((p / Wo) % Ho);
      const int n = static_cast<int>(p / (static_cast<long long>(Wo) * Ho));
      const float* xb = x + static_cast<long long>(n) * 3 * H * W;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int tap = half * 16 + i;                     // tap = ci*9 + kr*3 + kc (27 real taps, 5 zero columns)
        float v = 0.f;
        if (tap < 27) {
          const int ci = tap / 9, kr = (tap % 9) / 3, kc = tap % 3;
          const int iy = oy * 2 - 1 + kr, ix = ox * 2 - 1 + kc;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(xb + (static_cast<long long>(ci) * H + iy) * W + ix);
        }
        xv[i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) { dy[i] = 0.f; xv[i] = 0.f; }
    }
    {
      float v8[8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = dy[q * 8 + i];
        split_store8(v8, my + q * 16, my + kPlane + q * 16);
#pragma unroll
        for (int i = 0; i < 8; ++i) v8[i] = xv[q * 8 + i];
        split_store8(v8, my + 2 * kPlane + q * 16, my + 3 * kPlane + q * 16);
      }
    }
    __syncthreads();
    // this warp: pixels warp*16 .. +15 of the tile (the MMA k dimension)
    const uint8_t* base = reinterpret_cast<const uint8_t*>(sm);
    uint32_t ah[2][4], al[2][4], bh[2][4], bl[2][4];       // b*[pair]: {b0, b1} of n-tile 2*pair, {b0, b1} of n-tile 2*pair+1
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      ldmatrix_x4_trans(ah[k2], base + a_off + k2 * 32);
      ldmatrix_x4_trans(al[k2], base + kPlane + a_off + k2 * 32);
      ldmatrix_x4_trans(bh[k2], base + 2 * kPlane + b_off + k2 * 32);
      ldmatrix_x4_trans(bl[k2], base + 3 * kPlane + b_off + k2 * 32);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t h0 = bh[nt >> 1][(nt & 1) * 2], h1 = bh[nt >> 1][(nt & 1) * 2 + 1];
        const uint32_t l0 = bl[nt >> 1][(nt & 1) * 2], l1 = bl[nt >> 1][(nt & 1) * 2 + 1];
        // fragment order of ldmatrix above: ah = {a0 (ch lo, pix lo), a1 (ch hi, pix lo), a2 (ch lo, pix hi), a3 (ch hi, pix hi)}
        mma16816_bf16(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], h0, h1);
        mma16816_bf16(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], l0, l1);
        mma16816_bf16(acc[mt][nt], al[mt][0], al[mt][1], al[mt][2], al[mt][3], h0, h1);
      }
  }
  // cross-warp sum of the [32 channels][32 taps] blocks, then the CTA's partial row
  __syncthreads();
  float* red = reinterpret_cast<float*>(sm);               // [4 warps][32][32] fp32 = 16 KiB <= 20 KiB
  const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float* r0 = red + (warp * 32 + mt * 16 + gq) * 32 + nt * 8 + 2 * tq;
      r0[0] = acc[mt][nt][0]; r0[1] = acc[mt][nt][1];
      r0[8 * 32] = acc[mt][nt][2]; r0[8 * 32 + 1] = acc[mt][nt][3];
    }
  __syncthreads();
  for (int i = tid; i < 32 * 27; i += kStemWgThreads) {
    const int c = i / 27, tap = i % 27;
    if (c < C) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) v += red[(w * 32 + c) * 32 + tap];
      part[(static_cast<long long>(blockIdx.x) * C + c) * 27 + tap] = v;
    }
  }
}

__global__ void sum_rows_kernel(const float* __restrict__ part, int rows, long long n, float* dst) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double a = 0.0;
  for (int r = 0; r < rows; ++r) a += static_cast<double>(part[r * n + i]);
  dst[i] = static_cast<float>(a);
}

// ------------------------------------------------------------------------------------------------
// dropout mask: counter-based hash, keep iff u >= p
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float drop_keep_scale(unsigned long long seed, unsigned long long idx, float p) {
  if (p <= 0.f) return 1.f;
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  const float u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
  return u >= p ? 1.f / (1.f - p) : 0.f;
}

// pooled[b, c] = dropout( mean_hw act(scale*h + shift) )
__global__ void pool_fwd_kernel(const __nv_bfloat16* __restrict__ h, long long ldh,
                                const float* __restrict__ scale, const float* __restrict__ shift,
                                int act, __nv_bfloat16* pooled, long long ldp, int B, int HW, int C,
                                float drop_p, unsigned long long seed, const unsigned long long* seed_dev) {
  if (seed_dev != nullptr) seed += *seed_dev;
  const int ncg = C / 8;
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= static_cast<long long>(B) * ncg) return;
  const int b = static_cast<int>(i / ncg), ch = static_cast<int>(i % ncg) * 8;
  float sc[8], sh[8], acc[8];
  load_coef8(scale, ch, C, sc, 1.f);
  load_coef8(shift, ch, C, sh, 0.f);
#pragma unroll
  for (int c = 0; c < 8; ++c) acc[c] = 0.f;
  for (int p = 0; p < HW; ++p) {
    const uint4 v = ldg_nc_v4(h + (static_cast<long long>(b) * HW + p) * ldh + ch);
    const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      acc[2 * e] += act_fwd(fmaf(sc[2 * e], bf16_lo(vi[e]), sh[2 * e]), act);
      acc[2 * e + 1] += act_fwd(fmaf(sc[2 * e + 1], bf16_hi(vi[e]), sh[2 * e + 1]), act);
    }
  }
  const float inv = 1.f / static_cast<float>(HW);
  float o[8];
#pragma unroll
  for (int c = 0; c < 8; ++c)
    o[c] = acc[c] * inv * drop_keep_scale(seed, static_cast<unsigned long long>(b) * C + ch + c, drop_p);
  stg_v4(pooled + static_cast<long long>(b) * ldp + ch,
         make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7])));
}

// gh[b,p,c] = dpooled[b,c] * keep/(1-p) / HW * act'(scale*h+shift); partials of sum gh, sum gh*xhat.
// CTA = 32 channel groups x 8 images; one partial row per blockIdx.x (8-image group).
__global__ void __launch_bounds__(256) pool_bwd_kernel(
    const __nv_bfloat16* __restrict__ dpooled, long long ldd, const __nv_bfloat16* __restrict__ h,
    long long ldh, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ rstd, int act, __nv_bfloat16* gh,
    long long ldg, float* stats_part, int B, int HW, int C, float drop_p, unsigned long long seed,
    const unsigned long long* seed_dev) {
  __shared__ float red[256 * 16];
  if (seed_dev != nullptr) seed += *seed_dev;
  const int ncg = C / 8;
  const int lcg = threadIdx.x % 32, li = threadIdx.x / 32;
  const int cg = blockIdx.y * 32 + lcg, ch = cg * 8;
  const int b = blockIdx.x * 8 + li;
  float s1[8], s2[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) { s1[c] = 0.f; s2[c] = 0.f; }
  if (cg < ncg && b < B) {
    float sc[8], sh[8], mu[8], rs[8], dp[8];
    load_coef8(scale, ch, C, sc, 1.f);
    load_coef8(shift, ch, C, sh, 0.f);
    load_coef8(mean, ch, C, mu, 0.f);
    load_coef8(rstd, ch, C, rs, 1.f);
    const uint4 dv = ldg_nc_v4(dpooled + static_cast<long long>(b) * ldd + ch);
    const uint32_t di[4] = {dv.x, dv.y, dv.z, dv.w};
    const float inv = 1.f / static_cast<float>(HW);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      dp[2 * e] = bf16_lo(di[e]) * inv * drop_keep_scale(seed, static_cast<unsigned long long>(b) * C + ch + 2 * e, drop_p);
      dp[2 * e + 1] = bf16_hi(di[e]) * inv * drop_keep_scale(seed, static_cast<unsigned long long>(b) * C + ch + 2 * e + 1, drop_p);
    }
    for (int p = 0; p < HW; ++p) {
      const long long row = static_cast<long long>(b) * HW + p;
      const uint4 v = ldg_nc_v4(h + row * ldh + ch);
      const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float hv = (e & 1) ? bf16_hi(vi[e >> 1]) : bf16_lo(vi[e >> 1]);
        const float gq = round_bf16(dp[e] * act_bwd(fmaf(sc[e], hv, sh[e]), act));
        o[e] = gq;
        s1[e] += gq;
        s2[e] = fmaf(gq, (hv - mu[e]) * rs[e], s2[e]);
      }
      stg_v4(gh + row * ldg + ch,
             make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7])));
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    red[threadIdx.x * 16 + c] = s1[c];
    red[threadIdx.x * 16 + 8 + c] = s2[c];
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 32 * 16; o += 256) {
    const int l = o / 16, e = o % 16;
    float a = 0.f;
    for (int r = 0; r < 8; ++r) a += red[(r * 32 + l) * 16 + e];
    const int c = (blockIdx.y * 32 + l) * 8 + (e & 7);
    if (c < C) stats_part[(static_cast<long long>(blockIdx.x) * 2 + (e >> 3)) * C + c] = a;
  }
}

// ------------------------------------------------------------------------------------------------
// label-smoothed cross entropy, one warp per sample
// ------------------------------------------------------------------------------------------------
__global__ void ce_loss_kernel(const float* __restrict__ logits, long long ldl,
                               const long long* __restrict__ labels, int B, int NC, float eps,
                               float grad_scale, float* loss, __nv_bfloat16* dlogits, long long ldd,
                               int ndpad) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  const float* row = logits + static_cast<long long>(warp) * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < NC; c += 32) mx = fmaxf(mx, row[c]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float se = 0.f, sl = 0.f;
  for (int c = lane; c < NC; c += 32) {
    se += expf(row[c] - mx);
    sl += row[c];
  }
  for (int o = 16; o > 0; o >>= 1) {
    se += __shfl_xor_sync(0xffffffffu, se, o);
    sl += __shfl_xor_sync(0xffffffffu, sl, o);
  }
  const float lse = mx + logf(se);
  const int y = static_cast<int>(labels[warp]);
  // loss = -sum_c t_c * (x_c - lse), t = (1-eps)*onehot + eps/NC
  if (lane == 0 && loss != nullptr)
    loss[warp] = -((1.f - eps) * (row[y] - lse) + eps / NC * (sl - NC * lse));
  if (dlogits != nullptr) {
    __nv_bfloat16* drow = dlogits + static_cast<long long>(warp) * ldd;
    for (int c = lane; c < ndpad; c += 32) {
      float gq = 0.f;
      if (c < NC) {
        const float pc = expf(row[c] - lse);
        const float t = (c == y ? 1.f - eps : 0.f) + eps / NC;
        gq = (pc - t) * grad_scale;
      }
      drow[c] = __float2bfloat16_rn(gq);
    }
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* dst,
                                   long long ldd, long long rows, int cols, int cols_pad) {
  const long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (i >= rows * cols_pad) return;
  const long long r = i / cols_pad;
  const int c = static_cast<int>(i % cols_pad);
  dst[r * ldd + c] = __float2bfloat16_rn(c < cols ? src[r * lds + c] : 0.f);
}

}  // namespace an

using namespace an;

extern "C" int an_stem_fwd_stats_rows(int B, int Ho, int Wo) {
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  return static_cast<int>((npix + kStemThreads * kStemChunks - 1) / (kStemThreads * kStemChunks));
}

extern "C" int an_stem_fwd(const float* x, const float* w, void* y, long long ldy, float* stats_part,
                           int B, int H, int W, int C, void* stream) {
  AN_REQUIRE(x && w && y && C % 8 == 0 && C <= 32 && ldy % 8 == 0, "an_stem_fwd: C=%d must be a multiple of 8 and <= 32", C);
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  stem_fwd_kernel<<<an_stem_fwd_stats_rows(B, Ho, Wo), kStemThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      x, w, reinterpret_cast<__nv_bfloat16*>(y), stats_part, B, H, W, C, Ho, Wo, ldy);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_stem_wgrad_rows(int B, int Ho, int Wo) {
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  const long long tiles = (npix + kStemWgPix - 1) / kStemWgPix;
  long long ctas = 4LL * sm_count();
  if (ctas > tiles) ctas = tiles;
  const long long tpc = (tiles + ctas - 1) / ctas;
  return static_cast<int>((tiles + tpc - 1) / tpc);
}

extern "C" int an_stem_wgrad(const float* x, const void* g, long long ldg, const void* sraw,
                             long long lds, const float* ca, const float* cb, const float* cc,
                             float* partials, float* dw, int B, int H, int W, int C, void* stream) {
  AN_REQUIRE(x && g && sraw && ca && cb && cc && partials && dw && C % 8 == 0 && C <= 32, "an_stem_wgrad: bad arguments");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long npix = static_cast<long long>(B) * Ho * Wo;
  const long long tiles = (npix + kStemWgPix - 1) / kStemWgPix;
  const int rows = an_stem_wgrad_rows(B, Ho, Wo);
  const int tpc = static_cast<int>((tiles + rows - 1) / rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  stem_wgrad_kernel<<<rows, kStemWgThreads, 0, st>>>(x, reinterpret_cast<const __nv_bfloat16*>(g), ldg,
                                           reinterpret_cast<const __nv_bfloat16*>(sraw), lds, ca, cb, cc,
                                           partials, B, H, W, C, Ho, Wo, tpc);
  AN_CHECK_LAUNCH();
  const long long n = static_cast<long long>(C) * 27;
  sum_rows_kernel<<<static_cast<int>((n + 127) / 128), 128, 0, st>>>(partials, rows, n, dw);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_pool_fwd(const void* h, long long ldh, const float* scale, const float* shift, int act,
                           void* pooled, long long ldp, int B, int HW, int C, float drop_p,
                           unsigned long long seed, const unsigned long long* seed_dev, void* stream) {
  AN_REQUIRE(h && pooled && C % 8 == 0 && ldh % 8 == 0 && ldp % 8 == 0, "an_pool_fwd: bad arguments");
  const long long n = static_cast<long long>(B) * (C / 8);
  pool_fwd_kernel<<<static_cast<int>((n + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(h), ldh, scale, shift, act, reinterpret_cast<__nv_bfloat16*>(pooled),
      ldp, B, HW, C, drop_p, seed, seed_dev);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_pool_bwd_stats_rows(int B) { return (B + 7) / 8; }

extern "C" int an_pool_bwd(const void* dpooled, long long ldd, const void* h, long long ldh,
                           const float* scale, const float* shift, const float* mean, const float* rstd,
                           int act, void* gh, long long ldg, float* stats_part, int B, int HW, int C,
                           float drop_p, unsigned long long seed, const unsigned long long* seed_dev,
                           void* stream) {
  AN_REQUIRE(dpooled && h && gh && stats_part && C % 8 == 0, "an_pool_bwd: bad arguments");
  dim3 grid((B + 7) / 8, (C / 8 + 31) / 32);
  pool_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dpooled), ldd, reinterpret_cast<const __nv_bfloat16*>(h), ldh, scale,
      shift, mean, rstd, act, reinterpret_cast<__nv_bfloat16*>(gh), ldg, stats_part, B, HW, C, drop_p, seed, seed_dev);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_ce_loss(const float* logits, long long ldl, const long long* labels, int B, int NC,
                          float label_smoothing, float grad_scale, float* loss, void* dlogits,
                          long long ldd, int nd_pad, void* stream) {
  AN_REQUIRE(logits && labels && B > 0 && NC > 0, "an_ce_loss: bad arguments");
  const int warps_per_block = 4;
  ce_loss_kernel<<<(B + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0,
                   static_cast<cudaStream_t>(stream)>>>(logits, ldl, labels, B, NC, label_smoothing, grad_scale,
                                                        loss, reinterpret_cast<__nv_bfloat16*>(dlogits), ldd, nd_pad);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_f32_to_bf16(const float* src, long long lds, void* dst, long long ldd, long long rows,
                              int cols, int cols_pad, void* stream) {
  AN_REQUIRE(src && dst && rows > 0 && cols > 0 && cols_pad >= cols, "an_f32_to_bf16: bad arguments");
  const long long n = rows * cols_pad;
  f32_to_bf16_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, lds, reinterpret_cast<__nv_bfloat16*>(dst), ldd, rows, cols, cols_pad);
  AN_CHECK_LAUNCH();
  return 0;
}
