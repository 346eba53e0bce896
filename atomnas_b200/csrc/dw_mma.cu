// Stride-1 depthwise convolution on the warp-level tensor cores (mma.sync m16n8k16 bf16 -> f32, SASS
// HMMA).  Reference call site: ConvBNReLU(h, h, k, 1, groups=h), models/mobilenet_base.py:330-336,
// and its autograd backward.
//
// The FFMA2 row-streaming kernels (dw2.cu) top out at ~45 % of the CUDA-core FMA rate (measured
// ceiling of their instruction mix: 26 TFMA/s, scripts/probe/fma_probe.cu) -- k = 5 / 7 layers are
// FMA-bound there, not HBM-bound.  A depthwise convolution has no reduction over channels, but along
// the image width it IS a banded (Toeplitz) matrix product, and 16 images of a batch share that
// matrix:
//
//   forward / data gradient, per channel c and kernel row kh
//       D[img 0..15][w_out 0..7] += A[img 0..15][w_in 0..15] * T_kh[w_in 0..15][w_out 0..7]
//       A = one input row of 16 images (w_in = w_out0 - P + 0..15),  T_kh[i][j] = w[c][kh][i - j]
//   weight gradient, per channel c and kernel row kh
//       D[w_in 0..15][w_out 0..7] += U^T[w_in][img 0..15] * DD[img 0..15][w_out 0..7]
//       dW[c][kh][kw] = sum of the kw-th diagonal of D (i - j == kw), summed over rows / tiles / CTAs
//
// Band utilisation is k/16 (44 % at k = 7), yet HMMA sustains ~245 dense TMAC/s (11-12 clk per MMA and
// SM sub-partition next to the fragment-building instructions, scripts/probe/hmma_probe.cu) against
// 26 TFMA/s on the FMA pipe: k = 5 / 7 become HBM-bound.
//
//   work split  conv: the 16 MMA rows are 8 images x 2 ADJACENT column tiles (rows 0-7 = tile A, 8-15 =
//               tile B of the same images, so the window columns 8..15 of tile A are columns 0..7 of tile
//               B: 3 distinct operand registers instead of 4, and 8 x 23 instead of 16 x 15 pixel rows
//               fetched per 4096 outputs).  CTA = 8 images x 16 output columns x 32 channels x the full
//               height, 8 warps = 8 channel quads.  wgrad: 16 images (the MMA k) x 8 output columns x 32
//               channels.  A warp keeps its 4 channels' Toeplitz fragments (k rows x 2 registers each)
//               and a ring of k output rows (4 accumulator registers each) in registers; an input row
//               is read from shared memory ONCE and feeds the k output rows it touches.
//   staging     one image row per stage: 4-D TMA box {32 ch, 16 + k - 1 (+1 if even) columns, 1, 8
//               images}, zero / NaN filled outside the tensor (= padding; also batch / channel
//               remainders).  The TMA unit issues ~0.3 line requests per clock and SM whatever the row
//               length (measured: 32-byte and 64-byte rows stream at the same ~1240 clk per 400 / 340
//               lines), so the box is kept as narrow as the window allows.  SWIZZLE_64B + an odd pixel
//               pitch spread the 8-byte channel quad of the 16 pixels a half warp reads over 8 bank
//               groups (2-way conflicts, the minimum for 8-byte reads of one quad).  Thread 0 refills the
//               ring D - 1 steps ahead.
//   operand     forward / weight gradient apply the producer's BatchNorm + activation on load
//               (fp32 FFMA2, one rounding to bf16 = the tensor-core operand; `u` is never written to
//               HBM); the Toeplitz operand is the fp32 master tap rounded to bf16 (what an autocast
//               convolution multiplies with).  fp32 accumulation.
//   epilogue    a finished output row goes through a double-buffered shared tile and leaves as
//               16-byte stores covering whole 32-byte sectors; BN partial sums of the rounded output
//               are kept per lane and reduced once per CTA -> [parts][2][C].
//   wgrad       accumulators are the 16 x 8 correlation blocks (k of them per channel); the diagonals are
//               summed in a fixed order at the end (deterministic) -> partial row [k*k][C].
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <set>

#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {
namespace dwm {

constexpr int kThreads = 256;
constexpr int kImgs = 16;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  rc = *reinterpret_cast<const unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  return make_float2(__uint_as_float(__byte_perm(w, 0u, 0x1044)), __uint_as_float(__byte_perm(w, 0u, 0x3244)));
}
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));   // NaN operand -> the other operand
  return d;
}
__device__ __forceinline__ uint32_t min_bf16x2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("min.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ uint2 lds64(const uint8_t* p) { return *reinterpret_cast<const uint2*>(p); }
// TMA SWIZZLE_32B (m = 1) / SWIZZLE_64B (m = 3): 16-byte chunk index ^= address bits 7..
__device__ __forceinline__ uint32_t swz(uint32_t o, uint32_t m) { return o ^ (((o >> 7) & m) << 4); }

template <bool SWISH>
__device__ __forceinline__ float2 act_grad2(const float2 pre, const float lo, const float hi) {
  if (SWISH) {
    const float sx = 1.f / (1.f + __expf(-pre.x)), sy = 1.f / (1.f + __expf(-pre.y));
    return make_float2(sx * (1.f + pre.x * (1.f - sx)), sy * (1.f + pre.y * (1.f - sy)));
  }
  return make_float2((pre.x > lo && pre.x < hi) ? 1.f : 0.f, (pre.y > lo && pre.y < hi) ? 1.f : 0.f);
}

// Producer BatchNorm + activation on a bf16x2 word -> bf16x2 tensor-core operand.
//   XF 1: ReLU (NaN-filled padding becomes 0);  XF 2: identity / ReLU6 / Swish (padding masked by the caller)
template <int XF>
__device__ __forceinline__ uint32_t xform(const uint32_t w, const float2 xs, const float2 xh, const uint32_t hi6,
                                          const int clamp) {
  float2 v = fma2(xs, unpack2(w), xh);
  if (XF == 2 && clamp == 2) v = make_float2(v.x / (1.f + __expf(-v.x)), v.y / (1.f + __expf(-v.y)));   // Swish
  uint32_t pk = pack_bf16(v.x, v.y);
  if (XF == 1) return max_bf16x2(pk, 0u);
  if (clamp == 1) pk = min_bf16x2(max_bf16x2(pk, 0u), hi6);
  return pk;
}

struct Params {
  CUtensorMap tm_in;                  // streamed operand rows: x / dd (conv), producer output (wgrad)
  CUtensorMap tm_aux;                 // data gradient: eraw rows; weight gradient: dd rows
  const float* w;                     // [k*k][C] tap-major fp32
  __nv_bfloat16* out;
  long long ldo;
  const float *x_scale, *x_shift;
  const float *e_scale, *e_shift, *e_mean, *e_rstd;
  float* stats_part;                  // [parts][2][C]
  float* wpart;                       // [parts][k*k][C]
  int act, B, C, H, W;
  int depth, pwb, pwa;
  int in_bytes, stage_bytes, in_tx, aux_tx, smem_bytes;
  int ragged;
};

constexpr int kCC = 32;               // channels per CTA = 8 warps x 4
constexpr int kPix = kCC * 2;         // bytes per staged pixel
constexpr uint32_t kSwz = 3u;         // SWIZZLE_64B

__host__ __device__ constexpr int cmod(int a, int b) { return ((a % b) + b) % b; }

__device__ __forceinline__ void mma16816_z(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}

// mbarrier wait without the diagnostic printf of common.cuh (it is inlined 2 x K times here): a protocol bug
// still traps instead of hanging the GPU.
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();
}

// Shared-memory prologue of both kernels: zero the whole dynamic region (fragment loads read up to
// one pixel past a tile -- multiplied by zero Toeplitz rows, but they must be finite), barriers.
__device__ __forceinline__ void prologue(uint8_t* dsm, int bytes, uint64_t* full, uint64_t* empty, int D,
                                         const CUtensorMap* t0, const CUtensorMap* t1) {
  for (int o = threadIdx.x * 16; o < bytes; o += kThreads * 16) *reinterpret_cast<uint4*>(dsm + o) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 8);
    }
    mbar_fence_init();
    fence_proxy_async_smem();
    tma_prefetch_desc(t0);
    if (t1 != nullptr) tma_prefetch_desc(t1);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward (MODE 0) and data gradient (MODE 1: ReLU-family window, MODE 2: Swish) -- the data
// gradient is the same convolution of dd with the flipped taps, times act'(bn(eraw)).
// MMA rows 0-7: images n0 + g at output columns x0 + 0..7; rows 8-15: the same images at x0 + 8..15.
// ------------------------------------------------------------------------------------------------
template <int K, int MODE, int XF>
__global__ void __launch_bounds__(kThreads, 1) dwm_conv(const __grid_constant__ Params p) {
  extern __shared__ uint8_t dsm_raw[];
  constexpr int P = (K - 1) / 2;
  constexpr bool EBN = MODE != 0, SWISH = MODE == 2;
  constexpr int kOImg = 16 * kPix + 16;                     // staged output image pitch (bytes)
  constexpr int kOst = 8 * kOImg;
  uint8_t* dsm = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int D = p.depth;
  uint8_t* ost = dsm + (size_t)D * p.stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ost + 2 * kOst);
  uint64_t* empty = full + D;
  float* red = reinterpret_cast<float*>(empty + D);         // [8 warps][8]
  prologue(dsm, p.smem_bytes, full, empty, D, &p.tm_in, EBN ? &p.tm_aux : nullptr);
  const int c0 = blockIdx.y * kCC, x0 = blockIdx.x * 16, n0 = blockIdx.z * 8;
  const int ch = c0 + 4 * q;
  const bool q_ok = ch < p.C;
  const int nsteps = p.H + P;
  const bool producer = threadIdx.x == 0;
  const uint32_t tx = static_cast<uint32_t>(p.in_tx + (EBN ? p.aux_tx : 0));
  auto issue = [&](int s, int st) {
    uint8_t* sp = dsm + (size_t)st * p.stage_bytes;
    mbar_expect_tx(&full[st], tx);
    tma_load_4d(sp, &p.tm_in, &full[st], c0, x0 - P, s, n0);
    if (EBN) tma_load_4d(sp + p.in_bytes, &p.tm_aux, &full[st], c0, x0, s - P, n0);
  };
  if (producer)
    for (int s = 0; s < D - 1 && s < nsteps; ++s) issue(s, s);

  // Toeplitz fragments of the warp's 4 channels: B[k][n] = w[kh][k - n], k = 2t (+1) (+8), n = g
  uint32_t bf[4][K][2];
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
    for (int kh = 0; kh < K; ++kh)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int dd = 2 * t + e + 8 * h - g;
          const int tap = (MODE == 0) ? (kh * K + dd) : (K * K - 1 - (kh * K + dd));
          v[e] = (q_ok && dd >= 0 && dd < K) ? __ldg(p.w + (long long)tap * p.C + ch + c4) : 0.f;
        }
        bf[c4][kh][h] = pack_bf16(v[0], v[1]);
      }
  // on-load BatchNorm + activation (forward)
  float2 xs01 = make_float2(1.f, 1.f), xs23 = xs01, xh01 = make_float2(0.f, 0.f), xh23 = xh01;
  int xclamp = 0;
  uint32_t hi6 = 0x7f807f80u;                               // +inf, +inf
  if (XF != 0) {
    if (q_ok) {
      xs01 = make_float2(p.x_scale[ch], p.x_scale[ch + 1]); xs23 = make_float2(p.x_scale[ch + 2], p.x_scale[ch + 3]);
      xh01 = make_float2(p.x_shift[ch], p.x_shift[ch + 1]); xh23 = make_float2(p.x_shift[ch + 2], p.x_shift[ch + 3]);
    }
    xclamp = (p.act == ACT_RELU || p.act == ACT_RELU6) ? 1 : (p.act == ACT_SWISH ? 2 : 0);
    if (p.act == ACT_RELU6) hi6 = 0x40c040c0u;              // 6.0, 6.0
  }
  // data-gradient epilogue constants (no producer BN: es 0, eh 1 = inside every window)
  float2 es01 = make_float2(0.f, 0.f), es23 = es01, eh01 = make_float2(1.f, 1.f), eh23 = eh01;
  float2 er01 = es01, er23 = es01, em01 = es01, em23 = es01;
  const bool e_bn = EBN && p.e_scale != nullptr;
  if (e_bn && q_ok) {
    es01 = make_float2(p.e_scale[ch], p.e_scale[ch + 1]); es23 = make_float2(p.e_scale[ch + 2], p.e_scale[ch + 3]);
    eh01 = make_float2(p.e_shift[ch], p.e_shift[ch + 1]); eh23 = make_float2(p.e_shift[ch + 2], p.e_shift[ch + 3]);
    er01 = make_float2(p.e_rstd[ch], p.e_rstd[ch + 1]); er23 = make_float2(p.e_rstd[ch + 2], p.e_rstd[ch + 3]);
    em01 = make_float2(-p.e_mean[ch] * er01.x, -p.e_mean[ch + 1] * er01.y);
    em23 = make_float2(-p.e_mean[ch + 2] * er23.x, -p.e_mean[ch + 3] * er23.y);
  }
  const float elo = (e_bn && (p.act == ACT_RELU || p.act == ACT_RELU6)) ? 0.f : -INFINITY;
  const float ehi = (e_bn && p.act == ACT_RELU6) ? 6.f : INFINITY;

  // shared-memory offsets of this lane's pixels: image g, window columns 2t, 2t+1 (+8, +16)
  const uint32_t pE = static_cast<uint32_t>(g * p.pwb + 2 * t);
  const uint32_t oE = swz(pE * kPix + q * 8, kSwz), oO = swz((pE + 1) * kPix + q * 8, kSwz);
  // eraw (data gradient) and output staging: image g, tile columns 2t, 2t+1 (+8)
  const uint32_t pX = static_cast<uint32_t>(g * p.pwa + 2 * t);
  const uint32_t oXe = p.in_bytes + swz(pX * kPix + q * 8, kSwz), oXo = p.in_bytes + swz((pX + 1) * kPix + q * 8, kSwz);
  const uint32_t oS = static_cast<uint32_t>(g * kOImg + 2 * t * kPix + q * 8);
  // validity bits.  imask (XF 2): bit = odd | colgroup << 1 (3 groups);  omask: bit = odd | tile << 1
  uint32_t imask = 0, omask = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int col = x0 - P + 2 * t + (i & 1) + 8 * (i >> 1);
    if (col >= 0 && col < p.W && n0 + g < p.B) imask |= 1u << i;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int col = x0 + 2 * t + (i & 1) + 8 * (i >> 1);
    if (col < p.W && n0 + g < p.B) omask |= 1u << i;
  }
  const bool imasked = XF == 2 && imask != 0x3fu;
  const bool omasked = p.ragged != 0 && omask != 0xfu;

  float acc[K][4][4];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[s][c4][i] = 0.f;
  float2 st1a = make_float2(0.f, 0.f), st1b = st1a, st2a = st1a, st2b = st1a;

  // write-back of a staged row: 8 images x 16 columns x 64 bytes = 512 16-byte chunks, 2 per thread
  // (chunk id = tid, tid + 256: image id >> 6, column (id & 63) >> 2, channel octet id & 3)
  __nv_bfloat16* wb_dst[2];
  uint32_t wb_src[2];
  bool wb_ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int id = threadIdx.x + j * kThreads;
    const int img = id >> 6, within = id & 63, col = within >> 2, part = within & 3;
    wb_ok[j] = n0 + img < p.B && x0 + col < p.W && c0 + part * 8 < p.C;
    wb_dst[j] = p.out + ((long long)(n0 + img) * p.H * p.W + x0 + col) * p.ldo + c0 + part * 8;
    wb_src[j] = static_cast<uint32_t>(img * kOImg + within * 16);
  }
  const long long wb_row = (long long)p.W * p.ldo;

  int st = 0;                                                // r mod D
  uint32_t ph = 0;                                           // (r / D) & 1
  // The ring period K is unrolled so that every accumulator slot is a compile-time register.
#pragma unroll 1
  for (int cb = 0; cb < nsteps; cb += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int r = cb + u;
      if (r < nsteps) {
        if (producer) {
          const int sn = r + D - 1;
          if (sn < nsteps) {
            const int pst = (st == 0) ? D - 1 : st - 1;      // the stage step r - 1 used = where step sn goes
            if (r >= 1) wait_bar(&empty[pst], (st == 0) ? (ph ^ 1u) : ph);
            issue(sn, pst);
          }
        }
        __syncwarp();
        wait_bar(&full[st], ph);
        const uint8_t* sp = dsm + (size_t)st * p.stage_bytes;
        if (r < p.H) {
          uint32_t a[4][4];
          uint32_t x[4][3];                                  // [channel][column group]: window columns 2t, 2t+1 (+8 cg)
#pragma unroll
          for (int cg = 0; cg < 3; ++cg) {
            uint2 e = lds64(sp + oE + cg * (8 * kPix)), o = lds64(sp + oO + cg * (8 * kPix));
            if (XF != 0) {
              e.x = xform<XF>(e.x, xs01, xh01, hi6, xclamp); e.y = xform<XF>(e.y, xs23, xh23, hi6, xclamp);
              o.x = xform<XF>(o.x, xs01, xh01, hi6, xclamp); o.y = xform<XF>(o.y, xs23, xh23, hi6, xclamp);
              if (XF == 2 && imasked) {
                if (!((imask >> (cg * 2)) & 1u)) e = make_uint2(0u, 0u);
                if (!((imask >> (cg * 2 + 1)) & 1u)) o = make_uint2(0u, 0u);
              }
            }
            x[0][cg] = __byte_perm(e.x, o.x, 0x5410); x[1][cg] = __byte_perm(e.x, o.x, 0x7632);
            x[2][cg] = __byte_perm(e.y, o.y, 0x5410); x[3][cg] = __byte_perm(e.y, o.y, 0x7632);
          }
          // rows 0-7 = tile A (window columns 0..15), rows 8-15 = tile B (window columns 8..23)
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) { a[c4][0] = x[c4][0]; a[c4][1] = x[c4][1]; a[c4][2] = x[c4][1]; a[c4][3] = x[c4][2]; }
#pragma unroll
          for (int kh = 0; kh < K; ++kh) {
            const int slot = cmod(u + P - kh, K);          // output row r + P - kh; kh = 0 opens it (C = 0)
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              if (kh == 0) mma16816_z(acc[slot][c4], a[c4], bf[c4][kh][0], bf[c4][kh][1]);
              else mma16816(acc[slot][c4], a[c4], bf[c4][kh][0], bf[c4][kh][1]);
            }
          }
        }
        const int od = r - P;                                // the output row finished by this step
        const int sd = cmod(u - P, K);
        if (od >= 0) {
          uint8_t* ob = ost + (od & 1) * kOst;
#pragma unroll
          for (int i = 0; i < 4; ++i) {                      // i = odd column | tile << 1
            float2 v01 = make_float2(acc[sd][0][i], acc[sd][1][i]), v23 = make_float2(acc[sd][2][i], acc[sd][3][i]);
            float2 x01 = make_float2(0.f, 0.f), x23 = x01;
            if (EBN) {
              const uint2 e = lds64(sp + ((i & 1) ? oXo : oXe) + (i >> 1) * (8 * kPix));
              x01 = unpack2(e.x); x23 = unpack2(e.y);
              v01 = mul2(v01, act_grad2<SWISH>(fma2(es01, x01, eh01), elo, ehi));
              v23 = mul2(v23, act_grad2<SWISH>(fma2(es23, x23, eh23), elo, ehi));
            }
            uint32_t lo = pack_bf16(v01.x, v01.y), hi = pack_bf16(v23.x, v23.y);
            if (omasked && !((omask >> i) & 1u)) { lo = 0u; hi = 0u; }
            const float2 f01 = unpack2(lo), f23 = unpack2(hi);       // statistics of the STORED (rounded) tensor
            st1a = add2(st1a, f01); st1b = add2(st1b, f23);
            st2a = fma2(f01, EBN ? fma2(x01, er01, em01) : f01, st2a);
            st2b = fma2(f23, EBN ? fma2(x23, er23, em23) : f23, st2b);
            *reinterpret_cast<uint2*>(ob + oS + (i & 1) * kPix + (i >> 1) * (8 * kPix)) = make_uint2(lo, hi);
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < 2; ++j)
            if (wb_ok[j]) *reinterpret_cast<uint4*>(wb_dst[j] + od * wb_row) = *reinterpret_cast<const uint4*>(ob + wb_src[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == D) { st = 0; ph ^= 1u; }
      }
    }
  }
  if (p.stats_part != nullptr) {
    float sv[8] = {st1a.x, st1a.y, st1b.x, st1b.y, st2a.x, st2a.y, st2b.x, st2b.y};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) sv[i] += __shfl_xor_sync(0xffffffffu, sv[i], o);
    if (lane == 0)
#pragma unroll
      for (int i = 0; i < 8; ++i) red[q * 8 + i] = sv[i];
    __syncthreads();
    if (threadIdx.x < 2 * kCC) {
      const int s = threadIdx.x / kCC, cidx = threadIdx.x % kCC;
      if (c0 + cidx < p.C) {
        const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
        p.stats_part[(part * 2 + s) * p.C + c0 + cidx] = red[(cidx >> 2) * 8 + s * 4 + (cidx & 3)];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient.  Step s streams dd row s (ring slot s % K) and the operand row r = s - P; the
// pair (r, kh) meets dd row r + P - kh = s - kh.  MMA: m = window column (16), k = image (16),
// n = output column (8).
// ------------------------------------------------------------------------------------------------
template <int K, int XF>
__global__ void __launch_bounds__(kThreads, 1) dwm_wgrad(const __grid_constant__ Params p) {
  extern __shared__ uint8_t dsm_raw[];
  constexpr int P = (K - 1) / 2;
  uint8_t* dsm = dsm_raw + ((1024u - (smem_u32(dsm_raw) & 1023u)) & 1023u);
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int D = p.depth;
  const size_t ring = (size_t)D * p.stage_bytes, scr_bytes = (size_t)8 * K * 4 * 128 * 4;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (ring > scr_bytes ? ring : scr_bytes));
  uint64_t* empty = full + D;
  prologue(dsm, p.smem_bytes, full, empty, D, &p.tm_in, &p.tm_aux);
  const int c0 = blockIdx.y * kCC, x0 = blockIdx.x * 8, n0 = blockIdx.z * kImgs;
  const int ch = c0 + 4 * q;
  const bool q_ok = ch < p.C;
  const int nsteps = p.H + P;
  const bool producer = threadIdx.x == 0;
  const uint32_t tx = static_cast<uint32_t>(p.in_tx + p.aux_tx);
  auto issue = [&](int s, int st) {
    uint8_t* sp = dsm + (size_t)st * p.stage_bytes;
    mbar_expect_tx(&full[st], tx);
    tma_load_4d(sp, &p.tm_in, &full[st], c0, x0 - P, s - P, n0);
    tma_load_4d(sp + p.in_bytes, &p.tm_aux, &full[st], c0, x0, s, n0);
  };
  if (producer)
    for (int s = 0; s < D - 1 && s < nsteps; ++s) issue(s, s);

  float2 xs01 = make_float2(1.f, 1.f), xs23 = xs01, xh01 = make_float2(0.f, 0.f), xh23 = xh01;
  int xclamp = 0;
  uint32_t hi6 = 0x7f807f80u;
  if (XF != 0) {
    if (q_ok) {
      xs01 = make_float2(p.x_scale[ch], p.x_scale[ch + 1]); xs23 = make_float2(p.x_scale[ch + 2], p.x_scale[ch + 3]);
      xh01 = make_float2(p.x_shift[ch], p.x_shift[ch + 1]); xh23 = make_float2(p.x_shift[ch + 2], p.x_shift[ch + 3]);
    }
    xclamp = (p.act == ACT_RELU || p.act == ACT_RELU6) ? 1 : (p.act == ACT_SWISH ? 2 : 0);
    if (p.act == ACT_RELU6) hi6 = 0x40c040c0u;
  }
  // A = U^T: rows = window columns g / g+8, k = images 2t, 2t+1 (+8)
  const uint32_t pE = static_cast<uint32_t>(2 * t * p.pwb + g);
  const uint32_t oE = swz(pE * kPix + q * 8, kSwz), oO = swz((pE + p.pwb) * kPix + q * 8, kSwz);
  const uint32_t dI = 8u * p.pwb * kPix, dC = 8u * kPix;
  // B = DD: k = images 2t, 2t+1 (+8), n = output column g of the tile
  const uint32_t pB = static_cast<uint32_t>(2 * t * p.pwa + g);
  const uint32_t oBe = p.in_bytes + swz(pB * kPix + q * 8, kSwz), oBo = p.in_bytes + swz((pB + p.pwa) * kPix + q * 8, kSwz);
  const uint32_t dBI = 8u * p.pwa * kPix;
  uint32_t imask = 0;                                       // bit = odd image | img8 << 1 | col8 << 2
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = x0 - P + g + 8 * (i >> 2), img = n0 + 2 * t + (i & 1) + 8 * ((i >> 1) & 1);
    if (col >= 0 && col < p.W && img < p.B) imask |= 1u << i;
  }
  const bool imasked = XF == 2 && imask != 0xffu;

  float acc[K][4][4];
  uint32_t bd[K][4][2];
#pragma unroll
  for (int s = 0; s < K; ++s)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[s][c4][i] = 0.f;
      bd[s][c4][0] = 0u; bd[s][c4][1] = 0u;
    }

  int st = 0;
  uint32_t ph = 0;
#pragma unroll 1
  for (int cb = 0; cb < nsteps; cb += K) {
#pragma unroll
    for (int u = 0; u < K; ++u) {
      const int s = cb + u;
      if (s < nsteps) {
        if (producer) {
          const int sn = s + D - 1;
          if (sn < nsteps) {
            const int pst = (st == 0) ? D - 1 : st - 1;
            if (s >= 1) wait_bar(&empty[pst], (st == 0) ? (ph ^ 1u) : ph);
            issue(sn, pst);
          }
        }
        __syncwarp();
        wait_bar(&full[st], ph);
        const uint8_t* sp = dsm + (size_t)st * p.stage_bytes;
#pragma unroll
        for (int i8 = 0; i8 < 2; ++i8) {                     // dd row s -> ring slot u
          const uint2 e = lds64(sp + oBe + i8 * dBI), o = lds64(sp + oBo + i8 * dBI);
          bd[u][0][i8] = __byte_perm(e.x, o.x, 0x5410); bd[u][1][i8] = __byte_perm(e.x, o.x, 0x7632);
          bd[u][2][i8] = __byte_perm(e.y, o.y, 0x5410); bd[u][3][i8] = __byte_perm(e.y, o.y, 0x7632);
        }
        const int r = s - P;
        if (r >= 0 && r < p.H) {
          uint32_t a[4][4];
#pragma unroll
          for (int i8 = 0; i8 < 2; ++i8)
#pragma unroll
            for (int c8 = 0; c8 < 2; ++c8) {
              uint2 e = lds64(sp + oE + i8 * dI + c8 * dC), o = lds64(sp + oO + i8 * dI + c8 * dC);
              if (XF != 0) {
                e.x = xform<XF>(e.x, xs01, xh01, hi6, xclamp); e.y = xform<XF>(e.y, xs23, xh23, hi6, xclamp);
                o.x = xform<XF>(o.x, xs01, xh01, hi6, xclamp); o.y = xform<XF>(o.y, xs23, xh23, hi6, xclamp);
                if (XF == 2 && imasked) {
                  if (!((imask >> (i8 * 2 + c8 * 4)) & 1u)) e = make_uint2(0u, 0u);
                  if (!((imask >> (i8 * 2 + c8 * 4 + 1)) & 1u)) o = make_uint2(0u, 0u);
                }
              }
              const int ri = i8 * 2 + c8;                    // a0: (col g, imgs lo) a1: (col g+8, lo) a2: (g, hi) a3: (g+8, hi)
              a[0][ri] = __byte_perm(e.x, o.x, 0x5410); a[1][ri] = __byte_perm(e.x, o.x, 0x7632);
              a[2][ri] = __byte_perm(e.y, o.y, 0x5410); a[3][ri] = __byte_perm(e.y, o.y, 0x7632);
            }
#pragma unroll
          for (int kh = 0; kh < K; ++kh) {
            const int slot = cmod(u - kh, K);                // dd row s - kh
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) mma16816(acc[kh][c4], a[c4], bd[slot][c4][0], bd[slot][c4][1]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (++st == D) { st = 0; ph ^= 1u; }
      }
    }
  }
  // every warp is done with the ring: reuse it for the 16 x 8 correlation blocks, then sum diagonals
  __syncthreads();
  float* scr = reinterpret_cast<float*>(dsm);
#pragma unroll
  for (int kh = 0; kh < K; ++kh)
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      float* b = scr + (((size_t)q * K + kh) * 4 + c4) * 128;
      b[g * 8 + 2 * t] = acc[kh][c4][0]; b[g * 8 + 2 * t + 1] = acc[kh][c4][1];
      b[(g + 8) * 8 + 2 * t] = acc[kh][c4][2]; b[(g + 8) * 8 + 2 * t + 1] = acc[kh][c4][3];
    }
  __syncthreads();
  for (int i = threadIdx.x; i < K * K * kCC; i += kThreads) {
    const int tap = i / kCC, cidx = i % kCC, kh = tap / K, kw = tap % K;
    const float* b = scr + (((size_t)(cidx >> 2) * K + kh) * 4 + (cidx & 3)) * 128;
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < 8; ++n) v += b[(n + kw) * 8 + n];
    if (c0 + cidx < p.C) {
      const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
      p.wpart[(part * K * K + tap) * p.C + c0 + cidx] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Cfg {
  int pwb, pwa, depth, gx, gy, gz, parts;
  int in_bytes, aux_bytes, stage_bytes, in_tx, aux_tx, imgs;
  bool ragged;
  size_t smem;
};

static inline int align1k(int v) { return (v + 1023) & ~1023; }
static inline int odd_up(int v) { return v | 1; }

// kind 0: forward, 3: data gradient (8 images x 16 columns per CTA); 1: weight gradient (16 images x 8 columns)
static Cfg make_cfg(int kind, int B, int C, int H, int W, int k) {
  Cfg c;
  const int P = (k - 1) / 2;
  const int tile_w = (kind == 1) ? 8 : 16;
  c.imgs = (kind == 1) ? 16 : 8;
  c.pwb = odd_up(tile_w + k - 1);             // window columns, odd pixel pitch per image
  c.pwa = tile_w + 1;
  c.in_tx = c.imgs * c.pwb * kPix;
  c.in_bytes = align1k(c.in_tx);
  c.aux_tx = (kind != 0) ? c.imgs * c.pwa * kPix : 0;
  c.aux_bytes = align1k(c.aux_tx);
  c.stage_bytes = c.in_bytes + c.aux_bytes;
  const int ost = (kind != 1) ? 2 * 8 * (16 * kPix + 16) : 0;
  const int fixed = ost + 2 * 12 * 8 + 8 * 8 * 4 + 256;
  int depth = (200 * 1024 - fixed) / c.stage_bytes;
  if (depth > 10) depth = 10;
  if (depth > H + P) depth = H + P;
  if (depth < 2) depth = 2;
  c.depth = depth;
  size_t ring = (size_t)depth * c.stage_bytes;
  if (kind == 1) {
    const size_t scr = (size_t)8 * k * 4 * 128 * 4;
    if (scr > ring) ring = scr;
  }
  c.smem = ring + fixed;                      // the kernels zero exactly this much (after the 1 KB alignment slack)
  c.gx = (W + tile_w - 1) / tile_w;
  c.gy = (C + kCC - 1) / kCC;
  c.gz = (B + c.imgs - 1) / c.imgs;
  c.parts = c.gx * c.gz;
  c.ragged = (B % c.imgs) != 0 || (W % tile_w) != 0;
  return c;
}

typedef void (*Kernel)(const Params);

static int launch(Kernel kernel, const Params& p, const Cfg& c, cudaStream_t st) {
  AN_REQUIRE(kernel != nullptr, "depthwise MMA kernel: unsupported kernel size");
  AN_REQUIRE(c.smem + 1024 <= 227 * 1024, "depthwise MMA kernel: %zu B of shared memory needed", c.smem);
  static std::mutex mu;
  static std::set<const void*> configured;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured.count(reinterpret_cast<const void*>(kernel))) {
      AN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      configured.insert(reinterpret_cast<const void*>(kernel));
    }
  }
  dim3 grid(c.gx, c.gy, c.gz);
  kernel<<<grid, kThreads, c.smem + 1024, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

static void fill(Params& p, const Cfg& c, int B, int C, int H, int W) {
  p.B = B; p.C = C; p.H = H; p.W = W;
  p.depth = c.depth; p.pwb = c.pwb; p.pwa = c.pwa;
  p.in_bytes = c.in_bytes; p.stage_bytes = c.stage_bytes; p.in_tx = c.in_tx; p.aux_tx = c.aux_tx;
  p.smem_bytes = static_cast<int>(c.smem) & ~15;
  p.ragged = c.ragged ? 1 : 0;
}

static int xf_mode(const float* scale, int act) {
  if (scale == nullptr) return 0;
  return act == ACT_RELU ? 1 : 2;
}

template <int K>
static Kernel conv_kernel(int mode, int xf) {
  if (mode == 1) return dwm_conv<K, 1, 0>;
  if (mode == 2) return dwm_conv<K, 2, 0>;
  switch (xf) {
    case 0: return dwm_conv<K, 0, 0>;
    case 1: return dwm_conv<K, 0, 1>;
    default: return dwm_conv<K, 0, 2>;
  }
}
template <int K>
static Kernel wgrad_kernel(int xf) {
  switch (xf) {
    case 0: return dwm_wgrad<K, 0>;
    case 1: return dwm_wgrad<K, 1>;
    default: return dwm_wgrad<K, 2>;
  }
}

// Which shapes run here (stride 1, at least one 8-image tile).  Measured on B200 at batch 256
// (profiles/r02_dw_mma_sweep.json): the banded MMA beats the FFMA2 row-streaming kernels for k = 7 at
// widths >= 14 (forward / data gradient 1.1-1.4x, weight gradient 1.3-2.0x) and for the k = 5 weight
// gradient except at width 28; k = 3 and the 7 x 7 maps stay on dw2.cu.  AN_DW_MMA = digits forces those
// kernel sizes onto this path for every width ("0" = never) -- the A/B switch of scripts/bench_kernels.py.
static const char* forced() {
  static const char* sel = getenv("AN_DW_MMA");
  return sel;
}
static bool forced_has(int k) {
  for (const char* c = forced(); *c; ++c)
    if (*c - '0' == k) return true;
  return false;
}
bool use_conv(int k, int stride, int B, int W) {
  if (stride != 1 || B < 8) return false;
  if (forced() != nullptr) return forced_has(k);
  return k == 7 && W >= 14;
}
bool use_wgrad(int k, int stride, int B, int W) {
  if (stride != 1 || B < 8) return false;
  if (forced() != nullptr) return forced_has(k);
  return (k == 7 && W >= 7) || (k == 5 && W >= 14 && W != 28);
}

int fwd_rows(int B, int C, int H, int W, int k) { return make_cfg(0, B, C, H, W, k).parts; }
int dgrad_rows(int B, int C, int H, int W, int k) { return make_cfg(3, B, C, H, W, k).parts; }
int wgrad_rows(int B, int C, int H, int W, int k) { return make_cfg(1, B, C, H, W, k).parts; }

int fwd(const an_dw_fwd_args* a, cudaStream_t st) {
  const Cfg c = make_cfg(0, a->B, a->C, a->H, a->W, a->k);
  Params p;
  memset(&p, 0, sizeof(p));
  const int xf = xf_mode(a->scale, a->act);
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->x, a->C, a->W, a->H, a->B, a->ldx, kCC, c.pwb, 1, c.imgs, xf == 1, kPix)) return e;
  fill(p, c, a->B, a->C, a->H, a->W);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldo = a->ldy;
  p.x_scale = a->scale; p.x_shift = a->shift;
  p.stats_part = a->stats_part; p.act = a->act;
  Kernel kn = nullptr;
  switch (a->k) {
    case 3: kn = conv_kernel<3>(0, xf); break;
    case 5: kn = conv_kernel<5>(0, xf); break;
    case 7: kn = conv_kernel<7>(0, xf); break;
  }
  return launch(kn, p, c, st);
}

int bwd_data(const an_dw_bwd_args* a, cudaStream_t st) {
  const Cfg c = make_cfg(3, a->B, a->C, a->H, a->W, a->k);
  Params p;
  memset(&p, 0, sizeof(p));
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->g, a->C, a->W, a->H, a->B, a->ldg, kCC, c.pwb, 1, c.imgs, false, kPix)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_aux, a->eraw, a->C, a->W, a->H, a->B, a->lde, kCC, c.pwa, 1, c.imgs, false, kPix)) return e;
  fill(p, c, a->B, a->C, a->H, a->W);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->ge); p.ldo = a->ldge;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  const int mode = (a->e_scale != nullptr && a->act == ACT_SWISH) ? 2 : 1;
  Kernel kn = nullptr;
  switch (a->k) {
    case 3: kn = conv_kernel<3>(mode, 0); break;
    case 5: kn = conv_kernel<5>(mode, 0); break;
    case 7: kn = conv_kernel<7>(mode, 0); break;
  }
  return launch(kn, p, c, st);
}

int wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st) {
  const Cfg c = make_cfg(1, a->B, a->C, a->H, a->W, a->k);
  Params p;
  memset(&p, 0, sizeof(p));
  const int xf = xf_mode(a->e_scale, a->act);
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->eraw, a->C, a->W, a->H, a->B, a->lde, kCC, c.pwb, 1, c.imgs, xf == 1, kPix)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_aux, a->g, a->C, a->W, a->H, a->B, a->ldg, kCC, c.pwa, 1, c.imgs, false, kPix)) return e;
  fill(p, c, a->B, a->C, a->H, a->W);
  p.wpart = partials;
  p.x_scale = a->e_scale; p.x_shift = a->e_shift; p.act = a->act;
  Kernel kn = nullptr;
  switch (a->k) {
    case 3: kn = wgrad_kernel<3>(xf); break;
    case 5: kn = wgrad_kernel<5>(xf); break;
    case 7: kn = wgrad_kernel<7>(xf); break;
  }
  return launch(kn, p, c, st);
}

}  // namespace dwm
}  // namespace an
