// Depthwise convolution kernels, row-streaming design, round 2 (forward both strides, stride-1 data
// gradient, weight gradient both strides).  Reference call site: ConvBNReLU(h, h, k, stride,
// groups=h), models/mobilenet_base.py:330-336 (pad :133), and its autograd backward.
//
//   layout   NHWC bf16.  A warp owns 64 consecutive channels (lane = one bf16x2 channel pair, so
//            every shared / global access of a warp is one contiguous 128-byte line) x TC output
//            columns x the full image height, streamed top to bottom.
//   staging  a producer warp streams the CTA's column strip through a ring of shared-memory
//            stages with 4-D TMA boxes {64 ch, strip width + halo, U rows, 1 image}; columns
//            outside the image are zero-filled by the TMA unit (= the convolution padding).  Only
//            REAL input rows are processed (rows above / below the image contribute nothing and
//            are never multiplied); every stage runs the same branch-free code.
//   operand  forward / weight gradient read the PRODUCER's raw output and apply its BatchNorm +
//            activation on load (scale*x + shift, clamp / swish), so the activated tensor `u` is
//            never materialised in HBM (one write + two reads of the 6x-expanded tensor saved per
//            block); with scale == NULL the operand is used as is.
//   math     all k*k taps of the channel pair live in registers (float2 each); an input row is
//            read from shared memory ONCE and feeds the k (stride 1) / (k+1)/2 (stride 2) output
//            rows it touches, which are kept as a register ring of partial sums; packed
//            fma.rn.f32x2 (FFMA2), fp32 accumulation.  The unroll period equals the ring period,
//            so every ring slot is a compile-time register.
//   occupancy k=7 needs ~230 registers (1 CTA of 8 warps per SM); k=5 is capped at 128 (2 CTAs),
//            k=3 at 80 (3 CTAs), which is what lifts the HBM-bound k=3 / k=5 layers.
//
//   dw2_conv<K,S,0,TC,XF>   forward            y  = dwconv_S(act(bn(x)))            + BN partials of y
//   dw2_conv<K,1,1|2,TC,0>  data gradient s=1  ge = dwconv_1(dd, flipped) * act'(.) + partials for BN_e
//   dw2_wgrad<K,S,TC,XF>    weight gradient    dW[c][kr][kc] = sum dd * act(bn(x))
//   (stride-2 data gradient: dw_dgrad2_stream in dw_stream.cu)
#include <cstdlib>
#include <mutex>
#include <set>
#include <type_traits>

#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

int dw_stream_dgrad2_rows(int B, int C, int H, int W, int k);
int dw_stream_dgrad2(const an_dw_bwd_args* a, cudaStream_t st);
namespace dwm {                        // tensor-core path for stride 1 (dw_mma.cu)
bool use_conv(int k, int stride, int B, int W);
bool use_wgrad(int k, int stride, int B, int W);
int fwd_rows(int B, int C, int H, int W, int k);
int dgrad_rows(int B, int C, int H, int W, int k);
int wgrad_rows(int B, int C, int H, int W, int k);
int fwd(const an_dw_fwd_args* a, cudaStream_t st);
int bwd_data(const an_dw_bwd_args* a, cudaStream_t st);
int wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st);
}  // namespace dwm

namespace dw2 {

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
// bf16x2 word -> two fp32, on the ALU pipe (PRMT), leaving the FMA pipe to the FFMA2 stream
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  return make_float2(__uint_as_float(__byte_perm(w, 0u, 0x1044)), __uint_as_float(__byte_perm(w, 0u, 0x3244)));
}

// d act / d pre as a multiplier pair.  ReLU / ReLU6 / identity are the window lo < pre < hi.
template <bool SWISH>
__device__ __forceinline__ float2 act_grad2(const float2 pre, const float lo, const float hi) {
  if (SWISH) {
    const float sx = 1.f / (1.f + __expf(-pre.x)), sy = 1.f / (1.f + __expf(-pre.y));
    return make_float2(sx * (1.f + pre.x * (1.f - sx)), sy * (1.f + pre.y * (1.f - sy)));
  }
  return make_float2((pre.x > lo && pre.x < hi) ? 1.f : 0.f, (pre.y > lo && pre.y < hi) ? 1.f : 0.f);
}

// Producer BatchNorm + activation applied to a freshly loaded operand pair.
//   XF 1: ReLU (one max per lane);  XF 2: run-time activation (identity / ReLU6 clamp / Swish)
template <int XF>
__device__ __forceinline__ float2 xform(const float2 x, const float2 xs, const float2 xh, const float lo,
                                        const float hi, const bool swish) {
  float2 v = fma2(xs, x, xh);
  if (XF == 1) return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f));
  if (swish) return make_float2(v.x / (1.f + __expf(-v.x)), v.y / (1.f + __expf(-v.y)));
  return make_float2(fminf(fmaxf(v.x, lo), hi), fminf(fmaxf(v.y, lo), hi));
}

struct Params {
  CUtensorMap tm_in;                  // streamed operand: x (forward / wgrad) or dd (data gradient)
  CUtensorMap tm_dd;                  // wgrad: dd rows, box {64, TC*nwc, R, 1}; data gradient: eraw rows
  CUtensorMap tm_in2, tm_dd2;         // the same with 2-image boxes (tail channel blocks, see `tail`)
  const float* w;                     // [k*k][C] tap-major fp32
  __nv_bfloat16* out;                 // forward: y; data gradient: ge
  long long ldo;
  const __nv_bfloat16* eraw;          // data gradient: producer raw output at the output position
  long long lde;
  const float *x_scale, *x_shift;     // on-load BN of the streamed operand (forward / wgrad) or NULL
  const float *e_scale, *e_shift, *e_mean, *e_rstd;   // data gradient epilogue
  float* stats_part;                  // [parts][2][C]
  float* wpart;                       // wgrad: [parts][k*k][C]
  int act;
  int B, C;
  int H, W;                           // streamed tensor's spatial size
  int Ho, Wo;                         // written tensor's spatial size (wgrad: the dd size)
  int nwc, PW, depth, nch, ipb;
  int tail;                           // the last channel block has <= 32 channels: its CTAs put TWO images
                                      // in one warp (lanes 16-31 = the next image), halving their count
};

constexpr int kThreads = 256;         // 7 compute warps + the producer warp

template <int K, int S, int TC>
struct Geom {
  static constexpr int P = (K - 1) / 2;
  static constexpr int U = (S == 1) ? K : K + 1;            // rows per stage == unroll period
  static constexpr int R = (S == 1) ? K : (K + 1) / 2;      // accumulator ring (output rows in flight)
  static constexpr int NX = S * (TC - 1) + K;               // input words per row for TC output columns
};
__host__ __device__ constexpr int min_blocks(int K) { return K == 7 ? 1 : (K == 5 ? 2 : 3); }
__host__ __device__ constexpr int cfloor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int cmod(int a, int b) { return ((a % b) + b) % b; }

__device__ __forceinline__ void ring_setup(uint8_t* dsm, int D, size_t stage_bytes, uint64_t*& full, uint64_t*& empty,
                                           int ncomp, const CUtensorMap* t0, const CUtensorMap* t1) {
  full = reinterpret_cast<uint64_t*>(dsm + (size_t)D * stage_bytes);
  empty = full + D;
  if (threadIdx.x == 0) {
    if (smem_u32(dsm) & 127u) __trap();
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncomp);
    }
    mbar_fence_init();
    tma_prefetch_desc(t0);
    if (t1 != nullptr) tma_prefetch_desc(t1);
  }
  __syncthreads();
}

// Cross-warp sum of the per-thread statistic pairs -> one partial row of this CTA.
__device__ __forceinline__ void write_stats(const Params& p, float* red, int warp, int lane, int ncomp, int cb0,
                                            const float2 st1, const float2 st2) {
  red[(warp * 2 + 0) * 64 + 2 * lane] = st1.x;
  red[(warp * 2 + 0) * 64 + 2 * lane + 1] = st1.y;
  red[(warp * 2 + 1) * 64 + 2 * lane] = st2.x;
  red[(warp * 2 + 1) * 64 + 2 * lane + 1] = st2.y;
  named_bar_sync(1, ncomp * 32);
  for (int o = threadIdx.x; o < 128; o += ncomp * 32) {
    const int s = o / 64, cidx = o % 64;
    float a = 0.f;
    for (int wv = 0; wv < ncomp; ++wv) a += red[(wv * 2 + s) * 64 + cidx];
    if (cb0 + cidx < p.C) {
      const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
      p.stats_part[(part * 2 + s) * p.C + cb0 + cidx] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// forward (S = 1, 2) and stride-1 data gradient.  Input row r = c*U + u (r = 0 is the first image
// row); it feeds output rows o with S*o + kr - P == r; output row o is complete after input row
// S*o + P, i.e. nch*U >= S*(Ho-1) + P + 1 rows are walked, the last P of them (beyond the image)
// only finish and store rows.
// ------------------------------------------------------------------------------------------------
template <int K, int S, int MODE, int TC, int XF, int MB, bool RAGGED>
__global__ void __launch_bounds__(kThreads, MB) dw2_conv(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t dsm[];
  using G = Geom<K, S, TC>;
  constexpr int P = G::P, U = G::U, R = G::R, NX = G::NX;
  constexpr bool EBN = MODE != 0, SWISH = MODE == 2;
  const int ncomp = p.nwc, PW = p.PW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb0 = blockIdx.y * 64;
  // Tail channel block (<= 32 channels = 16 lanes): a warp carries TWO images, lanes 16-31 the
  // second one.  Such CTAs own 2*ipb images (half of them have nothing to do), load two-image TMA
  // boxes and use half the ring depth (same shared-memory footprint).
  const bool tail = p.tail != 0 && (p.C - cb0) <= 32;
  const int nim = tail ? 2 : 1;
  const int D = tail ? p.depth / 2 : p.depth;
  const uint32_t img_bytes = static_cast<uint32_t>(U * PW * 128);
  const uint32_t eimg_bytes = EBN ? static_cast<uint32_t>(U * TC * p.nwc * 128) : 0u;   // U rows of eraw under the strip
  const uint32_t in_bytes = img_bytes * nim;
  const uint32_t stage_bytes = (img_bytes + eimg_bytes) * nim;
  uint64_t *full, *empty;
  ring_setup(dsm, p.depth, img_bytes + eimg_bytes, full, empty, ncomp, tail ? &p.tm_in2 : &p.tm_in,
             EBN ? (tail ? &p.tm_dd2 : &p.tm_dd) : nullptr);
  float* red = reinterpret_cast<float*>(empty + p.depth);               // [ncomp][2][64]
  const int ox0 = blockIdx.x * TC * p.nwc;
  const int n0 = blockIdx.z * p.ipb * nim;
  const int ngrp = (n0 < p.B) ? min(p.ipb, (p.B - n0 + nim - 1) / nim) : 0;     // image (pair) iterations

  if (warp == ncomp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      const CUtensorMap* tin = tail ? &p.tm_in2 : &p.tm_in;
      const CUtensorMap* tdd = tail ? &p.tm_dd2 : &p.tm_dd;
      uint32_t it = 0;
      for (int gi = 0; gi < ngrp; ++gi)
        for (int c = 0; c < p.nch; ++c, ++it) {
          const int st = it % D;
          mbar_wait(&empty[st], ((it / D) & 1) ^ 1);
          mbar_expect_tx(&full[st], stage_bytes);
          uint8_t* sp = dsm + (size_t)st * stage_bytes;
          tma_load_4d(sp, tin, &full[st], cb0, S * ox0 - P, c * U, n0 + gi * nim);
          if (EBN)       // output rows c*U-P .. c*U+U-1-P finish in this stage: their eraw rows
            tma_load_4d(sp + in_bytes, tdd, &full[st], cb0, ox0, c * U - P, n0 + gi * nim);
        }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int half = tail ? (lane >> 4) : 0;             // which image of the pair this lane works on
  const int clane = tail ? (lane & 15) : lane;         // channel-pair index inside the block
  const int ch = cb0 + 2 * clane;
  const bool ch_ok = ch < p.C;
  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    const int src = (MODE == 0) ? t : (K * K - 1 - t);                // data gradient: flipped taps
    w[t] = ch_ok ? make_float2(__ldg(p.w + (long long)src * p.C + ch), __ldg(p.w + (long long)src * p.C + ch + 1))
                 : make_float2(0.f, 0.f);
  }
  // on-load BatchNorm + activation of the streamed operand (forward)
  float2 xs = make_float2(1.f, 1.f), xh = make_float2(0.f, 0.f);
  float xlo = -INFINITY, xhi = INFINITY;
  bool xswish = false;
  if (XF != 0) {
    if (ch_ok) {
      xs = make_float2(p.x_scale[ch], p.x_scale[ch + 1]);
      xh = make_float2(p.x_shift[ch], p.x_shift[ch + 1]);
    }
    if (p.act == ACT_RELU || p.act == ACT_RELU6) xlo = 0.f;
    if (p.act == ACT_RELU6) xhi = 6.f;
    xswish = p.act == ACT_SWISH;
  }
  // data-gradient epilogue: no producer BN (non-expanding block) = es 0, eh 1 (inside every window)
  float2 es = make_float2(0.f, 0.f), eh = make_float2(1.f, 1.f), emr = es, er = es;
  const bool e_bn = EBN && p.e_scale != nullptr;
  if (e_bn && ch_ok) {
    es = make_float2(p.e_scale[ch], p.e_scale[ch + 1]);
    eh = make_float2(p.e_shift[ch], p.e_shift[ch + 1]);
    er = make_float2(p.e_rstd[ch], p.e_rstd[ch + 1]);
    emr = make_float2(-p.e_mean[ch] * er.x, -p.e_mean[ch + 1] * er.y);
  }
  const float elo = (e_bn && (p.act == ACT_RELU || p.act == ACT_RELU6)) ? 0.f : -INFINITY;
  const float ehi = (e_bn && p.act == ACT_RELU6) ? 6.f : INFINITY;
  float2 st1 = make_float2(0.f, 0.f), st2 = st1;
  float2 acc[R][TC];
#pragma unroll
  for (int s = 0; s < R; ++s)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[s][j] = make_float2(0.f, 0.f);

  const int oxw = ox0 + TC * warp;                                      // first output column of the warp
  const int ncols = RAGGED ? min(TC, p.Wo - oxw) : TC;
  // XF 2 (identity / ReLU6 / Swish on load): input columns outside the image must stay zero after the
  // affine transform -> per-column validity bits.  XF 1 (ReLU) needs none: the tensor map fills
  // out-of-image elements with NaN and max(NaN, 0) = 0.
  uint32_t cmask = 0;
  if (XF == 2) {
#pragma unroll
    for (int cc = 0; cc < NX; ++cc) {
      const int col = S * oxw - P + cc;
      if (col >= 0 && col < p.W) cmask |= 1u << cc;
    }
  }
  const bool edge = XF == 2 && cmask != ((1u << NX) - 1u);
  const bool col_live = (oxw < p.Wo) && ch_ok;
  const uint32_t* pbase = reinterpret_cast<const uint32_t*>(dsm) + half * (img_bytes / 4) + (S * TC * warp) * 32 + clane;
  const uint32_t* ebase = reinterpret_cast<const uint32_t*>(dsm) + half * (eimg_bytes / 4) + (TC * warp) * 32 + clane;
  const int rowstride = PW * 32;                                        // words
  const int erowstride = TC * p.nwc * 32;
  const long long ldo2 = p.ldo * 2;                                     // bytes between output columns
  const long long orow2 = (long long)p.Wo * ldo2;                       // bytes per output row
  uint32_t it = 0;
  for (int gi = 0; gi < ngrp; ++gi) {
    const int n = n0 + gi * nim + half;
    const bool live = col_live && n < p.B;
    char* op = reinterpret_cast<char*>(p.out + ((long long)n * p.Ho * p.Wo + oxw) * p.ldo + ch);   // output row 0
    for (int c = 0; c < p.nch; ++c, ++it) {
      const int st = it % D;
      mbar_wait(&full[st], (it / D) & 1);
      if (live) {
        const uint32_t* prow = pbase + (size_t)st * (stage_bytes / 4);
        const uint32_t* erow = ebase + ((size_t)st * stage_bytes + in_bytes) / 4;
        uint32_t xr[NX];
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[cc * 32];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = c * U + u;
          float2 x[NX];
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) x[cc] = unpack2(xr[cc]);
          if (u + 1 < U) {                                  // software pipeline: next row's words
            prow += rowstride;
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[cc * 32];
          }
          // output row finished by this input row (static in u): S*o + P == r
          const bool completes = cmod(u - P, S) == 0;
          const int dq_done = cfloor_div(u - P, S);
          const int o_done = c * R + dq_done;
          const bool emit = completes && o_done >= 0 && o_done < p.Ho;
          uint32_t xraw[TC];
#pragma unroll
          for (int j = 0; j < TC; ++j) xraw[j] = 0u;
          if (EBN && emit) {                                // row u of the stage's eraw box
#pragma unroll
            for (int j = 0; j < TC; ++j) xraw[j] = erow[u * erowstride + j * 32];
          }
          if (r < p.H) {                                    // real image row (warp-uniform)
            if (XF != 0) {
#pragma unroll
              for (int cc = 0; cc < NX; ++cc) x[cc] = xform<XF>(x[cc], xs, xh, xlo, xhi, xswish);
              if (XF == 2 && edge) {
#pragma unroll
                for (int cc = 0; cc < NX; ++cc)
                  if (!((cmask >> cc) & 1u)) x[cc] = make_float2(0.f, 0.f);
              }
            }
#pragma unroll
            for (int kr = 0; kr < K; ++kr) {
              if (cmod(u + P - kr, S) == 0) {
                const int slot = cmod(cfloor_div(u + P - kr, S), R);
#pragma unroll
                for (int kc = 0; kc < K; ++kc)
#pragma unroll
                  for (int j = 0; j < TC; ++j)
                    acc[slot][j] = fma2(w[kr * K + kc], x[S * j + kc], acc[slot][j]);
              }
            }
          }
          if (completes) {
            const int slot = cmod(dq_done, R);
            if (emit) {
#pragma unroll
              for (int j = 0; j < TC; ++j) {
                if (!RAGGED || j < ncols) {
                  float2 v = acc[slot][j];
                  float2 xe = make_float2(0.f, 0.f);
                  if (EBN) {
                    xe = unpack2(xraw[j]);
                    v = mul2(v, act_grad2<SWISH>(fma2(es, xe, eh), elo, ehi));
                  }
                  const uint32_t pk = pack_bf16(v.x, v.y);
                  const float2 vr = unpack2(pk);
                  st1 = add2(st1, vr);
                  st2 = fma2(vr, EBN ? fma2(xe, er, emr) : vr, st2);
                  *reinterpret_cast<uint32_t*>(op + j * ldo2) = pk;
                }
              }
              op += orow2;
            }
#pragma unroll
            for (int j = 0; j < TC; ++j) acc[slot][j] = make_float2(0.f, 0.f);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (gi + 1 < ngrp) {              // partial sums of rows below the image never complete: clear
#pragma unroll
      for (int s = 0; s < R; ++s)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[s][j] = make_float2(0.f, 0.f);
    }
  }
  if (p.stats_part != nullptr) {
    if (tail) {                       // both halves of the warp hold statistics of the same channels
      st1.x += __shfl_down_sync(0xffffffffu, st1.x, 16);
      st1.y += __shfl_down_sync(0xffffffffu, st1.y, 16);
      st2.x += __shfl_down_sync(0xffffffffu, st2.x, 16);
      st2.y += __shfl_down_sync(0xffffffffu, st2.y, 16);
    }
    write_stats(p, red, warp, lane, ncomp, cb0, st1, st2);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[c][kr][kc] = sum_{n,o,j} dd[n][o][j][c] * u[n][S*o + kr - P][S*j + kc - P][c]
// streamed over the u rows r = c*U + u - P; the dd row o = (r + P) / S enters a register ring of R
// rows when r + P is a multiple of S (rows of dd beyond the image are zero-filled by TMA).
// ------------------------------------------------------------------------------------------------
template <int K, int S, int TC, int XF>
__global__ void __launch_bounds__(kThreads, ((K == 3 || K == 5) && TC == 4) ? 2 : 1) dw2_wgrad(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) uint8_t dsm[];
  using G = Geom<K, S, TC>;
  constexpr int P = G::P, U = G::U, R = G::R, NX = G::NX;
  const int ncomp = p.nwc, PW = p.PW, TW = TC * p.nwc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cb0 = blockIdx.y * 64;
  const bool tail = p.tail != 0 && (p.C - cb0) <= 32;         // two images per warp (see dw2_conv)
  const int nim = tail ? 2 : 1;
  const int D = tail ? p.depth / 2 : p.depth;
  const uint32_t u_img = static_cast<uint32_t>(U * PW * 128), d_img = static_cast<uint32_t>(R * TW * 128);
  const uint32_t u_bytes = u_img * nim;
  const uint32_t stage_bytes = (u_img + d_img) * nim;
  const size_t red_bytes = (size_t)ncomp * K * K * 64 * 4;
  const size_t ring_bytes = (size_t)p.depth * (u_img + d_img);
  // barriers live after max(ring, reduction scratch): the scratch reuses the ring after the stream
  uint8_t* bars = dsm + (ring_bytes > red_bytes ? ring_bytes : red_bytes);
  uint64_t* full = reinterpret_cast<uint64_t*>(bars);
  uint64_t* empty = full + p.depth;
  if (threadIdx.x == 0) {
    if (smem_u32(dsm) & 127u) __trap();
    for (int s = 0; s < p.depth; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncomp);
    }
    mbar_fence_init();
    tma_prefetch_desc(tail ? &p.tm_in2 : &p.tm_in);
    tma_prefetch_desc(tail ? &p.tm_dd2 : &p.tm_dd);
  }
  __syncthreads();
  float* red = reinterpret_cast<float*>(dsm);
  const int ox0 = blockIdx.x * TW;                                      // dd column origin of the strip
  const int n0 = blockIdx.z * p.ipb * nim;
  const int ngrp = (n0 < p.B) ? min(p.ipb, (p.B - n0 + nim - 1) / nim) : 0;

  if (warp == ncomp) {
    if (lane == 0) {
      const CUtensorMap* tin = tail ? &p.tm_in2 : &p.tm_in;
      const CUtensorMap* tdd = tail ? &p.tm_dd2 : &p.tm_dd;
      uint32_t it = 0;
      for (int gi = 0; gi < ngrp; ++gi)
        for (int c = 0; c < p.nch; ++c, ++it) {
          const int st = it % D;
          mbar_wait(&empty[st], ((it / D) & 1) ^ 1);
          mbar_expect_tx(&full[st], stage_bytes);
          uint8_t* sp = dsm + (size_t)st * stage_bytes;
          tma_load_4d(sp, tin, &full[st], cb0, S * ox0 - P, c * U - P, n0 + gi * nim);
          tma_load_4d(sp + u_bytes, tdd, &full[st], cb0, ox0, c * R, n0 + gi * nim);
        }
    }
    return;
  }

  const int half = tail ? (lane >> 4) : 0;
  const int clane = tail ? (lane & 15) : lane;
  const int ch = cb0 + 2 * clane;
  const bool ch_ok = ch < p.C;
  float2 xs = make_float2(1.f, 1.f), xh = make_float2(0.f, 0.f);
  float xlo = -INFINITY, xhi = INFINITY;
  bool xswish = false;
  if (XF != 0) {
    if (ch_ok) {
      xs = make_float2(p.x_scale[ch], p.x_scale[ch + 1]);
      xh = make_float2(p.x_shift[ch], p.x_shift[ch + 1]);
    }
    if (p.act == ACT_RELU || p.act == ACT_RELU6) xlo = 0.f;
    if (p.act == ACT_RELU6) xhi = 6.f;
    xswish = p.act == ACT_SWISH;
  }
  float2 acc[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);
  float2 dr[R][TC];                                  // dd rows in flight, already unpacked
#pragma unroll
  for (int s = 0; s < R; ++s)
#pragma unroll
    for (int j = 0; j < TC; ++j) dr[s][j] = make_float2(0.f, 0.f);
  const int oxw = ox0 + TC * warp;
  uint32_t cmask = 0;                      // XF 1 (ReLU): NaN-filled padding, no mask (see dw2_conv)
  if (XF == 2) {
#pragma unroll
    for (int cc = 0; cc < NX; ++cc) {
      const int col = S * oxw - P + cc;
      if (col >= 0 && col < p.W) cmask |= 1u << cc;
    }
  }
  const bool edge = XF == 2 && cmask != ((1u << NX) - 1u);
  const bool col_live = (oxw < p.Wo) && ch_ok;
  const uint32_t* ubase = reinterpret_cast<const uint32_t*>(dsm) + half * (u_img / 4) + (S * TC * warp) * 32 + clane;
  const uint32_t* dbase = reinterpret_cast<const uint32_t*>(dsm) + half * (d_img / 4) + (TC * warp) * 32 + clane;
  const int rowstride = PW * 32, drow = TW * 32;
  uint32_t it = 0;
  for (int gi = 0; gi < ngrp; ++gi) {
    const bool live = col_live && (n0 + gi * nim + half) < p.B;
    for (int c = 0; c < p.nch; ++c, ++it) {
      const int st = it % D;
      mbar_wait(&full[st], (it / D) & 1);
      if (live) {
        const uint32_t* urow = ubase + (size_t)st * (stage_bytes / 4);
        const uint32_t* ddrow = dbase + ((size_t)st * stage_bytes + u_bytes) / 4;
        uint32_t xr[NX];
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) xr[cc] = urow[cc * 32];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int r = c * U + u - P;
          float2 x[NX];
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) x[cc] = unpack2(xr[cc]);
          if (u + 1 < U) {
            urow += rowstride;
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) xr[cc] = urow[cc * 32];
          }
          if (u % S == 0) {                         // dd row o = c*R + u/S enters the register ring
#pragma unroll
            for (int j = 0; j < TC; ++j) dr[(u / S) % R][j] = unpack2(ddrow[(u / S) * drow + j * 32]);
          }
          if (r >= 0 && r < p.H) {                  // real image row (warp-uniform)
            if (XF != 0) {
#pragma unroll
              for (int cc = 0; cc < NX; ++cc) x[cc] = xform<XF>(x[cc], xs, xh, xlo, xhi, xswish);
              if (XF == 2 && edge) {
#pragma unroll
                for (int cc = 0; cc < NX; ++cc)
                  if (!((cmask >> cc) & 1u)) x[cc] = make_float2(0.f, 0.f);
              }
            }
#pragma unroll
            for (int kr = 0; kr < K; ++kr) {
              if (cmod(u - kr, S) == 0) {           // S*o + kr - P == r  <=>  o = c*R + (u - kr)/S
                const int slot = cmod(cfloor_div(u - kr, S), R);
#pragma unroll
                for (int j = 0; j < TC; ++j)
#pragma unroll
                  for (int kc = 0; kc < K; ++kc)
                    acc[kr * K + kc] = fma2(dr[slot][j], x[S * j + kc], acc[kr * K + kc]);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
    if (gi + 1 < ngrp) {                            // dd rows of this image must not pair with the next
#pragma unroll
      for (int s = 0; s < R; ++s)
#pragma unroll
        for (int j = 0; j < TC; ++j) dr[s][j] = make_float2(0.f, 0.f);
    }
  }
  if (tail) {                                       // both halves accumulated the same channels
#pragma unroll
    for (int t = 0; t < K * K; ++t) {
      acc[t].x += __shfl_down_sync(0xffffffffu, acc[t].x, 16);
      acc[t].y += __shfl_down_sync(0xffffffffu, acc[t].y, 16);
    }
  }
  // every stage has been consumed by this warp; wait for the others before reusing the ring
  named_bar_sync(1, ncomp * 32);
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    red[((size_t)warp * K * K + t) * 64 + 2 * lane] = acc[t].x;
    red[((size_t)warp * K * K + t) * 64 + 2 * lane + 1] = acc[t].y;
  }
  named_bar_sync(1, ncomp * 32);
  for (int o = threadIdx.x; o < K * K * 64; o += ncomp * 32) {
    const int t = o / 64, cidx = o % 64;
    float a = 0.f;
    for (int wv = 0; wv < ncomp; ++wv) a += red[((size_t)wv * K * K + t) * 64 + cidx];
    if (cb0 + cidx < p.C) {
      const long long part = (long long)blockIdx.z * gridDim.x + blockIdx.x;
      p.wpart[(part * K * K + t) * p.C + cb0 + cidx] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct Cfg {
  int tc, nwc, strips, ipb, gz, nch, PW, depth, threads, parts, minb;
  bool ragged, tail;
  size_t smem;
};

// AN_DW_OCC=0 selects the one-CTA-per-SM builds of the k=3 / k=5 kernels (no register cap); the
// default caps registers so that 3 / 2 CTAs are resident (A/B switch for scripts/bench_kernels.py).
static bool low_occupancy() {
  static const bool v = [] {
    const char* e = getenv("AN_DW_OCC");
    return e != nullptr && e[0] == '0';
  }();
  return v;
}

// kind 0: forward (Ht x Wt = output size), 3: stride-1 data gradient (+ eraw rows per stage),
// 1: weight gradient (dd size).
// 7 column-group warps of TC columns each cover the widths of every valid AtomNAS feature map
// (112/56/28: 4 columns per thread, 14: 2, 7: 1) without idle threads; other widths take 4 (or 2)
// columns per thread with a ragged last group.
static Cfg make_cfg(int kind, int B, int C, int Ht, int Wt, int Hin, int k, int S) {
  Cfg c;
  if (Wt % 28 == 0) { c.tc = 4; c.nwc = 7; }
  else if (Wt == 14) { c.tc = 2; c.nwc = 7; }
  else if (Wt == 7) { c.tc = 1; c.nwc = 7; }
  else if (Wt <= 14) { c.tc = 2; c.nwc = (Wt + 1) / 2; }
  else { c.tc = 4; c.nwc = (Wt + 3) / 4; if (c.nwc > 7) c.nwc = 7; }
  const int TW = c.tc * c.nwc;
  c.strips = (Wt + TW - 1) / TW;
  c.ragged = (Wt % TW) != 0;
  const int cblocks = (C + 63) / 64;
  const int P = (k - 1) / 2;
  const int U = (S == 1) ? k : k + 1, R = (S == 1) ? k : (k + 1) / 2;
  c.PW = S * (TW - 1) + k;
  size_t stage = (size_t)U * c.PW * 128;
  int rows;
  if (kind == 1) {
    stage += (size_t)R * TW * 128;
    rows = Hin + P;                                    // u rows -P .. Hin-1
    c.minb = ((k == 3 || k == 5) && c.tc == 4) ? 2 : 1;
  } else {
    rows = S * (Ht - 1) + P + 1;                       // input rows 0 .. S*(Ho-1)+P
    if (kind == 3) stage += (size_t)U * TW * 128;      // producer raw output rows under the strip
    c.minb = (low_occupancy() || c.tc != 4 || c.ragged) ? 1 : min_blocks(k);
  }
  c.nch = (rows + U - 1) / U;
  const size_t budget = (size_t)(200 * 1024) / c.minb - 4096;
  int depth = static_cast<int>(budget / stage);
  if (depth > 6) depth = 6;
  if (depth > c.nch + 1) depth = c.nch + 1;
  if (depth < 2) depth = 2;
  c.depth = depth;
  // tail channel block with <= 32 channels: two images per warp, needs >= 2 two-image stages
  const int crem = C % 64;
  c.tail = crem != 0 && crem <= 32 && depth >= 4 && B >= 2 && getenv("AN_DW_NOTAIL") == nullptr;
  size_t ring = (size_t)depth * stage;
  if (kind == 1) {
    const size_t red = (size_t)c.nwc * k * k * 64 * 4;
    if (red > ring) ring = red;
  }
  c.smem = ring + 2 * depth * 8 + (size_t)c.nwc * 2 * 64 * 4 + 128;
  // ~4 waves of resident CTAs; each CTA keeps its weights / accumulators across `ipb` images
  long long ctas1 = (long long)c.strips * cblocks * B;
  int ipb = static_cast<int>(ctas1 / (4LL * sm_count() * c.minb));
  if (ipb < 1) ipb = 1;
  if (ipb > 16) ipb = 16;
  c.ipb = ipb;
  c.gz = (B + ipb - 1) / ipb;
  c.parts = c.strips * c.gz;
  c.threads = (c.nwc + 1) * 32;
  return c;
}

typedef void (*Kernel)(const Params);

static int launch(Kernel kernel, const Params& p, const Cfg& c, cudaStream_t st) {
  AN_REQUIRE(kernel != nullptr, "depthwise kernel: unsupported (k, stride, columns) combination");
  AN_REQUIRE(c.smem <= 220 * 1024, "depthwise kernel: %zu B of shared memory needed", c.smem);
  static std::mutex mu;
  static std::set<const void*> configured;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured.count(reinterpret_cast<const void*>(kernel))) {
      AN_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      configured.insert(reinterpret_cast<const void*>(kernel));
    }
  }
  dim3 grid(c.strips, (p.C + 63) / 64, c.gz);
  kernel<<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

static void fill(Params& p, const Cfg& c, int B, int C) {
  p.B = B; p.C = C;
  p.nwc = c.nwc; p.PW = c.PW; p.depth = c.depth; p.nch = c.nch; p.ipb = c.ipb;
  p.tail = c.tail ? 1 : 0;
}

template <int K, int S, int MODE, int XF>
static Kernel conv_kernel(int tc, bool ragged) {
  if (ragged) {                      // widths that are not a multiple of the strip: guarded stores
    switch (tc) {
      case 4: return dw2_conv<K, S, MODE, 4, XF, 1, true>;
      case 2: return dw2_conv<K, S, MODE, 2, XF, 1, true>;
      case 1: return dw2_conv<K, S, MODE, 1, XF, 1, true>;
    }
    return nullptr;
  }
  switch (tc) {
    case 4: return low_occupancy() ? dw2_conv<K, S, MODE, 4, XF, 1, false> : dw2_conv<K, S, MODE, 4, XF, min_blocks(K), false>;
    case 2: return dw2_conv<K, S, MODE, 2, XF, 1, false>;
    case 1: return dw2_conv<K, S, MODE, 1, XF, 1, false>;
  }
  return nullptr;
}
template <int K, int S, int MODE>
static Kernel conv_kernel_xf(int tc, int xf, bool ragged) {
  if (MODE != 0) return conv_kernel<K, S, MODE, 0>(tc, ragged);
  switch (xf) {
    case 0: return conv_kernel<K, S, 0, 0>(tc, ragged);
    case 1: return conv_kernel<K, S, 0, 1>(tc, ragged);
    default: return conv_kernel<K, S, 0, 2>(tc, ragged);
  }
}
template <int K, int S, int XF>
static Kernel wgrad_kernel(int tc) {
  switch (tc) {
    case 4: return dw2_wgrad<K, S, 4, XF>;
    case 2: return dw2_wgrad<K, S, 2, XF>;
    case 1: return dw2_wgrad<K, S, 1, XF>;
  }
  return nullptr;
}
template <int K, int S>
static Kernel wgrad_kernel_xf(int tc, int xf) {
  switch (xf) {
    case 0: return wgrad_kernel<K, S, 0>(tc);
    case 1: return wgrad_kernel<K, S, 1>(tc);
    default: return wgrad_kernel<K, S, 2>(tc);
  }
}

static int xf_mode(const float* scale, int act) {
  if (scale == nullptr) return 0;
  return act == ACT_RELU ? 1 : 2;
}

static int fwd(const an_dw_fwd_args* a, cudaStream_t st) {
  const int P = (a->k - 1) / 2, S = a->stride;
  const int Ho = (a->H + 2 * P - a->k) / S + 1, Wo = (a->W + 2 * P - a->k) / S + 1;
  const Cfg c = make_cfg(0, a->B, a->C, Ho, Wo, a->H, a->k, S);
  Params p;
  memset(&p, 0, sizeof(p));
  const int U = (S == 1) ? a->k : a->k + 1;
  const int xf = xf_mode(a->scale, a->act);
  // ReLU on load: out-of-image elements are NaN-filled and become 0 through max(., 0)
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->x, a->C, a->W, a->H, a->B, a->ldx, 64, c.PW, U, 1, xf == 1)) return e;
  if (c.tail)
    if (int e = make_tmap_bf16_nhwc(&p.tm_in2, a->x, a->C, a->W, a->H, a->B, a->ldx, 64, c.PW, U, 2, xf == 1)) return e;
  fill(p, c, a->B, a->C);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldo = a->ldy;
  p.x_scale = a->scale; p.x_shift = a->shift;
  p.stats_part = a->stats_part; p.act = a->act;
  p.H = a->H; p.W = a->W; p.Ho = Ho; p.Wo = Wo;
  Kernel kn = nullptr;
  switch (a->k * 10 + S) {
    case 31: kn = conv_kernel_xf<3, 1, 0>(c.tc, xf, c.ragged); break;
    case 51: kn = conv_kernel_xf<5, 1, 0>(c.tc, xf, c.ragged); break;
    case 71: kn = conv_kernel_xf<7, 1, 0>(c.tc, xf, c.ragged); break;
    case 32: kn = conv_kernel_xf<3, 2, 0>(c.tc, xf, c.ragged); break;
    case 52: kn = conv_kernel_xf<5, 2, 0>(c.tc, xf, c.ragged); break;
    case 72: kn = conv_kernel_xf<7, 2, 0>(c.tc, xf, c.ragged); break;
  }
  return launch(kn, p, c, st);
}

static int bwd_data1(const an_dw_bwd_args* a, cudaStream_t st) {      // stride 1: output size == dd size
  const Cfg c = make_cfg(3, a->B, a->C, a->H, a->W, a->H, a->k, 1);
  Params p;
  memset(&p, 0, sizeof(p));
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->g, a->C, a->W, a->H, a->B, a->ldg, 64, c.PW, a->k, 1)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.tc * c.nwc, a->k, 1)) return e;
  if (c.tail) {
    if (int e = make_tmap_bf16_nhwc(&p.tm_in2, a->g, a->C, a->W, a->H, a->B, a->ldg, 64, c.PW, a->k, 2)) return e;
    if (int e = make_tmap_bf16_nhwc(&p.tm_dd2, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.tc * c.nwc, a->k, 2)) return e;
  }
  fill(p, c, a->B, a->C);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->ge); p.ldo = a->ldge;
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw); p.lde = a->lde;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  p.H = a->H; p.W = a->W; p.Ho = a->H; p.Wo = a->W;
  const bool sw = a->e_scale != nullptr && a->act == ACT_SWISH;
  Kernel kn = nullptr;
  switch (a->k) {
    case 3: kn = sw ? conv_kernel<3, 1, 2, 0>(c.tc, c.ragged) : conv_kernel<3, 1, 1, 0>(c.tc, c.ragged); break;
    case 5: kn = sw ? conv_kernel<5, 1, 2, 0>(c.tc, c.ragged) : conv_kernel<5, 1, 1, 0>(c.tc, c.ragged); break;
    case 7: kn = sw ? conv_kernel<7, 1, 2, 0>(c.tc, c.ragged) : conv_kernel<7, 1, 1, 0>(c.tc, c.ragged); break;
  }
  return launch(kn, p, c, st);
}

static int wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st) {
  const int P = (a->k - 1) / 2, S = a->stride;
  const int Ho = (a->H + 2 * P - a->k) / S + 1, Wo = (a->W + 2 * P - a->k) / S + 1;
  const Cfg c = make_cfg(1, a->B, a->C, Ho, Wo, a->H, a->k, S);
  Params p;
  memset(&p, 0, sizeof(p));
  const int U = (S == 1) ? a->k : a->k + 1, R = (S == 1) ? a->k : (a->k + 1) / 2;
  const int xf = xf_mode(a->e_scale, a->act);
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.PW, U, 1, xf == 1)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->g, a->C, Wo, Ho, a->B, a->ldg, 64, c.tc * c.nwc, R, 1)) return e;
  if (c.tail) {
    if (int e = make_tmap_bf16_nhwc(&p.tm_in2, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.PW, U, 2, xf == 1)) return e;
    if (int e = make_tmap_bf16_nhwc(&p.tm_dd2, a->g, a->C, Wo, Ho, a->B, a->ldg, 64, c.tc * c.nwc, R, 2)) return e;
  }
  fill(p, c, a->B, a->C);
  p.wpart = partials;
  p.x_scale = a->e_scale; p.x_shift = a->e_shift; p.act = a->act;
  p.H = a->H; p.W = a->W; p.Ho = Ho; p.Wo = Wo;
  Kernel kn = nullptr;
  switch (a->k * 10 + S) {
    case 31: kn = wgrad_kernel_xf<3, 1>(c.tc, xf); break;
    case 51: kn = wgrad_kernel_xf<5, 1>(c.tc, xf); break;
    case 71: kn = wgrad_kernel_xf<7, 1>(c.tc, xf); break;
    case 32: kn = wgrad_kernel_xf<3, 2>(c.tc, xf); break;
    case 52: kn = wgrad_kernel_xf<5, 2>(c.tc, xf); break;
    case 72: kn = wgrad_kernel_xf<7, 2>(c.tc, xf); break;
  }
  return launch(kn, p, c, st);
}

// dst[c][tap] (master layout [C,1,k,k]) = sum_rows part[row][tap][c]
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int rows, int taps, int C, float* dst) {
  // block = 32 consecutive (tap, channel) outputs x 8 row lanes (fixed order -> deterministic)
  __shared__ double red[8][33];
  const int ex = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ex;
  const int n = taps * C;
  double a = 0.0;
  if (i < n) {
    int r = rl;
    for (; r + 24 < rows; r += 32) {              // 4 independent loads in flight per lane
      const float v0 = part[(long long)r * n + i], v1 = part[(long long)(r + 8) * n + i];
      const float v2 = part[(long long)(r + 16) * n + i], v3 = part[(long long)(r + 24) * n + i];
      a += static_cast<double>(v0);
      a += static_cast<double>(v1);
      a += static_cast<double>(v2);
      a += static_cast<double>(v3);
    }
    for (; r < rows; r += 8) a += static_cast<double>(part[(long long)r * n + i]);
  }
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

}  // namespace dw2
}  // namespace an

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace an;

static inline void out_hw(int H, int W, int k, int s, int* Ho, int* Wo) {
  const int P = (k - 1) / 2;
  *Ho = (H + 2 * P - k) / s + 1;
  *Wo = (W + 2 * P - k) / s + 1;
}

extern "C" int an_dw_on_tensor_cores(int which, int k, int stride, int B, int W) {
  return (which == 0 ? dwm::use_conv(k, stride, B, W) : dwm::use_wgrad(k, stride, B, W)) ? 1 : 0;
}

extern "C" int an_dw_fwd_stats_rows(int B, int C, int Ho, int Wo, int k, int stride, int /*fused*/) {
  if (dwm::use_conv(k, stride, B, Wo)) return dwm::fwd_rows(B, C, Ho, Wo, k);
  return dw2::make_cfg(0, B, C, Ho, Wo, Ho * stride, k, stride).parts;
}

extern "C" int an_dw_fwd(const an_dw_fwd_args* a, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_fwd: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_fwd: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && a->ldx % 8 == 0 && a->ldy % 8 == 0, "an_dw_fwd: channels/pitches must be multiples of 8");
  AN_REQUIRE(a->scale == nullptr || a->shift != nullptr, "an_dw_fwd: scale without shift");
  if (dwm::use_conv(a->k, a->stride, a->B, a->W)) return dwm::fwd(a, static_cast<cudaStream_t>(stream));
  return dw2::fwd(a, static_cast<cudaStream_t>(stream));
}

extern "C" int an_dw_bwd_data_stats_rows(int B, int C, int H, int W, int k, int stride, int /*fused*/) {
  if (stride == 2) return dw_stream_dgrad2_rows(B, C, H, W, k);
  if (dwm::use_conv(k, stride, B, W)) return dwm::dgrad_rows(B, C, H, W, k);
  return dw2::make_cfg(3, B, C, H, W, H, k, 1).parts;
}

extern "C" int an_dw_bwd_data(const an_dw_bwd_args* a, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_bwd_data: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_bwd_data: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && a->ldg % 8 == 0 && a->lde % 8 == 0 && a->ldge % 8 == 0, "an_dw_bwd_data: bad layout");
  AN_REQUIRE(a->stride == 1 || (a->H % 2 == 0 && a->W % 2 == 0), "an_dw_bwd_data: stride 2 needs even H, W");
  AN_REQUIRE(a->ca == nullptr, "an_dw_bwd_data: pass the BN-backward-applied gradient dd in `g` (an_affine_act); "
                               "the on-load dd transform was removed with the tile kernels");
  if (a->stride == 2) return dw_stream_dgrad2(a, static_cast<cudaStream_t>(stream));
  if (dwm::use_conv(a->k, a->stride, a->B, a->W)) return dwm::bwd_data(a, static_cast<cudaStream_t>(stream));
  return dw2::bwd_data1(a, static_cast<cudaStream_t>(stream));
}

extern "C" int an_dw_bwd_weight_rows(int B, int C, int H, int W, int k, int stride, int /*fused*/) {
  int Ho, Wo;
  out_hw(H, W, k, stride, &Ho, &Wo);
  if (dwm::use_wgrad(k, stride, B, W)) return dwm::wgrad_rows(B, C, H, W, k);
  return dw2::make_cfg(1, B, C, Ho, Wo, H, k, stride).parts;
}

extern "C" int an_dw_bwd_weight(const an_dw_bwd_args* a, float* partials, float* dw, void* stream) {
  AN_REQUIRE(a->k == 3 || a->k == 5 || a->k == 7, "an_dw_bwd_weight: kernel size %d unsupported", a->k);
  AN_REQUIRE(a->stride == 1 || a->stride == 2, "an_dw_bwd_weight: stride %d unsupported", a->stride);
  AN_REQUIRE(a->C % 8 == 0 && partials && dw, "an_dw_bwd_weight: bad arguments");
  AN_REQUIRE(a->ca == nullptr, "an_dw_bwd_weight: pass the BN-backward-applied gradient dd in `g` (an_affine_act)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (int rc = dwm::use_wgrad(a->k, a->stride, a->B, a->W) ? dwm::wgrad(a, partials, st) : dw2::wgrad(a, partials, st)) return rc;
  const int taps = a->k * a->k;
  dw2::wgrad_reduce_kernel<<<(taps * a->C + 31) / 32, 256, 0, st>>>(
      partials, an_dw_bwd_weight_rows(a->B, a->C, a->H, a->W, a->k, a->stride, 0), taps, a->C, dw);
  AN_CHECK_LAUNCH();
  return 0;
}
