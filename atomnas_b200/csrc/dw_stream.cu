// Depthwise convolution kernels, "row-streaming" design (stride 1 and 2; forward, data gradient,
// weight gradient).  Reference call site: ConvBNReLU(h, h, k, stride, groups=h),
// models/mobilenet_base.py:330-336, and its autograd backward.
//
//   operand  = NHWC bf16, already materialised (activated for forward, BN-backward-applied for the
//              gradients).  Depthwise k=5/7 is FMA-issue bound on B200 (25/49 FMA per 4 bytes of
//              traffic), so the kernels spend their issue slots on packed FFMA2 and nothing else:
//   staging  = a dedicated producer warp streams the column strip of the image(s) through a ring
//              of shared-memory stages with 4-D TMA boxes {64 ch, strip width + halo, U rows, NI
//              images}; out-of-image coordinates are zero-filled by the TMA unit, which IS the
//              convolution padding.  Only the CURRENT input row is ever read from shared memory:
//   thread   = one channel PAIR (a bf16x2 word; a warp's 32 lanes = 64 consecutive channels, so
//              every shared/global access of a warp is one contiguous 128-byte line) x 4 output
//              columns x the full image height, streamed top to bottom;
//   weights  = all k*k taps of the pair in registers (float2 each);
//   math     = fma.rn.f32x2, fp32 accumulation in a register ring of the partially accumulated
//              output rows (stride 1: k rows, stride 2: (k+1)/2 rows); the unroll period equals
//              the ring period, so every ring slot is a compile-time register;
//   warps    = NWC column groups x NI images (small feature maps put several images in one CTA).
//
//   dw_conv_stream<K,S,0>  forward           y  = dwconv_S(u)                   + BN partials of y
//   dw_conv_stream<K,1,1>  data gradient s=1 ge = dwconv_1(dd, flipped) * act'  + partials for BN_e
//   dw_dgrad2_stream<K>    data gradient s=2 ge = dwconv_transposed_2(dd) * act'+ partials for BN_e
//   dw_wgrad_stream<K,S>   weight gradient   dW[c][kr][kc] = sum dd * u
#include <mutex>
#include <set>
#include <type_traits>

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
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
// bf16x2 word -> two fp32, on the ALU pipe (PRMT), leaving the FMA pipe to the FFMA2 stream
__device__ __forceinline__ float2 unpack2(uint32_t w) {
  return make_float2(__uint_as_float(__byte_perm(w, 0u, 0x1044)), __uint_as_float(__byte_perm(w, 0u, 0x3244)));
}

__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  unsigned long long ra, rb, rd;
  ra = *reinterpret_cast<const unsigned long long*>(&a);
  rb = *reinterpret_cast<const unsigned long long*>(&b);
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
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

// One finished output row of a thread: activation derivative (data gradients), bf16 rounding,
// BN partial sums of the ROUNDED values, store; the accumulators are cleared for ring reuse.
//   EBN: v *= act'(es*xe + eh);  st1 += v;  st2 += v * xhat,  xhat = xe*er + emr  (emr = -mean*rstd)
//   else                         st1 += v;  st2 += v * v
template <int TC, bool EBN, bool SWISH>
__device__ __forceinline__ void emit_row(float2 (&a)[TC], const uint32_t (&xraw)[TC], __nv_bfloat16* op,
                                         const long long ldo, const int ncols, const float2 es, const float2 eh,
                                         const float2 emr, const float2 er, const float lo, const float hi,
                                         float2& st1, float2& st2) {
#pragma unroll
  for (int j = 0; j < TC; ++j) {
    float2 v = a[j];
    a[j] = make_float2(0.f, 0.f);
    if (j < ncols) {
      float2 xe = make_float2(0.f, 0.f);
      if (EBN) {
        xe = unpack2(xraw[j]);
        v = mul2(v, act_grad2<SWISH>(fma2(es, xe, eh), lo, hi));
      }
      const uint32_t pk = pack_bf16(v.x, v.y);
      const float2 vr = unpack2(pk);
      st1 = add2(st1, vr);
      st2 = fma2(vr, EBN ? fma2(xe, er, emr) : vr, st2);
      *reinterpret_cast<uint32_t*>(op + j * ldo) = pk;
    }
  }
}

struct DwStreamParams {
  CUtensorMap tm_in;                  // streamed operand: u (forward / wgrad) or dd (data gradients)
  CUtensorMap tm_dd;                  // wgrad: dd rows, box {64, TC*nwc, R, ni}; data gradients: the producer's
                                      // raw output rows under the finishing output rows, box {64, TC*nwc, ER, ni}
  const float* w;                     // [k*k][C] tap-major fp32
  __nv_bfloat16* out;                 // forward: y; data gradient: ge
  long long ldo;
  const __nv_bfloat16* eraw;          // data gradient: producer raw output at the output position
  long long lde;
  const float *e_scale, *e_shift, *e_mean, *e_rstd;
  float* stats_part;                  // [parts][2][C]
  float* wpart;                       // wgrad: [parts][k*k][C]
  int act;
  int B, C;
  int Hs, Ws;                         // streamed tensor's spatial size
  int Ho, Wo;                         // written tensor's spatial size (wgrad: the dd size)
  int nwc, ni, PW, depth, nch, ipb, groups;
};

constexpr int kMaxThreads = 256;      // 7 compute warps + the producer warp: 2 warps per scheduler, 255 registers

template <int K, int S, int TC>
struct StreamGeom {
  static constexpr int P = (K - 1) / 2;
  static constexpr int U = (S == 1) ? K : K + 1;            // patch rows per stage == unroll period
  static constexpr int R = (S == 1) ? K : (K + 1) / 2;      // accumulator ring (output rows in flight)
  static constexpr int NX = S * (TC - 1) + K;               // input words per row for TC output columns
};

__device__ __forceinline__ int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int cfloor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
__host__ __device__ constexpr int cmod(int a, int b) { return ((a % b) + b) % b; }

// ------------------------------------------------------------------------------------------------
// forward (S = 1, 2) and stride-1 data gradient
// ------------------------------------------------------------------------------------------------
template <int K, int S, int MODE, int TC>
__global__ void __launch_bounds__(kMaxThreads, 1) dw_conv_stream(const __grid_constant__ DwStreamParams p) {
  extern __shared__ __align__(128) uint8_t dsm[];
  using G = StreamGeom<K, S, TC>;
  constexpr int P = G::P, U = G::U, R = G::R, NX = G::NX;
  const int nwc = p.nwc, ni = p.ni, ncomp = nwc * ni, D = p.depth, PW = p.PW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t img_bytes = static_cast<uint32_t>(U * PW * 128);
  const uint32_t in_bytes = img_bytes * ni;
  const uint32_t e_img = (MODE != 0) ? static_cast<uint32_t>(U * TC * nwc * 128) : 0u;    // U rows of eraw
  const uint32_t stage_bytes = in_bytes + e_img * ni;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (size_t)D * stage_bytes);
  uint64_t* empty = full + D;
  float* red = reinterpret_cast<float*>(empty + D);                     // [ncomp][2][64]
  const int cb0 = blockIdx.y * 64;
  const int ox0 = blockIdx.x * TC * nwc;
  const int g0 = blockIdx.z * p.ipb;
  const int ngroups = min(p.ipb, p.groups - g0);

  if (threadIdx.x == 0) {
    if (smem_u32(dsm) & 127u) __trap();
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncomp);
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
    if (MODE != 0) tma_prefetch_desc(&p.tm_dd);
  }
  __syncthreads();

  if (warp == ncomp) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int gi = 0; gi < ngroups; ++gi)
        for (int c = 0; c < p.nch; ++c, ++it) {
          const int st = it % D;
          mbar_wait(&empty[st], ((it / D) & 1) ^ 1);
          mbar_expect_tx(&full[st], stage_bytes);
          tma_load_4d(dsm + (size_t)st * stage_bytes, &p.tm_in, &full[st], cb0, S * ox0 - P, c * U - P,
                      (g0 + gi) * ni);
          if (MODE != 0)          // output rows c*U-(K-1) .. c*U finish in this stage: their eraw rows
            tma_load_4d(dsm + (size_t)st * stage_bytes + in_bytes, &p.tm_dd, &full[st], cb0, ox0, c * U - (K - 1),
                        (g0 + gi) * ni);
        }
    }
    return;
  }

  // -------------------------------------------------------------------- consumers
  const int slot_img = warp / nwc, g = warp % nwc;
  const int ch = cb0 + 2 * lane;
  const bool ch_ok = ch < p.C;
  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t) {
    const int src = (MODE == 0) ? t : (K * K - 1 - t);                // data gradient: flipped taps
    w[t] = ch_ok ? make_float2(__ldg(p.w + (long long)src * p.C + ch), __ldg(p.w + (long long)src * p.C + ch + 1))
                 : make_float2(0.f, 0.f);
  }
  // no producer BN (non-expanding block): es = 0, eh = 1 puts every element inside the window
  float2 es = make_float2(0.f, 0.f), eh = make_float2(1.f, 1.f), emr = es, er = es;
  const bool e_bn = (MODE != 0) && p.e_scale != nullptr;
  if (e_bn && ch_ok) {
    es = make_float2(p.e_scale[ch], p.e_scale[ch + 1]);
    eh = make_float2(p.e_shift[ch], p.e_shift[ch + 1]);
    er = make_float2(p.e_rstd[ch], p.e_rstd[ch + 1]);
    emr = make_float2(-p.e_mean[ch] * er.x, -p.e_mean[ch + 1] * er.y);
  }
  const float lo = (e_bn && (p.act == ACT_RELU || p.act == ACT_RELU6)) ? 0.f : -INFINITY;
  const float hi = (e_bn && p.act == ACT_RELU6) ? 6.f : INFINITY;
  float2 st1 = make_float2(0.f, 0.f), st2 = st1;
  float2 acc[R][TC];
#pragma unroll
  for (int s = 0; s < R; ++s)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[s][j] = make_float2(0.f, 0.f);

  const uint32_t* pbase = reinterpret_cast<const uint32_t*>(dsm) + (size_t)slot_img * (img_bytes / 4) +
                          (S * TC * g) * 32 + lane;
  const uint32_t* ebase = reinterpret_cast<const uint32_t*>(dsm + in_bytes) + (size_t)slot_img * (e_img / 4) +
                          (TC * g) * 32 + lane;
  const int erowstride = TC * nwc * 32;
  const int rowstride = PW * 32;                                        // words
  const int oxw = ox0 + TC * g;                                         // first output column of the warp
  const bool warp_live = oxw < p.Wo;
  const int ncols = min(TC, p.Wo - oxw);
  uint32_t it = 0;
  for (int gi = 0; gi < ngroups; ++gi) {
    const int n = (g0 + gi) * ni + slot_img;
    const bool n_ok = n < p.B;
    const bool live = warp_live && n_ok && ch_ok;
    for (int c = 0; c < p.nch; ++c, ++it) {
      const int st = it % D;
      mbar_wait(&full[st], (it / D) & 1);
      const uint32_t* prow = pbase + (size_t)st * (stage_bytes / 4);
      // one stage = U patch rows.  GUARD = false is the branch-free body for stages whose rows all
      // feed (and finish) valid output rows -- every stage of a large map except its first and last
      auto stage_rows = [&](auto guard_tag) {
        constexpr bool GUARD = decltype(guard_tag)::value;
        uint32_t xr[NX];
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[cc * 32];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float2 x[NX];
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) x[cc] = unpack2(xr[cc]);
          if (u + 1 < U) {                                  // software pipeline: next row's words
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[(u + 1) * rowstride + cc * 32];
          }
          // output row completed by this patch row (static in u): S*o + K-1 == c*U + u
          const bool completes = cmod(u - (K - 1), S) == 0;
          const int o_done = c * R + cfloor_div(u - (K - 1), S);
          const bool emit = completes && (!GUARD || (o_done >= 0 && o_done < p.Ho));
          uint32_t xraw[TC];
#pragma unroll
          for (int j = 0; j < TC; ++j) xraw[j] = 0u;
          if (MODE != 0 && e_bn && emit) {               // S == 1: row u of the stage's eraw box
            const uint32_t* ep = ebase + (size_t)st * (stage_bytes / 4) + u * erowstride;
#pragma unroll
            for (int j = 0; j < TC; ++j) xraw[j] = ep[j * 32];
          }
#pragma unroll
          for (int kr = 0; kr < K; ++kr) {
            if (cmod(u - kr, S) == 0) {
              const int dq = cfloor_div(u - kr, S);
              const int slot = cmod(dq, R);
              const int o = c * R + dq;
              if (!GUARD || (o >= 0 && o < p.Ho)) {
#pragma unroll
                for (int kc = 0; kc < K; ++kc)
#pragma unroll
                  for (int j = 0; j < TC; ++j)
                    acc[slot][j] = fma2(w[kr * K + kc], x[S * j + kc], acc[slot][j]);
              }
            }
          }
          if (emit) {
            const int slot = cmod(cfloor_div(u - (K - 1), S), R);
            __nv_bfloat16* op = p.out + (((long long)n * p.Ho + o_done) * p.Wo + oxw) * p.ldo + ch;
            emit_row<TC, MODE != 0, MODE == 2>(acc[slot], xraw, op, p.ldo, ncols, es, eh, emr, er, lo, hi, st1, st2);
          }
        }
      };
      if (live) {
        // every (u, kr) of the stage maps to an output row in [0, Ho)?
        const bool interior = (c * R + cfloor_div(-(K - 1), S) >= 0) && (c * R + cfloor_div(U - 1, S) < p.Ho);
        if (interior) stage_rows(std::false_type{});
        else stage_rows(std::true_type{});
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }

  if (p.stats_part != nullptr) {
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
}

// ------------------------------------------------------------------------------------------------
// stride-2 data gradient (transposed convolution), streamed over the dd rows
//   du[y][x] = sum_{kr,kc : (y+P-kr), (x+P-kc) even} dd[(y+P-kr)/2][(x+P-kc)/2] * w[kr][kc]
// ------------------------------------------------------------------------------------------------
template <int K, int TC>
struct Dgrad2Geom {
  static constexpr int P = (K - 1) / 2;
  static constexpr int TLO = -(P / 2);                       // first dd row / column offset that contributes
  static constexpr int THI = (TC - 1 + P) / 2;
  static constexpr int NX = THI - TLO + 1;                   // dd words per row for TC du columns
  static constexpr int RB = K + 1;                           // du rows in flight (ring)
  static constexpr int U = (K + 1) / 2;                      // dd rows per stage == unroll period
};

template <int K, int TC, bool SWISH>
__global__ void __launch_bounds__(kMaxThreads, 1) dw_dgrad2_stream(const __grid_constant__ DwStreamParams p) {
  extern __shared__ __align__(128) uint8_t dsm[];
  using G = Dgrad2Geom<K, TC>;
  constexpr int P = G::P, TLO = G::TLO, NX = G::NX, RB = G::RB, U = G::U;
  const int nwc = p.nwc, ni = p.ni, ncomp = nwc * ni, D = p.depth, PW = p.PW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t img_bytes = static_cast<uint32_t>(U * PW * 128);
  const uint32_t in_bytes = img_bytes * ni;
  const uint32_t e_img = static_cast<uint32_t>(RB * TC * nwc * 128);     // the RB du rows finishing per stage
  const uint32_t stage_bytes = in_bytes + e_img * ni;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (size_t)D * stage_bytes);
  uint64_t* empty = full + D;
  float* red = reinterpret_cast<float*>(empty + D);
  const int cb0 = blockIdx.y * 64;
  const int ox0 = blockIdx.x * TC * nwc;                     // du column origin of the strip (even)
  const int g0 = blockIdx.z * p.ipb;
  const int ngroups = min(p.ipb, p.groups - g0);

  if (threadIdx.x == 0) {
    if (smem_u32(dsm) & 127u) __trap();
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncomp);
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
    tma_prefetch_desc(&p.tm_dd);
  }
  __syncthreads();

  if (warp == ncomp) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int gi = 0; gi < ngroups; ++gi)
        for (int c = 0; c < p.nch; ++c, ++it) {
          const int st = it % D;
          mbar_wait(&empty[st], ((it / D) & 1) ^ 1);
          mbar_expect_tx(&full[st], stage_bytes);
          tma_load_4d(dsm + (size_t)st * stage_bytes, &p.tm_in, &full[st], cb0, ox0 / 2 + TLO, c * U + TLO,
                      (g0 + gi) * ni);
          tma_load_4d(dsm + (size_t)st * stage_bytes + in_bytes, &p.tm_dd, &full[st], cb0, ox0,
                      c * RB + 2 * TLO - P, (g0 + gi) * ni);
        }
    }
    return;
  }

  const int slot_img = warp / nwc, g = warp % nwc;
  const int ch = cb0 + 2 * lane;
  const bool ch_ok = ch < p.C;
  float2 w[K * K];
#pragma unroll
  for (int t = 0; t < K * K; ++t)
    w[t] = ch_ok ? make_float2(__ldg(p.w + (long long)t * p.C + ch), __ldg(p.w + (long long)t * p.C + ch + 1))
                 : make_float2(0.f, 0.f);
  float2 es = make_float2(0.f, 0.f), eh = make_float2(1.f, 1.f), emr = es, er = es;
  const bool e_bn = p.e_scale != nullptr;
  if (e_bn && ch_ok) {
    es = make_float2(p.e_scale[ch], p.e_scale[ch + 1]);
    eh = make_float2(p.e_shift[ch], p.e_shift[ch + 1]);
    er = make_float2(p.e_rstd[ch], p.e_rstd[ch + 1]);
    emr = make_float2(-p.e_mean[ch] * er.x, -p.e_mean[ch + 1] * er.y);
  }
  const float lo = (e_bn && (p.act == ACT_RELU || p.act == ACT_RELU6)) ? 0.f : -INFINITY;
  const float hi = (e_bn && p.act == ACT_RELU6) ? 6.f : INFINITY;
  float2 st1 = make_float2(0.f, 0.f), st2 = st1;
  float2 acc[RB][TC];
#pragma unroll
  for (int s = 0; s < RB; ++s)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[s][j] = make_float2(0.f, 0.f);

  const uint32_t* pbase = reinterpret_cast<const uint32_t*>(dsm) + (size_t)slot_img * (img_bytes / 4) +
                          ((TC / 2) * g) * 32 + lane;
  const uint32_t* ebase = reinterpret_cast<const uint32_t*>(dsm + in_bytes) + (size_t)slot_img * (e_img / 4) +
                          (TC * g) * 32 + lane;
  const int erowstride = TC * nwc * 32;
  const int rowstride = PW * 32;
  const int oxw = ox0 + TC * g;
  const bool warp_live = oxw < p.Wo;
  const int ncols = min(TC, p.Wo - oxw);
  uint32_t it = 0;
  for (int gi = 0; gi < ngroups; ++gi) {
    const int n = (g0 + gi) * ni + slot_img;
    const bool live = warp_live && (n < p.B) && ch_ok;
    for (int c = 0; c < p.nch; ++c, ++it) {
      const int st = it % D;
      mbar_wait(&full[st], (it / D) & 1);
      const uint32_t* prow = pbase + (size_t)st * (stage_bytes / 4);
      if (live) {
        uint32_t xr[NX];
#pragma unroll
        for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[cc * 32];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float2 x[NX];
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) x[cc] = unpack2(xr[cc]);
          if (u + 1 < U) {
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) xr[cc] = prow[(u + 1) * rowstride + cc * 32];
          }
          // dd row r = c*U + u + TLO feeds du rows y = 2r - P + kr; rows kr = 0, 1 complete here
          const int ybase = c * RB + 2 * u + 2 * TLO - P;              // du row of kr = 0
          uint32_t xraw[2][TC];
#pragma unroll
          for (int j = 0; j < TC; ++j) xraw[0][j] = xraw[1][j] = 0u;
          if (e_bn) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int y = ybase + e;
              if (y >= 0 && y < p.Ho) {                  // row 2u+e of the stage's eraw box
                const uint32_t* ep = ebase + (size_t)st * (stage_bytes / 4) + (2 * u + e) * erowstride;
#pragma unroll
                for (int j = 0; j < TC; ++j) xraw[e][j] = ep[j * 32];
              }
            }
          }
#pragma unroll
          for (int kr = 0; kr < K; ++kr) {
            const int y = ybase + kr;
            const int slot = cmod(2 * u + 2 * TLO - P + kr, RB);
            if (y >= 0 && y < p.Ho) {
#pragma unroll
              for (int kc = 0; kc < K; ++kc)
#pragma unroll
                for (int j = 0; j < TC; ++j)
                  if (cmod(j + P - kc, 2) == 0)
                    acc[slot][j] = fma2(w[kr * K + kc], x[(j + P - kc) / 2 - TLO], acc[slot][j]);
            }
          }
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int y = ybase + e;
            const int slot = cmod(2 * u + 2 * TLO - P + e, RB);
            if (y >= 0 && y < p.Ho) {
              __nv_bfloat16* op = p.out + (((long long)n * p.Ho + y) * p.Wo + oxw) * p.ldo + ch;
              emit_row<TC, true, SWISH>(acc[slot], xraw[e], op, p.ldo, ncols, es, eh, emr, er, lo, hi, st1, st2);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[st]);
    }
  }

  if (p.stats_part != nullptr) {
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
}

// ------------------------------------------------------------------------------------------------
// weight gradient: dW[c][kr][kc] = sum_{n,o,j} dd[n][o][j][c] * u[n][S*o + kr - P][S*j + kc - P][c]
// streamed over the u rows; the R dd rows that pair with the current u row live in registers
// ------------------------------------------------------------------------------------------------
template <int K, int S, int TC>
__global__ void __launch_bounds__(kMaxThreads, 1) dw_wgrad_stream(const __grid_constant__ DwStreamParams p) {
  extern __shared__ __align__(128) uint8_t dsm[];
  using G = StreamGeom<K, S, TC>;
  constexpr int P = G::P, U = G::U, R = G::R, NX = G::NX;
  const int nwc = p.nwc, ni = p.ni, ncomp = nwc * ni, D = p.depth, PW = p.PW, TW = TC * nwc;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t u_img = static_cast<uint32_t>(U * PW * 128), d_img = static_cast<uint32_t>(R * TW * 128);
  const uint32_t u_bytes = u_img * ni, d_bytes = d_img * ni;
  const uint32_t stage_bytes = u_bytes + d_bytes;
  const size_t red_bytes = (size_t)ncomp * K * K * 64 * 4;
  const size_t ring_bytes = (size_t)D * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(dsm + (ring_bytes > red_bytes ? ring_bytes : red_bytes));
  uint64_t* empty = full + D;
  float* red = reinterpret_cast<float*>(dsm);                           // reuses the ring after the stream
  const int cb0 = blockIdx.y * 64;
  const int ox0 = blockIdx.x * TW;                                      // dd column origin of the strip
  const int g0 = blockIdx.z * p.ipb;
  const int ngroups = min(p.ipb, p.groups - g0);

  if (threadIdx.x == 0) {
    if (smem_u32(dsm) & 127u) __trap();
    for (int s = 0; s < D; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ncomp);
    }
    mbar_fence_init();
    tma_prefetch_desc(&p.tm_in);
    tma_prefetch_desc(&p.tm_dd);
  }
  __syncthreads();

  if (warp == ncomp) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int gi = 0; gi < ngroups; ++gi)
        for (int c = 0; c < p.nch; ++c, ++it) {
          const int st = it % D;
          mbar_wait(&empty[st], ((it / D) & 1) ^ 1);
          mbar_expect_tx(&full[st], stage_bytes);
          uint8_t* sp = dsm + (size_t)st * stage_bytes;
          tma_load_4d(sp, &p.tm_in, &full[st], cb0, S * ox0 - P, c * U - P, (g0 + gi) * ni);
          tma_load_4d(sp + u_bytes, &p.tm_dd, &full[st], cb0, ox0, c * R, (g0 + gi) * ni);
        }
    }
  } else {
    const int slot_img = warp / nwc, g = warp % nwc;
    const int ch = cb0 + 2 * lane;
    const bool ch_ok = ch < p.C;
    float2 acc[K * K];
#pragma unroll
    for (int t = 0; t < K * K; ++t) acc[t] = make_float2(0.f, 0.f);
    uint32_t dr[R][TC];
#pragma unroll
    for (int s = 0; s < R; ++s)
#pragma unroll
      for (int j = 0; j < TC; ++j) dr[s][j] = 0u;
    const uint32_t* ubase = reinterpret_cast<const uint32_t*>(dsm) + (size_t)slot_img * (u_img / 4) + (S * TC * g) * 32 + lane;
    const uint32_t* dbase = reinterpret_cast<const uint32_t*>(dsm + u_bytes) + (size_t)slot_img * (d_img / 4) + (TC * g) * 32 + lane;
    const int rowstride = PW * 32, drow = TW * 32;
    const bool warp_live = (ox0 + TC * g < p.Wo) && ch_ok;
    uint32_t it = 0;
    for (int gi = 0; gi < ngroups; ++gi) {
      for (int c = 0; c < p.nch; ++c, ++it) {
        const int st = it % D;
        mbar_wait(&full[st], (it / D) & 1);
        const uint32_t* urow = ubase + (size_t)st * (stage_bytes / 4);
        const uint32_t* ddrow = dbase + (size_t)st * (stage_bytes / 4);
        auto stage_rows = [&](auto guard_tag) {        // GUARD = false: branch-free interior stage
          constexpr bool GUARD = decltype(guard_tag)::value;
          uint32_t xr[NX];
#pragma unroll
          for (int cc = 0; cc < NX; ++cc) xr[cc] = urow[cc * 32];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float2 x[NX];
#pragma unroll
            for (int cc = 0; cc < NX; ++cc) x[cc] = unpack2(xr[cc]);
            if (u + 1 < U) {
#pragma unroll
              for (int cc = 0; cc < NX; ++cc) xr[cc] = urow[(u + 1) * rowstride + cc * 32];
            }
            if (u % S == 0) {                       // dd row o = c*R + u/S enters the register ring
#pragma unroll
              for (int j = 0; j < TC; ++j) dr[u / S][j] = ddrow[(u / S) * drow + j * 32];
            }
#pragma unroll
            for (int kr = 0; kr < K; ++kr) {
              if (cmod(u - kr, S) == 0) {
                const int dq = cfloor_div(u - kr, S);
                const int slot = cmod(dq, R);
                const int o = c * R + dq;
                if (!GUARD || (o >= 0 && o < p.Ho)) {
                  float2 d[TC];
#pragma unroll
                  for (int j = 0; j < TC; ++j) d[j] = unpack2(dr[slot][j]);
#pragma unroll
                  for (int j = 0; j < TC; ++j)
#pragma unroll
                    for (int kc = 0; kc < K; ++kc)
                      acc[kr * K + kc] = fma2(d[j], x[S * j + kc], acc[kr * K + kc]);
                }
              }
            }
          }
        };
        if (warp_live) {
          const bool interior = (c * R + cfloor_div(-(K - 1), S) >= 0) && (c * R + cfloor_div(U - 1, S) < p.Ho);
          if (interior) stage_rows(std::false_type{});
          else stage_rows(std::true_type{});
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
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
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct StreamCfg {
  int tc, nwc, ni, strips, groups, ipb, gz, nch, PW, depth, threads, parts;
  size_t smem;
};

// kind 0: forward / stride-1 data gradient (Ht x Wt = output size), 1: weight gradient (dd size),
// 2: stride-2 data gradient (du size).  7 column-group warps of TC columns each cover the widths of
// every valid AtomNAS feature map (112/56/28: 4 columns per thread, 14: 2, 7: 1) without idle
// threads; other widths fall back to 4 columns per thread with a ragged last group.
static StreamCfg stream_cfg(int kind, int B, int C, int Ht, int Wt, int k, int S) {
  StreamCfg c;
  if (Wt % 28 == 0) { c.tc = 4; c.nwc = 7; }
  else if (Wt == 14) { c.tc = 2; c.nwc = 7; }
  else if (Wt == 7 && kind != 2) { c.tc = 1; c.nwc = 7; }
  else if (Wt <= 14 && kind != 2) { c.tc = 2; c.nwc = (Wt + 1) / 2; }
  else { c.tc = 4; c.nwc = (Wt + 3) / 4; if (c.nwc > 7) c.nwc = 7; }
  c.ni = 1;
  const int TW = c.tc * c.nwc;
  c.strips = (Wt + TW - 1) / TW;
  c.groups = (B + c.ni - 1) / c.ni;
  const int cblocks = (C + 63) / 64;
  int U, rows;
  size_t stage;
  if (kind == 2) {
    const int P = (k - 1) / 2, TLO = -(P / 2), THI = (c.tc - 1 + P) / 2, NX = THI - TLO + 1;
    U = (k + 1) / 2;
    c.PW = (c.tc / 2) * (c.nwc - 1) + NX;
    rows = (Ht - 1 + P) / 2 - TLO + 1;                 // dd rows TLO .. floor((H-1+P)/2)
    stage = (size_t)U * c.PW * 128 * c.ni + (size_t)(k + 1) * TW * 128 * c.ni;
  } else {
    U = (S == 1) ? k : k + 1;
    const int R = (S == 1) ? k : (k + 1) / 2;
    c.PW = S * (TW - 1) + k;
    rows = S * (Ht - 1) + k;                           // patch rows feeding Ht output rows
    stage = (size_t)U * c.PW * 128 * c.ni;
    if (kind == 1) stage += (size_t)R * TW * 128 * c.ni;
    if (kind == 3) stage += (size_t)U * TW * 128 * c.ni;           // stride-1 data gradient: eraw rows
  }
  c.nch = (rows + U - 1) / U;
  int depth = static_cast<int>((96 * 1024) / stage);
  if (depth > 6) depth = 6;
  if (depth < 2) depth = 2;
  c.depth = depth;
  size_t ring = (size_t)depth * stage;
  if (kind == 1) {
    const size_t red = (size_t)c.nwc * c.ni * k * k * 64 * 4;
    if (red > ring) ring = red;
  }
  c.smem = ring + 2 * depth * 8 + (size_t)c.nwc * c.ni * 2 * 64 * 4 + 128;
  // ~4 waves of CTAs; each CTA keeps its weights / accumulators across `ipb` image groups
  long long ctas1 = (long long)c.strips * cblocks * c.groups;
  int ipb = static_cast<int>(ctas1 / (4LL * sm_count()));
  if (ipb < 1) ipb = 1;
  if (ipb > 16) ipb = 16;
  c.ipb = ipb;
  c.gz = (c.groups + ipb - 1) / ipb;
  c.parts = c.strips * c.gz;
  c.threads = (c.nwc * c.ni + 1) * 32;
  return c;
}

static void fill_common(DwStreamParams& p, const StreamCfg& c, int B, int C) {
  p.B = B; p.C = C;
  p.nwc = c.nwc; p.ni = c.ni; p.PW = c.PW; p.depth = c.depth; p.nch = c.nch; p.ipb = c.ipb; p.groups = c.groups;
}

typedef void (*StreamKernel)(const DwStreamParams);

static int launch_stream(StreamKernel kernel, const DwStreamParams& p, const StreamCfg& c, cudaStream_t st) {
  AN_REQUIRE(kernel != nullptr, "depthwise stream kernel: unsupported (k, stride, columns) combination");
  AN_REQUIRE(c.smem <= 220 * 1024, "depthwise stream kernel: %zu B of shared memory needed", c.smem);
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

template <int K, int S, int MODE>
static StreamKernel conv_kernel(int tc) {
  switch (tc) {
    case 4: return dw_conv_stream<K, S, MODE, 4>;
    case 2: return dw_conv_stream<K, S, MODE, 2>;
    case 1: return dw_conv_stream<K, S, MODE, 1>;
  }
  return nullptr;
}
template <int K>
static StreamKernel dgrad2_kernel(int tc, bool swish) {
  switch (tc) {
    case 4: return swish ? dw_dgrad2_stream<K, 4, true> : dw_dgrad2_stream<K, 4, false>;
    case 2: return swish ? dw_dgrad2_stream<K, 2, true> : dw_dgrad2_stream<K, 2, false>;
  }
  return nullptr;
}
template <int K, int S>
static StreamKernel wgrad_kernel(int tc) {
  switch (tc) {
    case 4: return dw_wgrad_stream<K, S, 4>;
    case 2: return dw_wgrad_stream<K, S, 2>;
    case 1: return dw_wgrad_stream<K, S, 1>;
  }
  return nullptr;
}

int dw_stream_fwd_rows(int B, int C, int Ho, int Wo, int k, int stride) {
  return stream_cfg(0, B, C, Ho, Wo, k, stride).parts;
}
int dw_stream_dgrad_rows(int B, int C, int H, int W, int k, int stride) {
  return stream_cfg(stride == 1 ? 3 : 2, B, C, H, W, k, stride).parts;
}
int dw_stream_wgrad_rows(int B, int C, int Ho, int Wo, int k, int stride) {
  return stream_cfg(1, B, C, Ho, Wo, k, stride).parts;
}

int dw_stream_fwd(const an_dw_fwd_args* a, cudaStream_t st) {
  const int P = (a->k - 1) / 2, S = a->stride;
  const int Ho = (a->H + 2 * P - a->k) / S + 1, Wo = (a->W + 2 * P - a->k) / S + 1;
  const StreamCfg c = stream_cfg(0, a->B, a->C, Ho, Wo, a->k, S);
  DwStreamParams p;
  memset(&p, 0, sizeof(p));
  const int U = (S == 1) ? a->k : a->k + 1;
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->x, a->C, a->W, a->H, a->B, a->ldx, 64, c.PW, U, c.ni)) return e;
  fill_common(p, c, a->B, a->C);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->y); p.ldo = a->ldy;
  p.stats_part = a->stats_part; p.act = a->act;
  p.Hs = a->H; p.Ws = a->W; p.Ho = Ho; p.Wo = Wo;
  StreamKernel kn = nullptr;
  switch (a->k * 10 + S) {
    case 31: kn = conv_kernel<3, 1, 0>(c.tc); break;
    case 51: kn = conv_kernel<5, 1, 0>(c.tc); break;
    case 71: kn = conv_kernel<7, 1, 0>(c.tc); break;
    case 32: kn = conv_kernel<3, 2, 0>(c.tc); break;
    case 52: kn = conv_kernel<5, 2, 0>(c.tc); break;
    case 72: kn = conv_kernel<7, 2, 0>(c.tc); break;
  }
  return launch_stream(kn, p, c, st);
}

int dw_stream_bwd_data(const an_dw_bwd_args* a, cudaStream_t st) {
  const int P = (a->k - 1) / 2, S = a->stride;
  const int Ho = (a->H + 2 * P - a->k) / S + 1, Wo = (a->W + 2 * P - a->k) / S + 1;   // dd size
  const StreamCfg c = stream_cfg(S == 1 ? 3 : 2, a->B, a->C, a->H, a->W, a->k, S);
  DwStreamParams p;
  memset(&p, 0, sizeof(p));
  const int U = (S == 1) ? a->k : (a->k + 1) / 2;
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->g, a->C, Wo, Ho, a->B, a->ldg, 64, c.PW, U, c.ni)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.tc * c.nwc,
                                  S == 1 ? a->k : a->k + 1, c.ni)) return e;
  fill_common(p, c, a->B, a->C);
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->ge); p.ldo = a->ldge;
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw); p.lde = a->lde;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  p.Hs = Ho; p.Ws = Wo; p.Ho = a->H; p.Wo = a->W;
  const bool sw = a->e_scale != nullptr && a->act == ACT_SWISH;
  StreamKernel kn = nullptr;
  switch (a->k * 10 + S) {
    case 31: kn = sw ? conv_kernel<3, 1, 2>(c.tc) : conv_kernel<3, 1, 1>(c.tc); break;
    case 51: kn = sw ? conv_kernel<5, 1, 2>(c.tc) : conv_kernel<5, 1, 1>(c.tc); break;
    case 71: kn = sw ? conv_kernel<7, 1, 2>(c.tc) : conv_kernel<7, 1, 1>(c.tc); break;
    case 32: kn = dgrad2_kernel<3>(c.tc, sw); break;
    case 52: kn = dgrad2_kernel<5>(c.tc, sw); break;
    case 72: kn = dgrad2_kernel<7>(c.tc, sw); break;
  }
  return launch_stream(kn, p, c, st);
}

int dw_stream_wgrad(const an_dw_bwd_args* a, float* partials, cudaStream_t st) {
  const int P = (a->k - 1) / 2, S = a->stride;
  const int Ho = (a->H + 2 * P - a->k) / S + 1, Wo = (a->W + 2 * P - a->k) / S + 1;
  const StreamCfg c = stream_cfg(1, a->B, a->C, Ho, Wo, a->k, S);
  DwStreamParams p;
  memset(&p, 0, sizeof(p));
  const int U = (S == 1) ? a->k : a->k + 1, R = (S == 1) ? a->k : (a->k + 1) / 2;
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.PW, U, c.ni)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->g, a->C, Wo, Ho, a->B, a->ldg, 64, c.tc * c.nwc, R, c.ni)) return e;
  fill_common(p, c, a->B, a->C);
  p.wpart = partials;
  p.Hs = a->H; p.Ws = a->W; p.Ho = Ho; p.Wo = Wo;
  StreamKernel kn = nullptr;
  switch (a->k * 10 + S) {
    case 31: kn = wgrad_kernel<3, 1>(c.tc); break;
    case 51: kn = wgrad_kernel<5, 1>(c.tc); break;
    case 71: kn = wgrad_kernel<7, 1>(c.tc); break;
    case 32: kn = wgrad_kernel<3, 2>(c.tc); break;
    case 52: kn = wgrad_kernel<5, 2>(c.tc); break;
    case 72: kn = wgrad_kernel<7, 2>(c.tc); break;
  }
  return launch_stream(kn, p, c, st);
}

}  // namespace an
