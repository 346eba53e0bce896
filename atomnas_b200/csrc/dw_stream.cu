// Depthwise stride-2 DATA GRADIENT (transposed convolution), row-streaming design.  The forward,
// stride-1 data gradient and weight gradient kernels live in dw2.cu.  Reference call site:
// ConvBNReLU(h, h, k, 2, groups=h), models/mobilenet_base.py:330-336, autograd backward.
//
//   operand  = dd, NHWC bf16, materialised (BN-backward applied).  A producer warp streams the
//              CTA's column strip through a ring of shared-memory stages with 4-D TMA boxes;
//              thread = one channel pair x TC du columns x the full height; k*k weights in
//              registers; a ring of k+1 partially accumulated du rows; parity-matched taps only
//              (k*k/4 FMA per output); packed fma.rn.f32x2.
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

// k = 3 fits 128 registers: two CTAs per SM (4 warps per scheduler) hide the shared-memory and
// store latencies the k = 5 / 7 builds (1 CTA; k = 5 spills at 128 registers) have to eat.
__host__ __device__ constexpr int dgrad2_blocks(int K) { return K == 3 ? 2 : 1; }   // k = 5 at 2 CTAs: 0.78 -> 1.16 ms (spills)

template <int K, int TC, bool SWISH>
__global__ void __launch_bounds__(kMaxThreads, dgrad2_blocks(K)) dw_dgrad2_stream(const __grid_constant__ DwStreamParams p) {
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
// host side (stride-2 data gradient only; everything else lives in dw2.cu)
// ------------------------------------------------------------------------------------------------
struct StreamCfg {
  int tc, nwc, ni, strips, groups, ipb, gz, nch, PW, depth, threads, parts;
  size_t smem;
};

static StreamCfg dgrad2_cfg(int B, int C, int Ht, int Wt, int k) {
  StreamCfg c;
  if (Wt % 28 == 0) { c.tc = 4; c.nwc = 7; }
  else if (Wt == 14) { c.tc = 2; c.nwc = 7; }
  else { c.tc = 4; c.nwc = (Wt + 3) / 4; if (c.nwc > 7) c.nwc = 7; }
  c.ni = 1;
  const int TW = c.tc * c.nwc;
  c.strips = (Wt + TW - 1) / TW;
  c.groups = B;
  const int cblocks = (C + 63) / 64;
  const int P = (k - 1) / 2, TLO = -(P / 2), THI = (c.tc - 1 + P) / 2, NX = THI - TLO + 1;
  const int U = (k + 1) / 2;
  c.PW = (c.tc / 2) * (c.nwc - 1) + NX;
  const int rows = (Ht - 1 + P) / 2 - TLO + 1;                 // dd rows TLO .. floor((H-1+P)/2)
  const size_t stage = (size_t)U * c.PW * 128 + (size_t)(k + 1) * TW * 128;
  c.nch = (rows + U - 1) / U;
  const int minb = dgrad2_blocks(k);
  int depth = static_cast<int>(((200 * 1024) / minb - 4096) / stage);
  if (depth > 6) depth = 6;
  if (depth > c.nch + 1) depth = c.nch + 1;
  if (depth < 2) depth = 2;
  c.depth = depth;
  c.smem = (size_t)depth * stage + 2 * depth * 8 + (size_t)c.nwc * 2 * 64 * 4 + 128;
  long long ctas1 = (long long)c.strips * cblocks * c.groups;
  int ipb = static_cast<int>(ctas1 / (4LL * sm_count() * minb));
  if (ipb < 1) ipb = 1;
  if (ipb > 16) ipb = 16;
  c.ipb = ipb;
  c.gz = (c.groups + ipb - 1) / ipb;
  c.parts = c.strips * c.gz;
  c.threads = (c.nwc + 1) * 32;
  return c;
}

typedef void (*StreamKernel)(const DwStreamParams);

template <int K>
static StreamKernel dgrad2_kernel(int tc, bool swish) {
  switch (tc) {
    case 4: return swish ? dw_dgrad2_stream<K, 4, true> : dw_dgrad2_stream<K, 4, false>;
    case 2: return swish ? dw_dgrad2_stream<K, 2, true> : dw_dgrad2_stream<K, 2, false>;
  }
  return nullptr;
}

int dw_stream_dgrad2_rows(int B, int C, int H, int W, int k) { return dgrad2_cfg(B, C, H, W, k).parts; }

int dw_stream_dgrad2(const an_dw_bwd_args* a, cudaStream_t st) {
  const int Ho = a->H / 2, Wo = a->W / 2;                       // dd size (even H, W)
  const StreamCfg c = dgrad2_cfg(a->B, a->C, a->H, a->W, a->k);
  DwStreamParams p;
  memset(&p, 0, sizeof(p));
  const int U = (a->k + 1) / 2;
  if (int e = make_tmap_bf16_nhwc(&p.tm_in, a->g, a->C, Wo, Ho, a->B, a->ldg, 64, c.PW, U, c.ni)) return e;
  if (int e = make_tmap_bf16_nhwc(&p.tm_dd, a->eraw, a->C, a->W, a->H, a->B, a->lde, 64, c.tc * c.nwc, a->k + 1, c.ni)) return e;
  p.B = a->B; p.C = a->C;
  p.nwc = c.nwc; p.ni = c.ni; p.PW = c.PW; p.depth = c.depth; p.nch = c.nch; p.ipb = c.ipb; p.groups = c.groups;
  p.w = a->w;
  p.out = reinterpret_cast<__nv_bfloat16*>(a->ge); p.ldo = a->ldge;
  p.eraw = reinterpret_cast<const __nv_bfloat16*>(a->eraw); p.lde = a->lde;
  p.e_scale = a->e_scale; p.e_shift = a->e_shift; p.e_mean = a->e_mean; p.e_rstd = a->e_rstd;
  p.stats_part = a->stats_part; p.act = a->act;
  p.Hs = Ho; p.Ws = Wo; p.Ho = a->H; p.Wo = a->W;
  const bool sw = a->e_scale != nullptr && a->act == ACT_SWISH;
  StreamKernel kn = nullptr;
  switch (a->k) {
    case 3: kn = dgrad2_kernel<3>(c.tc, sw); break;
    case 5: kn = dgrad2_kernel<5>(c.tc, sw); break;
    case 7: kn = dgrad2_kernel<7>(c.tc, sw); break;
  }
  AN_REQUIRE(kn != nullptr, "depthwise stride-2 data gradient: unsupported (k, columns) combination");
  AN_REQUIRE(c.smem <= 220 * 1024, "depthwise stride-2 data gradient: %zu B of shared memory needed", c.smem);
  static std::mutex mu;
  static std::set<const void*> configured;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!configured.count(reinterpret_cast<const void*>(kn))) {
      AN_CHECK_CUDA(cudaFuncSetAttribute(kn, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      configured.insert(reinterpret_cast<const void*>(kn));
    }
  }
  dim3 grid(c.strips, (p.C + 63) / 64, c.gz);
  kn<<<grid, c.threads, c.smem, st>>>(p);
  AN_CHECK_LAUNCH();
  return 0;
}

}  // namespace an
