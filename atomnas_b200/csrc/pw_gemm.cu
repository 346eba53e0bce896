// Pointwise (1x1) convolutions of the AtomNAS atomic block as tcgen05 tensor-core GEMMs.
//
//   pw_gemm_kernel  : D[M,N] = Aeff[M,K] * W[N,K]^T        (forward conv, and dgrad with W^T)
//   pw_wgrad_kernel : dW[cb,cs] = sum_m Peff[m,cb] * q[m,cs] (weight gradient, contraction over
//                     pixels, both operands MN-major straight out of the NHWC activations)
//
// Reference call sites replaced: nn.Conv2d(k=1) in ConvBNReLU (models/mobilenet_base.py:135-141,
// used at :315-320), the linear projection (:338) with the python branch `sum` (:378), the head
// conv (models/mobilenet_supernet.py:148-153) and nn.Linear (:160-163), plus their autograd
// backward.  Training-mode BatchNorm (:142, :342) is split around the GEMM: the producer layer's
// BN-apply + activation runs as a prologue on the A tile in shared memory, the statistics of this
// layer's output are reduced in the epilogue (see DESIGN.md "kernel pattern").
//
// Structure (both kernels): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2-5 = operand prologue (in-place transform of the 128B-swizzled
// tile), warps 6-9 = epilogue (tcgen05.ld -> registers -> global).  mbarrier pipelines:
// full/xf/empty per smem stage, tfull/tempty per TMEM accumulator buffer.
#include <cstdlib>

#include "common.cuh"
#include "../../include/atomnas_b200.h"

namespace an {

constexpr int kBM = 128;            // rows (pixels) per tile == UMMA M
constexpr int kBK = 64;             // bf16 elements per 128-byte swizzle row
constexpr int kMaxBN = 128;         // max UMMA N of the forward/dgrad kernel
constexpr int kTileA = kBM * kBK * 2;      // 16 KiB
constexpr int kTileB = kMaxBN * kBK * 2;   // 16 KiB
constexpr int kThreads = 320;            // weight-gradient kernel
constexpr int kGemmThreads = 448;        // forward / dgrad kernel: + 4 more operand-prologue warps
constexpr int kMaxGroupCols = 1024;        // columns whose BN partials one CTA keeps in smem
constexpr int kTmemCols = 512;             // 2 x 128 accumulator columns + 128 + 16 statistics columns
constexpr int kTmemStat1 = 256;            // D1 = Y^T Y (or G^T R): the diagonal holds the second moments
constexpr int kTmemStat2 = 384;            // D2 = Y^T 1: column sums
constexpr int kGroupBytes = kBM * 128;     // one staged 64-column group of an output tile: 128 rows x 128 B
constexpr int kMaxStaging = 4;            // staged output (and operand) tiles a CTA may keep in flight
constexpr int kBarBytes = 384;
constexpr int kAuxBytes = kBarBytes + 4 * 2 * kMaxBN * 4 + 2 * kMaxGroupCols * 4 + kMaxBN * 8;

// Diagnostic build only (-DAN_GEMM_TRACE): per-tile clock stamps of the TMA producer, the store / statistics thread
// and epilogue warp 6 of CTA (0,0), printed by that CTA when it finishes.
#ifdef AN_GEMM_TRACE
__device__ unsigned int g_gemm_trace[3][64][8];
#define AN_GT(role, slot) do { if (blockIdx.x == 0 && blockIdx.y == 0 && tc < 64) g_gemm_trace[role][tc][slot] = static_cast<unsigned int>(clock64()); } while (0)
#else
#define AN_GT(role, slot) do { } while (0)
#endif

struct GemmDev {
  CUtensorMap tm_x1, tm_x2, tm_w, tm_y, tm_r;
  int M, N, K;
  int block_n, stages, m_tiles, n_tiles, tiles_per_group;
  int a_mode, a_act;
  const float *a1, *a2, *a3;
  int epi_mode;
  void* y;
  long long ldy;
  int y_is_f32;
  const float* bias;
  float* stats_part;
  const __nv_bfloat16* r;
  long long ldr;
  const float *r_scale, *r_shift, *r_mean, *r_rstd;
  int r_act;
  // staged epilogue (bf16 output, no bias): tile -> swizzled smem -> TMA store; BN partials by MMA
  int staged, nbuf, ytile_bytes, off_y, off_r, off_ones, off_aux;
  int resident, off_w;   // weight tiles of the CTA's column group stay in shared memory (plain A operand, K <= 64)
};

// In-place prologue on one 16-byte chunk (8 channels) of a swizzled operand tile.
__device__ __forceinline__ uint4 transform_chunk(uint4 v, uint4 w, int mode, int act,
                                                 const float* c1, const float* c2,
                                                 const float* c3) {
  const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
  const uint32_t wi[4] = {w.x, w.y, w.z, w.w};
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float lo = fmaf(c1[2 * i], bf16_lo(vi[i]), c3[2 * i]);
    float hi = fmaf(c1[2 * i + 1], bf16_hi(vi[i]), c3[2 * i + 1]);
    if (mode == 2) {
      lo = fmaf(c2[2 * i], bf16_lo(wi[i]), lo);
      hi = fmaf(c2[2 * i + 1], bf16_hi(wi[i]), hi);
    } else {
      lo = act_fwd(lo, act);
      hi = act_fwd(hi, act);
    }
    o[i] = pack_bf16(lo, hi);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(kGemmThreads, 1) pw_gemm_kernel(const __grid_constant__ GemmDev p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KiB alignment by OFFSET (not by pointer casts): the compiler keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int a_bytes = (p.a_mode == 2) ? 2 * kTileA : kTileA;
  const int stage_bytes = a_bytes + kTileB;
  uint8_t* aux = smem + p.off_aux;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* xf = full + 8;
  uint64_t* empty = xf + 8;
  uint64_t* tfull = empty + 8;
  uint64_t* tempty = tfull + 2;
  uint64_t* sfull = tempty + 2;
  uint64_t* rfull = sfull + 1;             // up to kMaxStaging staged tiles in flight (p.nbuf)
  uint64_t* rempty = rfull + kMaxStaging;
  uint64_t* yfull = rempty + kMaxStaging;  // staged output tile written by every epilogue thread
  uint64_t* yempty = yfull + kMaxStaging;  // ... and free again (its TMA store has read it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(yempty + kMaxStaging);
  float* stat_warp = reinterpret_cast<float*>(aux + kBarBytes);   // [4][2][kMaxBN] (direct epilogue only)
  float* stat_acc = stat_warp + 4 * 2 * kMaxBN;                   // [2][kMaxGroupCols]
  float2* colpar = reinterpret_cast<float2*>(stat_acc + 2 * kMaxGroupCols);   // [kMaxBN]
  const bool has_r = p.staged && p.r != nullptr && p.epi_mode != 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_x1);
    tma_prefetch_desc(&p.tm_w);
    if (p.a_mode == 2) tma_prefetch_desc(&p.tm_x2);
    if (p.staged) tma_prefetch_desc(&p.tm_y);
    if (has_r) tma_prefetch_desc(&p.tm_r);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&xf[s], p.resident ? 1 : 8);   // one elected arrival per prologue warp (resident mode: xf[0] = weights landed)
        mbar_init(&empty[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull[b], 1);
        mbar_init(&tempty[b], p.a_mode == 0 ? 8 : 4);   // one elected arrival per epilogue warp
      }
      for (int b = 0; b < kMaxStaging; ++b) {
        mbar_init(&rfull[b], 1);
        mbar_init(&rempty[b], 1);
        mbar_init(&yfull[b], p.a_mode == 0 ? 8 : 4);
        mbar_init(&yempty[b], 1);
      }
      mbar_init(sfull, 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kTmemCols);
  }
  for (int i = threadIdx.x; i < 2 * kMaxGroupCols; i += kGemmThreads) stat_acc[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int nt_begin = blockIdx.y * p.tiles_per_group;
  const int nt_end = min(p.n_tiles, nt_begin + p.tiles_per_group);
  const int kblocks = (p.K + kBK - 1) / kBK;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && p.resident) {
      // Resident mode (expand forward / projection dgrad of the wide layers: K <= 64, plain A operand).  The
      // TMA unit sustains only ~0.3 box rows per clock (measured), and a 128 x 128 tile costs
      // 256 store rows (+ 256 rows of the saved raw tensor): reloading the 128 A rows and 128 weight rows
      // for every column tile of an M tile was 40 % of the rows.  Here the group's weight tiles are
      // loaded once per CTA and the A tile once per M tile.
      uint32_t it = 0, tc = 0;
      const int ntl = nt_end - nt_begin;
      mbar_expect_tx(&xf[0], static_cast<uint32_t>(ntl * p.block_n * 128));
      for (int j = 0; j < ntl; ++j)
        tma_load_2d(smem + p.off_w + j * kTileB, &p.tm_w, &xf[0], 0, (nt_begin + j) * p.block_n);
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
        const int s = it % S;
        mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
        mbar_expect_tx(&full[s], static_cast<uint32_t>(kTileA));
        tma_load_2d(smem + s * kTileA, &p.tm_x1, &full[s], 0, mt * kBM);
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          AN_GT(0, 0);
          if (has_r) {
            const int rb = tc % p.nbuf;
            const int ncol0 = nt * p.block_n;
            const int ng = min(p.block_n / 64, (p.N - ncol0 + 63) / 64);
            mbar_wait(&rempty[rb], ((tc / p.nbuf) & 1) ^ 1);
            AN_GT(0, 1);
            mbar_expect_tx(&rfull[rb], static_cast<uint32_t>(ng * kGroupBytes));
            for (int gq = 0; gq < ng; ++gq)
              tma_load_2d(smem + p.off_r + rb * p.ytile_bytes + gq * kGroupBytes, &p.tm_r, &rfull[rb],
                          ncol0 + gq * 64, mt * kBM);
          }
          AN_GT(0, 2);
          AN_GT(0, 3);
          AN_GT(0, 4);
        }
      }
    } else if (lane == 0) {
      uint32_t it = 0, tc = 0;
      const uint32_t tx = static_cast<uint32_t>(a_bytes + p.block_n * 128);
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          AN_GT(0, 0);
          if (has_r) {                       // epilogue operand tile (saved raw tensor / residual)
            const int rb = tc % p.nbuf;
            const int ncol0 = nt * p.block_n;
            const int ng = min(p.block_n / 64, (p.N - ncol0 + 63) / 64);
            mbar_wait(&rempty[rb], ((tc / p.nbuf) & 1) ^ 1);
            AN_GT(0, 1);
            mbar_expect_tx(&rfull[rb], static_cast<uint32_t>(ng * kGroupBytes));
            for (int gq = 0; gq < ng; ++gq)
              tma_load_2d(smem + p.off_r + rb * p.ytile_bytes + gq * kGroupBytes, &p.tm_r, &rfull[rb],
                          ncol0 + gq * 64, mt * kBM);
          }
          AN_GT(0, 2);
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&empty[s], ph ^ 1);
            if (kb == 0) AN_GT(0, 3);
            mbar_expect_tx(&full[s], tx);
            uint8_t* st = smem + s * stage_bytes;
            tma_load_2d(st, &p.tm_x1, &full[s], kb * kBK, mt * kBM);
            if (p.a_mode == 2) tma_load_2d(st + kTileA, &p.tm_x2, &full[s], kb * kBK, mt * kBM);
            tma_load_2d(st + a_bytes, &p.tm_w, &full[s], kb * kBK, nt * p.block_n);
          }
          AN_GT(0, 4);
        }
    } else if (lane == 1 && p.staged) {
      // -------------------------------------------------------------- tile store / statistics issuer
      // A thread of its own (never the electing lane of a warp that also executes .sync.aligned
      // tcgen05.ld): waits for a staged tile, issues its TMA stores and the statistics MMAs.
      const bool want_stats = p.stats_part != nullptr;
      uint8_t* ones = smem + p.off_ones;
      uint32_t tc = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          const int yb = tc % p.nbuf;
          const uint32_t yph = (tc / p.nbuf) & 1;
          const int ncol0 = nt * p.block_n;
          const int ncols_valid = min(p.block_n, p.N - ncol0);
          uint8_t* yt = smem + p.off_y + yb * p.ytile_bytes;
          const uint8_t* rt = smem + p.off_r + yb * p.ytile_bytes;
          AN_GT(1, 0);
          mbar_wait(&yfull[yb], yph);
          AN_GT(1, 1);
          const int ng = (ncols_valid + 63) / 64;
          for (int gq = 0; gq < ng; ++gq)
            tma_store_2d(&p.tm_y, yt + gq * kGroupBytes, ncol0 + gq * 64, mt * kBM);
          tma_store_commit();
          AN_GT(1, 2);
          // keep nbuf - 1 stores in flight: the store issued nbuf - 1 tiles ago has read its smem (a 32 KiB tile
          // store occupies the TMA unit for ~2000 clk).  The buffer is handed back BEFORE the statistics MMAs are
          // issued: measured, the epilogue waiting for a free staging tile costs more than its (deferred) read of
          // the statistics waiting for the MMAs (expand forward 1.77 vs 1.96 ms over the sweep).
          switch (p.nbuf) {
            case 4: tma_store_wait_read<3>(); break;
            case 3: tma_store_wait_read<2>(); break;
            case 2: tma_store_wait_read<1>(); break;
            default: tma_store_wait_read<0>(); break;
          }
          if (tc + 1 >= static_cast<uint32_t>(p.nbuf)) mbar_arrive(&yempty[(tc + 1 - p.nbuf) % p.nbuf]);
          if (want_stats) {
            tc_fence_after();
            const uint32_t a_addr = smem_u32(yt);
            const uint32_t b_addr = (p.epi_mode == 1) ? smem_u32(rt) : a_addr;
            const uint32_t o_addr = smem_u32(ones);
            const uint32_t idesc1 = umma_idesc_bf16(128, p.block_n, 1, 1);
            const uint32_t idesc2 = umma_idesc_bf16(128, 16, 1, 1);
#pragma unroll
            for (int k = 0; k < 8; ++k)          // 16 tile rows (the contraction index) per MMA
              umma_bf16(tmem_base + kTmemStat1, umma_desc_sw128(a_addr + k * 2048, kGroupBytes, 1024),
                        umma_desc_sw128(b_addr + k * 2048, kGroupBytes, 1024), idesc1, k > 0 ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem_base + kTmemStat2, umma_desc_sw128(a_addr + k * 2048, kGroupBytes, 1024),
                        umma_desc_sw128(o_addr + k * 2048, kGroupBytes, 1024), idesc2, k > 0 ? 1u : 0u);
            umma_commit(sfull);
            if (has_r) umma_commit(&rempty[yb]);           // operand tile free once the MMAs have read it
          } else if (has_r) {
            mbar_arrive(&rempty[yb]);                       // every epilogue thread is past its reads
          }
          AN_GT(1, 4);
        }
      tma_store_wait<0>();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && p.resident) {
      uint32_t it = 0, tc = 0;
      const uint32_t idesc = umma_idesc_bf16(kBM, p.block_n, 0, 0);
      const int nk = min(4, (p.K + 15) / 16);
      mbar_wait(&xf[0], 0);                      // the group's weight tiles
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x, ++it) {
        const int s = it % S;
        mbar_wait(&full[s], (it / S) & 1);
        const uint32_t a_addr = smem_u32(smem + s * kTileA);
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
          mbar_wait(&tempty[buf], tph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kMaxBN;
          const uint32_t b_addr = smem_u32(smem + p.off_w + (nt - nt_begin) * kTileB);
          for (int k = 0; k < nk; ++k)
            umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, 16, 1024), umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc,
                      k > 0 ? 1u : 0u);
          if (nt == nt_end - 1) umma_commit(&empty[s]);
          umma_commit(&tfull[buf]);
        }
      }
    } else if (lane == 0) {
      uint32_t it = 0, tc = 0;
      const uint32_t idesc = umma_idesc_bf16(kBM, p.block_n, 0, 0);
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
          mbar_wait(&tempty[buf], tph ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kMaxBN;
          for (int kb = 0; kb < kblocks; ++kb, ++it) {
            const int s = it % S;
            const uint32_t ph = (it / S) & 1;
            mbar_wait(&full[s], ph);
            if (p.a_mode != 0) mbar_wait(&xf[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
            const uint32_t b_addr = a_addr + a_bytes;
            const int nk = min(4, (p.K - kb * kBK + 15) / 16);
            for (int k = 0; k < nk; ++k)
              umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32, 16, 1024),
                        umma_desc_sw128(b_addr + k * 32, 16, 1024), idesc,
                        (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty[s]);
            if (kb == kblocks - 1) umma_commit(&tfull[buf]);
          }
        }
    }
  } else if ((warp < 6 || warp >= 10) && p.a_mode != 0) {
    // ------------------------------------------------------------------ A-operand prologue
    // 8 warps (2-5 and 10-13): the K-heavy GEMMs (projection forward, expand dgrad) are bound by
    // this in-place transform, not by their small epilogue
    const int t = (warp < 6 ? warp - 2 : warp - 6) * 32 + lane;      // 0 .. 255
    const int cp = t & 7, r0 = t >> 3;
    const int cl = cp ^ (r0 & 7);            // logical 16B chunk held at physical chunk cp
    uint32_t it = 0;
    for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
      for (int nt = nt_begin; nt < nt_end; ++nt)
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const int s = it % S;
          const uint32_t ph = (it / S) & 1;
          const int kch = kb * kBK + cl * 8;
          float c1[8], c2[8], c3[8];
          load_coef8(p.a1, kch, p.K, c1, 0.f);
          load_coef8(p.a3, kch, p.K, c3, 0.f);
          if (p.a_mode == 2) load_coef8(p.a2, kch, p.K, c2, 0.f);
          mbar_wait(&full[s], ph);
          uint8_t* a = smem + s * stage_bytes + r0 * 128 + cp * 16;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4* ptr = reinterpret_cast<uint4*>(a + i * 32 * 128);
            const uint4 v = *ptr;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (p.a_mode == 2) w = *reinterpret_cast<const uint4*>(a + i * 32 * 128 + kTileA);
            // rows past M were zero-filled by TMA and must stay zero (the affine shift would revive
            // them and they would leak into the epilogue's column statistics)
            const bool live = mt * kBM + r0 + i * 32 < p.M;
            *ptr = live ? transform_chunk(v, w, p.a_mode, p.a_act, c1, c2, c3) : make_uint4(0, 0, 0, 0);
          }
          fence_proxy_async_smem();
          __syncwarp();                          // 256 per-thread arrivals on one mbarrier serialise
          if (lane == 0) mbar_arrive(&xf[s]);
        }
  } else if (warp < 10) {
    // ------------------------------------------------------------------ epilogue
    // Warps 6-9 always; warps 2-5 join when there is no operand prologue (a_mode 0: exactly the
    // output-heavy GEMMs -- expand forward, projection dgrad).  Each TMEM lane quarter is then
    // served by two warps that split the tile's column chunks.
    const int nhalves = (p.a_mode == 0) ? 2 : 1;
    const int half = (warp < 6) ? 1 : 0;
    const int n_epi = 128 * nhalves;
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int et = q * 32 + lane;              // row within the tile
    const int ep_tid = half * 128 + et;
    const bool want_stats = p.stats_part != nullptr;

    if (p.staged) {
      // ---------------- staged epilogue: TMEM -> registers -> swizzled smem tile -> TMA store.
      // BN partial sums come from the tensor core: with the bf16 tile Y (128 rows x block_n) in
      // shared memory as an MN-major operand, D1 = Y^T B (B = Y, or the saved raw tensor R for the
      // backward) has the per-column second moments on its diagonal and D2 = Y^T 1 the column sums.
      uint8_t* ones = smem + p.off_ones;
      if (want_stats) {
        for (int i = ep_tid; i < kGroupBytes / 16; i += n_epi)
          reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
        fence_proxy_async_smem();
        named_bar_sync(1, n_epi);
      }
      const bool swish = p.r_act == ACT_SWISH;
      const float lo = (p.r_act == ACT_RELU || p.r_act == ACT_RELU6) ? 0.f : -INFINITY;
      const float hi = (p.r_act == ACT_RELU6) ? 6.f : INFINITY;
      float4* colpar4 = reinterpret_cast<float4*>(colpar);          // per column PAIR: {s0, s1, h0, h1}
      int prev_gc = 0, prev_valid = 0;
      auto read_stats = [&](uint32_t t, int gc, int nvalid) {
        mbar_wait(sfull, t & 1);
        __syncwarp();
        tc_fence_after();
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        if (nhalves == 1 || half == 1) {                 // column sums
          const float s1 = tmem_ld1(t_lane + kTmemStat2);
          if (et < nvalid) stat_acc[gc] += s1;
        }
        if ((nhalves == 1 || half == 0) && q * 32 < p.block_n) {   // diagonal of D1
          float d[32];
          tmem_ld32(t_lane + kTmemStat1 + q * 32, d);
#pragma unroll
          for (int b = 16; b >= 1; b >>= 1) {
            const bool up = (lane & b) != 0;
#pragma unroll
            for (int i = 0; i < b; ++i) d[i] = up ? d[i + b] : d[i];
          }
          if (et < nvalid) stat_acc[kMaxGroupCols + gc] += d[0];
        }
        tc_fence_before();
      };
      uint32_t tc = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
          const int yb = tc % p.nbuf;
          const uint32_t yph = (tc / p.nbuf) & 1;
          const int ncol0 = nt * p.block_n;
          const int ncols_valid = min(p.block_n, p.N - ncol0);
          const int nchunks = (ncols_valid + 31) / 32;
          const int c_begin = (nhalves == 1 || half == 0) ? 0 : (nchunks + 1) / 2;
          const int c_end = (nhalves == 1) ? nchunks : (half == 0 ? (nchunks + 1) / 2 : nchunks);
          uint8_t* yt = smem + p.off_y + yb * p.ytile_bytes;
          const uint8_t* rt = smem + p.off_r + yb * p.ytile_bytes;
          if (p.epi_mode == 1) {
            for (int c = ep_tid; c < p.block_n / 2; c += n_epi) {
              const int col = ncol0 + 2 * c;
              colpar4[c] = (col < p.N) ? make_float4(__ldg(p.r_scale + col), __ldg(p.r_scale + col + 1),
                                                      __ldg(p.r_shift + col), __ldg(p.r_shift + col + 1))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            named_bar_sync(1, n_epi);
          }
          if (warp == 6 && lane == 0) AN_GT(2, 0);
          if (has_r) mbar_wait(&rfull[yb], yph);
          if (warp == 6 && lane == 0) AN_GT(2, 1);
          mbar_wait(&yempty[yb], yph ^ 1);       // the store that last used this staging buffer is done
          if (warp == 6 && lane == 0) AN_GT(2, 2);
          mbar_wait(&tfull[buf], tph);
          if (warp == 6 && lane == 0) AN_GT(2, 3);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kMaxBN;
          if (c_begin >= c_end) {                // ragged tile: nothing to drain for this half
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[buf]);
          }
          for (int ch = c_begin; ch < c_end; ++ch) {
            float v[32];
            tmem_ld32(t_row + ch * 32, v);
            if (ch == c_end - 1) {               // accumulator is in registers: free it for the next MMA
              tc_fence_before();                 // before the (slower) staging of this chunk
              __syncwarp();
              if (lane == 0) mbar_arrive(&tempty[buf]);
            }
            const int goff = (ch >> 1) * kGroupBytes + et * 128;
            const int cbase = (ch & 1) * 4;                      // first 16-byte chunk of these 32 columns
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const int phys = ((cbase + k4) ^ (et & 7)) * 16;
              const float* vv = v + 8 * k4;
              uint32_t o[4];
              if (p.epi_mode == 0 || !has_r) {
#pragma unroll
                for (int i = 0; i < 4; ++i) o[i] = pack_bf16(vv[2 * i], vv[2 * i + 1]);
              } else {
                const uint4 rr = *reinterpret_cast<const uint4*>(rt + goff + phys);
                const uint32_t ri[4] = {rr.x, rr.y, rr.z, rr.w};
                if (p.epi_mode == 2) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    o[i] = pack_bf16(vv[2 * i] + bf16_lo(ri[i]), vv[2 * i + 1] + bf16_hi(ri[i]));
                } else {
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 cpv = colpar4[ch * 16 + k4 * 4 + i];
                    const float p0 = fmaf(cpv.x, bf16_lo(ri[i]), cpv.z);
                    const float p1 = fmaf(cpv.y, bf16_hi(ri[i]), cpv.w);
                    float g0, g1;
                    if (swish) {
                      g0 = vv[2 * i] * act_bwd(p0, ACT_SWISH);
                      g1 = vv[2 * i + 1] * act_bwd(p1, ACT_SWISH);
                    } else {
                      g0 = (p0 > lo && p0 < hi) ? vv[2 * i] : 0.f;
                      g1 = (p1 > lo && p1 < hi) ? vv[2 * i + 1] : 0.f;
                    }
                    o[i] = pack_bf16(g0, g1);
                  }
                }
              }
              *reinterpret_cast<uint4*>(yt + goff + phys) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          if (warp == 6 && lane == 0) AN_GT(2, 4);
          // Statistics of the PREVIOUS tile: its MMAs were issued when that tile was handed over and have
          // long finished while this tile was drained (waiting for them right after the hand-over cost
          // ~1800 clk per tile: profiles/r02_gemm_tiletrace_*_before.txt).  They are read before this
          // tile's hand-over, so the store thread cannot overwrite the statistics columns earlier.
          if (want_stats && tc > 0) read_stats(tc - 1, prev_gc, prev_valid);
          prev_gc = (nt - nt_begin) * p.block_n + et;
          prev_valid = ncols_valid;
          fence_proxy_async_smem();              // tile (generic-proxy writes) -> TMA store / UMMA reads
          __syncwarp();
          if (lane == 0) mbar_arrive(&yfull[yb]);     // one arrival per warp: per-thread arrivals serialise
          if (warp == 6 && lane == 0) AN_GT(2, 5);
          if (warp == 6 && lane == 0) AN_GT(2, 6);
        }
      if (want_stats && tc > 0) read_stats(tc - 1, prev_gc, prev_valid);
      if (want_stats) {
        named_bar_sync(1, n_epi);
        const int gcols = (nt_end - nt_begin) * p.block_n;
        for (int c = ep_tid; c < gcols; c += n_epi) {
          const int col = nt_begin * p.block_n + c;
          if (col < p.N) {
            p.stats_part[(static_cast<long long>(blockIdx.x) * 2 + 0) * p.N + col] = stat_acc[c];
            p.stats_part[(static_cast<long long>(blockIdx.x) * 2 + 1) * p.N + col] = stat_acc[kMaxGroupCols + c];
          }
        }
      }
    } else {
      // ---------------- direct epilogue (fp32 output / bias: the classifier): registers -> global
      const int nchunks = p.block_n / 16;
      const int c_begin = (nhalves == 1) ? 0 : (half == 0 ? 0 : (nchunks + 1) / 2);
      const int c_end = (nhalves == 1) ? nchunks : (half == 0 ? (nchunks + 1) / 2 : nchunks);
      uint32_t tc = 0;
      for (int mt = blockIdx.x; mt < p.m_tiles; mt += gridDim.x)
        for (int nt = nt_begin; nt < nt_end; ++nt, ++tc) {
          const uint32_t buf = tc & 1, tph = (tc >> 1) & 1;
          const long long row = static_cast<long long>(mt) * kBM + et;
          const bool row_ok = row < p.M;
          const int ncol0 = nt * p.block_n;
          mbar_wait(&tfull[buf], tph);
          tc_fence_after();
          const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * kMaxBN;
          for (int j = c_begin; j < c_end; ++j) {
            float v[16];
            tmem_ld16(t_row + j * 16, v);
            const int col0 = ncol0 + j * 16;
            float s1[16], s2[16];
            if (p.bias != nullptr) {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (col0 + c < p.N) v[c] += __ldg(p.bias + col0 + c);
            }
            if (!p.y_is_f32) {
#pragma unroll
              for (int c = 0; c < 16; ++c) v[c] = round_bf16(v[c]);
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const float x = (row_ok && col0 + c < p.N) ? v[c] : 0.f;
              s1[c] = x;
              s2[c] = x * x;
            }
            if (row_ok) {
              if (p.y_is_f32) {
                float* yp = reinterpret_cast<float*>(p.y) + row * p.ldy + col0;
#pragma unroll
                for (int g = 0; g < 4; ++g)
                  if (col0 + g * 4 < p.N)
                    *reinterpret_cast<float4*>(yp + g * 4) =
                        make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
              } else {
                __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(p.y) + row * p.ldy + col0;
#pragma unroll
                for (int g = 0; g < 2; ++g)
                  if (col0 + g * 8 < p.N)
                    stg_v4(yp + g * 8,
                           make_uint4(pack_bf16(v[8 * g], v[8 * g + 1]), pack_bf16(v[8 * g + 2], v[8 * g + 3]),
                                      pack_bf16(v[8 * g + 4], v[8 * g + 5]), pack_bf16(v[8 * g + 6], v[8 * g + 7])));
              }
            }
            if (want_stats) {
              const float t1 = warp_colsum16(s1, lane);
              const float t2 = warp_colsum16(s2, lane);
              if ((lane & 1) == 0) {
                stat_warp[(q * 2 + 0) * kMaxBN + j * 16 + (lane >> 1)] = t1;
                stat_warp[(q * 2 + 1) * kMaxBN + j * 16 + (lane >> 1)] = t2;
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[buf]);
          if (want_stats) {
            named_bar_sync(1, n_epi);
            if (ep_tid < p.block_n) {
              const int gc = (nt - nt_begin) * p.block_n + ep_tid;
              float a0 = 0.f, a1 = 0.f;
#pragma unroll
              for (int w = 0; w < 4; ++w) {
                a0 += stat_warp[(w * 2 + 0) * kMaxBN + ep_tid];
                a1 += stat_warp[(w * 2 + 1) * kMaxBN + ep_tid];
              }
              stat_acc[gc] += a0;
              stat_acc[kMaxGroupCols + gc] += a1;
            }
            named_bar_sync(1, n_epi);
          }
        }
      if (want_stats) {
        const int gcols = (nt_end - nt_begin) * p.block_n;
        for (int c = ep_tid; c < gcols; c += n_epi) {
          const int col = nt_begin * p.block_n + c;
          if (col < p.N) {
            p.stats_part[(static_cast<long long>(blockIdx.x) * 2 + 0) * p.N + col] = stat_acc[c];
            p.stats_part[(static_cast<long long>(blockIdx.x) * 2 + 1) * p.N + col] =
                stat_acc[kMaxGroupCols + c];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef AN_GEMM_TRACE
  if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned int (*g)[64][8] = g_gemm_trace;
    for (int t = 8; t < 28; ++t)
      printf("tile %2d | producer: period %5u wait-rempty %5u issue-r %5u wait-empty %5u issue-AB %5u | store: period %5u wait-yfull %5u "
             "issue-store %5u stats-mma %5u wait-read %5u | epi: period %5u wait-rfull %5u wait-yempty %5u wait-tfull %5u drain %5u arrive %5u stats %5u\n",
             t, g[0][t][0] - g[0][t - 1][0], g[0][t][1] - g[0][t][0], g[0][t][2] - g[0][t][1], g[0][t][3] - g[0][t][2], g[0][t][4] - g[0][t][3],
             g[1][t][0] - g[1][t - 1][0], g[1][t][1] - g[1][t][0], g[1][t][2] - g[1][t][1], g[1][t][3] - g[1][t][2], g[1][t][4] - g[1][t][3],
             g[2][t][0] - g[2][t - 1][0], g[2][t][1] - g[2][t][0], g[2][t][2] - g[2][t][1], g[2][t][3] - g[2][t][2], g[2][t][4] - g[2][t][3],
             g[2][t][5] - g[2][t][4], g[2][t][6] - g[2][t][5]);
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
constexpr int kWgBox = 64 * 64 * 2;          // one TMA box: 64 rows x 64 channels bf16 = 8 KiB
constexpr int kWgMaxQN = 256;
constexpr int kWgTmemCols = 256;

struct WgradDev {
  CUtensorMap tm_p1, tm_p2, tm_q;
  int Mrows, Cbig, Csmall;
  int qt;                 // small-side tile width (UMMA N), multiple of 16
  int qboxes;             // ceil(qt / 64)
  int stages;
  int p_mode, p_act;
  const float *a1, *a2, *a3;
  float* partials;
  int kblocks_total, kblocks_per_split;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
pw_wgrad_kernel(const __grid_constant__ WgradDev p) {
  extern __shared__ uint8_t smem_raw[];
  // 1 KiB alignment by OFFSET (not by pointer casts): the compiler keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;
  const int p_bytes = (p.p_mode == 2) ? 4 * kWgBox : 2 * kWgBox;
  const int stage_bytes = p_bytes + 4 * kWgBox;
  uint8_t* aux = smem + S * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(aux);
  uint64_t* xf = full + 8;
  uint64_t* empty = xf + 8;
  uint64_t* tfull = empty + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_p1);
    tma_prefetch_desc(&p.tm_q);
    if (p.p_mode == 2) tma_prefetch_desc(&p.tm_p2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&xf[s], 8);                  // one elected arrival per prologue warp
        mbar_init(&empty[s], 1);
      }
      mbar_init(&tfull[0], 1);
      mbar_fence_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, kWgTmemCols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int bb = blockIdx.x;                 // 128-channel block of the big side
  const int split = blockIdx.y;
  const int st = blockIdx.z;                 // tile of the small side
  const int kb_begin = split * p.kblocks_per_split;
  const int kb_end = min(p.kblocks_total, kb_begin + p.kblocks_per_split);
  const int nkb = kb_end - kb_begin;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx = static_cast<uint32_t>(p_bytes + p.qboxes * kWgBox);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        mbar_expect_tx(&full[s], tx);
        uint8_t* sp = smem + s * stage_bytes;
        const int row = (kb_begin + i) * 64;
        tma_load_2d(sp, &p.tm_p1, &full[s], bb * 128, row);
        tma_load_2d(sp + kWgBox, &p.tm_p1, &full[s], bb * 128 + 64, row);
        if (p.p_mode == 2) {
          tma_load_2d(sp + 2 * kWgBox, &p.tm_p2, &full[s], bb * 128, row);
          tma_load_2d(sp + 3 * kWgBox, &p.tm_p2, &full[s], bb * 128 + 64, row);
        }
        for (int b = 0; b < p.qboxes; ++b)
          tma_load_2d(sp + p_bytes + b * kWgBox, &p.tm_q, &full[s], st * p.qt + b * 64, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(128, p.qt, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&full[s], ph);
        if (p.p_mode != 0) mbar_wait(&xf[s], ph);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * stage_bytes);
        const uint32_t b_addr = a_addr + p_bytes;
#pragma unroll
        for (int k = 0; k < 4; ++k)      // 16 pixel rows per MMA: 16 x 128 B = 2048 B
          umma_bf16(tmem_base, umma_desc_sw128(a_addr + k * 2048, kWgBox, 1024),
                    umma_desc_sw128(b_addr + k * 2048, kWgBox, 1024), idesc,
                    (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty[s]);
        if (i == nkb - 1) umma_commit(&tfull[0]);
      }
    }
  } else if (warp < 6 || warp >= 10) {
    if (p.p_mode != 0) {               // 8 operand-prologue warps (2-5, 10-13)
      const int t = (warp < 6 ? warp - 2 : warp - 6) * 32 + lane;
      const int cp = t & 7, r0 = t >> 3;
      const int cl = cp ^ (r0 & 7);
      float c1[2][8], c2[2][8], c3[2][8];
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int ch = bb * 128 + b * 64 + cl * 8;
        load_coef8(p.a1, ch, p.Cbig, c1[b], 0.f);
        load_coef8(p.a3, ch, p.Cbig, c3[b], 0.f);
        load_coef8(p.p_mode == 2 ? p.a2 : nullptr, ch, p.Cbig, c2[b], 0.f);
      }
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (i / S) & 1;
        mbar_wait(&full[s], ph);
        uint8_t* a = smem + s * stage_bytes + r0 * 128 + cp * 16;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint4* ptr = reinterpret_cast<uint4*>(a + b * kWgBox + j * 32 * 128);
            const uint4 v = *ptr;
            uint4 w = make_uint4(0, 0, 0, 0);
            if (p.p_mode == 2)
              w = *reinterpret_cast<const uint4*>(a + (b + 2) * kWgBox + j * 32 * 128);
            *ptr = transform_chunk(v, w, p.p_mode, p.p_act, c1[b], c2[b], c3[b]);
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&xf[s]);
      }
    }
  } else {
    const int q = warp & 3;
    const int cb = bb * 128 + q * 32 + lane;
    mbar_wait(&tfull[0], 0);
    tc_fence_after();
    const int nchunks = p.qt / 16;
    for (int j = 0; j < nchunks; ++j) {
      float v[16];
      tmem_ld16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + j * 16, v);
      const int col0 = st * p.qt + j * 16;
      if (cb < p.Cbig) {
        float* dst = p.partials + (static_cast<long long>(split) * p.Cbig + cb) * p.Csmall + col0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (col0 + g * 4 < p.Csmall)
            *reinterpret_cast<float4*>(dst + g * 4) =
                make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kWgTmemCols);
  }
}

// dst[(cb-row0)*sb + cs*ss] = sum_s partials[s][cb][cs] for cb in [row0, row0+nrows), cs < ncols
__global__ void wgrad_reduce_kernel(const float* __restrict__ partials, int splits, int Cbig,
                                    int Csmall, int row0, int nrows, int ncols,
                                    float* __restrict__ dst, long long sb, long long ss) {
  // block = 32 consecutive outputs x 8 split lanes (fixed summation order -> deterministic)
  __shared__ float red[8][33];
  const long long full = static_cast<long long>(Cbig) * Csmall;
  const long long n = static_cast<long long>(nrows) * ncols;
  const int ex = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (long long base = blockIdx.x * 32LL; base < n; base += gridDim.x * 32LL) {
    const long long i = base + ex;
    float acc = 0.f;
    long long cb = 0, cs = 0;
    if (i < n) {
      cb = i / ncols; cs = i % ncols;
      const long long src = (row0 + cb) * Csmall + cs;
      for (int s = sl; s < splits; s += 8) acc += partials[s * full + src];
    }
    red[sl][ex] = acc;
    __syncthreads();
    if (sl == 0 && i < n) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][ex];
      dst[cb * sb + cs * ss] = a;
    }
    __syncthreads();
  }
}

// The same reduction for up to 4 row ranges of ONE partial buffer in one launch (the branch-fused
// weight gradient is un-fused into the per-branch master tensors): blockIdx.y = range.
struct WgSegs {
  int row0[4], nrows[4];
  float* dst[4];
  long long sb[4], ss[4];
};
__global__ void wgrad_reduce_multi_kernel(const float* __restrict__ partials, int splits, int Cbig, int Csmall,
                                          int ncols, const __grid_constant__ WgSegs sg) {
  __shared__ float red[8][33];
  const int q = blockIdx.y;
  const int row0 = sg.row0[q], nrows = sg.nrows[q];
  float* __restrict__ dst = sg.dst[q];
  const long long sb = sg.sb[q], ss = sg.ss[q];
  const long long full = static_cast<long long>(Cbig) * Csmall;
  const long long n = static_cast<long long>(nrows) * ncols;
  const int ex = threadIdx.x & 31, sl = threadIdx.x >> 5;
  for (long long base = blockIdx.x * 32LL; base < n; base += gridDim.x * 32LL) {
    const long long i = base + ex;
    float acc = 0.f;
    long long cb = 0, cs = 0;
    if (i < n) {
      cb = i / ncols; cs = i % ncols;
      const long long src = (row0 + cb) * Csmall + cs;
      for (int s = sl; s < splits; s += 8) acc += partials[s * full + src];
    }
    red[sl][ex] = acc;
    __syncthreads();
    if (sl == 0 && i < n) {
      float a = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) a += red[k][ex];
      dst[cb * sb + cs * ss] = a;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct GemmPlan {
  int block_n, n_tiles, m_tiles, groups, tiles_per_group, grid_x;
};

static GemmPlan plan_gemm(int M, int N) {
  GemmPlan g;
  const int sms = sm_count();
  // tiles are whole 64-column groups (the unit of the staged epilogue and of its TMA stores)
  g.block_n = (N > 64) ? kMaxBN : 64;
  g.n_tiles = (N + g.block_n - 1) / g.block_n;
  g.m_tiles = (M + kBM - 1) / kBM;
  const int max_tpg = kMaxGroupCols / g.block_n;
  int groups = (g.n_tiles + max_tpg - 1) / max_tpg;
  if (g.m_tiles < sms) {
    int want = sms / g.m_tiles;
    if (want > g.n_tiles) want = g.n_tiles;
    if (want > groups) groups = want;
  }
  g.tiles_per_group = (g.n_tiles + groups - 1) / groups;
  g.groups = (g.n_tiles + g.tiles_per_group - 1) / g.tiles_per_group;
  int gx = sms / g.groups;
  if (gx < 1) gx = 1;
  if (gx > g.m_tiles) gx = g.m_tiles;
  g.grid_x = gx;
  return g;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace an

using namespace an;

extern "C" int an_pw_gemm_stats_rows(int M, int N) { return plan_gemm(M, N).grid_x; }

extern "C" int an_pw_gemm(const an_pw_gemm_args* a, void* stream) {
  AN_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "an_pw_gemm: empty problem %d %d %d", a->M, a->N, a->K);
  AN_REQUIRE(a->K % 8 == 0 && a->N % 8 == 0, "an_pw_gemm: K=%d and N=%d must be multiples of 8", a->K, a->N);
  AN_REQUIRE(a->ldx1 % 8 == 0 && a->ldw % 8 == 0 && a->ldy % 8 == 0, "an_pw_gemm: pitches must be multiples of 8");
  AN_REQUIRE(aligned16(a->x1) && aligned16(a->w) && aligned16(a->y), "an_pw_gemm: pointers must be 16-byte aligned");
  AN_REQUIRE(a->a_mode >= 0 && a->a_mode <= 2 && a->epi_mode >= 0 && a->epi_mode <= 2, "an_pw_gemm: bad mode");
  if (a->a_mode != 0) AN_REQUIRE(a->a1 && a->a3, "an_pw_gemm: a_mode %d needs a1/a3", a->a_mode);
  if (a->a_mode == 2) AN_REQUIRE(a->x2 && a->a2 && a->ldx2 % 8 == 0 && aligned16(a->x2), "an_pw_gemm: a_mode 2 needs x2/a2");
  if (a->epi_mode == 1)
    AN_REQUIRE(a->r && a->r_scale && a->r_shift && a->r_mean && a->r_rstd && a->ldr % 8 == 0 && aligned16(a->r),
               "an_pw_gemm: epi_mode 1 needs r and its BN vectors");
  if (a->epi_mode == 2 && a->r) AN_REQUIRE(a->ldr % 8 == 0 && aligned16(a->r), "an_pw_gemm: bad residual");
  AN_REQUIRE(!(a->y_is_f32 && a->epi_mode != 0), "an_pw_gemm: f32 output only with epi_mode 0");

  const GemmPlan g = plan_gemm(a->M, a->N);
  GemmDev d;
  memset(&d, 0, sizeof(d));
  if (int e = make_tmap_bf16_2d(&d.tm_x1, a->x1, a->K, a->M, a->ldx1, kBK, kBM)) return e;
  if (a->a_mode == 2) {
    if (int e = make_tmap_bf16_2d(&d.tm_x2, a->x2, a->K, a->M, a->ldx2, kBK, kBM)) return e;
  }
  if (int e = make_tmap_bf16_2d(&d.tm_w, a->w, a->K, a->N, a->ldw, kBK, g.block_n)) return e;
  d.M = a->M; d.N = a->N; d.K = a->K;
  d.block_n = g.block_n; d.m_tiles = g.m_tiles; d.n_tiles = g.n_tiles;
  d.tiles_per_group = g.tiles_per_group;
  d.a_mode = a->a_mode; d.a_act = a->a_act; d.a1 = a->a1; d.a2 = a->a2; d.a3 = a->a3;
  d.epi_mode = a->epi_mode; d.y = a->y; d.ldy = a->ldy; d.y_is_f32 = a->y_is_f32;
  d.bias = a->bias; d.stats_part = a->stats_part;
  d.r = reinterpret_cast<const __nv_bfloat16*>(a->r); d.ldr = a->ldr;
  d.r_scale = a->r_scale; d.r_shift = a->r_shift; d.r_mean = a->r_mean; d.r_rstd = a->r_rstd;
  d.r_act = a->r_act;

  // shared memory: [stages][output tiles][operand tiles][ones][barriers + statistics]
  const int a_bytes = (a->a_mode == 2) ? 2 * kTileA : kTileA;
  const int stage_bytes = a_bytes + kTileB;
  d.staged = (!a->y_is_f32 && a->bias == nullptr) ? 1 : 0;
  const bool has_r = d.staged && a->r != nullptr && a->epi_mode != 0;
  d.ytile_bytes = (g.block_n / 64) * kGroupBytes;
  // staged tiles in flight: as many as shared memory allows next to >= min(3, k-blocks + 1) operand stages
  {
    const int kblocks = (a->K + kBK - 1) / kBK;
    const int want_stages = kblocks + 1 < 3 ? kblocks + 1 : 3;
    const int per_buf = d.ytile_bytes * ((d.staged && a->r != nullptr && a->epi_mode != 0) ? 2 : 1);
    const int avail = 227 * 1024 - 1024 - kAuxBytes - kGroupBytes - want_stages * stage_bytes;
    int nb = d.staged ? avail / per_buf : 2;
    if (nb > kMaxStaging) nb = kMaxStaging;
    if (nb < 2 && a->a_mode != 2) {           // two tiles in flight matter more than a third operand stage
      const int avail2 = avail + (want_stages - 2) * stage_bytes;
      if (avail2 / per_buf >= 2) nb = 2;
    }
    if (nb < 1) nb = 1;
    if (getenv("AN_PW_NBUF") != nullptr) { const int e = atoi(getenv("AN_PW_NBUF")); if (e >= 1 && e < nb) nb = e; }
    d.nbuf = nb;
  }
  // resident weights: plain A operand, one k-block, the column group's weight tiles + 2 A stages + two staged
  // tiles fit (AN_PW_RESIDENT=0 switches it off for A/B runs)
  {
    static const bool allow = [] { const char* e = getenv("AN_PW_RESIDENT"); return e == nullptr || e[0] != '0'; }();
    const int per_buf = d.ytile_bytes * (has_r ? 2 : 1);
    const int need = 2 * kTileA + g.tiles_per_group * kTileB + 2 * per_buf + kGroupBytes + kAuxBytes + 1024;
    d.resident = (allow && d.staged && a->a_mode == 0 && a->K <= kBK && need <= 227 * 1024) ? 1 : 0;
    if (d.resident) {
      int nb = (227 * 1024 - (need - 2 * per_buf)) / per_buf;
      if (nb > kMaxStaging) nb = kMaxStaging;
      d.nbuf = nb < 2 ? 2 : nb;                  // what fits next to the resident weights (>= 2 by `need`)
    }
  }
  int fixed = 0;
  if (d.staged) {
    if (int e = make_tmap_bf16_2d(&d.tm_y, a->y, a->N, a->M, a->ldy, 64, kBM)) return e;
    fixed += d.nbuf * d.ytile_bytes;
    if (has_r) {
      if (int e = make_tmap_bf16_2d(&d.tm_r, a->r, a->N, a->M, a->ldr, 64, kBM)) return e;
      fixed += d.nbuf * d.ytile_bytes;
    }
    // statistics MMA: B operand of the column sums = one group of bf16 ones.  (Its A operand always
    // spans two groups; with block_n = 64 the second one is whatever follows the tile -- in bounds,
    // and only ever feeds accumulator rows that are never read.)
    if (a->stats_part != nullptr) fixed += kGroupBytes;
  }
  const int budget = 227 * 1024 - 1024 - kAuxBytes - fixed;
  if (d.resident) {
    d.stages = 2;                                // A tiles only (one per M tile); the weight tiles follow them
    d.off_w = d.stages * kTileA;
    d.off_y = d.off_w + g.tiles_per_group * kTileB;
  } else {
    d.stages = budget / stage_bytes;
    if (d.stages > 4) d.stages = 4;
    AN_REQUIRE(d.stages >= 2, "an_pw_gemm: shared-memory budget too small (%d B for stages)", budget);
    d.off_y = d.stages * stage_bytes;
  }
  d.off_r = d.off_y + (d.staged ? d.nbuf * d.ytile_bytes : 0);
  d.off_ones = d.off_r + (has_r ? d.nbuf * d.ytile_bytes : 0);
  d.off_aux = d.off_ones + ((d.staged && a->stats_part != nullptr) ? kGroupBytes : 0);
  const int smem = d.off_aux + kAuxBytes + 1024;
  static int configured = 0;
  if (configured < smem) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(pw_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = 227 * 1024;
  }
  dim3 grid(g.grid_x, g.groups, 1);
  pw_gemm_kernel<<<grid, kGemmThreads, smem, static_cast<cudaStream_t>(stream)>>>(d);
  AN_CHECK_LAUNCH();
  return 0;
}

namespace an {
struct WgradPlan {
  int big_blocks, qt, qboxes, small_tiles, kblocks_total, kblocks_per_split, splits;
};
static WgradPlan plan_wgrad(int Mrows, int Cbig, int Csmall) {
  WgradPlan w;
  const int sms = sm_count();
  w.big_blocks = (Cbig + 127) / 128;
  const int st0 = (Csmall + kWgMaxQN - 1) / kWgMaxQN;
  w.qt = (((Csmall + st0 - 1) / st0) + 15) / 16 * 16;
  w.small_tiles = (Csmall + w.qt - 1) / w.qt;
  w.qboxes = (w.qt + 63) / 64;
  w.kblocks_total = (Mrows + 63) / 64;
  int want = (2 * sms) / (w.big_blocks * w.small_tiles);
  if (want < 1) want = 1;
  // keep each split at >= 8 k-blocks (512 pixel rows) so the TMEM drain is amortised
  int max_splits = (w.kblocks_total + 7) / 8;
  if (max_splits < 1) max_splits = 1;
  if (want > max_splits) want = max_splits;
  w.kblocks_per_split = (w.kblocks_total + want - 1) / want;
  w.splits = (w.kblocks_total + w.kblocks_per_split - 1) / w.kblocks_per_split;
  return w;
}
}  // namespace an

extern "C" int an_pw_wgrad_splits(int Mrows, int Cbig, int Csmall) {
  return plan_wgrad(Mrows, Cbig, Csmall).splits;
}

extern "C" int an_pw_wgrad(const an_pw_wgrad_args* a, void* stream) {
  AN_REQUIRE(a->Mrows > 0 && a->Cbig > 0 && a->Csmall > 0, "an_pw_wgrad: empty problem");
  AN_REQUIRE(a->Cbig % 8 == 0 && a->Csmall % 8 == 0, "an_pw_wgrad: channel counts must be multiples of 8");
  AN_REQUIRE(a->ldp1 % 8 == 0 && a->ldq % 8 == 0 && aligned16(a->p1) && aligned16(a->q), "an_pw_wgrad: bad operand layout");
  AN_REQUIRE(a->p_mode >= 0 && a->p_mode <= 2, "an_pw_wgrad: bad p_mode");
  if (a->p_mode != 0) AN_REQUIRE(a->a1 && a->a3, "an_pw_wgrad: p_mode needs a1/a3");
  if (a->p_mode == 2) AN_REQUIRE(a->p2 && a->a2 && a->ldp2 % 8 == 0 && aligned16(a->p2), "an_pw_wgrad: p_mode 2 needs p2/a2");
  const WgradPlan w = plan_wgrad(a->Mrows, a->Cbig, a->Csmall);
  AN_REQUIRE(a->splits == w.splits, "an_pw_wgrad: splits %d != planned %d", a->splits, w.splits);
  AN_REQUIRE(a->partials != nullptr && aligned16(a->partials), "an_pw_wgrad: partials buffer missing");

  WgradDev d;
  memset(&d, 0, sizeof(d));
  if (int e = make_tmap_bf16_2d(&d.tm_p1, a->p1, a->Cbig, a->Mrows, a->ldp1, 64, 64)) return e;
  if (a->p_mode == 2) {
    if (int e = make_tmap_bf16_2d(&d.tm_p2, a->p2, a->Cbig, a->Mrows, a->ldp2, 64, 64)) return e;
  }
  if (int e = make_tmap_bf16_2d(&d.tm_q, a->q, a->Csmall, a->Mrows, a->ldq, 64, 64)) return e;
  d.Mrows = a->Mrows; d.Cbig = a->Cbig; d.Csmall = a->Csmall;
  d.qt = w.qt; d.qboxes = w.qboxes;
  d.stages = 3;
  d.p_mode = a->p_mode; d.p_act = a->p_act; d.a1 = a->a1; d.a2 = a->a2; d.a3 = a->a3;
  d.partials = a->partials;
  d.kblocks_total = w.kblocks_total; d.kblocks_per_split = w.kblocks_per_split;
  const int p_bytes = (a->p_mode == 2) ? 4 * kWgBox : 2 * kWgBox;
  const int smem = d.stages * (p_bytes + 4 * kWgBox) + 256 + 1024;
  static bool configured = false;
  if (!configured) {
    AN_CHECK_CUDA(cudaFuncSetAttribute(pw_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  dim3 grid(w.big_blocks, w.splits, w.small_tiles);
  pw_wgrad_kernel<<<grid, kGemmThreads, smem, static_cast<cudaStream_t>(stream)>>>(d);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_wgrad_reduce(const float* partials, int splits, int Cbig, int Csmall, int row0,
                               int nrows, int ncols, float* dst, long long dst_sb, long long dst_ss,
                               void* stream) {
  AN_REQUIRE(partials && dst && splits > 0 && row0 >= 0 && nrows > 0 && row0 + nrows <= Cbig && ncols > 0 &&
                 ncols <= Csmall, "an_wgrad_reduce: bad arguments");
  const long long n = static_cast<long long>(nrows) * ncols;
  int blocks = static_cast<int>((n + 31) / 32);
  if (blocks > 16 * sm_count()) blocks = 16 * sm_count();
  wgrad_reduce_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(partials, splits, Cbig, Csmall, row0, nrows, ncols, dst, dst_sb, dst_ss);
  AN_CHECK_LAUNCH();
  return 0;
}

extern "C" int an_wgrad_reduce_multi(const float* partials, int splits, int Cbig, int Csmall, int ncols,
                                     const an_wg_seg* segs, int nseg, void* stream) {
  AN_REQUIRE(partials && segs && splits > 0 && nseg >= 1 && nseg <= 4 && ncols > 0 && ncols <= Csmall,
             "an_wgrad_reduce_multi: bad arguments");
  WgSegs sg;
  long long nmax = 0;
  for (int i = 0; i < 4; ++i) {
    const an_wg_seg& q = segs[i < nseg ? i : 0];
    AN_REQUIRE(q.dst != nullptr && q.row0 >= 0 && q.nrows > 0 && q.row0 + q.nrows <= Cbig, "an_wgrad_reduce_multi: bad segment %d", i);
    sg.row0[i] = q.row0; sg.nrows[i] = q.nrows; sg.dst[i] = q.dst; sg.sb[i] = q.dst_sb; sg.ss[i] = q.dst_ss;
    if (i < nseg && static_cast<long long>(q.nrows) * ncols > nmax) nmax = static_cast<long long>(q.nrows) * ncols;
  }
  int blocks = static_cast<int>((nmax + 31) / 32);
  if (blocks > 8 * sm_count()) blocks = 8 * sm_count();
  dim3 grid(blocks, nseg);
  wgrad_reduce_multi_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(partials, splits, Cbig, Csmall, ncols, sg);
  AN_CHECK_LAUNCH();
  return 0;
}
