// Probe: D = Y^T B on tcgen05 with both operands MN-major from one swizzled smem tile.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../atomnas_b200/csrc/common.cuh"
using namespace an;
__global__ void __launch_bounds__(128, 1) probe(float* out, int N, int same) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* yt = smem;                 // 2 groups x 16 KiB
  uint8_t* bt = smem + 32768;         // second tile (used when same == 0)
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (warp == 0) { if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); } __syncwarp(); tmem_alloc(slot, 512); }
  // Y[r][c] = ((r*7 + c*3) % 5) - 2 ; B[r][c] = ((r*3 + c*5) % 7) - 3
  for (int c8 = 0; c8 < 16; ++c8) {     // 16 chunks of 8 columns = 128 columns
    uint32_t o[4], ob[4];
    for (int i = 0; i < 4; ++i) {
      const int c = c8 * 8 + 2 * i;
      o[i] = pack_bf16(float(((t * 7 + c * 3) % 5) - 2), float(((t * 7 + (c + 1) * 3) % 5) - 2));
      ob[i] = pack_bf16(float(((t * 3 + c * 5) % 7) - 3), float(((t * 3 + (c + 1) * 5) % 7) - 3));
    }
    const int grp = c8 >> 3, j = c8 & 7;
    const int off = grp * 16384 + t * 128 + ((j ^ (t & 7)) * 16);
    *reinterpret_cast<uint4*>(yt + off) = make_uint4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<uint4*>(bt + off) = make_uint4(ob[0], ob[1], ob[2], ob[3]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *slot;
  if (t == 0) {
    const uint32_t a = smem_u32(yt), b = same ? a : smem_u32(bt);
    const uint32_t idesc = umma_idesc_bf16(128, N, 1, 1);
    for (int k = 0; k < 8; ++k)
      umma_bf16(tb + 256, umma_desc_sw128(a + k * 2048, 16384, 1024), umma_desc_sw128(b + k * 2048, 16384, 1024), idesc, k > 0 ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int ch = 0; ch < N / 32; ++ch) {
    float v[32];
    tmem_ld32(tb + (static_cast<uint32_t>(warp * 32) << 16) + 256 + ch * 32, v);
    for (int i = 0; i < 32; ++i) out[(size_t)t * N + ch * 32 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tb, 512); }
}
int main() {
  for (int same = 0; same < 2; ++same)
    for (int N : {64, 128}) {
      float* d; cudaMalloc(&d, 128 * N * 4);
      cudaMemset(d, 0xff, 128 * N * 4);
      cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      probe<<<1, 128, 80 * 1024>>>(d, N, same);
      cudaError_t e = cudaDeviceSynchronize();
      float* h = (float*)malloc(128 * N * 4);
      cudaMemcpy(h, d, 128 * N * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0; int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int r = 0; r < 128; ++r) {
          double y = ((r * 7 + m * 3) % 5) - 2;
          double b = same ? (((r * 7 + n * 3) % 5) - 2) : (((r * 3 + n * 5) % 7) - 3);
          ref += y * b;
        }
        double err = fabs(ref - h[m * N + n]);
        if (err > 1e-3) { if (bad < 5) printf("  mismatch m=%d n=%d got %f ref %f\n", m, n, h[m * N + n], ref); ++bad; }
        if (err > maxerr) maxerr = err;
      }
      printf("same=%d N=%d status=%s bad=%d maxerr=%g\n", same, N, cudaGetErrorString(e), bad, maxerr);
      cudaFree(d); free(h);
    }
  return 0;
}
