// Warp-level tensor-core (mma.sync m16n8k16 bf16 -> f32, SASS HMMA) issue-rate probe for sm_100a (B200):
// the ceiling of a depthwise convolution written as banded (Toeplitz) MMAs, one channel per MMA.
// Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_probe hmma_probe.cu
// Prints MMAs per clock per SM (clock64) and dense TMAC/s (events) for
//   hmma     : back-to-back MMAs, NA independent accumulators, operands fixed in registers
//   hmma_mix : 14 MMAs : 4 LDS.32 : 4 PRMT : 12 ALU (the A-fragment build of the banded depthwise loop)
// at 4 / 8 / 16 / 32 resident warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int ITERS = 4096;

template <int MODE, int THREADS, int NA>
__global__ void __launch_bounds__(THREADS) probe(float* out, long long* cycles, const uint32_t* in) {
  __shared__ uint32_t sm[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0x3c003c00u + i;
  __syncthreads();
  const uint32_t s = in[threadIdx.x & 31];
  float acc[NA][4];
#pragma unroll
  for (int i = 0; i < NA; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  uint32_t a[2][4], b[7][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) a[i][j] = 0x3c003c00u + s + i * 4 + j;
#pragma unroll
  for (int i = 0; i < 7; ++i) { b[i][0] = 0x3c003c00u + s + i; b[i][1] = 0x3c003c00u + s + 2 * i; }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 1) {
      // rebuild the two A fragments from shared memory: 4 LDS.32 -> scale/shift/relu on the pairs -> 4 PRMT
      uint32_t l[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) l[j] = sm[(threadIdx.x * 4 + j + it) & 4095];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float lo = __uint_as_float(l[j] << 16), hi = __uint_as_float(l[j] & 0xffff0000u);
        lo = fmaxf(fmaf(lo, 1.0001f, 1e-6f), 0.f);
        hi = fmaxf(fmaf(hi, 0.9999f, 1e-6f), 0.f);
        l[j] = __byte_perm(__float_as_uint(lo), __float_as_uint(hi), 0x7632);
      }
      a[0][0] = __byte_perm(l[0], l[1], 0x5410); a[1][0] = __byte_perm(l[0], l[1], 0x7632);
      a[0][1] = __byte_perm(l[2], l[3], 0x5410); a[1][1] = __byte_perm(l[2], l[3], 0x7632);
    }
    // 14 MMAs: 2 channels x 7 kernel rows; accumulator i of NA (ring of output rows)
#pragma unroll
    for (int r = 0; r < 14; ++r) mma16816(acc[r % NA], a[r & 1], b[r % 7]);
  }
  const long long t1 = clock64();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < NA; ++i) r += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int THREADS, int NA>
int run(const char* name, int sms, float* out, long long* cyc, const uint32_t* in) {
  cudaEvent_t ea, eb;
  CK(cudaEventCreate(&ea)); CK(cudaEventCreate(&eb));
  probe<MODE, THREADS, NA><<<sms, THREADS>>>(out, cyc, in);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(ea));
  probe<MODE, THREADS, NA><<<sms, THREADS>>>(out, cyc, in);
  CK(cudaEventRecord(eb));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, ea, eb));
  long long h[1024];
  CK(cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
  const double mmas_per_sm = (double)ITERS * 14 * (THREADS / 32);
  printf("%-9s warps/SM=%2d acc=%2d  MMA/clk/SM=%6.3f  clk/MMA/SMSP=%6.2f  dense MAC/clk/SM=%7.1f  TMAC/s=%7.2f  (%.3f ms)\n", name,
         THREADS / 32, NA, mmas_per_sm / mx, mx / (mmas_per_sm / 4.0), mmas_per_sm * 2048 / mx,
         mmas_per_sm * 2048 * sms / (ms * 1e-3) / 1e12, ms);
  return 0;
}

template <int THREADS>
int run_all(int sms, float* out, long long* cyc, const uint32_t* in) {
  if (run<0, THREADS, 14>("hmma", sms, out, cyc, in)) return 1;
  if (run<0, THREADS, 7>("hmma", sms, out, cyc, in)) return 1;
  if (run<0, THREADS, 2>("hmma", sms, out, cyc, in)) return 1;
  if (run<1, THREADS, 14>("hmma_mix", sms, out, cyc, in)) return 1;
  return 0;
}

int main() {
  int sms = 0;
  CK(cudaSetDevice(0));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  float* out; uint32_t* in; long long* cyc;
  CK(cudaMalloc(&out, sizeof(float) * sms * 1024));
  CK(cudaMalloc(&in, sizeof(uint32_t) * 32));
  CK(cudaMemset(in, 0, sizeof(uint32_t) * 32));
  CK(cudaMalloc(&cyc, sizeof(long long) * 1024));
  printf("SMs=%d  mma.sync.m16n8k16 bf16: 2048 MAC per MMA\n", sms);
  if (run_all<128>(sms, out, cyc, in)) return 1;
  if (run_all<256>(sms, out, cyc, in)) return 1;
  if (run_all<512>(sms, out, cyc, in)) return 1;
  if (run_all<1024>(sms, out, cyc, in)) return 1;
  return 0;
}
