// Host-side plumbing shared by all entry points: error string, device query, TMA descriptors.
#include "common.cuh"
#include "../../include/atomnas_b200.h"

#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <string.h>

namespace an {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    (void)cudaGetLastError();
    return 148;   // B200; only reached without a device (planning helpers stay callable on CPU)
  }
  cached = n;
  return n;
}

// cuTensorMapEncodeTiled is a driver API; it is looked up through the runtime so that the
// library has no link-time dependency on libcuda (it must load on a GPU-less build box).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t inner, uint64_t outer,
                      uint64_t ld_elems, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  AN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver / no sm_100 GPU)");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AN_REQUIRE(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu ld=%llu box=%ux%u",
             (int)r, base, (unsigned long long)inner, (unsigned long long)outer,
             (unsigned long long)ld_elems, box_inner, box_outer);
  return 0;
}

int make_tmap_bf16_nhwc(CUtensorMap* tm, const void* base, uint64_t C, uint64_t W, uint64_t H,
                        uint64_t B, uint64_t ld_elems, uint32_t box_c, uint32_t box_w, uint32_t box_h,
                        uint32_t box_n, bool oob_nan, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  AN_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver / no sm_100 GPU)");
  cuuint64_t dims[4] = {C, W, H, B};
  cuuint64_t strides[3] = {ld_elems * 2, W * ld_elems * 2, H * W * ld_elems * 2};
  cuuint32_t box[4] = {box_c, box_w, box_h, box_n};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                      : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  oob_nan ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(4d) failed (%d): C=%llu W=%llu H=%llu B=%llu ld=%llu box=%ux%ux%u",
             (int)r, (unsigned long long)C, (unsigned long long)W, (unsigned long long)H, (unsigned long long)B,
             (unsigned long long)ld_elems, box_c, box_w, box_h);
  return 0;
}

}  // namespace an

extern "C" const char* an_last_error(void) { return an::get_last_error(); }
extern "C" int an_version(void) { return 100; }
extern "C" int an_sm_count(void) { return an::sm_count(); }
extern "C" long long an_launch_count(void) { return an::g_launches.load(); }
