// Host helpers: error plumbing and TMA tensor-map construction.
#include <cstdarg>
#include <cstdio>
#include <map>
#include <mutex>
#include <utility>
#include <cudaTypedefs.h>

#include "kernels.h"

namespace asr {

static thread_local char g_err[512] = "";
unsigned long long g_kernel_launches = 0;

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

int ensure_dyn_smem(const void* fn, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  int dev = 0;
  ASR_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = done[{dev, fn}];
  if (have == 0) have = 48 * 1024;      // the default limit needs no opt-in
  if (bytes > have) {
    ASR_CUDA_OK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return 0;
}

int device_props(int* n_sm, int* max_smem_optin) {
  static std::mutex mu;
  static std::map<int, std::pair<int, int>> cache;
  int dev = 0;
  ASR_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(dev);
  if (it == cache.end()) {
    int sm = 0, smem = 0;
    ASR_CUDA_OK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
    ASR_CUDA_OK(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    it = cache.emplace(dev, std::make_pair(sm, smem)).first;
  }
  if (n_sm) *n_sm = it->second.first;
  if (max_smem_optin) *max_smem_optin = it->second.second;
  return 0;
}

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes);
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, base, rank, dims, strides_bytes, box, elem_strides, swizzle_bytes);
}
int make_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  return make_tmap_any(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, elem_strides, swizzle_bytes);
}
static int make_tmap_any(CUtensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                         const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides, int swizzle_bytes) {
  auto fn = encode_fn();
  if (!fn) return set_error(-101, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(-102, "TMA base pointer not 16-byte aligned");
  for (int i = 1; i < rank; ++i)
    if (strides_bytes[i] % 16 != 0) return set_error(-102, "TMA stride %d (%llu B) not a multiple of 16", i,
                                                     (unsigned long long)strides_bytes[i]);
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(-103, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace asr
