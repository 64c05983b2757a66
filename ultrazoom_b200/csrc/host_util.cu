// host_util.cu -- error plumbing, device queries and the tensor-map encoder (driver entry point
// fetched through the runtime so the library has no link-time dependency on libcuda).
#include <stdarg.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace mz {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return MZ_ERR_CUDA;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int encode_tmap(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, void* gaddr, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return MZ_ERR_CUDA;
  }
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], es[5];
  for (uint32_t i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  CUresult r = enc(map, dt, rank, gaddr, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %u, dims %llu/%llu/%llu/%llu, box %u/%u/%u/%u)",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return MZ_ERR_CUDA;
  }
  return MZ_OK;
}

int sm_count(int device) {
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return n;
}

}  // namespace mz

extern "C" {

const char* mz_last_error(void) { return mz::g_err; }

int mz_abi_version(void) { return MZ_ABI_VERSION; }

int mz_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
  }
  return ok;
}

// Channel counts are padded to a multiple of 16 (one UMMA k-step); beyond 64 to a multiple of 32 so that a K chunk
// is at least a 64-byte swizzled row (108 -> 128 rather than 112 = 7 x 16).
int mz_padded_channels(int32_t c) {
  if (c <= 0) return 0;
  const int p16 = ((c + 15) / 16) * 16;
  return p16 <= 64 ? p16 : ((c + 31) / 32) * 32;
}

int mz_zb_pitch(int32_t cp) {
  // MZ_ZB_PITCH64=1 restores the 128-byte rows of an earlier layout for the 48-channel model (diagnostic)
  static const bool wide = getenv("MZ_ZB_PITCH64") != nullptr;
  return (cp == 48 && wide) ? 64 : cp;
}

}  // extern "C"
