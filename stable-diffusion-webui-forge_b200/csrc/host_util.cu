// b200forge — host-side helpers: error string, device queries, TMA descriptor encoding.
#include "host_util.h"
#include <stdlib.h>

namespace b200 {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int sms = -1;
  if (sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    sms = v;
  }
  return sms;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

PFN_tmapEncodeTiled tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle) {
  PFN_tmapEncodeTiled enc = tmap_encoder();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return B200_ENODEVICE;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base address %p is not 16-byte aligned", base);
    return B200_EINVAL;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes", (unsigned long long)gstr[i - 1], i);
        return B200_EINVAL;
      }
    }
    if (bx[i] == 0 || bx[i] > 256) {
      set_error("TMA box dim %d = %u out of range", i, bx[i]);
      return B200_EINVAL;
    }
  }
  CUresult r = enc(out, dtype == B200_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                   (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu, box %u %u)", (int)r, rank,
              (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0), bx[0], rank > 1 ? bx[1] : 0);
    return B200_ECUDA;
  }
  return B200_OK;
}

}  // namespace b200

extern "C" {

int b200_version(void) { return 100; }

const char* b200_last_error(void) { return b200::g_err; }

int b200_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    b200::set_error("no CUDA device visible");
    return B200_ENODEVICE;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  int major = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    b200::set_error("device compute capability %d.x is not sm_100", major);
    return B200_ENODEVICE;
  }
  return B200_OK;
}

int b200_num_sms(void) { return b200::num_sms(); }

int b200_fill_zero(void* p, size_t bytes, b200_stream_t s) {
  cudaError_t e = cudaMemsetAsync(p, 0, bytes, static_cast<cudaStream_t>(s));
  if (e != cudaSuccess) {
    b200::set_error("cudaMemsetAsync: %s", cudaGetErrorString(e));
    return B200_ECUDA;
  }
  return B200_OK;
}

}  // extern "C"
