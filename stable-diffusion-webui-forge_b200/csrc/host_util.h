// b200forge — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200forge.h"

namespace b200 {

void set_error(const char* fmt, ...);
int num_sms();

// cuTensorMapEncodeTiled fetched through the runtime so that the library has no link-time
// dependency on libcuda.so (it must dlopen on a GPU-less build box).
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
PFN_tmapEncodeTiled tmap_encoder();

// rank-N tiled tensor map over 16-bit elements, 128-byte swizzle, zero fill out of bounds.
// dims/box are innermost-first; strides_bytes has rank-1 entries (dims 1..rank-1).
int make_tmap(CUtensorMap* out, int dtype, const void* base, int rank, const uint64_t* dims,
              const uint64_t* strides_bytes, const uint32_t* box,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B);

// Programmatic dependent launch (B200_PDL=0 turns it off): a kernel launched through launch_pdl() may start — CTA scheduling,
// barrier init, TMEM allocation, tensor-map prefetch — while its predecessor on the stream is still draining, and blocks in
// `griddepcontrol.wait` until the predecessor grid has completed and flushed.  Every kernel launched this way MUST execute
// pdl_wait() (common.cuh) on every CTA before it touches global memory: the transitive ordering A -> B -> C relies on B not
// completing before A.  Inside a stream capture the edge becomes a programmatic dependency of the CUDA graph.
bool pdl_enabled();

template <typename... KA, typename... A>
static inline cudaError_t launch_pdl(void (*kern)(KA...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                     A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KA>(args)...);
}

#define B200_CHECK_ARG(cond, ...)  \
  do {                             \
    if (!(cond)) {                 \
      b200::set_error(__VA_ARGS__); \
      return B200_EINVAL;          \
    }                              \
  } while (0)

#define B200_CHECK_LAUNCH(name)                                                  \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      b200::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));   \
      return B200_ECUDA;                                                         \
    }                                                                            \
  } while (0)

}  // namespace b200
