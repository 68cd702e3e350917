// cuTensorMapEncodeTiled wrapper shared by the extensions that build TMA descriptors (conv kernels, hardware probes).
#include <cuda.h>
#include <cuda_runtime.h>

#include <mutex>
#include <stdexcept>
#include <string>

#include "conv_api.h"

namespace b200 {

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || p == nullptr)
      throw std::runtime_error("b200ddl: cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    fn = reinterpret_cast<EncodeFn>(p);
  });
  return fn;
}

// dims/strides innermost first; strides in BYTES for dims 1..rank-1.
CUtensorMap encode_bf16(void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                               const uint32_t* box) {
  CUtensorMap m;
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, ptr, gd, gs, bx, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string msg = "b200ddl: cuTensorMapEncodeTiled failed, code " + std::to_string(int(r)) + " rank " +
                      std::to_string(rank) + " dims";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(dims[i]);
    msg += " box";
    for (int i = 0; i < rank; ++i) msg += " " + std::to_string(box[i]);
    throw std::runtime_error(msg);
  }
  return m;
}

}  // namespace b200
