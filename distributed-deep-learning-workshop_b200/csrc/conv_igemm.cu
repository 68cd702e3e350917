// Host side of the tcgen05 implicit-GEMM convolution: tensor-map encoding, plan object, launch.
#include <mutex>
#include <stdexcept>
#include <vector>

#include "conv_igemm.cuh"
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


template <int BN, int ST, bool RB = false, bool HL = false>
static void launch_t(const ConvPlanRaw& pl, cudaStream_t s) {
  auto kern = conv_igemm_kernel<BN, ST, RB, HL>;
  using L = ConvSmem<BN, ST, RB, HL>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) throw std::runtime_error(std::string("conv_igemm: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  TmapArray4 a;
  for (int i = 0; i < 4; ++i) a.m[i] = pl.tmA[i];
  kern<<<pl.grid, ST ? 384 : 256, L::kTotal, s>>>(a, pl.tmB, pl.tmD, pl.tmY, pl.p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("conv_igemm launch: ") + cudaGetErrorString(e));
}

template <int ST>
static void launch_n(const ConvPlanRaw& pl, cudaStream_t s) {
  if (pl.block_n == 256) launch_t<256, ST>(pl, s);
  else if (pl.block_n == 128) launch_t<128, ST>(pl, s);
  else if (pl.res_b && pl.halo) launch_t<64, ST, true, true>(pl, s);
  else if (pl.res_b) launch_t<64, ST, true>(pl, s);
  else launch_t<64, ST>(pl, s);
}

void conv_plan_launch(const ConvPlanRaw& pl, cudaStream_t s) {
  if (pl.stats == 1) launch_n<1>(pl, s);
  else if (pl.stats == 2) launch_n<2>(pl, s);
  else if (pl.stats == 3) {
    // block-gradient merge epilogue: flat 1x1 plans only, no resident-filter / halo variants
    if (pl.block_n == 256) launch_t<256, 3>(pl, s);
    else if (pl.block_n == 128) launch_t<128, 3>(pl, s);
    else launch_t<64, 3>(pl, s);
  } else launch_n<0>(pl, s);
}

}  // namespace b200
