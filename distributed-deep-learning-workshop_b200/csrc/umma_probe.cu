// Hardware probe: does a SWIZZLE_128B K-major UMMA descriptor whose start address is shifted by whole 128-byte
// rows (not 1024-byte aligned) read the rows TMA wrote there?  And does the descriptor's base_offset field matter?
// D[m, n] = sum_k T[m + shift, k] * B[n, k]   for a 160-row tile T loaded by TMA, M = 128, N = K = 64.
// The answer decides whether a 3x3 convolution can reuse ONE halo tile for its horizontal taps (row-shifted
// descriptors) instead of re-loading the tile per tap.
// Second question (sbo_bytes != 1024): may the 8-row groups of the M = 128 operand sit at a stride that is not 1024
// bytes?  Output row 8g + r then reads tile row shift + g * (sbo_bytes / 128) + r.  A 2-D halo tile (8 output pixels per
// image row inside a 10-pixel-wide halo row) needs a stride of 1280 bytes.
#include <stdexcept>
#include <string>

#include "conv_api.h"
#include "ptx.cuh"

namespace b200 {

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmT, const __grid_constant__ CUtensorMap tmB, float* out,
                  int shift, int use_base_offset, int t_rows, int sbo_bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int t_bytes = (t_rows * 128 + 1023) & ~1023;
  uint8_t* sT = smem;                 // t_rows x 128 B (1024-aligned)
  uint8_t* sB = smem + t_bytes;       // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + t_bytes + 8 * 1024);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, t_rows * 128 + 64 * 128);
    tma_load_2d(sT, &tmT, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sT) + shift * 128;
    uint64_t da = umma_desc_sw128(a_addr, 16, static_cast<uint32_t>(sbo_bytes));
    if (use_base_offset) da |= static_cast<uint64_t>((a_addr >> 7) & 7u) << 49;
    const uint64_t db = umma_desc_sw128(smem_u32(sB), 16, 1024);
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0 ? 1u : 0u);
    umma_commit(done);
  }
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t r[32];
  const int row = warp * 32 + lane;
  for (int h = 0; h < 2; ++h) {
    tmem_ld_32x32b_x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + h * 32, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) out[row * 64 + h * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

void umma_probe_launch(const CUtensorMap& tmT, const CUtensorMap& tmB, float* out, int shift, int use_base_offset,
                       int t_rows, int sbo_bytes, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  const int t_bytes = (t_rows * 128 + 1023) & ~1023;
  umma_probe_kernel<<<1, 128, t_bytes + 9 * 1024 + 64, s>>>(tmT, tmB, out, shift, use_base_offset, t_rows, sbo_bytes);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("umma_probe: ") + cudaGetErrorString(e));
}

}  // namespace b200
