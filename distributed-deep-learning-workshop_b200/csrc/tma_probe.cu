// TMA load-throughput probe: how fast can ONE SM's TMA engine fill shared memory for a given box shape?
//
// A persistent kernel (one CTA per SM) whose producer thread issues `loads_per_cta` box loads of a [rows x 64] bf16
// SWIZZLE_128B box into a ring of stages and whose consumer thread does nothing but release each stage as soon as it
// is full: no MMA, no epilogue.  The time per load is therefore the load pipeline itself (issue + L2 + TMA engine).
// Used to decide between many small boxes (one per filter tap) and fewer large halo boxes for the 3x3 convolutions.
#include <stdexcept>
#include <string>

#include "conv_api.h"
#include "ptx.cuh"

namespace b200 {


__global__ void __launch_bounds__(128, 1)
tma_probe_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ TmaProbeParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int stage_bytes = (p.box_bytes + 1023) & ~1023;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 16;
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t s0 = smem_u32(smem), full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const int positions = p.mode == 0 ? p.m_tiles : p.tiles_w * p.tiles_h * p.tiles_n;
      int pos = (static_cast<int>(blockIdx.x) * 977) % positions;  // CTAs start at different places
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < p.loads_per_cta; ++i) {
        mbar_wait_u32(empty0 + stage * 8, phase ^ 1);
        mbar_arrive_expect_tx_u32(full0 + stage * 8, static_cast<uint32_t>(p.box_bytes));
        if (p.mode == 0) {
          tma_load_2d_u32(s0 + stage * stage_bytes, &tm, full0 + stage * 8, 0, pos * 128);
        } else {
          const int tw = pos % p.tiles_w;
          const int rest = pos / p.tiles_w;
          const int th = rest % p.tiles_h;
          const int tn = rest / p.tiles_h;
          tma_load_4d_u32(s0 + stage * stage_bytes, &tm, full0 + stage * 8, 0, tw * p.bw + p.dw, th * p.bh + p.dh,
                          tn * p.bn);
        }
        pos += gridDim.x;
        if (pos >= positions) pos -= positions * (pos / positions);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < p.loads_per_cta; ++i) {
        mbar_wait_u32(full0 + stage * 8, phase);
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty0 + stage * 8) : "memory");
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  }
}

void tma_probe_launch(const CUtensorMap& tm, const TmaProbeParams& p, int grid, cudaStream_t s) {
  const int stage_bytes = (p.box_bytes + 1023) & ~1023;
  const int smem_bytes = p.stages * stage_bytes + 256;
  if (smem_bytes > 232448) throw std::runtime_error("tma_probe: stages * box exceed 227 KB");
  cudaError_t e = cudaFuncSetAttribute(tma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) throw std::runtime_error(std::string("tma_probe: ") + cudaGetErrorString(e));
  tma_probe_kernel<<<grid, 128, smem_bytes, s>>>(tm, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("tma_probe launch: ") + cudaGetErrorString(e));
}

}  // namespace b200
