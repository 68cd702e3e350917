// Weight-gradient convolution kernel for sm_100a (tcgen05 + TMEM + TMA), split-K over pixels.
//
//   dW[tap][co][ci] += sum_px dY[px, co] * X_tap[px, ci]
//
// Both operands are "MN-major" for the tensor core (the reduction index, the pixel, is the slow smem dimension):
// the very same TMA boxes the forward kernel loads ([P pixels] x [64 channels], SWIZZLE_128B) are consumed
// directly with a_major = b_major = MN in the instruction descriptor, so there is no transpose pass.
// A unit of work = (pixel chunk, 128-wide co block, group of <= 6 (tap, ci-block) chunks of one filter row);
// the accumulators (up to 384 fp32 columns) live in TMEM for the whole pixel chunk and are flushed once with
// 16-byte vector atomics (red.global.add.v4.f32) into the fp32 gradient buffer that the all-reduce consumes.
#include <stdexcept>
#include <string>

#include "conv_api.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kWgMaxStages = 8;

__global__ void __launch_bounds__(256, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX0,
                  const __grid_constant__ CUtensorMap tmX1, const __grid_constant__ CUtensorMap tmX2,
                  const __grid_constant__ CUtensorMap tmX3, const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  const int slot_bytes = p.P * 128;
  const int stage_bytes = (p.a_chunks + p.G) * slot_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;                   // [stages]
  uint64_t* empty_bar = bars + kWgMaxStages;   // [stages]
  uint64_t* tfull_bar = bars + 2 * kWgMaxStages;
  uint64_t* tempty_bar = bars + 2 * kWgMaxStages + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kWgMaxStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    tma_prefetch_desc(&tmX0);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 4);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int combos = p.co_blocks * p.groups;
  const int n_acc = p.G / p.acc_chunks;
  const int acc_cols = 64 * p.acc_chunks;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx = static_cast<uint32_t>(stage_bytes);
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int pxc = unit / combos;
      const int combo = unit - pxc * combos;
      const int cob = combo / p.groups;
      const int g = combo - cob * p.groups;
      const int r = g / p.cgroups;
      const int cbg = g - r * p.cgroups;
      const int it0 = pxc * p.iters_per_chunk;
      const int it1 = min(it0 + p.iters_per_chunk, p.iters_total);
      for (int it = it0; it < it1; ++it) {
        int w0 = 0, h0 = 0, n0 = 0;
        if (p.mode == 1) {
          const int tw = it % p.tiles_w;
          const int rest = it / p.tiles_w;
          const int th = rest % p.tiles_h;
          w0 = tw * p.bw;
          h0 = th * p.bh;
          n0 = (rest / p.tiles_h) * p.bn;
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], tx);
        uint8_t* base = smem + stage * stage_bytes;
        for (int i = 0; i < p.a_chunks; ++i) {
          const int c = (cob * 2 + i) * 64;
          if (p.mode == 0)
            tma_load_2d(base + i * slot_bytes, &tmDY, &full_bar[stage], c, it * p.P);
          else
            tma_load_4d(base + i * slot_bytes, &tmDY, &full_bar[stage], c, w0, h0, n0);
        }
        uint8_t* bbase = base + p.a_chunks * slot_bytes;
        for (int j = 0; j < p.G; ++j) {
          const int cbi = j / p.S;
          const int s = j - cbi * p.S;
          const int tap = r * p.S + s;
          const int c = (cbg * p.cb_per_group + cbi) * 64;
          const CUtensorMap* xm = &tmX0;
          if (p.tap_map[tap] == 1) xm = &tmX1;
          if (p.tap_map[tap] == 2) xm = &tmX2;
          if (p.tap_map[tap] == 3) xm = &tmX3;
          if (p.mode == 0)
            tma_load_2d(bbase + j * slot_bytes, xm, &full_bar[stage], c, it * p.P);
          else
            tma_load_4d(bbase + j * slot_bytes, xm, &full_bar[stage], c, w0 + p.tap_dw[tap], h0 + p.tap_dh[tap], n0);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------------ MMA issuer
    const uint32_t idesc = umma_idesc_bf16(128, acc_cols, 1, 1);  // both operands MN-major
    const uint32_t a_lbo = (p.a_chunks == 2) ? slot_bytes : 0;    // Cout == 64: rows 64..127 mirror rows 0..63
    int stage = 0;
    uint32_t phase = 0;
    uint32_t uphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int pxc = unit / combos;
      const int it0 = pxc * p.iters_per_chunk;
      const int it1 = min(it0 + p.iters_per_chunk, p.iters_total);
      mbar_wait(tempty_bar, uphase ^ 1);
      tc_fence_after();
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t abase = smem_u32(smem + stage * stage_bytes);
        const uint32_t bbase = abase + p.a_chunks * slot_bytes;
        const int ksteps = p.P / 16;
        for (int ks = 0; ks < ksteps; ++ks) {
          // 16 pixels = two 8-row swizzle atoms = 2048 B along the (slow) K dimension
          const uint64_t da = umma_desc_sw128(abase + ks * 2048, a_lbo, 1024);
          for (int a = 0; a < n_acc; ++a) {
            const uint64_t db = umma_desc_sw128(bbase + a * p.acc_chunks * slot_bytes + ks * 2048, slot_bytes, 1024);
            umma_bf16(tmem_base + a * acc_cols, da, db, idesc, (it != it0 || ks != 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(tfull_bar);
      uphase ^= 1;
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: TMEM -> vector atomics
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    uint32_t uphase = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      const int pxc = unit / combos;
      const int combo = unit - pxc * combos;
      const int cob = combo / p.groups;
      const int g = combo - cob * p.groups;
      const int r = g / p.cgroups;
      const int cbg = g - r * p.cgroups;
      const int co = cob * 128 + row;
      mbar_wait(tfull_bar, uphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      for (int j = 0; j < p.G; ++j) {
        const int cbi = j / p.S;
        const int s = j - cbi * p.S;
        const int tap = r * p.S + s;
        const int ci0 = (cbg * p.cb_per_group + cbi) * 64;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(taddr + j * 64 + h * 32, v);
          tmem_ld_wait();
          if (co < p.cout) {
            float4* dst = reinterpret_cast<float4*>(p.dw + (static_cast<size_t>(tap) * p.cout + co) * p.cin + ci0 + h * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 f = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                     __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
              atomicAdd(dst + q, f);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar);
      uphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

void wgrad_plan_launch(const WgradPlanRaw& pl, cudaStream_t s) {
  static bool attr_set = false;
  const int smem_bytes = pl.p.stages * (pl.p.a_chunks + pl.p.G) * pl.p.P * 128 + 256;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess)
      throw std::runtime_error(std::string("conv_wgrad: cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    attr_set = true;
  }
  conv_wgrad_kernel<<<pl.grid, 256, smem_bytes, s>>>(pl.tmDY, pl.tmX[0], pl.tmX[1], pl.tmX[2], pl.tmX[3], pl.p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string("conv_wgrad launch: ") + cudaGetErrorString(e));
}

}  // namespace b200
