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

  // provably warp-uniform (see conv_igemm.cuh): keeps the single-thread MMA loop free of per-instruction ELECT loops
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
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

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one warp; lane j issues chunk j)
    // All per-chunk parameters (tensor map, channel coordinate, tap offset) are hoisted out of the k loop and the
    // <= 8 loads of a stage are issued by 8 lanes in parallel: a single thread doing the index arithmetic for every
    // load was the bottleneck of this kernel (ncu: tensor pipe 26 % active, everything else idle).
    const int n_loads = p.a_chunks + p.G;
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
      // this lane's load
      const CUtensorMap* my_map = &tmDY;
      int my_c = 0, my_dw = 0, my_dh = 0;
      if (lane < p.a_chunks) {
        my_c = (cob * 2 + lane) * 64;
      } else if (lane < n_loads) {
        const int j = lane - p.a_chunks;
        const int cbi = j / p.S;
        const int sx = j - cbi * p.S;
        const int tap = r * p.S + sx;
        my_c = (cbg * p.cb_per_group + cbi) * 64;
        my_dw = p.tap_dw[tap];
        my_dh = p.tap_dh[tap];
        my_map = &tmX0;
        if (p.tap_map[tap] == 1) my_map = &tmX1;
        if (p.tap_map[tap] == 2) my_map = &tmX2;
        if (p.tap_map[tap] == 3) my_map = &tmX3;
      }
      // pixel-box coordinates advance incrementally (no div/mod in the loop)
      int tw = 0, th = 0, tn = 0;
      if (p.mode == 1) {
        tw = it0 % p.tiles_w;
        const int rest = it0 / p.tiles_w;
        th = rest % p.tiles_h;
        tn = rest / p.tiles_h;
      }
      for (int it = it0; it < it1; ++it) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], tx);
        __syncwarp();
        if (lane < n_loads) {
          uint8_t* dst = smem + stage * stage_bytes + lane * slot_bytes;
          if (p.mode == 0)
            tma_load_2d(dst, my_map, &full_bar[stage], my_c, it * p.P);
          else
            tma_load_4d(dst, my_map, &full_bar[stage], my_c, tw * p.bw + my_dw, th * p.bh + my_dh, tn * p.bn);
        }
        if (++tw == p.tiles_w) {
          tw = 0;
          if (++th == p.tiles_h) {
            th = 0;
            ++tn;
          }
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------------------------------------------------------- MMA issuer (one elected thread)
      // elect.sync + integer barrier addresses: with `lane == 0` every tcgen05.mma / commit below was wrapped in an
      // ELECT / PLOP3 / BRA.U.ANY serialisation loop (~10 extra instructions per MMA, 8-24 MMAs per stage).
      const uint32_t idesc = umma_idesc_bf16(128, acc_cols, 1, 1);  // both operands MN-major
      const uint32_t a_lbo = (p.a_chunks == 2) ? slot_bytes : 0;    // Cout == 64: rows 64..127 mirror rows 0..63
      // descriptors differ only in the 14-bit start-address field: build one per stage up front, then add offsets
      const uint64_t da0 = umma_desc_sw128(smem_u32(smem), a_lbo, 1024);
      const uint64_t db0 = umma_desc_sw128(smem_u32(smem) + p.a_chunks * slot_bytes, slot_bytes, 1024);
      const uint32_t stage_step = static_cast<uint32_t>(stage_bytes) >> 4;
      const uint32_t acc_step = static_cast<uint32_t>(p.acc_chunks * slot_bytes) >> 4;
      const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
      const uint32_t tfull_a = smem_u32(tfull_bar), tempty_a = smem_u32(tempty_bar);
      const int ksteps = p.P / 16;
      const int num_units = p.num_units, iters_per_chunk = p.iters_per_chunk, iters_total = p.iters_total, stages = p.stages;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t uphase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const int pxc = unit / combos;
        const int it0 = pxc * iters_per_chunk;
        const int it1 = min(it0 + iters_per_chunk, iters_total);
        mbar_wait_u32(tempty_a, uphase ^ 1);
        tc_fence_after();
        for (int it = it0; it < it1; ++it) {
          mbar_wait_u32(full0 + stage * 8, phase);
          tc_fence_after();
          const uint64_t da_s = da0 + stage * stage_step;
          const uint64_t db_s = db0 + stage * stage_step;
          for (int ks = 0; ks < ksteps; ++ks) {
            // 16 pixels = two 8-row swizzle atoms = 2048 B along the (slow) K dimension -> +128 in the address field
            const uint32_t acc_flag = (it != it0 || ks != 0) ? 1u : 0u;
            for (int a = 0; a < n_acc; ++a)
              umma_bf16(tmem_base + a * acc_cols, da_s + ks * 128, db_s + a * acc_step + ks * 128, idesc, acc_flag);
          }
          umma_commit_u32(empty0 + stage * 8);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_u32(tfull_a);
        uphase ^= 1;
      }
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
