// ResNet stem (7x7 stride-2 pad-3 conv on a 3-channel uint8 image) on the tensor cores, forward and weight gradient.
//
// A 3-channel pixel is 3 bytes - below TMA's 16-byte granule - so the im2col operand is built by threads:
// 256 "builder" threads gather each output pixel's 7x7x3 = 147-byte patch straight from the uint8 batch
// (L1/L2 resident), apply the input normalisation (x/127.5 - 1: the reference's `preprocess_input`, C13), convert
// to bf16 and write a [128 pixels x 192 k] tile into shared memory in the canonical SWIZZLE_128B layout
// (3 k-blocks of 128 rows x 128 B).  That single tile is
//   * the K-major A operand of the forward GEMM   y[px, co]  = sum_k patch[px, k] * W[co, k]      (tcgen05, M=128, N=64)
//   * the MN-major B operand of the weight-gradient GEMM  dW[co, k] = sum_px dy[px, co] * patch[px, k]  (N=192)
// so the preprocess kernel, the bf16 copy of the input and cuDNN's padded-NHWC stem kernels all disappear.
// Forward fuses the BatchNorm statistics like conv_igemm.cuh; wgrad keeps its 64x192 fp32 accumulator in TMEM for
// the CTA's whole pixel range and flushes once with atomics.
#include <stdexcept>
#include <string>

#include "conv_api.h"
#include "ptx.cuh"

namespace b200 {

constexpr int kStemK = 192;        // 147 padded to 3 x 64
constexpr int kStemStages = 3;
constexpr int kStemTileBytes = 3 * 128 * 128;  // 48 KB im2col tile

__device__ __forceinline__ void stem_build_rows(const StemParams& p, uint8_t* tile, int tile_idx, int bt) {
  // builder thread bt in [0, 256): pixel row m = bt & 127, k-chunks [half*12, half*12 + 12)
  const int m = bt & 127;
  const int half = bt >> 7;
  const int64_t px = (int64_t)tile_idx * 128 + m;
  const bool live = px < p.M;
  int n = 0, ho = 0, wo = 0;
  if (live) {
    const int hw = p.Ho * p.Wo;
    n = (int)(px / hw);
    const int rem = (int)(px - (int64_t)n * hw);
    ho = rem / p.Wo;
    wo = rem - ho * p.Wo;
  }
  const int h0 = 2 * ho - 3, w0 = 2 * wo - 3;
  const uint8_t* img = p.x + (int64_t)n * p.H * p.W * 3;
#pragma unroll
  for (int cc = 0; cc < 12; ++cc) {
    uint32_t w32[4] = {0u, 0u, 0u, 0u};
    if (half == 0) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = cc * 8 + e;  // 0..95 < 147
        const int r = k / 21, j = k % 21, s = j / 3;
        const int h = h0 + r, w = w0 + s;
        float f = 0.f;
        if (live && h >= 0 && h < p.H && w >= 0 && w < p.W) f = fmaf((float)img[((int64_t)h * p.W + w0) * 3 + j], p.mul, p.add);
        const uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(f));
        w32[e >> 1] |= b << (16 * (e & 1));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int k = 96 + cc * 8 + e;
        if (k < 147) {
          const int r = k / 21, j = k % 21, s = j / 3;
          const int h = h0 + r, w = w0 + s;
          float f = 0.f;
          if (live && h >= 0 && h < p.H && w >= 0 && w < p.W) f = fmaf((float)img[((int64_t)h * p.W + w0) * 3 + j], p.mul, p.add);
          const uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16(f));
          w32[e >> 1] |= b << (16 * (e & 1));
        }
      }
    }
    const int chunk = half * 12 + cc;    // 16-byte chunk index along K, 0..23
    const int kb = chunk >> 3;           // k-block (64 k each)
    const int c8 = chunk & 7;
    *reinterpret_cast<uint4*>(tile + kb * (128 * 128) + m * 128 + ((c8 ^ (m & 7)) << 4)) =
        make_uint4(w32[0], w32[1], w32[2], w32[3]);
  }
}

// ------------------------------------------------------------------------------------------------ forward
struct StemFwdSmem {
  static constexpr int kB = 3 * 64 * 128;            // weights: 3 k-blocks of [64 co x 64 k]
  static constexpr int kStaging = 2 * 128 * 128;
  static constexpr int kTotal = kStemStages * kStemTileBytes + kB + kStaging + 256 + 2 * 64 * 4;
};

__global__ void __launch_bounds__(512, 1)
stem_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStemStages * kStemTileBytes;
  uint8_t* sStage = sB + StemFwdSmem::kB;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + StemFwdSmem::kStaging);
  uint64_t* full_bar = bars;                // [3]  builders -> MMA
  uint64_t* empty_bar = bars + 3;           // [3]  MMA -> builders
  uint64_t* tfull_bar = bars + 6;           // [2]
  uint64_t* tempty_bar = bars + 8;          // [2]
  uint64_t* w_bar = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);
  float* sStat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [2][64]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmY);
    for (int i = 0; i < kStemStages; ++i) {
      mbar_init(&full_bar[i], 8);   // one arrival per builder warp
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 128; i += blockDim.x) sStat[i] = 0.f;
  if (warp == 2) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // weights stay resident for the whole kernel
    mbar_arrive_expect_tx(w_bar, StemFwdSmem::kB);
    for (int kb = 0; kb < 3; ++kb) tma_load_2d(sB + kb * (64 * 128), &tmW, w_bar, kb * 64, 0);
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, 64, 0, 0);
    mbar_wait(w_bar, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 64;
#pragma unroll
      for (int kb = 0; kb < 3; ++kb) {
        const uint64_t da = umma_desc_sw128(smem_u32(sA + stage * kStemTileBytes + kb * (128 * 128)), 16, 1024);
        const uint64_t db = umma_desc_sw128(smem_u32(sB + kb * (64 * 128)), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);
      umma_commit(&tfull_bar[acc]);
      if (++stage == kStemStages) { stage = 0; phase ^= 1; }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && warp < 8) {
    // ---------------------------------------------------------------- epilogue: TMEM -> bf16 -> TMA store (+ stats)
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int etid = threadIdx.x - 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ctr = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++ctr) {
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      uint8_t* sbuf = sStage + (ctr & 1) * (128 * 128);
      if (etid == 0) tma_store_wait_read<1>();
      named_bar_sync(1, 128);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * 64;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + h * 32, r);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
          v.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
          v.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
          v.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
          *reinterpret_cast<uint4*>(sbuf + row * 128 + (((h * 4 + j) ^ (row & 7)) << 4)) = v;
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(2, 128);
      if (etid == 0) {
        tma_store_2d(&tmY, sbuf, 0, tile * 128);
        tma_store_commit();
      }
      if (p.stat_sum != nullptr) {
        const int cp = etid & 31, rg = etid >> 5;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
        const int64_t left = p.M - (int64_t)tile * 128;
        const int valid = left < 128 ? (int)left : 128;
        const int r_end = min(rg * 32 + 32, valid);
#pragma unroll 8
        for (int r = rg * 32; r < r_end; ++r) {
          const uint32_t w = *reinterpret_cast<const uint32_t*>(sbuf + r * 128 + (((cp >> 2) ^ (r & 7)) << 4) + ((cp & 3) << 2));
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
          s0 += f.x; s1 += f.y;
          q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
        }
        atomicAdd(&sStat[cp * 2], s0);
        atomicAdd(&sStat[cp * 2 + 1], s1);
        atomicAdd(&sStat[64 + cp * 2], q0);
        atomicAdd(&sStat[64 + cp * 2 + 1], q1);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.stat_sum != nullptr) {
      named_bar_sync(3, 128);
      if (etid < 64) atomicAdd(p.stat_sum + etid, sStat[etid]);
      else atomicAdd(p.stat_sqsum + (etid - 64), sStat[etid]);
    }
    if (etid == 0) tma_store_wait_all<0>();
  } else if (warp >= 8) {
    // ---------------------------------------------------------------- im2col builders (256 threads)
    const int bt = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      stem_build_rows(p, sA + stage * kStemTileBytes, tile, bt);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      if (++stage == kStemStages) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
struct StemWgSmem {
  static constexpr int kStage = kStemTileBytes + 128 * 128;   // im2col tile + dY tile
  static constexpr int kTotal = kStemStages * kStage + 256;
};

__global__ void __launch_bounds__(512, 1)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ StemParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStemStages * StemWgSmem::kStage);
  uint64_t* full_bar = bars;      // [3]  8 builder warps + TMA producer
  uint64_t* empty_bar = bars + 3; // [3]
  uint64_t* tfull_bar = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmDY);
    for (int i = 0; i < kStemStages; ++i) {
      mbar_init(&full_bar[i], 9);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // contiguous tile range per CTA
  const int per = (p.num_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per;
  const int t1 = min(t0 + per, p.num_tiles);

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      mbar_arrive_expect_tx(&full_bar[stage], 128 * 128);
      tma_load_2d(smem + stage * StemWgSmem::kStage + kStemTileBytes, &tmDY, &full_bar[stage], 0, tile * 128);
      if (++stage == kStemStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = umma_idesc_bf16(128, kStemK, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t b0 = smem_u32(smem + stage * StemWgSmem::kStage);
      const uint32_t a0 = b0 + kStemTileBytes;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t da = umma_desc_sw128(a0 + ks * 2048, 0, 1024);              // 64 co; rows 64..127 mirror
        const uint64_t db = umma_desc_sw128(b0 + ks * 2048, 128 * 128, 1024);      // 3 chunks of 64 k
        umma_bf16(tmem_base, da, db, idesc, (tile != t0 || ks != 0) ? 1u : 0u);
      }
      umma_commit(&empty_bar[stage]);
      if (++stage == kStemStages) { stage = 0; phase ^= 1; }
    }
    umma_commit(tfull_bar);
  } else if (warp >= 4 && warp < 6) {
    // rows 0..63 of the accumulator = output channels
    if (t1 > t0) {
      mbar_wait(tfull_bar, 0);
      tc_fence_after();
      const int co = (warp - 4) * 32 + lane;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>((warp - 4) * 32) << 16);
#pragma unroll 1
      for (int c32 = 0; c32 < kStemK / 32; ++c32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c32 * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int k = c32 * 32 + i;
          if (k < 147) atomicAdd(p.dw + ((k / 3) * 64 + co) * 3 + (k % 3), __uint_as_float(v[i]));
        }
      }
    }
  } else if (warp >= 8) {
    const int bt = threadIdx.x - 256;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = t0; tile < t1; ++tile) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      stem_build_rows(p, smem + stage * StemWgSmem::kStage, tile, bt);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_bar[stage]);
      if (++stage == kStemStages) { stage = 0; phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

static void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

void stem_fwd_launch(const StemPlanRaw& pl, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    check(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, StemFwdSmem::kTotal),
          "stem_fwd attr");
    attr = true;
  }
  stem_fwd_kernel<<<pl.grid, 512, StemFwdSmem::kTotal, s>>>(pl.tmW, pl.tmY, pl.p);
  check(cudaGetLastError(), "stem_fwd launch");
}
void stem_wgrad_launch(const StemPlanRaw& pl, cudaStream_t s) {
  static bool attr = false;
  if (!attr) {
    check(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, StemWgSmem::kTotal),
          "stem_wgrad attr");
    attr = true;
  }
  stem_wgrad_kernel<<<pl.grid, 512, StemWgSmem::kTotal, s>>>(pl.tmY, pl.p);
  check(cudaGetLastError(), "stem_wgrad launch");
}

}  // namespace b200
