// Classifier-head and stem-tail kernels: the bandwidth-bound glue around the two GEMM-shaped
// ends of the network.  The FC layer itself runs on the tcgen05 implicit-GEMM kernels (a 1x1 "convolution" over a 1x1
// image, csrc/conv_igemm.cuh / conv_wgrad.cu); what lives here is everything the reference's Dense + softmax head
// (P1/02:175) and MaxPool backward need besides the GEMMs:
//   * softmax_ce_head: bf16 logits + fp32 bias -> loss, accuracy, bf16 dlogits (padded columns zero) in one pass;
//   * fc_bias_grad: column sums of dlogits;
//   * stem_pool_bn_bwd: 3x3/2 max-pool backward (argmax gather) fused with the stem BatchNorm+ReLU backward - the
//     112x112x64 pooled-gradient tensor is never written: pass 0 reduces sum(dz), sum(dz*y), pass 1 writes dy;
//     both gather from a shared-memory tile of (g1 + g2, argmax) so the <= 4 windows per pixel hit smem, not L2;
//   * pack_stem_weight: fp32 [49, 64, 3] master -> bf16 [64, 192] GEMM operand.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ops_api.h"

namespace b200 {

namespace {

struct alignas(16) bf8 {
  __nv_bfloat162 v[4];
};
__device__ __forceinline__ void unpack8h(const bf8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf8 pack8h(const float (&f)[8]) {
  bf8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

}  // namespace

// ------------------------------------------------------------------------------------------------ softmax CE head
// One warp per row.  logit[k] = bf16 logits16[row, k] + bias[k] for k < K (row pitch ld >= K, padded columns ignored).
__global__ void softmax_ce_head_kernel(const __nv_bfloat16* __restrict__ logits16, int ld,
                                       const float* __restrict__ bias, const int64_t* __restrict__ labels,
                                       float* __restrict__ logits32, __nv_bfloat16* __restrict__ dlogits16,
                                       float* __restrict__ loss_rows, float* __restrict__ stats, int B, int K,
                                       float grad_scale) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= B) return;
  const __nv_bfloat16* row = logits16 + (int64_t)warp * ld;
  const int label = (int)labels[warp];
  float mx = -3.0e38f;
  int amax = 0;
  for (int k = lane; k < K; k += 32) {
    const float v = __bfloat162float(row[k]) + bias[k];
    if (logits32 != nullptr) logits32[(int64_t)warp * K + k] = v;
    if (v > mx) { mx = v; amax = k; }
  }
  for (int o = 16; o; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
    if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
  }
  float se = 0.f;
  for (int k = lane; k < K; k += 32) se += __expf(__bfloat162float(row[k]) + bias[k] - mx);
  for (int o = 16; o; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
  const float lse = mx + __logf(se);
  const float inv = 1.f / se;
  if (dlogits16 != nullptr) {
    __nv_bfloat16* drow = dlogits16 + (int64_t)warp * ld;
    for (int k = lane; k < ld; k += 32) {
      float d = 0.f;
      if (k < K) {
        const float pk = __expf(__bfloat162float(row[k]) + bias[k] - mx) * inv;
        d = (pk - (k == label ? 1.f : 0.f)) * grad_scale;
      }
      drow[k] = __float2bfloat16(d);
    }
  }
  if (lane == 0) {
    const float l = lse - (__bfloat162float(row[label]) + bias[label]);
    if (loss_rows != nullptr) loss_rows[warp] = l;
    atomicAdd(stats, l);
    atomicAdd(stats + 1, amax == label ? 1.f : 0.f);
  }
}
void softmax_ce_head(const void* logits16, int ld, const float* bias, const int64_t* labels, float* logits32,
                     void* dlogits16, float* loss_rows, float* stats, int B, int K, float grad_scale, cudaStream_t s) {
  const int threads = 128;
  const int blocks = (B * 32 + threads - 1) / threads;
  softmax_ce_head_kernel<<<blocks, threads, 0, s>>>((const __nv_bfloat16*)logits16, ld, bias, labels, logits32,
                                                    (__nv_bfloat16*)dlogits16, loss_rows, stats, B, K, grad_scale);
}

// dbias[k] = sum_b dlogits16[b, k]   (thread per column: coalesced across k for every row)
__global__ void fc_bias_grad_kernel(const __nv_bfloat16* __restrict__ dl, int ld, float* __restrict__ dbias, int B, int K) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int b = 0;
  for (; b + 4 <= B; b += 4) {
    a0 += __bfloat162float(dl[(int64_t)(b + 0) * ld + k]);
    a1 += __bfloat162float(dl[(int64_t)(b + 1) * ld + k]);
    a2 += __bfloat162float(dl[(int64_t)(b + 2) * ld + k]);
    a3 += __bfloat162float(dl[(int64_t)(b + 3) * ld + k]);
  }
  for (; b < B; ++b) a0 += __bfloat162float(dl[(int64_t)b * ld + k]);
  dbias[k] = (a0 + a1) + (a2 + a3);
}
void fc_bias_grad(const void* dlogits16, int ld, float* dbias, int B, int K, cudaStream_t s) {
  fc_bias_grad_kernel<<<(K + 127) / 128, 128, 0, s>>>((const __nv_bfloat16*)dlogits16, ld, dbias, B, K);
}

// ------------------------------------------------------------------------------------------------ stem weight pack
// fp32 [49, 64, 3] (tap, co, c) -> bf16 [64, 192] with k = tap*3 + c (columns 147.. are zero)
__global__ void pack_stem_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 64 * 192) return;
  const int co = i / 192, k = i - co * 192;
  float v = 0.f;
  if (k < 147) v = w[((k / 3) * 64 + co) * 3 + (k % 3)];
  out[i] = __float2bfloat16(v);
}
void pack_stem_weight(const float* w, void* out, cudaStream_t s) {
  pack_stem_weight_kernel<<<(64 * 192 + 255) / 256, 256, 0, s>>>(w, (__nv_bfloat16*)out);
}

// ------------------------------------------------------------------------------------------------ stem tail backward
// Backward of  p = maxpool3x3/2( relu( y*scale + shift ) )  for the 64-channel stem, without materialising the pooled
// gradient on the 2Ho x 2Wo grid:
//   da[h, w]  = sum over the <= 4 windows containing (h, w) whose saved argmax is this position of (g1 + g2)[ho, wo]
//   dz        = da * [y*scale + shift > 0]
//   PASS 0:   sum_dz += dz, sum_dzy += dz * y          (per channel; one flush of atomics per CTA)
//   PASS 1:   dy = cA*dz + cB*y + cC
// A CTA walks 8x8 pooled tiles (16x16 input pixels): the 9x9 windows it needs are staged in shared memory as
// fp32 (g1 + g2) + the argmax byte, so the gather reads smem; y / dy move as coalesced 16-byte vectors.
constexpr int kPoolTile = 8;
constexpr int kPoolWin = kPoolTile + 1;

template <int PASS>
__global__ void __launch_bounds__(256)
stem_pool_bn_bwd_kernel(const uint8_t* __restrict__ idx, const __nv_bfloat16* __restrict__ g1,
                        const __nv_bfloat16* __restrict__ g2, const __nv_bfloat16* __restrict__ y,
                        const float* __restrict__ scale, const float* __restrict__ shift,
                        const float* __restrict__ cA, const float* __restrict__ cB, const float* __restrict__ cC,
                        __nv_bfloat16* __restrict__ dy, float* __restrict__ sum_dz, float* __restrict__ sum_dzy, int N,
                        int Ho, int Wo) {
  constexpr int C = 64;
  __shared__ __align__(16) float gS[kPoolWin * kPoolWin][C];
  __shared__ __align__(8) uint8_t iS[kPoolWin * kPoolWin][C];
  __shared__ float red[2][8][C];
  const int H = 2 * Ho, W = 2 * Wo;
  const int tiles_h = Ho / kPoolTile, tiles_w = Wo / kPoolTile;
  const int num_tiles = N * tiles_h * tiles_w;
  const int t = threadIdx.x;
  const int cv = t & 7;
  const int pl = t >> 3;  // 0..31: pixel lane; pixels p = pl + 32*i
  float sc[8], sh[8], A[8], B[8], Cc[8], t0[8], t1[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale[cv * 8 + k];
    sh[k] = shift[cv * 8 + k];
    t0[k] = t1[k] = 0.f;
    if (PASS == 1) {
      A[k] = cA[cv * 8 + k];
      B[k] = cB[cv * 8 + k];
      Cc[k] = cC[cv * 8 + k];
    }
  }
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int tb = tile % tiles_w;
    const int rest = tile / tiles_w;
    const int ta = rest % tiles_h;
    const int n = rest / tiles_h;
    const int a0 = ta * kPoolTile, b0 = tb * kPoolTile;
    __syncthreads();  // the previous tile's gather is done with gS / iS
    for (int item = t; item < kPoolWin * kPoolWin * 8; item += 256) {
      const int win = item >> 3, v8 = item & 7;
      const int wi = win / kPoolWin, wj = win - wi * kPoolWin;
      const int ho = a0 + wi, wo = b0 + wj;
      float gv[8];
      uint2 pk = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);  // position 255 never matches
#pragma unroll
      for (int k = 0; k < 8; ++k) gv[k] = 0.f;
      if (ho < Ho && wo < Wo) {
        const int64_t o = ((((int64_t)n * Ho + ho) * Wo + wo) * C) + v8 * 8;
        unpack8h(*reinterpret_cast<const bf8*>(g1 + o), gv);
        if (g2 != nullptr) {
          float hv[8];
          unpack8h(*reinterpret_cast<const bf8*>(g2 + o), hv);
#pragma unroll
          for (int k = 0; k < 8; ++k) gv[k] += hv[k];
        }
        pk = *reinterpret_cast<const uint2*>(idx + o);
      }
      *reinterpret_cast<float4*>(&gS[win][v8 * 8]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
      *reinterpret_cast<float4*>(&gS[win][v8 * 8 + 4]) = make_float4(gv[4], gv[5], gv[6], gv[7]);
      *reinterpret_cast<uint2*>(&iS[win][v8 * 8]) = pk;
    }
    __syncthreads();
    const int pw = pl & 15;
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
      const int ph = (pl >> 4) + 2 * i;
      const int h = 2 * a0 + ph, w = 2 * b0 + pw;
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      // local windows containing local row ph: wi in [ph/2, (ph+1)/2], position r = ph - (2*wi - 1)
      for (int wi = ph >> 1; wi <= ((ph + 1) >> 1); ++wi) {
        const int r = ph - (2 * wi - 1);
        for (int wj = pw >> 1; wj <= ((pw + 1) >> 1); ++wj) {
          const uint32_t pos = r * 3 + (pw - (2 * wj - 1));
          const int win = wi * kPoolWin + wj;
          const uint2 pk = *reinterpret_cast<const uint2*>(&iS[win][cv * 8]);
          const float4 ga = *reinterpret_cast<const float4*>(&gS[win][cv * 8]);
          const float4 gb = *reinterpret_cast<const float4*>(&gS[win][cv * 8 + 4]);
          const float gv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t a = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xFFu;
            acc[k] += (a == pos) ? gv[k] : 0.f;
          }
        }
      }
      const int64_t off = ((((int64_t)n * H + h) * W + w) * C) + cv * 8;
      float yy[8];
      unpack8h(*reinterpret_cast<const bf8*>(y + off), yy);
      float dz[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // bf16 rounding of the pooled gradient: same values as the materialised path (maxpool_bwd -> bf16 tensor)
        const float a = bf16_round(acc[k]);
        dz[k] = fmaf(yy[k], sc[k], sh[k]) > 0.f ? a : 0.f;
      }
      if (PASS == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          t0[k] += dz[k];
          t1[k] = fmaf(dz[k], yy[k], t1[k]);
        }
      } else {
        float o8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) o8[k] = fmaf(A[k], dz[k], fmaf(B[k], yy[k], Cc[k]));
        *reinterpret_cast<bf8*>(dy + off) = pack8h(o8);
      }
    }
  }
  if (PASS == 0) {
    // lanes with equal (lane & 7) own the same 8 channels: fold the 4 pixel lanes of a warp, then the 8 warps
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      t0[k] += __shfl_xor_sync(0xffffffffu, t0[k], 8);
      t0[k] += __shfl_xor_sync(0xffffffffu, t0[k], 16);
      t1[k] += __shfl_xor_sync(0xffffffffu, t1[k], 8);
      t1[k] += __shfl_xor_sync(0xffffffffu, t1[k], 16);
    }
    const int warp = t >> 5, lane = t & 31;
    __syncthreads();
    if (lane < 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        red[0][warp][lane * 8 + k] = t0[k];
        red[1][warp][lane * 8 + k] = t1[k];
      }
    }
    __syncthreads();
    if (t < 2 * C) {
      const int which = t / C, c = t % C;
      float a = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) a += red[which][wv][c];
      atomicAdd((which == 0 ? sum_dz : sum_dzy) + c, a);
    }
  }
}

// pass 0: reduce (sum_dz, sum_dzy accumulate);  pass 1: apply (dy written)
void stem_pool_bn_bwd(int pass, const void* idx, const void* g1, const void* g2, const void* y, const float* scale,
                      const float* shift, const float* cA, const float* cB, const float* cC, void* dy, float* sum_dz,
                      float* sum_dzy, int N, int Ho, int Wo, cudaStream_t s) {
  const int tiles = N * (Ho / kPoolTile) * (Wo / kPoolTile);
  int grid = 148 * 4;
  if (grid > tiles) grid = tiles;
  if (pass == 0)
    stem_pool_bn_bwd_kernel<0><<<grid, 256, 0, s>>>((const uint8_t*)idx, (const __nv_bfloat16*)g1,
                                                    (const __nv_bfloat16*)g2, (const __nv_bfloat16*)y, scale, shift,
                                                    cA, cB, cC, (__nv_bfloat16*)dy, sum_dz, sum_dzy, N, Ho, Wo);
  else
    stem_pool_bn_bwd_kernel<1><<<grid, 256, 0, s>>>((const uint8_t*)idx, (const __nv_bfloat16*)g1,
                                                    (const __nv_bfloat16*)g2, (const __nv_bfloat16*)y, scale, shift,
                                                    cA, cB, cC, (__nv_bfloat16*)dy, sum_dz, sum_dzy, N, Ho, Wo);
}

}  // namespace b200
