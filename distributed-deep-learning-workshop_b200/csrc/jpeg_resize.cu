// Batched resize for the GPU JPEG path (csrc/jpeg_decode.cpp): every image of the batch has its OWN decoded size; one
// launch resizes all of them to [n, OH, OW, 3] uint8 (HWC).  The filter is the one PIL's `resize(..., BILINEAR)` applies
// in the CPU loader (models/preprocess.py decode_image): a triangle filter whose support is stretched by the down-scale
// factor (i.e. anti-aliased when shrinking, plain 2-tap bilinear when enlarging), pixel centres at +0.5 - so a model sees
// the same pixels (up to rounding) whichever decode path fed it.
//   table[i] = (byte offset of image i in `src`, width, height, row pitch in bytes), interleaved RGB
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

__global__ void __launch_bounds__(256)
resize_triangle_batched_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ table,
                               uint8_t* __restrict__ out, int OH, int OW) {
  const int n = blockIdx.y;
  const int64_t off = table[n * 4 + 0];
  const int W = (int)table[n * 4 + 1], H = (int)table[n * 4 + 2];
  const int64_t pitch = table[n * 4 + 3];
  const uint8_t* img = src + off;
  const float sx = (float)W / OW, sy = (float)H / OH;
  const float fsx = fmaxf(sx, 1.f), fsy = fmaxf(sy, 1.f);   // filter scale: support = 1.0 * filterscale (PIL precompute_coeffs)
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < OH * OW; p += gridDim.x * blockDim.x) {
    const int ox = p % OW, oy = p / OW;
    const float cx = (ox + 0.5f) * sx, cy = (oy + 0.5f) * sy;
    int x0 = (int)(cx - fsx + 0.5f), x1 = (int)(cx + fsx + 0.5f);
    int y0 = (int)(cy - fsy + 0.5f), y1 = (int)(cy + fsy + 0.5f);
    x0 = max(x0, 0); y0 = max(y0, 0);
    x1 = min(x1, W); y1 = min(y1, H);
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, wsum = 0.f;
    for (int y = y0; y < y1; ++y) {
      const float wy = fmaxf(0.f, 1.f - fabsf((y - cy + 0.5f) / fsy));
      if (wy == 0.f) continue;
      const uint8_t* row = img + (int64_t)y * pitch;
      for (int x = x0; x < x1; ++x) {
        const float w = wy * fmaxf(0.f, 1.f - fabsf((x - cx + 0.5f) / fsx));
        const uint8_t* px = row + x * 3;
        acc0 = fmaf(w, (float)px[0], acc0);
        acc1 = fmaf(w, (float)px[1], acc1);
        acc2 = fmaf(w, (float)px[2], acc2);
        wsum += w;
      }
    }
    const float inv = wsum > 0.f ? 1.f / wsum : 0.f;
    uint8_t* o = out + ((int64_t)n * OH * OW + p) * 3;
    o[0] = (uint8_t)fminf(fmaxf(acc0 * inv + 0.5f, 0.f), 255.f);
    o[1] = (uint8_t)fminf(fmaxf(acc1 * inv + 0.5f, 0.f), 255.f);
    o[2] = (uint8_t)fminf(fmaxf(acc2 * inv + 0.5f, 0.f), 255.f);
  }
}

void resize_triangle_batched(const uint8_t* src, const int64_t* table_dev, uint8_t* out, int n, int OH, int OW,
                             cudaStream_t s) {
  if (n <= 0) return;
  int bx = (OH * OW + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid(bx, n);
  resize_triangle_batched_kernel<<<grid, 256, 0, s>>>(src, table_dev, out, OH, OW);
}

}  // namespace b200
