// Fused mask post-processing (reference: segment_anything/sam_mask_decoder_head.py:99-135 ==
// modeling/sam.py:133-162, threshold sam.py:19, and utils/seg_refinement.py:111):
//
//   low-res logits 256x256 --bilinear--> 1024x1024 --crop[:in_h,:in_w]--bilinear--> out_h x out_w --(> thr)--> bool
//   (+ optional 'nearest-exact' resample of the bool mask to the U-Net grid, e.g. 384x224)
//
// The reference materialises both fp32 intermediates (4 MiB + 4*H0*W0 bytes per mask); here every output
// pixel composes the two align_corners=False bilinear stages on the fly from the 256 KiB logit plane
// (L1/L2 resident), so HBM sees 256 KiB in and 1 byte per output pixel out.
#include "common.cuh"
#include "kernels.h"

namespace b200sam {

namespace {

struct UpParams {
  const float* low;  // [n, L, L]
  int L;             // 256
  int S;             // 1024 (encoder input size)
  int in_h, in_w;    // crop in the S x S frame
  int out_h, out_w;  // native resolution
  float s1;          // L / S
  float sy2, sx2;    // in_h / out_h, in_w / out_w
  float thr;
};

// torch area_pixel_compute_source_index(align_corners=False, cubic=False): max(0, scale*(dst+0.5)-0.5)
B200SAM_DEVINL void src_index(float scale, int dst, int in_size, int& i0, int& i1, float& w0, float& w1) {
  float r = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  r = r < 0.0f ? 0.0f : r;
  i0 = static_cast<int>(r);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  w1 = r - static_cast<float>(i0);
  w0 = 1.0f - w1;
}

// value of the (virtual) S x S stage-1 image at (Y, X)
B200SAM_DEVINL float stage1(const UpParams& p, const float* __restrict__ plane, int Y, int X) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(p.s1, Y, p.L, y0, y1, wy0, wy1);
  src_index(p.s1, X, p.L, x0, x1, wx0, wx1);
  const float a = __ldg(plane + y0 * p.L + x0), b = __ldg(plane + y0 * p.L + x1);
  const float c = __ldg(plane + y1 * p.L + x0), d = __ldg(plane + y1 * p.L + x1);
  return wy0 * (wx0 * a + wx1 * b) + wy1 * (wx0 * c + wx1 * d);
}

B200SAM_DEVINL float final_value(const UpParams& p, const float* __restrict__ plane, int oy, int ox) {
  int y0, y1, x0, x1;
  float wy0, wy1, wx0, wx1;
  src_index(p.sy2, oy, p.in_h, y0, y1, wy0, wy1);
  src_index(p.sx2, ox, p.in_w, x0, x1, wx0, wx1);
  const float a = stage1(p, plane, y0, x0), b = stage1(p, plane, y0, x1);
  const float c = stage1(p, plane, y1, x0), d = stage1(p, plane, y1, x1);
  return wy0 * (wx0 * a + wx1 * b) + wy1 * (wx0 * c + wx1 * d);
}

// one thread = 4 horizontally adjacent output pixels -> one 32-bit store of 4 mask bytes
__global__ void __launch_bounds__(256) upscale_threshold_kernel(UpParams p, uint8_t* __restrict__ mask_out,
                                                                float* __restrict__ logits_out) {
  const int n = blockIdx.z;
  const float* plane = p.low + static_cast<size_t>(n) * p.L * p.L;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y;
  const int ox0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (oy >= p.out_h || ox0 >= p.out_w) return;
  const size_t obase = (static_cast<size_t>(n) * p.out_h + oy) * p.out_w + ox0;
  float v[4];
  const int cnt = min(4, p.out_w - ox0);
#pragma unroll
  for (int k = 0; k < 4; ++k) v[k] = k < cnt ? final_value(p, plane, oy, ox0 + k) : 0.0f;
  if (mask_out != nullptr) {
    if (cnt == 4 && ((obase & 3) == 0)) {
      const uint32_t w = (v[0] > p.thr ? 1u : 0u) | (v[1] > p.thr ? 0x100u : 0u) | (v[2] > p.thr ? 0x10000u : 0u) |
                         (v[3] > p.thr ? 0x1000000u : 0u);
      *reinterpret_cast<uint32_t*>(mask_out + obase) = w;
    } else {
      for (int k = 0; k < cnt; ++k) mask_out[obase + k] = v[k] > p.thr;
    }
  }
  if (logits_out != nullptr)
    for (int k = 0; k < cnt; ++k) logits_out[obase + k] = v[k];
}

// 'nearest-exact' resample of the thresholded native-resolution mask: src = floor((dst + 0.5) * in / out)
__global__ void __launch_bounds__(256) nearest_exact_kernel(UpParams p, uint8_t* __restrict__ small_out, int sh, int sw,
                                                            float ny, float nx) {
  const int n = blockIdx.z;
  const float* plane = p.low + static_cast<size_t>(n) * p.L * p.L;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (y >= sh || x >= sw) return;
  const int sy = min(static_cast<int>(floorf((static_cast<float>(y) + 0.5f) * ny)), p.out_h - 1);
  const int sx = min(static_cast<int>(floorf((static_cast<float>(x) + 0.5f) * nx)), p.out_w - 1);
  small_out[(static_cast<size_t>(n) * sh + y) * sw + x] = final_value(p, plane, sy, sx) > p.thr;
}

}  // namespace

int upscale_threshold(const float* low_res, int n, int low, int img_size, int in_h, int in_w, int out_h, int out_w,
                      float thresh, uint8_t* mask_out, float* logits_out, uint8_t* small_out, int small_h,
                      int small_w, cudaStream_t stream) {
  B200SAM_REQUIRE(n >= 0 && low > 0 && img_size > 0, "upscale: bad sizes n=%d low=%d img=%d", n, low, img_size);
  B200SAM_REQUIRE(in_h > 0 && in_w > 0 && in_h <= img_size && in_w <= img_size && out_h > 0 && out_w > 0,
                  "upscale: bad crop/output size in=(%d,%d) out=(%d,%d)", in_h, in_w, out_h, out_w);
  B200SAM_REQUIRE(n <= 65535, "upscale: at most 65535 masks per launch, got %d", n);
  if (n == 0) return 0;
  UpParams p;
  p.low = low_res;
  p.L = low;
  p.S = img_size;
  p.in_h = in_h; p.in_w = in_w; p.out_h = out_h; p.out_w = out_w;
  p.s1 = static_cast<float>(low) / static_cast<float>(img_size);
  p.sy2 = static_cast<float>(in_h) / static_cast<float>(out_h);
  p.sx2 = static_cast<float>(in_w) / static_cast<float>(out_w);
  p.thr = thresh;
  if (mask_out != nullptr || logits_out != nullptr) {
    dim3 block(64, 4);
    dim3 grid((out_w + 255) / 256, (out_h + 3) / 4, n);
    upscale_threshold_kernel<<<grid, block, 0, stream>>>(p, mask_out, logits_out);
  }
  if (small_out != nullptr) {
    B200SAM_REQUIRE(small_h > 0 && small_w > 0, "upscale: bad nearest-exact size (%d,%d)", small_h, small_w);
    dim3 block(32, 8);
    dim3 grid((small_w + 31) / 32, (small_h + 7) / 8, n);
    nearest_exact_kernel<<<grid, block, 0, stream>>>(p, small_out, small_h, small_w,
                                                     static_cast<float>(out_h) / static_cast<float>(small_h),
                                                     static_cast<float>(out_w) / static_cast<float>(small_w));
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
