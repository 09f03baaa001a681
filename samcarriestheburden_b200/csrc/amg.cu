// Mask statistics of the upstream SamPredictor / automatic-mask-generator surface:
//   calculate_stability_score (reference: segment_anything/utils/amg.py:154-176): IoU of the masks obtained by
//     thresholding the logits at thr + off and thr - off = count(x > thr + off) / count(x > thr - off), the two int32
//     counts divided in fp32 (0 / 0 = NaN, like the reference); the caller forms the two thresholds (in double, rounded
//     to fp32 once, which is what torch does with the Python scalar);
//   batched_mask_to_box (amg.py:303-346): XYXY box of the non-zero pixels of each mask, [0, 0, 0, 0] for an empty mask.
// One pass over the data each: grid = (chunks, masks), per-block reduction, integer atomics into a scratch slot per
// mask, then a finalize kernel.  Integer results: bit-exact.
#include "common.cuh"
#include "kernels.h"
#include <limits.h>
#include <algorithm>

namespace b200sam {

namespace {

constexpr int AMG_THREADS = 256;

__global__ void amg_init_kernel(int32_t* scratch, int n, int box) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (box) {
    scratch[4 * i + 0] = INT_MAX; scratch[4 * i + 1] = INT_MAX; scratch[4 * i + 2] = -1; scratch[4 * i + 3] = -1;
  } else {
    scratch[2 * i] = 0; scratch[2 * i + 1] = 0;
  }
}

__global__ void __launch_bounds__(AMG_THREADS) stability_count_kernel(const float* __restrict__ x, size_t HW, float hi,
                                                                      float lo, int32_t* __restrict__ scratch) {
  const float* p = x + static_cast<size_t>(blockIdx.y) * HW;
  int c_hi = 0, c_lo = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * AMG_THREADS + threadIdx.x; i < HW;
       i += static_cast<size_t>(gridDim.x) * AMG_THREADS) {
    const float v = __ldg(p + i);
    c_hi += v > hi;
    c_lo += v > lo;
  }
  c_hi = __reduce_add_sync(0xffffffffu, c_hi);
  c_lo = __reduce_add_sync(0xffffffffu, c_lo);
  if ((threadIdx.x & 31) == 0 && (c_hi | c_lo)) {
    if (c_hi) atomicAdd(&scratch[2 * blockIdx.y], c_hi);
    if (c_lo) atomicAdd(&scratch[2 * blockIdx.y + 1], c_lo);
  }
}

__global__ void stability_finalize_kernel(const int32_t* __restrict__ scratch, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = __fdiv_rn(static_cast<float>(scratch[2 * i]), static_cast<float>(scratch[2 * i + 1]));
}

__global__ void __launch_bounds__(AMG_THREADS) box_scan_kernel(const uint8_t* __restrict__ m, int H, int W,
                                                               int32_t* __restrict__ scratch) {
  const uint8_t* p = m + static_cast<size_t>(blockIdx.y) * H * W;
  int x0 = INT_MAX, y0 = INT_MAX, x1 = -1, y1 = -1;
  // a thread walks whole rows segments: row = blockIdx.x + k * gridDim.x, columns strided by the block
  for (int y = blockIdx.x; y < H; y += gridDim.x) {
    const uint8_t* row = p + static_cast<size_t>(y) * W;
    for (int x = threadIdx.x; x < W; x += AMG_THREADS) {
      if (row[x]) {
        x0 = min(x0, x); x1 = max(x1, x);
        y0 = min(y0, y); y1 = max(y1, y);
      }
    }
  }
  x0 = __reduce_min_sync(0xffffffffu, x0); y0 = __reduce_min_sync(0xffffffffu, y0);
  x1 = __reduce_max_sync(0xffffffffu, x1); y1 = __reduce_max_sync(0xffffffffu, y1);
  if ((threadIdx.x & 31) == 0 && x1 >= 0) {
    int32_t* s = scratch + 4 * blockIdx.y;
    atomicMin(&s[0], x0); atomicMin(&s[1], y0); atomicMax(&s[2], x1); atomicMax(&s[3], y1);
  }
}

__global__ void box_finalize_kernel(const int32_t* __restrict__ scratch, int n, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t* s = scratch + 4 * i;
  const bool empty = s[2] < s[0] || s[3] < s[1];
  out[4 * i + 0] = empty ? 0 : s[0];
  out[4 * i + 1] = empty ? 0 : s[1];
  out[4 * i + 2] = empty ? 0 : s[2];
  out[4 * i + 3] = empty ? 0 : s[3];
}

}  // namespace

int stability_score(const float* logits, int n, int H, int W, float threshold_hi, float threshold_lo,
                    float* score_out, int32_t* scratch, cudaStream_t stream) {
  B200SAM_REQUIRE(n >= 0 && H > 0 && W > 0, "stability_score: bad sizes n=%d H=%d W=%d", n, H, W);
  if (n == 0) return 0;
  B200SAM_REQUIRE(logits != nullptr && score_out != nullptr && scratch != nullptr, "stability_score: null pointer");
  B200SAM_REQUIRE(n <= 65535, "stability_score: at most 65535 masks per call, got %d", n);
  const size_t HW = static_cast<size_t>(H) * W;
  amg_init_kernel<<<(n + 127) / 128, 128, 0, stream>>>(scratch, n, 0);
  const unsigned chunks = static_cast<unsigned>(std::min<size_t>(64, (HW + AMG_THREADS * 16 - 1) / (AMG_THREADS * 16)));
  dim3 grid(chunks, n);
  stability_count_kernel<<<grid, AMG_THREADS, 0, stream>>>(logits, HW, threshold_hi, threshold_lo, scratch);
  stability_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(scratch, n, score_out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mask_to_box(const uint8_t* masks, int n, int H, int W, long long* boxes_out, int32_t* scratch, cudaStream_t stream) {
  B200SAM_REQUIRE(n >= 0 && H > 0 && W > 0, "mask_to_box: bad sizes n=%d H=%d W=%d", n, H, W);
  if (n == 0) return 0;
  B200SAM_REQUIRE(masks != nullptr && boxes_out != nullptr && scratch != nullptr, "mask_to_box: null pointer");
  B200SAM_REQUIRE(n <= 65535, "mask_to_box: at most 65535 masks per call, got %d", n);
  amg_init_kernel<<<(n + 127) / 128, 128, 0, stream>>>(scratch, n, 1);
  dim3 grid(static_cast<unsigned>(std::min(H, 64)), n);
  box_scan_kernel<<<grid, AMG_THREADS, 0, stream>>>(masks, H, W, scratch);
  box_finalize_kernel<<<(n + 127) / 128, 128, 0, stream>>>(scratch, n, boxes_out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
