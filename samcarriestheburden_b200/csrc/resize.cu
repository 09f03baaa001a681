// Image ingest on the GPU: the antialiased bilinear uint8 resize of
// ResizeLongestSide.apply_image (reference: segment_anything/utils/transforms.py:26-31 = torchvision `resize` of a PIL
// image, called from SamPredictor.set_image predictor.py:54-57 and scripts/generate_img_embeddings.py:43-45).
//
// The arithmetic is Pillow's ImagingResample (src/libImaging/Resample.c, un-vendored dependency of the reference,
// pillow pinned in environment.yml) with the bilinear (triangle) filter, restated from its published algorithm:
//   per axis: scale = in / out, filterscale = max(scale, 1), support = filterscale, ksize = ceil(support) * 2 + 1;
//   per output sample: center = (o + 0.5) * scale, taps [xmin, xmin + n) = [trunc(center - support + 0.5),
//   trunc(center + support + 0.5)) clipped to the image, weights triangle((x - center + 0.5) / filterscale) normalised
//   to sum 1 in double precision, then quantised to 22-bit fixed point (round half away from zero);
//   a pass computes clip8((2^21 + sum_k pixel_k * coeff_k) >> 22); horizontal pass first, its uint8 result feeds the
//   vertical pass; a pass whose size does not change is skipped.
// The coefficient tables are built on the host in double precision (resize_coeffs_host, the same operations in the same
// order as Pillow's precompute_coeffs / normalize_coeffs_8bpc); the passes are integer, so the result is bit-exact.
#include "common.cuh"
#include "kernels.h"
#include <cmath>
#include <vector>

namespace b200sam {

namespace {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

B200SAM_DEVINL uint8_t clip8(int v) {
  v >>= RS_PRECISION_BITS;
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [H, W, C] -> out [H, out_w, C]; one thread per output byte (x * C + c fastest: coalesced stores)
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ in, int H, int W, int C,
                                                       const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, int ksize, int out_w,
                                                       uint8_t* __restrict__ out) {
  const int row_bytes = out_w * C;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (j >= row_bytes) return;
  const int xx = j / C, c = j - xx * C;
  const int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
  const uint8_t* src = in + (static_cast<size_t>(y) * W + xmin) * C + c;
  const int32_t* k = kk + static_cast<size_t>(xx) * ksize;
  int ss = 1 << (RS_PRECISION_BITS - 1);
  for (int t = 0; t < n; ++t) ss += static_cast<int>(src[static_cast<size_t>(t) * C]) * __ldg(k + t);
  out[static_cast<size_t>(y) * row_bytes + j] = clip8(ss);
}

// in [H, W, C] -> out [out_h, W, C] (chw = 0) or [C, out_h, W] (chw = 1, what the encoder's preprocess kernel reads)
__global__ void __launch_bounds__(256) resize_v_kernel(const uint8_t* __restrict__ in, int H, int W, int C,
                                                       const int32_t* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, int ksize, int out_h,
                                                       uint8_t* __restrict__ out, int chw) {
  const int row_bytes = W * C;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int yy = blockIdx.y;
  if (j >= row_bytes) return;
  int ss = 1 << (RS_PRECISION_BITS - 1);
  if (bounds != nullptr) {
    const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
    const uint8_t* src = in + static_cast<size_t>(ymin) * row_bytes + j;
    const int32_t* k = kk + static_cast<size_t>(yy) * ksize;
    for (int t = 0; t < n; ++t) ss += static_cast<int>(src[static_cast<size_t>(t) * row_bytes]) * __ldg(k + t);
  } else {  // no vertical resampling: layout conversion only
    ss += static_cast<int>(in[static_cast<size_t>(yy) * row_bytes + j]) << RS_PRECISION_BITS;
  }
  const uint8_t v = clip8(ss);
  if (chw) {
    const int x = j / C, c = j - x * C;
    out[(static_cast<size_t>(c) * out_h + yy) * W + x] = v;
  } else {
    out[static_cast<size_t>(yy) * row_bytes + j] = v;
  }
}

}  // namespace

int resize_ksize(int in_size, int out_size) {
  if (in_size <= 0 || out_size <= 0) return 0;
  double filterscale = static_cast<double>(static_cast<float>(in_size)) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  return static_cast<int>(std::ceil(support)) * 2 + 1;
}

int resize_coeffs_host(int in_size, int out_size, int32_t* bounds, int32_t* kk) {
  B200SAM_REQUIRE(in_size > 0 && out_size > 0 && bounds != nullptr && kk != nullptr, "resize_coeffs: bad arguments");
  const double scale = static_cast<double>(static_cast<float>(in_size) - 0.0f) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  std::vector<double> pre(static_cast<size_t>(ksize));
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = 0.0 + (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    int x = 0;
    for (; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;  // bilinear (triangle) filter
      pre[x] = w;
      ww += w;
    }
    for (x = 0; x < xmax; ++x)
      if (ww != 0.0) pre[x] /= ww;
    for (; x < ksize; ++x) pre[x] = 0.0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
    for (x = 0; x < ksize; ++x) {
      const double v = pre[x];
      kk[static_cast<size_t>(xx) * ksize + x] =
          static_cast<int32_t>(v < 0.0 ? -0.5 + v * (1 << RS_PRECISION_BITS) : 0.5 + v * (1 << RS_PRECISION_BITS));
    }
  }
  return 0;
}

int resize_u8(const uint8_t* in, int H, int W, int C, const int32_t* xbounds, const int32_t* xkk, int xksize,
              const int32_t* ybounds, const int32_t* ykk, int yksize, int out_h, int out_w, uint8_t* tmp, uint8_t* out,
              int out_chw, cudaStream_t stream) {
  B200SAM_REQUIRE(in != nullptr && out != nullptr && H > 0 && W > 0 && C > 0 && out_h > 0 && out_w > 0,
                  "resize: bad arguments (H=%d W=%d C=%d out=%dx%d)", H, W, C, out_h, out_w);
  B200SAM_REQUIRE(H <= 65535 && out_h <= 65535, "resize: at most 65535 rows, got %d -> %d", H, out_h);
  const bool horiz = xbounds != nullptr, vert = ybounds != nullptr;
  B200SAM_REQUIRE(horiz || out_w == W, "resize: width changes (%d -> %d) but no horizontal coefficients", W, out_w);
  B200SAM_REQUIRE(vert || out_h == H, "resize: height changes (%d -> %d) but no vertical coefficients", H, out_h);
  B200SAM_REQUIRE(!horiz || (xkk != nullptr && xksize > 0), "resize: missing horizontal coefficient table");
  B200SAM_REQUIRE(!vert || (ykk != nullptr && yksize > 0), "resize: missing vertical coefficient table");
  const uint8_t* vsrc = in;
  if (horiz) {
    // Pillow runs the horizontal pass first; if nothing follows (same height, HWC out) it writes the result directly
    const bool last = !vert && !out_chw;
    uint8_t* hdst = last ? out : tmp;
    B200SAM_REQUIRE(hdst != nullptr, "resize: tmp buffer [H, out_w, C] required");
    dim3 grid((out_w * C + 255) / 256, H);
    resize_h_kernel<<<grid, 256, 0, stream>>>(in, H, W, C, xbounds, xkk, xksize, out_w, hdst);
    if (last) { B200SAM_CHECK_CUDA(cudaGetLastError()); return 0; }
    vsrc = tmp;
  }
  dim3 grid((out_w * C + 255) / 256, out_h);
  resize_v_kernel<<<grid, 256, 0, stream>>>(vsrc, H, out_w, C, vert ? ybounds : nullptr, ykk, yksize, out_h, out,
                                            out_chw);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
