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

// ---------------------------------------------------------------------------------------------------------------
// U-Net ingest: cv2.resize(uint8 grey, (W, H), INTER_LINEAR) of scripts/save_refined_segmentations.py:63
// followed by `.float() / 255` (:64) and `(img - IMG_MEAN) / IMG_STD` (:67).  OpenCV is an un-vendored dependency of the
// reference (environment.yml: opencv 4.9); its published uint8 linear path (modules/imgproc/src/resize.cpp, resizeGeneric_
// with HResizeLinear / VResizeLinear<uchar, int, short, FixedPtCast<int, uchar, 22>>) is restated:
//   per axis: f = (float)((d + 0.5) * (double)(in / out) - 0.5), s = floor(f), f -= s; 11-bit weights
//   saturate_cast<short>((1 - f) * 2048), saturate_cast<short>(f * 2048) (round half to even);
//   x axis: s < 0 -> (s, f) = (0, 0); s >= in - 1 -> (in - 1, 0); y axis: the two ROWS are clamped to the image but the
//   weights are kept; horizontal pass exact in int32; vertical pass
//   uchar((((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2).
// Tables are built on the host (cvresize_coeffs_host, IEEE double / float in OpenCV's order); the passes are integer, so
// the uint8 result is bit-exact (tests/golden/cv2resize_golden.npz was written by cv2 itself).
namespace {

__global__ void __launch_bounds__(256) cvresize_linear_kernel(const uint8_t* __restrict__ in, int H, int W,
                                                              const int32_t* __restrict__ xi, const int32_t* __restrict__ xw,
                                                              const int32_t* __restrict__ yi, const int32_t* __restrict__ yw,
                                                              int out_h, int out_w, uint8_t* __restrict__ out_u8,
                                                              float* __restrict__ out_norm, float mean, float sd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int n = blockIdx.z;
  if (x >= out_w) return;
  const uint8_t* img = in + static_cast<size_t>(n) * H * W;
  const int x0 = __ldg(xi + 2 * x), x1 = __ldg(xi + 2 * x + 1), a0 = __ldg(xw + 2 * x), a1 = __ldg(xw + 2 * x + 1);
  const int y0 = __ldg(yi + 2 * y), y1 = __ldg(yi + 2 * y + 1), b0 = __ldg(yw + 2 * y), b1 = __ldg(yw + 2 * y + 1);
  const uint8_t* r0 = img + static_cast<size_t>(y0) * W;
  const uint8_t* r1 = img + static_cast<size_t>(y1) * W;
  const int s0 = static_cast<int>(r0[x0]) * a0 + static_cast<int>(r0[x1]) * a1;
  const int s1 = static_cast<int>(r1[x0]) * a0 + static_cast<int>(r1[x1]) * a1;
  int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
  v = v < 0 ? 0 : (v > 255 ? 255 : v);
  const size_t o = (static_cast<size_t>(n) * out_h + y) * out_w + x;
  if (out_u8 != nullptr) out_u8[o] = static_cast<uint8_t>(v);
  if (out_norm != nullptr)  // fp32: u8 / 255, then (v - mean) / std, each op rounded like torch's elementwise kernels
    out_norm[o] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean), sd);
}

}  // namespace

int cvresize_coeffs_host(int in_size, int out_size, int clamp_weights, int32_t* idx2, int32_t* w2) {
  B200SAM_REQUIRE(in_size > 0 && out_size > 0 && idx2 != nullptr && w2 != nullptr, "cvresize_coeffs: bad arguments");
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  for (int d = 0; d < out_size; ++d) {
    volatile double pos = (d + 0.5) * scale;  // volatile: no fused multiply-add contraction across the two operations
    float f = static_cast<float>(pos - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    if (clamp_weights) {
      if (s < 0) { s = 0; f = 0.0f; }
      if (s >= in_size - 1) { s = in_size - 1; f = 0.0f; }
    }
    auto sat16 = [](float v) {
      const long r = std::lrint(v);  // round half to even (default rounding mode), like cvRound
      return static_cast<int32_t>(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
    };
    w2[2 * d] = sat16((1.0f - f) * 2048.0f);
    w2[2 * d + 1] = sat16(f * 2048.0f);
    const int i0 = s < 0 ? 0 : (s > in_size - 1 ? in_size - 1 : s);
    const int i1 = s + 1 < 0 ? 0 : (s + 1 > in_size - 1 ? in_size - 1 : s + 1);
    idx2[2 * d] = i0;
    idx2[2 * d + 1] = i1;
  }
  return 0;
}

int cvresize_linear_u8(const uint8_t* in, int n, int H, int W, const int32_t* xi, const int32_t* xw, const int32_t* yi,
                       const int32_t* yw, int out_h, int out_w, uint8_t* out_u8, float* out_norm, float mean, float sd,
                       cudaStream_t stream) {
  B200SAM_REQUIRE(in != nullptr && n > 0 && H > 0 && W > 0 && out_h > 0 && out_w > 0 && xi && xw && yi && yw,
                  "cvresize: bad arguments (n=%d H=%d W=%d out=%dx%d)", n, H, W, out_h, out_w);
  B200SAM_REQUIRE(out_u8 != nullptr || out_norm != nullptr, "cvresize: no output requested");
  B200SAM_REQUIRE(out_h <= 65535 && n <= 65535, "cvresize: at most 65535 output rows / images");
  B200SAM_REQUIRE(out_norm == nullptr || sd != 0.0f, "cvresize: std must be non-zero");
  dim3 grid((out_w + 255) / 256, out_h, n);
  cvresize_linear_kernel<<<grid, 256, 0, stream>>>(in, H, W, xi, xw, yi, yw, out_h, out_w, out_u8, out_norm, mean, sd);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// MedSAM ingest (scripts/generate_img_embeddings.py:49-62, the reference's default sam_type): cv2.resize(RGB uint8,
// (1024, 1024), INTER_CUBIC) -> (x - min) / clip(max - min, 1e-8) in float64 -> float32 [3, S, S].  OpenCV's published
// (IPP-less) uint8 cubic path: interpolateCubic (A = -0.75, fp32), 11-bit weights, HResizeCubic in int32 with clamped
// tap indices, VResizeCubic's vector body for uchar in fp32: ((S3 b3 + S2 b2) + S1 b1) + S0 b0, b = w / 2^22, round half
// to even, saturate.  The grey image is resized once; its three identical channels are written by the normalise pass.
namespace {

__global__ void __launch_bounds__(256) cvresize_cubic_kernel(const uint8_t* __restrict__ in, int H, int W,
                                                             const int32_t* __restrict__ xi, const int32_t* __restrict__ xw,
                                                             const int32_t* __restrict__ yi, const int32_t* __restrict__ yw,
                                                             int out_h, int out_w, uint8_t* __restrict__ out,
                                                             int32_t* __restrict__ minmax) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  int v = 0;
  if (x < out_w) {
    int xs[4], ws[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { xs[k] = __ldg(xi + 4 * x + k); ws[k] = __ldg(xw + 4 * x + k); }
    float acc = 0.0f;
#pragma unroll
    for (int k = 3; k >= 0; --k) {  // S3 first, like the nested multiply-adds of OpenCV's vector body
      const uint8_t* r = in + static_cast<size_t>(__ldg(yi + 4 * y + k)) * W;
      const int s = static_cast<int>(r[xs[0]]) * ws[0] + static_cast<int>(r[xs[1]]) * ws[1] +
                    static_cast<int>(r[xs[2]]) * ws[2] + static_cast<int>(r[xs[3]]) * ws[3];
      const float b = __fmul_rn(static_cast<float>(__ldg(yw + 4 * y + k)), 1.0f / (2048.0f * 2048.0f));
      const float t = __fmul_rn(static_cast<float>(s), b);
      acc = k == 3 ? t : __fadd_rn(t, acc);
    }
    v = __float2int_rn(acc);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);
    out[static_cast<size_t>(y) * out_w + x] = static_cast<uint8_t>(v);
  }
  if (minmax != nullptr) {
    int lo = x < out_w ? v : 255, hi = x < out_w ? v : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
  }
}

__global__ void minmax_init_kernel(int32_t* minmax) { minmax[0] = 255; minmax[1] = 0; }

// out[c, y, x] = float((double)(u8 - min) / max((double)(max - min), 1e-8)), c = 0..2
__global__ void __launch_bounds__(256) minmax_normalize3_kernel(const uint8_t* __restrict__ in, size_t n,
                                                                const int32_t* __restrict__ minmax,
                                                                float* __restrict__ out) {
  const int lo = minmax[0], hi = minmax[1];
  const double den = fmax(static_cast<double>(hi - lo), 1e-8);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = __double2float_rn(static_cast<double>(static_cast<int>(in[i]) - lo) / den);
    out[i] = v;
    out[n + i] = v;
    out[2 * n + i] = v;
  }
}

}  // namespace

int cvresize_cubic_coeffs_host(int in_size, int out_size, int32_t* idx4, int32_t* w4) {
  B200SAM_REQUIRE(in_size > 0 && out_size > 0 && idx4 != nullptr && w4 != nullptr, "cvresize_cubic_coeffs: bad arguments");
  const double scale = static_cast<double>(in_size) / static_cast<double>(out_size);
  for (int d = 0; d < out_size; ++d) {
    volatile double pos = (d + 0.5) * scale;
    const float f0 = static_cast<float>(pos - 0.5);
    const int s = static_cast<int>(std::floor(f0));
    // interpolateCubic, every operation rounded to fp32 (volatile: no contraction, no excess precision)
    volatile float x = f0 - static_cast<float>(s);
    const float A = -0.75f;
    volatile float x1 = x + 1.0f, u = 1.0f - x;
    volatile float t;
    volatile float c[4];
    t = A * x1; t = t - 5.0f * A; t = t * x1; t = t + 8.0f * A; t = t * x1; c[0] = t - 4.0f * A;
    t = (A + 2.0f) * x; t = t - (A + 3.0f); t = t * x; t = t * x; c[1] = t + 1.0f;
    t = (A + 2.0f) * u; t = t - (A + 3.0f); t = t * u; t = t * u; c[2] = t + 1.0f;
    t = 1.0f - c[0]; t = t - c[1]; c[3] = t - c[2];
    for (int k = 0; k < 4; ++k) {
      volatile float scaled = c[k] * 2048.0f;
      const long r = std::lrint(scaled);
      w4[4 * d + k] = static_cast<int32_t>(r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
      const int i = s - 1 + k;
      idx4[4 * d + k] = i < 0 ? 0 : (i > in_size - 1 ? in_size - 1 : i);
    }
  }
  return 0;
}

int medsam_preprocess(const uint8_t* gray, int H, int W, const int32_t* xi, const int32_t* xw, const int32_t* yi,
                      const int32_t* yw, int size, uint8_t* tmp_u8, int32_t* minmax, float* out3, cudaStream_t stream) {
  B200SAM_REQUIRE(gray && xi && xw && yi && yw && tmp_u8 && minmax && out3 && H > 0 && W > 0 && size > 0 && size <= 65535,
                  "medsam_preprocess: bad arguments (H=%d W=%d size=%d)", H, W, size);
  minmax_init_kernel<<<1, 1, 0, stream>>>(minmax);
  dim3 grid((size + 255) / 256, size);
  cvresize_cubic_kernel<<<grid, 256, 0, stream>>>(gray, H, W, xi, xw, yi, yw, size, size, tmp_u8, minmax);
  const size_t n = static_cast<size_t>(size) * size;
  minmax_normalize3_kernel<<<148 * 8, 256, 0, stream>>>(tmp_u8, n, minmax, out3);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
