// U-Net inference: the producer of the masks that the SAM refinement consumes
// (reference: custom_arcitecture/classic_u_net.py:9-119, called from scripts/save_refined_segmentations.py:62-69).
//
//   DoubleConv = (conv3x3 pad 1, no bias -> InstanceNorm2d(affine, eps 1e-5) -> LeakyReLU(0.01)) x 2
//   Down = MaxPool2d(2) + DoubleConv;  Up = ConvTranspose2d(k2, s2) + cat([skip, up]) + DoubleConv;  OutConv = conv1x1
//
// Activations are NHWC fp32 ([B*H*W, C] row-major, optionally a channel slice of a wider concat buffer).  Every 3x3
// convolution with Cin % 64 == 0 is an IMPLICIT GEMM on the tcgen05 kernel: the input is written once as a [hi | lo] bf16
// split operand on a pixel grid with a one-pixel zero border (unet_pad_split_kernel, 4*Cin bytes per pixel), the GEMM runs
// over that padded grid and its TMA producer fetches the A tile of tap (ky, kx) as the same 2-D box shifted by
// (ky-1)*(W+2) + (kx-1) rows (gemm_tcgen05.cu, conv mode) - no im2col matrix (36*Cin bytes per pixel) is ever written.
// 3-way bf16 split operands (x = hi + lo; [hi|lo|hi] x [hi|hi|lo], fp32 accumulate, ~2^-17 relative error - the same
// fp32-grade path as the mask decoder); the weights are split once at create time.  The first convolution (one input
// channel, 9 MACs per output) is a direct fp32 kernel.  InstanceNorm: per-(image, channel) sum / sum of squares in fp64
// over the interior rows of the padded grid, (mean, rstd) finalised once per channel, normalise + LeakyReLU written into
// the consumer's (un-padded) buffer.  ConvTranspose2d(k2, s2) = GEMM + depth-to-space; OutConv = 1x1 GEMM + sigmoid.
#include "common.cuh"
#include "kernels.h"
#include "decoder_ops.h"
#include <algorithm>
#include <string>
#include <vector>

namespace b200sam {

namespace {

inline size_t al256(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

B200SAM_DEVINL void split8(const float (&f)[8], uint4& H, uint4& Lo) {
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(f[2 * i]), h1 = __float2bfloat16_rn(f[2 * i + 1]);
    __nv_bfloat162 hp, lp;
    hp.x = h0; hp.y = h1;
    lp = __floats2bfloat162_rn(f[2 * i] - __bfloat162float(h0), f[2 * i + 1] - __bfloat162float(h1));
    hi[i] = *reinterpret_cast<uint32_t*>(&hp);
    lo[i] = *reinterpret_cast<uint32_t*>(&lp);
  }
  H = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  Lo = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// in: fp32 [B, H, W, *] (row pitch ld_in floats, C channels used) -> out bf16 [B*H*W, 2*Kp] = [hi | lo], column of
// tap (ky, kx), channel c = (ky*3 + kx)*C + c, zero for out-of-image taps and for the padding columns 9C..Kp-1.
__global__ void __launch_bounds__(256) unet_im2col3x3_split_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                                                   int ld_in, int Kp, __nv_bfloat16* __restrict__ out) {
  const int vec_per_row = Kp / 8;
  const size_t total = static_cast<size_t>(B) * H * W * vec_per_row;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec_per_row);
    const size_t m = idx / vec_per_row;
    const int x = static_cast<int>(m % W);
    const int y = static_cast<int>((m / W) % H);
    const size_t bimg = m / (static_cast<size_t>(W) * H);
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int k0 = v * 8;
    if (C % 8 == 0) {
      if (k0 < 9 * C) {
        const int tap = k0 / C, c = k0 - tap * C;
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float4* src = reinterpret_cast<const float4*>(in + ((bimg * H + yy) * W + xx) * ld_in + c);
          const float4 a = src[0], b = src[1];
          f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        }
      }
    } else {  // few input channels (the first convolution: C = 1): element-wise
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = k0 + i;
        if (k < 9 * C) {
          const int tap = k / C, c = k - tap * C;
          const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) f[i] = in[((bimg * H + yy) * W + xx) * ld_in + c];
        }
      }
    }
    uint4 Hh, Lo;
    split8(f, Hh, Lo);
    __nv_bfloat16* o = out + m * (2 * static_cast<size_t>(Kp)) + k0;
    *reinterpret_cast<uint4*>(o) = Hh;
    *reinterpret_cast<uint4*>(o + Kp) = Lo;
  }
}

// First convolution (one input channel): direct fp32 FMAs, one thread per (pixel, 4 output channels); wt fp32 [Cout, Kp]
// tap-major.  The grid stride is a multiple of Cout / 4, so a thread keeps its 4 x 9 weights in registers for all its
// pixels.  9 MACs per output: HBM-bound on the output write.
__global__ void __launch_bounds__(256) unet_conv3x3_direct_kernel(const float* __restrict__ in, int B, int H, int W,
                                                                  int ld_in, const float* __restrict__ wt, int Kp, int Cout,
                                                                  float* __restrict__ out) {
  const int q = Cout / 4;
  const size_t total = static_cast<size_t>(B) * H * W * q;
  const size_t first = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const int co = static_cast<int>(first % q) * 4;  // invariant: the stride below is a multiple of q
  float wr[4][9];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[i][t] = __ldg(wt + static_cast<size_t>(co + i) * Kp + t);
  for (size_t idx = first; idx < total; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t m = idx / q;
    const int x = static_cast<int>(m % W);
    const int y = static_cast<int>((m / W) % H);
    const size_t bimg = m / (static_cast<size_t>(W) * H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      const float v = ok ? __ldg(in + ((bimg * H + (ok ? yy : y)) * W + (ok ? xx : x)) * ld_in) : 0.0f;
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(v, wr[i][t], acc[i]);
    }
    *reinterpret_cast<float4*>(out + m * Cout + co) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}

// Operand of the implicit 3x3 convolution (gemm_tcgen05.cu, conv mode): in fp32 [B, H, W, *] (row pitch ld_in floats,
// C channels, C % 8 == 0) -> out bf16 [B * (H+2) * (W+2), 2*C] = [hi | lo] on a pixel grid with a one-pixel ZERO border.
// 4*C bytes written per pixel instead of the 36*C of the explicit im2col matrix.
__global__ void __launch_bounds__(256) unet_pad_split_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                                             int ld_in, __nv_bfloat16* __restrict__ out) {
  const int vec_per_row = C / 8;
  const int Hp = H + 2, Wp = W + 2;
  const size_t total = static_cast<size_t>(B) * Hp * Wp * vec_per_row;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec_per_row);
    const size_t pm = idx / vec_per_row;
    const int xp = static_cast<int>(pm % Wp);
    const int yp = static_cast<int>((pm / Wp) % Hp);
    const size_t bimg = pm / (static_cast<size_t>(Wp) * Hp);
    uint4 Hh = make_uint4(0, 0, 0, 0), Lo = make_uint4(0, 0, 0, 0);
    if (xp >= 1 && xp <= W && yp >= 1 && yp <= H) {
      const float4* src = reinterpret_cast<const float4*>(in + ((bimg * H + (yp - 1)) * W + (xp - 1)) * ld_in + 8 * v);
      const float4 a = src[0], b = src[1];
      const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      split8(f, Hh, Lo);
    }
    __nv_bfloat16* o = out + pm * (2 * static_cast<size_t>(C)) + 8 * v;
    *reinterpret_cast<uint4*>(o) = Hh;
    *reinterpret_cast<uint4*>(o + C) = Lo;
  }
}

// per-(image, channel) sum and sum of squares over the H*W rows of x [B*HW, C] (fp64 accumulators, pre-zeroed).
// W > 0: x lives on the zero-bordered (H+2) x (W+2) grid of the implicit convolution; only interior rows count.
// A block covers rows [r0, r1) of one image and 4 * min(C / 4, 256) channels: thread = (float4 channel group, row
// sub-sequence), so all 256 threads stream 16-byte loads whatever C is (C = 64: 16 groups x 16 row sub-sequences);
// short fp32 runs, fp64 across runs, shared-memory tree over the row sub-sequences, one fp64 atomic pair per channel.
__global__ void __launch_bounds__(256) unet_in_stats_kernel(const float* __restrict__ x, int HW, int C, int rows_per_block,
                                                            double* __restrict__ stats /*[B, C, 2]*/, int W) {
  __shared__ double sh[256][8];
  const int b = blockIdx.z;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(HW, r0 + rows_per_block);
  const int c4n = C >> 2;                  // float4 groups per row
  const int gpb = min(c4n, 256);           // groups handled by this block
  const int nsub = 256 / gpb;              // row sub-sequences
  const int g = threadIdx.x % gpb, sub = threadIdx.x / gpb;
  const int c4 = blockIdx.x * gpb + g;
  double ds[4] = {0.0, 0.0, 0.0, 0.0}, dq[4] = {0.0, 0.0, 0.0, 0.0};
  if (c4 < c4n && sub < nsub) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
    int n = 0;
    const int Wp = W + 2;
    const size_t img_rows = W > 0 ? static_cast<size_t>(HW / W + 2) * Wp : static_cast<size_t>(HW);
    const float* px = x + (static_cast<size_t>(b) * img_rows) * C + 4 * c4;
    int yy = 0, xx = 0;
    if (W > 0) { yy = (r0 + sub) / W; xx = (r0 + sub) - yy * W; }
    for (int r = r0 + sub; r < r1; r += nsub) {
      size_t row = static_cast<size_t>(r);
      if (W > 0) {
        row = static_cast<size_t>(yy + 1) * Wp + (xx + 1);
        xx += nsub;
        while (xx >= W) { xx -= W; ++yy; }
      }
      const float4 v = __ldg(reinterpret_cast<const float4*>(px + row * C));
      s4[0] += v.x; s4[1] += v.y; s4[2] += v.z; s4[3] += v.w;
      q4[0] = fmaf(v.x, v.x, q4[0]); q4[1] = fmaf(v.y, v.y, q4[1]);
      q4[2] = fmaf(v.z, v.z, q4[2]); q4[3] = fmaf(v.w, v.w, q4[3]);
      if (++n == 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { ds[i] += s4[i]; dq[i] += q4[i]; s4[i] = q4[i] = 0.0f; }
        n = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { ds[i] += s4[i]; dq[i] += q4[i]; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) { sh[threadIdx.x][i] = ds[i]; sh[threadIdx.x][4 + i] = dq[i]; }
  __syncthreads();
  // threads 0 .. gpb*8-1: one (group, component) each, summing over the row sub-sequences
  for (int t = threadIdx.x; t < gpb * 8; t += 256) {
    const int gg = t >> 3, comp = t & 7;
    double acc = 0.0;
    for (int k = 0; k < nsub; ++k) acc += sh[k * gpb + gg][comp];
    const int cc = 4 * (blockIdx.x * gpb + gg) + (comp & 3);
    if (cc < C) atomicAdd(&stats[(static_cast<size_t>(b) * C + cc) * 2 + (comp >> 2)], acc);
  }
}

// (sum, sum of squares) -> (mean, 1 / sqrt(var + eps)) in fp64, rounded to fp32 once per (image, channel);
// biased variance, eps 1e-5 (nn.InstanceNorm2d defaults)
__global__ void __launch_bounds__(256) unet_in_finalize_kernel(const double* __restrict__ stats, int HW, int n,
                                                               float2* __restrict__ mr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = stats[2 * i] / HW;
  double var = stats[2 * i + 1] / HW - mean * mean;
  var = var < 0.0 ? 0.0 : var;
  mr[i] = make_float2(static_cast<float>(mean), static_cast<float>(1.0 / sqrt(var + 1e-5)));
}

// y = LeakyReLU_0.01((x - mean) * rstd * gamma + beta)
__global__ void __launch_bounds__(256) unet_in_apply_kernel(const float* __restrict__ x, const float2* __restrict__ mr,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int HW, int C, size_t total4, float* __restrict__ y, int ld_y,
                                                            int W /* > 0: x on the zero-bordered grid */) {
  const int c4n = C / 4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % c4n) * 4;
    const size_t m = idx / c4n;
    const size_t b = m / HW;
    size_t xrow = m;
    if (W > 0) {
      const int rem = static_cast<int>(m - b * HW);
      const int yy = rem / W, xx = rem - yy * W;
      xrow = (b * (HW / W + 2) + yy + 1) * (W + 2) + xx + 1;
    }
    const float4 v = *reinterpret_cast<const float4*>(x + xrow * C + c);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float in[4] = {v.x, v.y, v.z, v.w};
    const float gm[4] = {g4.x, g4.y, g4.z, g4.w}, bt[4] = {b4.x, b4.y, b4.z, b4.w};
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 s = __ldg(mr + b * C + c + i);
      const float t = (in[i] - s.x) * s.y * gm[i] + bt[i];
      o[i] = t >= 0.0f ? t : 0.01f * t;
    }
    *reinterpret_cast<float4*>(y + m * ld_y + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// MaxPool2d(2) on NHWC fp32 (input row pitch ld_in): [B, H, W, C] -> [B, H/2, W/2, C]
__global__ void __launch_bounds__(256) unet_maxpool_kernel(const float* __restrict__ in, int H, int W, int C, int ld_in,
                                                           size_t total4, float* __restrict__ out) {
  const int c4n = C / 4, Wo = W / 2, Ho = H / 2;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % c4n) * 4;
    size_t m = idx / c4n;
    const int xo = static_cast<int>(m % Wo);
    m /= Wo;
    const int yo = static_cast<int>(m % Ho);
    const size_t b = m / Ho;
    const float* p = in + ((b * H + 2 * yo) * W + 2 * xo) * ld_in + c;
    const float4 a = *reinterpret_cast<const float4*>(p), bq = *reinterpret_cast<const float4*>(p + ld_in);
    const float4 cq = *reinterpret_cast<const float4*>(p + static_cast<size_t>(W) * ld_in);
    const float4 d = *reinterpret_cast<const float4*>(p + static_cast<size_t>(W) * ld_in + ld_in);
    float4 o;
    o.x = fmaxf(fmaxf(a.x, bq.x), fmaxf(cq.x, d.x));
    o.y = fmaxf(fmaxf(a.y, bq.y), fmaxf(cq.y, d.y));
    o.z = fmaxf(fmaxf(a.z, bq.z), fmaxf(cq.z, d.z));
    o.w = fmaxf(fmaxf(a.w, bq.w), fmaxf(cq.w, d.w));
    *reinterpret_cast<float4*>(out + ((b * Ho + yo) * Wo + xo) * C + c) = o;
  }
}

// ConvTranspose2d(k2, s2) output of the GEMM, [B*H*W, 4*Co] with column (dy*2+dx)*Co + co, scattered to NHWC
// [B, 2H, 2W, *] at channel offset c_off of a buffer with row pitch ld_out (the second half of the skip concat)
__global__ void __launch_bounds__(256) unet_depth_to_space_kernel(const float* __restrict__ in, int H, int W, int Co,
                                                                  size_t total4, float* __restrict__ out, int ld_out,
                                                                  int c_off) {
  const int c4n = Co / 4;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total4;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % c4n) * 4;
    size_t r = idx / c4n;
    const int q = static_cast<int>(r & 3);
    const size_t m = r >> 2;
    const int x = static_cast<int>(m % W);
    const int y = static_cast<int>((m / W) % H);
    const size_t b = m / (static_cast<size_t>(W) * H);
    const float4 v = *reinterpret_cast<const float4*>(in + m * (4 * static_cast<size_t>(Co)) + q * Co + c);
    *reinterpret_cast<float4*>(out + ((b * 2 * H + 2 * y + (q >> 1)) * 2 * W + 2 * x + (q & 1)) * ld_out + c_off + c) = v;
  }
}

// OutConv logits [B*HW, ldl] (first n_classes columns valid) -> NCHW logits and / or sigmoid probabilities
__global__ void __launch_bounds__(256) unet_head_kernel(const float* __restrict__ logits, int HW, int n_classes, int ldl,
                                                        size_t total, float* __restrict__ logits_out,
                                                        float* __restrict__ probs_out) {
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % HW);
    size_t r = idx / HW;
    const int c = static_cast<int>(r % n_classes);
    const size_t b = r / n_classes;
    const float v = logits[(b * HW + p) * ldl + c];
    if (logits_out != nullptr) logits_out[idx] = v;
    if (probs_out != nullptr) probs_out[idx] = 1.0f / (1.0f + expf(-v));  // torch.sigmoid
  }
}

unsigned grid_for(size_t total) {
  size_t g = (total + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  return static_cast<unsigned>(g < 1 ? 1 : g);
}

// ---------------------------------------------------------------- weight table
struct ConvSpec { int cin, cout; };

struct Arch {
  // DoubleConv list in execution order: inc, down1..4, up1..4 (each: conv a: cin->mid, conv b: mid->cout)
  ConvSpec dc[9][2];
  int up_cin[4];  // ConvTranspose2d cin (cout = cin / 2)
};

Arch make_arch(int n_channels, int n_last) {
  Arch a;
  const int ch[5] = {64, 128, 256, 512, 1024};
  a.dc[0][0] = {n_channels, 64}; a.dc[0][1] = {64, 64};
  for (int i = 1; i < 5; ++i) { a.dc[i][0] = {ch[i - 1], ch[i]}; a.dc[i][1] = {ch[i], ch[i]}; }
  for (int i = 0; i < 4; ++i) {
    const int cin = ch[4 - i], cout = i < 3 ? ch[3 - i] : n_last;
    a.up_cin[i] = cin;
    a.dc[5 + i][0] = {cin, cout};  // cat([skip cin/2, up cin/2]) -> cout
    a.dc[5 + i][1] = {cout, cout};
  }
  return a;
}

const char* kDcPrefix[9] = {"inc.double_conv", "down1.maxpool_conv.1.double_conv", "down2.maxpool_conv.1.double_conv",
                            "down3.maxpool_conv.1.double_conv", "down4.maxpool_conv.1.double_conv",
                            "up1.conv.double_conv", "up2.conv.double_conv", "up3.conv.double_conv", "up4.conv.double_conv"};

std::vector<std::string> build_names() {
  std::vector<std::string> n;
  for (int i = 0; i < 9; ++i) {
    const std::string p = kDcPrefix[i];
    n.push_back(p + ".0.weight|conv3x3_tap");  // fp32 [Cout, Kp]: column (ky*3+kx)*Cin + c, zero padded to Kp
    n.push_back(p + ".1.weight");
    n.push_back(p + ".1.bias");
    n.push_back(p + ".3.weight|conv3x3_tap");
    n.push_back(p + ".4.weight");
    n.push_back(p + ".4.bias");
  }
  for (int i = 1; i <= 4; ++i) {
    n.push_back("up" + std::to_string(i) + ".up.weight|convT");  // fp32 [(dy*2+dx)*Cout + co, Cin]
    n.push_back("up" + std::to_string(i) + ".up.bias|repeat4");
  }
  n.push_back("outc.conv.weight|pad_rows8");  // fp32 [round_up(n_classes, 8), n_last]
  n.push_back("outc.conv.bias|pad8");
  return n;
}
const std::vector<std::string>& names() {
  static std::vector<std::string> n = build_names();
  return n;
}
enum : int { W_DC = 0, W_UP = 54, W_OUT = 62, W_COUNT = 64 };

}  // namespace

struct UNetCtx {
  int n_channels, n_classes, n_last, ncp;  // ncp = n_classes rounded up to 8
  Arch arch;
  std::vector<const float*> w;  // borrowed fp32 pointers in unet_weight_name() order
  __nv_bfloat16* wsplit;        // owned: [hi|hi|lo] splits of all GEMM weights
  const __nv_bfloat16* ws_dc[9][2];
  const __nv_bfloat16* ws_up[4];
  const __nv_bfloat16* ws_out;
};

int unet_conv_kp(int cin) { return round_up(9 * cin, 64); }

int unet_weight_count() { return W_COUNT; }
const char* unet_weight_name(int i) {
  if (i < 0 || i >= static_cast<int>(names().size())) return nullptr;
  return names()[i].c_str();
}

int unet_create(int n_channels, int n_classes, int n_last, const void* const* weights, int n, UNetCtx** out,
                cudaStream_t stream) {
  B200SAM_REQUIRE(n == W_COUNT && static_cast<int>(names().size()) == W_COUNT, "unet_create: expected %d weights, got %d",
                  W_COUNT, n);
  B200SAM_REQUIRE(n_channels >= 1 && n_classes >= 1 && n_last >= 8 && n_last % 8 == 0,
                  "unet_create: bad configuration (channels %d, classes %d, last %d)", n_channels, n_classes, n_last);
  for (int i = 0; i < n; ++i) B200SAM_REQUIRE(weights[i] != nullptr, "unet_create: weight %d (%s) is null", i, names()[i].c_str());
  UNetCtx* u = new UNetCtx();
  u->n_channels = n_channels; u->n_classes = n_classes; u->n_last = n_last; u->ncp = round_up(n_classes, 8);
  u->arch = make_arch(n_channels, n_last);
  u->w.assign(reinterpret_cast<const float* const*>(weights), reinterpret_cast<const float* const*>(weights) + n);
  struct Item { int idx, N, K; const __nv_bfloat16** dst; };
  std::vector<Item> items;
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 2; ++j)
      items.push_back({W_DC + i * 6 + j * 3, u->arch.dc[i][j].cout, unet_conv_kp(u->arch.dc[i][j].cin), &u->ws_dc[i][j]});
  for (int i = 0; i < 4; ++i) items.push_back({W_UP + 2 * i, 4 * (u->arch.up_cin[i] / 2), u->arch.up_cin[i], &u->ws_up[i]});
  items.push_back({W_OUT, u->ncp, n_last, &u->ws_out});
  size_t total = 0;
  for (const Item& it : items) total += static_cast<size_t>(it.N) * 3 * it.K;
  u->wsplit = nullptr;
  if (cudaMalloc(&u->wsplit, total * sizeof(__nv_bfloat16)) != cudaSuccess) {
    set_last_error("unet_create: cudaMalloc of the split weights failed");
    delete u;
    return 1;
  }
  size_t off = 0;
  for (const Item& it : items) {
    *it.dst = u->wsplit + off;
    if (int rc = split3_bf16(u->w[it.idx], nullptr, 0, u->wsplit + off, it.N, it.K, 1, stream)) {
      cudaFree(u->wsplit);
      delete u;
      return rc;
    }
    off += static_cast<size_t>(it.N) * 3 * it.K;
  }
  *out = u;
  return 0;
}

void unet_destroy(UNetCtx* u) {
  if (u == nullptr) return;
  if (u->wsplit) cudaFree(u->wsplit);
  delete u;
}

namespace {

struct UWork {
  float* cat[4];     // concat buffers of the 4 skip levels: [B*HW_l, 2*C_l] (skip | upsampled)
  float* x5;         // bottleneck [B*HW/256, 1024]
  float *t0, *t1;    // conv outputs / pooled inputs (largest: [B*HW, 64] and [B*HW/4, 128] ...)
  __nv_bfloat16* col;  // im2col / split operand
  double* stats;     // [B, 1024, 2]
  float2* mr;        // [B, 1024] (mean, rstd)
  size_t total;
};

UWork carve_unet(uint8_t* base, const UNetCtx* u, int B, int H, int W) {
  UWork w;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base + off;
    off += al256(bytes);
    return p;
  };
  const size_t HW = static_cast<size_t>(H) * W;
  const int ch[4] = {64, 128, 256, 512};
  for (int l = 0; l < 4; ++l) w.cat[l] = reinterpret_cast<float*>(take(B * (HW >> (2 * l)) * 2 * ch[l] * sizeof(float)));
  w.x5 = reinterpret_cast<float*>(take(B * (HW >> 8) * 1024 * sizeof(float)));
  // conv outputs: max over layers of M * Cout; ConvT outputs M_in * 4 * Cout have the same size as the next level's map
  size_t tmax = 0, cmax = 0;
  for (int i = 0; i < 9; ++i) {
    const int lvl = i < 5 ? i : 8 - i;  // resolution level of the DoubleConv
    const size_t M = B * (HW >> (2 * lvl));
    const size_t Mp = static_cast<size_t>(B) * ((H >> lvl) + 2) * ((W >> lvl) + 2);  // implicit-conv outputs: padded grid
    for (int j = 0; j < 2; ++j) {
      tmax = std::max(tmax, Mp * u->arch.dc[i][j].cout);
      cmax = std::max(cmax, Mp * 2 * static_cast<size_t>(u->arch.dc[i][j].cin));
      cmax = std::max(cmax, M * 2 * static_cast<size_t>(unet_conv_kp(u->arch.dc[i][j].cin)));
    }
  }
  for (int i = 0; i < 4; ++i) {
    const size_t Min = B * (HW >> (2 * (4 - i)));
    tmax = std::max(tmax, Min * 4 * (u->arch.up_cin[i] / 2));
    cmax = std::max(cmax, Min * 2 * static_cast<size_t>(u->arch.up_cin[i]));
  }
  tmax = std::max(tmax, B * HW * static_cast<size_t>(u->ncp));
  cmax = std::max(cmax, B * HW * 2 * static_cast<size_t>(u->n_last));
  w.t0 = reinterpret_cast<float*>(take(tmax * sizeof(float)));
  w.t1 = reinterpret_cast<float*>(take(tmax * sizeof(float)));
  w.col = reinterpret_cast<__nv_bfloat16*>(take(cmax * sizeof(__nv_bfloat16)));
  w.stats = reinterpret_cast<double*>(take(static_cast<size_t>(B) * 1024 * 2 * sizeof(double)));
  w.mr = reinterpret_cast<float2*>(take(static_cast<size_t>(B) * 1024 * sizeof(float2)));
  w.total = off;
  return w;
}

#define TRY(x) do { if (int _rc = (x)) return _rc; } while (0)

// out[M, N] fp32 = split(A)[M, 3K] . W'[N, 3K]^T (+ bias)
int split_gemm(const __nv_bfloat16* As, const __nv_bfloat16* Ws, const float* bias, float* out, size_t M, int N, int K,
               cudaStream_t s) {
  B200SAM_REQUIRE(M < (1ull << 31), "unet: too many rows (%zu)", M);
  GemmArgs g;
  g.A = As; g.B = Ws; g.out = out; g.bias = bias; g.residual = nullptr;
  g.M = static_cast<int>(M); g.N = N; g.K = 3 * K; g.lda = 2 * K; g.ldb = 3 * K; g.ldo = N; g.ldr = 0; g.res_row_mod = 0;
  g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 2 * K;
  return gemm_bf16_tn(g, s);
}

// conv3x3 (no bias) + InstanceNorm + LeakyReLU: in [B,H,W,cin] (pitch ld_in) -> out [B,H,W,cout] (pitch ld_out)
int conv_in_lrelu(const UNetCtx* u, const UWork& w, const float* in, int ld_in, int B, int H, int W, int cin, int cout,
                  const __nv_bfloat16* ws, const float* wf32, const float* gamma, const float* beta, float* tmp, float* out,
                  int ld_out, cudaStream_t s) {
  (void)u;
  const size_t HW = static_cast<size_t>(H) * W, M = B * HW;
  const int Kp = unet_conv_kp(cin);
  const bool implicit = cin % 64 == 0;  // every convolution but the first (cin = n_channels = 1)
  if (implicit) {
    // zero-bordered [hi | lo] operand + implicit-convolution GEMM over the padded grid (no im2col matrix)
    const size_t Mp = static_cast<size_t>(B) * (H + 2) * (W + 2);
    B200SAM_REQUIRE(Mp < (1ull << 31) && Kp == 9 * cin, "unet: bad implicit convolution shape");
    unet_pad_split_kernel<<<grid_for(Mp * (cin / 8)), 256, 0, s>>>(in, B, H, W, cin, ld_in, w.col);
    GemmArgs g;
    g.A = w.col; g.B = ws; g.out = tmp; g.bias = nullptr; g.residual = nullptr;
    g.M = static_cast<int>(Mp); g.N = cout; g.K = 27 * cin; g.lda = 2 * cin; g.ldb = 27 * cin; g.ldo = cout; g.ldr = 0;
    g.res_row_mod = 0; g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 0; g.conv_cin = cin; g.conv_wp = W + 2;
    TRY(gemm_bf16_tn(g, s));
  } else if (cin == 1 && wf32 != nullptr && 256 % (cout / 4) == 0) {
    unet_conv3x3_direct_kernel<<<grid_for(M * (cout / 4)), 256, 0, s>>>(in, B, H, W, ld_in, wf32, Kp, cout, tmp);
  } else {
    unet_im2col3x3_split_kernel<<<grid_for(M * (Kp / 8)), 256, 0, s>>>(in, B, H, W, cin, ld_in, Kp, w.col);
    TRY(split_gemm(w.col, ws, nullptr, tmp, M, cout, Kp, s));
  }
  const int Wg = implicit ? W : 0;
  B200SAM_CHECK_CUDA(cudaMemsetAsync(w.stats, 0, static_cast<size_t>(B) * cout * 2 * sizeof(double), s));
  const int rpb = static_cast<int>(std::min<size_t>(512, std::max<size_t>(64, HW / 32)));  // >= ~250 blocks per launch
  dim3 gs((cout / 4 + 255) / 256, static_cast<unsigned>((HW + rpb - 1) / rpb), B);
  unet_in_stats_kernel<<<gs, 256, 0, s>>>(tmp, static_cast<int>(HW), cout, rpb, w.stats, Wg);
  const size_t total4 = M * (cout / 4);
  unet_in_finalize_kernel<<<(B * cout + 255) / 256, 256, 0, s>>>(w.stats, static_cast<int>(HW), B * cout, w.mr);
  unet_in_apply_kernel<<<grid_for(total4), 256, 0, s>>>(tmp, w.mr, gamma, beta, static_cast<int>(HW), cout, total4, out,
                                                       ld_out, Wg);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

size_t unet_workspace_bytes(const UNetCtx* u, int B, int H, int W) {
  if (u == nullptr || B <= 0 || H <= 0 || W <= 0) return 0;
  return carve_unet(nullptr, u, B, H, W).total + 256;
}

int unet_forward(const UNetCtx* u, const float* image, int B, int H, int W, float* logits_out, float* probs_out,
                 void* workspace, size_t workspace_bytes, cudaStream_t s) {
  B200SAM_REQUIRE(u != nullptr && image != nullptr && B > 0, "unet_forward: null argument");
  B200SAM_REQUIRE(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "unet_forward: H and W must be multiples of 16 (got %d x %d)", H, W);
  B200SAM_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                  "unet_forward: workspace must be non-null and 256-byte aligned");
  UWork w = carve_unet(static_cast<uint8_t*>(workspace), u, B, H, W);
  B200SAM_REQUIRE(w.total <= workspace_bytes, "unet_forward: workspace too small (%zu < %zu)", workspace_bytes, w.total);
  const float* const* Wt = u->w.data();
  const Arch& a = u->arch;
  const int ch[5] = {64, 128, 256, 512, 1024};
  auto dc = [&](int i, const float* in, int ld_in, int h, int wd, float* out, int ld_out) -> int {
    const float* const* P = Wt + W_DC + i * 6;
    // conv a -> t1 (pitch = mid channels), conv b -> out
    TRY(conv_in_lrelu(u, w, in, ld_in, B, h, wd, a.dc[i][0].cin, a.dc[i][0].cout, u->ws_dc[i][0], P[0], P[1], P[2], w.t0, w.t1,
                      a.dc[i][0].cout, s));
    TRY(conv_in_lrelu(u, w, w.t1, a.dc[i][0].cout, B, h, wd, a.dc[i][1].cin, a.dc[i][1].cout, u->ws_dc[i][1], P[3], P[4], P[5],
                      w.t0, out, ld_out, s));
    return 0;
  };
  // ---- encoder: x1..x4 land in the first half of their level's concat buffer (torch.cat([x2, x1], dim=1): skip first)
  // image [B, n_channels, H, W] with n_channels == 1 is already NHWC
  B200SAM_REQUIRE(u->n_channels == 1, "unet_forward: only single-channel input is supported (got %d)", u->n_channels);
  TRY(dc(0, image, 1, H, W, w.cat[0], 2 * ch[0]));
  for (int l = 1; l <= 4; ++l) {
    const int h = H >> l, wd = W >> l;
    const size_t total4 = static_cast<size_t>(B) * h * wd * (ch[l - 1] / 4);
    unet_maxpool_kernel<<<grid_for(total4), 256, 0, s>>>(w.cat[l - 1], H >> (l - 1), W >> (l - 1), ch[l - 1], 2 * ch[l - 1],
                                                        total4, w.t1);
    // the pooled map is consumed from t1 by the im2col of conv a, whose output goes to t0 and then (normalised) back to
    // t1: the im2col has finished reading t1 by then (stream order)
    float* out = l < 4 ? w.cat[l] : w.x5;
    TRY(dc(l, w.t1, ch[l - 1], h, wd, out, l < 4 ? 2 * ch[l] : ch[4]));
  }
  // ---- decoder
  const float* cur = w.x5;
  int cur_ld = ch[4];
  for (int i = 0; i < 4; ++i) {
    const int lvl = 3 - i;                        // level of the skip connection / output resolution
    const int h = H >> (lvl + 1), wd = W >> (lvl + 1);  // input resolution of the transposed convolution
    const int cin = a.up_cin[i], co = cin / 2;
    const size_t Min = static_cast<size_t>(B) * h * wd;
    B200SAM_REQUIRE(cur_ld == cin, "unet_forward: internal pitch mismatch");
    TRY(split3_bf16(cur, nullptr, 0, w.col, Min, cin, 0, s));
    TRY(split_gemm(w.col, u->ws_up[i], Wt[W_UP + 2 * i + 1], w.t0, Min, 4 * co, cin, s));
    const size_t total4 = Min * 4 * (co / 4);
    unet_depth_to_space_kernel<<<grid_for(total4), 256, 0, s>>>(w.t0, h, wd, co, total4, w.cat[lvl], 2 * ch[lvl], ch[lvl]);
    // DoubleConv on the concat buffer; the level's output stays in t1 (conv b's im2col has consumed conv a's output
    // from t1 before the normalised result of conv b is written there: stream order)
    const int cout = a.dc[5 + i][1].cout;
    TRY(dc(5 + i, w.cat[lvl], 2 * ch[lvl], 2 * h, 2 * wd, w.t1, cout));
    cur = w.t1;
    cur_ld = cout;
  }
  // ---- OutConv (1x1) + sigmoid
  {
    const size_t HW = static_cast<size_t>(H) * W, M = B * HW;
    TRY(split3_bf16(cur, nullptr, 0, w.col, M, u->n_last, 0, s));
    TRY(split_gemm(w.col, u->ws_out, Wt[W_OUT + 1], w.t0, M, u->ncp, u->n_last, s));
    const size_t total = static_cast<size_t>(B) * u->n_classes * HW;
    unet_head_kernel<<<grid_for(total), 256, 0, s>>>(w.t0, static_cast<int>(HW), u->n_classes, u->ncp, total, logits_out,
                                                    probs_out);
  }
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
