// Memory-bound helpers around the encoder GEMMs: pixel normalisation + patch im2col, LayerNorm,
// the neck's 3x3 im2col and the final LayerNorm2d -> NCHW store.  All are one-pass, vectorised.
#include "common.cuh"
#include "kernels.h"

namespace b200sam {

namespace {

// ------------------------------------------------------------------ preprocess + patchify
// Reference: Sam.preprocess (segment_anything/modeling/sam.py:164-174) = (x - mean) / std, zero pad to
// 1024^2 AFTER normalisation; PatchEmbed conv k16/s16 (image_encoder.py:387-395) becomes a GEMM whose
// A operand row = token (py*64+px) and column = c*256 + ky*16 + kx.
template <bool IS_U8, bool F16>
__global__ void __launch_bounds__(256) preprocess_patchify_kernel(const void* __restrict__ img, int B, int h, int w,
                                                                  float m0, float m1, float m2, float s0, float s1,
                                                                  float s2, __nv_bfloat16* __restrict__ out) {
  // one thread = 8 consecutive kx of one (token, c, ky)
  const size_t total = static_cast<size_t>(B) * 4096 * 96;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % 96);  // 16-byte group inside the 768-wide row
    const size_t tok = idx / 96;
    const int b = static_cast<int>(tok >> 12);
    const int t = static_cast<int>(tok & 4095);
    const int py = t >> 6, px = t & 63;
    const int c = g >> 5;
    const int ky = (g >> 1) & 15;
    const int kx0 = (g & 1) * 8;
    const int y = py * 16 + ky;
    const int x0 = px * 16 + kx0;
    const float mean = c == 0 ? m0 : (c == 1 ? m1 : m2);
    const float sd = c == 0 ? s0 : (c == 1 ? s1 : s2);
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.0f;
    if (y < h) {
      const size_t base = (static_cast<size_t>(b) * 3 + c) * h * static_cast<size_t>(w) + static_cast<size_t>(y) * w;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int x = x0 + i;
        if (x < w) {
          float p;
          if constexpr (IS_U8) p = static_cast<float>(reinterpret_cast<const uint8_t*>(img)[base + x]);
          else p = reinterpret_cast<const float*>(img)[base + x];
          v[i] = __fdiv_rn(p - mean, sd);
        }
      }
    }
    uint4 o;
    o.x = pack_op16x2<F16>(v[0], v[1]);
    o.y = pack_op16x2<F16>(v[2], v[3]);
    o.z = pack_op16x2<F16>(v[4], v[5]);
    o.w = pack_op16x2<F16>(v[6], v[7]);
    *reinterpret_cast<uint4*>(out + tok * 768 + g * 8) = o;
  }
}

// ------------------------------------------------------------------ LayerNorm over rows
// One warp per row, the row lives in registers (D <= 1280, D % 128 == 0): mean, then centred variance
// (two-pass, fp32) like torch's CPU kernel; eps inside the sqrt (image_encoder.py:168,180 via nn.LayerNorm).
template <int NV, int OUT_KIND>
__global__ void __launch_bounds__(256) layernorm_rows_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int M, int D,
                                                             void* __restrict__ y, int reverse) {
  const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w0 >= M) return;
  // reverse: the first blocks take the LAST rows - the ones the producing GEMM wrote most recently (L2 resident)
  const int warp = reverse ? M - 1 - w0 : w0;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(warp) * D);
  float4 v[NV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  s = warp_sum(s);
  const float mean = s / static_cast<float>(D);
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  q = warp_sum(q);
  const float rstd = 1.0f / sqrtf(q / static_cast<float>(D) + eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = g4[lane + 32 * i];
    const float4 b = b4[lane + 32 * i];
    float4 o;
    o.x = (v[i].x - mean) * rstd * g.x + b.x;
    o.y = (v[i].y - mean) * rstd * g.y + b.y;
    o.z = (v[i].z - mean) * rstd * g.z + b.z;
    o.w = (v[i].w - mean) * rstd * g.w + b.w;
    if constexpr (OUT_KIND != 0) {
      uint2 p;
      p.x = pack_op16x2<OUT_KIND == 2>(o.x, o.y);
      p.y = pack_op16x2<OUT_KIND == 2>(o.z, o.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + static_cast<size_t>(warp) * D)[lane + 32 * i] = p;
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + static_cast<size_t>(warp) * D)[lane + 32 * i] = o;
    }
  }
}

template <int NV>
int launch_ln(const float* x, const float* gamma, const float* beta, float eps, int M, int D, void* y, int out_kind,
              cudaStream_t stream, int reverse) {
  const int warps_per_block = 8;
  const int blocks = (M + warps_per_block - 1) / warps_per_block;
  if (out_kind == 2)
    layernorm_rows_kernel<NV, 2><<<blocks, 256, 0, stream>>>(x, gamma, beta, eps, M, D, y, reverse);
  else if (out_kind == 1)
    layernorm_rows_kernel<NV, 1><<<blocks, 256, 0, stream>>>(x, gamma, beta, eps, M, D, y, reverse);
  else
    layernorm_rows_kernel<NV, 0><<<blocks, 256, 0, stream>>>(x, gamma, beta, eps, M, D, y, reverse);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ neck 3x3 im2col (pad 1)
// in [B*4096, C] bf16 token-major (NHWC) -> out [B*4096, 9*C], column = (ky*3+kx)*C + c.
__global__ void __launch_bounds__(256) im2col3x3_kernel(const __nv_bfloat16* __restrict__ in, int B, int C,
                                                        __nv_bfloat16* __restrict__ out) {
  const int vec_per_tap = C / 8;
  const size_t total = static_cast<size_t>(B) * 4096 * 9 * vec_per_tap;
  for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vec_per_tap);
    size_t r = idx / vec_per_tap;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int b = static_cast<int>(r >> 12);
    const int t = static_cast<int>(r & 4095);
    const int y = (t >> 6) + tap / 3 - 1;
    const int x = (t & 63) + tap % 3 - 1;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (y >= 0 && y < 64 && x >= 0 && x < 64)
      val = *reinterpret_cast<const uint4*>(in + (static_cast<size_t>(b) * 4096 + y * 64 + x) * C + v * 8);
    *reinterpret_cast<uint4*>(out + r * (9 * static_cast<size_t>(C)) + tap * C + v * 8) = val;
  }
}

// ------------------------------------------------------------------ final LayerNorm2d + NCHW store
// LayerNorm2d (common.py:31-43) normalises over channels at each pixel == row LN in token-major layout.
// Block = 32 tokens x C(=256) channels, transposed through shared memory so the NCHW store is coalesced.
__global__ void __launch_bounds__(256) layernorm_to_nchw_kernel(const float* __restrict__ x,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, int C,
                                                                float* __restrict__ out) {
  extern __shared__ float tile[];  // [C][33]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t tok0 = static_cast<size_t>(blockIdx.x) * 32;  // 32 consecutive tokens, same image
  const int per_lane = C / 32;                               // 8 for C = 256
  for (int tt = warp; tt < 32; tt += 8) {
    const float* xr = x + (tok0 + tt) * C;
    float v[16];
    float s = 0.0f;
    for (int i = 0; i < per_lane; ++i) { v[i] = xr[lane + 32 * i]; s += v[i]; }
    s = warp_sum(s);
    const float mean = s / static_cast<float>(C);
    float q = 0.0f;
    for (int i = 0; i < per_lane; ++i) { const float d = v[i] - mean; q += d * d; }
    q = warp_sum(q);
    const float rstd = 1.0f / sqrtf(q / static_cast<float>(C) + eps);
    for (int i = 0; i < per_lane; ++i) {
      const int c = lane + 32 * i;
      tile[c * 33 + tt] = (v[i] - mean) * rstd * gamma[c] + beta[c];
    }
  }
  __syncthreads();
  const int b = static_cast<int>(tok0 >> 12);
  const int t0 = static_cast<int>(tok0 & 4095);
  for (int c = warp; c < C; c += 8)
    out[(static_cast<size_t>(b) * C + c) * 4096 + t0 + lane] = tile[c * 33 + lane];
}

template <bool F16>
__global__ void __launch_bounds__(256) f32_to_op16_kernel(const float* __restrict__ in, uint32_t* __restrict__ out,
                                                          size_t npairs) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < npairs;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(in)[i];
    out[i] = pack_op16x2<F16>(v.x, v.y);
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return static_cast<int>(g < cap ? (g == 0 ? 1 : g) : cap);
}

}  // namespace

int preprocess_patchify(const void* img, int is_u8, int B, int h, int w, const float* mean3, const float* std3,
                        __nv_bfloat16* out, int out_f16, cudaStream_t stream) {
  B200SAM_REQUIRE(B > 0 && h > 0 && w > 0 && h <= 1024 && w <= 1024, "preprocess: bad shape B=%d h=%d w=%d", B, h, w);
  const size_t total = static_cast<size_t>(B) * 4096 * 96;
  const int grid = grid_for(total, 256);
  using KernelFn = void (*)(const void*, int, int, int, float, float, float, float, float, float, __nv_bfloat16*);
  static const KernelFn table[2][2] = {{preprocess_patchify_kernel<false, false>, preprocess_patchify_kernel<false, true>},
                                       {preprocess_patchify_kernel<true, false>, preprocess_patchify_kernel<true, true>}};
  table[is_u8 ? 1 : 0][out_f16 ? 1 : 0]<<<grid, 256, 0, stream>>>(img, B, h, w, mean3[0], mean3[1], mean3[2], std3[0],
                                                                   std3[1], std3[2], out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_rows(const float* x, const float* gamma, const float* beta, float eps, int M, int D, void* y,
                   int out_bf16, cudaStream_t stream, int reverse) {  // out_bf16 = out_kind: 0 fp32, 1 bf16, 2 fp16
  B200SAM_REQUIRE(M > 0 && D % 128 == 0 && D <= 2048, "layernorm: D=%d must be a multiple of 128 and <= 2048", D);
  switch (D / 128) {
    case 1: return launch_ln<1>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    case 2: return launch_ln<2>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    case 6: return launch_ln<6>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    case 8: return launch_ln<8>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    case 10: return launch_ln<10>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    case 16: return launch_ln<16>(x, gamma, beta, eps, M, D, y, out_bf16, stream, reverse);
    default: break;
  }
  set_last_error("layernorm: unsupported D=%d (supported: 128,256,768,1024,1280,2048)", D);
  return 2;
}

int im2col3x3_tokens(const __nv_bfloat16* in, int B, int C, __nv_bfloat16* out, cudaStream_t stream) {
  B200SAM_REQUIRE(B > 0 && C % 8 == 0, "im2col3x3: C=%d must be a multiple of 8", C);
  const size_t total = static_cast<size_t>(B) * 4096 * 9 * (C / 8);
  im2col3x3_kernel<<<grid_for(total, 256), 256, 0, stream>>>(in, B, C, out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int layernorm_to_nchw(const float* x, const float* gamma, const float* beta, float eps, int B, int C, float* out,
                      cudaStream_t stream) {
  B200SAM_REQUIRE(B > 0 && C % 32 == 0 && C <= 512, "layernorm_to_nchw: C=%d unsupported", C);
  const int smem = C * 33 * 4;
  layernorm_to_nchw_kernel<<<B * 128, 256, smem, stream>>>(x, gamma, beta, eps, C, out);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int f32_to_op16(const float* in, __nv_bfloat16* out, size_t n, int out_f16, cudaStream_t stream) {
  if (n == 0) return 0;
  B200SAM_REQUIRE(n % 2 == 0, "f32_to_op16: element count must be even");
  if (out_f16)
    f32_to_op16_kernel<true><<<grid_for(n / 2, 256), 256, 0, stream>>>(in, reinterpret_cast<uint32_t*>(out), n / 2);
  else
    f32_to_op16_kernel<false><<<grid_for(n / 2, 256), 256, 0, stream>>>(in, reinterpret_cast<uint32_t*>(out), n / 2);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
