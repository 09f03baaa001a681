// Internal host-side launch interface shared by the .cu translation units and api.cu.
// Everything here takes raw device pointers + a stream; no allocation, no synchronisation.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace b200sam {

void set_last_error(const char* fmt, ...);
const char* get_last_error();

// error plumbing: launchers return 0 on success and leave a message for b200sam_last_error() otherwise
#define B200SAM_CHECK_CUDA(expr)                                                        \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      b200sam::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,             \
                              cudaGetErrorString(_e));                                  \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)
#define B200SAM_REQUIRE(cond, ...)                                                      \
  do {                                                                                  \
    if (!(cond)) {                                                                      \
      b200sam::set_last_error(__VA_ARGS__);                                             \
      return 2;                                                                         \
    }                                                                                   \
  } while (0)


// ---- gemm_tcgen05.cu -------------------------------------------------------------------------
struct GemmArgs {
  const __nv_bfloat16* A;  // [M, K] row-major, pitch lda (16-bit elements: bf16, or fp16 with op_f16)
  const __nv_bfloat16* B;  // [N, K] row-major (nn.Linear weight), pitch ldb
  void* out;               // 16-bit or fp32 [M, N], pitch ldo
  const float* bias;       // [N] or null
  const float* residual;   // fp32, pitch ldr, or null (only with fp32 out)
  int M, N, K;
  int lda, ldb, ldo, ldr;
  int res_row_mod;  // >0: residual row index = row % res_row_mod
  long long out_plane = 0;  // != 0: every 128-column tile j writes its own [M, 128] output at out + j * out_plane (fp32, ldo = 128)
  long long res_plane = 0;  //       and reads its residual table at residual + j * res_plane
  int gelu;         // activation after bias: 0 none, 1 exact erf GELU, 2 ReLU
  int out_kind;     // 0: fp32 output, 1: bf16, 2: fp16 (a 16-bit output has the operands' format)
  int max_ctas;     // <= 0: one CTA per SM; > 0 caps the persistent grid
  int a_wrap = 0;   // > 0: A has only a_wrap columns; k >= a_wrap reads column k - a_wrap ([hi|lo|hi] stored as [hi|lo])
  int conv_cin = 0;  // > 0: implicit 3x3 convolution over a zero-bordered pixel grid (see gemm_tcgen05.cu); K = 27 * conv_cin
  int conv_wp = 0;   //      padded row length W + 2
  int reverse_m = 0; // 1: walk the row blocks last-to-first (start with what the producer kernel left in L2)
  // fused mask-decoder upscaler epilogues (gemm_epilogue.cuh): 0 none, 1 LN2d(64)+GELU+split (N = 256), 2 GELU+hyper dot (N = 128)
  int epi_mode = 0;
  const float* aux0 = nullptr;  // mode 1: LN gamma [64]; mode 2: hyper [prompts, 4, 32]
  const float* aux1 = nullptr;  // mode 1: LN beta [64]
  int tok0 = 0, ntok = 0;       // mode 2
  int use_pair = 0;             // CTA-pair kernel: 0 = follow gemm_pair_enabled(), 1 = force, -1 = never
  int op_f16 = 0;               // A / B are fp16 instead of bf16 (the ViT encoder's default)
  // LayerNorm folded into the GEMMs around it (gemm_epilogue.cuh): producer side (fp32 out) ...
  void* xh = nullptr;             // 16-bit copy of out, pitch ldo
  float* rowstat_out = nullptr;   // [M, N/64, 2] partial (sum, sum sq) of out per 64-column part (N % 128 == 0)
  // ... consumer side (16-bit out): out = rstd * (acc - mean * colsum[n]) + bias[n]
  const float* rowstat_in = nullptr;  // [M, nparts_in, 2]
  const float* colsum = nullptr;      // [N]
  int nparts_in = 0;
  int ln_dim = 0;       // number of elements the statistics cover (= K)
  float ln_eps = 0.0f;
};
int gemm_bf16_tn(const GemmArgs& g, cudaStream_t stream);
// CTA-pair (tcgen05 cta_group::2, 256 x 256 tiles) variant for the encoder's plain linears (gemm_pair.cu);
// gemm_bf16_tn dispatches to it when enabled (B200SAM_GEMM_PAIR != 0 or gemm_pair_set_mode) and eligible
bool gemm_pair_enabled();
void gemm_pair_set_mode(int mode);  // -1: environment decides, 0: off, 1: on
bool gemm_pair_eligible(const GemmArgs& g);
int gemm_pair_max_clusters();
int gemm_f16_tn_pair(const GemmArgs& g, cudaStream_t stream);

// ---- encoder_ops.cu --------------------------------------------------------------------------
// image [B,3,h,w] (uint8 or fp32) -> normalised, zero-padded 1024^2, im2col'd bf16 [B*4096, 768]
int preprocess_patchify(const void* img, int is_u8, int B, int h, int w, const float* mean3, const float* std3,
                        __nv_bfloat16* out, int out_f16, cudaStream_t stream);
// y = LN(x) over the last dim (fp32 statistics, biased variance); x fp32 [M,D]; y: out_kind 0 fp32, 1 bf16, 2 fp16
int layernorm_rows(const float* x, const float* gamma, const float* beta, float eps, int M, int D, void* y,
                   int out_kind, cudaStream_t stream, int reverse = 0);
// 3x3/pad1 im2col over a 64x64 token grid: in bf16 [B*4096, C] -> out bf16 [B*4096, 9*C] (tap-major)
int im2col3x3_tokens(const __nv_bfloat16* in, int B, int C, __nv_bfloat16* out, cudaStream_t stream);
// LayerNorm2d over channels + token-major -> NCHW transpose: in fp32 [B*4096, C] -> out fp32 [B,C,64,64]
int layernorm_to_nchw(const float* x, const float* gamma, const float* beta, float eps, int B, int C, float* out,
                      cudaStream_t stream);
int f32_to_op16(const float* in, __nv_bfloat16* out, size_t n, int out_f16, cudaStream_t stream);

// ---- attention_tc.cu -------------------------------------------------------------------------
struct AttnArgs {
  const __nv_bfloat16* qkv;       // [B*4096, 3*D] 16-bit (q | k | v, each heads x hd)
  const __nv_bfloat16* qkv_bias;  // [3*D] 16-bit (K/V of zero-padded window tokens)
  const __nv_bfloat16* rel_h;     // [2S-1, hd] 16-bit
  const __nv_bfloat16* rel_w;     // [2S-1, hd] 16-bit
  __nv_bfloat16* out;             // [B*4096, D] 16-bit
  int B, heads, hd;
  int reverse = 0;  // walk the images last-to-first (start with what the qkv GEMM left in L2)
  int f16 = 0;      // all 16-bit tensors are fp16 instead of bf16
};
int window_attention_tc(const AttnArgs& a, cudaStream_t stream);  // 14x14 windows on tcgen05 / TMEM
int global_attention_tc(const AttnArgs& a, cudaStream_t stream);  // full 4096x4096 on tcgen05 / TMEM

// ---- prompt_extract.cu -----------------------------------------------------------------------
int prompt_extract(const uint8_t* masks, int n_img, int C, int H, int W, int32_t* seeds, int32_t* boxes,
                   uint8_t* has_seed, uint8_t* has_box, int32_t* scratch, cudaStream_t stream);
size_t prompt_extract_scratch_bytes(int n_img, int C);

// ---- upscale.cu ------------------------------------------------------------------------------
int upscale_threshold(const float* low_res, int n, int low, int img_size, int in_h, int in_w, int out_h, int out_w,
                      float thresh, uint8_t* mask_out, float* logits_out, uint8_t* small_out, int small_h,
                      int small_w, cudaStream_t stream);

// ---- ccl.cu ----------------------------------------------------------------------------------
// prob [n_planes, H, W] fp32 -> out = prob * (winning 8-connected component of prob > threshold), per plane
size_t ccl_scratch_bytes(int n_planes, int H, int W);
// planes_per_call: planes the reference labels in ONE call (its label 0 = pixel 0 of the call's first plane is background);
// <= 0: all n_planes are one call
int ccl_select(const float* prob, int n_planes, int planes_per_call, int H, int W, float threshold, int by_area, float* out, void* scratch,
               cudaStream_t stream);
// flat (0/1 structuring element) grey-scale dilation / erosion, out-of-image taps ignored
int morph_flat(const float* in, int n_planes, int H, int W, const uint8_t* se, int kh, int kw, int origin_y,
               int origin_x, int dilate, float* out, cudaStream_t stream);

// ---- resize.cu -------------------------------------------------------------------------------
// Pillow-exact antialiased bilinear uint8 resize (ResizeLongestSide.apply_image); coefficient tables built on the host
int resize_ksize(int in_size, int out_size);
int resize_coeffs_host(int in_size, int out_size, int32_t* bounds, int32_t* kk);
int resize_u8(const uint8_t* in, int H, int W, int C, const int32_t* xbounds, const int32_t* xkk, int xksize,
              const int32_t* ybounds, const int32_t* ykk, int yksize, int out_h, int out_w, uint8_t* tmp, uint8_t* out,
              int out_chw, cudaStream_t stream);

// OpenCV-exact uint8 INTER_LINEAR resize (+ optional /255 and (x - mean) / std) for the U-Net ingest; tables: idx [out, 2]
// (the two taps), w [out, 2] (11-bit weights); clamp_weights = 1 for the x axis, 0 for the y axis
int cvresize_coeffs_host(int in_size, int out_size, int clamp_weights, int32_t* idx2, int32_t* w2);
int cvresize_linear_u8(const uint8_t* in, int n, int H, int W, const int32_t* xi, const int32_t* xw, const int32_t* yi,
                       const int32_t* yw, int out_h, int out_w, uint8_t* out_u8, float* out_norm, float mean, float sd,
                       cudaStream_t stream);

// MedSAM ingest: OpenCV-exact uint8 INTER_CUBIC resize of a grey image to size x size, min-max normalisation in float64,
// three identical fp32 channels [3, size, size]; tables: idx [size, 4] clamped taps, w [size, 4] 11-bit weights
int cvresize_cubic_coeffs_host(int in_size, int out_size, int32_t* idx4, int32_t* w4);
int medsam_preprocess(const uint8_t* gray, int H, int W, const int32_t* xi, const int32_t* xw, const int32_t* yi,
                      const int32_t* yw, int size, uint8_t* tmp_u8, int32_t* minmax, float* out3, cudaStream_t stream);

// ---- amg.cu ----------------------------------------------------------------------------------
// stability score (count(x > thr + off) / count(x > thr - off) per mask) and XYXY boxes of bool masks (amg.py)
int stability_score(const float* logits, int n, int H, int W, float threshold_hi, float threshold_lo,
                    float* score_out, int32_t* scratch, cudaStream_t stream);
int mask_to_box(const uint8_t* masks, int n, int H, int W, long long* boxes_out, int32_t* scratch, cudaStream_t stream);

// ---- unet.cu -----------------------------------------------------------------------------------
struct UNetCtx;
int unet_weight_count();
const char* unet_weight_name(int i);
int unet_conv_kp(int cin);  // padded K of a 3x3 convolution with cin input channels
int unet_create(int n_channels, int n_classes, int n_last, const void* const* weights, int n, UNetCtx** out,
                cudaStream_t stream);
void unet_destroy(UNetCtx* u);
size_t unet_workspace_bytes(const UNetCtx* u, int B, int H, int W);
// image [B, 1, H, W] fp32 (already normalised) -> logits / sigmoid probabilities [B, n_classes, H, W] (either may be null)
int unet_forward(const UNetCtx* u, const float* image, int B, int H, int W, float* logits_out, float* probs_out,
                 void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace b200sam
