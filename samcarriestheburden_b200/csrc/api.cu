// extern "C" boundary (include/b200sam.h): argument validation + error plumbing, then the launchers.
#include <stdarg.h>
#include <stdio.h>
#include "../../include/b200sam.h"
#include "decoder.h"
#include "encoder.h"
#include "kernels.h"
#include "launch.h"

namespace b200sam {
namespace {
thread_local char g_err[1024] = "";
}
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_err; }
}  // namespace b200sam

using namespace b200sam;

struct b200sam_encoder { Encoder* impl; };
struct b200sam_decoder { Decoder* impl; };
struct b200sam_unet { UNetCtx* impl; };

static EncoderConfig to_cfg(const b200sam_encoder_config* c) {
  EncoderConfig e;
  e.embed_dim = c->embed_dim; e.depth = c->depth; e.num_heads = c->num_heads;
  e.global_mask_lo = c->global_attn_mask; e.out_chans = c->out_chans;
  e.operand_format = c->operand_format; e.flags = c->flags;
  return e;
}

extern "C" {

const char* b200sam_last_error(void) { return get_last_error(); }
int b200sam_abi_version(void) { return 2; }

int b200sam_encoder_weight_count(const b200sam_encoder_config* cfg) {
  if (!cfg) return -1;
  return encoder_weight_count(to_cfg(cfg));
}
const char* b200sam_encoder_weight_name(const b200sam_encoder_config* cfg, int i) {
  if (!cfg) return nullptr;
  return encoder_weight_name(to_cfg(cfg), i);
}
size_t b200sam_encoder_workspace_bytes(const b200sam_encoder_config* cfg, int batch) {
  if (!cfg) return 0;
  return encoder_workspace_bytes(to_cfg(cfg), batch);
}
int b200sam_encoder_create(const b200sam_encoder_config* cfg, const void* const* weights, int n_weights,
                           b200sam_encoder** out) {
  if (!cfg || !weights || !out) { set_last_error("encoder_create: null argument"); return 2; }
  Encoder* e = nullptr;
  if (int rc = encoder_create(to_cfg(cfg), weights, n_weights, &e)) return rc;
  *out = new b200sam_encoder{e};
  return 0;
}
void b200sam_encoder_destroy(b200sam_encoder* enc) {
  if (!enc) return;
  encoder_destroy(enc->impl);
  delete enc;
}
int b200sam_encoder_forward(const b200sam_encoder* enc, const void* image, int is_u8, int batch, int h, int w,
                            const float* mean3, const float* std3, float* embedding_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  if (!enc || !mean3 || !std3) { set_last_error("encoder_forward: null argument"); return 2; }
  return encoder_forward(enc->impl, image, is_u8, batch, h, w, mean3, std3, embedding_out, workspace, workspace_bytes,
                         static_cast<cudaStream_t>(stream));
}

size_t b200sam_prompt_extract_scratch_bytes(int n_img, int n_classes) {
  if (n_img < 0 || n_classes < 0) return 0;
  return prompt_extract_scratch_bytes(n_img, n_classes);
}
int b200sam_prompt_extract(const uint8_t* masks, int n_img, int n_classes, int H, int W, int32_t* seeds,
                           int32_t* boxes, uint8_t* has_seed, uint8_t* has_box, void* scratch, void* stream) {
  if (n_img > 0 && n_classes > 0 && (!masks || !seeds || !boxes || !has_seed || !has_box || !scratch)) {
    set_last_error("prompt_extract: null argument");
    return 2;
  }
  return prompt_extract(masks, n_img, n_classes, H, W, seeds, boxes, has_seed, has_box,
                        static_cast<int32_t*>(scratch), static_cast<cudaStream_t>(stream));
}

int b200sam_decoder_weight_count(void) { return decoder_weight_count(); }
const char* b200sam_decoder_weight_name(int i) { return decoder_weight_name(i); }
size_t b200sam_decoder_workspace_bytes(int n_prompts, int n_points) {
  return decoder_workspace_bytes(1, n_prompts, n_points);
}
size_t b200sam_decoder_workspace_bytes_batch(int n_images, int n_prompts, int n_points) {
  return decoder_workspace_bytes(n_images, n_prompts, n_points);
}
int b200sam_decoder_create(const void* const* weights, int n_weights, b200sam_decoder** out, void* stream) {
  if (!weights || !out) { set_last_error("decoder_create: null argument"); return 2; }
  Decoder* d = nullptr;
  if (int rc = decoder_create(weights, n_weights, &d, static_cast<cudaStream_t>(stream))) return rc;
  *out = new b200sam_decoder{d};
  return 0;
}
void b200sam_decoder_destroy(b200sam_decoder* dec) {
  if (!dec) return;
  decoder_destroy(dec->impl);
  delete dec;
}
int b200sam_decoder_copy_dense_pe(const b200sam_decoder* dec, float* out, void* stream) {
  if (!dec || !out) { set_last_error("decoder_copy_dense_pe: null argument"); return 2; }
  B200SAM_CHECK_CUDA(cudaMemcpyAsync(out, decoder_dense_pe(dec->impl), 4096 * 256 * sizeof(float),
                                     cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}
int b200sam_decode(const b200sam_decoder* dec, const float* embedding, int n_prompts, int n_points,
                   const float* coords, const int32_t* labels, const float* mask_prev, int multimask,
                   float* low_res_out, float* iou_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dec) { set_last_error("decode: null decoder"); return 2; }
  return b200sam_decode_batch(dec, embedding, 1, nullptr, n_prompts, n_points, coords, labels, mask_prev, multimask,
                              low_res_out, iou_out, workspace, workspace_bytes, stream);
}
int b200sam_decode_batch(const b200sam_decoder* dec, const float* embeddings, int n_images, const int32_t* image_of,
                         int n_prompts, int n_points, const float* coords, const int32_t* labels,
                         const float* mask_prev, int multimask, float* low_res_out, float* iou_out, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (!dec) { set_last_error("decode: null decoder"); return 2; }
  DecodeArgs a;
  a.emb = embeddings; a.n_images = n_images; a.image_of = image_of; a.NB = n_prompts; a.Np = n_points;
  a.coords = coords; a.labels = labels; a.mask_prev = mask_prev;
  a.img_w = 1024.0f; a.img_h = 1024.0f; a.multimask = multimask; a.low_res_out = low_res_out; a.iou_out = iou_out;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return decoder_forward(dec->impl, a, static_cast<cudaStream_t>(stream));
}

int b200sam_prompt_encode(const b200sam_decoder* dec, const float* coords, const int32_t* labels, int n_prompts,
                          int n_points, const float* mask_prev, float* tokens_tmp, int32_t* ntok_tmp, float* sparse_out,
                          float* dense_tok_out, void* stream) {
  if (!dec) { set_last_error("prompt_encode: null decoder"); return 2; }
  return prompt_encode(dec->impl, coords, labels, n_prompts, n_points, mask_prev, tokens_tmp, ntok_tmp, sparse_out,
                       dense_tok_out, static_cast<cudaStream_t>(stream));
}
int b200sam_decode_embedded(const b200sam_decoder* dec, const float* embeddings, int n_images, const int32_t* image_of,
                            int n_prompts, int n_sparse, const float* sparse, const float* dense_tok, int multimask,
                            float* low_res_out, float* iou_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!dec) { set_last_error("decode_embedded: null decoder"); return 2; }
  if (!dense_tok || (n_sparse > 0 && !sparse)) { set_last_error("decode_embedded: null embeddings"); return 2; }
  DecodeArgs a;
  a.emb = embeddings; a.n_images = n_images; a.image_of = image_of; a.NB = n_prompts; a.Np = n_sparse;
  a.coords = nullptr; a.labels = nullptr; a.mask_prev = nullptr; a.sparse_tokens = sparse; a.dense_tok = dense_tok;
  a.img_w = 1024.0f; a.img_h = 1024.0f; a.multimask = multimask; a.low_res_out = low_res_out; a.iou_out = iou_out;
  a.workspace = workspace; a.workspace_bytes = workspace_bytes;
  return decoder_forward(dec->impl, a, static_cast<cudaStream_t>(stream));
}

int b200sam_upscale_threshold(const float* low_res, int n, int low, int img_size, int in_h, int in_w, int out_h,
                              int out_w, float threshold, uint8_t* mask_out, float* logits_out, uint8_t* small_out,
                              int small_h, int small_w, void* stream) {
  if (n > 0 && !low_res) { set_last_error("upscale: null input"); return 2; }
  return upscale_threshold(low_res, n, low, img_size, in_h, in_w, out_h, out_w, threshold, mask_out, logits_out,
                           small_out, small_h, small_w, static_cast<cudaStream_t>(stream));
}

int b200sam_unet_weight_count(void) { return unet_weight_count(); }
const char* b200sam_unet_weight_name(int i) { return unet_weight_name(i); }
int b200sam_unet_conv_kp(int cin) { return unet_conv_kp(cin); }
int b200sam_unet_create(int n_channels, int n_classes, int n_last_channel, const void* const* weights, int n_weights,
                        b200sam_unet** out, void* stream) {
  if (!weights || !out) { set_last_error("unet_create: null argument"); return 2; }
  UNetCtx* u = nullptr;
  if (int rc = unet_create(n_channels, n_classes, n_last_channel, weights, n_weights, &u, static_cast<cudaStream_t>(stream)))
    return rc;
  *out = new b200sam_unet{u};
  return 0;
}
void b200sam_unet_destroy(b200sam_unet* u) {
  if (!u) return;
  unet_destroy(u->impl);
  delete u;
}
size_t b200sam_unet_workspace_bytes(const b200sam_unet* u, int batch, int H, int W) {
  return u ? unet_workspace_bytes(u->impl, batch, H, W) : 0;
}
int b200sam_unet_forward(const b200sam_unet* u, const float* image, int batch, int H, int W, float* logits_out,
                         float* probs_out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!u) { set_last_error("unet_forward: null handle"); return 2; }
  return unet_forward(u->impl, image, batch, H, W, logits_out, probs_out, workspace, workspace_bytes,
                      static_cast<cudaStream_t>(stream));
}

int b200sam_resize_ksize(int in_size, int out_size) { return resize_ksize(in_size, out_size); }
int b200sam_resize_coeffs_host(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host) {
  return resize_coeffs_host(in_size, out_size, bounds_host, kk_host);
}
int b200sam_resize_u8(const uint8_t* image, int H, int W, int C, const int32_t* xbounds, const int32_t* xkk, int xksize,
                      const int32_t* ybounds, const int32_t* ykk, int yksize, int out_h, int out_w, uint8_t* tmp,
                      uint8_t* out, int out_chw, void* stream) {
  return resize_u8(image, H, W, C, xbounds, xkk, xksize, ybounds, ykk, yksize, out_h, out_w, tmp, out, out_chw,
                   static_cast<cudaStream_t>(stream));
}

int b200sam_cvresize_coeffs_host(int in_size, int out_size, int clamp_weights, int32_t* idx2_host, int32_t* w2_host) {
  return cvresize_coeffs_host(in_size, out_size, clamp_weights, idx2_host, w2_host);
}
int b200sam_cvresize_linear_u8(const uint8_t* image, int n, int H, int W, const int32_t* xidx, const int32_t* xw,
                               const int32_t* yidx, const int32_t* yw, int out_h, int out_w, uint8_t* out_u8,
                               float* out_norm, float mean, float std, void* stream) {
  return cvresize_linear_u8(image, n, H, W, xidx, xw, yidx, yw, out_h, out_w, out_u8, out_norm, mean, std,
                            static_cast<cudaStream_t>(stream));
}

int b200sam_cvresize_cubic_coeffs_host(int in_size, int out_size, int32_t* idx4_host, int32_t* w4_host) {
  return cvresize_cubic_coeffs_host(in_size, out_size, idx4_host, w4_host);
}
int b200sam_medsam_preprocess(const uint8_t* gray, int H, int W, const int32_t* xidx, const int32_t* xw, const int32_t* yidx,
                              const int32_t* yw, int size, uint8_t* resized_u8, int32_t* minmax, float* out3, void* stream) {
  return medsam_preprocess(gray, H, W, xidx, xw, yidx, yw, size, resized_u8, minmax, out3, static_cast<cudaStream_t>(stream));
}

int b200sam_stability_score(const float* logits, int n, int H, int W, float threshold_hi, float threshold_lo,
                            float* score_out, int32_t* scratch, void* stream) {
  return stability_score(logits, n, H, W, threshold_hi, threshold_lo, score_out, scratch,
                         static_cast<cudaStream_t>(stream));
}
int b200sam_mask_to_box(const uint8_t* masks, int n, int H, int W, int64_t* boxes_out, int32_t* scratch, void* stream) {
  return mask_to_box(masks, n, H, W, reinterpret_cast<long long*>(boxes_out), scratch, static_cast<cudaStream_t>(stream));
}

size_t b200sam_ccl_scratch_bytes(int n_planes, int H, int W) {
  if (n_planes < 0 || H <= 0 || W <= 0) return 0;
  return ccl_scratch_bytes(n_planes, H, W);
}
int b200sam_ccl_select(const float* prob, int n_planes, int planes_per_call, int H, int W, float threshold, int by_area, float* out,
                       void* scratch, void* stream) {
  return ccl_select(prob, n_planes, planes_per_call, H, W, threshold, by_area, out, scratch, static_cast<cudaStream_t>(stream));
}
int b200sam_morph_flat(const float* in, int n_planes, int H, int W, const uint8_t* se, int kh, int kw, int origin_y,
                       int origin_x, int dilate, float* out, void* stream) {
  return morph_flat(in, n_planes, H, W, se, kh, kw, origin_y, origin_x, dilate, out, static_cast<cudaStream_t>(stream));
}

static int gemm16(const void* A, const void* W, void* out, const float* bias, const float* residual, int M, int N, int K,
                  int lda, int ldb, int ldo, int ldr, int res_row_mod, int gelu, int out_16bit, int max_ctas, int f16,
                  void* stream) {
  if (!A || !W || !out) { set_last_error("gemm: null argument"); return 2; }
  GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A); g.B = static_cast<const __nv_bfloat16*>(W); g.out = out;
  g.bias = bias; g.residual = residual; g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldo = ldo; g.ldr = ldr;
  g.res_row_mod = res_row_mod; g.gelu = gelu; g.out_kind = out_16bit ? (f16 ? 2 : 1) : 0; g.max_ctas = max_ctas;
  g.op_f16 = f16;
  return gemm_bf16_tn(g, static_cast<cudaStream_t>(stream));
}
int b200sam_gemm_bf16(const void* A, const void* W, void* out, const float* bias, const float* residual, int M, int N,
                      int K, int lda, int ldb, int ldo, int ldr, int res_row_mod, int gelu, int out_bf16, int max_ctas,
                      void* stream) {
  return gemm16(A, W, out, bias, residual, M, N, K, lda, ldb, ldo, ldr, res_row_mod, gelu, out_bf16, max_ctas, 0, stream);
}
int b200sam_gemm_f16(const void* A, const void* W, void* out, const float* bias, const float* residual, int M, int N,
                     int K, int lda, int ldb, int ldo, int ldr, int res_row_mod, int gelu, int out_f16, int max_ctas,
                     void* stream) {
  return gemm16(A, W, out, bias, residual, M, N, K, lda, ldb, ldo, ldr, res_row_mod, gelu, out_f16, max_ctas, 1, stream);
}
int b200sam_gemm_ln_residual(const void* A, const void* W, const float* bias, const float* residual, float* out,
                             void* out16, float* rowstat_out, int M, int N, int K, int operand_format, void* stream) {
  if (!A || !W || !out || !out16 || !rowstat_out) { set_last_error("gemm_ln_residual: null argument"); return 2; }
  GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A); g.B = static_cast<const __nv_bfloat16*>(W); g.out = out;
  g.bias = bias; g.residual = residual; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldo = N; g.ldr = N;
  g.res_row_mod = 0; g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.op_f16 = operand_format == 1;
  g.xh = out16; g.rowstat_out = rowstat_out;
  return gemm_bf16_tn(g, static_cast<cudaStream_t>(stream));
}
int b200sam_gemm_ln_folded(const void* A, const void* W_folded, const float* bias_folded, const float* colsum,
                           const float* rowstat_in, int nparts, float eps, void* out16, int M, int N, int K, int gelu,
                           int operand_format, void* stream) {
  if (!A || !W_folded || !bias_folded || !colsum || !rowstat_in || !out16) {
    set_last_error("gemm_ln_folded: null argument");
    return 2;
  }
  GemmArgs g;
  g.A = static_cast<const __nv_bfloat16*>(A); g.B = static_cast<const __nv_bfloat16*>(W_folded); g.out = out16;
  g.bias = bias_folded; g.residual = nullptr; g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldo = N; g.ldr = 0;
  g.res_row_mod = 0; g.gelu = gelu; g.op_f16 = operand_format == 1; g.out_kind = g.op_f16 ? 2 : 1; g.max_ctas = 0;
  g.rowstat_in = rowstat_in; g.colsum = colsum; g.nparts_in = nparts; g.ln_dim = K; g.ln_eps = eps;
  return gemm_bf16_tn(g, static_cast<cudaStream_t>(stream));
}
int b200sam_layernorm(const float* x, const float* gamma, const float* beta, float eps, int M, int D, void* y,
                      int out_kind, void* stream) {
  if (!x || !gamma || !beta || !y) { set_last_error("layernorm: null argument"); return 2; }
  if (out_kind < 0 || out_kind > 2) { set_last_error("layernorm: out_kind %d (0 fp32, 1 bf16, 2 fp16)", out_kind); return 2; }
  return layernorm_rows(x, gamma, beta, eps, M, D, y, out_kind, static_cast<cudaStream_t>(stream));
}
int b200sam_encoder_attention(const void* qkv, const void* qkv_bias16, const void* rel_h16, const void* rel_w16,
                              void* out, int batch, int heads, int hd, int global_attn, int operand_format,
                              void* stream) {
  AttnArgs a;
  a.qkv = static_cast<const __nv_bfloat16*>(qkv); a.qkv_bias = static_cast<const __nv_bfloat16*>(qkv_bias16);
  a.rel_h = static_cast<const __nv_bfloat16*>(rel_h16); a.rel_w = static_cast<const __nv_bfloat16*>(rel_w16);
  a.out = static_cast<__nv_bfloat16*>(out); a.B = batch; a.heads = heads; a.hd = hd; a.f16 = operand_format == 1;
  if (operand_format != 0 && operand_format != 1) { set_last_error("encoder_attention: operand_format %d", operand_format); return 2; }
  if (global_attn == 1) return global_attention_tc(a, static_cast<cudaStream_t>(stream));
  if (global_attn == 0) return window_attention_tc(a, static_cast<cudaStream_t>(stream));
  set_last_error("encoder_attention: global_attn must be 0 (14x14 windows) or 1 (global), got %d", global_attn);
  return 2;
}
int b200sam_preprocess_patchify(const void* image, int is_u8, int batch, int h, int w, const float* mean3,
                                const float* std3, void* out16, int operand_format, void* stream) {
  if (!image || !mean3 || !std3 || !out16) { set_last_error("preprocess: null argument"); return 2; }
  return preprocess_patchify(image, is_u8, batch, h, w, mean3, std3, static_cast<__nv_bfloat16*>(out16),
                             operand_format == 1, static_cast<cudaStream_t>(stream));
}
int b200sam_set_gemm_pair(int mode) { gemm_pair_set_mode(mode); return 0; }
int b200sam_gemm_pair_max_clusters(void) { return gemm_pair_max_clusters(); }
int b200sam_timing_start(int capacity) { return timing_start(capacity); }
int b200sam_timing_stop(int* kinds_host, double* work_host, int* dims3_host, float* ms_host, int capacity, int* n_out_host) {
  return timing_stop(kinds_host, work_host, dims3_host, ms_host, capacity, n_out_host);
}
int b200sam_linear_f32(const float* A, const float* A2, int a2_row_mod, const float* W, const float* bias,
                       const float* residual, float* out, int M, int N, int K, int act, void* stream) {
  if (!A || !W || !out) { set_last_error("linear_f32: null argument"); return 2; }
  LinearArgs p;
  p.A = A; p.A2 = A2; p.W = W; p.bias = bias; p.residual = residual; p.out = out; p.M = M; p.N = N; p.K = K;
  p.lda = K; p.lda2 = K; p.ldo = N; p.ldr = N; p.a2_row_mod = a2_row_mod; p.act = act;
  return linear_f32(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
