// Decoder context: borrowed fp32 weight pointers (owned by the Python modules) + the cached dense PE.
#pragma once
#include <vector>
#include "decoder_ops.h"

namespace b200sam {

struct Decoder {
  std::vector<const float*> w;  // in decoder_weight_name() order
  float* pe_tok;                // [4096, 256] token-major dense positional encoding (owned)
  // hi/hi/lo bf16 splits ([N, 3K]) of the weights of the image-side linears (owned, one allocation):
  // per layer l: t2i k, t2i v, i2t q, i2t out, mlp 1, mlp 2; then final k, final v, upscale ConvT1, upscale ConvT2
  __nv_bfloat16* wsplit;
  const __nv_bfloat16* ws_t2i_k[2];
  const __nv_bfloat16* ws_t2i_v[2];
  const __nv_bfloat16* ws_i2t_q[2];
  const __nv_bfloat16* ws_i2t_o[2];
  const __nv_bfloat16* ws_mlp1[2];  // token-side MLP (transformer.py:172-174): 256 -> 2048 (ReLU) -> 256
  const __nv_bfloat16* ws_mlp2[2];
  const __nv_bfloat16* ws_fin_k;
  const __nv_bfloat16* ws_fin_v;
  const __nv_bfloat16* ws_up1;
  const __nv_bfloat16* ws_up2;
  // positional-encoding terms of the projections that consume keys + pe, folded into per-token tables (owned):
  // (keys + pe) W^T + b = keys W^T + (pe W^T + b); pek_* = pe W^T + b, fp32 [4096, 128], added by the GEMM epilogue as a
  // residual indexed by row % 4096, so only split(keys) is ever materialised (not split(keys + pe) as well)
  // The k, v and image-side q projections of a layer read the same split(keys): they run as ONE tall-tile GEMM with N = 384
  // whose three 128-column tiles write three output planes (gemm "planes"); the tables are therefore laid out per layer as
  // [t2i k | zeros (v has a plain bias) | i2t q], then [final k | zeros], 4096 x 128 floats each, and the biases as
  // [0 | v bias | 0] per layer, [0 | v bias] for the final attention.
  float* pek;  // 8 tables
  const float* pek_t2i_k[2];
  const float* pek_i2t_q[2];
  const float* pek_fin_k;
  float* bias_kvq;  // 2 x 384 + 256 floats (owned)
  const float* bias_kvq_l[2];
  const float* bias_fin_kv;
};

struct DecodeArgs {
  const float* emb;        // [n_images, 256, 64, 64] fp32 image embeddings (NCHW)
  int n_images;            // >= 1
  const int* image_of;     // [NB] device: image index of every prompt; null = all prompts belong to image 0
  int NB;                  // prompts in this batch (of all images)
  int Np;                  // sparse point slots per prompt (incl. pad / box corners)
  const float* coords;     // [NB, Np, 2] (x, y) in the encoder input frame
  const int* labels;       // [NB, Np]: -1 pad, 0 neg, 1 pos, 2/3 box corners, -2 absent slot (trailing; ragged batches)
  const float* mask_prev;  // [NB, 256, 256] logits of a previous pass, or null
  // standalone MaskDecoder.forward (mask_decoder.py:71-110): caller-supplied embeddings instead of prompts
  const float* sparse_tokens = nullptr;  // [NB, Np, 256] sparse prompt embeddings (coords / labels unused)
  const float* dense_tok = nullptr;      // [NB, 4096, 256] token-major dense prompt embeddings (mask_prev unused)
  float img_w, img_h;      // prompt_encoder.input_image_size (W, H) = (1024, 1024)
  int multimask;           // 0: mask token 0 only, 1: tokens 1..3
  float* low_res_out;      // [NB, 1|3, 256, 256]
  float* iou_out;          // [NB, 1|3]
  void* workspace;
  size_t workspace_bytes;
};

int decoder_weight_count();
const char* decoder_weight_name(int i);
size_t decoder_workspace_bytes(int n_images, int NB, int Np);
int decoder_create(const void* const* weights, int n, Decoder** out, cudaStream_t stream);
void decoder_destroy(Decoder* d);
const float* decoder_dense_pe(const Decoder* d);
int decoder_forward(const Decoder* d, const DecodeArgs& a, cudaStream_t stream);
// PromptEncoder.forward alone (prompt_encoder.py:128-168): sparse_out [NB, Np, 256], dense_tok_out [NB, 4096, 256]
// (token-major); tokens_tmp [NB, 5 + Np, 256] and ntok_tmp [NB] are scratch
int prompt_encode(const Decoder* d, const float* coords, const int* labels, int NB, int Np, const float* mask_prev,
                  float* tokens_tmp, int* ntok_tmp, float* sparse_out, float* dense_tok_out, cudaStream_t stream);

}  // namespace b200sam
