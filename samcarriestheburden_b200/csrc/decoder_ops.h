// Launchers of the fp32 decoder kernels (decoder_ops.cu); used by decoder.cu and the op-level C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "kernels.h"

namespace b200sam {

struct LinearArgs {
  const float* A;    // [M, K], pitch lda
  const float* A2;   // optional addend to A (e.g. positional encoding), pitch lda2
  const float* W;    // [N, K] (nn.Linear layout)
  const float* bias; // [N] or null
  const float* residual;  // [M, N] pitch ldr or null; added AFTER the activation
  float* out;        // [M, N], pitch ldo
  int M, N, K;
  int lda, lda2, ldo, ldr;
  int a2_row_mod;    // >0: A2 row = m % a2_row_mod
  int act;           // 0 none, 1 ReLU, 2 GELU(erf)
};
int linear_f32(const LinearArgs& p, cudaStream_t stream);

// q [NB,Tq,heads*dh], k/v [NB,Tk,heads*dh] -> out [NB,Tq,heads*dh]; dh in {16, 32}
// part: optional scratch of NB*heads*ATTN_FEWQ_SPLITS*Tq*(dh+2) floats enabling the key-split path for long Tk
constexpr int ATTN_FEWQ_SPLITS = 16;
// tk_valid: optional per-batch count of valid keys (<= Tk; Tk stays the row pitch) for ragged prompt batches
int attn_few_queries(const float* q, const float* k, const float* v, float* out, int NB, int Tq, int Tk, int heads,
                     int dh, float* part, const int* tk_valid, cudaStream_t stream, const int* kv_of = nullptr);
// q [NB,Nq,128], k/v [NB,Tk<=32,128] (8 heads x 16) -> out [NB,Nq,128]
// split_out != null: write the bf16 [hi | lo] split operand [NB*Nq, 256] instead of the fp32 `out`
int attn_few_keys(const float* q, const float* k, const float* v, float* out, int NB, int Nq, int Tk,
                  const int* tk_valid, __nv_bfloat16* split_out, cudaStream_t stream, const int* q_of = nullptr);
int dense_pe_tokens(const float* G, float* pe, cudaStream_t stream);
int prompt_tokens(const float* coords, const int* labels, int NB, int Np, const float* G, const float* point_emb,
                  const float* not_a_point, const float* iou_token, const float* mask_tokens, float img_w, float img_h,
                  float* tokens, int* ntok, cudaStream_t stream);
// in [n_images, 256, 4096] NCHW -> out [n_images, 4096, 256] token-major
int nchw_to_tokens(const float* in, float* out, int n_images, cudaStream_t stream);
// image_of: optional [NB] index of the image (row block of emb_tok) each prompt belongs to; null = image 0
// also emits the first layer's split operands sa = split(keys + pe), sb = split(keys) ([NB*4096, 512] bf16 each)
int keys_init(const float* emb_tok, const float* no_mask, float* keys, int NB, const int* image_of, const float* pe,
              __nv_bfloat16* sa, __nv_bfloat16* sb, cudaStream_t stream, const float* dense_tok = nullptr);
// tokens [NB, 5 + Ns, 256] = [iou_token; 4 mask tokens; sparse[b]] for caller-supplied sparse embeddings [NB, Ns, 256]
int tokens_from_sparse(const float* sparse, int NB, int Ns, const float* iou_token, const float* mask_tokens, float* tokens,
                       int* ntok, cudaStream_t stream);
int broadcast_row256(const float* v, float* out, size_t rows, cudaStream_t stream);
// in-place LayerNorm (eps 1e-5) of the [M, 256] image-side keys fused with sa = split(keys + pe[row % 4096]), sb = split(keys)
int ln256_keys_split(float* keys, const float* gamma, const float* beta, const float* pe, size_t M, __nv_bfloat16* sa,
                     __nv_bfloat16* sb, cudaStream_t stream);
int mask_downscale_keys(const float* mask, const float* const* w10, const float* emb_tok, float* keys, int NB,
                        const int* image_of, __nv_bfloat16* sb, cudaStream_t stream);
// split_out != null: write the bf16 [hi | lo] split operand [ngroups, 128] instead of updating x in place
int ln64_gelu(float* x, const float* g, const float* b, size_t ngroups, __nv_bfloat16* split_out, cudaStream_t stream);
int mlp3_tokens(const float* hs, int NB, int T, const float* const* w15, const float* const* b15, float* hyper,
                float* iou, cudaStream_t stream);
int mask_dot(const float* up, const float* hyper, int NB, int tok0, int ntok, float* masks, cudaStream_t stream);
// fp32 [M,K] (+ optional addend) -> bf16 hi/lo split operand: mode 0 activations [M,2K] = [hi|lo], mode 1 weights
// [M,3K] = [hi|hi|lo]
int split3_bf16(const float* x, const float* x2, int x2_row_mod, __nv_bfloat16* out, size_t M, int K, int mode,
                cudaStream_t stream);
int add_rows(const float* a, const float* b, float* out, size_t n, cudaStream_t stream);

}  // namespace b200sam
