/* b200sam C ABI — B200 (sm_100a) kernels for the SAM pseudo-label refinement hot path.
 *
 * The reference (multimodallearning/SamCarriesTheBurden) is pure Python/PyTorch and has no FFI boundary;
 * each entry point below names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns all memory;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work (no allocation in hot calls, no
 *     synchronisation) except *_create, which may allocate small constant tables;
 *   - return 0 on success; non-zero on error with a message available from b200sam_last_error()
 *     (thread-local).  There is no CPU fallback: without a CUDA device the compute calls fail.
 */
#ifndef B200SAM_H
#define B200SAM_H
#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define B200SAM_API __attribute__((visibility("default")))
#else
#define B200SAM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

B200SAM_API const char* b200sam_last_error(void);
B200SAM_API int b200sam_abi_version(void);

/* ---------------------------------------------------------------- image encoder
 * Replaces Sam.preprocess + ImageEncoderViT.forward
 * (segment_anything/modeling/sam.py:164-174, modeling/image_encoder.py:106-116) as called from
 * SamPredictor.set_torch_image (segment_anything/predictor.py:62-90) and
 * scripts/generate_img_embeddings.py:45,62. */
typedef struct b200sam_encoder_config {
  int embed_dim;        /* 768 | 1024 | 1280            (build_sam.py:14-44) */
  int depth;            /* 12 | 24 | 32 */
  int num_heads;        /* 12 | 16 | 16 */
  int global_attn_mask; /* bit i set: block i is a global-attention block */
  int out_chans;        /* 256 */
  int operand_format;   /* 16-bit tensor-core operand format: 0 = bf16, 1 = fp16 (11 significand bits; what the
                           reference's own mixed-precision run uses: torch.cuda.amp.autocast defaults to fp16,
                           seg_processing/hpo_bce_unet_sam_postprocess.py:44).  fp32 accumulate, fp32 residual stream,
                           fp32 LayerNorm / softmax statistics in both. */
  int flags;            /* B200SAM_ENC_LN_FUSED: norm1 / norm2 (image_encoder.py:168,180) folded into the GEMMs */
} b200sam_encoder_config;
#define B200SAM_OPERAND_BF16 0
#define B200SAM_OPERAND_FP16 1
#define B200SAM_ENC_LN_FUSED 1
typedef struct b200sam_encoder b200sam_encoder;

B200SAM_API int b200sam_encoder_weight_count(const b200sam_encoder_config* cfg);
/* "state_dict key|packing[|norm prefix]" of weight slot i; packing: f32 | op16 | op16_flat | op16_tap | f32_tokens |
 * none (unused slot, any non-null pointer) and, with B200SAM_ENC_LN_FUSED, for the linear `key` that follows the
 * LayerNorm `norm prefix`: fold_w = op16(gamma * W), fold_s[n] = sum_k fold_w[n,k] (fp32), fold_c = beta . W^T + b (fp32);
 * op16 = cfg->operand_format. */
B200SAM_API const char* b200sam_encoder_weight_name(const b200sam_encoder_config* cfg, int i);
B200SAM_API size_t b200sam_encoder_workspace_bytes(const b200sam_encoder_config* cfg, int batch);
B200SAM_API int b200sam_encoder_create(const b200sam_encoder_config* cfg, const void* const* weights, int n_weights,
                           b200sam_encoder** out);
B200SAM_API void b200sam_encoder_destroy(b200sam_encoder* enc);
/* image: [batch,3,h,w] uint8 (is_u8=1) or float32, un-normalised, long side <= 1024;
 * embedding_out: [batch,256,64,64] float32 == SamPredictor.features */
B200SAM_API int b200sam_encoder_forward(const b200sam_encoder* enc, const void* image, int is_u8, int batch, int h, int w,
                            const float* pixel_mean3_host, const float* pixel_std3_host, float* embedding_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- prompt extraction
 * Replaces PromptExtractor._extract_seeds / _extract_box / masked_non_overlapping_label_areas
 * (segment_anything/utils/prompt_utils.py:34-67).  masks: [n_img,C,H,W] bool bytes.
 * seeds: [n_img,C,2] (x,y); boxes: [n_img,C,4] (xmin,ymin,xmax,ymax); has_seed/has_box: [n_img,C] (0/1). */
B200SAM_API size_t b200sam_prompt_extract_scratch_bytes(int n_img, int n_classes);
B200SAM_API int b200sam_prompt_extract(const uint8_t* masks, int n_img, int n_classes, int H, int W, int32_t* seeds,
                           int32_t* boxes, uint8_t* has_seed, uint8_t* has_box, void* scratch, void* stream);

/* ---------------------------------------------------------------- prompt encoder + mask decoder
 * Replaces PromptEncoder.forward + MaskDecoder.forward (modeling/prompt_encoder.py:128-168,
 * modeling/mask_decoder.py:71-149, modeling/transformer.py:62-240) for ALL prompts of one image at once,
 * i.e. the per-class loop of SAMSegRefiner.refine (utils/seg_refinement.py:105-109) and
 * SAMMaskDecoderHead.predict_mask (segment_anything/sam_mask_decoder_head.py:79-96). */
typedef struct b200sam_decoder b200sam_decoder;
B200SAM_API int b200sam_decoder_weight_count(void);
/* "state_dict key[|packing]" of weight slot i (all float32); packing: cat4 | convT | repeat4 */
B200SAM_API const char* b200sam_decoder_weight_name(int i);
B200SAM_API size_t b200sam_decoder_workspace_bytes(int n_prompts, int n_points);
B200SAM_API int b200sam_decoder_create(const void* const* weights, int n_weights, b200sam_decoder** out, void* stream);
B200SAM_API void b200sam_decoder_destroy(b200sam_decoder* dec);
/* copy the dense positional encoding, token-major [4096,256], into out (PromptEncoder.get_dense_pe,
 * prompt_encoder.py:62-71; NCHW view = out.view(64,64,256).permute(2,0,1)) */
B200SAM_API int b200sam_decoder_copy_dense_pe(const b200sam_decoder* dec, float* out, void* stream);
/* embedding: [256,64,64]; coords: [n_prompts,n_points,2] (x,y) in the 1024 input frame; labels:
 * [n_prompts,n_points] with -1 padding point, 0 negative, 1 positive, 2/3 box corners; mask_prev: optional
 * [n_prompts,256,256] logits; outputs low_res: [n_prompts,(multimask?3:1),256,256], iou: [n_prompts,(3|1)] */
B200SAM_API int b200sam_decode(const b200sam_decoder* dec, const float* embedding, int n_prompts, int n_points,
                   const float* coords, const int32_t* labels, const float* mask_prev, int multimask,
                   float* low_res_out, float* iou_out, void* workspace, size_t workspace_bytes, void* stream);
/* The same for the prompts of SEVERAL images in one pass, i.e. additionally the per-image loop of
 * scripts/save_refined_segmentations.py:60-80.  embeddings: [n_images,256,64,64]; image_of: [n_prompts] image
 * index of every prompt (may be NULL when n_images == 1).  Prompts may carry different numbers of points:
 * n_points is the slot count, unused TRAILING slots have label -2 and are masked out of every attention, so each
 * prompt's result equals what it gets in a call of its own. */
B200SAM_API size_t b200sam_decoder_workspace_bytes_batch(int n_images, int n_prompts, int n_points);
B200SAM_API int b200sam_decode_batch(const b200sam_decoder* dec, const float* embeddings, int n_images,
                         const int32_t* image_of, int n_prompts, int n_points, const float* coords,
                         const int32_t* labels, const float* mask_prev, int multimask, float* low_res_out,
                         float* iou_out, void* workspace, size_t workspace_bytes, void* stream);

/* The two halves on their own, for callers of the reference's standalone module API:
 * b200sam_prompt_encode = PromptEncoder.forward (modeling/prompt_encoder.py:128-168): sparse_out [n_prompts,n_points,256]
 * (points, pad point and box corners already assembled by the caller as coords / labels like for b200sam_decode) and
 * dense_tok_out [n_prompts,4096,256] = the dense embedding TOKEN-MAJOR (NCHW view: .view(n,64,64,256).permute(0,3,1,2));
 * tokens_tmp [n_prompts,5+n_points,256] and ntok_tmp [n_prompts] are scratch.
 * b200sam_decode_embedded = MaskDecoder.forward (modeling/mask_decoder.py:71-110) on caller-supplied sparse
 * [n_prompts,n_sparse,256] and token-major dense [n_prompts,4096,256] embeddings; the image positional encoding is the
 * model's own dense PE (b200sam_decoder_copy_dense_pe). */
B200SAM_API int b200sam_prompt_encode(const b200sam_decoder* dec, const float* coords, const int32_t* labels, int n_prompts,
                          int n_points, const float* mask_prev, float* tokens_tmp, int32_t* ntok_tmp, float* sparse_out,
                          float* dense_tok_out, void* stream);
B200SAM_API int b200sam_decode_embedded(const b200sam_decoder* dec, const float* embeddings, int n_images,
                            const int32_t* image_of, int n_prompts, int n_sparse, const float* sparse,
                            const float* dense_tok, int multimask, float* low_res_out, float* iou_out, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- mask post-processing
 * Replaces postprocess_masks + threshold (segment_anything/sam_mask_decoder_head.py:99-135,
 * modeling/sam.py:133-162) and the nearest-exact resample of utils/seg_refinement.py:111.
 * low_res: [n,low,low]; any of mask_out [n,out_h,out_w] (0/1 bytes), logits_out [n,out_h,out_w] float32,
 * small_out [n,small_h,small_w] may be NULL. */
B200SAM_API int b200sam_upscale_threshold(const float* low_res, int n, int low, int img_size, int in_h, int in_w, int out_h,
                              int out_w, float threshold, uint8_t* mask_out, float* logits_out, uint8_t* small_out,
                              int small_h, int small_w, void* stream);

/* ---------------------------------------------------------------- connected-component pre-processing
 * Replaces remove_all_but_one_connected_component (utils/segmentation_preprocessing.py:7-52, called from
 * SegEnhance.enhance, utils/seg_refinement.py:64-72) for n_planes = images x classes probability planes at once:
 * out = prob * (winning 8-connected component of prob > threshold); by_area = 1: 'largest', 0: 'highest_probability'.
 * Labels follow kornia.contrib.connected_components after convergence (component label = its largest pixel index), so
 * the component labelled 0 - a lone pixel at index 0 of the FIRST plane of a reference call - is background there and
 * here; planes_per_call = planes of one reference call (C when n_planes = images x C; <= 0: all planes are one call). */
B200SAM_API size_t b200sam_ccl_scratch_bytes(int n_planes, int H, int W);
B200SAM_API int b200sam_ccl_select(const float* prob, int n_planes, int planes_per_call, int H, int W, float threshold,
                       int by_area, float* out, void* scratch, void* stream);
/* Flat grey-scale dilation (dilate = 1) / erosion (0) with a 0/1 structuring element se [kh,kw] (device) anchored at
 * (origin_y, origin_x); replaces kornia.morphology.dilation / erosion as used by SegEnhance (seg_refinement.py:44-62). */
B200SAM_API int b200sam_morph_flat(const float* in, int n_planes, int H, int W, const uint8_t* se, int kh, int kw,
                       int origin_y, int origin_x, int dilate, float* out, void* stream);

/* ---------------------------------------------------------------- mask statistics
 * Replace calculate_stability_score (segment_anything/utils/amg.py:154-176; logits [n,H,W] float32 -> score [n] =
 * count(x > threshold_hi) / count(x > threshold_lo) with threshold_hi/lo = mask_threshold +/- threshold_offset formed
 * by the caller, int32 counts divided in fp32) and batched_mask_to_box (amg.py:303-346;
 * masks [n,H,W] 0/1 bytes -> int64 XYXY boxes [n,4], zeros for an empty mask).  scratch: int32 [n,2] / [n,4]. */
B200SAM_API int b200sam_stability_score(const float* logits, int n, int H, int W, float threshold_hi, float threshold_lo,
                            float* score_out, int32_t* scratch, void* stream);
B200SAM_API int b200sam_mask_to_box(const uint8_t* masks, int n, int H, int W, int64_t* boxes_out, int32_t* scratch, void* stream);

/* ---------------------------------------------------------------- image ingest
 * Replaces ResizeLongestSide.apply_image (segment_anything/utils/transforms.py:26-31: torchvision `resize` of a PIL
 * image = Pillow's antialiased bilinear ImagingResample in 22-bit fixed point) as called from SamPredictor.set_image
 * (predictor.py:54-57) and scripts/generate_img_embeddings.py:43-45.  Bit-exact with Pillow.
 * b200sam_resize_coeffs_host is a pure HOST function: it fills bounds_host [out_size,2] (first tap, tap count) and
 * kk_host [out_size, b200sam_resize_ksize(in,out)] (fixed-point weights); the caller uploads them once per size pair.
 * b200sam_resize_u8: image [H,W,C] uint8 -> out [out_h,out_w,C] (out_chw = 0) or [C,out_h,out_w] (out_chw = 1, the
 * layout b200sam_encoder_forward reads).  A NULL bounds table skips that pass (size unchanged on that axis);
 * tmp [H,out_w,C] is needed when the horizontal pass is followed by the vertical one. */
B200SAM_API int b200sam_resize_ksize(int in_size, int out_size);
B200SAM_API int b200sam_resize_coeffs_host(int in_size, int out_size, int32_t* bounds_host, int32_t* kk_host);
B200SAM_API int b200sam_resize_u8(const uint8_t* image, int H, int W, int C, const int32_t* xbounds, const int32_t* xkk,
                      int xksize, const int32_t* ybounds, const int32_t* ykk, int yksize, int out_h, int out_w,
                      uint8_t* tmp, uint8_t* out, int out_chw, void* stream);

/* U-Net ingest: replaces cv2.resize(grey uint8, (W, H), interpolation=cv2.INTER_LINEAR) + `.float() / 255` +
 * `(img - IMG_MEAN) / IMG_STD` of scripts/save_refined_segmentations.py:62-67 (OpenCV's 11-bit fixed-point uint8 path,
 * bit-exact with cv2).  b200sam_cvresize_coeffs_host is a pure HOST function: idx2_host [out_size,2] = the two taps,
 * w2_host [out_size,2] = their weights; clamp_weights = 1 for the x axis, 0 for the y axis (OpenCV clamps the horizontal
 * taps with weight (1,0) but only the ROW indices vertically).  image: [n,H,W] uint8; out_u8 [n,out_h,out_w] and / or
 * out_norm [n,out_h,out_w] float32 = ((u8 / 255) - mean) / std; either may be NULL. */
B200SAM_API int b200sam_cvresize_coeffs_host(int in_size, int out_size, int clamp_weights, int32_t* idx2_host, int32_t* w2_host);
B200SAM_API int b200sam_cvresize_linear_u8(const uint8_t* image, int n, int H, int W, const int32_t* xidx, const int32_t* xw,
                               const int32_t* yidx, const int32_t* yw, int out_h, int out_w, uint8_t* out_u8,
                               float* out_norm, float mean, float std, void* stream);

/* MedSAM ingest (the reference's default sam_type): replaces cv2.resize(img, (1024, 1024), interpolation=cv2.INTER_CUBIC)
 * + min-max normalisation of scripts/generate_img_embeddings.py:49-62 (OpenCV's own uint8 cubic path, i.e. without Intel
 * IPP; the result feeds b200sam_encoder_forward as float32 with mean 0 / std 1, bypassing Sam.preprocess like the
 * reference).  gray: [H,W] uint8 (its RGB replication has three identical channels); tables from the HOST function
 * (idx4 [size,4] clamped taps, w4 [size,4]); resized_u8 [size,size] and minmax [2] are outputs / scratch;
 * out3: [3,size,size] float32 in [0,1]. */
B200SAM_API int b200sam_cvresize_cubic_coeffs_host(int in_size, int out_size, int32_t* idx4_host, int32_t* w4_host);
B200SAM_API int b200sam_medsam_preprocess(const uint8_t* gray, int H, int W, const int32_t* xidx, const int32_t* xw,
                              const int32_t* yidx, const int32_t* yw, int size, uint8_t* resized_u8, int32_t* minmax,
                              float* out3, void* stream);

/* ---------------------------------------------------------------- U-Net inference
 * Replaces UNet.forward (custom_arcitecture/classic_u_net.py:81-119, bilinear = False) as called from
 * scripts/save_refined_segmentations.py:67-69 (+ the torch.sigmoid that follows).  Weight table like the SAM handles:
 * "state_dict key|packing" with packing conv3x3_tap (fp32 [Cout, Kp], column (ky*3+kx)*Cin + c, zero padded to
 * Kp = b200sam_unet_conv_kp(Cin)), convT, repeat4, pad_rows8 / pad8 (rows / elements zero padded to a multiple of 8).
 * image: [batch,1,H,W] float32, already normalised, H and W multiples of 16; outputs [batch,n_classes,H,W]. */
typedef struct b200sam_unet b200sam_unet;
B200SAM_API int b200sam_unet_weight_count(void);
B200SAM_API const char* b200sam_unet_weight_name(int i);
B200SAM_API int b200sam_unet_conv_kp(int cin);
B200SAM_API int b200sam_unet_create(int n_channels, int n_classes, int n_last_channel, const void* const* weights,
                        int n_weights, b200sam_unet** out, void* stream);
B200SAM_API void b200sam_unet_destroy(b200sam_unet* u);
B200SAM_API size_t b200sam_unet_workspace_bytes(const b200sam_unet* u, int batch, int H, int W);
B200SAM_API int b200sam_unet_forward(const b200sam_unet* u, const float* image, int batch, int H, int W, float* logits_out,
                         float* probs_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------- building blocks (exposed for parity tests)
 * D[M,N] = A[M,K] W[N,K]^T (+bias) (+GELU) (+residual[row % res_row_mod]); 16-bit operands (bf16 / fp16 entry point),
 * fp32 accumulate (tcgen05); out_16bit = 1: output in the operands' format, 0: fp32 (with optional fp32 residual). */
B200SAM_API int b200sam_gemm_bf16(const void* A, const void* W, void* out, const float* bias, const float* residual, int M,
                      int N, int K, int lda, int ldb, int ldo, int ldr, int res_row_mod, int gelu, int out_bf16,
                      int max_ctas, void* stream);
B200SAM_API int b200sam_gemm_f16(const void* A, const void* W, void* out, const float* bias, const float* residual, int M,
                     int N, int K, int lda, int ldb, int ldo, int ldr, int res_row_mod, int gelu, int out_f16,
                     int max_ctas, void* stream);
/* The two halves of a LayerNorm folded into the GEMMs around it (Block.forward, image_encoder.py:166-182):
 * _ln_residual: out = A W^T + bias + residual (fp32 [M,N], may alias residual), out16 = its 16-bit copy,
 *               rowstat_out [M, N/64, 2] = per-row partial (sum, sum of squares) of out per 64-column part (N % 128 == 0);
 * _ln_folded:   out16 [M,N] = act( rstd * (A W_folded^T - mean * colsum) + bias_folded ) with (mean, rstd) of every row
 *               of A's fp32 original from rowstat_in [M, nparts, 2] (K elements per row, eps inside the sqrt). */
B200SAM_API int b200sam_gemm_ln_residual(const void* A, const void* W, const float* bias, const float* residual, float* out,
                             void* out16, float* rowstat_out, int M, int N, int K, int operand_format, void* stream);
B200SAM_API int b200sam_gemm_ln_folded(const void* A, const void* W_folded, const float* bias_folded, const float* colsum,
                           const float* rowstat_in, int nparts, float eps, void* out16, int M, int N, int K, int gelu,
                           int operand_format, void* stream);
/* out_kind: 0 = fp32, 1 = bf16, 2 = fp16 */
B200SAM_API int b200sam_layernorm(const float* x, const float* gamma, const float* beta, float eps, int M, int D, void* y,
                      int out_kind, void* stream);
/* qkv: [B*4096, 3*heads*hd] 16-bit; out: [B*4096, heads*hd] 16-bit; global_attn: 0 = 14x14 windows, 1 = global (both on
 * tcgen05 / TMEM); operand_format as in b200sam_encoder_config */
B200SAM_API int b200sam_encoder_attention(const void* qkv, const void* qkv_bias16, const void* rel_h16, const void* rel_w16,
                              void* out, int batch, int heads, int hd, int global_attn, int operand_format,
                              void* stream);
B200SAM_API int b200sam_preprocess_patchify(const void* image, int is_u8, int batch, int h, int w, const float* mean3_host,
                                const float* std3_host, void* out16, int operand_format, void* stream);
/* The encoder's plain linears run on the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles over the two SMs of a TPC)
 * unless B200SAM_GEMM_PAIR=0; b200sam_set_gemm_pair overrides the environment at run time (-1: environment, 0: single-CTA
 * kernel, != 0: pair kernel) for A/B measurements and parity tests.  _max_clusters = co-resident CTA pairs on the current
 * device (SMs whose TPC partner is fused off cannot host one). */
B200SAM_API int b200sam_set_gemm_pair(int mode);
B200SAM_API int b200sam_gemm_pair_max_clusters(void);
/* In-run kernel timing for the bench's roofline block: between _start and _stop every launch of the tcgen05 GEMM
 * (kind 0; work = 2 M N K; dims = M, N, K), the windowed (kind 1) and the global (kind 2) attention kernel (work =
 * algorithmic FLOPs; dims = batch, heads, head dim) is bracketed by CUDA events on its own stream.  _stop synchronises
 * on the recorded events and fills the HOST arrays (any may be NULL) with up to `capacity` records. */
B200SAM_API int b200sam_timing_start(int capacity);
B200SAM_API int b200sam_timing_stop(int* kinds_host, double* work_host, int* dims3_host, float* ms_host, int capacity,
                        int* n_out_host);
B200SAM_API int b200sam_linear_f32(const float* A, const float* A2, int a2_row_mod, const float* W, const float* bias,
                       const float* residual, float* out, int M, int N, int K, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200SAM_H */
