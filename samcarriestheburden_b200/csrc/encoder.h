// Encoder context: ViT hyper-parameters + borrowed, pre-packed weight pointers.
#pragma once
#include <vector>
#include "kernels.h"

namespace b200sam {

struct EncoderConfig {
  int embed_dim;     // 768 / 1024 / 1280
  int depth;         // 12 / 24 / 32
  int num_heads;     // 12 / 16 / 16
  int global_mask_lo;  // bit i set -> block i uses global attention (blocks 0..31)
  int out_chans;     // 256
  int operand_format;  // 16-bit MMA operand format: 0 = bf16, 1 = fp16 (default of the Python layer)
  int flags;           // ENC_FLAG_*
};
enum : int { ENC_FLAG_LN_FUSED = 1 };  // norm1 / norm2 folded into the GEMMs around them (no LayerNorm launches)

struct Encoder {
  EncoderConfig cfg;
  std::vector<const void*> w;  // in encoder_weight_name() order
};

int encoder_weight_count(const EncoderConfig& c);
// "state_dict key|packing[|norm prefix]": packing in {f32, op16 (the 16-bit operand format), op16_flat (conv weight
// flattened to [out, -1]), op16_tap (3x3 conv as [out, (ky*3+kx)*Cin + c]), f32_tokens (pos_embed as [4096, D]),
// fold_w / fold_s / fold_c (LayerNorm folded into the linear `key`, see encoder.cu), none (unused slot)}
const char* encoder_weight_name(const EncoderConfig& c, int i);
size_t encoder_workspace_bytes(const EncoderConfig& c, int B);
int encoder_create(const EncoderConfig& c, const void* const* weights, int n, Encoder** out);
void encoder_destroy(Encoder* e);
// img: [B,3,h,w] uint8 or fp32 (ResizeLongestSide output, un-normalised); out: [B,out_chans,64,64] fp32
int encoder_forward(const Encoder* e, const void* img, int is_u8, int B, int h, int w, const float* mean3,
                    const float* std3, float* out, void* workspace, size_t workspace_bytes, cudaStream_t stream);

}  // namespace b200sam
