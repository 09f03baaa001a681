// Host-side orchestration of the ViT image encoder (reference: ImageEncoderViT.forward,
// segment_anything/modeling/image_encoder.py:106-116; Block.forward :166-182).  Launch sequence per block:
//   qkv GEMM -> fused (window|global) attention -> proj GEMM(+residual) -> lin1 GEMM(+GELU) -> lin2 GEMM(+residual)
// with norm1 / norm2 folded into the GEMMs around them (ENC_FLAG_LN_FUSED: the residual GEMMs also emit a 16-bit copy
// of x and per-row partial sums, the qkv / lin1 epilogues apply (mean, rstd)); without the flag LN1 / LN2 are separate
// launches.  The residual stream stays fp32 in HBM; MMA operands are 16-bit (fp16 by default, bf16 selectable).
#include "encoder.h"
#include <cstdlib>
#include <string>

namespace b200sam {

namespace {

enum : int { G_PATCH_W = 0, G_PATCH_B, G_POS, G_BLOCK0 };
enum : int { B_N1W = 0, B_N1B, B_QKVW, B_QKVB, B_QKVS, B_QKVB16, B_RELH, B_RELW, B_PROJW, B_PROJB, B_N2W, B_N2B, B_L1W,
             B_L1B, B_L1S, B_L2W, B_L2B, B_STRIDE };
enum : int { N_C0W = 0, N_L1W, N_L1B, N_C2W, N_L3W, N_L3B, N_COUNT };

inline bool ln_fused(const EncoderConfig& c) { return (c.flags & ENC_FLAG_LN_FUSED) != 0; }

// "state_dict key|packing[|norm prefix]".  op16 = the encoder's 16-bit operand format (cfg.operand_format).
// LayerNorm folding (flag ENC_FLAG_LN_FUSED): the linear that follows a LayerNorm carries gamma in its weights and beta
// in its bias;  fold_w = op16(gamma * W),  fold_s[n] = sum_k fold_w[n,k] (fp32),  fold_c = beta . W^T + b (fp32).
std::string name_of(const EncoderConfig& c, int i) {
  const std::string ie = "image_encoder.";
  const bool fused = ln_fused(c);
  if (i == G_PATCH_W) return ie + "patch_embed.proj.weight|op16_flat";
  if (i == G_PATCH_B) return ie + "patch_embed.proj.bias|f32";
  if (i == G_POS) return ie + "pos_embed|f32_tokens";
  const int nb = c.depth * B_STRIDE;
  if (i < G_BLOCK0 + nb) {
    const int b = (i - G_BLOCK0) / B_STRIDE, k = (i - G_BLOCK0) % B_STRIDE;
    const std::string p = ie + "blocks." + std::to_string(b) + ".";
    switch (k) {
      case B_N1W: return p + "norm1.weight|f32";
      case B_N1B: return p + "norm1.bias|f32";
      case B_QKVW: return fused ? p + "attn.qkv|fold_w|" + p + "norm1" : p + "attn.qkv.weight|op16";
      case B_QKVB: return fused ? p + "attn.qkv|fold_c|" + p + "norm1" : p + "attn.qkv.bias|f32";
      case B_QKVS: return fused ? p + "attn.qkv|fold_s|" + p + "norm1" : p + "attn.qkv.bias|none";
      case B_QKVB16: return p + "attn.qkv.bias|op16";
      case B_RELH: return p + "attn.rel_pos_h|op16";
      case B_RELW: return p + "attn.rel_pos_w|op16";
      case B_PROJW: return p + "attn.proj.weight|op16";
      case B_PROJB: return p + "attn.proj.bias|f32";
      case B_N2W: return p + "norm2.weight|f32";
      case B_N2B: return p + "norm2.bias|f32";
      case B_L1W: return fused ? p + "mlp.lin1|fold_w|" + p + "norm2" : p + "mlp.lin1.weight|op16";
      case B_L1B: return fused ? p + "mlp.lin1|fold_c|" + p + "norm2" : p + "mlp.lin1.bias|f32";
      case B_L1S: return fused ? p + "mlp.lin1|fold_s|" + p + "norm2" : p + "mlp.lin1.bias|none";
      case B_L2W: return p + "mlp.lin2.weight|op16";
      case B_L2B: return p + "mlp.lin2.bias|f32";
    }
  }
  const int k = i - G_BLOCK0 - nb;
  switch (k) {
    case N_C0W: return ie + "neck.0.weight|op16_flat";
    case N_L1W: return ie + "neck.1.weight|f32";
    case N_L1B: return ie + "neck.1.bias|f32";
    case N_C2W: return ie + "neck.2.weight|op16_tap";
    case N_L3W: return ie + "neck.3.weight|f32";
    case N_L3B: return ie + "neck.3.bias|f32";
  }
  return "";
}

inline size_t align_up(size_t x) { return (x + 1023) & ~static_cast<size_t>(1023); }

struct Workspace {
  float* x;             // [M, D] fp32 residual stream
  __nv_bfloat16* xn;    // [M, D] 16-bit LN output / 16-bit copy of x  (neck: fp32 [M,256] conv3x3 output)
  __nv_bfloat16* qkv;   // [M, 3D]                                (neck: fp32 [M,256] conv1x1 output)
  __nv_bfloat16* att;   // [M, D] attention output                (neck: 16-bit [M,256] LN output)
  __nv_bfloat16* h;     // [M, 4D] MLP hidden; also patch im2col [M,768] and neck im2col [M,2304]
  float* stat;          // [M, D/64, 2] per-row partial (sum, sum of squares) of x (LayerNorm folding)
  size_t total;
};

Workspace carve(uint8_t* base, const EncoderConfig& c, int B) {
  Workspace w;
  const size_t M = static_cast<size_t>(B) * 4096, D = c.embed_dim;
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base + off; off += align_up(bytes); return p; };
  w.x = reinterpret_cast<float*>(take(M * D * 4));
  w.xn = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
  w.qkv = reinterpret_cast<__nv_bfloat16*>(take(M * 3 * D * 2));
  w.att = reinterpret_cast<__nv_bfloat16*>(take(M * D * 2));
  w.h = reinterpret_cast<__nv_bfloat16*>(take(M * 4 * D * 2));
  w.stat = reinterpret_cast<float*>(take(M * (D / 64) * 2 * 4));
  w.total = off;
  return w;
}

#define TRY(x) do { if (int _rc = (x)) return _rc; } while (0)

}  // namespace

int encoder_weight_count(const EncoderConfig& c) { return G_BLOCK0 + c.depth * B_STRIDE + N_COUNT; }

const char* encoder_weight_name(const EncoderConfig& c, int i) {
  static thread_local std::string buf;
  if (i < 0 || i >= encoder_weight_count(c)) return nullptr;
  buf = name_of(c, i);
  return buf.c_str();
}

size_t encoder_workspace_bytes(const EncoderConfig& c, int B) {
  if (B <= 0) return 0;
  return carve(nullptr, c, B).total + 1024;
}

int encoder_create(const EncoderConfig& c, const void* const* weights, int n, Encoder** out) {
  B200SAM_REQUIRE(c.embed_dim > 0 && c.embed_dim % 128 == 0 && c.num_heads > 0 && c.embed_dim % c.num_heads == 0,
                  "encoder_create: bad embed_dim=%d num_heads=%d", c.embed_dim, c.num_heads);
  const int hd = c.embed_dim / c.num_heads;
  B200SAM_REQUIRE(hd == 64 || hd == 80, "encoder_create: head dim %d unsupported (64 or 80)", hd);
  B200SAM_REQUIRE(c.depth > 0 && c.depth <= 32, "encoder_create: depth %d unsupported (1..32)", c.depth);
  B200SAM_REQUIRE(c.out_chans == 256, "encoder_create: out_chans must be 256, got %d", c.out_chans);
  B200SAM_REQUIRE(c.operand_format == 0 || c.operand_format == 1,
                  "encoder_create: operand_format %d unknown (0 = bf16, 1 = fp16)", c.operand_format);
  B200SAM_REQUIRE((c.flags & ~ENC_FLAG_LN_FUSED) == 0, "encoder_create: unknown flags 0x%x", c.flags);
  B200SAM_REQUIRE(n == encoder_weight_count(c), "encoder_create: expected %d weight pointers, got %d",
                  encoder_weight_count(c), n);
  for (int i = 0; i < n; ++i)
    B200SAM_REQUIRE(weights[i] != nullptr, "encoder_create: weight %d (%s) is null", i, name_of(c, i).c_str());
  Encoder* e = new Encoder();
  e->cfg = c;
  e->w.assign(weights, weights + n);
  *out = e;
  return 0;
}

void encoder_destroy(Encoder* e) { delete e; }

int encoder_forward(const Encoder* e, const void* img, int is_u8, int B, int h, int w, const float* mean3,
                    const float* std3, float* out, void* workspace, size_t workspace_bytes, cudaStream_t s) {
  const EncoderConfig& c = e->cfg;
  B200SAM_REQUIRE(B > 0 && img != nullptr && out != nullptr, "encoder_forward: bad arguments");
  B200SAM_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
                  "encoder_forward: workspace must be non-null and 1024-byte aligned");
  Workspace ws = carve(reinterpret_cast<uint8_t*>(workspace), c, B);
  B200SAM_REQUIRE(ws.total <= workspace_bytes, "encoder_forward: workspace too small (%zu < %zu)", workspace_bytes,
                  ws.total);
  const int D = c.embed_dim, M = B * 4096;
  const int f16 = c.operand_format == 1;
  const int k16 = f16 ? 2 : 1;  // out_kind of the 16-bit tensors
  const bool fused = ln_fused(c);
  const int nparts = D / 64;  // row statistics come in 64-column parts (gemm_epilogue.cuh)
  const void* const* W = e->w.data();
  auto F = [&](int i) { return reinterpret_cast<const float*>(W[i]); };
  auto H = [&](int i) { return reinterpret_cast<const __nv_bfloat16*>(W[i]); };

  // D[M,N] = A[M,K] W[N,K]^T with the encoder's operand format
  auto base = [&](const __nv_bfloat16* A, int wi, void* o, const float* bias, int N, int K, int dir) {
    GemmArgs g;
    g.A = A; g.B = H(wi); g.out = o; g.bias = bias; g.residual = nullptr;
    g.M = M; g.N = N; g.K = K; g.lda = K; g.ldb = K; g.ldo = N; g.ldr = 0; g.res_row_mod = 0;
    g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.reverse_m = dir; g.op_f16 = f16;
    return g;
  };
  // residual-stream producer: x = A W^T + b + residual (fp32, in place); with LayerNorm folding also the 16-bit copy of x
  // and its per-row partial sums for the next linear
  auto gemm_residual = [&](const __nv_bfloat16* A, int wi, int bi, const float* res, int res_mod, int K, int dir) {
    GemmArgs g = base(A, wi, ws.x, F(bi), D, K, dir);
    g.residual = res; g.ldr = D; g.res_row_mod = res_mod;
    if (fused) { g.xh = ws.xn; g.rowstat_out = ws.stat; }
    return gemm_bf16_tn(g, s);
  };
  // 16-bit consumer of LN(x): plain (A = LN output) or folded (A = 16-bit x, statistics applied in the epilogue)
  auto gemm_after_ln = [&](int wi, int bi, int si, void* o, int N, int gelu, int dir) {
    GemmArgs g = base(ws.xn, wi, o, F(bi), N, D, dir);
    g.out_kind = k16; g.gelu = gelu;
    if (fused) { g.rowstat_in = ws.stat; g.colsum = F(si); g.nparts_in = nparts; g.ln_dim = D; g.ln_eps = 1e-6f; }
    return gemm_bf16_tn(g, s);
  };

  // patch embedding (+bias +abs pos embed) : image_encoder.py:107-109
  TRY(preprocess_patchify(img, is_u8, B, h, w, mean3, std3, ws.h, f16, s));
  TRY(gemm_residual(ws.h, G_PATCH_W, G_PATCH_B, F(G_POS), 4096, 768, 0));

  AttnArgs at;
  at.qkv = ws.qkv; at.out = ws.att; at.B = B; at.heads = c.num_heads; at.hd = D / c.num_heads; at.f16 = f16;
  // Traversal direction ("boustrophedon"): the residual stream (4*D bytes per token, 168 MB at batch 8) and the MLP
  // hidden (8*D) are larger than the 126 MB L2, so a consumer that walks the rows in the SAME order as its producer finds
  // its first rows already evicted.  Each LayerNorm / GEMM below therefore starts at the end its producer finished at
  // (results are identical; only the order of the row blocks changes).  B200SAM_FORWARD_ONLY=1 disables it (A/B timing).
  static const bool kBoustrophedon = std::getenv("B200SAM_FORWARD_ONLY") == nullptr;
  int x_dir = 0;  // direction in which ws.x / its 16-bit copy was last written (patch embedding: forward)
  for (int b = 0; b < c.depth; ++b) {
    const int o = G_BLOCK0 + b * B_STRIDE;
    int qkv_dir, proj_dir, l1_dir, l2_dir;
    if (fused) {
      // lin2(d) -> qkv(!d) -> attention(d) -> proj(!d) -> lin1(d) -> lin2(!d): every kernel starts where its producer ended
      qkv_dir = kBoustrophedon ? !x_dir : 0;
      proj_dir = qkv_dir;
      l1_dir = kBoustrophedon ? !proj_dir : 0;
      l2_dir = kBoustrophedon ? !l1_dir : 0;
    } else {
      const int ln1_dir = kBoustrophedon ? !x_dir : 0;
      qkv_dir = kBoustrophedon ? !ln1_dir : 0;
      TRY(layernorm_rows(ws.x, F(o + B_N1W), F(o + B_N1B), 1e-6f, M, D, ws.xn, k16, s, ln1_dir));
      proj_dir = 0;  // proj runs forward (its A operand, the 2*D-byte attention output, fits the L2 either way)
      l1_dir = 0; l2_dir = kBoustrophedon ? 1 : 0;  // LN2 reverse, lin1 forward, lin2 reverse
    }
    TRY(gemm_after_ln(o + B_QKVW, o + B_QKVB, o + B_QKVS, ws.qkv, 3 * D, 0, qkv_dir));
    at.qkv_bias = H(o + B_QKVB16); at.rel_h = H(o + B_RELH); at.rel_w = H(o + B_RELW);
    at.reverse = kBoustrophedon ? !qkv_dir : 0;
    if ((c.global_mask_lo >> b) & 1) TRY(global_attention_tc(at, s));
    else TRY(window_attention_tc(at, s));
    TRY(gemm_residual(ws.att, o + B_PROJW, o + B_PROJB, ws.x, 0, D, proj_dir));
    if (!fused)
      TRY(layernorm_rows(ws.x, F(o + B_N2W), F(o + B_N2B), 1e-6f, M, D, ws.xn, k16, s, kBoustrophedon ? 1 : 0));
    TRY(gemm_after_ln(o + B_L1W, o + B_L1B, o + B_L1S, ws.h, 4 * D, 1, l1_dir));
    {
      GemmArgs g = base(ws.h, o + B_L2W, ws.x, F(o + B_L2B), D, 4 * D, l2_dir);
      g.residual = ws.x; g.ldr = D;
      if (fused) { g.xh = ws.xn; g.rowstat_out = ws.stat; }
      TRY(gemm_bf16_tn(g, s));
    }
    x_dir = l2_dir;
  }

  // neck: conv1x1 -> LayerNorm2d -> conv3x3(pad 1) -> LayerNorm2d (image_encoder.py:88-104), NCHW fp32 out
  const int o = G_BLOCK0 + c.depth * B_STRIDE;
  const int C = c.out_chans;
  float* n0 = reinterpret_cast<float*>(ws.qkv);
  float* n2 = reinterpret_cast<float*>(ws.xn);
  const __nv_bfloat16* xh = ws.xn;  // folded: the last lin2 already wrote the 16-bit copy of x
  if (!fused) {
    TRY(f32_to_op16(ws.x, ws.att, static_cast<size_t>(M) * D, f16, s));
    xh = ws.att;
  }
  {
    GemmArgs g = base(xh, o + N_C0W, n0, nullptr, C, D, 0);
    TRY(gemm_bf16_tn(g, s));
  }
  TRY(layernorm_rows(n0, F(o + N_L1W), F(o + N_L1B), 1e-6f, M, C, ws.att, k16, s));
  TRY(im2col3x3_tokens(ws.att, B, C, ws.h, s));
  {
    GemmArgs g = base(ws.h, o + N_C2W, n2, nullptr, C, 9 * C, 0);
    TRY(gemm_bf16_tn(g, s));
  }
  TRY(layernorm_to_nchw(n2, F(o + N_L3W), F(o + N_L3B), 1e-6f, B, C, out, s));
  return 0;
}

}  // namespace b200sam
