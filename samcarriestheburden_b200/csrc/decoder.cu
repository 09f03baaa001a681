// Host-side orchestration of the prompt encoder + mask decoder for ALL prompts of one or several images in
// one batched pass (reference loop being replaced: utils/seg_refinement.py:105-109 calling
// sam_mask_decoder_head.py:79-96 once per class with B=1, inside the per-image loop of
// scripts/save_refined_segmentations.py:60-80).  Prompts of different images may carry different numbers of
// sparse points (label -2 = absent slot): absent token slots are masked out as attention keys, so every prompt
// computes exactly what it would compute alone.  Pure launch sequencing: no allocation, no
// synchronisation; every buffer lives in the caller-provided workspace.
#include "decoder.h"
#include <string>
#include <vector>

namespace b200sam {

namespace {

const char* kAttnParts[8] = {"q_proj.weight", "q_proj.bias", "k_proj.weight", "k_proj.bias",
                             "v_proj.weight", "v_proj.bias", "out_proj.weight", "out_proj.bias"};

std::vector<std::string> build_names() {
  std::vector<std::string> n;
  const std::string pe = "prompt_encoder.";
  n.push_back(pe + "pe_layer.positional_encoding_gaussian_matrix");
  n.push_back(pe + "point_embeddings|cat4");  // [4,256] = cat(point_embeddings.{0..3}.weight)
  n.push_back(pe + "not_a_point_embed.weight");
  n.push_back(pe + "no_mask_embed.weight");
  const char* md[5] = {"0", "1", "3", "4", "6"};
  for (int i = 0; i < 5; ++i) {
    n.push_back(pe + "mask_downscaling." + md[i] + ".weight");
    n.push_back(pe + "mask_downscaling." + md[i] + ".bias");
  }
  const std::string dec = "mask_decoder.";
  n.push_back(dec + "iou_token.weight");
  n.push_back(dec + "mask_tokens.weight");
  for (int l = 0; l < 2; ++l) {
    const std::string L = dec + "transformer.layers." + std::to_string(l) + ".";
    for (int i = 0; i < 8; ++i) n.push_back(L + "self_attn." + kAttnParts[i]);
    n.push_back(L + "norm1.weight"); n.push_back(L + "norm1.bias");
    for (int i = 0; i < 8; ++i) n.push_back(L + "cross_attn_token_to_image." + kAttnParts[i]);
    n.push_back(L + "norm2.weight"); n.push_back(L + "norm2.bias");
    n.push_back(L + "mlp.lin1.weight"); n.push_back(L + "mlp.lin1.bias");
    n.push_back(L + "mlp.lin2.weight"); n.push_back(L + "mlp.lin2.bias");
    n.push_back(L + "norm3.weight"); n.push_back(L + "norm3.bias");
    n.push_back(L + "norm4.weight"); n.push_back(L + "norm4.bias");
    for (int i = 0; i < 8; ++i) n.push_back(L + "cross_attn_image_to_token." + kAttnParts[i]);
  }
  for (int i = 0; i < 8; ++i) n.push_back(dec + "transformer.final_attn_token_to_image." + kAttnParts[i]);
  n.push_back(dec + "transformer.norm_final_attn.weight");
  n.push_back(dec + "transformer.norm_final_attn.bias");
  n.push_back(dec + "output_upscaling.0.weight|convT");  // [(dy*2+dx)*Cout+co, ci]
  n.push_back(dec + "output_upscaling.0.bias|repeat4");
  n.push_back(dec + "output_upscaling.1.weight");
  n.push_back(dec + "output_upscaling.1.bias");
  n.push_back(dec + "output_upscaling.3.weight|convT");
  n.push_back(dec + "output_upscaling.3.bias|repeat4");
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 3; ++j) {
      const std::string M = dec + "output_hypernetworks_mlps." + std::to_string(i) + ".layers." + std::to_string(j);
      n.push_back(M + ".weight");
      n.push_back(M + ".bias");
    }
  for (int j = 0; j < 3; ++j) {
    const std::string M = dec + "iou_prediction_head.layers." + std::to_string(j);
    n.push_back(M + ".weight");
    n.push_back(M + ".bias");
  }
  return n;
}

const std::vector<std::string>& names() {
  static std::vector<std::string> n = build_names();
  return n;
}

// offsets into the table (must mirror build_names)
enum : int {
  W_GAUSS = 0, W_POINT4 = 1, W_NOT_A_POINT = 2, W_NO_MASK = 3, W_MASKDOWN = 4 /*10*/, W_IOU_TOKEN = 14,
  W_MASK_TOKENS = 15, W_LAYER0 = 16, LAYER_STRIDE = 36,
  L_SELF = 0, L_N1 = 8, L_T2I = 10, L_N2 = 18, L_MLP = 20, L_N3 = 24, L_N4 = 26, L_I2T = 28,
  W_FINAL = W_LAYER0 + 2 * LAYER_STRIDE, W_NF = W_FINAL + 8, W_UP = W_NF + 2 /*6*/, W_HYPER = W_UP + 6 /*24*/,
  W_IOUHEAD = W_HYPER + 24 /*6*/, W_COUNT = W_IOUHEAD + 6
};

__global__ void slice_iou_kernel(const float* __restrict__ iou4, float* __restrict__ out, int NB, int tok0, int ntok) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NB * ntok) out[i] = iou4[(i / ntok) * 4 + tok0 + (i % ntok)];
}

inline size_t align_up(size_t x) { return (x + 255) & ~static_cast<size_t>(255); }

struct Workspace {
  float *emb_tok, *keys, *kbuf, *vbuf, *qibuf;
  float *tokens, *queries, *tq, *tk, *tv, *ta, *th, *hyper, *iou4, *part;
  int* ntok;               // [NB] valid tokens per prompt (5 + present sparse points)
  __nv_bfloat16 *sa, *sb;  // hi/lo split operands [Mi, 512] bf16 each
  size_t total;
};

Workspace carve(uint8_t* base, int n_images, int NB, int T) {
  Workspace w;
  size_t off = 0;
  auto take = [&](size_t nfloat) {
    float* p = reinterpret_cast<float*>(base + off);
    off += align_up(nfloat * sizeof(float));
    return p;
  };
  // (image-side buffers also hold one block per IMAGE in the shared first layer: size them for max(prompts, images))
  const size_t Mi = static_cast<size_t>(NB > n_images ? NB : n_images) * 4096, Mt = static_cast<size_t>(NB) * T;
  w.emb_tok = take(static_cast<size_t>(n_images) * 4096 * 256);
  w.ntok = reinterpret_cast<int*>(take(static_cast<size_t>(NB)));
  w.keys = take(Mi * 256);
  w.kbuf = take(Mi * 128);
  w.vbuf = take(Mi * 128);
  w.qibuf = take(Mi * 128);
  w.tokens = take(Mt * 256);
  w.queries = take(Mt * 256);
  w.tq = take(Mt * 256);
  w.tk = take(Mt * 256);
  w.tv = take(Mt * 256);
  w.ta = take(Mt * 256);
  w.th = take(Mt * 2048);
  w.hyper = take(static_cast<size_t>(NB) * 128);
  w.iou4 = take(static_cast<size_t>(NB) * 4);
  w.part = take(static_cast<size_t>(NB) * 8 * ATTN_FEWQ_SPLITS * T * 18);
  w.sa = reinterpret_cast<__nv_bfloat16*>(take(Mi * 512 / 2));
  w.sb = reinterpret_cast<__nv_bfloat16*>(take(Mi * 512 / 2));
  w.total = off;
  return w;
}

int lin(const float* A, const float* A2, int a2mod, const float* W, const float* b, const float* res, float* out, int M,
        int N, int K, int act, cudaStream_t s) {
  LinearArgs p;
  p.A = A; p.A2 = A2; p.W = W; p.bias = b; p.residual = res; p.out = out;
  p.M = M; p.N = N; p.K = K; p.lda = K; p.lda2 = K; p.ldo = N; p.ldr = N; p.a2_row_mod = a2mod; p.act = act;
  return linear_f32(p, s);
}

// out[M,N] fp32 = [hi|lo|hi][M,3K] . W'[N,3K]^T + bias (+GELU) (+residual): tcgen05 GEMM on the split operands
// (A is stored as [hi|lo] with pitch 2K; the kernel's A loader wraps the third K segment back to hi)
int tc_lin(const __nv_bfloat16* As, const __nv_bfloat16* Ws, const float* b, const float* res, float* out, int M, int N,
           int K, int gelu, cudaStream_t s, int res_row_mod = 0) {
  GemmArgs g;
  g.A = As; g.B = Ws; g.out = out; g.bias = b; g.residual = res;
  g.M = M; g.N = N; g.K = 3 * K; g.lda = 2 * K; g.ldb = 3 * K; g.ldo = N; g.ldr = N; g.res_row_mod = res_row_mod;
  g.gelu = gelu; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 2 * K;
  return gemm_bf16_tn(g, s);
}

// several 128-wide projections of the same split operand in ONE launch: Ws = their weights back to back ([n_planes * 128,
// 3K]), bias = their biases back to back, tables = their per-token (row % 4096) residual tables one `tab_plane` apart,
// out0 = the first output [M, 128], the others one `out_plane` apart (elements)
int tc_lin_planes(const __nv_bfloat16* As, const __nv_bfloat16* Ws, const float* bias, const float* tables, size_t tab_plane,
                  float* out0, size_t out_plane, int n_planes, int M, int K, cudaStream_t s) {
  GemmArgs g;
  g.A = As; g.B = Ws; g.out = out0; g.bias = bias; g.residual = tables;
  g.M = M; g.N = 128 * n_planes; g.K = 3 * K; g.lda = 2 * K; g.ldb = 3 * K; g.ldo = 128; g.ldr = 128; g.res_row_mod = 4096;
  g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 2 * K;
  g.out_plane = static_cast<long long>(out_plane);
  g.res_plane = static_cast<long long>(tab_plane);
  return gemm_bf16_tn(g, s);
}

#define TRY(x) do { if (int _rc = (x)) return _rc; } while (0)

}  // namespace

int decoder_weight_count() { return W_COUNT; }
const char* decoder_weight_name(int i) {
  if (i < 0 || i >= static_cast<int>(names().size())) return nullptr;
  return names()[i].c_str();
}

size_t decoder_workspace_bytes(int n_images, int NB, int Np) {
  if (n_images <= 0 || NB <= 0 || Np < 0) return 0;
  return carve(nullptr, n_images, NB, 5 + Np).total + 256;
}

int decoder_create(const void* const* weights, int n, Decoder** out, cudaStream_t stream) {
  B200SAM_REQUIRE(n == W_COUNT && static_cast<int>(names().size()) == W_COUNT,
                  "decoder_create: expected %d weight pointers, got %d", W_COUNT, n);
  for (int i = 0; i < n; ++i) B200SAM_REQUIRE(weights[i] != nullptr, "decoder_create: weight %d (%s) is null", i,
                                              names()[i].c_str());
  Decoder* d = new Decoder();
  d->w.assign(reinterpret_cast<const float* const*>(weights), reinterpret_cast<const float* const*>(weights) + n);
  d->pe_tok = nullptr;
  if (cudaMalloc(&d->pe_tok, 4096 * 256 * sizeof(float)) != cudaSuccess) {
    set_last_error("decoder_create: cudaMalloc of the dense PE table failed");
    delete d;
    return 1;
  }
  if (int rc = dense_pe_tokens(d->w[W_GAUSS], d->pe_tok, stream)) { cudaFree(d->pe_tok); delete d; return rc; }
  // hi/hi/lo splits of the image-side weights ([N, 3K] bf16)
  struct Item { int idx, N, K; const __nv_bfloat16** dst; };
  std::vector<Item> items;
  for (int l = 0; l < 2; ++l) {
    const int L = W_LAYER0 + l * LAYER_STRIDE;
    items.push_back({L + L_T2I + 2, 128, 256, &d->ws_t2i_k[l]});
    items.push_back({L + L_T2I + 4, 128, 256, &d->ws_t2i_v[l]});
    items.push_back({L + L_I2T + 0, 128, 256, &d->ws_i2t_q[l]});
    items.push_back({L + L_I2T + 6, 256, 128, &d->ws_i2t_o[l]});
    items.push_back({L + L_MLP + 0, 2048, 256, &d->ws_mlp1[l]});
    items.push_back({L + L_MLP + 2, 256, 2048, &d->ws_mlp2[l]});
  }
  items.push_back({W_FINAL + 2, 128, 256, &d->ws_fin_k});
  items.push_back({W_FINAL + 4, 128, 256, &d->ws_fin_v});
  items.push_back({W_UP + 0, 256, 256, &d->ws_up1});
  items.push_back({W_UP + 4, 128, 64, &d->ws_up2});
  size_t total = 0;
  for (const Item& it : items) total += static_cast<size_t>(it.N) * 3 * it.K;
  d->wsplit = nullptr;
  if (cudaMalloc(&d->wsplit, total * sizeof(__nv_bfloat16)) != cudaSuccess) {
    set_last_error("decoder_create: cudaMalloc of the split weights failed");
    cudaFree(d->pe_tok);
    delete d;
    return 1;
  }
  size_t off = 0;
  for (const Item& it : items) {
    *it.dst = d->wsplit + off;
    if (int rc = split3_bf16(d->w[it.idx], nullptr, 0, d->wsplit + off, it.N, it.K, 1, stream)) {
      cudaFree(d->wsplit); cudaFree(d->pe_tok); delete d; return rc;
    }
    off += static_cast<size_t>(it.N) * 3 * it.K;
  }
  // pe W^T + b tables of the five projections that take keys + pe (fp32 linear, once), interleaved with zero tables so that
  // the fused k | v | q launch finds them one plane apart (decoder.h)
  constexpr size_t TAB = static_cast<size_t>(4096) * 128;
  d->pek = nullptr;
  d->bias_kvq = nullptr;
  if (cudaMalloc(&d->pek, 8 * TAB * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d->bias_kvq, (2 * 384 + 256) * sizeof(float)) != cudaSuccess) {
    set_last_error("decoder_create: cudaMalloc of the positional-encoding tables failed");
    if (d->pek) cudaFree(d->pek);
    cudaFree(d->wsplit); cudaFree(d->pe_tok); delete d;
    return 1;
  }
  {
    const float* const* Wt = d->w.data();
    cudaError_t e = cudaMemsetAsync(d->pek, 0, 8 * TAB * sizeof(float), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d->bias_kvq, 0, (2 * 384 + 256) * sizeof(float), stream);
    struct PeItem { int w_idx, b_idx; const float** dst; size_t slot; };
    std::vector<PeItem> pis;
    for (int l = 0; l < 2; ++l) {
      const int L = W_LAYER0 + l * LAYER_STRIDE;
      pis.push_back({L + L_T2I + 2, L + L_T2I + 3, &d->pek_t2i_k[l], static_cast<size_t>(3 * l)});
      pis.push_back({L + L_I2T + 0, L + L_I2T + 1, &d->pek_i2t_q[l], static_cast<size_t>(3 * l + 2)});
      d->bias_kvq_l[l] = d->bias_kvq + 384 * l;
      if (e == cudaSuccess)  // v bias into the middle third
        e = cudaMemcpyAsync(d->bias_kvq + 384 * l + 128, Wt[L + L_T2I + 5], 128 * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    }
    pis.push_back({W_FINAL + 2, W_FINAL + 3, &d->pek_fin_k, 6});
    d->bias_fin_kv = d->bias_kvq + 768;
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(d->bias_kvq + 768 + 128, Wt[W_FINAL + 5], 128 * sizeof(float), cudaMemcpyDeviceToDevice, stream);
    if (e != cudaSuccess) {
      set_last_error("decoder_create: preparing the fused-projection tables failed: %s", cudaGetErrorString(e));
      cudaFree(d->bias_kvq); cudaFree(d->pek); cudaFree(d->wsplit); cudaFree(d->pe_tok); delete d;
      return 1;
    }
    for (const PeItem& it : pis) {
      float* dstp = d->pek + it.slot * TAB;
      *it.dst = dstp;
      if (int rc = lin(d->pe_tok, nullptr, 0, Wt[it.w_idx], Wt[it.b_idx], nullptr, dstp, 4096, 128, 256, 0, stream)) {
        cudaFree(d->bias_kvq); cudaFree(d->pek); cudaFree(d->wsplit); cudaFree(d->pe_tok); delete d; return rc;
      }
    }
  }
  *out = d;
  return 0;
}

void decoder_destroy(Decoder* d) {
  if (d == nullptr) return;
  if (d->bias_kvq) cudaFree(d->bias_kvq);
  if (d->pek) cudaFree(d->pek);
  if (d->pe_tok) cudaFree(d->pe_tok);
  if (d->wsplit) cudaFree(d->wsplit);
  delete d;
}

const float* decoder_dense_pe(const Decoder* d) { return d->pe_tok; }

int prompt_encode(const Decoder* d, const float* coords, const int* labels, int NB, int Np, const float* mask_prev,
                  float* tokens_tmp, int* ntok_tmp, float* sparse_out, float* dense_tok_out, cudaStream_t s) {
  B200SAM_REQUIRE(NB > 0 && Np >= 0 && Np <= 27, "prompt_encode: bad shape NB=%d Np=%d", NB, Np);
  B200SAM_REQUIRE(dense_tok_out != nullptr && (Np == 0 || (coords && labels && tokens_tmp && ntok_tmp && sparse_out)),
                  "prompt_encode: null pointer argument");
  const float* const* W = d->w.data();
  if (Np > 0) {
    const int T = 5 + Np;
    TRY(prompt_tokens(coords, labels, NB, Np, W[W_GAUSS], W[W_POINT4], W[W_NOT_A_POINT], W[W_IOU_TOKEN], W[W_MASK_TOKENS],
                      1024.0f, 1024.0f, tokens_tmp, ntok_tmp, s));
    B200SAM_CHECK_CUDA(cudaMemcpy2DAsync(sparse_out, static_cast<size_t>(Np) * 256 * sizeof(float), tokens_tmp + 5 * 256,
                                         static_cast<size_t>(T) * 256 * sizeof(float),
                                         static_cast<size_t>(Np) * 256 * sizeof(float), NB, cudaMemcpyDeviceToDevice, s));
  }
  if (mask_prev != nullptr)
    TRY(mask_downscale_keys(mask_prev, W + W_MASKDOWN, nullptr, dense_tok_out, NB, nullptr, nullptr, s));
  else
    TRY(broadcast_row256(W[W_NO_MASK], dense_tok_out, static_cast<size_t>(NB) * 4096, s));
  return 0;
}

int decoder_forward(const Decoder* d, const DecodeArgs& a, cudaStream_t s) {
  B200SAM_REQUIRE(a.NB > 0 && a.Np >= 0 && a.n_images > 0, "decode: bad prompt batch NB=%d Np=%d images=%d", a.NB, a.Np,
                  a.n_images);
  B200SAM_REQUIRE(a.n_images == 1 || a.image_of != nullptr, "decode: image_of is required for more than one image");
  B200SAM_REQUIRE(a.NB <= 65535 && a.NB <= (1 << 30) / 16384, "decode: at most 65535 prompts per call, got %d", a.NB);
  const int NB = a.NB, T = 5 + a.Np;
  B200SAM_REQUIRE(T <= 32, "decode: at most 27 sparse prompt tokens per prompt supported, got %d", a.Np);
  B200SAM_REQUIRE(a.Np == 0 || a.sparse_tokens != nullptr || (a.coords != nullptr && a.labels != nullptr),
                  "decode: coords/labels missing");
  B200SAM_REQUIRE(a.emb != nullptr && a.low_res_out != nullptr && a.iou_out != nullptr, "decode: null in/out pointer");
  B200SAM_REQUIRE(a.workspace != nullptr && (reinterpret_cast<uintptr_t>(a.workspace) & 255) == 0,
                  "decode: workspace must be non-null and 256-byte aligned");
  Workspace w = carve(reinterpret_cast<uint8_t*>(a.workspace), a.n_images, NB, T);
  B200SAM_REQUIRE(w.total <= a.workspace_bytes, "decode: workspace too small (%zu < %zu)", a.workspace_bytes, w.total);
  B200SAM_REQUIRE(w.vbuf - w.kbuf == w.qibuf - w.vbuf, "decode: k / v / q output planes must be equally spaced");
  const float* const* W = d->w.data();
  const int Mi = NB * 4096, Mt = NB * T;
  const bool share0 = a.mask_prev == nullptr && a.dense_tok == nullptr && a.image_of != nullptr && a.n_images < NB;
  const float* pe = d->pe_tok;

  // ---- prompt encoder (prompt_encoder.py:128-168) + output tokens (mask_decoder.py:120-122)
  if (a.sparse_tokens != nullptr)
    TRY(tokens_from_sparse(a.sparse_tokens, NB, a.Np, W[W_IOU_TOKEN], W[W_MASK_TOKENS], w.tokens, w.ntok, s));
  else
    TRY(prompt_tokens(a.coords, a.labels, NB, a.Np, W[W_GAUSS], W[W_POINT4], W[W_NOT_A_POINT], W[W_IOU_TOKEN],
                      W[W_MASK_TOKENS], a.img_w, a.img_h, w.tokens, w.ntok, s));
  TRY(nchw_to_tokens(a.emb, w.emb_tok, a.n_images, s));
  // the producers of the image-side keys also emit sb = split(keys), the bf16 [hi | lo] operand of EVERY image-side
  // projection: the ones that take keys + pe (k of the token->image attentions, q of the image->token attention) add
  // their positional term as the per-token table pek_* = pe W^T + b in the GEMM epilogue (residual row = row % 4096)
  if (a.dense_tok != nullptr) {
    TRY(keys_init(w.emb_tok, W[W_NO_MASK], w.keys, NB, a.image_of, pe, nullptr, w.sb, s, a.dense_tok));
  } else if (a.mask_prev != nullptr) {
    TRY(mask_downscale_keys(a.mask_prev, W + W_MASKDOWN, w.emb_tok, w.keys, NB, a.image_of, w.sb, s));
  } else {
    // Without mask prompts the image-side keys emb + no_mask are the same for all prompts of an image until the first
    // image->token block updates them: the three image-side projections of layer 0 then run once per IMAGE (shared)
    // and the attention kernels pick the image's block through image_of.
    TRY(keys_init(w.emb_tok, W[W_NO_MASK], w.keys, NB, a.image_of, pe, nullptr, share0 ? nullptr : w.sb, s));
  }
  B200SAM_CHECK_CUDA(cudaMemcpyAsync(w.queries, w.tokens, static_cast<size_t>(Mt) * 256 * sizeof(float),
                                     cudaMemcpyDeviceToDevice, s));

  // ---- two-way transformer (transformer.py:62-106, :151-182)
  for (int l = 0; l < 2; ++l) {
    const float* const* L = W + W_LAYER0 + l * LAYER_STRIDE;
    const float* const* SA = L + L_SELF;
    const float* qpe = l == 0 ? nullptr : w.tokens;  // layer 0 skips the PE and REPLACES the queries
    TRY(lin(w.queries, qpe, 0, SA[0], SA[1], nullptr, w.tq, Mt, 256, 256, 0, s));
    TRY(lin(w.queries, qpe, 0, SA[2], SA[3], nullptr, w.tk, Mt, 256, 256, 0, s));
    TRY(lin(w.queries, nullptr, 0, SA[4], SA[5], nullptr, w.tv, Mt, 256, 256, 0, s));
    TRY(attn_few_queries(w.tq, w.tk, w.tv, w.ta, NB, T, T, 8, 32, nullptr, w.ntok, s));
    TRY(lin(w.ta, nullptr, 0, SA[6], SA[7], l == 0 ? nullptr : w.queries, w.queries, Mt, 256, 256, 0, s));
    TRY(layernorm_rows(w.queries, L[L_N1], L[L_N1 + 1], 1e-5f, Mt, 256, w.queries, 0, s));

    const float* const* TI = L + L_T2I;
    TRY(lin(w.queries, w.tokens, 0, TI[0], TI[1], nullptr, w.tq, Mt, 128, 256, 0, s));
    // image-side projections on the tensor cores (3-way bf16 split operands, fp32 accumulate); keys are constant
    // until the end of the layer, so the two split operands also serve the image->token query projection
    const bool shared = share0 && l == 0;
    const int Mp = shared ? a.n_images * 4096 : Mi;            // rows of the image-side projections
    const __nv_bfloat16* sbp = w.sb;
    if (shared) {  // split(emb + no_mask) per image, in the (still unused) sa buffer
      TRY(split3_bf16(w.emb_tok, W[W_NO_MASK], 1, w.sa, static_cast<size_t>(Mp), 256, 0, s));
      sbp = w.sa;
    }
    const int* blk_of = shared ? a.image_of : nullptr;
    // k | v | image-side q in one launch: split(keys) is read from HBM once instead of three times
    TRY(tc_lin_planes(sbp, d->ws_t2i_k[l], d->bias_kvq_l[l], d->pek_t2i_k[l], static_cast<size_t>(4096) * 128, w.kbuf,
                      static_cast<size_t>(w.vbuf - w.kbuf), 3, Mp, 256, s));
    TRY(attn_few_queries(w.tq, w.kbuf, w.vbuf, w.ta, NB, T, 4096, 8, 16, w.part, nullptr, s, blk_of));
    TRY(lin(w.ta, nullptr, 0, TI[6], TI[7], w.queries, w.queries, Mt, 256, 128, 0, s));
    TRY(layernorm_rows(w.queries, L[L_N2], L[L_N2 + 1], 1e-5f, Mt, 256, w.queries, 0, s));

    // token-side MLP on the tensor cores as well (3-way bf16 split operands like the image side; the scalar fp32 kernel
    // needed 48 - 140 us per linear for these [prompts x tokens, 256 / 2048] shapes).  The two split operands live at the
    // head of sa, which is free between the k | v | q projection above and the image->token attention below.
    {
      __nv_bfloat16* sq = w.sa;
      __nv_bfloat16* sth = w.sa + ((static_cast<size_t>(Mt) * 512 + 127) & ~static_cast<size_t>(127));
      TRY(split3_bf16(w.queries, nullptr, 0, sq, static_cast<size_t>(Mt), 256, 0, s));
      TRY(tc_lin(sq, d->ws_mlp1[l], L[L_MLP + 1], nullptr, w.th, Mt, 2048, 256, 2, s));   // + ReLU
      TRY(split3_bf16(w.th, nullptr, 0, sth, static_cast<size_t>(Mt), 2048, 0, s));
      TRY(tc_lin(sth, d->ws_mlp2[l], L[L_MLP + 3], w.queries, w.queries, Mt, 256, 2048, 0, s));
    }
    TRY(layernorm_rows(w.queries, L[L_N3], L[L_N3 + 1], 1e-5f, Mt, 256, w.queries, 0, s));

    const float* const* IT = L + L_I2T;  // image tokens are the queries here
    TRY(lin(w.queries, w.tokens, 0, IT[2], IT[3], nullptr, w.tk, Mt, 128, 256, 0, s));
    TRY(lin(w.queries, nullptr, 0, IT[4], IT[5], nullptr, w.tv, Mt, 128, 256, 0, s));
    TRY(attn_few_keys(w.qibuf, w.tk, w.tv, nullptr, NB, 4096, T, w.ntok, w.sa, s, blk_of));  // -> split(attention out)
    TRY(tc_lin(w.sa, d->ws_i2t_o[l], IT[7], w.keys, w.keys, Mi, 256, 128, 0, s));
    TRY(ln256_keys_split(w.keys, L[L_N4], L[L_N4 + 1], pe, Mi, nullptr, w.sb, s));  // norm4 + split(keys)
  }
  {
    const float* const* F = W + W_FINAL;
    TRY(lin(w.queries, w.tokens, 0, F[0], F[1], nullptr, w.tq, Mt, 128, 256, 0, s));
    TRY(tc_lin_planes(w.sb, d->ws_fin_k, d->bias_fin_kv, d->pek_fin_k, static_cast<size_t>(4096) * 128, w.kbuf,
                      static_cast<size_t>(w.vbuf - w.kbuf), 2, Mi, 256, s));
    TRY(attn_few_queries(w.tq, w.kbuf, w.vbuf, w.ta, NB, T, 4096, 8, 16, w.part, nullptr, s));
    TRY(lin(w.ta, nullptr, 0, F[6], F[7], w.queries, w.queries, Mt, 256, 128, 0, s));
    TRY(layernorm_rows(w.queries, W[W_NF], W[W_NF + 1], 1e-5f, Mt, 256, w.queries, 0, s));
  }

  // ---- hypernetwork heads, then the upscaler with its element-wise stages fused into the two ConvT GEMM epilogues
  //      (mask_decoder.py:137-147): ConvT1 + LayerNorm2d(64) + GELU -> split operand; ConvT2 + GELU + hyper dot -> masks
  const float* const* U = W + W_UP;
  {
    const float* wt[15];
    const float* bs[15];
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 3; ++j) { wt[i * 3 + j] = W[W_HYPER + (i * 3 + j) * 2]; bs[i * 3 + j] = W[W_HYPER + (i * 3 + j) * 2 + 1]; }
    for (int j = 0; j < 3; ++j) { wt[12 + j] = W[W_IOUHEAD + 2 * j]; bs[12 + j] = W[W_IOUHEAD + 2 * j + 1]; }
    TRY(mlp3_tokens(w.queries, NB, T, wt, bs, w.hyper, w.iou4, s));
  }
  const int tok0 = a.multimask ? 1 : 0, ntok = a.multimask ? 3 : 1;  // mask_decoder.py:101-107
  {
    // w.sb still holds the split of the final keys (they do not change after the last image->token block)
    GemmArgs g;
    g.A = w.sb; g.B = d->ws_up1; g.out = w.sa; g.bias = U[1]; g.residual = nullptr;
    g.M = Mi; g.N = 256; g.K = 3 * 256; g.lda = 2 * 256; g.ldb = 3 * 256; g.ldo = 0; g.ldr = 0; g.res_row_mod = 0;
    g.gelu = 0; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 2 * 256;
    g.epi_mode = 1; g.aux0 = U[2]; g.aux1 = U[3];
    TRY(gemm_bf16_tn(g, s));                                                           // ConvT 256->64 + LN2d + GELU
  }
  {
    GemmArgs g;
    g.A = w.sa; g.B = d->ws_up2; g.out = a.low_res_out; g.bias = U[5]; g.residual = nullptr;
    g.M = Mi * 4; g.N = 128; g.K = 3 * 64; g.lda = 2 * 64; g.ldb = 3 * 64; g.ldo = 0; g.ldr = 0; g.res_row_mod = 0;
    g.gelu = 1; g.out_kind = 0; g.max_ctas = 0; g.a_wrap = 2 * 64;
    g.epi_mode = 2; g.aux0 = w.hyper; g.tok0 = tok0; g.ntok = ntok;
    TRY(gemm_bf16_tn(g, s));                                                           // ConvT 64->32 + GELU + mask dot
  }
  slice_iou_kernel<<<(NB * ntok + 127) / 128, 128, 0, s>>>(w.iou4, a.iou_out, NB, tok0, ntok);
  B200SAM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200sam
