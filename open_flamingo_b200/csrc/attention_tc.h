// Internal interface of the tcgen05 / TMA attention cores (attention_tc.cu).  The extern "C" entry points in
// attention.cu (media-masked cores: gated cross-attention, Perceiver, ViT) and attention_dense.cu (LM self-attention)
// fill this description and try the tensor-core path first; the mma.sync kernels in those files remain as the
// path for layouts TMA cannot describe (and as the A/B reference with OFK_ATTN_LEGACY=1).
#pragma once

namespace ofk {
namespace tc {

struct Args {
  const void *q, *k, *v, *o, *d_o;
  void *out, *dq, *dk, *dv;
  float* lse;      // [batch, heads, nq], log2 domain: m + log2(l); 0 for rows without any mass
  float* delta;    // [batch, heads, nq] scratch (backward)
  int batch, heads, hd, nq, nk;
  long long q_bs, ldq, k_bs, ldk, v_bs, ldv, o_bs, ldo, dq_bs, lddq, dk_bs, lddk, dv_bs, lddv;
  float scale;
  int dense;                       // 0: media rules (mask_mode / text_time / kpm), 1: dense rules (causal / mask / slopes)
  int mask_mode; const int* text_time; int kpm;
  int causal; const unsigned char* mask; const float* slopes; const int* pure_causal;
  void* workspace; long long workspace_bytes;   // backward: fp32 dQ accumulator (see bwd_workspace_bytes)
  void* stream;
  int q_tile_limit;                // forward: > 0 = only the first q_tile_limit 128-query tiles (the caller covers the tail rows)
};

// true when the tensor-core kernels can run this problem (TMA-describable layout, supported head_dim); depends only
// on shapes / strides / pointers' alignment, so forward and backward of the same tensors take the same decision.
bool fwd_supported(const Args& a);
bool bwd_supported(const Args& a);
long long bwd_workspace_bytes(int batch, int heads, int hd, int nq, int nk);
int fwd(const Args& a);   // 0 or OFK_ERR_*
int bwd(const Args& a);
bool legacy_forced();     // OFK_ATTN_LEGACY=1

}  // namespace tc
}  // namespace ofk
