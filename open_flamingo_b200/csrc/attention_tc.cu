// Attention cores on the 5th-generation tensor cores (sm_100a): TMA-staged Q/K/V/dO tiles (128-byte swizzle),
// tcgen05.mma for every contraction with the accumulators in TMEM, softmax on a lane-per-row read-out of TMEM
// (no shuffles), P / dS handed back to the tensor core through swizzled shared memory.
//
// Two mask families, one kernel template each for forward and backward (head_dim 64 or 128):
//   media rules (DENSE = false) -- MaskedCrossAttention core helpers.py:190-232 (text_time == / >= media index,
//       zero rows helpers.py:223-229, uniform rows for the masked_fill + softmax of a fully masked row),
//       PerceiverAttention core helpers.py:55-64 and the ViT MHA core (mask_mode 0);
//   dense rules (DENSE = true)  -- the frozen LM's self-attention (HF MptAttention, reached via
//       flamingo_lm.py:63-65): causal + ALiBi + optional byte mask with masked_fill(finfo.min) semantics.
// Semantics, arguments and outputs are those of the mma.sync kernels in attention.cu / attention_dense.cu, which
// stay as the path for layouts TMA cannot describe; LSE is log2-domain (m + log2 l) in both families.
//
// Forward  (CTA = 128 queries of one (batch, head); 192 threads; 2 CTAs / SM; 64-key tiles)
//   warp 4 lane 0 : TMA producer (Q once; a 2-stage ring of K tiles, V tiles) + TMEM alloc (256 cols)
//   warp 5 lane 0 : MMA issuer   S_j = Q K_j^T into a DOUBLE-BUFFERED S (S_{j+1} is issued while the softmax warps work on
//                   S_j) ; O += P_j V_j (A = P from smem, B = V through the MN-major descriptor)
//   warps 0-3     : thread = query row = TMEM lane.  One TMEM read of the tile's 64 scores into registers, mask / scale,
//                   p = exp2(s - m) as bf16 into the K-major SW128 layout; the running output in TMEM is rescaled lazily
//                   (only when a row maximum grew by more than 2^8); finally O / l -> global, LSE.
// Backward (CTA = 128 keys of one (batch, head), loop over the query tiles that can see them; 320 threads; 1 CTA / SM)
//   warp 8 lane 0 : TMA (K, V once; Q_i, dO_i per query tile) + TMEM alloc (512 cols)
//   warp 9 lane 0 : MMA  S^T = K Q_i^T, dP^T = V dO_i^T  ->  [threads]  ->  dV += P^T dO_i, dK += dS^T Q_i,
//                   dQ_i(partial) = dS K   (dS read through the MN-major descriptor from the same dS^T tile)
//   warps 0-7     : thread = key row; P^T = exp2(S^T - lse_q), dS^T = P^T (dP^T - delta_q) scale, both to smem as
//                   bf16; dQ partials go to an fp32 accumulator with red.global.add (or straight to bf16 dQ when a
//                   single key tile covers all keys); dK / dV are drained once at the end.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "attention_tc.h"
#include "ofk_internal.h"
#include "ofk_ptx.cuh"

namespace ofk {
namespace tc {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float MASKED = -30000.0f;   // log2-domain stand-in for masked_fill(finfo.min): exp2(MASKED - m) == 0 for real m

struct Params {
  __nv_bfloat16 *out, *dq, *dk, *dv;
  float* lse;
  const float* delta;
  float* dq32;                      // [batch * nq, heads * HD] fp32 accumulator (backward, unless dq_direct)
  const int* text_time;
  const unsigned char* mask;
  const float* slopes;
  const int* pure_causal;
  int batch, heads, nq, nk;
  long long o_bs, ldo, dq_bs, lddq, dk_bs, lddk, dv_bs, lddv;
  float scale;
  int mask_mode, kpm, causal, dq_direct;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- media mask rules (see attention.cu)
// kind 0: normal masked row; 1: zero row / row outside the problem; 2: uniform row (S = 0 over all keys, no
// gradient to q / k); 3: unmasked
struct MediaRow { int tt, kind; };
__device__ __forceinline__ MediaRow classify(const Params& p, int b, int row) {
  MediaRow r; r.tt = 0; r.kind = 3;
  if (row >= p.nq) { r.kind = 1; return r; }
  if (p.mask_mode == 0) return r;
  const int tt = p.text_time[(long long)b * p.nq + row];
  const int n_media = p.nk / p.kpm;
  r.tt = tt;
  const bool has = (p.mask_mode == 1) ? (tt >= 1 && tt <= n_media) : (tt >= 1);
  if (has) r.kind = 0;
  else if (p.mask_mode == 1 && tt == 0) r.kind = 1;
  else r.kind = 2;
  return r;
}
__device__ __forceinline__ bool media_allowed(int mask_mode, int kind, int tt, int media) {
  if (kind >= 2) return true;
  if (kind == 1) return false;
  return mask_mode == 1 ? (tt == media) : (tt >= media);
}

// ================================================================================================ forward
constexpr int FWD_THREADS = 192;

template <bool DENSE>
struct RowCtx;
template <>
struct RowCtx<false> { int kind, tt; float z; };
template <>
struct RowCtx<true> { int valid, causal_last; const unsigned char* mrow; float sl2, slope2; };

// log2-domain scores of 16 consecutive keys [kg, kg + 16) of this thread's query row.
template <bool DENSE>
__device__ __forceinline__ void scores16(const Params& p, const RowCtx<DENSE>& rc, const uint32_t* acc, int kg, float* s) {
  if constexpr (!DENSE) {
    const bool ok = rc.kind != 1 && (p.mask_mode == 0 || media_allowed(p.mask_mode, rc.kind, rc.tt, kg / p.kpm + 1));
    if (ok && kg + 16 <= p.nk) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = __uint_as_float(acc[i]) * rc.z;
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = (ok && kg + i < p.nk) ? __uint_as_float(acc[i]) * rc.z : -INFINITY;
    }
  } else {
    if (!rc.valid) {
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = -INFINITY;
      return;
    }
    if (rc.mrow == nullptr && kg + 15 <= rc.causal_last && kg + 16 <= p.nk) {   // fully visible group
      const float kb = rc.slope2 * (float)kg;
#pragma unroll
      for (int i = 0; i < 16; ++i) s[i] = fmaf(__uint_as_float(acc[i]), rc.sl2, fmaf(rc.slope2, (float)i, kb));
      return;
    }
    uint32_t mb[4] = {0u, 0u, 0u, 0u};   // 16 mask bytes
    if (rc.mrow != nullptr) {
      if (((p.nk | kg) & 15) == 0 && (reinterpret_cast<uintptr_t>(rc.mrow) & 15) == 0 && kg + 16 <= p.nk) {
        const uint4 u = *reinterpret_cast<const uint4*>(rc.mrow + kg);
        mb[0] = u.x; mb[1] = u.y; mb[2] = u.z; mb[3] = u.w;
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (kg + i < p.nk && rc.mrow[kg + i]) mb[i >> 2] |= 1u << ((i & 3) * 8);
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int key = kg + i;
      const bool masked = key > rc.causal_last || ((mb[i >> 2] >> ((i & 3) * 8)) & 0xffu) != 0;
      const float v = masked ? MASKED : fmaf(__uint_as_float(acc[i]), rc.sl2, rc.slope2 * (float)key);
      s[i] = key < p.nk ? v : -INFINITY;
    }
  }
}

// ================================================================================================ forward
// Forward kernel (second generation; the first one used 128-key tiles, two passes over S and a single S buffer and was
// 25 % slower, profiles/r02_attention_by_shape.md): 64-key tiles for both head dims, the S accumulator double-buffered in TMEM (S_{j+1} = Q K_{j+1}^T
// is issued while the softmax warps work on S_j), a two-stage K ring, the scores of a tile held in registers (one TMEM
// read, one pass), and FA4-style lazy rescaling: the running output in TMEM is only rescaled when some row's maximum
// grew by more than 2^8 since the value the accumulators are expressed in (p <= 256 is harmless in fp32 / bf16), so
// most tiles skip the TMEM read-modify-write and the exponentials of tile j overlap P V_{j-1}.
template <int HD>
struct Fwd2Cfg {
  static constexpr int BKT = 64;
  static constexpr int ATOMS = HD / 64;
  static constexpr int Q_ATOM = 128 * 128;
  static constexpr int KV_ATOM = BKT * 128;            // 8 KiB: 64 keys x 64 head-dim columns
  static constexpr int Q_BYTES = ATOMS * Q_ATOM;
  static constexpr int KV_BYTES = ATOMS * KV_ATOM;
  static constexpr int NKST = 2;
  static constexpr int NVST = HD == 64 ? 2 : 1;        // shared memory: 64 KiB (hd 64) / 96 KiB (hd 128) -> 2 CTAs / SM
  static constexpr int P_BYTES = Q_ATOM;               // [128 queries x 64 keys] bf16, one swizzle atom
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + Q_BYTES;
  static constexpr int OFF_V = OFF_K + NKST * KV_BYTES;
  static constexpr int OFF_P = OFF_V + NVST * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + P_BYTES;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = 256;           // S[0]: 0..63, S[1]: 64..127, O: 128..128+HD
  static constexpr uint32_t O_COL = 128;
};
constexpr float RESCALE_THRESHOLD = 8.0f;              // log2 units

template <int HD, bool DENSE>
__global__ void __launch_bounds__(FWD_THREADS, 2)
attn_fwd2_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                    const __grid_constant__ CUtensorMap tma_v, const Params p) {
  using C = Fwd2Cfg<HD>;
  constexpr int BKT = C::BKT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + C::OFF_Q;
  uint8_t* sK = smem + C::OFF_K;
  uint8_t* sV = smem + C::OFF_V;
  uint8_t* sP = smem + C::OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [NVST]
  uint64_t* v_empty = bars + 5;    // [NVST]
  uint64_t* s_full = bars + 7;     // [2]
  uint64_t* p_full = bars + 9;
  uint64_t* o_full = bars + 10;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
  int* s_range = reinterpret_cast<int*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) { s_range[0] = 1 << 30; s_range[1] = -1; s_range[2] = 0; }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&k_full[i], 1); mbar_init(&s_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1); }
    // softmax warps whose 32 query rows all lie outside the problem (64 Perceiver latents in a 128-row tile, the ragged
    // last tile of a sequence) do nothing: they neither arrive on p_full nor read TMEM; their P rows stay unwritten,
    // which only affects their own (never stored) output rows
    mbar_init(p_full, 32 * min(4, (p.nq - q0 + 31) / 32)); mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); }
    __syncwarp();
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }
  __syncthreads();

  bool causal = false;
  const unsigned char* mask = nullptr;
  if constexpr (DENSE) {
    causal = p.causal != 0; mask = p.mask;
    if (p.pure_causal != nullptr && *p.pure_causal != 0) { mask = nullptr; causal = true; }
  }
  RowCtx<DENSE> rc;
  if constexpr (!DENSE) {
    rc.kind = 1; rc.tt = 0; rc.z = 0.f;
    if (warp < 4) {
      const MediaRow r = classify(p, b, q0 + threadIdx.x);
      rc.kind = r.kind; rc.tt = r.tt;
      rc.z = r.kind == 2 ? 0.f : p.scale * LOG2E;
      if (p.mask_mode != 0) {
        if (r.kind == 0) { atomicMin(&s_range[0], r.tt); atomicMax(&s_range[1], r.tt); }
        if (r.kind == 2) s_range[2] = 1;
      }
    }
  } else {
    const int row = q0 + (int)threadIdx.x;
    rc.valid = warp < 4 && row < p.nq;
    rc.causal_last = causal ? row + (p.nk - p.nq) : 0x7fffffff;
    rc.mrow = (mask != nullptr && rc.valid) ? mask + ((long long)b * p.nq + row) * p.nk : nullptr;
    rc.sl2 = p.scale * LOG2E;
    rc.slope2 = p.slopes != nullptr ? p.slopes[h] * LOG2E : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  int lo = 0, hi = p.nk;
  if constexpr (!DENSE) {
    if (p.mask_mode != 0 && s_range[2] == 0) {
      if (s_range[1] < 0) { lo = 0; hi = 0; }
      else {
        hi = min(p.nk, s_range[1] * p.kpm);
        lo = p.mask_mode == 1 ? max(0, (s_range[0] - 1) * p.kpm) : 0;
      }
    }
  } else {
    if (causal && mask == nullptr) hi = max(0, min(p.nk, min(q0 + 127, p.nq - 1) + (p.nk - p.nq) + 1));
  }
  const int t_lo = lo / BKT, t_hi = (hi + BKT - 1) / BKT;
  const int nsteps = max(0, t_hi - t_lo);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0 && nsteps > 0) {
      mbar_arrive_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a) tma_load_2d(sQ + a * C::Q_ATOM, &tma_q, q_full, h * HD + 64 * a, b * p.nq + q0);
      for (int it = 0; it < nsteps; ++it) {
        const int key0 = (t_lo + it) * BKT;
        const int kb = it & 1;
        if (it >= 2) mbar_wait(&s_full[kb], ((it - 2) >> 1) & 1);        // S_{it-2} retired: K stage kb is free
        mbar_arrive_expect_tx(&k_full[kb], C::KV_BYTES);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a)
          tma_load_2d(sK + kb * C::KV_BYTES + a * C::KV_ATOM, &tma_k, &k_full[kb], h * HD + 64 * a, b * p.nk + key0);
        const int vb = it % C::NVST;
        if (it >= C::NVST) mbar_wait(&v_empty[vb], ((it - C::NVST) / C::NVST) & 1);   // P V_{it-NVST} retired
        mbar_arrive_expect_tx(&v_full[vb], C::KV_BYTES);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a)
          tma_load_2d(sV + vb * C::KV_BYTES + a * C::KV_ATOM, &tma_v, &v_full[vb], h * HD + 64 * a, b * p.nk + key0);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nsteps > 0) {
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ap = smem_u32(sP);
      auto issue_s = [&](int it) {                                  // S_it -> TMEM buffer it & 1, K stage it & 1
        const int key0 = (t_lo + it) * BKT;
        const int n_eff = min(BKT, ((p.nk - key0) + 15) & ~15);
        const uint32_t idesc = make_idesc_bf16(128, n_eff, 0, 0);
        const uint32_t kb = ak + (it & 1) * C::KV_BYTES;
        mbar_wait(&k_full[it & 1], (it >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk)
          umma_bf16(tmem_base + (it & 1) * BKT, make_smem_desc_sw128(aq + (kk >> 2) * C::Q_ATOM + (kk & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(kb + (kk >> 2) * C::KV_ATOM + (kk & 3) * 32, 0, 1024), idesc, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[it & 1]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      if (nsteps > 1) issue_s(1);
      for (int it = 0; it < nsteps; ++it) {
        const int key0 = (t_lo + it) * BKT;
        const int n_eff = min(BKT, ((p.nk - key0) + 15) & ~15);
        const int vb = it % C::NVST;
        mbar_wait(p_full, it & 1);                                  // P_it in smem, S_it consumed, O rescaled
        mbar_wait(&v_full[vb], (it / C::NVST) & 1);
        tc_fence_after();
        constexpr uint32_t idesc_o = make_idesc_bf16(128, HD, 0, 1);
        const uint32_t vbase = av + vb * C::KV_BYTES;
        for (int ks = 0; ks < n_eff / 16; ++ks)
          umma_bf16(tmem_base + C::O_COL, make_smem_desc_sw128(ap + ks * 32, 0, 1024),
                    make_smem_desc_sw128(vbase + ks * 2048, C::KV_ATOM, 1024), idesc_o, (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(&v_empty[vb]);
        if (it + 2 < nsteps) issue_s(it + 2);                       // S buffer / K stage it & 1 are free again
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / output (thread = query row = TMEM lane) =====================
    const int row = q0 + (int)threadIdx.x;
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool warp_active = q0 + warp * 32 < p.nq;                  // warp-uniform
    float m_used = -INFINITY, l = 0.f;          // the accumulators (O in TMEM, l) are expressed relative to m_used
    if (warp_active) {
    const uint32_t p_row = smem_u32(sP) + threadIdx.x * 128;
    const int sw = threadIdx.x & 7;
    for (int it = 0; it < nsteps; ++it) {
      const int key0 = (t_lo + it) * BKT;
      const int n_eff = min(BKT, ((p.nk - key0) + 15) & ~15);
      const uint32_t t_s = t_row + (it & 1) * BKT;
      mbar_wait(&s_full[it & 1], (it >> 1) & 1);
      tc_fence_after();
      uint32_t acc[64];
      tmem_ld16(t_s, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
      if (n_eff > 16) tmem_ld16(t_s + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
      if (n_eff > 32) tmem_ld16(t_s + 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[32]));
      if (n_eff > 48) tmem_ld16(t_s + 48, *reinterpret_cast<uint32_t(*)[16]>(&acc[48]));
      tmem_ld_wait();
      float s[64];
      float mx = -INFINITY;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g * 16 < n_eff) {
          scores16<DENSE>(p, rc, &acc[g * 16], key0 + g * 16, &s[g * 16]);
#pragma unroll
          for (int i = 0; i < 16; ++i) mx = fmaxf(mx, s[g * 16 + i]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) s[g * 16 + i] = -INFINITY;
        }
      }
      const bool grow = mx > m_used + RESCALE_THRESHOLD || (m_used == -INFINITY && mx > -INFINITY);
      const float m_next = grow ? mx : m_used;
      const float sub = m_next == -INFINITY ? 0.f : m_next;
      const float corr = grow ? ex2_approx(m_used - sub) : 1.0f;     // m_used = -inf -> 0
      float rs = 0.f;
      uint32_t w[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2_approx(s[i] - sub), p1 = ex2_approx(s[i + 1] - sub);
        rs += p0 + p1;
        w[i >> 1] = pack_bf16x2(p0, p1);
      }
      l = l * corr + rs;
      m_used = m_next;
      if (it > 0) {
        mbar_wait(o_full, (it - 1) & 1);                             // P V_{it-1} retired: P buffer free, O complete
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t o32[32];
            tmem_ld16(t_row + C::O_COL + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o32[0]));
            tmem_ld16(t_row + C::O_COL + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o32[16]));
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o32[i] = __float_as_uint(__uint_as_float(o32[i]) * corr);
            tmem_st16(t_row + C::O_COL + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o32[0]));
            tmem_st16(t_row + C::O_COL + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o32[16]));
          }
          tmem_st_wait();
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j * 8 < n_eff) st_shared_v4(p_row + ((j ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    if (nsteps > 0) {
      mbar_wait(o_full, (nsteps - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    __nv_bfloat16* op = p.out + b * p.o_bs + (long long)row * p.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t o32[32];
      if (nsteps > 0) {   // block-uniform: tcgen05.ld is .sync.aligned, it must never sit under a per-row condition
        tmem_ld16(t_row + C::O_COL + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o32[0]));
        tmem_ld16(t_row + C::O_COL + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o32[16]));
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o32[i] = 0u;
      }
      if (row < p.nq) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(o32[8 * j]) * inv, __uint_as_float(o32[8 * j + 1]) * inv);
          v.y = pack_bf16x2(__uint_as_float(o32[8 * j + 2]) * inv, __uint_as_float(o32[8 * j + 3]) * inv);
          v.z = pack_bf16x2(__uint_as_float(o32[8 * j + 4]) * inv, __uint_as_float(o32[8 * j + 5]) * inv);
          v.w = pack_bf16x2(__uint_as_float(o32[8 * j + 6]) * inv, __uint_as_float(o32[8 * j + 7]) * inv);
          if (!(l > 0.f)) v = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(op + c * 32 + 8 * j) = v;
        }
      }
    }
    if (row < p.nq && p.lse != nullptr) p.lse[((long long)b * p.heads + h) * p.nq + row] = l > 0.f ? m_used + log2f(l) : 0.f;
    }   // warp_active
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ================================================================================================ backward
// delta[b, h, q] = sum_d dO[q, d] * O[q, d]
template <int HD>
__global__ void delta_kernel(const __nv_bfloat16* o, const __nv_bfloat16* d_o, float* delta, int batch, int heads, int nq,
                             long long o_bs, long long ldo) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long total = (long long)batch * heads * nq;
  if (row >= total) return;
  const int lane = threadIdx.x & 31;
  const int qi = (int)(row % nq), h = (int)((row / nq) % heads), b = (int)(row / ((long long)nq * heads));
  const __nv_bfloat16* op = o + b * o_bs + (long long)qi * ldo + h * HD;
  const __nv_bfloat16* dop = d_o + b * o_bs + (long long)qi * ldo + h * HD;
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 64; ++c) {
    const uint32_t a = *reinterpret_cast<const uint32_t*>(op + c * 64 + 2 * lane);
    const uint32_t d = *reinterpret_cast<const uint32_t*>(dop + c * 64 + 2 * lane);
    v += bf16_lo(a) * bf16_lo(d) + bf16_hi(a) * bf16_hi(d);
  }
  v = warp_sum(v);
  if (lane == 0) delta[row] = v;
}

// dq(bf16, strided) = dq32(fp32 [batch * nq, cols])
__global__ void dq_convert_kernel(const float* src, __nv_bfloat16* dq, int nq, int cols, long long dq_bs, long long lddq,
                                  long long total8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total8) return;
  const int c8 = cols >> 3;
  const long long r = i / c8;
  const int c = (int)(i - r * c8) * 8;
  const float4 a = *reinterpret_cast<const float4*>(src + r * cols + c);
  const float4 bq = *reinterpret_cast<const float4*>(src + r * cols + c + 4);
  uint4 v;
  v.x = pack_bf16x2(a.x, a.y); v.y = pack_bf16x2(a.z, a.w); v.z = pack_bf16x2(bq.x, bq.y); v.w = pack_bf16x2(bq.z, bq.w);
  const long long b = r / nq, qi = r - b * nq;
  *reinterpret_cast<uint4*>(dq + b * dq_bs + qi * lddq + c) = v;
}

template <int HD>
struct BwdCfg {
  static constexpr int ATOMS = HD / 64;
  static constexpr int ATOM = 128 * 128;              // [128 rows x 64 bf16] swizzled block
  static constexpr int TILE = ATOMS * ATOM;           // a [128 x HD] operand tile
  static constexpr int NQBUF = HD == 64 ? 2 : 1;      // Q / dO ring depth (shared memory: 160 KB / 192 KB)
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE;
  static constexpr int OFF_Q = OFF_V + TILE;
  static constexpr int OFF_DO = OFF_Q + NQBUF * TILE;
  static constexpr int OFF_PT = OFF_DO + NQBUF * TILE;      // P^T  [128 keys x 128 queries] bf16, two 64-query atoms
  static constexpr int OFF_DST = OFF_PT + 2 * ATOM;         // dS^T, same layout
  static constexpr int OFF_ROW = OFF_DST + 2 * ATOM;        // lse[128], delta[128], tt[128], kind[128]
  static constexpr int OFF_QLIST = OFF_ROW + 4 * 128 * 4;   // query-tile list (MAXQT ints) + count
  static constexpr int MAXQT = 126;
  static constexpr int OFF_BAR = OFF_QLIST + (MAXQT + 2) * 4;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t ST_COL = 0, DPT_COL = 128, DK_COL = 256, DV_COL = 256 + HD;
  static constexpr bool DQ_ALIAS = HD == 128;               // dQ partial re-uses the S^T columns
  static constexpr uint32_t DQ_COL = DQ_ALIAS ? 0 : 256 + 2 * HD;
};
constexpr int BWD_THREADS = 320;

template <int HD, bool DENSE>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                   const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_do, const Params p) {
  using C = BwdCfg<HD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + C::OFF_K;
  uint8_t* sV = smem + C::OFF_V;
  uint8_t* sQ = smem + C::OFF_Q;
  uint8_t* sdO = smem + C::OFF_DO;
  uint8_t* sPT = smem + C::OFF_PT;
  uint8_t* sdST = smem + C::OFF_DST;
  float* s_lse = reinterpret_cast<float*>(smem + C::OFF_ROW);
  float* s_del = s_lse + 128;
  int* s_meta = reinterpret_cast<int*>(s_del + 128);   // media rules: kind | text_time << 2 per query row
  int* s_qlist = reinterpret_cast<int*>(smem + C::OFF_QLIST);
  int* s_nqt = s_qlist + C::MAXQT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;      // [NQBUF]
  uint64_t* a_done = bars + 3;        // S^T, dP^T in TMEM
  uint64_t* pds_full = bars + 4;      // P^T, dS^T in smem (256 arrivals)
  uint64_t* b_done = bars + 5;        // dV, dK, dQ partial MMAs retired
  uint64_t* dq_drained = bars + 6;    // dQ partial read out of TMEM (256 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  // one "buffer free" barrier per Q / dO ring slot: a parity wait may lag its barrier by at most one phase, and a
  // slot's barrier cannot complete again before the refill this wait gates (b_done itself can run one phase ahead)
  uint64_t* qdo_empty = bars + 10;    // [NQBUF]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (warp == 9 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < C::NQBUF; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(a_done, 1); mbar_init(pds_full, 256); mbar_init(b_done, 1); mbar_init(dq_drained, 256);
    fence_barrier_init();
  }
  if (warp == 8) {
    if (lane == 0) { tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); tma_prefetch_desc(&tma_do); }
    __syncwarp();
    tmem_alloc(tmem_ptr, C::TMEM_COLS);
    tmem_relinquish();
  }

  bool causal = false;
  const unsigned char* mask = nullptr;
  if constexpr (DENSE) {
    causal = p.causal != 0; mask = p.mask;
    if (p.pure_causal != nullptr && *p.pure_causal != 0) { mask = nullptr; causal = true; }
  }
  // ---- the query tiles that can see this key tile (block-uniform list in shared memory)
  const int n_qtiles_all = (p.nq + 127) / 128;
  {
    int count = 0;
    const int k_last = min(p.nk, k0 + 128) - 1;
    for (int qt = 0; qt < n_qtiles_all; ++qt) {
      int need = 0;
      if constexpr (DENSE) {
        // causal without an explicit mask: rows q with q + (nk - nq) >= k0 exist in this tile?  (with a mask every
        // tile is processed: a fully masked row attends uniformly to ALL keys)
        need = (causal && mask == nullptr) ? (min(p.nq - 1, qt * 128 + 127) + (p.nk - p.nq) >= k0) : 1;
      } else {
        if (p.mask_mode == 0) need = 1;
        else if (threadIdx.x < 128) {
          const MediaRow r = classify(p, b, qt * 128 + (int)threadIdx.x);
          if (r.kind == 2) need = 1;
          else if (r.kind == 0) {
            const int m_lo = k0 / p.kpm + 1, m_hi = k_last / p.kpm + 1;
            need = p.mask_mode == 1 ? (r.tt >= m_lo && r.tt <= m_hi) : (r.tt >= m_lo);
          }
        }
      }
      need = __syncthreads_or(need);
      if (need) {
        if (threadIdx.x == 0 && count < C::MAXQT) s_qlist[count] = qt;
        ++count;
      } else if (p.dq_direct == 1 && threadIdx.x < 256) {
        // single-key-tile mode writes dQ straight from the accumulator, so a query tile nobody visits (e.g. 128 rows
        // before the first <image>) must get its zeros here; with the fp32 accumulator the memset provides them
        constexpr int PIECES = HD / 8;                             // 16-byte pieces per row
        for (int i = threadIdx.x; i < 128 * PIECES; i += 256) {
          const int r = qt * 128 + i / PIECES;
          if (r < p.nq)
            *reinterpret_cast<uint4*>(p.dq + b * p.dq_bs + (long long)r * p.lddq + h * HD + (i % PIECES) * 8) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    if (threadIdx.x == 0) *s_nqt = min(count, C::MAXQT);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int nqt = *s_nqt;

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0 && nqt > 0) {
      mbar_arrive_expect_tx(kv_full, 2 * C::TILE);
#pragma unroll
      for (int a = 0; a < C::ATOMS; ++a) {
        tma_load_2d(sK + a * C::ATOM, &tma_k, kv_full, h * HD + 64 * a, b * p.nk + k0);
        tma_load_2d(sV + a * C::ATOM, &tma_v, kv_full, h * HD + 64 * a, b * p.nk + k0);
      }
      for (int idx = 0; idx < nqt; ++idx) {
        const int buf = idx % C::NQBUF;
        if (idx >= C::NQBUF) mbar_wait(&qdo_empty[buf], (idx / C::NQBUF - 1) & 1);   // B(idx - NQBUF) finished reading this slot
        const int q0 = s_qlist[idx] * 128;
        mbar_arrive_expect_tx(&qdo_full[buf], 2 * C::TILE);
#pragma unroll
        for (int a = 0; a < C::ATOMS; ++a) {
          tma_load_2d(sQ + buf * C::TILE + a * C::ATOM, &tma_q, &qdo_full[buf], h * HD + 64 * a, b * p.nq + q0);
          tma_load_2d(sdO + buf * C::TILE + a * C::ATOM, &tma_do, &qdo_full[buf], h * HD + 64 * a, b * p.nq + q0);
        }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0 && nqt > 0) {
      const uint32_t ak = smem_u32(sK), av = smem_u32(sV), apt = smem_u32(sPT), adst = smem_u32(sdST);
      mbar_wait(kv_full, 0);
      for (int idx = 0; idx < nqt; ++idx) {
        const int buf = idx % C::NQBUF;
        const int q0 = s_qlist[idx] * 128;
        const int n_eff = min(128, ((p.nq - q0) + 15) & ~15);     // valid query columns, rounded up to 16
        const uint32_t aq = smem_u32(sQ + buf * C::TILE), ado = smem_u32(sdO + buf * C::TILE);
        mbar_wait(&qdo_full[buf], (idx / C::NQBUF) & 1);
        if (C::DQ_ALIAS && idx > 0) mbar_wait(dq_drained, (idx - 1) & 1);
        tc_fence_after();
        {  // A phase: S^T = K Q^T, dP^T = V dO^T   (M = 128 keys, N = n_eff queries, K = HD; all K-major)
          const uint32_t idesc = make_idesc_bf16(128, n_eff, 0, 0);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem_base + C::ST_COL, make_smem_desc_sw128(ak + (kk >> 2) * C::ATOM + (kk & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(aq + (kk >> 2) * C::ATOM + (kk & 3) * 32, 0, 1024), idesc, kk > 0 ? 1u : 0u);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16(tmem_base + C::DPT_COL, make_smem_desc_sw128(av + (kk >> 2) * C::ATOM + (kk & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(ado + (kk >> 2) * C::ATOM + (kk & 3) * 32, 0, 1024), idesc, kk > 0 ? 1u : 0u);
          umma_commit(a_done);
        }
        mbar_wait(pds_full, idx & 1);
        tc_fence_after();
        {  // B phase
          constexpr uint32_t idesc_kv = make_idesc_bf16(128, HD, 0, 1);   // A = P^T / dS^T (K-major), B = dO / Q (MN-major)
          for (int ks = 0; ks < n_eff / 16; ++ks)                          // dV += P^T dO   (K = queries)
            umma_bf16(tmem_base + C::DV_COL, make_smem_desc_sw128(apt + (ks >> 2) * C::ATOM + (ks & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(ado + ks * 2048, C::ATOM, 1024), idesc_kv, (idx > 0 || ks > 0) ? 1u : 0u);
          for (int ks = 0; ks < n_eff / 16; ++ks)                          // dK += dS^T Q
            umma_bf16(tmem_base + C::DK_COL, make_smem_desc_sw128(adst + (ks >> 2) * C::ATOM + (ks & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(aq + ks * 2048, C::ATOM, 1024), idesc_kv, (idx > 0 || ks > 0) ? 1u : 0u);
          constexpr uint32_t idesc_q = make_idesc_bf16(128, HD, 1, 1);    // dQ = dS K: A = dS (MN-major view of dS^T), B = K (MN-major)
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)                                   // K = 128 keys
            umma_bf16(tmem_base + C::DQ_COL, make_smem_desc_sw128(adst + ks * 2048, C::ATOM, 1024),
                      make_smem_desc_sw128(ak + ks * 2048, C::ATOM, 1024), idesc_q, ks > 0 ? 1u : 0u);
          umma_commit(b_done);
          umma_commit(&qdo_empty[buf]);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== compute warps: thread = key row (TMEM lane), 64 of the 128 query columns =====================
    const int quarter = warp & 3, half = warp >> 2;
    const int krow = quarter * 32 + lane;                          // key row inside the tile
    const int key = k0 + krow;
    const bool key_ok = key < p.nk;
    const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const float sl2 = p.scale * LOG2E;
    const int media_k = (!DENSE && p.mask_mode != 0) ? key / p.kpm + 1 : 0;
    float slope_key = 0.f;
    if constexpr (DENSE) slope_key = (p.slopes != nullptr ? p.slopes[h] * LOG2E : 0.f) * (float)key;
    const int off = p.nk - p.nq;
    const uint32_t pt_row = smem_u32(sPT) + krow * 128, dst_row = smem_u32(sdST) + krow * 128;
    const int sw = krow & 7;
    const int tid = (int)threadIdx.x;                             // 0..255

    for (int idx = 0; idx < nqt; ++idx) {
      const int q0 = s_qlist[idx] * 128;
      const int n_eff = min(128, ((p.nq - q0) + 15) & ~15);
      // ---- stage this query tile's row data (the previous tile's readers are all past b_done(idx-1)).  Rows that
      //      contribute nothing (outside the problem, zero rows) get lse = +inf, so exp2(s - lse) is exactly 0.
      if (tid < 128) {
        const int row = q0 + tid;
        const bool ok = row < p.nq;
        const long long ri = ((long long)b * p.heads + h) * p.nq + row;
        float lse_v = ok ? p.lse[ri] : INFINITY;
        if constexpr (!DENSE) {
          const MediaRow r = classify(p, b, row);
          if (r.kind == 1) lse_v = INFINITY;
          s_meta[tid] = r.kind | (r.tt << 2);
        }
        s_lse[tid] = lse_v;
        s_del[tid] = ok ? p.delta[ri] : 0.f;
      }
      named_bar_sync(1, 256);
      mbar_wait(a_done, idx & 1);
      tc_fence_after();
      // ---- P^T and dS^T for columns [half * 64, half * 64 + 64)
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        const int col0 = half * 64 + c * 32;                       // first query column of this chunk
        if (col0 >= n_eff) break;                                  // warp-uniform
        uint32_t sv[32], dpv[32];
        tmem_ld16(t_row + C::ST_COL + col0, *reinterpret_cast<uint32_t(*)[16]>(&sv[0]));
        tmem_ld16(t_row + C::DPT_COL + col0, *reinterpret_cast<uint32_t(*)[16]>(&dpv[0]));
        const bool two = col0 + 16 < n_eff;
        if (two) {
          tmem_ld16(t_row + C::ST_COL + col0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[16]));
          tmem_ld16(t_row + C::DPT_COL + col0 + 16, *reinterpret_cast<uint32_t(*)[16]>(&dpv[16]));
        }
        tmem_ld_wait();
        uint32_t wp[16], wd[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float pv[4] = {0.f, 0.f, 0.f, 0.f}, dsv[4] = {0.f, 0.f, 0.f, 0.f};
          if (i < 16 || two) {
            const int cc = col0 + i;                               // 4 query columns: broadcast 16-byte reads of their row data
            const float4 L4 = *reinterpret_cast<const float4*>(&s_lse[cc]);
            const float4 D4 = *reinterpret_cast<const float4*>(&s_del[cc]);
            const float lse4[4] = {L4.x, L4.y, L4.z, L4.w}, del4[4] = {D4.x, D4.y, D4.z, D4.w};
            int meta4[4] = {0, 0, 0, 0};
            if constexpr (!DENSE) {
              const int4 M4 = *reinterpret_cast<const int4*>(&s_meta[cc]);
              meta4[0] = M4.x; meta4[1] = M4.y; meta4[2] = M4.z; meta4[3] = M4.w;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float sraw = __uint_as_float(sv[i + e]), dpraw = __uint_as_float(dpv[i + e]);
              float pe, de;
              if constexpr (!DENSE) {
                const int kind = meta4[e] & 3, tt = meta4[e] >> 2;
                const bool ok = key_ok && ((kind & 2) != 0 || (p.mask_mode == 1 ? tt == media_k : tt >= media_k));
                const bool uni = kind == 2;                        // uniform row: S = 0 over all keys, no grad to q / k
                pe = ok ? ex2_approx(fmaf(sraw, uni ? 0.f : sl2, -lse4[e])) : 0.f;
                de = uni ? 0.f : pe * (dpraw - del4[e]) * p.scale;
              } else {
                const int qrow = q0 + cc + e;
                bool masked = causal && key > qrow + off;
                if (mask != nullptr && !masked && key_ok && lse4[e] < INFINITY)
                  masked = mask[((long long)b * p.nq + qrow) * p.nk + key] != 0;
                const float sc = masked ? MASKED : fmaf(sraw, sl2, slope_key);
                pe = key_ok ? ex2_approx(sc - lse4[e]) : 0.f;
                de = masked ? 0.f : pe * (dpraw - del4[e]) * p.scale;
              }
              pv[e] = pe; dsv[e] = de;
            }
          }
          wp[i >> 1] = pack_bf16x2(pv[0], pv[1]); wp[(i >> 1) + 1] = pack_bf16x2(pv[2], pv[3]);
          wd[i >> 1] = pack_bf16x2(dsv[0], dsv[1]); wd[(i >> 1) + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
        // 32 query columns = four 16-byte pieces of this key row inside the 64-query atom `half`
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < 2 || two) {
            const int piece = c * 4 + j;
            st_shared_v4(pt_row + half * C::ATOM + ((piece ^ sw) << 4), wp[4 * j], wp[4 * j + 1], wp[4 * j + 2], wp[4 * j + 3]);
            st_shared_v4(dst_row + half * C::ATOM + ((piece ^ sw) << 4), wd[4 * j], wd[4 * j + 1], wd[4 * j + 2], wd[4 * j + 3]);
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      // ---- dQ partial of this query tile: rows = queries (lanes), this warp's HD / 2 columns
      mbar_wait(b_done, idx & 1);
      tc_fence_after();
      {
        const int qrow = q0 + krow;                                // lane = query row in the dQ accumulator
#pragma unroll 1
        for (int c = 0; c < HD / 64; ++c) {
          const int col = half * (HD / 2) + c * 32;
          uint32_t acc[32];
          tmem_ld16(t_row + C::DQ_COL + col, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
          tmem_ld16(t_row + C::DQ_COL + col + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
          tmem_ld_wait();
          if (qrow < p.nq) {
            if (p.dq_direct) {
              __nv_bfloat16* g = p.dq + b * p.dq_bs + (long long)qrow * p.lddq + h * HD + col;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(acc[8 * j]), __uint_as_float(acc[8 * j + 1]));
                v.y = pack_bf16x2(__uint_as_float(acc[8 * j + 2]), __uint_as_float(acc[8 * j + 3]));
                v.z = pack_bf16x2(__uint_as_float(acc[8 * j + 4]), __uint_as_float(acc[8 * j + 5]));
                v.w = pack_bf16x2(__uint_as_float(acc[8 * j + 6]), __uint_as_float(acc[8 * j + 7]));
                if (p.dq_direct == 1) {
                  *reinterpret_cast<uint4*>(g + 8 * j) = v;
                } else {   // 2..4 key tiles: accumulate straight into the (pre-zeroed) bf16 dQ, 8 elements per red
                  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(g + 8 * j), "r"(v.x), "r"(v.y),
                               "r"(v.z), "r"(v.w)
                               : "memory");
                }
              }
            } else {
              float* g = p.dq32 + ((long long)b * p.nq + qrow) * ((long long)p.heads * HD) + h * HD + col;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g + 4 * j), "f"(__uint_as_float(acc[4 * j])),
                             "f"(__uint_as_float(acc[4 * j + 1])), "f"(__uint_as_float(acc[4 * j + 2])),
                             "f"(__uint_as_float(acc[4 * j + 3]))
                             : "memory");
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(dq_drained);
    }
    // ---- dK, dV: rows = keys (lanes), this warp's HD / 2 columns
    {
      __nv_bfloat16* dkp = p.dk + b * p.dk_bs + (long long)key * p.lddk + h * HD;
      __nv_bfloat16* dvp = p.dv + b * p.dv_bs + (long long)key * p.lddv + h * HD;
#pragma unroll 1
      for (int t = 0; t < 2; ++t) {
        __nv_bfloat16* g = t == 0 ? dkp : dvp;
        const uint32_t tcol = t == 0 ? C::DK_COL : C::DV_COL;
#pragma unroll 1
        for (int c = 0; c < HD / 64; ++c) {
          const int col = half * (HD / 2) + c * 32;
          uint32_t acc[32];
          if (nqt > 0) {
            tmem_ld16(t_row + tcol + col, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
            tmem_ld16(t_row + tcol + col + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
            tmem_ld_wait();
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0u;
          }
          if (key_ok) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(acc[8 * j]), __uint_as_float(acc[8 * j + 1]));
              v.y = pack_bf16x2(__uint_as_float(acc[8 * j + 2]), __uint_as_float(acc[8 * j + 3]));
              v.z = pack_bf16x2(__uint_as_float(acc[8 * j + 4]), __uint_as_float(acc[8 * j + 5]));
              v.w = pack_bf16x2(__uint_as_float(acc[8 * j + 6]), __uint_as_float(acc[8 * j + 7]));
              *reinterpret_cast<uint4*>(g + col + 8 * j) = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ================================================================================================ host side
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

static int g_legacy = -1;
static long long g_tc_launches = 0;
bool legacy_forced() {
  if (g_legacy < 0) { const char* e = getenv("OFK_ATTN_LEGACY"); g_legacy = (e && atoi(e) != 0) ? 1 : 0; }
  return g_legacy == 1;
}

static bool common_supported(const Args& a) {
  if (legacy_forced()) return false;
  if (a.hd != 64 && a.hd != 128) return false;
  if (a.batch <= 0 || a.heads <= 0 || a.nq <= 0 || a.nk <= 0) return false;
  if (a.batch > 65535 || a.heads > 65535) return false;
  // one flat 2-D tensor map per operand: batches must be stacked rows of the same [rows, heads * hd] view
  if (a.q_bs != (long long)a.nq * a.ldq || a.k_bs != (long long)a.nk * a.ldk || a.v_bs != (long long)a.nk * a.ldv) return false;
  if (!aligned16(a.q) || !aligned16(a.k) || !aligned16(a.v) || (a.ldq | a.ldk | a.ldv) % 8 != 0) return false;
  if ((long long)a.batch * a.nq >= (1LL << 31) || (long long)a.batch * a.nk >= (1LL << 31)) return false;
  if (!a.dense && a.mask_mode != 0 && (a.text_time == nullptr || a.kpm <= 0 || a.kpm % 16 != 0 || a.nk % a.kpm != 0)) return false;
  return true;
}
bool fwd_supported(const Args& a) {
  if (!common_supported(a)) return false;
  return aligned16(a.out) && a.ldo % 8 == 0 && a.o_bs % 8 == 0;
}
// dQ gets one partial per 128-key tile.  One tile: stored directly.  2..4 tiles (the LM at T_txt <= 512, 2..5 images of
// 64 latents): bf16 reductions (red.global.add.noftz.bf16x2) into the zero-filled dQ itself -- at most three extra bf16
// roundings, no scratch, no conversion pass.  More tiles (the 4160-key Perceiver isolation shape): fp32 accumulator.
constexpr int BF16_ATOMIC_MAX_KEYS = 512;
long long bwd_workspace_bytes(int batch, int heads, int hd, int nq, int nk) {
  if (nk <= BF16_ATOMIC_MAX_KEYS) return 0;
  return (long long)batch * nq * heads * hd * 4;
}
bool bwd_supported(const Args& a) {
  if (!common_supported(a)) return false;
  // The backward CTA owns 128 KEYS and visits the query tiles that see them.  With at most 64 query rows per (batch,
  // head) -- the 64 Perceiver latents, a decode step -- there is a single, half-empty query tile: every CTA runs one
  // iteration and is all prologue / epilogue (measured on the 4160-key isolation shape: 1164 us against 778 us for the
  // 64-row mma.sync kernels, profiles/r02_attention_by_shape.md).  Those problems keep the mma.sync backward.
  if (a.nq <= 64) return false;
  if (a.o_bs != (long long)a.nq * a.ldo || !aligned16(a.d_o) || a.ldo % 8 != 0) return false;
  if (!aligned16(a.dq) || !aligned16(a.dk) || !aligned16(a.dv)) return false;
  if ((a.lddq | a.lddk | a.lddv | a.dq_bs | a.dk_bs | a.dv_bs) % 8 != 0) return false;
  if ((a.nq + 127) / 128 > BwdCfg<64>::MAXQT) return false;
  if (a.nk > 128 && a.nk <= BF16_ATOMIC_MAX_KEYS && a.dq_bs != (long long)a.nq * a.lddq) return false;   // 2-D memset of dQ
  const long long need = bwd_workspace_bytes(a.batch, a.heads, a.hd, a.nq, a.nk);
  if (need > 0 && (a.workspace == nullptr || a.workspace_bytes < need || !aligned16(a.workspace))) return false;
  return true;
}

static void fill_params(const Args& a, Params& p) {
  p.out = (__nv_bfloat16*)a.out; p.dq = (__nv_bfloat16*)a.dq; p.dk = (__nv_bfloat16*)a.dk; p.dv = (__nv_bfloat16*)a.dv;
  p.lse = a.lse; p.delta = a.delta; p.dq32 = (float*)a.workspace;
  p.text_time = a.text_time; p.mask = a.mask; p.slopes = a.slopes; p.pure_causal = a.pure_causal;
  p.batch = a.batch; p.heads = a.heads; p.nq = a.nq; p.nk = a.nk;
  p.o_bs = a.o_bs; p.ldo = a.ldo; p.dq_bs = a.dq_bs; p.lddq = a.lddq; p.dk_bs = a.dk_bs; p.lddk = a.lddk;
  p.dv_bs = a.dv_bs; p.lddv = a.lddv; p.scale = a.scale;
  p.mask_mode = a.dense ? 0 : a.mask_mode; p.kpm = a.kpm > 0 ? a.kpm : 64; p.causal = a.causal; p.dq_direct = 0;
}

template <int HD, bool DENSE>
static int launch_fwd(const Args& a) {
  using C = Fwd2Cfg<HD>;
  CUtensorMap tq, tk, tv;
  int rc = ofk_tensor_map_bf16(a.q, a.ldq, a.batch * a.nq, a.heads * HD, 64, 128, &tq);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(a.k, a.ldk, a.batch * a.nk, a.heads * HD, 64, C::BKT, &tk);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(a.v, a.ldv, a.batch * a.nk, a.heads * HD, 64, C::BKT, &tv);
  if (rc) return rc;
  auto kern = attn_fwd2_tc_kernel<HD, DENSE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  Params p;
  fill_params(a, p);
  int q_tiles = (a.nq + 127) / 128;
  if (a.q_tile_limit > 0 && a.q_tile_limit < q_tiles) q_tiles = a.q_tile_limit;
  dim3 grid(q_tiles, a.heads, a.batch);
  kern<<<grid, FWD_THREADS, C::SMEM, (cudaStream_t)a.stream>>>(tq, tk, tv, p);
  OFK_CHECK_LAUNCH();
  ++g_tc_launches;
  return 0;
}

int fwd(const Args& a) {
  if (a.hd == 64) return a.dense ? launch_fwd<64, true>(a) : launch_fwd<64, false>(a);
  return a.dense ? launch_fwd<128, true>(a) : launch_fwd<128, false>(a);
}

template <int HD, bool DENSE>
static int launch_bwd(const Args& a) {
  using C = BwdCfg<HD>;
  cudaStream_t stream = (cudaStream_t)a.stream;
  CUtensorMap tq, tk, tv, tdo;
  int rc = ofk_tensor_map_bf16(a.q, a.ldq, a.batch * a.nq, a.heads * HD, 64, 128, &tq);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(a.k, a.ldk, a.batch * a.nk, a.heads * HD, 64, 128, &tk);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(a.v, a.ldv, a.batch * a.nk, a.heads * HD, 64, 128, &tv);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(a.d_o, a.ldo, a.batch * a.nq, a.heads * HD, 64, 128, &tdo);
  if (rc) return rc;
  auto kern = attn_bwd_tc_kernel<HD, DENSE>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    if (e != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  Params p;
  fill_params(a, p);
  const long long rows = (long long)a.batch * a.heads * a.nq;
  delta_kernel<HD><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>((const __nv_bfloat16*)a.o, (const __nv_bfloat16*)a.d_o,
                                                                   a.delta, a.batch, a.heads, a.nq, a.o_bs, a.ldo);
  OFK_CHECK_LAUNCH();
  const long long ws = bwd_workspace_bytes(a.batch, a.heads, HD, a.nq, a.nk);
  p.dq_direct = a.nk <= 128 ? 1 : (ws == 0 ? 2 : 0);
  if (p.dq_direct == 2) {
    cudaError_t e = cudaMemset2DAsync(a.dq, (size_t)a.lddq * 2, 0, (size_t)a.heads * HD * 2, (size_t)a.batch * a.nq, stream);
    if (e != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e));
  } else if (ws > 0) {
    cudaError_t e = cudaMemsetAsync(a.workspace, 0, (size_t)ws, stream);
    if (e != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e));
  }
  dim3 grid((a.nk + 127) / 128, a.heads, a.batch);
  kern<<<grid, BWD_THREADS, C::SMEM, stream>>>(tq, tk, tv, tdo, p);
  OFK_CHECK_LAUNCH();
  ++g_tc_launches;
  if (ws > 0) {
    const int cols = a.heads * HD;
    const long long total8 = (long long)a.batch * a.nq * (cols / 8);
    dq_convert_kernel<<<(unsigned)((total8 + 255) / 256), 256, 0, stream>>>((const float*)a.workspace, (__nv_bfloat16*)a.dq,
                                                                            a.nq, cols, a.dq_bs, a.lddq, total8);
    OFK_CHECK_LAUNCH();
  }
  return 0;
}

int bwd(const Args& a) {
  if (a.hd == 64) return a.dense ? launch_bwd<64, true>(a) : launch_bwd<64, false>(a);
  return a.dense ? launch_bwd<128, true>(a) : launch_bwd<128, false>(a);
}

}  // namespace tc
}  // namespace ofk

extern "C" long long ofk_attn_bwd_workspace_bytes(int batch, int heads, int head_dim, int nq, int nk) {
  return ofk::tc::bwd_workspace_bytes(batch, heads, head_dim, nq, nk);
}
extern "C" int ofk_attn_force_legacy(int on) {
  const int prev = ofk::tc::legacy_forced() ? 1 : 0;
  ofk::tc::g_legacy = on ? 1 : 0;
  return prev;
}
extern "C" long long ofk_attn_tc_launch_count(void) { return ofk::tc::g_tc_launches; }
