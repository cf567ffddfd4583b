// EXPERIMENTAL (round-1 preparation, NOT yet validated on hardware, not on any default path): the attention
// forward core of attention.cu (head_dim 64: gated masked cross-attention, Perceiver attention, ViT attention) on
// the 5th-generation tensor cores -- what BASELINE.json's north_star asks for ("TMA staging of Q/K/V tiles ...,
// tcgen05 tensor-core MMA for the QK^T and PV contractions").  Same semantics, arguments and outputs as
// ofk_attn_fwd (helpers.py:55-64, :190-232; see attention.cu for the mask rules); reached only through
// ofk_attn_fwd_tc, which ops.attn_fwd calls when OFK_ATTN_TC=1.
//
// One CTA = one (batch, head, 128-query tile); 2 CTAs per SM overlap each other's phases.
//   warp 4 (one thread): TMA producer -- Q tile once, then one K tile + one V tile (128 keys x 64, 16 KiB each,
//                        128B swizzle) per key step; also allocates TMEM (256 columns: S = 0..127, O_j = 128..191)
//   warp 5 (one thread): MMA issuer   -- S = Q K_j^T   (M 128, N 128, K 64 : 4 x tcgen05.mma, K-major A and B)
//                                        O_j = P_j V_j (M 128, N 64, K 128 : 8 x tcgen05.mma, A = P from smem,
//                                        B = V_j through the MN-major descriptor the dgrad GEMMs use)
//   warps 0-3: softmax, lane = query row (TMEM lane), so row max / row sum need no shuffles:
//              pass 1 over S (tcgen05.ld) -> masked, scaled row max; pass 2 -> p = exp2(s - m), row sum, P as bf16
//              into shared memory in the canonical K-major SW128 layout (2 blocks of 64 keys); then
//              o = o * corr + O_j with the running output held in 64 registers (no TMEM read-modify-write).
// Barriers (mbarrier, one phase per key step): kv_full (TMA bytes) -> s_full (commit) -> p_full (128 softmax
// threads) -> o_full (commit; the same commit also frees K/V: kv_empty) -> o_empty (128 threads).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ofk_internal.h"
#include "ofk_ptx.cuh"

namespace ofk {
namespace tc {

constexpr int BQ = 128, BK = 128, HD = 64;
constexpr int THREADS = 192;                       // 4 softmax warps + TMA warp + MMA warp
constexpr int TILE_BYTES = 128 * 128;              // 128 rows x 64 bf16
constexpr int P_BYTES = 2 * TILE_BYTES;            // 128 rows x 128 keys bf16 = two K-major blocks of 64 keys
constexpr int SMEM_BYTES = 3 * TILE_BYTES + P_BYTES + 256 + 1024;
constexpr uint32_t TMEM_COLS = 256;
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
  __nv_bfloat16* out;
  float* lse;
  const int* text_time;
  int batch, heads, nq, nk;
  long long o_bs, ldo;
  float scale;
  int mask_mode, kpm;
};

// kind 0: normal masked row; 1: zero row; 2: uniform row; 3: unmasked (see attention.cu::classify_row)
struct Row { int tt, kind; };
__device__ __forceinline__ Row classify(const Params& p, int b, int row) {
  Row r; r.tt = 0; r.kind = 3;
  if (p.mask_mode == 0) return r;
  if (row >= p.nq) { r.kind = 1; return r; }
  const int tt = p.text_time[(long long)b * p.nq + row];
  const int n_media = p.nk / p.kpm;
  r.tt = tt;
  const bool has = (p.mask_mode == 1) ? (tt >= 1 && tt <= n_media) : (tt >= 1);
  if (has) r.kind = 0;
  else if (p.mask_mode == 1 && tt == 0) r.kind = 1;
  else r.kind = 2;
  return r;
}
__device__ __forceinline__ bool allowed(const Params& p, const Row& r, int media) {
  if (r.kind >= 2) return true;
  if (r.kind == 1) return false;
  return p.mask_mode == 1 ? (r.tt == media) : (r.tt >= media);
}

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                   const __grid_constant__ CUtensorMap tma_v, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 2 * TILE_BYTES;
  uint8_t* sP = smem + 3 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * TILE_BYTES + P_BYTES);
  uint64_t* kv_full = bars + 0;
  uint64_t* kv_empty = bars + 1;
  uint64_t* s_full = bars + 2;
  uint64_t* p_full = bars + 3;
  uint64_t* o_full = bars + 4;
  uint64_t* o_empty = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);
  int* s_range = reinterpret_cast<int*>(bars + 10);   // [0] min tt, [1] max tt, [2] any uniform, [3] any normal/unmasked

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) { s_range[0] = 1 << 30; s_range[1] = -1; s_range[2] = 0; s_range[3] = 0; }
  if (warp == 5 && lane == 0) {
    mbar_init(kv_full, 1); mbar_init(kv_empty, 1); mbar_init(s_full, 1);
    mbar_init(p_full, 128); mbar_init(o_full, 1); mbar_init(o_empty, 128);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) { tma_prefetch_desc(&tma_q); tma_prefetch_desc(&tma_k); tma_prefetch_desc(&tma_v); }
    __syncwarp();
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  __syncthreads();

  // ---- key range this query tile needs (block-uniform): the same rule as attention.cu::block_key_range
  Row my; my.tt = 0; my.kind = 1;
  if (warp < 4) {
    my = classify(p, b, q0 + threadIdx.x);
    if (my.kind == 0) { atomicMin(&s_range[0], my.tt); atomicMax(&s_range[1], my.tt); }
    if (my.kind == 2) s_range[2] = 1;
    if (my.kind == 0 || my.kind == 3) s_range[3] = 1;
    if (q0 + (int)threadIdx.x >= p.nq) my.kind = 1;          // tile tail: nothing to compute, nothing to store
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  int lo = 0, hi = p.nk;
  if (p.mask_mode != 0 && s_range[2] == 0) {
    if (s_range[1] < 0) { lo = 0; hi = 0; }                   // only zero rows
    else {
      hi = min(p.nk, s_range[1] * p.kpm);
      lo = p.mask_mode == 1 ? max(0, (s_range[0] - 1) * p.kpm) : 0;
    }
  }
  const int t_lo = lo / BK, t_hi = (hi + BK - 1) / BK;       // key steps [t_lo, t_hi)
  const int nsteps = max(0, t_hi - t_lo);

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int it = 0; it < nsteps; ++it) {
        if (it > 0) mbar_wait(kv_empty, (it - 1) & 1);
        const int key0 = (t_lo + it) * BK;
        mbar_arrive_expect_tx(kv_full, (it == 0 ? 3 : 2) * TILE_BYTES);
        if (it == 0) tma_load_2d(sQ, &tma_q, kv_full, h * HD, b * p.nq + q0);
        tma_load_2d(sK, &tma_k, kv_full, h * HD, b * p.nk + key0);
        tma_load_2d(sV, &tma_v, kv_full, h * HD, b * p.nk + key0);
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);
      const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ap = smem_u32(sP);
      for (int it = 0; it < nsteps; ++it) {
        mbar_wait(kv_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)       // S = Q K^T : both operands K-major, 32 bytes per 16-wide k step
          umma_bf16(tmem_base, make_smem_desc_sw128(aq + k * 32, 0, 1024), make_smem_desc_sw128(ak + k * 32, 0, 1024),
                    idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, it & 1);              // P_j is in shared memory (and S_j has been consumed)
        if (it > 0) mbar_wait(o_empty, (it - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)       // O_j = P V : A = P (K-major, block k/4), B = V (MN-major, 16 key rows / step)
          umma_bf16(tmem_base + 128, make_smem_desc_sw128(ap + (k >> 2) * TILE_BYTES + (k & 3) * 32, 0, 1024),
                    make_smem_desc_sw128(av + k * 2048, 64 * 128, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(o_full);
        umma_commit(kv_empty);
      }
    }
    __syncwarp();
  } else {
    // ===================== softmax / output (lane = query row) =====================
    const int row = q0 + threadIdx.x;                         // threadIdx.x in [0, 128)
    const uint32_t t_row = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float z = my.kind == 2 ? 0.f : p.scale * LOG2E;     // uniform rows: S = 0 over every key
    float o[HD];
#pragma unroll
    for (int i = 0; i < HD; ++i) o[i] = 0.f;
    float m = -INFINITY, l = 0.f;
    const uint32_t p_row = smem_u32(sP) + threadIdx.x * 128;  // this row inside a K-major block (8-row groups of 1 KiB)
    const int sw = threadIdx.x & 7;
    for (int it = 0; it < nsteps; ++it) {
      const int key0 = (t_lo + it) * BK;
      mbar_wait(s_full, it & 1);
      tc_fence_after();
      // ---- pass 1: masked / scaled row maximum
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < BK / 32; ++c) {
        uint32_t acc[32];
        tmem_ld16(t_row + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        tmem_ld16(t_row + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {                          // kpm % 16 == 0: one media per 16 keys
          const int kg = key0 + c * 32 + g * 16;
          const bool ok = my.kind != 1 && (p.mask_mode == 0 || allowed(p, my, kg / p.kpm + 1));
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float s = (ok && kg + i < p.nk) ? __uint_as_float(acc[g * 16 + i]) * z : -INFINITY;
            mx = fmaxf(mx, s);
          }
        }
      }
      const float m_new = fmaxf(m, mx);
      const float sub = m_new == -INFINITY ? 0.f : m_new;
      const float corr = ex2_approx(m - sub);                  // m = -inf -> 0
      m = m_new;
      // ---- pass 2: p = exp2(s - m), row sum, P (bf16) into the K-major SW128 layout
      float rs = 0.f;
#pragma unroll 1
      for (int c = 0; c < BK / 32; ++c) {
        uint32_t acc[32];
        tmem_ld16(t_row + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        tmem_ld16(t_row + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int kg = key0 + c * 32 + g * 16;
          const bool ok = my.kind != 1 && (p.mask_mode == 0 || allowed(p, my, kg / p.kpm + 1));
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            const float s0 = (ok && kg + i < p.nk) ? __uint_as_float(acc[g * 16 + i]) * z : -INFINITY;
            const float s1 = (ok && kg + i + 1 < p.nk) ? __uint_as_float(acc[g * 16 + i + 1]) * z : -INFINITY;
            const float p0 = ex2_approx(s0 - sub), p1 = ex2_approx(s1 - sub);
            rs += p0 + p1;
            w[(g * 16 + i) >> 1] = pack_bf16x2(p0, p1);
          }
        }
        // 32 keys = 64 bytes = four 16-byte chunks; chunk index inside the 64-key block: (c & 1) * 4 + j
        const uint32_t blk = p_row + (c >> 1) * TILE_BYTES;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int chunk = (c & 1) * 4 + j;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(blk + ((chunk ^ sw) << 4)), "r"(w[4 * j]),
                       "r"(w[4 * j + 1]), "r"(w[4 * j + 2]), "r"(w[4 * j + 3]) : "memory");
        }
      }
      l = l * corr + rs;
      fence_proxy_async_smem();                                // generic-proxy writes of P -> visible to the MMA
      tc_fence_before();
      mbar_arrive(p_full);
      // ---- o = o * corr + O_j
      mbar_wait(o_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t acc[32];
        tmem_ld16(t_row + 128 + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&acc[0]));
        tmem_ld16(t_row + 128 + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&acc[16]));
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c * 32 + i] = fmaf(o[c * 32 + i], corr, __uint_as_float(acc[i]));
      }
      tc_fence_before();
      mbar_arrive(o_empty);
    }
    // ---- normalise and store this row (64 bf16 = 128 contiguous bytes); log2-domain LSE as in attention.cu
    if (row < p.nq) {
      const float inv = l > 0.f ? 1.f / l : 0.f;
      __nv_bfloat16* op = p.out + b * p.o_bs + (long long)row * p.ldo + h * HD;
#pragma unroll
      for (int j = 0; j < HD / 8; ++j) {
        uint4 v;
        v.x = pack_bf16x2(o[8 * j] * inv, o[8 * j + 1] * inv);
        v.y = pack_bf16x2(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
        v.z = pack_bf16x2(o[8 * j + 4] * inv, o[8 * j + 5] * inv);
        v.w = pack_bf16x2(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
        *reinterpret_cast<uint4*>(op + 8 * j) = v;
      }
      if (p.lse != nullptr) p.lse[((long long)b * p.heads + h) * p.nq + row] = l > 0.f ? m + log2f(l) : 0.f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace tc
}  // namespace ofk

extern "C" int ofk_attn_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse, int batch, int heads,
                               int nq, int nk, long long q_bstride, long long ldq, long long k_bstride, long long ldk,
                               long long v_bstride, long long ldv, long long o_bstride, long long ldo, float scale,
                               int mask_mode, const int* text_time, int keys_per_media, void* stream_) {
  using namespace ofk::tc;
  if (!q || !k || !v || !o) return ofk_set_error(OFK_ERR_ARG, "attention(tc): null pointer");
  if (batch <= 0 || heads <= 0 || nq <= 0 || nk <= 0) return ofk_set_error(OFK_ERR_ARG, "attention(tc): empty problem");
  if (mask_mode != 0 && (!text_time || keys_per_media <= 0 || keys_per_media % 16 != 0 || nk % keys_per_media != 0))
    return ofk_set_error(OFK_ERR_ARG, "attention(tc): media mask needs text_time and keys_per_media % 16 == 0 dividing nk");
  // one flat 2-D tensor map per operand: batches must be stacked rows of the same [rows, heads * 64] view
  if (q_bstride != (long long)nq * ldq || k_bstride != (long long)nk * ldk || v_bstride != (long long)nk * ldv)
    return ofk_set_error(OFK_ERR_ARG, "attention(tc): batch stride must equal rows * row stride");
  if ((ldo % 8) != 0 || (o_bstride % 8) != 0 || (reinterpret_cast<uintptr_t>(o) & 15))
    return ofk_set_error(OFK_ERR_ALIGN, "attention(tc): output rows must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  int rc = ofk_tensor_map_bf16(q, ldq, batch * nq, heads * HD, HD, 128, &tq);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(k, ldk, batch * nk, heads * HD, HD, 128, &tk);
  if (rc) return rc;
  rc = ofk_tensor_map_bf16(v, ldv, batch * nk, heads * HD, HD, 128, &tv);
  if (rc) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return ofk_set_error(OFK_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  Params p;
  p.out = (__nv_bfloat16*)o; p.lse = lse; p.text_time = text_time;
  p.batch = batch; p.heads = heads; p.nq = nq; p.nk = nk; p.o_bs = o_bstride; p.ldo = ldo;
  p.scale = scale; p.mask_mode = mask_mode; p.kpm = keys_per_media > 0 ? keys_per_media : 64;
  dim3 grid((nq + BQ - 1) / BQ, heads, batch);
  attn_fwd_tc_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream_>>>(tq, tk, tv, p);
  OFK_CHECK_LAUNCH();
  return 0;
}
