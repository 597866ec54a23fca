// dn_kvquant.cuh -- 8/4-bit affine KV cache (group 64 along head_dim), the reference API's DEFAULT kv_bits
// (src/dnet/api/models.py:316,342; built by src/dnet/utils/model.py:505-554 as mlx_lm QuantizedKVCache).
//
// Semantics restated from mlx / mlx_lm (not vendored in the reference: PARITY UNPINNED, oracle/llama_oracle.py
// `mlx_affine_quantize`, `LlamaOracle.sdpa_quantized`):
//   append    every new K / V row (post-RoPE, bf16) is quantised per group of 64 dims:
//               edge = bound with the larger magnitude; scale = max((max-min)/(2^b-1), 1e-7) signed so that
//               edge/scale >= 0; q0 = rint(edge/scale); if q0 != 0 { scale = edge/q0; bias = edge } else bias = 0
//               code = clamp(rint((w - bias)/scale), 0, 2^b-1)  with the fp32 scale/bias; scale, bias stored as bf16
//   attention mlx_lm quantized_scaled_dot_product_attention:
//               q' = bf16(q * hd^-1/2);  s_j = bf16(sum_d q'_d (scale*code + bias));  mask;
//               p = bf16(softmax_fp32(s));  out = bf16(sum_j p_j (scale*code + bias))
//             i.e. scores and probabilities are ROUNDED TO bf16 (quantized_matmul returns the input dtype), which an
//             online-softmax kernel cannot reproduce: the kernels below are two-pass (scores first, then P.V).
//
// Storage (replaces the bf16 page layout; one "unit" per (page, K|V, kv head)):
//   [PAGE rows x HD*BITS/8 bytes of codes][PAGE x 2 bf16 scales][PAGE x 2 bf16 biases]
//   8 bit: 8192 + 256 + 256 = 8704 B (bf16: 16384 B);  4 bit: 4096 + 512 = 4608 B.
// A lane owns dims 4*lane .. 4*lane+3 of a row (lanes 0-15 = group 0, 16-31 = group 1), so its 4 codes are one
// 32-bit word (8 bit) or one 16-bit word (4 bit) and a warp-wide load of a row is fully coalesced.
#pragma once
#include "dn_kernels.cuh"

namespace dn {

template <int BITS> struct KvQ {
  static constexpr int ROW_BYTES = HD * BITS / 8;                       // 128 / 64
  static constexpr int CODES_BYTES = PAGE * ROW_BYTES;                  // 8192 / 4096
  static constexpr int UNIT_BYTES = CODES_BYTES + PAGE * 2 * 2 * 2;     // + scales + biases (bf16 x 2 groups)
  static constexpr float NBINS = (float)((1 << BITS) - 1);
};
__host__ __device__ inline int kvq_unit_bytes(int bits) { return PAGE * (HD * bits / 8) + PAGE * 8; }

template <int BITS>
__device__ __forceinline__ const unsigned char* kvq_unit(const unsigned char* layer_pool, int page, int kv, int kvh, int n_kv) {
  return layer_pool + (((size_t)page * 2 + kv) * n_kv + kvh) * (size_t)KvQ<BITS>::UNIT_BYTES;
}
// this lane's 4 codes of row `tok` as floats
template <int BITS>
__device__ __forceinline__ void kvq_load_codes(const unsigned char* unit, int tok, int lane, float (&c)[4]) {
  if (BITS == 8) {
    const uint32_t w = __ldcg(reinterpret_cast<const uint32_t*>(unit + (size_t)tok * 128) + lane);
    c[0] = (float)(w & 0xffu); c[1] = (float)((w >> 8) & 0xffu); c[2] = (float)((w >> 16) & 0xffu); c[3] = (float)(w >> 24);
  } else {
    const uint32_t w = __ldcg(reinterpret_cast<const unsigned short*>(unit + (size_t)tok * 64) + lane);
    c[0] = (float)(w & 0xfu); c[1] = (float)((w >> 4) & 0xfu); c[2] = (float)((w >> 8) & 0xfu); c[3] = (float)((w >> 12) & 0xfu);
  }
}
// (scale, bias) of this lane's group for row `tok`
template <int BITS>
__device__ __forceinline__ void kvq_load_sb(const unsigned char* unit, int tok, int lane, float& sc, float& bi) {
  const unsigned short* s = reinterpret_cast<const unsigned short*>(unit + KvQ<BITS>::CODES_BYTES);
  const int g = lane >> 4;
  sc = __uint_as_float((uint32_t)__ldcg(s + tok * 2 + g) << 16);
  bi = __uint_as_float((uint32_t)__ldcg(s + PAGE * 2 + tok * 2 + g) << 16);
}

// Quantise one 128-dim row held as 4 values per lane (warp-collective).  Returns this lane's 4 codes and the
// bf16-rounded (scale, bias) of its group -- the values dequantisation uses.
template <int BITS>
__device__ __forceinline__ void kvq_quantise_row(const float (&w)[4], float (&code)[4], float& sc_bf, float& bi_bf) {
  float mx = fmaxf(fmaxf(w[0], w[1]), fmaxf(w[2], w[3]));
  float mn = fminf(fminf(w[0], w[1]), fminf(w[2], w[3]));
#pragma unroll
  for (int s = 8; s >= 1; s >>= 1) {       // lanes 0-15 and 16-31 reduce separately (xor stays inside a half warp)
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, s));
  }
  const bool side = fabsf(mn) > fabsf(mx);
  float scale = fmaxf(__fdiv_rn(__fsub_rn(mx, mn), KvQ<BITS>::NBINS), 1e-7f);
  scale = side ? scale : -scale;
  const float edge = side ? mn : mx;
  const float q0 = rintf(__fdiv_rn(edge, scale));
  float bias = 0.f;
  if (q0 != 0.f) { scale = __fdiv_rn(edge, q0); bias = edge; }
#pragma unroll
  for (int i = 0; i < 4; ++i) code[i] = fminf(fmaxf(rintf(__fdiv_rn(__fsub_rn(w[i], bias), scale)), 0.f), KvQ<BITS>::NBINS);
  sc_bf = bf16r(scale);
  bi_bf = bf16r(bias);
}
template <int BITS>
__device__ __forceinline__ void kvq_store_row(unsigned char* unit, int tok, int lane, const float (&code)[4], float sc_bf, float bi_bf) {
  if (BITS == 8) {
    const uint32_t wd = (uint32_t)code[0] | ((uint32_t)code[1] << 8) | ((uint32_t)code[2] << 16) | ((uint32_t)code[3] << 24);
    reinterpret_cast<uint32_t*>(unit + (size_t)tok * 128)[lane] = wd;
  } else {
    const uint32_t wd = (uint32_t)code[0] | ((uint32_t)code[1] << 4) | ((uint32_t)code[2] << 8) | ((uint32_t)code[3] << 12);
    reinterpret_cast<unsigned short*>(unit + (size_t)tok * 64)[lane] = (unsigned short)wd;
  }
  if ((lane & 15) == 0) {
    unsigned short* s = reinterpret_cast<unsigned short*>(unit + KvQ<BITS>::CODES_BYTES);
    const int g = lane >> 4;
    s[tok * 2 + g] = (unsigned short)(__float_as_uint(sc_bf) >> 16);
    s[PAGE * 2 + tok * 2 + g] = (unsigned short)(__float_as_uint(bi_bf) >> 16);
  }
}

// bf16 staging "pool" the unchanged append kernels write into (same page layout as the bf16 pool; logical page p
// lives in staging page p % KVQ_STAGE_PAGES -- a chunk of <= 512 tokens spans at most 9 logical pages)
constexpr int KVQ_STAGE_PAGES = 16;
__device__ __forceinline__ const bf16* kvq_stage_row(const bf16* stage, int pos, int kv, int kvh, int n_kv) {
  const int sp = (pos / PAGE) % KVQ_STAGE_PAGES;
  return stage + (((size_t)sp * 2 + kv) * n_kv + kvh) * (PAGE * HD) + (size_t)(pos % PAGE) * HD;
}
__device__ __forceinline__ void kvq_read_stage4(const bf16* row, int lane, float (&w)[4]) {
  const uint2 u = __ldcg(reinterpret_cast<const uint2*>(row + lane * 4));
  w[0] = bf_lo(u.x); w[1] = bf_hi(u.x); w[2] = bf_lo(u.y); w[3] = bf_hi(u.y);
}

// quantise + append T staged rows: grid (T, 2 * n_kv), one warp each
template <int BITS>
__global__ void __launch_bounds__(32) k_kv_quant_append(const bf16* __restrict__ stage, unsigned char* __restrict__ layer_pool,
                                                        const int32_t* __restrict__ block_table, const StepState* __restrict__ st,
                                                        int n_kv) {
  const int t = blockIdx.x, kv = blockIdx.y / n_kv, kvh = blockIdx.y % n_kv, lane = threadIdx.x;
  const int pos = st->pos + t;
  float w[4], code[4], sc, bi;
  kvq_read_stage4(kvq_stage_row(stage, pos, kv, kvh, n_kv), lane, w);
  kvq_quantise_row<BITS>(w, code, sc, bi);
  unsigned char* unit = const_cast<unsigned char*>(kvq_unit<BITS>(layer_pool, block_table[pos / PAGE], kv, kvh, n_kv));
  kvq_store_row<BITS>(unit, pos % PAGE, lane, code, sc, bi);
}

// ---------------------------------------------------------------------------------
// generic quantised attention for T >= 1 (per-op decode path, small chunks, prefill chunks): CTA = (q head, query
// token), 4 warps take 32-key tiles round-robin.  Two passes: (A) bf16-rounded scores into shared memory + max,
// (B) p = bf16(exp(s - M) / L), out = sum_j p_j (scale*code + bias).  Every sum has a fixed order -> deterministic.
// Shared memory: kv_len floats (+ small scratch) -> contexts up to ~48K tokens.
// ---------------------------------------------------------------------------------
constexpr int AQ_WARPS = 4;
template <int BITS>
__global__ void __launch_bounds__(AQ_WARPS * 32) k_attn_q(const bf16* __restrict__ q, const unsigned char* __restrict__ layer_pool,
                                                         const int32_t* __restrict__ block_table, const StepState* __restrict__ st,
                                                         bf16* __restrict__ out, int n_heads, int n_kv) {
  extern __shared__ float aq_smem[];
  const int head = blockIdx.x, t = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = n_heads / n_kv, kvh = head / G;
  const int kv_len = st->pos + t + 1;                 // causal: query t sees positions <= pos0 + t
  const int n_tiles = (kv_len + 31) >> 5;
  float* sc_s = aq_smem;                              // [n_tiles * 32]
  float* red = aq_smem + n_tiles * 32;                // [AQ_WARPS][132]
  const float scale = 0.08838834764831845f;
  float qv[4], qsum;
  {
    const uint2 u = *reinterpret_cast<const uint2*>(q + (size_t)t * n_heads * HD + head * HD + lane * 4);
    qv[0] = bf16r(__fmul_rn(bf_lo(u.x), scale)); qv[1] = bf16r(__fmul_rn(bf_hi(u.x), scale));     // queries *= scale, in bf16
    qv[2] = bf16r(__fmul_rn(bf_lo(u.y), scale)); qv[3] = bf16r(__fmul_rn(bf_hi(u.y), scale));
    qsum = (qv[0] + qv[1]) + (qv[2] + qv[3]);
  }
  // ---- pass A: scores
  float m = -INFINITY;
  for (int tile = warp; tile < n_tiles; tile += AQ_WARPS) {
    const int t0 = tile << 5, nt = min(32, kv_len - t0);
    const unsigned char* ku = kvq_unit<BITS>(layer_pool, block_table[t0 / PAGE], 0, kvh, n_kv);
    float sc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      sc[j] = 0.f;
      if (j < nt) {
        float c[4], s1, b1;
        kvq_load_codes<BITS>(ku, (t0 % PAGE) + j, lane, c);
        kvq_load_sb<BITS>(ku, (t0 % PAGE) + j, lane, s1, b1);
        const float dot = fmaf(qv[0], c[0], fmaf(qv[1], c[1], fmaf(qv[2], c[2], qv[3] * c[3])));
        sc[j] = fmaf(s1, dot, b1 * qsum);
      }
    }
    transpose_reduce32(sc, lane);
    const float s = (lane < nt) ? bf16r(sc[0]) : -INFINITY;
    sc_s[t0 + lane] = s;
    m = fmaxf(m, warp_max(s));
  }
  if (lane == 0) red[warp] = m;
  __syncthreads();
  float M = red[0];
#pragma unroll
  for (int w = 1; w < AQ_WARPS; ++w) M = fmaxf(M, red[w]);
  __syncthreads();
  // denominator: each warp sums its tiles (same tiles as pass A), then the warps in fixed order
  float l = 0.f;
  for (int tile = warp; tile < n_tiles; tile += AQ_WARPS) {
    const float s = sc_s[(tile << 5) + lane];
    l += warp_sum(s == -INFINITY ? 0.f : exp2f((s - M) * LOG2E));
  }
  if (lane == 0) red[warp] = l;
  __syncthreads();
  float Lsum = 0.f;
#pragma unroll
  for (int w = 0; w < AQ_WARPS; ++w) Lsum += red[w];
  const float invL = 1.0f / Lsum;
  __syncthreads();
  // ---- pass B: P.V with bf16 probabilities
  float o[4] = {0.f, 0.f, 0.f, 0.f}, ob = 0.f;
  for (int tile = warp; tile < n_tiles; tile += AQ_WARPS) {
    const int t0 = tile << 5, nt = min(32, kv_len - t0);
    const unsigned char* vu = kvq_unit<BITS>(layer_pool, block_table[t0 / PAGE], 1, kvh, n_kv);
    const float sl = sc_s[t0 + lane];
    const float pl = (lane < nt) ? bf16r(exp2f((sl - M) * LOG2E) * invL) : 0.f;
    for (int j = 0; j < nt; ++j) {
      const float pj = __shfl_sync(0xffffffffu, pl, j);
      float c[4], s1, b1;
      kvq_load_codes<BITS>(vu, (t0 % PAGE) + j, lane, c);
      kvq_load_sb<BITS>(vu, (t0 % PAGE) + j, lane, s1, b1);
      const float ws = pj * s1;
      o[0] = fmaf(ws, c[0], o[0]); o[1] = fmaf(ws, c[1], o[1]); o[2] = fmaf(ws, c[2], o[2]); o[3] = fmaf(ws, c[3], o[3]);
      ob = fmaf(pj, b1, ob);
    }
  }
  float* wr = red + warp * 132;
  __syncthreads();
  *reinterpret_cast<float4*>(wr + lane * 4) = make_float4(o[0] + ob, o[1] + ob, o[2] + ob, o[3] + ob);
  __syncthreads();
  if (threadIdx.x < HD) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < AQ_WARPS; ++w) acc += red[w * 132 + threadIdx.x];
    out[(size_t)t * n_heads * HD + head * HD + threadIdx.x] = __float2bfloat16_rn(acc);
  }
}

}  // namespace dn
