// dn_kernels.cuh -- sm_100a device code for the dnet shard decode forward.
//
// Every kernel here is HBM-bandwidth bound integer/float streaming work (decode GEMV,
// arithmetic intensity ~1 FLOP/B), so the design rules are coalescing, loads in
// flight, grid = multiple of the SM count and zero redundant DRAM traffic -- not
// tensor cores.  Shapes follow mlx_lm.models.llama as instantiated by the reference
// (core/models/llama.py:33-46); rounding points follow oracle/llama_oracle.py.
//
// "slice GEMV" layout used by all five weight-streaming kernels:
//   * grid = CTAS_PER_SM x #SMs persistent-style CTAs; CTA b owns a contiguous range of
//     output rows (units of ALIGN rows), so every SM streams the same number of bytes
//     (+-1 row) regardless of the matrix height;
//   * inside a CTA, warp w owns a contiguous K-slice (chunks of 256 elements = one
//     16-byte load per lane); it walks the CTA's rows in blocks of 32/T rows, keeping
//     32 fp32 partial sums per lane, so the activation chunk is read from shared memory
//     once per 32 rows and weights are the only global traffic;
//   * a 31-shuffle transpose-reduction leaves lane i with the warp total of value i, the
//     8 warp partials are summed through shared memory in fixed order (deterministic),
//     and warp 0 runs the fused epilogue (RoPE + paged-KV append, residual add,
//     SwiGLU, argmax/logsumexp partials).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dn {

typedef __nv_bfloat16 bf16;

constexpr int NW = 8;                  // warps per GEMV CTA
constexpr int GEMV_THREADS = NW * 32;
constexpr int CTAS_PER_SM = 2;
constexpr int PAGE = 64;               // tokens per KV page
constexpr int HD = 128;                // head_dim (Llama-3 / Qwen2.5 / Mixtral)
constexpr int PART_STRIDE = 132;       // attention partial: 128 o + m + l (+pad)
constexpr float LOG2E = 1.4426950408889634f;

struct StepState {   // per-nonce device state read by the kernels -> graphs replay unchanged
  int32_t pos;       // cache.offset: tokens already in the KV (mlx_lm KVCache.offset)
  int32_t token;     // last sampled token (input of the next embed)
  int32_t pad[2];
};

// ---------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float bf16r(float x) {  // round to storage dtype and back
  return __bfloat162float(__float2bfloat16_rn(x));
}
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// streaming weight load: read-only path, do not allocate in L1 (weights are read once)
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// Programmatic dependent launch: the next kernel's CTAs may start (and prefetch weights,
// which no kernel ever writes) while this one drains; they block here before touching
// anything a predecessor produced.  No-ops when launched without the PDL attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, s));
  return v;
}

// 32 values per lane -> lane i ends with (sum over lanes of value i) in v[0]. 31 shuffles.
__device__ __forceinline__ void transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int i = 0; i < s; ++i) {
      const float keep = up ? v[i + s] : v[i];
      const float send = up ? v[i] : v[i + s];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
}

__device__ __forceinline__ float dot8(const uint4& w, const float (&x)[8], float acc) {
  acc = fmaf(bf_lo(w.x), x[0], acc);
  acc = fmaf(bf_hi(w.x), x[1], acc);
  acc = fmaf(bf_lo(w.y), x[2], acc);
  acc = fmaf(bf_hi(w.y), x[3], acc);
  acc = fmaf(bf_lo(w.z), x[4], acc);
  acc = fmaf(bf_hi(w.z), x[5], acc);
  acc = fmaf(bf_lo(w.w), x[6], acc);
  acc = fmaf(bf_hi(w.w), x[7], acc);
  return acc;
}

// ---------------------------------------------------------------------------------
// prologues: stage the (normalised) activation as bf16 in shared memory  xs[T][K]
// ---------------------------------------------------------------------------------
// plain copy of a bf16 activation [T][K]
template <int T>
__device__ __forceinline__ void stage_copy(bf16* xs, const bf16* __restrict__ src, int K) {
  const int n16 = T * K / 8;
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(xs);
  for (int i = threadIdx.x; i < n16; i += GEMV_THREADS) d4[i] = s4[i];
  __syncthreads();
}

// mx.fast.rms_norm: fp32 math, y = T(T(x * rsqrt(mean(x^2) + eps)) * w)  (oracle rms_norm)
// scratch: NW floats of shared memory
template <int T>
__device__ __forceinline__ void stage_rmsnorm(bf16* xs, float* scratch, const bf16* __restrict__ src,
                                              const bf16* __restrict__ w, int K, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int t = 0; t < T; ++t) {
    const bf16* x = src + (size_t)t * K;
    float ss = 0.f;
    for (int i = threadIdx.x * 8; i < K; i += GEMV_THREADS * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + i);
      const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y),
                          bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
    ss = warp_sum(ss);
    __syncthreads();  // scratch reuse across t
    if (lane == 0) scratch[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += scratch[i];
    const float inv = 1.0f / sqrtf(tot / (float)K + eps);
    for (int i = threadIdx.x * 8; i < K; i += GEMV_THREADS * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(x + i);
      const uint4 g = *reinterpret_cast<const uint4*>(w + i);
      const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y),
                          bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
      const float gw[8] = {bf_lo(g.x), bf_hi(g.x), bf_lo(g.y), bf_hi(g.y),
                           bf_lo(g.z), bf_hi(g.z), bf_lo(g.w), bf_hi(g.w)};
      __align__(16) bf16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        o[j] = __float2bfloat16_rn(__fmul_rn(bf16r(__fmul_rn(f[j], inv)), gw[j]));
      *reinterpret_cast<uint4*>(xs + (size_t)t * K + i) = *reinterpret_cast<const uint4*>(o);
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------
// the slice-GEMV skeleton
// ---------------------------------------------------------------------------------
// Op contract:
//   static constexpr int ALIGN;        rows are partitioned over CTAs in units of ALIGN
//   int K;  int nrows;                 reduction length (multiple of 256), virtual rows
//   __device__ const bf16* row(int r)  weight row pointer of virtual row r
//   __device__ void prologue<T>(bf16* xs, float* scratch, float* ost)
//   __device__ void epilogue<T>(int rb, int nv, int lane, float v, float* ost)  (warp 0 only;
//       lane = r_local*T + t; all 32 lanes call it, including invalid ones)
//   __device__ void finish(float* ost)  once per CTA after the row loop (all threads)
//   ost: 96 floats of shared memory private to the op (running argmax state of the head)
template <int T, class Op>
__global__ void __launch_bounds__(GEMV_THREADS, CTAS_PER_SM) k_gemv(Op op, int l2_prefetch_bytes) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = op.K;
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  float* red = reinterpret_cast<float*>(smem_raw + (size_t)T * K * sizeof(bf16));  // [2][NW][32]
  float* scratch = red + 2 * NW * 32;                                               // [NW]
  float* ost = scratch + NW;                                                        // [96] op state
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NROW = 32 / T;

  const int units = op.nrows / Op::ALIGN;
  const int r0 = (int)(((long long)units * blockIdx.x) / gridDim.x) * Op::ALIGN;
  const int r1 = (int)(((long long)units * (blockIdx.x + 1)) / gridDim.x) * Op::ALIGN;
  const int nchunks = K >> 8;
  const int c0 = (nchunks * warp) / NW, c1 = (nchunks * (warp + 1)) / NW;

  pdl_launch_dependents();
  // Weights are immutable, so pull this CTA's first rows toward L2 while the predecessor
  // kernel is still draining (PDL) and while the prologue below runs.
  if (l2_prefetch_bytes > 0) {
    const int rowbytes = K * 2;
    int budget_rows = l2_prefetch_bytes / rowbytes;
    if (budget_rows > r1 - r0) budget_rows = r1 - r0;
    const int lines_per_row = rowbytes >> 7;
    for (int r = warp; r < budget_rows; r += NW) {
      const char* base = reinterpret_cast<const char*>(op.row(r0 + r));
      for (int ln = lane; ln < lines_per_row; ln += 32) prefetch_l2(base + ((size_t)ln << 7));
    }
  }
  pdl_wait();
  op.template prologue<T>(xs, scratch, ost);

  int buf = 0;
#pragma unroll 1
  for (int rb = r0; rb < r1; rb += NROW, buf ^= 1) {
    const int nv = min(NROW, r1 - rb);
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int c = c0; c < c1; ++c) {
      const int koff = (c << 8) + (lane << 3);
      float xf[T][8];
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)t * K + koff);
        xf[t][0] = bf_lo(xv.x); xf[t][1] = bf_hi(xv.x); xf[t][2] = bf_lo(xv.y); xf[t][3] = bf_hi(xv.y);
        xf[t][4] = bf_lo(xv.z); xf[t][5] = bf_hi(xv.z); xf[t][6] = bf_lo(xv.w); xf[t][7] = bf_hi(xv.w);
      }
      constexpr int BATCH = (NROW < 8) ? NROW : 8;
#pragma unroll
      for (int b0 = 0; b0 < NROW; b0 += BATCH) {
        uint4 wv[BATCH];
#pragma unroll
        for (int b = 0; b < BATCH; ++b) {
          wv[b] = make_uint4(0u, 0u, 0u, 0u);
          if (b0 + b < nv) wv[b] = ldg_stream(op.row(rb + b0 + b) + koff);
        }
#pragma unroll
        for (int b = 0; b < BATCH; ++b)
#pragma unroll
          for (int t = 0; t < T; ++t) acc[(b0 + b) * T + t] = dot8(wv[b], xf[t], acc[(b0 + b) * T + t]);
      }
    }
    transpose_reduce32(acc, lane);
    red[(buf * NW + warp) * 32 + lane] = acc[0];
    __syncthreads();
    if (warp == 0) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < NW; ++w) v += red[(buf * NW + w) * 32 + lane];
      op.template epilogue<T>(rb, nv, lane, v, ost);
    }
  }
  op.finish(ost);
}

// ---------------------------------------------------------------------------------
// Op 1: RMSNorm -> q/k/v projection -> RoPE -> paged-KV append
//   virtual row vr = 2*task + which; task = slot*64 + d; slot in [0, n_heads + 2*n_kv):
//   q heads, then k heads, then v heads; which selects dim d (0) or d+64 (1), the
//   rotate-half partner, so a RoPE pair lands in adjacent lanes of warp 0.
// ---------------------------------------------------------------------------------
struct OpQKV {
  static constexpr int ALIGN = 2;
  int K, nrows;
  const bf16 *x, *ln_w;              // [T][H], [H]
  const bf16 *wq, *wk, *wv;          // [n_heads*128][H], [n_kv*128][H] x2
  const bf16 *bq, *bk, *bv;          // optional biases
  bf16* q_out;                       // [T][n_heads*128]
  bf16* kv_pool;                     // this layer's pages [page][2][n_kv][PAGE][HD]
  const int32_t* block_table;
  const StepState* st;
  const float* inv_freq;             // [64]
  int n_heads, n_kv;
  float eps;

  __device__ __forceinline__ const bf16* row(int vr) const {
    const int task = vr >> 1, which = vr & 1;
    const int slot = task >> 6, d = (task & 63) + (which << 6);
    const bf16* base;
    int hrow;
    if (slot < n_heads) { base = wq; hrow = slot; }
    else if (slot < n_heads + n_kv) { base = wk; hrow = slot - n_heads; }
    else { base = wv; hrow = slot - n_heads - n_kv; }
    return base + ((size_t)hrow * HD + d) * K;
  }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float* scratch, float*) const {
    stage_rmsnorm<T>(xs, scratch, x, ln_w, K, eps);
  }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float* ost) const {
    const int r_local = lane / T, t = lane % T;
    const int vr = rb + r_local;
    const bool valid = r_local < nv;
    const int task = vr >> 1, which = vr & 1;
    const int slot = task >> 6, d = task & 63;
    int kind = 0, hrow = slot;  // 0 q, 1 k, 2 v
    if (slot >= n_heads + n_kv) { kind = 2; hrow = slot - n_heads - n_kv; }
    else if (slot >= n_heads) { kind = 1; hrow = slot - n_heads; }
    const int dim = d + (which << 6);
    const bf16* bias = kind == 0 ? bq : (kind == 1 ? bk : bv);
    if (valid && bias != nullptr) v += __bfloat162float(bias[hrow * HD + dim]);
    const float y = bf16r(v);                                   // Linear output in T
    const float yp = __shfl_xor_sync(0xffffffffu, y, T);        // rotate-half partner
    if (!valid) return;
    const int pos = st->pos + t;
    float o = y;
    if (kind != 2) {
      // mx.fast.rope (traditional=False): fp32 rotation of the bf16 values, one rounding
      const float theta = __fmul_rn((float)pos, inv_freq[d]);
      float sn, cs;
      sincosf(theta, &sn, &cs);
      o = which == 0 ? __fsub_rn(__fmul_rn(y, cs), __fmul_rn(yp, sn))
                     : __fadd_rn(__fmul_rn(yp, sn), __fmul_rn(y, cs));
      o = bf16r(o);
    }
    if (kind == 0) {
      q_out[(size_t)t * n_heads * HD + hrow * HD + dim] = __float2bfloat16_rn(o);
    } else {
      const int page = block_table[pos / PAGE];
      const size_t off = (((size_t)page * 2 + (kind - 1)) * n_kv + hrow) * (PAGE * HD) +
                         (size_t)(pos % PAGE) * HD + dim;
      kv_pool[off] = __float2bfloat16_rn(o);
    }
  }
  __device__ __forceinline__ void finish(float*) const {}
};

// ---------------------------------------------------------------------------------
// Op 2: attention output -> o_proj -> + residual          h = T(x + T(a W_o^T))
// ---------------------------------------------------------------------------------
struct OpOProj {
  static constexpr int ALIGN = 1;
  int K, nrows;
  const bf16 *a, *w, *resid;   // [T][K], [H][K], [T][H]
  bf16* out;                   // [T][H]
  __device__ __forceinline__ const bf16* row(int r) const { return w + (size_t)r * K; }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float*, float*) const { stage_copy<T>(xs, a, K); }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float* ost) const {
    const int r_local = lane / T, t = lane % T;
    if (r_local >= nv) return;
    const size_t idx = (size_t)t * nrows + rb + r_local;
    const float o = bf16r(v);
    out[idx] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(resid[idx]), o));
  }
  __device__ __forceinline__ void finish(float*) const {}
};

// ---------------------------------------------------------------------------------
// Op 3: RMSNorm -> gate/up -> SwiGLU with per-primitive bf16 temporaries
//   virtual row vr = 2*i + which (0 gate, 1 up)
// ---------------------------------------------------------------------------------
struct OpGateUp {
  static constexpr int ALIGN = 2;
  int K, nrows;                // nrows = 2*ffn
  const bf16 *x, *ln_w, *wg, *wu;
  bf16* act;                   // [T][ffn]
  float eps;
  __device__ __forceinline__ const bf16* row(int vr) const {
    return ((vr & 1) ? wu : wg) + (size_t)(vr >> 1) * K;
  }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float* scratch, float*) const {
    stage_rmsnorm<T>(xs, scratch, x, ln_w, K, eps);
  }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float* ost) const {
    const int r_local = lane / T, t = lane % T;
    const int vr = rb + r_local;
    const float y = bf16r(v);
    const float u = __shfl_xor_sync(0xffffffffu, y, T);
    if (r_local >= nv || (vr & 1)) return;
    const float s = bf16r(1.0f / (1.0f + expf(-y)));   // T(sigmoid(g))
    const float a = bf16r(__fmul_rn(y, s));            // T(g * s)
    const float m = __fmul_rn(a, u);                   // T(a * u) on store
    act[(size_t)t * (nrows >> 1) + (vr >> 1)] = __float2bfloat16_rn(m);
  }
  __device__ __forceinline__ void finish(float*) const {}
};

// ---------------------------------------------------------------------------------
// Op 4: down_proj -> + residual -> (cast to wire dtype == T)    out = T(h + T(m W_d^T))
// ---------------------------------------------------------------------------------
typedef OpOProj OpDown;  // identical dataflow: a=[T][ffn], w=[H][ffn], resid=h, out=x

// ---------------------------------------------------------------------------------
// Sparse MoE FFN (mixtral; BASELINE configs[4]).  Restated from mlx_lm.models.mixtral.MixtralSparseMoeBlock +
// switch_layers.SwitchGLU (oracle/llama_oracle.py moe_block has the rounding points); the reference hosts MoE
// families through the same operator API (src/dnet/core/models/gpt_oss.py, deepseek_v2.py).  One token at a time
// (T = 1): the experts a token uses are only known on the device, so the expert ops take a device table of expert
// weight pointers and a device index instead of weight pointers -- the launch sequence is fixed, i.e. capturable.
//   OpRouter      xn = RMSNorm(h);  logits[e] = T(xn . Wr[e])
//   k_moe_select  the k largest logits (ties: lowest index), scores = T(softmax over those k, fp32 math)
//   OpGateUpMoe   act_j = SwiGLU of expert sel[j]           (j = 0..k-1; same epilogue as the dense OpGateUp)
//   OpDownMoe     y_j = T(act_j Wd[sel[j]]^T); term_j = T(y_j * score_j); acc = T(acc + term_j) in selection order;
//                 the last j writes out = T(h + acc)
// ---------------------------------------------------------------------------------
struct OpRouter {
  static constexpr int ALIGN = 1;
  int K, nrows;                  // hidden, n_experts
  const bf16 *x, *ln_w, *w;      // [H], [H], [E][H]
  float* logits;                 // [E]
  float eps;
  __device__ __forceinline__ const bf16* row(int r) const { return w + (size_t)r * K; }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float* scratch, float*) const {
    stage_rmsnorm<T>(xs, scratch, x, ln_w, K, eps);
  }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float*) const {
    const int r_local = lane / T, t = lane % T;
    if (r_local < nv) logits[(size_t)t * nrows + rb + r_local] = bf16r(v);
  }
  __device__ __forceinline__ void finish(float*) const {}
};

constexpr int MOE_MAX_E = 64, MOE_MAX_K = 8;
__global__ void __launch_bounds__(32) k_moe_select(const float* __restrict__ logits, int E, int k, int32_t* __restrict__ sel,
                                                   float* __restrict__ score) {
  if (threadIdx.x != 0) return;
  float v[MOE_MAX_K];
  int id[MOE_MAX_K];
  unsigned long long taken = 0ull;
  for (int j = 0; j < k; ++j) {                    // selection in descending order, first index wins a tie
    float best = -INFINITY; int bi = 0;
    for (int e = 0; e < E; ++e)
      if (!((taken >> e) & 1ull) && logits[e] > best) { best = logits[e]; bi = e; }
    taken |= 1ull << bi;
    v[j] = best; id[j] = bi;
  }
  float sum = 0.f, ex[MOE_MAX_K];
  for (int j = 0; j < k; ++j) { ex[j] = expf(v[j] - v[0]); sum += ex[j]; }
  for (int j = 0; j < k; ++j) { sel[j] = id[j]; score[j] = bf16r(ex[j] / sum); }
}

struct OpGateUpMoe {
  static constexpr int ALIGN = 2;
  int K, nrows;                        // hidden, 2*ffn
  const bf16 *x, *ln_w;
  const bf16* const* table;            // device: [3][E] expert pointers (gate, up, down)
  const int32_t* sel; int j, E;
  bf16* act;                           // [ffn] of selection j
  float eps;
  __device__ __forceinline__ const bf16* row(int vr) const {
    // (clamped: with programmatic dependent launch the L2 pre-touch may run before k_moe_select has written sel)
    const bf16* base = table[((vr & 1) ? E : 0) + min(max(__ldg(sel + j), 0), E - 1)];
    return base + (size_t)(vr >> 1) * K;
  }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float* scratch, float*) const {
    stage_rmsnorm<T>(xs, scratch, x, ln_w, K, eps);
  }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float*) const {
    const int r_local = lane / T, t = lane % T;
    const int vr = rb + r_local;
    const float y = bf16r(v);
    const float u = __shfl_xor_sync(0xffffffffu, y, T);
    if (r_local >= nv || (vr & 1)) return;
    const float s = bf16r(1.0f / (1.0f + expf(-y)));
    const float a = bf16r(__fmul_rn(y, s));
    act[(size_t)t * (nrows >> 1) + (vr >> 1)] = __float2bfloat16_rn(__fmul_rn(a, u));
  }
  __device__ __forceinline__ void finish(float*) const {}
};

struct OpDownMoe {
  static constexpr int ALIGN = 1;
  int K, nrows;                        // ffn, hidden
  const bf16* a;                       // [ffn] activation of selection j
  const bf16* const* table;
  const int32_t* sel; const float* score; int j, k, E;
  float* ybuf;                         // [H] running T-rounded sum of the weighted expert outputs
  const bf16* resid; bf16* out;        // [H]
  __device__ __forceinline__ const bf16* row(int r) const {
    return table[2 * E + min(max(__ldg(sel + j), 0), E - 1)] + (size_t)r * K;
  }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float*, float*) const { stage_copy<T>(xs, a, K); }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float*) const {
    const int r_local = lane / T;
    if (r_local >= nv) return;
    const int idx = rb + r_local;
    const float term = bf16r(__fmul_rn(bf16r(v), __ldg(score + j)));
    const float acc = j == 0 ? term : bf16r(__fadd_rn(ybuf[idx], term));
    if (j + 1 < k) ybuf[idx] = acc;
    else out[idx] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(resid[idx]), acc));
  }
  __device__ __forceinline__ void finish(float*) const {}
};

// ---------------------------------------------------------------------------------
// Op 5: final RMSNorm (last position) -> lm_head -> bf16 logits -> greedy sample
//   Sampler.sample with temperature 0: argmax (first maximal index), logprob =
//   T(v - T(logsumexp(v)))     (reference core/decoding/sampler.py:33-52)
// ---------------------------------------------------------------------------------
struct HeadPartial { float m; float l; int idx; int pad; };

struct OpHead {
  static constexpr int ALIGN = 1;
  int K, nrows;                  // hidden, vocab
  const bf16 *x, *ln_w, *w;      // x = last position row [H]
  bf16* logits_bf16;             // [V] (always written)
  float* logits_f32;             // [V] or null: accumulator before rounding
  HeadPartial* partials;         // [gridDim.x]
  unsigned int* ticket;
  int32_t* token_out;            // device-accessible, may be null
  float* logprob_out;            // device-accessible, may be null
  StepState* st;                 // may be null
  float eps;
  // per-lane running (max, argmax, sum-exp) of warp 0 lives in ost: [0,32) m, [32,64) l, [64,96) idx
  __device__ __forceinline__ const bf16* row(int r) const { return w + (size_t)r * K; }
  template <int T>
  __device__ __forceinline__ void prologue(bf16* xs, float* scratch, float* ost) const {
    static_assert(T == 1, "head runs on the last position only");
    if (threadIdx.x < 32) {
      ost[threadIdx.x] = -INFINITY; ost[32 + threadIdx.x] = 0.f;
      reinterpret_cast<int*>(ost)[64 + threadIdx.x] = 0x7fffffff;
    }
    stage_rmsnorm<1>(xs, scratch, x, ln_w, K, eps);
  }
  template <int T>
  __device__ __forceinline__ void epilogue(int rb, int nv, int lane, float v, float* ost) const {
    if (lane >= nv) return;
    const int r = rb + lane;
    const float lg = bf16r(v);
    logits_bf16[r] = __float2bfloat16_rn(v);
    if (logits_f32 != nullptr) logits_f32[r] = v;
    float* sm = ost; float* sl = ost + 32; int* si = reinterpret_cast<int*>(ost) + 64;
    const float m = sm[lane];
    if (lg > m) {          // strict: rows grow within a lane, so the first maximum is kept
      sl[lane] = sl[lane] * exp2f((m - lg) * LOG2E) + 1.0f;
      sm[lane] = lg; si[lane] = r;
    } else {
      sl[lane] += exp2f((lg - m) * LOG2E);
    }
  }
  __device__ __forceinline__ void finish(float* ost) const {
    __shared__ int s_last;
    float* sm = ost; float* sl = ost + 32; int* si = reinterpret_cast<int*>(ost) + 64;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
      float m = sm[lane], l = sl[lane]; int idx = si[lane];
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, s);
        const float ol = __shfl_xor_sync(0xffffffffu, l, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        const float nm = fmaxf(m, om);
        const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
        const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
        l = a + b;
        if (om > m || (om == m && oi < idx)) idx = oi;
        m = nm;
      }
      if (lane == 0) {
        HeadPartial p; p.m = m; p.l = l; p.idx = idx; p.pad = 0;
        partials[blockIdx.x] = p;
        __threadfence();
        const unsigned int old = atomicAdd(ticket, 1u);
        s_last = (old == gridDim.x - 1) ? 1 : 0;
      }
    }
    __syncthreads();
    if (!s_last || threadIdx.x >= 32) return;
    __threadfence();
    // last CTA: fixed-order reduction over the per-CTA partials (deterministic)
    float m = -INFINITY, l = 0.f; int idx = 0x7fffffff;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      const float om = __ldcg(&partials[i].m), ol = __ldcg(&partials[i].l);
      const int oi = __ldcg(&partials[i].idx);
      const float nm = fmaxf(m, om);
      const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
      const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
      l = a + b;
      if (om > m || (om == m && oi < idx)) idx = oi;
      m = nm;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, m, s);
      const float ol = __shfl_xor_sync(0xffffffffu, l, s);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
      const float nm = fmaxf(m, om);
      const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
      const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
      l = a + b;
      if (om > m || (om == m && oi < idx)) idx = oi;
      m = nm;
    }
    if (lane == 0) {
      const float lse = bf16r(m + logf(l));            // T(logsumexp(v))
      const float lp = bf16r(__fsub_rn(m, lse));        // T(v[tok] - lse); v[tok] == m
      if (token_out != nullptr) *token_out = idx;
      if (logprob_out != nullptr) *logprob_out = lp;
      if (st != nullptr) st->token = idx;
      *ticket = 0u;                                      // re-arm for the next replay
      __threadfence_system();
    }
  }
};

// ---------------------------------------------------------------------------------
// paged-KV decode attention (flash-decoding, split over KV pages, last-CTA combine)
//   grid (n_kv * nsplit, T), block = G warps; warp w serves q head kvh*G + w.
//   mx.fast.scaled_dot_product_attention semantics: fp32 scores/softmax/PV on bf16 q,k,v,
//   one rounding to bf16; causal for a T-token chunk: query t sees positions <= pos0 + t.
// ---------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(G * 32) k_attn(const bf16* __restrict__ q, const bf16* __restrict__ kv_pool,
                                                const int32_t* __restrict__ block_table,
                                                const StepState* __restrict__ st, float* __restrict__ part,
                                                unsigned int* __restrict__ tickets, bf16* __restrict__ out,
                                                int n_heads, int n_kv, int nsplit) {
  __shared__ __align__(16) bf16 Ks[PAGE * HD];
  __shared__ __align__(16) bf16 Vs[PAGE * HD];
  __shared__ float ps[G][32];
  __shared__ int s_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kvh = blockIdx.x / nsplit, sp = blockIdx.x % nsplit, t = blockIdx.y;
  const int head = kvh * G + warp;
  pdl_launch_dependents();
  pdl_wait();
  const int kv_len = st->pos + t + 1;
  const int npages = (kv_len + PAGE - 1) / PAGE;
  const int pps = (npages + nsplit - 1) / nsplit;
  const int p0 = sp * pps, p1 = min(p0 + pps, npages);

  const float scale = 0.08838834764831845f;  // 128^-0.5
  float qv[4];
  {
    const uint2 u = *reinterpret_cast<const uint2*>(q + (size_t)t * n_heads * HD + head * HD + lane * 4);
    qv[0] = __fmul_rn(bf_lo(u.x), scale); qv[1] = __fmul_rn(bf_hi(u.x), scale);
    qv[2] = __fmul_rn(bf_lo(u.y), scale); qv[3] = __fmul_rn(bf_hi(u.y), scale);
  }
  float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};

  for (int p = p0; p < p1; ++p) {
    const int phys = block_table[p];
    const uint4* ksrc = reinterpret_cast<const uint4*>(kv_pool + (((size_t)phys * 2 + 0) * n_kv + kvh) * (PAGE * HD));
    const uint4* vsrc = reinterpret_cast<const uint4*>(kv_pool + (((size_t)phys * 2 + 1) * n_kv + kvh) * (PAGE * HD));
    const int ntok = min(PAGE, kv_len - p * PAGE);
    const int n16 = ntok * HD / 8;
    for (int i = threadIdx.x; i < n16; i += G * 32) {
      reinterpret_cast<uint4*>(Ks)[i] = ksrc[i];
      reinterpret_cast<uint4*>(Vs)[i] = vsrc[i];
    }
    __syncthreads();
#pragma unroll 1
    for (int h0 = 0; h0 < ntok; h0 += 32) {
      const int nt = min(32, ntok - h0);
      float sc[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        sc[j] = 0.f;
        if (j < nt) {
          const uint2 u = *reinterpret_cast<const uint2*>(Ks + (h0 + j) * HD + lane * 4);
          sc[j] = fmaf(qv[0], bf_lo(u.x), fmaf(qv[1], bf_hi(u.x), fmaf(qv[2], bf_lo(u.y), qv[3] * bf_hi(u.y))));
        }
      }
      transpose_reduce32(sc, lane);
      const bool valid = lane < nt;
      const float s = valid ? sc[0] : -INFINITY;
      const float m_new = fmaxf(m, warp_max(s));
      const float pj = valid ? exp2f((s - m_new) * LOG2E) : 0.f;
      const float corr = exp2f((m - m_new) * LOG2E);   // m == -inf -> 0
      l = l * corr + warp_sum(pj);
#pragma unroll
      for (int d = 0; d < 4; ++d) o[d] *= corr;
      m = m_new;
      ps[warp][lane] = pj;
      __syncwarp();
      for (int j = 0; j < nt; ++j) {
        const float pw = ps[warp][j];
        const uint2 u = *reinterpret_cast<const uint2*>(Vs + (h0 + j) * HD + lane * 4);
        o[0] = fmaf(pw, bf_lo(u.x), o[0]); o[1] = fmaf(pw, bf_hi(u.x), o[1]);
        o[2] = fmaf(pw, bf_lo(u.y), o[2]); o[3] = fmaf(pw, bf_hi(u.y), o[3]);
      }
      __syncwarp();
    }
    __syncthreads();
  }
  // partial for (t, head, split)
  float* pp = part + (((size_t)t * n_heads + head) * nsplit + sp) * PART_STRIDE;
  *reinterpret_cast<float4*>(pp + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
  if (lane == 0) { pp[128] = m; pp[129] = l; }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int old = atomicAdd(&tickets[t * n_kv + kvh], 1u);
    s_last = (old == (unsigned)nsplit - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last CTA of this (token, kv head): merge the splits in fixed order
  const float* hp = part + (((size_t)t * n_heads + head) * nsplit) * PART_STRIDE;
  float M = -INFINITY;
  for (int s2 = 0; s2 < nsplit; ++s2) M = fmaxf(M, __ldcg(hp + (size_t)s2 * PART_STRIDE + 128));
  float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s2 = 0; s2 < nsplit; ++s2) {
    const float ms = __ldcg(hp + (size_t)s2 * PART_STRIDE + 128);
    if (ms == -INFINITY) continue;
    const float wgt = exp2f((ms - M) * LOG2E);
    L = fmaf(__ldcg(hp + (size_t)s2 * PART_STRIDE + 129), wgt, L);
    const float4 ov = __ldcg(reinterpret_cast<const float4*>(hp + (size_t)s2 * PART_STRIDE + lane * 4));
    acc[0] = fmaf(ov.x, wgt, acc[0]); acc[1] = fmaf(ov.y, wgt, acc[1]);
    acc[2] = fmaf(ov.z, wgt, acc[2]); acc[3] = fmaf(ov.w, wgt, acc[3]);
  }
  const float invL = 1.0f / L;
  __align__(8) bf16 ob[4];
#pragma unroll
  for (int d = 0; d < 4; ++d) ob[d] = __float2bfloat16_rn(acc[d] * invL);
  *reinterpret_cast<uint2*>(out + (size_t)t * n_heads * HD + head * HD + lane * 4) = *reinterpret_cast<const uint2*>(ob);
  if (threadIdx.x == 0) tickets[t * n_kv + kvh] = 0u;
}

// ---------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------
// embed: row gather (reference core/models/llama.py:56-57); ids device int32[T]
__global__ void __launch_bounds__(256) k_embed(const int32_t* __restrict__ ids, const bf16* __restrict__ table,
                                               bf16* __restrict__ out, int H, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  int id = ids[blockIdx.x];
  if (id < 0) id = 0;
  if (id >= vocab) id = vocab - 1;
  const uint4* src = reinterpret_cast<const uint4*>(table + (size_t)id * H);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)blockIdx.x * H);
  for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}

__global__ void k_advance(StepState* st, int T) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) st->pos += T;
}
__global__ void k_set_state(StepState* st, int pos, int token, int set_pos, int set_token) {
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) {
    if (set_pos) st->pos = pos;
    if (set_token) st->token = token;
  }
}

// ring hop flags (system scope: the flag lives on the receiving GPU, written over NVLink)
__global__ void k_flag_set(uint32_t* flag, uint32_t seq) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
  }
}
// decode hop: payload (<= 64 KiB) and flag in one launch -- peer stores over NVLink, then a
// system-scope release of the sequence number.  No copy engine is involved, so a hop can never
// queue behind an unrelated cudaMemcpy (copy-engine queues are shared across streams).
__global__ void __launch_bounds__(512) k_hop_send(uint4* __restrict__ dst, const uint4* __restrict__ src, int n16,
                                                  uint32_t* flag, uint32_t seq) {
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(seq) : "memory");
}
__global__ void k_flag_wait(const uint32_t* flag, uint32_t seq, unsigned long long timeout_ns, uint32_t* err) {
  if (threadIdx.x != 0) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if ((int32_t)(v - seq) >= 0) break;
    __nanosleep(100);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > timeout_ns) {   // never hang the GPU: flag the error and fall through
      if (err != nullptr) atomicExch(err, 1u);
      break;
    }
  }
}

// Direct measurement of the ring hop: a token (payload of `n16` uint4 + sequence flag) travels `iters` times
// around the ring of GPUs, every rank running this kernel on one lane nobody else uses.  Each rank waits for
// its own flag, copies the payload from its slot into the successor's slot with peer stores, fences at system
// scope and releases the successor's flag -- exactly what the fused hop at the end of k_shard_step does.  The
// origin rank times the whole trip on its own globaltimer (no cross-GPU clock offset involved):
// one hop = elapsed / (iters * ring size).  Bounded spins: a dead peer ends the kernel with *out_ns = 0.
__global__ void __launch_bounds__(512) k_hop_ring_probe(const uint4* __restrict__ own_slot, const uint32_t* own_flag,
                                                        uint4* __restrict__ next_slot, uint32_t* next_flag, int n16,
                                                        uint32_t base, int iters, int is_origin,
                                                        unsigned long long timeout_ns, unsigned long long* out_ns) {
  __shared__ int dead;
  if (threadIdx.x == 0) dead = 0;
  __syncthreads();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i <= iters; ++i) {
    // origin: iteration i first waits for the return of token i (none for i == 0); the others wait for token i+1
    const uint32_t want = base + (uint32_t)(is_origin ? i : i + 1);
    if (!(is_origin && i == 0) && !(!is_origin && i == iters)) {
      if (threadIdx.x == 0) {
        unsigned long long ts;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ts));
        for (;;) {
          uint32_t v;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(own_flag) : "memory");
          if ((int32_t)(v - want) >= 0) break;
          unsigned long long tn;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tn));
          if (tn - ts > timeout_ns) { dead = 1; break; }
        }
      }
      __syncthreads();
      if (dead) { if (threadIdx.x == 0 && out_ns) *out_ns = 0ull; return; }
    }
    if (i == iters) break;                       // origin: last return received; others: all tokens forwarded
    for (int j = threadIdx.x; j < n16; j += blockDim.x) next_slot[j] = __ldcg(own_slot + j);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(next_flag), "r"(base + (uint32_t)i + 1u) : "memory");
  }
  if (threadIdx.x == 0 && out_ns != nullptr) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    *out_ns = t1 - t0;
  }
}

}  // namespace dn
