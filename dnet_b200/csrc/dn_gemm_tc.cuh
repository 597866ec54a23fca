// dn_gemm_tc.cuh -- prefill GEMMs on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged
// by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B), warp-specialised: warp 0 = TMA producer, warp 1 =
// single-thread tcgen05.mma issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld).
//
//   Y[T, N] = X[T, K] . W[N, K]^T        bf16 operands, fp32 accumulate in TMEM
//
// "swap-AB" orientation: the WEIGHT tile is the MMA's A operand (M = 128 output features per CTA
// tile, one TMEM lane each) and the TOKEN tile is B (N = BN tokens = TMEM columns), because prefill
// chunks have few tokens and many output features; both operands are K-major in shared memory,
// exactly what TMA produces from the row-major [rows, K] tensors.  Each CTA walks n-tiles
// persistently; per k-step (64 columns = one 128-byte swizzle span) four UMMA_K=16 instructions.
//
// Epilogues (fused, same rounding points as the decode kernels / oracle):
//   EPI_STORE  : y = T(acc + bias)                        -> Y[t][col0 + n]
//   EPI_RESID  : y = T(resid[t][n] + T(acc))              -> Y[t][n]
//   EPI_SWIGLU : two weight tiles (gate, up) share the token tile; g = T(acc_g), u = T(acc_u),
//                s = T(sigmoid g), a = T(g s), m = T(a u)  -> Y[t][n]
#pragma once
#include <cuda.h>
#include "dn_megakernel.cuh"

namespace dn {

enum { EPI_STORE = 0, EPI_RESID = 1, EPI_SWIGLU = 2 };

constexpr int TC_BM = 128;        // weight rows per tile (MMA M)
constexpr int TC_BK = 64;         // k-step: 64 bf16 = 128 bytes = one swizzle span
constexpr int TC_THREADS = 256;

struct TcParams {
  int T, N, K;                    // tokens, output features, reduction
  int ldy, col0;                  // output leading dimension (elements) and column offset
  bf16* Y;
  const bf16* resid;              // [T][ldr]
  int ldr;
  const bf16* bias;               // [N] or null
  unsigned int* err;
};

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T   (kind::f16: bf16 x bf16 -> fp32)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//  [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major) |
//  [32,46) stride byte offset >> 4 = 1024 B between 8-row groups | [46,48) version = 1 (sm_100) |
//  [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t tc_smem_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) at [4,6), a/b format BF16 (1)
// at [7,10)/[10,13), a/b K-major (0) at 15/16, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t tc_instr_desc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 16 consecutive fp32 columns of this warp's TMEM quarter
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

template <int BN, int EPI, int STAGES>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapW2,
          const __grid_constant__ CUtensorMap mapX, const TcParams p) {
  constexpr int NA = (EPI == EPI_SWIGLU) ? 2 : 1;
  constexpr int A_BYTES = TC_BM * TC_BK * 2;      // 16 KB
  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = NA * A_BYTES + B_BYTES;
  // two accumulator sets when they fit the 512 TMEM columns: the epilogue of tile i drains one while
  // the MMAs of tile i+1 fill the other
  constexpr int NACC = (2 * NA * BN <= 512) ? 2 : 1;
  constexpr int ACC_COLS = NA * BN;
  constexpr uint32_t TMEM_COLS = (NACC * ACC_COLS <= 32) ? 32 : (NACC * ACC_COLS <= 64) ? 64 : (NACC * ACC_COLS <= 128) ? 128 : (NACC * ACC_COLS <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;          // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.N + TC_BM - 1) / TC_BM;
  const int t_tiles = (p.T + BN - 1) / BN;
  const int total_tiles = n_tiles * t_tiles;
  const int k_steps = p.K / TC_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }   // 4: one arrival per epilogue warp
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {                // TMEM allocation (whole warp), address lands in shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tt = tile % t_tiles, nt = tile / t_tiles;     // token tiles of one weight tile run back to back (weights hit L2)
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1u, p.err);
          mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
          unsigned char* sa = smem + (size_t)stage * STAGE_BYTES;
          tma_load_2d(sa, &mapW, ks * TC_BK, nt * TC_BM, &full[stage]);
          if (NA == 2) tma_load_2d(sa + A_BYTES, &mapW2, ks * TC_BK, nt * TC_BM, &full[stage]);
          tma_load_2d(sa + NA * A_BYTES, &mapX, ks * TC_BK, tt * BN, &full[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      const uint32_t idesc = tc_instr_desc(TC_BM, BN);
      int stage = 0; uint32_t phase = 0;
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const uint32_t ab = (NACC == 2) ? (it & 1u) : 0u;              // accumulator set of this tile
        const uint32_t acc_phase = ((NACC == 2) ? (it >> 1) : it) & 1u;
        const uint32_t tacc = tmem_base + ab * ACC_COLS;
        mbar_wait(&tmem_empty[ab], acc_phase ^ 1u, p.err); // epilogue drained this accumulator set
        tc_fence_after();
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(&full[stage], phase, p.err);
          tc_fence_after();
          unsigned char* sa = smem + (size_t)stage * STAGE_BYTES;
          const uint64_t bdesc = tc_smem_desc(sa + NA * A_BYTES);
#pragma unroll
          for (int a = 0; a < NA; ++a) {
            const uint64_t adesc = tc_smem_desc(sa + a * A_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)      // +32 bytes (>>4 = 2) per UMMA_K inside the swizzle span
              tc_mma(tacc + a * BN, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ks | k) ? 1u : 0u);
          }
          tc_commit(&empty[stage]);                   // smem stage reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(&tmem_full[ab]);                    // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int ew = warp - 4;                          // TMEM lanes [32 ew, 32 ew + 32)
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int tt = tile % t_tiles, nt = tile / t_tiles;     // token tiles of one weight tile run back to back (weights hit L2)
      const int n = nt * TC_BM + ew * 32 + lane;      // output feature owned by this thread
      const uint32_t ab = (NACC == 2) ? (it & 1u) : 0u;
      const uint32_t acc_phase = ((NACC == 2) ? (it >> 1) : it) & 1u;
      mbar_wait(&tmem_full[ab], acc_phase, p.err);
      tc_fence_after();
      const uint32_t trow = tmem_base + ab * ACC_COLS + ((uint32_t)(ew * 32) << 16);
      float bias = 0.f;
      if (EPI == EPI_STORE && p.bias != nullptr && n < p.N) bias = __bfloat162float(p.bias[n]);
#pragma unroll 1
      for (int c = 0; c < BN; c += 16) {
        float v[16], v2[16];
        tc_ld16(trow + c, v);
        if (EPI == EPI_SWIGLU) tc_ld16(trow + BN + c, v2);
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int t = tt * BN + c + j;
            if (t < p.T) {
              float y;
              if (EPI == EPI_STORE) {
                y = v[j] + bias;
              } else if (EPI == EPI_RESID) {
                y = __fadd_rn(__bfloat162float(p.resid[(size_t)t * p.ldr + n]), bf16r(v[j]));
              } else {
                const float g = bf16r(v[j]), u = bf16r(v2[j]);
                const float s = bf16r(1.0f / (1.0f + expf(-g)));
                y = __fmul_rn(bf16r(__fmul_rn(g, s)), u);
              }
              p.Y[(size_t)t * p.ldy + p.col0 + n] = __float2bfloat16_rn(y);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------
// prefill helpers around the GEMMs
// ---------------------------------------------------------------------------------
// RMSNorm of T rows (same arithmetic as stage_rmsnorm): one CTA per token
__global__ void __launch_bounds__(256) k_rmsnorm_rows(const bf16* __restrict__ x, const bf16* __restrict__ w, bf16* __restrict__ out,
                                                      int K, float eps) {
  __shared__ float scratch[8];
  const bf16* xr = x + (size_t)blockIdx.x * K;
  bf16* o = out + (size_t)blockIdx.x * K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < K; i += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) scratch[warp] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += scratch[i];
  const float inv = 1.0f / sqrtf(tot / (float)K + eps);
  for (int i = threadIdx.x * 8; i < K; i += 256 * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(xr + i);
    const uint4 g = *reinterpret_cast<const uint4*>(w + i);
    const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
    const float gw[8] = {bf_lo(g.x), bf_hi(g.x), bf_lo(g.y), bf_hi(g.y), bf_lo(g.z), bf_hi(g.z), bf_lo(g.w), bf_hi(g.w)};
    __align__(16) bf16 ob[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) ob[j] = __float2bfloat16_rn(__fmul_rn(bf16r(__fmul_rn(f[j], inv)), gw[j]));
    *reinterpret_cast<uint4*>(o + i) = *reinterpret_cast<const uint4*>(ob);
  }
}

// RoPE on q (in place in the qkv buffer -> q_out) and k, and KV append, for T tokens.
// qkv: [T][(n_heads + 2 n_kv) * 128] bf16 (Linear outputs already rounded to bf16)
__global__ void __launch_bounds__(128) k_rope_append(const bf16* __restrict__ qkv, bf16* __restrict__ q_out, bf16* __restrict__ kv_pool,
                                                     const int32_t* __restrict__ block_table, const StepState* __restrict__ st,
                                                     const float* __restrict__ inv_freq, int n_heads, int n_kv) {
  const int t = blockIdx.y, slot = blockIdx.x;            // slot: q heads, then k heads, then v heads
  const int pos = st->pos + t;
  const int ld = (n_heads + 2 * n_kv) * HD;
  const bf16* src = qkv + (size_t)t * ld + (size_t)slot * HD;
  const int d = threadIdx.x;                              // 0..127
  const float y = __bfloat162float(src[d]);
  int kind = 0, hrow = slot;
  if (slot >= n_heads + n_kv) { kind = 2; hrow = slot - n_heads - n_kv; }
  else if (slot >= n_heads) { kind = 1; hrow = slot - n_heads; }
  float o = y;
  if (kind != 2) {
    const int dd = d & 63;
    const float yp = __bfloat162float(src[d ^ 64]);
    const float theta = __fmul_rn((float)pos, inv_freq[dd]);
    float sn, cs;
    sincosf(theta, &sn, &cs);
    o = (d < 64) ? __fsub_rn(__fmul_rn(y, cs), __fmul_rn(yp, sn)) : __fadd_rn(__fmul_rn(yp, sn), __fmul_rn(y, cs));
    o = bf16r(o);
  }
  if (kind == 0) {
    q_out[(size_t)t * n_heads * HD + hrow * HD + d] = __float2bfloat16_rn(o);
  } else {
    const int page = block_table[pos / PAGE];
    const size_t off = (((size_t)page * 2 + (kind - 1)) * n_kv + hrow) * (PAGE * HD) + (size_t)(pos % PAGE) * HD + d;
    kv_pool[off] = __float2bfloat16_rn(o);
  }
}

// causal prefill attention over the paged KV: CTA = (kv head, group of QT query tokens); warp w
// serves q head kvh*G + w for each of the QT tokens; K/V tiles are staged once per CTA.
template <int G, int QT>
__global__ void __launch_bounds__(G * 32) k_attn_prefill(const bf16* __restrict__ q, const bf16* __restrict__ kv_pool,
                                                        const int32_t* __restrict__ block_table, const StepState* __restrict__ st,
                                                        bf16* __restrict__ out, int n_heads, int n_kv, int T) {
  __shared__ __align__(16) bf16 Ks[PAGE * HD];
  __shared__ __align__(16) bf16 Vs[PAGE * HD];
  __shared__ float ps[G][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int kvh = blockIdx.x, t0 = blockIdx.y * QT;
  const int head = kvh * G + warp;
  const int pos0 = st->pos;
  const int nq = min(QT, T - t0);
  const int kv_max = pos0 + t0 + nq;                     // keys visible to the last query of this CTA
  const float scale = 0.08838834764831845f;
  float qv[QT][4], m[QT], l[QT], o[QT][4];
#pragma unroll
  for (int i = 0; i < QT; ++i) {
    m[i] = -INFINITY; l[i] = 0.f;
    o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    qv[i][0] = qv[i][1] = qv[i][2] = qv[i][3] = 0.f;
    if (i < nq) {
      const uint2 u = *reinterpret_cast<const uint2*>(q + (size_t)(t0 + i) * n_heads * HD + head * HD + lane * 4);
      qv[i][0] = __fmul_rn(bf_lo(u.x), scale); qv[i][1] = __fmul_rn(bf_hi(u.x), scale);
      qv[i][2] = __fmul_rn(bf_lo(u.y), scale); qv[i][3] = __fmul_rn(bf_hi(u.y), scale);
    }
  }
  const int npages = (kv_max + PAGE - 1) / PAGE;
  for (int pg = 0; pg < npages; ++pg) {
    const int phys = block_table[pg];
    const uint4* ksrc = reinterpret_cast<const uint4*>(kv_pool + (((size_t)phys * 2 + 0) * n_kv + kvh) * (PAGE * HD));
    const uint4* vsrc = reinterpret_cast<const uint4*>(kv_pool + (((size_t)phys * 2 + 1) * n_kv + kvh) * (PAGE * HD));
    const int ntok = min(PAGE, kv_max - pg * PAGE);
    for (int i = threadIdx.x; i < ntok * HD / 8; i += G * 32) {
      reinterpret_cast<uint4*>(Ks)[i] = ksrc[i];
      reinterpret_cast<uint4*>(Vs)[i] = vsrc[i];
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
      if (qi >= nq) continue;
      const int kv_len = pos0 + t0 + qi + 1;              // causal: query t sees positions <= pos0 + t
      const int vis = min(ntok, kv_len - pg * PAGE);
      if (vis <= 0) continue;
#pragma unroll 1
      for (int h0 = 0; h0 < vis; h0 += 32) {
        const int nt = min(32, vis - h0);
        float sc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          sc[j] = 0.f;
          if (j < nt) {
            const uint2 u = *reinterpret_cast<const uint2*>(Ks + (h0 + j) * HD + lane * 4);
            sc[j] = fmaf(qv[qi][0], bf_lo(u.x), fmaf(qv[qi][1], bf_hi(u.x), fmaf(qv[qi][2], bf_lo(u.y), qv[qi][3] * bf_hi(u.y))));
          }
        }
        transpose_reduce32(sc, lane);
        const bool valid = lane < nt;
        const float s = valid ? sc[0] : -INFINITY;
        const float m_new = fmaxf(m[qi], warp_max(s));
        const float pj = valid ? exp2f((s - m_new) * LOG2E) : 0.f;
        const float corr = exp2f((m[qi] - m_new) * LOG2E);
        l[qi] = l[qi] * corr + warp_sum(pj);
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) o[qi][dd] *= corr;
        m[qi] = m_new;
        ps[warp][lane] = pj;
        __syncwarp();
        for (int j = 0; j < nt; ++j) {
          const float pw = ps[warp][j];
          const uint2 u = *reinterpret_cast<const uint2*>(Vs + (h0 + j) * HD + lane * 4);
          o[qi][0] = fmaf(pw, bf_lo(u.x), o[qi][0]); o[qi][1] = fmaf(pw, bf_hi(u.x), o[qi][1]);
          o[qi][2] = fmaf(pw, bf_lo(u.y), o[qi][2]); o[qi][3] = fmaf(pw, bf_hi(u.y), o[qi][3]);
        }
        __syncwarp();
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int qi = 0; qi < QT; ++qi) {
    if (qi >= nq) continue;
    const float invL = 1.0f / l[qi];
    __align__(8) bf16 ob[4];
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) ob[dd] = __float2bfloat16_rn(o[qi][dd] * invL);
    *reinterpret_cast<uint2*>(out + (size_t)(t0 + qi) * n_heads * HD + head * HD + lane * 4) = *reinterpret_cast<const uint2*>(ob);
  }
}


// =================================================================================================
// causal prefill attention on tcgen05 (flash-attention forward over the paged KV)
//
// CTA = (kv head, block of TB = 128/G query tokens): the 128 MMA rows are the G query heads of the
// group x TB tokens (row = g*TB + t), so one K/V tile serves the whole GQA group.
//   S  = Q . K^T      A = Q [128 x 128 d] (K-major, written once with the 128B swizzle),
//                     B = K tile [128 tokens x 128 d] (K-major, TMA SWIZZLE_128B, 2 pages)
//   O += P . V        A = P [128 x 128 tokens] (K-major, written by the softmax threads),
//                     B = V tile [128 tokens x 128 d] as MN-major operand (rows = k), same TMA boxes
// Accumulators live in TMEM (S: columns 0-127, P.V: columns 128-255); the 128 softmax threads own
// one row each (TMEM lane = row): fp32 scores, causal mask, online max / sum exactly as the
// CUDA-core kernel, and the running output o[128] stays in registers (o = o*corr + P.V per tile),
// so nothing is read-modify-written in TMEM.  P is fed to the tensor core as P_hi + P_lo (two bf16
// terms, 16 mantissa bits): the oracle keeps P in fp32, and a single bf16 P would add a rounding
// point the reference does not have.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4-7 = softmax.
// =================================================================================================
struct AtParams {
  const bf16* q;                 // [T][n_heads][128] (RoPE applied, unscaled)
  bf16* out;                     // [T][n_heads][128]
  const int32_t* block_table;
  const StepState* st;
  int n_heads, n_kv, T;
  unsigned int* err;
};

// MN-major, SWIZZLE_128B operand: 64 contiguous elements along N per 128-byte row, 8 k-rows per
// 1024-byte atom; LBO = bytes between N atoms (dims 0-63 -> 64-127), SBO = bytes between k groups
__device__ __forceinline__ uint64_t tc_smem_desc_mn(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// byte offset of the 16-byte chunk (row r, chunk c of 16) inside a [2 halves][128 rows][128 B] swizzled tile
__device__ __forceinline__ uint32_t at_swz(int r, int c) {
  return (uint32_t)((c >> 3) * 16384 + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}

constexpr int AT_TILE = 128 * 128 * 2;     // one [128 x 128] bf16 tile = 32 KiB (two 64-column halves)
constexpr int AT_SMEM = 5 * AT_TILE + 1024 + 256;

template <int G>
__global__ void __launch_bounds__(256, 1) k_attn_prefill_tc(const __grid_constant__ CUtensorMap mapKV, const AtParams p) {
  constexpr int TB = 128 / G;
  extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* sQ = smem;
  unsigned char* sK = smem + AT_TILE;
  unsigned char* sV = smem + 2 * AT_TILE;
  unsigned char* sPh = smem + 3 * AT_TILE;
  unsigned char* sPl = smem + 4 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * AT_TILE);
  uint64_t *k_full = bars, *k_empty = bars + 1, *v_full = bars + 2, *v_empty = bars + 3, *s_full = bars + 4, *s_empty = bars + 5,
           *p_full = bars + 6, *pv_full = bars + 7, *pv_empty = bars + 8;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvh = blockIdx.x, t0 = blockIdx.y * TB;
  const int pos0 = p.st->pos;
  const int nq = min(TB, p.T - t0);                         // query tokens of this CTA
  const int kv_max = pos0 + t0 + nq;                        // keys visible to its last query
  const int n_tiles = (kv_max + 127) >> 7;

  if (threadIdx.x == 0) {
    mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(s_empty, 4); mbar_init(p_full, 4); mbar_init(pv_full, 1); mbar_init(pv_empty, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_base_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Q tile: rows (g, t) <- q[t0 + t][kvh*G + g][:], zero rows beyond the chunk; written with the TMA 128B swizzle
  for (int i = threadIdx.x; i < 128 * 16; i += 256) {
    const int r = i >> 4, c = i & 15;
    const int g = r / TB, t = r % TB;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (g < G && t < nq) v = *reinterpret_cast<const uint4*>(p.q + ((size_t)(t0 + t) * p.n_heads + kvh * G + g) * HD + c * 8);
    *reinterpret_cast<uint4*>(sQ + at_swz(r, c)) = v;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const uint32_t tS = tmem_base, tPV = tmem_base + 128;

  if (warp == 0) {
    if (lane == 0) {
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        const int np = (kv_max + PAGE - 1) / PAGE;
        const int pg0 = p.block_table[2 * j];
        const int pg1 = (2 * j + 1 < np) ? p.block_table[2 * j + 1] : pg0;     // unused half: any valid page (masked)
        const int rk0 = ((pg0 * 2 + 0) * p.n_kv + kvh) * PAGE, rk1 = ((pg1 * 2 + 0) * p.n_kv + kvh) * PAGE;
        const int rv0 = ((pg0 * 2 + 1) * p.n_kv + kvh) * PAGE, rv1 = ((pg1 * 2 + 1) * p.n_kv + kvh) * PAGE;
        mbar_wait(k_empty, ph ^ 1u, p.err);
        mbar_arrive_expect_tx(k_full, AT_TILE);
        tma_load_2d(sK, &mapKV, 0, rk0, k_full);
        tma_load_2d(sK + 16384, &mapKV, 64, rk0, k_full);
        tma_load_2d(sK + 8192, &mapKV, 0, rk1, k_full);
        tma_load_2d(sK + 16384 + 8192, &mapKV, 64, rk1, k_full);
        mbar_wait(v_empty, ph ^ 1u, p.err);
        mbar_arrive_expect_tx(v_full, AT_TILE);
        tma_load_2d(sV, &mapKV, 0, rv0, v_full);
        tma_load_2d(sV + 16384, &mapKV, 64, rv0, v_full);
        tma_load_2d(sV + 8192, &mapKV, 0, rv1, v_full);
        tma_load_2d(sV + 16384 + 8192, &mapKV, 64, rv1, v_full);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = tc_instr_desc(128, 128);
      const uint32_t idesc_pv = tc_instr_desc(128, 128) | (1u << 16);       // B operand MN-major
      for (int j = 0; j < n_tiles; ++j) {
        const uint32_t ph = (uint32_t)(j & 1);
        // ---- S = Q K^T
        mbar_wait(k_full, ph, p.err);
        mbar_wait(s_empty, ph ^ 1u, p.err);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint64_t adesc = tc_smem_desc(sQ + h * 16384), bdesc = tc_smem_desc(sK + h * 16384);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc_mma(tS, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, (h | k) ? 1u : 0u);
        }
        tc_commit(k_empty);
        tc_commit(s_full);
        // ---- PV = (P_hi + P_lo) V
        mbar_wait(p_full, ph, p.err);
        mbar_wait(v_full, ph, p.err);
        mbar_wait(pv_empty, ph ^ 1u, p.err);
        tc_fence_after();
        const uint64_t vdesc = tc_smem_desc_mn(sV, 16384, 1024);
#pragma unroll
        for (int part = 0; part < 2; ++part) {
          const unsigned char* sP = part ? sPl : sPh;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t adesc = tc_smem_desc(sP + (kk >> 2) * 16384) + (uint64_t)(2 * (kk & 3));
            tc_mma(tPV, adesc, vdesc + (uint64_t)(kk * (2048 >> 4)), idesc_pv, (part | kk) ? 1u : 0u);
          }
        }
        tc_commit(v_empty);
        tc_commit(pv_full);
      }
    }
  } else if (warp >= 4) {
    const int qd = warp - 4;                           // TMEM lane quarter
    const int r = qd * 32 + lane;                      // this thread's row
    const int g = r / TB, t = r % TB;
    const bool row_ok = g < G && t < nq;
    const int limit = pos0 + t0 + t;                   // last visible key position (causal)
    const float scale = 0.08838834764831845f;
    const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
    float m = -INFINITY, l = 0.f;
    float o[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) o[i] = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t ph = (uint32_t)(j & 1);
      const int tok0 = j << 7;
      mbar_wait(s_full, ph, p.err);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[16];
        tc_ld16(tS + lane_off + c * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool vis = row_ok && (tok0 + c * 16 + i) <= limit;
          mx = fmaxf(mx, vis ? v[i] * scale : -INFINITY);
        }
      }
      const float m_new = fmaxf(m, mx);
      const float corr = (m_new == -INFINITY) ? 1.0f : exp2f((m - m_new) * LOG2E);
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[16];
        tc_ld16(tS + lane_off + c * 16, v);
        __align__(16) bf16 hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const bool vis = row_ok && (tok0 + c * 16 + i) <= limit;
          const float pe = vis ? exp2f((v[i] * scale - m_new) * LOG2E) : 0.f;
          lsum += pe;
          hi[i] = __float2bfloat16_rn(pe);
          lo[i] = __float2bfloat16_rn(pe - __bfloat162float(hi[i]));
        }
        *reinterpret_cast<uint4*>(sPh + at_swz(r, 2 * c)) = *reinterpret_cast<const uint4*>(hi);
        *reinterpret_cast<uint4*>(sPh + at_swz(r, 2 * c + 1)) = *reinterpret_cast<const uint4*>(hi + 8);
        *reinterpret_cast<uint4*>(sPl + at_swz(r, 2 * c)) = *reinterpret_cast<const uint4*>(lo);
        *reinterpret_cast<uint4*>(sPl + at_swz(r, 2 * c + 1)) = *reinterpret_cast<const uint4*>(lo + 8);
      }
      l = l * corr + lsum;
      m = m_new;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(p_full); mbar_arrive(s_empty); }
      mbar_wait(pv_full, ph, p.err);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float v[16];
        tc_ld16(tPV + lane_off + c * 16, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c * 16 + i] = fmaf(o[c * 16 + i], corr, v[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pv_empty);
    }
    if (row_ok) {
      const float invL = 1.0f / l;
      bf16* dst = p.out + ((size_t)(t0 + t) * p.n_heads + kvh * G + g) * HD;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        __align__(16) bf16 ob[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ob[i] = __float2bfloat16_rn(o[c * 8 + i] * invL);
        *reinterpret_cast<uint4*>(dst + c * 8) = *reinterpret_cast<const uint4*>(ob);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
  }
}

}  // namespace dn
