// dn_megakernel.cuh -- the whole single-token shard step as ONE persistent kernel.
//
// Why: ncu on the per-op kernels (profiles/r01_*) shows every weight-streaming kernel moves
// exactly its algorithmic bytes but is latency bound (59% long-scoreboard stalls, 16 warps/SM)
// and pays ~10-14 us of launch + prologue + ramp + tail per launch, 160 launches per token.
// This kernel removes both (DESIGN.md section 3.2 has the measured history):
//   * one CTA per SM (grid = #SMs, cooperative launch), 12 warps = 3 warpgroups.
//   * warps 8-9 are PRODUCERS: they stream this CTA's share of every weight matrix of every layer
//     (and the lm_head) through a shared-memory ring with TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx, 16 rows x 1024 columns per stage).  No registers are tied up by loads in flight,
//     and because weights are immutable the producers never wait for a phase boundary; they only
//     bound the bytes outstanding (2-3 stages), because a deeper queue delays every barrier poll.
//   * warps 0-7 are CONSUMERS: they wait on a stage's full-barrier and feed their 128-column slice
//     of the 16 rows to the tensor cores (ldmatrix + mma.sync m16n8k16, activation replicated over
//     the n columns, fp32 accumulate), merge the 8 slices per row block in fixed order and run the
//     same fused epilogues as the per-op kernels (RoPE + paged-KV append, residual, SwiGLU,
//     argmax/logsumexp).  setmaxnreg gives them 224 registers, the data-movement warpgroup 56.
//   * phases that depend on activations produced by other CTAs are separated by a grid barrier
//     (monotonic counter, release-arrive, a poller warp acquire-polls; bounded spins so a bug can
//     never hang the GPU); five per layer.  While they wait the consumers keep draining the ring
//     into TENSOR MEMORY (tcgen05.st, 8 stages per warp) and read the fragments back afterwards,
//     so HBM streams through the barriers.
//   * attention: one CTA per (q head, split), K/V rows global -> registers, fp32 online softmax.
// Rounding points and summation structure per output row are fixed -> deterministic.
#pragma once
#include "dn_kernels.cuh"
#include "dn_kvquant.cuh"

namespace dn {

constexpr int MK_CW = 8;                        // consumer warps
constexpr int MK_CTHREADS = MK_CW * 32;         // 256
constexpr int MK_PW = 2;                        // producer warps (alternate ring stages)
constexpr int MK_THREADS = MK_CTHREADS + 4 * 32;   // + one warpgroup: 2 producers, the L2 prefetch warp, the barrier poller
// register re-allocation (setmaxnreg works per warpgroup): the data-movement warpgroup gives most
// of its registers to the 8 consumer warps.  8*32*224 + 4*32*56 = 64512 = 168*384
constexpr int MK_REGS_CONSUMER = 224, MK_REGS_PRODUCER = 56;
// the pool is what the CTA got at launch (168 regs x 384 threads), NOT the whole file: asking for more blocks forever
static_assert(8 * 32 * MK_REGS_CONSUMER + 4 * 32 * MK_REGS_PRODUCER <= 168 * MK_THREADS, "setmaxnreg split exceeds the CTA's launch allocation");
constexpr int MK_ROWS = 16;                     // rows per ring stage = the M of one mma tile
constexpr int MK_MAX_SEG = 1024;                // bf16 columns per row per stage: one 2 KiB TMA op per row
// rows sit 16 bytes further apart than their payload, so the 8 row addresses of an ldmatrix fall
// into 8 different bank groups (a 2 KiB pitch would put them all on the same 4 banks)
constexpr int MK_ROW_PITCH = MK_MAX_SEG * 2 + 16;   // 2064
constexpr int MK_STAGE_BYTES = MK_ROWS * MK_ROW_PITCH;   // 33024
constexpr int MK_MAX_STAGES = 8;
// r01 measurement: with 512-byte bulk copies the step ran at 2.2 TB/s (one TMA op per ~66
// cycles per SM is the limit, not bytes), so a stage is 16 rows x 1024 columns = 16 ops of 2 KiB.
#ifndef MK_OPT_PRE
#define MK_OPT_PRE 1   // issue the epilogue's residual load before the row block's stages
#endif
constexpr int MK_DBG_WORDS = 32;   // per (CTA, layer): 15 phase stamps, [16..19] producer-blocked ns, [24..27] consumer-wait ns
constexpr unsigned long long MK_TIMEOUT_NS = 4000000000ull;  // bounded spins

enum { MK_W_Q = 0, MK_W_K, MK_W_V, MK_W_O, MK_W_GATE, MK_W_UP, MK_W_DOWN, MK_W_LN1, MK_W_LN2, MK_W_QB, MK_W_KB, MK_W_VB, MK_W_N };

struct MkLayer {
  const bf16* w[MK_W_N];
  bf16* kv_pool;          // this layer's pages
};

struct MkParams {
  const MkLayer* layers;  // device array, execution order
  int n_layers;
  int H, FFN, n_heads, n_kv, vocab, nsplit;
  float eps;
  const bf16* x_in;       // [H]; ignored when embed != nullptr
  const bf16* embed;      // first shard: x_in = embed[st->token]
  bf16 *xa, *xb, *hbuf, *qbuf, *attn, *act;
  bf16* x_out;            // [H] result of the last layer (may alias x_in)
  const int32_t* block_table;
  StepState* st;
  const float* inv_freq;
  float* part;
  unsigned int* tickets;
  const bf16 *norm_w, *head_w;
  bf16* logits_bf16;
  float* logits_f32;
  HeadPartial* head_part;
  unsigned int* head_ticket;
  int32_t* token_out;
  float* logprob_out;
  int do_head, advance;
  unsigned int *bar_count, *bar_epoch, *err;   // monotonic arrival counter, its value at launch start
  int n_stages;
  int pf_depth;           // L2 prefetch look-ahead of the producer, in ring stages
  int flags;              // bit0: two-word (count + generation) grid barrier; bit1: nested-loop producer without L2 prefetch
  unsigned int* bar_gen;
  // fused ring hop (all optional): wait for the predecessor's flag before touching the input,
  // publish the result into the successor's slot + flag at the end -- no separate hop kernels
  const uint32_t* wait_flag; uint32_t wait_seq;
  const int32_t* token_in;      // first shard: token id lives here (the hop slot) instead of st->token
  void* send_dst; uint32_t* send_flag; uint32_t send_seq;
  int park;                  // 1: consumer warps park ready ring stages in tensor memory while they wait at a grid barrier
  int inflight_hi;           // cap while the consumers are starving for weights (>= inflight)
  int inflight;              // producer: at most this many ring stages with loads outstanding (0 = whole ring)
  int attn_chunk;            // minimum tokens per attention split (multiple of 32)
  const int* bounds;         // optional [4 phases][grid+1] row boundaries (calibrated partition), else equal split
  unsigned long long* dbg;   // optional [grid][n_layers][16] globaltimer stamps of CTA thread 0 (mk_debug)
  int scratch_bytes;      // shared scratch (activation vector / attention tiles)
  int attn_tc;               // 1: long-context attention phase on mma.sync with K/V tiles staged in shared memory (mk_attention_tc)
  int attn_last;             // plain attention, S > 1: the last-arriving split CTA merges its head (else every CTA merges all heads while staging o_proj)
  int attn_single;           // plain attention: contexts up to this many tokens keep one CTA per head (no split merge)
  // quantised KV (dn_kvquant.cuh): 0 = bf16 pages; 4 / 8 = packed pages, two-pass attention
  int kv_bits;
  bf16* kv_stage;            // bf16 staging pool the q/k/v epilogue writes the new K/V row into
  float* sc_buf;             // [n_heads][sc_stride] bf16-rounded scores of the current layer
  int sc_stride;
  unsigned int* head_tk;     // [n_heads] monotonic arrival counters: the splits of a head exchange (max, sum)
  // ---- tensor-parallel lm_head over the ring (dn_shard_step_tp).  Every shard holds
  //      vocab/S rows of the lm_head; the last shard broadcasts the final hidden state of a token to all
  //      shards over NVLink, each computes (max, sum-exp, argmax) of its slice and stores that partial into the
  //      head shard's table, which merges them -- so no shard streams the whole 1 GB head.
  const bf16* hp_x;            // head part of THIS launch: final hidden state of the due nonce (local slot); null = none
  const uint32_t* hp_wait_flag; uint32_t hp_seq;   // arrival flag of hp_x (released by the last shard's broadcast)
  int hp_row0, head_rows;      // vocabulary rows of the local slice: [hp_row0, hp_row0 + head_rows) (head_w points at the slice)
  float* hp_dst; uint32_t* hp_dst_flag;            // this shard's entry of the head shard's partial table + its flag
  int bc_n; bf16* bc_dst[16]; uint32_t* bc_flag[16]; uint32_t bc_seq;   // last shard: where x_out is broadcast to
  int mg_n; const float* mg_part; const uint32_t* mg_flags; uint32_t mg_seq;   // head shard: partial table [16][4], flags 64 B apart
  StepState* mg_st; int32_t* mg_token_out; float* mg_logprob_out;
  int32_t* mg_slot; uint32_t* mg_slot_flag; uint32_t mg_slot_seq;      // own lane slot: the token for the next step's token_in
};

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded wait: on timeout flag the error and fall through (garbage results, never a hang)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, unsigned int* err) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = gtimer();
  unsigned it = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++it & 1023u) == 0 && gtimer() - t0 > MK_TIMEOUT_NS) {
      atomicExch(err, 2u);
      return;
    }
  }
}
// debug variant: also accumulates the time spent waiting into *acc_ns (mk_debug only)
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, unsigned int* err, unsigned long long* acc_ns) {
  if (acc_ns == nullptr) { mbar_wait(bar, parity, err); return; }
  {   // test_wait never suspends (try_wait may block for a hardware time slice and hide the wait)
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) return;
  }
  const unsigned long long t0 = gtimer();
  mbar_wait(bar, parity, err);
  *acc_ns += gtimer() - t0;
}

// ---------------------------------------------------------------------------------
// tensor memory as a second-level weight buffer.  While the consumers sit in a grid barrier the
// ring would fill up and HBM would go idle; instead every consumer warp keeps draining ready
// stages: it runs the same 8 ldmatrix it would run to consume its column slice of the stage and
// parks the 32 fragment registers in its own TMEM region (lanes of its quarter, 32 columns per
// stage, 8 stages per warp = all 256 KiB), then frees the ring slot.  After the barrier the warp
// reads the fragments back (tcgen05.ld) in the same order and feeds them to the same mma sequence,
// so results do not depend on where a stage waited.
// ---------------------------------------------------------------------------------
constexpr int MK_PARK_SLOTS = 8;
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {     // never suspends
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void st_release_cta(unsigned int* sp, unsigned int v) {
  asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(sp)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_cta(const unsigned int* sp) {
  unsigned int v;
  asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(sp)) : "memory");
  return v;
}
// L2 eviction policies: weights are read exactly once per step, so demand loads are marked
// evict-first (they must not push KV pages, activations or prefetched tiles out of L2) and
// look-ahead prefetches evict-last (they must survive until the ring asks for them)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
// TMA bulk copy global -> shared, completion counted on an mbarrier
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
// bulk L2 prefetch: pulls the bytes toward L2 without occupying a ring stage
__device__ __forceinline__ void tma_prefetch_l2(const void* src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void cbar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(MK_CTHREADS) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_gpu(const unsigned int* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------
// phase description shared by producer and consumers (identical traversal order)
// ---------------------------------------------------------------------------------
enum { PH_QKV = 0, PH_O = 1, PH_GU = 2, PH_DOWN = 3, PH_HEAD = 4 };

struct MkPhase {
  int K, nrows, align, seg;   // seg: columns per row per stage (largest of 1024/512/256 dividing K)
};
__device__ __forceinline__ MkPhase mk_phase(const MkParams& p, int ph) {
  MkPhase d;
  switch (ph) {
    case PH_QKV: d.K = p.H; d.nrows = (p.n_heads + 2 * p.n_kv) * HD; d.align = 2; break;
    case PH_O: d.K = p.n_heads * HD; d.nrows = p.H; d.align = 1; break;
    case PH_GU: d.K = p.H; d.nrows = 2 * p.FFN; d.align = 2; break;
    case PH_DOWN: d.K = p.FFN; d.nrows = p.H; d.align = 1; break;
    default: d.K = p.H; d.nrows = p.head_rows; d.align = 1; break;
  }
  d.seg = (d.K % 1024 == 0) ? 1024 : ((d.K % 512 == 0) ? 512 : 256);
  return d;
}
__device__ __forceinline__ const bf16* mk_row(const MkParams& p, const MkLayer& L, int ph, int vr, int K) {
  switch (ph) {
    case PH_QKV: {
      const int task = vr >> 1, which = vr & 1;
      const int slot = task >> 6, d = (task & 63) + (which << 6);
      if (slot < p.n_heads) return L.w[MK_W_Q] + ((size_t)slot * HD + d) * K;
      if (slot < p.n_heads + p.n_kv) return L.w[MK_W_K] + ((size_t)(slot - p.n_heads) * HD + d) * K;
      return L.w[MK_W_V] + ((size_t)(slot - p.n_heads - p.n_kv) * HD + d) * K;
    }
    case PH_O: return L.w[MK_W_O] + (size_t)vr * K;
    case PH_GU: return ((vr & 1) ? L.w[MK_W_UP] : L.w[MK_W_GATE]) + (size_t)(vr >> 1) * K;
    case PH_DOWN: return L.w[MK_W_DOWN] + (size_t)vr * K;
    default: return p.head_w + (size_t)vr * K;
  }
}
__device__ __forceinline__ void mk_range(const MkParams& p, int ph, const MkPhase& d, int& r0, int& r1) {
  if (p.bounds != nullptr && ph < PH_HEAD) {     // per-SM calibrated partition (dn_step_set_bounds)
    r0 = p.bounds[ph * (gridDim.x + 1) + blockIdx.x];
    r1 = p.bounds[ph * (gridDim.x + 1) + blockIdx.x + 1];
    return;
  }
  const int units = d.nrows / d.align;
  r0 = (int)(((long long)units * blockIdx.x) / gridDim.x) * d.align;
  r1 = (int)(((long long)units * (blockIdx.x + 1)) / gridDim.x) * d.align;
}

struct MkRing {
  volatile unsigned int* starve;   // set by consumer warp 0 while it waits for weights (ring empty), cleared at a grid barrier
  unsigned char* data;   // n_stages * MK_STAGE_BYTES
  uint64_t* full;        // [n_stages]
  uint64_t* empty;       // [n_stages]
  int n_stages;
  int stage;
  uint32_t phase;
  __device__ __forceinline__ void advance() {
    if (++stage == n_stages) { stage = 0; phase ^= 1u; }
  }
};

// ---------------------------------------------------------------------------------
// producers: MK_PW warps walk this CTA's ring stages in execution order (layers x {QKV, O,
// GATE/UP, DOWN} x row blocks x K segments, then the head); warp `which` feeds the stages
// whose running index has that parity, so two TMA issue loops run concurrently (one warp
// issuing sixteen 2 KiB bulk copies per stage sustains ~56 GB/s per SM -- enough for the HBM
// share but not for refilling the ring from L2 after a stall).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void mk_produce_phase(const MkParams& p, const MkLayer& L, int ph, MkRing& ring, int lane,
                                                 int which, unsigned int& idx, int li) {
  const uint64_t pol = l2_policy_evict_first();
  unsigned long long stall = 0;
  unsigned long long* accp = (p.dbg != nullptr && which == 0) ? &stall : nullptr;
  const MkPhase d = mk_phase(p, ph);
  int r0, r1;
  mk_range(p, ph, d, r0, r1);
  const int nseg = d.K / d.seg;
  const uint32_t rowbytes = (uint32_t)d.seg * 2u;
  for (int rb = r0; rb < r1; rb += MK_ROWS) {
    const int nv = min(MK_ROWS, r1 - rb);
    const bf16* src = (lane < nv) ? mk_row(p, L, ph, rb + lane, d.K) : nullptr;
    for (int sg = 0; sg < nseg; ++sg) {
      if ((int)(idx % MK_PW) == which) {
        const int cap = (*ring.starve != 0u) ? p.inflight_hi : p.inflight;
        if (cap > 0 && idx >= (unsigned)cap) {
          // bound the bytes in flight (not the buffered bytes): a deep queue of outstanding bulk loads
          // is what the grid barrier's polls and every staging load have to wait behind
          const unsigned int j = idx - (unsigned)cap;
          mbar_wait(&ring.full[j % (unsigned)ring.n_stages], (j / (unsigned)ring.n_stages) & 1u, p.err);
        }
        mbar_wait_dbg(&ring.empty[ring.stage], ring.phase ^ 1u, p.err, accp);
        if (lane == 0) mbar_arrive_expect_tx(&ring.full[ring.stage], (uint32_t)nv * rowbytes);
        __syncwarp();
        if (lane < nv)
          tma_bulk_g2s(ring.data + (size_t)ring.stage * MK_STAGE_BYTES + (size_t)lane * MK_ROW_PITCH,
                       src + (size_t)sg * d.seg, rowbytes, &ring.full[ring.stage], pol);
      }
      ++idx;
      ring.advance();
    }
  }
  // mk_debug: ns producer 0 spent blocked on a full ring while producing this (layer, phase)
  if (accp != nullptr && lane == 0 && ph < 4) p.dbg[((size_t)blockIdx.x * p.n_layers + li) * MK_DBG_WORDS + 16 + ph] = stall;
}
__device__ __forceinline__ void mk_producer(const MkParams& p, MkRing& ring, int lane, int which) {
  unsigned int idx = 0;
  if (p.hp_x != nullptr) {       // tensor-parallel head part runs FIRST (its input is another nonce's, already here)
    const MkLayer L0 = p.layers[0];
    mk_produce_phase(p, L0, PH_HEAD, ring, lane, which, idx, 0);
  }
  for (int li = 0; li < p.n_layers; ++li) {
    const MkLayer L = p.layers[li];
    mk_produce_phase(p, L, PH_QKV, ring, lane, which, idx, li);
    mk_produce_phase(p, L, PH_O, ring, lane, which, idx, li);
    mk_produce_phase(p, L, PH_GU, ring, lane, which, idx, li);
    mk_produce_phase(p, L, PH_DOWN, ring, lane, which, idx, li);
  }
  if (p.do_head) {
    const MkLayer L0 = p.layers[0];
    mk_produce_phase(p, L0, PH_HEAD, ring, lane, which, idx, 0);
  }
}

// L2 prefetch warp: walks the same sequence `limit` stages ahead of consumption and issues bulk
// L2 prefetches, so HBM keeps streaming into L2 while the ring is full (consumers inside a grid
// barrier / staging / attention) and the ring then refills at L2 speed.
__device__ __forceinline__ void mk_prefetch_phase(const MkParams& p, const MkLayer& L, int ph, int lane, unsigned int& n,
                                                  volatile unsigned int* consumed, unsigned int limit, bool& alive) {
  const uint64_t pol = l2_policy_evict_last();
  const MkPhase d = mk_phase(p, ph);
  int r0, r1;
  mk_range(p, ph, d, r0, r1);
  // One prefetch op covers up to 8 KiB of a row (4 ring stages), not one 2 KiB segment: the TMA
  // unit retires ~1 bulk op per 46 cycles whatever its size, and a prefetch per segment doubles
  // the op count (measured: look-ahead >= 16 stages dropped the step to 4.9 ms).
  const int piece_cols = 4 * d.seg;                       // columns per prefetch op
  const int npieces = (d.K + piece_cols - 1) / piece_cols;
  for (int rb = r0; rb < r1 && alive; rb += MK_ROWS) {
    const int nv = min(MK_ROWS, r1 - rb);
    const bf16* src = (lane < nv) ? mk_row(p, L, ph, rb + lane, d.K) : nullptr;
    for (int pc = 0; pc < npieces; ++pc) {
      const int cols = min(piece_cols, d.K - pc * piece_cols);
      if ((int)(n - *consumed) > (int)limit) {
        const unsigned long long t0 = gtimer();
        while ((int)(n - *consumed) > (int)limit) {
          __nanosleep(64);
          if (gtimer() - t0 > MK_TIMEOUT_NS) { alive = false; return; }
        }
      }
      if (src != nullptr) tma_prefetch_l2(src + (size_t)pc * piece_cols, (uint32_t)cols * 2u, pol);
      n += (unsigned int)(cols / d.seg);                  // ring stages covered by this piece
    }
  }
}
__device__ __forceinline__ void mk_prefetcher(const MkParams& p, int lane, volatile unsigned int* consumed) {
  if (p.pf_depth <= 0) return;
  unsigned int n = 0;
  bool alive = true;
  const unsigned int limit = (unsigned int)(p.pf_depth + p.n_stages);
  if (p.hp_x != nullptr) {
    const MkLayer L0 = p.layers[0];
    mk_prefetch_phase(p, L0, PH_HEAD, lane, n, consumed, limit, alive);
  }
  for (int li = 0; li < p.n_layers && alive; ++li) {
    const MkLayer L = p.layers[li];
    mk_prefetch_phase(p, L, PH_QKV, lane, n, consumed, limit, alive);
    mk_prefetch_phase(p, L, PH_O, lane, n, consumed, limit, alive);
    mk_prefetch_phase(p, L, PH_GU, lane, n, consumed, limit, alive);
    mk_prefetch_phase(p, L, PH_DOWN, lane, n, consumed, limit, alive);
  }
  if (p.do_head && alive) {
    const MkLayer L0 = p.layers[0];
    mk_prefetch_phase(p, L0, PH_HEAD, lane, n, consumed, limit, alive);
  }
}

// ---------------------------------------------------------------------------------
// consumer: one GEMV phase on the tensor cores.  A ring stage holds 16 weight rows x seg columns;
// consumer warp w owns the column slice [w*seg/8, (w+1)*seg/8) of all 16 rows and feeds it to
// mma.m16n8k16 (A = weights via ldmatrix, B = the activation slice replicated over the 8 n columns,
// fp32 accumulators).  The bf16 x bf16 products are exact and accumulate in fp32, as in the scalar
// formulation; per stage a warp issues 8 ldmatrix + 16 LDS.32 + 8 mma instead of ~170 ALU ops,
// which is what lets the phase run at the HBM rate (the scalar loop was the bottleneck: the ring
// stayed full).  At the end of a 16-row block the 8 per-warp partial sums of each row are added in
// fixed warp order (deterministic) by the block's epilogue warp (rotating), whose lanes 0..15 own
// rows rb+lane; the RoPE / SwiGLU partner row sits in the adjacent lane.
//   pre(vr, owner)  -> value   issued before the block's stages (hides the epilogue's global load)
//   epi(vr, v, owner, value)   called by all 32 lanes of the epilogue warp
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// per-warp consumer position: stages [cons_n, ring_n) of this warp's column slice are parked in TMEM
struct MkCons {
  uint32_t tmem;          // this warp's TMEM region (lane quarter | column half)
  bool park;              // parking needs seg == 1024 in every phase (one fragment geometry)
  unsigned int cons_n;    // stages consumed so far
  unsigned int ring_n;    // stages taken out of the shared-memory ring (consumed or parked)
  uint32_t a_lane;        // ldmatrix lane offset at seg == 1024
};
__device__ __forceinline__ bool mk_try_park(MkRing& ring, MkCons& cs, int lane) {
  if (cs.ring_n - cs.cons_n >= (unsigned)MK_PARK_SLOTS) return false;
  if (!mbar_test_wait(&ring.full[ring.stage], ring.phase)) return false;
  uint32_t r[32];
  const uint32_t a_base = smem_u32(ring.data) + (uint32_t)ring.stage * MK_STAGE_BYTES + cs.a_lane;
#pragma unroll
  for (int j = 0; j < 8; ++j) ldmatrix_x4(a_base + j * 32, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  tmem_st32(cs.tmem + (cs.ring_n % MK_PARK_SLOTS) * 32u, r);
  __syncwarp();
  if (lane == 0) mbar_arrive(&ring.empty[ring.stage]);
  ring.advance();
  ++cs.ring_n;
  return true;
}

// ---------------------------------------------------------------------------------
// grid-wide barrier between dependent phases.  All CTAs are co-resident (cooperative launch,
// grid == #SMs).  The arrival counter only ever grows: barrier k of this launch is passed when it
// reaches base + (k+1)*grid, base being the value the previous launch published in bar_epoch.
// Consumer thread 0 release-arrives and posts a request to the poller warp, which acquire-polls
// the counter and publishes completion in shared memory; the 8 consumer warps meanwhile park ready
// ring stages in TMEM, so the weight stream keeps flowing through the wait.
// Cross-CTA activations are read with ld.global.cg (L2), so no L1 invalidation is needed.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void mk_grid_barrier(const MkParams& p, unsigned int& k, MkRing& ring, MkCons& cs, unsigned int* bar_req,
                                                unsigned int* bar_done, int lane) {
  cbar_sync();          // every consumer thread's writes precede thread 0's release (cumulativity)
  if (threadIdx.x == 0) {
    *ring.starve = 0u;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar_count) : "memory");
    if (!(p.flags & 8)) st_release_cta(bar_req, k + 1u);
  }
  ++k;
  if (p.flags & 8) return;     // flags bit3: timing experiment (streaming ceiling): no grid-wide wait
  const unsigned long long t0 = gtimer();
  unsigned it = 0;
  for (;;) {
    unsigned int done = 0;
    if (lane == 0) done = ld_acquire_cta(bar_done);
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done == k) break;
    const bool parked = cs.park && mk_try_park(ring, cs, lane);
    if (!parked && (++it & 1023u) == 0 && gtimer() - t0 > MK_TIMEOUT_NS) { atomicExch(p.err, 3u); break; }
  }
  __syncwarp();
}
// the poller warp (one lane): serves the consumers' barrier requests until told to exit
__device__ __forceinline__ void mk_barrier_poller(const MkParams& p, unsigned int* bar_req, unsigned int* bar_done) {
  unsigned int served = 0;
  for (;;) {
    unsigned int req;
    const unsigned long long t0 = gtimer();
    unsigned it = 0;
    while ((req = ld_acquire_cta(bar_req)) == served) {
      __nanosleep(32);
      if ((++it & 4095u) == 0 && gtimer() - t0 > 8ull * MK_TIMEOUT_NS) return;
    }
    if (req == 0xffffffffu) return;
    const unsigned int target = bar_req[2] + req * gridDim.x;      // launch base, stored by consumer thread 0 before its first request
    if ((int)(ld_acquire_gpu(p.bar_count) - target) < 0) {
      const unsigned long long t1 = gtimer();
      unsigned it2 = 0;
      while ((int)(ld_acquire_gpu(p.bar_count) - target) < 0) {
        if ((++it2 & 255u) == 0 && gtimer() - t1 > MK_TIMEOUT_NS) { atomicExch(p.err, 3u); break; }
      }
    }
    st_release_cta(bar_done, req);
    served = req;
  }
}

template <class Pre, class Epi>
__device__ __forceinline__ void mk_consume(const MkParams& p, int ph, int li, MkRing& ring, const bf16* xs, float* red2, int cw, int lane,
                                           volatile unsigned int* consumed, MkCons& cs, unsigned int& nblk, Pre pre, Epi epi) {
  const MkPhase d = mk_phase(p, ph);
  int r0, r1;
  mk_range(p, ph, d, r0, r1);
  const int nseg = d.K / d.seg;
  const int slice = d.seg >> 3;                 // columns per warp per stage: 128 / 64 / 32
  const int ksteps = slice >> 4;
  unsigned long long waited = 0;
  unsigned long long* accp = (p.dbg != nullptr && cw == 0) ? &waited : nullptr;
  // ldmatrix: lane l addresses row (l&7) + 8*((l>>3)&1) of the 16, 16-byte chunk (l>>4) of the k16 step
  const uint32_t a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * MK_ROW_PITCH + (lane >> 4) * 16 + cw * slice * 2);
  const uint32_t ring_base = smem_u32(ring.data);
  const int t = lane & 3, g = lane >> 2;
  for (int rb = r0; rb < r1; rb += MK_ROWS) {
    const int nv = min(MK_ROWS, r1 - rb);
    const int ew = (int)(nblk & (MK_CW - 1)), buf = (int)(nblk & 1u);
    const bool owner = (cw == ew) && lane < nv;
    const auto pv = pre(rb + lane, owner);
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
    for (int sg = 0; sg < nseg; ++sg) {
      const uint32_t* xw = reinterpret_cast<const uint32_t*>(xs + (size_t)sg * d.seg + cw * slice) + t;
      if (cs.cons_n != cs.ring_n) {
        // this stage was parked in TMEM during a barrier: same fragments, same mma order
        uint32_t r[32];
        tmem_ld32(cs.tmem + (cs.cons_n % MK_PARK_SLOTS) * 32u, r);
        if (!(p.flags & 4)) {
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            mma_bf16_16816(c0, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], xw[j * 8], xw[j * 8 + 4]);
            mma_bf16_16816(c1, r[4 * j + 4], r[4 * j + 5], r[4 * j + 6], r[4 * j + 7], xw[j * 8 + 8], xw[j * 8 + 12]);
          }
        }
      } else {
        if (cw == 0 && !mbar_test_wait(&ring.full[ring.stage], ring.phase)) { if (lane == 0) *ring.starve = 1u; }
        mbar_wait_dbg(&ring.full[ring.stage], ring.phase, p.err, accp);
        if (!(p.flags & 4)) {     // flags bit2: timing experiment, skip the math (results are garbage)
          const uint32_t a_base = ring_base + (uint32_t)ring.stage * MK_STAGE_BYTES + a_lane;
          if (ksteps == 8) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              uint32_t a0, a1, a2, a3, e0, e1, e2, e3;
              ldmatrix_x4(a_base + j * 32, a0, a1, a2, a3);
              ldmatrix_x4(a_base + j * 32 + 32, e0, e1, e2, e3);
              mma_bf16_16816(c0, a0, a1, a2, a3, xw[j * 8], xw[j * 8 + 4]);
              mma_bf16_16816(c1, e0, e1, e2, e3, xw[j * 8 + 8], xw[j * 8 + 12]);
            }
          } else {
            for (int j = 0; j < ksteps; ++j) {
              uint32_t a0, a1, a2, a3;
              ldmatrix_x4(a_base + j * 32, a0, a1, a2, a3);
              mma_bf16_16816(c0, a0, a1, a2, a3, xw[j * 8], xw[j * 8 + 4]);
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&ring.empty[ring.stage]);
        ring.advance();
        ++cs.ring_n;
      }
      ++cs.cons_n;
      if (threadIdx.x == 0) *consumed = cs.cons_n;      // progress signal for the L2 prefetch warp
    }
    // every column of the accumulator tile holds the same dot product: lanes with t == 0 publish
    // rows g (c[0]) and g + 8 (c[2]) of this warp's column slice
    if (t == 0) {
      float* r = red2 + (buf * MK_CW + cw) * MK_ROWS;
      r[g] = c0[0] + c1[0];
      r[g + 8] = c0[2] + c1[2];
    }
    cbar_sync();
    if (cw == ew) {
      float v = 0.f;
      if (lane < MK_ROWS) {
#pragma unroll
        for (int w = 0; w < MK_CW; ++w) v += red2[(buf * MK_CW + w) * MK_ROWS + lane];
      }
      epi(rb + lane, v, owner, pv);
    }
    ++nblk;
  }
  // mk_debug: ns consumer warp 0 waited for weights (ring empty = HBM-bound time) in this phase
  if (accp != nullptr && lane == 0 && ph < 4) p.dbg[((size_t)blockIdx.x * p.n_layers + li) * MK_DBG_WORDS + 24 + ph] = waited;
}

// stage a bf16 vector [K] from global (produced by other CTAs: bypass L1) into shared memory
__device__ __forceinline__ void mk_stage_copy(bf16* xs, const bf16* src, int K) {
  const int tid = threadIdx.x;
  for (int i = tid; i < K / 8; i += MK_CTHREADS)
    reinterpret_cast<uint4*>(xs)[i] = __ldcg(reinterpret_cast<const uint4*>(src) + i);
  cbar_sync();
}
// RMSNorm staging (same arithmetic and summation structure as stage_rmsnorm<1>); x and w are
// fetched once and kept in registers across the block reduction (K <= 8192), so the phase
// start costs one L2 round trip instead of two.
__device__ __forceinline__ void mk_stage_rmsnorm(bf16* xs, float* scratch, const bf16* src, const bf16* w, int K, float eps) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int MAXIT = 4;
  uint4 gv[MAXIT];
  float ss = 0.f;
#pragma unroll
  for (int it = 0; it < MAXIT; ++it) {
    const int i = tid * 8 + it * MK_CTHREADS * 8;
    if (i < K) {
      const uint4 v = __ldcg(reinterpret_cast<const uint4*>(src + i));
      gv[it] = *reinterpret_cast<const uint4*>(w + i);
      *reinterpret_cast<uint4*>(xs + i) = v;           // raw x parked in shared memory, normalised in place below
      const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
  }
  ss = warp_sum(ss);
  cbar_sync();
  if (lane == 0) scratch[warp] = ss;
  cbar_sync();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < MK_CW; ++i) tot += scratch[i];
  const float inv = 1.0f / sqrtf(tot / (float)K + eps);
#pragma unroll
  for (int it = 0; it < MAXIT; ++it) {
    const int i = tid * 8 + it * MK_CTHREADS * 8;
    if (i < K) {
      const uint4 v = *reinterpret_cast<const uint4*>(xs + i), g = gv[it];
      const float f[8] = {bf_lo(v.x), bf_hi(v.x), bf_lo(v.y), bf_hi(v.y), bf_lo(v.z), bf_hi(v.z), bf_lo(v.w), bf_hi(v.w)};
      const float gw[8] = {bf_lo(g.x), bf_hi(g.x), bf_lo(g.y), bf_hi(g.y), bf_lo(g.z), bf_hi(g.z), bf_lo(g.w), bf_hi(g.w)};
      __align__(16) bf16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __float2bfloat16_rn(__fmul_rn(bf16r(__fmul_rn(f[j], inv)), gw[j]));
      *reinterpret_cast<uint4*>(xs + i) = *reinterpret_cast<const uint4*>(o);
    }
  }
  cbar_sync();
}

// attention geometry shared by the attention phase and the o_proj staging.  The context is cut
// into 32-token tiles; one CTA serves one (q head, CTA split) and its 8 consumer warps take the
// split's tiles round-robin, merging through shared memory.  A second CTA split per head is only
// added once every warp already has `attn_chunk` tokens, so short contexts need no cross-CTA
// merge at all (S == 1: the CTA writes the normalised head output directly).
// The attention phase exists in four forms; the step kernel is instantiated once per form (template MODE) so that the
// common one (16-bit KV, short/medium contexts) carries none of the others' code: measured, the all-in-one kernel was
// 0.5 % slower at a 128-token context for that alone.
constexpr int MK_ATT_PLAIN = 0;   // 16-bit pages, CUDA-core tiles straight from global memory
constexpr int MK_ATT_TC = 1;      // 16-bit pages, long contexts: shared-memory tiles + mma.sync for the GQA group
constexpr int MK_ATT_Q8 = 2;      // 8-bit affine-quantised pages, two-pass
constexpr int MK_ATT_Q4 = 3;      // 4-bit
constexpr int ATC_TOK = 16;                           // tokens per warp tile of the tensor-core attention
constexpr int ATC_PITCH = HD * 2 + 16;                // 272-byte rows: the 8 row addresses of an ldmatrix hit 8 different bank groups
constexpr int ATC_TILE_BYTES = ATC_TOK * ATC_PITCH;   // 4,352
constexpr int ATC_WARP_BYTES = 2 * ATC_TILE_BYTES;    // K + V
constexpr int ATC_WARPS = 6;                          // warps that own tiles; all 8 keep draining the weight ring into tensor memory
constexpr int ATC_SMEM_BYTES = ATC_WARPS * ATC_WARP_BYTES;   // 52,224: lives in the activation scratch (dead during attention)
__device__ __forceinline__ void mk_attn_geometry(const MkParams& p, int kv_len, int& S, int& tps) {
  if (p.attn_tc) {
    // tensor-core path: one CTA per (kv head, split) serves the whole GQA group; tiles of 16 tokens
    const int n_tiles = (kv_len + ATC_TOK - 1) / ATC_TOK;
    int smax = (int)gridDim.x / p.n_kv;
    smax = max(1, min(smax, p.nsplit));
    const int per_cta = ATC_WARPS * 4;                // a second split once every tile warp has 4 tiles (64 tokens)
    int s = (n_tiles + per_cta - 1) / per_cta;
    s = max(1, min(s, smax));
    tps = (n_tiles + s - 1) / s;
    S = (n_tiles + tps - 1) / tps;
    return;
  }
  const int n_tiles = (kv_len + 31) >> 5;
  if (n_tiles <= (p.attn_single >> 5)) {
    // short contexts: ONE CTA per head even when its warps need a second or third tile -- splitting a head over CTAs
    // costs the merge while staging o_proj (measured 3.5 us per layer against 0.85 us for the plain copy), an extra
    // tile pass costs ~1 us
    S = 1; tps = n_tiles;
    return;
  }
  int smax = (int)gridDim.x / p.n_heads;
  smax = max(1, min(smax, min(p.nsplit, 8)));
  const int per_cta = MK_CW * max(1, p.attn_chunk >> 5);
  int s = (n_tiles + per_cta - 1) / per_cta;
  s = max(1, min(s, smax));
  tps = (n_tiles + s - 1) / s;
  S = (n_tiles + tps - 1) / tps;
}

// ---------------------------------------------------------------------------------
// attention phase: K and V rows go global -> registers (one 256-byte row per warp-wide load,
// 32 rows in flight), no shared-memory staging and no CTA-wide sync inside the tile loop
// ---------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ void mk_attention(const MkParams& p, const MkLayer& L, unsigned char* scratch, int cw, int lane,
                                             int S, int tps, int tile_first, int phys_first) {
  const int kv_len = p.st->pos + 1;
  const int n_tiles = (kv_len + 31) >> 5;
  float* wpart = reinterpret_cast<float*>(scratch);           // [MK_CW][132]: o[128], m, l per warp
  const float scale = 0.08838834764831845f;
  for (int task = blockIdx.x; task < p.n_heads * S; task += gridDim.x) {     // CTA-uniform
    const int head = task / S, sp = task % S;
    const int kvh = head / G;
    const int tile1 = min(n_tiles, (sp + 1) * tps);
    float qv[4];
    {
      const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p.qbuf + head * HD + lane * 4));
      qv[0] = __fmul_rn(bf_lo(u.x), scale); qv[1] = __fmul_rn(bf_hi(u.x), scale);
      qv[2] = __fmul_rn(bf_lo(u.y), scale); qv[3] = __fmul_rn(bf_hi(u.y), scale);
    }
    float m = -INFINITY, l = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int tile = sp * tps + cw; tile < tile1; tile += MK_CW) {
      const int t0 = tile << 5;
      const int nt = min(32, kv_len - t0);
      // a 32-token tile never straddles a 64-token page; the first tile's page was looked up at kernel start
      const int phys = (task == (int)blockIdx.x && tile == tile_first) ? phys_first : p.block_table[t0 / PAGE];
      const bf16* kbase = L.kv_pool + (((size_t)phys * 2) * p.n_kv + kvh) * (PAGE * HD) + (size_t)(t0 % PAGE) * HD + lane * 4;
      const bf16* vbase = kbase + (size_t)p.n_kv * (PAGE * HD);
      const int pf_tile = tile + MK_CW;                         // one tile ahead (two ahead measured no better)
      if (lane == 0 && pf_tile < tile1) {
        // pull this warp's next tile toward L2 while this one is computed (no registers, one lane, two bulk
        // ops): measured +2.7 % at an 8K context, +6.4 % at 32K
        const int tn = pf_tile << 5;
        const int pn = p.block_table[tn / PAGE];
        const bf16* kn = L.kv_pool + (((size_t)pn * 2) * p.n_kv + kvh) * (PAGE * HD) + (size_t)(tn % PAGE) * HD;
        const uint32_t nb = (uint32_t)min(32, kv_len - tn) * HD * 2u;
        const uint64_t polk = l2_policy_evict_last();
        tma_prefetch_l2(kn, nb, polk);
        tma_prefetch_l2(kn + (size_t)p.n_kv * (PAGE * HD), nb, polk);
      }
      float sc[32];
      uint2 kr[32], vr[32];                                     // 64 independent 8-byte loads in flight per lane
#pragma unroll
      for (int j = 0; j < 32; ++j)
        kr[j] = (j < nt) ? __ldcg(reinterpret_cast<const uint2*>(kbase + j * HD)) : make_uint2(0u, 0u);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        vr[j] = (j < nt) ? __ldcg(reinterpret_cast<const uint2*>(vbase + j * HD)) : make_uint2(0u, 0u);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        sc[j] = fmaf(qv[0], bf_lo(kr[j].x), fmaf(qv[1], bf_hi(kr[j].x), fmaf(qv[2], bf_lo(kr[j].y), qv[3] * bf_hi(kr[j].y))));
      transpose_reduce32(sc, lane);
      const bool valid = lane < nt;
      const float s = valid ? sc[0] : -INFINITY;
      const float m_new = fmaxf(m, warp_max(s));
      const float pj = valid ? exp2f((s - m_new) * LOG2E) : 0.f;
      const float corr = exp2f((m - m_new) * LOG2E);
      l = l * corr + warp_sum(pj);
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) o[dd] *= corr;
      m = m_new;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float pw = __shfl_sync(0xffffffffu, pj, j);
        o[0] = fmaf(pw, bf_lo(vr[j].x), o[0]); o[1] = fmaf(pw, bf_hi(vr[j].x), o[1]);
        o[2] = fmaf(pw, bf_lo(vr[j].y), o[2]); o[3] = fmaf(pw, bf_hi(vr[j].y), o[3]);
      }
    }
    *reinterpret_cast<float4*>(wpart + cw * 132 + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
    if (lane == 0) { wpart[cw * 132 + 128] = m; wpart[cw * 132 + 129] = l; }
    cbar_sync();
    if (threadIdx.x < HD) {
      // merge the warps in fixed order -> deterministic
      const int d = threadIdx.x;
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < MK_CW; ++w) M = fmaxf(M, wpart[w * 132 + 128]);
      float Ls = 0.f, acc = 0.f;
#pragma unroll
      for (int w = 0; w < MK_CW; ++w) {
        const float mw = wpart[w * 132 + 128];
        if (mw != -INFINITY) {
          const float e = exp2f((mw - M) * LOG2E);
          Ls = fmaf(wpart[w * 132 + 129], e, Ls);
          acc = fmaf(wpart[w * 132 + d], e, acc);
        }
      }
      if (S == 1) {
        p.attn[head * HD + d] = __float2bfloat16_rn(acc * (1.0f / Ls));
      } else {
        float* pp = p.part + ((size_t)head * p.nsplit + sp) * PART_STRIDE;
        pp[d] = acc;
        if (d == 0) { pp[128] = M; pp[129] = Ls; }
      }
    }
    if (S > 1 && p.attn_last) {
      // the last split CTA of the head to arrive merges the head's S partials (split order -> the result does not depend
      // on who merges) and writes the normalised bf16 head: the o_proj staging of all CTAs becomes a plain copy
      __threadfence();
      cbar_sync();
      int* flag = reinterpret_cast<int*>(wpart + MK_CW * 132);
      if (threadIdx.x == 0) {
        const unsigned int old = atomicAdd(p.tickets + head, 1u);
        *flag = (old == (unsigned)S - 1u) ? 1 : 0;
        if (*flag) p.tickets[head] = 0u;                       // re-armed for the next layer / launch
      }
      cbar_sync();
      if (*flag && threadIdx.x < HD) {
        __threadfence();
        const int d = threadIdx.x;
        const float* hp = p.part + ((size_t)head * p.nsplit) * PART_STRIDE;
        float Mg = -INFINITY;
        for (int s2 = 0; s2 < S; ++s2) Mg = fmaxf(Mg, __ldcg(hp + (size_t)s2 * PART_STRIDE + 128));
        float Lg = 0.f, acc = 0.f;
        for (int s2 = 0; s2 < S; ++s2) {
          const float ms = __ldcg(hp + (size_t)s2 * PART_STRIDE + 128);
          if (ms == -INFINITY) continue;
          const float wgt = exp2f((ms - Mg) * LOG2E);
          Lg = fmaf(__ldcg(hp + (size_t)s2 * PART_STRIDE + 129), wgt, Lg);
          acc = fmaf(__ldcg(hp + (size_t)s2 * PART_STRIDE + d), wgt, acc);
        }
        p.attn[head * HD + d] = __float2bfloat16_rn(acc * (1.0f / Lg));
      }
    }
    cbar_sync();
  }
}

// ---------------------------------------------------------------------------------
// attention phase on the tensor cores, for long contexts (bf16 pages).  The CUDA-core phase above is instruction
// bound (~23 instructions per token and head, K/V read once per q head of a GQA group): at an 8K context it costs
// ~29 us per layer against 5 us of KV bytes.  Here one CTA serves a (kv head, split) for the whole GQA group:
//   * each of the 8 warps takes 16-token tiles round-robin and stages its K and V tile (2 x 4 KiB, contiguous
//     inside a page) into its own shared-memory buffer with cp.async -- the buffers alias the activation scratch,
//     which is dead between the q/k/v and o_proj phases;
//   * S = Q K^T with mma.sync m16n8k16: A = Q of the group's G heads (rows >= G are zero), fragments held in
//     registers for the whole phase; B = K rows through ldmatrix;
//   * fp32 online softmax in the accumulator-fragment layout (a row's 16 scores sit in one quad);
//   * O += P V with P re-used straight from the S accumulators as the A operand, split into bf16 hi + lo terms (the
//     oracle keeps P in fp32), B = V through ldmatrix.trans;
//   * warps are merged in fixed order; with S > 1 splits the (o, m, l) partials go through `part` exactly like the
//     CUDA-core phase, so the o_proj staging merge is shared.
// K/V bytes are read once per GROUP and ~50 tensor instructions replace ~1,500 ALU instructions per tile and group.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816_top(float& c0, float& c1, float& z0, float& z1, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  // rows 8..15 of A are zero (a1 = a3 = 0): their accumulators z0, z1 stay zero and are shared by every tile
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c0), "+f"(c1), "+f"(z0), "+f"(z1) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <int G>
__device__ __forceinline__ void mk_attention_tc(const MkParams& p, const MkLayer& L, unsigned char* scratch, int cw, int lane,
                                                int S, int tps, MkRing& ring, MkCons& cs, unsigned int* done_ctr) {
  const int kv_len = p.st->pos + 1;
  const int n_tiles = (kv_len + ATC_TOK - 1) / ATC_TOK;
  const float scale = 0.08838834764831845f;
  const int task = blockIdx.x;
  if (task >= p.n_kv * S) return;                              // CTA-uniform
  const int kvh = task / S, sp = task % S;
  const int tile0 = sp * tps, tile1 = min(n_tiles, (sp + 1) * tps);
  const int g = lane >> 2, t = lane & 3;
  unsigned char* kt = scratch + (size_t)cw * ATC_WARP_BYTES;
  unsigned char* vt = kt + ATC_TILE_BYTES;
  const uint32_t kt_s = smem_u32(kt), vt_s = smem_u32(vt);
  // Q fragments of the group's heads (row g = head kvh*G + g), all 8 k-steps of the 128 dims
  uint32_t qa0[8], qa2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    qa0[j] = 0u; qa2[j] = 0u;
    if (g < G) {
      const bf16* qrow = p.qbuf + (size_t)(kvh * G + g) * HD + 16 * j + 2 * t;
      qa0[j] = __ldcg(reinterpret_cast<const uint32_t*>(qrow));
      qa2[j] = __ldcg(reinterpret_cast<const uint32_t*>(qrow + 8));
    }
  }
  float o[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) { o[i][0] = 0.f; o[i][1] = 0.f; }
  float m = -INFINITY, l = 0.f, z0 = 0.f, z1 = 0.f;
  // While a tile is in flight every warp drains ready weight stages into its tensor-memory slice (the same parking the
  // grid barriers use): the phase is latency bound (one L2 / HBM round trip per tile), so HBM keeps streaming the
  // o_proj / gate / up weights underneath it instead of idling with a full ring.
#pragma unroll 1
  for (int tile = tile0 + cw; cw < ATC_WARPS && tile < tile1; tile += ATC_WARPS) {
    const int tk0 = tile * ATC_TOK, nt = min(ATC_TOK, kv_len - tk0);
    const int phys = p.block_table[tk0 / PAGE];
    const unsigned char* kg = reinterpret_cast<const unsigned char*>(L.kv_pool + (((size_t)phys * 2) * p.n_kv + kvh) * (PAGE * HD) + (size_t)(tk0 % PAGE) * HD);
    const unsigned char* vg = kg + (size_t)p.n_kv * (PAGE * HD) * 2;
    __syncwarp();                                              // the previous tile's ldmatrix reads are done
#pragma unroll
    for (int c = 0; c < 8; ++c) {                              // 256 chunks of 16 bytes per tile, 8 per lane
      const int chunk = c * 32 + lane, row = chunk >> 4, col = chunk & 15;
      const uint32_t so = (uint32_t)(row * ATC_PITCH + col * 16);
      if (row < nt) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kt_s + so), "l"(kg + (size_t)row * (HD * 2) + col * 16) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vt_s + so), "l"(vg + (size_t)row * (HD * 2) + col * 16) : "memory");
      } else {                                                 // rows past the context: finite zeros (masked P = 0 must not meet junk)
        *reinterpret_cast<uint4*>(kt + so) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(vt + so) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (cs.park) { if (mk_try_park(ring, cs, lane)) mk_try_park(ring, cs, lane); }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- S = Q K^T for 16 tokens: two n-tiles of 8 tokens
    float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
    {
      const uint32_t ka = kt_s + (uint32_t)(((lane & 7) + ((lane >> 4) & 1) * 8) * ATC_PITCH + ((lane >> 3) & 1) * 16);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(ka + j * 32, b0, b1, b2, b3);
        mma_bf16_16816_top(s0a, s0b, z0, z1, qa0[j], qa2[j], b0, b1);
        mma_bf16_16816_top(s1a, s1b, z0, z1, qa0[j], qa2[j], b2, b3);
      }
    }
    // row g holds tokens {2t, 2t+1} (n-tile 0) and {8+2t, 9+2t} (n-tile 1)
    float sv[4] = {s0a * scale, s0b * scale, s1a * scale, s1b * scale};
    const int tk[4] = {2 * t, 2 * t + 1, 8 + 2 * t, 9 + 2 * t};
#pragma unroll
    for (int i = 0; i < 4; ++i) if (tk[i] >= nt) sv[i] = -INFINITY;
    float mt = fmaxf(fmaxf(sv[0], sv[1]), fmaxf(sv[2], sv[3]));
    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 1));
    mt = fmaxf(mt, __shfl_xor_sync(0xffffffffu, mt, 2));
    const float m_new = fmaxf(m, mt);
    float pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) pv[i] = (sv[i] == -INFINITY) ? 0.f : exp2f((sv[i] - m_new) * LOG2E);
    const float corr = exp2f((m - m_new) * LOG2E);             // m == -inf -> 0
    float ls = (pv[0] + pv[1]) + (pv[2] + pv[3]);
    ls += __shfl_xor_sync(0xffffffffu, ls, 1);
    ls += __shfl_xor_sync(0xffffffffu, ls, 2);
    l = l * corr + ls;
    m = m_new;
#pragma unroll
    for (int i = 0; i < 16; ++i) { o[i][0] *= corr; o[i][1] *= corr; }
    // P as the A operand: hi + lo bf16 terms of the fp32 probabilities
    const uint32_t ph0 = pack_bf16x2(pv[0], pv[1]), ph2 = pack_bf16x2(pv[2], pv[3]);
    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&ph0), h2 = *reinterpret_cast<const __nv_bfloat162*>(&ph2);
    const uint32_t pl0 = pack_bf16x2(pv[0] - __low2float(h0), pv[1] - __high2float(h0));
    const uint32_t pl2 = pack_bf16x2(pv[2] - __low2float(h2), pv[3] - __high2float(h2));
    {
      const uint32_t va = vt_s + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * ATC_PITCH + ((lane >> 4) & 1) * 16);
#pragma unroll
      for (int i = 0; i < 8; ++i) {                            // 16 dims per iteration = two n-tiles
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(va + i * 32, b0, b1, b2, b3);
        mma_bf16_16816_top(o[2 * i][0], o[2 * i][1], z0, z1, ph0, ph2, b0, b1);
        mma_bf16_16816_top(o[2 * i][0], o[2 * i][1], z0, z1, pl0, pl2, b0, b1);
        mma_bf16_16816_top(o[2 * i + 1][0], o[2 * i + 1][1], z0, z1, ph0, ph2, b2, b3);
        mma_bf16_16816_top(o[2 * i + 1][0], o[2 * i + 1][1], z0, z1, pl0, pl2, b2, b3);
      }
    }
  }
  // ---- every warp keeps parking until the tile warps are done (a blocking barrier would stop the ring draining:
  //      a ring slot is only released once all 8 warps took their slice of it)
  if (cw < ATC_WARPS) { __syncwarp(); if (lane == 0) atomicAdd(done_ctr, 1u); }
  if (cs.park) {
    const unsigned long long t0 = gtimer();
    unsigned it = 0;
    for (;;) {
      unsigned int dn = 0;
      if (lane == 0) dn = ld_acquire_cta(done_ctr);
      dn = __shfl_sync(0xffffffffu, dn, 0);
      if (dn >= (unsigned)ATC_WARPS) break;
      const bool parked = mk_try_park(ring, cs, lane);
      if (!parked && (++it & 1023u) == 0 && gtimer() - t0 > MK_TIMEOUT_NS) { atomicExch(p.err, 3u); break; }
    }
  }
  // ---- merge the tile warps (fixed order), per head of the group
  cbar_sync();                                                 // every warp is done with its tile buffers
  if (threadIdx.x == 0) *done_ctr = 0u;                        // re-armed for the next layer (ordered by the barriers below)
  float* wpart = reinterpret_cast<float*>(scratch);            // [ATC_WARPS][G][132] floats (<= 25.3 KB)
  if (cw < ATC_WARPS && g < G) {
    float* row = wpart + ((size_t)cw * G + g) * 132;
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<float2*>(row + 8 * i + 2 * t) = make_float2(o[i][0], o[i][1]);
    if (t == 0) { row[128] = m; row[129] = l; }
  }
  cbar_sync();
  for (int e = threadIdx.x; e < G * HD; e += MK_CTHREADS) {
    const int hg = e / HD, d = e % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATC_WARPS; ++w) M = fmaxf(M, wpart[((size_t)w * G + hg) * 132 + 128]);
    float Ls = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < ATC_WARPS; ++w) {
      const float mw = wpart[((size_t)w * G + hg) * 132 + 128];
      if (mw != -INFINITY) {
        const float ew = exp2f((mw - M) * LOG2E);
        Ls = fmaf(wpart[((size_t)w * G + hg) * 132 + 129], ew, Ls);
        acc = fmaf(wpart[((size_t)w * G + hg) * 132 + d], ew, acc);
      }
    }
    const int head = kvh * G + hg;
    if (S == 1) {
      p.attn[head * HD + d] = __float2bfloat16_rn(acc * (1.0f / Ls));
    } else {
      float* pp = p.part + ((size_t)head * p.nsplit + sp) * PART_STRIDE;
      pp[d] = acc;
      if (d == 0) { pp[128] = M; pp[129] = Ls; }
    }
  }
  if (S > 1) {
    // The last split CTA of this kv head to arrive merges the S partials of its G heads (split order -> the result
    // does not depend on who merges) and writes the normalised bf16 heads, so the o_proj staging of all 148 CTAs is a
    // plain 8 KiB copy instead of every CTA merging n_heads x S partials (measured 12.7 us per layer at S = 18).
    __threadfence();
    cbar_sync();
    int* flag = reinterpret_cast<int*>(wpart + ATC_WARPS * G * 132);
    if (threadIdx.x == 0) {
      const unsigned int old = atomicAdd(p.tickets + kvh, 1u);
      *flag = (old == (unsigned)S - 1u) ? 1 : 0;
      if (*flag) p.tickets[kvh] = 0u;                          // re-armed for the next layer / launch
    }
    cbar_sync();
    if (*flag) {
      __threadfence();
      for (int e = threadIdx.x; e < G * HD; e += MK_CTHREADS) {
        const int hg = e / HD, d = e % HD, head = kvh * G + hg;
        const float* hp = p.part + ((size_t)head * p.nsplit) * PART_STRIDE;
        float Mg = -INFINITY;
        for (int s2 = 0; s2 < S; ++s2) Mg = fmaxf(Mg, __ldcg(hp + (size_t)s2 * PART_STRIDE + 128));
        float Lg = 0.f, acc = 0.f;
        for (int s2 = 0; s2 < S; ++s2) {
          const float ms = __ldcg(hp + (size_t)s2 * PART_STRIDE + 128);
          if (ms == -INFINITY) continue;
          const float wgt = exp2f((ms - Mg) * LOG2E);
          Lg = fmaf(__ldcg(hp + (size_t)s2 * PART_STRIDE + 129), wgt, Lg);
          acc = fmaf(__ldcg(hp + (size_t)s2 * PART_STRIDE + d), wgt, acc);
        }
        p.attn[head * HD + d] = __float2bfloat16_rn(acc * (1.0f / Lg));
      }
    }
  }
  cbar_sync();
}

// ---------------------------------------------------------------------------------
// attention phase over a QUANTISED paged KV (dn_kvquant.cuh has the semantics).  Same task mapping as mk_attention
// (one CTA per (q head, split), 8 warps take the split's 32-token tiles round-robin) but two passes, because
// mlx_lm's quantised attention rounds scores AND probabilities to bf16:
//   A  s_j = bf16(q'.dequant(K_j)) -> score scratch (global, L2-resident); per-warp max; CTA max and sum(exp);
//      with S > 1 splits the CTAs of a head exchange (max, sum) through `part` + a monotonic arrival counter
//   B  p_j = bf16(exp(s_j - M) / L); o += p_j * dequant(V_j); fixed-order sums -> deterministic
// The new token's K/V row was written as bf16 into the staging pool by the q/k/v epilogue; the warp that owns the
// last tile quantises it on the fly (every CTA of the GQA group computes the same codes) and the group's first
// head also stores the packed row into the pool.
// ---------------------------------------------------------------------------------
template <int G, int BITS>
__device__ __forceinline__ void mk_attention_q(const MkParams& p, const MkLayer& L, unsigned char* scratch, int cw, int lane,
                                               int S, int tps, int li, unsigned int tk_base) {
  const int pos = p.st->pos, kv_len = pos + 1;
  const int n_tiles = (kv_len + 31) >> 5;
  float* wpart = reinterpret_cast<float*>(scratch);           // [MK_CW][132]
  const float scale = 0.08838834764831845f;
  unsigned char* pool = reinterpret_cast<unsigned char*>(L.kv_pool);
  const int task = blockIdx.x;
  if (task >= p.n_heads * S) return;                           // CTA-uniform: no task for this CTA
  const int head = task / S, sp = task % S, kvh = head / G;
  const int tile0 = sp * tps, tile1 = min(n_tiles, (sp + 1) * tps);
  float* scb = p.sc_buf + (size_t)head * p.sc_stride;
  float qv[4], qsum;
  {
    const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p.qbuf + head * HD + lane * 4));
    qv[0] = bf16r(__fmul_rn(bf_lo(u.x), scale)); qv[1] = bf16r(__fmul_rn(bf_hi(u.x), scale));
    qv[2] = bf16r(__fmul_rn(bf_lo(u.y), scale)); qv[3] = bf16r(__fmul_rn(bf_hi(u.y), scale));
    qsum = (qv[0] + qv[1]) + (qv[2] + qv[3]);
  }
  // the new token lives in the last tile; its owner quantises it from the staging pool
  const int last_tile = n_tiles - 1, jnew = pos & 31;
  const bool owns_new = last_tile >= tile0 && last_tile < tile1 && ((last_tile - tile0) % MK_CW) == cw;
  float kc[4] = {0.f, 0.f, 0.f, 0.f}, vc[4] = {0.f, 0.f, 0.f, 0.f}, ksc = 0.f, kbi = 0.f, vsc = 0.f, vbi = 0.f;
  if (owns_new) {
    float w[4];
    kvq_read_stage4(kvq_stage_row(p.kv_stage, pos, 0, kvh, p.n_kv), lane, w);
    kvq_quantise_row<BITS>(w, kc, ksc, kbi);
    kvq_read_stage4(kvq_stage_row(p.kv_stage, pos, 1, kvh, p.n_kv), lane, w);
    kvq_quantise_row<BITS>(w, vc, vsc, vbi);
    if (head % G == 0) {
      const int pg = p.block_table[pos / PAGE];
      kvq_store_row<BITS>(const_cast<unsigned char*>(kvq_unit<BITS>(pool, pg, 0, kvh, p.n_kv)), pos % PAGE, lane, kc, ksc, kbi);
      kvq_store_row<BITS>(const_cast<unsigned char*>(kvq_unit<BITS>(pool, pg, 1, kvh, p.n_kv)), pos % PAGE, lane, vc, vsc, vbi);
    }
  }
  // ---- pass A
  float m = -INFINITY;
#pragma unroll 1
  for (int tile = tile0 + cw; tile < tile1; tile += MK_CW) {
    const int t0 = tile << 5, nt = min(32, kv_len - t0);
    const unsigned char* ku = kvq_unit<BITS>(pool, p.block_table[t0 / PAGE], 0, kvh, p.n_kv);
    const bool has_new = tile == last_tile;
    float sc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      sc[j] = 0.f;
      if (j < nt) {
        float c[4], s1, b1;
        if (has_new && j == jnew) { c[0] = kc[0]; c[1] = kc[1]; c[2] = kc[2]; c[3] = kc[3]; s1 = ksc; b1 = kbi; }
        else { kvq_load_codes<BITS>(ku, (t0 % PAGE) + j, lane, c); kvq_load_sb<BITS>(ku, (t0 % PAGE) + j, lane, s1, b1); }
        const float dot = fmaf(qv[0], c[0], fmaf(qv[1], c[1], fmaf(qv[2], c[2], qv[3] * c[3])));
        sc[j] = fmaf(s1, dot, b1 * qsum);
      }
    }
    transpose_reduce32(sc, lane);
    const float s = (lane < nt) ? bf16r(sc[0]) : -INFINITY;
    scb[t0 + lane] = s;
    m = fmaxf(m, warp_max(s));
  }
  if (lane == 0) wpart[cw * 132 + 128] = m;
  cbar_sync();
  float M = -INFINITY;
#pragma unroll
  for (int w = 0; w < MK_CW; ++w) M = fmaxf(M, wpart[w * 132 + 128]);
  float l = 0.f;
#pragma unroll 1
  for (int tile = tile0 + cw; tile < tile1; tile += MK_CW) {
    const float s = scb[(tile << 5) + lane];                   // this thread's own store
    l += warp_sum(s == -INFINITY ? 0.f : exp2f((s - M) * LOG2E));
  }
  if (lane == 0) wpart[cw * 132 + 129] = l;
  cbar_sync();
  float Lsum = 0.f;
#pragma unroll
  for (int w = 0; w < MK_CW; ++w) Lsum += wpart[w * 132 + 129];
  if (S > 1) {
    // exchange (max, sum) among the S co-resident CTAs of this head; merged in split order -> deterministic
    float* pp = p.part + ((size_t)head * p.nsplit + sp) * PART_STRIDE;
    if (threadIdx.x == 0) {
      pp[130] = M; pp[131] = Lsum;
      __threadfence();
      atomicAdd(p.head_tk + head, 1u);
      const unsigned int target = tk_base + (unsigned int)S * (unsigned int)(li + 1);
      const unsigned long long t0 = gtimer();
      unsigned it = 0;
      while ((int)(ld_acquire_gpu(p.head_tk + head) - target) < 0) {
        if ((++it & 255u) == 0 && gtimer() - t0 > MK_TIMEOUT_NS) { atomicExch(p.err, 3u); break; }
      }
    }
    cbar_sync();
    float Mg = -INFINITY;
    for (int s2 = 0; s2 < S; ++s2) Mg = fmaxf(Mg, __ldcg(p.part + ((size_t)head * p.nsplit + s2) * PART_STRIDE + 130));
    float Lg = 0.f;
    for (int s2 = 0; s2 < S; ++s2) {
      const float ms = __ldcg(p.part + ((size_t)head * p.nsplit + s2) * PART_STRIDE + 130);
      if (ms != -INFINITY) Lg = fmaf(__ldcg(p.part + ((size_t)head * p.nsplit + s2) * PART_STRIDE + 131), exp2f((ms - Mg) * LOG2E), Lg);
    }
    M = Mg; Lsum = Lg;
  }
  const float invL = 1.0f / Lsum;
  // ---- pass B
  float o[4] = {0.f, 0.f, 0.f, 0.f}, ob = 0.f;
#pragma unroll 1
  for (int tile = tile0 + cw; tile < tile1; tile += MK_CW) {
    const int t0 = tile << 5, nt = min(32, kv_len - t0);
    const unsigned char* vu = kvq_unit<BITS>(pool, p.block_table[t0 / PAGE], 1, kvh, p.n_kv);
    const bool has_new = tile == last_tile;
    const float sl = scb[t0 + lane];
    const float pl = (lane < nt) ? bf16r(exp2f((sl - M) * LOG2E) * invL) : 0.f;
    for (int j = 0; j < nt; ++j) {
      const float pj = __shfl_sync(0xffffffffu, pl, j);
      float c[4], s1, b1;
      if (has_new && j == jnew) { c[0] = vc[0]; c[1] = vc[1]; c[2] = vc[2]; c[3] = vc[3]; s1 = vsc; b1 = vbi; }
      else { kvq_load_codes<BITS>(vu, (t0 % PAGE) + j, lane, c); kvq_load_sb<BITS>(vu, (t0 % PAGE) + j, lane, s1, b1); }
      const float ws = pj * s1;
      o[0] = fmaf(ws, c[0], o[0]); o[1] = fmaf(ws, c[1], o[1]); o[2] = fmaf(ws, c[2], o[2]); o[3] = fmaf(ws, c[3], o[3]);
      ob = fmaf(pj, b1, ob);
    }
  }
  cbar_sync();                                                 // everyone is done with wpart[..][128/129]
  *reinterpret_cast<float4*>(wpart + cw * 132 + lane * 4) = make_float4(o[0] + ob, o[1] + ob, o[2] + ob, o[3] + ob);
  cbar_sync();
  if (threadIdx.x < HD) {
    const int d = threadIdx.x;
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < MK_CW; ++w) acc += wpart[w * 132 + d];
    if (S == 1) p.attn[head * HD + d] = __float2bfloat16_rn(acc);
    else p.part[((size_t)head * p.nsplit + sp) * PART_STRIDE + d] = acc;
  }
  cbar_sync();
}

// o_proj staging.  Normally a plain copy of the normalised heads: a head that fits one CTA wrote them itself, and with
// split heads the last-arriving split CTA merged them inside the attention phase (attn_last, the default; measured
// 3.5 us -> 1.0 us per layer for this staging at a 2K context).  With attn_last = 0 (kept for A/B runs) every CTA
// merges all heads' splits here, redundantly, from L2 (nact x 132 floats per head, fixed order -> deterministic);
// quantised KV sums its already-normalised partials.
template <int MODE>
__device__ __forceinline__ void mk_stage_attn_merge(bf16* xs, const MkParams& p) {
  const int kv_len = p.st->pos + 1;
  int nact, tps;
  mk_attn_geometry(p, kv_len, nact, tps);
  if (nact == 1 || MODE == MK_ATT_TC || (MODE == MK_ATT_PLAIN && p.attn_last)) {  // the attention CTAs already wrote the normalised heads
    mk_stage_copy(xs, p.attn, p.n_heads * HD);
    return;
  }
  if (MODE == MK_ATT_Q8 || MODE == MK_ATT_Q4) {
    // quantised KV: every split's partial is already normalised by the head's global (max, sum): out = bf16(sum of partials)
    for (int i = threadIdx.x; i < p.n_heads * 32; i += MK_CTHREADS) {
      const int head = i >> 5, l4 = i & 31;
      const float* hp = p.part + ((size_t)head * p.nsplit) * PART_STRIDE + l4 * 4;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int s2 = 0; s2 < nact; ++s2) {          // fixed order -> deterministic
        const float4 v = __ldcg(reinterpret_cast<const float4*>(hp + (size_t)s2 * PART_STRIDE));
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
      }
      __align__(8) bf16 ob[4] = {__float2bfloat16_rn(a0), __float2bfloat16_rn(a1), __float2bfloat16_rn(a2), __float2bfloat16_rn(a3)};
      *reinterpret_cast<uint2*>(xs + head * HD + l4 * 4) = *reinterpret_cast<const uint2*>(ob);
    }
    cbar_sync();
    return;
  }
  // (1) all (m, l) pairs in one parallel round trip -> shared memory (behind the activation vector)
  float* ml = reinterpret_cast<float*>(xs + p.n_heads * HD);            // [n_heads][nact][2]
  for (int i = threadIdx.x; i < p.n_heads * nact; i += MK_CTHREADS) {
    const int head = i / nact, s2 = i % nact;
    const float2 v = __ldcg(reinterpret_cast<const float2*>(p.part + ((size_t)head * p.nsplit + s2) * PART_STRIDE + 128));
    ml[2 * i] = v.x; ml[2 * i + 1] = v.y;
  }
  cbar_sync();
  // (2) each thread merges 2 float4 groups per round and fetches 4 splits of both at once
  //     (8 independent 16-byte loads in flight; more would spill in the GEMV hot loops)
  const int ngrp = p.n_heads * 32;
  for (int g0 = threadIdx.x; g0 < ngrp; g0 += MK_CTHREADS * 2) {
    float acc[2][4], Lsum[2], M[2];
    const float* hp[2];
    const float* hml[2];
    bool ok[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int gidx = g0 + q * MK_CTHREADS;
      ok[q] = gidx < ngrp;
      const int head = ok[q] ? (gidx >> 5) : 0, l4 = gidx & 31;
      hp[q] = p.part + ((size_t)head * p.nsplit) * PART_STRIDE + l4 * 4;
      hml[q] = ml + 2 * head * nact;
      float mm = -INFINITY;
      for (int s2 = 0; s2 < nact; ++s2) mm = fmaxf(mm, hml[q][2 * s2]);
      M[q] = mm; Lsum[q] = 0.f;
      acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.f;
    }
#pragma unroll 1
    for (int s0 = 0; s0 < nact; s0 += 4) {
      float4 ov[2][4];
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (ok[q] && s0 + j < nact) ov[q][j] = __ldcg(reinterpret_cast<const float4*>(hp[q] + (size_t)(s0 + j) * PART_STRIDE));
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (ok[q] && s0 + j < nact) {
            const float ms = hml[q][2 * (s0 + j)];
            if (ms != -INFINITY) {                     // fixed order over splits -> deterministic
              const float w2 = exp2f((ms - M[q]) * LOG2E);
              Lsum[q] = fmaf(hml[q][2 * (s0 + j) + 1], w2, Lsum[q]);
              acc[q][0] = fmaf(ov[q][j].x, w2, acc[q][0]); acc[q][1] = fmaf(ov[q][j].y, w2, acc[q][1]);
              acc[q][2] = fmaf(ov[q][j].z, w2, acc[q][2]); acc[q][3] = fmaf(ov[q][j].w, w2, acc[q][3]);
            }
          }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (!ok[q]) continue;
      const int gidx = g0 + q * MK_CTHREADS;
      const float invL = 1.0f / Lsum[q];
      __align__(8) bf16 ob[4];
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) ob[dd] = __float2bfloat16_rn(acc[q][dd] * invL);
      *reinterpret_cast<uint2*>(xs + (gidx >> 5) * HD + (gidx & 31) * 4) = *reinterpret_cast<const uint2*>(ob);
    }
  }
  cbar_sync();
}

// ---------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------
template <int G, int MODE>
__global__ void __launch_bounds__(MK_THREADS, 1) k_shard_step(const MkParams p) {
  constexpr bool QUANT = MODE == MK_ATT_Q8 || MODE == MK_ATT_Q4;
  extern __shared__ __align__(1024) unsigned char smem[];
  // layout: [ring n_stages*16K][scratch scratch_bytes][full[12]][empty[12]][red floats]
  MkRing ring;
  ring.data = smem;
  ring.n_stages = p.n_stages;
  unsigned char* scratch = smem + (size_t)p.n_stages * MK_STAGE_BYTES;
  ring.full = reinterpret_cast<uint64_t*>(scratch + p.scratch_bytes);
  ring.empty = ring.full + MK_MAX_STAGES;
  float* red = reinterpret_cast<float*>(ring.empty + MK_MAX_STAGES);   // [64] misc scratch
  ring.stage = 0;
  ring.phase = 0;
  ring.starve = reinterpret_cast<volatile unsigned int*>(red + 60);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.n_stages; ++i) {
      mbar_init(&ring.full[i], 1);
      mbar_init(&ring.empty[i], MK_CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  volatile unsigned int* consumed = reinterpret_cast<volatile unsigned int*>(red + 48);
  unsigned int* tmem_slot = reinterpret_cast<unsigned int*>(red + 56);
  unsigned int* bar_req = reinterpret_cast<unsigned int*>(red + 57);
  unsigned int* bar_done = reinterpret_cast<unsigned int*>(red + 58);
  if (threadIdx.x == 0) { *consumed = 0u; *bar_req = 0u; *bar_done = 0u; *tmem_slot = 0u; red[60] = 0.f; red[61] = 0.f; }
  if (p.park && warp == 0) {
    // all 512 TMEM columns: 8 consumer warps x 8 parked stages x 32 columns (one CTA per SM, so nothing else wants them)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp >= MK_CW) {
    // ===== PRODUCERS / PREFETCHER / BARRIER POLLER: never wait for activations =====
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(MK_REGS_PRODUCER));
    if (warp < MK_CW + MK_PW) mk_producer(p, ring, lane, warp - MK_CW);
    else if (warp == MK_CW + MK_PW) mk_prefetcher(p, lane, consumed);
    else if (lane == 0) mk_barrier_poller(p, bar_req, bar_done);
    return;
  }
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(MK_REGS_CONSUMER));

  // ===== CONSUMERS =====
  const int cw = warp;
  bf16* xs = reinterpret_cast<bf16*>(scratch);
  unsigned int bar_k = 0, nblk = 0;
  MkCons cs;
  cs.park = p.park != 0;
  cs.tmem = *reinterpret_cast<volatile unsigned int*>(tmem_slot) + ((uint32_t)(32 * (cw & 3)) << 16) + (uint32_t)((cw >> 2) * 256);
  cs.cons_n = 0; cs.ring_n = 0;
  cs.a_lane = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * MK_ROW_PITCH + (lane >> 4) * 16 + cw * 256);
  float* red2 = red + 192;                                     // [2][MK_CW][MK_ROWS] row-block partial sums
  const unsigned int bar_base = *reinterpret_cast<volatile const unsigned int*>(p.bar_epoch);
  if (threadIdx.x == 0) bar_req[2] = bar_base;                 // red[59]: read by the poller after its first acquire of bar_req
  if (p.hp_x != nullptr) {
    // ===== tensor-parallel head part: (max, sum-exp, argmax) of this shard's vocabulary slice for the due nonce =====
    if (threadIdx.x == 0 && p.hp_wait_flag != nullptr) {
      uint32_t v;
      const unsigned long long t0 = gtimer();
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.hp_wait_flag) : "memory");
        if ((int32_t)(v - p.hp_seq) >= 0) break;
        if (gtimer() - t0 > 5ull * MK_TIMEOUT_NS) { atomicExch(p.err, 4u); break; }
      }
      __threadfence();
    }
    cbar_sync();
    mk_stage_rmsnorm(xs, red, p.hp_x, p.norm_w, p.H, p.eps);
    float hm = -INFINITY, hl = 0.f;
    int hi = 0x7fffffff;
    mk_consume(p, PH_HEAD, 0, ring, xs, red2, cw, lane, consumed, cs, nblk, [&](int, bool) { return 0; },
               [&](int vr, float v, bool owner, int) {
      if (!owner) return;
      const float lg = bf16r(v);
      p.logits_bf16[p.hp_row0 + vr] = __float2bfloat16_rn(v);
      const int gi = p.hp_row0 + vr;
      if (lg > hm) { hl = hl * exp2f((hm - lg) * LOG2E) + 1.0f; hm = lg; hi = gi; }
      else hl += exp2f((lg - hm) * LOG2E);
    });
    auto merge = [](float& m, float& l, int& idx, float om, float ol, int oi) {
      const float nm = fmaxf(m, om);
      const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
      const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
      l = a + b;
      if (om > m || (om == m && oi < idx)) idx = oi;
      m = nm;
    };
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, hm, s);
      const float ol = __shfl_xor_sync(0xffffffffu, hl, s);
      const int oi = __shfl_xor_sync(0xffffffffu, hi, s);
      merge(hm, hl, hi, om, ol, oi);
    }
    int* redi = reinterpret_cast<int*>(red);
    if (lane == 0) { red[cw] = hm; red[8 + cw] = hl; redi[16 + cw] = hi; }
    cbar_sync();
    if (threadIdx.x == 0) {
      float m = red[0], l = red[8]; int idx = redi[16];
      for (int w = 1; w < MK_CW; ++w) merge(m, l, idx, red[w], red[8 + w], redi[16 + w]);
      HeadPartial hp; hp.m = m; hp.l = l; hp.idx = idx; hp.pad = 0;
      p.head_part[blockIdx.x] = hp;
      __threadfence();
      const unsigned int old = atomicAdd(p.head_ticket, 1u);
      redi[32] = (old == gridDim.x - 1) ? 1 : 0;
    }
    cbar_sync();
    if (redi[32] && warp == 0) {
      __threadfence();
      float m = -INFINITY, l = 0.f; int idx = 0x7fffffff;
      for (int i = lane; i < (int)gridDim.x; i += 32)
        merge(m, l, idx, __ldcg(&p.head_part[i].m), __ldcg(&p.head_part[i].l), __ldcg(&p.head_part[i].idx));
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, s);
        const float ol = __shfl_xor_sync(0xffffffffu, l, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        merge(m, l, idx, om, ol, oi);
      }
      if (lane == 0) {
        *p.head_ticket = 0u;
        // this shard's partial -> the head shard's table (peer store over NVLink), then its flag
        volatile float* d = p.hp_dst;
        d[0] = m; d[1] = l; d[2] = __int_as_float(idx); d[3] = 0.f;
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.hp_dst_flag), "r"(p.hp_seq) : "memory");
      }
    }
    cbar_sync();
  }
  if (p.wait_flag != nullptr) {
    // the weights of this step are already streaming into the ring while we wait for the hop
    if (threadIdx.x == 0) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_flag) : "memory");
      if ((int32_t)(v - p.wait_seq) < 0) {
        const unsigned long long t0 = gtimer();
        for (;;) {
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.wait_flag) : "memory");
          if ((int32_t)(v - p.wait_seq) >= 0) break;
          if (gtimer() - t0 > 5ull * MK_TIMEOUT_NS) { atomicExch(p.err, 4u); break; }
        }
      }
      __threadfence();
    }
    cbar_sync();
  }
  const int pos = p.st->pos;
  // per-step constants of the q/k/v epilogue: RoPE (cos, sin) per pair index and the page of `pos`
  float* rope_cs = red + 64;                                   // [64][2]
  if (threadIdx.x < HD / 2) {
    float sn, cs;
    sincosf(__fmul_rn((float)pos, p.inv_freq[threadIdx.x]), &sn, &cs);
    rope_cs[2 * threadIdx.x] = cs; rope_cs[2 * threadIdx.x + 1] = sn;
  }
  const int page_pos = p.block_table[pos / PAGE];
  // attention geometry is fixed for the whole step
  int att_S, att_tps;
  mk_attn_geometry(p, pos + 1, att_S, att_tps);
  const int att_tile0 = ((int)blockIdx.x % att_S) * att_tps + cw;        // this warp's first tile (if the CTA has an attention task)
  const bool att_has = MODE == MK_ATT_PLAIN && (int)blockIdx.x < p.n_heads * att_S && att_tile0 < min((pos + 32) >> 5, ((int)blockIdx.x % att_S + 1) * att_tps);
  // quantised KV: value of this CTA's head counter before any arrival of this launch (arrivals happen after the first grid barrier)
  const unsigned int tk_base = (QUANT && (int)blockIdx.x < p.n_heads * att_S) ? __ldcg(p.head_tk + (int)blockIdx.x / att_S) : 0u;
  const int att_phys0 = att_has ? p.block_table[(att_tile0 << 5) / PAGE] : 0;
  cbar_sync();
  const int tok_in = p.token_in != nullptr ? __ldcg(p.token_in) : p.st->token;
  const bf16* cur = p.embed != nullptr ? p.embed + (size_t)min(max(tok_in, 0), p.vocab - 1) * p.H : p.x_in;

#define MK_STAMP(i) do { if (p.dbg != nullptr && threadIdx.x == 0) p.dbg[((size_t)blockIdx.x * p.n_layers + li) * MK_DBG_WORDS + (i)] = gtimer(); } while (0)
  for (int li = 0; li < p.n_layers; ++li) {
    const MkLayer L = p.layers[li];
    bf16* nxt = (li == p.n_layers - 1) ? p.x_out : ((li & 1) ? p.xb : p.xa);
    MK_STAMP(0);
    if (att_has && lane == 0) {
      // pull this warp's K and V tile (8 KiB each, contiguous inside the page) toward L2 now; the
      // loads after the next grid barrier then hit L2 instead of HBM
      const int kvh0 = ((int)blockIdx.x / att_S) / G;
      const int nt0 = min(32, pos + 1 - (att_tile0 << 5));
      const bf16* kb = L.kv_pool + (((size_t)att_phys0 * 2) * p.n_kv + kvh0) * (PAGE * HD) + (size_t)((att_tile0 << 5) % PAGE) * HD;
      const uint64_t polk = l2_policy_evict_last();
      tma_prefetch_l2(kb, (uint32_t)nt0 * HD * 2u, polk);
      tma_prefetch_l2(kb + (size_t)p.n_kv * (PAGE * HD), (uint32_t)nt0 * HD * 2u, polk);
    }

    // ---- P1: RMSNorm -> q/k/v -> RoPE -> paged-KV append
    mk_stage_rmsnorm(xs, red, cur, L.w[MK_W_LN1], p.H, p.eps);
    MK_STAMP(1);
    mk_consume(p, PH_QKV, li, ring, xs, red2, cw, lane, consumed, cs, nblk, [&](int, bool) { return 0; },
               [&](int vr, float v, bool owner, int) {
      const int task = vr >> 1, which = vr & 1;
      const int slot = task >> 6, d = task & 63;
      int kind = 0, hrow = slot;
      if (slot >= p.n_heads + p.n_kv) { kind = 2; hrow = slot - p.n_heads - p.n_kv; }
      else if (slot >= p.n_heads) { kind = 1; hrow = slot - p.n_heads; }
      const int dim = d + (which << 6);
      const bf16* bias = L.w[kind == 0 ? MK_W_QB : (kind == 1 ? MK_W_KB : MK_W_VB)];
      if (owner && bias != nullptr) v += __bfloat162float(bias[hrow * HD + dim]);
      const float y = bf16r(v);
      const float yp = __shfl_xor_sync(0xffffffffu, y, 1);
      if (!owner) return;
      float o = y;
      if (kind != 2) {
        const float cs = rope_cs[2 * d], sn = rope_cs[2 * d + 1];     // per-step table (same sincosf(pos * inv_freq[d]))
        o = which == 0 ? __fsub_rn(__fmul_rn(y, cs), __fmul_rn(yp, sn)) : __fadd_rn(__fmul_rn(yp, sn), __fmul_rn(y, cs));
        o = bf16r(o);
      }
      if (kind == 0) {
        p.qbuf[hrow * HD + dim] = __float2bfloat16_rn(o);
      } else {
        // quantised KV: the bf16 row goes to the staging pool; the attention phase quantises and appends it
        bf16* wpool = QUANT ? p.kv_stage : L.kv_pool;
        const int wpage = QUANT ? (pos / PAGE) % KVQ_STAGE_PAGES : page_pos;
        const size_t off = (((size_t)wpage * 2 + (kind - 1)) * p.n_kv + hrow) * (PAGE * HD) + (size_t)(pos % PAGE) * HD + dim;
        wpool[off] = __float2bfloat16_rn(o);
      }
    });
    MK_STAMP(2);
    mk_grid_barrier(p, bar_k, ring, cs, bar_req, bar_done, lane);
    MK_STAMP(3);

    // ---- P2: paged-KV attention (split over pages; splits are merged while staging P3)
    if constexpr (MODE == MK_ATT_Q8) mk_attention_q<G, 8>(p, L, scratch, cw, lane, att_S, att_tps, li, tk_base);
    else if constexpr (MODE == MK_ATT_Q4) mk_attention_q<G, 4>(p, L, scratch, cw, lane, att_S, att_tps, li, tk_base);
    else if constexpr (MODE == MK_ATT_TC) mk_attention_tc<G>(p, L, scratch, cw, lane, att_S, att_tps, ring, cs, reinterpret_cast<unsigned int*>(red + 61));
    else mk_attention<G>(p, L, scratch, cw, lane, att_S, att_tps, att_tile0, att_phys0);
    MK_STAMP(4);
    mk_grid_barrier(p, bar_k, ring, cs, bar_req, bar_done, lane);
    MK_STAMP(5);

    // ---- P3: merge attention splits -> o_proj + residual
    mk_stage_attn_merge<MODE>(xs, p);
    MK_STAMP(6);
    mk_consume(p, PH_O, li, ring, xs, red2, cw, lane, consumed, cs, nblk,
               [&](int vr, bool owner) -> unsigned short { return (MK_OPT_PRE && owner) ? __ldcg(reinterpret_cast<const unsigned short*>(cur) + vr) : (unsigned short)0; },
               [&](int vr, float v, bool owner, unsigned short xb_) {
      if (!owner) return;
      if (!MK_OPT_PRE) xb_ = __ldcg(reinterpret_cast<const unsigned short*>(cur) + vr);
      const float o = bf16r(v);
      p.hbuf[vr] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(__ushort_as_bfloat16(xb_)), o));
    });
    MK_STAMP(7);
    mk_grid_barrier(p, bar_k, ring, cs, bar_req, bar_done, lane);
    MK_STAMP(8);

    // ---- P4: RMSNorm -> gate/up -> SwiGLU
    mk_stage_rmsnorm(xs, red, p.hbuf, L.w[MK_W_LN2], p.H, p.eps);
    MK_STAMP(9);
    mk_consume(p, PH_GU, li, ring, xs, red2, cw, lane, consumed, cs, nblk, [&](int, bool) { return 0; },
               [&](int vr, float v, bool owner, int) {
      const float y = bf16r(v);
      const float u = __shfl_xor_sync(0xffffffffu, y, 1);
      if (!owner || (vr & 1)) return;
      const float s = bf16r(1.0f / (1.0f + expf(-y)));
      const float a = bf16r(__fmul_rn(y, s));
      p.act[vr >> 1] = __float2bfloat16_rn(__fmul_rn(a, u));
    });
    MK_STAMP(10);
    mk_grid_barrier(p, bar_k, ring, cs, bar_req, bar_done, lane);
    MK_STAMP(11);

    // ---- P5: down_proj + residual (+ cast to wire dtype == bf16 store)
    mk_stage_copy(xs, p.act, p.FFN);
    MK_STAMP(12);
    mk_consume(p, PH_DOWN, li, ring, xs, red2, cw, lane, consumed, cs, nblk,
               [&](int vr, bool owner) -> unsigned short { return (MK_OPT_PRE && owner) ? __ldcg(reinterpret_cast<const unsigned short*>(p.hbuf) + vr) : (unsigned short)0; },
               [&](int vr, float v, bool owner, unsigned short hb) {
      if (!owner) return;
      if (!MK_OPT_PRE) hb = __ldcg(reinterpret_cast<const unsigned short*>(p.hbuf) + vr);
      const float o = bf16r(v);
      nxt[vr] = __float2bfloat16_rn(__fadd_rn(__bfloat162float(__ushort_as_bfloat16(hb)), o));
    });
    MK_STAMP(13);
    mk_grid_barrier(p, bar_k, ring, cs, bar_req, bar_done, lane);
    MK_STAMP(14);
    cur = nxt;
  }

  if (p.do_head) {
    // ---- final RMSNorm -> lm_head -> bf16 logits -> greedy sample
    mk_stage_rmsnorm(xs, red, cur, p.norm_w, p.H, p.eps);
    float hm = -INFINITY, hl = 0.f;
    int hi = 0x7fffffff;
    mk_consume(p, PH_HEAD, 0, ring, xs, red2, cw, lane, consumed, cs, nblk, [&](int, bool) { return 0; },
               [&](int vr, float v, bool owner, int) {
      if (!owner) return;
      const float lg = bf16r(v);
      p.logits_bf16[vr] = __float2bfloat16_rn(v);
      if (p.logits_f32 != nullptr) p.logits_f32[vr] = v;
      if (lg > hm) { hl = hl * exp2f((hm - lg) * LOG2E) + 1.0f; hm = lg; hi = vr; }
      else hl += exp2f((lg - hm) * LOG2E);
    });
    // merge (m, l, idx) over the warp, then over the 8 consumer warps, then over CTAs
    auto merge = [](float& m, float& l, int& idx, float om, float ol, int oi) {
      const float nm = fmaxf(m, om);
      const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
      const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
      l = a + b;
      if (om > m || (om == m && oi < idx)) idx = oi;
      m = nm;
    };
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, hm, s);
      const float ol = __shfl_xor_sync(0xffffffffu, hl, s);
      const int oi = __shfl_xor_sync(0xffffffffu, hi, s);
      merge(hm, hl, hi, om, ol, oi);
    }
    int* redi = reinterpret_cast<int*>(red);
    if (lane == 0) { red[cw] = hm; red[8 + cw] = hl; redi[16 + cw] = hi; }
    cbar_sync();
    if (threadIdx.x == 0) {
      float m = red[0], l = red[8]; int idx = redi[16];
      for (int w = 1; w < MK_CW; ++w) merge(m, l, idx, red[w], red[8 + w], redi[16 + w]);
      HeadPartial hp; hp.m = m; hp.l = l; hp.idx = idx; hp.pad = 0;
      p.head_part[blockIdx.x] = hp;
      __threadfence();
      const unsigned int old = atomicAdd(p.head_ticket, 1u);
      redi[32] = (old == gridDim.x - 1) ? 1 : 0;
    }
    cbar_sync();
    if (redi[32] && warp == 0) {
      __threadfence();
      float m = -INFINITY, l = 0.f; int idx = 0x7fffffff;
      for (int i = lane; i < (int)gridDim.x; i += 32)
        merge(m, l, idx, __ldcg(&p.head_part[i].m), __ldcg(&p.head_part[i].l), __ldcg(&p.head_part[i].idx));
#pragma unroll
      for (int s = 16; s >= 1; s >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, m, s);
        const float ol = __shfl_xor_sync(0xffffffffu, l, s);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, s);
        merge(m, l, idx, om, ol, oi);
      }
      if (lane == 0) {
        const float lse = bf16r(m + logf(l));
        const float lp = bf16r(__fsub_rn(m, lse));
        // host-visible result (token_out / logprob_out may be pinned host memory polled WITHOUT a stream sync,
        // shard/token_tap.py): logprob first, system fence, then the token.  A timed-out bounded wait makes the
        // step's results invalid: the host then sees -(1000 + code) instead of a token id.
        const unsigned int ecode = *reinterpret_cast<volatile unsigned int*>(p.err);
        if (p.logprob_out != nullptr) *reinterpret_cast<volatile float*>(p.logprob_out) = lp;
        p.st->token = idx;
        *p.head_ticket = 0u;
        if (p.send_dst != nullptr) *reinterpret_cast<volatile int32_t*>(p.send_dst) = idx;   // token hop to shard 0
        __threadfence_system();
        if (p.token_out != nullptr) *reinterpret_cast<volatile int32_t*>(p.token_out) = ecode ? -(1000 + (int)ecode) : idx;
        __threadfence_system();
        if (p.send_flag != nullptr)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.send_flag), "r"(p.send_seq) : "memory");
      }
    }
  }
  if (!p.do_head && p.send_dst != nullptr && blockIdx.x == 0) {
    // activation hop: every CTA passed the barrier after the last down-proj, so x_out is complete;
    // CTA 0 stores it into the successor's slot over NVLink and releases the sequence flag
    const uint4* src4 = reinterpret_cast<const uint4*>(p.x_out);
    uint4* dst4 = reinterpret_cast<uint4*>(p.send_dst);
    for (int i = threadIdx.x; i < p.H / 8; i += MK_CTHREADS) dst4[i] = __ldcg(src4 + i);
    __threadfence_system();
    cbar_sync();
    if (threadIdx.x == 0)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.send_flag), "r"(p.send_seq) : "memory");
  }
  if (p.bc_n > 0 && p.n_layers > 0 && blockIdx.x == 0) {
    // last shard, tensor-parallel head: broadcast the final hidden state to every shard's slot (itself included)
    const uint4* src4 = reinterpret_cast<const uint4*>(p.x_out);
    for (int d = 0; d < p.bc_n; ++d) {
      uint4* dst4 = reinterpret_cast<uint4*>(p.bc_dst[d]);
      for (int i = threadIdx.x; i < p.H / 8; i += MK_CTHREADS) dst4[i] = __ldcg(src4 + i);
    }
    __threadfence_system();
    cbar_sync();
    if (threadIdx.x < p.bc_n)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.bc_flag[threadIdx.x]), "r"(p.bc_seq) : "memory");
  }
  if (p.mg_n > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    // head shard: every shard's partial of the due nonce has landed (they ran their head parts while this launch
    // computed its layers); merge in shard order -> token, logprob; hand the token to the next step (own lane slot)
    float m = -INFINITY, l = 0.f; int idx = 0x7fffffff;
    for (int i = 0; i < p.mg_n; ++i) {
      uint32_t v;
      const unsigned long long t0 = gtimer();
      for (;;) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.mg_flags + 16 * i) : "memory");
        if ((int32_t)(v - p.mg_seq) >= 0) break;
        if (gtimer() - t0 > 5ull * MK_TIMEOUT_NS) { atomicExch(p.err, 4u); break; }
      }
      const float om = __ldcg(p.mg_part + 4 * i), ol = __ldcg(p.mg_part + 4 * i + 1);
      const int oi = __float_as_int(__ldcg(p.mg_part + 4 * i + 2));
      const float nm = fmaxf(m, om);
      const float a = (m == -INFINITY) ? 0.f : l * exp2f((m - nm) * LOG2E);
      const float b = (om == -INFINITY) ? 0.f : ol * exp2f((om - nm) * LOG2E);
      l = a + b;
      if (om > m || (om == m && oi < idx)) idx = oi;
      m = nm;
    }
    const float lse = bf16r(m + logf(l));
    const float lp = bf16r(__fsub_rn(m, lse));
    const unsigned int ecode = *reinterpret_cast<volatile unsigned int*>(p.err);
    if (p.mg_st != nullptr) p.mg_st->token = idx;
    if (p.mg_logprob_out != nullptr) *reinterpret_cast<volatile float*>(p.mg_logprob_out) = lp;
    if (p.mg_slot != nullptr) *reinterpret_cast<volatile int32_t*>(p.mg_slot) = idx;
    __threadfence_system();
    if (p.mg_token_out != nullptr) *reinterpret_cast<volatile int32_t*>(p.mg_token_out) = ecode ? -(1000 + (int)ecode) : idx;
    if (p.mg_slot_flag != nullptr)
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.mg_slot_flag), "r"(p.mg_slot_seq) : "memory");
    __threadfence_system();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    if (p.advance) p.st->pos = pos + 1;
    *p.bar_epoch = bar_base + bar_k * gridDim.x;   // every CTA passed bar_k barriers; next launch starts here
  }
  if (threadIdx.x == 0) st_release_cta(bar_req, 0xffffffffu);          // poller warp: exit
  if (p.park) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cbar_sync();                                                         // every warp is done with its TMEM region
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
      const unsigned int tbase = *reinterpret_cast<volatile unsigned int*>(tmem_slot);
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
    }
  }
}

}  // namespace dn
