/*
 * dnet_b200.h -- C ABI of libdnet_b200.so: the B200-native shard forward of dnet's
 * pipelined ring (per-layer decode forward, end-shard sampling, ring hop, layer swap).
 *
 * The reference (firstbatchxyz/dnet @ e76f54a) has no C/FFI seam on this path: its
 * seams are Python plug-in points (ComputePolicy registry, BaseRingModel operator
 * API).  Every entry point below is what a binding for those seams would call; the
 * reference interface each one replaces is cited as file:line relative to
 * /root/reference/.  INTEGRATION.md shows the reference-side ctypes stub.
 *
 * Conventions
 *   - plain C: opaque handles, raw device pointers, sizes; no torch / C++ types.
 *   - every function returns 0 on success or a negative DN_E* code and never throws;
 *     dn_last_error() returns a thread-local human-readable message.
 *   - the caller owns every buffer and stream it passes in.  Functions taking a
 *     stream are asynchronous with respect to it and contain no hidden host sync.
 *   - one compute thread per shard, as in the reference (shard/runtime.py:364-372);
 *     handles are thread-compatible, not thread-safe.
 *   - activations are row-major [T, hidden] bf16 (the reference's (1,T,H) array with
 *     the unit batch dim dropped); weights are HF/MLX layout [out, in] row-major bf16,
 *     y = x W^T.
 */
#ifndef DNET_B200_H
#define DNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DN_OK 0
#define DN_EINVAL (-22)   /* bad argument / unsupported shape */
#define DN_ENOMEM (-12)   /* device / pinned allocation failed */
#define DN_ENOENT (-2)    /* layer not hosted / not bound */
#define DN_ECUDA (-5)     /* a CUDA runtime call failed; see dn_last_error() */
#define DN_ETIME (-62)    /* hop wait timed out */
#define DN_ENOSPC (-28)   /* KV capacity exceeded */

#define DN_DTYPE_BF16 0

/* index of each tensor in the dev_ptrs array handed to dn_bind_layer
 * (reference utils/model.py:27-43 key regexes) */
enum {
  DN_W_Q = 0, DN_W_K, DN_W_V, DN_W_O, DN_W_GATE, DN_W_UP, DN_W_DOWN,
  DN_W_LN1, DN_W_LN2,
  DN_W_QB, DN_W_KB, DN_W_VB,   /* optional q/k/v bias (qwen2); NULL for llama */
  DN_W_COUNT
};

typedef struct dn_model dn_model;   /* one shard's model slice + scratch */
typedef struct dn_kv dn_kv;         /* one nonce's paged KV + step state  */
typedef struct dn_graph dn_graph;   /* an instantiated CUDA graph         */
typedef void* dn_stream;            /* cudaStream_t                       */
typedef void* dn_event;             /* cudaEvent_t                        */

/* mlx_lm.models.llama.ModelArgs as consumed by reference core/models/llama.py:33 */
typedef struct dn_model_cfg {
  int32_t hidden;          /* hidden_size */
  int32_t n_heads;         /* num_attention_heads */
  int32_t n_kv_heads;      /* num_key_value_heads */
  int32_t head_dim;        /* must be 128 */
  int32_t ffn;             /* intermediate_size */
  int32_t vocab;           /* vocab_size */
  int32_t n_layers_total;  /* num_hidden_layers of the whole model */
  float rms_eps;           /* rms_norm_eps */
  int32_t tie_embeddings;  /* lm_head = embed_tokens (core/models/llama.py:62-66) */
  int32_t dtype;           /* DN_DTYPE_BF16 */
  int32_t wire_dtype;      /* must equal dtype (mixed wire dtype) */
  int32_t kv_page_tokens;  /* 64 */
  int32_t kv_pool_pages;   /* pages in the shard's KV pool (shared by all nonces) */
  int32_t kv_bits;         /* 0: 16-bit KV (mlx_lm KVCache); 4 / 8: affine-quantised KV (QuantizedKVCache,
                              reference utils/model.py:505-554 -- the API's default kv_bits) */
  int32_t kv_group;        /* quantisation group along head_dim: 64 */
  int32_t n_experts;       /* 0: dense MLP; > 0: sparse MoE FFN (mixtral num_local_experts; BASELINE configs[4]) */
  int32_t top_k;           /* experts per token (num_experts_per_tok) */
} dn_model_cfg;

/* ---- process / device -------------------------------------------------- */
int dn_init(int device);                         /* cudaSetDevice + capability check (sm_100) */
const char* dn_last_error(void);
const char* dn_version(void);
int dn_set_option(const char* key, int64_t value);
/* process-wide tuning / experiment switches; results never depend on them:
 *   "pdl" 0/1 (per-op path: programmatic dependent launch)   "l2_prefetch_kb" (per-op path)
 *   "tc_prefill" 0/1  tensor-core prefill chunks of 16..512 tokens (default 1)
 *   "tc_attn" 0/1     prefill attention on tcgen05 instead of CUDA cores (default 1)
 *   step kernel: "park" 0/1 TMEM parking during grid barriers (1), "inflight" ring stages with loads
 *   outstanding (2), "inflight_hi" cap while the consumers starve for weights (3), "attn_chunk" tokens per
 *   warp before a head is split over a second CTA (32), "pf_depth" L2 look-ahead stages (0),
 *   "mk_debug" 0/1 phase stamps for dn_step_debug, "mk_flags" timing experiments (bit2 skip math,
 *   bit3 skip grid barriers: garbage results) */
int64_t dn_launch_count(void);                   /* kernels launched by this library so far
                                                    (graph replays count their nodes) */
int dn_device_sm_count(void);

/* ---- model slice: replaces BaseRingModel ctor + load_weights / unload_layers
 *      (reference core/models/llama.py:20-54, core/models/base.py:111-195,474-486).
 *      The model BORROWS weight memory (owned by the caller's WeightCache). */
int dn_model_create(const dn_model_cfg* cfg, const int32_t* abs_layers, int n_layers,
                    const float* inv_freq_host /* head_dim/2 fp32 */, dn_model** out);
int dn_model_destroy(dn_model* m);
int dn_bind_layer(dn_model* m, int abs_layer, const void* const* dev_ptrs /* DN_W_COUNT */);
/* sparse MoE layers (cfg.n_experts > 0; mlx_lm.models.mixtral via the same BaseRingModel seam the reference uses for
 * its MoE families, core/models/gpt_oss.py, deepseek_v2.py): router [E][H] and per-expert gate / up / down matrices.
 * dn_bind_layer still binds the attention tensors and norms (its gate/up/down entries may be NULL for such a model). */
int dn_bind_layer_experts(dn_model* m, int abs_layer, const void* router, const void* const* gate /* [E] */,
                          const void* const* up /* [E] */, const void* const* down /* [E] */, int n_experts);
int dn_unbind_layer(dn_model* m, int abs_layer);
int dn_layer_is_bound(dn_model* m, int abs_layer);
/* embed_tokens / final norm / lm_head (reference shard/runtime.py:263-273); any may be NULL */
int dn_bind_api(dn_model* m, const void* embed, const void* norm, const void* head);
int dn_model_max_chunk(dn_model* m);             /* largest small chunk (1,2,4) accepted by dn_layer_forward */
int dn_model_max_prefill_chunk(dn_model* m);     /* 16..this many tokens run on the tcgen05/TMA prefill GEMMs (0: none) */

/* ---- per-nonce KV: replaces make_cache + mlx_lm KVCache
 *      (reference utils/model.py:470-555, shard/runtime.py:374-396) */
int dn_kv_create(dn_model* m, int max_tokens, dn_kv** out);
int dn_kv_free(dn_kv* kv);
int dn_kv_reset(dn_kv* kv, dn_stream s);          /* offset <- 0 */
int dn_kv_offset(dn_kv* kv);                      /* host mirror of cache.offset */
int dn_kv_advance(dn_kv* kv, int T, dn_stream s); /* offset += T (device + host mirror) */
int dn_kv_seek(dn_kv* kv, int pos, dn_stream s);  /* offset <- pos (chunked prefill across windows) */
int dn_kv_set_token(dn_kv* kv, int32_t token, dn_stream s); /* device step-state token */
int dn_kv_note_advance(dn_kv* kv, int T);          /* host mirror only: a graph replay advanced the offset */
void* dn_kv_token_ptr(dn_kv* kv);                 /* device int32*: step-state token */

/* ---- operators (BaseRingModel API, reference core/models/base.py:20-73) ---- */
/* embed: core/models/llama.py:56-57.  ids: device int32[T]. x_out: [T,hidden] bf16 */
int dn_embed(dn_model* m, const int32_t* ids_dev, int T, void* x_out, dn_stream s);
/* apply_single_layer + the policy's cast to wire dtype
 * (core/models/llama.py:76-102, shard/policies/fit_in_memory.py:102-109).
 * x_inout [T,hidden] bf16, updated in place.  KV positions are kv.offset..+T-1;
 * the caller advances the offset once per message with dn_kv_advance. */
int dn_layer_forward(dn_model* m, int abs_layer, void* x_inout, int T, dn_kv* kv, dn_stream s);
/* the window loop of FitInMemoryPolicy.process (fit_in_memory.py:78-124) in one call */
int dn_window_forward(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, int T,
                      dn_kv* kv, dn_stream s);
/* The whole single-token step of this shard -- [embed] + a contiguous run of local layers +
 * [final norm + lm_head + greedy sample] + [offset += 1] -- as ONE persistent cooperative
 * kernel (TMA-fed weight ring, grid barriers between phases; dn_megakernel.cuh).  Same
 * semantics as dn_embed + dn_window_forward + dn_head_sample_greedy + dn_kv_advance with T=1
 * (FitInMemoryPolicy.process for a decode message, fit_in_memory.py:34-209).  With
 * embed_from_token the input is embed_tokens[step-state token] and x_inout is output only. */
int dn_shard_step(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                  int embed_from_token, int do_head, int32_t* token_out, float* logprob_out,
                  float* logits_f32_out, int advance, dn_stream s);
/* dn_shard_step with the ring hop fused into the kernel (compute + peer stores over NVLink in one
 * launch): waits for wait_flag >= wait_seq before reading its input, publishes its result into the
 * successor's slot + flag at the end.  Replaces RingAdapter._send_activation / ingress for the
 * tensor bytes (shard/adapters/ring.py:161-206,265-299) without any separate hop kernel. */
int dn_shard_step_hop(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                      int embed_from_token, int do_head, int32_t* token_out, float* logprob_out, int advance,
                      const uint32_t* wait_flag, uint32_t wait_seq, const int32_t* token_in,
                      void* send_dst, uint32_t* send_flag, uint32_t send_seq, dn_stream s);
/* ---- tensor-parallel lm_head over the ring: every shard holds vocab/S rows of the head (dn_bind_head_slice); the
 *      last shard broadcasts a token's final hidden state to every shard over NVLink, each computes (max, sum-exp,
 *      argmax) of its slice and stores it into the head shard's table, which merges the S partials into the token.
 *      Replaces normalize + lm_project + Sampler.sample(temperature 0) on the end shard (fit_in_memory.py:134-157)
 *      for the on-device decode schedule; all pointers except hp_x / mg_part / mg_flags may be peer (IPC) memory. */
typedef struct dn_tp_args {
  const void* hp_x;                 /* head part of THIS launch: [hidden] bf16 final hidden state, local; NULL = none */
  const uint32_t* hp_wait_flag;     /* its arrival flag (>= hp_seq) */
  uint32_t hp_seq;
  void* hp_dst;                     /* 4 floats (m, l, idx bits, pad): this shard's entry of the head shard's table */
  uint32_t* hp_dst_flag;            /* released with hp_seq */
  int32_t bc_n;                     /* last shard: number of destinations of the final hidden state (ring size) */
  void* bc_dst[16];
  uint32_t* bc_flag[16];
  uint32_t bc_seq;
  int32_t mg_n;                     /* head shard: number of partials to merge at the end of the launch (ring size) */
  const void* mg_part;              /* [16][4] floats, local */
  const uint32_t* mg_flags;         /* one flag per 64 bytes, local */
  uint32_t mg_seq;
  dn_kv* mg_kv;                     /* the due nonce's step state receives the token */
  int32_t* mg_token_out;            /* host-visible (pinned) token / logprob, as dn_shard_step */
  float* mg_logprob_out;
  void* mg_slot;                    /* own lane slot: token for the nonce's next step (token_in) ... */
  uint32_t* mg_slot_flag;           /* ... and its flag, released with mg_slot_seq */
  uint32_t mg_slot_seq;
} dn_tp_args;
int dn_bind_head_slice(dn_model* m, const void* slice /* [nrows, hidden] bf16 */, int row0, int nrows);
int dn_shard_step_tp(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv, int embed_from_token,
                     int advance, const uint32_t* wait_flag, uint32_t wait_seq, const int32_t* token_in,
                     void* send_dst, uint32_t* send_flag, uint32_t send_seq, const dn_tp_args* tp, dn_stream s);
int dn_step_error(dn_model* m, dn_stream s);
/* 0, or the code of a timed-out bounded wait inside k_shard_step (2 ring, 3 grid barrier, 4 hop flag).  The word is
 * sticky and the kernel reports it to the host as token_out = -(1000 + code); dn_step_error_clear resets it
 * (asynchronous on s) once the caller has failed the request. */
int dn_step_error_clear(dn_model* m, dn_stream s);
/* parity bisection hook: copy a scratch buffer of the per-op path's LAST chunk to the host (0 q after RoPE,
 * 1 attention output, 2 h = x + o_proj, 3 SwiGLU output); synchronises s */
int dn_debug_scratch(dn_model* m, int which, void* host_out, size_t bytes, dn_stream s);
/* per-SM row partition of the step kernel's four weight phases, [4][sms+1] (NULL: equal split); see
 * dnet_b200.shard.calibrate */
int dn_step_set_bounds(dn_model* m, const int32_t* bounds_host);
/* measurement hook: per-phase globaltimer stamps [sm][layer][16] of the last step (option mk_debug=1) */
int dn_step_debug(dn_model* m, unsigned long long* out_host, size_t max_words, dn_stream s);      /* 0, or the code of a timed-out in-kernel wait */
/* measurement hooks for bench.py: per-kernel device times (CUDA events on stream s between
 * the five launches of one layer: qkv+rope+append, attention, o_proj, gate/up, down) */
int dn_layer_forward_timed(dn_model* m, int abs_layer, void* x_inout, int T, dn_kv* kv, dn_stream s,
                           float ms_out[5]);
int dn_head_timed(dn_model* m, const void* x, int T, dn_stream s, float* ms_out);
/* normalize + lm_project on the LAST position + Sampler.sample with temperature==0
 * (fit_in_memory.py:134-157, core/decoding/sampler.py:15-65).  token_out / logprob_out:
 * device-accessible (device or mapped pinned) int32 / float.  If kv != NULL the token
 * is also written to the nonce's step state (feeds the next dn_embed). */
int dn_head_sample_greedy(dn_model* m, const void* x, int T, dn_kv* kv, int32_t* token_out,
                          float* logprob_out, dn_stream s);
/* fp32 last-position logits BEFORE the bf16 rounding (parity / stochastic sampling) and
 * the bf16-rounded logits the reference's Linear returns; either pointer may be NULL */
int dn_head_logits(dn_model* m, const void* x, int T, float* logits_f32_out,
                   void* logits_bf16_out, dn_stream s);

/* ---- CUDA graph capture of a window / step (replaces the per-layer Python loop +
 *      mx.eval per window, fit_in_memory.py:102-113) */
int dn_graph_begin(dn_stream s);
int dn_graph_end(dn_stream s, dn_graph** out);
int dn_graph_launch(dn_graph* g, dn_stream s);
int dn_graph_destroy(dn_graph* g);
int dn_graph_num_nodes(dn_graph* g);

/* ---- ring hop data plane: replaces RingAdapter._send_activation + gRPC
 *      StreamActivations for the tensor bytes (reference shard/adapters/ring.py:265-299,
 *      shard/grpc_servicer/servicer.py:129-161).  Slots live on the RECEIVING GPU. */
int dn_hop_alloc(size_t bytes, void** dev_ptr);             /* cudaMalloc, IPC-exportable, zeroed */
int dn_hop_free(void* dev_ptr);
int dn_hop_export(void* dev_ptr, uint8_t handle_out[64]);   /* cudaIpcGetMemHandle */
int dn_hop_import(const uint8_t handle[64], void** dev_ptr);/* cudaIpcOpenMemHandle */
int dn_hop_close(void* imported_ptr);
int dn_enable_peer(int peer_device);                        /* same-process multi-GPU */
/* copy bytes into the peer slot on stream s, then publish seq to *dst_flag (system scope) */
int dn_hop_send(void* dst_slot, const void* src, size_t bytes, uint32_t* dst_flag, uint32_t seq,
                dn_stream s);
/* make stream s wait until *flag >= seq (device-side spin, bounded by timeout_ms; on timeout
 * the kernel sets *err_flag (device uint32, may be NULL) and returns so the GPU never hangs) */
int dn_hop_wait(const uint32_t* flag, uint32_t seq, uint32_t timeout_ms, uint32_t* err_flag,
                dn_stream s);

/* measurement hook for the ring-hop latency BASELINE.json's metric names: a (bytes + flag) token makes `iters` trips
 * around the ring on a lane no request uses, every rank running this on its stream; the origin's *out_ns
 * (device-accessible) receives the elapsed ns of all trips on ITS clock: one hop = out_ns / (iters * ring size). */
int dn_hop_ring_probe(const void* own_slot, const uint32_t* own_flag, void* next_slot, uint32_t* next_flag,
                      size_t bytes, uint32_t base_seq, int iters, int is_origin, uint32_t timeout_ms,
                      unsigned long long* out_ns, dn_stream s);

/* ---- layer swap: replaces LayerManager.load_layer_to_gpu / WeightCache materialise
 *      (reference utils/layer_manager.py:229-282, core/memory/weight_cache.py:68-196) */
int dn_pinned_alloc(size_t bytes, void** host_ptr);         /* cudaHostAlloc (portable) */
int dn_pinned_free(void* host_ptr);
int dn_device_alloc(size_t bytes, void** dev_ptr);
int dn_device_free(void* dev_ptr);
/* stage one layer slot from pinned host memory on the prefetch stream; record done */
int dn_slot_prefetch(void* dst_dev, const void* src_pinned, size_t bytes, dn_stream prefetch,
                     dn_event done);
int dn_stream_create(dn_stream* out, int high_priority);
int dn_stream_destroy(dn_stream s);
int dn_stream_sync(dn_stream s);
int dn_stream_wait_event(dn_stream s, dn_event e);
int dn_event_create(dn_event* out, int timing);
int dn_event_destroy(dn_event e);
int dn_event_record(dn_event e, dn_stream s);
int dn_event_query(dn_event e);                              /* 1 done, 0 pending, <0 error */
int dn_event_sync(dn_event e);
int dn_event_elapsed_ms(dn_event a, dn_event b, float* ms);
int dn_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, dn_stream s);
int dn_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, dn_stream s);

#ifdef __cplusplus
}
#endif
#endif /* DNET_B200_H */
