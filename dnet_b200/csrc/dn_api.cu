// dn_api.cu -- C-ABI of libdnet_b200.so (see include/dnet_b200.h).
// Host-side bookkeeping only: which layer pointers are bound, the per-nonce paged KV,
// scratch buffers, launch geometry, CUDA graphs, the NVLink hop and the pinned->HBM
// layer-swap copies.  All arithmetic is in dn_kernels.cuh.
#include "dn_kernels.cuh"
#include "dn_kvquant.cuh"
#include "dn_megakernel.cuh"
#include "dn_gemm_tc.cuh"
#include "../../include/dnet_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_map>
#include <vector>

using namespace dn;

// ---------------------------------------------------------------------------------
// errors / options
// ---------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(...)                                                                           \
  do {                                                                                    \
    cudaError_t e_ = (__VA_ARGS__);                                                             \
    if (e_ != cudaSuccess)                                                                \
      return fail(DN_ECUDA, "%s failed: %s (%s:%d)", #__VA_ARGS__, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

static std::atomic<long long> g_launches{0};
static int g_pdl = 0;
static int g_l2_prefetch_kb = 64;
static int g_mk_flags = 0;
static int g_inflight_hi = 3;   // step kernel: cap while the consumers are starving for weights
static int g_park = 1;          // step kernel: park ready ring stages in tensor memory during grid barriers
static int g_inflight = 2;      // step kernel: ring stages with loads outstanding while the consumers are not starving (barriers, staging); measured caps 2/3/4/5/none = 357/381/374/369/366 tok/s static, 2-when-idle/3-when-starving +0.7 % on top
static int g_attn_chunk = 32;   // step kernel: tokens per warp before a head is split over a second CTA
static int g_attn_single = 512; // step kernel: up to this context one CTA per head runs several tile passes instead of splitting the head
static int g_attn_last = 1;     // step kernel, plain attention with split heads: last-arriving CTA merges (0 = every CTA merges while staging)
static int g_attn_tc = 1;       // step kernel: tensor-core attention phase (shared-memory K/V tiles + mma) for long contexts ...
static int g_attn_tc_min = 12288;   // ... from this many tokens of context on (bf16 KV)
static int g_mk_debug = 0;
static int g_pf_depth = 0;    // step kernel: L2 prefetch look-ahead in 32 KB ring stages (measured: <= +2% at 8,
                              // harmful beyond -- 148 SMs x depth x 32 KB must stay well inside one L2 partition)
static int g_sms = 0;
// Internal one-off copies / fills (block tables, layer pointer tables, flag zeroing) go through a NON-BLOCKING
// utility stream and are waited for on the host.  The legacy default stream is never used: a synchronous
// cudaMemcpy on it would wait for whatever else the process has queued there -- e.g. a collective kernel of the
// host framework spinning in a barrier -- and a shard must keep serving requests meanwhile.
static cudaStream_t g_util = nullptr;
static cudaError_t util_h2d(void* dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_util);
  return e != cudaSuccess ? e : cudaStreamSynchronize(g_util);
}
static cudaError_t util_fill0(void* dst, size_t bytes) {
  cudaError_t e = cudaMemsetAsync(dst, 0, bytes, g_util);
  return e != cudaSuccess ? e : cudaStreamSynchronize(g_util);
}
static int g_device = -1;
static bool g_capturing = false;
static long long g_capture_launches = 0;

struct dn_graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  long long kernels = 0;
  size_t nodes = 0;
};

template <typename... KArgs, typename... Args>
static cudaError_t launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                          bool pdl_ok, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (g_pdl && pdl_ok) ? 1 : 0;
  if (g_capturing) g_capture_launches++;
  else g_launches++;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------
struct LayerW {
  const bf16* w[DN_W_COUNT];
  bool bound = false;
  CUtensorMap tm[7];          // TMA descriptors of q,k,v,o,gate,up,down ([rows, K] bf16, box 128 x 64, SWIZZLE_128B)
  bool tm_ok = false;
  // sparse MoE layers (cfg.n_experts > 0): router [E][H] and a device table [3][E] of expert pointers (gate, up, down)
  const bf16* router = nullptr;
  const bf16** etable = nullptr;
  bool experts_bound = false;
};

constexpr int TPF_MAX = 512;   // tokens per tensor-core prefill chunk (weights are streamed once per chunk)
static int g_tc_prefill = 1;
static int g_gemm_bn256 = 0;   // prefill GEMMs: 256-token tiles (one accumulator set for SwiGLU); measured slower than 128-token tiles with two sets (30.7K vs 32.6K tok/s at 2048 tokens)
static int g_tc_attn = 1;      // prefill attention on tcgen05 (GQA groups dividing 128), else the CUDA-core kernel

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled g_encode = nullptr;

static int make_tmap(CUtensorMap* out, const void* base, int rows, int cols, int box_rows) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return fail(DN_ECUDA, "cuTensorMapEncodeTiled unavailable");
    g_encode = (PFN_encodeTiled)fn;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DN_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d", (int)r, rows, cols);
  return DN_OK;
}

struct dn_model {
  dn_model_cfg cfg;
  std::vector<int> abs_layers;
  std::unordered_map<int, int> abs2local;
  std::vector<LayerW> layers;
  const bf16 *embed = nullptr, *norm = nullptr, *head = nullptr;
  int tmax = 1, nsplit = 1, G = 1;
  // scratch (one compute stream per model, like the reference's single compute thread)
  bf16 *hbuf = nullptr, *qbuf = nullptr, *attn = nullptr, *act = nullptr, *logits_bf16 = nullptr;
  float* part = nullptr;
  unsigned int* tickets = nullptr;       // [tmax * n_kv] attention + [1] head
  HeadPartial* head_part = nullptr;
  float* inv_freq = nullptr;
  // sparse MoE scratch: router logits [E], selected experts [k] + scores [k], running weighted sum [H]
  float* moe_logits = nullptr; int32_t* moe_sel = nullptr; float* moe_score = nullptr; float* moe_y = nullptr;
  const bf16** moe_tables = nullptr;    // [local layers][3][E]
  StepState* null_state = nullptr;      // zeroed step state + block table for launches without a nonce (head-part only)
  int32_t* null_bt = nullptr;
  // paged KV pool: [local layer][page][2][n_kv][PAGE][HD]
  bf16* kv_pool = nullptr;
  size_t page_elems = 0, layer_elems = 0;
  std::vector<int> free_pages;
  size_t max_smem = 0;
  // megakernel state
  std::vector<MkLayer> mk_host;
  MkLayer* mk_dev = nullptr;
  bf16 *xa = nullptr, *xb = nullptr;
  // tensor-core prefill scratch ([TPF_MAX][...]) and the TMA descriptors of the activation buffers
  bf16 *pf_xn = nullptr, *pf_qkv = nullptr, *pf_q = nullptr, *pf_attn = nullptr, *pf_h = nullptr, *pf_act = nullptr;
  CUtensorMap tm_xn[4], tm_attn[4], tm_act[4];   // token-tile boxes of 32 / 64 / 128 / 256 rows
  std::vector<CUtensorMap> tm_kv;                // per local layer: KV pool as [pages*2*n_kv*64 rows][128] bf16, box 64 x 64
  bool tm_kv_ok = false;
  bool pf_ok = false;
  unsigned long long* mk_dbg = nullptr;
  size_t mk_dbg_words = 0;
  int* mk_bounds = nullptr;      // [4][sms+1] calibrated row partition of the step kernel (null: equal split)
  bool mk_bounds_on = false;
  unsigned int* mk_sync = nullptr;   // [0] barrier count, [1] generation, [2] error, [3] head ticket
  // quantised KV (cfg.kv_bits 4 / 8): bf16 staging pool the append kernels write into + its block table,
  // per-head score scratch of the two-pass attention, per-head arrival counters of the split exchange
  bf16* kvq_stage = nullptr;
  int32_t* kvq_stage_bt = nullptr;
  float* kvq_scores = nullptr;
  int kvq_score_stride = 0;
  unsigned int* kvq_head_tk = nullptr;
  size_t kv_layer_bytes = 0;         // bytes of one local layer's pages (bf16 or packed)
  // tensor-parallel lm_head: this shard's vocabulary slice (dn_bind_head_slice)
  const bf16* head_slice = nullptr;
  int head_row0 = 0, head_nrows = 0;
};

struct dn_kv {
  dn_model* m = nullptr;
  int max_tokens = 0;
  std::vector<int> pages;
  int32_t* block_table = nullptr;
  StepState* st = nullptr;
  int host_pos = 0;
};

static size_t gemv_smem(int T, int K) { return (size_t)T * K * 2 + (2 * NW * 32 + NW + 96) * sizeof(float); }

template <int T, class Op>
static cudaError_t launch_gemv(const Op& op, int units, cudaStream_t s) {
  const size_t smem = gemv_smem(T, op.K);
  int grid = CTAS_PER_SM * g_sms;
  if (grid > units) grid = units;
  if (grid < 1) grid = 1;
  return launch(k_gemv<T, Op>, dim3(grid), dim3(GEMV_THREADS), smem, s, true, op, g_l2_prefetch_kb * 1024);
}

template <class Op>
static cudaError_t launch_gemv_T(int T, const Op& op, int units, cudaStream_t s) {
  switch (T) {
    case 1: return launch_gemv<1, Op>(op, units, s);
    case 2: return launch_gemv<2, Op>(op, units, s);
    case 4: return launch_gemv<4, Op>(op, units, s);
    default: return cudaErrorInvalidValue;
  }
}

static cudaError_t launch_attn(dn_model* m, const bf16* q, const bf16* pool, const int32_t* bt, const StepState* st,
                               int T, cudaStream_t s) {
  dim3 grid(m->cfg.n_kv_heads * m->nsplit, T);
  const int nh = m->cfg.n_heads, nkv = m->cfg.n_kv_heads, ns = m->nsplit;
#define ATT(Gv)                                                                                          \
  case Gv:                                                                                               \
    return launch(k_attn<Gv>, grid, dim3(Gv * 32), 0, s, true, q, pool, bt, st, m->part, m->tickets, m->attn, nh, nkv, ns);
  switch (m->G) {
    ATT(1) ATT(2) ATT(4) ATT(5) ATT(7) ATT(8)
    default: return cudaErrorInvalidValue;
  }
#undef ATT
}

// ---------------------------------------------------------------------------------
// process / device
// ---------------------------------------------------------------------------------
template <int T, class Op>
static cudaError_t set_gemv_attr() {
  return cudaFuncSetAttribute(k_gemv<T, Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
// every instantiation gets its dynamic-smem limit at init, never inside a graph capture
static cudaError_t init_kernel_attrs() {
  cudaError_t e;
#define SETA(T, Op) if ((e = set_gemv_attr<T, Op>()) != cudaSuccess) return e;
  SETA(1, OpQKV) SETA(2, OpQKV) SETA(4, OpQKV)
  SETA(1, OpOProj) SETA(2, OpOProj) SETA(4, OpOProj)
  SETA(1, OpGateUp) SETA(2, OpGateUp) SETA(4, OpGateUp)
  SETA(1, OpHead)
  SETA(1, OpRouter) SETA(1, OpGateUpMoe) SETA(1, OpDownMoe)
#undef SETA
  // CUDA loads kernels lazily; loading can need a context-wide sync, so the first launch of a
  // kernel issued while k_flag_wait is spinning would deadlock until its timeout.  Touch every
  // kernel once here.
  cudaFuncAttributes fa;
#define PRE(k) if ((e = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return e;
  PRE(k_moe_select) PRE(k_hop_send) PRE(k_flag_set) PRE(k_flag_wait) PRE(k_embed) PRE(k_advance) PRE(k_set_state)
  PRE(k_attn<1>) PRE(k_attn<2>) PRE(k_attn<4>) PRE(k_attn<5>) PRE(k_attn<7>) PRE(k_attn<8>)
#define PRS(Gv) PRE((k_shard_step<Gv, MK_ATT_PLAIN>)) PRE((k_shard_step<Gv, MK_ATT_TC>)) PRE((k_shard_step<Gv, MK_ATT_Q8>)) PRE((k_shard_step<Gv, MK_ATT_Q4>))
  PRS(1) PRS(2) PRS(4) PRS(5) PRS(7) PRS(8)
#undef PRS
#undef PRE
#define ATA(Gv) if ((e = cudaFuncSetAttribute(k_attn_prefill_tc<Gv>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM)) != cudaSuccess) return e;
  ATA(1) ATA(2) ATA(4) ATA(8)
#undef ATA
  return cudaSuccess;
}

extern "C" const char* dn_last_error(void) { return g_err; }
extern "C" const char* dn_version(void) { return "dnet_b200 0.1 (sm_100a)"; }
extern "C" int64_t dn_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" int dn_device_sm_count(void) { return g_sms; }

extern "C" int dn_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(DN_ECUDA, "no CUDA device visible (%s): libdnet_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(DN_EINVAL, "device %d out of range [0,%d)", device, n);
  CK(cudaSetDevice(device));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) return fail(DN_EINVAL, "device %d is sm_%d%d; this library is built for sm_100a only", device, p.major, p.minor);
  g_sms = p.multiProcessorCount;
  g_device = device;
  if (!g_util) CK(cudaStreamCreateWithFlags(&g_util, cudaStreamNonBlocking));
  CK(init_kernel_attrs());
  return DN_OK;
}

extern "C" int dn_set_option(const char* key, int64_t value) {
  if (!key) return fail(DN_EINVAL, "null option key");
  if (!strcmp(key, "pdl")) { g_pdl = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "l2_prefetch_kb")) { g_l2_prefetch_kb = (int)value; return DN_OK; }
  if (!strcmp(key, "tc_prefill")) { g_tc_prefill = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "gemm_bn256")) { g_gemm_bn256 = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "tc_attn")) { g_tc_attn = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "attn_single")) { g_attn_single = (int)value; return DN_OK; }
  if (!strcmp(key, "attn_last")) { g_attn_last = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "attn_tc")) { g_attn_tc = value ? 1 : 0; return DN_OK; }
  if (!strcmp(key, "attn_tc_min")) { g_attn_tc_min = (int)value; return DN_OK; }
  if (!strcmp(key, "mk_debug")) { g_mk_debug = (int)value; return DN_OK; }
  if (!strcmp(key, "inflight_hi")) { g_inflight_hi = value < 0 ? 0 : (int)value; return DN_OK; }
  if (!strcmp(key, "park")) { g_park = value != 0; return DN_OK; }
  if (!strcmp(key, "inflight")) { g_inflight = value < 0 ? 0 : (int)value; return DN_OK; }
  if (!strcmp(key, "attn_chunk")) { g_attn_chunk = value < 32 ? 32 : (int)((value + 31) / 32 * 32); return DN_OK; }
  if (!strcmp(key, "mk_flags")) { g_mk_flags = (int)value; return DN_OK; }
  if (!strcmp(key, "pf_depth")) { g_pf_depth = value < 0 ? 0 : (int)value; return DN_OK; }
  return fail(DN_EINVAL, "unknown option '%s'", key);
}

// ---------------------------------------------------------------------------------
// model
// ---------------------------------------------------------------------------------
extern "C" int dn_model_create(const dn_model_cfg* cfg, const int32_t* abs_layers, int n_layers,
                               const float* inv_freq_host, dn_model** out) {
  if (!cfg || !out || (n_layers > 0 && !abs_layers) || !inv_freq_host) return fail(DN_EINVAL, "null argument");
  if (g_device < 0) return fail(DN_EINVAL, "dn_init() has not been called");
  if (cfg->head_dim != HD) return fail(DN_EINVAL, "head_dim %d unsupported (must be 128)", cfg->head_dim);
  if (cfg->dtype != DN_DTYPE_BF16 || cfg->wire_dtype != cfg->dtype)
    return fail(DN_EINVAL, "only bf16 weights with a bf16 wire dtype are supported (set DNET_TRANSPORT_WIRE_DTYPE=bf16)");
  if (cfg->hidden % 256 || cfg->ffn % 256 || (cfg->n_heads * HD) % 256)
    return fail(DN_EINVAL, "hidden/ffn/q_dim must be multiples of 256 (got %d/%d/%d)", cfg->hidden, cfg->ffn, cfg->n_heads * HD);
  if (cfg->n_kv_heads <= 0 || cfg->n_heads % cfg->n_kv_heads) return fail(DN_EINVAL, "n_heads must be a multiple of n_kv_heads");
  const int G = cfg->n_heads / cfg->n_kv_heads;
  if (!(G == 1 || G == 2 || G == 4 || G == 5 || G == 7 || G == 8)) return fail(DN_EINVAL, "GQA group %d unsupported", G);
  if (cfg->kv_page_tokens != PAGE) return fail(DN_EINVAL, "kv_page_tokens must be %d", PAGE);
  dn_model* m = new (std::nothrow) dn_model();
  if (!m) return fail(DN_ENOMEM, "host allocation failed");
  m->cfg = *cfg;
  m->G = G;
  for (int i = 0; i < n_layers; ++i) {
    m->abs2local[abs_layers[i]] = i;
    m->abs_layers.push_back(abs_layers[i]);
  }
  m->layers.resize(n_layers);
  const int kmax = cfg->ffn > cfg->hidden ? cfg->ffn : cfg->hidden;
  int tmax = 4;
  while (tmax > 1 && gemv_smem(tmax, kmax) > 200 * 1024) tmax >>= 1;
  m->tmax = tmax;
  int ns = g_sms / cfg->n_kv_heads;
  if (ns < 1) ns = 1;
  if (ns > 32) ns = 32;
  m->nsplit = ns;
  const int H = cfg->hidden, qd = cfg->n_heads * HD;
  CK(cudaMalloc(&m->hbuf, (size_t)tmax * H * 2));
  CK(cudaMalloc(&m->qbuf, (size_t)tmax * qd * 2));
  CK(cudaMalloc(&m->attn, (size_t)tmax * qd * 2));
  CK(cudaMalloc(&m->act, (size_t)tmax * cfg->ffn * 2));
  CK(cudaMalloc(&m->logits_bf16, (size_t)cfg->vocab * 2));
  if (cfg->n_experts > 0) {
    if (cfg->n_experts > MOE_MAX_E || cfg->top_k < 1 || cfg->top_k > MOE_MAX_K || cfg->top_k > cfg->n_experts || tmax < cfg->top_k) {
      dn_model_destroy(m);
      return fail(DN_EINVAL, "MoE: %d experts / top-%d unsupported (at most %d experts, top-%d)", cfg->n_experts, cfg->top_k, MOE_MAX_E, MOE_MAX_K);
    }
    CK(cudaMalloc(&m->moe_logits, (size_t)cfg->n_experts * sizeof(float)));
    CK(cudaMalloc(&m->moe_sel, (size_t)MOE_MAX_K * sizeof(int32_t)));
    CK(cudaMalloc(&m->moe_score, (size_t)MOE_MAX_K * sizeof(float)));
    CK(cudaMalloc(&m->moe_y, (size_t)H * sizeof(float)));
    CK(cudaMalloc(&m->moe_tables, (size_t)(n_layers > 0 ? n_layers : 1) * 3 * cfg->n_experts * sizeof(void*)));
    CK(util_fill0(m->moe_sel, MOE_MAX_K * sizeof(int32_t)));
    for (int i = 0; i < n_layers; ++i) m->layers[i].etable = m->moe_tables + (size_t)i * 3 * cfg->n_experts;
  }
  CK(cudaMalloc(&m->part, (size_t)tmax * cfg->n_heads * ns * PART_STRIDE * sizeof(float)));
  CK(cudaMalloc(&m->tickets, ((size_t)tmax * cfg->n_kv_heads + 4) * sizeof(unsigned int)));
  CK(util_fill0(m->tickets, ((size_t)tmax * cfg->n_kv_heads + 4) * sizeof(unsigned int)));
  CK(cudaMalloc(&m->head_part, (size_t)CTAS_PER_SM * g_sms * sizeof(HeadPartial)));
  CK(cudaMalloc(&m->inv_freq, (HD / 2) * sizeof(float)));
  CK(util_h2d(m->inv_freq, inv_freq_host, (HD / 2) * sizeof(float)));
  if (cfg->kv_bits != 0 && cfg->kv_bits != 4 && cfg->kv_bits != 8) { dn_model_destroy(m); return fail(DN_EINVAL, "kv_bits must be 0 (16-bit), 4 or 8"); }
  if (cfg->kv_bits != 0 && cfg->kv_group != 64) { dn_model_destroy(m); return fail(DN_EINVAL, "quantised KV needs group size 64 (got %d)", cfg->kv_group); }
  if (cfg->kv_bits != 0 && cfg->n_heads > g_sms) { dn_model_destroy(m); return fail(DN_EINVAL, "quantised KV: more q heads than SMs"); }
  // elements of the bf16 layout, or the same number of BYTES / 2 for the packed layout (units are even-sized)
  m->page_elems = cfg->kv_bits ? (size_t)2 * cfg->n_kv_heads * kvq_unit_bytes(cfg->kv_bits) / 2 : (size_t)2 * cfg->n_kv_heads * PAGE * HD;
  m->layer_elems = m->page_elems * (size_t)cfg->kv_pool_pages;
  m->kv_layer_bytes = m->layer_elems * 2;
  if (n_layers > 0 && cfg->kv_pool_pages > 0) {
    cudaError_t e = cudaMalloc(&m->kv_pool, m->layer_elems * n_layers * 2);
    if (e != cudaSuccess) {
      fail(DN_ENOMEM, "KV pool of %zu bytes: %s", m->layer_elems * n_layers * 2, cudaGetErrorString(e));
      dn_model_destroy(m);
      return DN_ENOMEM;
    }
  }
  if (m->kv_pool) {
    // never-written rows must be finite: the tcgen05 attention multiplies masked P (= 0) with whatever V holds
    CK(util_fill0(m->kv_pool, m->layer_elems * n_layers * 2));
    m->tm_kv.resize(n_layers);
    m->tm_kv_ok = cfg->kv_bits == 0;       // the tcgen05 prefill attention reads bf16 pages through TMA
    const long long rows = (long long)cfg->kv_pool_pages * 2 * cfg->n_kv_heads * PAGE;
    for (int i = 0; i < n_layers && m->tm_kv_ok; ++i)
      m->tm_kv_ok = rows < (1ll << 31) && make_tmap(&m->tm_kv[i], m->kv_pool + (size_t)i * m->layer_elems, (int)rows, HD, PAGE) == DN_OK;
  }
  for (int p = cfg->kv_pool_pages - 1; p >= 0; --p) m->free_pages.push_back(p);
  if (cfg->kv_bits) {
    const size_t stage_elems = (size_t)KVQ_STAGE_PAGES * 2 * cfg->n_kv_heads * PAGE * HD;
    CK(cudaMalloc(&m->kvq_stage, stage_elems * 2));
    CK(util_fill0(m->kvq_stage, stage_elems * 2));
    std::vector<int32_t> bt(65536);
    for (size_t i = 0; i < bt.size(); ++i) bt[i] = (int32_t)(i % KVQ_STAGE_PAGES);
    CK(cudaMalloc(&m->kvq_stage_bt, bt.size() * sizeof(int32_t)));
    CK(util_h2d(m->kvq_stage_bt, bt.data(), bt.size() * sizeof(int32_t)));
    CK(cudaMalloc(&m->kvq_head_tk, (size_t)cfg->n_heads * sizeof(unsigned int)));
    CK(util_fill0(m->kvq_head_tk, (size_t)cfg->n_heads * sizeof(unsigned int)));
  }
  m->mk_host.resize(n_layers > 0 ? n_layers : 1);
  memset(m->mk_host.data(), 0, m->mk_host.size() * sizeof(MkLayer));
  for (int i = 0; i < n_layers; ++i) m->mk_host[i].kv_pool = m->kv_pool ? m->kv_pool + (size_t)i * m->layer_elems : nullptr;
  CK(cudaMalloc(&m->mk_dev, m->mk_host.size() * sizeof(MkLayer)));
  CK(util_h2d(m->mk_dev, m->mk_host.data(), m->mk_host.size() * sizeof(MkLayer)));
  CK(cudaMalloc(&m->xa, (size_t)H * 2));
  CK(cudaMalloc(&m->xb, (size_t)H * 2));
  CK(cudaMalloc(&m->mk_sync, 64));
  CK(util_fill0(m->mk_sync, 64));
  CK(cudaMalloc(&m->null_state, sizeof(StepState)));
  CK(util_fill0(m->null_state, sizeof(StepState)));
  CK(cudaMalloc(&m->null_bt, 64));
  CK(util_fill0(m->null_bt, 64));
  {
    const int qkvd = (cfg->n_heads + 2 * cfg->n_kv_heads) * HD;
    CK(cudaMalloc(&m->pf_xn, (size_t)TPF_MAX * H * 2));
    CK(cudaMalloc(&m->pf_qkv, (size_t)TPF_MAX * qkvd * 2));
    CK(cudaMalloc(&m->pf_q, (size_t)TPF_MAX * qd * 2));
    CK(cudaMalloc(&m->pf_attn, (size_t)TPF_MAX * qd * 2));
    CK(cudaMalloc(&m->pf_h, (size_t)TPF_MAX * H * 2));
    CK(cudaMalloc(&m->pf_act, (size_t)TPF_MAX * cfg->ffn * 2));
    CK(util_fill0(m->pf_xn, (size_t)TPF_MAX * H * 2));
    CK(util_fill0(m->pf_attn, (size_t)TPF_MAX * qd * 2));
    CK(util_fill0(m->pf_act, (size_t)TPF_MAX * cfg->ffn * 2));
    m->pf_ok = (H % 128 == 0) && (qd % 128 == 0) && (cfg->ffn % 128 == 0) && ((cfg->n_kv_heads * HD) % 128 == 0) &&
               true;
    for (int b = 0; b < 4 && m->pf_ok; ++b) {
      const int box = 32 << b;
      m->pf_ok = make_tmap(&m->tm_xn[b], m->pf_xn, TPF_MAX, H, box) == DN_OK &&
                 make_tmap(&m->tm_attn[b], m->pf_attn, TPF_MAX, qd, box) == DN_OK &&
                 make_tmap(&m->tm_act[b], m->pf_act, TPF_MAX, cfg->ffn, box) == DN_OK;
    }
  }
  *out = m;
  return DN_OK;
}

extern "C" int dn_model_destroy(dn_model* m) {
  if (!m) return DN_OK;
  cudaFree(m->hbuf); cudaFree(m->qbuf); cudaFree(m->attn); cudaFree(m->act); cudaFree(m->logits_bf16);
  cudaFree(m->part); cudaFree(m->tickets); cudaFree(m->head_part); cudaFree(m->inv_freq); cudaFree(m->kv_pool);
  cudaFree(m->mk_bounds); cudaFree(m->pf_xn); cudaFree(m->pf_qkv); cudaFree(m->pf_q); cudaFree(m->pf_attn); cudaFree(m->pf_h); cudaFree(m->pf_act);
  cudaFree(m->mk_dbg); cudaFree(m->mk_dev); cudaFree(m->xa); cudaFree(m->xb); cudaFree(m->mk_sync);
  cudaFree(m->null_state); cudaFree(m->null_bt);
  cudaFree(m->moe_logits); cudaFree(m->moe_sel); cudaFree(m->moe_score); cudaFree(m->moe_y); cudaFree(m->moe_tables);
  cudaFree(m->kvq_stage); cudaFree(m->kvq_stage_bt); cudaFree(m->kvq_scores); cudaFree(m->kvq_head_tk);
  delete m;
  return DN_OK;
}

extern "C" int dn_model_max_chunk(dn_model* m) { return m ? m->tmax : 0; }
// largest chunk for the tensor-core prefill path (0 = unavailable); chunks must be >= 16 tokens
extern "C" int dn_model_max_prefill_chunk(dn_model* m) { return (m && m->pf_ok && g_tc_prefill) ? TPF_MAX : 0; }

extern "C" int dn_bind_layer(dn_model* m, int abs_layer, const void* const* dev_ptrs) {
  if (!m || !dev_ptrs) return fail(DN_EINVAL, "null argument");
  auto it = m->abs2local.find(abs_layer);
  if (it == m->abs2local.end()) return fail(DN_ENOENT, "layer %d not hosted on this model instance", abs_layer);
  LayerW& L = m->layers[it->second];
  for (int i = 0; i < DN_W_COUNT; ++i) {
    L.w[i] = static_cast<const bf16*>(dev_ptrs[i]);
    const bool dense_mlp = i == DN_W_GATE || i == DN_W_UP || i == DN_W_DOWN;
    if (i <= DN_W_LN2 && L.w[i] == nullptr && !(dense_mlp && m->cfg.n_experts > 0))
      return fail(DN_EINVAL, "layer %d: tensor %d is null", abs_layer, i);
    if (((uintptr_t)L.w[i]) & 15) return fail(DN_EINVAL, "layer %d: tensor %d is not 16-byte aligned", abs_layer, i);
  }
  L.bound = true;
  {
    const dn_model_cfg& c = m->cfg;
    const int H = c.hidden, qd = c.n_heads * HD, kd = c.n_kv_heads * HD;
    const int rows[7] = {qd, kd, kd, H, c.ffn, c.ffn, H};
    const int cols[7] = {H, H, H, qd, H, H, c.ffn};
    L.tm_ok = m->pf_ok;
    const int n_tm = m->cfg.n_experts > 0 ? 4 : 7;       // MoE layers: the expert FFN runs token by token, no GEMM descriptors
    for (int i = 0; i < n_tm && L.tm_ok; ++i) L.tm_ok = make_tmap(&L.tm[i], L.w[i], rows[i], cols[i], TC_BM) == DN_OK;
  }
  MkLayer& ML = m->mk_host[it->second];
  for (int i = 0; i < DN_W_COUNT; ++i) ML.w[i] = L.w[i];
  CK(util_h2d(m->mk_dev + it->second, &ML, sizeof(MkLayer)));
  return DN_OK;
}

// MoE layers: the router and the experts' gate / up / down matrices ([ffn][H], [ffn][H], [H][ffn] each, any placement).
// The pointer table is copied to the device here; dn_bind_layer binds the layer's attention tensors and norms.
extern "C" int dn_bind_layer_experts(dn_model* m, int abs_layer, const void* router, const void* const* gate,
                                     const void* const* up, const void* const* down, int n_experts) {
  if (!m || !router || !gate || !up || !down) return fail(DN_EINVAL, "null argument");
  if (m->cfg.n_experts <= 0 || n_experts != m->cfg.n_experts) return fail(DN_EINVAL, "model has %d experts, got %d", m->cfg.n_experts, n_experts);
  auto it = m->abs2local.find(abs_layer);
  if (it == m->abs2local.end()) return fail(DN_ENOENT, "layer %d not hosted on this model instance", abs_layer);
  LayerW& L = m->layers[it->second];
  std::vector<const void*> tab((size_t)3 * n_experts);
  for (int e = 0; e < n_experts; ++e) {
    tab[e] = gate[e]; tab[n_experts + e] = up[e]; tab[2 * n_experts + e] = down[e];
    if (!gate[e] || !up[e] || !down[e]) return fail(DN_EINVAL, "layer %d: expert %d has a null matrix", abs_layer, e);
    if ((((uintptr_t)gate[e]) | ((uintptr_t)up[e]) | ((uintptr_t)down[e])) & 15) return fail(DN_EINVAL, "layer %d: expert %d is not 16-byte aligned", abs_layer, e);
  }
  if (((uintptr_t)router) & 15) return fail(DN_EINVAL, "layer %d: router is not 16-byte aligned", abs_layer);
  CK(util_h2d(L.etable, tab.data(), tab.size() * sizeof(void*)));
  L.router = static_cast<const bf16*>(router);
  L.experts_bound = true;
  return DN_OK;
}

extern "C" int dn_unbind_layer(dn_model* m, int abs_layer) {
  if (!m) return fail(DN_EINVAL, "null argument");
  auto it = m->abs2local.find(abs_layer);
  if (it == m->abs2local.end()) return fail(DN_ENOENT, "layer %d not hosted on this model instance", abs_layer);
  m->layers[it->second].bound = false;
  m->layers[it->second].experts_bound = false;
  memset(m->mk_host[it->second].w, 0, sizeof(m->mk_host[it->second].w));
  memset(m->layers[it->second].w, 0, sizeof(m->layers[it->second].w));
  return DN_OK;
}

extern "C" int dn_layer_is_bound(dn_model* m, int abs_layer) {
  if (!m) return 0;
  auto it = m->abs2local.find(abs_layer);
  return (it != m->abs2local.end() && m->layers[it->second].bound) ? 1 : 0;
}

extern "C" int dn_bind_api(dn_model* m, const void* embed, const void* norm, const void* head) {
  if (!m) return fail(DN_EINVAL, "null argument");
  if ((((uintptr_t)embed) | ((uintptr_t)norm) | ((uintptr_t)head)) & 15) return fail(DN_EINVAL, "api tensors must be 16-byte aligned");
  m->embed = static_cast<const bf16*>(embed);
  m->norm = static_cast<const bf16*>(norm);
  m->head = static_cast<const bf16*>(head);
  if (m->cfg.tie_embeddings && !m->head) m->head = m->embed;
  return DN_OK;
}

// ---------------------------------------------------------------------------------
// per-nonce KV
// ---------------------------------------------------------------------------------
extern "C" int dn_kv_free(dn_kv* kv);
extern "C" int dn_kv_create(dn_model* m, int max_tokens, dn_kv** out) {
  if (!m || !out || max_tokens <= 0) return fail(DN_EINVAL, "bad argument");
  const int np = (max_tokens + PAGE - 1) / PAGE;
  if ((int)m->free_pages.size() < np)
    return fail(DN_ENOSPC, "KV pool exhausted: need %d pages, %zu free", np, m->free_pages.size());
  dn_kv* kv = new (std::nothrow) dn_kv();
  if (!kv) return fail(DN_ENOMEM, "host allocation failed");
  kv->m = m;
  kv->max_tokens = np * PAGE;
  for (int i = 0; i < np; ++i) { kv->pages.push_back(m->free_pages.back()); m->free_pages.pop_back(); }
  cudaError_t e = cudaMalloc(&kv->block_table, (size_t)np * sizeof(int32_t));
  if (e == cudaSuccess) e = util_h2d(kv->block_table, kv->pages.data(), (size_t)np * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&kv->st, sizeof(StepState));
  if (e == cudaSuccess) e = util_fill0(kv->st, sizeof(StepState));
  if (e == cudaSuccess && m->cfg.kv_bits && m->kvq_score_stride < kv->max_tokens) {
    // per-head score scratch of the two-pass quantised attention grows with the longest context a nonce may reach
    float* nb = nullptr;
    const int stride = (kv->max_tokens + 31) / 32 * 32;
    e = cudaMalloc(&nb, (size_t)m->cfg.n_heads * stride * sizeof(float));
    if (e == cudaSuccess) {
      if (m->kvq_scores) { cudaDeviceSynchronize(); cudaFree(m->kvq_scores); }
      m->kvq_scores = nb;
      m->kvq_score_stride = stride;
    }
  }
  if (e != cudaSuccess) {      // give everything back: the pool must not shrink on a failed create
    dn_kv_free(kv);
    return fail(DN_ECUDA, "dn_kv_create: %s", cudaGetErrorString(e));
  }
  *out = kv;
  return DN_OK;
}

extern "C" int dn_kv_free(dn_kv* kv) {
  if (!kv) return DN_OK;
  for (int p : kv->pages) kv->m->free_pages.push_back(p);
  cudaFree(kv->block_table);
  cudaFree(kv->st);
  delete kv;
  return DN_OK;
}

extern "C" int dn_kv_offset(dn_kv* kv) { return kv ? kv->host_pos : -1; }
extern "C" void* dn_kv_token_ptr(dn_kv* kv) { return kv ? (void*)&kv->st->token : nullptr; }

extern "C" int dn_kv_reset(dn_kv* kv, dn_stream s) {
  if (!kv) return fail(DN_EINVAL, "null kv");
  CK(launch(k_set_state, dim3(1), dim3(32), 0, (cudaStream_t)s, false, kv->st, 0, 0, 1, 0));
  kv->host_pos = 0;
  return DN_OK;
}

extern "C" int dn_kv_advance(dn_kv* kv, int T, dn_stream s) {
  if (!kv || T < 0) return fail(DN_EINVAL, "bad argument");
  CK(launch(k_advance, dim3(1), dim3(32), 0, (cudaStream_t)s, true, kv->st, T));
  if (!g_capturing) kv->host_pos += T;
  return DN_OK;
}

extern "C" int dn_kv_seek(dn_kv* kv, int pos, dn_stream s) {
  if (!kv || pos < 0 || pos > kv->max_tokens) return fail(DN_EINVAL, "bad seek position");
  CK(launch(k_set_state, dim3(1), dim3(32), 0, (cudaStream_t)s, false, kv->st, pos, 0, 1, 0));
  if (!g_capturing) kv->host_pos = pos;
  return DN_OK;
}

extern "C" int dn_kv_set_token(dn_kv* kv, int32_t token, dn_stream s) {
  if (!kv) return fail(DN_EINVAL, "null kv");
  CK(launch(k_set_state, dim3(1), dim3(32), 0, (cudaStream_t)s, false, kv->st, 0, (int)token, 0, 1));
  return DN_OK;
}

// host mirror bookkeeping for graph replays (a replay advances the device offset by the
// captured T without passing through dn_kv_advance)
extern "C" int dn_kv_note_advance(dn_kv* kv, int T) {
  if (!kv) return fail(DN_EINVAL, "null kv");
  kv->host_pos += T;
  return DN_OK;
}

// ---------------------------------------------------------------------------------
// operators
// ---------------------------------------------------------------------------------
extern "C" int dn_embed(dn_model* m, const int32_t* ids_dev, int T, void* x_out, dn_stream s) {
  if (!m || !ids_dev || !x_out || T <= 0) return fail(DN_EINVAL, "bad argument");
  if (!m->embed) return fail(DN_ENOENT, "embed_tokens not bound on this shard");
  CK(launch(k_embed, dim3(T), dim3(256), 0, (cudaStream_t)s, true, ids_dev, m->embed, (bf16*)x_out, m->cfg.hidden, m->cfg.vocab));
  return DN_OK;
}

// quantised KV: after the (unchanged) append kernels wrote the T new K/V rows as bf16 into the staging pool, quantise
// them into the packed pool and run the two-pass quantised attention (dn_kvquant.cuh)
static int kvq_append_and_attend(dn_model* m, int li, const bf16* q, bf16* attn_out, int T, dn_kv* kv, cudaStream_t s) {
  const dn_model_cfg& c = m->cfg;
  unsigned char* pool = reinterpret_cast<unsigned char*>(m->kv_pool) + (size_t)li * m->kv_layer_bytes;
  // host mirror: longest context any of the T queries sees (a captured graph is replayed at growing contexts)
  const int kv_len_max = g_capturing ? kv->max_tokens : kv->host_pos + T;
  const size_t smem = ((size_t)((kv_len_max + 31) / 32 * 32) + AQ_WARPS * 132) * sizeof(float);
  if (smem > 200 * 1024) return fail(DN_EINVAL, "quantised-KV attention outside the step kernel supports contexts up to ~50K tokens (got %d)", kv_len_max);
  static size_t smem_set[2] = {0, 0};
  const int bi = c.kv_bits == 8 ? 1 : 0;
  if (smem > smem_set[bi]) {
    const size_t want = smem > 48 * 1024 ? 200 * 1024 : 48 * 1024;
    cudaError_t e = c.kv_bits == 8 ? cudaFuncSetAttribute(k_attn_q<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want)
                                   : cudaFuncSetAttribute(k_attn_q<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
    if (e != cudaSuccess) return fail(DN_ECUDA, "k_attn_q smem attribute: %s", cudaGetErrorString(e));
    smem_set[bi] = want;
  }
  const dim3 ga(T, 2 * c.n_kv_heads), gq(c.n_heads, T);
  if (c.kv_bits == 8) {
    CK(launch(k_kv_quant_append<8>, ga, dim3(32), 0, s, false, (const bf16*)m->kvq_stage, pool, (const int32_t*)kv->block_table, (const StepState*)kv->st, c.n_kv_heads));
    CK(launch(k_attn_q<8>, gq, dim3(AQ_WARPS * 32), smem, s, false, q, (const unsigned char*)pool, (const int32_t*)kv->block_table, (const StepState*)kv->st, attn_out, c.n_heads, c.n_kv_heads));
  } else {
    CK(launch(k_kv_quant_append<4>, ga, dim3(32), 0, s, false, (const bf16*)m->kvq_stage, pool, (const int32_t*)kv->block_table, (const StepState*)kv->st, c.n_kv_heads));
    CK(launch(k_attn_q<4>, gq, dim3(AQ_WARPS * 32), smem, s, false, q, (const unsigned char*)pool, (const int32_t*)kv->block_table, (const StepState*)kv->st, attn_out, c.n_heads, c.n_kv_heads));
  }
  return DN_OK;
}

// sparse MoE FFN of one layer, token by token (dn_kernels.cuh "Sparse MoE FFN"): h = post-attention residual [T][H],
// out = T(h + moe(RMSNorm(h))) [T][H].  4 + 2k launches per token, all with device-side expert indirection.
static int moe_ffn(dn_model* m, const LayerW& L, const bf16* h, bf16* out, int T, cudaStream_t s) {
  const dn_model_cfg& c = m->cfg;
  if (!L.experts_bound) return fail(DN_ENOENT, "MoE layer has no experts bound (dn_bind_layer_experts)");
  const int H = c.hidden, E = c.n_experts, k = c.top_k;
  for (int t = 0; t < T; ++t) {
    OpRouter r;
    r.K = H; r.nrows = E; r.x = h + (size_t)t * H; r.ln_w = L.w[DN_W_LN2]; r.w = L.router; r.logits = m->moe_logits; r.eps = c.rms_eps;
    CK((launch_gemv<1, OpRouter>(r, E, s)));
    CK(launch(k_moe_select, dim3(1), dim3(32), 0, s, false, (const float*)m->moe_logits, E, k, m->moe_sel, m->moe_score));
    for (int j = 0; j < k; ++j) {
      OpGateUpMoe g;
      g.K = H; g.nrows = 2 * c.ffn; g.x = h + (size_t)t * H; g.ln_w = L.w[DN_W_LN2]; g.table = L.etable; g.sel = m->moe_sel;
      g.j = j; g.E = E; g.act = m->act + (size_t)j * c.ffn; g.eps = c.rms_eps;
      CK((launch_gemv<1, OpGateUpMoe>(g, c.ffn, s)));
    }
    for (int j = 0; j < k; ++j) {
      OpDownMoe d;
      d.K = c.ffn; d.nrows = H; d.a = m->act + (size_t)j * c.ffn; d.table = L.etable; d.sel = m->moe_sel; d.score = m->moe_score;
      d.j = j; d.k = k; d.E = E; d.ybuf = m->moe_y; d.resid = h + (size_t)t * H; d.out = out + (size_t)t * H;
      CK((launch_gemv<1, OpDownMoe>(d, H, s)));
    }
  }
  return DN_OK;
}

static int layer_forward(dn_model* m, int abs_layer, bf16* x, int T, dn_kv* kv, cudaStream_t s, cudaEvent_t* evs = nullptr) {
  auto it = m->abs2local.find(abs_layer);
  if (it == m->abs2local.end()) return fail(DN_ENOENT, "Layer %d not hosted on this model instance", abs_layer);
  const int li = it->second;
  const LayerW& L = m->layers[li];
  if (!L.bound) return fail(DN_ENOENT, "layer %d has no weights bound", abs_layer);
  const dn_model_cfg& c = m->cfg;
  const int H = c.hidden, qd = c.n_heads * HD;
  bf16* pool = m->kv_pool + (size_t)li * m->layer_elems;

  OpQKV q;
  q.K = H; q.nrows = (c.n_heads + 2 * c.n_kv_heads) * HD;
  q.x = x; q.ln_w = L.w[DN_W_LN1];
  q.wq = L.w[DN_W_Q]; q.wk = L.w[DN_W_K]; q.wv = L.w[DN_W_V];
  q.bq = L.w[DN_W_QB]; q.bk = L.w[DN_W_KB]; q.bv = L.w[DN_W_VB];
  const bool kvq = c.kv_bits != 0;
  q.q_out = m->qbuf; q.kv_pool = kvq ? m->kvq_stage : pool; q.block_table = kvq ? m->kvq_stage_bt : kv->block_table; q.st = kv->st;
  q.inv_freq = m->inv_freq; q.n_heads = c.n_heads; q.n_kv = c.n_kv_heads; q.eps = c.rms_eps;
  if (evs) CK(cudaEventRecord(evs[0], s));
  CK(launch_gemv_T(T, q, q.nrows / 2, s));
  if (evs) CK(cudaEventRecord(evs[1], s));

  if (kvq) { const int rc = kvq_append_and_attend(m, li, m->qbuf, m->attn, T, kv, s); if (rc) return rc; }
  else CK(launch_attn(m, m->qbuf, pool, kv->block_table, kv->st, T, s));
  if (evs) CK(cudaEventRecord(evs[2], s));

  OpOProj o;
  o.K = qd; o.nrows = H; o.a = m->attn; o.w = L.w[DN_W_O]; o.resid = x; o.out = m->hbuf;
  CK(launch_gemv_T(T, o, o.nrows, s));
  if (evs) CK(cudaEventRecord(evs[3], s));

  if (c.n_experts > 0) {
    const int rc = moe_ffn(m, L, m->hbuf, x, T, s);
    if (rc) return rc;
    if (evs) { CK(cudaEventRecord(evs[4], s)); CK(cudaEventRecord(evs[5], s)); }
    return DN_OK;
  }
  OpGateUp g;
  g.K = H; g.nrows = 2 * c.ffn; g.x = m->hbuf; g.ln_w = L.w[DN_W_LN2];
  g.wg = L.w[DN_W_GATE]; g.wu = L.w[DN_W_UP]; g.act = m->act; g.eps = c.rms_eps;
  CK(launch_gemv_T(T, g, c.ffn, s));
  if (evs) CK(cudaEventRecord(evs[4], s));

  OpDown d;
  d.K = c.ffn; d.nrows = H; d.a = m->act; d.w = L.w[DN_W_DOWN]; d.resid = m->hbuf; d.out = x;
  CK(launch_gemv_T(T, d, d.nrows, s));
  if (evs) CK(cudaEventRecord(evs[5], s));
  return DN_OK;
}

static int head_common(dn_model* m, const void* x, int T, dn_kv* kv, int32_t* token_out, float* logprob_out,
                       float* logits_f32, cudaStream_t s);

// ---- tensor-core prefill path (16 <= T <= TPF_MAX): tcgen05 GEMMs fed by TMA (dn_gemm_tc.cuh)
template <int BN, int EPI, int STAGES>
static cudaError_t launch_tc(const CUtensorMap& w, const CUtensorMap& w2, const CUtensorMap& x, const TcParams& p, cudaStream_t s) {
  constexpr int NA = (EPI == EPI_SWIGLU) ? 2 : 1;
  constexpr size_t smem = (size_t)STAGES * (NA * TC_BM * TC_BK * 2 + BN * TC_BK * 2) + 1024 + 256;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(k_gemm_tc<BN, EPI, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const int tiles = ((p.N + TC_BM - 1) / TC_BM) * ((p.T + BN - 1) / BN);
  int grid = tiles < g_sms ? tiles : g_sms;
  if (g_capturing) g_capture_launches++; else g_launches++;
  k_gemm_tc<BN, EPI, STAGES><<<grid, TC_THREADS, smem, s>>>(w, w2, x, p);
  return cudaGetLastError();
}
// token-tile width: the widest tile that still yields >= ~100 CTAs (a 4096-row matrix has only 32
// row tiles; narrower token tiles re-read W from L2, not from HBM -- concurrent CTAs share the tile)
template <int EPI>
static cudaError_t gemm_tc(const CUtensorMap& w, const CUtensorMap& w2, const CUtensorMap* x /*[4]: 32,64,128,256*/,
                           const TcParams& p, cudaStream_t s) {
  const int n_tiles = (p.N + TC_BM - 1) / TC_BM;
  auto tiles = [&](int bn) { return n_tiles * ((p.T + bn - 1) / bn); };
  // 256-token tiles halve the L2 -> shared-memory re-reads of W at 512-token chunks (the GEMMs are L2-bound there)
  if (g_gemm_bn256 && p.T > 128 && tiles(256) >= 100) return launch_tc<256, EPI, (EPI == EPI_SWIGLU ? 3 : 4)>(w, w2, x[3], p, s);
  if (p.T > 64 && tiles(128) >= 100) return launch_tc<128, EPI, (EPI == EPI_SWIGLU ? 4 : 6)>(w, w2, x[2], p, s);
  if (p.T > 32 && tiles(64) >= 100) return launch_tc<64, EPI, (EPI == EPI_SWIGLU ? 5 : 8)>(w, w2, x[1], p, s);
  if (p.T > 32 && tiles(32) < 100 && tiles(64) * 2 > tiles(32)) return launch_tc<64, EPI, (EPI == EPI_SWIGLU ? 5 : 8)>(w, w2, x[1], p, s);
  return launch_tc<32, EPI, (EPI == EPI_SWIGLU ? 5 : 8)>(w, w2, x[0], p, s);
}

static int layer_forward_tc(dn_model* m, int abs_layer, bf16* x, int T, dn_kv* kv, cudaStream_t s) {
  auto it = m->abs2local.find(abs_layer);
  if (it == m->abs2local.end()) return fail(DN_ENOENT, "Layer %d not hosted on this model instance", abs_layer);
  const int li = it->second;
  LayerW& L = m->layers[li];
  if (!L.bound) return fail(DN_ENOENT, "layer %d has no weights bound", abs_layer);
  const dn_model_cfg& c = m->cfg;
  const int H = c.hidden, qd = c.n_heads * HD, kd = c.n_kv_heads * HD, qkvd = qd + 2 * kd;
  bf16* pool = m->kv_pool + (size_t)li * m->layer_elems;
  unsigned int* err = m->mk_sync + 2;
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.T = T; p.err = err;

  k_rmsnorm_rows<<<T, 256, 0, s>>>(x, L.w[DN_W_LN1], m->pf_xn, H, c.rms_eps);
  g_launches++;
  p.K = H; p.Y = m->pf_qkv; p.ldy = qkvd;
  p.N = qd; p.col0 = 0; p.bias = L.w[DN_W_QB];
  CK(gemm_tc<EPI_STORE>(L.tm[0], L.tm[0], m->tm_xn, p, s));
  p.N = kd; p.col0 = qd; p.bias = L.w[DN_W_KB];
  CK(gemm_tc<EPI_STORE>(L.tm[1], L.tm[1], m->tm_xn, p, s));
  p.col0 = qd + kd; p.bias = L.w[DN_W_VB];
  CK(gemm_tc<EPI_STORE>(L.tm[2], L.tm[2], m->tm_xn, p, s));
  p.bias = nullptr;
  const bool kvq = c.kv_bits != 0;
  k_rope_append<<<dim3(c.n_heads + 2 * c.n_kv_heads, T), 128, 0, s>>>(m->pf_qkv, m->pf_q, kvq ? m->kvq_stage : pool,
                                                                        kvq ? m->kvq_stage_bt : kv->block_table, kv->st, m->inv_freq,
                                                                        c.n_heads, c.n_kv_heads);
  g_launches++;
  if (kvq) {
    const int rc = kvq_append_and_attend(m, li, m->pf_q, m->pf_attn, T, kv, s);
    if (rc) return rc;
  } else if (g_tc_attn && m->tm_kv_ok && (m->G == 1 || m->G == 2 || m->G == 4 || m->G == 8)) {
    AtParams a;
    a.q = m->pf_q; a.out = m->pf_attn; a.block_table = kv->block_table; a.st = kv->st;
    a.n_heads = c.n_heads; a.n_kv = c.n_kv_heads; a.T = T; a.err = err;
    const int tb = 128 / m->G;
    dim3 grid(c.n_kv_heads, (T + tb - 1) / tb);
#define PFT(Gv) case Gv: k_attn_prefill_tc<Gv><<<grid, 256, AT_SMEM, s>>>(m->tm_kv[li], a); break;
    switch (m->G) { PFT(1) PFT(2) PFT(4) PFT(8) default: break; }
#undef PFT
    g_launches++;
  } else {
    dim3 grid(c.n_kv_heads, (T + 3) / 4);
#define PFA(Gv) case Gv: k_attn_prefill<Gv, 4><<<grid, Gv * 32, 0, s>>>(m->pf_q, pool, kv->block_table, kv->st, m->pf_attn, c.n_heads, c.n_kv_heads, T); break;
    switch (m->G) { PFA(1) PFA(2) PFA(4) PFA(5) PFA(7) PFA(8) default: return fail(DN_EINVAL, "GQA group unsupported"); }
#undef PFA
    g_launches++;
  }
  p.K = qd; p.N = H; p.Y = m->pf_h; p.ldy = H; p.col0 = 0; p.resid = x; p.ldr = H;
  CK(gemm_tc<EPI_RESID>(L.tm[3], L.tm[3], m->tm_attn, p, s));
  if (c.n_experts > 0) return moe_ffn(m, L, m->pf_h, x, T, s);
  k_rmsnorm_rows<<<T, 256, 0, s>>>(m->pf_h, L.w[DN_W_LN2], m->pf_xn, H, c.rms_eps);
  g_launches++;
  p.K = H; p.N = c.ffn; p.Y = m->pf_act; p.ldy = c.ffn; p.resid = nullptr;
  CK(gemm_tc<EPI_SWIGLU>(L.tm[4], L.tm[5], m->tm_xn, p, s));
  p.K = c.ffn; p.N = H; p.Y = x; p.ldy = H; p.resid = m->pf_h; p.ldr = H;
  CK(gemm_tc<EPI_RESID>(L.tm[6], L.tm[6], m->tm_act, p, s));
  CK(cudaGetLastError());
  return DN_OK;
}

static int check_fwd(dn_model* m, void* x, int T, dn_kv* kv) {
  if (!m || !x || !kv) return fail(DN_EINVAL, "null argument");
  if (kv->m != m) return fail(DN_EINVAL, "kv belongs to a different model");
  const bool tc_ok = g_tc_prefill && m->pf_ok && T >= 16 && T <= TPF_MAX;
  if (!tc_ok && (!(T == 1 || T == 2 || T == 4) || T > m->tmax))
    return fail(DN_EINVAL, "T=%d unsupported (1,2,4 up to %d, or 16..%d on the tensor-core prefill path)", T, m->tmax, TPF_MAX);
  if (!g_capturing && kv->host_pos + T > kv->max_tokens)
    return fail(DN_ENOSPC, "KV capacity exceeded: offset %d + %d > %d", kv->host_pos, T, kv->max_tokens);
  if (((uintptr_t)x) & 15) return fail(DN_EINVAL, "activation must be 16-byte aligned");
  return DN_OK;
}

extern "C" int dn_layer_forward(dn_model* m, int abs_layer, void* x_inout, int T, dn_kv* kv, dn_stream s) {
  int rc = check_fwd(m, x_inout, T, kv);
  if (rc) return rc;
  if (T >= 16) return layer_forward_tc(m, abs_layer, (bf16*)x_inout, T, kv, (cudaStream_t)s);
  return layer_forward(m, abs_layer, (bf16*)x_inout, T, kv, (cudaStream_t)s);
}

// Per-kernel device times of one layer (CUDA events between the five launches, on the
// launching stream; PDL is off for this call so each kernel is timed alone).  Used by
// bench.py for the live roofline numerator; synchronises the stream.
extern "C" int dn_layer_forward_timed(dn_model* m, int abs_layer, void* x_inout, int T, dn_kv* kv, dn_stream s,
                                      float ms_out[5]) {
  int rc = check_fwd(m, x_inout, T, kv);
  if (rc) return rc;
  if (!ms_out) return fail(DN_EINVAL, "null ms_out");
  if (g_capturing) return fail(DN_EINVAL, "cannot time inside a graph capture");
  cudaEvent_t ev[6];
  for (int i = 0; i < 6; ++i) CK(cudaEventCreate(&ev[i]));
  const int pdl = g_pdl;
  g_pdl = 0;
  rc = layer_forward(m, abs_layer, (bf16*)x_inout, T, kv, (cudaStream_t)s, ev);
  g_pdl = pdl;
  if (rc == DN_OK) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)s);
    if (e != cudaSuccess) rc = fail(DN_ECUDA, "sync: %s", cudaGetErrorString(e));
  }
  if (rc == DN_OK)
    for (int i = 0; i < 5; ++i) cudaEventElapsedTime(&ms_out[i], ev[i], ev[i + 1]);
  for (int i = 0; i < 6; ++i) cudaEventDestroy(ev[i]);
  return rc;
}

// same for the fused final-norm + lm_head + greedy-sample kernel
extern "C" int dn_head_timed(dn_model* m, const void* x, int T, dn_stream s, float* ms_out) {
  if (!ms_out) return fail(DN_EINVAL, "null ms_out");
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  const int pdl = g_pdl;
  g_pdl = 0;
  CK(cudaEventRecord(a, (cudaStream_t)s));
  int rc = head_common(m, x, T, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)s);
  g_pdl = pdl;
  if (rc) return rc;
  CK(cudaEventRecord(b, (cudaStream_t)s));
  CK(cudaStreamSynchronize((cudaStream_t)s));
  CK(cudaEventElapsedTime(ms_out, a, b));
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return DN_OK;
}

extern "C" int dn_window_forward(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, int T, dn_kv* kv, dn_stream s) {
  int rc = check_fwd(m, x_inout, T, kv);
  if (rc) return rc;
  if (n > 0 && !abs_layers) return fail(DN_EINVAL, "null layer list");
  for (int i = 0; i < n; ++i) {
    rc = T >= 16 ? layer_forward_tc(m, abs_layers[i], (bf16*)x_inout, T, kv, (cudaStream_t)s)
                 : layer_forward(m, abs_layers[i], (bf16*)x_inout, T, kv, (cudaStream_t)s);
    if (rc) return rc;
  }
  return DN_OK;
}

static int head_common(dn_model* m, const void* x, int T, dn_kv* kv, int32_t* token_out, float* logprob_out,
                       float* logits_f32, cudaStream_t s) {
  if (!m || !x || T <= 0) return fail(DN_EINVAL, "bad argument");
  if (!m->norm || !m->head) return fail(DN_ENOENT, "final norm / lm_head not bound on this shard");
  OpHead h;
  h.K = m->cfg.hidden; h.nrows = m->cfg.vocab;
  h.x = (const bf16*)x + (size_t)(T - 1) * m->cfg.hidden;  // last position only (result-identical)
  h.ln_w = m->norm; h.w = m->head; h.logits_bf16 = m->logits_bf16; h.logits_f32 = logits_f32;
  h.partials = m->head_part; h.ticket = m->tickets + (size_t)m->tmax * m->cfg.n_kv_heads;
  h.token_out = token_out; h.logprob_out = logprob_out; h.st = kv ? kv->st : nullptr; h.eps = m->cfg.rms_eps;
  CK(launch_gemv<1, OpHead>(h, h.nrows, s));
  return DN_OK;
}

extern "C" int dn_head_sample_greedy(dn_model* m, const void* x, int T, dn_kv* kv, int32_t* token_out,
                                     float* logprob_out, dn_stream s) {
  return head_common(m, x, T, kv, token_out, logprob_out, nullptr, (cudaStream_t)s);
}

extern "C" int dn_head_logits(dn_model* m, const void* x, int T, float* logits_f32_out, void* logits_bf16_out, dn_stream s) {
  int rc = head_common(m, x, T, nullptr, nullptr, nullptr, logits_f32_out, (cudaStream_t)s);
  if (rc) return rc;
  if (logits_bf16_out)
    CK(cudaMemcpyAsync(logits_bf16_out, m->logits_bf16, (size_t)m->cfg.vocab * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  return DN_OK;
}

// ---------------------------------------------------------------------------------
// the single-token shard step as one persistent kernel (dn_megakernel.cuh)
// ---------------------------------------------------------------------------------
static int g_mk_smem_set[4] = {0, 0, 0, 0};
template <int G, int MODE>
static cudaError_t launch_step_mode(const MkParams& p, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  if (!(g_mk_smem_set[MODE] & (1 << G))) {
    e = cudaFuncSetAttribute(k_shard_step<G, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    g_mk_smem_set[MODE] |= (1 << G);
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(g_sms);
  cfg.blockDim = dim3(MK_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident or the launch fails (never a deadlock)
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (g_capturing) g_capture_launches++;
  else g_launches++;
  return cudaLaunchKernelEx(&cfg, k_shard_step<G, MODE>, p);
}
// one instantiation per attention form (dn_megakernel.cuh MK_ATT_*): the launch picks it from the same two fields
// the kernel used to branch on
template <int G>
static cudaError_t launch_step(const MkParams& p, size_t smem, cudaStream_t s) {
  if (p.kv_bits == 8) return launch_step_mode<G, MK_ATT_Q8>(p, smem, s);
  if (p.kv_bits == 4) return launch_step_mode<G, MK_ATT_Q4>(p, smem, s);
  if (p.attn_tc) return launch_step_mode<G, MK_ATT_TC>(p, smem, s);
  return launch_step_mode<G, MK_ATT_PLAIN>(p, smem, s);
}

struct HopArgs {
  const uint32_t* wait_flag = nullptr; uint32_t wait_seq = 0;
  const int32_t* token_in = nullptr;
  void* send_dst = nullptr; uint32_t* send_flag = nullptr; uint32_t send_seq = 0;
  const dn_tp_args* tp = nullptr;
};
static int shard_step_impl(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                           int embed_from_token, int do_head, int32_t* token_out, float* logprob_out,
                           float* logits_f32_out, int advance, const HopArgs& hop, dn_stream s);

extern "C" int dn_shard_step(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                             int embed_from_token, int do_head, int32_t* token_out, float* logprob_out,
                             float* logits_f32_out, int advance, dn_stream s) {
  return shard_step_impl(m, abs_layers, n, x_inout, kv, embed_from_token, do_head, token_out, logprob_out,
                         logits_f32_out, advance, HopArgs(), s);
}

// dn_shard_step with the ring hop fused into the kernel: it spins (bounded) on wait_flag >= wait_seq
// before reading its input (x_inout, or token_in on the first shard) while the weight ring already
// fills, and at the end stores its result (activation, or the sampled token on the last shard) into
// the successor's slot send_dst and releases send_flag = send_seq at system scope.  Any of the three
// groups may be NULL.  Replaces dn_hop_wait + dn_shard_step + dn_hop_send (two launches per hop).
extern "C" int dn_shard_step_hop(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                                 int embed_from_token, int do_head, int32_t* token_out, float* logprob_out, int advance,
                                 const uint32_t* wait_flag, uint32_t wait_seq, const int32_t* token_in,
                                 void* send_dst, uint32_t* send_flag, uint32_t send_seq, dn_stream s) {
  HopArgs h;
  h.wait_flag = wait_flag; h.wait_seq = wait_seq; h.token_in = token_in;
  h.send_dst = send_dst; h.send_flag = send_flag; h.send_seq = send_seq;
  if (send_dst && !send_flag) return fail(DN_EINVAL, "send_dst without send_flag");
  return shard_step_impl(m, abs_layers, n, x_inout, kv, embed_from_token, do_head, token_out, logprob_out, nullptr,
                         advance, h, s);
}

static int shard_step_impl(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv,
                           int embed_from_token, int do_head, int32_t* token_out, float* logprob_out,
                           float* logits_f32_out, int advance, const HopArgs& hop, dn_stream s) {
  const dn_tp_args* tp = hop.tp;
  const bool bubble = tp != nullptr && n == 0;          // tensor-parallel head: a launch that only serves another nonce's head part
  if (!m || n < 0 || (n > 0 && !abs_layers)) return fail(DN_EINVAL, "bad argument");
  if (m->cfg.n_experts > 0) return fail(DN_EINVAL, "MoE models run on the per-op path (dn_window_forward), not in the persistent step kernel");
  if (!bubble && (!x_inout || !kv)) return fail(DN_EINVAL, "bad argument");
  if (kv && kv->m != m) return fail(DN_EINVAL, "kv belongs to a different model");
  if (n == 0 && !do_head && !(tp && (tp->hp_x || tp->mg_n > 0))) return fail(DN_EINVAL, "nothing to do");
  if (!g_capturing && n > 0 && kv->host_pos + 1 > kv->max_tokens) return fail(DN_ENOSPC, "KV capacity exceeded");
  if (tp && tp->hp_x && (!m->head_slice || !m->norm)) return fail(DN_ENOENT, "tensor-parallel head part without a bound lm_head slice / final norm");
  int first_local = 0;
  for (int i = 0; i < n; ++i) {
    auto it = m->abs2local.find(abs_layers[i]);
    if (it == m->abs2local.end()) return fail(DN_ENOENT, "Layer %d not hosted on this model instance", abs_layers[i]);
    if (!m->layers[it->second].bound) return fail(DN_ENOENT, "layer %d has no weights bound", abs_layers[i]);
    if (i == 0) first_local = it->second;
    else if (it->second != first_local + i) return fail(DN_EINVAL, "dn_shard_step needs a contiguous run of local layers");
  }
  if (embed_from_token && !m->embed) return fail(DN_ENOENT, "embed_tokens not bound on this shard");
  if (do_head && (!m->norm || !m->head)) return fail(DN_ENOENT, "final norm / lm_head not bound on this shard");
  const dn_model_cfg& c = m->cfg;
  MkParams p;
  memset(&p, 0, sizeof(p));
  p.layers = m->mk_dev + first_local;
  p.n_layers = n;
  p.H = c.hidden; p.FFN = c.ffn; p.n_heads = c.n_heads; p.n_kv = c.n_kv_heads; p.vocab = c.vocab; p.nsplit = m->nsplit;
  p.eps = c.rms_eps;
  p.x_in = (const bf16*)x_inout;
  p.embed = embed_from_token ? m->embed : nullptr;
  p.xa = m->xa; p.xb = m->xb; p.hbuf = m->hbuf; p.qbuf = m->qbuf; p.attn = m->attn; p.act = m->act;
  p.x_out = (bf16*)x_inout;
  p.block_table = kv ? kv->block_table : m->null_bt; p.st = kv ? kv->st : m->null_state; p.inv_freq = m->inv_freq;
  if (!kv && !m->null_state) return fail(DN_EINVAL, "no null step state");
  p.part = m->part; p.tickets = m->tickets;
  p.norm_w = m->norm; p.head_w = m->head; p.logits_bf16 = m->logits_bf16; p.logits_f32 = logits_f32_out;
  p.head_part = m->head_part; p.head_ticket = m->mk_sync + 3;
  p.token_out = token_out; p.logprob_out = logprob_out; p.do_head = do_head ? 1 : 0; p.advance = advance ? 1 : 0;
  p.bar_count = m->mk_sync; p.bar_epoch = m->mk_sync + 1; p.err = m->mk_sync + 2;
  p.pf_depth = g_pf_depth;
  p.attn_chunk = g_attn_chunk;
  p.attn_single = g_attn_single;
  p.attn_last = g_attn_last;
  p.inflight = g_inflight;
  p.inflight_hi = g_inflight_hi > g_inflight ? g_inflight_hi : g_inflight;
  // TMEM parking needs one fragment geometry (seg == 1024) in every phase
  p.park = (g_park && c.hidden % 1024 == 0 && c.ffn % 1024 == 0 && (c.n_heads * HD) % 1024 == 0) ? 1 : 0;
  p.head_rows = c.vocab;
  if (tp) {
    if (tp->hp_x) {
      p.hp_x = (const bf16*)tp->hp_x; p.hp_wait_flag = tp->hp_wait_flag; p.hp_seq = tp->hp_seq;
      p.hp_row0 = m->head_row0; p.head_rows = m->head_nrows; p.head_w = m->head_slice;
      p.hp_dst = (float*)tp->hp_dst; p.hp_dst_flag = tp->hp_dst_flag;
      if (!p.hp_dst || !p.hp_dst_flag) return fail(DN_EINVAL, "head part without a destination");
    }
    if (tp->bc_n < 0 || tp->bc_n > 16 || tp->mg_n < 0 || tp->mg_n > 16) return fail(DN_EINVAL, "ring larger than 16 shards");
    p.bc_n = tp->bc_n; p.bc_seq = tp->bc_seq;
    for (int i = 0; i < tp->bc_n; ++i) { p.bc_dst[i] = (bf16*)tp->bc_dst[i]; p.bc_flag[i] = tp->bc_flag[i]; }
    p.mg_n = tp->mg_n; p.mg_part = (const float*)tp->mg_part; p.mg_flags = tp->mg_flags; p.mg_seq = tp->mg_seq;
    p.mg_st = tp->mg_kv ? tp->mg_kv->st : nullptr; p.mg_token_out = tp->mg_token_out; p.mg_logprob_out = tp->mg_logprob_out;
    p.mg_slot = (int32_t*)tp->mg_slot; p.mg_slot_flag = tp->mg_slot_flag; p.mg_slot_seq = tp->mg_slot_seq;
  }
  p.wait_flag = hop.wait_flag; p.wait_seq = hop.wait_seq; p.token_in = hop.token_in;
  p.send_dst = hop.send_dst; p.send_flag = hop.send_flag; p.send_seq = hop.send_seq;
  p.flags = g_mk_flags;
  p.bounds = m->mk_bounds_on ? m->mk_bounds : nullptr;
  p.dbg = nullptr;
  if (g_mk_debug) {
    const size_t words = (size_t)g_sms * (n > 0 ? n : 1) * dn::MK_DBG_WORDS;
    if (m->mk_dbg_words < words) {
      cudaFree(m->mk_dbg);
      CK(cudaMalloc(&m->mk_dbg, words * 8));
      m->mk_dbg_words = words;
    }
    p.dbg = m->mk_dbg;
  }
  p.bar_gen = m->mk_sync + 4;
  p.kv_bits = c.kv_bits;
  p.kv_stage = m->kvq_stage;
  p.sc_buf = m->kvq_scores;
  p.sc_stride = m->kvq_score_stride;
  p.head_tk = m->kvq_head_tk;
  if (c.kv_bits && kv && (!m->kvq_scores || m->kvq_score_stride < kv->max_tokens)) return fail(DN_EINVAL, "quantised KV: score scratch smaller than the nonce's capacity");
  const int kmax = c.ffn > c.hidden ? c.ffn : c.hidden;
  int scratch = kmax * 2;
  const int attn_bytes = 8 * 132 * 4 + 64;                 // per-warp attention partials
  if (scratch < attn_bytes) scratch = attn_bytes;
  const int merge_bytes = c.n_heads * HD * 2 + c.n_heads * m->nsplit * 8 + 64;   // o_proj vector + (m,l) table
  if (scratch < merge_bytes) scratch = merge_bytes;
  // long contexts: attention on the tensor cores, its per-warp K/V tiles alias the activation scratch (one ring stage less)
  p.attn_tc = (g_attn_tc && c.kv_bits == 0 && kv && n > 0 && kv->host_pos + 1 >= g_attn_tc_min && m->G <= 8) ? 1 : 0;
  if (p.attn_tc && scratch < ATC_SMEM_BYTES) scratch = ATC_SMEM_BYTES;
  if (c.hidden > 8192) return fail(DN_EINVAL, "hidden > 8192 unsupported by the step kernel's RMSNorm staging");
  scratch = (scratch + 1023) / 1024 * 1024;
  const int tail = 2 * MK_MAX_STAGES * 8 + 64 * 4 + 128 * 4 + 2 * 8 * 16 * 4;   // barriers, misc scratch, RoPE table, row-block partials
  int stages = (227 * 1024 - scratch - tail) / MK_STAGE_BYTES;
  if (stages > MK_MAX_STAGES) stages = MK_MAX_STAGES;
  if (stages < 2) return fail(DN_EINVAL, "model too wide for the megakernel's shared-memory ring");
  p.n_stages = stages;
  // the cap waits on the full-barrier of an earlier use of a slot: it must stay below the ring depth (parity aliasing)
  // Safe caps with two producers taking alternate stages: the slot a producer waits on is re-armed
  // `stages - cap` stages later; that must be its own stage (even distance) or one the other producer
  // cannot reach first (cap <= stages/2), else the full-barrier can run two phases ahead of the waited
  // parity and the wait only returns one ring revolution later (measured: 6 tok/s at cap 5 of 6).
  auto safe_cap = [stages](int cap) {
    if (cap >= stages) cap = stages - 1;
    while (cap > stages / 2 && ((stages - cap) & 1)) --cap;
    return cap < 0 ? 0 : cap;
  };
  p.inflight = safe_cap(p.inflight);
  p.inflight_hi = safe_cap(p.inflight_hi);
  if (p.inflight_hi < p.inflight) p.inflight_hi = p.inflight;
  p.scratch_bytes = scratch;
  const size_t smem = (size_t)stages * MK_STAGE_BYTES + scratch + tail;
  if ((size_t)g_sms > (size_t)CTAS_PER_SM * g_sms) return fail(DN_EINVAL, "head partial buffer too small");
  cudaError_t e;
  switch (m->G) {
    case 1: e = launch_step<1>(p, smem, (cudaStream_t)s); break;
    case 2: e = launch_step<2>(p, smem, (cudaStream_t)s); break;
    case 4: e = launch_step<4>(p, smem, (cudaStream_t)s); break;
    case 5: e = launch_step<5>(p, smem, (cudaStream_t)s); break;
    case 7: e = launch_step<7>(p, smem, (cudaStream_t)s); break;
    case 8: e = launch_step<8>(p, smem, (cudaStream_t)s); break;
    default: return fail(DN_EINVAL, "GQA group %d unsupported", m->G);
  }
  if (e != cudaSuccess) return fail(DN_ECUDA, "k_shard_step launch: %s", cudaGetErrorString(e));
  if (advance && !g_capturing && kv) kv->host_pos += 1;
  return DN_OK;
}

// Tensor-parallel lm_head: bind this shard's vocabulary slice [row0, row0 + nrows) ([nrows, hidden] bf16, borrowed) --
// the final norm comes through dn_bind_api as before.
extern "C" int dn_bind_head_slice(dn_model* m, const void* slice, int row0, int nrows) {
  if (!m || !slice || row0 < 0 || nrows <= 0 || row0 + nrows > m->cfg.vocab) return fail(DN_EINVAL, "bad head slice");
  if (((uintptr_t)slice) & 15) return fail(DN_EINVAL, "head slice must be 16-byte aligned");
  m->head_slice = static_cast<const bf16*>(slice);
  m->head_row0 = row0; m->head_nrows = nrows;
  return DN_OK;
}

// dn_shard_step_hop with the lm_head tensor-parallel over the ring (dn_tp_args): optional head part of another
// nonce first, the layers, optional broadcast of the final hidden state (last shard), optional merge of the S
// partials into a token (head shard).  n == 0 with kv == NULL is a launch that only serves head part / merge.
extern "C" int dn_shard_step_tp(dn_model* m, const int32_t* abs_layers, int n, void* x_inout, dn_kv* kv, int embed_from_token,
                                int advance, const uint32_t* wait_flag, uint32_t wait_seq, const int32_t* token_in,
                                void* send_dst, uint32_t* send_flag, uint32_t send_seq, const dn_tp_args* tp, dn_stream s) {
  if (!tp) return fail(DN_EINVAL, "null tp args");
  HopArgs h;
  h.wait_flag = wait_flag; h.wait_seq = wait_seq; h.token_in = token_in;
  h.send_dst = send_dst; h.send_flag = send_flag; h.send_seq = send_seq;
  h.tp = tp;
  if (send_dst && !send_flag) return fail(DN_EINVAL, "send_dst without send_flag");
  return shard_step_impl(m, abs_layers, n, x_inout, kv, n > 0 ? embed_from_token : 0, 0, nullptr, nullptr, nullptr,
                         n > 0 ? advance : 0, h, s);
}

// phase timestamps of the last dn_shard_step run with option mk_debug=1: [sm][layer][16] ns
extern "C" int dn_step_debug(dn_model* m, unsigned long long* out_host, size_t max_words, dn_stream s) {
  if (!m || !out_host) return fail(DN_EINVAL, "null argument");
  if (!m->mk_dbg) return fail(DN_ENOENT, "no debug stamps recorded (set option mk_debug=1 first)");
  const size_t n = m->mk_dbg_words < max_words ? m->mk_dbg_words : max_words;
  CK(cudaMemcpyAsync(out_host, m->mk_dbg, n * 8, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  CK(cudaStreamSynchronize((cudaStream_t)s));
  return (int)(n / 16);
}

// Row partition of the four weight phases of k_shard_step over the SMs: bounds_host is
// [4][sms+1] (QKV, O, GATE/UP, DOWN), non-decreasing, bounds[ph][0] = 0, bounds[ph][sms] = rows of the
// phase (virtual rows: pairs interleaved for QKV and GATE/UP, so entries must be even there).
// Used to give faster SMs proportionally more rows (measured with option mk_debug); results do not
// depend on the partition (each output row has a fixed summation order).  NULL restores the equal split.
extern "C" int dn_step_set_bounds(dn_model* m, const int32_t* bounds_host) {
  if (!m) return fail(DN_EINVAL, "null model");
  if (!bounds_host) { m->mk_bounds_on = false; return DN_OK; }
  const dn_model_cfg& c = m->cfg;
  const int rows[4] = {(c.n_heads + 2 * c.n_kv_heads) * HD, c.hidden, 2 * c.ffn, c.hidden};
  const int align[4] = {2, 1, 2, 1};
  const int n = g_sms + 1;
  for (int ph = 0; ph < 4; ++ph) {
    const int32_t* b = bounds_host + ph * n;
    if (b[0] != 0 || b[g_sms] != rows[ph]) return fail(DN_EINVAL, "bounds of phase %d must span [0, %d]", ph, rows[ph]);
    for (int i = 0; i < g_sms; ++i)
      if (b[i + 1] < b[i] || (b[i] % align[ph])) return fail(DN_EINVAL, "bounds of phase %d not monotone / aligned at %d", ph, i);
  }
  if (!m->mk_bounds) CK(cudaMalloc(&m->mk_bounds, (size_t)4 * n * sizeof(int)));
  CK(cudaDeviceSynchronize());   // calibration only: no step kernel may be reading the old bounds
  CK(util_h2d(m->mk_bounds, bounds_host, (size_t)4 * n * sizeof(int)));
  m->mk_bounds_on = true;
  return DN_OK;
}

// 0 = clean; 2/3 = a bounded spin inside k_shard_step timed out (results of that step are invalid)
extern "C" int dn_step_error(dn_model* m, dn_stream s) {
  if (!m) return fail(DN_EINVAL, "null model");
  unsigned int v = 0;
  CK(cudaMemcpyAsync(&v, m->mk_sync + 2, 4, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  CK(cudaStreamSynchronize((cudaStream_t)s));
  return (int)v;
}

// debugging / parity bisection: copy one of the per-op path's scratch buffers of the LAST chunk to the host
// (which: 0 q after RoPE [tmax][n_heads*128], 1 attention output, 2 h = x + o_proj [tmax][H], 3 SwiGLU output [tmax][ffn])
extern "C" int dn_debug_scratch(dn_model* m, int which, void* host_out, size_t bytes, dn_stream s) {
  if (!m || !host_out) return fail(DN_EINVAL, "null argument");
  const bf16* src = which == 0 ? m->qbuf : which == 1 ? m->attn : which == 2 ? m->hbuf : which == 3 ? m->act : nullptr;
  if (!src) return fail(DN_EINVAL, "unknown scratch buffer %d", which);
  CK(cudaMemcpyAsync(host_out, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)s));
  CK(cudaStreamSynchronize((cudaStream_t)s));
  return DN_OK;
}

// the error word is sticky (every later step reports it): clear it once the failed request was dropped
extern "C" int dn_step_error_clear(dn_model* m, dn_stream s) {
  if (!m) return fail(DN_EINVAL, "null model");
  CK(cudaMemsetAsync(m->mk_sync + 2, 0, 4, (cudaStream_t)s));
  return DN_OK;
}

// ---------------------------------------------------------------------------------
// graphs
// ---------------------------------------------------------------------------------
extern "C" int dn_graph_begin(dn_stream s) {
  if (g_capturing) return fail(DN_EINVAL, "a capture is already in progress");
  CK(cudaStreamBeginCapture((cudaStream_t)s, cudaStreamCaptureModeThreadLocal));
  g_capturing = true;
  g_capture_launches = 0;
  return DN_OK;
}

extern "C" int dn_graph_end(dn_stream s, dn_graph** out) {
  if (!g_capturing) return fail(DN_EINVAL, "no capture in progress");
  g_capturing = false;
  dn_graph* g = new (std::nothrow) dn_graph();
  if (!g) return fail(DN_ENOMEM, "host allocation failed");
  cudaError_t e = cudaStreamEndCapture((cudaStream_t)s, &g->graph);
  if (e != cudaSuccess || !g->graph) { delete g; return fail(DN_ECUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); }
  e = cudaGraphInstantiate(&g->exec, g->graph, 0);
  if (e != cudaSuccess) { cudaGraphDestroy(g->graph); delete g; return fail(DN_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
  g->kernels = g_capture_launches;
  cudaGraphGetNodes(g->graph, nullptr, &g->nodes);
  *out = g;
  return DN_OK;
}

extern "C" int dn_graph_launch(dn_graph* g, dn_stream s) {
  if (!g) return fail(DN_EINVAL, "null graph");
  CK(cudaGraphLaunch(g->exec, (cudaStream_t)s));
  g_launches += g->kernels;
  return DN_OK;
}

extern "C" int dn_graph_destroy(dn_graph* g) {
  if (!g) return DN_OK;
  cudaGraphExecDestroy(g->exec);
  cudaGraphDestroy(g->graph);
  delete g;
  return DN_OK;
}

extern "C" int dn_graph_num_nodes(dn_graph* g) { return g ? (int)g->nodes : 0; }

// ---------------------------------------------------------------------------------
// ring hop
// ---------------------------------------------------------------------------------
extern "C" int dn_hop_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes == 0) return fail(DN_EINVAL, "bad argument");
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return fail(DN_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  CK(util_fill0(*dev_ptr, bytes));
  return DN_OK;
}
extern "C" int dn_hop_free(void* p) { if (p) CK(cudaFree(p)); return DN_OK; }
extern "C" int dn_hop_export(void* dev_ptr, uint8_t handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, dev_ptr));
  memcpy(handle_out, &h, 64);
  return DN_OK;
}
extern "C" int dn_hop_import(const uint8_t handle[64], void** dev_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return DN_OK;
}
extern "C" int dn_hop_close(void* p) { if (p) CK(cudaIpcCloseMemHandle(p)); return DN_OK; }
extern "C" int dn_enable_peer(int peer) {
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, g_device, peer));
  if (!can) return fail(DN_EINVAL, "device %d cannot access peer %d", g_device, peer);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return DN_OK; }
  CK(e);
  return DN_OK;
}
extern "C" int dn_hop_send(void* dst_slot, const void* src, size_t bytes, uint32_t* dst_flag, uint32_t seq, dn_stream s) {
  if (!dst_slot || !src || !dst_flag) return fail(DN_EINVAL, "null argument");
  if (bytes <= 65536 && (bytes % 16) == 0 && !((((uintptr_t)dst_slot) | ((uintptr_t)src)) & 15)) {
    CK(launch(k_hop_send, dim3(1), dim3(512), 0, (cudaStream_t)s, false, (uint4*)dst_slot, (const uint4*)src,
              (int)(bytes / 16), dst_flag, seq));
    return DN_OK;
  }
  CK(cudaMemcpyAsync(dst_slot, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  CK(launch(k_flag_set, dim3(1), dim3(32), 0, (cudaStream_t)s, false, dst_flag, seq));
  return DN_OK;
}
extern "C" int dn_hop_wait(const uint32_t* flag, uint32_t seq, uint32_t timeout_ms, uint32_t* err_flag, dn_stream s) {
  if (!flag) return fail(DN_EINVAL, "null argument");
  CK(launch(k_flag_wait, dim3(1), dim3(32), 0, (cudaStream_t)s, false, flag, seq,
            (unsigned long long)timeout_ms * 1000000ull, err_flag));
  return DN_OK;
}

// measurement hook: `iters` trips of a (bytes + flag) token around the ring on an otherwise unused lane; the origin
// rank's *out_ns (device-accessible) = elapsed ns for all trips, measured on its own clock (see k_hop_ring_probe)
extern "C" int dn_hop_ring_probe(const void* own_slot, const uint32_t* own_flag, void* next_slot, uint32_t* next_flag,
                                 size_t bytes, uint32_t base_seq, int iters, int is_origin, uint32_t timeout_ms,
                                 unsigned long long* out_ns, dn_stream s) {
  if (!own_slot || !own_flag || !next_slot || !next_flag || iters <= 0 || (bytes % 16)) return fail(DN_EINVAL, "bad argument");
  CK(launch(k_hop_ring_probe, dim3(1), dim3(512), 0, (cudaStream_t)s, false, (const uint4*)own_slot, own_flag, (uint4*)next_slot,
            next_flag, (int)(bytes / 16), base_seq, iters, is_origin ? 1 : 0, (unsigned long long)timeout_ms * 1000000ull, out_ns));
  return DN_OK;
}

// ---------------------------------------------------------------------------------
// layer swap / plumbing
// ---------------------------------------------------------------------------------
extern "C" int dn_pinned_alloc(size_t bytes, void** host_ptr) {
  if (!host_ptr || bytes == 0) return fail(DN_EINVAL, "bad argument");
  cudaError_t e = cudaHostAlloc(host_ptr, bytes, cudaHostAllocPortable);
  if (e != cudaSuccess) return fail(DN_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
  return DN_OK;
}
extern "C" int dn_pinned_free(void* p) { if (p) CK(cudaFreeHost(p)); return DN_OK; }
extern "C" int dn_device_alloc(size_t bytes, void** dev_ptr) {
  if (!dev_ptr || bytes == 0) return fail(DN_EINVAL, "bad argument");
  cudaError_t e = cudaMalloc(dev_ptr, bytes);
  if (e != cudaSuccess) return fail(DN_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  return DN_OK;
}
extern "C" int dn_device_free(void* p) { if (p) CK(cudaFree(p)); return DN_OK; }
extern "C" int dn_slot_prefetch(void* dst_dev, const void* src_pinned, size_t bytes, dn_stream prefetch, dn_event done) {
  if (!dst_dev || !src_pinned) return fail(DN_EINVAL, "null argument");
  CK(cudaMemcpyAsync(dst_dev, src_pinned, bytes, cudaMemcpyHostToDevice, (cudaStream_t)prefetch));
  if (done) CK(cudaEventRecord((cudaEvent_t)done, (cudaStream_t)prefetch));
  return DN_OK;
}
extern "C" int dn_stream_create(dn_stream* out, int high_priority) {
  if (!out) return fail(DN_EINVAL, "null argument");
  int lo = 0, hi = 0;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t s;
  CK(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
  *out = s;
  return DN_OK;
}
extern "C" int dn_stream_destroy(dn_stream s) { CK(cudaStreamDestroy((cudaStream_t)s)); return DN_OK; }
extern "C" int dn_stream_sync(dn_stream s) { CK(cudaStreamSynchronize((cudaStream_t)s)); return DN_OK; }
extern "C" int dn_stream_wait_event(dn_stream s, dn_event e) { CK(cudaStreamWaitEvent((cudaStream_t)s, (cudaEvent_t)e, 0)); return DN_OK; }
extern "C" int dn_event_create(dn_event* out, int timing) {
  if (!out) return fail(DN_EINVAL, "null argument");
  cudaEvent_t e;
  CK(cudaEventCreateWithFlags(&e, timing ? cudaEventDefault : cudaEventDisableTiming));
  *out = e;
  return DN_OK;
}
extern "C" int dn_event_destroy(dn_event e) { CK(cudaEventDestroy((cudaEvent_t)e)); return DN_OK; }
extern "C" int dn_event_record(dn_event e, dn_stream s) { CK(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)s)); return DN_OK; }
extern "C" int dn_event_query(dn_event e) {
  cudaError_t r = cudaEventQuery((cudaEvent_t)e);
  if (r == cudaSuccess) return 1;
  if (r == cudaErrorNotReady) { cudaGetLastError(); return 0; }
  return fail(DN_ECUDA, "cudaEventQuery: %s", cudaGetErrorString(r));
}
extern "C" int dn_event_sync(dn_event e) { CK(cudaEventSynchronize((cudaEvent_t)e)); return DN_OK; }
extern "C" int dn_event_elapsed_ms(dn_event a, dn_event b, float* ms) {
  CK(cudaEventElapsedTime(ms, (cudaEvent_t)a, (cudaEvent_t)b));
  return DN_OK;
}
extern "C" int dn_memcpy_h2d(void* dst, const void* src, size_t bytes, dn_stream s) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)s));
  return DN_OK;
}
extern "C" int dn_memcpy_d2h(void* dst, const void* src, size_t bytes, dn_stream s) {
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)s));   // dst may be pinned host or device
  return DN_OK;
}
