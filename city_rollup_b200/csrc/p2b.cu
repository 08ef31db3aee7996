// p2b.cu — C ABI (include/p2b.h) over the sm_100a kernels: contexts, device-resident handles,
// PolynomialBatch::{from_values,from_coeffs}, MerkleTree::new, Poseidon utilities, Challenger, FRI.
// No CPU fallback: every entry point fails with P2B_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/p2b.h"
#include "fri_kernels.cuh"
#include "fused_kernels.cuh"
#include "hash_kernels.cuh"
#include "ntt2_kernels.cuh"
#include "ntt_kernels.cuh"
#include "plonk_kernels.cuh"
#include "prover_kernels.cuh"

#define P2B_VERSION 100

static thread_local std::string g_init_error;

// a captured proof (see "CUDA-graph plans" below)
struct ProvePlan {
  enum { SEEN = 0, READY = 1, FAILED = -1 };
  uint64_t circuit_id = 0, cs_id = 0;
  p2b_fri_params fp{};
  size_t n_pis = 0;
  bool dev_src = false;
  int state = SEEN;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  // pinned host block the graph's copy nodes read: parameter words and pointer tables, bump-allocated during capture
  uint64_t* h_pin = nullptr;
  size_t pin_cap = 0, pin_used = 0;
  uint64_t *h_digest = nullptr, *h_pis = nullptr;  // slots inside h_pin, refreshed before every launch
  uint64_t* h_wires = nullptr;                     // pinned staging matrix for host witnesses
  uint64_t* h_proof = nullptr;                     // pinned: where the graph leaves the proof words
  uint64_t* d_wires_dst = nullptr;                 // destination of the witness copy node
  cudaGraphNode_t wires_node{};
  bool has_wires_node = false;
  const void* cur_src = nullptr;  // what the witness copy node currently reads
  size_t proof_len = 0, wires_words = 0;
  uint32_t n_kernels = 0;
  uint64_t last_use = 0;
};

struct p2b_ctx {
  int device = 0;
  int sm_count = 148;  // B200; read from the device at init (persistent-style grids are sized from it)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // Every context allocates from its OWN stream-ordered pool: with the shared default pool a block freed on one
  // context's stream is handed to another context's next allocation behind an internal event dependency, which
  // chains the streams of concurrently proving contexts together.
  cudaMemPool_t pool = nullptr;
  // second stream + events: host-to-device copies of a batch run ahead of the transforms of the previous
  // column group (batch_from_host)
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  // pinned double buffer for pageable host columns (plonky2's Vec<PolynomialValues>: one pageable allocation per
  // column): the worker thread memcpy's into one half while the DMA of the other half is in flight
  uint64_t* h_upload[2] = {nullptr, nullptr};
  cudaEvent_t upload_done[2] = {nullptr, nullptr};
  int upload_half = 0;
  // host waits: spin (cudaStreamSynchronize, lowest latency, one busy core per waiting thread) or sleep on a
  // blocking-sync event (p2b_set_blocking_sync: many contexts per host core)
  bool blocking_sync = false;
  // p2b_set_latency_mode: one proof at a time on this GPU — the Merkle trees fused from 2^15 digests (12 launches less per
  // proof) and the proof-of-work search on every SM (one stride of candidates covers the mean); see the setter
  bool latency_mode = false;
  cudaEvent_t ev_block = nullptr;
  bool poisoned = false;
  std::string err;
  uint64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  // twiddle tables (device)
  uint64_t* d_w12 = nullptr;
  uint64_t* d_rlo = nullptr;
  uint64_t* d_rhi = nullptr;
  // radix-16 path (ntt2_kernels.cuh): row-stage tables, and lazily built per-size tables kept for the
  // life of the context (plonky2's FftRootTable analogue): strided-step twiddles keyed by (log_m, d),
  // coset power tables keyed by (shift, log_n, rate_bits)
  uint64_t* d_t1 = nullptr;
  uint64_t* d_t2 = nullptr;
  std::map<uint64_t, uint64_t*> tw_cache;
  struct CpKey {
    uint64_t shift;
    uint32_t log_n, rate_bits;
    bool operator<(const CpKey& o) const {
      return shift != o.shift ? shift < o.shift : log_n != o.log_n ? log_n < o.log_n : rate_bits < o.rate_bits;
    }
  };
  std::map<CpKey, uint64_t*> cp_cache;
  size_t cp_cache_bytes = 0;
  // optional per-stage timing
  bool profiling = false;
  struct StageRec {
    int stage;
    cudaEvent_t a, b;
    uint64_t launches;
  };
  std::vector<StageRec> recs;
  std::vector<cudaEvent_t> ev_pool;
  int cur_stage = -1;
  cudaEvent_t cur_ev = nullptr;
  uint64_t cur_launch0 = 0;
  // small pinned staging buffer for accessor results
  uint64_t* h_stage = nullptr;
  size_t h_stage_bytes = 0;
  // device scratch words of the fused kernels: [0] the "CTAs done" counter of fusedk::k_tree_subtree (self-resetting),
  // [1] the best proof-of-work witness of the current search
  uint64_t* d_scratch = nullptr;
  // old-path (n < 2^12) coset power tables, keyed like cp_cache: FRI layers of every proof reuse the same few
  struct SmallCp {
    uint64_t *lo, *hi;
    uint32_t hi_count;
  };
  std::map<CpKey, SmallCp> small_cp_cache;
  // prove plans (CUDA graphs) and the proof submitted but not yet collected
  std::vector<ProvePlan*> plans;
  ProvePlan* cap = nullptr;  // non-null while prove_body runs under stream capture
  uint64_t plan_clock = 0;
  std::string plan_note;
  size_t pending_words = 0, pending_pow_index = 0;
  const uint64_t* pending_src = nullptr;
  uint64_t* h_proof = nullptr;  // pinned landing buffer of eagerly proven proofs
  size_t h_proof_bytes = 0;
  // completion of the last DMA that read caller-owned pinned memory (host-input entry points wait on it before
  // returning, so the caller may refill its buffers)
  cudaEvent_t ev_h2d = nullptr;
  bool h2d_event_pending = false;
  bool direct_src_pending = false;  // a replayed prove plan reads the caller's pinned witness (no separate upload event)
  nttk::Tables tables() const { return nttk::Tables{d_w12, d_rlo, d_rhi}; }
  ntt2::RootTables roots() const { return ntt2::RootTables{d_rlo, d_rhi}; }
};

struct p2b_tree {
  p2b_ctx* ctx = nullptr;
  size_t n_leaves = 0;
  uint32_t log_leaves = 0;
  uint32_t cap_height = 0;
  size_t leaf_len = 0;
  uint64_t* d_levels = nullptr;  // 2*n_leaves - 2^cap_height digests, leaf level first
  // leaves: either column-major (owned by a batch; col stride = n_leaves) or row-major (owned here)
  const uint64_t* d_leaves_cm = nullptr;
  uint64_t* d_leaves_rm = nullptr;
  bool owned_by_batch = false;
};

static uint64_t next_object_id() {
  static std::atomic<uint64_t> n{1};
  return n.fetch_add(1);
}

struct p2b_batch {
  p2b_ctx* ctx = nullptr;
  uint64_t id = next_object_id();  // never reused: prove plans are keyed by it
  bool view = false;               // a read-only view of another context's batch (p2b_batch_attach): owns nothing
  size_t n_cols = 0;
  uint32_t log_n = 0, rate_bits = 0, cap_height = 0;
  uint64_t* d_coeffs = nullptr;  // n_cols x n
  uint64_t* d_lde = nullptr;     // n_cols x (n << rate_bits), leaf order
  uint64_t* d_values = nullptr;  // n_cols x n values on H (P2B_KEEP_VALUES)
  p2b_tree tree;
};

struct p2b_circuit {
  p2b_ctx* ctx = nullptr;
  uint64_t id = next_object_id();
  p2b_circuit_desc d{};
  plonk::Gate* d_gates = nullptr;
  uint64_t* d_k_is = nullptr;
  uint64_t* d_zh = nullptr;  // ZeroPolyOnCoset: Z_H on the 2^mdb cosets of the quotient LDE, then the inverses
  uint32_t mdb = 0;          // log2(quotient_degree_factor)
  std::vector<uint32_t> gate_kinds;  // host copy of gates[g].kind
  // gate indices: the light gates evaluated directly, the heavy ones, then the EXTENDED gates (low constraint degree:
  // evaluated on a sub-coset and extended by NTT, plonk::k_quotient_gates) of class 1 (2 n points) and class 2 (4 n);
  // behind them the part list of k_quotient_combine (0 = permutation argument, 1 + g for every directly evaluated gate)
  uint32_t* d_gate_list = nullptr;
  uint32_t n_light = 0, n_heavy = 0, n_ext[3] = {0, 0, 0} /* index = log2(D) */, n_parts_direct = 0;
};

struct p2b_challenger {
  p2b_ctx* ctx = nullptr;
  uint64_t* d_state = nullptr;  // frik::ChallengerState (30 u64)
};

// ------------------------------------------------------------------------------------------------
static int fail(p2b_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) {
    ctx->err = buf;
    if (code == P2B_ERR_CUDA) ctx->poisoned = true;
  } else {
    g_init_error = buf;
  }
  return code;
}

static void plans_clear(p2b_ctx* ctx);
static void plans_forget(p2b_ctx* ctx, uint64_t circuit_id, uint64_t batch_id);

#define CU(ctx, call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? P2B_ERR_OOM : P2B_ERR_CUDA, "%s: %s (%s:%d)", #call, \
                  cudaGetErrorString(e__), __FILE__, __LINE__);                                    \
  } while (0)

#define CHECK_CTX(ctx)                                                        \
  do {                                                                        \
    if (!(ctx)) return P2B_ERR_INVALID;                                       \
    if ((ctx)->poisoned) return fail(ctx, P2B_ERR_CUDA, "context poisoned by an earlier CUDA error: %s", (ctx)->err.c_str()); \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                           \
    if (e__ != cudaSuccess) return fail(ctx, P2B_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e__)); \
  } while (0)

#define LAUNCH_CHECK(ctx)                                                                   \
  do {                                                                                      \
    (ctx)->launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess)                                                                 \
      return fail(ctx, P2B_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static inline unsigned cdiv(size_t a, size_t b) { return (unsigned)((a + b - 1) / b); }

static int dmalloc(p2b_ctx* ctx, uint64_t** p, size_t n_u64) {
  *p = nullptr;
  if (n_u64 == 0) n_u64 = 1;
  if (ctx->cap) {
    // under stream capture the allocation becomes a memory node of the graph (fixed address, owned by the graph)
    CU(ctx, cudaMallocAsync((void**)p, n_u64 * sizeof(uint64_t), ctx->stream));
    return P2B_OK;
  }
  CU(ctx, cudaMallocFromPoolAsync((void**)p, n_u64 * sizeof(uint64_t), ctx->pool, ctx->stream));
  return P2B_OK;
}
static void dfree(p2b_ctx* ctx, void* p) {
  if (p) cudaFreeAsync(p, ctx->stream);
}
// tables that outlive the call that builds them (caches): never from a plan's arena
static int dmalloc_persistent(p2b_ctx* ctx, uint64_t** p, size_t n_u64, bool persistent) {
  if (persistent && ctx->cap)
    return fail(ctx, P2B_ERR_UNSUPPORTED, "a table cache is cold during stream capture (the first proof of a shape runs eagerly to warm it)");
  return dmalloc(ctx, p, n_u64);
}
static bool trace_enabled() {
  static const bool v = getenv("P2B_TRACE") != nullptr;
  return v;
}

// ---- stage timing -----------------------------------------------------------------------------
enum { ST_H2D = 0, ST_INTT, ST_LDE, ST_LEAF, ST_TREE, ST_FRI, ST_TRANSCRIPT, ST_OTHER };
static cudaEvent_t prof_event(p2b_ctx* ctx) {
  if (!ctx->ev_pool.empty()) {
    cudaEvent_t e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
static void stage_end(p2b_ctx* ctx) {
  if (!ctx->profiling || ctx->cur_stage < 0) return;
  cudaEvent_t b = prof_event(ctx);
  cudaEventRecord(b, ctx->stream);
  ctx->recs.push_back({ctx->cur_stage, ctx->cur_ev, b, ctx->launches - ctx->cur_launch0});
  ctx->cur_stage = -1;
  ctx->cur_ev = nullptr;
}
static void stage_begin(p2b_ctx* ctx, int stage) {
  if (!ctx->profiling) return;
  stage_end(ctx);
  ctx->cur_stage = stage;
  ctx->cur_ev = prof_event(ctx);
  ctx->cur_launch0 = ctx->launches;
  cudaEventRecord(ctx->cur_ev, ctx->stream);
}

// every host wait on the context's stream goes through here
static cudaError_t ctx_sync(p2b_ctx* ctx) {
  if (!ctx->blocking_sync) return cudaStreamSynchronize(ctx->stream);
  if (!ctx->ev_block) {
    cudaError_t e = cudaEventCreateWithFlags(&ctx->ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(ctx->ev_block, ctx->stream);
  return e != cudaSuccess ? e : cudaEventSynchronize(ctx->ev_block);
}

extern "C" int p2b_profile_enable(p2b_ctx* ctx, int on) {
  CHECK_CTX(ctx);
  stage_end(ctx);
  ctx->profiling = on != 0;
  return P2B_OK;
}
extern "C" int p2b_profile_read(p2b_ctx* ctx, float* ms_out, uint64_t* count_out) {
  CHECK_CTX(ctx);
  if (!ms_out) return P2B_ERR_INVALID;
  stage_end(ctx);
  CU(ctx, ctx_sync(ctx));
  for (auto& r : ctx->recs) {
    float ms = 0;
    CU(ctx, cudaEventElapsedTime(&ms, r.a, r.b));
    ms_out[r.stage] += ms;
    if (count_out) count_out[r.stage] += r.launches;
    ctx->ev_pool.push_back(r.a);
    ctx->ev_pool.push_back(r.b);
  }
  ctx->recs.clear();
  return P2B_OK;
}

// ------------------------------------------------------------------------------------------------ context
extern "C" int p2b_version(void) { return P2B_VERSION; }

static int ctx_setup(p2b_ctx* ctx) {
  CU(ctx, cudaEventCreate(&ctx->ev0));
  CU(ctx, cudaEventCreate(&ctx->ev1));
  // keep freed blocks cached in the stream-ordered pool (no trimming at sync points)
  cudaMemPoolProps props{};
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = ctx->device;
  CU(ctx, cudaMemPoolCreate(&ctx->pool, &props));
  uint64_t thresh = UINT64_MAX;
  CU(ctx, cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  CU(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  CU(ctx, cudaFuncSetAttribute(nttk::k_ntt_cols, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  CU(ctx, cudaFuncSetAttribute(nttk::k_ntt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  int rc;
  if ((rc = dmalloc(ctx, &ctx->d_w12, 2048))) return rc;
  if ((rc = dmalloc(ctx, &ctx->d_rlo, 65536))) return rc;
  if ((rc = dmalloc(ctx, &ctx->d_rhi, 65536))) return rc;
  // G = 7^((p-1)/2^32): primitive 2^32-th root (plonky2 POWER_OF_TWO_GENERATOR); w_4096 = G^(2^20)
  const uint64_t G = 1753635133440165772ull;
  uint64_t g16 = G, w4096 = G;
  auto mulmod = [](uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_P); };
  for (int i = 0; i < 16; i++) g16 = mulmod(g16, g16);
  for (int i = 0; i < 20; i++) w4096 = mulmod(w4096, w4096);
  nttk::k_pow_table<<<cdiv(65536, 256), 256, 0, ctx->stream>>>(G, 65536, ctx->d_rlo);
  LAUNCH_CHECK(ctx);
  nttk::k_pow_table<<<cdiv(65536, 256), 256, 0, ctx->stream>>>(g16, 65536, ctx->d_rhi);
  LAUNCH_CHECK(ctx);
  nttk::k_pow_table<<<cdiv(2048, 256), 256, 0, ctx->stream>>>(w4096, 2048, ctx->d_w12);
  LAUNCH_CHECK(ctx);
  if ((rc = dmalloc(ctx, &ctx->d_t1, 4096))) return rc;
  if ((rc = dmalloc(ctx, &ctx->d_t2, 256))) return rc;
  ntt2::k_build_row_tables<<<16, 256, 0, ctx->stream>>>(ctx->roots(), ctx->d_t1, ctx->d_t2);
  LAUNCH_CHECK(ctx);
  if ((rc = dmalloc(ctx, &ctx->d_scratch, 4))) return rc;
  CU(ctx, cudaMemsetAsync(ctx->d_scratch, 0, 4 * sizeof(uint64_t), ctx->stream));
  ctx->h_stage_bytes = 1 << 20;
  CU(ctx, cudaMallocHost((void**)&ctx->h_stage, ctx->h_stage_bytes));
  CU(ctx, ctx_sync(ctx));
  ctx->launches = 0;
  return P2B_OK;
}

static int init_common(int device, void* stream, bool borrow, p2b_ctx** out) {
  if (!out) return P2B_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, P2B_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, P2B_ERR_INVALID, "device %d out of range (0..%d)", device, n - 1);
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, P2B_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  p2b_ctx* ctx = new (std::nothrow) p2b_ctx();
  if (!ctx) return fail(nullptr, P2B_ERR_OOM, "host allocation failed");
  ctx->device = device;
  if (const char* m = getenv("P2B_SYNC")) ctx->blocking_sync = m[0] == 'b' || m[0] == 'B';  // P2B_SYNC=block | spin
  if (const char* m = getenv("P2B_MODE")) ctx->latency_mode = m[0] == 'l' || m[0] == 'L';    // P2B_MODE=latency | throughput
  if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->sm_count <= 0)
    ctx->sm_count = 148;
  if (borrow) {
    ctx->stream = (cudaStream_t)stream;
  } else {
    e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete ctx;
      return fail(nullptr, P2B_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
    ctx->own_stream = true;
  }
  int rc = ctx_setup(ctx);
  if (rc != P2B_OK) {
    g_init_error = ctx->err;
    p2b_destroy(ctx);
    return rc;
  }
  *out = ctx;
  return P2B_OK;
}

extern "C" int p2b_init(int device, p2b_ctx** out) { return init_common(device, nullptr, false, out); }
extern "C" int p2b_init_on_stream(int device, void* cuda_stream, p2b_ctx** out) {
  return init_common(device, cuda_stream, true, out);
}

extern "C" void p2b_destroy(p2b_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  ctx_sync(ctx);
  plans_clear(ctx);
  if (ctx->h_proof) cudaFreeHost(ctx->h_proof);
  if (ctx->ev_h2d) cudaEventDestroy(ctx->ev_h2d);
  dfree(ctx, ctx->d_w12);
  dfree(ctx, ctx->d_rlo);
  dfree(ctx, ctx->d_rhi);
  dfree(ctx, ctx->d_t1);
  dfree(ctx, ctx->d_t2);
  for (auto& kv : ctx->tw_cache) dfree(ctx, kv.second);
  for (auto& kv : ctx->cp_cache) dfree(ctx, kv.second);
  for (auto& kv : ctx->small_cp_cache) {
    dfree(ctx, kv.second.lo);
    dfree(ctx, kv.second.hi);
  }
  dfree(ctx, ctx->d_scratch);
  if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
  for (int i = 0; i < 2; i++) {
    if (ctx->h_upload[i]) cudaFreeHost(ctx->h_upload[i]);
    if (ctx->upload_done[i]) cudaEventDestroy(ctx->upload_done[i]);
  }
  ctx_sync(ctx);
  for (auto& r : ctx->recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  if (ctx->cur_ev) cudaEventDestroy(ctx->cur_ev);
  if (ctx->ev_block) cudaEventDestroy(ctx->ev_block);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->copy_events) cudaEventDestroy(e);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* p2b_last_error(const p2b_ctx* ctx) { return ctx ? ctx->err.c_str() : g_init_error.c_str(); }

extern "C" int p2b_set_blocking_sync(p2b_ctx* ctx, int on) {
  CHECK_CTX(ctx);
  ctx->blocking_sync = on != 0;
  return P2B_OK;
}

// Throughput mode (default) sizes every launch for many proofs in flight on the GPU: what counts is the instructions a
// proof executes and the registers its CTAs hold.  Latency mode is for a worker that has the GPU to itself (the reference's
// one-job-at-a-time worker with no sibling processes on the device): dependent launch chains are shortened at the price of
// work — the Merkle trees are climbed by the fused subtree kernel from 2^15 digests up (12 launches less per 2^12-row
// proof; with 24 proofs in flight that costs 9 % throughput, profiles/r02_tree_fuse_sweep.txt) and the proof-of-work
// search runs one CTA per SM (one stride covers the mean of the 16-bit search; ~2x the necessary permutations).  Measured
// on one context at the City shape: 3.71 -> 3.35 ms per proof (269 -> 298 proofs/s).  Results are identical in both modes.
extern "C" int p2b_set_latency_mode(p2b_ctx* ctx, int on) {
  CHECK_CTX(ctx);
  if (ctx->pending_words) return fail(ctx, P2B_ERR_INVALID, "a submitted proof has not been collected yet");
  if (ctx->latency_mode != (on != 0)) {
    CU(ctx, ctx_sync(ctx));
    plans_clear(ctx);  // captured prove plans hold the launch configuration of the other mode
    ctx->latency_mode = on != 0;
  }
  return P2B_OK;
}

extern "C" int p2b_synchronize(p2b_ctx* ctx) {
  CHECK_CTX(ctx);
  CU(ctx, ctx_sync(ctx));
  return P2B_OK;
}

extern "C" int p2b_host_alloc(p2b_ctx* ctx, size_t bytes, void** out) {
  CHECK_CTX(ctx);
  if (!out) return P2B_ERR_INVALID;
  CU(ctx, cudaMallocHost(out, bytes ? bytes : 1));
  return P2B_OK;
}
extern "C" int p2b_host_free(p2b_ctx* ctx, void* p) {
  CHECK_CTX(ctx);
  if (p) CU(ctx, cudaFreeHost(p));
  return P2B_OK;
}
extern "C" uint64_t p2b_launch_count(const p2b_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int p2b_timer_start(p2b_ctx* ctx) {
  CHECK_CTX(ctx);
  CU(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  return P2B_OK;
}
extern "C" int p2b_timer_stop_ms(p2b_ctx* ctx, float* ms_out) {
  CHECK_CTX(ctx);
  if (!ms_out) return P2B_ERR_INVALID;
  CU(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  CU(ctx, cudaEventSynchronize(ctx->ev1));
  CU(ctx, cudaEventElapsedTime(ms_out, ctx->ev0, ctx->ev1));
  return P2B_OK;
}

extern "C" int p2b_timer_span_ms(p2b_ctx* first, p2b_ctx* last, float* ms_out) {
  CHECK_CTX(last);
  if (!first || !ms_out || first->device != last->device) return fail(last, P2B_ERR_INVALID, "contexts of one device expected");
  CU(last, cudaEventSynchronize(last->ev1));
  CU(last, cudaEventElapsedTime(ms_out, first->ev0, last->ev1));
  return P2B_OK;
}

// copy `n` u64 from device to a caller (pageable or pinned) buffer, synchronously
// Device -> caller's buffer.  The caller's memory is normally pageable (a Rust Vec, a numpy array), and a copy into
// pageable memory makes the driver spin inside cudaMemcpyAsync until everything queued before it has run — for
// p2b_prove that is the whole proof.  Small results (caps, challenges, a 130 KB proof) therefore land in the
// context's pinned staging buffer first, the wait goes through ctx_sync (which can sleep), then one memcpy.
static int d2h(p2b_ctx* ctx, uint64_t* dst, const uint64_t* d_src, size_t n) {
  const size_t bytes = n * sizeof(uint64_t);
  if (ctx->h_stage && bytes <= ctx->h_stage_bytes) {
    CU(ctx, cudaMemcpyAsync(ctx->h_stage, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, ctx_sync(ctx));
    memcpy(dst, ctx->h_stage, bytes);
    return P2B_OK;
  }
  CU(ctx, cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, ctx_sync(ctx));
  return P2B_OK;
}

// ------------------------------------------------------------------------------------------------ NTT driver
struct NttPlan {
  uint32_t lr, ls;
};
static NttPlan plan_for(uint32_t log_n) {
  NttPlan p;
  if (log_n <= 12) {
    p.lr = 0;
    p.ls = log_n;
  } else {
    p.ls = (log_n + 1) / 2;
    p.lr = log_n - p.ls;
  }
  return p;
}
static inline uint32_t u32min(uint32_t a, uint32_t b) { return a < b ? a : b; }


// ---- radix-16 path: 2^12 <= n <= 2^20 ------------------------------------------------------------
static bool ntt2_supported(uint32_t log_n) { return log_n >= 12 && log_n <= 20; }

static int ntt2_get_tw(p2b_ctx* ctx, uint32_t log_m, uint32_t d, const uint64_t** out) {
  uint64_t key = ((uint64_t)log_m << 8) | d;
  auto it = ctx->tw_cache.find(key);
  if (it == ctx->tw_cache.end()) {
    uint64_t* p = nullptr;
    int rc;
    if ((rc = dmalloc_persistent(ctx, &p, (size_t)1 << log_m, true))) return rc;
    ntt2::k_build_tw<<<cdiv((size_t)1 << log_m, 256), 256, 0, ctx->stream>>>(ctx->roots(), log_m, d, p);
    LAUNCH_CHECK(ctx);
    it = ctx->tw_cache.emplace(key, p).first;
  }
  *out = it->second;
  return P2B_OK;
}

static uint64_t h_mulmod(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_P); }
static uint64_t h_powmod(uint64_t a, uint64_t e) {
  uint64_t r = 1;
  a %= GL_P;
  while (e) {
    if (e & 1) r = h_mulmod(r, a);
    a = h_mulmod(a, a);
    e >>= 1;
  }
  return r;
}

// [2^rate_bits][n] powers of s_t = shift * w_{n << rate_bits}^t
static int ntt2_get_cp(p2b_ctx* ctx, uint64_t shift, uint32_t log_n, uint32_t rate_bits, const uint64_t** out) {
  p2b_ctx::CpKey key{shift % GL_P, log_n, rate_bits};
  auto it = ctx->cp_cache.find(key);
  if (it == ctx->cp_cache.end()) {
    const size_t n = (size_t)1 << log_n, bytes = (n << rate_bits) * 8;
    // bounded cache: FRI layers of different proofs reuse the same few shifts; drop everything if a caller
    // cycles through many distinct domains
    if (ctx->cap)
      return fail(ctx, P2B_ERR_UNSUPPORTED, "a table cache is cold during stream capture (the first proof of a shape runs eagerly to warm it)");
    if (ctx->cp_cache_bytes + bytes > ((size_t)4 << 30)) {
      plans_clear(ctx);  // captured graphs hold pointers into the tables that are about to go
      for (auto& kv : ctx->cp_cache) dfree(ctx, kv.second);
      ctx->cp_cache.clear();
      ctx->cp_cache_bytes = 0;
    }
    uint64_t* p = nullptr;
    int rc;
    if ((rc = dmalloc(ctx, &p, n << rate_bits))) return rc;
    const uint64_t G = 1753635133440165772ull;
    uint64_t wN = h_powmod(G, (uint64_t)1 << (32 - (log_n + rate_bits)));
    for (uint32_t t = 0; t < (1u << rate_bits); t++) {
      uint64_t st = h_mulmod(key.shift, h_powmod(wN, t));
      nttk::k_pow_table<<<cdiv(n, 256), 256, 0, ctx->stream>>>(st, n, p + (size_t)t * n);
      LAUNCH_CHECK(ctx);
    }
    it = ctx->cp_cache.emplace(key, p).first;
    ctx->cp_cache_bytes += bytes;
  }
  *out = it->second;
  return P2B_OK;
}

template <bool PRESCALE>
static int ntt2_launch_strided(p2b_ctx* ctx, uint32_t d, const ntt2::StridedParams& sp, size_t items, size_t n_cols) {
  dim3 grid(cdiv(items, 256), (unsigned)n_cols, 1);
  switch (d) {
    case 1: ntt2::k_strided<1, PRESCALE><<<grid, 256, 0, ctx->stream>>>(sp); break;
    case 2: ntt2::k_strided<2, PRESCALE><<<grid, 256, 0, ctx->stream>>>(sp); break;
    case 3: ntt2::k_strided<3, PRESCALE><<<grid, 256, 0, ctx->stream>>>(sp); break;
    default: ntt2::k_strided<4, PRESCALE><<<grid, 256, 0, ctx->stream>>>(sp); break;
  }
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

// Strided radix steps that bring every size-n transform of `buf` (n_cols columns of `total` contiguous
// elements, total a multiple of n) down to independent 4096-point rows.  The first step reads `src`.
// With cp != nullptr the first step is the LDE prescale step (src = coefficients, total = n << log_cosets).
static int ntt2_strided_steps(p2b_ctx* ctx, const uint64_t* src, size_t src_stride, uint64_t* buf, size_t buf_stride,
                              size_t n_cols, uint32_t log_n, size_t total, const uint64_t* cp, uint32_t log_cosets) {
  uint32_t log_m = log_n;
  bool first = true;
  int rc;
  while (log_m > 12) {
    uint32_t d = log_m - 12 >= 4 ? 4 : log_m - 12;
    ntt2::StridedParams sp{};
    if ((rc = ntt2_get_tw(ctx, log_m, d, &sp.tw))) return rc;
    sp.in = first ? src : buf;
    sp.in_col_stride = first ? src_stride : buf_stride;
    sp.out = buf;
    sp.out_col_stride = buf_stride;
    sp.out_coset_stride = (size_t)1 << log_n;
    sp.log_m = log_m;
    sp.log_n = log_n;
    sp.log_cosets = log_cosets;
    sp.cp = cp;
    if (first && cp) {
      if ((rc = ntt2_launch_strided<true>(ctx, d, sp, ((size_t)1 << log_n) >> d, n_cols))) return rc;
    } else {
      if ((rc = ntt2_launch_strided<false>(ctx, d, sp, (first ? (size_t)1 << log_n : total) >> d, n_cols))) return rc;
    }
    first = false;
    log_m -= d;
  }
  return P2B_OK;
}

static int run_intt2(p2b_ctx* ctx, const uint64_t* d_vals, uint64_t* d_coeffs, uint64_t* d_tmp, size_t n_cols,
                     uint32_t log_n, size_t tmp_poly_stride) {
  const size_t n = (size_t)1 << log_n;
  int rc;
  ntt2::RowParams rp{};
  rp.t1 = ctx->d_t1;
  rp.t2 = ctx->d_t2;
  rp.log_R = log_n - 12;
  rp.scale = GL_P - ((GL_P - 1) >> log_n);  // 1/n mod p
  rp.out = d_coeffs;
  rp.out_col_stride = n;
  if (log_n == 12) {
    rp.in = d_vals;
    rp.in_col_stride = n;
  } else {
    if ((rc = ntt2_strided_steps(ctx, d_vals, n, d_tmp, tmp_poly_stride, n_cols, log_n, n, nullptr, 0))) return rc;
    rp.in = d_tmp;
    rp.in_col_stride = tmp_poly_stride;
  }
  ntt2::k_row4096<2, false><<<dim3((unsigned)(n >> 12), (unsigned)n_cols, 1), 256, 0, ctx->stream>>>(rp);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

static int run_lde2(p2b_ctx* ctx, const uint64_t* d_coeffs, size_t in_stride, uint64_t* d_lde, size_t n_cols,
                    uint32_t log_n, uint32_t rate_bits, uint64_t shift) {
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  int rc;
  const uint64_t* cp = nullptr;
  if ((rc = ntt2_get_cp(ctx, shift, log_n, rate_bits, &cp))) return rc;
  ntt2::RowParams rp{};
  rp.t1 = ctx->d_t1;
  rp.t2 = ctx->d_t2;
  rp.out = d_lde;
  rp.out_col_stride = N;
  rp.out_coset_stride = n;
  if (log_n == 12) {
    rp.in = d_coeffs;
    rp.in_col_stride = in_stride;
    rp.cp = cp;
    rp.log_cosets = rate_bits;
    ntt2::k_row4096<0, true><<<dim3(1, (unsigned)n_cols, 1), 256, 0, ctx->stream>>>(rp);
    LAUNCH_CHECK(ctx);
    return P2B_OK;
  }
  if ((rc = ntt2_strided_steps(ctx, d_coeffs, in_stride, d_lde, N, n_cols, log_n, N, cp, rate_bits))) return rc;
  rp.in = d_lde;
  rp.in_col_stride = N;
  ntt2::k_row4096<0, false><<<dim3((unsigned)(N >> 12), (unsigned)n_cols, 1), 256, 0, ctx->stream>>>(rp);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

// values (n_cols x n, natural) -> coefficients (n_cols x n, natural).  d_tmp: n_cols x n scratch
// (only used when n > 2^12); may alias neither input nor output.
static int run_intt(p2b_ctx* ctx, const uint64_t* d_vals, uint64_t* d_coeffs, uint64_t* d_tmp, size_t n_cols,
                    uint32_t log_n, size_t tmp_poly_stride) {
  if (ntt2_supported(log_n)) return run_intt2(ctx, d_vals, d_coeffs, d_tmp, n_cols, log_n, tmp_poly_stride);
  const size_t n = (size_t)1 << log_n;
  NttPlan pl = plan_for(log_n);
  // 1/n mod p = p - (p-1)/n
  const uint64_t ninv = GL_P - ((GL_P - 1) >> log_n);
  const uint64_t scale = log_n == 0 ? 1 : ninv;
  nttk::RowsParams rp{};
  rp.log_n = log_n;
  rp.lr = pl.lr;
  rp.ls = pl.ls;
  rp.log_cosets = 0;
  rp.mode = 2;
  rp.prescale = 0;
  rp.scale = scale;
  rp.tb = ctx->tables();
  rp.out = d_coeffs;
  rp.out_poly_stride = n;
  rp.out_coset_stride = 0;
  if (pl.lr == 0) {
    rp.in = d_vals;
    rp.in_poly_stride = n;
    rp.log_TR = 0;
    size_t smem = n * 8;
    nttk::k_ntt_rows<<<dim3(1, (unsigned)n_cols, 1), 256, smem, ctx->stream>>>(rp);
    LAUNCH_CHECK(ctx);
    return P2B_OK;
  }
  nttk::ColsParams cp{};
  cp.in = d_vals;
  cp.in_poly_stride = n;
  cp.out = d_tmp;
  cp.out_poly_stride = tmp_poly_stride;
  cp.out_coset_stride = 0;
  cp.log_n = log_n;
  cp.lr = pl.lr;
  cp.ls = pl.ls;
  cp.log_T = u32min(pl.ls, (pl.lr >= 9 ? 13 : 12) - pl.lr);
  cp.log_cosets = 0;
  cp.natural_rows = 1;
  cp.prescale = 0;
  cp.tb = ctx->tables();
  {
    size_t smem = ((size_t)8 << pl.lr) << cp.log_T;
    dim3 grid((unsigned)(((size_t)1 << pl.ls) >> cp.log_T), (unsigned)n_cols, 1);
    nttk::k_ntt_cols<<<grid, 256, smem, ctx->stream>>>(cp);
    LAUNCH_CHECK(ctx);
  }
  rp.in = d_tmp;
  rp.in_poly_stride = tmp_poly_stride;
  rp.log_TR = u32min(pl.lr, (pl.ls >= 10 ? 13 : 12) - pl.ls);
  {
    size_t S = (size_t)1 << pl.ls, TR = (size_t)1 << rp.log_TR;
    size_t smem = TR * (S + (TR > 1 ? 1 : 0)) * 8;
    dim3 grid((unsigned)(((size_t)1 << pl.lr) >> rp.log_TR), (unsigned)n_cols, 1);
    nttk::k_ntt_rows<<<grid, 256, smem, ctx->stream>>>(rp);
    LAUNCH_CHECK(ctx);
  }
  return P2B_OK;
}

// coefficients (n_cols x n, natural, stride in_stride) -> coset evaluations on shift * w_{n<<rate_bits}^t * <w_n>,
// all 2^rate_bits cosets, written in leaf (bit-reversed) order: d_lde is n_cols x (n << rate_bits).
// `shift` is the coset shift of the whole domain (7 for PolynomialBatch; 7^(arity^l) for FRI layers).
static int run_lde(p2b_ctx* ctx, const uint64_t* d_coeffs, size_t in_stride, uint64_t* d_lde, size_t n_cols,
                   uint32_t log_n, uint32_t rate_bits, uint64_t shift) {
  if (ntt2_supported(log_n)) return run_lde2(ctx, d_coeffs, in_stride, d_lde, n_cols, log_n, rate_bits, shift);
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  const uint32_t n_cosets = 1u << rate_bits;
  NttPlan pl = plan_for(log_n);
  // power tables of s_t = shift * w_N^t
  nttk::CosetPow cpw{};
  uint64_t *d_lo = nullptr, *d_hi = nullptr;
  uint32_t hi_count = (uint32_t)(n > 4096 ? n >> 12 : 1);
  int rc;
  // small transforms (the FRI layers of every proof) keep their tables for the life of the context
  const bool cache_tables = log_n < 12;
  p2b_ctx::CpKey ckey{shift % GL_P, log_n, rate_bits};
  auto cit = cache_tables ? ctx->small_cp_cache.find(ckey) : ctx->small_cp_cache.end();
  const bool have_tables = cit != ctx->small_cp_cache.end();
  if (have_tables) {
    d_lo = cit->second.lo;
    d_hi = cit->second.hi;
  } else {
    if ((rc = dmalloc_persistent(ctx, &d_lo, (size_t)n_cosets * 4096, cache_tables))) return rc;
    if ((rc = dmalloc_persistent(ctx, &d_hi, (size_t)n_cosets * hi_count, cache_tables))) return rc;
  }
  auto mulmod = [](uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) % GL_P); };
  auto powmod = [&](uint64_t a, uint64_t e) {
    uint64_t r = 1;
    a %= GL_P;
    while (e) {
      if (e & 1) r = mulmod(r, a);
      a = mulmod(a, a);
      e >>= 1;
    }
    return r;
  };
  const uint64_t G = 1753635133440165772ull;
  const uint32_t log_N = log_n + rate_bits;
  uint64_t wN = powmod(G, (uint64_t)1 << (32 - log_N));
  for (uint32_t t = 0; t < n_cosets && !have_tables; t++) {
    uint64_t st = mulmod(shift % GL_P, powmod(wN, t));
    nttk::k_pow_table<<<cdiv(4096, 256), 256, 0, ctx->stream>>>(st, 4096, d_lo + (size_t)t * 4096);
    LAUNCH_CHECK(ctx);
    nttk::k_pow_table<<<cdiv(hi_count, 256), 256, 0, ctx->stream>>>(powmod(st, 4096), hi_count,
                                                                    d_hi + (size_t)t * hi_count);
    LAUNCH_CHECK(ctx);
  }
  if (cache_tables && !have_tables) ctx->small_cp_cache.emplace(ckey, p2b_ctx::SmallCp{d_lo, d_hi, hi_count});
  cpw.lo = d_lo;
  cpw.hi = d_hi;
  cpw.hi_count = hi_count;

  if (pl.lr == 0) {
    nttk::RowsParams rp{};
    rp.in = d_coeffs;
    rp.in_poly_stride = in_stride;
    rp.out = d_lde;
    rp.out_poly_stride = N;
    rp.out_coset_stride = n;
    rp.log_n = log_n;
    rp.lr = 0;
    rp.ls = log_n;
    rp.log_TR = 0;
    rp.log_cosets = rate_bits;
    rp.mode = 0;
    rp.prescale = 1;
    rp.scale = 1;
    rp.cp = cpw;
    rp.tb = ctx->tables();
    nttk::k_ntt_rows<<<dim3(1, (unsigned)n_cols, n_cosets), 256, n * 8, ctx->stream>>>(rp);
    LAUNCH_CHECK(ctx);
  } else {
    nttk::ColsParams cp{};
    cp.in = d_coeffs;
    cp.in_poly_stride = in_stride;
    cp.out = d_lde;
    cp.out_poly_stride = N;
    cp.out_coset_stride = n;
    cp.log_n = log_n;
    cp.lr = pl.lr;
    cp.ls = pl.ls;
    cp.log_T = u32min(pl.ls, (pl.lr >= 9 ? 13 : 12) - pl.lr);
    cp.log_cosets = rate_bits;
    cp.natural_rows = 0;
    cp.prescale = 1;
    cp.cp = cpw;
    cp.tb = ctx->tables();
    {
      size_t smem = ((size_t)8 << pl.lr) << cp.log_T;
      dim3 grid((unsigned)(((size_t)1 << pl.ls) >> cp.log_T), (unsigned)n_cols, n_cosets);
      nttk::k_ntt_cols<<<grid, 256, smem, ctx->stream>>>(cp);
      LAUNCH_CHECK(ctx);
    }
    // pass B in place over all N/S rows of every column
    nttk::RowsParams rp{};
    rp.in = d_lde;
    rp.in_poly_stride = N;
    rp.out = d_lde;
    rp.out_poly_stride = N;
    rp.out_coset_stride = 0;
    rp.log_n = log_n;
    rp.lr = pl.lr;
    rp.ls = pl.ls;
    uint32_t log_rows = pl.lr + rate_bits;
    rp.log_TR = u32min(log_rows, 12 - pl.ls);
    rp.log_cosets = 0;
    rp.mode = 0;
    rp.prescale = 0;
    rp.scale = 1;
    rp.tb = ctx->tables();
    size_t S = (size_t)1 << pl.ls, TR = (size_t)1 << rp.log_TR;
    size_t smem = TR * (S + (TR > 1 ? 1 : 0)) * 8;
    dim3 grid((unsigned)(((size_t)1 << log_rows) >> rp.log_TR), (unsigned)n_cols, 1);
    nttk::k_ntt_rows<<<grid, 256, smem, ctx->stream>>>(rp);
    LAUNCH_CHECK(ctx);
  }
  if (!cache_tables) {
    dfree(ctx, d_lo);
    dfree(ctx, d_hi);
  }
  return P2B_OK;
}

// ------------------------------------------------------------------------------------------------ Merkle
static size_t levels_len(size_t n_leaves, uint32_t cap_height) { return 2 * n_leaves - ((size_t)1 << cap_height); }
static size_t level_off(size_t n_leaves, uint32_t i) { return 2 * n_leaves - 2 * (n_leaves >> i); }

// leaf digests must already be in t->d_levels[0 .. n_leaves)
// widest level handled with 16 lanes per node; override with P2B_COOP_MAX_NODES for tuning
static size_t coop_max_nodes() {
  static const size_t v = [] {
    const char* e = getenv("P2B_COOP_MAX_NODES");
    return e ? (size_t)strtoull(e, nullptr, 10) : (size_t)2048;
  }();
  return v;
}
#define COOP_MAX_NODES coop_max_nodes()
// P2B_FUSED_TREE=0: one launch per level (the round-1 path), for A/B measurements
static bool fused_tree() {
  static const bool v = [] {
    const char* e = getenv("P2B_FUSED_TREE");
    return !e || e[0] != '0';
  }();
  return v;
}
// threads per CTA of the one-permutation-per-thread hash kernels (P2B_HASH_BLOCK: 64 / 128 / 256, tuning)
static unsigned hash_block() {
  static const unsigned v = [] {
    const char* e = getenv("P2B_HASH_BLOCK");
    const int b = e ? atoi(e) : 256;
    return (unsigned)(b == 64 || b == 128 ? b : 256);
  }();
  return v;
}
static int build_levels(p2b_ctx* ctx, p2b_tree* t) {
  uint32_t L = t->log_leaves - t->cap_height;
  if (fused_tree() && L > 0) {
    // wide levels at full throughput (one permutation per thread, the whole GPU), then everything from a level of at
    // most 2^fuse_log digests up to the cap in ONE launch (fusedk::k_tree_subtree)
    // P2B_TREE_FUSE_LOG: widest level (log2 digests) handed to the fused kernel.  A subtree CTA lives for all its
    // dependent levels (23 us each) with ever fewer busy warps while it holds 40k registers: an SM that hosts one cannot
    // host a leaf-hash CTA.  With many proofs in flight that costs throughput — measured at the City shape, 24 workers
    // (profiles/r02_tree_fuse_sweep.txt): fused from 2^15 digests 891 proofs/s, 2^13 952, 2^11 969, 2^9 969 — and one proof
    // alone loses little (271 -> 258 proofs/s for 12 more launches).  Default 2^11: one launch per level at full
    // throughput down to 2048 digests, then four subtree CTAs.
    static const uint32_t fuse_log = [] {
      const char* e = getenv("P2B_TREE_FUSE_LOG");
      const int v = e ? atoi(e) : 11;
      return (uint32_t)(v < 1 ? 1 : v > 24 ? 24 : v);
    }();
    const uint32_t fuse_log_eff = ctx->latency_mode && fuse_log < 15 ? 15 : fuse_log;
    uint32_t i = 0;
    while (i < L && (t->n_leaves >> i) > ((size_t)1 << fuse_log_eff)) {
      const size_t n_par = t->n_leaves >> (i + 1);
      hashk::k_tree_level<<<cdiv(n_par, hash_block()), hash_block(), 0, ctx->stream>>>(t->d_levels + 4 * level_off(t->n_leaves, i),
                                                                      t->d_levels + 4 * level_off(t->n_leaves, i + 1), n_par);
      LAUNCH_CHECK(ctx);
      i++;
    }
    if (i < L) {
      const uint32_t log_first = t->log_leaves - i;
      const unsigned grid = log_first > (uint32_t)fusedk::SUB_LOG ? 1u << (log_first - fusedk::SUB_LOG) : 1u;
      fusedk::k_tree_subtree<<<grid, 256, 0, ctx->stream>>>(t->d_levels, t->n_leaves, i, log_first, L - i,
                                                           (unsigned int*)ctx->d_scratch);
      LAUNCH_CHECK(ctx);
    }
    return P2B_OK;
  }
  for (uint32_t i = 0; i < L; i++) {
    size_t n_par = t->n_leaves >> (i + 1);
    const uint64_t* child = t->d_levels + 4 * level_off(t->n_leaves, i);
    uint64_t* parent = t->d_levels + 4 * level_off(t->n_leaves, i + 1);
    // wide levels are throughput bound (one permutation per thread); the narrow top of the tree is a chain of
    // dependent levels, where one warp per node cuts the latency of each
    if (n_par > COOP_MAX_NODES)
      hashk::k_tree_level<<<cdiv(n_par, 256), 256, 0, ctx->stream>>>(child, parent, n_par);
    else
      hashk::k_tree_level_coop<<<cdiv(n_par * 32, 256), 256, 0, ctx->stream>>>(child, parent, n_par);
    LAUNCH_CHECK(ctx);
  }
  return P2B_OK;
}

static int tree_from_colmajor(p2b_ctx* ctx, p2b_tree* t, const uint64_t* d_data, size_t n_leaves, uint32_t log_leaves,
                              size_t n_cols, uint32_t cap_height) {
  t->ctx = ctx;
  t->n_leaves = n_leaves;
  t->log_leaves = log_leaves;
  t->cap_height = cap_height;
  t->leaf_len = n_cols;
  t->d_leaves_cm = d_data;
  int rc;
  if ((rc = dmalloc(ctx, &t->d_levels, 4 * levels_len(n_leaves, cap_height)))) return rc;
  stage_begin(ctx, ST_LEAF);
  hashk::k_leaf_hash_colmajor<<<cdiv(n_leaves, hash_block()), hash_block(), 0, ctx->stream>>>(d_data, n_leaves, (uint32_t)n_cols, n_leaves,
                                                                          t->d_levels);
  LAUNCH_CHECK(ctx);
  stage_begin(ctx, ST_TREE);
  rc = build_levels(ctx, t);
  stage_end(ctx);
  return rc;
}

// ------------------------------------------------------------------------------------------------ PolynomialBatch
static int check_batch_args(p2b_ctx* ctx, const void* cols, size_t n_cols, uint32_t log_n, uint32_t rate_bits,
                            uint32_t cap_height, uint32_t flags, p2b_batch** out) {
  if (!cols || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_cols == 0 || n_cols > 65535) return fail(ctx, P2B_ERR_INVALID, "n_cols must be in 1..65535 (got %zu)", n_cols);
  if (log_n > 24) return fail(ctx, P2B_ERR_UNSUPPORTED, "log_n %u > 24", log_n);
  if (rate_bits > 6) return fail(ctx, P2B_ERR_UNSUPPORTED, "rate_bits %u > 6", rate_bits);
  if (cap_height > log_n + rate_bits)
    return fail(ctx, P2B_ERR_INVALID, "cap_height %u exceeds log2(#leaves) = %u", cap_height, log_n + rate_bits);
  if (flags & ~P2B_KEEP_VALUES)
    return fail(ctx, P2B_ERR_UNSUPPORTED, "blinding (salted) batches are not supported (flags=%u)", flags);
  return P2B_OK;
}

static int batch_build(p2b_ctx* ctx, uint64_t* d_in /* owned, n_cols x n */, bool is_values, size_t n_cols,
                       uint32_t log_n, uint32_t rate_bits, uint32_t cap_height, p2b_batch** out,
                       bool keep_values = false) {
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  p2b_batch* b = new (std::nothrow) p2b_batch();
  if (!b) {
    dfree(ctx, d_in);
    return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  }
  b->ctx = ctx;
  b->n_cols = n_cols;
  b->log_n = log_n;
  b->rate_bits = rate_bits;
  b->cap_height = cap_height;
  b->tree.owned_by_batch = true;
  int rc = dmalloc(ctx, &b->d_lde, n_cols * N);
  if (rc == P2B_OK) {
    if (is_values) {
      rc = dmalloc(ctx, &b->d_coeffs, n_cols * n);
      // scratch for the two-pass inverse transform: the first n words of every LDE column
      stage_begin(ctx, ST_INTT);
      if (rc == P2B_OK) rc = run_intt(ctx, d_in, b->d_coeffs, b->d_lde, n_cols, log_n, N);
      stage_end(ctx);
      if (keep_values)
        b->d_values = d_in;
      else
        dfree(ctx, d_in);
    } else {
      b->d_coeffs = d_in;
    }
  } else {
    dfree(ctx, d_in);
  }
  stage_begin(ctx, ST_LDE);
  if (rc == P2B_OK) rc = run_lde(ctx, b->d_coeffs, n, b->d_lde, n_cols, log_n, rate_bits, 7);
  stage_end(ctx);
  if (rc == P2B_OK) rc = tree_from_colmajor(ctx, &b->tree, b->d_lde, N, log_n + rate_bits, n_cols, cap_height);
  if (rc != P2B_OK) {
    p2b_batch_free(b);
    return rc;
  }
  *out = b;
  return P2B_OK;
}

static const size_t UPLOAD_HALF_WORDS = (size_t)1 << 20;  // 8 MiB per half

static bool host_ptr_is_pinned(const void* p) {
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // unregistered memory reports an error on old drivers: not sticky
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// Host columns [c0, c1) -> d_dst (column c at d_dst + c * n) on `stream`.  Pinned memory (p2b_host_alloc,
// cudaHostRegister) goes by direct DMA, one copy per run of contiguous columns; pageable memory is staged through
// the context's pinned double buffer: consecutive pageable columns — plonky2's Vec<PolynomialValues>, one allocation per
// column, wherever they lie — are packed into one half (their destinations are contiguous), one DMA per 8 MiB instead
// of a synchronous, internally staged cudaMemcpy (or an event wait + DMA) per column.
static int h2d_cols(p2b_ctx* ctx, cudaStream_t stream, const uint64_t* const* cols, size_t c0, size_t c1, size_t n,
                    uint64_t* d_dst) {
  size_t c = c0;
  while (c < c1) {
    if (!cols[c]) return fail(ctx, P2B_ERR_INVALID, "cols[%zu] is null", c);
    if (host_ptr_is_pinned(cols[c])) {
      size_t e = c + 1;
      while (e < c1 && cols[e] == cols[e - 1] + n) e++;
      CU(ctx, cudaMemcpyAsync(d_dst + c * n, cols[c], (e - c) * n * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
      // the DMA reads the caller's memory: remember where it ends (the host-input entry points wait for it)
      if (!ctx->ev_h2d) CU(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d, cudaEventDisableTiming));
      CU(ctx, cudaEventRecord(ctx->ev_h2d, stream));
      ctx->h2d_event_pending = true;
      c = e;
      continue;
    }
    // a run of pageable columns (it ends at the next pinned one): pack whole columns, or pieces of a long one, half by half
    size_t off = 0;  // words already sent of column c
    bool run = true;
    while (run && c < c1) {
      const int h = ctx->upload_half;
      if (!ctx->h_upload[h]) {
        CU(ctx, cudaMallocHost((void**)&ctx->h_upload[h], UPLOAD_HALF_WORDS * sizeof(uint64_t)));
        CU(ctx, cudaEventCreateWithFlags(&ctx->upload_done[h], cudaEventDisableTiming));
      } else {
        CU(ctx, cudaEventSynchronize(ctx->upload_done[h]));  // the previous DMA out of this half has finished
      }
      size_t filled = 0;
      const size_t first_c = c, first_off = off;
      while (c < c1 && filled < UPLOAD_HALF_WORDS) {
        if (off == 0 && c != first_c) {  // a column the run has not looked at yet
          if (!cols[c]) return fail(ctx, P2B_ERR_INVALID, "cols[%zu] is null", c);
          if (host_ptr_is_pinned(cols[c])) {
            run = false;
            break;
          }
        }
        const size_t take = (n - off) < (UPLOAD_HALF_WORDS - filled) ? (n - off) : (UPLOAD_HALF_WORDS - filled);
        memcpy(ctx->h_upload[h] + filled, cols[c] + off, take * sizeof(uint64_t));
        filled += take;
        off += take;
        if (off == n) c++, off = 0;
      }
      if (filled) {
        // the packed words are contiguous in the destination too (column-major, consecutive columns)
        CU(ctx, cudaMemcpyAsync(d_dst + first_c * n + first_off, ctx->h_upload[h], filled * sizeof(uint64_t),
                                cudaMemcpyHostToDevice, stream));
        CU(ctx, cudaEventRecord(ctx->upload_done[h], stream));
        ctx->upload_half ^= 1;
      }
    }
  }
  return P2B_OK;
}

static int upload_cols(p2b_ctx* ctx, const uint64_t* const* cols, size_t n_cols, size_t n, uint64_t** d_out) {
  int rc = dmalloc(ctx, d_out, n_cols * n);
  if (rc) return rc;
  return h2d_cols(ctx, ctx->stream, cols, 0, n_cols, n, *d_out);
}

static cudaEvent_t copy_event(p2b_ctx* ctx, size_t i) {
  while (ctx->copy_events.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    ctx->copy_events.push_back(e);
  }
  return ctx->copy_events[i];
}

// from_values / from_coeffs of host columns, pipelined: the columns are uploaded in groups on the copy stream and
// the transforms of group g (inverse NTT, LDE) run on the compute stream while group g+1 is still in flight, so
// that for a wide batch only the first group's copy and the leaf hashing are exposed.
static int batch_from_host_pipelined(p2b_ctx* ctx, const uint64_t* const* cols, size_t n_cols, uint32_t log_n,
                                     uint32_t rate_bits, uint32_t cap_height, bool is_values, bool keep_values,
                                     p2b_batch** out) {
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  const size_t group = 16;
  p2b_batch* b = new (std::nothrow) p2b_batch();
  if (!b) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  b->ctx = ctx;
  b->n_cols = n_cols;
  b->log_n = log_n;
  b->rate_bits = rate_bits;
  b->cap_height = cap_height;
  b->tree.owned_by_batch = true;
  uint64_t* d_in = nullptr;
  int rc = dmalloc(ctx, &d_in, n_cols * n);
  if (rc == P2B_OK) rc = dmalloc(ctx, &b->d_lde, n_cols * N);
  if (rc == P2B_OK && is_values) rc = dmalloc(ctx, &b->d_coeffs, n_cols * n);
  auto bail = [&](int code) {
    cudaStreamSynchronize(ctx->copy_stream);  // no copy may still target buffers that are about to be freed
    if (d_in != b->d_coeffs && d_in != b->d_values) dfree(ctx, d_in);
    p2b_batch_free(b);
    return code;
  };
  if (rc) return bail(rc);
  if (!is_values) b->d_coeffs = d_in;
  // the copy stream must not write before the (stream-ordered) allocations have happened
  cudaEvent_t ev0 = copy_event(ctx, 0);
  if (!ev0) return bail(fail(ctx, P2B_ERR_CUDA, "cudaEventCreate failed"));
  if (cudaEventRecord(ev0, ctx->stream) != cudaSuccess || cudaStreamWaitEvent(ctx->copy_stream, ev0, 0) != cudaSuccess)
    return bail(fail(ctx, P2B_ERR_CUDA, "event setup failed"));
  // The leaf sponge runs in column ranges as well (hashk::k_leaf_absorb_colmajor, state carried in HBM): the upload
  // is the slower side of the copy/transform pipeline, so without this the GPU idles for most of the transfer and
  // then hashes for 7 ms; with it the hashing of the first columns covers the rest of the upload.  The first range
  // is one upload group (16 columns: hashing starts after 0.15 ms of transfer instead of 0.45 ms at 2^16 rows), the
  // ranges then grow to 48 columns (6 permutations per leaf and launch) to keep the number of launch tails small.
  p2b_tree* t = &b->tree;
  t->ctx = ctx;
  t->n_leaves = N;
  t->log_leaves = log_n + rate_bits;
  t->cap_height = cap_height;
  t->leaf_len = n_cols;
  t->d_leaves_cm = b->d_lde;
  uint64_t* d_sponge = nullptr;
  rc = dmalloc(ctx, &t->d_levels, 4 * levels_len(N, cap_height));
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_sponge, 12 * N);
  if (rc) {
    dfree(ctx, d_sponge);
    return bail(rc);
  }
  size_t hash_span = group;
  size_t hashed = 0;
  stage_begin(ctx, ST_H2D);
  size_t gi = 0;
  for (size_t c0 = 0; c0 < n_cols; c0 += group, gi++) {
    const size_t c1 = c0 + group < n_cols ? c0 + group : n_cols;
    rc = h2d_cols(ctx, ctx->copy_stream, cols, c0, c1, n, d_in);
    if (rc) break;
    cudaEvent_t ev = copy_event(ctx, gi + 1);
    if (!ev || cudaEventRecord(ev, ctx->copy_stream) != cudaSuccess || cudaStreamWaitEvent(ctx->stream, ev, 0) != cudaSuccess) {
      rc = fail(ctx, P2B_ERR_CUDA, "event setup failed");
      break;
    }
    const size_t k = c1 - c0;
    if (is_values) {
      stage_begin(ctx, ST_INTT);
      rc = run_intt(ctx, d_in + c0 * n, b->d_coeffs + c0 * n, b->d_lde + c0 * N, k, log_n, N);
      if (rc) break;
    }
    stage_begin(ctx, ST_LDE);
    rc = run_lde(ctx, b->d_coeffs + c0 * n, n, b->d_lde + c0 * N, k, log_n, rate_bits, 7);
    if (rc) break;
    if (c1 == n_cols || c1 - hashed >= hash_span) {  // group boundaries are multiples of 16, hence of the rate 8
      stage_begin(ctx, ST_LEAF);
      hashk::k_leaf_absorb_colmajor<<<cdiv(N, 256), 256, 0, ctx->stream>>>(b->d_lde, N, (uint32_t)hashed, (uint32_t)c1,
                                                                         (uint32_t)n_cols, N, d_sponge, t->d_levels);
      ctx->launches++;
      if (cudaGetLastError() != cudaSuccess) {
        rc = fail(ctx, P2B_ERR_CUDA, "k_leaf_absorb_colmajor launch failed");
        break;
      }
      hashed = c1;
      if (hash_span < 48) hash_span = hash_span * 2 < 48 ? hash_span * 2 : 48;
    }
  }
  dfree(ctx, d_sponge);
  if (rc) {
    stage_end(ctx);
    return bail(rc);
  }
  stage_begin(ctx, ST_TREE);
  rc = build_levels(ctx, t);
  stage_end(ctx);
  if (rc) return bail(rc);
  if (is_values) {
    if (keep_values)
      b->d_values = d_in;
    else
      dfree(ctx, d_in);
  }
  *out = b;
  return P2B_OK;
}

static int batch_from_host(p2b_ctx* ctx, const uint64_t* const* cols, size_t n_cols, uint32_t log_n, uint32_t rate_bits,
                           uint32_t cap_height, uint32_t flags, bool is_values, p2b_batch** out) {
  CHECK_CTX(ctx);
  int rc = check_batch_args(ctx, cols, n_cols, log_n, rate_bits, cap_height, flags, out);
  if (rc) return rc;
  // wide batches of long columns: overlap the upload with the transforms, group by group
  if (n_cols >= 32 && log_n >= 14)
    return batch_from_host_pipelined(ctx, cols, n_cols, log_n, rate_bits, cap_height, is_values,
                                     is_values && (flags & P2B_KEEP_VALUES), out);
  uint64_t* d_in = nullptr;
  stage_begin(ctx, ST_H2D);
  rc = upload_cols(ctx, cols, n_cols, (size_t)1 << log_n, &d_in);
  stage_end(ctx);
  if (rc) {
    dfree(ctx, d_in);
    return rc;
  }
  return batch_build(ctx, d_in, is_values, n_cols, log_n, rate_bits, cap_height, out,
                     is_values && (flags & P2B_KEEP_VALUES));
}

static int batch_from_dev(p2b_ctx* ctx, const uint64_t* d_cols, size_t n_cols, uint32_t log_n, uint32_t rate_bits,
                          uint32_t cap_height, uint32_t flags, bool is_values, p2b_batch** out) {
  CHECK_CTX(ctx);
  int rc = check_batch_args(ctx, d_cols, n_cols, log_n, rate_bits, cap_height, flags, out);
  if (rc) return rc;
  const size_t n = (size_t)1 << log_n;
  uint64_t* d_in = nullptr;
  if (is_values) {
    // the inverse transform reads the caller's buffer directly; nothing to copy
    p2b_batch* b = nullptr;
    const size_t N = n << rate_bits;
    b = new (std::nothrow) p2b_batch();
    if (!b) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
    b->ctx = ctx;
    b->n_cols = n_cols;
    b->log_n = log_n;
    b->rate_bits = rate_bits;
    b->cap_height = cap_height;
    b->tree.owned_by_batch = true;
    rc = dmalloc(ctx, &b->d_lde, n_cols * N);
    if (rc == P2B_OK) rc = dmalloc(ctx, &b->d_coeffs, n_cols * n);
    if (rc == P2B_OK && (flags & P2B_KEEP_VALUES)) {
      rc = dmalloc(ctx, &b->d_values, n_cols * n);
      if (rc == P2B_OK) {
        cudaError_t e = cudaMemcpyAsync(b->d_values, d_cols, n_cols * n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream);
        if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
      }
    }
    stage_begin(ctx, ST_INTT);
    if (rc == P2B_OK) rc = run_intt(ctx, d_cols, b->d_coeffs, b->d_lde, n_cols, log_n, N);
    stage_begin(ctx, ST_LDE);
    if (rc == P2B_OK) rc = run_lde(ctx, b->d_coeffs, n, b->d_lde, n_cols, log_n, rate_bits, 7);
    stage_end(ctx);
    if (rc == P2B_OK) rc = tree_from_colmajor(ctx, &b->tree, b->d_lde, N, log_n + rate_bits, n_cols, cap_height);
    if (rc != P2B_OK) {
      p2b_batch_free(b);
      return rc;
    }
    *out = b;
    return P2B_OK;
  }
  rc = dmalloc(ctx, &d_in, n_cols * n);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(d_in, d_cols, n_cols * n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  return batch_build(ctx, d_in, false, n_cols, log_n, rate_bits, cap_height, out);
}

// Pinned sources are read by DMA after the call has enqueued its work: wait for the UPLOAD (not for the transforms) so
// that the caller may refill its buffers as soon as the entry point returns.
static int wait_for_upload(p2b_ctx* ctx, int rc) {
  if (ctx && ctx->h2d_event_pending) {
    ctx->h2d_event_pending = false;
    cudaError_t e = cudaEventSynchronize(ctx->ev_h2d);
    if (e != cudaSuccess && rc == P2B_OK) return fail(ctx, P2B_ERR_CUDA, "cudaEventSynchronize: %s", cudaGetErrorString(e));
  }
  return rc;
}
extern "C" int p2b_batch_from_values(p2b_ctx* ctx, const uint64_t* const* cols, size_t n_cols, uint32_t log_n,
                                     uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch** out) {
  return wait_for_upload(ctx, batch_from_host(ctx, cols, n_cols, log_n, rate_bits, cap_height, flags, true, out));
}
extern "C" int p2b_batch_from_coeffs(p2b_ctx* ctx, const uint64_t* const* cols, size_t n_cols, uint32_t log_n,
                                     uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch** out) {
  return wait_for_upload(ctx, batch_from_host(ctx, cols, n_cols, log_n, rate_bits, cap_height, flags, false, out));
}
extern "C" int p2b_batch_from_values_dev(p2b_ctx* ctx, const uint64_t* d_cols, size_t n_cols, uint32_t log_n,
                                         uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch** out) {
  return batch_from_dev(ctx, d_cols, n_cols, log_n, rate_bits, cap_height, flags, true, out);
}
extern "C" int p2b_batch_from_coeffs_dev(p2b_ctx* ctx, const uint64_t* d_cols, size_t n_cols, uint32_t log_n,
                                         uint32_t rate_bits, uint32_t cap_height, uint32_t flags, p2b_batch** out) {
  return batch_from_dev(ctx, d_cols, n_cols, log_n, rate_bits, cap_height, flags, false, out);
}

extern "C" void p2b_batch_free(p2b_batch* b) {
  if (!b) return;
  p2b_ctx* ctx = b->ctx;
  cudaSetDevice(ctx->device);
  if (!ctx->plans.empty() && !ctx->cap) plans_forget(ctx, 0, b->id);
  if (b->view) {
    delete b;
    return;
  }
  dfree(ctx, b->d_coeffs);
  dfree(ctx, b->d_lde);
  dfree(ctx, b->d_values);
  dfree(ctx, b->tree.d_levels);
  delete b;
}

// ---- circuit-data reuse -------------------------------------------------------------------------
// prover_data.constants_sigmas_commitment is built once per circuit (CircuitBuilder::build: minutes for the whole
// toolbox, city_rollup_core_worker_qbench/src/qbench.rs:21; at job time for the sighash wrappers,
// city_rollup_circuit/src/sighash_circuits/sighash_wrapper.rs:142-148) and read by every proof of that circuit.
// p2b_batch_attach shares ONE device copy between the contexts of a device; p2b_batch_export / import carry it across
// processes as coefficients (+ values) + cap, the LDE and the tree being recomputed on the device and checked
// against the stored cap.
extern "C" int p2b_batch_attach(p2b_ctx* ctx, const p2b_batch* src, p2b_batch** out) {
  CHECK_CTX(ctx);
  if (!src || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  if (src->ctx->device != ctx->device) return fail(ctx, P2B_ERR_INVALID, "the batch lives on device %d, this context on %d", src->ctx->device, ctx->device);
  p2b_batch* v = new (std::nothrow) p2b_batch();
  if (!v) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  const uint64_t id = v->id;
  *v = *src;  // device pointers are shared
  v->id = id;
  v->ctx = ctx;
  v->view = true;
  v->tree.ctx = ctx;
  v->tree.owned_by_batch = true;
  if (src->ctx != ctx) {
    // everything the owner has enqueued for this batch must be done before this context's stream reads it
    cudaEvent_t ev = nullptr;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventRecord(ev, src->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ev, 0);
    if (ev) cudaEventDestroy(ev);
    if (e != cudaSuccess) {
      delete v;
      return fail(ctx, P2B_ERR_CUDA, "p2b_batch_attach: %s", cudaGetErrorString(e));
    }
  }
  *out = v;
  return P2B_OK;
}

static const uint64_t P2B_EXPORT_MAGIC = 0x3148435441423250ull;  // "P2BATCH1"
extern "C" size_t p2b_batch_export_len(const p2b_batch* b) {
  if (!b) return 0;
  const size_t n = (size_t)1 << b->log_n;
  return 8 * (8 + ((size_t)4 << b->cap_height) + b->n_cols * n * (b->d_values ? 2 : 1));
}
extern "C" int p2b_batch_export(p2b_batch* b, uint8_t* out, size_t out_cap, size_t* written) {
  if (!b || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  const size_t need = p2b_batch_export_len(b);
  if (out_cap < need) return fail(ctx, P2B_ERR_INVALID, "export buffer too small: %zu < %zu bytes", out_cap, need);
  const size_t n = (size_t)1 << b->log_n, cap_words = (size_t)4 << b->cap_height;
  uint64_t hdr[8] = {P2B_EXPORT_MAGIC, 1, b->n_cols, b->log_n, b->rate_bits, b->cap_height, b->d_values ? 1u : 0u, 0};
  memcpy(out, hdr, sizeof hdr);
  uint64_t* w = (uint64_t*)(out + sizeof hdr);  // callers hand over malloc'ed (8-byte aligned) buffers; memcpy below otherwise
  std::vector<uint64_t> tmp(cap_words);
  int rc = p2b_tree_cap(&b->tree, tmp.data());
  if (rc) return rc;
  memcpy(w, tmp.data(), cap_words * 8);
  uint8_t* p = (uint8_t*)(w + cap_words);
  // large copies straight into the caller's buffer (pageable or pinned)
  CU(ctx, cudaMemcpyAsync(p, b->d_coeffs, b->n_cols * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (b->d_values) CU(ctx, cudaMemcpyAsync(p + b->n_cols * n * 8, b->d_values, b->n_cols * n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, ctx_sync(ctx));
  if (written) *written = need;
  return P2B_OK;
}
extern "C" int p2b_batch_import(p2b_ctx* ctx, const uint8_t* bytes, size_t n_bytes, p2b_batch** out) {
  CHECK_CTX(ctx);
  if (!bytes || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  uint64_t hdr[8];
  if (n_bytes < sizeof hdr) return fail(ctx, P2B_ERR_INVALID, "truncated batch export");
  memcpy(hdr, bytes, sizeof hdr);
  if (hdr[0] != P2B_EXPORT_MAGIC || hdr[1] != 1) return fail(ctx, P2B_ERR_INVALID, "not a p2b batch export (magic / version)");
  const size_t n_cols = hdr[2];
  const uint32_t log_n = (uint32_t)hdr[3], rate_bits = (uint32_t)hdr[4], cap_height = (uint32_t)hdr[5];
  const bool has_values = hdr[6] != 0;
  if (hdr[3] > 24 || hdr[4] > 6 || hdr[5] > 30 || n_cols == 0 || n_cols > 65535) return fail(ctx, P2B_ERR_INVALID, "corrupt batch export header");
  const size_t n = (size_t)1 << log_n, cap_words = (size_t)4 << cap_height;
  if (cap_height > log_n + rate_bits) return fail(ctx, P2B_ERR_INVALID, "corrupt batch export header");
  const size_t need = 8 * (8 + cap_words + n_cols * n * (has_values ? 2 : 1));
  if (n_bytes != need) return fail(ctx, P2B_ERR_INVALID, "batch export is %zu bytes, its header says %zu", n_bytes, need);
  const uint8_t* p = bytes + sizeof hdr + cap_words * 8;
  uint64_t *d_coeffs = nullptr, *d_values = nullptr;
  int rc = dmalloc(ctx, &d_coeffs, n_cols * n);
  if (rc) return rc;
  cudaError_t e = cudaMemcpyAsync(d_coeffs, p, n_cols * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess && has_values) {
    rc = dmalloc(ctx, &d_values, n_cols * n);
    if (rc == P2B_OK) e = cudaMemcpyAsync(d_values, p + n_cols * n * 8, n_cols * n * 8, cudaMemcpyHostToDevice, ctx->stream);
  }
  if (e != cudaSuccess || rc != P2B_OK) {
    dfree(ctx, d_coeffs);
    dfree(ctx, d_values);
    return rc != P2B_OK ? rc : fail(ctx, P2B_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
  }
  p2b_batch* b = nullptr;
  rc = batch_build(ctx, d_coeffs, false, n_cols, log_n, rate_bits, cap_height, &b);  // LDE + tree recomputed on the device
  if (rc != P2B_OK) {
    dfree(ctx, d_values);
    return rc;
  }
  b->d_values = d_values;
  std::vector<uint64_t> cap(cap_words);
  rc = p2b_tree_cap(&b->tree, cap.data());
  if (rc == P2B_OK && memcmp(cap.data(), bytes + sizeof hdr, cap_words * 8) != 0)
    rc = fail(ctx, P2B_ERR_INVALID, "batch export is corrupt: the recomputed Merkle cap differs from the stored one");
  if (rc != P2B_OK) {
    p2b_batch_free(b);
    return rc;
  }
  *out = b;
  return P2B_OK;
}

extern "C" size_t p2b_batch_n_cols(const p2b_batch* b) { return b ? b->n_cols : 0; }
extern "C" uint32_t p2b_batch_degree_log(const p2b_batch* b) { return b ? b->log_n : 0; }
extern "C" uint32_t p2b_batch_rate_bits(const p2b_batch* b) { return b ? b->rate_bits : 0; }
extern "C" p2b_tree* p2b_batch_tree(p2b_batch* b) { return b ? &b->tree : nullptr; }
extern "C" const uint64_t* p2b_batch_dev_lde(const p2b_batch* b) { return b ? b->d_lde : nullptr; }
extern "C" const uint64_t* p2b_batch_dev_coeffs(const p2b_batch* b) { return b ? b->d_coeffs : nullptr; }

extern "C" int p2b_batch_cap(p2b_batch* b, uint64_t* out) {
  if (!b) return P2B_ERR_INVALID;
  return p2b_tree_cap(&b->tree, out);
}
extern "C" int p2b_batch_coeffs(p2b_batch* b, size_t col, uint64_t* out) {
  if (!b || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  if (col >= b->n_cols) return fail(ctx, P2B_ERR_INVALID, "column %zu out of range", col);
  size_t n = (size_t)1 << b->log_n;
  return d2h(ctx, out, b->d_coeffs + col * n, n);
}
extern "C" int p2b_batch_values(p2b_batch* b, size_t col, uint64_t* out) {
  if (!b || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  if (!b->d_values) return fail(ctx, P2B_ERR_INVALID, "batch was not built from values with P2B_KEEP_VALUES");
  if (col >= b->n_cols) return fail(ctx, P2B_ERR_INVALID, "column %zu out of range", col);
  size_t n = (size_t)1 << b->log_n;
  return d2h(ctx, out, b->d_values + col * n, n);
}

extern "C" int p2b_batch_lde_col(p2b_batch* b, size_t col, uint64_t* out) {
  if (!b || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  if (col >= b->n_cols) return fail(ctx, P2B_ERR_INVALID, "column %zu out of range", col);
  const size_t N = (size_t)1 << (b->log_n + b->rate_bits);
  return d2h(ctx, out, b->d_lde + col * N, N);
}

// ------------------------------------------------------------------------------------------------ PLONK stages
// Degree (in units of n - 1) of the gate's UNFILTERED constraints as polynomials in x: wires and constant columns count
// 1 each.  Only gates whose evaluators are worth extending are listed; 0 = evaluate at every point of the quotient coset.
//   arithmetic: w w c0 (3); base sum: l (l - 1) (2); u32 gates: limb range checks prod_{x<4} (limb - x) (4);
//   interleave / uninterleave: b (b - 1) (2); comparison: prod_{x < 2^chunk_bits} (chunk - x), and eq * msd with
//   msd = inter + (1 - eq) diff (3); extension arithmetic / multiplication: m0 m1 c0 (3); reducing: acc * alpha (2)
static uint32_t gate_constraint_degree(const p2b_gate& gt) {
  switch (gt.kind) {
    case plonk::GATE_ARITHMETIC: return 3;
    case plonk::GATE_BASE_SUM: return 2;
    case plonk::GATE_U32_ARITHMETIC:
    case plonk::GATE_U32_ADD_MANY:
    case plonk::GATE_U32_SUBTRACTION:
    case plonk::GATE_U32_RANGE_CHECK: return 4;
    case plonk::GATE_U32_INTERLEAVE:
    case plonk::GATE_UNINTERLEAVE_TO_U32:
    case plonk::GATE_UNINTERLEAVE_TO_B32: return 2;
    case plonk::GATE_COMPARISON: {
      const uint32_t cb = (gt.p0 + gt.p1 - 1) / gt.p1, prod = 1u << cb;
      return prod > 3 ? prod : 3;
    }
    case plonk::GATE_ARITHMETIC_EXT:
    case plonk::GATE_MUL_EXT: return 3;
    case plonk::GATE_REDUCING:
    case plonk::GATE_REDUCING_EXT: return 2;
    default: return 0;
  }
}

extern "C" int p2b_circuit_new(p2b_ctx* ctx, const p2b_circuit_desc* desc, p2b_circuit** out) {
  CHECK_CTX(ctx);
  if (!desc || !out || !desc->gates || !desc->k_is) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  const p2b_circuit_desc& d = *desc;
  if (d.degree_bits > 24) return fail(ctx, P2B_ERR_UNSUPPORTED, "degree_bits %u > 24", d.degree_bits);
  if (d.num_challenges == 0 || d.num_challenges > (uint32_t)plonk::MAX_CHALLENGES)
    return fail(ctx, P2B_ERR_UNSUPPORTED, "num_challenges must be in 1..%d", plonk::MAX_CHALLENGES);
  if (d.num_routed_wires == 0 || d.num_routed_wires > d.num_wires) return fail(ctx, P2B_ERR_INVALID, "num_routed_wires");
  if (d.quotient_degree_factor == 0 || (d.quotient_degree_factor & (d.quotient_degree_factor - 1)))
    return fail(ctx, P2B_ERR_UNSUPPORTED, "quotient_degree_factor %u is not a power of two", d.quotient_degree_factor);
  if (d.num_partial_products + 1 != (d.num_routed_wires + d.quotient_degree_factor - 1) / d.quotient_degree_factor)
    return fail(ctx, P2B_ERR_INVALID, "num_partial_products %u does not match ceil(%u / %u) - 1", d.num_partial_products,
                d.num_routed_wires, d.quotient_degree_factor);
  if (d.num_selectors > d.num_constants) return fail(ctx, P2B_ERR_INVALID, "num_selectors > num_constants");
  if (d.n_gates == 0) return fail(ctx, P2B_ERR_INVALID, "no gates");
  for (uint32_t g = 0; g < d.n_gates; g++) {
    const p2b_gate& gt = d.gates[g];
    if (gt.kind >= plonk::GATE_KIND_COUNT) return fail(ctx, P2B_ERR_UNSUPPORTED, "gate %u: unknown kind %u", g, gt.kind);
    if (gt.selector_index >= d.num_selectors || gt.group_start > gt.row || gt.row >= gt.group_end)
      return fail(ctx, P2B_ERR_INVALID, "gate %u: bad selector layout", g);
    // wires / constants / constraints the gate touches must exist
    uint32_t wires = 0, consts = 0, cons = 0;
    switch (gt.kind) {
      case plonk::GATE_CONSTANT: wires = gt.p0, consts = gt.p0, cons = gt.p0; break;
      case plonk::GATE_PUBLIC_INPUT: wires = 4, cons = 4; break;
      case plonk::GATE_ARITHMETIC: wires = 4 * gt.p0, consts = 2, cons = gt.p0; break;
      case plonk::GATE_POSEIDON: wires = 135, cons = 123; break;
      case plonk::GATE_BASE_SUM: wires = 1 + gt.p0, cons = 1 + gt.p0; break;
      case plonk::GATE_U32_ARITHMETIC: wires = 38 * gt.p0, cons = 36 * gt.p0; break;
      case plonk::GATE_U32_ADD_MANY: wires = (gt.p0 + 3 + 18) * gt.p1, cons = 21 * gt.p1; break;
      case plonk::GATE_U32_SUBTRACTION: wires = 21 * gt.p0, cons = 19 * gt.p0; break;
      case plonk::GATE_U32_RANGE_CHECK: wires = 17 * gt.p0, cons = 17 * gt.p0; break;
      case plonk::GATE_U32_INTERLEAVE: wires = 34 * gt.p0, cons = 34 * gt.p0; break;
      case plonk::GATE_UNINTERLEAVE_TO_U32:
      case plonk::GATE_UNINTERLEAVE_TO_B32: wires = 67 * gt.p0, cons = 67 * gt.p0; break;
      case plonk::GATE_ARITHMETIC_EXT: wires = 8 * gt.p0, consts = 2, cons = 2 * gt.p0; break;
      case plonk::GATE_MUL_EXT: wires = 6 * gt.p0, consts = 1, cons = 2 * gt.p0; break;
      case plonk::GATE_REDUCING: wires = gt.p0 ? 6 + gt.p0 + 2 * (gt.p0 - 1) : 6, cons = 2 * gt.p0; break;
      case plonk::GATE_REDUCING_EXT: wires = gt.p0 ? 6 + 2 * gt.p0 + 2 * (gt.p0 - 1) : 6, cons = 2 * gt.p0; break;
      case plonk::GATE_RANDOM_ACCESS: {
        const uint32_t copies = gt.p1 & 0xFFFF, extra = gt.p1 >> 16;
        if (gt.p0 == 0 || gt.p0 > 6) return fail(ctx, P2B_ERR_UNSUPPORTED, "gate %u: RandomAccessGate bits %u", g, gt.p0);
        wires = (2 + (1u << gt.p0) + gt.p0) * copies + extra, consts = extra, cons = (gt.p0 + 2) * copies + extra;
        break;
      }
      case plonk::GATE_POSEIDON_MDS: wires = 48, cons = 24; break;
      case plonk::GATE_COSET_INTERPOLATION: {
        if (gt.p0 == 0 || gt.p0 > 6 || gt.p1 < 2 || gt.p1 > (1u << gt.p0))
          return fail(ctx, P2B_ERR_UNSUPPORTED, "gate %u: CosetInterpolationGate(%u, degree %u)", g, gt.p0, gt.p1);
        const uint32_t np = 1u << gt.p0, n_int = (np - 2) / (gt.p1 - 1);
        wires = 1 + 2 * np + 4 + 2 * (2 * n_int + 1), cons = 4 + 4 * n_int;
        break;
      }
      case plonk::GATE_COMPARISON: {
        if (gt.p1 == 0 || gt.p0 == 0 || (gt.p0 + gt.p1 - 1) / gt.p1 > 8)
          return fail(ctx, P2B_ERR_UNSUPPORTED, "gate %u: ComparisonGate(%u, %u)", g, gt.p0, gt.p1);
        const uint32_t cb = (gt.p0 + gt.p1 - 1) / gt.p1;
        wires = 4 + 5 * gt.p1 + cb + 1, cons = 6 + 5 * gt.p1 + cb;
        break;
      }
      default: break;
    }
    if (wires > d.num_wires || consts > d.num_constants - d.num_selectors || cons > d.num_gate_constraints)
      return fail(ctx, P2B_ERR_INVALID, "gate %u (kind %u) needs %u wires / %u constants / %u constraints", g, gt.kind,
                  wires, consts, cons);
  }
  p2b_circuit* c = new (std::nothrow) p2b_circuit();
  if (!c) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  c->ctx = ctx;
  c->d = d;
  c->d.gates = nullptr;
  c->d.k_is = nullptr;
  static_assert(sizeof(plonk::Gate) == sizeof(p2b_gate), "gate layout");
  int rc = dmalloc(ctx, (uint64_t**)&c->d_gates, (d.n_gates * sizeof(p2b_gate) + 7) / 8);
  if (rc == P2B_OK) rc = dmalloc(ctx, &c->d_k_is, d.num_routed_wires);
  while ((1u << c->mdb) < d.quotient_degree_factor) c->mdb++;
  // ZeroPolyOnCoset::new(degree_bits, mdb): Z_H(7 w^i) = 7^n w_{2^mdb}^(i mod 2^mdb) - 1, and the inverses
  std::vector<uint64_t> zh((size_t)2 << c->mdb);
  {
    const uint64_t G = 1753635133440165772ull;
    const uint32_t mdb = c->mdb;
    uint64_t g_pow_n = h_powmod(7, (uint64_t)1 << d.degree_bits), wr = mdb ? h_powmod(G, (uint64_t)1 << (32 - mdb)) : 1, xr = 1;
    for (size_t i = 0; i < ((size_t)1 << mdb); i++) {
      const uint64_t v = h_mulmod(g_pow_n, xr);
      zh[i] = v ? v - 1 : GL_P - 1;
      if (zh[i] == 0 && rc == P2B_OK) rc = fail(ctx, P2B_ERR_INVALID, "Z_H vanishes on the quotient coset");
      zh[((size_t)1 << mdb) + i] = h_powmod(zh[i], GL_P - 2);
      xr = h_mulmod(xr, wr);
    }
  }
  if (rc == P2B_OK) rc = dmalloc(ctx, &c->d_zh, zh.size());
  for (uint32_t g = 0; g < d.n_gates; g++) c->gate_kinds.push_back(d.gates[g].kind);
  std::vector<uint32_t> gate_list, part_list{0};
  // extension class of a gate: log2 of the smallest power of two >= its constraint degree, if that is below the
  // quotient degree factor (and the gate is worth two NTT columns); 0 = evaluate at every point
  static const bool ext_on = [] {
    const char* e = getenv("P2B_QUOT_EXT");
    return !e || atoi(e) != 0;
  }();
  auto ext_class = [&](const p2b_gate& gt) -> uint32_t {
    const uint32_t deg = gate_constraint_degree(gt);
    if (!ext_on || deg == 0) return 0;
    uint32_t ld = 1;
    while ((1u << ld) < deg) ld++;
    return ld < c->mdb && ld <= 2 ? ld : 0;
  };
  for (int pass = 0; pass < 4; pass++)  // light direct, heavy direct, extended class 1, extended class 2
    for (uint32_t g = 0; g < d.n_gates; g++) {
      const uint32_t k = d.gates[g].kind;
      const bool h = k == plonk::GATE_POSEIDON || k == plonk::GATE_POSEIDON_MDS || k == plonk::GATE_RANDOM_ACCESS ||
                     k == plonk::GATE_COSET_INTERPOLATION;
      if (k == plonk::GATE_NOOP) continue;  // no constraints, no part
      const uint32_t ec = ext_class(d.gates[g]);
      if (pass < 2) {
        if (ec == 0 && h == (pass == 1)) {
          gate_list.push_back(g);
          part_list.push_back(1 + g);
          (pass ? c->n_heavy : c->n_light)++;
        }
      } else if (ec == (uint32_t)(pass - 1)) {
        gate_list.push_back(g);
        c->n_ext[ec]++;
      }
    }
  c->n_parts_direct = (uint32_t)part_list.size();
  gate_list.insert(gate_list.end(), part_list.begin(), part_list.end());
  if (rc == P2B_OK) rc = dmalloc(ctx, (uint64_t**)&c->d_gate_list, (gate_list.size() + 2) / 2);
  if (rc == P2B_OK && !gate_list.empty()) {
    cudaError_t e = cudaMemcpyAsync(c->d_gate_list, gate_list.data(), gate_list.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "circuit upload: %s", cudaGetErrorString(e));
  }
  if (rc == P2B_OK) {
    // pageable sources: the copies are complete (staged) when cudaMemcpyAsync returns
    cudaError_t e = cudaMemcpyAsync(c->d_gates, d.gates, d.n_gates * sizeof(p2b_gate), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(c->d_zh, zh.data(), zh.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(c->d_k_is, d.k_is, d.num_routed_wires * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = ctx_sync(ctx);
    if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "circuit upload: %s", cudaGetErrorString(e));
  }
  if (rc != P2B_OK) {
    p2b_circuit_free(c);
    return rc;
  }
  *out = c;
  return P2B_OK;
}

extern "C" void p2b_circuit_free(p2b_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  if (!c->ctx->plans.empty()) plans_forget(c->ctx, c->id, 0);
  dfree(c->ctx, c->d_gates);
  dfree(c->ctx, c->d_k_is);
  dfree(c->ctx, c->d_zh);
  dfree(c->ctx, c->d_gate_list);
  delete c;
}

static int check_plonk_batches(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const p2b_batch* wires) {
  if (!c || !cs || !wires) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (c->ctx != ctx || cs->ctx != ctx || wires->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "handle of another context");
  if (cs->log_n != c->d.degree_bits || wires->log_n != c->d.degree_bits)
    return fail(ctx, P2B_ERR_INVALID, "batch degree does not match the circuit");
  if (cs->n_cols != (size_t)c->d.num_constants + c->d.num_routed_wires)
    return fail(ctx, P2B_ERR_INVALID, "constants_sigmas has %zu columns, expected %u", cs->n_cols,
                c->d.num_constants + c->d.num_routed_wires);
  if (wires->n_cols != c->d.num_wires)
    return fail(ctx, P2B_ERR_INVALID, "wires has %zu columns, expected %u", wires->n_cols, c->d.num_wires);
  if (cs->rate_bits != wires->rate_bits) return fail(ctx, P2B_ERR_INVALID, "rate_bits differ");
  return P2B_OK;
}

// n words of host data (pointer tables, parameters) -> device.  The source is pageable: cudaMemcpyAsync stages it
// before returning, so the caller may free it right away.
// Under stream capture the copy node must read memory that is still there at launch time: the words are parked in a
// slot of the plan's pinned block (returned through `slot`, so that the launcher can refresh per-proof inputs).
static int h2d_small(p2b_ctx* ctx, uint64_t* d_dst, const void* h_src, size_t n_u64, uint64_t** slot = nullptr) {
  if (ctx->cap) {
    ProvePlan* pl = ctx->cap;
    const size_t need = n_u64 ? n_u64 : 1;
    if (pl->pin_used + need > pl->pin_cap) return fail(ctx, P2B_ERR_UNSUPPORTED, "prove plan: pinned block too small");
    uint64_t* h = pl->h_pin + pl->pin_used;
    pl->pin_used += need;
    if (n_u64) memcpy(h, h_src, n_u64 * sizeof(uint64_t));
    if (slot) *slot = h;
    h_src = h;
  }
  if (n_u64 == 0) return P2B_OK;
  CU(ctx, cudaMemcpyAsync(d_dst, h_src, n_u64 * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  return P2B_OK;
}

// uploads n host field elements (reduced mod p) to a fresh device buffer
static int upload_felts(p2b_ctx* ctx, const uint64_t* h, size_t n, uint64_t** d_out) {
  std::vector<uint64_t> tmp(n ? n : 1);
  for (size_t i = 0; i < n; i++) tmp[i] = h[i] % GL_P;
  int rc = dmalloc(ctx, d_out, n);
  if (rc) return rc;
  // pageable source: staged before cudaMemcpyAsync returns
  CU(ctx, cudaMemcpyAsync(*d_out, tmp.data(), n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  return P2B_OK;
}

// d_betas / d_gammas: num_challenges elements each, on the device
static int zs_pp_core(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const p2b_batch* wires,
                      const uint64_t* d_betas, const uint64_t* d_gammas, uint32_t rate_bits, uint32_t cap_height,
                      p2b_batch** out) {
  *out = nullptr;
  int rc = check_plonk_batches(ctx, c, cs, wires);
  if (rc) return rc;
  if (!cs->d_values || !wires->d_values)
    return fail(ctx, P2B_ERR_INVALID, "constants_sigmas and wires must be built with P2B_KEEP_VALUES");
  const p2b_circuit_desc& d = c->d;
  const size_t n = (size_t)1 << d.degree_bits;
  const uint32_t nch = d.num_challenges, n_chunks = d.num_partial_products + 1;
  const size_t n_cols = (size_t)nch * (1 + d.num_partial_products);
  if (cap_height > d.degree_bits + rate_bits) return fail(ctx, P2B_ERR_INVALID, "cap_height too large");
  uint64_t *d_local = nullptr, *d_z = nullptr, *d_out = nullptr;
  if ((rc = dmalloc(ctx, &d_local, (size_t)nch * n_chunks * n))) return rc;
  if ((rc = dmalloc(ctx, &d_z, (size_t)nch * n))) {
    dfree(ctx, d_local);
    return rc;
  }
  if ((rc = dmalloc(ctx, &d_out, n_cols * n))) {
    dfree(ctx, d_local);
    dfree(ctx, d_z);
    return rc;
  }
  stage_begin(ctx, ST_OTHER);
  plonk::PpParams pp{};
  pp.wires = wires->d_values;
  pp.sigmas = cs->d_values + (size_t)d.num_constants * n;
  pp.k_is = c->d_k_is;
  pp.local = d_local;
  pp.betas = d_betas;
  pp.gammas = d_gammas;
  pp.log_n = d.degree_bits;
  pp.num_routed = d.num_routed_wires;
  pp.chunk = d.quotient_degree_factor;
  pp.n_chunks = n_chunks;
  pp.roots = ctx->roots();
  plonk::k_pp_rows<<<dim3(cdiv(n, 256), nch), 256, 0, ctx->stream>>>(pp);
  LAUNCH_CHECK(ctx);
  plonk::k_pp_scan<<<nch, 1024, 0, ctx->stream>>>(d_local, d.degree_bits, n_chunks, d_z);
  LAUNCH_CHECK(ctx);
  plonk::k_pp_finish<<<dim3(cdiv(n, 256), nch), 256, 0, ctx->stream>>>(d_local, d_z, d.degree_bits, n_chunks, nch, d_out);
  LAUNCH_CHECK(ctx);
  stage_end(ctx);
  dfree(ctx, d_local);
  dfree(ctx, d_z);
  return batch_build(ctx, d_out, true, n_cols, d.degree_bits, rate_bits, cap_height, out, true);
}

extern "C" int p2b_zs_partial_products_commit(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs,
                                              const p2b_batch* wires, const uint64_t* betas, const uint64_t* gammas,
                                              uint32_t rate_bits, uint32_t cap_height, p2b_batch** out) {
  CHECK_CTX(ctx);
  if (!out || !betas || !gammas || !c) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  const uint32_t nch = c->d.num_challenges;
  std::vector<uint64_t> h(2 * nch);
  for (uint32_t i = 0; i < nch; i++) h[i] = betas[i], h[nch + i] = gammas[i];
  uint64_t* d_ch = nullptr;
  int rc = upload_felts(ctx, h.data(), h.size(), &d_ch);
  if (rc == P2B_OK) rc = zs_pp_core(ctx, c, cs, wires, d_ch, d_ch + nch, rate_bits, cap_height, out);
  dfree(ctx, d_ch);
  return rc;
}

static bool d_gate_kinds_noop(const p2b_circuit* c, uint32_t g) { return c->gate_kinds[g] == plonk::GATE_NOOP; }
static uint32_t quotient_n_terms(const p2b_circuit_desc& d) {
  return d.num_challenges * (d.num_partial_products + 2) + d.num_gate_constraints;
}
// all challenge / hash arguments on the device
static int quotient_core(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const p2b_batch* wires,
                         const p2b_batch* zs, const uint64_t* d_pi_hash, const uint64_t* d_betas, const uint64_t* d_gammas,
                         const uint64_t* d_alphas, uint32_t rate_bits, uint32_t cap_height, p2b_batch** out,
                         const uint64_t* d_apow_ready = nullptr /* [challenge][n_terms] powers of alpha, if the caller built them */) {
  *out = nullptr;
  int rc = check_plonk_batches(ctx, c, cs, wires);
  if (rc) return rc;
  const p2b_circuit_desc& d = c->d;
  const uint32_t nch = d.num_challenges, npp = d.num_partial_products, mdb = c->mdb;
  if (!zs || zs->ctx != ctx || zs->log_n != d.degree_bits || zs->n_cols != (size_t)nch * (1 + npp) || zs->rate_bits != cs->rate_bits)
    return fail(ctx, P2B_ERR_INVALID, "zs_partial_products batch does not match the circuit");
  if (mdb > cs->rate_bits)
    return fail(ctx, P2B_ERR_INVALID, "quotient_degree_factor %u exceeds the LDE rate 2^%u", d.quotient_degree_factor, cs->rate_bits);
  const uint32_t log_lde = d.degree_bits + mdb;
  if (log_lde > 24) return fail(ctx, P2B_ERR_UNSUPPORTED, "quotient LDE of 2^%u points", log_lde);
  const size_t n = (size_t)1 << d.degree_bits, lde_size = (size_t)1 << log_lde;
  const uint32_t n_terms = nch * (npp + 2) + d.num_gate_constraints;
  uint64_t *d_apow = nullptr, *d_q = nullptr, *d_coeffs = nullptr, *d_tmp = nullptr, *d_parts = nullptr, *d_apow_g = nullptr;
  const uint32_t n_parts = 1 + d.n_gates;
  if (!d_apow_ready && (rc = dmalloc(ctx, &d_apow, (size_t)nch * n_terms))) return rc;
  if ((rc = dmalloc(ctx, &d_apow_g, (size_t)(d.num_gate_constraints + 1) * plonk::MAX_CHALLENGES))) {
    dfree(ctx, d_apow);
    return rc;
  }
  if ((rc = dmalloc(ctx, &d_parts, (size_t)n_parts * nch * lde_size))) {
    dfree(ctx, d_apow);
    dfree(ctx, d_apow_g);
    return rc;
  }
  if ((rc = dmalloc(ctx, &d_q, (size_t)nch * lde_size)) == P2B_OK) rc = dmalloc(ctx, &d_coeffs, (size_t)nch * lde_size);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_tmp, (size_t)nch * lde_size);
  // extended (low-degree) gates: values / coefficients / NTT scratch on the largest sub-coset, extended sums on all points
  const uint32_t n_ext = c->n_ext[1] + c->n_ext[2];
  uint64_t *d_ext_vals = nullptr, *d_ext_coeffs = nullptr, *d_ext_tmp = nullptr, *d_ext = nullptr;
  if (n_ext) {
    size_t sub = 0;
    for (uint32_t ld = 1; ld <= 2; ld++)
      if ((size_t)c->n_ext[ld] * nch * (n << ld) > sub) sub = (size_t)c->n_ext[ld] * nch * (n << ld);
    if (rc == P2B_OK) rc = dmalloc(ctx, &d_ext_vals, sub);
    if (rc == P2B_OK) rc = dmalloc(ctx, &d_ext_coeffs, sub);
    if (rc == P2B_OK) rc = dmalloc(ctx, &d_ext_tmp, sub);
    if (rc == P2B_OK) rc = dmalloc(ctx, &d_ext, (size_t)n_ext * nch * lde_size);
  }
  if (rc == P2B_OK) {
    stage_begin(ctx, ST_OTHER);
    if (!d_apow_ready) {
      plonk::k_build_apow<<<nch, 32, 0, ctx->stream>>>(d_alphas, n_terms, d_apow);
      ctx->launches++;
    }
    plonk::QuotientParams qp{};
    qp.cs_lde = cs->d_lde;
    qp.wires_lde = wires->d_lde;
    qp.zs_lde = zs->d_lde;
    qp.N = n << cs->rate_bits;
    qp.gates = c->d_gates;
    qp.k_is = c->d_k_is;
    qp.apow = d_apow_ready ? d_apow_ready : d_apow;
    qp.apow_gates = d_apow_g;
    plonk::k_interleave_apow<<<cdiv(d.num_gate_constraints, 256), 256, 0, ctx->stream>>>(qp.apow, n_terms, nch * (npp + 2), nch, d_apow_g);
    ctx->launches++;
    qp.zh = c->d_zh;
    qp.parts = d_parts;
    qp.betas = d_betas;
    qp.gammas = d_gammas;
    qp.pi_hash = d_pi_hash;
    qp.degree_bits = d.degree_bits;
    qp.mdb = mdb;
    qp.num_routed = d.num_routed_wires;
    qp.num_constants = d.num_constants;
    qp.num_selectors = d.num_selectors;
    qp.n_chal = nch;
    qp.chunk = d.quotient_degree_factor;
    qp.num_pp = npp;
    qp.n_gates = d.n_gates;
    qp.n_terms = n_terms;
    qp.roots = ctx->roots();
    plonk::k_quotient_perm<<<dim3(cdiv(lde_size, 128), 1), 128, 0, ctx->stream>>>(qp);
    LAUNCH_CHECK(ctx);
    // CTA order: gate-major.  The point-major order (P2B_QUOT_POINT_MAJOR=1) reads a large circuit's wires from HBM
    // once per kernel instead of once per gate, and is SLOWER: 5.05 against 3.54 ms at 2^16 rows with the recursion gate
    // set, 0.64 against 0.49 ms at 2^12 rows with the City set (profiles/r02_summary.md) — with every gate's code live
    // on every SM the kernels wait for instructions, not for HBM (4.6 GB at 2^16 rows is 0.7 ms of the 3.5).
    static const int force_pm = [] {
      const char* e = getenv("P2B_QUOT_POINT_MAJOR");
      return e ? atoi(e) : 0;
    }();
    qp.point_major = force_pm > 0 ? 1u : 0u;
    const unsigned pblocks = cdiv(lde_size, 128);
    if (c->n_light) {
      qp.list_len = c->n_light;
      const dim3 grid = qp.point_major ? dim3(pblocks * c->n_light, 1) : dim3(pblocks, c->n_light);
      plonk::k_quotient_gates<false, false><<<grid, 128, 0, ctx->stream>>>(qp, c->d_gate_list);
      LAUNCH_CHECK(ctx);
    }
    if (c->n_heavy) {
      qp.list_len = c->n_heavy;
      const dim3 grid = qp.point_major ? dim3(pblocks * c->n_heavy, 1) : dim3(pblocks, c->n_heavy);
      plonk::k_quotient_gates<true, false><<<grid, 128, 0, ctx->stream>>>(qp, c->d_gate_list + c->n_light);
      LAUNCH_CHECK(ctx);
    }
    // low-degree gates: unfiltered sums on the first D n leaves (natural order of that sub-coset), then
    // iNTT (D n) + NTT to all 2^mdb n points, leaf order (see plonk::k_quotient_gates)
    qp.point_major = 0;
    for (uint32_t ld = 1; ld <= 2 && rc == P2B_OK; ld++) {
      if (!c->n_ext[ld]) continue;
      const uint32_t log_pts = d.degree_bits + ld;
      const size_t pts = (size_t)1 << log_pts, cols = (size_t)c->n_ext[ld] * nch;
      const uint32_t first = ld == 1 ? 0 : c->n_ext[1];
      qp.list_len = c->n_ext[ld];
      qp.ext_out = d_ext_vals;
      qp.ext_log_pts = log_pts;
      plonk::k_quotient_gates<false, true><<<dim3(cdiv(pts, 128), c->n_ext[ld]), 128, 0, ctx->stream>>>(
          qp, c->d_gate_list + c->n_light + c->n_heavy + first);
      LAUNCH_CHECK(ctx);
      rc = run_intt(ctx, d_ext_vals, d_ext_coeffs, d_ext_tmp, cols, log_pts, pts);
      if (rc == P2B_OK) rc = run_lde(ctx, d_ext_coeffs, pts, d_ext + (size_t)first * nch * lde_size, cols, log_pts, mdb - ld, 1);
    }
    if (rc == P2B_OK) {
      const uint32_t* lists = c->d_gate_list + c->n_light + c->n_heavy;
      plonk::k_quotient_combine<<<cdiv(lde_size, 256), 256, 0, ctx->stream>>>(
          d_parts, lists + n_ext, c->n_parts_direct, d_ext, lists, n_ext, c->d_gates, cs->d_lde, qp.N, d.num_selectors, nch,
          log_lde, mdb, c->d_zh, d_q);
      LAUNCH_CHECK(ctx);
    }
    stage_end(ctx);
    // values.coset_ifft(7): ifft, then coefficient k / 7^k
    stage_begin(ctx, ST_INTT);
    if (rc == P2B_OK) rc = run_intt(ctx, d_q, d_coeffs, d_tmp, nch, log_lde, lde_size);
    if (rc == P2B_OK) {
      plonk::k_scale_by_powers<<<dim3(cdiv(lde_size, 256 * 16), nch), 256, 0, ctx->stream>>>(d_coeffs, lde_size,
                                                                                            h_powmod(7, GL_P - 2));
      LAUNCH_CHECK(ctx);
    }
    stage_end(ctx);
  }
  dfree(ctx, d_apow);
  dfree(ctx, d_apow_g);
  dfree(ctx, d_parts);
  dfree(ctx, d_q);
  dfree(ctx, d_tmp);
  dfree(ctx, d_ext_vals);
  dfree(ctx, d_ext_coeffs);
  dfree(ctx, d_ext_tmp);
  dfree(ctx, d_ext);
  if (rc != P2B_OK) {
    dfree(ctx, d_coeffs);
    return rc;
  }
  // [challenge][chunk][n] coefficients = num_challenges * quotient_degree_factor polynomials of degree < n
  return batch_build(ctx, d_coeffs, false, (size_t)nch * d.quotient_degree_factor, d.degree_bits, rate_bits, cap_height, out);
}

extern "C" int p2b_quotient_commit(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const p2b_batch* wires,
                                   const p2b_batch* zs, const uint64_t* pi_hash, const uint64_t* betas,
                                   const uint64_t* gammas, const uint64_t* alphas, uint32_t rate_bits,
                                   uint32_t cap_height, p2b_batch** out) {
  CHECK_CTX(ctx);
  if (!out || !betas || !gammas || !alphas || !pi_hash || !zs || !c) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  const uint32_t nch = c->d.num_challenges;
  std::vector<uint64_t> h(3 * nch + 4);
  for (uint32_t i = 0; i < nch; i++) h[i] = betas[i], h[nch + i] = gammas[i], h[2 * nch + i] = alphas[i];
  for (int i = 0; i < 4; i++) h[3 * nch + i] = pi_hash[i] >= GL_P ? pi_hash[i] - GL_P : pi_hash[i];  // the kernels subtract it
  uint64_t* d_ch = nullptr;
  int rc = upload_felts(ctx, h.data(), h.size(), &d_ch);
  if (rc == P2B_OK)
    rc = quotient_core(ctx, c, cs, wires, zs, d_ch + 3 * nch, d_ch, d_ch + nch, d_ch + 2 * nch, rate_bits, cap_height, out);
  dfree(ctx, d_ch);
  return rc;
}

extern "C" int p2b_batch_leaf(p2b_batch* b, size_t leaf_index, uint64_t* out) {
  if (!b) return P2B_ERR_INVALID;
  return p2b_tree_leaf(&b->tree, leaf_index, out);
}
extern "C" int p2b_batch_lde_values(p2b_batch* b, size_t index, size_t step, uint64_t* out) {
  if (!b) return P2B_ERR_INVALID;
  uint32_t bits = b->log_n + b->rate_bits;
  size_t i = index * step;
  if (i >> bits) return fail(b->ctx, P2B_ERR_INVALID, "index*step %zu out of range", i);
  size_t r = 0;
  for (uint32_t k = 0; k < bits; k++) r |= ((i >> k) & 1) << (bits - 1 - k);
  return p2b_tree_leaf(&b->tree, r, out);
}
extern "C" int p2b_batch_leaves(p2b_batch* b, uint64_t* out) {
  if (!b || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  size_t N = (size_t)1 << (b->log_n + b->rate_bits);
  uint64_t* d_rm = nullptr;
  int rc = dmalloc(ctx, &d_rm, N * b->n_cols);
  if (rc) return rc;
  dim3 grid(cdiv(N, 32), cdiv(b->n_cols, 32));
  hashk::k_transpose_to_rowmajor<<<grid, 256, 0, ctx->stream>>>(b->d_lde, N, (uint32_t)b->n_cols, N, d_rm);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d_rm, N * b->n_cols);
  dfree(ctx, d_rm);
  return rc;
}

// ------------------------------------------------------------------------------------------------ MerkleTree
extern "C" int p2b_merkle_new(p2b_ctx* ctx, const uint64_t* leaves, size_t n_leaves, size_t leaf_len,
                              uint32_t cap_height, p2b_tree** out) {
  CHECK_CTX(ctx);
  if (!leaves || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n_leaves == 0 || (n_leaves & (n_leaves - 1))) return fail(ctx, P2B_ERR_INVALID, "n_leaves must be a power of two");
  uint32_t log_leaves = 0;
  while (((size_t)1 << log_leaves) < n_leaves) log_leaves++;
  if (cap_height > log_leaves)
    return fail(ctx, P2B_ERR_INVALID, "cap_height %u exceeds log2(#leaves) = %u", cap_height, log_leaves);
  p2b_tree* t = new (std::nothrow) p2b_tree();
  if (!t) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  t->ctx = ctx;
  t->n_leaves = n_leaves;
  t->log_leaves = log_leaves;
  t->cap_height = cap_height;
  t->leaf_len = leaf_len;
  int rc = dmalloc(ctx, &t->d_leaves_rm, n_leaves * leaf_len);
  if (rc == P2B_OK) rc = dmalloc(ctx, &t->d_levels, 4 * levels_len(n_leaves, cap_height));
  if (rc == P2B_OK && leaf_len) {
    cudaError_t e = cudaMemcpyAsync(t->d_leaves_rm, leaves, n_leaves * leaf_len * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                    ctx->stream);
    if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
  }
  if (rc == P2B_OK) {
    hashk::k_leaf_hash_rowmajor<<<cdiv(n_leaves, 256), 256, 0, ctx->stream>>>(t->d_leaves_rm, leaf_len, n_leaves,
                                                                            t->d_levels);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e));
  }
  if (rc == P2B_OK) rc = build_levels(ctx, t);
  if (rc != P2B_OK) {
    p2b_tree_free(t);
    return rc;
  }
  *out = t;
  return P2B_OK;
}

extern "C" void p2b_tree_free(p2b_tree* t) {
  if (!t || t->owned_by_batch) return;
  p2b_ctx* ctx = t->ctx;
  cudaSetDevice(ctx->device);
  dfree(ctx, t->d_levels);
  dfree(ctx, t->d_leaves_rm);
  delete t;
}
extern "C" size_t p2b_tree_n_leaves(const p2b_tree* t) { return t ? t->n_leaves : 0; }
extern "C" uint32_t p2b_tree_cap_height(const p2b_tree* t) { return t ? t->cap_height : 0; }

extern "C" int p2b_tree_cap(p2b_tree* t, uint64_t* out) {
  if (!t || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = t->ctx;
  CHECK_CTX(ctx);
  uint32_t L = t->log_leaves - t->cap_height;
  return d2h(ctx, out, t->d_levels + 4 * level_off(t->n_leaves, L), (size_t)4 << t->cap_height);
}

extern "C" int p2b_tree_prove(p2b_tree* t, size_t leaf_index, uint64_t* out) {
  if (!t || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = t->ctx;
  CHECK_CTX(ctx);
  if (leaf_index >= t->n_leaves) return fail(ctx, P2B_ERR_INVALID, "leaf index %zu out of range", leaf_index);
  uint32_t L = t->log_leaves - t->cap_height;
  if (L == 0) return P2B_OK;
  uint64_t* d_tmp = nullptr;
  int rc = dmalloc(ctx, &d_tmp, 4 * (size_t)L);
  if (rc) return rc;
  hashk::k_gather_proof<<<1, 4 * L <= 32 ? 32 : 4 * ((L + 7) / 8) * 8, 0, ctx->stream>>>(t->d_levels, t->n_leaves, L,
                                                                                         leaf_index, d_tmp);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d_tmp, 4 * (size_t)L);
  dfree(ctx, d_tmp);
  return rc;
}

extern "C" int p2b_tree_digests(p2b_tree* t, uint64_t* out) {
  if (!t || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = t->ctx;
  CHECK_CTX(ctx);
  uint32_t L = t->log_leaves - t->cap_height;
  size_t total = 2 * (t->n_leaves - ((size_t)1 << t->cap_height));
  if (total == 0) return P2B_OK;
  uint64_t* d_tmp = nullptr;
  int rc = dmalloc(ctx, &d_tmp, 4 * total);
  if (rc) return rc;
  hashk::k_export_plonky2_digests<<<cdiv(total, 256), 256, 0, ctx->stream>>>(t->d_levels, t->n_leaves, L, d_tmp);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d_tmp, 4 * total);
  dfree(ctx, d_tmp);
  return rc;
}

extern "C" int p2b_tree_leaf(p2b_tree* t, size_t leaf_index, uint64_t* out) {
  if (!t || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = t->ctx;
  CHECK_CTX(ctx);
  if (leaf_index >= t->n_leaves) return fail(ctx, P2B_ERR_INVALID, "leaf index %zu out of range", leaf_index);
  if (t->leaf_len == 0) return P2B_OK;
  if (t->d_leaves_rm) return d2h(ctx, out, t->d_leaves_rm + leaf_index * t->leaf_len, t->leaf_len);
  uint64_t* d_tmp = nullptr;
  int rc = dmalloc(ctx, &d_tmp, t->leaf_len);
  if (rc) return rc;
  hashk::k_gather_row_colmajor<<<cdiv(t->leaf_len, 128), 128, 0, ctx->stream>>>(t->d_leaves_cm, t->n_leaves,
                                                                               (uint32_t)t->leaf_len, leaf_index, d_tmp);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d_tmp, t->leaf_len);
  dfree(ctx, d_tmp);
  return rc;
}

// ------------------------------------------------------------------------------------------------ Poseidon
extern "C" int p2b_poseidon_permute(p2b_ctx* ctx, uint64_t* states, size_t n) {
  CHECK_CTX(ctx);
  if (!states) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (n == 0) return P2B_OK;
  uint64_t* d = nullptr;
  int rc = dmalloc(ctx, &d, 12 * n);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(d, states, 12 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  hashk::k_permute_states<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d, n);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, states, d, 12 * n);
  dfree(ctx, d);
  return rc;
}

extern "C" int p2b_hash_no_pad(p2b_ctx* ctx, const uint64_t* in, size_t len, uint64_t* out) {
  CHECK_CTX(ctx);
  if ((!in && len) || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  uint64_t* d = nullptr;
  int rc = dmalloc(ctx, &d, len + 4);
  if (rc) return rc;
  if (len) CU(ctx, cudaMemcpyAsync(d + 4, in, len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  hashk::k_hash_no_pad_single<<<1, 32, 0, ctx->stream>>>(d + 4, len, d);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d, 4);
  dfree(ctx, d);
  return rc;
}

extern "C" int p2b_two_to_one(p2b_ctx* ctx, const uint64_t* left, const uint64_t* right, size_t n, uint64_t* out) {
  CHECK_CTX(ctx);
  if (!left || !right || !out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (n == 0) return P2B_OK;
  uint64_t* d = nullptr;
  int rc = dmalloc(ctx, &d, 12 * n);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(d, left, 4 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d + 4 * n, right, 4 * n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  hashk::k_two_to_one_pairs<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d, d + 4 * n, n, d + 8 * n);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d + 8 * n, 4 * n);
  dfree(ctx, d);
  return rc;
}

// ------------------------------------------------------------------------------------------------ Challenger
extern "C" int p2b_challenger_new(p2b_ctx* ctx, p2b_challenger** out) {
  CHECK_CTX(ctx);
  if (!out) return P2B_ERR_INVALID;
  *out = nullptr;
  p2b_challenger* c = new (std::nothrow) p2b_challenger();
  if (!c) return fail(ctx, P2B_ERR_OOM, "host allocation failed");
  c->ctx = ctx;
  int rc = dmalloc(ctx, &c->d_state, frik::CH_WORDS);
  if (rc) {
    delete c;
    return rc;
  }
  CU(ctx, cudaMemsetAsync(c->d_state, 0, frik::CH_WORDS * sizeof(uint64_t), ctx->stream));
  *out = c;
  return P2B_OK;
}
extern "C" void p2b_challenger_free(p2b_challenger* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  dfree(c->ctx, c->d_state);
  delete c;
}

// observe n elements that already live on the device
static int challenger_observe_dev(p2b_challenger* c, const uint64_t* d_elems, size_t n) {
  p2b_ctx* ctx = c->ctx;
  if (n == 0) return P2B_OK;
  frik::k_challenger_observe<<<1, 32, 0, ctx->stream>>>(c->d_state, d_elems, n);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

extern "C" int p2b_challenger_observe(p2b_challenger* c, const uint64_t* elems, size_t n) {
  if (!c) return P2B_ERR_INVALID;
  p2b_ctx* ctx = c->ctx;
  CHECK_CTX(ctx);
  if (n == 0) return P2B_OK;
  if (!elems) return fail(ctx, P2B_ERR_INVALID, "null argument");
  uint64_t* d = nullptr;
  int rc = dmalloc(ctx, &d, n);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(d, elems, n * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  rc = challenger_observe_dev(c, d, n);
  // the H2D source may be pageable and reused by the caller right after we return
  if (rc == P2B_OK) CU(ctx, ctx_sync(ctx));
  dfree(ctx, d);
  return rc;
}

extern "C" int p2b_challenger_observe_cap(p2b_challenger* c, p2b_tree* t) {
  if (!c || !t) return P2B_ERR_INVALID;
  p2b_ctx* ctx = c->ctx;
  CHECK_CTX(ctx);
  if (t->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "tree belongs to a different context");
  uint32_t L = t->log_leaves - t->cap_height;
  return challenger_observe_dev(c, t->d_levels + 4 * level_off(t->n_leaves, L), (size_t)4 << t->cap_height);
}

extern "C" int p2b_challenger_get(p2b_challenger* c, size_t n, uint64_t* out) {
  if (!c || (!out && n)) return P2B_ERR_INVALID;
  p2b_ctx* ctx = c->ctx;
  CHECK_CTX(ctx);
  if (n == 0) return P2B_OK;
  uint64_t* d = nullptr;
  int rc = dmalloc(ctx, &d, n);
  if (rc) return rc;
  frik::k_challenger_get<<<1, 32, 0, ctx->stream>>>(c->d_state, n, d);
  LAUNCH_CHECK(ctx);
  rc = d2h(ctx, out, d, n);
  dfree(ctx, d);
  return rc;
}

extern "C" int p2b_challenger_export(p2b_challenger* c, uint64_t* out30) {
  if (!c || !out30) return P2B_ERR_INVALID;
  CHECK_CTX(c->ctx);
  return d2h(c->ctx, out30, c->d_state, frik::CH_WORDS);
}
extern "C" int p2b_challenger_import(p2b_challenger* c, const uint64_t* in30) {
  if (!c || !in30) return P2B_ERR_INVALID;
  p2b_ctx* ctx = c->ctx;
  CHECK_CTX(ctx);
  // Challenger invariant: observe() duplexes as soon as the input buffer holds RATE = 8 elements, so a stored
  // buffer is always shorter; 8 would make the proof-of-work candidate land in a capacity lane
  if (in30[12] >= 8 || in30[21] > 8) return fail(ctx, P2B_ERR_INVALID, "input buffer length must be < 8, output buffer length <= 8");
  CU(ctx, cudaMemcpyAsync(c->d_state, in30, frik::CH_WORDS * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, ctx_sync(ctx));
  return P2B_OK;
}

// ------------------------------------------------------------------------------------------------ FRI
// One transcript step in one launch (fusedk::k_transcript): observe the given device segments in order, then squeeze
// n_out challenges into d_out.
struct TrSegs {
  const uint64_t* p[fusedk::TR_MAX_SEGS];
  size_t n[fusedk::TR_MAX_SEGS];
  uint32_t count = 0;
  void add(const uint64_t* ptr, size_t len) {
    if (len == 0) return;
    p[count] = ptr;
    n[count] = len;
    count++;
  }
};
static int transcript_step(p2b_ctx* ctx, uint64_t* d_state, bool reset, const TrSegs& segs, uint64_t* d_out, uint32_t n_out,
                           uint64_t scale_g = 0, uint64_t* pow_tab = nullptr, uint32_t pow_n = 0, uint64_t* ext_tab = nullptr,
                           uint32_t ext_n = 0, uint32_t subgroup_check_bits = 0, uint64_t* subgroup_flag = nullptr) {
  fusedk::TranscriptParams tp{};
  tp.state = d_state;
  tp.n_seg = segs.count;
  for (uint32_t i = 0; i < segs.count; i++) {
    tp.seg[i] = segs.p[i];
    tp.seg_len[i] = (uint32_t)segs.n[i];
  }
  tp.reset = reset ? 1u : 0u;
  tp.out = d_out;
  tp.n_out = n_out;
  tp.scale_g = scale_g;
  tp.pow_tab = pow_tab;
  tp.pow_n = pow_n;
  tp.ext_tab = ext_tab;
  tp.ext_n = ext_n;
  tp.subgroup_check_bits = subgroup_check_bits;
  tp.subgroup_flag = subgroup_flag;
  fusedk::k_transcript<<<1, 32, 0, ctx->stream>>>(tp);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}
static const uint64_t* tree_cap_ptr(const p2b_tree* t) {
  return t->d_levels + 4 * level_off(t->n_leaves, t->log_leaves - t->cap_height);
}

// fri_committed_trees on device-resident inputs.
//   d_coef      coefficients as two planes (c0 | c1) of `len` each (read only; the folds ping-pong between two
//               scratch buffers)
//   d_vals_nat  layer-0 values, interleaved, natural order (bit-reversed here), or nullptr
//   d_vals_leaf layer-0 values as two planes of `len` in leaf (bit-reversed) order, or nullptr
//   d_final     receives the final polynomial, interleaved, 2 * (len >> sum(arity) >> rate_bits) words
// On success `trees` owns the layer trees; on failure they are freed.  The final polynomial is observed.
static int fri_commit_core(p2b_ctx* ctx, const uint64_t* d_coef, const uint64_t* d_vals_nat, const uint64_t* d_vals_leaf,
                           size_t len, uint32_t log_len, const uint32_t* arity_bits, size_t n_layers, uint32_t rate_bits,
                           uint32_t cap_height, p2b_challenger* ch, std::vector<p2b_tree*>& trees, uint64_t* d_final) {
  uint64_t *d_fold[2] = {nullptr, nullptr}, *d_beta = nullptr, *d_planes = nullptr;
  const size_t fold_words = n_layers ? 2 * (len >> arity_bits[0]) : 1;
  int rc = dmalloc(ctx, &d_fold[0], fold_words);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_fold[1], fold_words);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_beta, 2);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_planes, fold_words);
  auto cleanup = [&](int code) {
    dfree(ctx, d_fold[0]);
    dfree(ctx, d_fold[1]);
    dfree(ctx, d_beta);
    dfree(ctx, d_planes);
    if (code != P2B_OK) {
      for (p2b_tree* t : trees) p2b_tree_free(t);
      trees.clear();
    }
    return code;
  };
  if (rc) return cleanup(rc);
#define LAUNCHF()                                                                                  \
  do {                                                                                             \
    ctx->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return cleanup(fail(ctx, P2B_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)
  size_t cur = len;
  uint32_t log_cur = log_len;
  uint64_t shift = 7;
  const uint64_t* src = d_coef;  // planes src | src + cur
  for (size_t l = 0; l < n_layers; l++) {
    const uint32_t ab = arity_bits[l];
    const size_t n_leaves = cur >> ab, leaf_len = (size_t)2 << ab;
    if (cap_height > log_cur - ab)
      return cleanup(fail(ctx, P2B_ERR_INVALID, "cap_height %u exceeds layer %zu height %u", cap_height, l, log_cur - ab));
    p2b_tree* t = new (std::nothrow) p2b_tree();
    if (!t) return cleanup(fail(ctx, P2B_ERR_OOM, "host allocation failed"));
    trees.push_back(t);
    t->ctx = ctx;
    t->n_leaves = n_leaves;
    t->log_leaves = log_cur - ab;
    t->cap_height = cap_height;
    t->leaf_len = leaf_len;
    if ((rc = dmalloc(ctx, &t->d_leaves_rm, 2 * cur))) return cleanup(rc);
    if ((rc = dmalloc(ctx, &t->d_levels, 4 * levels_len(n_leaves, cap_height)))) return cleanup(rc);
    stage_begin(ctx, ST_FRI);
    if (l == 0 && d_vals_nat) {
      // natural order interleaved -> leaf order (bit-reversed), row-major leaves
      frik::k_bitrev_ext<<<cdiv(cur, 256), 256, 0, ctx->stream>>>(d_vals_nat, log_cur, t->d_leaves_rm);
      LAUNCHF();
      if (n_leaves > COOP_MAX_NODES)
        hashk::k_leaf_hash_rowmajor<<<cdiv(n_leaves, 256), 256, 0, ctx->stream>>>(t->d_leaves_rm, leaf_len, n_leaves, t->d_levels);
      else
        hashk::k_leaf_hash_rowmajor_coop<<<cdiv(n_leaves * 32, 256), 256, 0, ctx->stream>>>(t->d_leaves_rm, leaf_len, n_leaves, t->d_levels);
      LAUNCHF();
    } else {
      // planes in leaf order: the caller's LDE (layer 0) or the coset NTT of the folded coefficients
      const uint64_t* pl = l == 0 ? d_vals_leaf : d_planes;
      // one thread per leaf for layers of >= 1024 leaves (arity >= 4: whole permutations of 4 extension elements), one
      // warp per leaf below that: the cooperative permutation costs ~10x the instructions and only pays where a handful
      // of leaves would otherwise leave the GPU waiting on one 23 us permutation chain
      if (n_leaves >= 1024) {
        fusedk::k_leaf_hash_planes<<<cdiv(n_leaves, 256), 256, 0, ctx->stream>>>(pl, pl + cur, ab, n_leaves, t->d_leaves_rm, t->d_levels);
        LAUNCHF();
      } else {
        fusedk::k_leaf_hash_planes_coop<<<cdiv(n_leaves * 32, 256), 256, 0, ctx->stream>>>(pl, pl + cur, ab, n_leaves, t->d_leaves_rm,
                                                                                           t->d_levels);
        LAUNCHF();
      }
    }
    if ((rc = build_levels(ctx, t))) return cleanup(rc);
    // observe_cap, beta = get_extension_challenge (device resident), one launch
    {
      TrSegs sg;
      sg.add(tree_cap_ptr(t), (size_t)4 << cap_height);
      if ((rc = transcript_step(ctx, ch->d_state, false, sg, d_beta, 2))) return cleanup(rc);
    }
    // fold: coeffs[i] = sum_j coeffs[i*arity + j] * beta^j, into the other scratch buffer
    uint64_t* dst = d_fold[l & 1];
    frik::k_fold_coeffs<<<cdiv(n_leaves, 256), 256, 0, ctx->stream>>>(src, src + cur, cur, ab, d_beta, dst, dst + n_leaves);
    LAUNCHF();
    src = dst;
    cur = n_leaves;
    log_cur -= ab;
    for (uint32_t k = 0; k < ab; k++) shift = h_mulmod(shift, shift);
    if (l + 1 < n_layers) {
      // values of the next layer = coset NTT (shift) of the folded coefficients, leaf order, 2 planes
      if ((rc = run_lde(ctx, src, cur, d_planes, 2, log_cur, 0, shift))) return cleanup(rc);
    }
    stage_end(ctx);
  }
  // final polynomial: first cur >> rate_bits coefficients (the rest are zero for a valid codeword)
  size_t n_final = cur >> rate_bits;
  frik::k_interleave<<<cdiv(n_final, 256), 256, 0, ctx->stream>>>(src, src + cur, n_final, d_final);
  LAUNCHF();
  {
    TrSegs sg;
    sg.add(d_final, 2 * n_final);
    if ((rc = transcript_step(ctx, ch->d_state, false, sg, nullptr, 0))) return cleanup(rc);
  }
  return cleanup(P2B_OK);
#undef LAUNCHF
}

static int check_fri_args(p2b_ctx* ctx, size_t len, const uint32_t* arity_bits, size_t n_layers, uint32_t rate_bits,
                          uint32_t* log_len_out) {
  if (len == 0 || (len & (len - 1))) return fail(ctx, P2B_ERR_INVALID, "len must be a power of two");
  uint32_t log_len = 0;
  while (((size_t)1 << log_len) < len) log_len++;
  uint32_t sum = 0;
  for (size_t l = 0; l < n_layers; l++) {
    if (arity_bits[l] == 0 || arity_bits[l] > 6) return fail(ctx, P2B_ERR_UNSUPPORTED, "arity_bits must be in 1..6");
    sum += arity_bits[l];
  }
  if (sum + rate_bits > log_len) return fail(ctx, P2B_ERR_INVALID, "reduction exceeds the polynomial length");
  if (log_len > 24) return fail(ctx, P2B_ERR_UNSUPPORTED, "len > 2^24");
  *log_len_out = log_len;
  return P2B_OK;
}

extern "C" int p2b_fri_commit(p2b_ctx* ctx, const uint64_t* coeffs_ext, const uint64_t* values_ext, size_t len,
                              const uint32_t* arity_bits, size_t n_layers, uint32_t rate_bits, uint32_t cap_height,
                              p2b_challenger* ch, p2b_tree** layers_out, uint64_t* final_poly_out) {
  CHECK_CTX(ctx);
  if (!coeffs_ext || !values_ext || !ch || !final_poly_out || (n_layers && (!arity_bits || !layers_out)))
    return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ch->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "challenger belongs to a different context");
  uint32_t log_len = 0;
  int rc = check_fri_args(ctx, len, arity_bits, n_layers, rate_bits, &log_len);
  if (rc) return rc;
  uint32_t sum = 0;
  for (size_t l = 0; l < n_layers; l++) {
    sum += arity_bits[l];
    layers_out[l] = nullptr;
  }
  const size_t n_final = (len >> sum) >> rate_bits;
  uint64_t *d_in = nullptr, *d_coef = nullptr, *d_vals = nullptr, *d_final = nullptr;
  rc = dmalloc(ctx, &d_in, 2 * len);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_coef, 2 * len);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_vals, 2 * len);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_final, 2 * n_final);
  std::vector<p2b_tree*> trees;
  if (rc == P2B_OK) {
    cudaError_t e = cudaMemcpyAsync(d_in, coeffs_ext, 2 * len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
      frik::k_deinterleave<<<cdiv(len, 256), 256, 0, ctx->stream>>>(d_in, len, d_coef, d_coef + len);
      ctx->launches++;
      e = cudaMemcpyAsync(d_vals, values_ext, 2 * len * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) rc = fail(ctx, P2B_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
  }
  if (rc == P2B_OK)
    rc = fri_commit_core(ctx, d_coef, d_vals, nullptr, len, log_len, arity_bits, n_layers, rate_bits, cap_height, ch, trees,
                         d_final);
  if (rc == P2B_OK) {
    cudaError_t e = cudaMemcpyAsync(final_poly_out, d_final, 2 * n_final * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = ctx_sync(ctx);
    if (e != cudaSuccess) {
      rc = fail(ctx, P2B_ERR_CUDA, "download: %s", cudaGetErrorString(e));
      for (p2b_tree* t : trees) p2b_tree_free(t);
      trees.clear();
    }
  }
  dfree(ctx, d_in);
  dfree(ctx, d_coef);
  dfree(ctx, d_vals);
  dfree(ctx, d_final);
  if (rc == P2B_OK)
    for (size_t l = 0; l < n_layers; l++) layers_out[l] = trees[l];
  return rc;
}

// fri_proof_of_work without a host round trip: ONE search launch (every thread walks its candidates in increasing
// order and leaves as soon as its candidate exceeds the best witness found so far, so the launch ends one grid stride
// after the first hit however large the range is), then fusedk::k_pow_finish observes the witness, squeezes the
// response and the n_queries query-index challenges (d_chal; may be null with n_queries == 0), and stores the witness
// at d_witness_out.
static int fri_pow_dev(p2b_ctx* ctx, p2b_challenger* ch, uint32_t pow_bits, uint64_t* d_witness_out, uint32_t n_queries,
                       uint64_t* d_chal) {
  if (pow_bits > 40) return fail(ctx, P2B_ERR_UNSUPPORTED, "pow_bits > 40");
  unsigned long long* d_best = (unsigned long long*)(ctx->d_scratch + 1);
  // The minimal witness is geometric with mean 2^pow_bits.  Every thread of a stride evaluates its candidate even when
  // the witness sits at the start of the stride (half a stride of wasted permutations on average), and the stride
  // after the hit has usually started before the hit is visible (k_pow_search drops it a quarter of the way in): about
  // 0.8 strides of waste.  With many proofs in flight only the permutation count matters, so the stride is an eighth
  // of the mean (32 CTAs at 16 bits: ~9 strides of 23 us, ~1.1x the necessary permutations; 64 CTAs without the
  // in-flight check did 1.37x, 148 CTAs ~2x), between 16 CTAs and one CTA per SM.  P2B_POW_BLOCKS overrides.
  const uint64_t sms = (uint64_t)ctx->sm_count;
  static const int force_blocks = [] {
    const char* e = getenv("P2B_POW_BLOCKS");
    return e ? atoi(e) : 0;
  }();
  uint64_t blocks = (((uint64_t)1 << pow_bits) / 8 + 255) / 256;
  blocks = blocks < 16 ? 16 : blocks > sms ? sms : blocks;
  if (ctx->latency_mode) blocks = sms;
  if (force_blocks > 0) blocks = (uint64_t)force_blocks;
  CU(ctx, cudaMemsetAsync(d_best, 0xFF, sizeof(uint64_t), ctx->stream));
  frik::k_pow_search<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ch->d_state, 0, GL_P, pow_bits, d_best);
  LAUNCH_CHECK(ctx);
  fusedk::k_pow_finish<<<1, 32, 0, ctx->stream>>>(ch->d_state, d_best, d_witness_out, n_queries, d_chal);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

extern "C" int p2b_fri_pow(p2b_ctx* ctx, p2b_challenger* ch, uint32_t pow_bits, uint64_t* witness_out) {
  CHECK_CTX(ctx);
  if (!ch || !witness_out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ch->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "challenger belongs to a different context");
  uint64_t* d_w = nullptr;
  int rc = dmalloc(ctx, &d_w, 1);
  if (rc) return rc;
  rc = fri_pow_dev(ctx, ch, pow_bits, d_w, 0, nullptr);
  if (rc == P2B_OK) rc = d2h(ctx, witness_out, d_w, 1);
  dfree(ctx, d_w);
  if (rc == P2B_OK && *witness_out == ~0ull) return fail(ctx, P2B_ERR_INVALID, "no proof-of-work witness found");
  return rc;
}

// ------------------------------------------------------------------------------------------------ openings + FRI proof
// d_point: extension point on the device; d_out: 2 * count words on the device
static int eval_ext_core(p2b_ctx* ctx, const p2b_batch* b, const uint64_t* d_point, size_t first, size_t count, uint64_t* d_out) {
  if (count == 0) return P2B_OK;
  const size_t n = (size_t)1 << b->log_n;
  provk::k_eval_polys_ext<<<(unsigned)count, 256, 0, ctx->stream>>>(b->d_coeffs + first * n, n, d_point, d_out);
  LAUNCH_CHECK(ctx);
  return P2B_OK;
}

extern "C" int p2b_batch_eval_ext(p2b_batch* b, const uint64_t* point, size_t first, size_t count, uint64_t* out) {
  if (!b || !point || !out) return P2B_ERR_INVALID;
  p2b_ctx* ctx = b->ctx;
  CHECK_CTX(ctx);
  if (first + count > b->n_cols || first + count < first) return fail(ctx, P2B_ERR_INVALID, "polynomial range out of bounds");
  if (count == 0) return P2B_OK;
  uint64_t *d_out = nullptr, *d_pt = nullptr;
  int rc = upload_felts(ctx, point, 2, &d_pt);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_out, 2 * count);
  if (rc == P2B_OK) rc = eval_ext_core(ctx, b, d_pt, first, count, d_out);
  if (rc == P2B_OK) rc = d2h(ctx, out, d_out, 2 * count);
  dfree(ctx, d_out);
  dfree(ctx, d_pt);
  return rc;
}

static int check_fri_params(p2b_ctx* ctx, const p2b_batch* const* oracles, size_t n_oracles, const p2b_fri_params* fp) {
  if (!oracles || !fp || n_oracles == 0 || n_oracles > 16) return fail(ctx, P2B_ERR_INVALID, "bad oracle list");
  if (fp->n_layers > P2B_MAX_FRI_LAYERS) return fail(ctx, P2B_ERR_INVALID, "too many FRI layers");
  for (size_t o = 0; o < n_oracles; o++) {
    if (!oracles[o]) return fail(ctx, P2B_ERR_INVALID, "null oracle");
    if (oracles[o]->log_n != oracles[0]->log_n || oracles[o]->rate_bits != fp->rate_bits)
      return fail(ctx, P2B_ERR_INVALID, "oracle %zu: degree / rate_bits mismatch", o);
  }
  return P2B_OK;
}

static size_t fri_proof_len_impl(const p2b_batch* const* oracles, size_t n_oracles, const p2b_fri_params* fp) {
  const uint32_t log_N = oracles[0]->log_n + fp->rate_bits;
  size_t per_query = 0;
  for (size_t o = 0; o < n_oracles; o++)
    per_query += oracles[o]->n_cols + 4 * (size_t)(log_N - oracles[o]->cap_height);
  uint32_t log_cur = log_N;
  for (uint32_t l = 0; l < fp->n_layers; l++) {
    const uint32_t ab = fp->reduction_arity_bits[l];
    log_cur -= ab;
    per_query += ((size_t)2 << ab) + 4 * (size_t)(log_cur - fp->cap_height);
  }
  const size_t n_final = ((size_t)1 << log_cur) >> fp->rate_bits;
  return (size_t)fp->n_layers * ((size_t)4 << fp->cap_height) + (size_t)fp->num_query_rounds * per_query + 2 * n_final + 1;
}

extern "C" size_t p2b_fri_proof_len(const p2b_batch* const* oracles, size_t n_oracles, const p2b_fri_params* fp) {
  if (!oracles || !fp || n_oracles == 0) return 0;
  uint32_t sum = 0;
  for (uint32_t l = 0; l < fp->n_layers && l < P2B_MAX_FRI_LAYERS; l++) sum += fp->reduction_arity_bits[l];
  if (fp->n_layers > P2B_MAX_FRI_LAYERS || sum + fp->rate_bits > oracles[0]->log_n + fp->rate_bits) return 0;
  return fri_proof_len_impl(oracles, n_oracles, fp);
}

// The general form (any number of opening batches): one launch per step, as in round 1.
// d_points: n_batches extension points on the device (2 words each); d_proof: fri_proof_len words on the device.
static int prove_openings_generic(p2b_ctx* ctx, const p2b_batch* const* oracles, size_t n_oracles,
                               const p2b_fri_batch* batches, size_t n_batches, const uint64_t* d_points,
                               p2b_challenger* ch, const p2b_fri_params* fp, uint64_t* d_proof, const TrSegs* pre) {
  if (!batches || !ch || !d_proof || n_batches == 0) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ch->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "challenger belongs to a different context");
  int rc = check_fri_params(ctx, oracles, n_oracles, fp);
  if (rc) return rc;
  const uint32_t log_n = oracles[0]->log_n, rate_bits = fp->rate_bits, log_N = log_n + rate_bits;
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  uint32_t log_len = 0;
  if ((rc = check_fri_args(ctx, N, fp->reduction_arity_bits, fp->n_layers, rate_bits, &log_len))) return rc;
  for (size_t o = 0; o < n_oracles; o++)
    if (oracles[o]->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "oracle of another context");
  const size_t proof_len = fri_proof_len_impl(oracles, n_oracles, fp);
  // polynomial pointer tables of every batch
  std::vector<std::vector<const uint64_t*>> tabs(n_batches);
  size_t max_m = 0;
  for (size_t bi = 0; bi < n_batches; bi++) {
    const p2b_fri_batch& fb = batches[bi];
    if (fb.n_ranges == 0 || fb.n_ranges > P2B_MAX_FRI_RANGES) return fail(ctx, P2B_ERR_INVALID, "batch %zu: bad range count", bi);
    for (uint32_t r = 0; r < fb.n_ranges; r++) {
      const auto& rg = fb.ranges[r];
      if (rg.oracle >= n_oracles || (size_t)rg.first + rg.count > oracles[rg.oracle]->n_cols)
        return fail(ctx, P2B_ERR_INVALID, "batch %zu range %u out of bounds", bi, r);
      for (uint32_t k = 0; k < rg.count; k++) tabs[bi].push_back(oracles[rg.oracle]->d_coeffs + (size_t)(rg.first + k) * n);
    }
    if (tabs[bi].size() > max_m) max_m = tabs[bi].size();
  }
  uint32_t sum_ab = 0;
  for (uint32_t l = 0; l < fp->n_layers; l++) sum_ab += fp->reduction_arity_bits[l];
  const size_t n_final = (N >> sum_ab) >> rate_bits;

  uint64_t *d_alpha = nullptr, *d_pw = nullptr, *d_ptrs = nullptr, *d_comp = nullptr, *d_quot = nullptr, *d_fin = nullptr;
  uint64_t *d_coef = nullptr, *d_vals = nullptr, *d_final = nullptr, *d_chal = nullptr;
  std::vector<p2b_tree*> trees;
  auto cleanup = [&](int code) {
    for (uint64_t* p : {d_alpha, d_pw, d_ptrs, d_comp, d_quot, d_fin, d_coef, d_vals, d_final, d_chal}) dfree(ctx, p);
    for (p2b_tree* t : trees) p2b_tree_free(t);
    return code;
  };
#define TRY(expr)                          \
  do {                                     \
    int rc__ = (expr);                     \
    if (rc__ != P2B_OK) return cleanup(rc__); \
  } while (0)
#define CUP(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return cleanup(fail(ctx, P2B_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)
#define LAUNCHP()                                                                                  \
  do {                                                                                             \
    ctx->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return cleanup(fail(ctx, P2B_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)
  TRY(dmalloc(ctx, &d_alpha, 2));
  TRY(dmalloc(ctx, &d_pw, 2 * (max_m + 1)));
  TRY(dmalloc(ctx, &d_ptrs, max_m));
  TRY(dmalloc(ctx, &d_comp, 2 * n));
  TRY(dmalloc(ctx, &d_quot, 2 * n));
  TRY(dmalloc(ctx, &d_fin, 2 * n));
  TRY(dmalloc(ctx, &d_coef, 2 * N));
  TRY(dmalloc(ctx, &d_vals, 2 * N));
  TRY(dmalloc(ctx, &d_final, 2 * (n_final ? n_final : 1)));
  TRY(dmalloc(ctx, &d_chal, fp->num_query_rounds ? fp->num_query_rounds : 1));

  stage_begin(ctx, ST_OTHER);
  // alpha = challenger.get_extension_challenge()
  {
    TrSegs none;
    TRY(transcript_step(ctx, ch->d_state, false, pre ? *pre : none, d_alpha, 2));
  }
  CUP(cudaMemsetAsync(d_fin, 0, 2 * n * sizeof(uint64_t), ctx->stream));
  for (size_t bi = 0; bi < n_batches; bi++) {
    const uint32_t m = (uint32_t)tabs[bi].size();
    // the table is pageable host memory: cudaMemcpyAsync stages it before returning
    CUP(cudaMemcpyAsync(d_ptrs, tabs[bi].data(), m * sizeof(uint64_t*), cudaMemcpyHostToDevice, ctx->stream));
    provk::k_ext_powers<<<1, 32, 0, ctx->stream>>>(d_alpha, m, d_pw);
    LAUNCHP();
    provk::k_reduce_polys<<<cdiv(n, 256), 256, 0, ctx->stream>>>((const uint64_t* const*)d_ptrs, m, n, d_pw, d_comp, d_comp + n);
    LAUNCHP();
    provk::k_divide_by_linear<<<1, 1024, 0, ctx->stream>>>(d_comp, d_comp + n, n, d_points + 2 * bi, d_quot, d_quot + n);
    LAUNCHP();
    provk::k_shift_add<<<cdiv(n, 256), 256, 0, ctx->stream>>>(d_fin, d_fin + n, d_quot, d_quot + n, n, d_pw + 2 * m);
    LAUNCHP();
  }
  // lde_final_poly = final_poly.lde(rate_bits) (zero padded coefficient planes); lde_final_values = coset_fft(7)
  CUP(cudaMemsetAsync(d_coef, 0, 2 * N * sizeof(uint64_t), ctx->stream));
  CUP(cudaMemcpyAsync(d_coef, d_fin, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  CUP(cudaMemcpyAsync(d_coef + N, d_fin + n, n * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  stage_begin(ctx, ST_LDE);
  TRY(run_lde(ctx, d_fin, n, d_vals, 2, log_n, rate_bits, 7));
  stage_end(ctx);
  // fri_proof: commit phase
  TRY(fri_commit_core(ctx, d_coef, nullptr, d_vals, N, log_N, fp->reduction_arity_bits, fp->n_layers, rate_bits,
                      fp->cap_height, ch, trees, d_final));
  TRY(fri_pow_dev(ctx, ch, fp->proof_of_work_bits, d_proof + (proof_len - 1), fp->num_query_rounds, d_chal));
  // query rounds: indices squeezed on the device, then one gather kernel per (oracle | layer, leaf | path)
  stage_begin(ctx, ST_OTHER);
  const uint32_t nq = fp->num_query_rounds;
  size_t off = 0;
  for (uint32_t l = 0; l < fp->n_layers; l++) {
    CUP(cudaMemcpyAsync(d_proof + off, trees[l]->d_levels + 4 * level_off(trees[l]->n_leaves, trees[l]->log_leaves - trees[l]->cap_height),
                        ((size_t)4 << fp->cap_height) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    off += (size_t)4 << fp->cap_height;
  }
  size_t per_query = 0;
  for (size_t o = 0; o < n_oracles; o++) per_query += oracles[o]->n_cols + 4 * (size_t)(log_N - oracles[o]->cap_height);
  {
    uint32_t lc = log_N;
    for (uint32_t l = 0; l < fp->n_layers; l++) {
      lc -= fp->reduction_arity_bits[l];
      per_query += ((size_t)2 << fp->reduction_arity_bits[l]) + 4 * (size_t)(lc - fp->cap_height);
    }
  }
  if (nq) {
    size_t qoff = off;
    for (size_t o = 0; o < n_oracles; o++) {
      const p2b_batch* b = oracles[o];
      const uint32_t L = log_N - b->cap_height;
      provk::k_query_leaf_colmajor<<<nq, 128, 0, ctx->stream>>>(b->d_lde, N, (uint32_t)b->n_cols, d_chal, log_N, d_proof + qoff, per_query);
      LAUNCHP();
      qoff += b->n_cols;
      provk::k_query_siblings<<<nq, 128, 0, ctx->stream>>>(b->tree.d_levels, N, L, d_chal, log_N, 0, d_proof + qoff, per_query);
      LAUNCHP();
      qoff += 4 * (size_t)L;
    }
    uint32_t shift = 0;
    for (uint32_t l = 0; l < fp->n_layers; l++) {
      const uint32_t ab = fp->reduction_arity_bits[l];
      shift += ab;
      const p2b_tree* t = trees[l];
      const uint32_t L = t->log_leaves - t->cap_height;
      provk::k_query_leaf_rowmajor<<<nq, 128, 0, ctx->stream>>>(t->d_leaves_rm, (uint32_t)t->leaf_len, d_chal, log_N, shift, d_proof + qoff, per_query);
      LAUNCHP();
      qoff += t->leaf_len;
      provk::k_query_siblings<<<nq, 128, 0, ctx->stream>>>(t->d_levels, t->n_leaves, L, d_chal, log_N, shift, d_proof + qoff, per_query);
      LAUNCHP();
      qoff += 4 * (size_t)L;
    }
  }
  off += (size_t)nq * per_query;
  CUP(cudaMemcpyAsync(d_proof + off, d_final, 2 * n_final * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
  off += 2 * n_final;
  off += 1;  // pow_witness: written by k_pow_finish
  stage_end(ctx);
  return cleanup(off == proof_len ? P2B_OK : fail(ctx, P2B_ERR_INVALID, "internal: proof length mismatch"));
#undef TRY
#undef CUP
#undef LAUNCHP
}

// PolynomialBatch::prove_openings on device-resident oracles.
// d_points: n_batches extension points on the device (2 words each); d_proof: fri_proof_len words on the device;
// pre: segments to observe before alpha is squeezed (p2b_prove hands over the openings here so that "observe the
// openings, squeeze alpha, build the alpha powers" is one launch).  No host synchronisation inside.
// The two-batch instance of every plonky2 circuit without lookups (everything at zeta, the Zs again at g zeta) takes
// the fused path below; anything else goes through prove_openings_generic.
static int prove_openings_core(p2b_ctx* ctx, const p2b_batch* const* oracles, size_t n_oracles,
                               const p2b_fri_batch* batches, size_t n_batches, const uint64_t* d_points,
                               p2b_challenger* ch, const p2b_fri_params* fp, uint64_t* d_proof, const TrSegs* pre = nullptr) {
  if (!batches || !ch || !d_proof || n_batches == 0) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (n_batches != 2 || n_oracles > (size_t)fusedk::Q_MAX_ORACLES)
    return prove_openings_generic(ctx, oracles, n_oracles, batches, n_batches, d_points, ch, fp, d_proof, pre);
  if (ch->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "challenger belongs to a different context");
  int rc = check_fri_params(ctx, oracles, n_oracles, fp);
  if (rc) return rc;
  const uint32_t log_n = oracles[0]->log_n, rate_bits = fp->rate_bits, log_N = log_n + rate_bits;
  const size_t n = (size_t)1 << log_n, N = n << rate_bits;
  uint32_t log_len = 0;
  if ((rc = check_fri_args(ctx, N, fp->reduction_arity_bits, fp->n_layers, rate_bits, &log_len))) return rc;
  for (size_t o = 0; o < n_oracles; o++)
    if (oracles[o]->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "oracle of another context");
  const size_t proof_len = fri_proof_len_impl(oracles, n_oracles, fp);
  // polynomial pointer tables of the two batches, back to back
  std::vector<const uint64_t*> tab;
  size_t m[2] = {0, 0};
  for (size_t bi = 0; bi < 2; bi++) {
    const p2b_fri_batch& fb = batches[bi];
    if (fb.n_ranges == 0 || fb.n_ranges > P2B_MAX_FRI_RANGES) return fail(ctx, P2B_ERR_INVALID, "batch %zu: bad range count", bi);
    for (uint32_t r = 0; r < fb.n_ranges; r++) {
      const auto& rg = fb.ranges[r];
      if (rg.oracle >= n_oracles || (size_t)rg.first + rg.count > oracles[rg.oracle]->n_cols)
        return fail(ctx, P2B_ERR_INVALID, "batch %zu range %u out of bounds", bi, r);
      for (uint32_t k = 0; k < rg.count; k++) tab.push_back(oracles[rg.oracle]->d_coeffs + (size_t)(rg.first + k) * n);
      m[bi] += rg.count;
    }
  }
  const size_t max_m = m[0] > m[1] ? m[0] : m[1];
  uint32_t sum_ab = 0;
  for (uint32_t l = 0; l < fp->n_layers; l++) sum_ab += fp->reduction_arity_bits[l];
  const size_t n_final = (N >> sum_ab) >> rate_bits;
  const size_t row_ctas = (n + 255) / 256;
  uint32_t slices = (uint32_t)((2 * (size_t)ctx->sm_count + row_ctas - 1) / row_ctas);
  slices = slices < 1 ? 1 : slices > 16 ? 16 : slices;
  if (slices > m[0]) slices = m[0] ? (uint32_t)m[0] : 1;

  uint64_t *d_alpha = nullptr, *d_pw = nullptr, *d_ptrs = nullptr, *d_part = nullptr, *d_comp = nullptr, *d_quot = nullptr;
  uint64_t *d_fin = nullptr, *d_coef = nullptr, *d_vals = nullptr, *d_final = nullptr, *d_chal = nullptr;
  std::vector<p2b_tree*> trees;
  auto cleanup = [&](int code) {
    for (uint64_t* q : {d_alpha, d_pw, d_ptrs, d_part, d_comp, d_quot, d_fin, d_coef, d_vals, d_final, d_chal}) dfree(ctx, q);
    for (p2b_tree* t : trees) p2b_tree_free(t);
    return code;
  };
#define TRY(expr)                          \
  do {                                     \
    int rc__ = (expr);                     \
    if (rc__ != P2B_OK) return cleanup(rc__); \
  } while (0)
#define LAUNCHP()                                                                                  \
  do {                                                                                             \
    ctx->launches++;                                                                               \
    cudaError_t e__ = cudaGetLastError();                                                          \
    if (e__ != cudaSuccess)                                                                        \
      return cleanup(fail(ctx, P2B_ERR_CUDA, "kernel launch: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)
  TRY(dmalloc(ctx, &d_alpha, 2));
  TRY(dmalloc(ctx, &d_pw, 2 * (max_m + 1)));
  TRY(dmalloc(ctx, &d_ptrs, tab.size()));
  TRY(dmalloc(ctx, &d_part, (size_t)slices * 2 * n));
  TRY(dmalloc(ctx, &d_comp, 4 * n));  // batch 1's reduction | batch 0's summed reduction
  TRY(dmalloc(ctx, &d_quot, 4 * n));
  TRY(dmalloc(ctx, &d_fin, 2 * n));
  TRY(dmalloc(ctx, &d_coef, 2 * N));
  TRY(dmalloc(ctx, &d_vals, 2 * N));
  TRY(dmalloc(ctx, &d_final, 2 * (n_final ? n_final : 1)));
  TRY(dmalloc(ctx, &d_chal, fp->num_query_rounds ? fp->num_query_rounds : 1));

  stage_begin(ctx, ST_OTHER);
  TRY(h2d_small(ctx, d_ptrs, tab.data(), tab.size()));
  // [observe the openings;] alpha = challenger.get_extension_challenge(); alpha^0 .. alpha^max_m
  {
    TrSegs none;
    TRY(transcript_step(ctx, ch->d_state, false, pre ? *pre : none, d_alpha, 2, 0, nullptr, 0, d_pw, (uint32_t)max_m));
  }
  // final_poly = sum over the batches of alpha-shifted (reduce_polys_base(batch) / (X - point))
  fusedk::k_reduce_polys2<<<dim3((unsigned)row_ctas, slices + 1), 256, 0, ctx->stream>>>(
      (const uint64_t* const*)d_ptrs, (uint32_t)m[0], (const uint64_t* const*)(d_ptrs + m[0]), (uint32_t)m[1], n, d_pw, slices, d_part,
      d_comp);
  LAUNCHP();
  fusedk::k_divide_by_linear2<<<2, 1024, 0, ctx->stream>>>(d_part, slices, d_comp, n, d_points, d_quot, d_comp + 2 * n);
  LAUNCHP();
  // lde_final_poly = final_poly.lde(rate_bits) (zero padded coefficient planes); lde_final_values = coset_fft(7)
  fusedk::k_final_poly_combine<<<cdiv(N, 256), 256, 0, ctx->stream>>>(d_quot, n, N, d_pw + 2 * m[1], d_fin, d_coef);
  LAUNCHP();
  stage_begin(ctx, ST_LDE);
  TRY(run_lde(ctx, d_fin, n, d_vals, 2, log_n, rate_bits, 7));
  stage_end(ctx);
  // fri_proof: commit phase, proof of work, query indices
  TRY(fri_commit_core(ctx, d_coef, nullptr, d_vals, N, log_N, fp->reduction_arity_bits, fp->n_layers, rate_bits,
                      fp->cap_height, ch, trees, d_final));
  const uint32_t nq = fp->num_query_rounds;
  TRY(fri_pow_dev(ctx, ch, fp->proof_of_work_bits, d_proof + (proof_len - 1), nq, d_chal));
  // proof words: commit-phase caps | query rounds | final polynomial | pow witness
  stage_begin(ctx, ST_OTHER);
  const size_t cap_words = (size_t)4 << fp->cap_height;
  fusedk::QueryParams qp{};
  size_t per_query = 0;
  for (size_t o = 0; o < n_oracles; o++) {
    const p2b_batch* b = oracles[o];
    qp.o_data[o] = b->d_lde;
    qp.o_levels[o] = b->tree.d_levels;
    qp.o_cols[o] = (uint32_t)b->n_cols;
    qp.o_L[o] = log_N - b->cap_height;
    qp.o_off[o] = (uint32_t)per_query;
    per_query += b->n_cols + 4 * (size_t)(log_N - b->cap_height);
  }
  {
    uint32_t shift = 0;
    for (uint32_t l = 0; l < fp->n_layers; l++) {
      const p2b_tree* t = trees[l];
      shift += fp->reduction_arity_bits[l];
      qp.l_leaves[l] = t->d_leaves_rm;
      qp.l_levels[l] = t->d_levels;
      qp.l_leaf_len[l] = (uint32_t)t->leaf_len;
      qp.l_L[l] = t->log_leaves - t->cap_height;
      qp.l_shift[l] = shift;
      qp.l_log_leaves[l] = t->log_leaves;
      qp.l_off[l] = (uint32_t)per_query;
      per_query += t->leaf_len + 4 * (size_t)(t->log_leaves - t->cap_height);
    }
  }
  size_t off = 0;
  fusedk::CopyParams cp{};
  for (uint32_t l = 0; l < fp->n_layers; l++) {
    cp.src[cp.n] = tree_cap_ptr(trees[l]);
    cp.dst[cp.n] = d_proof + off;
    cp.len[cp.n] = (uint32_t)cap_words;
    cp.n++;
    off += cap_words;
  }
  if (nq) {
    qp.n_oracles = (uint32_t)n_oracles;
    qp.n_layers = fp->n_layers;
    qp.log_lde = log_N;
    qp.N = N;
    qp.per_query = per_query;
    qp.chal = d_chal;
    qp.out = d_proof + off;
    fusedk::k_query_all<<<dim3(nq, (unsigned)(n_oracles + fp->n_layers)), 128, 0, ctx->stream>>>(qp);
    LAUNCHP();
  }
  off += (size_t)nq * per_query;
  cp.src[cp.n] = d_final;
  cp.dst[cp.n] = d_proof + off;
  cp.len[cp.n] = (uint32_t)(2 * n_final);
  cp.n++;
  off += 2 * n_final + 1;  // + pow_witness, written by k_pow_finish
  fusedk::k_copy_multi<<<cp.n, 256, 0, ctx->stream>>>(cp);
  LAUNCHP();
  stage_end(ctx);
  return cleanup(off == proof_len ? P2B_OK : fail(ctx, P2B_ERR_INVALID, "internal: proof length mismatch"));
#undef TRY
#undef LAUNCHP
}

extern "C" int p2b_prove_openings(p2b_ctx* ctx, const p2b_batch* const* oracles, size_t n_oracles,
                                  const p2b_fri_batch* batches, size_t n_batches, p2b_challenger* ch,
                                  const p2b_fri_params* fp, uint64_t* proof_out, size_t proof_cap) {
  CHECK_CTX(ctx);
  if (!batches || !ch || !proof_out || n_batches == 0) return fail(ctx, P2B_ERR_INVALID, "null argument");
  int rc = check_fri_params(ctx, oracles, n_oracles, fp);
  if (rc) return rc;
  uint32_t log_len = 0;
  if ((rc = check_fri_args(ctx, ((size_t)1 << oracles[0]->log_n) << fp->rate_bits, fp->reduction_arity_bits, fp->n_layers,
                           fp->rate_bits, &log_len)))
    return rc;
  const size_t proof_len = fri_proof_len_impl(oracles, n_oracles, fp);
  if (proof_cap < proof_len) return fail(ctx, P2B_ERR_INVALID, "proof buffer too small: %zu < %zu words", proof_cap, proof_len);
  std::vector<uint64_t> pts(2 * n_batches);
  for (size_t i = 0; i < n_batches; i++) pts[2 * i] = batches[i].point[0], pts[2 * i + 1] = batches[i].point[1];
  uint64_t *d_pts = nullptr, *d_proof = nullptr;
  rc = upload_felts(ctx, pts.data(), pts.size(), &d_pts);
  if (rc == P2B_OK) rc = dmalloc(ctx, &d_proof, proof_len);
  if (rc == P2B_OK) rc = prove_openings_core(ctx, oracles, n_oracles, batches, n_batches, d_pts, ch, fp, d_proof);
  if (rc == P2B_OK) rc = d2h(ctx, proof_out, d_proof, proof_len);
  dfree(ctx, d_pts);
  dfree(ctx, d_proof);
  return rc;
}

// ------------------------------------------------------------------------------------------------ prove
static size_t proof_len_impl(const p2b_circuit* c, const p2b_batch* cs, const p2b_fri_params* fp, size_t n_pis,
                             size_t* fri_len_out) {
  const p2b_circuit_desc& d = c->d;
  const uint32_t nch = d.num_challenges;
  // widths of the four oracles: constants|sigmas, wires, Zs|partial products, quotient chunks
  p2b_batch shape[4];
  for (int i = 0; i < 4; i++) shape[i].log_n = d.degree_bits, shape[i].rate_bits = fp->rate_bits, shape[i].cap_height = fp->cap_height;
  shape[0].cap_height = cs->cap_height;
  shape[0].n_cols = cs->n_cols;
  shape[1].n_cols = d.num_wires;
  shape[2].n_cols = (size_t)nch * (1 + d.num_partial_products);
  shape[3].n_cols = (size_t)nch * d.quotient_degree_factor;
  const p2b_batch* ptrs[4] = {&shape[0], &shape[1], &shape[2], &shape[3]};
  const size_t fri_len = fri_proof_len_impl(ptrs, 4, fp);
  if (fri_len_out) *fri_len_out = fri_len;
  const size_t n_open = cs->n_cols + d.num_wires + shape[2].n_cols + nch + shape[3].n_cols;
  return 3 * ((size_t)4 << fp->cap_height) + 2 * n_open + fri_len + n_pis;
}

extern "C" size_t p2b_proof_len(const p2b_circuit* c, const p2b_batch* cs, const p2b_fri_params* fp, size_t n_public_inputs) {
  if (!c || !cs || !fp || fp->n_layers > P2B_MAX_FRI_LAYERS) return 0;
  uint32_t sum = 0;
  for (uint32_t l = 0; l < fp->n_layers; l++) sum += fp->reduction_arity_bits[l];
  if (sum > c->d.degree_bits) return 0;
  return proof_len_impl(c, cs, fp, n_public_inputs, nullptr);
}

// ---- the body of a proof: everything is ENQUEUED on the context's stream, nothing waits for the device.  It runs
// either eagerly or under stream capture (ctx->cap != nullptr), in which case every host source it copies from is a
// slot of the plan's pinned block.  The proof words land in h_proof_dst (pinned host memory).
// wire_cols (host column pointers) or d_wires (device, column-major): exactly one is non-null
static int prove_body(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                      const uint64_t* const* wire_cols, const uint64_t* d_wires, const uint64_t* public_inputs,
                      size_t n_public_inputs, const p2b_fri_params* fp, uint64_t* h_proof_dst) {
  const p2b_circuit_desc& d = c->d;
  const uint32_t nch = d.num_challenges, rb = fp->rate_bits, caph = fp->cap_height;
  size_t fri_len = 0;
  const size_t proof_len = proof_len_impl(c, cs, fp, n_public_inputs, &fri_len);
  const size_t cap_words = (size_t)4 << caph;
  const size_t w_zs = (size_t)nch * (1 + d.num_partial_products), w_q = (size_t)nch * d.quotient_degree_factor;

  p2b_batch *wires = nullptr, *zs = nullptr, *qt = nullptr;
  p2b_challenger* ch = nullptr;
  uint64_t *d_small = nullptr, *d_proof = nullptr;
  auto cleanup = [&](int code) {
    p2b_batch_free(qt);
    p2b_batch_free(zs);
    p2b_batch_free(wires);
    p2b_challenger_free(ch);
    dfree(ctx, d_small);
    dfree(ctx, d_proof);
    return code;
  };
#define TRY(expr)                             \
  do {                                        \
    int rc__ = (expr);                        \
    if (rc__ != P2B_OK) return cleanup(rc__); \
  } while (0)
#define CUP(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return cleanup(fail(ctx, P2B_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__)); \
  } while (0)
  // small device block: digest[4] | pi_hash[4] | betas[nch] | gammas[nch] | alphas[nch] | zeta[2] | zeta_next[2] |
  // apow[nch * n_terms] (the powers of alpha the quotient's reduce_with_powers needs)
  const uint32_t n_terms = quotient_n_terms(d);
  std::vector<uint64_t> h_small(8 + 3 * nch + 4, 0);
  for (int i = 0; i < 4; i++) h_small[i] = circuit_digest[i] % GL_P;
  TRY(dmalloc(ctx, &d_small, h_small.size() + (size_t)nch * n_terms));
  TRY(h2d_small(ctx, d_small, h_small.data(), h_small.size(), ctx->cap ? &ctx->cap->h_digest : nullptr));
  uint64_t *d_digest = d_small, *d_pih = d_small + 4, *d_betas = d_small + 8, *d_gammas = d_betas + nch,
           *d_alphas = d_gammas + nch, *d_zeta = d_alphas + nch, *d_zeta_next = d_zeta + 2, *d_apow = d_zeta_next + 2;
  TRY(dmalloc(ctx, &d_proof, proof_len + 1));  // + one status word: "zeta lies in the subgroup"
  // P2B_TRACE=1: phase-by-phase wall clock of one proof on stderr (each mark synchronises, so the phases are
  // serialised GPU time + host time; a development aid, never on in measurements)
  const bool trace = trace_enabled() && !ctx->cap;
  auto now_us = [] {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
  };
  double t_prev = trace ? now_us() : 0, t_enq = t_prev;
  auto mark = [&](const char* what) {
    if (!trace) return;
    const double t_e = now_us();
    cudaStreamSynchronize(ctx->stream);
    const double t = now_us();
    fprintf(stderr, "[p2b_prove] %-22s enqueue %8.1f us, done +%8.1f us (phase %8.1f us)\n", what, t_e - t_enq, t - t_e, t - t_prev);
    t_prev = t_enq = t;
  };
  // public_inputs_hash = PoseidonHash::hash_no_pad(public_inputs); the public inputs go straight to their place at
  // the end of the proof
  uint64_t* d_pis = d_proof + (proof_len - n_public_inputs);
  {
    std::vector<uint64_t> pis(n_public_inputs ? n_public_inputs : 1);
    for (size_t i = 0; i < n_public_inputs; i++) pis[i] = public_inputs[i] % GL_P;
    TRY(h2d_small(ctx, d_pis, pis.data(), n_public_inputs, ctx->cap ? &ctx->cap->h_pis : nullptr));
  }
  hashk::k_hash_no_pad_single<<<1, 32, 0, ctx->stream>>>(d_pis, n_public_inputs, d_pih);
  ctx->launches++;
  mark("setup + pi hash");
  // wires commitment
  if (ctx->cap) {
    // under capture: ONE copy node into a buffer of the graph's own (host source = the plan's pinned staging matrix,
    // device source = the caller's matrix); its source is re-pointed before every launch
    const size_t n = (size_t)1 << d.degree_bits;
    uint64_t* d_in = nullptr;
    TRY(dmalloc(ctx, &d_in, (size_t)d.num_wires * n));
    ctx->cap->d_wires_dst = d_in;
    cudaError_t e = cudaMemcpyAsync(d_in, wire_cols ? ctx->cap->h_wires : d_wires, (size_t)d.num_wires * n * sizeof(uint64_t),
                                    wire_cols ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream);
    if (e != cudaSuccess) {
      dfree(ctx, d_in);
      return cleanup(fail(ctx, P2B_ERR_CUDA, "witness copy: %s", cudaGetErrorString(e)));
    }
    TRY(batch_build(ctx, d_in, true, d.num_wires, d.degree_bits, rb, caph, &wires, true));
  } else if (wire_cols) {
    TRY(batch_from_host(ctx, wire_cols, d.num_wires, d.degree_bits, rb, caph, P2B_KEEP_VALUES, true, &wires));
  } else {
    TRY(batch_from_dev(ctx, d_wires, d.num_wires, d.degree_bits, rb, caph, P2B_KEEP_VALUES, true, &wires));
  }
  mark("wires commit");
  TRY(p2b_challenger_new(ctx, &ch));
  {  // observe circuit digest, public inputs hash, wires cap; betas then gammas
    TrSegs sg;
    sg.add(d_digest, 4);
    sg.add(d_pih, 4);
    sg.add(tree_cap_ptr(&wires->tree), cap_words);
    TRY(transcript_step(ctx, ch->d_state, true, sg, d_betas, 2 * nch));
  }
  mark("transcript: betas");
  TRY(zs_pp_core(ctx, c, cs, wires, d_betas, d_gammas, rb, caph, &zs));
  mark("zs/pp + commit");
  {  // observe the Z / partial products cap; alphas and their powers
    TrSegs sg;
    sg.add(tree_cap_ptr(&zs->tree), cap_words);
    TRY(transcript_step(ctx, ch->d_state, false, sg, d_alphas, nch, 0, d_apow, n_terms));
  }
  mark("transcript: alphas");
  TRY(quotient_core(ctx, c, cs, wires, zs, d_pih, d_betas, d_gammas, d_alphas, rb, caph, &qt, d_apow));
  mark("quotient + commit");
  {  // observe the quotient cap; zeta, and zeta_next = g * zeta
    const uint64_t G = 1753635133440165772ull;
    const uint64_t g = d.degree_bits ? h_powmod(G, (uint64_t)1 << (32 - d.degree_bits)) : 1;
    TrSegs sg;
    sg.add(tree_cap_ptr(&qt->tree), cap_words);
    // plonky2: ensure!(zeta.exp_power_of_2(degree_bits) != F::Extension::ONE, "Opening point is in the subgroup.")
    TRY(transcript_step(ctx, ch->d_state, false, sg, d_zeta, 2, g, nullptr, 0, nullptr, 0, d.degree_bits + 1,
                        d_proof + proof_len));  // d_zeta | d_zeta_next are adjacent
  }
  mark("transcript: zeta");
  // proof layout: caps | openings (OpeningSet field order) | FRI proof | public inputs
  size_t off = 0;
  fusedk::CopyParams cpy{};
  for (p2b_batch* b : {wires, zs, qt}) {
    cpy.src[cpy.n] = tree_cap_ptr(&b->tree);
    cpy.dst[cpy.n] = d_proof + off;
    cpy.len[cpy.n] = (uint32_t)cap_words;
    cpy.n++;
    off += cap_words;
  }
  fusedk::k_copy_multi<<<cpy.n, 256, 0, ctx->stream>>>(cpy);
  ctx->launches++;
  uint64_t* o_constants = d_proof + off;  // constants | sigmas contiguous = all of constants_sigmas
  uint64_t* o_wires = o_constants + 2 * cs->n_cols;
  uint64_t* o_zs = o_wires + 2 * (size_t)d.num_wires;
  uint64_t* o_zs_next = o_zs + 2 * (size_t)nch;
  uint64_t* o_pp = o_zs_next + 2 * (size_t)nch;
  uint64_t* o_quot = o_pp + 2 * (w_zs - nch);
  {
    // OpeningSet::new: every opened polynomial in one launch (point 0 = zeta, point 1 = g zeta)
    fusedk::EvalParams ep{};
    const size_t n = (size_t)1 << d.degree_bits;
    struct R {
      const uint64_t* c;
      size_t count;
      uint32_t pt;
      uint64_t* o;
    } rs[6] = {{cs->d_coeffs, cs->n_cols, 0, o_constants},        {wires->d_coeffs, d.num_wires, 0, o_wires},
               {zs->d_coeffs, nch, 0, o_zs},                      {zs->d_coeffs, nch, 1, o_zs_next},
               {zs->d_coeffs + (size_t)nch * n, w_zs - nch, 0, o_pp}, {qt->d_coeffs, w_q, 0, o_quot}};
    uint32_t total = 0;
    for (int i = 0; i < 6; i++) {
      if (rs[i].count == 0) continue;
      const uint32_t k = ep.n_ranges++;
      ep.coeffs[k] = rs[i].c;
      ep.out[k] = rs[i].o;
      ep.point[k] = rs[i].pt;
      ep.first_cta[k] = total;
      total += (uint32_t)rs[i].count;
    }
    ep.first_cta[ep.n_ranges] = total;
    ep.n = n;
    ep.points = d_zeta;
    fusedk::k_eval_polys_multi<<<total, 256, 0, ctx->stream>>>(ep);
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) return cleanup(fail(ctx, P2B_ERR_CUDA, "k_eval_polys_multi launch failed"));
  }
  mark("openings");
  off += 2 * (cs->n_cols + d.num_wires + w_zs + nch + w_q);
  // challenger.observe_openings(&openings.to_fri_openings()): the zeta batch in FRI order, then zs_next — handed to
  // prove_openings, whose first launch observes them and squeezes alpha
  TrSegs open_segs;
  open_segs.add(o_constants, 2 * (cs->n_cols + d.num_wires + nch));
  open_segs.add(o_pp, 2 * (w_zs - nch + w_q));
  open_segs.add(o_zs_next, 2 * (size_t)nch);
  // FRI instance (CommonCircuitData::get_fri_instance): everything at zeta, the Zs again at g * zeta
  const p2b_batch* oracles[4] = {cs, wires, zs, qt};
  p2b_fri_batch fb[2] = {};
  fb[0].n_ranges = 4;
  for (uint32_t o = 0; o < 4; o++) fb[0].ranges[o] = {o, 0, (uint32_t)oracles[o]->n_cols};
  fb[1].n_ranges = 1;
  fb[1].ranges[0] = {2, 0, nch};
  TRY(check_fri_params(ctx, oracles, 4, fp));
  TRY(prove_openings_core(ctx, oracles, 4, fb, 2, d_zeta, ch, fp, d_proof + off, &open_segs));
  mark("prove_openings (FRI)");
  off += fri_len + n_public_inputs;
  if (off != proof_len) return cleanup(fail(ctx, P2B_ERR_INVALID, "internal: proof length mismatch"));
  CUP(cudaMemcpyAsync(h_proof_dst, d_proof, (proof_len + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  return cleanup(P2B_OK);
#undef TRY
#undef CUP
}

// ---- CUDA-graph plans -------------------------------------------------------------------------------------------
// A worker proves the same few circuits over and over (city_rollup_core_worker/src/actors/simple.rs:57-113), and a
// 2^12-row proof is ~50 dependent launches plus ~40 stream-ordered allocations: for a job this small the host side of
// every launch is visible.  The second proof of a given (circuit, constants_sigmas, FRI parameters, public-input count,
// witness location) on a context is therefore CAPTURED — prove_body runs once under stream capture, its allocations
// becoming memory nodes of the graph — and every later proof of that shape is one cudaGraphLaunch after the inputs
// have been put where the graph reads them: circuit digest, public inputs and pointer tables in the plan's pinned
// block, the witness through one copy node whose source is re-pointed before each launch.
// P2B_GRAPH=0 disables the plans (every proof runs prove_body eagerly).
static bool graphs_enabled() {
  static const bool v = [] {
    const char* e = getenv("P2B_GRAPH");
    return !e || e[0] != '0';
  }();
  return v;
}

static void plan_destroy(p2b_ctx* ctx, ProvePlan* pl) {
  if (!pl) return;
  if (pl->exec) {
    ctx_sync(ctx);  // a launch of this graph may still be in flight
    cudaGraphExecDestroy(pl->exec);
  }
  if (pl->graph) cudaGraphDestroy(pl->graph);
  if (pl->h_pin) cudaFreeHost(pl->h_pin);
  if (pl->h_wires) cudaFreeHost(pl->h_wires);
  if (pl->h_proof) cudaFreeHost(pl->h_proof);
  delete pl;
}

// drops the plans that refer to a circuit / batch that is being freed (ids are never reused)
static void plans_forget(p2b_ctx* ctx, uint64_t circuit_id, uint64_t batch_id) {
  for (size_t i = 0; i < ctx->plans.size();) {
    ProvePlan* pl = ctx->plans[i];
    if ((circuit_id && pl->circuit_id == circuit_id) || (batch_id && pl->cs_id == batch_id)) {
      plan_destroy(ctx, pl);
      ctx->plans.erase(ctx->plans.begin() + i);
    } else {
      i++;
    }
  }
}
static void plans_clear(p2b_ctx* ctx) {
  for (ProvePlan* pl : ctx->plans) plan_destroy(ctx, pl);
  ctx->plans.clear();
  cudaDeviceGraphMemTrim(ctx->device);
}

static ProvePlan* plan_find(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const p2b_fri_params* fp, size_t n_pis,
                            bool dev_src) {
  for (ProvePlan* pl : ctx->plans)
    if (pl->circuit_id == c->id && pl->cs_id == cs->id && pl->n_pis == n_pis && pl->dev_src == dev_src &&
        memcmp(&pl->fp, fp, sizeof(*fp)) == 0)
      return pl;
  return nullptr;
}

// capture prove_body into pl->exec.  Any failure leaves the plan in state FAILED (eager from then on) and the
// context healthy.
static int plan_capture(p2b_ctx* ctx, ProvePlan* pl, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                        const uint64_t* const* wire_cols, const uint64_t* d_wires, const uint64_t* public_inputs,
                        size_t n_public_inputs, const p2b_fri_params* fp) {
  pl->state = ProvePlan::FAILED;
  const size_t n = (size_t)1 << c->d.degree_bits;
  pl->wires_words = (size_t)c->d.num_wires * n;
  pl->proof_len = proof_len_impl(c, cs, fp, n_public_inputs, nullptr);
  pl->pin_cap = 4096 + 2 * (cs->n_cols + c->d.num_wires + 64) + n_public_inputs;
  if (cudaMallocHost((void**)&pl->h_pin, pl->pin_cap * sizeof(uint64_t)) != cudaSuccess ||
      cudaMallocHost((void**)&pl->h_proof, (pl->proof_len + 1) * sizeof(uint64_t)) != cudaSuccess ||
      (!pl->dev_src && cudaMallocHost((void**)&pl->h_wires, pl->wires_words * sizeof(uint64_t)) != cudaSuccess)) {
    cudaGetLastError();
    return fail(ctx, P2B_ERR_OOM, "pinned host allocation for a prove plan failed");
  }
  pl->pin_used = 0;
  const uint64_t launches0 = ctx->launches;
  if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return fail(ctx, P2B_ERR_UNSUPPORTED, "stream capture unavailable");
  }
  ctx->cap = pl;
  const bool was_poisoned = ctx->poisoned;
  int rc = prove_body(ctx, c, cs, circuit_digest, wire_cols, d_wires, public_inputs, n_public_inputs, fp, pl->h_proof);
  ctx->cap = nullptr;
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
  pl->n_kernels = (uint32_t)(ctx->launches - launches0);
  ctx->launches = launches0;  // nothing ran
  if (rc != P2B_OK || e != cudaSuccess || !g) {
    cudaGetLastError();
    if (g) cudaGraphDestroy(g);
    ctx->poisoned = was_poisoned;  // an error raised while capturing says nothing about the device
    return rc != P2B_OK ? rc : fail(ctx, P2B_ERR_UNSUPPORTED, "stream capture failed: %s", cudaGetErrorString(e));
  }
  pl->graph = g;
  if (cudaGraphInstantiate(&pl->exec, g, 0) != cudaSuccess) {
    cudaGetLastError();
    pl->exec = nullptr;
    return fail(ctx, P2B_ERR_UNSUPPORTED, "cudaGraphInstantiate failed");
  }
  // the witness copy node: the memcpy node that writes pl->d_wires_dst
  size_t n_nodes = 0;
  cudaGraphGetNodes(g, nullptr, &n_nodes);
  std::vector<cudaGraphNode_t> nodes(n_nodes);
  if (n_nodes) cudaGraphGetNodes(g, nodes.data(), &n_nodes);
  pl->has_wires_node = false;
  for (cudaGraphNode_t nd : nodes) {
    cudaGraphNodeType ty;
    if (cudaGraphNodeGetType(nd, &ty) != cudaSuccess || ty != cudaGraphNodeTypeMemcpy) continue;
    cudaMemcpy3DParms mp{};
    if (cudaGraphMemcpyNodeGetParams(nd, &mp) != cudaSuccess) continue;
    if (mp.dstPtr.ptr == (void*)pl->d_wires_dst) {
      pl->wires_node = nd;
      pl->has_wires_node = true;
      break;
    }
  }
  cudaGetLastError();
  if (!pl->has_wires_node) return fail(ctx, P2B_ERR_UNSUPPORTED, "prove plan: witness copy node not found");
  pl->state = ProvePlan::READY;
  return P2B_OK;
}

static bool cols_contiguous(const uint64_t* const* cols, size_t n_cols, size_t n) {
  for (size_t i = 1; i < n_cols; i++)
    if (cols[i] != cols[0] + i * n) return false;
  return true;
}

// one launch of a READY plan
static int plan_launch(p2b_ctx* ctx, ProvePlan* pl, const p2b_circuit* c, const uint64_t* circuit_digest,
                       const uint64_t* const* wire_cols, const uint64_t* d_wires, const uint64_t* public_inputs,
                       size_t n_public_inputs, bool caller_keeps_buffers) {
  const size_t n = (size_t)1 << c->d.degree_bits;
  for (int i = 0; i < 4; i++) pl->h_digest[i] = circuit_digest[i] % GL_P;
  for (size_t i = 0; i < n_public_inputs; i++) pl->h_pis[i] = public_inputs[i] % GL_P;
  const void* src;
  cudaMemcpyKind kind;
  if (pl->dev_src) {
    src = d_wires;
    kind = cudaMemcpyDeviceToDevice;
  } else {
    kind = cudaMemcpyHostToDevice;
    // a contiguous pinned matrix goes by DMA straight from the caller's memory when the caller keeps it untouched
    // until the proof is collected (the blocking p2b_prove); everything else is packed into the plan's pinned matrix
    if (caller_keeps_buffers && cols_contiguous(wire_cols, c->d.num_wires, n) && host_ptr_is_pinned(wire_cols[0])) {
      src = wire_cols[0];
      ctx->direct_src_pending = true;  // p2b_prove_upload_poll: the replayed graph reads the caller's memory
    } else {
      for (uint32_t w = 0; w < c->d.num_wires; w++) {
        if (!wire_cols[w]) return fail(ctx, P2B_ERR_INVALID, "cols[%u] is null", w);
        memcpy(pl->h_wires + (size_t)w * n, wire_cols[w], n * sizeof(uint64_t));
      }
      src = pl->h_wires;
    }
  }
  if (src != pl->cur_src) {
    CU(ctx, cudaGraphExecMemcpyNodeSetParams1D(pl->exec, pl->wires_node, pl->d_wires_dst, src, pl->wires_words * sizeof(uint64_t), kind));
    pl->cur_src = src;
  }
  CU(ctx, cudaGraphLaunch(pl->exec, ctx->stream));
  ctx->launches += pl->n_kernels;
  pl->last_use = ++ctx->plan_clock;
  return P2B_OK;
}

// p2b_prove_submit / p2b_prove_dev_submit: validate, pick eager or graph, enqueue.  The proof is then pending on the
// context until prove_collect.
static int prove_submit(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                        const uint64_t* const* wire_cols, const uint64_t* d_wires, const uint64_t* public_inputs,
                        size_t n_public_inputs, const p2b_fri_params* fp, bool caller_keeps_buffers) {
  CHECK_CTX(ctx);
  if (!c || !cs || !circuit_digest || (!wire_cols && !d_wires) || !fp || (n_public_inputs && !public_inputs))
    return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ctx->pending_words) return fail(ctx, P2B_ERR_INVALID, "a submitted proof has not been collected yet (p2b_prove_collect)");
  if (c->ctx != ctx || cs->ctx != ctx) return fail(ctx, P2B_ERR_INVALID, "handle of another context");
  if (fp->n_layers > P2B_MAX_FRI_LAYERS) return fail(ctx, P2B_ERR_INVALID, "too many FRI layers");
  if (cs->rate_bits != fp->rate_bits) return fail(ctx, P2B_ERR_INVALID, "constants_sigmas rate_bits differ from the FRI parameters");
  const p2b_circuit_desc& d = c->d;
  const uint32_t rb = fp->rate_bits, caph = fp->cap_height;
  {
    // validate the FRI parameters before any length is derived from them (unsigned underflow otherwise)
    uint32_t log_len = 0;
    int rc0 = check_fri_args(ctx, ((size_t)1 << d.degree_bits) << rb, fp->reduction_arity_bits, fp->n_layers, rb, &log_len);
    if (rc0) return rc0;
    uint32_t lc = log_len;
    for (uint32_t l = 0; l < fp->n_layers; l++) {
      lc -= fp->reduction_arity_bits[l];
      if (caph > lc) return fail(ctx, P2B_ERR_INVALID, "cap_height %u exceeds the height %u of FRI layer %u", caph, lc, l);
    }
    if (caph > d.degree_bits + rb || cs->cap_height > d.degree_bits + rb) return fail(ctx, P2B_ERR_INVALID, "cap_height too large");
  }
  if (cs->log_n != d.degree_bits || cs->n_cols != (size_t)d.num_constants + d.num_routed_wires)
    return fail(ctx, P2B_ERR_INVALID, "constants_sigmas (2^%u x %zu) does not match the circuit (2^%u x %u)", cs->log_n, cs->n_cols,
                d.degree_bits, d.num_constants + d.num_routed_wires);
  int rc = P2B_OK;
  const size_t proof_len = proof_len_impl(c, cs, fp, n_public_inputs, nullptr);
  const bool dev_src = d_wires != nullptr;
  // small proofs only: a large one is bound by its kernels, and its host upload is pipelined against the transforms
  const bool want_graph = graphs_enabled() && !ctx->profiling && !trace_enabled() && d.degree_bits <= 14;
  ProvePlan* pl = want_graph ? plan_find(ctx, c, cs, fp, n_public_inputs, dev_src) : nullptr;
  if (pl && pl->state == ProvePlan::SEEN) {
    rc = plan_capture(ctx, pl, c, cs, circuit_digest, wire_cols, d_wires, public_inputs, n_public_inputs, fp);
    if (rc != P2B_OK) {
      if (ctx->poisoned) return rc;
      ctx->plan_note = ctx->err;  // why this shape stays on the eager path (p2b_plan_info)
    }
  }
  if (pl && pl->state == ProvePlan::READY) {
    rc = plan_launch(ctx, pl, c, circuit_digest, wire_cols, d_wires, public_inputs, n_public_inputs, caller_keeps_buffers);
    if (rc) return rc;
    ctx->pending_src = pl->h_proof;
  } else {
    // eager
    if ((proof_len + 1) * sizeof(uint64_t) > ctx->h_proof_bytes) {
      if (ctx->h_proof) cudaFreeHost(ctx->h_proof);
      ctx->h_proof = nullptr;
      ctx->h_proof_bytes = 0;
      CU(ctx, cudaMallocHost((void**)&ctx->h_proof, (proof_len + 1) * sizeof(uint64_t)));
      ctx->h_proof_bytes = (proof_len + 1) * sizeof(uint64_t);
    }
    rc = prove_body(ctx, c, cs, circuit_digest, wire_cols, d_wires, public_inputs, n_public_inputs, fp, ctx->h_proof);
    if (rc) return rc;
    ctx->pending_src = ctx->h_proof;
    if (want_graph && !pl) {
      // first proof of this shape on this context: it has warmed the twiddle / coset tables; the next one is captured
      if (ctx->plans.size() >= 16) {  // bounded: evict the least recently used plan
        size_t lru = 0;
        for (size_t i = 1; i < ctx->plans.size(); i++)
          if (ctx->plans[i]->last_use < ctx->plans[lru]->last_use) lru = i;
        plan_destroy(ctx, ctx->plans[lru]);
        ctx->plans.erase(ctx->plans.begin() + lru);
      }
      ProvePlan* np = new (std::nothrow) ProvePlan();
      if (np) {
        np->circuit_id = c->id;
        np->cs_id = cs->id;
        np->fp = *fp;
        np->n_pis = n_public_inputs;
        np->dev_src = dev_src;
        np->state = ProvePlan::SEEN;
        np->last_use = ++ctx->plan_clock;
        ctx->plans.push_back(np);
      }
    }
  }
  ctx->pending_words = proof_len;
  ctx->pending_pow_index = proof_len - n_public_inputs - 1;
  return P2B_OK;
}

static int prove_collect(p2b_ctx* ctx, uint64_t* proof_out, size_t proof_cap) {
  CHECK_CTX(ctx);
  if (!proof_out) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (!ctx->pending_words) return fail(ctx, P2B_ERR_INVALID, "no submitted proof to collect");
  const size_t len = ctx->pending_words;
  if (proof_cap < len) return fail(ctx, P2B_ERR_INVALID, "proof buffer too small: %zu < %zu words", proof_cap, len);
  ctx->pending_words = 0;
  ctx->h2d_event_pending = false;  // everything enqueued has finished once the stream is idle
  ctx->direct_src_pending = false;
  CU(ctx, ctx_sync(ctx));
  memcpy(proof_out, ctx->pending_src, len * sizeof(uint64_t));
  if (ctx->pending_src[len] != 0) return fail(ctx, P2B_ERR_INVALID, "Opening point is in the subgroup.");  // plonky2's own failure
  if (proof_out[ctx->pending_pow_index] == ~0ull) return fail(ctx, P2B_ERR_INVALID, "no proof-of-work witness found");
  return P2B_OK;
}

extern "C" int p2b_prove(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                         const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                         const p2b_fri_params* fp, uint64_t* proof_out, size_t proof_cap) {
  if (ctx && (!wire_cols || !proof_out)) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ctx && c && cs && fp && fp->n_layers <= P2B_MAX_FRI_LAYERS) {
    const size_t need = p2b_proof_len(c, cs, fp, n_public_inputs);
    if (need && proof_cap < need) return fail(ctx, P2B_ERR_INVALID, "proof buffer too small: %zu < %zu words", proof_cap, need);
  }
  int rc = prove_submit(ctx, c, cs, circuit_digest, wire_cols, nullptr, public_inputs, n_public_inputs, fp, true);
  return rc ? rc : prove_collect(ctx, proof_out, proof_cap);
}
extern "C" int p2b_prove_dev(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                             const uint64_t* d_wire_values, const uint64_t* public_inputs, size_t n_public_inputs,
                             const p2b_fri_params* fp, uint64_t* proof_out, size_t proof_cap) {
  if (ctx && (!d_wire_values || !proof_out)) return fail(ctx, P2B_ERR_INVALID, "null argument");
  if (ctx && c && cs && fp && fp->n_layers <= P2B_MAX_FRI_LAYERS) {
    const size_t need = p2b_proof_len(c, cs, fp, n_public_inputs);
    if (need && proof_cap < need) return fail(ctx, P2B_ERR_INVALID, "proof buffer too small: %zu < %zu words", proof_cap, need);
  }
  int rc = prove_submit(ctx, c, cs, circuit_digest, nullptr, d_wire_values, public_inputs, n_public_inputs, fp, true);
  return rc ? rc : prove_collect(ctx, proof_out, proof_cap);
}
extern "C" int p2b_prove_submit(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                                const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                                const p2b_fri_params* fp) {
  if (ctx && !wire_cols) return fail(ctx, P2B_ERR_INVALID, "null argument");
  int rc = prove_submit(ctx, c, cs, circuit_digest, wire_cols, nullptr, public_inputs, n_public_inputs, fp, false);
  // the caller may refill its witness buffers as soon as this returns: wait for the upload (not for the proof)
  if (rc == P2B_OK && ctx->h2d_event_pending) {
    ctx->h2d_event_pending = false;
    CU(ctx, cudaEventSynchronize(ctx->ev_h2d));
  }
  return rc;
}
// The driver-thread form: no wait for the upload.  With many proofs in flight an upload can sit behind another context's
// kernels in a shared hardware queue for milliseconds; a thread that drives several contexts must not stall there.
extern "C" int p2b_prove_submit_nowait(p2b_ctx* ctx, const p2b_circuit* c, const p2b_batch* cs, const uint64_t* circuit_digest,
                                       const uint64_t* const* wire_cols, const uint64_t* public_inputs, size_t n_public_inputs,
                                       const p2b_fri_params* fp) {
  if (ctx && !wire_cols) return fail(ctx, P2B_ERR_INVALID, "null argument");
  // the caller keeps its buffers untouched until p2b_prove_upload_poll says otherwise: a contiguous pinned matrix is read
  // by DMA straight from the caller's memory (no 4.4 MB host copy into the plan's staging matrix inside submit)
  return prove_submit(ctx, c, cs, circuit_digest, wire_cols, nullptr, public_inputs, n_public_inputs, fp, true);
}
extern "C" int p2b_prove_upload_poll(p2b_ctx* ctx) {
  CHECK_CTX(ctx);
  if (ctx->direct_src_pending) {
    // the upload is a node of the replayed graph: no event of its own, the buffer is free when the proof has finished
    cudaError_t q = cudaStreamQuery(ctx->stream);
    if (q == cudaErrorNotReady) return 0;
    if (q != cudaSuccess) return fail(ctx, P2B_ERR_CUDA, "cudaStreamQuery: %s", cudaGetErrorString(q));
    ctx->direct_src_pending = false;
  }
  if (!ctx->h2d_event_pending) return 1;  // pageable columns were staged inside submit; nothing else reads host memory
  cudaError_t e = cudaEventQuery(ctx->ev_h2d);
  if (e == cudaErrorNotReady) return 0;
  if (e != cudaSuccess) return fail(ctx, P2B_ERR_CUDA, "cudaEventQuery: %s", cudaGetErrorString(e));
  ctx->h2d_event_pending = false;
  return 1;
}
extern "C" int p2b_prove_poll(p2b_ctx* ctx) {
  CHECK_CTX(ctx);
  if (!ctx->pending_words) return fail(ctx, P2B_ERR_INVALID, "no submitted proof");
  cudaError_t e = cudaStreamQuery(ctx->stream);
  if (e == cudaSuccess) return 1;
  if (e == cudaErrorNotReady) return 0;
  return fail(ctx, P2B_ERR_CUDA, "cudaStreamQuery: %s", cudaGetErrorString(e));
}
extern "C" int p2b_prove_collect(p2b_ctx* ctx, uint64_t* proof_out, size_t proof_cap) { return prove_collect(ctx, proof_out, proof_cap); }
extern "C" int p2b_plan_info(p2b_ctx* ctx, uint32_t* n_ready, uint32_t* n_seen, uint32_t* n_failed, uint32_t* kernels_per_launch) {
  CHECK_CTX(ctx);
  uint32_t r = 0, sn = 0, f = 0, k = 0;
  for (ProvePlan* pl : ctx->plans) {
    if (pl->state == ProvePlan::READY) r++, k = pl->n_kernels;
    else if (pl->state == ProvePlan::SEEN) sn++;
    else f++;
  }
  if (n_ready) *n_ready = r;
  if (n_seen) *n_seen = sn;
  if (n_failed) *n_failed = f;
  if (kernels_per_launch) *kernels_per_launch = k;
  if (f) ctx->err = "prove plan not captured: " + ctx->plan_note;
  return P2B_OK;
}
