// engine.cu -- device context, workspaces and kernel launch sequences behind the C ABI.
//
// This is the "thin C-ABI layer" between the plain-C host code (host/eip2537_host.c, which
// mirrors the control flow of /root/reference/src/eip2537.c: length checks, k=1 delegation,
// error conventions) and the sm_100a kernels (msm.cuh, pairing.cuh).  No torch types, no
// CPU arithmetic fallback: if CUDA is unavailable every entry point fails with
// EIP2537_MEMORY_ERROR and bls12_b200_last_error() says why.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "engine.h"
#include "msm.cuh"
#include "pairing.cuh"
#include "coop12.cuh"
#include "pairing_dot.cuh"
#include "pairing_coop.cuh"
#include "map.cuh"
#include "../../include/eip2537_b200.h"

using namespace b200;

// ------------------------------------------------------------------------------------------
// bookkeeping
// ------------------------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};
static thread_local char g_last_error[256] = "";
static std::atomic<int> g_forced_window{0};
static std::atomic<int> g_checked_msm{0};   // opt-in: subgroup-check MULTIEXP inputs (changes error codes vs the reference)
static long pairing_coop_default() {
  const char* e = getenv("B200_PAIRING_COOP_MAX");
  return e ? atol(e) : 128;   // measured crossover with the batch planner: ~160 calls (profiles/r01_bench.md)
}
static std::atomic<long> g_pairing_coop_max{pairing_coop_default()};

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess) {                                                                 \
      snprintf(g_last_error, sizeof g_last_error, "%s:%d %s: %s", __FILE__, __LINE__, #expr,    \
               cudaGetErrorString(err__));                                                      \
      return E_MEMORY;                                                                          \
    }                                                                                           \
  } while (0)

#define CUDA_TRY2(expr)                                                                         \
  do {                                                                                          \
    cudaError_t err__ = (expr);                                                                 \
    if (err__ != cudaSuccess) {                                                                 \
      snprintf(g_last_error, sizeof g_last_error, "%s:%d %s: %s", __FILE__, __LINE__, #expr,    \
               cudaGetErrorString(err__));                                                      \
      return EIP2537_MEMORY_ERROR;                                                              \
    }                                                                                           \
  } while (0)

#define LAUNCH(kernel, grid, block, stream, ...)                                                \
  do {                                                                                          \
    kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
    if (tl_pool) tl_pool->launches.fetch_add(1, std::memory_order_relaxed);                     \
  } while (0)

#define LAUNCH_SMEM(kernel, grid, block, smem, stream, ...)                                     \
  do {                                                                                          \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                 \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
    if (tl_pool) tl_pool->launches.fetch_add(1, std::memory_order_relaxed);                     \
  } while (0)

static inline unsigned blocks_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

struct Engine;
static thread_local Engine* tl_engine = nullptr;      // the workspace leased by this thread (Buffer::reserve, LAUNCH)
static void engine_quiesce(Engine* e);

struct Buffer {
  void* ptr = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return E_SUCCESS;
    if (ptr) {
      engine_quiesce(tl_engine);          // never free a buffer that an asynchronous submission may still be reading
      cudaFree(ptr);
    }
    ptr = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t err = cudaMalloc(&ptr, want);
    if (err != cudaSuccess) {
      (void)cudaGetLastError();           // the failed allocation must not be reported again by a later, valid call
      ptr = nullptr;
      snprintf(g_last_error, sizeof g_last_error, "cudaMalloc(%zu bytes): %s", want, cudaGetErrorString(err));
      return E_MEMORY;
    }
    cap = want;
    return E_SUCCESS;
  }
  void release() { if (ptr) cudaFree(ptr); ptr = nullptr; cap = 0; }
};

// Pinned staging ring for PAGEABLE caller memory (Go heap slices, Rust Vec): the caller's bytes are copied into
// pinned slots by a few host threads while earlier slots are in flight to the device.
struct StageRing {
  static constexpr int SLOTS = 4;
  static constexpr size_t SLOT_BYTES = 8u << 20;
  unsigned char* slot[SLOTS] = {};
  cudaEvent_t freed[SLOTS] = {};
  bool used[SLOTS] = {};
  int next = 0;
};

// One WORKSPACE: streams, events, grow-only HBM buffers, pinned result words.  A workspace serves one submission
// at a time; every device owns a small pool of them (DevicePool), so concurrent host threads run concurrently.
struct Engine {
  int device = -1;
  int sm_count = 148;
  struct DevicePool* pool = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;            // H2D copies of streamed MSM chunks (overlap with accumulation)
  cudaEvent_t ev_group[8] = {};
  cudaEvent_t ev_tail = nullptr;
  // timing of the last streamed MSM (events with timing): when the H2D copies took a large share of the call -- several
  // ranks sharing the host's memory bandwidth -- the next call uses a chunk schedule whose LAST chunks are small
  cudaEvent_t ev_copy[2] = {}, ev_span[2] = {};
  bool span_valid = false;
  bool copy_bound = false;
  // asynchronous (device-resident) submissions return while their kernels still use the buffers below: `busy` is
  // recorded on the caller's stream at the end of each of them; the next user of the workspace orders itself after it
  cudaEvent_t busy = nullptr;
  bool pending = false;
  cudaStream_t pending_stream = nullptr;
  // MSM workspaces
  Buffer raw, pts, digits, counts, offsets, block_sums, entries, buckets, nodes_a, nodes_b, partial, out, status, order, tasks, task_partials;
  Buffer pair_a, pair_b, pair_layout;      // batched-affine pair rounds (msm.cuh k_pair_round)
  Buffer gather;                             // multi-GPU: N records of {partial sum, status key}
  // pairing workspaces
  Buffer pr_raw, pr_offsets, pr_lines, pr_tasks, pr_g1, pr_g2, pr_status, pr_f, pr_outs, pr_errs, pr_slots, pr_park, pr_prog;
  int pr_prog_len = 0;
  unsigned char* h_out = nullptr;            // pinned: result bytes
  unsigned long long* h_status = nullptr;    // pinned
  StageRing ring;
  std::vector<Buffer*> all_buffers() {
    return {&pair_a, &pair_b, &pair_layout, &raw, &pts, &digits, &counts, &offsets, &block_sums, &entries, &buckets, &nodes_a, &nodes_b, &partial, &out, &status,
            &order, &tasks, &task_partials, &gather, &pr_raw, &pr_offsets, &pr_lines, &pr_tasks, &pr_g1, &pr_g2, &pr_status, &pr_f,
            &pr_outs, &pr_errs, &pr_slots, &pr_park, &pr_prog};
  }
};

static void engine_quiesce(Engine* e) {
  if (e && e->pending) { cudaEventSynchronize(e->busy); e->pending = false; }
}

static constexpr int MAX_DEVICES = 16;
struct DevicePool {
  int device = -1;
  int sm_count = 148;
  std::atomic<bool> ready{false};
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Engine*> idle;
  int created = 0;
  std::atomic<uint64_t> launches{0};
};
static DevicePool g_pools[MAX_DEVICES];
static std::mutex g_init_mu;
static thread_local DevicePool* tl_pool = nullptr;
static int max_workspaces() {
  static const int v = getenv("B200_WORKSPACES") ? atoi(getenv("B200_WORKSPACES")) : 4;
  return v < 1 ? 1 : (v > 16 ? 16 : v);
}

static void engine_destroy(Engine* e) {
  if (!e) return;
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (Buffer* b : e->all_buffers()) b->release();
  if (e->h_out) cudaFreeHost(e->h_out);
  if (e->h_status) cudaFreeHost(e->h_status);
  for (int k = 0; k < StageRing::SLOTS; k++) {
    if (e->ring.slot[k]) cudaFreeHost(e->ring.slot[k]);
    if (e->ring.freed[k]) cudaEventDestroy(e->ring.freed[k]);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->stream2) cudaStreamDestroy(e->stream2);
  for (int k = 0; k < 8; k++) if (e->ev_group[k]) cudaEventDestroy(e->ev_group[k]);
  if (e->ev_tail) cudaEventDestroy(e->ev_tail);
  if (e->busy) cudaEventDestroy(e->busy);
  for (int k = 0; k < 2; k++) { if (e->ev_copy[k]) cudaEventDestroy(e->ev_copy[k]); if (e->ev_span[k]) cudaEventDestroy(e->ev_span[k]); }
  delete e;
}

// create one workspace on the CURRENT device (== pool.device); partially created resources are destroyed on failure
static int engine_create(DevicePool& pool, Engine** out) {
  Engine* e = new Engine();
  e->device = pool.device; e->sm_count = pool.sm_count; e->pool = &pool;
  cudaError_t err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (err == cudaSuccess) {
    int lo = 0, hi = 0;
    err = cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (err == cudaSuccess) err = cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, hi);
  }
  for (int k = 0; k < 8 && err == cudaSuccess; k++) err = cudaEventCreateWithFlags(&e->ev_group[k], cudaEventDisableTiming);
  if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->ev_tail, cudaEventDisableTiming);
  if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e->busy, cudaEventDisableTiming);
  for (int k = 0; k < 2 && err == cudaSuccess; k++) { err = cudaEventCreate(&e->ev_copy[k]); if (err == cudaSuccess) err = cudaEventCreate(&e->ev_span[k]); }
  if (err == cudaSuccess) err = cudaMallocHost((void**)&e->h_out, 4096);
  if (err == cudaSuccess) err = cudaMallocHost((void**)&e->h_status, 64);
  if (err != cudaSuccess) {
    (void)cudaGetLastError();
    snprintf(g_last_error, sizeof g_last_error, "workspace creation on device %d: %s", pool.device, cudaGetErrorString(err));
    engine_destroy(e);
    return E_MEMORY;
  }
  *out = e;
  return E_SUCCESS;
}

static int pool_get(DevicePool** out, int device) {
  if (device < 0) CUDA_TRY(cudaGetDevice(&device));
  if (device >= MAX_DEVICES) { snprintf(g_last_error, sizeof g_last_error, "device %d out of range", device); return E_MEMORY; }
  DevicePool& p = g_pools[device];
  if (!p.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(g_init_mu);
    if (!p.ready.load(std::memory_order_relaxed)) {
      int count = 0;
      CUDA_TRY(cudaGetDeviceCount(&count));
      if (device >= count) { snprintf(g_last_error, sizeof g_last_error, "device %d of %d does not exist", device, count); return E_MEMORY; }
      CUDA_TRY(cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, device));
      p.device = device;
      p.ready.store(true, std::memory_order_release);
    }
  }
  *out = &p;
  return E_SUCCESS;
}

// RAII lease of one workspace of a device.  Host-buffer entry points run on the workspace's own stream and finish
// with a stream synchronize; device-resident entry points enqueue on the CALLER's stream and call submitted().
struct Lease {
  Engine* e = nullptr;
  DevicePool* pool = nullptr;
  int prev_device = -1;
  Engine* prev_tl_engine = nullptr;
  DevicePool* prev_tl_pool = nullptr;
  // `async_stream`: the stream an asynchronous submission will use (its identity lets back-to-back submissions on
  // one stream reuse a workspace without any wait: stream order already serialises them)
  int acquire(int device, bool is_async = false, cudaStream_t async_stream = nullptr) {
    (void)cudaGetLastError();        // a stale, non-sticky error of an earlier call must not fail this one
    int rc = pool_get(&pool, device);
    if (rc) return rc;
    CUDA_TRY(cudaGetDevice(&prev_device));
    if (prev_device != pool->device) CUDA_TRY(cudaSetDevice(pool->device));
    {
      std::unique_lock<std::mutex> lk(pool->mu);
      for (;;) {
        if (!pool->idle.empty()) {
          size_t pick = pool->idle.size() - 1;
          for (size_t i = 0; i < pool->idle.size(); i++) {       // prefer a workspace that needs no wait
            Engine* c = pool->idle[i];
            if (!c->pending || (is_async && c->pending_stream == async_stream)) { pick = i; break; }
          }
          e = pool->idle[pick];
          pool->idle.erase(pool->idle.begin() + pick);
          break;
        }
        if (pool->created < max_workspaces()) {
          pool->created++;
          lk.unlock();
          rc = engine_create(*pool, &e);
          if (rc) { lk.lock(); pool->created--; lk.unlock(); pool->cv.notify_one(); restore(); return rc; }
          break;
        }
        pool->cv.wait(lk);
      }
    }
    if (e->pending && !(is_async && e->pending_stream == async_stream)) engine_quiesce(e);
    prev_tl_engine = tl_engine; prev_tl_pool = tl_pool;
    tl_engine = e; tl_pool = pool;
    return E_SUCCESS;
  }
  // asynchronous submission finished enqueuing on `s`
  void submitted(cudaStream_t s) {
    if (!e) return;
    if (cudaEventRecord(e->busy, s) == cudaSuccess) { e->pending = true; e->pending_stream = s; }
    else { (void)cudaGetLastError(); cudaStreamSynchronize(s); e->pending = false; }
  }
  void restore() {
    if (prev_device >= 0 && pool && prev_device != pool->device) cudaSetDevice(prev_device);
    prev_device = -1;
  }
  ~Lease() {
    if (e) {
      tl_engine = prev_tl_engine; tl_pool = prev_tl_pool;
      { std::lock_guard<std::mutex> lk(pool->mu); pool->idle.push_back(e); }
      pool->cv.notify_one();
    }
    restore();
  }
};

// ------------------------------------------------------------------------------------------
// MSM
// ------------------------------------------------------------------------------------------
// Window width.  Measured on B200 (tools/sweep_window.py, device-resident, every c in 3..16 at
// n = 2^7 .. 2^20, G1 and G2): the optimum tracks log2(n) -- one thread per bucket needs both enough
// buckets to fill the GPU and short per-bucket chains -- and flattens at 15-16 where the bucket
// reduction starts to cost as much as it saves.  Within +-1 of the optimum the time changes by < 5 %.
static int choose_window(size_t n, bool g2) {
  int forced = g_forced_window.load();
  if (forced >= 2 && forced <= 16) return forced;
  int lg = 0;
  while ((n >> (lg + 1)) != 0) lg++;
  if (lg >= 20) return 16;
  // measured (tools/sweep_window.py, tools/sweep_window_small.py): ~log2 n, one less in the mid range where the
  // tail (buckets) still outweighs the accumulation (points); G2's tail is heavier, so it stays one below longer
  if (lg >= 11 && lg <= (g2 ? 16 : 14)) return lg - 1 > 15 ? 15 : lg - 1;
  return lg < 7 ? 7 : (lg > 15 ? 15 : lg);
}

// optional per-stage timing (bench.py's roofline leg): events around the stages of the last MSM
static std::atomic<int> g_profile{0};
// (developer / bench mode: bls12_b200_set_profile(1) is meant for ONE caller thread on one device at a time)
struct StageTimer {
  static constexpr int N = 8;
  cudaEvent_t ev[N] = {};
  bool made = false;
  int used = 0;
  std::mutex mu;
  void mark(int i, cudaStream_t s) {
    if (!g_profile.load()) return;
    std::lock_guard<std::mutex> lk(mu);
    if (!made) {
      for (int k = 0; k < N; k++)
        if (cudaEventCreate(&ev[k]) != cudaSuccess) { (void)cudaGetLastError(); return; }
      made = true;
    }
    if (cudaEventRecord(ev[i], s) != cudaSuccess) { (void)cudaGetLastError(); return; }
    if (i + 1 > used) used = i + 1;
  }
};
static StageTimer g_stage;
static StageTimer g_pstage;   // pairing: decode | lines | plan+accumulate | calls
static unsigned long long g_last_entries = 0, g_last_affine_adds = 0, g_last_walk_adds = 0;
static const uint32_t* g_work_counts = nullptr; static size_t g_work_nbt = 0; static int g_work_rounds = 0, g_work_device = 0;
// bucket sizes of the last profiled MULTIEXP -> D = non-zero digits, additions done pairwise in affine form, additions of the XYZZ walk
// (read back from the workspace on request: valid until the next call on that workspace; developer / bench mode, one caller)
static void msm_count_work() {
  if (!g_work_counts) return;
  std::vector<uint32_t> hc(g_work_nbt);
  int prev = 0; cudaGetDevice(&prev); cudaSetDevice(g_work_device);
  cudaMemcpy(hc.data(), g_work_counts, g_work_nbt * sizeof(uint32_t), cudaMemcpyDeviceToHost);
  cudaSetDevice(prev);
  unsigned long long D = 0, aff = 0, walk = 0;
  for (uint32_t m : hc) {
    D += m;
    for (int k = 0; k < g_work_rounds; k++) { aff += m / 2; m = (m + 1) / 2; }
    walk += m;
  }
  g_last_entries = D; g_last_affine_adds = aff; g_last_walk_adds = walk;
  g_work_counts = nullptr;
}

// The MSM runs in three phases so that a host-resident input can be STREAMED: bucket accumulation is
// additive, so each chunk of pairs is decoded, sorted and accumulated on top of the buckets left by
// the previous chunks while the next chunk is still crossing PCIe; the latency-bound tail (bucket
// reduction, window combine) runs once at the end.
template <class F>
struct MsmRun {
  MsmPlan plan;
  bool glv;
  size_t nbt, chunk_cap;
  Affine<F>* pts; int* digits; uint32_t *counts, *cursors, *offsets, *block_sums, *entries, *order, *bin_total, *bin_start;
  OrderCounters* oc; OverflowTask* tasks; BigBucket* big; XYZZ<F>* task_partials; XYZZ<F>* buckets;
  size_t max_tasks;
  int rounds;                                  // batched-affine pair rounds before the XYZZ walk (0 = none)
  Affine<F>*pair_a, *pair_b; uint32_t *counts_r, *offsets_r, *total_padded;
};

template <class F>
static int msm_begin(Engine& e, MsmRun<F>& r, size_t n_total, size_t chunk_n) {
  // checked MULTIEXP: subgroup-proven inputs allow the GLV split (msm.cuh glv_split): 2n points, 128-bit scalars
  r.glv = g_checked_msm.load() != 0 && !(getenv("B200_NO_GLV") && atoi(getenv("B200_NO_GLV")));
  if (r.glv) { n_total *= 2; chunk_n *= 2; }
  r.plan = make_plan(choose_window(n_total, sizeof(F) == sizeof(Fp2)), r.glv ? 128 : 256);
  const MsmPlan& plan = r.plan;
  r.nbt = (size_t)plan.nwin * plan.nb;   // total buckets, uniform layout
  r.chunk_cap = chunk_n;
  const size_t nbt = r.nbt;
  int rc;
  // offsets, cursors and sorted entries are 32-bit and a point index carries its sign in bit 31
  if (chunk_n >= ((size_t)1 << 31) || chunk_n * (size_t)plan.nwin > 0xFFFFFFFFull) {
    snprintf(g_last_error, sizeof g_last_error, "MULTIEXP of %zu pairs per chunk exceeds the 32-bit sort index: shard it (bls12_b200_init_multi) or split the call", chunk_n);
    return E_MEMORY;
  }
  if ((rc = e.pts.reserve(chunk_n * sizeof(Affine<F>)))) return rc;
  if ((rc = e.digits.reserve(chunk_n * plan.nwin * sizeof(int)))) return rc;
  if ((rc = e.counts.reserve(2 * nbt * sizeof(uint32_t)))) return rc;       // counts + cursors
  if ((rc = e.offsets.reserve(nbt * sizeof(uint32_t)))) return rc;
  if ((rc = e.block_sums.reserve((nbt / 1024 + 2) * sizeof(uint32_t)))) return rc;
  // Batched-affine pair rounds (msm.cuh k_pair_round) are OPT-IN (B200_AFFINE_ROUNDS = 1 or 2): bit-exact, but measured at
  // break-even on this pipeline (G1 2^20: accumulate stage 6.40 ms against 6.37 ms; profiles/r02_msm_tail.md).
  {
    static const int rounds_env = getenv("B200_AFFINE_ROUNDS") ? atoi(getenv("B200_AFFINE_ROUNDS")) : 0;
    r.rounds = rounds_env < 0 ? 0 : (rounds_env > 2 ? 2 : rounds_env);
  }
  if (r.rounds && chunk_n * (size_t)plan.nwin + nbt * (((size_t)1 << r.rounds) - 1) > 0xFFFFFFFFull) r.rounds = 0;   // 32-bit offsets incl. padding
  const size_t pad_slots = r.rounds ? nbt * (((size_t)1 << r.rounds) - 1) : 0;
  if ((rc = e.entries.reserve((chunk_n * plan.nwin + pad_slots + 4) * sizeof(uint32_t)))) return rc;
  if (r.rounds) {
    const size_t bound = chunk_n * plan.nwin + pad_slots;
    if ((rc = e.pair_a.reserve((bound / 2 + 1) * sizeof(Affine<F>)))) return rc;
    if (r.rounds > 1 && (rc = e.pair_b.reserve((bound / 4 + 1) * sizeof(Affine<F>)))) return rc;
    if ((rc = e.pair_layout.reserve((2 * nbt + 4) * sizeof(uint32_t)))) return rc;
  }
  if ((rc = e.buckets.reserve(nbt * sizeof(XYZZ<F>)))) return rc;
  r.max_tasks = (chunk_n * (size_t)plan.nwin) / 64 + 2;      // cap >= 64 entries per task
  if ((rc = e.order.reserve(nbt * sizeof(uint32_t) + 2 * ORDER_BINS * sizeof(uint32_t) + sizeof(OrderCounters)))) return rc;
  if ((rc = e.tasks.reserve(r.max_tasks * (sizeof(OverflowTask) + sizeof(BigBucket))))) return rc;
  if ((rc = e.task_partials.reserve(r.max_tasks * sizeof(XYZZ<F>)))) return rc;
  r.pts = (Affine<F>*)e.pts.ptr;
  r.digits = (int*)e.digits.ptr;
  r.counts = (uint32_t*)e.counts.ptr;
  r.cursors = r.counts + nbt;
  r.offsets = (uint32_t*)e.offsets.ptr;
  r.block_sums = (uint32_t*)e.block_sums.ptr;
  r.entries = (uint32_t*)e.entries.ptr;
  r.buckets = (XYZZ<F>*)e.buckets.ptr;
  r.order = (uint32_t*)e.order.ptr;
  r.bin_total = r.order + nbt;
  r.bin_start = r.bin_total + ORDER_BINS;
  r.oc = (OrderCounters*)(r.bin_start + ORDER_BINS);
  r.tasks = (OverflowTask*)e.tasks.ptr;
  r.big = (BigBucket*)(r.tasks + r.max_tasks);
  r.task_partials = (XYZZ<F>*)e.task_partials.ptr;
  r.pair_a = (Affine<F>*)e.pair_a.ptr;
  r.pair_b = (Affine<F>*)e.pair_b.ptr;
  r.counts_r = (uint32_t*)e.pair_layout.ptr;
  r.offsets_r = r.counts_r ? r.counts_r + nbt : nullptr;
  r.total_padded = r.counts_r ? r.offsets_r + nbt : nullptr;
  return E_SUCCESS;
}

// decode + digit recoding + counting sort + size-ordered bucket accumulation of one chunk of pairs
template <class F>
static int msm_feed(Engine& e, MsmRun<F>& r, const uint32_t* d_raw, size_t n, uint64_t index_base, bool first,
                    unsigned long long* d_status, cudaStream_t s) {
  const MsmPlan& plan = r.plan;
  const size_t nbt = r.nbt;
  if (first) g_stage.mark(0, s);
  CUDA_TRY(cudaMemsetAsync(r.counts, 0, 2 * nbt * sizeof(uint32_t), s));
  const int glv = r.glv ? 1 : 0;
  const size_t nv = r.glv ? 2 * n : n;      // virtual points fed to the bucket machinery
  LAUNCH(k_decode<F>, blocks_for(n, 128), 128, s, d_raw, n, r.pts, d_status, (size_t)index_base, glv);
  if (g_checked_msm.load()) {   // opt-in "checked MSM" (SURVEY.md 8(f)-4): reject points outside G1/G2 with code 2
    int rc2 = e.pr_status.reserve(n * sizeof(int));
    if (rc2) return rc2;
    LAUNCH(k_points_check<F>, blocks_for(n, 64), 64, s, d_raw, n, Wire<F>::PAIR_WORDS, 1, (int*)e.pr_status.ptr);
    LAUNCH(k_codes_to_status, blocks_for(n, 256), 256, s, (const int*)e.pr_status.ptr, n, (size_t)index_base, d_status);
  }
  LAUNCH(k_digits<F>, blocks_for(n, 256), 256, s, d_raw, n, r.pts, plan, r.digits, r.counts, glv);
  unsigned nblk = blocks_for(nbt, 1024);
  const int R = r.rounds;
  const uint32_t pad_mask = (1u << R) - 1;
  const size_t total_digits = nv * plan.nwin;
  const size_t entry_bound = total_digits + (R ? nbt * (size_t)pad_mask : 0);     // sorted entries incl. alignment padding
  if (R) CUDA_TRY(cudaMemsetAsync(r.entries, 0xFF, entry_bound * sizeof(uint32_t), s));   // ENTRY_NONE in the padding slots
  LAUNCH(k_scan_blocks, nblk, 1024, s, r.counts, r.offsets, r.block_sums, (uint32_t)nbt, pad_mask);
  LAUNCH(k_scan_sums, 1, 1024, s, r.block_sums, nblk);
  LAUNCH(k_scan_fix, nblk, 1024, s, r.offsets, r.block_sums, (uint32_t)nbt);
  LAUNCH(k_scatter, blocks_for(nv * plan.nwin, 256), 256, s, r.digits, nv, plan, r.offsets, r.cursors, r.entries);
  if (first) g_stage.mark(1, s);     // stage "accumulate" = pair rounds + bucket schedule + XYZZ walk
  // batched-affine pair rounds: the segments shrink 2^R-fold before the XYZZ walk
  const Affine<F>* walk_pts = r.pts;
  const uint32_t *walk_counts = r.counts, *walk_offsets = r.offsets;
  if (R) {
    LAUNCH(k_pair_total, 1, 1, s, r.counts, r.offsets, (uint32_t)nbt, R, r.total_padded);
    LAUNCH((k_pair_round<F, true, PAIR_BATCH>), blocks_for(entry_bound / 2 + 1, 128 * PAIR_BATCH), 128, s, r.pts, r.entries, r.total_padded, 0, r.pair_a);
    if (R > 1) LAUNCH((k_pair_round<F, false, PAIR_BATCH>), blocks_for(entry_bound / 4 + 1, 128 * PAIR_BATCH), 128, s, r.pair_a, (const uint32_t*)nullptr, r.total_padded, 1, r.pair_b);
    LAUNCH(k_pair_layout, blocks_for(nbt, 256), 256, s, r.counts, r.offsets, (uint32_t)nbt, R, r.counts_r, r.offsets_r);
    walk_pts = R > 1 ? r.pair_b : r.pair_a;
    walk_counts = r.counts_r; walk_offsets = r.offsets_r;
  }
  // size-ordered bucket schedule + overflow plan for oversized buckets
  uint32_t cap = (uint32_t)(4 * ((total_digits >> R) / nbt + 1) + 64);
  CUDA_TRY(cudaMemsetAsync(r.bin_total, 0, 2 * ORDER_BINS * sizeof(uint32_t) + sizeof(OrderCounters), s));
  LAUNCH(k_order_hist, nblk, 1024, s, walk_counts, walk_offsets, (uint32_t)nbt, cap, r.bin_total, r.oc, r.big, r.tasks);
  LAUNCH(k_order_scan, 1, 1024, s, r.bin_total, r.bin_start);
  LAUNCH(k_order_scatter, nblk, 1024, s, walk_counts, (uint32_t)nbt, r.bin_start, r.order);
  // multiply-loop unrolling in k_accumulate<Fp> (6 rows per iteration: -2.3 % on the kernel, profiles/r02_k1_probes.md) and the
  // register budget of k_accumulate<Fp2>; B200_ACC_ROWS / B200_ACC_G2_BLOCKS are developer switches
  static const int acc_rows = getenv("B200_ACC_ROWS") ? atoi(getenv("B200_ACC_ROWS")) : 6;
  static const int g2_blocks = getenv("B200_ACC_G2_BLOCKS") ? atoi(getenv("B200_ACC_G2_BLOCKS")) : 4;
  const unsigned acc_grid = blocks_for(nbt, 128);
  const int add_flag = first ? 0 : 1;
#define ACC_LAUNCH(K) LAUNCH(K, acc_grid, 128, s, walk_pts, r.entries, walk_offsets, walk_counts, r.order, (uint32_t)nbt, cap, add_flag, r.buckets)
  if (R) {
    if (sizeof(F) == sizeof(Fp)) ACC_LAUNCH((k_accumulate<F, 6, 4, true>));
    else                         ACC_LAUNCH((k_accumulate<F, 2, 4, true>));
  } else if (sizeof(F) == sizeof(Fp)) {
    if (acc_rows == 6)       ACC_LAUNCH((k_accumulate<F, 6, 4>));
    else if (acc_rows == 12) ACC_LAUNCH((k_accumulate<F, 12, 4>));
    else                     ACC_LAUNCH((k_accumulate<F, 2, 4>));
  } else {
    if (g2_blocks == 2)      ACC_LAUNCH((k_accumulate<F, 2, 2>));
    else if (g2_blocks == 3) ACC_LAUNCH((k_accumulate<F, 2, 3>));
    else                     ACC_LAUNCH((k_accumulate<F, 2, 4>));
  }
#undef ACC_LAUNCH
  const size_t tasks_bound = (total_digits >> R) / cap + 2;
  if (R) LAUNCH((k_accumulate_overflow<F, true>), blocks_for(tasks_bound, 128), 128, s, walk_pts, r.entries, r.tasks, r.oc, r.task_partials);
  else   LAUNCH((k_accumulate_overflow<F, false>), blocks_for(tasks_bound, 128), 128, s, walk_pts, r.entries, r.tasks, r.oc, r.task_partials);
  LAUNCH(k_merge_overflow<F>, blocks_for(tasks_bound * 32, 128), 128, s, r.big, r.oc, r.task_partials, r.buckets);
  if (first) {
    g_stage.mark(2, s);
    if (g_profile.load()) { g_work_counts = r.counts; g_work_nbt = nbt; g_work_rounds = R; g_work_device = e.device; }   // counted lazily (bls12_b200_last_msm_work)
  }
  CUDA_TRY(cudaGetLastError());
  return E_SUCCESS;
}

// Bucket reduction (leaf threads fold 2^L0_log buckets by running sums; the levels above only ADD vectors of partial sums and
// one walk per window scales them, msm.cuh k_reduce_level / k_window_finish), then the Horner combine of the window sums.
// (Overlapping this latency-bound tail with the accumulation of other windows on a second stream was tried and measured
// slower: 11.7 vs 10.7 ms for 2^20 G1 -- the dependent chains need the multiply pipe to themselves.)
template <class F>
static int msm_tail(Engine& e, MsmRun<F>& r, XYZZ<F>* d_partial, const unsigned long long* d_status, unsigned long long* d_status_copy,
                    cudaStream_t s) {
  const MsmPlan& plan = r.plan;
  int rc;
  static const int leaf_env = getenv("B200_LEAF_LOG") ? atoi(getenv("B200_LEAF_LOG")) : 0;
  // Buckets per leaf thread (log2), measured per size (profiles/r02_msm_tail.md): small MSMs have few buckets and the
  // leaf level is pure latency (2 dependent point additions per bucket), so short leaves win; large ones are
  // throughput-bound and 8 buckets per thread cost the fewest additions (14 per 8 buckets).
  const size_t total_buckets = (size_t)plan.nwin * plan.nb;
  const int L0_auto = total_buckets < (1u << 14) ? 1 : (total_buckets < (1u << 16) ? 2 : 3);
  int L0_want = leaf_env > 0 ? leaf_env : L0_auto;
  int L0_log = plan.log_nb < L0_want ? plan.log_nb : L0_want;
  size_t nodes_per_win = plan.nb >> L0_log;
  static const int inner_env = getenv("B200_INNER_LOG") ? atoi(getenv("B200_INNER_LOG")) : 0;
  // two children per level: one dependent addition per level and one more component per node (measured against four: 1.30 vs 1.42 ms)
  const ReduceLevels lv = make_reduce_levels(plan.log_nb, L0_log, inner_env > 0 ? inner_env : 1);
  // node vectors of XYZZ sums: the leaves write (s, w); level i adds 2^l_log - 1 components and has 2^l_log times fewer nodes
  const size_t node_words = (size_t)plan.nwin * (2 * nodes_per_win + REDUCE_MAX_LEVELS * 8) * sizeof(XYZZ<F>);
  if ((rc = e.nodes_a.reserve(node_words))) return rc;
  if ((rc = e.nodes_b.reserve(node_words + (size_t)plan.nwin * sizeof(XYZZ<F>)))) return rc;
  XYZZ<F>* cur = (XYZZ<F>*)e.nodes_a.ptr;
  XYZZ<F>* nxt = (XYZZ<F>*)e.nodes_b.ptr;
  LAUNCH(k_reduce_leaf<F>, blocks_for(plan.nwin * nodes_per_win, 128), 128, s, r.buckets,
         (uint32_t)(plan.nwin * nodes_per_win), 1 << L0_log, (Node<F>*)cur);
  int K = 2;
  for (int i = 0; i < lv.n; i++) {
    const int l_log = lv.l_log[i], K_out = K + (1 << l_log) - 1;
    const size_t out_per_win = nodes_per_win >> l_log;
    const size_t comps = (size_t)plan.nwin * out_per_win * K_out;
    // lane groups pay only where a level is latency-bound (a group spends 32 / 128 multiplication slots on an addition that costs a
    // thread 14 / 42): measured crossover, profiles/r02_msm_tail.md
    static const size_t coop_env = getenv("B200_REDUCE_COOP_BELOW") ? (size_t)atol(getenv("B200_REDUCE_COOP_BELOW")) : 0;
    const size_t coop_below = coop_env ? coop_env : (sizeof(F) == sizeof(Fp) ? 8192 : 6144);
    if (comps >= coop_below) LAUNCH((k_reduce_level<F, false>), blocks_for(comps, 128), 128, s, cur, K, (uint32_t)(plan.nwin * out_per_win), l_log, nxt);
    else                     LAUNCH((k_reduce_level<F, true>), blocks_for(comps * Coop<F>::LANES, 128), 128, s, cur, K, (uint32_t)(plan.nwin * out_per_win), l_log, nxt);
    XYZZ<F>* tmp = cur; cur = nxt; nxt = tmp;
    nodes_per_win = out_per_win;
    K = K_out;
  }
  Hom<F>* tw = (Hom<F>*)((char*)e.nodes_b.ptr + node_words);        // T_w of every window, behind the node vectors
  LAUNCH(k_window_finish<F>, (unsigned)plan.nwin, 8 * Coop<F>::LANES, s, cur, K, lv, tw);
  g_stage.mark(3, s);
  LAUNCH(k_window_combine<F>, 1, 32, s, tw, plan, d_partial, d_status, d_status_copy);
  g_stage.mark(4, s);
  CUDA_TRY(cudaGetLastError());
  return E_SUCCESS;
}

// device-resident input: one chunk
template <class F>
static int msm_pipeline(Engine& e, const uint32_t* d_raw, size_t n, uint64_t index_base, XYZZ<F>* d_partial,
                        unsigned long long* d_status, unsigned long long* d_status_copy, cudaStream_t s) {
  MsmRun<F> r;
  int rc;
  if ((rc = msm_begin<F>(e, r, n, n))) return rc;
  if ((rc = msm_feed<F>(e, r, d_raw, n, index_base, true, d_status, s))) return rc;
  return msm_tail<F>(e, r, d_partial, d_status, d_status_copy, s);
}

// ---- pageable caller memory ------------------------------------------------------------------
// Go passes heap slices (go/blst_eip2537.go:74), Rust passes Vec / stack arrays: ordinary pageable memory, which
// cudaMemcpyAsync moves at ~7 GB/s through the driver's own bounce buffer.  Here a small pool of host threads
// copies the caller's bytes into the workspace's pinned ring while earlier ring slots are in flight over PCIe.
static std::atomic<int> g_ngpu{1};      // bls12_b200_init_multi: devices a single large call is sharded over

struct CopyPool {
  struct Job { unsigned char* d; const unsigned char* s; size_t n; std::atomic<int>* left; };
  std::vector<std::thread> th;
  std::mutex mu;
  std::condition_variable cv;
  std::vector<Job> q;
  bool stop = false;
  int nthreads = 0;
  void start() {
    std::lock_guard<std::mutex> lk(mu);
    if (nthreads) return;
    // B200_COPY_THREADS per device in use (default 4), never more than half the host's hardware threads
    int want = (getenv("B200_COPY_THREADS") ? atoi(getenv("B200_COPY_THREADS")) : 4) * (g_ngpu.load() > 1 ? g_ngpu.load() : 1);
    int hw = (int)std::thread::hardware_concurrency();
    if (hw > 0 && want > hw / 2) want = hw / 2;
    if (want < 1) want = 1;
    nthreads = want;
    for (int i = 0; i < want; i++) th.emplace_back([this] { this->run(); });
  }
  void run() {
    for (;;) {
      Job j;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [this] { return stop || !q.empty(); });
        if (stop && q.empty()) return;
        j = q.back();
        q.pop_back();
      }
      memcpy(j.d, j.s, j.n);
      j.left->fetch_sub(1, std::memory_order_release);
    }
  }
  // dst <- src with every pool thread and the caller taking a share
  void copy(unsigned char* d, const unsigned char* s, size_t n) {
    start();
    const int share = nthreads < 4 ? nthreads : 4;      // one copy never takes more than 4 helpers: concurrent callers share the pool
    const int parts = share + 1;
    const size_t per = ((n / parts) + 4095) & ~(size_t)4095;
    std::atomic<int> left{0};
    size_t off = 0;
    {
      std::lock_guard<std::mutex> lk(mu);
      for (int i = 0; i < share && off + per < n; i++, off += per) { left.fetch_add(1); q.push_back(Job{d + off, s + off, per, &left}); }
    }
    cv.notify_all();
    memcpy(d + off, s + off, n - off);
    while (left.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv.notify_all();
    for (auto& t : th) if (t.joinable()) t.join();
  }
};
static CopyPool g_copy_pool;

static bool host_pointer_is_pageable(const void* p) {
  static const int force = getenv("B200_FORCE_STAGING") ? atoi(getenv("B200_FORCE_STAGING")) : 0;
  if (force) return force > 0;
  cudaPointerAttributes at;
  cudaError_t err = cudaPointerGetAttributes(&at, p);
  if (err != cudaSuccess) { (void)cudaGetLastError(); return true; }
  return at.type == cudaMemoryTypeUnregistered;
}

// host -> device copy of `bytes` on stream `cs`; pageable sources go through the pinned ring
static int h2d_copy(Engine& e, unsigned char* dst, const unsigned char* src, size_t bytes, bool pageable, cudaStream_t cs) {
  if (!pageable || bytes < (256u << 10)) {
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cs));
    return E_SUCCESS;
  }
  StageRing& r = e.ring;
  for (size_t off = 0; off < bytes; off += StageRing::SLOT_BYTES) {
    const size_t len = bytes - off < StageRing::SLOT_BYTES ? bytes - off : StageRing::SLOT_BYTES;
    const int k = r.next;
    r.next = (k + 1) % StageRing::SLOTS;
    if (!r.slot[k]) {
      CUDA_TRY(cudaMallocHost((void**)&r.slot[k], StageRing::SLOT_BYTES));
      CUDA_TRY(cudaEventCreateWithFlags(&r.freed[k], cudaEventDisableTiming));
    }
    if (r.used[k]) CUDA_TRY(cudaEventSynchronize(r.freed[k]));     // the slot's previous copy has left the host
    g_copy_pool.copy(r.slot[k], src + off, len);
    CUDA_TRY(cudaMemcpyAsync(dst + off, r.slot[k], len, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaEventRecord(r.freed[k], cs));
    r.used[k] = true;
  }
  return E_SUCCESS;
}

// host-resident input: stream it in chunks, copy of chunk i+1 overlapped with accumulation of chunk i.
// Leaves the XYZZ partial sum at dst_partial (default e.partial) and the status key in e.status (and a copy at
// dst_status if given) -- both may be PEER pointers on another GPU --, all work queued on e.stream.
// developer sweep: B200_STREAM_SCHEDULE=9 B200_STREAM_CUTS="0,2,6,16,32" (cut points in thirty-seconds of the input)
static const std::vector<int>& env_cuts() {
  static const std::vector<int> v = [] {
    std::vector<int> r;
    const char* e = getenv("B200_STREAM_CUTS");
    while (e && *e) { r.push_back(atoi(e)); const char* c = strchr(e, ','); e = c ? c + 1 : nullptr; }
    if (r.size() < 2 || r.front() != 0 || r.back() != 32) r.clear();
    for (size_t i = 1; i < r.size(); i++) if (r[i] <= r[i - 1]) { r.clear(); break; }
    return r;
  }();
  return v;
}
template <class F>
static int msm_stream_from_host(Engine& e, const unsigned char* in, size_t n, uint64_t index_base,
                                XYZZ<F>* dst_partial = nullptr, unsigned long long* dst_status = nullptr) {
  int rc;
  const size_t pair_bytes = Wire<F>::PAIR_WORDS * 4;
  if ((rc = e.raw.reserve(n * pair_bytes))) return rc;
  if ((rc = e.partial.reserve(sizeof(XYZZ<Fp2>)))) return rc;
  if ((rc = e.status.reserve(8))) return rc;
  cudaStream_t s = e.stream, cs = e.stream2;
  const bool pageable = host_pointer_is_pageable(in);
  // Chunk boundaries in 32nds of the input.  Geometric sizes: the first copy (which nothing can hide) is small, and
  // every later copy is shorter than the accumulation of the chunk before it.  B200_STREAM_EVEN=1 restores 8 equal chunks.
  static const int even_env = getenv("B200_STREAM_EVEN") ? atoi(getenv("B200_STREAM_EVEN")) : 0;
  static const int cuts_geo[] = {0, 1, 2, 4, 8, 16, 32};
  static const int cuts_sym[] = {0, 1, 3, 7, 15, 23, 28, 31, 32};       // 1 2 4 8 8 5 3 1: small FIRST and LAST chunks
  static const int cuts_even8[] = {0, 4, 8, 12, 16, 20, 24, 28, 32};
  static const int cuts_mid[] = {0, 4, 8, 16, 32};
  static const int cuts_one[] = {0, 32};
  // compute-bound regime (one rank per host): four chunks 1/16, 1/8, 5/16, 1/2 -- every chunk repeats the fixed costs of the
  // sort and of loading / storing the buckets, so fewer chunks win once the first one is small enough (measured at 2^20:
  // 107.2 M points/s end to end against 105.7 M with round 1's six geometric chunks; profiles/r02_bench.md)
  static const int cuts_q4[] = {0, 2, 6, 16, 32};
  static const int cuts_q3[] = {0, 4, 16, 32};                           // 2^18 .. 2^20 pairs: three chunks (G2 2^18: 25.8 M against 23.7 M)
  // Copy-bound regime (measured on the previous streamed call of this workspace: the copies took > 40 % of the call, e.g.
  // 8 ranks sharing the host's ~186 GB/s): what follows the LAST copy is exposed, so the last chunks must be small.
  static const int sched_env = getenv("B200_STREAM_SCHEDULE") ? atoi(getenv("B200_STREAM_SCHEDULE")) : 0;   // 1 geometric, 2 symmetric
  if (e.span_valid && !sched_env) {
    float copy_ms = 0, span_ms = 0;
    if (cudaEventElapsedTime(&copy_ms, e.ev_copy[0], e.ev_copy[1]) == cudaSuccess &&
        cudaEventElapsedTime(&span_ms, e.ev_span[0], e.ev_span[1]) == cudaSuccess && span_ms > 0)
      e.copy_bound = copy_ms > 0.4f * span_ms;
    else
      (void)cudaGetLastError();
  }
  e.span_valid = false;
  const int* cuts = cuts_one;
  int nchunks = 1;
  if (n >= (1u << 18)) {
    if (even_env) { cuts = cuts_even8; nchunks = 8; }
    else if (sched_env == 2 || (sched_env == 0 && e.copy_bound)) { cuts = cuts_sym; nchunks = 8; }
    else if (sched_env == 1) { cuts = cuts_geo; nchunks = 6; }
    else if (sched_env == 9 && env_cuts().size() >= 2) { cuts = env_cuts().data(); nchunks = (int)env_cuts().size() - 1; }
    else if (n >= (1u << 20)) { cuts = cuts_q4; nchunks = 4; }
    else { cuts = cuts_q3; nchunks = 3; }
  }
  else if (n >= (1u << 16)) { cuts = cuts_mid; nchunks = 4; }
  size_t chunk_cap = 0;
  for (int c = 0; c < nchunks; c++) {
    const size_t lo = n * cuts[c] / 32, hi = n * cuts[c + 1] / 32;
    if (hi - lo > chunk_cap) chunk_cap = hi - lo;
  }
  MsmRun<F> r;
  if ((rc = msm_begin<F>(e, r, n, chunk_cap))) return rc;
  CUDA_TRY(cudaMemsetAsync(e.status.ptr, 0xFF, 8, s));
  const bool timed = nchunks > 1 && n >= (1u << 18);
  if (timed) {
    CUDA_TRY(cudaEventRecord(e.ev_span[0], s));
    CUDA_TRY(cudaStreamWaitEvent(cs, e.ev_span[0], 0));       // the copies of this call start with the call
    CUDA_TRY(cudaEventRecord(e.ev_copy[0], cs));
  }
  bool first = true;
  for (int c = 0; c < nchunks; c++) {
    const size_t lo = n * cuts[c] / 32, hi = n * cuts[c + 1] / 32;
    if (lo >= hi) continue;
    unsigned char* dst = (unsigned char*)e.raw.ptr + lo * pair_bytes;
    if (nchunks == 1) {
      if ((rc = h2d_copy(e, dst, in + lo * pair_bytes, (hi - lo) * pair_bytes, pageable, s))) return rc;
    } else {
      if ((rc = h2d_copy(e, dst, in + lo * pair_bytes, (hi - lo) * pair_bytes, pageable, cs))) return rc;
      CUDA_TRY(cudaEventRecord(e.ev_group[c], cs));
      CUDA_TRY(cudaStreamWaitEvent(s, e.ev_group[c], 0));
    }
    if (timed && c == nchunks - 1) CUDA_TRY(cudaEventRecord(e.ev_copy[1], cs));
    if ((rc = msm_feed<F>(e, r, (const uint32_t*)dst, hi - lo, index_base + lo, first, (unsigned long long*)e.status.ptr, s))) return rc;
    first = false;
  }
  rc = msm_tail<F>(e, r, dst_partial ? dst_partial : (XYZZ<F>*)e.partial.ptr, (unsigned long long*)e.status.ptr, dst_status, s);
  if (!rc && timed) {
    CUDA_TRY(cudaEventRecord(e.ev_span[1], s));
    e.span_valid = true;       // every caller synchronises e.stream before the workspace is leased again
  }
  return rc;
}

// finish a host call on workspace e: partial sum(s) -> affine -> bytes; `out` is written only on success
// (eip2537.c:613, :701 encode last)
template <class F>
static int msm_finish_host(Engine& e, const ShardRecord* recs, int count, unsigned char* out) {
  int rc;
  const size_t out_bytes = Wire<F>::POINT_WORDS * 4;
  if ((rc = e.out.reserve(out_bytes + 16))) return rc;
  cudaStream_t s = e.stream;
  unsigned long long* d_st = (unsigned long long*)((unsigned char*)e.out.ptr + ((out_bytes + 15) & ~(size_t)15));
  if (recs) {
    LAUNCH(k_finalize_records<F>, 1, 32, s, recs, count, (uint32_t*)e.out.ptr, d_st);
  } else {
    LAUNCH(k_finalize<F>, 1, 32, s, (const XYZZ<F>*)e.partial.ptr, 1, (uint32_t*)e.out.ptr);
    d_st = (unsigned long long*)e.status.ptr;
  }
  CUDA_TRY(cudaMemcpyAsync(e.h_out, e.out.ptr, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(e.h_status, d_st, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  unsigned long long st = *e.h_status;
  if (st != STATUS_OK) return (int)(st & 0xFF);
  memcpy(out, e.h_out, out_bytes);
  return E_SUCCESS;
}

// ---- one process, several GPUs: the plain ABI call shards its pairs -------------------------
// bls12_b200_init_multi(G) makes a LARGE bls12_g{1,2}multiexp call split its pair range over G devices: one host
// thread and one stream per device, each device pulls its slice over its own PCIe link and runs the whole
// single-GPU pipeline; its LAST kernel (k_window_combine) stores the partial sum and the first-error key straight
// into device 0's gather buffer through NVLink peer memory (the one exchange step: G x 512 B), and device 0 sums,
// inverts once and encodes.  Error precedence is the minimum (global pair index << 8 | code) over the shards.
static bool g_peer_ok[MAX_DEVICES] = {};
static constexpr size_t MULTI_MIN_PAIRS_PER_GPU = (size_t)1 << 17;

template <class F>
static int msm_multi_host(const unsigned char* in, size_t n, unsigned char* out, int G) {
  Lease l0;
  int rc = l0.acquire(0);
  if (rc) return rc;
  Engine& e0 = *l0.e;
  if ((rc = e0.gather.reserve((size_t)G * sizeof(ShardRecord)))) return rc;
  ShardRecord* recs = (ShardRecord*)e0.gather.ptr;
  const size_t pair_bytes = Wire<F>::PAIR_WORDS * 4;
  std::vector<int> rcs(G, 0);
  std::vector<std::string> errs(G);
  auto shard = [&](int d) {
    const size_t lo = n * (size_t)d / G, hi = n * (size_t)(d + 1) / G;
    Lease ld;
    Engine* e = &e0;
    if (d != 0) {
      if ((rcs[d] = ld.acquire(d))) { errs[d] = g_last_error; return; }
      e = ld.e;
    }
    int r;
    if (d == 0 || g_peer_ok[d]) {
      r = msm_stream_from_host<F>(*e, in + lo * pair_bytes, hi - lo, lo, (XYZZ<F>*)recs[d].partial, &recs[d].status);
    } else {   // no peer mapping: stage locally, then a peer copy of the 512-byte record
      r = e->gather.reserve(sizeof(ShardRecord));
      ShardRecord* local = (ShardRecord*)e->gather.ptr;
      if (!r) r = msm_stream_from_host<F>(*e, in + lo * pair_bytes, hi - lo, lo, (XYZZ<F>*)local->partial, &local->status);
      if (!r && cudaMemcpyPeerAsync(&recs[d], 0, local, d, sizeof(ShardRecord), e->stream) != cudaSuccess) r = E_MEMORY;
    }
    if (!r && cudaStreamSynchronize(e->stream) != cudaSuccess) r = E_MEMORY;
    if (r) { rcs[d] = r; errs[d] = g_last_error; }
  };
  std::vector<std::thread> th;
  for (int d = 1; d < G; d++) th.emplace_back(shard, d);
  shard(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < G; d++)
    if (rcs[d]) { snprintf(g_last_error, sizeof g_last_error, "device %d: %s", d, errs[d].c_str()); return rcs[d]; }
  return msm_finish_host<F>(e0, recs, G, out);
}

template <class F>
static int msm_host_impl(const unsigned char* in, size_t n, unsigned char* out) {
  const int ngpu = g_ngpu.load();
  if (ngpu > 1) {
    size_t G = n / MULTI_MIN_PAIRS_PER_GPU;
    if (G > (size_t)ngpu) G = ngpu;
    if (G > 1) return msm_multi_host<F>(in, n, out, (int)G);
  }
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  if ((rc = msm_stream_from_host<F>(e, in, n, 0))) return rc;
  return msm_finish_host<F>(e, nullptr, 1, out);
}

extern "C" EIP2537_ERROR bls12_b200_init_multi(int ngpu) {
  int count = 0;
  CUDA_TRY2(cudaGetDeviceCount(&count));
  if (ngpu <= 0 || ngpu > count) ngpu = count;
  if (ngpu > MAX_DEVICES) ngpu = MAX_DEVICES;
  int prev = 0;
  CUDA_TRY2(cudaGetDevice(&prev));
  for (int d = 0; d < ngpu; d++) {
    DevicePool* p;
    int rc = pool_get(&p, d);
    if (rc) return (EIP2537_ERROR)rc;
    if (d == 0) continue;
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, d, 0) != cudaSuccess) { (void)cudaGetLastError(); can = 0; }
    if (can) {
      CUDA_TRY2(cudaSetDevice(d));
      cudaError_t err = cudaDeviceEnablePeerAccess(0, 0);
      if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) can = 0;
      (void)cudaGetLastError();
    }
    g_peer_ok[d] = can != 0;
  }
  CUDA_TRY2(cudaSetDevice(prev));
  g_ngpu.store(ngpu);
  return EIP2537_SUCCESS;
}
extern "C" int bls12_b200_multi_gpus(void) { return g_ngpu.load(); }
extern "C" uint64_t bls12_b200_device_launch_count(int device) {
  return device >= 0 && device < MAX_DEVICES ? g_pools[device].launches.load() : 0;
}

// multi-GPU sharding with host-resident shards: stream this rank's slice in, leave the partial sum and
// the status key in caller-provided DEVICE buffers (ready for an all-gather); synchronous
template <class F>
static int msm_partial_host_impl(const unsigned char* in, size_t n, uint64_t index_base, void* d_partial, uint64_t* d_status) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  if ((rc = msm_stream_from_host<F>(e, in, n, index_base, (XYZZ<F>*)d_partial, (unsigned long long*)d_status))) return rc;
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  return E_SUCCESS;
}
extern "C" EIP2537_ERROR bls12_b200_msm_partial_host(int group, const byte* in, size_t n, uint64_t index_base,
                                                     void* d_partial, uint64_t* d_status) {
  if (n == 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)(group == 1 ? msm_partial_host_impl<Fp>(in, n, index_base, d_partial, d_status)
                                    : msm_partial_host_impl<Fp2>(in, n, index_base, d_partial, d_status));
}

extern "C" int b200_msm_host(int group, const unsigned char* in, size_t n, unsigned char* out) {
  return group == 1 ? msm_host_impl<Fp>(in, n, out) : msm_host_impl<Fp2>(in, n, out);
}

static int check_device_args(const void* a, const void* b) {
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15)) {
    snprintf(g_last_error, sizeof g_last_error, "device pointers must be 16-byte aligned");
    return E_MEMORY;
  }
  return E_SUCCESS;
}

template <class F>
static int msm_partial_device_impl(const void* d_in, size_t n, uint64_t index_base, void* d_partial,
                                   uint64_t* d_status, void* stream) {
  int rc = check_device_args(d_in, d_partial);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  Lease lease;
  if ((rc = lease.acquire(-1, true, s))) return rc;
  rc = msm_pipeline<F>(*lease.e, (const uint32_t*)d_in, n, index_base, (XYZZ<F>*)d_partial, (unsigned long long*)d_status, nullptr, s);
  lease.submitted(s);
  return rc;
}

extern "C" EIP2537_ERROR bls12_b200_msm_partial_device(int group, const void* d_in, size_t n, uint64_t index_base,
                                                       void* d_partial, uint64_t* d_status, void* stream) {
  if (n == 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)(group == 1 ? msm_partial_device_impl<Fp>(d_in, n, index_base, d_partial, d_status, stream)
                                    : msm_partial_device_impl<Fp2>(d_in, n, index_base, d_partial, d_status, stream));
}

extern "C" EIP2537_ERROR bls12_b200_msm_combine_device(int group, const void* d_partials, int count, void* d_out,
                                                       void* stream) {
  DevicePool* p;
  int rc = pool_get(&p, -1);
  if (rc) return (EIP2537_ERROR)rc;
  (void)cudaGetLastError();
  cudaStream_t s = (cudaStream_t)stream;
  if (group == 1) LAUNCH(k_finalize<Fp>, 1, 32, s, (const XYZZ<Fp>*)d_partials, count, (uint32_t*)d_out);
  else            LAUNCH(k_finalize<Fp2>, 1, 32, s, (const XYZZ<Fp2>*)d_partials, count, (uint32_t*)d_out);
  CUDA_TRY2(cudaGetLastError());
  return EIP2537_SUCCESS;
}

template <class F>
static int msm_device_impl(const void* d_in, size_t n, void* d_out, uint64_t* d_status, cudaStream_t s) {
  int rc = check_device_args(d_in, d_out);
  if (rc) return rc;
  Lease lease;
  if ((rc = lease.acquire(-1, true, s))) return rc;
  Engine& e = *lease.e;
  if ((rc = e.partial.reserve(sizeof(XYZZ<Fp2>)))) return rc;      // the partial sum lives in THIS submission's workspace
  CUDA_TRY(cudaMemsetAsync(d_status, 0xFF, 8, s));
  rc = msm_pipeline<F>(e, (const uint32_t*)d_in, n, 0, (XYZZ<F>*)e.partial.ptr, (unsigned long long*)d_status, nullptr, s);
  if (!rc) {
    LAUNCH(k_finalize<F>, 1, 32, s, (const XYZZ<F>*)e.partial.ptr, 1, (uint32_t*)d_out);
    if (cudaGetLastError() != cudaSuccess) rc = E_MEMORY;
  }
  lease.submitted(s);
  return rc;
}
extern "C" EIP2537_ERROR bls12_b200_msm_device(int group, const void* d_in, size_t n, void* d_out,
                                               uint64_t* d_status, void* stream) {
  if (n == 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)(group == 1 ? msm_device_impl<Fp>(d_in, n, d_out, d_status, (cudaStream_t)stream)
                                    : msm_device_impl<Fp2>(d_in, n, d_out, d_status, (cudaStream_t)stream));
}

// ---- one process PER GPU (torchrun / MPI style): the exchange is ONE ncclAllGather issued from here ----------
// NCCL is loaded lazily with dlopen (libnccl.so.2, or B200_NCCL_LIB), so single-GPU consumers of the library
// (the Go / Rust packages) carry no NCCL dependency.
namespace nccl_dl {
typedef struct ncclComm* comm_t;
struct unique_id { char internal[128]; };
typedef int (*get_unique_id_fn)(unique_id*);
typedef int (*comm_init_rank_fn)(comm_t*, int, unique_id, int);
typedef int (*all_gather_fn)(const void*, void*, size_t, int, comm_t, cudaStream_t);
typedef int (*comm_destroy_fn)(comm_t);
typedef const char* (*get_error_string_fn)(int);
static void* handle = nullptr;
static get_unique_id_fn get_unique_id = nullptr;
static comm_init_rank_fn comm_init_rank = nullptr;
static all_gather_fn all_gather = nullptr;
static comm_destroy_fn comm_destroy = nullptr;
static get_error_string_fn get_error_string = nullptr;
static comm_t comm = nullptr;
static int world = 1, rank = 0;
static std::mutex mu;
static int load() {
  std::lock_guard<std::mutex> lk(mu);
  if (handle) return E_SUCCESS;
  const char* names[] = {getenv("B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) {
    if (!nm) continue;
    handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (handle) break;
  }
  if (!handle) { snprintf(g_last_error, sizeof g_last_error, "NCCL not found: %s", dlerror()); return E_MEMORY; }
  get_unique_id = (get_unique_id_fn)dlsym(handle, "ncclGetUniqueId");
  comm_init_rank = (comm_init_rank_fn)dlsym(handle, "ncclCommInitRank");
  all_gather = (all_gather_fn)dlsym(handle, "ncclAllGather");
  comm_destroy = (comm_destroy_fn)dlsym(handle, "ncclCommDestroy");
  get_error_string = (get_error_string_fn)dlsym(handle, "ncclGetErrorString");
  if (!get_unique_id || !comm_init_rank || !all_gather || !comm_destroy) {
    snprintf(g_last_error, sizeof g_last_error, "NCCL symbols missing");
    handle = nullptr;
    return E_MEMORY;
  }
  return E_SUCCESS;
}
static int fail(const char* what, int code) {
  snprintf(g_last_error, sizeof g_last_error, "%s: %s", what, get_error_string ? get_error_string(code) : "nccl error");
  return E_MEMORY;
}
}  // namespace nccl_dl

extern "C" EIP2537_ERROR bls12_b200_comm_unique_id(byte* id128) {
  int rc = nccl_dl::load();
  if (rc) return (EIP2537_ERROR)rc;
  nccl_dl::unique_id id;
  int r = nccl_dl::get_unique_id(&id);
  if (r) return (EIP2537_ERROR)nccl_dl::fail("ncclGetUniqueId", r);
  memcpy(id128, &id, 128);
  return EIP2537_SUCCESS;
}
extern "C" EIP2537_ERROR bls12_b200_comm_init(int world, int rank, const byte* id128) {
  int rc = nccl_dl::load();
  if (rc) return (EIP2537_ERROR)rc;
  DevicePool* p;
  if ((rc = pool_get(&p, -1))) return (EIP2537_ERROR)rc;
  nccl_dl::unique_id id;
  memcpy(&id, id128, 128);
  if (nccl_dl::comm) { nccl_dl::comm_destroy(nccl_dl::comm); nccl_dl::comm = nullptr; }
  int r = nccl_dl::comm_init_rank(&nccl_dl::comm, world, id, rank);
  if (r) return (EIP2537_ERROR)nccl_dl::fail("ncclCommInitRank", r);
  nccl_dl::world = world; nccl_dl::rank = rank;
  return EIP2537_SUCCESS;
}
extern "C" void bls12_b200_comm_destroy(void) {
  if (nccl_dl::comm) { nccl_dl::comm_destroy(nccl_dl::comm); nccl_dl::comm = nullptr; }
  nccl_dl::world = 1; nccl_dl::rank = 0;
}

// this rank's shard -> record {partial, status} -> ONE all-gather of world x 512 B -> every rank sums, inverts, encodes
template <class F>
static int msm_sharded_finish(Engine& e, uint32_t* d_out, unsigned long long* d_status, cudaStream_t s) {
  ShardRecord* recs = (ShardRecord*)e.gather.ptr;     // [0] = this rank's record, [1 .. world] = gathered
  if (nccl_dl::world > 1) {
    if (!nccl_dl::comm) { snprintf(g_last_error, sizeof g_last_error, "bls12_b200_comm_init was not called"); return E_MEMORY; }
    int r = nccl_dl::all_gather(recs, recs + 1, sizeof(ShardRecord), /*ncclUint8*/ 1, nccl_dl::comm, s);
    if (r) return nccl_dl::fail("ncclAllGather", r);
    LAUNCH(k_finalize_records<F>, 1, 32, s, recs + 1, nccl_dl::world, d_out, d_status);
  } else {
    LAUNCH(k_finalize_records<F>, 1, 32, s, recs, 1, d_out, d_status);
  }
  CUDA_TRY(cudaGetLastError());
  return E_SUCCESS;
}
template <class F>
static int msm_sharded_device_impl(const void* d_in, size_t n, uint64_t index_base, void* d_out, uint64_t* d_status, cudaStream_t s) {
  int rc = check_device_args(d_in, d_out);
  if (rc) return rc;
  Lease lease;
  if ((rc = lease.acquire(-1, true, s))) return rc;
  Engine& e = *lease.e;
  if ((rc = e.gather.reserve((size_t)(nccl_dl::world + 1) * sizeof(ShardRecord)))) return rc;
  ShardRecord* recs = (ShardRecord*)e.gather.ptr;
  if ((rc = e.status.reserve(8))) return rc;
  CUDA_TRY(cudaMemsetAsync(e.status.ptr, 0xFF, 8, s));
  rc = msm_pipeline<F>(e, (const uint32_t*)d_in, n, index_base, (XYZZ<F>*)recs[0].partial, (unsigned long long*)e.status.ptr, &recs[0].status, s);
  if (!rc) rc = msm_sharded_finish<F>(e, (uint32_t*)d_out, (unsigned long long*)d_status, s);
  lease.submitted(s);
  return rc;
}
extern "C" EIP2537_ERROR bls12_b200_msm_sharded_device(int group, const void* d_in, size_t n, uint64_t index_base, void* d_out,
                                                       uint64_t* d_status, void* stream) {
  if (n == 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)(group == 1 ? msm_sharded_device_impl<Fp>(d_in, n, index_base, d_out, d_status, (cudaStream_t)stream)
                                    : msm_sharded_device_impl<Fp2>(d_in, n, index_base, d_out, d_status, (cudaStream_t)stream));
}
template <class F>
static int msm_sharded_host_impl(const unsigned char* in, size_t n, uint64_t index_base, unsigned char* out) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t out_bytes = Wire<F>::POINT_WORDS * 4;
  if ((rc = e.gather.reserve((size_t)(nccl_dl::world + 1) * sizeof(ShardRecord)))) return rc;
  if ((rc = e.out.reserve(out_bytes + 16))) return rc;
  ShardRecord* recs = (ShardRecord*)e.gather.ptr;
  unsigned long long* d_st = (unsigned long long*)((unsigned char*)e.out.ptr + ((out_bytes + 15) & ~(size_t)15));
  if ((rc = msm_stream_from_host<F>(e, in, n, index_base, (XYZZ<F>*)recs[0].partial, &recs[0].status))) return rc;
  if ((rc = msm_sharded_finish<F>(e, (uint32_t*)e.out.ptr, d_st, e.stream))) return rc;
  CUDA_TRY(cudaMemcpyAsync(e.h_out, e.out.ptr, out_bytes, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaMemcpyAsync(e.h_status, d_st, 8, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY(cudaStreamSynchronize(e.stream));
  if (*e.h_status != STATUS_OK) return (int)(*e.h_status & 0xFF);
  memcpy(out, e.h_out, out_bytes);
  return E_SUCCESS;
}
extern "C" EIP2537_ERROR bls12_b200_msm_sharded_host(int group, const byte* in, size_t n, uint64_t index_base, byte* out) {
  if (n == 0) return EIP2537_INVALID_LENGTH;
  return (EIP2537_ERROR)(group == 1 ? msm_sharded_host_impl<Fp>(in, n, index_base, out) : msm_sharded_host_impl<Fp2>(in, n, index_base, out));
}

// ------------------------------------------------------------------------------------------
// batch of independent small MULTIEXP calls (host buffers)
// ------------------------------------------------------------------------------------------
template <class F>
static int msm_batch_impl(unsigned char* outs, int* errs, const unsigned char* in, const uint64_t* offsets, size_t n) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  constexpr size_t PAIR = Wire<F>::PAIR_WORDS * 4, OUT = Wire<F>::POINT_WORDS * 4;
  // calls with a bad length contribute no pairs (answered INVALID_LENGTH, eip2537.c:543 / :831)
  std::vector<unsigned long long> off(n + 1);
  size_t cursor = 0;
  bool any_bad = false;
  for (size_t i = 0; i < n; i++) {
    size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    off[i] = cursor;
    if (len == 0 || len % PAIR) any_bad = true; else cursor += len;
  }
  off[n] = cursor;
  const size_t total_pairs = cursor / PAIR;
  if ((rc = e.raw.reserve(cursor + 16)) || (rc = e.pr_offsets.reserve((n + 1) * 8)) || (rc = e.pr_outs.reserve(n * OUT)) ||
      (rc = e.pr_errs.reserve(n * 4)) || (rc = e.buckets.reserve(total_pairs * sizeof(XYZZ<F>))) ||
      (rc = e.pr_status.reserve(total_pairs * sizeof(int) + 16)))
    return rc;
  cudaStream_t s = e.stream;
  if (!any_bad) {
    CUDA_TRY(cudaMemcpyAsync(e.raw.ptr, in + offsets[0], cursor, cudaMemcpyHostToDevice, s));
  } else {
    for (size_t i = 0; i < n; i++) {
      size_t len = (size_t)(off[i + 1] - off[i]);
      if (len) CUDA_TRY(cudaMemcpyAsync((char*)e.raw.ptr + off[i], in + offsets[i], len, cudaMemcpyHostToDevice, s));
    }
  }
  CUDA_TRY(cudaMemcpyAsync(e.pr_offsets.ptr, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
  if (total_pairs)
    LAUNCH(k_batch_pair_mul<F>, blocks_for(total_pairs, 128), 128, s, (const uint32_t*)e.raw.ptr, total_pairs,
           (XYZZ<F>*)e.buckets.ptr, (int*)e.pr_status.ptr);
  LAUNCH(k_batch_call_sum<F>, blocks_for(n * 32, 128), 128, s, (const unsigned long long*)e.pr_offsets.ptr, n,
         (const XYZZ<F>*)e.buckets.ptr, (const int*)e.pr_status.ptr, (uint32_t*)e.pr_outs.ptr, (int*)e.pr_errs.ptr);
  CUDA_TRY(cudaMemcpyAsync(outs, e.pr_outs.ptr, n * OUT, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(errs, e.pr_errs.ptr, n * 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));   // `off` and the caller's buffers stay alive until here
  return E_SUCCESS;
}
extern "C" EIP2537_ERROR bls12_g1multiexp_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, const uint64_t* offsets, size_t n) {
  if (n == 0) return EIP2537_SUCCESS;
  return (EIP2537_ERROR)msm_batch_impl<Fp>(outs, (int*)errs, in, offsets, n);
}
extern "C" EIP2537_ERROR bls12_g2multiexp_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, const uint64_t* offsets, size_t n) {
  if (n == 0) return EIP2537_SUCCESS;
  return (EIP2537_ERROR)msm_batch_impl<Fp2>(outs, (int*)errs, in, offsets, n);
}

// ------------------------------------------------------------------------------------------
// single add (G1ADD / G2ADD: eip2537.c:434-471, :722-759) -- one thread; completes the ABI
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void k_add_points(const uint32_t* __restrict__ raw, uint32_t* __restrict__ out_words, unsigned long long* status) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  constexpr int PW = Wire<F>::POINT_WORDS;
  uint32_t w[PW];
  Affine<F> a, b;
  for (int i = 0; i < PW; i++) w[i] = raw[i];
  int ca = decode_point(a, w);
  for (int i = 0; i < PW; i++) w[i] = raw[PW + i];
  int cb = decode_point(b, w);
  if (ca) { *status = (0ull << 8) | (unsigned)ca; return; }
  if (cb) { *status = (1ull << 8) | (unsigned)cb; return; }
  XYZZ<F> acc = xyzz_from_affine(b);
  xyzz_madd(acc, a);
  Affine<F> r = xyzz_to_affine(acc);
  encode_point(w, r);
  for (int i = 0; i < PW; i++) out_words[i] = w[i];
}

template <class F>
static int add_host_impl(const unsigned char* in, unsigned char* out) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t in_bytes = 2 * Wire<F>::POINT_WORDS * 4, out_bytes = Wire<F>::POINT_WORDS * 4;
  if ((rc = e.raw.reserve(in_bytes))) return rc;
  if ((rc = e.out.reserve(out_bytes))) return rc;
  if ((rc = e.status.reserve(8))) return rc;
  cudaStream_t s = e.stream;
  CUDA_TRY(cudaMemcpyAsync(e.raw.ptr, in, in_bytes, cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemsetAsync(e.status.ptr, 0xFF, 8, s));
  LAUNCH(k_add_points<F>, 1, 32, s, (const uint32_t*)e.raw.ptr, (uint32_t*)e.out.ptr, (unsigned long long*)e.status.ptr);
  CUDA_TRY(cudaMemcpyAsync(e.h_out, e.out.ptr, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(e.h_status, e.status.ptr, 8, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  if (*e.h_status != STATUS_OK) return (int)(*e.h_status & 0xFF);
  memcpy(out, e.h_out, out_bytes);
  return E_SUCCESS;
}
extern "C" int b200_add_host(int group, const unsigned char* in, unsigned char* out) {
  return group == 1 ? add_host_impl<Fp>(in, out) : add_host_impl<Fp2>(in, out);
}

// ------------------------------------------------------------------------------------------
// MAP_FP_TO_G1 / MAP_FP2_TO_G2 (eip2537.c:1093-1163): n independent field elements, one thread each
// (map.cuh).  outs: n x (128 | 256) bytes, written only for elements whose code is 0; codes[n].
// ------------------------------------------------------------------------------------------
template <class F>
static int map_host_impl(unsigned char* outs, int* codes, const unsigned char* in, size_t n) {
  if (n == 0) return E_SUCCESS;
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t in_bytes = n * (Wire<F>::POINT_WORDS * 2), out_bytes = n * (Wire<F>::POINT_WORDS * 4);
  if ((rc = e.raw.reserve(in_bytes))) return rc;
  if ((rc = e.pts.reserve(out_bytes))) return rc;
  if ((rc = e.pr_errs.reserve(n * sizeof(int)))) return rc;
  cudaStream_t s = e.stream;
  CUDA_TRY(cudaMemcpyAsync(e.raw.ptr, in, in_bytes, cudaMemcpyHostToDevice, s));
  LAUNCH(k_map_to_group<F>, blocks_for(n, 64), 64, s, (const uint32_t*)e.raw.ptr, n, (uint32_t*)e.pts.ptr, (int*)e.pr_errs.ptr);
  if (n == 1) {          // single call: keep the caller's `out` untouched on error (eip2537.c:1118 encodes last)
    CUDA_TRY(cudaMemcpyAsync(e.h_out, e.pts.ptr, out_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(e.h_status, e.pr_errs.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    codes[0] = *reinterpret_cast<int*>(e.h_status);
    if (codes[0] == E_SUCCESS) memcpy(outs, e.h_out, out_bytes);
    return E_SUCCESS;
  }
  CUDA_TRY(cudaMemcpyAsync(outs, e.pts.ptr, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(codes, e.pr_errs.ptr, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return E_SUCCESS;
}
extern "C" int b200_map_host(int group, const unsigned char* in, unsigned char* out) {
  int code = E_MEMORY;
  int rc = group == 1 ? map_host_impl<Fp>(out, &code, in, 1) : map_host_impl<Fp2>(out, &code, in, 1);
  return rc ? rc : code;
}
extern "C" EIP2537_ERROR bls12_map_fp_to_g1_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, size_t n) {
  return (EIP2537_ERROR)map_host_impl<Fp>(outs, reinterpret_cast<int*>(errs), in, n);
}
extern "C" EIP2537_ERROR bls12_map_fp2_to_g2_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, size_t n) {
  return (EIP2537_ERROR)map_host_impl<Fp2>(outs, reinterpret_cast<int*>(errs), in, n);
}

// ------------------------------------------------------------------------------------------
// workload generators: out[i] = encode(k_i * generator)
// ------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_generator_mul(const uint32_t* __restrict__ scalars, size_t n, uint32_t* __restrict__ out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k[8];
  scalar_from_slot(k, scalars + 8 * i);
  Affine<F> g;
  const uint32_t* gc = sizeof(F) == sizeof(Fp) ? C_G1GEN() : C_G2GEN();
  uint32_t* gw = reinterpret_cast<uint32_t*>(&g);
  for (int t = 0; t < (int)(sizeof(Affine<F>) / 4); t++) gw[t] = gc[t];
  XYZZ<F> acc = xyzz_scalar_mul(g, k, 256);
  Affine<F> a = xyzz_to_affine(acc);
  uint32_t w[Wire<F>::POINT_WORDS];
  encode_point(w, a);
  for (int t = 0; t < Wire<F>::POINT_WORDS; t++) out[i * Wire<F>::POINT_WORDS + t] = w[t];
}

template <class F>
static int generator_mul_impl(unsigned char* out, const unsigned char* scalars, size_t n) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t out_bytes = n * Wire<F>::POINT_WORDS * 4;
  if ((rc = e.raw.reserve(32 * n))) return rc;
  if ((rc = e.pts.reserve(out_bytes))) return rc;
  cudaStream_t s = e.stream;
  CUDA_TRY(cudaMemcpyAsync(e.raw.ptr, scalars, 32 * n, cudaMemcpyHostToDevice, s));
  LAUNCH(k_generator_mul<F>, blocks_for(n, 128), 128, s, (const uint32_t*)e.raw.ptr, n, (uint32_t*)e.pts.ptr);
  CUDA_TRY(cudaMemcpyAsync(out, e.pts.ptr, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return E_SUCCESS;
}
extern "C" EIP2537_ERROR bls12_b200_g1_generator_mul(byte* out, const byte* scalars, size_t n) {
  return (EIP2537_ERROR)generator_mul_impl<Fp>(out, scalars, n);
}
extern "C" EIP2537_ERROR bls12_b200_g2_generator_mul(byte* out, const byte* scalars, size_t n) {
  return (EIP2537_ERROR)generator_mul_impl<Fp2>(out, scalars, n);
}

// ------------------------------------------------------------------------------------------
// pairing
// ------------------------------------------------------------------------------------------
// The exponent chain of final_exp() (pairing.cuh) as a program for k_pairing_final_dot6: three shared-memory buffers
// (0..2), FX_PARK_SLOTS values parked in HBM.  Input: the product of the call's chunk values in buffer 0; output in 0.
static std::vector<uint32_t> final_exp_program() {
  std::vector<uint32_t> p;
  auto emit = [&](uint32_t op, uint32_t d, uint32_t x, uint32_t y) { p.push_back(fx_encode(op, d, x, y)); };
  auto other = [](int a, int b) { return 3 - a - b; };
  // acc <- a^|z| conjugated (z < 0), a in buffer `a`; uses both other buffers; returns the buffer holding the result
  auto exp_z = [&](int a) {
    int acc = (a + 1) % 3, tmp = (a + 2) % 3;
    emit(FX_COPY, acc, a, 0);
    for (int i = 62; i >= 0; i--) {
      emit(FX_CYC, tmp, acc, acc); std::swap(acc, tmp);
      if ((B200_Z_ABS >> i) & 1) { emit(FX_MUL, tmp, acc, a); std::swap(acc, tmp); }
    }
    emit(FX_CONJ, 0, acc, 0);
    return acc;
  };
  // easy part: f = fin^((p^6 - 1)(p^2 + 1))
  emit(FX_INV, 1, 0, 0);            // 1 = fin^-1
  emit(FX_CONJ, 0, 0, 0);           // 0 = conj(fin)
  emit(FX_MUL, 2, 0, 1);            // 2 = fin^(p^6-1)
  emit(FX_COPY, 0, 2, 0); emit(FX_FROB2, 0, 0, 0);
  emit(FX_MUL, 1, 0, 2);            // 1 = f
  emit(FX_PARK, 0, 1, 0);           // slot 0 = f
  // hard part: (z-1)^2 (z+p) (z^2 + p^2 - 1) + 3
  int a = 1, acc, r;
  acc = exp_z(a); emit(FX_CONJ, 0, a, 0); r = other(a, acc); emit(FX_MUL, r, acc, a); a = r;      // f^(z-1)
  acc = exp_z(a); emit(FX_CONJ, 0, a, 0); r = other(a, acc); emit(FX_MUL, r, acc, a); a = r;      // ^(z-1)
  acc = exp_z(a); emit(FX_FROB1, 0, a, 0); r = other(a, acc); emit(FX_MUL, r, acc, a); a = r;     // ^(z+p)  -> h
  emit(FX_PARK, 0, a, 1);           // slot 1 = h
  acc = exp_z(a); a = acc;
  acc = exp_z(a);                   // h^(z^2)
  r = (acc + 1) % 3;
  emit(FX_UNPARK, r, 0, 1); emit(FX_FROB2, 0, r, 0);
  { int t = other(acc, r); emit(FX_MUL, t, acc, r); acc = t; }
  r = (acc + 1) % 3;
  emit(FX_UNPARK, r, 0, 1); emit(FX_CONJ, 0, r, 0);
  { int t = other(acc, r); emit(FX_MUL, t, acc, r); acc = t; }                                  // h^(z^2 + p^2 - 1)
  emit(FX_PARK, 0, acc, 1);         // slot 1 = that
  emit(FX_UNPARK, 0, 0, 0);         // 0 = f
  emit(FX_CYC, 1, 0, 0);            // 1 = f^2
  emit(FX_MUL, 2, 1, 0);            // 2 = f^3
  emit(FX_UNPARK, 0, 0, 1);
  emit(FX_MUL, 1, 0, 2);
  emit(FX_COPY, 0, 1, 0);           // result in buffer 0
  return p;
}

static std::atomic<const PairingPlanState*> g_last_plan{nullptr};
static std::atomic<int> g_last_plan_device{0};

static int pairing_batch_device_impl(Engine& e, const uint32_t* d_raw, const unsigned long long* d_offsets, size_t n_calls,
                                     size_t total_pairs, uint32_t* d_outs, int* d_errs, cudaStream_t s) {
  int rc;
  const size_t max_tasks = total_pairs;    // every chunk holds at least one pair
  if (total_pairs > (4u << 20)) { snprintf(g_last_error, sizeof g_last_error, "pairing batch too large: split it"); return E_MEMORY; }
  constexpr size_t PLAN_BYTES = 2048;
  static_assert(sizeof(PairingPlanState) <= PLAN_BYTES, "plan header");
  if ((rc = e.pr_g1.reserve(total_pairs * sizeof(G1Affine)))) return rc;
  if ((rc = e.pr_g2.reserve(total_pairs * sizeof(G2Affine)))) return rc;
  if ((rc = e.pr_status.reserve(total_pairs * sizeof(int) + total_pairs + 64))) return rc;
  // dot engine: pair slots are grouped by position inside the chunk, each group padded to a multiple of 32 slots
  const size_t slot_stride = ((total_pairs + 31) & ~(size_t)31) + 32 * PAIRING_MAX_CHUNK;
  if ((rc = e.pr_lines.reserve((size_t)ML_STEPS * slot_stride * sizeof(Line)))) return rc;
  if ((rc = e.pr_tasks.reserve(PLAN_BYTES + max_tasks * sizeof(PairingTask) + n_calls * sizeof(uint32_t) + 64))) return rc;
  if ((rc = e.pr_f.reserve(max_tasks * sizeof(Fp12)))) return rc;
  if ((rc = e.pr_slots.reserve(slot_stride * sizeof(uint32_t) + slot_stride + 64))) return rc;
  G1Affine* g1 = (G1Affine*)e.pr_g1.ptr;
  G2Affine* g2 = (G2Affine*)e.pr_g2.ptr;
  int* pstat = (int*)e.pr_status.ptr;
  unsigned char* skip = (unsigned char*)(pstat + total_pairs);
  Line* lines = (Line*)e.pr_lines.ptr;
  PairingPlanState* plan_state = (PairingPlanState*)e.pr_tasks.ptr;
  PairingTask* tasks = (PairingTask*)((char*)e.pr_tasks.ptr + PLAN_BYTES);
  uint32_t* call_first = (uint32_t*)(tasks + max_tasks);
  Fp12* f = (Fp12*)e.pr_f.ptr;
  uint32_t* slot_pair = (uint32_t*)e.pr_slots.ptr;
  unsigned char* skip_slot = (unsigned char*)(slot_pair + slot_stride);
  // Pairs per chunk: chosen on the device from the batch's own shape (pairing_choose_chunk); B200_PAIRING_CHUNK forces it.
  static const int forced_chunk = getenv("B200_PAIRING_CHUNK") ? atoi(getenv("B200_PAIRING_CHUNK")) : 0;
  // B200_PAIRING_ACC=thread selects round 1's thread-per-chunk accumulate (Fp12 in thread-local memory) for A/B runs
  static const bool use_dot = !(getenv("B200_PAIRING_ACC") && !strcmp(getenv("B200_PAIRING_ACC"), "thread"));
  const uint32_t wave_thread = (uint32_t)e.sm_count * 6 * 64;                     // k_pairing_accumulate: 6 blocks of 64 threads per SM
  // B200_DOT_WAVE_PCT: the planner's idea of "one wave" as a percentage of the resident chunk slots (developer sweep)
  static const int wave_pct = getenv("B200_DOT_WAVE_PCT") ? atoi(getenv("B200_DOT_WAVE_PCT")) : 180;
  const uint32_t fc = (uint32_t)(forced_chunk > 0 ? forced_chunk : 0);
  CUDA_TRY(cudaMemsetAsync(plan_state, 0, sizeof(PairingPlanState), s));
  g_last_plan.store(plan_state); g_last_plan_device.store(e.device);
  g_pstage.mark(0, s);
  const bool small_batch = (long)n_calls <= g_pairing_coop_max.load();
  const bool dot_path = use_dot && !small_batch;
  // few pairs in all: decode, both membership tests and the line functions of a pair by lane groups (pairing_coop.cuh)
  static const long pair_coop_max = getenv("B200_PAIR_COOP_MAX") ? atol(getenv("B200_PAIR_COOP_MAX")) : 2048;
  const bool pair_coop = small_batch && (long)total_pairs <= pair_coop_max;
  if (pair_coop) {
    if (total_pairs) LAUNCH(k_pairing_pair_coop, (unsigned)total_pairs, 64, s, d_raw, total_pairs, lines, skip, pstat);
  } else if (small_batch) {
    LAUNCH(k_pairing_decode_split, blocks_for(total_pairs, 32), 64, s, d_raw, total_pairs, g1, g2, pstat);
  } else if (dot_path) {     // G2 membership is decided by the line kernel (its walk of [|z|]Q is that test's ladder)
    LAUNCH(k_pairing_decode<false>, blocks_for(total_pairs, 64), 64, s, d_raw, total_pairs, g1, g2, pstat);
  } else {
    LAUNCH(k_pairing_decode<true>, blocks_for(total_pairs, 64), 64, s, d_raw, total_pairs, g1, g2, pstat);
  }
  g_pstage.mark(1, s);
  // Back ends for "multiply the lines into f, final exponentiation, is-one":
  //   warp-cooperative (coop12.cuh): one warp per call, Fp12 in shared memory -- lowest latency, small batches;
  //   dot engine (pairing_dot.cuh): three lanes per chunk, Fp12 in shared memory, one reduction per output
  //     component -- the throughput path;
  //   thread-per-chunk (pairing.cuh, round 1): kept for small batches of long calls and for A/B measurements.
  //   A lone warp walks its call's pairs one after the other (~0.55 ms per pair), so small batches of LONG calls
  //   (more than ~12 pairs per call) accumulate per chunk in parallel first and give the warp only the product of
  //   the chunk values and the final exponentiation.
  if (small_batch && total_pairs <= 12 * n_calls) {
    if (!pair_coop) LAUNCH(k_pairing_lines, blocks_for(total_pairs, 64), 64, s, g1, g2, pstat, total_pairs, lines, skip);
    g_pstage.mark(2, s);
    LAUNCH(k_pairing_count, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, pstat, plan_state, d_errs, (uint32_t)1);
    g_pstage.mark(3, s);
    LAUNCH(k_pairing_call_coop, (unsigned)n_calls, 32, s, n_calls, d_offsets, lines, skip, total_pairs, d_outs, d_errs);
  } else if (!use_dot || small_batch) {
    if (!pair_coop) LAUNCH(k_pairing_lines, blocks_for(total_pairs, 64), 64, s, g1, g2, pstat, total_pairs, lines, skip);
    g_pstage.mark(2, s);
    LAUNCH(k_pairing_count, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, pstat, plan_state, d_errs, (uint32_t)PAIRING_MAX_CHUNK_THREAD);
    LAUNCH(k_pairing_plan, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, d_errs, wave_thread, fc, (uint32_t)PAIRING_MAX_CHUNK_THREAD,
           (uint32_t)85, plan_state, tasks, call_first, (uint32_t*)nullptr);
    LAUNCH(k_pairing_accumulate, blocks_for(max_tasks, 64), 64, s, tasks, plan_state, lines, skip, total_pairs, f);
    g_pstage.mark(3, s);
    if (small_batch) LAUNCH(k_pairing_call_coop_chunks, (unsigned)n_calls, 32, s, n_calls, d_offsets, plan_state, call_first, f, d_outs, d_errs);
    else             LAUNCH(k_pairing_calls, blocks_for(n_calls, 64), 64, s, n_calls, d_offsets, plan_state, call_first, f, d_outs, d_errs);
  } else {
    static const int dot6_blocks = getenv("B200_DOT6_BLOCKS") ? atoi(getenv("B200_DOT6_BLOCKS")) : DOT6_BLOCKS_PER_SM;
    const uint32_t wave = (uint32_t)((uint64_t)e.sm_count * (dot6_blocks == 4 ? 4 : 3) * 32 * wave_pct / 100);
    CUDA_TRY(cudaMemsetAsync(slot_pair, 0xFF, slot_stride * sizeof(uint32_t), s));
    LAUNCH(k_pairing_count, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, pstat, plan_state, d_errs, (uint32_t)PAIRING_MAX_CHUNK);
    LAUNCH(k_pairing_early_fix, blocks_for(n_calls, 64), 64, s, d_offsets, n_calls, g2, pstat, d_errs);
    LAUNCH(k_pairing_plan, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, d_errs, wave, fc, (uint32_t)PAIRING_MAX_CHUNK,
           (uint32_t)97, plan_state, tasks, call_first, slot_pair);
    LAUNCH(k_pairing_lines_slots, blocks_for(slot_stride, 64), 64, s, g1, g2, plan_state, slot_pair, slot_stride, (uint32_t*)lines, skip_slot, pstat);
    LAUNCH(k_pairing_late_errs, blocks_for(n_calls, 128), 128, s, d_offsets, n_calls, pstat, d_errs);
    g_pstage.mark(2, s);
    const unsigned nb = blocks_for(max_tasks, 32);
    const uint32_t* lt = (const uint32_t*)lines;
    {  // shared-memory opt-ins (once per device): 54 KB of dynamic shared memory per block
      static std::atomic<unsigned> carved{0};
      if (!((carved.load() >> e.device) & 1u)) {
        cudaFuncSetAttribute(k_pairing_accumulate_dot6<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOT6_SMEM_BYTES);
        cudaFuncSetAttribute(k_pairing_accumulate_dot6<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, DOT6_SMEM_BYTES);
        cudaFuncSetAttribute(k_pairing_accumulate_dot6<4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(k_pairing_final_dot6, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * DOT6_F_WORDS * 4);
        (void)cudaGetLastError();
        carved.fetch_or(1u << e.device);
      }
    }
    if (dot6_blocks == 4) LAUNCH_SMEM(k_pairing_accumulate_dot6<4>, nb, DOT6_THREADS, DOT6_SMEM_BYTES, s, tasks, plan_state, lt, skip_slot, slot_stride, f);
    else                  LAUNCH_SMEM(k_pairing_accumulate_dot6<3>, nb, DOT6_THREADS, DOT6_SMEM_BYTES, s, tasks, plan_state, lt, skip_slot, slot_stride, f);
    g_pstage.mark(3, s);
    static const int final_v = getenv("B200_FINAL_V") ? atoi(getenv("B200_FINAL_V")) : 6;   // 1 = round 1's thread-per-call kernel
    if (final_v == 1) {
      LAUNCH(k_pairing_calls, blocks_for(n_calls, 64), 64, s, n_calls, d_offsets, plan_state, call_first, f, d_outs, d_errs);
    } else {
      if (!e.pr_prog_len) {
        static const std::vector<uint32_t> prog = final_exp_program();
        if ((rc = e.pr_prog.reserve(prog.size() * 4))) return rc;
        CUDA_TRY(cudaMemcpyAsync(e.pr_prog.ptr, prog.data(), prog.size() * 4, cudaMemcpyHostToDevice, s));
        e.pr_prog_len = (int)prog.size();
      }
      if ((rc = e.pr_park.reserve((size_t)FX_PARK_SLOTS * n_calls * sizeof(Fp12)))) return rc;
      LAUNCH_SMEM(k_pairing_final_dot6, blocks_for(n_calls, 32), DOT6_THREADS, 3 * DOT6_F_WORDS * 4, s, n_calls, d_offsets, plan_state, call_first,
                  f, (const uint32_t*)e.pr_prog.ptr, e.pr_prog_len, (Fp12*)e.pr_park.ptr, d_outs, d_errs);
    }
  }
  g_pstage.mark(4, s);
  CUDA_TRY(cudaGetLastError());
  return E_SUCCESS;
}

// pairs per chunk the planner chose for the last pairing batch (bench.py: algorithmic work of the accumulate kernel)
extern "C" int bls12_b200_last_pairing_chunk(void) {
  const PairingPlanState* p = g_last_plan.load();
  if (!p) return 0;
  uint32_t chunk = 0;
  if (cudaMemcpy(&chunk, &p->chunk, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return (int)chunk;
}

extern "C" EIP2537_ERROR bls12_b200_pairing_batch_device(const void* d_in, const uint64_t* d_offsets, size_t n,
                                                         size_t total_pairs, void* d_outs, int32_t* d_errs, void* stream) {
  if (n == 0) return EIP2537_SUCCESS;
  cudaStream_t s = (cudaStream_t)stream;
  Lease lease;
  int rc = lease.acquire(-1, true, s);
  if (rc) return (EIP2537_ERROR)rc;
  rc = pairing_batch_device_impl(*lease.e, (const uint32_t*)d_in, (const unsigned long long*)d_offsets, n,
                                 total_pairs, (uint32_t*)d_outs, (int*)d_errs, s);
  lease.submitted(s);
  return (EIP2537_ERROR)rc;
}

// n independent calls on ONE device (device < 0: the current one); host buffers, synchronous
static int pairing_batch_on_device(int device, unsigned char* outs, int* errs, const unsigned char* in, const uint64_t* offsets, size_t n) {
  // per-call length validation on the host (eip2537.c:1022-1024); bad calls become zero-pair calls on the device
  Lease lease;
  int rc = lease.acquire(device);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t total_bytes = (size_t)(offsets[n] - offsets[0]);
  size_t total_pairs = 0, cursor = 0;
  std::vector<unsigned long long> h_off(n + 1);      // compacted offsets: calls with an invalid length contribute no pairs
  bool any_bad = false;
  for (size_t i = 0; i < n; i++) {
    size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    h_off[i] = cursor;
    if (len == 0 || len % 384) any_bad = true;
    else { cursor += len; total_pairs += len / 384; }
  }
  h_off[n] = cursor;
  cudaStream_t s = e.stream;
  if ((rc = e.pr_raw.reserve(cursor + 16)) || (rc = e.pr_offsets.reserve((n + 1) * 8)) ||
      (rc = e.pr_outs.reserve(n * 32)) || (rc = e.pr_errs.reserve(n * 4))) return rc;
  const bool pageable = host_pointer_is_pageable(in);
  if (!any_bad) {
    if ((rc = h2d_copy(e, (unsigned char*)e.pr_raw.ptr, in + offsets[0], total_bytes, pageable, s))) return rc;
  } else {
    for (size_t i = 0; i < n; i++) {
      size_t len = (size_t)(h_off[i + 1] - h_off[i]);
      if (len) CUDA_TRY(cudaMemcpyAsync((char*)e.pr_raw.ptr + h_off[i], in + offsets[i], len, cudaMemcpyHostToDevice, s));
    }
  }
  CUDA_TRY(cudaMemcpyAsync(e.pr_offsets.ptr, h_off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, s));
  if (total_pairs) {
    if ((rc = pairing_batch_device_impl(e, (const uint32_t*)e.pr_raw.ptr, (const unsigned long long*)e.pr_offsets.ptr, n,
                                        total_pairs, (uint32_t*)e.pr_outs.ptr, (int*)e.pr_errs.ptr, s))) return rc;
    CUDA_TRY(cudaMemcpyAsync(outs, e.pr_outs.ptr, n * 32, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(errs, e.pr_errs.ptr, n * 4, cudaMemcpyDeviceToHost, s));
  }
  CUDA_TRY(cudaStreamSynchronize(s));          // h_off and the caller's buffers stay alive until here
  for (size_t i = 0; i < n; i++) {
    size_t len = (size_t)(offsets[i + 1] - offsets[i]);
    if (len == 0 || len % 384) { errs[i] = EIP2537_INVALID_LENGTH; memset(outs + 32 * i, 0, 32); }
    else if (errs[i] != EIP2537_SUCCESS) memset(outs + 32 * i, 0, 32);
  }
  return E_SUCCESS;
}

// Calls are independent units: after bls12_b200_init_multi(G) a large batch is cut into G contiguous groups of calls
// (balanced by pair count), one host thread per device, no collective.
extern "C" EIP2537_ERROR bls12_pairing_batch(byte* outs, EIP2537_ERROR* errs, const byte* in, const uint64_t* offsets, size_t n) {
  if (n == 0) return EIP2537_SUCCESS;
  int G = g_ngpu.load();
  if ((size_t)G > n / 256) G = (int)(n / 256);
  if (G <= 1) return (EIP2537_ERROR)pairing_batch_on_device(-1, outs, (int*)errs, in, offsets, n);
  std::vector<size_t> cut(G + 1, n);
  cut[0] = 0;
  const uint64_t total = offsets[n] - offsets[0];
  for (int d = 1; d < G; d++) {          // first call whose start is past d/G of the bytes
    const uint64_t want = offsets[0] + total * d / G;
    size_t lo = cut[d - 1], hi = n;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (offsets[mid] < want) lo = mid + 1; else hi = mid; }
    cut[d] = lo;
  }
  std::vector<int> rcs(G, 0);
  std::vector<std::string> msgs(G);
  auto work = [&](int d) {
    const size_t lo = cut[d], hi = cut[d + 1];
    if (lo >= hi) return;
    rcs[d] = pairing_batch_on_device(d, outs + 32 * lo, (int*)errs + lo, in, offsets + lo, hi - lo);
    if (rcs[d]) msgs[d] = g_last_error;
  };
  std::vector<std::thread> th;
  for (int d = 1; d < G; d++) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < G; d++)
    if (rcs[d]) { snprintf(g_last_error, sizeof g_last_error, "device %d: %s", d, msgs[d].c_str()); return (EIP2537_ERROR)rcs[d]; }
  return EIP2537_SUCCESS;
}

extern "C" int b200_pairing_host(const unsigned char* in, size_t k, unsigned char* out) {
  uint64_t offs[2] = {0, (uint64_t)k * 384};
  unsigned char res[32];
  EIP2537_ERROR err = EIP2537_SUCCESS;
  EIP2537_ERROR rc = bls12_pairing_batch(res, &err, in, offs, 1);
  if (rc) return rc;
  if (err) return err;
  memcpy(out, res, 32);
  return E_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// K1 microbenchmarks (roofline denominator: measured MAC32/s; Fp-mul throughput)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fp_chain(int iters, Fp* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_one(), y = fp_load_const(C_G1GEN());
  x.v[0] ^= (uint32_t)i;
  x = fp_reduce_once(x, 0);
  for (int k = 0; k < iters; k++) x = mul(x, y);
  // keep one result per block alive
  if (threadIdx.x == 0) out[blockIdx.x] = x;
  else if (x.v[0] == 0xdeadbeefu && x.v[5] == 0x12345u) out[blockIdx.x] = x;
}
// raw integer-multiply pipe probe: 16 independent 64-bit accumulators per thread, each
// acc = a*b + acc (mad.wide.u32 -> IMAD.WIDE.U32).  The multiplier changes every pass (xor with an
// accumulator word), otherwise ptxas hoists the loop-invariant products and the loop degenerates into
// 64-bit adds (an earlier version of this probe did exactly that and over-reported the peak 2x).
__global__ void __launch_bounds__(256) k_imad_peak(int iters, unsigned long long* out) {
  unsigned a[16], b = blockIdx.x * 40503u + 7;
  unsigned long long c[16];
#pragma unroll
  for (int u = 0; u < 16; u++) { c[u] = u + 1; a[u] = threadIdx.x * 2654435761u + 977u * u; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int u = 0; u < 16; u++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(c[u]) : "r"(a[u]), "r"(b));
      b ^= (unsigned)c[r];
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int u = 0; u < 16; u++) r ^= c[u];
  if (r == 0x123456789ull || threadIdx.x == 0) out[blockIdx.x] = r;
}
// same with separate 32-bit halves: lo = a*b + lo (IMAD), hi = mulhi(a,b) + hi (IMAD.HI), no carry between them
__global__ void __launch_bounds__(256) k_imad32_peak(int iters, unsigned long long* out) {
  unsigned a[16], b = blockIdx.x * 40503u + 7, lo[16], hi[16];
#pragma unroll
  for (int u = 0; u < 16; u++) { lo[u] = u + 1; hi[u] = 3 * u; a[u] = threadIdx.x * 2654435761u + 977u * u; }
  for (int k = 0; k < iters; k++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int u = 0; u < 16; u++) {
        asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[u]) : "r"(a[u]), "r"(b));
        asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(hi[u]) : "r"(a[u]), "r"(b));
      }
      b ^= lo[r];
    }
  }
  unsigned long long r = 0;
#pragma unroll
  for (int u = 0; u < 16; u++) r ^= lo[u] + ((unsigned long long)hi[u] << 32);
  if (r == 0x123456789ull || threadIdx.x == 0) out[blockIdx.x] = r;
}

// carry-chain probe: 4 independent 13-word accumulators, each fed by mad.lo.cc/madc.hi.cc rows
// (the IMAD.WIDE.U32.X form the Montgomery multiply uses): 24 MAC32 per iteration per thread
__global__ void __launch_bounds__(256) k_imad_carry_probe(int iters, unsigned long long* out) {
  uint32_t a[12], acc[4][13];
#pragma unroll
  for (int i = 0; i < 12; i++) a[i] = threadIdx.x * 2654435761u + i;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 13; i++) acc[k][i] = k + i;
  uint32_t b = blockIdx.x * 40503u + 7;
#ifdef __CUDA_ARCH__
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < 4; k++) b200::detail::mad_row(acc[k], a, b);
  }
#endif
  unsigned long long r = 0;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int i = 0; i < 13; i++) r += acc[k][i];
  if (r == 0x123456789ull || threadIdx.x == 0) out[blockIdx.x] = r;
}

// dependent multiplications x <- x*y with a PER-THREAD multiplier (y changes too), so the multiplier limbs live in
// vector registers as they do in the point formulas (k_fp_chain's constant multiplier sits in uniform registers)
template <int VARIANT>
__global__ void __launch_bounds__(256) k_fp_chain_v(int iters, Fp* out) {
  __shared__ uint32_t slots[12 * 256];
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_one(), y = fp_load_const(C_G1GEN());
  x.v[0] ^= (uint32_t)i; y.v[1] ^= (uint32_t)i;
  x = fp_reduce_once(x, 0);
  for (int k = 0; k < iters; k++) {
    Fp t = x;
#ifdef __CUDA_ARCH__
    if (VARIANT == 0)      t = mul(x, y);
    else if (VARIANT == 1) t = mul_unrolled<4>(x, y);
    else if (VARIANT == 2) t = mul_unrolled<6>(x, y);
    else if (VARIANT == 3) t = mul_unrolled<12>(x, y);
    else if (VARIANT == 4) t = mul_bsmem(x, y, slots + threadIdx.x, 256);
    else if (VARIANT == 5) t = mul_unrolled<2, true>(x, y);
    else                   t = mul_unrolled<6, true>(x, y);
#endif
    y = x; x = t;
  }
  if (threadIdx.x == 0) out[blockIdx.x] = x;
  else if (x.v[0] == 0xdeadbeefu && x.v[5] == 0x12345u) out[blockIdx.x] = x;
}

// dot engine in isolation: per iteration 6 double-width products + ONE reduction (1020 MAC32), operands in registers
template <class ACC>
__global__ void __launch_bounds__(256) k_dot_chain(int iters, Fp* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_one(), y = fp_load_const(C_G1GEN());
  x.v[0] ^= (uint32_t)i;
  x = fp_reduce_once(x, 0);
  for (int k = 0; k < iters; k++) {
    ACC A;
    dot::acc_zero(A);
#pragma unroll 1
    for (int t = 0; t < 6; t++) { dot::acc_product(A, x, y); y.v[0] ^= 1u; }
    x = dot::acc_reduce(A, 1);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = x;
  else if (x.v[0] == 0xdeadbeefu && x.v[5] == 0x12345u) out[blockIdx.x] = x;
}

// ONE double-width product + separated reduction per dependent step: the low-latency form of a field multiplication
// (the product rows are independent carry chains; only the reduction rows are serial)
template <class ACC>
__global__ void __launch_bounds__(256) k_ll_chain(int iters, Fp* out) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_one(), y = fp_load_const(C_G1GEN());
  x.v[0] ^= (uint32_t)i;
  x = fp_reduce_once(x, 0);
#pragma unroll 1
  for (int k = 0; k < iters; k++) {
    ACC A;
    dot::acc_zero(A);
    dot::acc_product(A, x, y);
    x = dot::acc_reduce(A, 1);
  }
  if (threadIdx.x == 0) out[blockIdx.x] = x;
  else if (x.v[0] == 0xdeadbeefu && x.v[5] == 0x12345u) out[blockIdx.x] = x;
}

extern "C" EIP2537_ERROR bls12_b200_fp_microbench(int mode, size_t n_threads, int iters, float* ms, byte* digest48) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return (EIP2537_ERROR)rc;
  Engine& e = *lease.e;
  unsigned nblk = blocks_for(n_threads, 256);
  if ((rc = e.pts.reserve((size_t)nblk * sizeof(Fp)))) return (EIP2537_ERROR)rc;
  cudaEvent_t t0, t1;
  CUDA_TRY2(cudaEventCreate(&t0));
  CUDA_TRY2(cudaEventCreate(&t1));
  cudaStream_t s = e.stream;
  for (int rep = 0; rep < 2; rep++) {   // first pass warms up
    CUDA_TRY2(cudaEventRecord(t0, s));
    if (mode == 0)      LAUNCH(k_fp_chain, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 1) LAUNCH(k_imad_peak, nblk, 256, s, iters, (unsigned long long*)e.pts.ptr);
    else if (mode == 2) LAUNCH(k_imad_carry_probe, nblk, 256, s, iters, (unsigned long long*)e.pts.ptr);
    else if (mode == 4) LAUNCH(k_dot_chain<dot::Acc>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 6) LAUNCH(k_fp_chain_v<0>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 7) LAUNCH(k_fp_chain_v<1>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 8) LAUNCH(k_fp_chain_v<2>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 9) LAUNCH(k_fp_chain_v<3>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 10) LAUNCH(k_fp_chain_v<4>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 11) LAUNCH(k_fp_chain_v<5>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 12) LAUNCH(k_fp_chain_v<6>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 5) LAUNCH(k_dot_chain<dot::Acc64>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 13) LAUNCH(k_ll_chain<dot::Acc64>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else if (mode == 14) LAUNCH(k_ll_chain<dot::Acc>, nblk, 256, s, iters, (Fp*)e.pts.ptr);
    else                LAUNCH(k_imad32_peak, nblk, 256, s, iters, (unsigned long long*)e.pts.ptr);
    CUDA_TRY2(cudaEventRecord(t1, s));
    CUDA_TRY2(cudaStreamSynchronize(s));
  }
  CUDA_TRY2(cudaEventElapsedTime(ms, t0, t1));
  if (digest48) CUDA_TRY2(cudaMemcpy(digest48, e.pts.ptr, 48, cudaMemcpyDeviceToHost));
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  return EIP2537_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// on-device self test: PTX carry-chain field ops vs portable 64-bit C++ on the same inputs
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ Fp st_add_ref(const Fp& a, const Fp& b, bool subtract) {
  const uint32_t* p = C_P();
  uint32_t t[12];
  uint64_t c = 0;
  if (!subtract) {
    for (int i = 0; i < 12; i++) { uint64_t s = (uint64_t)a.v[i] + b.v[i] + c; t[i] = (uint32_t)s; c = s >> 32; }
    uint32_t u[12]; uint64_t bw = 0;
    for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)t[i] - p[i] - bw; u[i] = (uint32_t)d; bw = (d >> 32) & 1; }
    Fp r; for (int i = 0; i < 12; i++) r.v[i] = bw ? t[i] : u[i];
    return r;
  }
  uint64_t bw = 0;
  for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)a.v[i] - b.v[i] - bw; t[i] = (uint32_t)d; bw = (d >> 32) & 1; }
  if (bw) { c = 0; for (int i = 0; i < 12; i++) { uint64_t s = (uint64_t)t[i] + p[i] + c; t[i] = (uint32_t)s; c = s >> 32; } }
  Fp r; for (int i = 0; i < 12; i++) r.v[i] = t[i];
  return r;
}
__global__ void __launch_bounds__(128) k_selftest(size_t n, unsigned long long* mismatches) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // two pseudo-random field elements (< p: top limb masked below p's top limb 0x1a0111ea)
  unsigned long long s = 0x9E3779B97F4A7C15ull * (i + 1);
  Fp a, b;
  for (int k = 0; k < 12; k++) {
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27; a.v[k] = (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27; b.v[k] = (uint32_t)((s * 0x2545F4914F6CDD1Dull) >> 32);
  }
  a.v[11] &= 0x0fffffffu; b.v[11] &= 0x0fffffffu;
  if ((i & 7) == 0) { a = fp_load_const(C_P()); a.v[0] -= 1; }            // p - 1
  if ((i & 7) == 1) { b = fp_zero(); }
  if ((i & 7) == 2) { b = a; }
  if (!eq(mul(a, b), mul_portable(a, b))) atomicAdd(&mismatches[0], 1ull);
  {  // fused a*b - c*d (one reduction) against the two-product form, including the extreme operands p-1 and 1
    Fp pm1 = fp_load_const(C_P()); pm1.v[0] -= 1;
    Fp one_raw = fp_zero(); one_raw.v[0] = 1;
    Fp c2 = (i & 3) == 0 ? one_raw : ((i & 3) == 1 ? pm1 : b);
    Fp d2 = (i & 4) ? pm1 : a;
    Fp a2 = (i & 8) ? pm1 : a, b2 = (i & 16) ? pm1 : b;
    if (!eq(mul_diff(a2, b2, c2, d2), sub(mul_portable(a2, b2), mul_portable(c2, d2)))) atomicAdd(&mismatches[0], 1ull);
  }
  {  // dot engine (dot12.cuh): 1..12 double-width products accumulated, ONE reduction, against the sum of portable products
    Fp pm1 = fp_load_const(C_P()); pm1.v[0] -= 1;
    dot::Acc A;
    dot::acc_zero(A);
    Fp sum = fp_zero(), aa = a, bb = b;
    const int nt = 1 + (int)(i % 12);
    for (int t = 0; t < nt; t++) {
      if ((i & 31) == 7) { aa = pm1; bb = pm1; }
      dot::acc_product(A, aa, bb);
      sum = add(sum, mul_portable(aa, bb));
      aa = add(aa, b); bb = sub(bb, a);
    }
    if (!eq(dot::acc_reduce(A, nt > 9 ? 2 : 1), sum)) atomicAdd(&mismatches[0], 1ull);
  }
  if (!eq(add(a, b), st_add_ref(a, b, false))) atomicAdd(&mismatches[1], 1ull);
  if (!eq(sub(a, b), st_add_ref(a, b, true))) atomicAdd(&mismatches[2], 1ull);
  Fp q = mul(a, b);
  if (!is_zero(a) && !is_zero(b) && (i & 63) == 5) { if (!eq(mul(mul(q, inv(b)), fp_load_const(C_RR())), mul(a, fp_load_const(C_RR())))) atomicAdd(&mismatches[3], 1ull); }
}
extern "C" EIP2537_ERROR bls12_b200_selftest(uint64_t* mismatches4, size_t n) {
  Lease lease;
  int rc = lease.acquire(-1);
  if (rc) return (EIP2537_ERROR)rc;
  Engine& e = *lease.e;
  if ((rc = e.status.reserve(64))) return (EIP2537_ERROR)rc;
  CUDA_TRY2(cudaMemsetAsync(e.status.ptr, 0, 32, e.stream));
  LAUNCH(k_selftest, blocks_for(n, 128), 128, e.stream, n, (unsigned long long*)e.status.ptr);
  CUDA_TRY2(cudaMemcpyAsync(mismatches4, e.status.ptr, 32, cudaMemcpyDeviceToHost, e.stream));
  CUDA_TRY2(cudaStreamSynchronize(e.stream));
  return EIP2537_SUCCESS;
}

// ------------------------------------------------------------------------------------------
// lifetime / introspection
// ------------------------------------------------------------------------------------------
extern "C" EIP2537_ERROR bls12_b200_init(int device) {
  Lease lease;                       // creates the device's first workspace (streams, events, pinned words)
  return (EIP2537_ERROR)lease.acquire(device);
}
// Releases every idle workspace of every device.  Must not race with calls in flight (the caller's contract, as
// for any library teardown); workspaces leased at this moment are returned to the pool afterwards and stay usable.
extern "C" void bls12_b200_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  bls12_b200_comm_destroy();
  int prev = 0;
  if (cudaGetDevice(&prev) != cudaSuccess) { (void)cudaGetLastError(); return; }
  for (int d = 0; d < MAX_DEVICES; d++) {
    DevicePool& p = g_pools[d];
    if (!p.ready.load(std::memory_order_acquire)) continue;
    std::vector<Engine*> idle;
    { std::lock_guard<std::mutex> lk2(p.mu); idle.swap(p.idle); p.created -= (int)idle.size(); }
    cudaSetDevice(p.device);
    for (Engine* e : idle) { engine_quiesce(e); engine_destroy(e); }
  }
  cudaSetDevice(prev);
  g_last_plan.store(nullptr);
}
extern "C" const char* bls12_b200_last_error(void) { return g_last_error; }
extern "C" uint64_t bls12_b200_launch_count(void) { return g_launches.load(); }
extern "C" void bls12_b200_set_window(int c) { g_forced_window.store(c); }
extern "C" void bls12_b200_set_checked_msm(int on) { g_checked_msm.store(on); }
extern "C" long bls12_b200_set_pairing_coop_max(long n_calls) {
  return g_pairing_coop_max.exchange(n_calls < 0 ? pairing_coop_default() : n_calls);
}

// batched point validation: device-resident and host-buffer forms
extern "C" EIP2537_ERROR bls12_b200_points_check_device(int group, const void* d_points, size_t n, size_t stride_bytes,
                                                        int check_subgroup, int32_t* d_codes, void* stream) {
  if (n == 0) return EIP2537_SUCCESS;
  if (((uintptr_t)d_points & 15) || (stride_bytes & 15)) {
    snprintf(g_last_error, sizeof g_last_error, "points and stride must be 16-byte aligned");
    return EIP2537_MEMORY_ERROR;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (group == 1) LAUNCH(k_points_check<Fp>, blocks_for(n, 64), 64, s, (const uint32_t*)d_points, n, (int)(stride_bytes / 4), check_subgroup, (int*)d_codes);
  else            LAUNCH(k_points_check<Fp2>, blocks_for(n, 64), 64, s, (const uint32_t*)d_points, n, (int)(stride_bytes / 4), check_subgroup, (int*)d_codes);
  CUDA_TRY2(cudaGetLastError());
  return EIP2537_SUCCESS;
}
static int points_check_on_device(int device, int group, const byte* points, size_t n, size_t stride_bytes, int check_subgroup, int32_t* codes) {
  Lease lease;
  int rc = lease.acquire(device);
  if (rc) return rc;
  Engine& e = *lease.e;
  const size_t bytes = n * stride_bytes;
  if ((rc = e.raw.reserve(bytes)) || (rc = e.pr_status.reserve(n * sizeof(int)))) return rc;
  cudaStream_t s = e.stream;
  if ((rc = h2d_copy(e, (unsigned char*)e.raw.ptr, points, bytes, host_pointer_is_pageable(points), s))) return rc;
  EIP2537_ERROR r = bls12_b200_points_check_device(group, e.raw.ptr, n, stride_bytes, check_subgroup, (int32_t*)e.pr_status.ptr, (void*)s);
  if (r) return r;
  CUDA_TRY(cudaMemcpyAsync(codes, e.pr_status.ptr, n * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return E_SUCCESS;
}
// independent per point: after bls12_b200_init_multi(G) a large array is cut into G contiguous ranges, one host thread
// and device each (SURVEY.md 8(e) "batched decode / subgroup checks": no exchange at all, the codes land in place)
extern "C" EIP2537_ERROR bls12_b200_points_check(int group, const byte* points, size_t n, size_t stride_bytes,
                                                 int check_subgroup, int32_t* codes) {
  if (n == 0) return EIP2537_SUCCESS;
  if (stride_bytes & 15) { snprintf(g_last_error, sizeof g_last_error, "stride must be a multiple of 16"); return EIP2537_MEMORY_ERROR; }
  int G = g_ngpu.load();
  if ((size_t)G > n / 4096) G = (int)(n / 4096);
  if (G <= 1) return (EIP2537_ERROR)points_check_on_device(-1, group, points, n, stride_bytes, check_subgroup, codes);
  std::vector<int> rcs(G, 0);
  std::vector<std::string> msgs(G);
  auto work = [&](int d) {
    const size_t lo = n * (size_t)d / G, hi = n * (size_t)(d + 1) / G;
    rcs[d] = points_check_on_device(d, group, points + lo * stride_bytes, hi - lo, stride_bytes, check_subgroup, codes + lo);
    if (rcs[d]) msgs[d] = g_last_error;
  };
  std::vector<std::thread> th;
  for (int d = 1; d < G; d++) th.emplace_back(work, d);
  work(0);
  for (auto& t : th) t.join();
  for (int d = 0; d < G; d++)
    if (rcs[d]) { snprintf(g_last_error, sizeof g_last_error, "device %d: %s", d, msgs[d].c_str()); return (EIP2537_ERROR)rcs[d]; }
  return EIP2537_SUCCESS;
}
// stage_ms4 = {decode + subgroup checks, line functions, chunked accumulate, product + final exp}
extern "C" EIP2537_ERROR bls12_b200_last_pairing_profile(float* stage_ms4) {
  if (g_pstage.used < 5) return EIP2537_EMPTY_INPUT;
  CUDA_TRY2(cudaEventSynchronize(g_pstage.ev[4]));
  for (int i = 0; i < 4; i++) CUDA_TRY2(cudaEventElapsedTime(&stage_ms4[i], g_pstage.ev[i], g_pstage.ev[i + 1]));
  return EIP2537_SUCCESS;
}
extern "C" size_t bls12_b200_partial_bytes(int group) { return group == 1 ? sizeof(XYZZ<Fp>) : sizeof(XYZZ<Fp2>); }
extern "C" void bls12_b200_set_profile(int on) { g_profile.store(on); }
// stage_ms[4] = {decode+digits+sort, accumulate, bucket reduce tree, window combine} of the last profiled MSM
extern "C" EIP2537_ERROR bls12_b200_last_msm_profile(float* stage_ms4, uint64_t* nonzero_digits) {
  if (g_stage.used < 5) return EIP2537_EMPTY_INPUT;
  CUDA_TRY2(cudaEventSynchronize(g_stage.ev[4]));
  for (int i = 0; i < 4; i++) CUDA_TRY2(cudaEventElapsedTime(&stage_ms4[i], g_stage.ev[i], g_stage.ev[i + 1]));
  msm_count_work();
  *nonzero_digits = g_last_entries;
  return EIP2537_SUCCESS;
}
// work3 = {non-zero digits D, additions done pairwise in affine form (6 Fp-mul each), additions of the XYZZ walk (10 / 28 Fp-mul)}
extern "C" EIP2537_ERROR bls12_b200_last_msm_work(uint64_t* work3) {
  msm_count_work();
  work3[0] = g_last_entries; work3[1] = g_last_affine_adds; work3[2] = g_last_walk_adds;
  return EIP2537_SUCCESS;
}
