// C ABI: lp_multi -- the path-tracing frame on several B200s of one box.
//
// The reference is single-device (one wgpu::Device, one queue [ref crates/standalone/src/
// lib.rs:220-231]); this is the contract's extension of its Renderer: the scene is REPLICATED
// on every GPU (SceneGPU::new_from_scene per device [ref crates/lib/src/scene.rs:151-187]), a
// frame's samples are split -- global rank g of W traces sample indices g, g+W, g+2W, ... of
// the SAME sample sequence one GPU would trace, so the union is the 1-GPU sample set -- and
// the only exchange step of the path is the sum of the FP32 SUM accumulators (W*H*4 floats:
// 33 MB at 1080p, 133 MB at 4K) to rank 0, once per batch, followed on rank 0 by the
// x 1/count -> tone map -> sRGB8 of BlitPass [ref renderer.rs:756-770].
//
// Two shapes of the same object:
//   lp_multi_create(ordinals, n)            ONE process drives n GPUs: ncclCommInitAll, one
//                                           host worker thread per device (launching a wave is
//                                           ~30 kernel launches; eight devices in a row from
//                                           one thread would serialise ~2 ms of launch work)
//   lp_multi_create_rank(ordinal, id, W, g) one process per GPU (torchrun, MPI): the caller
//                                           carries the 128-byte NCCL id from rank 0 to the
//                                           others; ncclCommInitRank
//
// Two implementations of the exchange step (lp_multi_set_reduce_mode):
//   LP_REDUCE_NCCL   ncclReduce(sum, fp32, root 0) in place on a dedicated communication
//                    stream + a second tiny ncclReduce of the ray counters (12 x u64), then
//                    tonemap on rank 0's communication stream.
//   LP_REDUCE_PEER   ONE kernel per GPU over NVLink peer memory: GPU g sums
//                    pixel slice g of every peer's accumulator with plain loads from the
//                    peers' HBM (a reduce-scatter: (W-1)/W of the image crosses each GPU's
//                    NVLink port instead of the whole image converging on rank 0), tone-maps
//                    it and stores both the FP32 sum and the sRGB8 bytes of its slice straight
//                    into rank 0's targets (the gather).  Sum, normalise, tone map and the
//                    transfers are the same instructions of the same kernel.  One process:
//                    the devices map each other (cudaDeviceEnablePeerAccess) and the streams
//                    are ordered by CUDA events.  One process per GPU: the accumulators are
//                    mapped through CUDA IPC handles (exchanged with an ncclAllGather when the
//                    targets are made) and the ranks' streams are ordered by two one-word
//                    ncclAllReduce barriers around the kernel.  No host synchronisation either way.
// Either way a frame in flight may TRACE its next batch while the exchange runs; only its
// accumulate kernel waits (lp_renderer::accum_guard).
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../host/api_common.hpp"
#include "../host/scene.hpp"
#include "api_gpu.cuh"
#include "renderer_state.cuh"
#include "tonemap.cuh"

using namespace lp;

static_assert(LP_MULTI_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "lp_multi id size == ncclUniqueId");
static_assert(sizeof(Counters) == 12 * sizeof(unsigned long long), "12 x u64 ray counters");

#define NCCL_CHECK(expr)                                                                  \
  do {                                                                                    \
    ncclResult_t _n = (expr);                                                             \
    if (_n != ncclSuccess && _n != ncclInProgress)                                        \
      return lp::fail(LP_ERR_NCCL, std::string(#expr) + ": " + ncclGetErrorString(_n));   \
  } while (0)

namespace {

constexpr int kMaxPeers = 16;

// One host thread per device: runs the closures lp_multi posts to it (enqueueing a wave,
// uploading a scene copy) so that the devices' launch work proceeds in parallel.
class Worker {
 public:
  Worker() : th_([this] { loop(); }) {}
  ~Worker() {
    {
      std::lock_guard<std::mutex> g(m_);
      quit_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  void post(std::function<lp_status()> fn) {
    {
      std::lock_guard<std::mutex> g(m_);
      task_ = std::move(fn);
      has_ = true;
      done_ = false;
    }
    cv_.notify_all();
  }
  lp_status wait(std::string &err) {
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return done_; });
    err = error_;
    return status_;
  }

 private:
  void loop() {
    for (;;) {
      std::function<lp_status()> fn;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return has_ || quit_; });
        if (quit_) return;
        fn = std::move(task_);
        has_ = false;
      }
      lp_status st;
      std::string err;
      try {
        st = fn();
        if (st != LP_OK) err = lp_last_error();  // thread-local of THIS thread
      } catch (const std::exception &e) {
        st = LP_ERR_INVALID_ARG;
        err = e.what();
      }
      {
        std::lock_guard<std::mutex> g(m_);
        status_ = st;
        error_ = err;
        done_ = true;
      }
      cv_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::function<lp_status()> task_;
  bool has_ = false, done_ = true, quit_ = false;
  lp_status status_ = LP_OK;
  std::string error_;
  std::thread th_;  // last: starts after the members above exist
};

struct PeerTable {  // kernel argument of the fused exchange
  const float4 *accum[kMaxPeers];
  const Counters *counters[kMaxPeers];
  float4 *root_accum;
  uchar4 *root_ldr;
  Counters *root_counters;
  uint32_t world, rank, n_pixels;
};

// LP_REDUCE_PEER: reduce-scatter + normalise + tone map + gather to rank 0 in one kernel.
// Rank g owns the pixels [g n / W, (g+1) n / W).  For each it loads the W accumulators (its own
// from local HBM, the others over NVLink) in RANK ORDER -- the sum is the same whatever GPU
// computes it -- and stores the sum and its sRGB8 bytes into rank 0's targets.  The slice of
// rank 0's own accumulator is read and written by one thread only (its owner), so the in-place
// result needs no staging copy.  Rank 0 also sums the ray counters.
//
// Rank 0's NVLink port is the hot spot of a reduce TO rank 0: it receives every other slice's
// result (20 B per pixel) whatever the algorithm.  From four GPUs on it therefore owns NO slice
// (its port would also have to carry its slice's inputs): the W-1 others split the image.
__global__ void __launch_bounds__(256) peer_reduce_tonemap_kernel(const PeerTable T) {
  const bool root_works = T.world < 4;
  const uint32_t workers = root_works ? T.world : T.world - 1;
  const uint32_t me = root_works ? T.rank : T.rank - 1;  // rank 0 idles when it owns no slice
  uint32_t lo = 0, hi = 0;
  if (root_works || T.rank > 0) {
    lo = (uint32_t)(((uint64_t)T.n_pixels * me) / workers);
    hi = (uint32_t)(((uint64_t)T.n_pixels * (me + 1)) / workers);
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) {
    float4 v[kMaxPeers];
    // plain loads: L1 is invalidated at every launch boundary and peer lines never enter the
    // local L2, so nothing stale can be read; all W loads are in flight before the first add
#pragma unroll
    for (int k = 0; k < kMaxPeers; ++k)
      if (k < (int)T.world) v[k] = T.accum[k][i];
    float4 s = v[0];
#pragma unroll
    for (int k = 1; k < kMaxPeers; ++k)
      if (k < (int)T.world) {
        s.x += v[k].x;
        s.y += v[k].y;
        s.z += v[k].z;
        s.w += v[k].w;
      }
    T.root_accum[i] = s;
    T.root_ldr[i] = tonemap_srgb8(s);
  }
  if (T.rank == 0 && blockIdx.x == 0 && threadIdx.x < 12) {
    unsigned long long c = 0;
    for (uint32_t k = 0; k < T.world; ++k)
      c += reinterpret_cast<const unsigned long long *>(T.counters[k])[threadIdx.x];
    reinterpret_cast<unsigned long long *>(T.root_counters)[threadIdx.x] = c;
  }
}

__global__ void __launch_bounds__(256) tonemap_reduced_kernel(const float4 *__restrict__ accum,
                                                              uchar4 *__restrict__ out, uint32_t n) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = tonemap_srgb8(accum[i]);
}

struct Lane {  // one local device
  int ordinal = 0, rank = 0;
  lp_device *dev = nullptr;
  lp_scene_gpu *sg = nullptr;
  lp_probe *probe = nullptr;
  lp_renderer *r = nullptr;
  ncclComm_t comm = nullptr;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_rendered = nullptr, ev_reduced = nullptr, ev_all_reduced = nullptr;
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  DevBuf<Counters> counters_red;  // rank 0: the reduced ray counters
  uint32_t samples_this_call = 0;
  std::unique_ptr<Worker> worker;
};

}  // namespace

struct lp_multi {
  int world = 1, first_rank = 0;
  bool single_process = true;
  std::vector<std::unique_ptr<Lane>> lanes;
  lp_render_config cfg{};
  lp_multi_reduce_mode mode = LP_REDUCE_AUTO;
  bool peer_ok = false;       // every GPU of the job can map every other one's targets
  // one process per GPU: the peers' targets mapped through CUDA IPC (index = global rank; the
  // own entries are the local pointers), re-made by lp_multi_resize
  std::vector<void *> ipc_accum, ipc_counters;
  void *ipc_root_ldr = nullptr, *ipc_root_counters_red = nullptr;
  std::vector<void *> ipc_opened;  // what cudaIpcCloseMemHandle has to see again
  DevBuf<int> barrier_word;
  bool timed = false;         // ev_t0/ev_t1 of the last reduce are recorded
  double reduce_ms_total = 0.0;
  uint64_t reduce_count = 0;
  bool has_root() const { return first_rank == 0; }
};

namespace {

// Runs fn(lane) on every lane's worker thread, waits for all, reports the first failure.
lp_status run_all(lp_multi *m, const std::function<lp_status(Lane &)> &fn, size_t from = 0) {
  for (size_t i = from; i < m->lanes.size(); ++i) {
    Lane *l = m->lanes[i].get();
    l->worker->post([l, &fn] {
      cudaError_t e = cudaSetDevice(l->ordinal);
      if (e != cudaSuccess) return fail(LP_ERR_CUDA, cudaGetErrorString(e));
      return fn(*l);
    });
  }
  lp_status first = LP_OK;
  std::string first_err;
  for (size_t i = from; i < m->lanes.size(); ++i) {
    std::string err;
    const lp_status st = m->lanes[i]->worker->wait(err);
    if (st != LP_OK && first == LP_OK) {
      first = st;
      first_err = "GPU " + std::to_string(m->lanes[i]->ordinal) + ": " + err;
    }
  }
  return first == LP_OK ? LP_OK : fail(first, first_err);
}

lp_status check_nccl_async(lp_multi *m) {
  for (auto &l : m->lanes) {
    if (!l->comm) continue;
    ncclResult_t async = ncclSuccess;
    const ncclResult_t q = ncclCommGetAsyncError(l->comm, &async);
    if (q != ncclSuccess || (async != ncclSuccess && async != ncclInProgress)) {
      const ncclResult_t bad = q != ncclSuccess ? q : async;
      ncclCommAbort(l->comm);  // the communicator is unusable from here on
      l->comm = nullptr;
      return fail(LP_ERR_NCCL, std::string("NCCL asynchronous error on rank ") +
                                   std::to_string(l->rank) + ": " + ncclGetErrorString(bad));
    }
  }
  return LP_OK;
}

lp_status lane_init(Lane &l, int ordinal, int rank) {
  l.ordinal = ordinal;
  l.rank = rank;
  lp_status st = lp_device_create(ordinal, &l.dev);
  if (st != LP_OK) return st;
  CUDA_CHECK(cudaSetDevice(ordinal));
  CUDA_CHECK(cudaStreamCreateWithFlags(&l.comm_stream, cudaStreamNonBlocking));
  for (cudaEvent_t *e : {&l.ev_rendered, &l.ev_reduced, &l.ev_all_reduced})
    CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreate(&l.ev_t0));
  CUDA_CHECK(cudaEventCreate(&l.ev_t1));
  CUDA_CHECK(l.counters_red.alloc(1));
  CUDA_CHECK(cudaMemsetAsync(l.counters_red.ptr, 0, sizeof(Counters), l.comm_stream));
  st = lp_renderer_new(l.dev, 2, 2, &l.r);  // sized by lp_multi_resize
  if (st != LP_OK) return st;
  l.worker.reset(new Worker());
  return LP_OK;
}

void lane_destroy(Lane &l) {
  l.worker.reset();
  cudaSetDevice(l.ordinal);
  if (l.comm_stream) cudaStreamSynchronize(l.comm_stream);
  if (l.comm) ncclCommDestroy(l.comm);
  if (l.r) lp_renderer_destroy(l.r);
  if (l.probe) lp_probe_destroy(l.probe);
  if (l.sg) lp_scene_gpu_destroy(l.sg);
  l.counters_red.release();
  for (cudaEvent_t e : {l.ev_rendered, l.ev_reduced, l.ev_all_reduced, l.ev_t0, l.ev_t1})
    if (e) cudaEventDestroy(e);
  if (l.comm_stream) cudaStreamDestroy(l.comm_stream);
  if (l.dev) lp_device_destroy(l.dev);
}

void ipc_close(lp_multi *m);

void multi_free(lp_multi *m) {
  if (!m) return;
  if (!m->lanes.empty() && m->lanes[0]->dev) {
    cudaSetDevice(m->lanes[0]->ordinal);
    if (m->lanes[0]->comm_stream) cudaStreamSynchronize(m->lanes[0]->comm_stream);
    ipc_close(m);
    m->barrier_word.release();
  }
  for (auto &l : m->lanes) lane_destroy(*l);
  delete m;
}

// interleaved share of a call's samples: rank g traces indices g, g+W, ...
uint32_t samples_for_rank(uint32_t total, uint32_t rank, uint32_t world) {
  return total > rank ? (total - rank + world - 1) / world : 0;
}

void ipc_close(lp_multi *m) {
  for (void *p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  m->ipc_opened.clear();
  m->ipc_accum.clear();
  m->ipc_counters.clear();
  m->ipc_root_ldr = m->ipc_root_counters_red = nullptr;
}

// One process per GPU: every rank publishes IPC handles of its SUM accumulator and ray counters
// (rank 0 also of its sRGB8 target and reduced counters), the handles travel with one
// ncclAllGather, every rank maps what it will read or write.  Collective.  peer_ok is the AND
// over all ranks (an ncclAllReduce), so the ranks always agree on the exchange they run.
lp_status ipc_exchange(lp_multi *m) {
  Lane &l = *m->lanes[0];
  CUDA_CHECK(cudaSetDevice(l.ordinal));
  CUDA_CHECK(cudaStreamSynchronize(l.comm_stream));
  ipc_close(m);
  m->peer_ok = false;
  if (m->world < 2) return LP_OK;
  struct Handles {
    cudaIpcMemHandle_t accum, counters, ldr, counters_red;
    int device, ok;
    char pad[8];
  };
  static_assert(sizeof(Handles) % 8 == 0, "gathered as bytes");
  Handles mine;
  std::memset(&mine, 0, sizeof(mine));
  mine.device = l.ordinal;
  mine.ok = cudaIpcGetMemHandle(&mine.accum, l.r->accum.ptr) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.counters, l.r->counters.ptr) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.ldr, l.r->ldr.ptr) == cudaSuccess &&
            cudaIpcGetMemHandle(&mine.counters_red, l.counters_red.ptr) == cudaSuccess;
  cudaGetLastError();
  DevBuf<unsigned char> send, recv;
  CUDA_CHECK(send.alloc(sizeof(Handles)));
  CUDA_CHECK(recv.alloc(sizeof(Handles) * (size_t)m->world));
  CUDA_CHECK(cudaMemcpyAsync(send.ptr, &mine, sizeof(mine), cudaMemcpyHostToDevice, l.comm_stream));
  NCCL_CHECK(ncclAllGather(send.ptr, recv.ptr, sizeof(Handles), ncclUint8, l.comm, l.comm_stream));
  std::vector<Handles> all((size_t)m->world);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), recv.ptr, sizeof(Handles) * (size_t)m->world,
                             cudaMemcpyDeviceToHost, l.comm_stream));
  CUDA_CHECK(cudaStreamSynchronize(l.comm_stream));
  int ok = 1;
  m->ipc_accum.assign((size_t)m->world, nullptr);
  m->ipc_counters.assign((size_t)m->world, nullptr);
  auto open = [&](const cudaIpcMemHandle_t &h) -> void * {
    void *p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
      return nullptr;
    }
    m->ipc_opened.push_back(p);
    return p;
  };
  for (int g = 0; g < m->world && ok; ++g) {
    if (!all[(size_t)g].ok) ok = 0;
    if (g == l.rank) {
      m->ipc_accum[(size_t)g] = l.r->accum.ptr;
      m->ipc_counters[(size_t)g] = l.r->counters.ptr;
      continue;
    }
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, l.ordinal, all[(size_t)g].device) != cudaSuccess || !can) ok = 0;
    if (!ok) break;
    m->ipc_accum[(size_t)g] = open(all[(size_t)g].accum);
    m->ipc_counters[(size_t)g] = open(all[(size_t)g].counters);
  }
  if (ok) {
    if (l.rank == 0) {
      m->ipc_root_ldr = l.r->ldr.ptr;
      m->ipc_root_counters_red = l.counters_red.ptr;
    } else {
      m->ipc_root_ldr = open(all[0].ldr);
      m->ipc_root_counters_red = open(all[0].counters_red);
    }
  }
  // every rank must have mapped everything, or all of them fall back to NCCL
  CUDA_CHECK(m->barrier_word.alloc(1));
  CUDA_CHECK(cudaMemcpyAsync(m->barrier_word.ptr, &ok, sizeof(int), cudaMemcpyHostToDevice,
                             l.comm_stream));
  NCCL_CHECK(ncclAllReduce(m->barrier_word.ptr, m->barrier_word.ptr, 1, ncclInt32, ncclMin, l.comm,
                           l.comm_stream));
  int all_ok = 0;
  CUDA_CHECK(cudaMemcpyAsync(&all_ok, m->barrier_word.ptr, sizeof(int), cudaMemcpyDeviceToHost,
                             l.comm_stream));
  CUDA_CHECK(cudaStreamSynchronize(l.comm_stream));
  if (!all_ok) ipc_close(m);
  m->peer_ok = all_ok != 0;
  return LP_OK;
}

lp_status apply_config(lp_multi *m, Lane &l) {
  lp_render_config c = m->cfg;
  const uint32_t s0 = std::max(1u, m->cfg.sample_stride);
  l.samples_this_call = samples_for_rank(m->cfg.spp_per_call, (uint32_t)l.rank, (uint32_t)m->world);
  c.spp_per_call = std::max(1u, l.samples_this_call);
  c.sample_offset = m->cfg.sample_offset + (uint32_t)l.rank * s0;
  c.sample_stride = (uint32_t)m->world * s0;
  return lp_renderer_set_config(l.r, &c);
}

}  // namespace

extern "C" {

LP_API lp_status lp_multi_unique_id(uint8_t id[LP_MULTI_ID_BYTES]) try {
  if (!id) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  ncclUniqueId u;
  NCCL_CHECK(ncclGetUniqueId(&u));
  std::memcpy(id, &u, LP_MULTI_ID_BYTES);
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_create(const int *cuda_ordinals, int n_devices, lp_multi **out) try {
  if (!out || n_devices < 1 || n_devices > kMaxPeers)
    return fail(LP_ERR_INVALID_ARG, "lp_multi_create: 1..16 devices");
  int available = 0;
  cudaError_t e = cudaGetDeviceCount(&available);
  if (e != cudaSuccess || available == 0)
    return fail(LP_ERR_CUDA, std::string("no CUDA device available (this library has no CPU "
                                         "fallback): ") + cudaGetErrorString(e));
  std::vector<int> ords(n_devices);
  for (int i = 0; i < n_devices; ++i) {
    ords[i] = cuda_ordinals ? cuda_ordinals[i] : i;
    if (ords[i] < 0 || ords[i] >= available) return fail(LP_ERR_INVALID_ARG, "bad CUDA ordinal");
    for (int k = 0; k < i; ++k)
      if (ords[k] == ords[i]) return fail(LP_ERR_INVALID_ARG, "duplicate CUDA ordinal");
  }
  lp_multi *m = new (std::nothrow) lp_multi();
  if (!m) return fail(LP_ERR_OOM, "out of host memory");
  m->world = n_devices;
  m->first_rank = 0;
  m->single_process = true;
  lp_render_config_default(&m->cfg);
  for (int i = 0; i < n_devices; ++i) {
    m->lanes.emplace_back(new Lane());
    const lp_status st = lane_init(*m->lanes.back(), ords[i], i);
    if (st != LP_OK) {
      const std::string msg = lp_last_error();
      multi_free(m);
      return fail(st, msg);
    }
  }
  if (n_devices > 1) {
    std::vector<ncclComm_t> comms(n_devices);
    const ncclResult_t nr = ncclCommInitAll(comms.data(), n_devices, ords.data());
    if (nr != ncclSuccess) {
      multi_free(m);
      return fail(LP_ERR_NCCL, std::string("ncclCommInitAll: ") + ncclGetErrorString(nr));
    }
    for (int i = 0; i < n_devices; ++i) m->lanes[i]->comm = comms[i];
    // NVLink peer mappings for LP_REDUCE_PEER: every device maps every other one
    bool ok = true;
    for (int i = 0; i < n_devices && ok; ++i)
      for (int k = 0; k < n_devices && ok; ++k) {
        if (i == k) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ords[i], ords[k]) != cudaSuccess || !can) ok = false;
      }
    for (int i = 0; i < n_devices && ok; ++i) {
      cudaSetDevice(ords[i]);
      for (int k = 0; k < n_devices && ok; ++k) {
        if (i == k) continue;
        const cudaError_t pe = cudaDeviceEnablePeerAccess(ords[k], 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (pe != cudaSuccess) {
          cudaGetLastError();
          ok = false;
        }
      }
    }
    m->peer_ok = ok;
  }
  *out = m;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_create_rank(int cuda_ordinal, const uint8_t id[LP_MULTI_ID_BYTES],
                                      int n_ranks, int rank, lp_multi **out) try {
  if (!out || !id || n_ranks < 1 || rank < 0 || rank >= n_ranks)
    return fail(LP_ERR_INVALID_ARG, "lp_multi_create_rank: bad rank / world size");
  lp_multi *m = new (std::nothrow) lp_multi();
  if (!m) return fail(LP_ERR_OOM, "out of host memory");
  m->world = n_ranks;
  m->first_rank = rank;
  m->single_process = false;
  lp_render_config_default(&m->cfg);
  m->lanes.emplace_back(new Lane());
  Lane &l = *m->lanes.back();
  lp_status st = lane_init(l, cuda_ordinal, rank);
  if (st != LP_OK) {
    const std::string msg = lp_last_error();
    multi_free(m);
    return fail(st, msg);
  }
  if (n_ranks > 1) {
    ncclUniqueId u;
    std::memcpy(&u, id, LP_MULTI_ID_BYTES);
    const ncclResult_t nr = ncclCommInitRank(&l.comm, n_ranks, u, rank);  // collective
    if (nr != ncclSuccess) {
      multi_free(m);
      return fail(LP_ERR_NCCL, std::string("ncclCommInitRank: ") + ncclGetErrorString(nr));
    }
  }
  *out = m;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_destroy(lp_multi *m) try {
  multi_free(m);
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_info(const lp_multi *m, int *world, int *first_rank, int *local_devices,
                               int *peer_access) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (world) *world = m->world;
  if (first_rank) *first_rank = m->first_rank;
  if (local_devices) *local_devices = (int)m->lanes.size();
  if (peer_access) *peer_access = m->peer_ok ? 1 : 0;  // the fused exchange is available
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_device(lp_multi *m, int local_index, lp_device **out) try {
  if (!m || !out || local_index < 0 || local_index >= (int)m->lanes.size())
    return fail(LP_ERR_INVALID_ARG, "bad argument");
  *out = m->lanes[local_index]->dev;
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_renderer(lp_multi *m, int local_index, lp_renderer **out) try {
  if (!m || !out || local_index < 0 || local_index >= (int)m->lanes.size())
    return fail(LP_ERR_INVALID_ARG, "bad argument");
  *out = m->lanes[local_index]->r;
  return LP_OK;
} LP_ABI_CATCH

// Replicated upload: one SceneGPU per local device.  The first copy runs alone (it builds the
// scene's derived host arrays -- TLAS, GPU layout, atlas -- which the others then only read),
// the rest in parallel on the devices' worker threads.
LP_API lp_status lp_multi_set_scene(lp_multi *m, lp_scene *scene, int device_build) try {
  if (!m || !scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  auto upload = [&](Lane &l) -> lp_status {
    lp_scene_gpu *g = nullptr;
    const lp_status st = device_build ? lp_scene_gpu_new_from_scene_lbvh(scene, l.dev, &g)
                                      : lp_scene_gpu_new_from_scene(scene, l.dev, &g);
    if (st != LP_OK) return st;
    CUDA_CHECK(cudaStreamSynchronize(l.dev->stream));
    CUDA_CHECK(cudaStreamSynchronize(l.dev->stream2));
    if (l.sg) lp_scene_gpu_destroy(l.sg);
    l.sg = g;
    return lp_renderer_set_resources(l.r, l.sg, l.probe);
  };
  if (device_build) {  // the device build re-packs the scene's atlas per call: one at a time
    for (auto &l : m->lanes) {
      CUDA_CHECK(cudaSetDevice(l->ordinal));
      const lp_status st = upload(*l);
      if (st != LP_OK) return st;
    }
    return LP_OK;
  }
  CUDA_CHECK(cudaSetDevice(m->lanes[0]->ordinal));
  const lp_status st0 = upload(*m->lanes[0]);
  if (st0 != LP_OK) return st0;
  return m->lanes.size() > 1 ? run_all(m, upload, 1) : LP_OK;
} LP_ABI_CATCH

// Instance::set_transform / small-table edits carried to every copy
// (lp_scene_gpu_update_instances per device; the host TLAS rebuild happens once).
LP_API lp_status lp_multi_update_instances(lp_multi *m, lp_scene *scene) try {
  if (!m || !scene) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  for (auto &l : m->lanes) {
    if (!l->sg) return fail(LP_ERR_INVALID_ARG, "lp_multi_set_scene has not been called");
    CUDA_CHECK(cudaSetDevice(l->ordinal));
    const lp_status st = lp_scene_gpu_update_instances(l->sg, scene);
    if (st != LP_OK) return st;
  }
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_set_probe(lp_multi *m, const uint8_t *rgbe8, uint32_t width,
                                    uint32_t height) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  return run_all(m, [&](Lane &l) -> lp_status {
    lp_probe *p = nullptr;
    if (rgbe8) {
      const lp_status st = lp_probe_new(l.dev, rgbe8, width, height, &p);
      if (st != LP_OK) return st;
    }
    CUDA_CHECK(cudaStreamSynchronize(l.dev->stream));
    CUDA_CHECK(cudaStreamSynchronize(l.dev->stream2));
    if (l.probe) lp_probe_destroy(l.probe);
    l.probe = p;
    return lp_renderer_set_resources(l.r, l.sg, l.probe);
  });
} LP_ABI_CATCH

LP_API lp_status lp_multi_resize(lp_multi *m, uint32_t width, uint32_t height,
                                 float downsample_factor) try {
  if (!m || !width || !height) return fail(LP_ERR_INVALID_ARG, "bad argument");
  if (!m->single_process && m->world > 1 && !m->ipc_opened.empty()) {
    // every rank drops its mappings of the peers' targets BEFORE any rank frees them
    Lane &l0 = *m->lanes[0];
    CUDA_CHECK(cudaSetDevice(l0.ordinal));
    CUDA_CHECK(cudaStreamSynchronize(l0.comm_stream));
    ipc_close(m);
    m->peer_ok = false;
    NCCL_CHECK(ncclAllReduce(m->barrier_word.ptr, m->barrier_word.ptr, 1, ncclInt32, ncclMin,
                             l0.comm, l0.comm_stream));
    CUDA_CHECK(cudaStreamSynchronize(l0.comm_stream));
  }
  const lp_status st = run_all(m, [&](Lane &l) -> lp_status {
    CUDA_CHECK(cudaStreamSynchronize(l.comm_stream));  // peers may still read the old targets
    const lp_status ds = lp_renderer_set_downsample_factor(l.r, downsample_factor);
    if (ds != LP_OK) return ds;
    l.r->accum_guard = nullptr;
    return lp_renderer_resize(l.r, l.sg, l.probe, width, height);
  });
  if (st != LP_OK) return st;
  // one process per GPU: the targets were re-made, so the peers' mappings are too (collective)
  if (!m->single_process && m->world > 1) return ipc_exchange(m);
  return LP_OK;
} LP_ABI_CATCH

// cfg.spp_per_call is the TOTAL number of samples per pixel one lp_multi_render traces over all
// ranks; cfg.sample_offset / sample_stride describe the sequence ONE GPU would trace.  Rank g
// receives offset + g * stride, stride * W and its share of the count.
LP_API lp_status lp_multi_set_config(lp_multi *m, const lp_render_config *cfg) try {
  if (!m || !cfg) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (cfg->spp_per_call < 1) return fail(LP_ERR_INVALID_ARG, "spp_per_call must be >= 1");
  m->cfg = *cfg;
  return run_all(m, [&](Lane &l) { return apply_config(m, l); });
} LP_ABI_CATCH

LP_API lp_status lp_multi_set_accumulate(lp_multi *m, int flag) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  for (auto &l : m->lanes) lp_renderer_set_accumulate(l->r, flag);
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_set_reduce_mode(lp_multi *m, lp_multi_reduce_mode mode) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (mode == LP_REDUCE_PEER && !(m->peer_ok || (m->single_process && m->world == 1)))
    return fail(LP_ERR_INVALID_ARG,
                "LP_REDUCE_PEER needs NVLink peer access between every pair of GPUs (and, with "
                "one process per GPU, lp_multi_resize to have mapped the peers' targets)");
  if (mode != LP_REDUCE_AUTO && mode != LP_REDUCE_NCCL && mode != LP_REDUCE_PEER)
    return fail(LP_ERR_INVALID_ARG, "unknown reduce mode");
  m->mode = mode;
  return LP_OK;
} LP_ABI_CATCH

// Renderer::raytrace on every local device, each tracing its share of cfg.spp_per_call
// (asynchronous: returns when the work is enqueued).  A rank whose share of this call is empty
// traces nothing and contributes zeros.
LP_API lp_status lp_multi_render(lp_multi *m, const float view_transform[16]) try {
  if (!m || !view_transform) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  auto trace = [&](Lane &l) -> lp_status {
    if (l.samples_this_call == 0) {
      if (!l.r->accumulate || l.r->samples_accumulated == 0) {
        if (l.r->accum_guard) {
          CUDA_CHECK(cudaStreamWaitEvent(l.dev->stream, l.r->accum_guard, 0));
          l.r->accum_guard = nullptr;
        }
        CUDA_CHECK(cudaMemsetAsync(l.r->accum.ptr, 0, l.r->accum.count * sizeof(float4),
                                   l.dev->stream));
      }
      return LP_OK;
    }
    return lp_renderer_raytrace(l.r, view_transform);
  };
  if (m->lanes.size() == 1) {
    CUDA_CHECK(cudaSetDevice(m->lanes[0]->ordinal));
    return trace(*m->lanes[0]);
  }
  return run_all(m, trace);
} LP_ABI_CATCH

// The exchange step.  Asynchronous: everything is enqueued on the communication streams,
// ordered after the tracing streams by events; lp_multi_read_* / lp_multi_synchronize wait.
LP_API lp_status lp_multi_reduce(lp_multi *m) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  // AUTO = NCCL: measured on 2 and 8 B200s the fused kernel ties (0.108 vs 0.104 ms at 1080p,
  // 2 GPUs) or loses (0.70 vs 0.29 ms at 4K, 8 GPUs) because a reduce TO rank 0 is bound by rank
  // 0's NVLink port either way and ncclReduce already runs at that bound; PEER is explicit
  const bool peer = m->world > 1 && m->peer_ok && m->mode == LP_REDUCE_PEER;
  if (m->mode == LP_REDUCE_PEER && m->world > 1 && !m->peer_ok)
    return fail(LP_ERR_INVALID_ARG, "LP_REDUCE_PEER: the peers' targets are not mapped");
  const uint32_t n_pixels = m->lanes[0]->r->width * m->lanes[0]->r->height;
  for (auto &l : m->lanes) {
    if (l->r->width * l->r->height != n_pixels)
      return fail(LP_ERR_INVALID_ARG, "renderers differ in size: call lp_multi_resize");
    CUDA_CHECK(cudaSetDevice(l->ordinal));
    CUDA_CHECK(cudaEventRecord(l->ev_rendered, l->dev->stream));
  }
  Lane *root = m->has_root() ? m->lanes[0].get() : nullptr;
  if (peer && !m->single_process) {
    // one process per GPU: same kernel over the IPC mappings; the ranks' communication streams
    // meet in a one-word all-reduce before it (every accumulator is complete) and after it
    // (every slice has landed on rank 0, every accumulator may be overwritten again)
    Lane &l = *m->lanes[0];
    CUDA_CHECK(cudaSetDevice(l.ordinal));
    CUDA_CHECK(cudaStreamWaitEvent(l.comm_stream, l.ev_rendered, 0));
    if (root) CUDA_CHECK(cudaEventRecord(l.ev_t0, l.comm_stream));
    NCCL_CHECK(ncclAllReduce(m->barrier_word.ptr, m->barrier_word.ptr, 1, ncclInt32, ncclMin, l.comm,
                             l.comm_stream));
    PeerTable T{};
    T.world = (uint32_t)m->world;
    T.rank = (uint32_t)l.rank;
    T.n_pixels = n_pixels;
    for (int k = 0; k < m->world; ++k) {
      T.accum[k] = static_cast<const float4 *>(m->ipc_accum[(size_t)k]);
      T.counters[k] = static_cast<const Counters *>(m->ipc_counters[(size_t)k]);
    }
    T.root_accum = static_cast<float4 *>(m->ipc_accum[0]);
    T.root_ldr = static_cast<uchar4 *>(m->ipc_root_ldr);
    T.root_counters = static_cast<Counters *>(m->ipc_root_counters_red);
    const uint32_t slice = n_pixels / (uint32_t)std::max(1, m->world - 1) + 1;
    const int blocks = (int)std::min<uint32_t>((slice + 255) / 256, (uint32_t)l.dev->sm_count * 8);
    peer_reduce_tonemap_kernel<<<blocks, 256, 0, l.comm_stream>>>(T);
    NCCL_CHECK(ncclAllReduce(m->barrier_word.ptr, m->barrier_word.ptr, 1, ncclInt32, ncclMin, l.comm,
                             l.comm_stream));
    if (root) CUDA_CHECK(cudaEventRecord(l.ev_t1, l.comm_stream));
    CUDA_CHECK(cudaEventRecord(l.ev_all_reduced, l.comm_stream));
    l.r->accum_guard = l.ev_all_reduced;
  } else if (peer && m->world > 1) {
    PeerTable T{};
    T.world = (uint32_t)m->world;
    T.n_pixels = n_pixels;
    for (int k = 0; k < m->world; ++k) {
      T.accum[k] = m->lanes[k]->r->accum.ptr;
      T.counters[k] = m->lanes[k]->r->counters.ptr;
    }
    T.root_accum = root->r->accum.ptr;
    T.root_ldr = root->r->ldr.ptr;
    T.root_counters = root->counters_red.ptr;
    for (auto &l : m->lanes) {
      CUDA_CHECK(cudaSetDevice(l->ordinal));
      for (auto &o : m->lanes)  // every peer's batch is complete before anyone reads it
        CUDA_CHECK(cudaStreamWaitEvent(l->comm_stream, o->ev_rendered, 0));
      if (l.get() == root) CUDA_CHECK(cudaEventRecord(l->ev_t0, l->comm_stream));
      T.rank = (uint32_t)l->rank;
      const uint32_t slice = n_pixels / (uint32_t)std::max(1, m->world - 1) + 1;
      const int blocks = (int)std::min<uint32_t>((slice + 255) / 256, (uint32_t)l->dev->sm_count * 8);
      peer_reduce_tonemap_kernel<<<blocks, 256, 0, l->comm_stream>>>(T);
      CUDA_CHECK(cudaEventRecord(l->ev_reduced, l->comm_stream));
    }
    for (auto &l : m->lanes) {  // a lane's accumulator is free again when EVERY peer has read it
      CUDA_CHECK(cudaSetDevice(l->ordinal));
      for (auto &o : m->lanes)
        if (o.get() != l.get()) CUDA_CHECK(cudaStreamWaitEvent(l->comm_stream, o->ev_reduced, 0));
      if (l.get() == root) CUDA_CHECK(cudaEventRecord(l->ev_t1, l->comm_stream));
      CUDA_CHECK(cudaEventRecord(l->ev_all_reduced, l->comm_stream));
      l->r->accum_guard = l->ev_all_reduced;
    }
  } else {
    for (auto &l : m->lanes) {
      CUDA_CHECK(cudaSetDevice(l->ordinal));
      CUDA_CHECK(cudaStreamWaitEvent(l->comm_stream, l->ev_rendered, 0));
      if (l.get() == root) CUDA_CHECK(cudaEventRecord(l->ev_t0, l->comm_stream));
    }
    if (m->world > 1) {
      NCCL_CHECK(ncclGroupStart());
      for (auto &l : m->lanes) {
        NCCL_CHECK(ncclReduce(l->r->accum.ptr, l->r->accum.ptr, (size_t)n_pixels * 4, ncclFloat32,
                              ncclSum, 0, l->comm, l->comm_stream));
        NCCL_CHECK(ncclReduce(l->r->counters.ptr, l->counters_red.ptr, 12, ncclUint64, ncclSum, 0,
                              l->comm, l->comm_stream));
      }
      NCCL_CHECK(ncclGroupEnd());
    } else {
      CUDA_CHECK(cudaMemcpyAsync(root->counters_red.ptr, root->r->counters.ptr, sizeof(Counters),
                                 cudaMemcpyDeviceToDevice, root->comm_stream));
    }
    if (root) {  // x 1/count -> tone map -> sRGB8, right behind the reduce on the same stream
      CUDA_CHECK(cudaSetDevice(root->ordinal));
      tonemap_reduced_kernel<<<root->dev->sm_count * 8, 256, 0, root->comm_stream>>>(
          root->r->accum.ptr, root->r->ldr.ptr, n_pixels);
      CUDA_CHECK(cudaEventRecord(root->ev_t1, root->comm_stream));
    }
    for (auto &l : m->lanes) {
      CUDA_CHECK(cudaSetDevice(l->ordinal));
      CUDA_CHECK(cudaEventRecord(l->ev_all_reduced, l->comm_stream));
      l->r->accum_guard = l->ev_all_reduced;
    }
  }
  CUDA_CHECK(cudaGetLastError());
  m->timed = root != nullptr;
  return check_nccl_async(m);
} LP_ABI_CATCH

// Makes every local device's TRACING stream (lp_device_stream) wait for the exchange step
// enqueued so far, without blocking the host: an event a caller records on that stream
// afterwards covers the reduce too (device-side timing of render + reduce).
LP_API lp_status lp_multi_join(lp_multi *m) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  for (auto &l : m->lanes) {
    CUDA_CHECK(cudaSetDevice(l->ordinal));
    CUDA_CHECK(cudaStreamWaitEvent(l->dev->stream, l->ev_all_reduced, 0));
  }
  return LP_OK;
} LP_ABI_CATCH

LP_API lp_status lp_multi_synchronize(lp_multi *m) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  for (auto &l : m->lanes) {
    CUDA_CHECK(cudaSetDevice(l->ordinal));
    CUDA_CHECK(cudaStreamSynchronize(l->dev->stream));
    CUDA_CHECK(cudaStreamSynchronize(l->dev->stream2));
    CUDA_CHECK(cudaStreamSynchronize(l->comm_stream));
  }
  if (m->timed && m->has_root()) {
    float ms = 0.f;
    Lane *root = m->lanes[0].get();
    if (cudaEventElapsedTime(&ms, root->ev_t0, root->ev_t1) == cudaSuccess) {
      m->reduce_ms_total += ms;
      m->reduce_count++;
    } else {
      cudaGetLastError();
    }
    m->timed = false;
  }
  return check_nccl_async(m);
} LP_ABI_CATCH

// Device time of the exchange step on rank 0's communication stream (from "every input is
// ready" to "the sRGB8 frame is complete"), summed over the lp_multi_reduce calls that were
// followed by a synchronising call before the next reduce.
LP_API lp_status lp_multi_reduce_time(lp_multi *m, double *total_ms, uint64_t *count, int reset) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const lp_status st = lp_multi_synchronize(m);
  if (st != LP_OK) return st;
  if (total_ms) *total_ms = m->reduce_ms_total;
  if (count) *count = m->reduce_count;
  if (reset) {
    m->reduce_ms_total = 0.0;
    m->reduce_count = 0;
  }
  return LP_OK;
} LP_ABI_CATCH

// Renderer::read_pixels of the reduced frame (rank 0 only): the bytes the exchange step
// already produced.
LP_API lp_status lp_multi_read_pixels(lp_multi *m, uint8_t *out, size_t cap) try {
  if (!m || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (!m->has_root()) return fail(LP_ERR_INVALID_ARG, "only rank 0 holds the reduced frame");
  Lane *root = m->lanes[0].get();
  const size_t n = (size_t)root->r->width * root->r->height;
  if (cap < n * 4) return fail(LP_ERR_READBACK, "output buffer too small");
  if (cudaSetDevice(root->ordinal) != cudaSuccess) return fail(LP_ERR_READBACK, "cudaSetDevice");
  cudaError_t e = cudaMemcpyAsync(out, root->r->ldr.ptr, n * 4, cudaMemcpyDeviceToHost,
                                  root->comm_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(root->comm_stream);
  if (e != cudaSuccess) return fail(LP_ERR_READBACK, cudaGetErrorString(e));
  return lp_multi_synchronize(m);  // also polls the communicators (LP_ERR_NCCL)
} LP_ABI_CATCH

// The reduced FP32 SUM target (alpha = total sample count) of rank 0.
LP_API lp_status lp_multi_read_accum_sum(lp_multi *m, float *out, size_t cap_floats) try {
  if (!m || !out) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  if (!m->has_root()) return fail(LP_ERR_INVALID_ARG, "only rank 0 holds the reduced frame");
  Lane *root = m->lanes[0].get();
  const size_t n = (size_t)root->r->width * root->r->height;
  if (cap_floats < n * 4) return fail(LP_ERR_READBACK, "output buffer too small");
  CUDA_CHECK(cudaSetDevice(root->ordinal));
  CUDA_CHECK(cudaMemcpyAsync(out, root->r->accum.ptr, n * sizeof(float4), cudaMemcpyDeviceToHost,
                             root->comm_stream));
  CUDA_CHECK(cudaStreamSynchronize(root->comm_stream));
  return lp_multi_synchronize(m);
} LP_ABI_CATCH

// Ray counters summed over all ranks by the LAST lp_multi_reduce (rank 0 only); reset != 0 also
// clears every local renderer's counters.
LP_API lp_status lp_multi_ray_counters(lp_multi *m, lp_ray_counters *out, int reset) try {
  if (!m) return fail(LP_ERR_INVALID_ARG, "NULL argument");
  const lp_status st = lp_multi_synchronize(m);
  if (st != LP_OK) return st;
  if (out) {
    if (!m->has_root()) return fail(LP_ERR_INVALID_ARG, "only rank 0 holds the reduced counters");
    Lane *root = m->lanes[0].get();
    Counters c;
    CUDA_CHECK(cudaSetDevice(root->ordinal));
    CUDA_CHECK(cudaMemcpy(&c, root->counters_red.ptr, sizeof(c), cudaMemcpyDeviceToHost));
    out->primary = c.rays[0];
    out->bounce = c.rays[1];
    out->shadow = c.rays[2];
    for (int k = 0; k < 3; ++k) {
      out->n_int[k] = c.n_int[k];
      out->n_tri[k] = c.n_tri[k];
      out->n_inst[k] = c.n_inst[k];
    }
  }
  if (reset)
    for (auto &l : m->lanes) {
      const lp_status rs = lp_renderer_ray_counters(l->r, nullptr, 1);
      if (rs != LP_OK) return rs;
    }
  return LP_OK;
} LP_ABI_CATCH

}  // extern "C"
