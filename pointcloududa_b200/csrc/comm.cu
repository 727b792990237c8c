// Multi-GPU exchange of the adversarial step (SURVEY.md §8b "pcuda_comm_init/allreduce/destroy", §8e).
//
// The reference is single-process; the only data-path exchange of the batch-sharded step is the sum of D4's
// parameter gradients (one flat fp32 bucket, 6.4 MB for the default network) in front of its SGD step
// (train_mscmrseg.py:329-330).  Two implementations behind one communicator object:
//
//   pcuda_comm_allreduce       ncclAllReduce on the caller's stream.  libnccl is bound with dlopen at the first
//                              pcuda_comm_* call (the copy torch has already loaded, else libnccl.so.2), so libpcuda.so
//                              itself has no link-time dependency on it.
//   pcuda_comm_allreduce_p2p   a two-shot all-reduce written for NVLink 5 / NVSwitch peer memory: every rank owns a
//                              "symmetric" region (cudaMalloc + cudaIpc handles exchanged once through NCCL), the
//                              bucket is PACKED STRAIGHT INTO it by pcuda_grad_sum_pack, and ONE kernel does
//                                 start barrier -> each rank sums its 1/R slice from all R regions in rank order
//                                 (P2P loads) -> stores the sum into every region's output buffer (P2P stores) ->
//                                 end barrier,
//                              block b of every rank pairing with block b of the peers through release/acquire
//                              flags in peer memory.  At 6.4 MB x 8 ranks that is 5.6 MB read + 5.6 MB written over
//                              NVLink per GPU and two flag round trips, against NCCL's ring / tree latency; every
//                              rank ends with bit-identical sums (one owner per element, fixed rank order).
//
// Both are plain launches on the caller's stream: legal under CUDA-graph capture, so the whole step — three D4
// passes, the exchange, SGD — replays as ONE graph per rank.  A peer that never arrives makes the spin loops give
// up after 30 s (pcuda_comm_status reports it) instead of hanging the GPU.
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include <mutex>

#include "pcuda_common.cuh"

namespace pcuda {
namespace {

// ---- NCCL, bound at run time ----------------------------------------------------------------------------
struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};
static Nccl g_nccl;
static std::mutex g_nccl_mu;

static const Nccl* nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.ok) return &g_nccl;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the copy already in the process (torch's)
    if (h) break;
  }
  for (int i = 0; !h && i < 2; ++i) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) { set_error("pcuda_comm: libnccl.so.2 not found (%s)", dlerror()); return nullptr; }
  g_nccl.handle = h;
#define PCUDA_NCCL_SYM(field, sym)                                                      \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, sym);                             \
  if (!g_nccl.field) { set_error("pcuda_comm: %s missing from libnccl", sym); return nullptr; }
  PCUDA_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  PCUDA_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  PCUDA_NCCL_SYM(AllReduce, "ncclAllReduce")
  PCUDA_NCCL_SYM(AllGather, "ncclAllGather")
  PCUDA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  PCUDA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  PCUDA_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef PCUDA_NCCL_SYM
  g_nccl.ok = true;
  return &g_nccl;
}

// ---- symmetric region -------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 128;
constexpr int kP2PThreads = 512;
constexpr unsigned long long kSpinLimitNs = 30000000000ull;  // 30 s: a peer that never arrives.  Generous on purpose: on the
                                                             // first step one rank may still be loading its CUDA modules
                                                             // (lazy loading, ~200 kernels) seconds after the other

// [ epoch[kMaxBlocks] | start[kMaxRanks][kMaxBlocks] | end[kMaxRanks][kMaxBlocks] | status[4] | pad ] [ in ] [ out ]
constexpr size_t kEpochOff = 0;
constexpr size_t kStartOff = kEpochOff + sizeof(unsigned) * kMaxBlocks;
constexpr size_t kEndOff = kStartOff + sizeof(unsigned) * kMaxRanks * kMaxBlocks;
constexpr size_t kStatusOff = kEndOff + sizeof(unsigned) * kMaxRanks * kMaxBlocks;
constexpr size_t kHeaderBytes = (kStatusOff + 64 + 1023) / 1024 * 1024;

// Small-message channels (cross-rank BatchNorm sums: <= 2048 doubles, dozens per step, latency is everything): per channel
// [ epoch | flags[kMaxRanks] | pad to 1 KB ] [ mailbox[2][kMaxRanks][kSmallMax] doubles ], placed behind the two data buffers.
// One-shot: every rank stores its values into slot [parity][rank] of every peer's mailbox, one flag round trip, every
// rank adds the R slots of its own mailbox in rank order (bit-identical everywhere).  The two parities alternate with
// the epoch, so a rank that is already one call ahead never overwrites what a peer still reads.  Concurrent streams
// (the three D4 passes of a step) use different channels.
constexpr int kSmallMax = 2048;
constexpr int kChannels = 8;
constexpr size_t kChanHeader = 1024;
constexpr size_t kChanBytes = kChanHeader + sizeof(double) * 2 * kMaxRanks * kSmallMax;

struct SmallArgs {
  char* peer[kMaxRanks];
  int rank, world, count;
  long long chan_off;       // byte offset of the channel inside every rank's region
  double* buf;
};

struct P2PArgs {
  char* peer[kMaxRanks];
  int rank, world;
  long long slice4;         // float4 elements per rank slice
  long long cap_bytes;      // bytes of one data buffer (in / out)
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// thread r < world: tell peer r that (this rank, block b) reached `epoch`, then wait for peer r's same message
__device__ __forceinline__ void pair_barrier(const P2PArgs& a, size_t flag_off, int b, unsigned epoch) {
  if (static_cast<int>(threadIdx.x) < a.world) {
    const int r = threadIdx.x;
    unsigned* theirs = reinterpret_cast<unsigned*>(a.peer[r] + flag_off) + a.rank * kMaxBlocks + b;
    const unsigned* mine = reinterpret_cast<const unsigned*>(a.peer[a.rank] + flag_off) + r * kMaxBlocks + b;
    __threadfence_system();
    st_release_sys(theirs, epoch);
    const unsigned long long t0 = global_ns();
    while (static_cast<int>(ld_acquire_sys(mine) - epoch) < 0) {
      if (global_ns() - t0 > kSpinLimitNs) {
        reinterpret_cast<unsigned*>(a.peer[a.rank] + kStatusOff)[0] = 1u;      // reported by pcuda_comm_status
        break;
      }
    }
  }
  __syncthreads();
}

// Neither kernel below lets its successor start early (no griddepcontrol.launch_dependents: pdl_wait() only).  A
// successor that is resident while this kernel spins on a PEER holds shared memory / thread slots on its SMs; with three
// concurrent branches per rank those waiting CTAs can cover every SM, the large-shared-memory GEMM another branch needs
// before ITS exchange cannot be placed, and the two ranks wait for each other until the spin limit (seen with cross-rank
// BatchNorm at cfg-3 on 2 GPUs).  Kernels waiting on LOCAL predecessors cannot form such a cycle.
__global__ void __launch_bounds__(kP2PThreads) p2p_allreduce_kernel(const __grid_constant__ P2PArgs a) {
  pdl_wait();
  const int b = blockIdx.x, nb = gridDim.x;
  __shared__ unsigned s_epoch;
  unsigned* epochs = reinterpret_cast<unsigned*>(a.peer[a.rank] + kEpochOff);
  if (threadIdx.x == 0) s_epoch = epochs[b] + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  // every peer's bucket is complete (its pack kernel precedes this kernel on its stream) once its block b is here
  pair_barrier(a, kStartOff, b, epoch);
  const long long per = (a.slice4 + nb - 1) / nb;
  const long long i0 = a.rank * a.slice4 + static_cast<long long>(b) * per;
  const long long i1 = min(i0 + per, (a.rank + 1) * a.slice4);
  for (long long i = i0 + threadIdx.x; i < i1; i += kP2PThreads) {
    float4 acc = reinterpret_cast<const float4*>(a.peer[0] + kHeaderBytes)[i];
#pragma unroll
    for (int r = 1; r < kMaxRanks; ++r) {
      if (r < a.world) {
        const float4 v = reinterpret_cast<const float4*>(a.peer[r] + kHeaderBytes)[i];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < a.world) reinterpret_cast<float4*>(a.peer[r] + kHeaderBytes + a.cap_bytes)[i] = acc;
  }
  __syncthreads();
  // part b of every rank's slice has landed in this rank's output buffer once every peer's block b is here
  pair_barrier(a, kEndOff, b, epoch);
  if (threadIdx.x == 0) epochs[b] = epoch;
}

__global__ void __launch_bounds__(1024) p2p_sum_f64_kernel(const __grid_constant__ SmallArgs a) {
  pdl_wait();
  char* mine = a.peer[a.rank] + a.chan_off;
  unsigned* ep = reinterpret_cast<unsigned*>(mine);
  const unsigned* flags_mine = ep + 8;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = ep[0] + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  const size_t slot = static_cast<size_t>(epoch & 1u) * kMaxRanks * kSmallMax;
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) {
    const double v = a.buf[i];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < a.world)
        reinterpret_cast<double*>(a.peer[r] + a.chan_off + kChanHeader)[slot + static_cast<size_t>(a.rank) * kSmallMax + i] = v;
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < a.world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(reinterpret_cast<unsigned*>(a.peer[r] + a.chan_off) + 8 + a.rank, epoch);
    const unsigned long long t0 = global_ns();
    while (static_cast<int>(ld_acquire_sys(flags_mine + r) - epoch) < 0) {
      if (global_ns() - t0 > kSpinLimitNs) {
        reinterpret_cast<unsigned*>(a.peer[a.rank] + kStatusOff)[0] = 1u;
        break;
      }
    }
  }
  __syncthreads();
  const double* box = reinterpret_cast<const double*>(mine + kChanHeader) + slot;
  for (int i = threadIdx.x; i < a.count; i += blockDim.x) {
    double sum = 0.0;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < a.world) sum += box[static_cast<size_t>(r) * kSmallMax + i];
    a.buf[i] = sum;
  }
  if (threadIdx.x == 0) ep[0] = epoch;
}

}  // namespace
}  // namespace pcuda

using namespace pcuda;

struct pcuda_comm {
  int rank = 0, world = 1, dev = 0;
  ncclComm_t nccl = nullptr;
  size_t cap_floats = 0;       // capacity of one data buffer
  char* local = nullptr;
  char* peer[kMaxRanks] = {};
  bool p2p = false;
  int blocks = 64;
  // small-message channels: one per caller stream, in order of first use (the same order on every rank)
  cudaStream_t chan_stream[kChannels] = {};
  int n_chan = 0;
};

extern "C" int pcuda_comm_unique_id(void* id_out, int bytes) {
  PCUDA_REQUIRE(id_out != nullptr, PCUDA_E_NULL, "comm_unique_id: NULL buffer");
  PCUDA_REQUIRE(bytes >= static_cast<int>(sizeof(ncclUniqueId)), PCUDA_E_SHAPE, "comm_unique_id: need %d bytes", static_cast<int>(sizeof(ncclUniqueId)));
  const Nccl* n = nccl();
  if (!n) return PCUDA_E_UNSUPPORTED;
  ncclUniqueId id;
  const ncclResult_t r = n->GetUniqueId(&id);
  if (r != ncclSuccess) return fail(1000 + static_cast<int>(r), "ncclGetUniqueId: %s", n->GetErrorString(r));
  memset(id_out, 0, bytes);
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

static int setup_p2p(pcuda_comm* c, size_t p2p_floats) {
  const Nccl* n = nccl();
  // capacity: a multiple of 4 * world floats, so that every rank's slice is a whole number of float4
  const size_t q = 4 * static_cast<size_t>(c->world);
  c->cap_floats = (p2p_floats + q - 1) / q * q;
  const size_t bytes = kHeaderBytes + 2 * c->cap_floats * sizeof(float) + kChannels * kChanBytes;
  int ok = 1;
  if (cudaMalloc(&c->local, bytes) != cudaSuccess) { cudaGetLastError(); c->local = nullptr; ok = 0; }
  cudaIpcMemHandle_t mine;
  memset(&mine, 0, sizeof(mine));
  if (ok) {
    cudaMemset(c->local, 0, bytes);
    if (cudaIpcGetMemHandle(&mine, c->local) != cudaSuccess) { cudaGetLastError(); ok = 0; }
  }
  // exchange (handle, ok) of every rank through NCCL
  struct Msg { cudaIpcMemHandle_t h; int ok; int pad[15]; };
  static_assert(sizeof(Msg) == 128, "Msg");
  Msg m; memset(&m, 0, sizeof(m)); m.h = mine; m.ok = ok;
  Msg* d = nullptr;
  Msg all[kMaxRanks];
  if (cudaMalloc(&d, sizeof(Msg) * (c->world + 1)) != cudaSuccess) return fail(static_cast<int>(cudaGetLastError()), "comm_init: cudaMalloc");
  cudaMemcpy(d + c->world, &m, sizeof(m), cudaMemcpyHostToDevice);
  const ncclResult_t r = n->AllGather(d + c->world, d, sizeof(Msg), ncclChar, c->nccl, nullptr);
  cudaError_t e = cudaStreamSynchronize(nullptr);
  if (r != ncclSuccess || e != cudaSuccess) { cudaFree(d); return fail(1000 + static_cast<int>(r), "comm_init: handle exchange failed"); }
  cudaMemcpy(all, d, sizeof(Msg) * c->world, cudaMemcpyDeviceToHost);
  cudaFree(d);
  for (int i = 0; i < c->world; ++i) ok = ok && all[i].ok;
  if (ok) {
    for (int i = 0; i < c->world && ok; ++i) {
      if (i == c->rank) { c->peer[i] = c->local; continue; }
      void* p = nullptr;
      if (cudaIpcOpenMemHandle(&p, all[i].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; }
      c->peer[i] = static_cast<char*>(p);
    }
  }
  // all ranks must agree (one rank without peer access disables the path everywhere)
  int* flag = nullptr;
  cudaMalloc(&flag, sizeof(int));
  cudaMemcpy(flag, &ok, sizeof(int), cudaMemcpyHostToDevice);
  n->AllReduce(flag, flag, 1, ncclInt, ncclMin, c->nccl, nullptr);
  cudaStreamSynchronize(nullptr);
  cudaMemcpy(&ok, flag, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(flag);
  c->p2p = ok != 0;
  return 0;
}

extern "C" int pcuda_comm_init(const void* unique_id, int rank, int world, size_t p2p_floats, pcuda_comm_t** out) {
  PCUDA_REQUIRE(unique_id && out, PCUDA_E_NULL, "comm_init: NULL argument");
  PCUDA_REQUIRE(world >= 1 && rank >= 0 && rank < world, PCUDA_E_SHAPE, "comm_init: rank %d of %d", rank, world);
  const Nccl* n = nccl();
  if (!n) return PCUDA_E_UNSUPPORTED;
  pcuda_comm* c = new pcuda_comm();
  c->rank = rank; c->world = world;
  cudaGetDevice(&c->dev);
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  const ncclResult_t r = n->CommInitRank(&c->nccl, world, id, rank);
  if (r != ncclSuccess) { delete c; return fail(1000 + static_cast<int>(r), "ncclCommInitRank: %s", n->GetErrorString(r)); }
  if (p2p_floats > 0 && world <= kMaxRanks) {
    if (int rc = setup_p2p(c, p2p_floats)) { n->CommDestroy(c->nccl); delete c; return rc; }
  }
  *out = c;
  return 0;
}

extern "C" int pcuda_comm_allreduce(pcuda_comm_t* c, float* buf, int64_t count, pcuda_stream_t stream) {
  PCUDA_REQUIRE(c && buf, PCUDA_E_NULL, "comm_allreduce: NULL argument");
  PCUDA_REQUIRE(count >= 0, PCUDA_E_SHAPE, "comm_allreduce: count %lld", static_cast<long long>(count));
  if (count == 0) return 0;
  const Nccl* n = nccl();
  if (!n) return PCUDA_E_UNSUPPORTED;
  const ncclResult_t r = n->AllReduce(buf, buf, static_cast<size_t>(count), ncclFloat, ncclSum, c->nccl, static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return fail(1000 + static_cast<int>(r), "ncclAllReduce: %s", n->GetErrorString(r));
  return 0;
}

extern "C" int pcuda_comm_allreduce_f64(pcuda_comm_t* c, double* buf, int64_t count, pcuda_stream_t stream) {
  PCUDA_REQUIRE(c && buf, PCUDA_E_NULL, "comm_allreduce_f64: NULL argument");
  PCUDA_REQUIRE(count >= 0, PCUDA_E_SHAPE, "comm_allreduce_f64: count %lld", static_cast<long long>(count));
  if (count == 0 || c->world == 1) return 0;
  if (c->p2p && count <= kSmallMax && !tuning(TUNE_COMM_NO_SMALL_P2P)) {
    // one-shot sum over NVLink peer memory (one ~5 us kernel instead of a ~25 us NCCL call); channel = caller stream
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int ch = -1;
    for (int i = 0; i < c->n_chan; ++i)
      if (c->chan_stream[i] == st) ch = i;
    if (ch < 0 && c->n_chan < kChannels) { ch = c->n_chan; c->chan_stream[c->n_chan++] = st; }
    if (ch >= 0) {
      SmallArgs a{};
      for (int i = 0; i < c->world; ++i) a.peer[i] = c->peer[i];
      a.rank = c->rank; a.world = c->world; a.count = static_cast<int>(count);
      a.chan_off = static_cast<long long>(kHeaderBytes + 2 * c->cap_floats * sizeof(float) + static_cast<size_t>(ch) * kChanBytes);
      a.buf = buf;
      const int threads = count <= 256 ? 256 : (count <= 512 ? 512 : 1024);
      PCUDA_LAUNCH(p2p_sum_f64_kernel, 1, threads, 0, st, a);
      count_launch(1);
      return check_launch("comm_allreduce_f64(p2p)");
    }
  }
  const Nccl* n = nccl();
  if (!n) return PCUDA_E_UNSUPPORTED;
  const ncclResult_t r = n->AllReduce(buf, buf, static_cast<size_t>(count), ncclDouble, ncclSum, c->nccl, static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return fail(1000 + static_cast<int>(r), "ncclAllReduce(f64): %s", n->GetErrorString(r));
  return 0;
}

extern "C" int pcuda_comm_allgather(pcuda_comm_t* c, const float* send, float* recv, int64_t count_per_rank, pcuda_stream_t stream) {
  PCUDA_REQUIRE(c && send && recv, PCUDA_E_NULL, "comm_allgather: NULL argument");
  PCUDA_REQUIRE(count_per_rank >= 0, PCUDA_E_SHAPE, "comm_allgather: count %lld", static_cast<long long>(count_per_rank));
  if (count_per_rank == 0) return 0;
  const Nccl* n = nccl();
  if (!n) return PCUDA_E_UNSUPPORTED;
  const ncclResult_t r = n->AllGather(send, recv, static_cast<size_t>(count_per_rank), ncclFloat, c->nccl, static_cast<cudaStream_t>(stream));
  if (r != ncclSuccess) return fail(1000 + static_cast<int>(r), "ncclAllGather: %s", n->GetErrorString(r));
  return 0;
}

namespace pcuda {
// for the BatchNorm statistics exchange inside pcuda_pointmlp_*_xf (pointmlp.cu)
int comm_world(pcuda_comm_t* c) { return c ? c->world : 1; }
int comm_sum_f64(pcuda_comm_t* c, double* buf, int64_t count, cudaStream_t st) {
  return pcuda_comm_allreduce_f64(c, buf, count, st);
}
}  // namespace pcuda

extern "C" int pcuda_comm_p2p_buffers(pcuda_comm_t* c, float** in, float** out, int64_t* capacity) {
  PCUDA_REQUIRE(c, PCUDA_E_NULL, "comm_p2p_buffers: NULL communicator");
  if (!c->p2p) {
    if (in) *in = nullptr;
    if (out) *out = nullptr;
    if (capacity) *capacity = 0;
    return 0;
  }
  if (in) *in = reinterpret_cast<float*>(c->local + kHeaderBytes);
  if (out) *out = reinterpret_cast<float*>(c->local + kHeaderBytes + c->cap_floats * sizeof(float));
  if (capacity) *capacity = static_cast<int64_t>(c->cap_floats);
  return 0;
}

extern "C" int pcuda_comm_allreduce_p2p(pcuda_comm_t* c, int64_t count, pcuda_stream_t stream) {
  PCUDA_REQUIRE(c, PCUDA_E_NULL, "comm_allreduce_p2p: NULL communicator");
  PCUDA_REQUIRE(c->p2p, PCUDA_E_UNSUPPORTED, "comm_allreduce_p2p: peer memory is not set up (p2p_floats = 0, > 8 ranks, or no peer access)");
  PCUDA_REQUIRE(count >= 0 && static_cast<size_t>(count) <= c->cap_floats, PCUDA_E_SHAPE, "comm_allreduce_p2p: count %lld exceeds the capacity %lld",
                static_cast<long long>(count), static_cast<long long>(c->cap_floats));
  if (count == 0) return 0;
  P2PArgs a{};
  for (int i = 0; i < c->world; ++i) a.peer[i] = c->peer[i];
  a.rank = c->rank; a.world = c->world;
  const long long q = 4ll * c->world;
  a.slice4 = (count + q - 1) / q;                  // float4 per rank slice (the tail beyond `count` is zero padding)
  a.cap_bytes = static_cast<long long>(c->cap_floats * sizeof(float));
  const int want = static_cast<int>((a.slice4 + kP2PThreads - 1) / kP2PThreads);
  const int tuned = tuning(TUNE_COMM_BLOCKS);
  const int cap = tuned > 0 ? tuned : c->blocks;
  const int blocks = std::max(1, std::min(std::min(want, cap), kMaxBlocks));
  PCUDA_LAUNCH(p2p_allreduce_kernel, blocks, kP2PThreads, 0, static_cast<cudaStream_t>(stream), a);
  count_launch(1);
  return check_launch("comm_allreduce_p2p");
}

extern "C" int pcuda_comm_status(pcuda_comm_t* c) {
  PCUDA_REQUIRE(c, PCUDA_E_NULL, "comm_status: NULL communicator");
  if (!c->p2p) return 0;
  unsigned st = 0;
  if (cudaMemcpy(&st, c->local + kStatusOff, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess) return static_cast<int>(cudaGetLastError());
  if (st != 0) return fail(PCUDA_E_UNSUPPORTED, "comm_allreduce_p2p: a peer did not arrive within the spin limit");
  return 0;
}

extern "C" int pcuda_comm_info(pcuda_comm_t* c, int* rank, int* world, int* p2p, int* nccl_version) {
  PCUDA_REQUIRE(c, PCUDA_E_NULL, "comm_info: NULL communicator");
  if (rank) *rank = c->rank;
  if (world) *world = c->world;
  if (p2p) *p2p = c->p2p ? 1 : 0;
  if (nccl_version) { const Nccl* n = nccl(); int v = 0; if (n) n->GetVersion(&v); *nccl_version = v; }
  return 0;
}

extern "C" int pcuda_comm_destroy(pcuda_comm_t* c) {
  if (!c) return 0;
  const Nccl* n = nccl();
  cudaDeviceSynchronize();
  for (int i = 0; i < c->world; ++i)
    if (c->p2p && i != c->rank && c->peer[i]) cudaIpcCloseMemHandle(c->peer[i]);
  if (c->nccl && n) {
    if (c->p2p) {          // every peer has unmapped this rank's region before it is freed
      int* flag = nullptr;
      if (cudaMalloc(&flag, sizeof(int)) == cudaSuccess) {
        cudaMemset(flag, 0, sizeof(int));
        n->AllReduce(flag, flag, 1, ncclInt, ncclSum, c->nccl, nullptr);
        cudaStreamSynchronize(nullptr);
        cudaFree(flag);
      }
    }
    n->CommDestroy(c->nccl);
  }
  if (c->local) cudaFree(c->local);
  delete c;
  return 0;
}
