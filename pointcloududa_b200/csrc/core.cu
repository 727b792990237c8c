// libpcuda core: version, error strings, per-process device facts, tuning knobs.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "pcuda_common.cuh"

namespace pcuda {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

static std::mutex g_mu;
static int g_sm_count[64];
static bool g_sm_known[64];

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_sm_known[dev]) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    g_sm_count[dev] = n;
    g_sm_known[dev] = true;
  }
  return g_sm_count[dev];
}

// Opt-in to > 48 KB of dynamic shared memory.  The attribute is per (function, device): remembered per device so
// that a process driving several GPUs sets it on each of them (a process-wide flag left every device but the
// first without it), and guarded by the context mutex (launches may come from several host threads).
static std::map<std::pair<const void*, int>, int> g_smem_optin;

cudaError_t smem_optin_impl(const void* fn, int bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(fn, dev);
  auto it = g_smem_optin.find(key);
  if (it != g_smem_optin.end() && it->second >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) g_smem_optin[key] = bytes;
  return e;
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n)); }

static std::atomic<int> g_tune[TUNE_NKEYS];
int tuning(int key) { return (key >= 0 && key < TUNE_NKEYS) ? g_tune[key].load() : 0; }

}  // namespace pcuda

extern "C" {

int pcuda_version(void) { return PCUDA_VERSION; }

const char* pcuda_last_error_string(void) { return pcuda::g_err; }

const char* pcuda_error_name(int code) {
  switch (code) {
    case PCUDA_OK: return "PCUDA_OK";
    case PCUDA_E_NULL: return "PCUDA_E_NULL";
    case PCUDA_E_SHAPE: return "PCUDA_E_SHAPE";
    case PCUDA_E_UNSUPPORTED: return "PCUDA_E_UNSUPPORTED";
    case PCUDA_E_ALIGN: return "PCUDA_E_ALIGN";
    case PCUDA_E_WORKSPACE: return "PCUDA_E_WORKSPACE";
    default: return code >= 1000 ? "NCCL error (1000 + ncclResult_t)" : (code > 0 ? cudaGetErrorName(static_cast<cudaError_t>(code)) : "PCUDA_E_?");
  }
}

int pcuda_sm_count(void) { return pcuda::sm_count(); }

uint64_t pcuda_launch_count(void) { return pcuda::g_launches.load(); }

// Tuning knobs for benchmarking kernel variants; not part of the reference-facing contract.
int pcuda_tune(int key, int value) {
  if (key < 0 || key >= pcuda::TUNE_NKEYS) return PCUDA_E_UNSUPPORTED;
  pcuda::g_tune[key].store(value);
  return 0;
}

}  // extern "C"
