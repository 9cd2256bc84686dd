// Launch accounting and optional per-kernel CUDA-event timing on the launching stream
// (bench.py uses it for the roofline line; disabled by default, zero cost when off).
#include <mutex>
#include <vector>

#include <cstdlib>
#include "erd_common.cuh"

namespace erd {

static const char* kKernelNames[kNumKernels] = {"ers_scan", "ers_flags", "ers_select", "atss_candidates", "atss_finalize",
                                                "pos_prepass", "nms_prep", "nms_mask", "nms_resolve", "nms_order", "upstream_check",
                                                "student_pass", "box_fix", "teacher_head"};

struct ProfState {
  std::mutex mu;
  unsigned int mask = 0;   // bit k: time kernel k
  unsigned long long launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending[kNumKernels];
  std::vector<cudaEvent_t> pool;
  cudaEvent_t ref = nullptr;   // erd_profile_mark
};
static ProfState g_prof;

// Inside a stream capture an event record must be an external event-record node to be timed
// after the graph has run.
static cudaError_t record(cudaEvent_t e, cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive)
    return cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
  return cudaEventRecord(e, st);
}

static cudaEvent_t take_event() {
  if (!g_prof.pool.empty()) {
    cudaEvent_t e = g_prof.pool.back();
    g_prof.pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

void prof_begin(int id, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  ++g_prof.launches;
  if (!(g_prof.mask >> id & 1u)) return;
  cudaEvent_t a = take_event(), b = take_event();
  record(a, st);
  g_prof.pending[id].push_back({a, b});
}

#ifdef ERD_DEV_ABLATE
bool ablated(int id) {
  static long mask = -1;
  if (mask < 0) { const char* e = getenv("ERD_ABLATE"); mask = e ? strtol(e, nullptr, 0) : 0; }
  return (mask >> id) & 1;
}
#endif

void prof_end(int id, cudaStream_t st) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!(g_prof.mask >> id & 1u)) return;
  if (!g_prof.pending[id].empty()) record(g_prof.pending[id].back().second, st);
}

}  // namespace erd

using namespace erd;

extern "C" {

int erd_profile_enable(unsigned int kernel_mask) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  g_prof.mask = kernel_mask;
  return ERD_OK;
}

unsigned long long erd_launch_count(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  return g_prof.launches;
}

// Timeline support: mark a reference point on `stream`; erd_profile_timeline then reports, for the
// most recent profiled launch of each kernel, its start and end relative to that mark (ms).
int erd_profile_mark(void* stream) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.ref) cudaEventCreate(&g_prof.ref);
  return record(g_prof.ref, (cudaStream_t)stream) == cudaSuccess ? ERD_OK : ERD_ERR_CUDA;
}

int erd_profile_timeline(float* start_ms, float* end_ms) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  if (!g_prof.ref) return ERD_ERR_NULL;
  for (int k = 0; k < kNumKernels; ++k) {
    start_ms[k] = end_ms[k] = -1.f;
    if (g_prof.pending[k].empty()) continue;
    auto& pr = g_prof.pending[k].back();
    if (cudaEventSynchronize(pr.second) != cudaSuccess) continue;
    cudaEventElapsedTime(&start_ms[k], g_prof.ref, pr.first);
    cudaEventElapsedTime(&end_ms[k], g_prof.ref, pr.second);
  }
  return ERD_OK;
}

int erd_profile_num_kernels(void) { return kNumKernels; }
const char* erd_profile_kernel_name(int id) { return (id >= 0 && id < kNumKernels) ? kKernelNames[id] : ""; }

// Sums the elapsed time of every profiled launch since the last call (blocks until those
// launches have finished).  total_ms / count are arrays of erd_profile_num_kernels().
int erd_profile_collect(float* total_ms, int* count) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int k = 0; k < kNumKernels; ++k) {
    float tot = 0.f;
    int n = 0;
    for (auto& pr : g_prof.pending[k]) {
      float ms = 0.f;
      if (cudaEventSynchronize(pr.second) == cudaSuccess && cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        tot += ms;
        ++n;
      }
      g_prof.pool.push_back(pr.first);
      g_prof.pool.push_back(pr.second);
    }
    g_prof.pending[k].clear();
    if (total_ms) total_ms[k] = tot;
    if (count) count[k] = n;
  }
  return ERD_OK;
}

}  // extern "C"
