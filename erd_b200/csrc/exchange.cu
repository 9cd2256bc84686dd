// The path's one collective -- reduce_mean of the two avg factors (mmdet/utils/dist_utils.py:
// 59-65; call sites gfl_head_increment_erd.py:390-391,406-407) -- as ONE tiny kernel over
// NVLink peer memory instead of an NCCL launch: 8 bytes per rank, so the cost is pure latency
// and sits on the chain in front of every sweep.  Each rank owns a symmetric buffer every peer
// has mapped; a step is
//   store (a0, a1) into slot[rank] of every peer's buffer, release-store the epoch behind it
//   acquire-spin on the W slots of the own buffer until they carry this epoch
//   sum a_r / W in rank order (identical bits on every rank)
// A peer that does not arrive within the spin bound (minutes) turns both factors into NaN.
// The epoch lives in device memory (the kernel increments it), so the launch is CUDA-graph
// replayable; slots are double-buffered by epoch parity, which is enough because no rank can
// finish step k+1 before every rank has contributed to it, i.e. has finished reading step k.
#include "erd_common.cuh"

namespace erd {

constexpr int kMaxRanks = 64;

struct ExchangeSlot {
  float a0, a1;
  unsigned int epoch, pad;
};
struct ExchangePeers {
  unsigned char* buf[kMaxRanks];
};
constexpr size_t kSlotBytes = sizeof(ExchangeSlot) * 2 * kMaxRanks;   // then: epoch counter, status

__global__ void __launch_bounds__(kMaxRanks) avg_exchange_kernel(float* __restrict__ avg, ExchangePeers peers, int rank,
                                                                 int world) {
  __shared__ unsigned int s_epoch;
  __shared__ float s_v[kMaxRanks][2];
  unsigned char* mine = peers.buf[rank];
  unsigned int* ctr = reinterpret_cast<unsigned int*>(mine + kSlotBytes);
  if (threadIdx.x == 0) {
    s_epoch = ctr[0] + 1u;
    ctr[0] = s_epoch;
  }
  __syncthreads();
  const unsigned int e = s_epoch;
  const int par = (int)(e & 1u);
  const int t = threadIdx.x;
  if (t < world) {
    ExchangeSlot* dst = reinterpret_cast<ExchangeSlot*>(peers.buf[t]) + par * kMaxRanks + rank;
    *reinterpret_cast<volatile float*>(&dst->a0) = avg[0];
    *reinterpret_cast<volatile float*>(&dst->a1) = avg[1];
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&dst->epoch), "r"(e) : "memory");
    const ExchangeSlot* src = reinterpret_cast<const ExchangeSlot*>(mine) + par * kMaxRanks + t;
    unsigned int seen = 0;
    long long spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(&src->epoch) : "memory");
    } while (seen != e && ++spins < (1ll << 28));   // bounded: a lost peer must not hang the GPU
    // A peer that never arrived is FATAL, not silent: both factors become NaN, so every loss of the
    // step is NaN (CheckInvalidLossHook / the caller's own check fires), and the status word says why.
    const bool lost = seen != e;
    if (lost) ctr[1] = 1u;
    s_v[t][0] = lost ? __int_as_float(0x7fc00000) : *reinterpret_cast<const volatile float*>(&src->a0);
    s_v[t][1] = lost ? __int_as_float(0x7fc00000) : *reinterpret_cast<const volatile float*>(&src->a1);
  }
  __syncthreads();
  if (t < 2) {
    const float w = (float)world;
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += s_v[r][t] / w;   // t.div_(world) then SUM, rank order
    avg[t] = s;
  }
}

}  // namespace erd

extern "C" {

size_t erd_avg_exchange_bytes(void) { return erd::kSlotBytes + 16; }

int erd_avg_exchange(float* avg, void* const* peer_bufs, int32_t rank, int32_t world, void* stream) {
  if (!avg || !peer_bufs || world < 1 || world > erd::kMaxRanks || rank < 0 || rank >= world) return -2;
  erd::ExchangePeers p;
  for (int i = 0; i < erd::kMaxRanks; ++i) p.buf[i] = i < world ? (unsigned char*)peer_bufs[i] : nullptr;
  for (int i = 0; i < world; ++i)
    if (!p.buf[i]) return -2;
  erd::avg_exchange_kernel<<<1, erd::kMaxRanks, 0, (cudaStream_t)stream>>>(avg, p, rank, world);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // extern "C"
