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

// (structures and the post / wait halves: erd_common.cuh -- inside a fused step the assignment prepass posts and
// the student pass waits, so that no launch of its own sits between them; this stand-alone kernel does both)
__global__ void __launch_bounds__(kMaxRanks) avg_exchange_kernel(float* __restrict__ avg, ExchangeInfo x) {
  __shared__ unsigned int s_epoch;
  __shared__ float s_v[kMaxRanks][2];
  exchange_post(x, avg[0], avg[1], &s_epoch);
  float a0, a1;
  exchange_wait(x, s_epoch, s_v, a0, a1);
  if (threadIdx.x == 0) {
    avg[0] = a0;
    avg[1] = a1;
  }
}

}  // namespace erd

extern "C" {

size_t erd_avg_exchange_bytes(void) { return erd::kExchangeSlotBytes + 16; }

int erd_avg_exchange(float* avg, void* const* peer_bufs, int32_t rank, int32_t world, void* stream) {
  if (!avg || !peer_bufs || world < 1 || world > erd::kMaxRanks || rank < 0 || rank >= world) return -2;
  erd::ExchangeInfo x;
  for (int i = 0; i < erd::kMaxRanks; ++i) x.peers.buf[i] = i < world ? (unsigned char*)peer_bufs[i] : nullptr;
  for (int i = 0; i < world; ++i)
    if (!x.peers.buf[i]) return -2;
  x.rank = rank;
  x.world = world;
  erd::avg_exchange_kernel<<<1, erd::kMaxRanks, 0, (cudaStream_t)stream>>>(avg, x);
  return cudaGetLastError() == cudaSuccess ? 0 : -4;
}

}  // extern "C"
