// One-shot all-reduce over NVLink peer memory for the tiny fp64 messages of synchronised BatchNorm (sum x, sum x^2
// forward; sum g, sum g*xhat backward; the loss normaliser): 2 x C doubles, 84 + 84 times per two-encoder step.
// A NCCL all-reduce of 1-8 KB is pure launch + protocol latency; here every rank stores its vector straight into a
// slot of every peer's symmetric buffer, raises a flag there, waits for the world's flags in its own buffer and sums
// the slots in RANK ORDER (so all ranks obtain bit-identical sums).  One CTA, no host involvement, graph-capturable.
//
// The reference is single-GPU; this is the exchange that makes N ranks on shards of a
// batch reproduce its full-batch BatchNorm.
//
// Symmetric buffer layout per rank (allocated and exchanged by the host: torch symmetric memory = plumbing):
//   [0, 1024)            flags   uint64 [2 parities][64 ranks]   (monotonic call numbers)
//   [1024, ...)          slots   double [2 parities][world][max_n]
// Call k (k = 1, 2, ...) uses parity k & 1.  A peer can be at most one call ahead of a rank that is still reading its
// slots (it needs that rank's flag for call k+1 to finish k+1), so two parities are enough and nothing is ever reset.
#include "common.cuh"

namespace adni {
extern void count_launch();
namespace {

constexpr int kPeerThreads = 256;
constexpr int kPeerMaxWorld = 64;
constexpr size_t kPeerFlagBytes = 2 * kPeerMaxWorld * sizeof(unsigned long long);

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_f64_kernel(double* __restrict__ data, int n,
                                                                          const unsigned long long* __restrict__ peers,
                                                                          unsigned long long* __restrict__ call_counter,
                                                                          int rank, int world, int max_n) {
  pdl_enter();
  const unsigned long long call = *call_counter + 1;
  const int par = static_cast<int>(call & 1ull);
  // 1. my vector into slot [par][rank] of every rank's buffer (peer stores travel over NVLink)
  for (int r = 0; r < world; r++) {
    double* dst = reinterpret_cast<double*>(peers[r] + kPeerFlagBytes) + (static_cast<size_t>(par) * world + rank) * max_n;
    for (int i = threadIdx.x; i < n; i += kPeerThreads) dst[i] = data[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag in every rank's buffer, 3. wait for every rank's flag in mine
  if (threadIdx.x < world) {
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(peers[threadIdx.x]) + par * kPeerMaxWorld + rank;
    st_release_sys(remote, call);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peers[rank]) + par * kPeerMaxWorld + threadIdx.x;
    const uint64_t t0 = global_timer_ns();
    uint32_t spins = 0;
    while (ld_acquire_sys(mine) < call) {
      if ((++spins & 1023u) == 0 && global_timer_ns() - t0 > 8000000000ull) {  // 8 s: a peer is gone
        printf("adni_b200: peer all-reduce timeout rank %d waiting for rank %d call %llu\n", rank, threadIdx.x, call);
        __trap();
      }
    }
  }
  __syncthreads();
  // 4. sum the world's slots in rank order (L1 is bypassed: the lines were written by peers)
  const double* slots = reinterpret_cast<const double*>(peers[rank] + kPeerFlagBytes) + static_cast<size_t>(par) * world * max_n;
  for (int i = threadIdx.x; i < n; i += kPeerThreads) {
    double s = 0.0;
    for (int r = 0; r < world; r++) s += __ldcg(slots + static_cast<size_t>(r) * max_n + i);
    data[i] = s;
  }
  if (threadIdx.x == 0) *call_counter = call;
}

}  // namespace
}  // namespace adni

using namespace adni;

extern "C" {

size_t adni_peer_buffer_bytes(int world, int max_n) {
  return kPeerFlagBytes + sizeof(double) * 2 * static_cast<size_t>(world) * static_cast<size_t>(max_n);
}

int adni_peer_allreduce_f64(double* data, int n, const void* peers, void* call_counter, int rank, int world, int max_n,
                            void* stream) {
  ADNI_REQUIRE(data && peers && call_counter && n > 0, ADNI_EINVAL, "peer_allreduce_f64: bad arguments");
  ADNI_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world, ADNI_EINVAL,
               "peer_allreduce_f64: rank %d / world %d out of range", rank, world);
  ADNI_REQUIRE(n <= max_n, ADNI_ENOTSUP, "peer_allreduce_f64: %d elements exceed the slot size %d", n, max_n);
  pdl_launch(peer_allreduce_f64_kernel, 1, kPeerThreads, 0, static_cast<cudaStream_t>(stream))(
      data, n, static_cast<const unsigned long long*>(peers), static_cast<unsigned long long*>(call_counter), rank, world,
      max_n);
  count_launch();
  ADNI_LAUNCH_CHECK("peer_allreduce_f64_kernel");
  return ADNI_OK;
}

}  // extern "C"
