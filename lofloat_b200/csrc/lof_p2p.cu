// lof_p2p.cu -- flag hand-shake for the GEMM's peer-to-peer operand exchange (lofloat_b200/sharding.py PeerPanels).
//
// The 2-D partitioned GEMM needs, per rank, the A row panel of its row group and the B^T column panel of its column: a pure operand
// broadcast.  With NCCL that is two collectives (~0.15 ms of launch / rendez-vous latency at 8192^3, more than the per-rank GEMM itself at
// N >= 4); here every rank maps its peers' operand shards through CUDA IPC once and PULLS what it needs over NVLink with plain copies on side
// streams, overlapped with the GEMM on its own shard.  What is left to do is ordering, and that is these three one-thread kernels on a
// 16-byte flag block per rank that lives in IPC-mapped memory:   flags[0] = epoch of the rank's published shards,  flags[1] = number of
// remote pulls of its shards that have completed (all epochs).
//   post(e, expected_done)  owner's stream, after the kernels that wrote the shards: wait until every pull of epoch e - 1 has finished
//                           (the shards may have been overwritten since), then publish epoch e;
//   wait(peer, e)           consumer's side stream, before the copies from that peer;
//   done(peer)              consumer's side stream, after its copies from that peer.
// Spins are bounded (~2 s of SM clocks): a protocol error raises the sticky flag flags[3] instead of hanging a shared GPU.
#include <cuda_runtime.h>
#include <stdint.h>

namespace lof {

namespace {

constexpr long long kSpinLimit = 4000000000ll;

__global__ void p2p_post_kernel(volatile int* flags, int epoch, int expected_done)
{
    const long long t0 = clock64();
    while (flags[1] < expected_done) {
        if (clock64() - t0 > kSpinLimit) { flags[3] = 1; break; }
    }
    __threadfence_system();
    flags[0] = epoch;
    __threadfence_system();
}

__global__ void p2p_wait_kernel(const volatile int* peer_flags, volatile int* my_flags, int epoch)
{
    const long long t0 = clock64();
    while (peer_flags[0] < epoch) {
        if (clock64() - t0 > kSpinLimit) { my_flags[3] = 1; break; }
    }
    __threadfence_system();
}

__global__ void p2p_done_kernel(int* peer_flags)
{
    __threadfence_system();
    atomicAdd_system(&peer_flags[1], 1);
}

}  // namespace

cudaError_t launch_p2p_post(int* flags, int epoch, int expected_done, cudaStream_t stream)
{
    p2p_post_kernel<<<1, 1, 0, stream>>>(flags, epoch, expected_done);
    return cudaGetLastError();
}
cudaError_t launch_p2p_wait(const int* peer_flags, int* my_flags, int epoch, cudaStream_t stream)
{
    p2p_wait_kernel<<<1, 1, 0, stream>>>(peer_flags, my_flags, epoch);
    return cudaGetLastError();
}
cudaError_t launch_p2p_done(int* peer_flags, cudaStream_t stream)
{
    p2p_done_kernel<<<1, 1, 0, stream>>>(peer_flags);
    return cudaGetLastError();
}

}  // namespace lof
