// Gradient all-reduce of the keyframe-parallel mapping mode as ONE kernel over NVLink peer memory.
//
// The flat fp32 gradient buffer of every rank lives in symmetric memory (same size on every GPU, mapped into every
// process; torch.distributed._symmetric_memory provides the allocation, the peer / multicast pointers and the signal
// pads -- plumbing).  This kernel is launched on the stream of the blend backward on every rank and does a two-shot
// all-reduce in place:
//   barrier      every rank's gradient kernels have finished (release / acquire flags in the peers' signal pads)
//   reduce       rank r owns the slice [r n / G, (r + 1) n / G): it loads the G copies of that slice and adds them --
//                with NVSwitch multicast (NVLS) ONE multimem.ld_reduce.add.v4.f32 per 16 bytes performs the G loads and
//                the reduction inside the switch; without multicast, G - 1 peer loads (ld.global over NVLink) per vector --
//   broadcast    and stores the sum to all G copies: ONE multimem.st.v4.f32 (the switch replicates it), or G peer stores
//   barrier      all slices have been written everywhere before any rank's next kernel reads the buffer.
// Per GPU and direction the multicast form moves n bytes over its NVLink ports (n / G reduced in + ... the switch does
// the fan-in / fan-out), against 2 (G - 1) / G n for a ring: for the 48 MB gradient of BASELINE config 2 on 8 GPUs that
// is the difference between ~0.25 ms (NCCL ring, measured) and ~0.1 ms.
//
// Kernels of different ranks wait for one another here (like every collective); each rank owns its GPU, the grid is
// sized to be fully resident, and every spin is bounded: a peer that never arrives turns into a trap (a CUDA error on the
// host), not a hung GPU.
#include "hs_common.cuh"

namespace hs {
namespace ar {

constexpr int MAX_WORLD = 16;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

struct Peers {
    uint32_t* pad[MAX_WORLD];      // signal pad of every rank (peer-mapped): [blocks][world] words per barrier slot
    float* buf[MAX_WORLD];         // gradient buffer of every rank (peer-mapped); unused with multicast
};

// Block-wise barrier across the ranks.  Block b of rank r raises word [b][r] in EVERY rank's pad to `epoch` and waits until
// its own pad holds `epoch` in the words [b][0..G).  Epochs increase monotonically from call to call (the pad is zeroed
// once at set-up), so no reset traffic is needed.
__device__ __forceinline__ void rank_barrier(const Peers& P, int rank, int world, uint32_t epoch) {
    __syncthreads();
    if (threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(P.pad[threadIdx.x] + (size_t)blockIdx.x * world + rank, epoch);
        const uint32_t* mine = P.pad[rank] + (size_t)blockIdx.x * world + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (global_ns() - t0 > 4000000000ull) __trap();      // 4 s: a peer never arrived
        }
    }
    __syncthreads();
}

template <bool MULTICAST>
__global__ void __launch_bounds__(512, 1) allreduce_sum_kernel(Peers P, float* mc, int rank, int world, size_t n4,
                                                               uint32_t epoch) {
    rank_barrier(P, rank, world, epoch);
    // this rank's slice, in 16-byte vectors
    const size_t per = (n4 + world - 1) / world;
    const size_t lo = (size_t)rank * per, hi = min(n4, lo + per);
    constexpr int U = 4;     // independent 16-byte vectors in flight per thread
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i0 = lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
        if (MULTICAST) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; u++)
                if (i0 + u * stride < hi)
                    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];\n"
                                 : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                 : "l"(reinterpret_cast<float4*>(mc) + i0 + u * stride) : "memory");
#pragma unroll
            for (int u = 0; u < U; u++)
                if (i0 + u * stride < hi)
                    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n"
                                 ::"l"(reinterpret_cast<float4*>(mc) + i0 + u * stride), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z),
                                 "f"(v[u].w) : "memory");
        } else {
            float4 acc[U];
#pragma unroll
            for (int u = 0; u < U; u++)
                acc[u] = (i0 + u * stride < hi) ? reinterpret_cast<const float4*>(P.buf[0])[i0 + u * stride] : make_float4(0.f, 0.f, 0.f, 0.f);
            for (int r = 1; r < world; r++) {      // fixed rank order: every rank computes bit-identical sums
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (i0 + u * stride < hi) {
                        const float4 v = reinterpret_cast<const float4*>(P.buf[r])[i0 + u * stride];
                        acc[u].x += v.x;
                        acc[u].y += v.y;
                        acc[u].z += v.z;
                        acc[u].w += v.w;
                    }
                }
            }
            for (int r = 0; r < world; r++) {
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (i0 + u * stride < hi) reinterpret_cast<float4*>(P.buf[r])[i0 + u * stride] = acc[u];
            }
        }
    }
    rank_barrier(P, rank, world, epoch + 1);
}

}  // namespace ar

int launch_allreduce_sum(void* multicast_ptr, void* const* peer_bufs, void* const* peer_pads, int rank, int world, size_t n,
                         unsigned epoch, int blocks, cudaStream_t stream) {
    using namespace ar;
    if (world < 2) return 0;
    if (world > MAX_WORLD || rank < 0 || rank >= world || peer_pads == nullptr || (multicast_ptr == nullptr && peer_bufs == nullptr)) {
        set_error("hs_allreduce_sum: bad arguments (world %d, rank %d)", world, rank);
        return 1;
    }
    if (n % 4 != 0) {
        set_error("hs_allreduce_sum: the element count must be a multiple of 4 (16-byte vectors), got %zu", n);
        return 1;
    }
    Peers P;
    for (int r = 0; r < MAX_WORLD; r++) {
        P.pad[r] = r < world ? (uint32_t*)peer_pads[r] : nullptr;
        P.buf[r] = (r < world && peer_bufs != nullptr) ? (float*)peer_bufs[r] : nullptr;
    }
    if (blocks <= 0) blocks = 128;
    if (multicast_ptr != nullptr)
        allreduce_sum_kernel<true><<<blocks, 512, 0, stream>>>(P, (float*)multicast_ptr, rank, world, n / 4, epoch);
    else
        allreduce_sum_kernel<false><<<blocks, 512, 0, stream>>>(P, nullptr, rank, world, n / 4, epoch);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs
