// Groundwork for a tcgen05 leaf-loss kernel: the smallest complete tcgen05 program --
// TMEM allocation, hand-encoded shared-memory / instruction descriptors for kind::tf32 with K-major operands in the
// no-swizzle canonical layout, one elected thread issuing tcgen05.mma, tcgen05.commit onto an mbarrier, tcgen05.ld of the
// accumulator -- checked against a host GEMM.  NOT part of the product and not validated on a GPU yet (the round's GPU
// budget ended): the next round's first GPU call runs it to pin the descriptor encodings before they go into a kernel.
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/tcgen05_probe tools/micro/tcgen05_probe.cu
//     timeout 20 tools/micro/tcgen05_probe          # prints max |D - A B^T| ; exit code 0 = match
//
// Hang safety: every wait is a bounded try_wait loop; on a time-out the kernel raises an error word and still frees TMEM.
//
// Three modes, each compared with the same host GEMM:
//   0  SS form, A and B K-major                      (W as the B operand; G^T X needs the next one)
//   1  SS form, A MN-major (A handed over as [K][M], M contiguous -- how the planar [channel][pixel] map is staged)
//   2  TS form: A written to TMEM with tcgen05.st (32 rows per warp, one 32-bit column per K element), B in shared memory
//
// D[M=128][N=64] (fp32, TMEM) = A[128][K] * B[64][K]^T, K = 32 tf32 elements = 4 MMAs of K = 8.
// Canonical no-swizzle K-major layout (cute::UMMA::LayoutType::SWIZZLE_NONE): 8-row x 16-byte "core matrices" (8 rows x 4
// tf32), stored as 128 contiguous bytes; LBO = byte distance between core matrices adjacent along K, SBO = byte distance
// between core matrices adjacent along M / N.  Here: [row group r/8][k chunk k/4][row r%8][k%4].
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32, UMMA_K = 8;
constexpr uint32_t LBO = 128;                  // next core matrix along K
constexpr uint32_t SBO = (K / 4) * 128;        // next 8-row group: all K chunks of the previous one

__host__ __device__ inline uint32_t canon_index(int r, int k) {    // float index of element (r, k)
    return (uint32_t)((r / 8) * (K / 4) * 32 + (k / 4) * 32 + (r % 8) * 4 + (k % 4));
}

// MN-major canonical no-swizzle layout: core matrix = 8 K-rows x 16 bytes (4 consecutive M elements);
// LBO = distance between 8-row groups along K, SBO = distance between 4-element groups along M.
constexpr uint32_t LBO_MN = 128, SBO_MN = (K / 8) * 128;
__host__ __device__ inline uint32_t canon_index_mn(int m, int k) {
    return (uint32_t)((m / 4) * (K / 8) * 32 + (k / 8) * 32 + (k % 8) * 4 + (m % 4));
}

__device__ inline uint64_t smem_desc(uint32_t smem_addr_bytes, uint32_t lbo = LBO, uint32_t sbo = SBO) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3fff);            // [0,14)  start address >> 4
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;                   // [16,30) leading-dimension byte offset >> 4
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;                   // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                       // [46,48) descriptor version 1 (sm_100)
    // base offset 0, lbo mode 0, layout type [61,64) = 0 (SWIZZLE_NONE)
    return d;
}

__device__ inline uint32_t instr_desc_tf32(bool a_mn_major = false) {
    uint32_t d = a_mn_major ? (1u << 15) : 0u;   // a_major: 0 = K-major, 1 = MN-major
    d |= 1u << 4;                      // c_format  = F32
    d |= 2u << 7;                      // a_format  = TF32
    d |= 2u << 10;                     // b_format  = TF32
    // a_major = b_major = 0 (K-major), no negate, dense
    d |= (uint32_t)(N >> 3) << 17;     // n_dim
    d |= (uint32_t)(M >> 4) << 24;     // m_dim
    return d;
}

__device__ inline bool mbar_wait(uint64_t* bar, uint32_t phase, int max_spins) {
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    for (int i = 0; i < max_spins; i++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(phase) : "memory");
        if (ok) return true;
    }
    return false;
}

__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int* __restrict__ err, int mode) {
    __shared__ __align__(128) float sA[M * K];
    __shared__ __align__(128) float sB[N * K];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128)
        sA[mode == 1 ? canon_index_mn(i / K, i % K) : canon_index(i / K, i % K)] = A[i];
    for (int i = tid; i < N * K; i += 128) sB[canon_index(i / K, i % K)] = B[i];
    if (tid == 0) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {                       // one warp allocates 128 TMEM columns: 64 fp32 accumulators + 32 of A (mode 2)
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(dst), "n"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t tmem_a = tmem + 64;     // mode 2: A[128 lanes][K = 32 columns]
    if (mode == 2) {                       // every warp stores the 32 rows it can address: 32 lanes x 8 columns per instruction
        for (int c0 = 0; c0 < K; c0 += 8) {
            uint32_t v[8];
            for (int q = 0; q < 8; q++) v[q] = __float_as_uint(A[(size_t)(warp * 32 + lane) * K + c0 + q]);
            const uint32_t addr = tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n"
                         :: "r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                         : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    }

    if (warp == 0) {
        if (lane == 0) {                   // a single thread issues the MMAs and the commit
            const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
            const uint32_t idesc = instr_desc_tf32(mode == 1);
            for (int k = 0; k < K / UMMA_K; k++) {
                // K-major: advancing K by 8 tf32 = 2 core matrices = 2 * LBO bytes; MN-major: one 8-row group = LBO_MN
                const uint64_t da = mode == 1 ? smem_desc(a0 + k * LBO_MN, LBO_MN, SBO_MN) : smem_desc(a0 + k * 2 * LBO);
                const uint64_t db = smem_desc(b0 + k * 2 * LBO);
                const uint32_t accumulate = k > 0 ? 1u : 0u;
                if (mode == 2) {
                    const uint32_t ta = tmem_a + (uint32_t)(k * UMMA_K);      // 8 columns of A per MMA
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                        :: "r"(tmem), "r"(ta), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                } else {
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                        :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
                }
            }
            const uint64_t bar_addr = (uint64_t)__cvta_generic_to_shared(&bar);
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                         :: "l"(bar_addr) : "memory");
        }
        __syncwarp();
    }
    const bool done = mbar_wait(&bar, 0, 1 << 22);
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    if (!done) {
        if (tid == 0) err[0] = 1;
    } else {
        // warp w reads TMEM lanes [32 w, 32 w + 32) = rows of D; 32x32b.x8: 8 consecutive columns per instruction
        for (int c0 = 0; c0 < N; c0 += 8) {
            uint32_t v[8];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            for (int q = 0; q < 8; q++) D[(size_t)(warp * 32 + lane) * N + c0 + q] = __uint_as_float(v[q]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(128));
}

int main() {
    float *hA = (float*)malloc(sizeof(float) * M * K), *hB = (float*)malloc(sizeof(float) * N * K),
          *hD = (float*)malloc(sizeof(float) * M * N);
    for (int i = 0; i < M * K; i++) hA[i] = (float)((i * 7 + 3) % 11 - 5);          // small integers: exact in tf32
    for (int i = 0; i < N * K; i++) hB[i] = (float)((i * 5 + 1) % 9 - 4);
    float *dA, *dB, *dD;
    int *dErr, hErr = 0;
    cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K); cudaMalloc(&dD, sizeof(float) * M * N);
    cudaMalloc(&dErr, sizeof(int));
    cudaMemcpy(dA, hA, sizeof(float) * M * K, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof(float) * N * K, cudaMemcpyHostToDevice);
    int failures = 0;
    const char* names[3] = {"SS, A K-major", "SS, A MN-major", "TS, A in TMEM"};
    for (int mode = 0; mode < 3; mode++) {
        cudaMemset(dD, 0, sizeof(float) * M * N); cudaMemset(dErr, 0, sizeof(int));
        probe_kernel<<<1, 128>>>(dA, dB, dD, dErr, mode);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d (%s): CUDA error: %s\n", mode, names[mode], cudaGetErrorString(e)); return 2; }
        cudaMemcpy(hD, dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hErr, dErr, sizeof(int), cudaMemcpyDeviceToHost);
        if (hErr) { printf("mode %d (%s): timed out waiting for the MMA commit\n", mode, names[mode]); failures++; continue; }
        double worst = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += (double)hA[m * K + k] * hB[n * K + k];
                const double d = fabs(ref - hD[m * N + n]);
                if (d > worst) worst = d;
            }
        printf("tcgen05 probe mode %d (%s): max |D - A B^T| = %g (%s)\n", mode, names[mode], worst,
               worst == 0 ? "OK" : "MISMATCH");
        failures += worst != 0;
    }
    return failures == 0 ? 0 : 1;
}
