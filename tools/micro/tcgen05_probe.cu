// tcgen05 probe: pins the hand-encoded descriptors of the leaf-loss kernel (hier_slam_b200/csrc/leaf_loss_tc.cu) against a host
// GEMM, one feature at a time -- TMEM allocation, shared-memory / instruction descriptors for kind::tf32 in the no-swizzle
// canonical layouts, one elected thread issuing tcgen05.mma, tcgen05.commit onto an mbarrier, tcgen05.ld / tcgen05.st.
// NOT part of the product.
//
//     nvcc -gencode arch=compute_100a,code=sm_100a -o tools/micro/tcgen05_probe tools/micro/tcgen05_probe.cu
//     timeout 60 tools/micro/tcgen05_probe          # one line per case; exit code 0 = all match
//
// Hang safety: every wait is a bounded try_wait loop; on a time-out the kernel raises an error word and still frees TMEM.
//
// D[M=128][N] (fp32, TMEM) = A[128][K] * B[N][K]^T.  Cases:
//   a_mode 0  A in shared memory, K-major          b_mode 0  B K-major   (B handed over as [N][K], K contiguous)
//   a_mode 1  A in shared memory, MN-major         b_mode 1  B MN-major  (B handed over as [K][N], N contiguous)
//   a_mode 2  A in TMEM (tcgen05.st), TS form
//   split 1   3xTF32: D = A_hi B_hi + A_lo B_hi + A_hi B_lo as three accumulating MMA sequences (hi = upper 19 bits)
// Canonical no-swizzle layouts (cute/atom/mma_traits_sm100.hpp, LayoutType::INTERLEAVE), T = 4 tf32 per 16 bytes:
//   K-major : ((8,m),(T,2)) : ((1T,SBO),(1,LBO))  -- 8 rows x 16 B core matrices, LBO between K chunks, SBO between row groups
//   MN-major: ((T,1,m),(8,k)) : ((1,T,SBO),(1T,LBO)) -- 8 K-rows x 16 B (4 MN elements), SBO between MN groups, LBO between K groups
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

constexpr int M = 128, UMMA_K = 8;

__host__ __device__ inline uint32_t idx_kmajor(int r, int k, int K) {      // float index of element (row r, k)
    return (uint32_t)((r / 8) * (K / 4) * 32 + (k / 4) * 32 + (r % 8) * 4 + (k % 4));
}
__host__ __device__ inline uint32_t idx_mnmajor(int m, int k, int K) {     // float index of element (mn index m, k)
    return (uint32_t)((m / 4) * (K / 8) * 32 + (k / 8) * 32 + (k % 8) * 4 + (m % 4));
}

// SWIZZLE_128B K-major (LayoutType::B128: Swizzle<3,4,3> o ((8,m),(T,2)):((8T,SBO),(1,T))): a row holds 32 consecutive K
// elements (128 B), 8 rows = one 1024-B swizzle atom in which the 16-byte chunk index is XORed with the row index; K extents
// beyond 32 are further [rows x 32] blocks.  The base must be 1024-byte aligned.
__host__ __device__ inline uint32_t idx_sw128(int r, int k, int rows) {
    const int kb = k / 32, kk = k % 32;
    return (uint32_t)(kb * rows * 32 + (r / 8) * 256 + (r % 8) * 32 + (((kk / 4) ^ (r % 8)) * 4) + (kk % 4));
}
__device__ inline uint64_t smem_desc_sw128(uint32_t smem_addr_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                                       // LBO: unused for swizzled K-major (convention: 1)
    d |= (uint64_t)(1024 >> 4) << 32;                             // SBO: 8-row group pitch
    d |= (uint64_t)1 << 46;                                       // version
    d |= (uint64_t)2 << 61;                                       // layout type SWIZZLE_128B
    return d;
}

__device__ inline uint64_t smem_desc(uint32_t smem_addr_bytes, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr_bytes >> 4) & 0x3fff);            // [0,14)  start address >> 4
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;                   // [16,30) leading-dimension byte offset >> 4
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;                   // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                       // [46,48) descriptor version 1 (sm_100)
    return d;                                                     // base offset 0, lbo mode 0, layout type SWIZZLE_NONE
}

__device__ inline uint32_t instr_desc_tf32(int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                      // c_format  = F32
    d |= 2u << 7;                      // a_format  = TF32
    d |= 2u << 10;                     // b_format  = TF32
    d |= (a_mn ? 1u : 0u) << 15;       // a_major: 0 = K-major, 1 = MN-major
    d |= (b_mn ? 1u : 0u) << 16;       // b_major
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
__device__ inline float hi_part(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// dynamic shared memory: sA[M*K] sAlo[M*K] sB[N*K] sBlo[N*K]
__global__ void __launch_bounds__(128) probe_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ D, int* __restrict__ err, int N, int K, int a_mode,
                                                    int b_mode, int split, int swap_mn) {
    extern __shared__ __align__(1024) float dsm[];
    float* sA = dsm;
    float* sAlo = sA + M * K;
    float* sB = sAlo + M * K;
    float* sBlo = sB + N * K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        const float v = A[i];
        const uint32_t o = a_mode == 3 ? idx_sw128(i / K, i % K, M) : a_mode == 1 ? idx_mnmajor(i / K, i % K, K) : idx_kmajor(i / K, i % K, K);
        sA[o] = split ? hi_part(v) : v;
        sAlo[o] = v - hi_part(v);
    }
    for (int i = tid; i < N * K; i += 128) {
        const float v = B[i];
        const uint32_t o = b_mode == 3 ? idx_sw128(i / K, i % K, N) : b_mode == 1 ? idx_mnmajor(i / K, i % K, K) : idx_kmajor(i / K, i % K, K);
        sB[o] = split ? hi_part(v) : v;
        sBlo[o] = v - hi_part(v);
    }
    if (tid == 0) {
        const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" :: "r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {                       // one warp allocates all 512 TMEM columns: N accumulators + K columns of A (mode 2)
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmem_base);
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(dst), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = tmem_base;
    const uint32_t tmem_a = tmem + 256;    // mode 2: A[128 lanes][K columns] (hi), A_lo behind it
    if (a_mode == 2) {                     // every warp stores the 32 rows it can address: 32 lanes x 8 columns per instruction
        for (int part = 0; part < (split ? 2 : 1); part++)
            for (int c0 = 0; c0 < K; c0 += 8) {
                uint32_t v[8];
                for (int q = 0; q < 8; q++) {
                    const float x = A[(size_t)(warp * 32 + lane) * K + c0 + q];
                    v[q] = __float_as_uint(part ? x - hi_part(x) : (split ? hi_part(x) : x));
                }
                const uint32_t addr = tmem_a + ((uint32_t)(warp * 32) << 16) + (uint32_t)(part * K + c0);
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
            const uint32_t idesc = instr_desc_tf32(N, a_mode == 1, b_mode == 1);
            // byte distances of the canonical layouts
            const uint32_t lbo_k = 128, sbo_k = (uint32_t)(K / 4) * 128;       // K-major
            uint32_t lbo_mn = 128, sbo_mn = (uint32_t)(K / 8) * 128;           // MN-major
            if (swap_mn) { const uint32_t t_ = lbo_mn; lbo_mn = sbo_mn; sbo_mn = t_; }   // experiment: the two fields exchanged
            bool first = true;
            for (int pass = 0; pass < (split ? 3 : 1); pass++) {               // hi*hi, lo*hi, hi*lo
                const float* pa = pass == 1 ? sAlo : sA;
                const float* pb = pass == 2 ? sBlo : sB;
                const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(pa), b0 = (uint32_t)__cvta_generic_to_shared(pb);
                for (int k = 0; k < K / UMMA_K; k++) {
                    // K-major: 8 tf32 along K = 2 core matrices = 2 LBO; MN-major: one 8-row K group = 1 LBO
                    // SWIZZLE_128B: K step of 8 tf32 = 32 bytes inside the 128-byte row; every 4th step enters the next [rows x 32] block
                    const uint64_t da3 = smem_desc_sw128(a0 + (k / 4) * M * 128 + (k % 4) * 32), db3 = smem_desc_sw128(b0 + (k / 4) * N * 128 + (k % 4) * 32);
                    const uint64_t da = a_mode == 3 ? da3 : a_mode == 1 ? smem_desc(a0 + k * 128, lbo_mn, sbo_mn) : smem_desc(a0 + k * 2 * lbo_k, lbo_k, sbo_k);
                    const uint64_t db = b_mode == 3 ? db3 : b_mode == 1 ? smem_desc(b0 + k * 128, lbo_mn, sbo_mn) : smem_desc(b0 + k * 2 * lbo_k, lbo_k, sbo_k);
                    const uint32_t accumulate = first ? 0u : 1u;
                    first = false;
                    if (a_mode == 2) {
                        const uint32_t ta = tmem_a + (uint32_t)((pass == 1 ? K : 0) + k * UMMA_K);      // 8 columns of A per MMA
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
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "n"(512));
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }
static float round_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u += 0x1000u; u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

struct Case { const char* name; int N, K, a_mode, b_mode, split, exact_ints, swap_mn; };

int main() {
    const Case cases[] = {
        {"SS  A K-major   B K-major   N=64  K=32 ints", 64, 32, 0, 0, 0, 1},
        {"TS  A in TMEM   B K-major   N=64  K=32 ints", 64, 32, 2, 0, 0, 1},
        {"SS  A K-major   B K-major   N=256 K=32 ints", 256, 32, 0, 0, 0, 1},
        {"SS  A K-major   B K-major   N=192 K=80 random, 3xTF32", 192, 80, 0, 0, 1, 0},
        {"TS  A in TMEM   B K-major   N=80  K=96 random, 3xTF32", 80, 96, 2, 0, 1, 0},
        {"SS  A K-major   B K-major   N=192 K=80 random, single tf32", 192, 80, 0, 0, 0, 0},
        {"SS  A SW128     B SW128     N=64  K=32 ints", 64, 32, 3, 3, 0, 1},
        {"SS  A SW128     B K-major   N=64  K=32 ints", 64, 32, 3, 0, 0, 1},
        {"SS  A SW128     B SW128     N=80  K=128 ints", 80, 128, 3, 3, 0, 1},
        {"SS  A SW128     B SW128     N=80  K=128 random, 3xTF32", 80, 128, 3, 3, 1, 0},
    };
    int failures = 0;
    for (const Case& c : cases) {
        const int N = c.N, K = c.K;
        float *hA = (float*)malloc(sizeof(float) * M * K), *hB = (float*)malloc(sizeof(float) * N * K),
              *hD = (float*)malloc(sizeof(float) * M * N);
        // A is handed over as [M][K]; B as [N][K] (b_mode only changes how the kernel lays it out in shared memory)
        uint32_t s = 12345u;
        auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)((s >> 8) & 0xffff) / 65536.0f - 0.5f; };
        for (int i = 0; i < M * K; i++) hA[i] = c.exact_ints ? (float)((i * 7 + 3) % 11 - 5) : rnd() * 3.1f;
        for (int i = 0; i < N * K; i++) hB[i] = c.exact_ints ? (float)((i * 5 + 1) % 9 - 4) : rnd() * 1.7f;
        float *dA, *dB, *dD;
        int *dErr, hErr = 0;
        cudaMalloc(&dA, sizeof(float) * M * K); cudaMalloc(&dB, sizeof(float) * N * K); cudaMalloc(&dD, sizeof(float) * M * N);
        cudaMalloc(&dErr, sizeof(int));
        cudaMemcpy(dA, hA, sizeof(float) * M * K, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, hB, sizeof(float) * N * K, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, sizeof(float) * M * N); cudaMemset(dErr, 0, sizeof(int));
        const size_t smem = sizeof(float) * (size_t)(2 * M * K + 2 * N * K);
        cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        probe_kernel<<<1, 128, smem>>>(dA, dB, dD, dErr, N, K, c.a_mode, c.b_mode, c.split, c.swap_mn);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: CUDA error: %s\n", c.name, cudaGetErrorString(e)); return 2; }
        cudaMemcpy(hD, dD, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hErr, dErr, sizeof(int), cudaMemcpyDeviceToHost);
        if (hErr) { printf("%s: timed out waiting for the MMA commit\n", c.name); failures++; continue; }
        double worst = 0, worst_trunc = 0, worst_round = 0, scale = 0;
        for (int m = 0; m < M; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0, rt = 0, rr = 0;
                for (int k = 0; k < K; k++) {
                    ref += (double)hA[m * K + k] * hB[n * K + k];
                    rt += (double)trunc_tf32(hA[m * K + k]) * trunc_tf32(hB[n * K + k]);
                    rr += (double)round_tf32(hA[m * K + k]) * round_tf32(hB[n * K + k]);
                }
                worst = fmax(worst, fabs(ref - hD[m * N + n]));
                worst_trunc = fmax(worst_trunc, fabs(rt - hD[m * N + n]));
                worst_round = fmax(worst_round, fabs(rr - hD[m * N + n]));
                scale = fmax(scale, fabs(ref));
            }
        const bool ok = c.exact_ints ? worst == 0 : (c.split ? worst < 2e-6 * scale * 4 : worst < 4e-3 * scale);
        printf("tcgen05 probe [%s]: max|D-ref| = %.3g (|ref|max %.3g), vs truncated-input ref %.3g, vs rounded-input ref %.3g  %s\n",
               c.name, worst, scale, worst_trunc, worst_round, ok ? "OK" : "MISMATCH");
        failures += !ok;
        if (!ok) {
            printf("    D[0][0..7]   :"); for (int n = 0; n < 8; n++) printf(" %g", hD[n]); printf("\n    ref[0][0..7] :");
            for (int n = 0; n < 8; n++) { double r = 0; for (int k = 0; k < K; k++) r += (double)hA[k] * hB[n * K + k]; printf(" %g", r); }
            printf("\n    D[1][0..3] D[8][0..3]:"); for (int n = 0; n < 4; n++) printf(" %g", hD[N + n]); for (int n = 0; n < 4; n++) printf(" %g", hD[8 * N + n]);
            printf("\n");
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dErr); free(hA); free(hB); free(hD);
    }
    return failures == 0 ? 0 : 1;
}
