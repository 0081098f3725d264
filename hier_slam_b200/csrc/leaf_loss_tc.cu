// Leaf-level semantic loss on the 5th-generation tensor cores (tcgen05 + TMEM + TMA bulk copies), sm_100a.
//
//     logits = Conv2d(S -> L, kernel 1)(sem);  loss = scale * sum_pixels CE(logits, leaf label);  dL/dsem
// (scripts/hierslam.py:975-984 / :1009-1016, MLP_func = torch.nn.Conv2d(num_semantic, num_semantic_class, 1), :1756;
// 550 classes for the ScanNet tree).  leaf_loss.cu holds the mma.sync generation of this loss; this file is the
// Blackwell-native pixel pass (logits, softmax statistics, loss, dL/dsem).  The weight gradient stays in leaf_loss.cu.
//
// One CTA owns 128 pixels = the 128 TMEM lanes: 8 softmax warps (two per quarter of the lanes, each working on half of
// the columns) + 1 issuer warp.  Per pixel tile:
//   X      the pixel rows [128 x K] (K = S + 1 "ones" channel for the bias, padded to 16) are written to TMEM with
//          tcgen05.st as hi / lo tf32 parts and stay there for all class chunks (A operand from TMEM, "TS" form);
//   Z_c    = X W_c^T for a chunk of 64 classes: tcgen05.mma kind::tf32, M = 128, N = 64, accumulator in TMEM, three
//          accumulating products X_hi W_hi + X_lo W_hi + X_hi W_lo (3xTF32: fp32-accurate);
//   P_c    = exp(Z_c - m) (online softmax with a lazily updated reference maximum m: the running sums are only rescaled
//          when the maximum grows by more than 8, which keeps the exponentials in range and the result exact), read with
//          tcgen05.ld -- one pixel row per thread, no shuffles -- and written back to TMEM as the A operand of
//   dX    += P_c W_c (M = 128, N = K, contraction over the chunk's classes), accumulated in TMEM across the chunks;
//   end    lse = m + log(l); loss; dL/dsem = scale (dX / l - W[label]) written planar, coalesced over the pixels.
// W reaches shared memory with cp.async.bulk (TMA, mbarrier complete_tx) from a workspace in which a small kernel has
// laid the weights out ONCE per call as UMMA canonical K-major operand tiles (hi and lo parts): B1_c = W_c [64 x K] for
// the logits, B2_c = W_c^T [K x 64] for dX.  Both rings are double buffered, so the copies of chunk c+1 run under the
// MMAs / softmax of chunk c; Z is double buffered in TMEM, so the logits MMAs of chunk c+1 run under the softmax of c.
// The [L, pixels] logits never exist in memory.
//
// Descriptor encodings (shared-memory matrix descriptor, instruction descriptor, canonical no-swizzle K-major layout)
// are the ones tools/micro/tcgen05_probe.cu validates against a host GEMM on the B200.
#include "hs_common.cuh"

namespace hs {
namespace leaftc {

constexpr int NC = 64;          // classes per chunk (N of the logits MMA, K of the dX MMA)
constexpr int KMAX = 80;        // padded channel count limit (S + 1 <= 80)
constexpr int TILE = 128;       // pixels per CTA tile = TMEM lanes

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol error must surface as a trap (a CUDA error on the host), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 28); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp (elect.sync): the form ptxas recognises as "exactly one active thread", so that the
// tcgen05.mma sequences below are emitted back to back instead of inside a per-instruction elect / branch waterfall
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 ::"l"((uint64_t)smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]   (TS form, kind::tf32)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};\n"
                 ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// shared-memory matrix descriptor, canonical K-major no-swizzle layout (8-row x 16-byte core matrices of 128 contiguous
// bytes; LBO = distance between core matrices adjacent along K, SBO = between 8-row groups)
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version (sm_100)
    return d;
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) {    // D fp32, A / B tf32, both K-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE >> 4) << 24);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// float index of element (row r, k) in a canonical K-major operand tile with K columns
__host__ __device__ inline uint32_t kmajor_index(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K >> 2) * 32 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3));
}

// ---- workspace layout ----------------------------------------------------------------------------------------------
// per chunk c: B1 hi [NC x K], B1 lo, B2 hi [K x NC], B2 lo  -- 4 NC K floats
__host__ __device__ inline size_t chunk_floats(int K) { return (size_t)4 * NC * K; }

// W [L x S] (+ bias) -> operand tiles.  One thread per (chunk, class j, channel s).
__global__ void __launch_bounds__(256) leaf_tc_prepare_kernel(const float* __restrict__ weight, const float* __restrict__ bias,
                                                              int S, int L, int K, int chunks, float* __restrict__ ws) {
    const int total = chunks * NC * K;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < total; e += gridDim.x * 256) {
        const int c = e / (NC * K), r = e - c * NC * K, j = r / K, s = r - j * K;
        const int cls = c * NC + j;
        float v = 0.f;
        if (cls < L) v = s < S ? weight[(size_t)cls * S + s] : (s == S && bias != nullptr ? bias[cls] : 0.f);
        const float hi = tf32_hi(v), lo = v - hi;
        float* base = ws + (size_t)c * chunk_floats(K);
        base[kmajor_index(j, s, K)] = hi;                                   // B1: rows = classes, K = channels
        base[(size_t)NC * K + kmajor_index(j, s, K)] = lo;
        base[(size_t)2 * NC * K + kmajor_index(s, j, NC)] = hi;             // B2: rows = channels, K = classes of the chunk
        base[(size_t)3 * NC * K + kmajor_index(s, j, NC)] = lo;
    }
}

// TMEM columns (fp32): X_hi [0,K) X_lo [K,2K) | Z0 Z1 (2 x NC) | P_hi P_lo (2 x NC) | dX [.., +K)   -> 2K + 4 NC + K <= 496
struct TmemMap {
    uint32_t x_hi, x_lo, z[2], p_hi, p_lo, dx;
};
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// four 16-column loads in flight, one wait
__device__ __forceinline__ void tmem_ld64(uint32_t addr, float (&v)[64]) {
    uint32_t r[64];
#pragma unroll
    for (int q = 0; q < 4; q++)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]), "=r"(r[16 * q + 4]),
              "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]), "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]),
              "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]), "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]),
              "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
            : "r"(addr + 16 * q));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; i++) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld8(uint32_t addr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, float (&v)[32]) {     // two 16-column loads in flight, one wait
    uint32_t r[32];
#pragma unroll
    for (int q = 0; q < 2; q++)
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
            : "=r"(r[16 * q + 0]), "=r"(r[16 * q + 1]), "=r"(r[16 * q + 2]), "=r"(r[16 * q + 3]), "=r"(r[16 * q + 4]),
              "=r"(r[16 * q + 5]), "=r"(r[16 * q + 6]), "=r"(r[16 * q + 7]), "=r"(r[16 * q + 8]), "=r"(r[16 * q + 9]),
              "=r"(r[16 * q + 10]), "=r"(r[16 * q + 11]), "=r"(r[16 * q + 12]), "=r"(r[16 * q + 13]),
              "=r"(r[16 * q + 14]), "=r"(r[16 * q + 15])
            : "r"(addr + 16 * q));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void pair_sync(int quarter) {      // the two warps that share a quarter of the TMEM lanes
    asm volatile("bar.sync %0, 64;\n" ::"r"(1 + quarter) : "memory");
}

constexpr int SOFTMAX_THREADS = 256;
constexpr int HALF = NC / 2;

// Warp roles: warps 0-7 = softmax / epilogue warps, warps 8 / 9 = issuers (one elected lane each issues the TMA bulk
// copies and tcgen05.mma of the logits / of dX; they never touch pixel data).  A warp can only address the TMEM lanes 32 (warp % 4) .. + 31, so the
// pixel rows of a tile are shared by TWO warps: thread (h, quarter, lane), h = warp / 4, owns pixel 32 quarter + lane
// together with its partner in the other half, and works on half h of everything that has a column index -- the channels
// of X and dX, the 32 classes [32 h, 32 h + 32) of every chunk.  The two exchange their chunk maxima / row sums through
// shared memory behind a 64-thread named barrier.  Roles meet only on mbarriers:
//   bar_x      (256 arrivals)  the tile's X rows are in TMEM                      softmax -> issuer
//   bar_z[2]   (commit)        logits of a chunk are complete in Z[slot]          issuer  -> softmax (and issuer: B1 slot free)
//   bar_zf[2]  (256 arrivals)  Z[slot] has been read into registers               softmax -> issuer
//   bar_p      (256 arrivals)  P of a chunk is in TMEM                            softmax -> issuer
//   bar_d      (commit)        dX MMAs of a chunk are complete                    issuer  -> softmax (and issuer: B2 slot free)
//   bar_t      (256 arrivals)  the tile's dX has been read out of TMEM            softmax -> issuer
//   bar_b1[2], bar_b2[2]       TMA full barriers (expect_tx)                      TMA     -> issuer
__global__ void __launch_bounds__(SOFTMAX_THREADS + 64, 1) leaf_ce_pixel_tc_kernel(
    const float* __restrict__ sem, const int* __restrict__ labels, const float* __restrict__ weight,
    const float* __restrict__ bias, const float* __restrict__ ws, int S, int L, int K, int chunks, size_t HW, float scale,
    float* __restrict__ loss, float* __restrict__ lse_out, float* __restrict__ grad_sem, int accumulate,
    long long* __restrict__ dbg) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sB1 = reinterpret_cast<float*>(smem_raw);                // [2][2][NC * K]  ring of (hi, lo)
    float* sB2 = sB1 + (size_t)4 * NC * K;                          // [2][2][K * NC]
    __shared__ __align__(8) uint64_t bar_b1[2], bar_b2[2], bar_z[2], bar_zf[2], bar_d, bar_p, bar_x, bar_t;
    __shared__ uint32_t s_tmem;
    __shared__ float s_loss[8];
    __shared__ float s_xchg[2][2][TILE];      // [parity][half][pixel]: chunk maxima / row sums / label logits of the partner
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int dbg_n = 0;
#define HS_STAMP() do { if (dbg != nullptr && blockIdx.x == 0 && tid == 0 && dbg_n < 60) dbg[dbg_n++] = clock64(); } while (0)
    const uint32_t tile_bytes = (uint32_t)(2 * NC * K * sizeof(float));        // hi + lo of one operand tile

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbar_init(&bar_b1[i], 1);
            mbar_init(&bar_b2[i], 1);
            mbar_init(&bar_z[i], 1);
            mbar_init(&bar_zf[i], SOFTMAX_THREADS);
        }
        mbar_init(&bar_d, 1);
        mbar_init(&bar_p, SOFTMAX_THREADS);
        mbar_init(&bar_x, SOFTMAX_THREADS);
        mbar_init(&bar_t, SOFTMAX_THREADS);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    TmemMap tm;
    tm.x_hi = tmem;
    tm.x_lo = tmem + K;
    tm.z[0] = tmem + 2 * K;
    tm.z[1] = tm.z[0] + NC;
    tm.p_hi = tm.z[1] + NC;
    tm.p_lo = tm.p_hi + NC;
    tm.dx = tm.p_lo + NC;
    const size_t n_tiles = (HW + TILE - 1) / TILE;

    if (warp == 8) {
        // ================================ logits issuer: B1 copies + Z MMAs ===========================================
        // (a single thread issues a tcgen05.mma in ~100 cycles -- uniform-datapath address arithmetic, elect, the
        // instruction itself -- while an M = 128, N = 64, K = 8 MMA is 32 cycles of tensor work, so the two MMA streams
        // of a chunk are issued by two different warps)
        if (elect_one()) {
            const uint32_t idesc_z = idesc_tf32(NC);
            const uint32_t b1_sbo = (uint32_t)(K >> 2) * 128;
            const int ksteps_z = K / 8;
            uint32_t g = 0, n_x = 0;              // chunks issued so far: ring slot = g & 1, phase = (g >> 1) & 1
            auto load_b1 = [&](uint32_t gg, int c) {
                const int slot = gg & 1;
                mbar_expect_tx(&bar_b1[slot], tile_bytes);
                tma_bulk_g2s(sB1 + (size_t)slot * 2 * NC * K, ws + (size_t)c * chunk_floats(K), tile_bytes, &bar_b1[slot]);
            };
            auto issue_z = [&](uint32_t gg) {     // logits of chunk gg into Z[gg & 1]:  X_hi W_hi + X_lo W_hi + X_hi W_lo
                const int slot = gg & 1;
                const uint32_t b_hi = smem_u32(sB1 + (size_t)slot * 2 * NC * K);
                const uint64_t d_hi = kmajor_desc(b_hi, 128, b1_sbo);
                const uint64_t d_lo = kmajor_desc(b_hi + (uint32_t)(NC * K * sizeof(float)), 128, b1_sbo);
                const uint32_t dz = tm.z[slot];
#pragma unroll 1
                for (int k = 0; k < ksteps_z; k++) mma_ts(dz, tm.x_hi + 8 * k, d_hi + (uint64_t)(16 * k), idesc_z, k > 0);
#pragma unroll 1
                for (int k = 0; k < ksteps_z; k++) mma_ts(dz, tm.x_lo + 8 * k, d_hi + (uint64_t)(16 * k), idesc_z, true);
#pragma unroll 1
                for (int k = 0; k < ksteps_z; k++) mma_ts(dz, tm.x_hi + 8 * k, d_lo + (uint64_t)(16 * k), idesc_z, true);
                tc_commit(&bar_z[slot]);
            };
            if (blockIdx.x < n_tiles) {           // the first tile's operand tiles; later tiles get theirs one tile ahead
                load_b1(0, 0);
                if (chunks > 1) load_b1(1, 1);
            }
            for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_x, n_x & 1);       // X rows of this tile are in TMEM
                n_x++;
                tc_fence_after();
                mbar_wait(&bar_b1[g & 1], (g >> 1) & 1);
                issue_z(g);
                for (int c = 0; c < chunks; c++, g++) {
                    if (c + 1 < chunks) {
                        // Z[(g+1)&1] was last used by chunk g-1: its registers have been read (bar_zf), and B1 has landed
                        if (g >= 1) mbar_wait(&bar_zf[(g + 1) & 1], ((g - 1) >> 1) & 1);
                        mbar_wait(&bar_b1[(g + 1) & 1], ((g + 1) >> 1) & 1);
                        tc_fence_after();
                        issue_z(g + 1);
                    }
                    mbar_wait(&bar_z[g & 1], (g >> 1) & 1);    // B1 slot of chunk g is free once its MMAs are complete
                    if (c + 2 < chunks) load_b1(g + 2, c + 2);
                }
                if (tile + gridDim.x < n_tiles) {  // every logits MMA of this tile is complete: the ring is free
                    load_b1(g, 0);
                    if (chunks > 1) load_b1(g + 1, 1);
                }
            }
        }
    } else if (warp == 9) {
        // ================================ dX issuer: B2 copies + dX MMAs ===============================================
        if (elect_one()) {
            const uint32_t idesc_dx = idesc_tf32(K);
            const uint32_t b2_sbo = (uint32_t)(NC >> 2) * 128;
            uint32_t g = 0, n_p = 0, n_d = 0, n_t = 0;
            auto load_b2 = [&](uint32_t gg, int c) {
                const int slot = gg & 1;
                mbar_expect_tx(&bar_b2[slot], tile_bytes);
                tma_bulk_g2s(sB2 + (size_t)slot * 2 * NC * K, ws + (size_t)c * chunk_floats(K) + (size_t)2 * NC * K,
                             tile_bytes, &bar_b2[slot]);
            };
            auto issue_dx = [&](uint32_t gg, bool first) {    // dX (+)= P_hi W_hi + P_lo W_hi + P_hi W_lo
                const int slot = gg & 1;
                const uint32_t b_hi = smem_u32(sB2 + (size_t)slot * 2 * NC * K);
                const uint64_t d_hi = kmajor_desc(b_hi, 128, b2_sbo);
                const uint64_t d_lo = kmajor_desc(b_hi + (uint32_t)(NC * K * sizeof(float)), 128, b2_sbo);
#pragma unroll
                for (int k = 0; k < NC / 8; k++) mma_ts(tm.dx, tm.p_hi + 8 * k, d_hi + (uint64_t)(16 * k), idesc_dx, !(first && k == 0));
#pragma unroll
                for (int k = 0; k < NC / 8; k++) mma_ts(tm.dx, tm.p_lo + 8 * k, d_hi + (uint64_t)(16 * k), idesc_dx, true);
#pragma unroll
                for (int k = 0; k < NC / 8; k++) mma_ts(tm.dx, tm.p_hi + 8 * k, d_lo + (uint64_t)(16 * k), idesc_dx, true);
                tc_commit(&bar_d);
            };
            if (blockIdx.x < n_tiles) load_b2(0, 0);
            for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                for (int c = 0; c < chunks; c++, g++) {
                    if (c >= 1) {                 // B2 slot of the previous chunk is free once ITS dX MMAs are complete
                        mbar_wait(&bar_d, n_d & 1);
                        n_d++;
                    }
                    if (c + 1 < chunks) load_b2(g + 1, c + 1);
                    if (c == 0 && tile != blockIdx.x) {       // the previous tile's dX has been read out
                        mbar_wait(&bar_t, n_t & 1);
                        n_t++;
                    }
                    mbar_wait(&bar_p, n_p & 1);               // P of this chunk is in TMEM
                    n_p++;
                    mbar_wait(&bar_b2[g & 1], (g >> 1) & 1);
                    tc_fence_after();
                    issue_dx(g, c == 0);
                }
                mbar_wait(&bar_d, n_d & 1);                   // last chunk's dX: both B2 slots are free
                n_d++;
                if (tile + gridDim.x < n_tiles) load_b2(g, 0);
            }
        }
    } else {
        // ================================ softmax / epilogue warps ===================================================
        const int h = warp >> 2, quarter = warp & 3;
        const int pix = quarter * 32 + lane;                            // pixel of the tile = TMEM lane
        const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;       // a warp addresses the lanes 32 (warp % 4) ..
        const int KH = K >> 1;                                          // channels of this half: [h KH, h KH + KH), KH % 8 == 0
        const int ch0 = h * KH;
        float loss_acc = 0.f;
        uint32_t g = 0, n_d = 0, xp = 0;
        for (size_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const size_t px = tile * TILE + pix;
            const bool in = px < HW;
            HS_STAMP();   // tile start
            // ---- this half's channels of the pixel row -> TMEM (hi / lo).  For a fixed channel the 32 lanes of a warp
            // read 32 consecutive pixels; all loads are issued before the first TMEM store (one memory latency).
            {
                float v[KMAX / 2];
#pragma unroll
                for (int q = 0; q < KMAX / 2; q++) {
                    const int s = ch0 + q;
                    v[q] = 0.f;
                    if (q < KH && in && s < S) v[q] = __ldg(sem + (size_t)s * HW + px);
                    else if (q < KH && in && s == S) v[q] = 1.f;
                }
#pragma unroll
                for (int q8 = 0; q8 < KMAX / 2; q8 += 8) {
                    if (q8 < KH) {
                        float hi[8], lo[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) {
                            hi[q] = tf32_hi(v[q8 + q]);
                            lo[q] = v[q8 + q] - hi[q];
                        }
                        tmem_st8(tm.x_hi + lane_off + ch0 + q8, hi);
                        tmem_st8(tm.x_lo + lane_off + ch0 + q8, lo);
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&bar_x);
            HS_STAMP();   // X stored
            const int lab_raw = in ? __ldg(labels + px) : -1;
            const bool use = lab_raw >= 0 && lab_raw < L;
            const int y = use ? lab_raw : 0;
            float m_ref = -1.0e30f, l_sum = 0.f;
            for (int c = 0; c < chunks; c++, g++) {
                const int slot = g & 1;
                mbar_wait(&bar_z[slot], (g >> 1) & 1);        // logits of chunk c are complete
                tc_fence_after();
                HS_STAMP();   // Z ready
                float z[HALF];
                tmem_ld32(tm.z[slot] + lane_off + h * HALF, z);
                tc_fence_before();
                mbar_arrive(&bar_zf[slot]);                   // Z[slot] may be overwritten by the logits of chunk c + 2
                const int cls0 = c * NC + h * HALF;
                float cm[4] = {-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
                if (cls0 + HALF <= L) {                       // warp-uniform: no class mask needed
#pragma unroll
                    for (int i = 0; i < HALF; i++) cm[i & 3] = fmaxf(cm[i & 3], z[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < HALF; i++) {
                        if (cls0 + i >= L) z[i] = -3.0e38f;   // exp() of it is 0 below
                        cm[i & 3] = fmaxf(cm[i & 3], z[i]);
                    }
                }
                const float my_max = fmaxf(fmaxf(cm[0], cm[1]), fmaxf(cm[2], cm[3]));
                s_xchg[xp][h][pix] = my_max;
                pair_sync(quarter);
                const float cmax = fmaxf(my_max, s_xchg[xp][h ^ 1][pix]);
                xp ^= 1;
                // lazily updated reference maximum: rescale only when the maximum grows by more than 8 (exp(8) = 2981:
                // far from fp32 overflow; the result stays exact because every term uses the same reference).  Both warps
                // of a pixel take the same decision from the same numbers.
                float rescale = 1.f;
                if (cmax > m_ref + 8.f) {
                    rescale = __expf(m_ref - cmax);           // 0 for the first chunk (m_ref = -1e30)
                    m_ref = cmax;
                }
#pragma unroll
                for (int i = 0; i < HALF; i++) z[i] = __expf(z[i] - m_ref);
                HS_STAMP();   // exponentials done
                // the previous chunk's dX MMAs read P (and wrote dX): complete before P is overwritten / dX rescaled
                if (c >= 1) {
                    mbar_wait(&bar_d, n_d & 1);
                    n_d++;
                    tc_fence_after();
                }
                l_sum *= rescale;
                // rare after the first chunk.  tcgen05.ld / st are warp-collective (.sync.aligned): the decision is made
                // per warp, lanes whose reference did not move multiply by 1; each half rescales its own dX columns
                if (__any_sync(0xffffffffu, c >= 1 && rescale != 1.f)) {
#pragma unroll 1
                    for (int q = 0; q < KH; q += 8) {
                        float d[8];
                        tmem_ld8(tm.dx + lane_off + ch0 + q, d);
#pragma unroll
                        for (int i = 0; i < 8; i++) d[i] *= rescale;
                        tmem_st8(tm.dx + lane_off + ch0 + q, d);
                    }
                }
                float ls[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int q = 0; q < HALF; q += 8) {
                    float h8[8], l8[8];
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        ls[i & 3] += z[q + i];
                        h8[i] = tf32_hi(z[q + i]);
                        l8[i] = z[q + i] - h8[i];
                    }
                    tmem_st8(tm.p_hi + lane_off + h * HALF + q, h8);
                    tmem_st8(tm.p_lo + lane_off + h * HALF + q, l8);
                }
                l_sum += (ls[0] + ls[1]) + (ls[2] + ls[3]);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&bar_p);
                HS_STAMP();   // P stored
            }
            // ---- tile epilogue: row sums of both halves, then dX
            s_xchg[xp][h][pix] = l_sum;
            pair_sync(quarter);
            const float l_tot = l_sum + s_xchg[xp][h ^ 1][pix];
            xp ^= 1;
            mbar_wait(&bar_d, n_d & 1);
            n_d++;
            tc_fence_after();
            HS_STAMP();   // last dX done
            const float sc = use ? scale : 0.f;
            const float inv = sc / l_tot;
            const float* wy = weight + (size_t)y * S;
            float zy_part = 0.f;                              // this half's share of the label's logit  <x, W[y]> (+ bias)
#pragma unroll 1
            for (int q = 0; q < KH; q += 8) {
                float d[8], wv[8], xv[8], old[8];
                tmem_ld8(tm.dx + lane_off + ch0 + q, d);
                if (q + 8 >= KH) {                            // last read of this tile's dX by this thread
                    tc_fence_before();
                    mbar_arrive(&bar_t);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int s = ch0 + q + i;
                    wv[i] = (s < S) ? __ldg(wy + s) : ((s == S && bias != nullptr) ? __ldg(bias + y) : 0.f);
                    xv[i] = (s < S && in) ? __ldg(sem + (size_t)s * HW + px) : (s == S ? 1.f : 0.f);
                    old[i] = (accumulate && s < S && in) ? grad_sem[(size_t)s * HW + px] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int s = ch0 + q + i;
                    zy_part = fmaf(xv[i], wv[i], zy_part);
                    if (s < S && in) grad_sem[(size_t)s * HW + px] = old[i] + (d[i] * inv - sc * wv[i]);
                }
            }
            s_xchg[xp][h][pix] = zy_part;
            pair_sync(quarter);
            if (h == 0 && in) {
                const float zy = zy_part + s_xchg[xp][1][pix];
                const float lse = m_ref + __logf(l_tot);
                lse_out[px] = lse;
                loss_acc += sc * (lse - zy);
            }
            xp ^= 1;
            HS_STAMP();   // outputs written
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
        if (lane == 0) s_loss[warp] = loss_acc;
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) atomicAdd(loss, s_loss[0] + s_loss[1] + s_loss[2] + s_loss[3]);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "n"(512));
#undef HS_STAMP
}

}  // namespace leaftc

static long long* g_leaf_tc_dbg = nullptr;   // debugging aid: clock64 stamps of CTA 0 (hs_leaf_tc_debug)
void leaf_tc_set_debug(long long* p) { g_leaf_tc_dbg = p; }

size_t leaf_tc_workspace_bytes(int S, int L) {
    const int K = (S + 1 + 15) & ~15;
    const int chunks = (L + leaftc::NC - 1) / leaftc::NC;
    return (size_t)chunks * leaftc::chunk_floats(K) * sizeof(float);
}

int launch_leaf_cross_entropy_tc(const float* sem, const int* labels, const float* weight, const float* bias, int S, int L,
                                 size_t HW, float scale, float* loss, float* lse, float* grad_sem, int accumulate,
                                 float* workspace, size_t workspace_bytes, cudaStream_t stream) {
    using namespace leaftc;
    if (HW == 0 || L <= 0) return 0;
    if (S < 1 || S + 1 > KMAX) {
        set_error("leaf cross-entropy (tcgen05): 1 <= S <= %d semantic channels supported, got %d", KMAX - 1, S);
        return 1;
    }
    const int K = (S + 1 + 15) & ~15;
    const int chunks = (L + NC - 1) / NC;
    if (workspace == nullptr || workspace_bytes < leaf_tc_workspace_bytes(S, L)) {
        set_error("leaf cross-entropy (tcgen05): workspace of %zu bytes required", leaf_tc_workspace_bytes(S, L));
        return 1;
    }
    if ((reinterpret_cast<uintptr_t>(workspace) & 127) != 0) {
        set_error("leaf cross-entropy (tcgen05): workspace must be 128-byte aligned");
        return 1;
    }
    const int total = chunks * NC * K;
    leaf_tc_prepare_kernel<<<(total + 255) / 256, 256, 0, stream>>>(weight, bias, S, L, K, chunks, workspace);
    HS_LAUNCH_OK(stream, false);
    const size_t smem = (size_t)8 * NC * K * sizeof(float);      // two rings of two (hi, lo) tiles
    auto k = leaf_ce_pixel_tc_kernel;
    HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int sms = 148;
    {
        int dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    const size_t n_tiles = (HW + TILE - 1) / TILE;
    const int grid = (int)(n_tiles < (size_t)sms ? n_tiles : (size_t)sms);
    k<<<grid, SOFTMAX_THREADS + 64, smem, stream>>>(sem, labels, weight, bias, workspace, S, L, K, chunks, HW, scale, loss, lse, grad_sem,
                                    accumulate, g_leaf_tc_dbg);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs
