// Forward alpha compositing of RGB + depth + median depth + silhouette + S semantic channels.
// Reference behaviour: cuda_rasterizer/forward.cu:400-538 (renderCUDA_SEM) and :261-398 (renderCUDA).
//
// What is kept from the reference: one 16x16 tile per CTA, one pixel per thread, front-to-back traversal of
// the tile's depth-sorted list, and the exact alpha / transmittance arithmetic (same expression shapes, accurate
// expf, IEEE division) so that the skip / stop decisions -- and therefore n_contrib and final_T -- agree bit for
// bit.  What is new: every per-Gaussian operand of the inner loop, including the (4+S)-float feature row
// [r g b depth s0..s(S-1)], is staged once per batch in shared memory and consumed with 128-bit broadcast
// loads (the reference re-reads colour and semantic rows from global memory for every contributing pixel,
// forward.cu:505-508).  The staging is done by the TMA unit: a small kernel packs one record per visible Gaussian
// ([conic opacity | xy | r g b depth | sem...], packed_row_floats(S) floats), and per batch of B list entries ONE warp
// issues B / 4 cp.async.bulk.tensor.2d ... tile::gather4 instructions -- each gathers the records of four Gaussians of
// the tile's sorted list straight into the shared-memory batch buffer and signals an mbarrier (complete_tx) -- instead of
// every thread issuing 6 + S / 2 cp.async with its own address arithmetic; batches are double buffered, the Gaussian
// ids (contiguous in the sorted list) run two batches ahead.  A conservative footprint box per Gaussian gives an
// 8-bit region mask, and a warp (one pixel region of the tile: an 8x4 block, hs_common.cuh) iterates only over the set bits of its ballot of 32 entries; the
// inner loop has no divergent branch (non-contributing lanes blend with weight 0, votes are warp-uniform); channels
// are blended in pairs with the packed FFMA2; and the strips that actually blended an entry are recorded
// (strip_hits) so that the backward visits exactly those (strip, entry) pairs.
#include "hs_common.cuh"
#include <cuda.h>
#include <cuda_pipeline.h>

#ifndef HS_FWD_U
#define HS_FWD_U 2      // Gaussians whose alpha is evaluated together (independent dependency chains per lane)
#endif
#ifndef HS_FWD_OCC
#define HS_FWD_OCC 3    // CTAs per SM the narrow instantiations (S <= 26) are compiled for
#endif
#ifndef HS_FWD_OCC_WIDE
#define HS_FWD_OCC_WIDE 1   // ... and S = 74 (2 CTAs per SM = 128 registers was measured slower: 427 vs 416 us at c5)
#endif

namespace hs {

template <int S>
struct FwdCfg {
    static constexpr int F = 4 + S;                 // feature row: r g b depth s...
    static constexpr int FS = (F + 3) & ~3;         // row stride in floats (16-B aligned rows)
    static constexpr int BATCH = (S <= 32) ? 128 : 64;
    static constexpr int RS = (8 + FS + 7) & ~7;    // packed record: [co(4) xy(2) pad(2) | feature row | pad] == packed_row_floats(S)
    // double-buffered record batches; ring of 3 id arrays; strip masks; hit words
    static constexpr size_t SMEM = (size_t)BATCH * (2 * RS * sizeof(float) + 3 * sizeof(int) + sizeof(uint32_t)) +
                                   2 * 8 * (BATCH / 32) * sizeof(uint32_t);
};

// one record per visible Gaussian: thread <-> (Gaussian, 16-byte chunk)
template <int S>
__global__ void __launch_bounds__(256) pack_rows_kernel(int P, const int* __restrict__ radii,
                                                        const float4* __restrict__ conic_opacity,
                                                        const float2* __restrict__ means2D, const float* __restrict__ depths,
                                                        const float* __restrict__ colors, const float* __restrict__ semantics,
                                                        float4* __restrict__ rows) {
    constexpr int RS4 = FwdCfg<S>::RS / 4;
    const size_t e = (size_t)blockIdx.x * 256 + threadIdx.x;
    const size_t g = e / RS4;
    const int c = (int)(e - g * RS4);
    if (g >= (size_t)P || radii[g] <= 0) return;          // culled Gaussians never appear in a tile list
    float4 v = {0.f, 0.f, 0.f, 0.f};
    if (c == 0) v = conic_opacity[g];
    else if (c == 1) {
        const float2 xy = means2D[g];
        v = {xy.x, xy.y, 0.f, 0.f};
    } else if (c == 2) v = {colors[3 * g], colors[3 * g + 1], colors[3 * g + 2], depths[g]};
    else {
        const int s0 = 4 * (c - 3);
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; q++) t[q] = (s0 + q < S) ? semantics[g * (size_t)(S > 0 ? S : 1) + s0 + q] : 0.f;
        v = {t[0], t[1], t[2], t[3]};
    }
    rows[e] = v;
}

__device__ __forceinline__ uint32_t fwd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// Bounded mbarrier wait (a protocol error becomes a trap, never a hung GPU).
__device__ __forceinline__ void fwd_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = fwd_smem_u32(bar);
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

// Packed FP32 FMA (sm_100 FFMA2): two IEEE fused multiply-adds per issue slot, bit-identical to two FFMAs.
__device__ __forceinline__ void ffma2(unsigned long long& acc, const unsigned long long a, const unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(const float lo, const float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(const unsigned long long v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}

template <int S, bool MASK>
__global__ void __launch_bounds__(256, (S <= 26 ? HS_FWD_OCC : S <= 74 ? HS_FWD_OCC_WIDE : 1)) blend_forward_kernel(
    const __grid_constant__ CUtensorMap rows_map, const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
    int W, int H, int grid_x, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_median, float* __restrict__ out_semantic, float* __restrict__ out_opacity,
    float* __restrict__ out_mask, uint8_t* __restrict__ strip_hits, int flags) {
    using Cfg = FwdCfg<S>;
    constexpr int B = Cfg::BATCH;
    constexpr int FS = Cfg::FS;
    constexpr int RS = Cfg::RS;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_rows2 = reinterpret_cast<float*>(smem_raw);             // [2][B][RS] packed records of the batch
    int* s_id3 = reinterpret_cast<int*>(s_rows2 + 2 * B * RS);       // [3][B]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_id3 + 3 * B);   // [B] bit w: Gaussian may touch warp w's pixel region
    uint32_t* s_hit2 = s_mask + B;   // [2][B/32][8] bit j%32 of word (j/32, w): warp w blended Gaussian j of the batch
    __shared__ __align__(8) uint64_t s_bar[2];                       // TMA full barriers of the two record buffers

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    const uint32_t px = tile_x * HS_TILE_X + HS_PX_X(warp, tid & 31);
    const uint32_t py = tile_y * HS_TILE_Y + HS_PX_Y(warp, tid & 31);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;
    const float tx0 = (float)(tile_x * HS_TILE_X), ty0 = (float)(tile_y * HS_TILE_Y);
    const bool cull = !(flags & HS_FLAG_NO_CULL);

    const uint2 range = ranges[tile_y * grid_x + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + B - 1) / B;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    static_assert((4 + S) % 2 == 0, "feature rows are blended two channels per FFMA2");
    constexpr int F2 = (4 + S) / 2;
    unsigned long long acc2[F2];   // packed channel pairs: (r g) (b depth) (s0 s1) ...
#pragma unroll
    for (int k = 0; k < F2; k++) acc2[k] = 0ull;
    float median_D = 15.0f;
    float M = 0.f;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(fwd_smem_u32(&s_bar[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(fwd_smem_u32(&s_bar[1])));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // Staging: the ids of a batch (contiguous in the sorted list) arrive with cp.async two batches ahead; the records of
    // batch r+1 are gathered by the TMA unit while batch r is blended: warp 0 issues one tile::gather4 per four entries
    // (lane l: entries 4l .. 4l+3; a ragged tail repeats its last id, the extra rows are never read).
    auto fetch_ids = [&](int r) {
        const int n = min(B, total - r * B);
        if (r < rounds && tid < n) __pipeline_memcpy_async(s_id3 + (r % 3) * B + tid, point_list + range.x + r * B + tid, 4);
    };
    auto gather_batch = [&](int r) {   // ids of batch r are already in s_id3[r % 3]; called by warp 0 only
        const int n = min(B, total - r * B);
        const int groups = (n + 3) >> 2;
        const int* ids = s_id3 + (r % 3) * B;
        uint64_t* bar = &s_bar[r & 1];
        if (tid == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                         ::"r"(fwd_smem_u32(bar)), "r"((uint32_t)(groups * 4 * RS * sizeof(float))) : "memory");
        __syncwarp();
        for (int l = tid; l < groups; l += 32) {
            const int e = 4 * l;
            const int i0 = ids[e], i1 = ids[min(e + 1, n - 1)], i2 = ids[min(e + 2, n - 1)], i3 = ids[min(e + 3, n - 1)];
            asm volatile(
                "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                ::"r"(fwd_smem_u32(s_rows2 + (size_t)((r & 1) * B + e) * RS)), "l"(&rows_map), "r"(0), "r"(i0), "r"(i1), "r"(i2),
                "r"(i3), "r"(fwd_smem_u32(bar)) : "memory");
        }
    };
    fetch_ids(0);
    fetch_ids(1);
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();                  // ids of batches 0 and 1 present, mbarriers initialised
    if (warp == 0 && rounds > 0) gather_batch(0);

    for (int i = 0; i <= rounds; i++) {
        __pipeline_wait_prior(0);
        if (i < rounds) fwd_mbar_wait(&s_bar[i & 1], (uint32_t)((i >> 1) & 1));      // the records of batch i have landed
        // ... and the ids of batch i+1 are present; also the block-wide early-out vote
        const bool all_done = __syncthreads_count(done) == 256;
        if (i > 0) {
            // which strips blended which entries of batch i-1: saved for the backward, which then visits exactly the
            // (strip, Gaussian) pairs that contributed
            const int np = min(B, total - (i - 1) * B);
            if (tid < np) {
                const uint32_t* h = s_hit2 + (((i - 1) & 1) * (B / 32) + (tid >> 5)) * 8;
                uint32_t m = 0;
#pragma unroll
                for (int w8 = 0; w8 < 8; w8++) m |= ((h[w8] >> (tid & 31)) & 1u) << w8;
                strip_hits[range.x + (i - 1) * B + tid] = (uint8_t)m;
            }
        }
        if (all_done || i == rounds) break;
        uint32_t* s_hit = s_hit2 + (i & 1) * (B / 32) * 8;
        const int nb = min(B, total - i * B);
        if (warp == 0 && i + 1 < rounds) gather_batch(i + 1);
        fetch_ids(i + 2);
        __pipeline_commit();
        const float* s_rows = s_rows2 + (size_t)(i & 1) * B * RS;
        if (tid < nb) {
            // which of the 8 warp regions (HS_REGION_W x HS_REGION_H pixels) can this Gaussian reach with alpha >= 1/255 ?
            uint32_t mk = 0xffu;
            if (cull) {
                const float* row = s_rows + (size_t)tid * RS;
                const float4 bx = footprint_box(*reinterpret_cast<const float2*>(row + 4), *reinterpret_cast<const float4*>(row));
                mk = 0;
                if (!(bx.x > tx0 + 15.f || bx.y < tx0)) {
#pragma unroll
                    for (int w8 = 0; w8 < 8; w8++) {
                        const float x0 = tx0 + (float)HS_REGION_X0(w8), y0 = ty0 + (float)HS_REGION_Y0(w8);
                        if (!(bx.x > x0 + (float)(HS_REGION_W - 1) || bx.y < x0) &&
                            !(bx.z > y0 + (float)(HS_REGION_H - 1) || bx.w < y0))
                            mk |= 1u << w8;
                    }
                }
            }
            s_mask[tid] = mk;
        }
        __syncthreads();

        // The inner loop is written without divergent branches: all 32 lanes of a warp stay together, a lane that
        // does not take a Gaussian (outside its footprint, alpha < 1/255, already saturated) blends it with weight 0
        // (x + 0 * f == x bit for bit for finite f), and the only branches are warp-uniform votes.  Gaussians that
        // survive the strip cull are taken U at a time so that the U alpha evaluations (independent of T) overlap.
        constexpr int U = HS_FWD_U;
#pragma unroll 1
        for (int k0 = 0; k0 < nb; k0 += 32) {
            const int jl = k0 + (tid & 31);
            uint32_t bits = __ballot_sync(0xffffffffu, jl < nb && ((s_mask[jl] >> warp) & 1));
            if (__all_sync(0xffffffffu, done)) bits = 0;
            uint32_t hit_bits = 0;   // warp-uniform: entries of this group of 32 that some lane blended
#pragma unroll 1
            while (bits) {
                int jj[U];
                float al[U];
                bool ok[U];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const bool have = bits != 0;
                    jj[u] = have ? k0 + __ffs(bits) - 1 : jj[0];
                    bits &= bits - 1;
                    const float* row = s_rows + jj[u] * RS;
                    const float2 xy = *reinterpret_cast<const float2*>(row + 4);
                    const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                    const float4 con_o = *reinterpret_cast<const float4*>(row);
                    const float power = gauss_power(d, con_o);
                    al[u] = min(0.99f, con_o.w * exp(power));
                    ok[u] = have && !(power > 0.0f) && !(al[u] < 1.0f / 255.0f);
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    bool v = ok[u] && !done;
                    const float test_T = T * (1 - al[u]);
                    if (v && test_T < 0.0001f) {
                        done = true;
                        v = false;
                    }
                    if (!__any_sync(0xffffffffu, v)) continue;   // warp-uniform
                    hit_bits |= 1u << (jj[u] - k0);
                    const float w = v ? al[u] * T : 0.f;
                    const ulonglong2* f4 = reinterpret_cast<const ulonglong2*>(s_rows + jj[u] * RS + 8);
                    float depth_j;
#pragma unroll
                    for (int q = 0; q < (4 + S + 3) / 4; q++) {
                        const ulonglong2 f = f4[q];
                        if (q == 0) depth_j = unpack2(f.y).y;
                        ffma2(acc2[2 * q], f.x, pack2(w, w));
                        if (2 * q + 1 < F2) ffma2(acc2[2 * q + 1], f.y, pack2(w, w));
                    }
                    if (MASK) M += w;
                    if (v && T > 0.5f && test_T < 0.5) median_D = depth_j;
                    T = v ? test_T : T;
                    last_contributor = v ? i * B + jj[u] + 1 : last_contributor;
                }
            }
            if ((tid & 31) == 0) s_hit[(k0 >> 5) * 8 + warp] = hit_bits;
        }
    }
    __pipeline_wait_prior(0);

    float acc[4 + S];
#pragma unroll
    for (int k = 0; k < F2; k++) {
        const float2 a = unpack2(acc2[k]);
        acc[2 * k] = a.x;
        acc[2 * k + 1] = a.y;
    }
    if (inside) {
        const size_t HW = (size_t)H * W;
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        out_color[pix_id] = acc[0];
        out_color[HW + pix_id] = acc[1];
        out_color[2 * HW + pix_id] = acc[2];
        out_depth[pix_id] = acc[3];
        out_median[pix_id] = median_D;
        out_opacity[pix_id] = 1 - T;
        if (MASK) out_mask[pix_id] = M;
#pragma unroll
        for (int ch = 0; ch < S; ch++) out_semantic[ch * HW + pix_id] = acc[4 + ch];
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

template <int S>
static int launch_fwd_t(int P, const int* radii, const Camera& cam, const GeomView& g, const BinningView& b,
                        const ImageView& img, const float* colors, const float* semantics, float* out_color,
                        float* out_semantic, float* out_depth, float* out_median, float* out_opacity, float* out_mask, int flags,
                        cudaStream_t stream, bool debug) {
    using Cfg = FwdCfg<S>;
    dim3 grid(cam.grid_x, cam.grid_y, 1);
    const size_t smem = Cfg::SMEM;
    static_assert(Cfg::RS % 8 == 0, "four packed records must be a multiple of 128 bytes");
    // 2-D tensor map over the packed records [P rows][RS floats]; box = one row: tile::gather4 fetches four rows per instruction
    CUtensorMap rows_map;
    {
        EncodeTiledFn enc = encode_tiled_fn();
        if (enc == nullptr) {
            set_error("cuTensorMapEncodeTiled is not available from this driver (TMA staging of the forward blend)");
            return 2;
        }
        const cuuint64_t dims[2] = {(cuuint64_t)Cfg::RS, (cuuint64_t)(P > 0 ? P : 1)};
        const cuuint64_t strides[1] = {(cuuint64_t)Cfg::RS * sizeof(float)};
        const cuuint32_t box[2] = {(cuuint32_t)Cfg::RS, 1}, estr[2] = {1, 1};
        const CUresult rc = enc(&rows_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g.rows, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) {
            set_error("cuTensorMapEncodeTiled failed (%d) for the packed records [%d][%d]", (int)rc, P, Cfg::RS);
            return 2;
        }
    }
    prof_begin(ST_BLEND_FWD, stream);
    if (S == 0 && out_mask != nullptr) {
        auto k = blend_forward_kernel<S, true>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(rows_map, img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, img.final_T, img.n_contrib,
                                       out_color, out_depth, out_median, out_semantic, out_opacity, out_mask, b.strip_hits, flags);
    } else {
        auto k = blend_forward_kernel<S, false>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(rows_map, img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, img.final_T, img.n_contrib,
                                       out_color, out_depth, out_median, out_semantic, out_opacity, out_mask, b.strip_hits, flags);
    }
    prof_end(ST_BLEND_FWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

template <int S>
static int launch_pack_t(int P, const int* radii, const GeomView& g, const float* colors, const float* semantics,
                         cudaStream_t stream) {
    if (P <= 0) return 0;
    const size_t chunks = (size_t)P * (FwdCfg<S>::RS / 4);
    pack_rows_kernel<S><<<(unsigned)((chunks + 255) / 256), 256, 0, stream>>>(
        P, radii, g.conic_opacity, g.means2D, g.depths, colors, semantics, reinterpret_cast<float4*>(g.rows));
    HS_LAUNCH_OK(stream, false);
    return 0;
}

// packs the per-Gaussian records the forward blend gathers with TMA (needs the per-Gaussian pass, not the binning: the
// caller may run it on a side stream next to the scatter / sort kernels)
int launch_pack_rows(int P, const int* radii, int S, const GeomView& g, const float* colors, const float* semantics,
                     cudaStream_t stream) {
    switch (S) {
        case 0: return launch_pack_t<0>(P, radii, g, colors, semantics, stream);
        case 16: return launch_pack_t<16>(P, radii, g, colors, semantics, stream);
        case 26: return launch_pack_t<26>(P, radii, g, colors, semantics, stream);
        case 32: return launch_pack_t<32>(P, radii, g, colors, semantics, stream);
        case 48: return launch_pack_t<48>(P, radii, g, colors, semantics, stream);
        case 64: return launch_pack_t<64>(P, radii, g, colors, semantics, stream);
        case 74: return launch_pack_t<74>(P, radii, g, colors, semantics, stream);
        case 102: return launch_pack_t<102>(P, radii, g, colors, semantics, stream);
        default:
            set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,32,48,64,74,102)", S);
            return 3;
    }
}

int launch_blend_forward(int P, const int* radii, int S, const Camera& cam, const GeomView& g, const BinningView& b,
                         const ImageView& img, const float* colors, const float* semantics, float* out_color, float* out_semantic,
                         float* out_depth, float* out_median, float* out_opacity, float* out_mask, int flags,
                         cudaStream_t stream, bool debug) {
#define HS_FWD_CASE(SV)                                                                                          \
    case SV:                                                                                                     \
        return launch_fwd_t<SV>(P, radii, cam, g, b, img, colors, semantics, out_color, out_semantic, out_depth, \
                                out_median, out_opacity, out_mask, flags, stream, debug);
    switch (S) {
        HS_FWD_CASE(0)
        HS_FWD_CASE(16)
        HS_FWD_CASE(26)
        HS_FWD_CASE(32)
        HS_FWD_CASE(48)
        HS_FWD_CASE(64)
        HS_FWD_CASE(74)
        HS_FWD_CASE(102)
        default:
            set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,32,48,64,74,102)", S);
            return 3;
    }
#undef HS_FWD_CASE
}

}  // namespace hs
