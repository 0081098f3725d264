// Forward alpha compositing of RGB + depth + median depth + silhouette + S semantic channels.
// Reference behaviour: cuda_rasterizer/forward.cu:400-538 (renderCUDA_SEM) and :261-398 (renderCUDA).
//
// What is kept from the reference: one 16x16 tile per CTA, one pixel per thread, front-to-back traversal of
// the tile's depth-sorted list, and the exact alpha / transmittance arithmetic (same expression shapes, accurate
// expf, IEEE division) so that the skip / stop decisions -- and therefore n_contrib and final_T -- agree bit for
// bit.  What is new: every per-Gaussian operand of the inner loop, including the (4+S)-float feature row
// [r g b depth s0..s(S-1)], is staged once per batch in shared memory and consumed with 128-bit broadcast
// loads (the reference re-reads colour and semantic rows from global memory for every contributing pixel,
// forward.cu:505-508), batches double-buffered with cp.async; a conservative footprint box per Gaussian gives an
// 8-bit region mask, and a warp (one pixel region of the tile: an 8x4 block, hs_common.cuh) iterates only over the set bits of its ballot of 32 entries; the
// inner loop has no divergent branch (non-contributing lanes blend with weight 0, votes are warp-uniform); channels
// are blended in pairs with the packed FFMA2; and the strips that actually blended an entry are recorded
// (strip_hits) so that the backward visits exactly those (strip, entry) pairs.
#include "hs_common.cuh"
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
    static constexpr int SCH = (S % 4 == 0) ? 4 : (S % 2 == 0) ? 2 : 1;   // floats per cp.async of a semantic row
    // double-buffered staging: conic+opacity, feature rows, centres; ring of 3 id arrays; strip masks
    static constexpr size_t SMEM = (size_t)BATCH * (2 * sizeof(float4) + 2 * FS * sizeof(float) + 2 * sizeof(float2) +
                                                    3 * sizeof(int) + sizeof(uint32_t)) + 2 * 8 * (BATCH / 32) * sizeof(uint32_t);
};

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
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float2* __restrict__ means2D, const float* __restrict__ colors, const float* __restrict__ depths,
    const float* __restrict__ semantics, const float4* __restrict__ conic_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_median, float* __restrict__ out_semantic, float* __restrict__ out_opacity,
    float* __restrict__ out_mask, uint8_t* __restrict__ strip_hits, int flags) {
    using Cfg = FwdCfg<S>;
    constexpr int B = Cfg::BATCH;
    constexpr int FS = Cfg::FS;
    constexpr int SCH = Cfg::SCH;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_co2 = reinterpret_cast<float4*>(smem_raw);             // [2][B]
    float* s_feat2 = reinterpret_cast<float*>(s_co2 + 2 * B);        // [2][B][FS]
    float2* s_xy2 = reinterpret_cast<float2*>(s_feat2 + 2 * B * FS); // [2][B]
    int* s_id3 = reinterpret_cast<int*>(s_xy2 + 2 * B);              // [3][B]
    uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_id3 + 3 * B);   // [B] bit w: Gaussian may touch warp w's pixel region
    uint32_t* s_hit2 = s_mask + B;   // [2][B/32][8] bit j%32 of word (j/32, w): warp w blended Gaussian j of the batch

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

    // Asynchronous staging: the per-Gaussian records of batch r+1 are gathered with cp.async while batch r is
    // blended; the Gaussian ids (first level of the gather) run two batches ahead.
    auto fetch_ids = [&](int r) {
        const int n = min(B, total - r * B);
        if (r < rounds && tid < n) __pipeline_memcpy_async(s_id3 + (r % 3) * B + tid, point_list + range.x + r * B + tid, 4);
    };
    auto gather_batch = [&](int r) {   // ids of batch r are already in s_id3[r % 3]
        const int n = min(B, total - r * B);
        const int* ids = s_id3 + (r % 3) * B;
        const int bo = (r & 1) * B;
        if (tid < n) {
            const int id = ids[tid];
            __pipeline_memcpy_async(s_xy2 + bo + tid, means2D + id, 8);
            __pipeline_memcpy_async(s_co2 + bo + tid, conic_opacity + id, 16);
            float* f = s_feat2 + (size_t)(bo + tid) * FS;
            __pipeline_memcpy_async(f, colors + 3 * (size_t)id, 4);
            __pipeline_memcpy_async(f + 1, colors + 3 * (size_t)id + 1, 4);
            __pipeline_memcpy_async(f + 2, colors + 3 * (size_t)id + 2, 4);
            __pipeline_memcpy_async(f + 3, depths + id, 4);
        }
        if (S > 0) {
            constexpr int PER = (S > 0 ? S : 1) / SCH;   // chunks per row
            if (SCH > 1 && !(flags & HS_FLAG_SEM_UNALIGNED)) {
                for (int e = tid; e < n * PER; e += 256) {
                    const int j = e / PER, c = (e - j * PER) * SCH;
                    __pipeline_memcpy_async(s_feat2 + (size_t)(bo + j) * FS + 4 + c,
                                            semantics + (size_t)ids[j] * S + c, 4 * SCH);
                }
            } else {
                for (int e = tid; e < n * S; e += 256) {
                    const int j = e / (S > 0 ? S : 1), c = e - j * S;
                    __pipeline_memcpy_async(s_feat2 + (size_t)(bo + j) * FS + 4 + c,
                                            semantics + (size_t)ids[j] * S + c, 4);
                }
            }
        }
    };
    fetch_ids(0);
    fetch_ids(1);
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    gather_batch(0);
    __pipeline_commit();

    for (int i = 0; i <= rounds; i++) {
        __pipeline_wait_prior(0);
        // batch i has landed and the ids of batch i+1 are present; also the block-wide early-out vote
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
        if (i + 1 < rounds) gather_batch(i + 1);
        fetch_ids(i + 2);
        __pipeline_commit();
        const float2* s_xy = s_xy2 + (i & 1) * B;
        const float4* s_co = s_co2 + (i & 1) * B;
        const float* s_feat = s_feat2 + (size_t)(i & 1) * B * FS;
        if (tid < nb) {
            // which of the 8 warp regions (HS_REGION_W x HS_REGION_H pixels) can this Gaussian reach with alpha >= 1/255 ?
            uint32_t mk = 0xffu;
            if (cull) {
                const float4 bx = footprint_box(s_xy[tid], s_co[tid]);
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
                    const float2 xy = s_xy[jj[u]];
                    const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                    const float4 con_o = s_co[jj[u]];
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
                    const ulonglong2* f4 = reinterpret_cast<const ulonglong2*>(s_feat + jj[u] * FS);
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

template <int S>
static int launch_fwd_t(const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                        const float* colors, const float* semantics, float* out_color, float* out_semantic,
                        float* out_depth, float* out_median, float* out_opacity, float* out_mask, int flags,
                        cudaStream_t stream, bool debug) {
    dim3 grid(cam.grid_x, cam.grid_y, 1);
    const size_t smem = FwdCfg<S>::SMEM;
    if (reinterpret_cast<uintptr_t>(semantics) % (4 * FwdCfg<S>::SCH) != 0) flags |= HS_FLAG_SEM_UNALIGNED;
    prof_begin(ST_BLEND_FWD, stream);
    if (S == 0 && out_mask != nullptr) {
        auto k = blend_forward_kernel<S, true>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask, b.strip_hits, flags);
    } else {
        auto k = blend_forward_kernel<S, false>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask, b.strip_hits, flags);
    }
    prof_end(ST_BLEND_FWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

int launch_blend_forward(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                         const float* colors, const float* semantics, float* out_color, float* out_semantic,
                         float* out_depth, float* out_median, float* out_opacity, float* out_mask, int flags,
                         cudaStream_t stream, bool debug) {
#define HS_FWD_CASE(SV)                                                                                      \
    case SV:                                                                                                 \
        return launch_fwd_t<SV>(cam, g, b, img, colors, semantics, out_color, out_semantic, out_depth,       \
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
