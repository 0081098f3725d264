// Forward alpha compositing of RGB + depth + median depth + silhouette + S semantic channels.
// Reference behaviour: cuda_rasterizer/forward.cu:400-538 (renderCUDA_SEM) and :261-398 (renderCUDA).
//
// What is kept from the reference: one 16x16 tile per CTA, one pixel per thread, front-to-back traversal of
// the tile's depth-sorted list, and the exact alpha / transmittance arithmetic (same expression shapes, accurate
// expf, IEEE division) so that the skip / stop decisions -- and therefore n_contrib and final_T -- agree bit for
// bit.  What is new: every per-Gaussian operand of the inner loop, including the (4+S)-float feature row
// [r g b depth s0..s(S-1)], is staged once per batch in shared memory and consumed with 128-bit broadcast
// loads (the reference re-reads colour and semantic rows from global memory for every contributing pixel,
// forward.cu:505-508); a conservative footprint box per Gaussian lets a whole warp (a 16x2 pixel strip) skip a
// Gaussian with one shared load and four compares; and culled / finished tiles leave after one vote.
#include "hs_common.cuh"

namespace hs {

template <int S>
struct FwdCfg {
    static constexpr int F = 4 + S;                 // feature row: r g b depth s...
    static constexpr int FS = (F + 3) & ~3;         // row stride in floats (16-B aligned rows)
    static constexpr int BATCH = (S <= 32) ? 128 : 64;
    static constexpr size_t SMEM = (size_t)BATCH * (sizeof(float2) + 2 * sizeof(float4) + sizeof(int) + FS * sizeof(float));
};

template <int S, bool MASK>
__global__ void __launch_bounds__(256) blend_forward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float2* __restrict__ means2D, const float* __restrict__ colors, const float* __restrict__ depths,
    const float* __restrict__ semantics, const float4* __restrict__ conic_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_median, float* __restrict__ out_semantic, float* __restrict__ out_opacity,
    float* __restrict__ out_mask, int flags) {
    using Cfg = FwdCfg<S>;
    constexpr int B = Cfg::BATCH;
    constexpr int FS = Cfg::FS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_co = reinterpret_cast<float4*>(smem_raw);              // [B]
    float4* s_box = s_co + B;                                        // [B]
    float* s_feat = reinterpret_cast<float*>(s_box + B);             // [B][FS]
    float2* s_xy = reinterpret_cast<float2*>(s_feat + B * FS);       // [B]
    int* s_id = reinterpret_cast<int*>(s_xy + B);                    // [B]

    const int tid = threadIdx.x;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    const uint32_t px = tile_x * HS_TILE_X + (tid & 15);
    const uint32_t py = tile_y * HS_TILE_Y + (tid >> 4);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;
    // the 16x2 pixel strip of this warp, for the footprint test
    const float wx0 = (float)(tile_x * HS_TILE_X), wx1 = wx0 + 15.f;
    const float wy0 = (float)(tile_y * HS_TILE_Y + ((tid >> 5) << 1)), wy1 = wy0 + 1.f;
    const bool cull = !(flags & HS_FLAG_NO_CULL);

    const uint2 range = ranges[tile_y * grid_x + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + B - 1) / B;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float acc[4 + S];
#pragma unroll
    for (int k = 0; k < 4 + S; k++) acc[k] = 0.f;
    float median_D = 15.0f;
    float M = 0.f;

    for (int i = 0; i < rounds; i++) {
        if (__syncthreads_count(done) == 256) break;
        const int nb = min(B, total - i * B);
        if (tid < nb) {
            const int id = point_list[range.x + i * B + tid];
            s_id[tid] = id;
            const float2 xy = means2D[id];
            const float4 co = conic_opacity[id];
            s_xy[tid] = xy;
            s_co[tid] = co;
            s_box[tid] = footprint_box(xy, co);
            float* f = s_feat + tid * FS;
            f[0] = __ldg(colors + 3 * (size_t)id);
            f[1] = __ldg(colors + 3 * (size_t)id + 1);
            f[2] = __ldg(colors + 3 * (size_t)id + 2);
            f[3] = depths[id];
        }
        if (S > 0) {
            __syncthreads();
            for (int e = tid; e < nb * S; e += 256) {
                const int j = e / (S > 0 ? S : 1), c = e - j * S;
                s_feat[j * FS + 4 + c] = __ldg(semantics + (size_t)s_id[j] * S + c);
            }
        }
        __syncthreads();

        for (int j = 0; !done && j < nb; j++) {
            if (cull) {
                const float4 bx = s_box[j];
                if (bx.x > wx1 || bx.y < wx0 || bx.z > wy1 || bx.w < wy0) continue;  // warp-uniform
            }
            const float2 xy = s_xy[j];
            const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
            const float4 con_o = s_co[j];
            const float power = gauss_power(d, con_o);
            if (power > 0.0f) continue;
            const float alpha = min(0.99f, con_o.w * exp(power));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = T * (1 - alpha);
            if (test_T < 0.0001f) {
                done = true;
                continue;
            }
            const float w = alpha * T;
            const float4* f4 = reinterpret_cast<const float4*>(s_feat + j * FS);
#pragma unroll
            for (int q = 0; q < (4 + S + 3) / 4; q++) {
                const float4 v = f4[q];
                acc[4 * q] = fmaf(v.x, w, acc[4 * q]);
                if (4 * q + 1 < 4 + S) acc[4 * q + 1] = fmaf(v.y, w, acc[4 * q + 1]);
                if (4 * q + 2 < 4 + S) acc[4 * q + 2] = fmaf(v.z, w, acc[4 * q + 2]);
                if (4 * q + 3 < 4 + S) acc[4 * q + 3] = fmaf(v.w, w, acc[4 * q + 3]);
            }
            if (MASK) M += w;
            if (T > 0.5f && test_T < 0.5) median_D = f4[0].w;
            T = test_T;
            last_contributor = i * B + j + 1;
        }
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
    prof_begin(ST_BLEND_FWD, stream);
    if (S == 0 && out_mask != nullptr) {
        auto k = blend_forward_kernel<S, true>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask, flags);
    } else {
        auto k = blend_forward_kernel<S, false>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask, flags);
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
        HS_FWD_CASE(74)
        HS_FWD_CASE(102)
        default:
            set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,74,102)", S);
            return 3;
    }
#undef HS_FWD_CASE
}

}  // namespace hs
