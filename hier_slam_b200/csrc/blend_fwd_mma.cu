// Forward alpha compositing, tensor-core version (default for S <= 74; reference behaviour:
// cuda_rasterizer/forward.cu:400-538 / :261-398).
//
// The blend  out[pix][c] = sum_j w[pix,j] * f[j][c]  (w = alpha * T, c = r g b depth s0..s(S-1))  is a per-tile
// contraction over the tile's Gaussian list.  Each warp owns a 16x2 pixel strip and walks the list 8 Gaussians at a time:
//   (a) alpha of the 8 Gaussians, branch-free (eight independent dependency chains per lane);
//   (b) the sequential transmittance recurrence with the reference's exact arithmetic and stop rules -- this is what
//       makes n_contrib / final_T bit-identical -- producing the 8 blend weights of the lane's pixel;
//   (c) the weights are transposed through a per-warp shared-memory tile and
//       D[32 pixels][8 channels] += A[32 x 8 Gaussians] * B[8 Gaussians x 8 channels] runs on the tensor cores
//       (mma.sync m16n8k8 TF32, 3xTF32 split: a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulate), with the feature
//       rows staged once per batch in shared memory in a bank-conflict-free layout.
// A group whose 8 Gaussians touch none of the warp's live pixels costs one vote.
#include "hs_common.cuh"

namespace hs {

template <int S>
struct FwdMmaCfg {
    static constexpr int F = 4 + S;                     // r g b depth s...
    static constexpr int NT = (F + 7) / 8;              // channel n-tiles
    static constexpr int FS = 8 * NT + ((8 - (8 * NT) % 32 + 32) % 32);  // feature row stride == 8 (mod 32)
    static constexpr int BATCH = (S <= 32) ? 128 : 64;
    static constexpr int WT = 40;                       // row stride of the per-warp weight tile [8][32] (== 8 mod 32)
    static constexpr size_t SMEM = (size_t)BATCH * (sizeof(float2) + sizeof(float4) + sizeof(int) + FS * sizeof(float)) +
                                   (size_t)8 * 8 * WT * sizeof(float);
};

__device__ __forceinline__ void fmma_tf32(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                          const uint32_t a3, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ftf32_lo(const float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

template <int S, bool MASK>
__global__ void __launch_bounds__(256, (S <= 32 ? 2 : 1)) blend_forward_mma_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float2* __restrict__ means2D, const float* __restrict__ colors, const float* __restrict__ depths,
    const float* __restrict__ semantics, const float4* __restrict__ conic_opacity, float* __restrict__ final_T,
    uint32_t* __restrict__ n_contrib, float* __restrict__ out_color, float* __restrict__ out_depth,
    float* __restrict__ out_median, float* __restrict__ out_semantic, float* __restrict__ out_opacity,
    float* __restrict__ out_mask) {
    using Cfg = FwdMmaCfg<S>;
    constexpr int B = Cfg::BATCH, FS = Cfg::FS, NT = Cfg::NT, WT = Cfg::WT, F = Cfg::F;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_co = reinterpret_cast<float4*>(smem_raw);              // [B]
    float* s_feat = reinterpret_cast<float*>(s_co + B);              // [B][FS]
    float* s_wt = s_feat + B * FS;                                   // [8 warps][8][WT]
    float2* s_xy = reinterpret_cast<float2*>(s_wt + 8 * 8 * WT);     // [B]
    int* s_id = reinterpret_cast<int*>(s_xy + B);                    // [B]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qk = lane & 3, qn = lane >> 2;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    const uint32_t px = tile_x * HS_TILE_X + (lane & 15);
    const uint32_t py = tile_y * HS_TILE_Y + 2 * warp + (lane >> 4);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    bool done = !inside;

    const uint2 range = ranges[tile_y * grid_x + tile_x];
    const int total = (int)(range.y - range.x);
    const int rounds = (total + B - 1) / B;

    float T = 1.0f;
    uint32_t last_contributor = 0;
    float median_D = 15.0f;
    float M = 0.f;
    float acc[2][NT][4];   // D fragments: pixel 16 mt + lane/4 (+8), channel 8 nt + 2 (lane%4) (+1)
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++)
#pragma unroll
            for (int r = 0; r < 4; r++) acc[mt][nt][r] = 0.f;
    float* wt = s_wt + warp * 8 * WT;

    for (int i = 0; i < rounds; i++) {
        if (__syncthreads_count(done) == 256) break;
        const int nb = min(B, total - i * B);
        if (tid < nb) {
            const int id = point_list[range.x + i * B + tid];
            s_id[tid] = id;
            s_xy[tid] = means2D[id];
            s_co[tid] = conic_opacity[id];
            float* f = s_feat + tid * FS;
            f[0] = __ldg(colors + 3 * (size_t)id);
            f[1] = __ldg(colors + 3 * (size_t)id + 1);
            f[2] = __ldg(colors + 3 * (size_t)id + 2);
            f[3] = depths[id];
        }
        if (tid < B) {   // everything the MMA reads must be finite: padding channels, and the rows that complete
            float* f = s_feat + tid * FS;   // the last group of 8 (their weights are zero, but 0 * NaN is NaN)
#pragma unroll
            for (int c = F; c < 8 * NT; c++) f[c] = 0.f;
            if (tid >= nb && tid < ((nb + 7) & ~7)) {
#pragma unroll
                for (int c = 0; c < F; c++) f[c] = 0.f;
            }
        }
        if (S > 0) {
            __syncthreads();
            for (int e = tid; e < nb * S; e += 256) {
                const int j = e / (S > 0 ? S : 1), c = e - j * S;
                s_feat[j * FS + 4 + c] = __ldg(semantics + (size_t)s_id[j] * S + c);
            }
        }
        __syncthreads();

#pragma unroll 1
        for (int g0 = 0; g0 < nb; g0 += 8) {
            // (a) alpha of 8 Gaussians, branch-free
            float og[8];
            uint32_t abits = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int j = min(g0 + u, nb - 1);
                const float2 xy = s_xy[j];
                const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                const float4 con_o = s_co[j];
                const float power = gauss_power(d, con_o);
                const float o_g = con_o.w * exp(power);
                const bool valid = (g0 + u < nb) && !(power > 0.0f) && !(min(0.99f, o_g) < 1.0f / 255.0f);
                og[u] = o_g;
                abits |= (valid ? 1u : 0u) << u;
            }
            if (done) abits = 0;
            if (__reduce_or_sync(0xffffffffu, abits) == 0) continue;   // warp-uniform
            // (b) transmittance recurrence, reference arithmetic: test_T = T (1 - alpha); stop when test_T < 1e-4
            float w[8];
            uint32_t cbits = 0;
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const bool act = ((abits >> u) & 1) && !done;
                const float alpha = min(0.99f, og[u]);
                const float test_T = T * (1 - alpha);
                const bool stop = act && (test_T < 0.0001f);
                const bool contrib = act && !stop;
                done = done || stop;
                w[u] = contrib ? alpha * T : 0.f;
                if (contrib && T > 0.5f && test_T < 0.5) median_D = s_feat[(g0 + u) * FS + 3];
                T = contrib ? test_T : T;
                last_contributor = contrib ? (uint32_t)(i * B + g0 + u + 1) : last_contributor;
                cbits |= (contrib ? 1u : 0u) << u;
                if (MASK) M += w[u];
            }
            if (__reduce_or_sync(0xffffffffu, cbits) == 0) continue;   // warp-uniform
            // (c) transpose the weights through shared memory and blend on the tensor cores
#pragma unroll
            for (int u = 0; u < 8; u++) wt[u * WT + lane] = w[u];
            __syncwarp();
            uint32_t ah[2][4], al[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                // A fragment: rows = pixels 16 mt + qn (+8), cols = Gaussians qk (+4)
                const float a0 = wt[qk * WT + 16 * mt + qn], a1 = wt[qk * WT + 16 * mt + qn + 8];
                const float a2 = wt[(qk + 4) * WT + 16 * mt + qn], a3 = wt[(qk + 4) * WT + 16 * mt + qn + 8];
                ah[mt][0] = __float_as_uint(a0); ah[mt][1] = __float_as_uint(a1);
                ah[mt][2] = __float_as_uint(a2); ah[mt][3] = __float_as_uint(a3);
                al[mt][0] = __float_as_uint(ftf32_lo(a0)); al[mt][1] = __float_as_uint(ftf32_lo(a1));
                al[mt][2] = __float_as_uint(ftf32_lo(a2)); al[mt][3] = __float_as_uint(ftf32_lo(a3));
            }
            const float* fb = s_feat + (g0 + qk) * FS + qn;   // B fragment: rows = Gaussians qk (+4), col = channel 8 nt + qn
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
                const float b0 = fb[8 * nt], b1 = fb[4 * FS + 8 * nt];
                const uint32_t bh0 = __float_as_uint(b0), bh1 = __float_as_uint(b1);
                const uint32_t bl0 = __float_as_uint(ftf32_lo(b0)), bl1 = __float_as_uint(ftf32_lo(b1));
#pragma unroll
                for (int mt = 0; mt < 2; mt++) {
                    fmma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bh0, bh1);
                    fmma_tf32(acc[mt][nt], al[mt][0], al[mt][1], al[mt][2], al[mt][3], bh0, bh1);
                    fmma_tf32(acc[mt][nt], ah[mt][0], ah[mt][1], ah[mt][2], ah[mt][3], bl0, bl1);
                }
            }
            __syncwarp();   // wt is rewritten by the next group
        }
    }

    const size_t HW = (size_t)H * W;
    if (inside) {
        final_T[pix_id] = T;
        n_contrib[pix_id] = last_contributor;
        out_median[pix_id] = median_D;
        out_opacity[pix_id] = 1 - T;
        if (MASK) out_mask[pix_id] = M;
    }
    // D fragments -> channel-planar images
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const uint32_t y = tile_y * HS_TILE_Y + 2 * warp + mt;
        if (y >= (uint32_t)H) continue;
#pragma unroll
        for (int half = 0; half < 2; half++) {
            const uint32_t x = tile_x * HS_TILE_X + qn + 8 * half;
            if (x >= (uint32_t)W) continue;
            const size_t pid = (size_t)W * y + x;
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int c = 8 * nt + 2 * qk + e;
                    const float v = acc[mt][nt][2 * half + e];
                    if (c < 3) out_color[(size_t)c * HW + pid] = v;
                    else if (c == 3) out_depth[pid] = v;
                    else if (c < F) out_semantic[(size_t)(c - 4) * HW + pid] = v;
                }
        }
    }
}

template <int S>
static int launch_fwd_mma_t(const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                            const float* colors, const float* semantics, float* out_color, float* out_semantic,
                            float* out_depth, float* out_median, float* out_opacity, float* out_mask,
                            cudaStream_t stream, bool debug) {
    dim3 grid(cam.grid_x, cam.grid_y, 1);
    const size_t smem = FwdMmaCfg<S>::SMEM;
    prof_begin(ST_BLEND_FWD, stream);
    if (S == 0 && out_mask != nullptr) {
        auto k = blend_forward_mma_kernel<S, true>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask);
    } else {
        auto k = blend_forward_mma_kernel<S, false>;
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, g.means2D, colors, g.depths,
                                       semantics, g.conic_opacity, img.final_T, img.n_contrib, out_color, out_depth,
                                       out_median, out_semantic, out_opacity, out_mask);
    }
    prof_end(ST_BLEND_FWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

int launch_blend_forward_mma(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                             const float* colors, const float* semantics, float* out_color, float* out_semantic,
                             float* out_depth, float* out_median, float* out_opacity, float* out_mask,
                             cudaStream_t stream, bool debug) {
#define HS_FWDM_CASE(SV)                                                                                        \
    case SV:                                                                                                    \
        return launch_fwd_mma_t<SV>(cam, g, b, img, colors, semantics, out_color, out_semantic, out_depth,      \
                                    out_median, out_opacity, out_mask, stream, debug);
    switch (S) {
        HS_FWDM_CASE(0)
        HS_FWDM_CASE(16)
        HS_FWDM_CASE(26)
        HS_FWDM_CASE(74)
        default:
            set_error("tensor-core blend forward: S=%d is not instantiated (built: 0,16,26,74)", S);
            return 3;
    }
#undef HS_FWDM_CASE
}

}  // namespace hs
