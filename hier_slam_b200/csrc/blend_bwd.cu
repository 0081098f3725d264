// Backward of the alpha compositing: per-pixel upstream gradients -> per-Gaussian gradients of colour,
// semantics, depth, opacity, 2D mean and conic.
// Reference behaviour: cuda_rasterizer/backward.cu:669-899 (renderCUDA_SEM) and :472-666 (renderCUDA),
// including quirks Q1-Q4 of SURVEY.md section 8a.
//
// The reference issues 11+S global float atomics per (pixel, Gaussian) contribution.  Here:
//   * the per-channel "colour behind" recurrences (accum_rec[ch], last_color[ch], ...; 2(5+S) registers) are
//     collapsed into ONE scalar recurrence on q_j = sum_f f_j * dL/dout_f -- the recurrence is linear, so
//     sum_f (f_j - accum_f) * dL_f == q_j - accum_q.  This removes the register pressure that makes the
//     reference spill at S >= 74;
//   * the K = S+10 per-Gaussian partial sums of a warp are combined with a butterfly (transpose) reduction:
//     K-1 shuffles instead of 5K, leaving lane l with the warp total of value l;
//   * warp totals are accumulated in a shared-memory tile [batch][K] with shared-memory reductions, and each
//     (tile, Gaussian) pair is flushed ONCE to global memory with coalesced atomics (lanes = channels);
//   * a warp skips a Gaussian with one vote when none of its 32 pixels is touched, tiles start at the deepest
//     contributor of any of their pixels, and upstream-gradient planes that autograd did not materialise
//     (None) are never read.
#include "hs_common.cuh"

namespace hs {

template <int S>
struct BwdCfg {
    static constexpr int K = S + 10;                  // sem[S] rgb[3] depth opacity mean2D[2] conic[3]
    static constexpr int NF = 5;                      // staged feature row: r g b depth (pad) ; +S in exact mode
    static constexpr int BATCH = 64;
};

// ---- butterfly reduction -------------------------------------------------------------------------------
// Reduce N (power of two <= 32) per-lane values across the 32 lanes; afterwards v[0] of lane l holds the warp
// total of value (l mod N).
template <int N>
__device__ __forceinline__ void warp_reduce_transpose(float* v, const int lane) {
#pragma unroll
    for (int o = 16; o >= N; o >>= 1) {
#pragma unroll
        for (int i = 0; i < N; i++) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    }
#pragma unroll
    for (int n = N; n > 1; n >>= 1) {
        const int o = n >> 1;
        const bool upper = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; i++) {
            const float send = upper ? v[i] : v[i + n / 2];
            const float keep = upper ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
}

template <int REM>
struct Pow2Ceil {
    static constexpr int value = REM <= 1 ? 1 : REM <= 2 ? 2 : REM <= 4 ? 4 : REM <= 8 ? 8 : REM <= 16 ? 16 : 32;
};

// Reduce K values; add the totals into acc[0..K) (shared memory).
template <int K>
__device__ __forceinline__ void warp_reduce_to_smem(float* v, float* acc, const int lane) {
    constexpr int FULL = K / 32;
    constexpr int REM = K - FULL * 32;
#pragma unroll
    for (int c = 0; c < FULL; c++) {
        warp_reduce_transpose<32>(v + 32 * c, lane);
        atomicAdd(acc + 32 * c + lane, v[32 * c]);
    }
    if (REM > 0) {
        constexpr int N = Pow2Ceil<REM>::value;
        float t[N];
#pragma unroll
        for (int i = 0; i < N; i++) t[i] = (i < REM) ? v[32 * FULL + i] : 0.f;
        warp_reduce_transpose<N>(t, lane);
        if (lane < REM) atomicAdd(acc + 32 * FULL + lane, t[0]);
    }
}

template <int S, bool EXACT>
__global__ void __launch_bounds__(256) blend_backward_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float* __restrict__ bg_color, const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
    const float* __restrict__ colors, const float* __restrict__ semantics, const float* __restrict__ depths,
    const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dpixels_sem, const float* __restrict__ dL_dpixel_depths,
    const float* __restrict__ dL_dpixel_medians, const float* __restrict__ dL_dpixel_opacitys,
    float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic2D, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolors, float* __restrict__ dL_dsemantics, float* __restrict__ dL_ddepths) {
    using Cfg = BwdCfg<S>;
    constexpr int B = Cfg::BATCH;
    constexpr int K = Cfg::K;
    constexpr int NSEM = EXACT ? S : 0;
    constexpr int FS = (4 + NSEM + 3) & ~3;  // staged row: r g b depth [sem...]
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_co = reinterpret_cast<float4*>(smem_raw);        // [B]
    float* s_feat = reinterpret_cast<float*>(s_co + B);        // [B][FS]
    float* s_acc = s_feat + B * FS;                            // [B][K]
    float2* s_xy = reinterpret_cast<float2*>(s_acc + B * K);   // [B]
    int* s_id = reinterpret_cast<int*>(s_xy + B);              // [B]
    int* s_touched = s_id + B;                                 // [B]
    __shared__ int s_maxc;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y;
    const uint32_t px = tile_x * HS_TILE_X + (tid & 15);
    const uint32_t py = tile_y * HS_TILE_Y + (tid >> 4);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint2 range = ranges[tile_y * grid_x + tile_x];
    const size_t HW = (size_t)H * W;

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

    // tile-wide deepest contributor: entries behind it are touched by no pixel
    if (tid == 0) s_maxc = 0;
    __syncthreads();
    {
        int m = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0) atomicMax(&s_maxc, m);
    }
    __syncthreads();
    const int total = min(s_maxc, (int)(range.y - range.x));
    if (total <= 0) return;

    float dL_rgb[3] = {0.f, 0.f, 0.f};
    float dL_depth = 0.f, dL_median = 0.f, dL_op = 0.f;
    float dL_sem[S > 0 ? S : 1];
#pragma unroll
    for (int k = 0; k < (S > 0 ? S : 1); k++) dL_sem[k] = 0.f;
    if (inside) {
        if (dL_dpixels) {
#pragma unroll
            for (int i = 0; i < 3; i++) dL_rgb[i] = dL_dpixels[i * HW + pix_id];
        }
        if (dL_dpixel_depths) dL_depth = dL_dpixel_depths[pix_id];
        if (dL_dpixel_medians) dL_median = dL_dpixel_medians[pix_id];
        if (dL_dpixel_opacitys) dL_op = dL_dpixel_opacitys[pix_id];
        if (S > 0 && dL_dpixels_sem) {
#pragma unroll
            for (int i = 0; i < S; i++) dL_sem[i] = dL_dpixels_sem[i * HW + pix_id];
        }
    }
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_rgb[i];

    float last_alpha = 0.f, last_q = 0.f, accum_q = 0.f;
    const float ddelx_dx = 0.5 * W;
    const float ddely_dy = 0.5 * H;

    const int rounds = (total + B - 1) / B;
    for (int i = 0; i < rounds; i++) {
        __syncthreads();  // previous batch fully flushed
        const int nb = min(B, total - i * B);
        // entry j of this batch is list position (total - 1 - i*B - j): back to front
        if (tid < nb) {
            const int id = point_list[range.x + (total - 1 - i * B - tid)];
            s_id[tid] = id;
            s_xy[tid] = means2D[id];
            s_co[tid] = conic_opacity[id];
            s_touched[tid] = 0;
            float* f = s_feat + tid * FS;
            f[0] = __ldg(colors + 3 * (size_t)id);
            f[1] = __ldg(colors + 3 * (size_t)id + 1);
            f[2] = __ldg(colors + 3 * (size_t)id + 2);
            f[3] = depths[id];
        }
        for (int e = tid; e < nb * K; e += 256) s_acc[e] = 0.f;
        if (NSEM > 0) {
            __syncthreads();
            for (int e = tid; e < nb * NSEM; e += 256) {
                const int j = e / (NSEM > 0 ? NSEM : 1), c = e - j * NSEM;
                s_feat[j * FS + 4 + c] = __ldg(semantics + (size_t)s_id[j] * S + c);
            }
        }
        __syncthreads();

        for (int j = 0; j < nb; j++) {
            const int gi = total - 1 - i * B - j;  // position in the tile list
            const float2 xy = s_xy[j];
            const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
            const float4 con_o = s_co[j];
            const float power = gauss_power(d, con_o);
            const float G = exp(power);
            const float alpha = min(0.99f, con_o.w * G);
            const bool active = (gi < last_contributor) && !(power > 0.0f) && !(alpha < 1.0f / 255.0f);
            if (!__any_sync(0xffffffffu, active)) continue;

            float v[K];
            if (active) {
                const float test_T = T / (1.f - alpha);
                const float w = alpha * test_T;
                const float* f = s_feat + j * FS;
                // q_j = sum_f f_j dL_f over colour, depth, silhouette (constant 1) [+ semantics in exact mode]
                float q = f[0] * dL_rgb[0] + f[1] * dL_rgb[1] + f[2] * dL_rgb[2];
                q = fmaf(f[3], dL_depth, q);
                q += dL_op;
                if (NSEM > 0) {
#pragma unroll
                    for (int c = 0; c < NSEM; c++) q = fmaf(f[4 + c], dL_sem[c], q);
                }
                accum_q = last_alpha * last_q + (1.f - last_alpha) * accum_q;
                last_q = q;
                float dL_dalpha = (q - accum_q) * test_T;
#pragma unroll
                for (int c = 0; c < S; c++) v[c] = w * dL_sem[c];
                v[S + 0] = w * dL_rgb[0];
                v[S + 1] = w * dL_rgb[1];
                v[S + 2] = w * dL_rgb[2];
                float gd = w * dL_depth;
                if (test_T > 0.5f && T < 0.5) gd += dL_median;  // the Gaussian that crossed T = 0.5 (quirk Q4)
                v[S + 3] = gd;
                T = test_T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                const float dL_dG = con_o.w * dL_dalpha;
                const float gdx = G * d.x;
                const float gdy = G * d.y;
                const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
                const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
                v[S + 4] = fmaf(G, dL_dalpha, w * dL_op);  // opacity (both terms; quirk Q2)
                v[S + 5] = dL_dG * dG_ddelx * ddelx_dx;
                v[S + 6] = dL_dG * dG_ddely * ddely_dy;
                v[S + 7] = -0.5f * gdx * d.x * dL_dG;
                v[S + 8] = -0.5f * gdx * d.y * dL_dG;
                v[S + 9] = -0.5f * gdy * d.y * dL_dG;
            } else {
#pragma unroll
                for (int c = 0; c < K; c++) v[c] = 0.f;
            }
            warp_reduce_to_smem<K>(v, s_acc + j * K, lane);
            if (lane == 0) s_touched[j] = 1;
        }
        __syncthreads();

        // flush: one warp per touched Gaussian, lanes = consecutive channels -> coalesced atomics
        for (int j = tid >> 5; j < nb; j += 8) {
            if (!s_touched[j]) continue;
            const size_t id = (size_t)s_id[j];
            const float* a = s_acc + j * K;
            for (int c = lane; c < S; c += 32) atomicAdd(dL_dsemantics + id * S + c, a[c]);
            if (lane < 3) atomicAdd(dL_dcolors + id * 3 + lane, a[S + lane]);
            else if (lane == 3) atomicAdd(dL_ddepths + id, a[S + 3]);
            else if (lane == 4) atomicAdd(dL_dopacity + id, a[S + 4]);
            else if (lane < 7) atomicAdd(dL_dmean2D + id * 3 + (lane - 5), a[S + lane]);
            else if (lane < 9) atomicAdd(dL_dconic2D + id * 4 + (lane - 7), a[S + lane]);
            else if (lane == 9) atomicAdd(dL_dconic2D + id * 4 + 3, a[S + 9]);
        }
    }
}

template <int S, bool EXACT>
static int launch_bwd_t(const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                        const float* bg, const float* colors, const float* semantics, const float* dL_color,
                        const float* dL_sem, const float* dL_depth, const float* dL_median, const float* dL_opacity,
                        float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors,
                        float* dL_dsemantics, float* dL_ddepths, cudaStream_t stream, bool debug) {
    constexpr int B = BwdCfg<S>::BATCH;
    constexpr int K = BwdCfg<S>::K;
    constexpr int NSEM = EXACT ? S : 0;
    constexpr int FS = (4 + NSEM + 3) & ~3;
    const size_t smem = (size_t)B * (sizeof(float4) + FS * sizeof(float) + K * sizeof(float) + sizeof(float2) + 2 * sizeof(int));
    auto k = blend_backward_kernel<S, EXACT>;
    HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cam.grid_x, cam.grid_y, 1);
    prof_begin(ST_BLEND_BWD, stream);
    k<<<grid, 256, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, bg, g.means2D, g.conic_opacity,
                                   colors, semantics, g.depths, img.final_T, img.n_contrib, dL_color, dL_sem, dL_depth,
                                   dL_median, dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics,
                                   dL_ddepths);
    prof_end(ST_BLEND_BWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

int launch_blend_backward(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                          const float* bg, const float* colors, const float* semantics, const float* dL_color,
                          const float* dL_sem, const float* dL_depth, const float* dL_median,
                          const float* dL_opacity, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                          float* dL_dcolors, float* dL_dsemantics, float* dL_ddepths, int flags,
                          cudaStream_t stream, bool debug) {
    // No upstream semantic gradient (tracking: the pose loss uses colour and depth only, scripts/hierslam.py:780-796):
    // dL/dsemantics is identically zero and the semantic channels cannot reach dL/dalpha in either Q1 mode, so the
    // non-semantic instantiation computes exactly the same gradients; dL_dsemantics keeps the caller's zeros.
    if (dL_sem == nullptr) S = 0;
    const bool exact = (flags & HS_FLAG_SEM_ALPHA_EXACT) != 0 && S > 0;
    if (!exact && !(flags & HS_FLAG_BWD_SHUFFLE) && (S <= 74 || S == 102))  // S = 102: two 51-channel passes
        return launch_blend_backward_mma(S, cam, g, b, img, bg, colors, dL_color, dL_sem, dL_depth, dL_median,
                                         dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics,
                                         dL_ddepths, stream, debug);
#define HS_BWD_CASE(SV)                                                                                          \
    case SV:                                                                                                     \
        if (exact)                                                                                               \
            return launch_bwd_t<SV, (SV > 0)>(cam, g, b, img, bg, colors, semantics, dL_color, dL_sem, dL_depth, \
                                              dL_median, dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity,         \
                                              dL_dcolors, dL_dsemantics, dL_ddepths, stream, debug);             \
        return launch_bwd_t<SV, false>(cam, g, b, img, bg, colors, semantics, dL_color, dL_sem, dL_depth,        \
                                       dL_median, dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors,    \
                                       dL_dsemantics, dL_ddepths, stream, debug);
    switch (S) {
        HS_BWD_CASE(0)
        HS_BWD_CASE(16)
        HS_BWD_CASE(26)
        HS_BWD_CASE(32)
        HS_BWD_CASE(48)
        HS_BWD_CASE(64)
        HS_BWD_CASE(74)
        HS_BWD_CASE(102)
        default:
            set_error("semantic channel count S=%d is not instantiated (built: 0,16,26,32,48,64,74,102)", S);
            return 3;
    }
#undef HS_BWD_CASE
}

}  // namespace hs
