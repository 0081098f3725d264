// Backward of the alpha compositing, tensor-core version (default path; reference behaviour:
// cuda_rasterizer/backward.cu:669-899 / :472-666 with quirk Q1 in its observable form, see blend_bwd.cu for the
// SIMT version that also implements the 'exact' semantic gradient).
//
// Per (tile, Gaussian j) the backward needs sums over the tile's pixels.  All of them are contractions:
//     dL/dfeat[j][c]  = sum_pix w[j,pix] * dL/dout[pix][c]          c in {sem 0..S-1, r, g, b, depth, silhouette}
//     moments[j][m]   = sum_pix g[j,pix] * phi_m(pix)               phi = {1, x, y, x^2, x y, y^2} (tile-centred)
// with w = alpha * T_front and g = dL/dG * G.  The 2D-mean, conic and opacity gradients are closed-form functions of
// the six moments (dx = x_j - x, so sum g dx^2 = x_j^2 M0 - 2 x_j M1 + M3, ...).  Each warp owns one pixel region of the tile (an 8x4 block; "strip" below, hs_common.cuh):
//   phase 1 (SIMT, one pixel per lane): walk, back to front and with the reference's alpha / T arithmetic, the list
//            entries that the forward blended into this strip (strip_hits: ~37 % of the (strip, entry) pairs), and write
//            w and g into two per-warp shared-memory matrices [16][32] whose rows are packed (row r <-> r-th hit entry);
//   phase 2 (tensor cores): D[16 Gaussians][8 channels] += A[16][8 pixels] * B[8 pixels][8 channels] with
//            mma.sync.m16n8k8 TF32 and the 3xTF32 split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate), so the
//            result keeps fp32 accuracy (the moment basis is exactly representable in TF32, two products suffice);
//   then     every warp scatters its D rows to the entries' rows of a warp-private partial tile (no shared-memory
//            atomics: fp32 atomicAdd on shared memory is a CAS loop, ATOMS.CAST.SPIN), and once per round of 32 entries
//            all 256 threads sum the partials of the warps that wrote a row and flush each (tile, Gaussian, channel)
//            ONCE to global memory with coalesced atomics.
// This replaces the (S+10)-value warp shuffle reduction of the SIMT version -- 64 % of a step in round 1 -- by
// ~8 tensor instructions per (warp, Gaussian).
#include "hs_common.cuh"
#include <cuda_pipeline.h>

#ifndef HS_BWD_OCC_S0
#define HS_BWD_OCC_S0 3   // CTAs per SM the S = 0 instantiation (tracking, non-semantic maps) is compiled for: 85 registers, 605 -> 521 us at c2
#endif
#ifndef HS_BWD_U
#define HS_BWD_U 2      // entries whose alpha is evaluated together (independent dependency chains per lane)
#endif

namespace hs {

template <int S>
struct MmaCfg {
    static constexpr int NF = S + 5;                 // sem[S] r g b depth silhouette
    static constexpr int NBF = (NF + 7) / 8;         // feature n-tiles
    static constexpr int KA = 8 * NBF + 8;           // + one n-tile of moments
    static constexpr bool B_IN_REGS = (NBF <= 4);    // dL fragments live in registers for S <= 27
    static constexpr int DS = 8 * NBF + ((8 - (8 * NBF) % 32 + 32) % 32);  // smem dL row stride == 8 (mod 32)
    static constexpr int BATCH = B_IN_REGS ? 32 : 24;  // tile-list entries staged per round (<= 32: one ballot word; 24 keeps S = 74 inside 227 KB)
    static constexpr int WS = 36;                    // row stride of the per-warp w / g matrices (conflict-free)
};

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                         const uint32_t a3, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// TF32 keeps the upper 19 bits; the remainder is exactly representable in fp32.
__device__ __forceinline__ float tf32_lo(const float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

// NW = warps per CTA: a CTA owns a 16 x (2 NW) pixel slab of a tile (NW = 8: the whole tile; NW = 4: half a tile,
// two CTAs walk the same Gaussian list).  Smaller CTAs give the register / shared-memory limited kernel a finer
// occupancy granularity and cheaper barriers.
// The transmittance in front of an entry is reconstructed as T / (1 - alpha) (backward.cu:815).  A division is ~45 cycles
// of dependent instructions on the loop-carried chain of this latency-bound loop, so the reciprocal of (1 - alpha) is
// evaluated next to alpha (off the chain: rcp.approx + one Newton step) and the chain is
//   EXACT_T  (a median-depth gradient is present: the reconstructed T decides which Gaussian receives it, quirk Q4):
//            q = T r; rem = fma(-d, q, T); T' = fma(r, rem, q) -- the quotient-correction sequence of the IEEE division,
//            bit-identical to T / d on d in [0.01, 1], T in (0, 1] (tools/micro/div_exact.cu: 2^33 pairs, 0 mismatches);
//   !EXACT_T (every Hier-SLAM loss: no gradient reaches the median depth): T' = T r, one multiply per entry (<= 1 ulp).
template <int S, int NW, bool EXACT_T>
__global__ void __launch_bounds__(32 * NW, (MmaCfg<S>::B_IN_REGS ? (S == 0 ? HS_BWD_OCC_S0 : 2) * (8 / NW) : 1)) blend_backward_mma_kernel(
    const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list, int W, int H, int grid_x,
    const float* __restrict__ bg_color, const float2* __restrict__ means2D, const float4* __restrict__ conic_opacity,
    const float* __restrict__ colors, const float* __restrict__ depths, const float* __restrict__ final_Ts,
    const uint32_t* __restrict__ n_contrib, const uint8_t* __restrict__ strip_hits,
    const float* __restrict__ dL_dpixels,
    const float* __restrict__ dL_dpixels_sem, const float* __restrict__ dL_dpixel_depths,
    const float* __restrict__ dL_dpixel_medians, const float* __restrict__ dL_dpixel_opacitys,
    float* __restrict__ dL_dmean2D, float* __restrict__ dL_dconic2D, float* __restrict__ dL_dopacity,
    float* __restrict__ dL_dcolors, float* __restrict__ dL_dsemantics, float* __restrict__ dL_ddepths,
    int sem_stride) {
    using Cfg = MmaCfg<S>;
    constexpr int B = Cfg::BATCH, NF = Cfg::NF, NBF = Cfg::NBF, KA = Cfg::KA, WS = Cfg::WS, DS = Cfg::DS;
    constexpr bool BREG = Cfg::B_IN_REGS;
    constexpr int U = HS_BWD_U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // per-Gaussian staging is double-buffered and filled with cp.async one round ahead (ids two rounds ahead)
    float4* s_co2 = reinterpret_cast<float4*>(smem_raw);         // [2][B] conic + opacity
    float4* s_feat2 = s_co2 + 2 * B;                             // [2][B] r g b depth
    float* s_part = reinterpret_cast<float*>(s_feat2 + 2 * B);   // [NW warps][B][KA] warp-private partial sums
    float* s_w = s_part + NW * B * KA;                           // [NW warps][16][WS]
    float* s_g = s_w + NW * 16 * WS;                             // [NW warps][16][WS]
    float2* s_xy2 = reinterpret_cast<float2*>(s_g + NW * 16 * WS); // [2][B]
    int* s_id3 = reinterpret_cast<int*>(s_xy2 + 2 * B);          // [3][B] ring of Gaussian ids
    uint32_t* s_valid = reinterpret_cast<uint32_t*>(s_id3 + 3 * B);   // [NW] rows of s_part each warp wrote this round
    uint32_t* s_hit3 = s_valid + 8;                              // [3][B] strips the forward blended each entry into
    float* s_dL = reinterpret_cast<float*>(s_hit3 + 3 * B);      // [32 NW][DS]   (only when !BREG)
    __shared__ int s_maxc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int SLABS = 8 / NW;                                // CTAs per tile
    const int tile_x = blockIdx.x, tile_y = blockIdx.y / SLABS;
    const int region = (blockIdx.y % SLABS) * NW + warp;         // this warp's pixel region of the tile (hs_common.cuh)
    const uint32_t px = tile_x * HS_TILE_X + HS_PX_X(region, lane);
    const uint32_t py = tile_y * HS_TILE_Y + HS_PX_Y(region, lane);
    const uint32_t pix_id = W * py + px;
    const float2 pixf = {(float)px, (float)py};
    const bool inside = px < (uint32_t)W && py < (uint32_t)H;
    const uint2 range = ranges[tile_y * grid_x + tile_x];
    const size_t HW = (size_t)H * W;

    const float T_final = inside ? final_Ts[pix_id] : 0;
    float T = T_final;
    const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

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

    // upstream-gradient plane of feature channel c (nullptr = treated as zero): sem[S] r g b depth silhouette
    auto plane_of = [&](int c) -> const float* {
        if (c < S) return dL_dpixels_sem ? dL_dpixels_sem + (size_t)c * HW : nullptr;
        if (c < S + 3) return dL_dpixels ? dL_dpixels + (size_t)(c - S) * HW : nullptr;
        if (c == S + 3) return dL_dpixel_depths;
        if (c == S + 4) return dL_dpixel_opacitys;
        return nullptr;
    };
    // global pixel index of pixel p (0..31) of this warp's pixel region, -1 outside the image
    auto pixel_of = [&](int p) -> int {
        const uint32_t x = tile_x * HS_TILE_X + HS_PX_X(region, p), y = tile_y * HS_TILE_Y + HS_PX_Y(region, p);
        return (x < (uint32_t)W && y < (uint32_t)H) ? (int)(W * y + x) : -1;
    };

    // this lane's own pixel: what the alpha-gradient recurrence needs
    float dL_rgb[3] = {0.f, 0.f, 0.f}, dL_depth = 0.f, dL_median = 0.f, dL_op = 0.f;
    if (inside) {
        if (dL_dpixels) {
#pragma unroll
            for (int i = 0; i < 3; i++) dL_rgb[i] = dL_dpixels[i * HW + pix_id];
        }
        if (dL_dpixel_depths) dL_depth = dL_dpixel_depths[pix_id];
        if (dL_dpixel_medians) dL_median = dL_dpixel_medians[pix_id];
        if (dL_dpixel_opacitys) dL_op = dL_dpixel_opacitys[pix_id];
    }
    float bg_dot_dpixel = 0;
#pragma unroll
    for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_rgb[i];
    const bool bg_any = __any_sync(0xffffffffu, bg_dot_dpixel != 0.f);   // false for the black background SLAM uses

    // B operands (k = pixel, n = channel).  Fragment element (kb, nb, h): pixel 8 kb + lane%4 + 4 h, channel 8 nb + lane/4.
    const int qk = lane & 3, qn = lane >> 2;
    float bfrag[BREG ? NBF * 8 : 1];
    if (BREG) {
        int pid[8];
#pragma unroll
        for (int kb = 0; kb < 4; kb++)
#pragma unroll
            for (int h = 0; h < 2; h++) pid[kb * 2 + h] = pixel_of(8 * kb + qk + 4 * h);
#pragma unroll
        for (int nb = 0; nb < NBF; nb++) {
            const float* pl = plane_of(8 * nb + qn);
#pragma unroll
            for (int e = 0; e < 8; e++) bfrag[nb * 8 + e] = (pl != nullptr && pid[e] >= 0) ? __ldg(pl + pid[e]) : 0.f;
        }
    } else {
        float* my = s_dL + (size_t)(warp * 32) * DS;
        const int id = pixel_of(lane);
        constexpr int LU = (8 * NBF) % 16 == 0 ? 16 : 8;   // independent loads in flight per lane
#pragma unroll 1
        for (int c0 = 0; c0 < 8 * NBF; c0 += LU) {
            float v[LU];
#pragma unroll
            for (int u = 0; u < LU; u++) {
                const float* pl = plane_of(c0 + u);
                v[u] = (pl != nullptr && id >= 0) ? __ldg(pl + id) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < LU; u++) my[lane * DS + c0 + u] = v[u];
        }
    }
    // moment basis for this lane's fragment pixels: column qn of {1, x, y, x^2, xy, y^2, 0, 0}, tile-centred
    float mfrag[8];
#pragma unroll
    for (int kb = 0; kb < 4; kb++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int p = 8 * kb + qk + 4 * h;
            const float x = (float)HS_PX_X(region, p) - 7.5f, y = (float)HS_PX_Y(region, p) - 7.5f;
            mfrag[kb * 2 + h] = qn == 0 ? 1.f : qn == 1 ? x : qn == 2 ? y : qn == 3 ? x * x : qn == 4 ? x * y
                                : qn == 5 ? y * y : 0.f;
        }

    // flush destinations of the feature columns this lane owns (column c = 32 k + lane): pointer + row stride
    float* feat_dst[(NBF * 8 + 31) / 32];
    int feat_stride[(NBF * 8 + 31) / 32];
#pragma unroll
    for (int k = 0; k < (NBF * 8 + 31) / 32; k++) {
        const int c = 32 * k + lane;
        feat_dst[k] = c < S ? dL_dsemantics + c : c < S + 3 ? dL_dcolors + (c - S) : c == S + 3 ? dL_ddepths : dL_dopacity;
        feat_stride[k] = c < S ? sem_stride : c < S + 3 ? 3 : 1;   // sem_stride: row length of dL_dsemantics (>= S)
    }
    float* wm = s_w + warp * 16 * WS;
    float* gm = s_g + warp * 16 * WS;
    float last_alpha = 0.f, last_q = 0.f, accum_q = 0.f;

    const int rounds = (total + B - 1) / B;
    // list position of entry t of round r (back to front), and the asynchronous gathers of one round
    auto list_pos = [&](int r, int t) { return range.x + (total - 1 - r * B - t); };
    auto gather_round = [&](int r) {   // ids of round r must already be in s_id3[r % 3]
        const int n = min(B, total - r * B);
        if (tid < n) {
            const int id = s_id3[(r % 3) * B + tid];
            const int bf = (r & 1) * B + tid;
            __pipeline_memcpy_async(s_xy2 + bf, means2D + id, 8);
            __pipeline_memcpy_async(s_co2 + bf, conic_opacity + id, 16);
            float* f = reinterpret_cast<float*>(s_feat2 + bf);
            __pipeline_memcpy_async(f, colors + 3 * (size_t)id, 4);
            __pipeline_memcpy_async(f + 1, colors + 3 * (size_t)id + 1, 4);
            __pipeline_memcpy_async(f + 2, colors + 3 * (size_t)id + 2, 4);
            __pipeline_memcpy_async(f + 3, depths + id, 4);
        }
    };
    auto fetch_ids = [&](int r) {
        const int n = min(B, total - r * B);
        if (r < rounds && tid < n) {
            __pipeline_memcpy_async(s_id3 + (r % 3) * B + tid, point_list + list_pos(r, tid), 4);
            s_hit3[(r % 3) * B + tid] = strip_hits[list_pos(r, tid)];
        }
    };
    const int strip = region;                                // the forward's strip_hits are indexed by the region number
    fetch_ids(0);
    fetch_ids(1);
    __pipeline_commit();
    __pipeline_wait_prior(0);
    __syncthreads();
    gather_round(0);
    __pipeline_commit();
    for (int i = 0; i < rounds; i++) {
        __pipeline_wait_prior(0);
        __syncthreads();  // staging of round i landed, ids of round i+1 present, previous flush finished
        const int nb_ = min(B, total - i * B);
        if (i + 1 < rounds) gather_round(i + 1);
        fetch_ids(i + 2);
        __pipeline_commit();
        const float2* s_xy = s_xy2 + (i & 1) * B;
        const float4* s_co = s_co2 + (i & 1) * B;
        const float4* s_feat = s_feat2 + (i & 1) * B;
        const int* s_id = s_id3 + (i % 3) * B;
        const uint32_t* s_hit = s_hit3 + (i % 3) * B;
        // Entries of this round that the forward blended into this warp's strip: exactly the entries with a contributing
        // lane.  Every other entry has act == false in all lanes and is not even evaluated; the hit entries are packed
        // into full 16-row MMA tiles, row r of a tile <-> entry `my_ent` of lane r.
        uint32_t todo = __ballot_sync(0xffffffffu, lane < nb_ && ((s_hit[lane] >> strip) & 1u));
        const uint32_t valid_rows = todo;   // rows of this warp's partial tile that get written this round
#pragma unroll 1   // keep the loops rolled: the kernel must stay inside the instruction cache
        while (todo) {
            // ---------------- phase 1: up to 16 entries, back to front ------------------------------------------------
            // (a) alpha of 8 entries at a time, branch-free: eight independent dependency chains per lane (ILP);
            // (b) the sequential transmittance / colour-behind recurrences.
            int my_ent = 0, nrow = 0;
#pragma unroll 1
            for (int part = 0; part < 16 / U && todo != 0; part++) {
                float oG[U], rinv[U];
                int ent[U];
                uint32_t abits = 0, have_bits = 0;
#pragma unroll
                for (int u = 0; u < U; u++) {
                    const bool have = todo != 0;
                    ent[u] = have ? __ffs(todo) - 1 : ent[0];
                    todo &= todo - 1;
                    have_bits |= (have ? 1u : 0u) << u;
                    const int j = ent[u];
                    const float2 xy = s_xy[j];
                    const float2 d = {xy.x - pixf.x, xy.y - pixf.y};
                    const float4 con_o = s_co[j];
                    const float power = gauss_power(d, con_o);
                    const float og = con_o.w * exp(power);
                    const int gi = total - 1 - i * B - j;
                    const bool active = have && (gi < last_contributor) && !(power > 0.0f) &&
                                        !(min(0.99f, og) < 1.0f / 255.0f);
                    oG[u] = og;
                    {
                        const float dd = 1.f - min(0.99f, og);
                        float r0;
                        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(dd));
                        rinv[u] = EXACT_T ? fmaf(r0, fmaf(-dd, r0, 1.0f), r0) : r0;
                    }
                    abits |= (active ? 1u : 0u) << u;
                }
#pragma unroll
                for (int u = 0; u < U; u++) {
                    if (!((have_bits >> u) & 1)) continue;   // warp-uniform
                    const int row = nrow++;
                    const int j = ent[u];
                    my_ent = (lane == row) ? j : my_ent;
                    const bool act = (abits >> u) & 1;
                    const float alpha = min(0.99f, oG[u]);
                    // same IEEE division as the reference (backward.cu:815): the reconstructed transmittance decides
                    // which Gaussian receives the median-depth gradient, so it is kept bit-identical
                    float test_T = T * rinv[u];
                    if (EXACT_T) test_T = fmaf(rinv[u], fmaf(alpha - 1.f, test_T, T), test_T);   // == T / (1 - alpha)
                    const float4 f = s_feat[j];
                    float q = f.x * dL_rgb[0] + f.y * dL_rgb[1] + f.z * dL_rgb[2];
                    q = fmaf(f.w, dL_depth, q);
                    q += dL_op;
                    const float acc_new = last_alpha * last_q + (1.f - last_alpha) * accum_q;
                    float dL_dalpha = (q - acc_new) * test_T;
                    if (bg_any) dL_dalpha -= __fdividef(T_final, 1.f - alpha) * bg_dot_dpixel;   // warp-uniform branch
                    if (EXACT_T && act && test_T > 0.5f && T < 0.5 && dL_median != 0.f)
                        atomicAdd(dL_ddepths + s_id[j], dL_median);   // the Gaussian that crossed T = 0.5 (quirk Q4)
                    wm[row * WS + lane] = act ? alpha * test_T : 0.f;
                    gm[row * WS + lane] = act ? dL_dalpha * oG[u] : 0.f;   // dL/dG * G = (o dL/dalpha) G
                    accum_q = act ? acc_new : accum_q;
                    last_q = act ? q : last_q;
                    last_alpha = act ? alpha : last_alpha;
                    T = act ? test_T : T;
                }
            }
            __syncwarp();
            // ---------------- phase 2: D[16 x 8] tiles on the tensor cores ---------------------------------------
            const bool v0 = qn < nrow, v1 = qn + 8 < nrow;
            float acc[NBF + 1][4];
#pragma unroll
            for (int nb = 0; nb <= NBF; nb++)
#pragma unroll
                for (int r = 0; r < 4; r++) acc[nb][r] = 0.f;
#pragma unroll
            for (int kb = 0; kb < 4; kb++) {
                const int col = 8 * kb + qk;
                const float a0 = v0 ? wm[qn * WS + col] : 0.f, a1 = v1 ? wm[(qn + 8) * WS + col] : 0.f;
                const float a2 = v0 ? wm[qn * WS + col + 4] : 0.f, a3 = v1 ? wm[(qn + 8) * WS + col + 4] : 0.f;
                const uint32_t ah0 = __float_as_uint(a0), ah1 = __float_as_uint(a1), ah2 = __float_as_uint(a2),
                               ah3 = __float_as_uint(a3);
                const uint32_t al0 = __float_as_uint(tf32_lo(a0)), al1 = __float_as_uint(tf32_lo(a1)),
                               al2 = __float_as_uint(tf32_lo(a2)), al3 = __float_as_uint(tf32_lo(a3));
#pragma unroll
                for (int nb = 0; nb < NBF; nb++) {
                    float b0, b1;
                    if (BREG) {
                        b0 = bfrag[(nb * 4 + kb) * 2];
                        b1 = bfrag[(nb * 4 + kb) * 2 + 1];
                    } else {
                        const float* my = s_dL + (size_t)(warp * 32 + col) * DS + 8 * nb + qn;
                        b0 = my[0];
                        b1 = my[4 * DS];
                    }
                    const uint32_t bh0 = __float_as_uint(b0), bh1 = __float_as_uint(b1);
                    mma_tf32(acc[nb], ah0, ah1, ah2, ah3, bh0, bh1);
                    mma_tf32(acc[nb], al0, al1, al2, al3, bh0, bh1);
                    mma_tf32(acc[nb], ah0, ah1, ah2, ah3, __float_as_uint(tf32_lo(b0)), __float_as_uint(tf32_lo(b1)));
                }
                // moments: the basis is exact in TF32, two products suffice
                const float e0 = v0 ? gm[qn * WS + col] : 0.f, e1 = v1 ? gm[(qn + 8) * WS + col] : 0.f;
                const float e2 = v0 ? gm[qn * WS + col + 4] : 0.f, e3 = v1 ? gm[(qn + 8) * WS + col + 4] : 0.f;
                const uint32_t mb0 = __float_as_uint(mfrag[kb * 2]), mb1 = __float_as_uint(mfrag[kb * 2 + 1]);
                mma_tf32(acc[NBF], __float_as_uint(e0), __float_as_uint(e1), __float_as_uint(e2), __float_as_uint(e3),
                         mb0, mb1);
                mma_tf32(acc[NBF], __float_as_uint(tf32_lo(e0)), __float_as_uint(tf32_lo(e1)),
                         __float_as_uint(tf32_lo(e2)), __float_as_uint(tf32_lo(e3)), mb0, mb1);
            }
            // D fragment: rows qn / qn+8, columns 2 qk, 2 qk + 1 of each n-tile -> row `entry` of this warp's private
            // partial tile (rows this warp does not write are masked out of the reduction by s_valid).
            const int e0 = __shfl_sync(0xffffffffu, my_ent, qn), e1 = __shfl_sync(0xffffffffu, my_ent, qn + 8);
            float* pt0 = s_part + (size_t)(warp * B + e0) * KA + 2 * qk;
            float* pt1 = s_part + (size_t)(warp * B + e1) * KA + 2 * qk;
#pragma unroll
            for (int nb = 0; nb <= NBF; nb++) {
                if (v0) *reinterpret_cast<float2*>(pt0 + 8 * nb) = make_float2(acc[nb][0], acc[nb][1]);
                if (v1) *reinterpret_cast<float2*>(pt1 + 8 * nb) = make_float2(acc[nb][2], acc[nb][3]);
            }
            __syncwarp();  // wm / gm are rewritten by the next group
        }
        if (lane == 0) s_valid[warp] = valid_rows;
        __syncthreads();

        // ---------------- reduce the 8 warp-private tiles and flush: thread <-> (Gaussian j, column c) -----------------
        // consecutive threads own consecutive columns -> coalesced global atomics; the 8 moment columns of a row sit in
        // 8 consecutive lanes of one warp, so the closed-form 2D-mean / conic / opacity terms gather them by shuffle.
        const float cx = (float)(tile_x * HS_TILE_X) + 7.5f, cy = (float)(tile_y * HS_TILE_Y) + 7.5f;
        uint32_t any_rows = 0, vrow[NW];
#pragma unroll
        for (int wv = 0; wv < NW; wv++) {
            vrow[wv] = s_valid[wv];
            any_rows |= vrow[wv];
        }
        // pass A, feature columns: lane <-> column (destination pointer and row stride are fixed per lane, so the loop
        // body is divergence-free), warp <-> Gaussian row.
#pragma unroll 1
        for (int j = warp; j < nb_; j += NW) {
            if (!((any_rows >> j) & 1)) continue;   // warp-uniform
            const size_t id = (size_t)s_id[j];
#pragma unroll
            for (int cb = 0; cb < NBF * 8; cb += 32) {
                const int c = cb + lane;
                if (c < NF) {
                    float p8[NW];
#pragma unroll
                    for (int wv = 0; wv < NW; wv++) p8[wv] = ((vrow[wv] >> j) & 1u) ? s_part[(size_t)(wv * B + j) * KA + c] : 0.f;
#pragma unroll
                    for (int st = 1; st < NW; st <<= 1)
#pragma unroll
                        for (int wv = 0; wv + st < NW; wv += 2 * st) p8[wv] += p8[wv + st];
                    if (p8[0] != 0.f) atomicAdd(feat_dst[cb / 32] + id * feat_stride[cb / 32], p8[0]);
                }
            }
        }
        // pass B, moments: 8 consecutive lanes <-> the moment columns of one Gaussian
#pragma unroll 1
        for (int e = tid; e < nb_ * 8; e += 32 * NW) {
            const int j = e >> 3, m = e & 7;
            float p8[NW];
#pragma unroll
            for (int wv = 0; wv < NW; wv++)
                p8[wv] = ((vrow[wv] >> j) & 1u) ? s_part[(size_t)(wv * B + j) * KA + 8 * NBF + m] : 0.f;
#pragma unroll
            for (int st = 1; st < NW; st <<= 1)
#pragma unroll
                for (int wv = 0; wv + st < NW; wv += 2 * st) p8[wv] += p8[wv + st];
            const float sum = p8[0];
            const int base = lane & ~7;
            const float M0 = __shfl_sync(0xffffffffu, sum, base), M1 = __shfl_sync(0xffffffffu, sum, base + 1);
            const float M2 = __shfl_sync(0xffffffffu, sum, base + 2), M3 = __shfl_sync(0xffffffffu, sum, base + 3);
            const float M4 = __shfl_sync(0xffffffffu, sum, base + 4), M5 = __shfl_sync(0xffffffffu, sum, base + 5);
            if (!((any_rows >> j) & 1) || m >= 6 || M0 == 0.f && M1 == 0.f && M2 == 0.f && M3 == 0.f && M4 == 0.f && M5 == 0.f)
                continue;   // all-zero moments: a channel-split pass without alpha gradients (S = 102), or nothing to add
            const size_t id = (size_t)s_id[j];
            const float2 xy = s_xy[j];
            const float4 co = s_co[j];
            const float xj = xy.x - cx, yj = xy.y - cy;
            const float sgdx = xj * M0 - M1, sgdy = yj * M0 - M2;  // sum g dx, sum g dy
            if (m == 0) atomicAdd(dL_dopacity + id, M0 / co.w);           // sum_pix G dL/dalpha
            else if (m == 1) atomicAdd(dL_dmean2D + id * 3, -(co.x * sgdx + co.y * sgdy) * (0.5f * W));
            else if (m == 2) atomicAdd(dL_dmean2D + id * 3 + 1, -(co.z * sgdy + co.y * sgdx) * (0.5f * H));
            else if (m == 3) atomicAdd(dL_dconic2D + id * 4, -0.5f * (xj * xj * M0 - 2.f * xj * M1 + M3));
            else if (m == 4) atomicAdd(dL_dconic2D + id * 4 + 1, -0.5f * (xj * yj * M0 - xj * M2 - yj * M1 + M4));
            else atomicAdd(dL_dconic2D + id * 4 + 3, -0.5f * (yj * yj * M0 - 2.f * yj * M2 + M5));
        }
    }
}

template <int S, int NW, bool EXACT_T>
static int launch_bwd_mma_t(const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                            const float* bg, const float* colors, const float* dL_color, const float* dL_sem,
                            const float* dL_depth, const float* dL_median, const float* dL_opacity, float* dL_dmean2D,
                            float* dL_dconic, float* dL_dopacity, float* dL_dcolors, float* dL_dsemantics,
                            float* dL_ddepths, int sem_stride, cudaStream_t stream, bool debug) {
    using Cfg = MmaCfg<S>;
    size_t smem = (size_t)Cfg::BATCH * (4 * sizeof(float4) + NW * Cfg::KA * sizeof(float) + 2 * sizeof(float2) + 6 * sizeof(int)) +
                  (size_t)2 * NW * 16 * Cfg::WS * sizeof(float) + 8 * sizeof(uint32_t);
    if (!Cfg::B_IN_REGS) smem += (size_t)32 * NW * Cfg::DS * sizeof(float);
    auto k = blend_backward_mma_kernel<S, NW, EXACT_T>;
    HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(cam.grid_x, cam.grid_y * (8 / NW), 1);
    prof_begin(ST_BLEND_BWD, stream);
    k<<<grid, 32 * NW, smem, stream>>>(img.ranges, b.point_list, cam.W, cam.H, cam.grid_x, bg, g.means2D, g.conic_opacity,
                                   colors, g.depths, img.final_T, img.n_contrib, b.strip_hits, dL_color, dL_sem, dL_depth, dL_median,
                                   dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics,
                                   dL_ddepths, sem_stride);
    prof_end(ST_BLEND_BWD, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

int launch_blend_backward_mma(int S, const Camera& cam, const GeomView& g, const BinningView& b, const ImageView& img,
                              const float* bg, const float* colors, const float* dL_color, const float* dL_sem,
                              const float* dL_depth, const float* dL_median, const float* dL_opacity,
                              float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolors,
                              float* dL_dsemantics, float* dL_ddepths, cudaStream_t stream, bool debug) {
// warps per CTA: whole-tile CTAs everywhere (half-tile CTAs, NW = 4, measured slower: 856 vs 804 us at c2)
#define HS_BWD_NW(SV) 8
#define HS_BWDM_CASE(SV)                                                                                        \
    case SV:                                                                                                    \
        if (dL_median != nullptr)                                                                               \
            return launch_bwd_mma_t<SV, HS_BWD_NW(SV), true>(cam, g, b, img, bg, colors, dL_color, dL_sem, dL_depth, dL_median, \
                                        dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics,     \
                                        dL_ddepths, SV, stream, debug);                                                \
        return launch_bwd_mma_t<SV, HS_BWD_NW(SV), false>(cam, g, b, img, bg, colors, dL_color, dL_sem, dL_depth, dL_median,   \
                                    dL_opacity, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics,         \
                                    dL_ddepths, SV, stream, debug);
    if (S == 102) {
        // Replica's flat 102-class map: two passes over 51 channels each.  The first carries everything that depends on
        // the alpha gradients (colour, depth, silhouette, moments, median routing) plus dL/dsemantics[:, 0:51]; the second
        // has no upstream gradient except the semantic planes 51..101, so its alpha gradients, moments and colour columns
        // are exactly zero (skipped by the flush) and only dL/dsemantics[:, 51:102] is produced.  In the reference-observable
        // mode the semantic channels never reach dL/dalpha (quirk Q1), which is what makes the split exact.
        const size_t HW = (size_t)cam.W * cam.H;
        int rc;
        if (dL_median != nullptr)
            rc = launch_bwd_mma_t<51, 8, true>(cam, g, b, img, bg, colors, dL_color, dL_sem, dL_depth, dL_median, dL_opacity,
                                               dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics, dL_ddepths, 102,
                                               stream, debug);
        else
            rc = launch_bwd_mma_t<51, 8, false>(cam, g, b, img, bg, colors, dL_color, dL_sem, dL_depth, dL_median, dL_opacity,
                                                dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics, dL_ddepths, 102,
                                                stream, debug);
        if (rc || dL_sem == nullptr) return rc;
        return launch_bwd_mma_t<51, 8, false>(cam, g, b, img, bg, colors, nullptr, dL_sem + 51 * HW, nullptr, nullptr, nullptr,
                                              dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolors, dL_dsemantics + 51, dL_ddepths,
                                              102, stream, debug);
    }
    switch (S) {
        HS_BWDM_CASE(0)
        HS_BWDM_CASE(16)
        HS_BWDM_CASE(26)
        HS_BWDM_CASE(32)
        HS_BWDM_CASE(48)
        HS_BWDM_CASE(64)
        HS_BWDM_CASE(74)
        default:
            set_error("tensor-core blend backward: S=%d is not instantiated (built: 0,16,26,32,48,64,74)", S);
            return 3;
    }
#undef HS_BWDM_CASE
}

}  // namespace hs
