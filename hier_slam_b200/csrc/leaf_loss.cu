// Leaf-level semantic loss of Hier-SLAM's tree encoding:
//     logits = Conv2d(S -> L, kernel 1)(sem);  loss = scale * sum_pixels CE(logits, leaf label)
// (scripts/hierslam.py:975-984 / :1009-1016: MLP_func = torch.nn.Conv2d(num_semantic, num_semantic_class, 1), :1756).
// In torch this materialises the [L,H,W] logits (333 MB at 1200x680 with Replica's 102 leaves, 676 MB for ScanNet's
// 550), their log-softmax and both gradients.  Here nothing of size L x pixels ever exists: the per-pixel work is a
// chain of small GEMMs whose operands live in shared memory and registers, evaluated with mma.sync.m16n8k8 TF32 in the
// 3xTF32 split (a_hi b_hi + a_lo b_hi + a_hi b_lo, fp32 accumulate: fp32-accurate).
//
//   leaf_ce_pixel_kernel   one warp owns 16-pixel row tiles.  Pass 1: Z = X W^T tile by tile (8 classes at a time) with an
//                          online max / sum per pixel -> logsumexp, loss, the label's logit.  Pass 2: Z again,
//                          G = scale (softmax - onehot) in the accumulator layout, and dX += G W with G fed back as
//                          the A operand WITHOUT a shuffle: the contraction index (the class) may be permuted freely, so
//                          accumulator column 2t / 2t+1 of a thread is declared to be k = t / t+4 and the B fragment
//                          reads the matching rows of W.  The bias rides along as an extra "ones" channel of X.
//   leaf_ce_weight_kernel  dW = G^T X contracts over PIXELS, which the accumulator layout spreads over the wrong lane
//                          index, so G goes through a warp-private 16x16 shared tile.  A CTA owns a group of classes
//                          (its accumulators stay in registers for the whole kernel), recomputes only that group's
//                          logits from the stored logsumexp, and adds its block of dW once at the end.
#include "hs_common.cuh"
#include <cuda_pipeline.h>

namespace hs {
namespace leaf {

__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
          "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
// TF32 keeps the upper 19 bits of an fp32 (the tensor core ignores the rest); the remainder is exact in fp32.
__device__ __forceinline__ float lo_part(const float x) { return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
template <int N> __device__ __forceinline__ void lo_parts(float (&lo)[N], const float (&x)[N]) {
#pragma unroll
    for (int i = 0; i < N; i++) lo[i] = lo_part(x[i]);
}
// d (+ dc) += a b.  P3: fp32-accurate 3xTF32 (a_hi b_hi -> d; a_lo b_hi + a_hi b_lo -> dc, a second, independent
// accumulator chain that the caller adds at the end); !P3: one TF32 product (what torch's cuDNN convolution does with
// its default allow_tf32 = True).
template <bool P3>
__device__ __forceinline__ void mma_p(float (&d)[4], float (&dc)[4], const float (&a)[4], const float (&al)[4],
                                      const float (&b)[2], const float (&bl)[2]) {
    mma_tf32(d, a, b);
    if (P3) {
        mma_tf32(dc, al, b);
        mma_tf32(dc, a, bl);
    }
}

constexpr int NWARP = 8;

template <int NJ, int MTV = (NJ <= 4 ? 2 : 1)> struct Cfg {
    static constexpr int SP = 8 * NJ;                 // padded channel count: S + 1 ("ones" channel = bias) rounded to 8
    static constexpr int MT = MTV;                    // 16-pixel row tiles per warp
    static constexpr int BPX = NWARP * 16 * MT;       // pixels per block iteration
    static constexpr int XS = BPX + 8;                // row stride of Xs: == 8 (mod 32) -> conflict-free A-fragment loads
    static constexpr int WS = SP + 4;                 // row stride of Ws: == 4, 20 or 28 (mod 32) -> conflict-free B loads
    static constexpr int LC = NJ <= 4 ? 128 : 64;     // classes per staged weight chunk of the pixel kernel (2 CTAs / SM)
    static constexpr int LG = NJ <= 4 ? 64 : 32;      // classes per CTA of the weight-gradient kernel
    static constexpr int GS = 20;                     // row stride of the warp-private G tiles (conflict-free)
    static constexpr int OCC = 2;                     // CTAs per SM of the pixel kernel
};

// Xs[s][p] = sem[s][px0 + p] for s < S, 1 for s == S (bias channel), 0 above; pixels beyond the image are 0.  Full
// 4-pixel groups of real channels travel with 16-byte cp.async (the caller commits / waits), the rest is stored directly.
template <int NJ, int MTV = (NJ <= 4 ? 2 : 1)>
__device__ __forceinline__ void stage_pixels(float* Xs, const float* __restrict__ sem, int S, size_t HW, size_t px0,
                                             bool vec_ok) {
    using C = Cfg<NJ, MTV>;
    for (int idx = threadIdx.x; idx < C::SP * (C::BPX / 4); idx += 32 * NWARP) {
        const int s = idx / (C::BPX / 4), p = (idx % (C::BPX / 4)) * 4;
        const size_t px = px0 + p;
        float* dst = Xs + s * C::XS + p;
        if (s < S && vec_ok && px + 3 < HW) {
            __pipeline_memcpy_async(dst, sem + (size_t)s * HW + px, 16);
        } else {
#pragma unroll
            for (int q = 0; q < 4; q++)
                dst[q] = px + q < HW ? (s < S ? sem[(size_t)s * HW + px + q] : (s == S ? 1.f : 0.f)) : 0.f;
        }
    }
}

// Ws[l][s] = weight[c0 + l][s] for s < S, bias for s == S, 0 elsewhere and for classes >= L.
template <int NJ>
__device__ __forceinline__ void stage_weights(float* Ws, const float* __restrict__ weight, const float* __restrict__ bias,
                                              int S, int L, int c0, int rows) {
    using C = Cfg<NJ>;
    for (int idx = threadIdx.x; idx < rows * C::SP; idx += 32 * NWARP) {
        const int l = idx / C::SP, s = idx % C::SP, cls = c0 + l;
        float v = 0.f;
        if (cls < L) v = s < S ? weight[(size_t)cls * S + s] : (s == S && bias != nullptr ? bias[cls] : 0.f);
        Ws[l * C::WS + s] = v;
    }
}

// z[mt][n][16 px x 8 classes] = X W^T for the MT row tiles of a warp (first pixel pxl) and the two class tiles
// 2 np, 2 np + 1 of the staged weights.  A fragments are shared by the two class tiles, B fragments by the row tiles;
// 4 MT (8 MT with the split correction accumulators) independent mma chains are in flight.
template <int NJ, bool P3, bool SPLIT, int MTV = (NJ <= 4 ? 2 : 1)>
__device__ __forceinline__ void logits_pair(float (&z)[MTV][2][4], const float* Xs, const float* Ws, int pxl,
                                            int np, int g, int t) {
    using C = Cfg<NJ, MTV>;
    float zc[SPLIT ? C::MT : 1][2][4];            // SPLIT: the two correction products run as a second accumulator chain
#pragma unroll
    for (int mt = 0; mt < C::MT; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                z[mt][n][q] = 0.f;
                if (SPLIT) zc[mt][n][q] = 0.f;
            }
#pragma unroll
    for (int k = 0; k < NJ; k++) {
        float b[2][2], bl[2][2];
#pragma unroll
        for (int n = 0; n < 2; n++) {
            const float* wb = Ws + (16 * np + 8 * n + g) * C::WS + 8 * k + t;
            b[n][0] = wb[0];
            b[n][1] = wb[4];
            if (P3) lo_parts(bl[n], b[n]);
        }
#pragma unroll
        for (int mt = 0; mt < C::MT; mt++) {
            const float* xa = Xs + (8 * k + t) * C::XS + pxl + 16 * mt + g;
            const float a[4] = {xa[0], xa[8], xa[4 * C::XS], xa[4 * C::XS + 8]};
            float al[4];
            if (P3) lo_parts(al, a);
#pragma unroll
            for (int n = 0; n < 2; n++) mma_p<P3>(z[mt][n], SPLIT ? zc[mt][n] : z[mt][n], a, al, b[n], bl[n]);
        }
    }
    if (P3 && SPLIT) {
#pragma unroll
        for (int mt = 0; mt < C::MT; mt++)
#pragma unroll
            for (int n = 0; n < 2; n++)
#pragma unroll
                for (int q = 0; q < 4; q++) z[mt][n][q] += zc[mt][n][q];
    }
}

template <int NJ, bool P3>
__global__ void __launch_bounds__(32 * NWARP, Cfg<NJ>::OCC) leaf_ce_pixel_kernel(
    const float* __restrict__ sem, const int* __restrict__ labels, const float* __restrict__ weight,
    const float* __restrict__ bias, int S, int L, size_t HW, float scale, float* __restrict__ loss,
    float* __restrict__ lse_out, float* __restrict__ grad_sem, int accumulate) {
    using C = Cfg<NJ>;
    constexpr int MT = C::MT, LC = C::LC;
    extern __shared__ float smem[];
    float* Xbuf = smem;                           // [2][SP][XS]  double buffered pixel blocks
    float* Ws = smem + 2 * C::SP * C::XS;         // [LC][WS]
    __shared__ float s_part[NWARP];
    const bool resident = L <= LC;                // all classes fit one chunk (Replica: 102): staged once per CTA
    const bool vec_ok = (HW & 3) == 0 && (reinterpret_cast<size_t>(sem) & 15) == 0;
    if (resident) stage_weights<NJ>(Ws, weight, bias, S, L, 0, LC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pxl = warp * 16 * MT;
    const size_t n_blocks = (HW + C::BPX - 1) / C::BPX;
    float loss_acc = 0.f;
    int buf = 0;
    if (blockIdx.x < n_blocks) stage_pixels<NJ>(Xbuf, sem, S, HW, (size_t)blockIdx.x * C::BPX, vec_ok);
    __pipeline_commit();
    for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, buf ^= 1) {
        const size_t px0 = blk * C::BPX;
        const float* Xs = Xbuf + buf * C::SP * C::XS;
        __pipeline_wait_prior(0);
        __syncthreads();                          // this block's pixels are visible; the other buffer (and Ws) is free
        if (blk + gridDim.x < n_blocks)
            stage_pixels<NJ>(Xbuf + (buf ^ 1) * C::SP * C::XS, sem, S, HW, (blk + gridDim.x) * C::BPX, vec_ok);
        __pipeline_commit();
        int y[MT][2];
        float sc[MT][2], m[MT][2], sum[MT][2], zy[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const size_t px = px0 + pxl + 16 * mt + g + 8 * r;
                const int lab = px < HW ? labels[px] : -1;
                const bool use = lab >= 0 && lab < L;           // torch's ignore_index (-100); out-of-range is ignored too
                y[mt][r] = use ? lab : -1;
                sc[mt][r] = use ? scale : 0.f;
                m[mt][r] = -1.0e30f;
                sum[mt][r] = 0.f;
                zy[mt][r] = 0.f;
            }
        // ---- pass 1: online max / sum over the classes
        for (int c0 = 0; c0 < L; c0 += LC) {
            if (!resident) {
                __syncthreads();
                stage_weights<NJ>(Ws, weight, bias, S, L, c0, LC);
                __syncthreads();
            }
            const int npairs = (min(LC, L - c0) + 15) >> 4;
            for (int np = 0; np < npairs; np++) {
                float z[MT][2][4];
                logits_pair<NJ, P3, true>(z, Xs, Ws, pxl, np, g, t);
#pragma unroll
                for (int n = 0; n < 2; n++) {
                    const int l0 = c0 + 16 * np + 8 * n + 2 * t;
#pragma unroll
                    for (int mt = 0; mt < MT; mt++)
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const float za = l0 < L ? z[mt][n][2 * r] : -3.0e38f;
                            const float zb = l0 + 1 < L ? z[mt][n][2 * r + 1] : -3.0e38f;
                            const float mn = fmaxf(m[mt][r], fmaxf(za, zb));
                            sum[mt][r] = sum[mt][r] * __expf(m[mt][r] - mn) + __expf(za - mn) + __expf(zb - mn);
                            m[mt][r] = mn;
                            if (l0 == y[mt][r]) zy[mt][r] = za;
                            if (l0 + 1 == y[mt][r]) zy[mt][r] = zb;
                        }
                }
            }
        }
        float lse[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                float mm = m[mt][r], ss = sum[mt][r], zz = zy[mt][r];
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {              // the four lanes of a quad share a pixel row
                    const float mo = __shfl_xor_sync(0xffffffffu, mm, o), so = __shfl_xor_sync(0xffffffffu, ss, o);
                    const float mn = fmaxf(mm, mo);
                    ss = ss * __expf(mm - mn) + so * __expf(mo - mn);
                    mm = mn;
                    zz += __shfl_xor_sync(0xffffffffu, zz, o);  // exactly one lane holds the label's logit
                }
                lse[mt][r] = mm + __logf(ss);
                const size_t px = px0 + pxl + 16 * mt + g + 8 * r;
                if (t == 0 && px < HW) {
                    lse_out[px] = lse[mt][r];
                    loss_acc += sc[mt][r] * (lse[mt][r] - zz);
                }
            }
        // ---- pass 2: G = scale (softmax - onehot), dX += G W
        float dx[MT][NJ][4];                      // MT NJ independent accumulator chains: no split needed here
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int j = 0; j < NJ; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) dx[mt][j][q] = 0.f;
        for (int c0 = 0; c0 < L; c0 += LC) {
            if (!resident) {
                __syncthreads();
                stage_weights<NJ>(Ws, weight, bias, S, L, c0, LC);
                __syncthreads();
            }
            const int npairs = (min(LC, L - c0) + 15) >> 4;
            for (int np = 0; np < npairs; np++) {
                float z[MT][2][4];
                logits_pair<NJ, P3, false>(z, Xs, Ws, pxl, np, g, t);
#pragma unroll
                for (int n = 0; n < 2; n++) {
                    const int l0 = c0 + 16 * np + 8 * n + 2 * t;
                    // A operand = G with the class index permuted: k = t <-> class 2t, k = t+4 <-> class 2t+1, i.e.
                    // a0 = G(row g, 2t), a1 = G(row g+8, 2t), a2 = G(row g, 2t+1), a3 = G(row g+8, 2t+1)
                    float ga[MT][4], gl[MT][4];
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                        for (int r = 0; r < 2; r++) {
                            const float pa = l0 < L ? __expf(z[mt][n][2 * r] - lse[mt][r]) : 0.f;
                            const float pb = l0 + 1 < L ? __expf(z[mt][n][2 * r + 1] - lse[mt][r]) : 0.f;
                            ga[mt][r] = sc[mt][r] * (pa - (l0 == y[mt][r] ? 1.f : 0.f));
                            ga[mt][2 + r] = sc[mt][r] * (pb - (l0 + 1 == y[mt][r] ? 1.f : 0.f));
                        }
                        if (P3) lo_parts(gl[mt], ga[mt]);
                    }
                    const float* w0 = Ws + (16 * np + 8 * n + 2 * t) * C::WS + g;
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
                        const float b[2] = {w0[8 * j], w0[C::WS + 8 * j]};
                        float bl[2];
                        if (P3) lo_parts(bl, b);
#pragma unroll
                        for (int mt = 0; mt < MT; mt++) mma_p<P3>(dx[mt][j], dx[mt][j], ga[mt], gl[mt], b, bl);
                    }
                }
            }
        }
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int j = 0; j < NJ; j++)
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int s = 8 * j + 2 * t + (q & 1);
                    const size_t px = px0 + pxl + 16 * mt + g + 8 * (q >> 1);
                    if (s < S && px < HW) {
                        float* dst = grad_sem + (size_t)s * HW + px;
                        const float v = dx[mt][j][q];
                        *dst = accumulate ? *dst + v : v;
                    }
                }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) s_part[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; w++) tot += s_part[w];
        atomicAdd(loss, tot);
    }
}

// One-pass variant for class counts whose logits fit a warp's registers (L <= 16 LT, e.g. Replica's 102 leaves): a warp
// owns ONE 16-pixel row tile, keeps all its logits (4 LT floats... 8 LT per thread) after the first contraction, and the
// softmax / gradient pass reuses them instead of recomputing Z: 24 instead of 36 MMAs per (row tile, 16 classes).
template <int NJ, bool P3, int LT>
__global__ void __launch_bounds__(32 * NWARP, 2) leaf_ce_pixel_onepass_kernel(
    const float* __restrict__ sem, const int* __restrict__ labels, const float* __restrict__ weight,
    const float* __restrict__ bias, int S, int L, size_t HW, float scale, float* __restrict__ loss,
    float* __restrict__ lse_out, float* __restrict__ grad_sem, int accumulate) {
    using C = Cfg<NJ, 1>;
    extern __shared__ float smem[];
    float* Xbuf = smem;                           // [2][SP][XS]
    float* Ws = smem + 2 * C::SP * C::XS;         // [16 LT][WS], resident
    __shared__ float s_part[NWARP];
    const bool vec_ok = (HW & 3) == 0 && (reinterpret_cast<size_t>(sem) & 15) == 0;
    stage_weights<NJ>(Ws, weight, bias, S, L, 0, 16 * LT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pxl = warp * 16;
    const int npairs = (L + 15) >> 4;
    const size_t n_blocks = (HW + C::BPX - 1) / C::BPX;
    float loss_acc = 0.f;
    int buf = 0;
    if (blockIdx.x < n_blocks) stage_pixels<NJ, 1>(Xbuf, sem, S, HW, (size_t)blockIdx.x * C::BPX, vec_ok);
    __pipeline_commit();
    for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, buf ^= 1) {
        const size_t px0 = blk * C::BPX;
        const float* Xs = Xbuf + buf * C::SP * C::XS;
        __pipeline_wait_prior(0);
        __syncthreads();
        if (blk + gridDim.x < n_blocks)
            stage_pixels<NJ, 1>(Xbuf + (buf ^ 1) * C::SP * C::XS, sem, S, HW, (blk + gridDim.x) * C::BPX, vec_ok);
        __pipeline_commit();
        int y[2];
        float sc[2], m[2] = {-1.0e30f, -1.0e30f}, zy[2] = {0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const size_t px = px0 + pxl + g + 8 * r;
            const int lab = px < HW ? labels[px] : -1;
            const bool use = lab >= 0 && lab < L;
            y[r] = use ? lab : -1;
            sc[r] = use ? scale : 0.f;
        }
        float z[LT][1][2][4];
#pragma unroll
        for (int np = 0; np < LT; np++) {
            if (np >= npairs) continue;                         // warp-uniform
            logits_pair<NJ, P3, true, 1>(z[np], Xs, Ws, pxl, np, g, t);
#pragma unroll
            for (int n = 0; n < 2; n++) {
                const int l0 = 16 * np + 8 * n + 2 * t;
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    if (l0 >= L) z[np][0][n][2 * r] = -3.0e38f;
                    if (l0 + 1 >= L) z[np][0][n][2 * r + 1] = -3.0e38f;
                    m[r] = fmaxf(m[r], fmaxf(z[np][0][n][2 * r], z[np][0][n][2 * r + 1]));
                    if (l0 == y[r]) zy[r] = z[np][0][n][2 * r];
                    if (l0 + 1 == y[r]) zy[r] = z[np][0][n][2 * r + 1];
                }
            }
        }
        float lse[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            m[r] = fmaxf(m[r], __shfl_xor_sync(0xffffffffu, m[r], 1));
            m[r] = fmaxf(m[r], __shfl_xor_sync(0xffffffffu, m[r], 2));
            zy[r] += __shfl_xor_sync(0xffffffffu, zy[r], 1);
            zy[r] += __shfl_xor_sync(0xffffffffu, zy[r], 2);
        }
        float sum[2] = {0.f, 0.f};
#pragma unroll
        for (int np = 0; np < LT; np++) {
            if (np >= npairs) continue;
#pragma unroll
            for (int n = 0; n < 2; n++)
#pragma unroll
                for (int q = 0; q < 4; q++) {                   // z becomes exp(z - max); masked classes give 0
                    z[np][0][n][q] = __expf(z[np][0][n][q] - m[q >> 1]);
                    sum[q >> 1] += z[np][0][n][q];
                }
        }
        float inv[2];
#pragma unroll
        for (int r = 0; r < 2; r++) {
            sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 1);
            sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], 2);
            lse[r] = m[r] + __logf(sum[r]);
            inv[r] = sc[r] / sum[r];
            const size_t px = px0 + pxl + g + 8 * r;
            if (t == 0 && px < HW) {
                lse_out[px] = lse[r];
                loss_acc += sc[r] * (lse[r] - zy[r]);
            }
        }
        float dx[NJ][4];
#pragma unroll
        for (int j = 0; j < NJ; j++) dx[j][0] = dx[j][1] = dx[j][2] = dx[j][3] = 0.f;
#pragma unroll
        for (int np = 0; np < LT; np++) {
            if (np >= npairs) continue;
#pragma unroll
            for (int n = 0; n < 2; n++) {
                const int l0 = 16 * np + 8 * n + 2 * t;
                float ga[4], gl[4];
#pragma unroll
                for (int r = 0; r < 2; r++) {                   // a0 = G(g, 2t), a1 = G(g+8, 2t), a2 = G(g, 2t+1), a3 = G(g+8, 2t+1)
                    ga[r] = z[np][0][n][2 * r] * inv[r] - (l0 == y[r] ? sc[r] : 0.f);
                    ga[2 + r] = z[np][0][n][2 * r + 1] * inv[r] - (l0 + 1 == y[r] ? sc[r] : 0.f);
                }
                if (P3) lo_parts(gl, ga);
                const float* w0 = Ws + (16 * np + 8 * n + 2 * t) * C::WS + g;
#pragma unroll
                for (int j = 0; j < NJ; j++) {
                    const float b[2] = {w0[8 * j], w0[C::WS + 8 * j]};
                    float bl[2];
                    if (P3) lo_parts(bl, b);
                    mma_p<P3>(dx[j], dx[j], ga, gl, b, bl);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NJ; j++)
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int s = 8 * j + 2 * t + (q & 1);
                const size_t px = px0 + pxl + g + 8 * (q >> 1);
                if (s < S && px < HW) {
                    float* dst = grad_sem + (size_t)s * HW + px;
                    *dst = accumulate ? *dst + dx[j][q] : dx[j][q];
                }
            }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
    if (lane == 0) s_part[warp] = loss_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NWARP; w++) tot += s_part[w];
        atomicAdd(loss, tot);
    }
}

template <int NJ, bool P3>
__global__ void __launch_bounds__(32 * NWARP, 2) leaf_ce_weight_kernel(
    const float* __restrict__ sem, const int* __restrict__ labels, const float* __restrict__ weight,
    const float* __restrict__ bias, const float* __restrict__ lse_in, int S, int L, size_t HW, float scale,
    float* __restrict__ grad_weight, float* __restrict__ grad_bias) {
    using C = Cfg<NJ>;
    constexpr int MT = C::MT, GS = C::GS;
    constexpr int NCH = C::LG / 16;               // 16-class chunks per CTA
    extern __shared__ float smem[];
    float* Xbuf = smem;                           // [2][SP][XS]  (reused as the dW reduction buffer [LG][SP] at the end)
    float* Ws = Xbuf + 2 * C::SP * C::XS;         // [LG][WS]
    float* Gt = Ws + C::LG * C::WS + (threadIdx.x >> 5) * MT * 16 * GS;   // [MT][16 classes][GS]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pxl = warp * 16 * MT;
    const int cls0 = blockIdx.y * C::LG;
    const bool vec_ok = (HW & 3) == 0 && (reinterpret_cast<size_t>(sem) & 15) == 0;
    const size_t n_blocks = (HW + C::BPX - 1) / C::BPX;
    float acc[NCH][NJ][4];
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int j = 0; j < NJ; j++) acc[c][j][0] = acc[c][j][1] = acc[c][j][2] = acc[c][j][3] = 0.f;
    stage_weights<NJ>(Ws, weight, bias, S, L, cls0, C::LG);
    int buf = 0;
    if (blockIdx.x < n_blocks) stage_pixels<NJ>(Xbuf, sem, S, HW, (size_t)blockIdx.x * C::BPX, vec_ok);
    __pipeline_commit();
    for (size_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, buf ^= 1) {
        const size_t px0 = blk * C::BPX;
        const float* Xs = Xbuf + buf * C::SP * C::XS;
        __pipeline_wait_prior(0);
        __syncthreads();
        if (blk + gridDim.x < n_blocks)
            stage_pixels<NJ>(Xbuf + (buf ^ 1) * C::SP * C::XS, sem, S, HW, (blk + gridDim.x) * C::BPX, vec_ok);
        __pipeline_commit();
        int y[MT][2];
        float sc[MT][2], lse[MT][2];
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                const size_t px = px0 + pxl + 16 * mt + g + 8 * r;
                const int lab = px < HW ? labels[px] : -1;
                const bool use = lab >= 0 && lab < L;
                y[mt][r] = use ? lab : -1;
                sc[mt][r] = use ? scale : 0.f;
                lse[mt][r] = px < HW ? lse_in[px] : 0.f;
            }
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            if (cls0 + 16 * c >= L) continue;                   // warp-uniform
            float z[MT][2][4];
            logits_pair<NJ, P3, false>(z, Xs, Ws, pxl, c, g, t);
#pragma unroll
            for (int n = 0; n < 2; n++) {
                const int l0 = cls0 + 16 * c + 8 * n + 2 * t;
#pragma unroll
                for (int mt = 0; mt < MT; mt++)
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const float pa = l0 < L ? __expf(z[mt][n][2 * r] - lse[mt][r]) : 0.f;
                        const float pb = l0 + 1 < L ? __expf(z[mt][n][2 * r + 1] - lse[mt][r]) : 0.f;
                        float* gt = Gt + mt * 16 * GS + (8 * n + 2 * t) * GS + g + 8 * r;
                        gt[0] = sc[mt][r] * (pa - (l0 == y[mt][r] ? 1.f : 0.f));
                        gt[GS] = sc[mt][r] * (pb - (l0 + 1 == y[mt][r] ? 1.f : 0.f));
                    }
            }
            __syncwarp();
            // dW[16 classes x SP] += G^T[16 classes x 16 px] X[16 px x SP], per row tile and 8-pixel k step
#pragma unroll
            for (int mt = 0; mt < MT; mt++)
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const float* ga = Gt + mt * 16 * GS + g * GS + 8 * kk + t;
                    const float a[4] = {ga[0], ga[8 * GS], ga[4], ga[8 * GS + 4]};
                    float al[4];
                    if (P3) lo_parts(al, a);
#pragma unroll
                    for (int j = 0; j < NJ; j++) {
                        const float* xb = Xs + (8 * j + g) * C::XS + pxl + 16 * mt + 8 * kk + t;
                        const float b[2] = {xb[0], xb[4]};
                        float bl[2];
                        if (P3) lo_parts(bl, b);
                        mma_p<P3>(acc[c][j], acc[c][j], a, al, b, bl);
                    }
                }
            __syncwarp();
        }
    }
    // ---- CTA reduction of the eight warps' partial blocks, then one atomic per (class, channel)
    __pipeline_wait_prior(0);
    __syncthreads();
    float* red = Xbuf;                            // [LG][SP]
    for (int idx = threadIdx.x; idx < C::LG * C::SP; idx += 32 * NWARP) red[idx] = 0.f;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NCH; c++)
#pragma unroll
        for (int j = 0; j < NJ; j++)
#pragma unroll
            for (int q = 0; q < 4; q++)
                atomicAdd(&red[(16 * c + g + 8 * (q >> 1)) * C::SP + 8 * j + 2 * t + (q & 1)], acc[c][j][q]);
    __syncthreads();
    for (int idx = threadIdx.x; idx < C::LG * C::SP; idx += 32 * NWARP) {
        const int l = cls0 + idx / C::SP, s = idx % C::SP;
        if (l >= L) continue;
        if (s < S) atomicAdd(&grad_weight[(size_t)l * S + s], red[idx]);
        else if (s == S && grad_bias != nullptr) atomicAdd(&grad_bias[l], red[idx]);
    }
}

// pixel_pass = false: only the weight-gradient kernel (lse comes from an earlier pixel pass, e.g. the tcgen05 kernel)
template <int NJ, bool P3>
static int launch_t(const float* sem, const int* labels, const float* weight, const float* bias, int S, int L, size_t HW,
                    float scale, float* loss, float* lse, float* grad_sem, int accumulate, float* grad_weight,
                    float* grad_bias, cudaStream_t stream, bool pixel_pass = true) {
    using C = Cfg<NJ>;
    const size_t n_blocks = (HW + C::BPX - 1) / C::BPX;
    constexpr int LT = 7;                     // one-pass variant: up to 112 classes
    if (!pixel_pass) {
    } else if (NJ <= 4 && L <= 16 * LT) {
        using C1 = Cfg<NJ, 1>;
        auto k = leaf_ce_pixel_onepass_kernel<NJ, P3, (NJ <= 4 ? LT : 1)>;
        const size_t sh = (size_t)(2 * C1::SP * C1::XS + 16 * LT * C1::WS) * sizeof(float);
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        const int grid = (int)min((size_t)148 * 2, (HW + C1::BPX - 1) / C1::BPX);
        k<<<grid, 32 * NWARP, sh, stream>>>(sem, labels, weight, bias, S, L, HW, scale, loss, lse, grad_sem, accumulate);
        HS_LAUNCH_OK(stream, false);
    } else {
        auto k = leaf_ce_pixel_kernel<NJ, P3>;
        const size_t sh = (size_t)(2 * C::SP * C::XS + C::LC * C::WS) * sizeof(float);
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        const int grid = (int)min((size_t)148 * C::OCC, n_blocks);
        k<<<grid, 32 * NWARP, sh, stream>>>(sem, labels, weight, bias, S, L, HW, scale, loss, lse, grad_sem, accumulate);
        HS_LAUNCH_OK(stream, false);
    }
    if (grad_weight != nullptr) {
        auto k = leaf_ce_weight_kernel<NJ, P3>;
        const size_t sh = (size_t)(2 * C::SP * C::XS + C::LG * C::WS + NWARP * C::MT * 16 * C::GS) * sizeof(float);
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        const int groups = (L + C::LG - 1) / C::LG;
        int gx = (148 * 2 + groups - 1) / groups;
        if ((size_t)gx > n_blocks) gx = (int)n_blocks;
        k<<<dim3(gx, groups), 32 * NWARP, sh, stream>>>(sem, labels, weight, bias, lse, S, L, HW, scale, grad_weight,
                                                         grad_bias);
        HS_LAUNCH_OK(stream, false);
    }
    return 0;
}

}  // namespace leaf

int launch_leaf_weight_grad(const float* sem, const int* labels, const float* weight, const float* bias, const float* lse,
                            int S, int L, size_t HW, float scale, float* grad_weight, float* grad_bias, int single_tf32,
                            cudaStream_t stream) {
    if (HW == 0 || L <= 0 || grad_weight == nullptr) return 0;
    if (S < 1 || S > 79) {
        set_error("leaf cross-entropy: 1 <= S <= 79 semantic channels supported, got %d", S);
        return 1;
    }
    const int nj = (S + 1 + 7) / 8;
    float* l = const_cast<float*>(lse);
#define HS_LEAFW_CASE(NJV)                                                                                              \
    return single_tf32 ? leaf::launch_t<NJV, false>(sem, labels, weight, bias, S, L, HW, scale, nullptr, l, nullptr, 0,  \
                                                     grad_weight, grad_bias, stream, false)                              \
                       : leaf::launch_t<NJV, true>(sem, labels, weight, bias, S, L, HW, scale, nullptr, l, nullptr, 0,   \
                                                    grad_weight, grad_bias, stream, false)
    if (nj <= 3) { HS_LEAFW_CASE(3); }
    if (nj <= 4) { HS_LEAFW_CASE(4); }
    HS_LEAFW_CASE(10);
#undef HS_LEAFW_CASE
}

int launch_leaf_cross_entropy(const float* sem, const int* labels, const float* weight, const float* bias, int S, int L,
                              size_t HW, float scale, float* loss, float* lse, float* grad_sem, int accumulate,
                              float* grad_weight, float* grad_bias, int single_tf32, cudaStream_t stream) {
    if (HW == 0 || L <= 0) return 0;
    if (S < 1 || S > 79) {
        set_error("leaf cross-entropy: 1 <= S <= 79 semantic channels supported, got %d", S);
        return 1;
    }
    const int nj = (S + 1 + 7) / 8;           // S channels + the bias channel, in 8-wide k steps
#define HS_LEAF_CASE(NJV)                                                                                               \
    return single_tf32 ? leaf::launch_t<NJV, false>(sem, labels, weight, bias, S, L, HW, scale, loss, lse, grad_sem,      \
                                                     accumulate, grad_weight, grad_bias, stream)                          \
                       : leaf::launch_t<NJV, true>(sem, labels, weight, bias, S, L, HW, scale, loss, lse, grad_sem,       \
                                                    accumulate, grad_weight, grad_bias, stream)
    if (nj <= 3) { HS_LEAF_CASE(3); }
    if (nj <= 4) { HS_LEAF_CASE(4); }
    HS_LEAF_CASE(10);
#undef HS_LEAF_CASE
}

}  // namespace hs
