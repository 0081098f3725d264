// Parameter maintenance on the flat parameter buffer.
//
// adam_flat_kernel   torch.optim.Adam's update (scripts/hierslam.py:411-417: per-name learning rates, eps = 1e-15 in
//                    mapping) for ALL parameter tensors in one pass over the flat buffers of
//                    hier_slam_b200.mapping.FlatParams: 16 B read + 12 B written per element, where torch's
//                    multi-tensor path runs seven elementwise kernels over every tensor.  The arithmetic follows
//                    torch/optim/adam.py::_multi_tensor_adam operation by operation (lerp, mul, addcmul, sqrt, div, add,
//                    addcdiv with float scalars), so the result is the same up to the contraction choices of the
//                    compiler (tests compare against torch.optim.Adam).
// compaction         Gaussian pruning (utils/slam_external.py:142-164 remove_points: `tensor[to_keep]` for every
//                    parameter and both Adam moments, each a nonzero() + gather with a host sync): one keep-mask scan
//                    (block counts -> block offsets -> ordered source-row list) and one gather per buffer over all
//                    segments.  Order-preserving, bit-exact.
#include "hs_common.cuh"

namespace hs {

struct AdamSegments {
    int n;
    unsigned long long end[HS_MAX_SEGMENTS];      // exclusive end (in floats) of each segment, ascending
    float neg_step_size[HS_MAX_SEGMENTS];         // -(lr / bias_correction1), rounded to float like torch's scalar
};

__global__ void __launch_bounds__(256) adam_flat_kernel(float4* __restrict__ param, const float4* __restrict__ grad,
                                                        float4* __restrict__ exp_avg, float4* __restrict__ exp_avg_sq,
                                                        size_t n4, AdamSegments seg, float w1, float beta2, float w2,
                                                        float bc2_sqrt, float eps, const float* __restrict__ dev_scalars) {
    // dev_scalars (device-side step counter, hs_adam_step_device): [segment] -> -(lr / bias_correction1), [HS_MAX_SEGMENTS]
    // -> sqrt(bias_correction2), written by adam_scalars_kernel just before this launch
    if (dev_scalars != nullptr) bc2_sqrt = dev_scalars[HS_MAX_SEGMENTS];
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
        int s = 0;
        while (s < seg.n - 1 && 4 * i >= seg.end[s]) s++;
        const float nss = dev_scalars != nullptr ? dev_scalars[s] : seg.neg_step_size[s];
        const float4 g4 = grad[i];
        float4 p4 = param[i], m4 = exp_avg[i], v4 = exp_avg_sq[i];
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        float p[4] = {p4.x, p4.y, p4.z, p4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            m[q] = __fmaf_rn(w1, __fsub_rn(g[q], m[q]), m[q]);                       // exp_avg.lerp_(grad, 1 - beta1)
            v[q] = __fmul_rn(v[q], beta2);                                           // exp_avg_sq.mul_(beta2)
            v[q] = __fmaf_rn(w2, __fmul_rn(g[q], g[q]), v[q]);                       // .addcmul_(grad, grad, 1 - beta2)
            const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v[q]), bc2_sqrt), eps);
            p[q] = __fmaf_rn(nss, __fdiv_rn(m[q], denom), p[q]);                     // param.addcdiv_(exp_avg, denom, -step_size)
        }
        param[i] = make_float4(p[0], p[1], p[2], p[3]);
        exp_avg[i] = make_float4(m[0], m[1], m[2], m[3]);
        exp_avg_sq[i] = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// step counter on the device (CUDA-graph capturable optimiser step): ++step, then the bias-correction scalars in double like
// the host path / torch's Python scalars
struct AdamLrs {
    double lr[HS_MAX_SEGMENTS];
};
__global__ void adam_scalars_kernel(int* __restrict__ step_counter, float* __restrict__ scalars, AdamLrs lrs, int n_seg,
                                    double beta1, double beta2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int step = step_counter[0] + 1;
    step_counter[0] = step;
    const double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
    for (int s = 0; s < n_seg; s++) scalars[s] = (float)((lrs.lr[s] / bc1) * -1.0);
    scalars[HS_MAX_SEGMENTS] = (float)sqrt(bc2);
}

int launch_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, size_t n, int n_seg,
                     const unsigned long long* seg_end, const double* seg_lr, double beta1, double beta2, double eps, int step,
                     cudaStream_t stream, int* step_counter, float* dev_scalars) {
    if (n == 0) return 0;
    if (step_counter != nullptr) step = 1;      // the real count lives on the device
    if (n_seg < 1 || n_seg > HS_MAX_SEGMENTS || (n & 3) != 0 || step < 1) {
        set_error("adam: 1..%d segments, a multiple of 4 elements and step >= 1 required", HS_MAX_SEGMENTS);
        return 1;
    }
    // the scalars exactly as torch/optim/adam.py computes them in Python doubles before they reach the float kernels
    const double b1 = beta1, b2 = beta2;
    const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
    AdamSegments seg;
    seg.n = n_seg;
    for (int s = 0; s < n_seg; s++) {
        if ((seg_end[s] & 3) != 0) {
            set_error("adam: segment boundaries must be multiples of 4 floats");
            return 1;
        }
        seg.end[s] = seg_end[s];
        seg.neg_step_size[s] = (float)((seg_lr[s] / bc1) * -1.0);
    }
    const size_t n4 = n / 4;
    int blocks = (int)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (step_counter != nullptr) {
        if (dev_scalars == nullptr) {
            set_error("adam: the device-side step counter needs a scalar scratch of %d floats", HS_MAX_SEGMENTS + 1);
            return 1;
        }
        AdamLrs lrs;
        for (int s = 0; s < HS_MAX_SEGMENTS; s++) lrs.lr[s] = s < n_seg ? seg_lr[s] : 0.0;
        adam_scalars_kernel<<<1, 32, 0, stream>>>(step_counter, dev_scalars, lrs, n_seg, b1, b2);
        HS_LAUNCH_OK(stream, false);
    }
    adam_flat_kernel<<<blocks, 256, 0, stream>>>((float4*)param, (const float4*)grad, (float4*)exp_avg, (float4*)exp_avg_sq, n4,
                                                 seg, (float)(1.0 - b1), (float)b2, (float)(1.0 - b2), (float)sqrt(bc2), (float)eps,
                                                 step_counter != nullptr ? dev_scalars : nullptr);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

// ---- keep-mask compaction -----------------------------------------------------------------------------------------
constexpr int CB = 1024;      // rows per block of the scan kernels

__global__ void __launch_bounds__(256) keep_count_kernel(const uint8_t* __restrict__ keep, int P, unsigned* __restrict__ counts) {
    __shared__ unsigned s_part[8];
    const int base = blockIdx.x * CB;
    unsigned c = 0;
#pragma unroll
    for (int q = 0; q < CB / 256; q++) {
        const int i = base + q * 256 + threadIdx.x;
        c += (i < P && keep[i] != 0) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) t += s_part[w];
        counts[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block counts in place; total -> counts[nb]
__global__ void __launch_bounds__(1024) keep_offsets_kernel(unsigned* __restrict__ counts, int nb) {
    __shared__ unsigned s_warp[32];
    __shared__ unsigned s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? counts[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const unsigned incl = x + ((threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0u) + s_carry;
        if (i < nb) counts[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) counts[nb] = s_carry;
}

// src_row[new index] = old index, in ascending old-index order (what tensor[mask] produces)
__global__ void __launch_bounds__(256) keep_rows_kernel(const uint8_t* __restrict__ keep, int P,
                                                        const unsigned* __restrict__ offsets, unsigned* __restrict__ src_row) {
    __shared__ unsigned s_warp[8];
    const int base = blockIdx.x * CB;
    unsigned run = offsets[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < CB / 256; q++) {
        const int i = base + q * 256 + threadIdx.x;
        const bool k = i < P && keep[i] != 0;
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; w++) {
            before += w < warp ? s_warp[w] : 0u;
            total += s_warp[w];
        }
        if (k) src_row[run + before + __popc(bal & ((1u << lane) - 1u))] = (unsigned)i;
        run += total;
        __syncthreads();
    }
}

struct RowSegments {
    int n;
    unsigned long long src_off[HS_MAX_SEGMENTS], dst_off[HS_MAX_SEGMENTS];   // in floats
    int width[HS_MAX_SEGMENTS];                                              // floats per row
};

// a row is spread over wp = 2^shift (<= 32) neighbouring threads, so there is no division and rows are read / written
// with consecutive addresses
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                          const unsigned* __restrict__ src_row, int rows, RowSegments seg) {
    for (int s = 0; s < seg.n; s++) {
        const int w = seg.width[s];
        int shift = 0;
        while ((1 << shift) < w && shift < 5) shift++;
        const int wp = 1 << shift, rpb = 256 >> shift;
        const float* in = src + seg.src_off[s];
        float* out = dst + seg.dst_off[s];
        for (size_t r = (size_t)blockIdx.x * rpb + (threadIdx.x >> shift); r < (size_t)rows; r += (size_t)gridDim.x * rpb) {
            const size_t from = (size_t)src_row[r] * w, to = r * w;
            for (int c = threadIdx.x & (wp - 1); c < w; c += wp) out[to + c] = in[from + c];
        }
    }
}

size_t compact_scratch_bytes(int P) {
    const size_t nb = ((size_t)P + CB - 1) / CB;
    return (nb + 1 + (size_t)P) * sizeof(unsigned);
}

// scratch: [nb + 1] block offsets (last = number of kept rows) followed by [P] source rows
int launch_compact_plan(const uint8_t* keep, int P, unsigned* scratch, cudaStream_t stream) {
    if (P <= 0) return 0;
    const int nb = (P + CB - 1) / CB;
    keep_count_kernel<<<nb, 256, 0, stream>>>(keep, P, scratch);
    HS_LAUNCH_OK(stream, false);
    keep_offsets_kernel<<<1, 1024, 0, stream>>>(scratch, nb);
    HS_LAUNCH_OK(stream, false);
    keep_rows_kernel<<<nb, 256, 0, stream>>>(keep, P, scratch, scratch + nb + 1);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

int launch_compact_gather(const float* src, float* dst, const unsigned* scratch, int P, int rows, int n_seg,
                          const unsigned long long* src_off, const unsigned long long* dst_off, const int* width,
                          cudaStream_t stream) {
    if (rows <= 0 || n_seg <= 0) return 0;
    if (n_seg > HS_MAX_SEGMENTS) {
        set_error("compaction: at most %d segments", HS_MAX_SEGMENTS);
        return 1;
    }
    RowSegments seg;
    seg.n = n_seg;
    size_t most = 0;
    for (int s = 0; s < n_seg; s++) {
        seg.src_off[s] = src_off[s];
        seg.dst_off[s] = dst_off[s];
        seg.width[s] = width[s];
        most = max(most, (size_t)rows * width[s]);
    }
    const int nb = (P + CB - 1) / CB;
    int blocks = (int)((most + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    gather_rows_kernel<<<blocks, 256, 0, stream>>>(src, dst, scratch + nb + 1, rows, seg);
    HS_LAUNCH_OK(stream, false);
    return 0;
}

}  // namespace hs
