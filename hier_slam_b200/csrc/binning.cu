// Binning: offsets scan, (tile | depth) key duplication, key sort, tile ranges, and the opaque state-buffer
// layouts.  Reference behaviour: cuda_rasterizer/rasterizer_impl.cu:70-138 (duplicateWithKeys,
// identifyTileRanges), :544-585 (scan, sort, memset) and :155-194 (state chunks).
//
// Parity contract: unsorted keys/values, sorted keys, point_list and ranges are bit-exact with the reference
// (stable ascending sort on key bits [0, 32 + getHigherMsb(tiles)) keeps equal keys in Gaussian order).
#include "hs_common.cuh"
#include <cub/cub.cuh>
#include <stdarg.h>
#include <mutex>
#include <string>
#include <vector>

namespace hs {

static thread_local std::string g_last_error;
void count_lib_call();
void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}
const char* last_error() { return g_last_error.c_str(); }

// ---- profiling / launch accounting ---------------------------------------------------------------------
// Every kernel launch made while profiling is on gets its own CUDA event pair on the launching stream;
// hs_profile_read() waits for all of them and returns per-stage totals and counts, then recycles the events.
struct ProfRec {
    int stage;
    cudaEvent_t e0, e1;
};
static bool g_prof_on = false;
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_recs;
static std::vector<ProfRec> g_free;

static long long g_launches = 0;      // kernels written in this library
static long long g_lib_calls = 0;     // CUB device-wide primitives (scan, sort)
void count_launch(int n) { g_launches += n; }
void count_lib_call() { g_lib_calls += 1; }
long long launches() { return g_launches; }
long long lib_calls() { return g_lib_calls; }
void prof_enable(bool on) { g_prof_on = on; }
void prof_begin(int stage, cudaStream_t stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    if (!g_free.empty()) {
        r = g_free.back();
        g_free.pop_back();
    } else {
        cudaEventCreate(&r.e0);
        cudaEventCreate(&r.e1);
    }
    r.stage = stage;
    cudaEventRecord(r.e0, stream);
    g_recs.push_back(r);

}
void prof_end(int stage, cudaStream_t stream) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = (int)g_recs.size() - 1; i >= 0; i--) {
        if (g_recs[i].stage == stage) {
            cudaEventRecord(g_recs[i].e1, stream);
            break;
        }
    }
}
int prof_read(float* total_ms, int* count) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int i = 0; i < ST_COUNT; i++) {
        total_ms[i] = 0.f;
        count[i] = 0;
    }
    int rc = 0;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.e1) != cudaSuccess || cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) {
            rc = 2;
            cudaGetLastError();
        } else {
            total_ms[r.stage] += ms;
            count[r.stage] += 1;
        }
        g_free.push_back(r);
    }
    g_recs.clear();
    return rc;
}

// ---- state layouts ------------------------------------------------------------------------------------
template <typename T>
static void carve(char*& p, T*& ptr, size_t count) {
    size_t off = align_up(reinterpret_cast<size_t>(p));
    ptr = reinterpret_cast<T*>(off);
    p = reinterpret_cast<char*>(ptr + count);
}

int geom_view(char* base, size_t P, GeomView* v) {
    char* p = base;
    carve(p, v->depths, P);
    carve(p, v->means2D, P);
    carve(p, v->conic_opacity, P);
    carve(p, v->tiles_touched, P);
    carve(p, v->point_offsets, P);
    carve(p, v->rgb, 3 * P);
    carve(p, v->clamped, P);
    v->scan_temp_bytes = 0;
    if (P > 0) {
        HS_CUDA_OK(cub::DeviceScan::InclusiveSum(nullptr, v->scan_temp_bytes, v->tiles_touched, v->point_offsets, (int)P));
    }
    carve(p, v->scan_temp, v->scan_temp_bytes);
    v->total_bytes = (size_t)(p - base) + HS_ALIGN;
    v->rows = reinterpret_cast<float*>(base + align_up(v->total_bytes));
    return 0;
}

int image_view(char* base, size_t N, size_t tiles, ImageView* v) {
    char* p = base;
    carve(p, v->final_T, N);
    carve(p, v->n_contrib, N);
    carve(p, v->ranges, tiles);
    carve(p, v->tile_count, tiles * HS_CTR_STRIDE);
    carve(p, v->info, 4);
    v->total_bytes = (size_t)(p - base) + HS_ALIGN;
    return 0;
}

int binning_view(char* base, size_t R, BinningView* v) {
    char* p = base;
    carve(p, v->point_list, R);
    carve(p, v->point_list_unsorted, R);
    carve(p, v->keys, R);
    carve(p, v->keys_unsorted, R);
    carve(p, v->strip_hits, R);
    v->sort_temp_bytes = 0;
    if (R > 0) {
        HS_CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, v->sort_temp_bytes, v->keys_unsorted, v->keys,
                                                   v->point_list_unsorted, v->point_list, (int)R));
    }
    carve(p, v->sort_temp, v->sort_temp_bytes);
    v->total_bytes = (size_t)(p - base) + HS_ALIGN;
    return 0;
}

int launch_scan(int P, const GeomView& g, cudaStream_t stream, bool debug) {
    if (P <= 0) return 0;
    size_t tb = g.scan_temp_bytes;
    prof_begin(ST_SCAN, stream);
    HS_CUDA_OK(cub::DeviceScan::InclusiveSum(g.scan_temp, tb, g.tiles_touched, g.point_offsets, P, stream));
    prof_end(ST_SCAN, stream);
    count_lib_call();
    HS_CUDA_OK(cudaGetLastError());
    if (debug) HS_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

// rasterizer_impl.cu:35-50
static uint32_t higher_msb(uint32_t n) {
    uint32_t msb = sizeof(n) * 4;
    uint32_t step = msb;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step;
        else msb -= step;
    }
    if (n >> msb) msb++;
    return msb;
}

__device__ __forceinline__ void get_rect_dev(const float2 p, int max_radius, uint2& rect_min, uint2& rect_max,
                                             unsigned gx, unsigned gy) {  // auxiliary.h:46-56
    rect_min = {min(gx, (unsigned)max((int)0, (int)((p.x - max_radius) / HS_TILE_X))),
                min(gy, (unsigned)max((int)0, (int)((p.y - max_radius) / HS_TILE_Y)))};
    rect_max = {min(gx, (unsigned)max((int)0, (int)((p.x + max_radius + HS_TILE_X - 1) / HS_TILE_X))),
                min(gy, (unsigned)max((int)0, (int)((p.y + max_radius + HS_TILE_Y - 1) / HS_TILE_Y)))};
}

// One thread per Gaussian like the reference, but a Gaussian that covers many tiles is expanded by its
// whole warp (coalesced 8-byte / 4-byte stores) instead of by one serial thread.  Output order is identical:
// instance n of Gaussian idx sits at offsets[idx-1] + n, y-major / x-minor inside the rect.
__global__ void __launch_bounds__(256) duplicate_kernel(int P, const float2* __restrict__ points_xy,
                                                        const float* __restrict__ depths,
                                                        const uint32_t* __restrict__ offsets,
                                                        uint64_t* __restrict__ keys_unsorted,
                                                        uint32_t* __restrict__ values_unsorted,
                                                        const int* __restrict__ radii, unsigned gx, unsigned gy) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t off = 0, w = 0, count = 0, dbits = 0;
    uint2 rmin = {0, 0}, rmax = {0, 0};
    if (idx < P && radii[idx] > 0) {
        off = (idx == 0) ? 0 : offsets[idx - 1];
        get_rect_dev(points_xy[idx], radii[idx], rmin, rmax, gx, gy);
        w = rmax.x - rmin.x;
        count = w * (rmax.y - rmin.y);
        dbits = __float_as_uint(depths[idx]);
    }
    const uint32_t kSmall = 4;
    if (count > 0 && count <= kSmall) {
        for (uint32_t n = 0; n < count; n++) {
            uint32_t y = rmin.y + n / w, x = rmin.x + n % w;
            uint64_t key = y * gx + x;
            key <<= 32;
            key |= dbits;
            keys_unsorted[off + n] = key;
            values_unsorted[off + n] = idx;
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, count > kSmall);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t s_off = __shfl_sync(0xffffffffu, off, src);
        const uint32_t s_w = __shfl_sync(0xffffffffu, w, src);
        const uint32_t s_count = __shfl_sync(0xffffffffu, count, src);
        const uint32_t s_db = __shfl_sync(0xffffffffu, dbits, src);
        const uint32_t s_x0 = __shfl_sync(0xffffffffu, rmin.x, src);
        const uint32_t s_y0 = __shfl_sync(0xffffffffu, rmin.y, src);
        const uint32_t s_idx = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) + src;
        for (uint32_t n = lane; n < s_count; n += 32) {
            uint32_t y = s_y0 + n / s_w, x = s_x0 + n % s_w;
            uint64_t key = y * gx + x;
            key <<= 32;
            key |= s_db;
            keys_unsorted[s_off + n] = key;
            values_unsorted[s_off + n] = s_idx;
        }
    }
}

// rasterizer_impl.cu:116-138
__global__ void tile_ranges_kernel(int L, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= L) return;
    const uint32_t currtile = keys[idx] >> 32;
    if (idx == 0) ranges[currtile].x = 0;
    else {
        const uint32_t prevtile = keys[idx - 1] >> 32;
        if (currtile != prevtile) {
            ranges[prevtile].y = idx;
            ranges[currtile].x = idx;
        }
    }
    if (idx == L - 1) ranges[currtile].y = L;
}

// ---- tile-bucket binning (default path) ----------------------------------------------------------------------
// The reference sorts all R (tile | depth) keys globally (6 radix passes over 12 R bytes each).  The same sorted list
// is obtained with one counting pass and one short sort per tile:
//   1. preprocess_kernel counts the instances of every tile (tile_count, atomics);
//   2. tile_scan_kernel turns the counts into [start, end) ranges (== identifyTileRanges, empty tiles keep (0,0)),
//      scatter cursors, num_rendered and the longest tile list;
//   3. scatter_kernel writes (depth bits << 32 | Gaussian id) of every instance into its tile's segment, in arbitrary
//      order (cursor atomics);
//   4. tile_sort_kernel sorts each segment in shared memory on the whole 64-bit word.  Depths are positive floats, so
//      their bit patterns order like the values; the id in the low word reproduces the tie order of the reference's
//      STABLE sort over a list that is generated in Gaussian order.  All words of a segment are distinct, so the
//      result does not depend on the scatter order: point_list and the sorted keys are bit-identical to the
//      reference's (tests/test_gpu_parity.py).
// Capacity mode (r_cap > 0, HS_ASYNC_BINNING): the host does not read the counts back but has sized the binning buffer
// for r_cap instances and the per-tile sort for lists of at most tile_cap entries.  If the frame needs more, info[3] is
// raised, every range is cleared (the frame renders empty, memory-safely) and the caller repeats the call synchronously.
__global__ void __launch_bounds__(1024) tile_scan_kernel(int tiles, uint32_t* __restrict__ tile_count,
                                                         uint2* __restrict__ ranges, uint32_t* __restrict__ info,
                                                         uint32_t r_cap, uint32_t tile_cap) {
    constexpr int IT = 4;   // consecutive tiles per thread: 4096 tiles per sweep, their loads are independent
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_small[32];
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_overflow;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    uint32_t my_max = 0, my_small = 0;
    __syncthreads();
    for (int base = 0; base < tiles; base += 1024 * IT) {
        const int t0 = base + tid * IT;
        uint32_t c[IT];
#pragma unroll
        for (int k = 0; k < IT; k++) c[k] = (t0 + k < tiles) ? tile_count[(size_t)(t0 + k) * HS_CTR_STRIDE] : 0;
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            sum += c[k];
            my_max = max(my_max, c[k]);
            my_small += (c[k] > 0 && c[k] <= HS_TILE_SORT_SMALL) ? 1u : 0u;
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += v;
            }
            s_warp[lane] = w;   // inclusive scan of the warp totals
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        uint32_t start = carry + (warp > 0 ? s_warp[warp - 1] : 0) + incl - sum;
#pragma unroll
        for (int k = 0; k < IT; k++) {
            if (t0 + k < tiles) {
                ranges[t0 + k] = c[k] > 0 ? make_uint2(start, start + c[k]) : make_uint2(0u, 0u);
                tile_count[(size_t)(t0 + k) * HS_CTR_STRIDE] = start;   // scatter cursor
            }
            start += c[k];
        }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
        my_small += __shfl_xor_sync(0xffffffffu, my_small, o);
    }
    if (lane == 0) {
        s_warp[warp] = my_max;
        s_small[warp] = my_small;
    }
    __syncthreads();
    if (tid == 0) {
        uint32_t m = 0, ns = 0;
        for (int w = 0; w < 32; w++) {
            m = max(m, s_warp[w]);
            ns += s_small[w];
        }
        info[0] = s_carry;   // num_rendered
        info[1] = m;         // longest tile list
        info[2] = ns;        // non-empty tiles with at most HS_TILE_SORT_SMALL entries
        s_overflow = (r_cap > 0 && (s_carry > r_cap || m > tile_cap)) ? 1u : 0u;
        info[3] = s_overflow;
    }
    __syncthreads();
    if (s_overflow)
        for (int t = tid; t < tiles; t += 1024) ranges[t] = make_uint2(0u, 0u);
}

int launch_tile_scan(const Camera& cam, const ImageView& img, uint32_t r_cap, uint32_t tile_cap, cudaStream_t stream,
                     bool debug) {
    prof_begin(ST_SCAN, stream);
    tile_scan_kernel<<<1, 1024, 0, stream>>>(cam.grid_x * cam.grid_y, img.tile_count, img.ranges, img.info, r_cap,
                                             tile_cap);
    prof_end(ST_SCAN, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

__global__ void __launch_bounds__(256) scatter_kernel(int P, const float2* __restrict__ points_xy,
                                                      const float* __restrict__ depths,
                                                      uint32_t* __restrict__ cursor, uint64_t* __restrict__ seg,
                                                      const int* __restrict__ radii, unsigned gx, unsigned gy,
                                                      const uint32_t* __restrict__ info) {
    if (info[3] != 0) return;   // capacity mode: the frame does not fit the binning buffer (tile_scan_kernel)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t w = 0, count = 0, dbits = 0;
    uint2 rmin = {0, 0}, rmax = {0, 0};
    if (idx < P && radii[idx] > 0) {
        get_rect_dev(points_xy[idx], radii[idx], rmin, rmax, gx, gy);
        w = rmax.x - rmin.x;
        count = w * (rmax.y - rmin.y);
        dbits = __float_as_uint(depths[idx]);
    }
    const uint32_t kSmall = 4;
    if (count > 0 && count <= kSmall) {
        const uint64_t word = ((uint64_t)dbits << 32) | (uint32_t)idx;
        for (uint32_t n = 0; n < count; n++) {
            const uint32_t tile = (rmin.y + n / w) * gx + rmin.x + n % w;
            seg[atomicAdd(cursor + (size_t)tile * HS_CTR_STRIDE, 1u)] = word;
        }
    }
    unsigned big = __ballot_sync(0xffffffffu, count > kSmall);
    while (big) {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        const uint32_t s_w = __shfl_sync(0xffffffffu, w, src);
        const uint32_t s_count = __shfl_sync(0xffffffffu, count, src);
        const uint32_t s_db = __shfl_sync(0xffffffffu, dbits, src);
        const uint32_t s_x0 = __shfl_sync(0xffffffffu, rmin.x, src);
        const uint32_t s_y0 = __shfl_sync(0xffffffffu, rmin.y, src);
        const uint32_t s_idx = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) + src;
        const uint64_t word = ((uint64_t)s_db << 32) | s_idx;
        for (uint32_t n = lane; n < s_count; n += 32) {
            const uint32_t tile = (s_y0 + n / s_w) * gx + s_x0 + n % s_w;
            seg[atomicAdd(cursor + (size_t)tile * HS_CTR_STRIDE, 1u)] = word;
        }
    }
}

// One CTA per tile; handles the tiles whose list length n satisfies lo < n <= hi (the host launches one size class
// with small CTAs and, only if longer lists exist, a second one with 512 threads and a large shared buffer).
// Bitonic network on 64-bit words, padded to a power of two with ~0.  A thread keeps 2^Q words in registers and runs
// Q consecutive strides of the network on them per shared-memory round trip (the words of one item are closed under
// those strides), which cuts the shared-memory traffic and the barriers by ~Q compared with one stride per pass.
// Every phase ends with one pass over strides 8, 4, 2, 1 (16 consecutive words per thread); the passes above it have
// strides >= 16, where consecutive lanes touch consecutive words.  With one pad word per 16 both access patterns are
// free of bank conflicts.
__device__ __forceinline__ uint32_t sort_pad(uint32_t i) { return i + (i >> 4); }

__device__ __forceinline__ void cmpx(uint64_t& a, uint64_t& b, bool up) {
    const bool sw = (a > b) == up;
    const uint64_t x = sw ? b : a, y = sw ? a : b;
    a = x;
    b = y;
}

// strides jh, jh/2, ..., jh >> (Q-1) of phase k
template <int Q>
__device__ __forceinline__ void bitonic_pass(uint64_t* s, uint32_t n2, uint32_t k, uint32_t jh, int tid, int T) {
    constexpr int E = 1 << Q;
    const uint32_t sl = jh >> (Q - 1);
    const int lsl = 31 - __clz(sl);
    for (uint32_t u = tid; u < (n2 >> Q); u += T) {
        const uint32_t base = ((u >> lsl) << (lsl + Q)) | (u & (sl - 1));
        const bool up = (base & k) == 0;
        uint64_t v[E];
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = s[sort_pad(base + e * sl)];
#pragma unroll
        for (int st = Q - 1; st >= 0; st--)
#pragma unroll
            for (int e = 0; e < E; e++)
                if (!(e & (1 << st))) cmpx(v[e], v[e | (1 << st)], up);
#pragma unroll
        for (int e = 0; e < E; e++) s[sort_pad(base + e * sl)] = v[e];
    }
}

// phases k = 2, 4, 8, 16 on 16 consecutive words per thread
__device__ __forceinline__ void bitonic_first16(uint64_t* s, uint32_t n2, int tid, int T) {
    for (uint32_t u = tid; u < (n2 >> 4); u += T) {
        const uint32_t base = u << 4;
        uint64_t v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = s[sort_pad(base + e)];
#pragma unroll
        for (int k = 2; k <= 16; k <<= 1)
#pragma unroll
            for (int st = k >> 1; st > 0; st >>= 1)
#pragma unroll
                for (int e = 0; e < 16; e++)
                    if (!(e & st)) cmpx(v[e], v[e | st], ((base + e) & k) == 0);
#pragma unroll
        for (int e = 0; e < 16; e++) s[sort_pad(base + e)] = v[e];
    }
}

#ifndef HS_SORT_T_SMALL
#define HS_SORT_T_SMALL 128   // threads per CTA for tile lists of up to HS_TILE_SORT_SMALL entries
#endif

template <int T>
__global__ void __launch_bounds__(T) tile_sort_kernel(const uint2* __restrict__ ranges, const uint64_t* __restrict__ seg,
                                                      uint32_t* __restrict__ point_list, uint64_t* __restrict__ keys,
                                                      uint32_t lo, uint32_t hi) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
    const uint32_t tile = blockIdx.x;
    const uint2 range = ranges[tile];
    const uint32_t n = range.y - range.x;
    if (n <= lo || n > hi) return;
    uint32_t n2 = 16;
    while (n2 < n) n2 <<= 1;
    const int tid = threadIdx.x;
    for (uint32_t i = tid; i < n2; i += T) s[sort_pad(i)] = i < n ? seg[range.x + i] : ~0ull;
    __syncthreads();
    bitonic_first16(s, n2, tid, T);
    __syncthreads();
    int m = 5;   // log2(k)
    for (uint32_t k = 32; k <= n2; k <<= 1, m++) {
        uint32_t j = k >> 1;            // strides k/2 ... 16 in groups of <= 3, then 8 4 2 1
        const int rem = (m - 4) % 3;
        if (rem == 1) {
            bitonic_pass<1>(s, n2, k, j, tid, T);
            j >>= 1;
            __syncthreads();
        } else if (rem == 2) {
            bitonic_pass<2>(s, n2, k, j, tid, T);
            j >>= 2;
            __syncthreads();
        }
        for (; j >= 64; j >>= 3) {
            bitonic_pass<3>(s, n2, k, j, tid, T);
            __syncthreads();
        }
        bitonic_pass<4>(s, n2, k, 8, tid, T);
        __syncthreads();
    }
    const uint64_t hi_word = (uint64_t)tile << 32;
    for (uint32_t i = tid; i < n; i += T) {
        const uint64_t v = s[sort_pad(i)];
        point_list[range.x + i] = (uint32_t)v;
        keys[range.x + i] = hi_word | (v >> 32);   // the reference's sorted key: tile << 32 | depth bits
    }
}

int launch_tile_binning(int P, int R, int max_tile, int n_small, const Camera& cam, const int* radii, const GeomView& g,
                        const BinningView& b, const ImageView& img, cudaStream_t stream, bool debug) {
    if (P <= 0 || R <= 0) return 0;
    const int tiles = cam.grid_x * cam.grid_y;
    prof_begin(ST_DUPLICATE, stream);
    scatter_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.means2D, g.depths, img.tile_count, b.keys_unsorted, radii,
                                                       cam.grid_x, cam.grid_y, img.info);
    prof_end(ST_DUPLICATE, stream);
    HS_LAUNCH_OK(stream, debug);
    prof_begin(ST_SORT, stream);
    const uint32_t small_cap = HS_TILE_SORT_SMALL;
    if (n_small > 0) {
        uint32_t cap = 16;
        while (cap < (uint32_t)max_tile && cap < small_cap) cap <<= 1;
        tile_sort_kernel<HS_SORT_T_SMALL><<<tiles, HS_SORT_T_SMALL, (cap + cap / 16) * sizeof(uint64_t), stream>>>(
            img.ranges, b.keys_unsorted, b.point_list, b.keys, 0u, small_cap);
        count_launch();
    }
    if ((uint32_t)max_tile > small_cap) {
        uint32_t cap = small_cap;
        while (cap < (uint32_t)max_tile) cap <<= 1;
        auto k = tile_sort_kernel<512>;
        const size_t smem = (size_t)(cap + cap / 16) * sizeof(uint64_t);
        HS_CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<tiles, 512, smem, stream>>>(img.ranges, b.keys_unsorted, b.point_list, b.keys, small_cap,
                                         (uint32_t)HS_TILE_SORT_MAX);
        count_launch();
    }
    prof_end(ST_SORT, stream);
    HS_CUDA_OK(cudaGetLastError());
    if (debug) HS_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

int launch_binning(int P, int R, const Camera& cam, const int* radii, const GeomView& g, const BinningView& b,
                   const ImageView& img, cudaStream_t stream, bool debug) {
    const int tiles = cam.grid_x * cam.grid_y;
    HS_CUDA_OK(cudaMemsetAsync(img.ranges, 0, sizeof(uint2) * (size_t)tiles, stream));
    if (P <= 0 || R <= 0) return 0;
    prof_begin(ST_DUPLICATE, stream);
    duplicate_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.means2D, g.depths, g.point_offsets, b.keys_unsorted,
                                                         b.point_list_unsorted, radii, cam.grid_x, cam.grid_y);
    prof_end(ST_DUPLICATE, stream);
    HS_LAUNCH_OK(stream, debug);
    const int bit = (int)higher_msb((uint32_t)tiles);
    size_t tb = b.sort_temp_bytes;
    prof_begin(ST_SORT, stream);
    HS_CUDA_OK(cub::DeviceRadixSort::SortPairs(b.sort_temp, tb, b.keys_unsorted, b.keys, b.point_list_unsorted,
                                               b.point_list, R, 0, 32 + bit, stream));
    prof_end(ST_SORT, stream);
    count_lib_call();
    HS_CUDA_OK(cudaGetLastError());
    if (debug) HS_CUDA_OK(cudaStreamSynchronize(stream));
    prof_begin(ST_RANGES, stream);
    tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, b.keys, img.ranges);
    prof_end(ST_RANGES, stream);
    HS_LAUNCH_OK(stream, debug);
    return 0;
}

}  // namespace hs
