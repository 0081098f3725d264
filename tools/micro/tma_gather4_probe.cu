// TMA tile::gather4 probe (sm_100a): which box shape does the instruction want, and how do the four gathered rows land in
// shared memory?  Table [64 rows][40 floats], value = 100 row + col; rows {5, 17, 3, 60} are gathered.
//   nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -o tools/micro/tma_gather4_probe tools/micro/tma_gather4_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap map, float* out, int* err, int expect_bytes) {
    __shared__ __align__(128) float buf[4 * 40 * 4];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(buf);
    for (int i = threadIdx.x; i < 640; i += blockDim.x) buf[i] = -1.f;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(b), "r"(expect_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];\n"
                     ::"r"(d), "l"(&map), "r"(0), "r"(5), "r"(17), "r"(3), "r"(60), "r"(b) : "memory");
    }
    bool ok = false;
    for (int i = 0; i < (1 << 22) && !ok; i++) {
        uint32_t p;
        asm volatile("{\n\t.reg .pred q;\n\tmbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0;\n\tselp.u32 %0, 1, 0, q;\n\t}\n" : "=r"(p) : "r"(b) : "memory");
        ok = p != 0;
    }
    if (!ok && threadIdx.x == 0) err[0] = 1;
    __syncthreads();
    for (int i = threadIdx.x; i < 640; i += blockDim.x) out[i] = buf[i];
}

int main() {
    EncodeTiled enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q) != cudaSuccess || enc == nullptr) {
        printf("no cuTensorMapEncodeTiled\n");
        return 2;
    }
    const int R = 64, C = 40;
    float* h = (float*)malloc(sizeof(float) * R * C);
    for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) h[r * C + c] = 100.f * r + c;
    float *d, *dout; int* derr;
    cudaMalloc(&d, sizeof(float) * R * C); cudaMalloc(&dout, sizeof(float) * 640); cudaMalloc(&derr, 4);
    cudaMemcpy(d, h, sizeof(float) * R * C, cudaMemcpyHostToDevice);
    const cuuint32_t boxes[2][2] = {{(cuuint32_t)C, 1}, {(cuuint32_t)C, 4}};
    for (int v = 0; v < 2; v++) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R}, strides[1] = {(cuuint64_t)C * 4};
        cuuint32_t es[2] = {1, 1};
        CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, boxes[v], es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box {%u,%u}: encode rc=%d\n", boxes[v][0], boxes[v][1], (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        cudaMemset(derr, 0, 4);
        probe<<<1, 128>>>(map, dout, derr, 4 * C * 4);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); return 3; }
        float o[640]; int herr;
        cudaMemcpy(o, dout, sizeof(o), cudaMemcpyDeviceToHost); cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
        printf("  timeout=%d  smem[0,1,39,40,41,80,120,159,160]: %g %g %g %g %g %g %g %g %g\n", herr, o[0], o[1], o[39], o[40], o[41], o[80],
               o[120], o[159], o[160]);
    }
    return 0;
}
