// Dev probe: semantics of cp.async.bulk.tensor.2d tile::gather4 on sm_100a (no PTX manual in the image).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gather4_probe gather4_probe.cu -lcuda && ./gather4_probe <box_rows>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap map, uint16_t* out, int r0, int r1, int r2, int r3, int c0, int expect) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    uint8_t* dst = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) dst[i] = 0xEE;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(expect) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                smem_u32(dst)),
            "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
            : "memory");
        uint32_t ok = 0, spins = 0;
        while (!ok && ++spins < (1u << 22)) {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
        }
        out[2048] = (uint16_t)ok;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(dst)[i];
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
    const int rows = 1024, cols = 256;
    std::vector<uint16_t> h(rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) h[r * cols + c] = (uint16_t)(r * 16 + (c / 8) % 16 + ((c % 8) << 12));  // row, 16B-chunk id, elt in chunk
    uint16_t *d, *o;
    cudaMalloc(&d, h.size() * 2);
    cudaMalloc(&o, 4200 * 2);
    cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    cuInit(0);
    CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode box_rows=%d -> %d\n", box_rows, (int)r);
    if (r != CUDA_SUCCESS) return 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    k<<<1, 128, 8192>>>(map, o, 5, 100, 7, 900, 64, 512);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 2;
    std::vector<uint16_t> res(4200);
    cudaMemcpy(res.data(), o, 4200 * 2, cudaMemcpyDeviceToHost);
    printf("barrier completed: %d\n", res[2048]);
    for (int line = 0; line < 8; ++line) {  // 128-byte lines of the destination: print (row, chunk) of each 16-byte chunk
        printf("line %d:", line);
        for (int ch = 0; ch < 8; ++ch) {
            uint16_t v = res[line * 64 + ch * 8];
            if (v == 0xEEEE) printf("  ----");
            else printf("  r%03d.c%02d", (v & 0xfff) / 16, v & 15);
        }
        printf("\n");
    }
    return 0;
}
