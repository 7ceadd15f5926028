// Dev probe: how fast does one SM's TMA unit deliver 2-D tile loads as a function of the box shape?
// Emulates the operand traffic of k_linear_tc: per K chunk one activation box [box_rows x CB bytes] from a large [N, K] fp32
// tensor plus W boxes [box_rows x CB bytes] from a small (L2-resident) weight tensor, ring of stages, no compute.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate_probe tma_rate_probe.cu -lcuda && ./tma_rate_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}

struct P { int tiles, kchunks, cb, wboxes, stages, box_rows; };

__global__ void __launch_bounds__(64) k(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mw, P p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t full[8];
    const uint32_t box = (uint32_t)p.box_rows * p.cb, stage_bytes = box * (1 + p.wboxes);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const int epc = p.cb / 4;
    int issued = 0, done = 0;
    const int my_tiles = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total = my_tiles * p.kchunks;
    while (done < total) {
        while (issued < total && issued - done < p.stages) {
            const int s = issued % p.stages;
            const int t = (int)blockIdx.x + (issued / p.kchunks) * (int)gridDim.x, c = issued % p.kchunks;
            uint8_t* st = smem + (size_t)s * stage_bytes;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(stage_bytes) : "memory");
            tma_load_2d(st, &ma, &full[s], c * epc, t * p.box_rows);
            for (int w = 0; w < p.wboxes; ++w) tma_load_2d(st + (size_t)(1 + w) * box, &mw, &full[s], c * epc, 0);
            ++issued;
        }
        mbar_wait(&full[done % p.stages], (uint32_t)((done / p.stages) & 1));
        ++done;
    }
}

static int make_map(CUtensorMap* m, void* base, long rows, long cols, int cb, int box_rows) {
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {(cuuint32_t)(cb / 4), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return (int)cuTensorMapEncodeTiled(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       cb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B),
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main() {
    const long N = 200000, K = 256;
    float *a, *w;
    cudaMalloc(&a, N * K * 4);
    cudaMalloc(&w, 256 * K * 4);
    cudaMemset(a, 0, N * K * 4);
    cudaMemset(w, 0, 256 * K * 4);
    cuInit(0);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("cb box_rows wboxes stages ctas/SM :  us   GB/s(all)  GB/s(activations)  clk/row@1.9GHz per SM\n");
    for (int ctas : {1, 2})
        for (int cb : {64, 128})
            for (int box_rows : {128, 64})
                for (int wboxes : {0, 2}) {
                    const int stages = 4;
                    CUtensorMap ma, mw;
                    if (make_map(&ma, a, N, K, cb, box_rows) || make_map(&mw, w, 256, K, cb, box_rows)) { printf("map failed\n"); return 1; }
                    P p{(int)((N + box_rows - 1) / box_rows), (int)(K * 4 / cb), cb, wboxes, stages, box_rows};
                    const size_t smem = (size_t)stages * box_rows * cb * (1 + wboxes) + 1024;
                    if (smem * ctas > 220 * 1024) continue;
                    float best = 1e9;
                    for (int it = 0; it < 5; ++it) {
                        cudaEventRecord(e0);
                        k<<<148 * ctas, 64, smem>>>(ma, mw, p);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        if (ms < best) best = ms;
                    }
                    if (cudaGetLastError() != cudaSuccess) { printf("kernel error\n"); return 2; }
                    const double bytes_a = (double)N * K * 4, bytes_all = bytes_a * (1 + wboxes);
                    const double rows = (double)p.tiles * p.kchunks * box_rows * (1 + wboxes);
                    printf("%3d %4d %d %d %d : %7.1f  %8.1f  %8.1f  %6.2f\n", cb, box_rows, wboxes, stages, ctas, best * 1e3, bytes_all / best / 1e6, bytes_a / best / 1e6,
                           best * 1e-3 * 1.9e9 / (rows / 148.0));
                }
    return 0;
}
