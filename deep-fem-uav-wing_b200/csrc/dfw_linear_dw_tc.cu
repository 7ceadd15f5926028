// (d3) weight gradients on the tensor cores:  dW[Hout, K] = g_y^T . a   (reduction over the nodes)
//
// Replaces autograd's `grad.T @ input` for lin_l / lin_r (reference: loss.backward(), train_gnn.py:57).
// Both operands are consumed in their natural row-major [node, feature] layout - no transposed copies:
// a TMA box [32 nodes x 128 B of features] (SWIZZLE_128B) IS the canonical MN-major UMMA tile
//   ((8,n),(8,k)) : ((1,LBO),(8,SBO))   in 16-byte units  (cute::UMMA::make_umma_desc<Major::MN>)
// with k = node (8 rows of 128 B = one 1024-byte swizzle atom, SBO = 1024 B) and MN = feature
// (one box = 128 B of features, LBO = box pitch).  So g_y^T is the A operand (M = 128 output features),
// the layer input is the B operand (N = 128 input features), both `major = MN` in the instruction descriptor.
// fp32 uses the same 3xTF32 split as the forward (both operands are activations, so both are split in
// shared memory by the converter warps); bf16 uses kind::f16.
// Each CTA owns one [128 x 128] tile of dW and one slice of the nodes (split-K), keeps the accumulator in
// TMEM for the whole slice and writes one partial tile; a fixed-order second pass sums the slices
// (deterministic, no float atomics).
#include <cuda.h>

#include <algorithm>
#include <cstring>

#include "dfw_common.cuh"
#include "dfw_linear_tc.cuh"
#include "dfw_tc_common.cuh"

namespace dfw {
namespace tc {

constexpr int kDwThreads = 192;
// nodes (K) per pipeline stage (measured: 16-node stages x 6 are slower, 126 vs 117 us - the per-stage barrier and
// descriptor overhead outweighs the shorter convert/MMA share of the round trip)
constexpr int kDwNodes = 32;
constexpr int kDwMaxStages = 4;
constexpr int kDwTile = 128;

struct DwMaps {
    CUtensorMap g;
    CUtensorMap a[2];
};

struct DwArgs {
    int64_t N;
    int64_t nodes_per_split;  // multiple of kDwNodes
    int tiles_j[2];           // 128-feature column tiles of a1 / a2
    int k[2];
    int Hout;
    int stages;
    float* part;              // [splits][tiles][128][128]
};

// MN-major descriptor: LBO = pitch between 128-byte feature blocks, SBO = pitch between k-groups.
//  16-bit operands: SWIZZLE_128B (layout type 2), k-group = 8 node rows (1024 B).
//  32-bit operands: the only MN-major layout UMMA accepts for tf32 is SWIZZLE_128B with a 32-byte swizzle
//  atom (layout type 1, cute Layout_MN_SW128_32B_Atom: Swizzle<2,5,2>, 4 rows x 128 B), k-group = 4 node rows
//  (512 B); TMA writes it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
template <bool TF32>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    constexpr uint64_t sbo = TF32 ? 512 : 1024;
    constexpr uint64_t layout = TF32 ? 1 : 2;
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | (layout << 61);
}

template <typename T, bool TF32>
__global__ void __launch_bounds__(kDwThreads) k_dw_tc(const __grid_constant__ DwMaps maps, const DwArgs p) {
    constexpr int EPB = kChunkBytes / (int)sizeof(T);       // features per 128-byte block: 32 fp32 / 64 bf16
    constexpr int NB = kDwTile / EPB;                       // feature blocks per 128-feature tile: 4 / 2
    constexpr uint32_t BOX = kDwNodes * kChunkBytes;        // 4 KB: [32 nodes x 128 B]
    constexpr uint32_t OPER = NB * BOX;                     // one operand tile per stage: 16 KB / 8 KB
    constexpr uint32_t STAGE = (TF32 ? 4 : 2) * OPER;       // g_hi, [g_lo], a_hi, [a_lo]
    constexpr uint32_t G_HI = 0, G_LO = OPER, A_HI = TF32 ? 2 * OPER : OPER, A_LO = 3 * OPER;
    constexpr int KSTEP_NODES = TF32 ? 8 : 16;              // UMMA_K
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET into the shared window: the pointer keeps its address space, so the compiler
    // emits LDS/STS instead of generic LD/ST for the converters and the epilogue staging
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * STAGE);
    uint64_t* empty = full + kDwMaxStages;
    uint64_t* conv = empty + kDwMaxStages;
    uint64_t* accum_full = conv + kDwMaxStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(accum_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_j = p.tiles_j[0] + p.tiles_j[1];
    const int ti = blockIdx.x / tiles_j;
    int tj = blockIdx.x % tiles_j;
    const int which = tj >= p.tiles_j[0];
    if (which) tj -= p.tiles_j[0];
    const int64_t node_beg = (int64_t)blockIdx.y * p.nodes_per_split;
    const int64_t node_end = min(p.N, node_beg + p.nodes_per_split);
    const int nsteps = node_end > node_beg ? (int)((node_end - node_beg + kDwNodes - 1) / kDwNodes) : 0;
    // ragged tiles: only the 128-byte feature blocks that exist are fetched.  Missing M blocks (Hout < 128) are
    // zero-filled once (the MMA always covers M = 128); missing N blocks simply shrink the instruction's N.
    const int mb_valid = min(NB, (p.Hout - ti * kDwTile + EPB - 1) / EPB);
    const int nb_valid = min(NB, (p.k[which] - tj * kDwTile + EPB - 1) / EPB);
    const int n_cols = nb_valid * EPB;
    const bool stacked = TF32 && nb_valid == NB;  // see the MMA issuer
    if (mb_valid < NB) {
        for (int s = 0; s < p.stages; ++s) {
            uint4* z = reinterpret_cast<uint4*>(smem + (size_t)s * STAGE + G_HI + mb_valid * BOX);
            const int n16 = (NB - mb_valid) * (int)BOX / 16;
            for (int i = threadIdx.x; i < n16; i += kDwThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
            if (TF32) {
                uint4* zl = reinterpret_cast<uint4*>(smem + (size_t)s * STAGE + G_LO + mb_valid * BOX);
                for (int i = threadIdx.x; i < n16; i += kDwThreads) zl[i] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        fence_proxy_async();
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.g);
        prefetch_tmap(&maps.a[which]);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&conv[s], 128);
        }
        mbar_init(accum_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(TF32 ? 512u : 128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_tc_before();
    __syncthreads();
    fence_tc_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nsteps; ++it) {
                const int node0 = (int)(node_beg + (int64_t)it * kDwNodes);
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* st = smem + (size_t)stage * STAGE;
                mbar_arrive_expect_tx(&full[stage], (uint32_t)(mb_valid + nb_valid) * BOX);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    if (b < mb_valid) tma_load_2d(st + G_HI + b * BOX, &maps.g, &full[stage], ti * kDwTile + b * EPB, node0);
                    if (b < nb_valid) tma_load_2d(st + A_HI + b * BOX, &maps.a[which], &full[stage], tj * kDwTile + b * EPB, node0);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // F32 accumulate | a/b format | A and B MN-major (bits 15, 16) | N = 128 | M = 128
            const uint32_t fmt = TF32 ? 2u : 1u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n_cols >> 3) << 17) |
                                   ((uint32_t)(kDwTile >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, accumulate = 0;
            uint32_t acc_main[2] = {0u, 0u}, acc_cross = 0u;  // fp32: two hi*hi accumulators + one for the cross terms
            // full-width tiles: a_lo sits right behind a_hi (same block pitch), so ONE N = 256 instruction yields
            // g_hi.a_hi (main) and g_hi.a_lo (cross) - two instructions per k-step instead of three (the MMA issue
            // rate, ~140 clk per instruction whatever N, and the operand reads bound this kernel).  TMEM columns:
            // [main0 | cross0 | main1 | cross1], k-steps alternate between the two pairs.
            const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * kDwTile) >> 3) << 17);
            int kstep = 0;
            for (int it = 0; it < nsteps; ++it) {
                mbar_wait(TF32 ? &conv[stage] : &full[stage], phase);
                fence_tc_after();
                const uint32_t st = smem_u32(smem + (size_t)stage * STAGE);
#pragma unroll
                for (int ks = 0; ks < kDwNodes / KSTEP_NODES; ++ks) {
                    const uint32_t koff = ks * KSTEP_NODES * kChunkBytes;  // 8 (16) node rows of 128 B
                    const uint64_t g_hi = make_desc_mn<TF32>(st + G_HI + koff, BOX);
                    const uint64_t a_hi = make_desc_mn<TF32>(st + A_HI + koff, BOX);
                    if (TF32) {
                        const uint64_t g_lo = make_desc_mn<TF32>(st + G_LO + koff, BOX);
                        const uint64_t a_lo = make_desc_mn<TF32>(st + A_LO + koff, BOX);
                        const int m = kstep & 1;
                        if (stacked) {
                            umma<TF32>(tmem_base + m * 2 * kDwTile, g_hi, a_hi, idesc2, acc_main[m]);
                            umma<TF32>(tmem_base + m * 2 * kDwTile + kDwTile, g_lo, a_hi, idesc, 1u);
                        } else {
                            umma<TF32>(tmem_base + 2 * kDwTile, g_lo, a_hi, idesc, acc_cross);
                            umma<TF32>(tmem_base + 2 * kDwTile, g_hi, a_lo, idesc, 1u);
                            umma<TF32>(tmem_base + m * kDwTile, g_hi, a_hi, idesc, acc_main[m]);
                            acc_cross = 1u;
                        }
                        acc_main[m] = 1u;
                        ++kstep;
                    } else {
                        umma<TF32>(tmem_base, g_hi, a_hi, idesc, accumulate);
                    }
                    accumulate = 1u;
                }
                umma_commit(&empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(accum_full);
        }
    } else {
        const int et = threadIdx.x - 64;
        if (TF32) {
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < nsteps; ++it) {
                mbar_wait(&full[stage], phase);
                uint8_t* st = smem + (size_t)stage * STAGE;
#pragma unroll
                for (int op = 0; op < 2; ++op) {  // g then a: position-preserving split, swizzle-agnostic
                    float4* hi = reinterpret_cast<float4*>(st + (op ? A_HI : G_HI));
                    float4* lo = reinterpret_cast<float4*>(st + (op ? A_LO : G_LO));
#pragma unroll
                    for (int i = 0; i < (int)(OPER / 16) / 128; ++i) {
                        const int idx = et + i * 128;
                        lo[idx] = tf32_lo(hi[idx]);
                    }
                }
                fence_proxy_async();
                mbar_arrive(&conv[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
        // epilogue: thread (= output feature row of the tile) writes its 128 partial sums
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int tiles = gridDim.x;
        float* dst = p.part + (((int64_t)blockIdx.y * tiles + blockIdx.x) * kDwTile + r) * kDwTile;
        if (nsteps > 0) {
            mbar_wait(accum_full, 0);
            fence_tc_after();
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
            const bool two_main = nsteps * (kDwNodes / KSTEP_NODES) >= 2;
#pragma unroll
            for (int c0 = 0; c0 < kDwTile; c0 += 32) {
                if (c0 >= n_cols) break;  // columns beyond the instruction's N were never written
                float v[32];
                tmem_ld32(t_row + c0, v);
                if (TF32) {  // round-to-nearest sum of the accumulators (see tmem_combine)
                    float w[32];
                    if (stacked) {  // [main0 | cross0 | main1 | cross1]; a stage is 4 k-steps, so both pairs exist
#pragma unroll
                        for (int a = 1; a < 4; ++a) {
                            tmem_ld32(t_row + a * kDwTile + c0, w);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += w[j];
                        }
                    } else {
                        if (two_main) {
                            tmem_ld32(t_row + kDwTile + c0, w);
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] += w[j];
                        }
                        tmem_ld32(t_row + 2 * kDwTile + c0, w);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += w[j];
                    }
                }
#pragma unroll
                for (int g = 0; g < 8; ++g)
                    *reinterpret_cast<float4*>(dst + c0 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
            }
        } else {
#pragma unroll
            for (int g = 0; g < kDwTile / 4; ++g) *reinterpret_cast<float4*>(dst + 4 * g) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    fence_tc_before();
    __syncthreads();
    if (warp == 2) {
        fence_tc_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TF32 ? 512u : 128u) : "memory");
    }
}

// column sums of g_y (bias gradient): two-pass, fixed order.  Pass 1: block b sums rows b, b+G, b+2G, ...
// (each warp reads whole 128..1024-byte rows, 4 rows in flight per thread); pass 2 sums the G partials in order.
constexpr int kColsumBlocks = kNumSMs * 4;
template <typename T>
__global__ void __launch_bounds__(256) k_colsum_partial(const T* __restrict__ g, int64_t N, int H, float* __restrict__ part) {
    __shared__ float red[8][257];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // warp w of block b handles rows (b*8 + w) + k * (gridDim.x*8); lane owns columns lane + 32*j
    for (int64_t r = (int64_t)blockIdx.x * 8 + wid; r < N; r += (int64_t)gridDim.x * 8) {
        const T* row = g + r * H;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (lane + 32 * j < H) acc[j] += to_f32(row[lane + 32 * j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (lane + 32 * j < H) red[wid][lane + 32 * j] = acc[j];
    __syncthreads();
    for (int c = threadIdx.x; c < H; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][c];
        part[(int64_t)blockIdx.x * H + c] = s;
    }
}
__global__ void __launch_bounds__(256) k_colsum_final(const float* __restrict__ part, int blocks, int H, float* __restrict__ out,
                                                       int accumulate) {
    // one warp per column, lanes stride over the partials, fixed shuffle tree (deterministic)
    const int lane = threadIdx.x & 31;
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= H) return;
    float s = 0.f;
    for (int b = lane; b < blocks; b += 32) s += part[(int64_t)b * H + c];
    s = warp_sum(s);
    if (lane == 0) out[c] = accumulate ? out[c] + s : s;
}

}  // namespace tc

bool dw_tc_eligible(int64_t N, int64_t Hout, int64_t k1, int64_t k2, int dtype, const void* g, const void* a1, const void* a2) {
    const int epb = dtype == DFW_F32 ? 32 : 64;  // features per 128-byte block
    if (N < 1 || N >= (1LL << 31) || Hout % epb || Hout > 256 || Hout < 32) return false;
    if (k1 % epb || k1 < 32 || (a2 && (k2 % epb || k2 < 32))) return false;
    return aligned16(g) && aligned16(a1) && (!a2 || aligned16(a2));
}

// Partial layout is identical to the SIMT split-K path ([split][tile][128][128]) so the same second pass is used.
int dw_tc_launch(const void* g_y, const void* a1, int64_t k1, const void* a2, int64_t k2, int64_t N, int64_t Hout, int dtype,
                 float* part, int splits, int64_t nodes_per_split, cudaStream_t s) {
    using namespace tc;
    const bool tf32 = dtype == DFW_F32;
    const int e = tf32 ? 4 : 2;
    DwMaps maps;
    memset(&maps, 0, sizeof(maps));
    if (make_map(&maps.g, g_y, N, Hout, e, kDwNodes, tf32 ? kMapSw128Atom32 : kMapSw128)) return 1;
    if (make_map(&maps.a[0], a1, N, k1, e, kDwNodes, tf32 ? kMapSw128Atom32 : kMapSw128)) return 1;
    if (a2 && make_map(&maps.a[1], a2, N, k2, e, kDwNodes, tf32 ? kMapSw128Atom32 : kMapSw128)) return 1;
    DwArgs p{};
    p.N = N;
    p.nodes_per_split = nodes_per_split;
    p.tiles_j[0] = (int)((k1 + kDwTile - 1) / kDwTile);
    p.tiles_j[1] = a2 ? (int)((k2 + kDwTile - 1) / kDwTile) : 0;
    p.k[0] = (int)k1;
    p.k[1] = (int)k2;
    p.Hout = (int)Hout;
    p.part = part;
    const uint32_t stage = (tf32 ? 4u : 2u) * (uint32_t)(kDwTile * e / kChunkBytes) * kDwNodes * kChunkBytes;
    p.stages = tf32 ? 3 : 4;  // 3 x 64 KB (fp32) / 4 x 16 KB (bf16)
    const size_t smem = (size_t)p.stages * stage + 8 * (3 * kDwMaxStages + 1) + 16 + 1024;
    dim3 grid((unsigned)(((Hout + kDwTile - 1) / kDwTile) * (p.tiles_j[0] + p.tiles_j[1])), (unsigned)splits, 1);
    if (tf32) {
        auto kern = k_dw_tc<float, true>;
        DFW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kDwThreads, smem, s>>>(maps, p);
    } else {
        auto kern = k_dw_tc<__nv_bfloat16, false>;
        DFW_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kDwThreads, smem, s>>>(maps, p);
    }
    DFW_LAUNCH_CHECK();
    return 0;
}

size_t colsum_ws_bytes(int64_t H) { return align_up(sizeof(float) * (size_t)tc::kColsumBlocks * (size_t)H, 256); }

int colsum_launch(const void* g_y, int64_t N, int64_t H, int dtype, float* part, float* dbias, int accumulate, cudaStream_t s) {
    using namespace tc;
    if (H > 256) {
        set_error("colsum: H=%lld > 256", (long long)H);
        return 1;
    }
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((N + 7) / 8, kColsumBlocks));
    if (dtype == DFW_F32) k_colsum_partial<float><<<blocks, 256, 0, s>>>((const float*)g_y, N, (int)H, part);
    else k_colsum_partial<__nv_bfloat16><<<blocks, 256, 0, s>>>((const __nv_bfloat16*)g_y, N, (int)H, part);
    DFW_LAUNCH_CHECK();
    k_colsum_final<<<(unsigned)((H * 32 + 255) / 256), 256, 0, s>>>(part, blocks, (int)H, dbias, accumulate);
    DFW_LAUNCH_CHECK();
    return 0;
}

}  // namespace dfw
