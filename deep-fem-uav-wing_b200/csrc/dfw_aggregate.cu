// (b) deterministic segmented neighbour aggregation over a CSR (no atomics).
//
// Replaces PyG's  x.index_select(0, src) -> zeros.scatter_add_(0, dst, .) -> / clamp(count,1)
// behind SAGEConv(aggr='mean') (reference call site src/deep_fem_uav_wing/gnn/model.py:90),
// and - with the transposed CSR and row_scale = NULL - the backward of that mean.
//
// Layout: one group of LANES lanes owns one destination row; every lane moves 16-byte vectors,
// so a group reads a whole source row with one coalesced request (512 B row = one warp-wide
// 128-bit load).  Column indices of a row are fetched LANES at a time (coalesced) and broadcast
// by shuffle; neighbour rows are fetched four at a time before accumulation (memory-level
// parallelism).  Accumulation is fp32 in CSR order -> bit-reproducible.
// Roofline: HBM.  Algorithmic bytes per launch  A_min = 2*N*H*b + 4*E + 4*(N+1)  (DESIGN.md).
#include <algorithm>

#include "dfw_common.cuh"

namespace dfw {
namespace {

template <typename T, int LANES, int VPL>
__global__ void __launch_bounds__(256) k_aggregate(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                    const float* __restrict__ row_scale, const T* __restrict__ x,
                                                    const T* __restrict__ addend, T* __restrict__ out, int64_t N,
                                                    int nvec /*16B vectors per row*/) {
    using V = Vec16<T>;
    constexpr int EPV = V::N;
    constexpr int GROUPS = 32 / LANES;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t row = warp_global * GROUPS + lane / LANES;
    const bool row_ok = row < N;

    int beg = 0, deg = 0;
    if (row_ok) {
        beg = __ldg(rowptr + row);
        deg = __ldg(rowptr + row + 1) - beg;
    }
    int maxdeg = deg;
    if (GROUPS > 1) {
#pragma unroll
        for (int o = 16; o >= LANES; o >>= 1) maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, o));
    }

    float acc[VPL][EPV];
#pragma unroll
    for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int i = 0; i < EPV; ++i) acc[v][i] = 0.f;

    bool vec_ok[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) vec_ok[v] = (sub + v * LANES) < nvec;

    const uint4* xv = reinterpret_cast<const uint4*>(x);

    for (int base = 0; base < maxdeg; base += LANES) {
        const int mine = (base + sub < deg) ? __ldg(col + beg + base + sub) : -1;
        const int m = min(LANES, maxdeg - base);
        for (int j0 = 0; j0 < m; j0 += 4) {
            int c[4];
            uint4 buf[4][VPL];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                // shuffle source (j0+u) % LANES stays inside the group; entries past m are -1 or ignored
                int cc = __shfl_sync(0xffffffffu, mine, (j0 + u) % LANES, LANES);
                c[u] = (j0 + u < m) ? cc : -1;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    if (c[u] >= 0 && vec_ok[v]) buf[u][v] = __ldg(xv + (int64_t)c[u] * nvec + sub + v * LANES);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (c[u] >= 0) {
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        if (vec_ok[v]) {
                            V t;
                            t.v = *reinterpret_cast<decltype(t.v)*>(&buf[u][v]);
                            float f[EPV];
                            t.to_float(f);
#pragma unroll
                            for (int i = 0; i < EPV; ++i) acc[v][i] += f[i];
                        }
                    }
                }
            }
        }
    }

    if (row_ok) {
        const float sc = row_scale ? __ldg(row_scale + row) : 1.0f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (vec_ok[v]) {
                const int64_t off = row * nvec + sub + v * LANES;
                float f[EPV];
                if (addend) {
                    V a;
                    uint4 raw = __ldg(reinterpret_cast<const uint4*>(addend) + off);
                    a.v = *reinterpret_cast<decltype(a.v)*>(&raw);
                    a.to_float(f);
#pragma unroll
                    for (int i = 0; i < EPV; ++i) f[i] = fmaf(acc[v][i], sc, f[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < EPV; ++i) f[i] = acc[v][i] * sc;
                }
                V o;
                o.from_float(f);
                reinterpret_cast<decltype(o.v)*>(out)[off] = o.v;
            }
        }
    }
}

template <typename T, int LANES, int VPL>
int launch(const int32_t* rowptr, const int32_t* col, const float* row_scale, const void* x, const void* addend,
           void* out, int64_t N, int nvec, cudaStream_t s) {
    constexpr int GROUPS = 32 / LANES;
    const int threads = 256;
    const int64_t rows_per_block = (int64_t)(threads / 32) * GROUPS;
    const int64_t blocks = (N + rows_per_block - 1) / rows_per_block;
    if (blocks == 0) return 0;
    k_aggregate<T, LANES, VPL><<<(unsigned)blocks, threads, 0, s>>>(rowptr, col, row_scale, (const T*)x, (const T*)addend,
                                                                    (T*)out, N, nvec);
    DFW_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int dispatch(const int32_t* rowptr, const int32_t* col, const float* row_scale, const void* x, const void* addend,
             void* out, int64_t N, int nvec, cudaStream_t s) {
#define DFW_AGG(L, V) return launch<T, L, V>(rowptr, col, row_scale, x, addend, out, N, nvec, s)
    if (nvec <= 4) DFW_AGG(4, 1);
    if (nvec <= 8) DFW_AGG(8, 1);
    if (nvec <= 16) DFW_AGG(16, 1);
    if (nvec <= 32) DFW_AGG(32, 1);
    if (nvec <= 64) DFW_AGG(32, 2);
    if (nvec <= 128) DFW_AGG(32, 4);
#undef DFW_AGG
    set_error("dfw_sage_aggregate: row of %d x 16 bytes is wider than the supported 2048 bytes", nvec);
    return 1;
}

}  // namespace
}  // namespace dfw

extern "C" int dfw_sage_aggregate(const int32_t* rowptr, const int32_t* col, const float* row_scale, const void* x,
                                  const void* addend, void* out, int64_t N, int64_t H, int dtype,
                                  dfw_stream_t stream) {
    using namespace dfw;
    DFW_REQUIRE(N >= 0 && H > 0, "dfw_sage_aggregate: bad shape N=%lld H=%lld", (long long)N, (long long)H);
    DFW_REQUIRE(dtype == DFW_F32 || dtype == DFW_BF16, "dfw_sage_aggregate: unknown dtype %d", dtype);
    const int64_t row_bytes = H * (dtype == DFW_F32 ? 4 : 2);
    DFW_REQUIRE(row_bytes % 16 == 0, "dfw_sage_aggregate: H*sizeof(dtype) = %lld must be a multiple of 16",
                (long long)row_bytes);
    if (N == 0) return 0;
    DFW_REQUIRE(rowptr && col && x && out, "dfw_sage_aggregate: null pointer");
    DFW_REQUIRE(aligned16(x) && aligned16(out) && (!addend || aligned16(addend)),
                "dfw_sage_aggregate: x/out/addend must be 16-byte aligned");
    DFW_REQUIRE(N * (row_bytes / 16) < (1LL << 40), "dfw_sage_aggregate: tensor too large");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = (int)(row_bytes / 16);
    if (dtype == DFW_F32) return dispatch<float>(rowptr, col, row_scale, x, addend, out, N, nvec, s);
    return dispatch<__nv_bfloat16>(rowptr, col, row_scale, x, addend, out, N, nvec, s);
}
