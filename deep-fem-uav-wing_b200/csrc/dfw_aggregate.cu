// (b) deterministic segmented neighbour aggregation over a CSR (no atomics).
//
// Replaces PyG's  x.index_select(0, src) -> zeros.scatter_add_(0, dst, .) -> / clamp(count,1)
// behind SAGEConv(aggr='mean') (reference call site src/deep_fem_uav_wing/gnn/model.py:90),
// and - with the transposed CSR and row_scale = NULL - the backward of that mean.
//
// Design ("row-block edge streaming"):
//  * a group of LANES lanes (a full warp for 512-byte rows) owns a block of R consecutive destination
//    rows and STREAMS the block's contiguous edge range: column indices are fetched LANES at a time
//    (coalesced, L1-bypassing) and broadcast by shuffle; the source rows of U = 8 edges are requested
//    back to back (8 x 512 B in flight per warp whatever the degree) before any of them is consumed;
//  * row boundaries are group-uniform branches inside the stream, so short rows (degree ~6 surface
//    meshes) cost no extra dependent round trip: rowptr and col are each read once per block;
//  * every lane moves 16-byte vectors; fp32 accumulation in CSR order (bit-reproducible) using packed
//    FADD2 for fp32 rows and FHADD.BF16 (f32 += bf16, no unpack) for bf16 rows - the bf16 config is
//    instruction-issue bound, not bandwidth bound, on a naive kernel (profiles/r01_ncu_*_v1.csv).
// Roofline: HBM.  Algorithmic bytes per launch  A_min = 2*N*H*b + 4*E + 4*(N+1)  (DESIGN.md).
#include <algorithm>

#include "dfw_common.cuh"

namespace dfw {
namespace {


__device__ __forceinline__ uint4 ldg_nc_v4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream_s32(const int* p) {
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// acc (fp32) += one 16-byte vector of T
template <typename T>
struct Acc;
template <>
struct Acc<float> {
    static constexpr int N = 4;
    uint64_t a[2];  // two packed f32x2
    __device__ __forceinline__ void zero() { a[0] = a[1] = 0ull; }
    __device__ __forceinline__ void add(const uint4& v) {
        uint64_t lo = ((uint64_t)v.y << 32) | v.x, hi = ((uint64_t)v.w << 32) | v.z;
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a[0]) : "l"(lo));
        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(a[1]) : "l"(hi));
    }
    __device__ __forceinline__ void get(float* f) const {
        f[0] = __uint_as_float((uint32_t)a[0]); f[1] = __uint_as_float((uint32_t)(a[0] >> 32));
        f[2] = __uint_as_float((uint32_t)a[1]); f[3] = __uint_as_float((uint32_t)(a[1] >> 32));
    }
};
template <>
struct Acc<__nv_bfloat16> {
    static constexpr int N = 8;
    float a[8];
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.f;
    }
    __device__ __forceinline__ void add(const uint4& v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi;
            asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w[i]));
            asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(a[2 * i]) : "h"(lo));      // FHADD.BF16: f32 += bf16
            asm("add.rn.f32.bf16 %0, %1, %0;" : "+f"(a[2 * i + 1]) : "h"(hi));
        }
    }
    __device__ __forceinline__ void get(float* f) const {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = a[i];
    }
};

constexpr int kAggThreads = 1024;

// Persistent: one 32-warp CTA per SM.  Rows are cut into CHUNKS of 32 row-blocks; chunk k belongs to CTA
// k mod gridDim.x and a CTA's warps take the blocks of its chunks in order from a CTA-local counter.  So
//  * inside an SM, ~32*R neighbouring rows are in flight together: the overlapping neighbourhoods of nearby
//    rows are served by L1 (mesh numberings are banded: a source row is requested by several nearby
//    destination rows) - L1 hit rate 15 % -> 57 % on the 2M-node lattice;
//  * across the chip, all SMs advance through the node array as ONE wavefront (148 chunks wide), so the
//    longer-range reuse (neighbouring lattice planes) stays inside the 126 MB L2 and DRAM sees each row once.
// (A contiguous band per SM gives the first property but loses the second: DRAM reads tripled.)
template <typename T, int LANES, int VPL>
__global__ void __launch_bounds__(kAggThreads, 1) k_aggregate(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                              const float* __restrict__ row_scale, const T* __restrict__ x,
                                                              const T* __restrict__ addend, T* __restrict__ out, int64_t N,
                                                              int nvec /*16B vectors per row*/, int R /*rows per group, < LANES*/) {
    using V = Vec16<T>;
    constexpr int EPV = V::N;
    constexpr int GROUPS = 32 / LANES;
    constexpr int kU = VPL >= 4 ? 2 : (VPL == 2 ? 4 : 8);  // source rows in flight per group (8 x 16 B per lane)
    __shared__ int s_next;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LANES;
    const unsigned gmask = LANES == 32 ? 0xffffffffu : (((1u << LANES) - 1u) << (lane / LANES * LANES));
    constexpr int kBPC = kAggThreads / 32;  // row-blocks (one per warp) per chunk
    const int64_t unit = (int64_t)R * GROUPS;  // rows per warp step
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    // Lanes past the end of a row (only when nvec < LANES*VPL) load a clamped, valid vector and never
    // store: every load below is UNCONDITIONAL, so the compiler keeps kU independent requests in flight.
    bool vec_ok[VPL];
    int voff[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) {
        vec_ok[v] = (sub + v * LANES) < nvec;
        voff[v] = min(sub + v * LANES, nvec - 1);
    }
    const uint4* xv = reinterpret_cast<const uint4*>(x);
    const uint4* av = addend ? reinterpret_cast<const uint4*>(addend) : nullptr;
    uint4* ov = reinterpret_cast<uint4*>(out);

  for (;;) {
    int blk = 0;
    if (lane == 0) blk = atomicAdd(&s_next, 1);
    blk = __shfl_sync(0xffffffffu, blk, 0);
    const int64_t chunk = (int64_t)blockIdx.x + (int64_t)(blk / kBPC) * gridDim.x;
    const int64_t w0 = (chunk * kBPC + blk % kBPC) * unit;
    if (chunk * kBPC * unit >= N) break;  // this CTA's chunks are exhausted (warp-uniform)
    const int64_t r0 = w0 + (int64_t)(lane / LANES) * R;
    if (r0 >= N) continue;  // ragged end of the last chunk (group-uniform)
    const int nrows = (int)min((int64_t)R, N - r0);

    // rowptr[r0 .. r0+nrows] -> one value per lane
    const int rp = sub <= nrows ? __ldg(rowptr + r0 + sub) : 0;
    const float rsc = (row_scale && sub < nrows) ? __ldg(row_scale + r0 + sub) : 1.0f;  // one scale per lane, no load at flush time
    const int e_beg = __shfl_sync(gmask, rp, 0, LANES);
    const int e_end = __shfl_sync(gmask, rp, nrows, LANES);

    Acc<T> acc[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) acc[v].zero();
    uint4 ad[VPL];
    int cur = 0;                                          // current row inside the block
    int cur_end = __shfl_sync(gmask, rp, 1, LANES);       // its end edge
    auto fetch_addend = [&](int r) {
        if (av) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) ad[v] = ldg_nc_v4(av + (r0 + r) * nvec + voff[v]);
        }
    };
    auto flush = [&]() {  // write row `cur`, start the next one
        const int64_t row = r0 + cur;
        const float sc = __shfl_sync(gmask, rsc, cur, LANES);
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            if (vec_ok[v]) {
                float f[EPV];
                acc[v].get(f);
                if (av) {
                    V a;
                    a.v = *reinterpret_cast<decltype(a.v)*>(&ad[v]);
                    float g[EPV];
                    a.to_float(g);
#pragma unroll
                    for (int i = 0; i < EPV; ++i) f[i] = fmaf(f[i], sc, g[i]);
                } else {
#pragma unroll
                    for (int i = 0; i < EPV; ++i) f[i] *= sc;
                }
                V o;
                o.from_float(f);
                ov[row * nvec + voff[v]] = *reinterpret_cast<uint4*>(&o.v);
            }
            acc[v].zero();
        }
        ++cur;
        if (cur < nrows) {
            cur_end = __shfl_sync(gmask, rp, cur + 1, LANES);
            fetch_addend(cur);
        }
    };
    fetch_addend(0);

    for (int eb = e_beg; eb < e_end; eb += LANES) {
        const int n = min(LANES, e_end - eb);
        // lanes past the end of the edge range repeat the last valid column (a harmless L1 hit)
        const int mine = ldg_stream_s32(col + eb + min(sub, n - 1));
        for (int j0 = 0; j0 < n; j0 += kU) {
            uint4 buf[kU][VPL];
            const int m = min(kU, n - j0);
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int c = __shfl_sync(gmask, mine, min(j0 + u, n - 1), LANES);
#pragma unroll
                for (int v = 0; v < VPL; ++v) buf[u][v] = ldg_nc_v4(xv + (int64_t)c * nvec + voff[v]);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (u < m) {
                    const int e = eb + j0 + u;
                    while (e == cur_end) flush();  // group-uniform; also steps over empty rows
#pragma unroll
                    for (int v = 0; v < VPL; ++v) acc[v].add(buf[u][v]);
                }
            }
        }
    }
    while (cur < nrows) flush();  // last row and trailing empty rows
  }
}

template <typename T, int LANES, int VPL>
int launch(const int32_t* rowptr, const int32_t* col, const float* row_scale, const void* x, const void* addend,
           void* out, int64_t N, int64_t E, int nvec, cudaStream_t s) {
    constexpr int GROUPS = 32 / LANES;
    // rows per group: aim at ~64 edges per group so that rowptr/col latency is amortised, keep R < LANES
    const double deg = N > 0 ? (double)E / (double)N : 0.0;
    int R = (int)(96.0 / (deg + 1.0));
    R = std::max(1, std::min(R, std::min(16, LANES - 1)));
    const int64_t chunk_rows = (int64_t)R * GROUPS * (kAggThreads / 32);
    const int64_t chunks = (N + chunk_rows - 1) / chunk_rows;
    const int64_t blocks = std::min<int64_t>(chunks, kNumSMs);
    if (blocks == 0) return 0;
    k_aggregate<T, LANES, VPL><<<(unsigned)blocks, kAggThreads, 0, s>>>(rowptr, col, row_scale, (const T*)x, (const T*)addend,
                                                                        (T*)out, N, nvec, R);
    DFW_LAUNCH_CHECK();
    return 0;
}

template <typename T>
int dispatch(const int32_t* rowptr, const int32_t* col, const float* row_scale, const void* x, const void* addend,
             void* out, int64_t N, int64_t E, int nvec, cudaStream_t s) {
#define DFW_AGG(L, V) return launch<T, L, V>(rowptr, col, row_scale, x, addend, out, N, E, nvec, s)
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
                                  const void* addend, void* out, int64_t N, int64_t E, int64_t H, int dtype,
                                  dfw_stream_t stream) {
    using namespace dfw;
    DFW_REQUIRE(N >= 0 && E >= 0 && H > 0, "dfw_sage_aggregate: bad shape N=%lld E=%lld H=%lld", (long long)N, (long long)E,
                (long long)H);
    DFW_REQUIRE(dtype == DFW_F32 || dtype == DFW_BF16, "dfw_sage_aggregate: unknown dtype %d", dtype);
    const int64_t row_bytes = H * (dtype == DFW_F32 ? 4 : 2);
    DFW_REQUIRE(row_bytes % 16 == 0, "dfw_sage_aggregate: H*sizeof(dtype) = %lld must be a multiple of 16",
                (long long)row_bytes);
    if (N == 0) return 0;
    DFW_REQUIRE(rowptr && (col || E == 0) && x && out, "dfw_sage_aggregate: null pointer");
    DFW_REQUIRE(aligned16(x) && aligned16(out) && (!addend || aligned16(addend)),
                "dfw_sage_aggregate: x/out/addend must be 16-byte aligned");
    DFW_REQUIRE(N * (row_bytes / 16) < (1LL << 40), "dfw_sage_aggregate: tensor too large");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int nvec = (int)(row_bytes / 16);
    if (dtype == DFW_F32) return dispatch<float>(rowptr, col, row_scale, x, addend, out, N, E, nvec, s);
    return dispatch<__nv_bfloat16>(rowptr, col, row_scale, x, addend, out, N, E, nvec, s);
}
