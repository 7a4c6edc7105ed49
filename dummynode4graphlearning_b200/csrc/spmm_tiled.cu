// libdn4gl.so -- K1, tiled variant: sum aggregation as a warp-specialised producer/consumer pipeline.
//
// A mini-batch is block-diagonal: every neighbour of a row of graph g is a row of graph g.  Rows are cut into TILES
// (dn4gl_make_row_tiles): tile k starts at the first graph boundary inside [kC, (k+1)C), or at kC itself when a graph is
// longer than the window (a "cut" tile).  A persistent CTA walks tiles t = blockIdx.x, += gridDim.x through a ring of
// STAGES shared-memory buffers:
//   * warp 0 (one elected lane) is the PRODUCER: for every tile it issues three bulk asynchronous copies on the TMA
//     engine's 1-D path (cp.async.bulk ... mbarrier::complete_tx; SASS UBLKCP) -- the tile's feature rows x[r0:r1)
//     (one contiguous slab), its row_ptr slice and its col slice -- into the next free stage, as soon as the consumers
//     have released it (empty mbarrier);
//   * warps 1..15 are CONSUMERS: they wait on the stage's full mbarrier, resolve every neighbour index against shared
//     memory, accumulate in CSR order and stream the output rows to HBM with 128-bit stores.
// DRAM therefore sees each feature row once as a streaming read and each output row once as a streaming write (the
// algorithmic bytes of SURVEY.md 8(d)); there is no dependent row_ptr -> col -> x gather chain on DRAM (the latency
// chain that bounded the per-row kernel, profiles/r1a), and the loads of tile k+1 overlap the arithmetic of tile k.
//
// Correctness never depends on the tiling: an index outside the staged window, a col position beyond the staged
// slice, or a tile larger than a stage falls back to global loads inside the same kernel.
//
// Rows with more than SPLIT neighbours:
//   * in a graph-aligned (self-contained) tile all neighbours are in shared memory: the sub-groups of the row's warp
//     split the list and combine with a fixed butterfly (deterministic);
//   * in a cut tile (graph longer than the window) the neighbours are in HBM: such rows are listed once per tiling
//     (dn4gl_make_row_tiles) and run FIRST as "virtual tiles": all 15 consumer warps of a CTA stride one row's list
//     and combine through the idle stage buffer in a fixed order.
// Per-row accumulation order for rows with <= SPLIT neighbours is CSR order with separately rounded adds, i.e.
// bit-identical to the per-row kernel and to the sequential oracle.
#include "common.cuh"

constexpr int TP_THREADS = 512;
constexpr int TP_NCW = TP_THREADS / 32 - 1;   // consumer warps
constexpr int TP_SPLIT = 64;                  // rows above this many neighbours are split (== graph.py HEAVY_THRESHOLD)
constexpr int TP_MAX_STAGES = 4;
constexpr uint32_t TP_BULK_CHUNK = 32768u;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(TP_NCW * 32) : "memory"); }

// shared-memory carve-up of ONE stage: [cap_rows][DV] float4 feature rows | col slice | row_ptr slice
struct TileCfg {
    int stages, cap_rows, cap_nnz;
    uint32_t stage_bytes, off_col, off_rp;
};
__host__ __device__ inline TileCfg tile_cfg(int smem_bytes, int stages, int DV, int nnz_per_row) {
    TileCfg L;
    L.stages = stages;
    L.stage_bytes = (static_cast<uint32_t>(smem_bytes) / stages) & ~127u;
    const int per_row = DV * 16 + nnz_per_row * 4 + 4;
    int cap = (static_cast<int>(L.stage_bytes) - 96) / per_row;   // 96: alignment pads of the two index slices
    if (cap < 0) cap = 0;
    L.cap_rows = cap;
    L.cap_nnz = cap * nnz_per_row;
    L.off_col = static_cast<uint32_t>(cap) * DV * 16u;
    L.off_rp = L.off_col + ((((static_cast<uint32_t>(L.cap_nnz) + 8u) * 4u) + 15u) & ~15u);
    return L;
}

template <int LANES, int VEC, bool STAGED>
struct TileView {
    const int32_t *__restrict__ row_ptr;
    const int32_t *__restrict__ col;
    const float4 *__restrict__ x;
    const float4 *sx;
    const int32_t *s_col;
    const int32_t *s_rp;
    int r0, r1, e0, a0, c0, nst;
    static constexpr int DV = LANES * VEC;
    __device__ __forceinline__ int rp(int i) const { return STAGED ? s_rp[i - a0] : __ldg(row_ptr + i); }
    __device__ __forceinline__ int colv(int q) const {
        return (STAGED && q - e0 < nst) ? s_col[q - c0] : __ldg(col + q);
    }
    __device__ __forceinline__ float4 xv(int c, int j) const {
        return (STAGED && c >= r0 && c < r1) ? sx[static_cast<size_t>(c - r0) * DV + j]
                                             : ldg4(x + static_cast<int64_t>(c) * DV + j);
    }
};

// all rows of one tile, consumer warp cw; see the header comment for the three row classes
template <int LANES, int VEC, bool STAGED>
__device__ __forceinline__ void process_tile(const TileView<LANES, VEC, STAGED> &tv, float4 *__restrict__ out,
                                             float self_scale, bool cut, int cw, int lane) {
    constexpr int RPW = 32 / LANES;   // rows per warp iteration
    constexpr int DV = LANES * VEC;
    constexpr int U = (VEC == 1) ? 4 : 2;
    const int sub = lane / LANES, sl = lane % LANES;
    for (int rb = tv.r0 + cw * RPW; rb < tv.r1; rb += TP_NCW * RPW) {
        const int row = rb + sub;
        const bool valid = row < tv.r1;
        const int beg = valid ? tv.rp(row) : 0, end = valid ? tv.rp(row + 1) : 0;
        const bool big = valid && (end - beg > TP_SPLIT);
        const bool seq = valid && (!big || (LANES == 32 && !cut));
        if (seq) {
            float4 acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = zero4();
            for (int p = beg; p < end; p += U) {
                int c[U];
#pragma unroll
                for (int u = 0; u < U; ++u) c[u] = (p + u < end) ? tv.colv(p + u) : -1;
                float4 v[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (c[u] < 0) continue;
#pragma unroll
                    for (int k = 0; k < VEC; ++k) v[u][k] = tv.xv(c[u], sl + k * LANES);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (c[u] >= 0) add4(acc[k], v[u][k]);
            }
            if (self_scale != 0.f) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) axpy4_rn(acc[k], self_scale, tv.xv(row, sl + k * LANES));
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) out[static_cast<int64_t>(row) * DV + sl + k * LANES] = acc[k];
        }
        if constexpr (LANES < 32) {
            // long rows of this warp iteration, one after the other, split over the RPW sub-groups
            unsigned m = __ballot_sync(0xffffffffu, big && !cut);
            while (m) {
                const int src_lane = __ffs(m) - 1;   // first lane of the owning sub-group
                m &= ~(((LANES == 32) ? 0xffffffffu : ((1u << LANES) - 1u)) << src_lane);
                const int brow = rb + src_lane / LANES;
                const int bbeg = __shfl_sync(0xffffffffu, beg, src_lane), bend = __shfl_sync(0xffffffffu, end, src_lane);
                float4 acc = zero4();
                for (int p = bbeg + sub; p < bend; p += RPW * 4) {
                    int c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = (p + u * RPW < bend) ? tv.colv(p + u * RPW) : -1;
                    float4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (c[u] >= 0) v[u] = tv.xv(c[u], sl);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (c[u] >= 0) add4(acc, v[u]);
                }
#pragma unroll
                for (int o = LANES; o < 32; o <<= 1) {
                    float4 t;
                    t.x = __shfl_xor_sync(0xffffffffu, acc.x, o); t.y = __shfl_xor_sync(0xffffffffu, acc.y, o);
                    t.z = __shfl_xor_sync(0xffffffffu, acc.z, o); t.w = __shfl_xor_sync(0xffffffffu, acc.w, o);
                    add4(acc, t);
                }
                if (sub == 0) {
                    if (self_scale != 0.f) axpy4_rn(acc, self_scale, tv.xv(brow, sl));
                    out[static_cast<int64_t>(brow) * DV + sl] = acc;
                }
            }
        }
    }
}

// one listed long row of a cut tile, reduced by all consumer warps of the CTA through `scratch` (an idle stage)
template <int LANES, int VEC>
__device__ __forceinline__ void process_heavy_row(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                                  const float4 *__restrict__ x, float4 *__restrict__ out, int row,
                                                  float self_scale, float4 *scratch, int cw, int lane) {
    constexpr int RPW = 32 / LANES;
    constexpr int DV = LANES * VEC;
    constexpr int SG = TP_NCW * RPW;   // sub-groups in the CTA
    constexpr int UH = (VEC == 4) ? 2 : 4;
    const int sub = lane / LANES, sl = lane % LANES;
    const int sg = cw * RPW + sub;
    const int beg = __ldg(row_ptr + row), end = __ldg(row_ptr + row + 1);
    float4 acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = zero4();
    for (int p = beg + sg; p < end; p += SG * UH) {
        int c[UH];
#pragma unroll
        for (int u = 0; u < UH; ++u) c[u] = (p + u * SG < end) ? __ldg(col + p + u * SG) : -1;
        float4 v[UH][VEC];
#pragma unroll
        for (int u = 0; u < UH; ++u)
            if (c[u] >= 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) v[u][k] = ldg4(x + static_cast<int64_t>(c[u]) * DV + sl + k * LANES);
            }
#pragma unroll
        for (int u = 0; u < UH; ++u)
            if (c[u] >= 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) add4(acc[k], v[u][k]);
            }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) scratch[static_cast<size_t>(sg) * DV + sl + k * LANES] = acc[k];
    consumer_bar_sync();
    if (cw == 0) {
        // sub-group `sub` adds the partials sub, sub+RPW, ... in ascending order, then a fixed butterfly over sub-groups
#pragma unroll
        for (int k = 0; k < VEC; ++k) acc[k] = zero4();
        for (int g = sub; g < SG; g += RPW) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) add4(acc[k], scratch[static_cast<size_t>(g) * DV + sl + k * LANES]);
        }
#pragma unroll
        for (int o = LANES; o < 32; o <<= 1) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float4 t;
                t.x = __shfl_xor_sync(0xffffffffu, acc[k].x, o); t.y = __shfl_xor_sync(0xffffffffu, acc[k].y, o);
                t.z = __shfl_xor_sync(0xffffffffu, acc[k].z, o); t.w = __shfl_xor_sync(0xffffffffu, acc[k].w, o);
                add4(acc[k], t);
            }
        }
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (self_scale != 0.f) axpy4_rn(acc[k], self_scale, ldg4(x + static_cast<int64_t>(row) * DV + sl + k * LANES));
                out[static_cast<int64_t>(row) * DV + sl + k * LANES] = acc[k];
            }
        }
    }
    // the scratch words were written through the generic proxy and the stage is refilled through the async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(TP_THREADS, 1)
spmm_pipe_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const float4 *__restrict__ x,
                 float4 *__restrict__ out, const int4 *__restrict__ tiles, int num_tiles,
                 const int32_t *__restrict__ heavy_list, const int32_t *__restrict__ heavy_count, float self_scale,
                 int smem_bytes, int stages, int nnz_per_row) {
    constexpr int DV = LANES * VEC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[TP_MAX_STAGES], empty_bar[TP_MAX_STAGES];
    const TileCfg L = tile_cfg(smem_bytes, stages, DV, nnz_per_row);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], TP_NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int H = (heavy_list != nullptr && heavy_count != nullptr) ? __ldg(heavy_count) : 0;
    const int total = H + num_tiles;

    if (warp == 0) {
        // ------------------------------------------------------------------ producer (one lane)
        if (lane != 0) return;
        int s = 0;
        uint32_t use = 0;   // how many times stage s has been filled before
        for (int t = blockIdx.x; t < total; t += gridDim.x) {
            if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1u);
            bool issued = false;
            if (t >= H) {
                const int4 td = __ldg(tiles + (t - H));
                const int r0 = td.x, r1 = td.y, e0 = td.z, e1 = td.w & 0x7fffffff;
                const int rows = r1 - r0;
                if (rows > 0 && rows <= L.cap_rows) {
                    unsigned char *base = smem_raw + static_cast<size_t>(s) * L.stage_bytes;
                    const uint32_t xbytes = static_cast<uint32_t>(rows) * DV * 16u;
                    const int a0 = r0 & ~3;
                    const uint32_t rp_bytes = static_cast<uint32_t>(((r1 + 1 - a0) + 3) & ~3) * 4u;
                    const int c0 = e0 & ~3;
                    const int nst = min(e1 - e0, L.cap_nnz);
                    const uint32_t col_bytes = nst > 0 ? static_cast<uint32_t>(((e0 + nst - c0) + 3) & ~3) * 4u : 0u;
                    mbar_expect_tx(&full_bar[s], xbytes + rp_bytes + col_bytes);
                    bulk_g2s(base + L.off_rp, row_ptr + a0, rp_bytes, &full_bar[s]);
                    if (col_bytes) bulk_g2s(base + L.off_col, col + c0, col_bytes, &full_bar[s]);
                    const char *src = reinterpret_cast<const char *>(x + static_cast<int64_t>(r0) * DV);
                    for (uint32_t off = 0; off < xbytes; off += TP_BULK_CHUNK) {
                        const uint32_t n = xbytes - off < TP_BULK_CHUNK ? xbytes - off : TP_BULK_CHUNK;
                        bulk_g2s(base + off, src + off, n, &full_bar[s]);
                    }
                    issued = true;
                }
            }
            if (!issued) mbar_arrive(&full_bar[s]);
            if (++s == stages) { s = 0; ++use; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int cw = warp - 1;
    int s = 0;
    uint32_t use = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
        unsigned char *base = smem_raw + static_cast<size_t>(s) * L.stage_bytes;
        if (t < H) {
            const int row = __ldg(heavy_list + t);
            mbar_wait(&full_bar[s], use & 1u);
            process_heavy_row<LANES, VEC>(row_ptr, col, x, out, row, self_scale, reinterpret_cast<float4 *>(base), cw, lane);
        } else {
            const int4 td = __ldg(tiles + (t - H));
            const int rows = td.y - td.x;
            mbar_wait(&full_bar[s], use & 1u);
            if (rows > 0) {
                const bool cut = heavy_list != nullptr && td.w < 0;
                if (rows <= L.cap_rows) {
                    TileView<LANES, VEC, true> tv;
                    tv.row_ptr = row_ptr; tv.col = col; tv.x = x;
                    tv.sx = reinterpret_cast<const float4 *>(base);
                    tv.s_col = reinterpret_cast<const int32_t *>(base + L.off_col);
                    tv.s_rp = reinterpret_cast<const int32_t *>(base + L.off_rp);
                    tv.r0 = td.x; tv.r1 = td.y; tv.e0 = td.z; tv.a0 = td.x & ~3; tv.c0 = td.z & ~3;
                    tv.nst = min((td.w & 0x7fffffff) - td.z, L.cap_nnz);
                    process_tile<LANES, VEC, true>(tv, out, self_scale, cut, cw, lane);
                } else {
                    TileView<LANES, VEC, false> tv;
                    tv.row_ptr = row_ptr; tv.col = col; tv.x = x;
                    tv.sx = nullptr; tv.s_col = nullptr; tv.s_rp = nullptr;
                    tv.r0 = td.x; tv.r1 = td.y; tv.e0 = td.z; tv.a0 = 0; tv.c0 = 0; tv.nst = 0;
                    process_tile<LANES, VEC, false>(tv, out, self_scale, cut, cw, lane);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == stages) { s = 0; ++use; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// tiling.  boundary(k) = first graph start in [kC, (k+1)C) if there is one (aligned), else kC (cut inside a graph
// longer than the window); boundary(T) = N.  tile k = [boundary(k), boundary(k+1)), desc = {r0, r1, e0, e1 | cut<<31}.
__device__ __forceinline__ int tile_boundary(const int32_t *__restrict__ seg_ptr, int B, int C, int N, int k, int T,
                                             bool *aligned) {
    if (k >= T) { *aligned = true; return N; }
    const int64_t target = static_cast<int64_t>(k) * C;
    int lo = 0, hi = B;   // first g in [0, B] with seg_ptr[g] >= target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (seg_ptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    const int gs = seg_ptr[lo];
    if (gs < target + C) { *aligned = true; return gs; }
    *aligned = false;
    return static_cast<int>(target);
}

__global__ void make_row_tiles_kernel(const int32_t *__restrict__ seg_ptr, int B, int C, const int32_t *__restrict__ row_ptr,
                                      int N, int4 *__restrict__ tiles, int T) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    bool a0, a1;
    const int r0 = tile_boundary(seg_ptr, B, C, N, k, T, &a0);
    const int r1 = tile_boundary(seg_ptr, B, C, N, k + 1, T, &a1);
    const int e0 = row_ptr[r0], e1 = row_ptr[r1];
    const bool cut = !(a0 && a1);
    tiles[k] = make_int4(r0, r1, e0, e1 | (cut ? static_cast<int>(0x80000000u) : 0));
}

// rows with more than TP_SPLIT neighbours that live in cut tiles -> heavy_list (order irrelevant, see kernel)
__global__ void collect_cut_heavy_kernel(const int32_t *__restrict__ row_ptr, int N, int C, const int4 *__restrict__ tiles,
                                         int T, int32_t *__restrict__ heavy_list, int cap, int32_t *__restrict__ heavy_count) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= N) return;
    if (row_ptr[r + 1] - row_ptr[r] <= TP_SPLIT) return;
    int k = r / C;
    if (k >= T) k = T - 1;
    if (r < tiles[k].x) --k;
    if (tiles[k].w < 0) {
        const int i = atomicAdd(heavy_count, 1);
        if (i < cap) heavy_list[i] = r;
    }
}

extern "C" int dn4gl_make_row_tiles(const int32_t *seg_ptr, int32_t B, int32_t window_rows, const int32_t *row_ptr,
                                    int64_t N, int32_t *tile_desc, int32_t num_tiles, int32_t *heavy_list,
                                    int32_t heavy_cap, int32_t *heavy_count, void *stream) {
    DN_ARG(seg_ptr && row_ptr && tile_desc && B >= 0 && window_rows > 0 && num_tiles >= 0 && N >= 0 && N < (1ll << 31));
    DN_ARG(static_cast<int64_t>(num_tiles) * window_rows >= N && aligned16(tile_desc));
    cudaStream_t st = as_stream(stream);
    int launched = 0;
    if (heavy_count) DN_CUDA(cudaMemsetAsync(heavy_count, 0, sizeof(int32_t), st));
    if (num_tiles > 0) {
        make_row_tiles_kernel<<<(num_tiles + 255) / 256, 256, 0, st>>>(seg_ptr, B, window_rows, row_ptr, static_cast<int>(N),
                                                                      reinterpret_cast<int4 *>(tile_desc), num_tiles);
        ++launched;
        if (heavy_list && heavy_count && N > 0) {
            collect_cut_heavy_kernel<<<static_cast<unsigned>(ceil_div64(N, 256)), 256, 0, st>>>(
                row_ptr, static_cast<int>(N), window_rows, reinterpret_cast<const int4 *>(tile_desc), num_tiles, heavy_list,
                heavy_cap, heavy_count);
            ++launched;
        }
    }
    if (launched) DN_LAUNCHED_N(launched);
    return DN4GL_OK;
}

/* rows of width D that one pipeline stage holds for (smem_bytes, stages, nnz_per_row) */
extern "C" int32_t dn4gl_spmm_tiled_cap_rows(int32_t D, int32_t smem_bytes, int32_t stages, int32_t nnz_per_row) {
    if (D <= 0 || D % 4 != 0 || stages < 1 || stages > TP_MAX_STAGES || nnz_per_row < 1) return 0;
    return tile_cfg(smem_bytes, stages, D / 4, nnz_per_row).cap_rows;
}

template <int LANES, int VEC>
static int launch_pipe(const int32_t *row_ptr, const int32_t *col, const float *x, float *out, const int32_t *tile_desc,
                       int num_tiles, const int32_t *heavy_list, const int32_t *heavy_count, int heavy_cap,
                       float self_scale, int smem_bytes, int stages, int nnz_per_row, cudaStream_t st) {
    static int attr_done = 0;
    if (attr_done < smem_bytes) {
        if (cudaFuncSetAttribute(spmm_pipe_kernel<LANES, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
            cudaSuccess)
            return -1;
        attr_done = smem_bytes;
    }
    const int per_sm = 1;   // one persistent CTA per SM: 15 consumer warps, up to 128 registers, the whole shared memory as ring
    int64_t want = static_cast<int64_t>(num_tiles) + (heavy_list ? heavy_cap : 0);
    int grid = static_cast<int>(want < dn4gl_num_sms() * per_sm ? want : dn4gl_num_sms() * per_sm);
    if (grid < 1) grid = 1;
    spmm_pipe_kernel<LANES, VEC><<<grid, TP_THREADS, smem_bytes, st>>>(
        row_ptr, col, reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(out),
        reinterpret_cast<const int4 *>(tile_desc), num_tiles, heavy_list, heavy_count, self_scale, smem_bytes, stages,
        nnz_per_row);
    return 0;
}

extern "C" int dn4gl_spmm_tiled_f32(const int32_t *row_ptr, const int32_t *col, const float *x, float *out, int64_t N,
                                    int32_t D, float self_scale, const int32_t *tile_desc, int32_t num_tiles,
                                    const int32_t *heavy_list, const int32_t *heavy_count, int32_t heavy_cap,
                                    int32_t smem_bytes, int32_t stages, int32_t nnz_per_row, void *stream) {
    DN_ARG(N >= 0 && D > 0 && D % 4 == 0 && num_tiles >= 0 && smem_bytes >= 16 * 1024 && smem_bytes <= 220 * 1024);
    DN_ARG(stages >= 1 && stages <= TP_MAX_STAGES && nnz_per_row >= 1 && nnz_per_row <= 64);
    if (N == 0 || num_tiles == 0) return DN4GL_OK;
    DN_ARG(row_ptr && col && x && out && tile_desc && aligned16(x) && aligned16(out) && aligned16(row_ptr) &&
           aligned16(col) && aligned16(tile_desc));
    DN_ARG((heavy_list == nullptr) == (heavy_count == nullptr));
    {   // the CTA-wide reduction of listed rows uses one stage as scratch: 15 warps x (32/LANES) partial rows
        const int dv = D / 4, lanes = dv < 32 ? dv : 32;
        const int64_t scratch = static_cast<int64_t>(TP_NCW) * (32 / lanes) * dv * 16;
        DN_ARG(heavy_list == nullptr || scratch <= ((smem_bytes / stages) & ~127));
    }
    cudaStream_t st = as_stream(stream);
    int rc = 0;
#define PIPE_CASE(L, V)                                                                                              \
    rc = launch_pipe<L, V>(row_ptr, col, x, out, tile_desc, num_tiles, heavy_list, heavy_count, heavy_cap, self_scale, \
                           smem_bytes, stages, nnz_per_row, st);                                                     \
    break
    switch (D / 4) {
        case 4: PIPE_CASE(4, 1);
        case 8: PIPE_CASE(8, 1);
        case 16: PIPE_CASE(16, 1);
        case 32: PIPE_CASE(32, 1);
        case 64: PIPE_CASE(32, 2);
        case 128: PIPE_CASE(32, 4);
        default:
            dn4gl_set_error("dn4gl_spmm_tiled_f32: unsupported D=%d (supported: 16,32,64,128,256,512)", D);
            return DN4GL_EINVAL;
    }
#undef PIPE_CASE
    if (rc != 0) {
        dn4gl_set_error("dn4gl_spmm_tiled_f32: cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d) failed", smem_bytes);
        return DN4GL_ECUDA;
    }
    DN_LAUNCHED();
    return DN4GL_OK;
}
