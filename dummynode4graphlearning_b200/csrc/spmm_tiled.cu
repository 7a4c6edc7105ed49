// libdn4gl.so -- K1, tiled variant: sum aggregation with the neighbour rows staged through shared memory.
//
// A mini-batch is block-diagonal: every neighbour of a row of graph g is a row of graph g.  Rows are therefore cut into
// TILES of whole consecutive graphs (dn4gl_make_row_tiles); one CTA owns a tile, pulls the tile's feature rows
// x[r0:r1) -- one contiguous slab -- into shared memory with bulk asynchronous copies (cp.async.bulk ... mbarrier
// complete_tx, i.e. the TMA engine's 1-D path; SASS: UBLKCP) and then resolves every neighbour index against shared
// memory.  DRAM sees each feature row exactly once as a streaming read and each output row once as a streaming write:
// the algorithmic bytes of SURVEY.md section 8(d).  There are no dependent DRAM gathers (the latency chain that bounded
// the per-row kernel, profiles/r1a) and dummy rows (degree = graph size) cost shared-memory reads only.
// Indices that fall outside the staged window (never for graph-aligned tiles) and tiles larger than the shared-memory
// budget fall back to global loads inside the same kernel, so the result is correct for ANY CSR; tiles are a
// performance contract, not a correctness one.  Accumulation order per row = CSR order with separately rounded adds
// (bit-identical to the per-row kernel and to the sequential oracle) except rows above HEAVY_SPLIT neighbours, which
// the whole CTA reduces with a fixed tree.
#include "common.cuh"

constexpr int TILED_THREADS = 512;
constexpr int HEAVY_SPLIT = 96;       // rows with more neighbours than this are reduced by the whole CTA
constexpr int NNZ_PER_ROW_BUDGET = 8; // shared-memory slots reserved for column indices, per staged row

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

// shared-memory carve-up for a given budget: rows of x, their column indices, row_ptr slice, hub-row list
struct TileSmem {
    int cap_rows, cap_nnz;
    size_t off_col, off_rp, off_heavy, total;
};
__host__ __device__ inline TileSmem tile_smem_layout(int smem_bytes, int DV) {
    TileSmem L;
    const int per_row = DV * 16 + NNZ_PER_ROW_BUDGET * 4 + 8;
    L.cap_rows = (smem_bytes - 64) / per_row;
    L.cap_nnz = L.cap_rows * NNZ_PER_ROW_BUDGET;
    L.off_col = static_cast<size_t>(L.cap_rows) * DV * 16;
    L.off_rp = L.off_col + static_cast<size_t>(L.cap_nnz) * 4;
    L.off_heavy = L.off_rp + static_cast<size_t>(L.cap_rows + 1) * 4;
    L.total = L.off_heavy + static_cast<size_t>(L.cap_rows) * 4;
    return L;
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(TILED_THREADS)
spmm_tiled_kernel(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ col, const float4 *__restrict__ x,
                  float4 *__restrict__ out, const int32_t *__restrict__ tile_ptr, int num_tiles, float self_scale,
                  int smem_bytes) {
    constexpr int SUBS = TILED_THREADS / LANES;
    constexpr int DV = LANES * VEC;          // float4 per row
    constexpr int U = (VEC == 1) ? 8 : (VEC == 2 ? 4 : 2);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const TileSmem L = tile_smem_layout(smem_bytes, DV);
    float4 *sx = reinterpret_cast<float4 *>(smem_raw);                       // [cap_rows][DV] staged feature rows
    int32_t *s_col = reinterpret_cast<int32_t *>(smem_raw + L.off_col);      // [cap_nnz]
    int32_t *s_rp = reinterpret_cast<int32_t *>(smem_raw + L.off_rp);        // [cap_rows + 1]
    int32_t *heavy_list = reinterpret_cast<int32_t *>(smem_raw + L.off_heavy);  // [cap_rows]
    __shared__ uint64_t bar;
    __shared__ int heavy_n;
    __shared__ float4 part[TILED_THREADS * VEC];
    const int sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int r0 = tile_ptr[t], r1 = tile_ptr[t + 1];
        const int rows = r1 - r0;
        if (rows <= 0) continue;
        const bool staged = rows <= L.cap_rows;
        if (threadIdx.x == 0) heavy_n = 0;
        if (staged && threadIdx.x == 0) {
            // the previous tile's generic-proxy reads of sx are ordered before this async-proxy write by the
            // __syncthreads that closes the loop body plus this proxy fence
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            const uint32_t bytes = static_cast<uint32_t>(rows) * DV * 16u;
            mbar_expect_tx(&bar, bytes);
            const char *src = reinterpret_cast<const char *>(x + static_cast<int64_t>(r0) * DV);
            char *dst = reinterpret_cast<char *>(sx);
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
                bulk_g2s(dst + off, src + off, n, &bar);
            }
        }
        // ---- stage the CSR slice of the tile with coalesced loads while the bulk copy is in flight
        const int e0 = __ldg(row_ptr + r0), e1 = __ldg(row_ptr + r1);
        const int nnz_staged = min(e1 - e0, L.cap_nnz);
        if (staged) {
            for (int i = threadIdx.x; i <= rows; i += TILED_THREADS) s_rp[i] = __ldg(row_ptr + r0 + i);
            for (int i = threadIdx.x; i < nnz_staged; i += TILED_THREADS) s_col[i] = __ldg(col + e0 + i);
        }
        __syncthreads();
        if (staged) {
            mbar_wait(&bar, phase);
            phase ^= 1;
        }
        // ---- per-row pass: sub-group `sub` takes rows r0+sub, r0+sub+SUBS, ...
        for (int row = r0 + sub; row < r1; row += SUBS) {
            const int beg = staged ? s_rp[row - r0] : __ldg(row_ptr + row);
            const int end = staged ? s_rp[row - r0 + 1] : __ldg(row_ptr + row + 1);
            if (staged && end - beg > HEAVY_SPLIT) {   // hub row (dummy node): deferred to the CTA-wide pass
                if (lane == 0) heavy_list[atomicAdd(&heavy_n, 1)] = row;
                continue;
            }
            float4 acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = zero4();
            for (int p = beg; p < end; p += U) {
                int c[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int q = p + u;
                    c[u] = (q < end) ? ((staged && q - e0 < nnz_staged) ? s_col[q - e0] : __ldg(col + q)) : -1;
                }
                float4 v[U][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (c[u] < 0) continue;
                    const bool in_win = staged && c[u] >= r0 && c[u] < r1;
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        v[u][k] = in_win ? sx[static_cast<size_t>(c[u] - r0) * DV + lane + k * LANES]
                                         : ldg4(x + static_cast<int64_t>(c[u]) * DV + lane + k * LANES);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int k = 0; k < VEC; ++k)
                        if (c[u] >= 0) add4(acc[k], v[u][k]);
            }
            if (self_scale != 0.f) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float4 s = staged ? sx[static_cast<size_t>(row - r0) * DV + lane + k * LANES]
                                      : ldg4(x + static_cast<int64_t>(row) * DV + lane + k * LANES);
                    axpy4_rn(acc[k], self_scale, s);
                }
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) out[static_cast<int64_t>(row) * DV + lane + k * LANES] = acc[k];
        }
        __syncthreads();
        // ---- hub rows: the whole CTA strides the neighbour list, fixed-shape tree over sub-groups
        const int nh = heavy_n;
        for (int h = 0; h < nh; ++h) {
            const int row = heavy_list[h];
            const int beg = s_rp[row - r0], end = s_rp[row - r0 + 1];
            float4 acc[VEC];
#pragma unroll
            for (int k = 0; k < VEC; ++k) acc[k] = zero4();
            for (int p = beg + sub; p < end; p += SUBS) {
                const int c = (p - e0 < nnz_staged) ? s_col[p - e0] : __ldg(col + p);
                const bool in_win = c >= r0 && c < r1;
#pragma unroll
                for (int k = 0; k < VEC; ++k)
                    add4(acc[k], in_win ? sx[static_cast<size_t>(c - r0) * DV + lane + k * LANES]
                                        : ldg4(x + static_cast<int64_t>(c) * DV + lane + k * LANES));
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) part[(sub * VEC + k) * LANES + lane] = acc[k];
            __syncthreads();
#pragma unroll
            for (int s = SUBS / 2; s >= 1; s >>= 1) {
                if (sub < s) {
#pragma unroll
                    for (int k = 0; k < VEC; ++k) {
                        float4 a = part[(sub * VEC + k) * LANES + lane];
                        add4(a, part[((sub + s) * VEC + k) * LANES + lane]);
                        part[(sub * VEC + k) * LANES + lane] = a;
                    }
                }
                __syncthreads();
            }
            if (sub == 0) {
#pragma unroll
                for (int k = 0; k < VEC; ++k) {
                    float4 a = part[k * LANES + lane];
                    if (self_scale != 0.f) axpy4_rn(a, self_scale, sx[static_cast<size_t>(row - r0) * DV + lane + k * LANES]);
                    out[static_cast<int64_t>(row) * DV + lane + k * LANES] = a;
                }
            }
            __syncthreads();
        }
        __syncthreads();   // all reads of sx / s_col done before the next tile overwrites them
    }
}

// tile k = graphs whose first row lies in [k*C, (k+1)*C): tile_ptr[k] = first graph start >= k*C (seg_ptr[B] at the end)
__global__ void make_row_tiles_kernel(const int32_t *__restrict__ seg_ptr, int B, int C, int32_t *__restrict__ tile_ptr,
                                      int T) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > T) return;
    int64_t target = static_cast<int64_t>(k) * C;
    int lo = 0, hi = B;   // first g in [0, B] with seg_ptr[g] >= target
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (seg_ptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    tile_ptr[k] = seg_ptr[lo];
}

extern "C" int dn4gl_make_row_tiles(const int32_t *seg_ptr, int32_t B, int32_t window_rows, int32_t *tile_ptr,
                                    int32_t num_tiles, void *stream) {
    DN_ARG(seg_ptr && tile_ptr && B >= 0 && window_rows > 0 && num_tiles >= 0);
    make_row_tiles_kernel<<<(num_tiles + 1 + 255) / 256, 256, 0, as_stream(stream)>>>(seg_ptr, B, window_rows, tile_ptr,
                                                                                     num_tiles);
    DN_LAUNCHED();
    return DN4GL_OK;
}

/* rows of width D that one tile may hold for a given dynamic shared-memory budget (for dn4gl_make_row_tiles) */
extern "C" int32_t dn4gl_spmm_tiled_cap_rows(int32_t D, int32_t smem_bytes) {
    if (D <= 0 || D % 4 != 0) return 0;
    return tile_smem_layout(smem_bytes, D / 4).cap_rows;
}

template <int LANES, int VEC>
static int launch_tiled(const int32_t *row_ptr, const int32_t *col, const float *x, float *out,
                        const int32_t *tile_ptr, int num_tiles, float self_scale, int smem_bytes, cudaStream_t st) {
    constexpr int DV = LANES * VEC;
    static int attr_done = 0;
    if (attr_done < smem_bytes) {
        if (cudaFuncSetAttribute(spmm_tiled_kernel<LANES, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) !=
            cudaSuccess)
            return -1;
        attr_done = smem_bytes;
    }
    int grid = num_tiles < dn4gl_num_sms() * 2 ? num_tiles : dn4gl_num_sms() * 2;
    if (grid < 1) grid = 1;
    spmm_tiled_kernel<LANES, VEC><<<grid, TILED_THREADS, smem_bytes, st>>>(
        row_ptr, col, reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(out), tile_ptr, num_tiles,
        self_scale, smem_bytes);
    return 0;
}

extern "C" int dn4gl_spmm_tiled_f32(const int32_t *row_ptr, const int32_t *col, const float *x, float *out, int64_t N,
                                    int32_t D, float self_scale, const int32_t *tile_ptr, int32_t num_tiles,
                                    int32_t smem_bytes, void *stream) {
    DN_ARG(N >= 0 && D > 0 && D % 4 == 0 && num_tiles >= 0 && smem_bytes >= 16 * 1024 && smem_bytes <= 190 * 1024);
    if (N == 0 || num_tiles == 0) return DN4GL_OK;
    DN_ARG(row_ptr && col && x && out && tile_ptr && aligned16(x) && aligned16(out));
    cudaStream_t st = as_stream(stream);
    int rc = 0;
    switch (D / 4) {
        case 4: rc = launch_tiled<4, 1>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        case 8: rc = launch_tiled<8, 1>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        case 16: rc = launch_tiled<16, 1>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        case 32: rc = launch_tiled<32, 1>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        case 64: rc = launch_tiled<32, 2>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        case 128: rc = launch_tiled<32, 4>(row_ptr, col, x, out, tile_ptr, num_tiles, self_scale, smem_bytes, st); break;
        default:
            dn4gl_set_error("dn4gl_spmm_tiled_f32: unsupported D=%d (supported: 16,32,64,128,256,512)", D);
            return DN4GL_EINVAL;
    }
    if (rc != 0) {
        dn4gl_set_error("dn4gl_spmm_tiled_f32: cudaFuncSetAttribute(MaxDynamicSharedMemorySize=%d) failed", smem_bytes);
        return DN4GL_ECUDA;
    }
    DN_LAUNCHED();
    return DN4GL_OK;
}
