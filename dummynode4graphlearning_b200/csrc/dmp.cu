// libdn4gl.so -- K4 / K5: DMPNN dual message passing (subgraph_isomorphism/models/dmpnn.py:111-156)
// restructured by linearity so the per-edge work is pure gather / segment-sum (HBM-bound) and the
// dense weights are applied by node-/edge-level GEMMs outside:
//   node:  agg = S_rev @ W_out - S_fwd @ W_in,   S_*[v] = sum_{e in in(v), rev_e = *} ef[e]
//   edge:  out = T[:, :D] + c_e * T[:, D:] + msg_e + b,   T = ef @ [W_eloop | W_src - W_dst],
//          msg_e = rev_e ? P[src]-Q[dst] : P[dst]-Q[src],  [P|Q] = h @ [W_dst | W_src],
//          c_e = 2 * (1 + log2(1 + out_deg[dst_e]))
#include "common.cuh"

// -------------------------------------------------------------------------------------------
// K4 forward.  Sub-group of LANES lanes per node, in-list walked in CSR (= edge id) order; each
// edge row of ef is read exactly once over the whole launch.
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
dmp_node_agg_kernel(const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid,
                    const uint8_t *__restrict__ is_rev, const float4 *__restrict__ ef, float4 *__restrict__ S,
                    int64_t N, int heavy_thr) {
    constexpr int ROWS = 256 / LANES;
    constexpr int DV = LANES * VEC;
    constexpr int U = (VEC == 1) ? 4 : 2;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * ROWS + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    if (row >= N) return;
    const int beg = __ldg(in_ptr + row), end = __ldg(in_ptr + row + 1);
    if (heavy_thr > 0 && end - beg > heavy_thr) return;
    float4 ar[VEC], af[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { ar[k] = zero4(); af[k] = zero4(); }
    for (int p = beg; p < end; p += U) {
        int e[U];
        bool rv[U];
        float4 v[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            e[u] = (p + u < end) ? __ldg(in_eid + p + u) : -1;
            rv[u] = (e[u] >= 0) ? (is_rev != nullptr && is_rev[e[u]] != 0) : false;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (e[u] >= 0) v[u][k] = ldg4(ef + static_cast<int64_t>(e[u]) * DV + lane + k * LANES);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (e[u] >= 0) { if (rv[u]) add4(ar[k], v[u][k]); else add4(af[k], v[u][k]); }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        S[row * (2 * DV) + lane + k * LANES] = ar[k];
        S[row * (2 * DV) + DV + lane + k * LANES] = af[k];
    }
}

template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
dmp_node_agg_heavy(const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid,
                   const uint8_t *__restrict__ is_rev, const float4 *__restrict__ ef, float4 *__restrict__ S,
                   const int32_t *__restrict__ heavy_rows, const int32_t *__restrict__ heavy_count) {
    constexpr int SUBS = 256 / LANES;
    constexpr int DV = LANES * VEC;
    __shared__ float4 part[2 * 256 * VEC];
    const int sub = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const int n_heavy = *heavy_count;
    for (int h = blockIdx.x; h < n_heavy; h += gridDim.x) {
        const int64_t row = heavy_rows[h];
        const int beg = in_ptr[row], end = in_ptr[row + 1];
        float4 ar[VEC], af[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) { ar[k] = zero4(); af[k] = zero4(); }
        for (int p = beg + sub; p < end; p += SUBS) {
            int e = __ldg(in_eid + p);
            bool rv = is_rev != nullptr && is_rev[e] != 0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                float4 v = ldg4(ef + static_cast<int64_t>(e) * DV + lane + k * LANES);
                if (rv) add4(ar[k], v); else add4(af[k], v);
            }
        }
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            part[((sub * 2 + 0) * VEC + k) * LANES + lane] = ar[k];
            part[((sub * 2 + 1) * VEC + k) * LANES + lane] = af[k];
        }
        __syncthreads();
#pragma unroll
        for (int s = SUBS / 2; s >= 1; s >>= 1) {
            if (sub < s) {
#pragma unroll
                for (int q = 0; q < 2 * VEC; ++q) {
                    float4 a = part[(sub * 2 * VEC + q) * LANES + lane];
                    add4(a, part[((sub + s) * 2 * VEC + q) * LANES + lane]);
                    part[(sub * 2 * VEC + q) * LANES + lane] = a;
                }
            }
            __syncthreads();
        }
        if (sub == 0) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                S[row * (2 * DV) + lane + k * LANES] = part[(0 * VEC + k) * LANES + lane];
                S[row * (2 * DV) + DV + lane + k * LANES] = part[(1 * VEC + k) * LANES + lane];
            }
        }
        __syncthreads();
    }
}

extern "C" int dn4gl_dmp_node_agg_f32(const int32_t *in_ptr, const int32_t *in_eid, const uint8_t *is_rev,
                                      const float *ef, float *S, int64_t N, int32_t D, const int32_t *heavy_rows,
                                      const int32_t *heavy_count, int32_t heavy_threshold, void *stream) {
    DN_ARG(N >= 0 && D > 0 && D % 4 == 0);
    if (N == 0) return DN4GL_OK;
    DN_ARG(in_ptr && in_eid && ef && S && aligned16(ef) && aligned16(S));
    cudaStream_t st = as_stream(stream);
    const bool heavy = heavy_rows && heavy_count && heavy_threshold > 0;
    const int thr = heavy ? heavy_threshold : 0;
    const unsigned hgrid = dn4gl_num_sms() * 4;
#define DMP_CASE(L, V)                                                                                          \
    dmp_node_agg_kernel<L, V><<<static_cast<unsigned>(ceil_div64(N, 256 / L)), 256, 0, st>>>(                   \
        in_ptr, in_eid, is_rev, reinterpret_cast<const float4 *>(ef), reinterpret_cast<float4 *>(S), N, thr);   \
    if (heavy)                                                                                                  \
        dmp_node_agg_heavy<L, V><<<hgrid, 256, 0, st>>>(in_ptr, in_eid, is_rev,                                 \
                                                        reinterpret_cast<const float4 *>(ef),                   \
                                                        reinterpret_cast<float4 *>(S), heavy_rows, heavy_count); \
    break
    switch (D / 4) {
        case 4: DMP_CASE(4, 1);
        case 8: DMP_CASE(8, 1);
        case 16: DMP_CASE(16, 1);
        case 32: DMP_CASE(32, 1);
        case 64: DMP_CASE(32, 2);
        default:
            dn4gl_set_error("dn4gl_dmp_node_agg_f32: unsupported D=%d (supported: 16,32,64,128,256)", D);
            return DN4GL_EINVAL;
    }
#undef DMP_CASE
    DN_LAUNCHED_N(heavy ? 2 : 1);
    return DN4GL_OK;
}

// K4 backward: per-edge gather of the destination's gradient half.
__global__ void dmp_node_agg_bwd_kernel(const int32_t *__restrict__ dst, const uint8_t *__restrict__ is_rev,
                                        const float4 *__restrict__ gS, float4 *__restrict__ gef, int64_t total,
                                        int DV) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t e = i / DV;
    int c = static_cast<int>(i - e * DV);
    bool rv = is_rev != nullptr && is_rev[e] != 0;
    gef[i] = ldg4(gS + static_cast<int64_t>(dst[e]) * (2 * DV) + (rv ? 0 : DV) + c);
}

extern "C" int dn4gl_dmp_node_agg_bwd_f32(const int32_t *dst, const uint8_t *is_rev, const float *gS, float *gef,
                                          int64_t E, int32_t D, void *stream) {
    DN_ARG(E >= 0 && D > 0 && D % 4 == 0);
    if (E == 0) return DN4GL_OK;
    DN_ARG(dst && gS && gef && aligned16(gS) && aligned16(gef));
    int DV = D / 4;
    int64_t total = E * DV;
    dmp_node_agg_bwd_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, as_stream(stream)>>>(
        dst, is_rev, reinterpret_cast<const float4 *>(gS), reinterpret_cast<float4 *>(gef), total, DV);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// -------------------------------------------------------------------------------------------
// K5 forward: one thread per (edge, float4 column).
__device__ __forceinline__ float edge_coeff(const int32_t *__restrict__ out_deg, int d) {
    // dmpnn.py:144-146: d = log2(1 + out_deg[dst]); coefficient 2 * (1 + d)
    float deg = static_cast<float>(out_deg[d]);
    return 2.f * (1.f + log2f(1.f + deg));
}

__global__ void dmp_edge_update_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                                       const uint8_t *__restrict__ is_rev, const int32_t *__restrict__ out_deg,
                                       const float4 *__restrict__ PQ, const float4 *__restrict__ T,
                                       const float4 *__restrict__ bias, float4 *__restrict__ out, int64_t total,
                                       int DV) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t e = i / DV;
    int c = static_cast<int>(i - e * DV);
    int s = src[e], d = dst[e];
    bool rv = is_rev != nullptr && is_rev[e] != 0;
    int a = rv ? s : d, b = rv ? d : s;  // msg = P[a] - Q[b]
    float4 P = ldg4(PQ + static_cast<int64_t>(a) * (2 * DV) + c);
    float4 Q = ldg4(PQ + static_cast<int64_t>(b) * (2 * DV) + DV + c);
    float4 t0 = ldg4(T + e * (2 * DV) + c);
    float4 t1 = ldg4(T + e * (2 * DV) + DV + c);
    float cf = edge_coeff(out_deg, d);
    // reference order (dmpnn.py:147-149): (eloop + add) + agg (+ bias)
    float4 r;
    r.x = __fadd_rn(__fadd_rn(t0.x, __fmul_rn(cf, t1.x)), __fsub_rn(P.x, Q.x));
    r.y = __fadd_rn(__fadd_rn(t0.y, __fmul_rn(cf, t1.y)), __fsub_rn(P.y, Q.y));
    r.z = __fadd_rn(__fadd_rn(t0.z, __fmul_rn(cf, t1.z)), __fsub_rn(P.z, Q.z));
    r.w = __fadd_rn(__fadd_rn(t0.w, __fmul_rn(cf, t1.w)), __fsub_rn(P.w, Q.w));
    if (bias) {
        float4 bb = ldg4(bias + c);
        r.x = __fadd_rn(r.x, bb.x); r.y = __fadd_rn(r.y, bb.y); r.z = __fadd_rn(r.z, bb.z); r.w = __fadd_rn(r.w, bb.w);
    }
    out[i] = r;
}

extern "C" int dn4gl_dmp_edge_update_f32(const int32_t *src, const int32_t *dst, const uint8_t *is_rev,
                                         const int32_t *out_deg, const float *PQ, const float *T, const float *bias,
                                         float *out, int64_t E, int32_t D, void *stream) {
    DN_ARG(E >= 0 && D > 0 && D % 4 == 0);
    if (E == 0) return DN4GL_OK;
    DN_ARG(src && dst && out_deg && PQ && T && out && aligned16(PQ) && aligned16(T) && aligned16(out));
    DN_ARG(bias == nullptr || aligned16(bias));
    int DV = D / 4;
    int64_t total = E * DV;
    dmp_edge_update_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, as_stream(stream)>>>(
        src, dst, is_rev, out_deg, reinterpret_cast<const float4 *>(PQ), reinterpret_cast<const float4 *>(T),
        reinterpret_cast<const float4 *>(bias), reinterpret_cast<float4 *>(out), total, DV);
    DN_LAUNCHED();
    return DN4GL_OK;
}

__global__ void dmp_edge_bwd_T_kernel(const int32_t *__restrict__ dst, const int32_t *__restrict__ out_deg,
                                      const float4 *__restrict__ g, float4 *__restrict__ gT, int64_t total, int DV) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int64_t e = i / DV;
    int c = static_cast<int>(i - e * DV);
    float cf = edge_coeff(out_deg, dst[e]);
    float4 v = ldg4(g + i);
    gT[e * (2 * DV) + c] = v;
    gT[e * (2 * DV) + DV + c] = make_float4(cf * v.x, cf * v.y, cf * v.z, cf * v.w);
}

extern "C" int dn4gl_dmp_edge_update_bwd_T_f32(const int32_t *dst, const int32_t *out_deg, const float *g, float *gT,
                                               int64_t E, int32_t D, void *stream) {
    DN_ARG(E >= 0 && D > 0 && D % 4 == 0);
    if (E == 0) return DN4GL_OK;
    DN_ARG(dst && out_deg && g && gT && aligned16(g) && aligned16(gT));
    int DV = D / 4;
    int64_t total = E * DV;
    dmp_edge_bwd_T_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, as_stream(stream)>>>(
        dst, out_deg, reinterpret_cast<const float4 *>(g), reinterpret_cast<float4 *>(gT), total, DV);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// K5 backward wrt [P|Q]: deterministic per-node reduction over the node's in- and out-lists.
template <int LANES, int VEC>
__global__ void __launch_bounds__(256)
dmp_edge_bwd_PQ_kernel(const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid,
                       const int32_t *__restrict__ out_ptr, const int32_t *__restrict__ out_eid,
                       const uint8_t *__restrict__ is_rev, const float4 *__restrict__ g, float4 *__restrict__ gPQ,
                       int64_t N) {
    constexpr int ROWS = 256 / LANES;
    constexpr int DV = LANES * VEC;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * ROWS + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    if (row >= N) return;
    float4 gp[VEC], gq[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) { gp[k] = zero4(); gq[k] = zero4(); }
    // in-edges (v is the destination): !rev -> P[v] (+g),  rev -> Q[v] (-g)
    for (int p = __ldg(in_ptr + row), end = __ldg(in_ptr + row + 1); p < end; ++p) {
        int e = __ldg(in_eid + p);
        bool rv = is_rev != nullptr && is_rev[e] != 0;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float4 v = ldg4(g + static_cast<int64_t>(e) * DV + lane + k * LANES);
            if (rv) sub4(gq[k], v); else add4(gp[k], v);
        }
    }
    // out-edges (v is the source): rev -> P[v] (+g),  !rev -> Q[v] (-g)
    for (int p = __ldg(out_ptr + row), end = __ldg(out_ptr + row + 1); p < end; ++p) {
        int e = __ldg(out_eid + p);
        bool rv = is_rev != nullptr && is_rev[e] != 0;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
            float4 v = ldg4(g + static_cast<int64_t>(e) * DV + lane + k * LANES);
            if (rv) add4(gp[k], v); else sub4(gq[k], v);
        }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
        gPQ[row * (2 * DV) + lane + k * LANES] = gp[k];
        gPQ[row * (2 * DV) + DV + lane + k * LANES] = gq[k];
    }
}

extern "C" int dn4gl_dmp_edge_update_bwd_PQ_f32(const int32_t *in_ptr, const int32_t *in_eid, const int32_t *out_ptr,
                                                const int32_t *out_eid, const uint8_t *is_rev, const float *g,
                                                float *gPQ, int64_t N, int32_t D, void *stream) {
    DN_ARG(N >= 0 && D > 0 && D % 4 == 0);
    if (N == 0) return DN4GL_OK;
    DN_ARG(in_ptr && in_eid && out_ptr && out_eid && g && gPQ && aligned16(g) && aligned16(gPQ));
    cudaStream_t st = as_stream(stream);
#define PQ_CASE(L, V)                                                                                        \
    dmp_edge_bwd_PQ_kernel<L, V><<<static_cast<unsigned>(ceil_div64(N, 256 / L)), 256, 0, st>>>(             \
        in_ptr, in_eid, out_ptr, out_eid, is_rev, reinterpret_cast<const float4 *>(g),                       \
        reinterpret_cast<float4 *>(gPQ), N);                                                                 \
    break
    switch (D / 4) {
        case 4: PQ_CASE(4, 1);
        case 8: PQ_CASE(8, 1);
        case 16: PQ_CASE(16, 1);
        case 32: PQ_CASE(32, 1);
        case 64: PQ_CASE(32, 2);
        default:
            dn4gl_set_error("dn4gl_dmp_edge_update_bwd_PQ_f32: unsupported D=%d", D);
            return DN4GL_EINVAL;
    }
#undef PQ_CASE
    DN_LAUNCHED();
    return DN4GL_OK;
}
