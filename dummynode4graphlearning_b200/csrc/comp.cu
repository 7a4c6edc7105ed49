// libdn4gl.so -- CompGCN composition (subgraph_isomorphism/models/compgcn.py:214-240), SURVEY.md 8(f) rank 1.
//
// The reference's message is  norm_e * comp(h[src e], ef[e]) @ (rev_e ? W_out : W_in)  summed over the in-edges of a node.
// By linearity the weights move out of the sum (one node-level GEMM on [S_rev | S_fwd], like K4), and every edge
// normalisation of compgcn.py:190-209 factorises over the endpoints: norm_e = a[src e] * b[dst e]
// ("in": a = 1, b = innorm; "out": a = outnorm, b = 1; "both": a = sqrt(outnorm), b = sqrt(innorm)).  What is left per
// edge is the composition itself:
//     C[e] = a[src e] * comp(h[src e], ef[e]),   comp = h - r ("sub") | h * r ("mult")
// one streaming pass over the edge features with an L2-resident gather of the source rows; the segment sum of C is
// dn4gl_dmp_node_agg_f32.  Backward: gEF edge-parallel, gH as a deterministic per-node sum over the out-list (CSR by
// source, items ascending by edge id) -- no float atomics.  ("corr", the circular correlation, stays on torch's FFT.)
// HBM-bound: forward bytes 4D(2E) + 4E + gathers of h (N rows, L2); backward 4D(3E) + 4D N.
#include "common.cuh"

enum { COMP_SUB = 0, COMP_MULT = 1 };

__device__ __forceinline__ float4 mul4(const float4 &a, const float4 &b) {
    return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
__device__ __forceinline__ float4 scale4(const float4 &a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }

// one thread per (edge, float4 column)
template <int OP>
__global__ void __launch_bounds__(256) comp_edge_kernel(const int32_t *__restrict__ src, const float *__restrict__ a,
                                                        const float4 *__restrict__ h, const float4 *__restrict__ ef,
                                                        float4 *__restrict__ C, int64_t total, int DV) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t e = i / DV;
    const int c = static_cast<int>(i - e * DV);
    const int u = __ldg(src + e);
    const float4 hv = ldg4(h + static_cast<int64_t>(u) * DV + c), r = ldg4(ef + i);
    float4 v;
    if (OP == COMP_SUB) { v = hv; sub4(v, r); } else { v = mul4(hv, r); }
    if (a != nullptr) v = scale4(v, __ldg(a + u));
    C[i] = v;
}

// gEF[e] = a[src] * (-gC[e])  (sub)   |   a[src] * gC[e] * h[src]  (mult)
template <int OP>
__global__ void __launch_bounds__(256) comp_edge_bwd_ef_kernel(const int32_t *__restrict__ src, const float *__restrict__ a,
                                                               const float4 *__restrict__ h, const float4 *__restrict__ gC,
                                                               float4 *__restrict__ gEF, int64_t total, int DV) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int64_t e = i / DV;
    const int c = static_cast<int>(i - e * DV);
    const int u = __ldg(src + e);
    float4 g = ldg4(gC + i);
    if (a != nullptr) g = scale4(g, __ldg(a + u));
    if (OP == COMP_SUB) g = make_float4(-g.x, -g.y, -g.z, -g.w);
    else g = mul4(g, ldg4(h + static_cast<int64_t>(u) * DV + c));
    gEF[i] = g;
}

// gH[u] = a[u] * sum_{e in out(u)} gC[e]  (sub)   |   a[u] * sum gC[e] * ef[e]  (mult); sub-group of LANES lanes per node,
// out-list walked in CSR (= edge id) order with 4 edges in flight
template <int OP, int LANES, int VEC>
__global__ void __launch_bounds__(256) comp_edge_bwd_h_kernel(const int32_t *__restrict__ out_ptr, const int32_t *__restrict__ out_eid,
                                                              const float *__restrict__ a, const float4 *__restrict__ ef,
                                                              const float4 *__restrict__ gC, float4 *__restrict__ gH, int64_t N) {
    constexpr int ROWS = 256 / LANES;
    constexpr int DV = LANES * VEC;
    constexpr int U = (VEC == 1) ? 4 : 2;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * ROWS + threadIdx.x / LANES;
    const int lane = threadIdx.x % LANES;
    if (row >= N) return;
    const int beg = __ldg(out_ptr + row), end = __ldg(out_ptr + row + 1);
    float4 acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = zero4();
    for (int p = beg; p < end; p += U) {
        int e[U];
        float4 g[U][VEC], r[U][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = (p + u < end) ? __ldg(out_eid + p + u) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (e[u] >= 0) {
                    g[u][k] = ldg4(gC + static_cast<int64_t>(e[u]) * DV + lane + k * LANES);
                    if (OP == COMP_MULT) r[u][k] = ldg4(ef + static_cast<int64_t>(e[u]) * DV + lane + k * LANES);
                }
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < VEC; ++k)
                if (e[u] >= 0) add4(acc[k], OP == COMP_MULT ? mul4(g[u][k], r[u][k]) : g[u][k]);
    }
    const float s = (a != nullptr) ? __ldg(a + row) : 1.f;
#pragma unroll
    for (int k = 0; k < VEC; ++k) gH[row * DV + lane + k * LANES] = scale4(acc[k], s);
}

static bool comp_args_ok(int64_t n, int32_t D, int32_t op) { return n >= 0 && D > 0 && D % 4 == 0 && (op == COMP_SUB || op == COMP_MULT); }

extern "C" int dn4gl_comp_edge_f32(const int32_t *src, const float *src_scale, const float *h, const float *ef, float *C,
                                   int64_t E, int32_t D, int32_t op, void *stream) {
    DN_ARG(comp_args_ok(E, D, op));
    if (E == 0) return DN4GL_OK;
    DN_ARG(src && h && ef && C && aligned16(h) && aligned16(ef) && aligned16(C));
    const int DV = D / 4;
    const int64_t total = E * DV;
    const unsigned grid = static_cast<unsigned>(ceil_div64(total, 256));
    auto H = reinterpret_cast<const float4 *>(h);
    auto R = reinterpret_cast<const float4 *>(ef);
    auto O = reinterpret_cast<float4 *>(C);
    if (op == COMP_SUB) comp_edge_kernel<COMP_SUB><<<grid, 256, 0, as_stream(stream)>>>(src, src_scale, H, R, O, total, DV);
    else comp_edge_kernel<COMP_MULT><<<grid, 256, 0, as_stream(stream)>>>(src, src_scale, H, R, O, total, DV);
    DN_LAUNCHED();
    return DN4GL_OK;
}

extern "C" int dn4gl_comp_edge_bwd_ef_f32(const int32_t *src, const float *src_scale, const float *h, const float *gC,
                                          float *gEF, int64_t E, int32_t D, int32_t op, void *stream) {
    DN_ARG(comp_args_ok(E, D, op));
    if (E == 0) return DN4GL_OK;
    DN_ARG(src && gC && gEF && aligned16(gC) && aligned16(gEF) && (op == COMP_SUB || (h && aligned16(h))));
    const int DV = D / 4;
    const int64_t total = E * DV;
    const unsigned grid = static_cast<unsigned>(ceil_div64(total, 256));
    auto H = reinterpret_cast<const float4 *>(h);
    auto G = reinterpret_cast<const float4 *>(gC);
    auto O = reinterpret_cast<float4 *>(gEF);
    if (op == COMP_SUB) comp_edge_bwd_ef_kernel<COMP_SUB><<<grid, 256, 0, as_stream(stream)>>>(src, src_scale, H, G, O, total, DV);
    else comp_edge_bwd_ef_kernel<COMP_MULT><<<grid, 256, 0, as_stream(stream)>>>(src, src_scale, H, G, O, total, DV);
    DN_LAUNCHED();
    return DN4GL_OK;
}

extern "C" int dn4gl_comp_edge_bwd_h_f32(const int32_t *out_ptr, const int32_t *out_eid, const float *src_scale,
                                         const float *ef, const float *gC, float *gH, int64_t N, int32_t D, int32_t op,
                                         void *stream) {
    DN_ARG(comp_args_ok(N, D, op));
    if (N == 0) return DN4GL_OK;
    DN_ARG(out_ptr && out_eid && gC && gH && aligned16(gC) && aligned16(gH) && (op == COMP_SUB || (ef && aligned16(ef))));
    cudaStream_t st = as_stream(stream);
    auto R = reinterpret_cast<const float4 *>(ef);
    auto G = reinterpret_cast<const float4 *>(gC);
    auto O = reinterpret_cast<float4 *>(gH);
#define COMP_CASE(L, V)                                                                                              \
    if (op == COMP_SUB)                                                                                              \
        comp_edge_bwd_h_kernel<COMP_SUB, L, V><<<static_cast<unsigned>(ceil_div64(N, 256 / L)), 256, 0, st>>>(       \
            out_ptr, out_eid, src_scale, R, G, O, N);                                                                \
    else                                                                                                             \
        comp_edge_bwd_h_kernel<COMP_MULT, L, V><<<static_cast<unsigned>(ceil_div64(N, 256 / L)), 256, 0, st>>>(      \
            out_ptr, out_eid, src_scale, R, G, O, N);                                                                \
    break
    switch (D / 4) {
        case 1: COMP_CASE(1, 1);
        case 2: COMP_CASE(2, 1);
        case 4: COMP_CASE(4, 1);
        case 8: COMP_CASE(8, 1);
        case 16: COMP_CASE(16, 1);
        case 32: COMP_CASE(32, 1);
        case 64: COMP_CASE(32, 2);
        default:
            dn4gl_set_error("dn4gl_comp_edge_bwd_h_f32: unsupported D=%d (supported: 4,8,16,32,64,128,256)", D);
            return DN4GL_EINVAL;
    }
#undef COMP_CASE
    DN_LAUNCHED();
    return DN4GL_OK;
}
