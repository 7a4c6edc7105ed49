// libdn4gl.so -- match-weight targets from ground-truth subisomorphisms: the step before the
// hot path when --match_weights node / edge is on.  Replaces the numba loops compute_nodeseq_subisoweights /
// compute_edgeseq_subisoweights (subgraph_isomorphism/dataset.py:54-108) as called per sample by
// GraphAdjDataset.calculate_node_weights / calculate_edge_weights (dataset.py:1491-1520), batched over a mini-batch.
//
// Layout: sample b owns S_b subisomorphisms, each a row of np_b graph-LOCAL node ids; all rows are concatenated in
// `values` (int32) with val_ptr[b] = first element of sample b, and np_b = pattern nodes of sample b.
// Integer atomics only (commutative and exact), so the result is deterministic.
#include "common.cuh"

// node weights: histogram of every subisomorphism entry (dataset.py:55-61)
__global__ void subiso_node_weights_kernel(int B, const int32_t *__restrict__ val_ptr, const int32_t *__restrict__ values,
                                           const int32_t *__restrict__ g_node_ptr, int32_t *__restrict__ w, int64_t total) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = segment_of(val_ptr, B, i);
    atomicAdd(w + g_node_ptr[b] + values[i], 1);
}

// the reference groups CONSECUTIVE pattern edges with equal (u, v) into dict entries keyed by (u, v)
// (dataset.py:79-90): a later run with the same key overwrites an earlier one, so an edge only counts if its run is the
// last one carrying its key.  active[e] = 1 for those edges.
__global__ void pattern_active_edges_kernel(int B, const int32_t *__restrict__ p_edge_ptr, const int32_t *__restrict__ p_src,
                                            const int32_t *__restrict__ p_dst, int32_t *__restrict__ active, int64_t Ep) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= Ep) return;
    const int b = segment_of(p_edge_ptr, B, e);
    const int e1 = p_edge_ptr[b + 1];
    const int u = p_src[e], v = p_dst[e];
    int j = static_cast<int>(e) + 1;
    while (j < e1 && p_src[j] == u && p_dst[j] == v) ++j;     // end of this edge's run
    int ok = 1;
    for (; j < e1; ++j)
        if (p_src[j] == u && p_dst[j] == v) { ok = 0; break; }
    active[e] = ok;
}

// edge weights (dataset.py:92-107): one thread per (subisomorphism, pattern edge).  The graph's out-lists are sorted by
// (dst, edge id) -- the reference's all_edges(order="srcdst") -- so the edges u' -> v' form one run inside row u'.
__global__ void subiso_edge_weights_kernel(int B, const int32_t *__restrict__ work_ptr, const int32_t *__restrict__ val_ptr,
                                           const int32_t *__restrict__ values, const int32_t *__restrict__ p_node_ptr,
                                           const int32_t *__restrict__ p_edge_ptr, const int32_t *__restrict__ p_src,
                                           const int32_t *__restrict__ p_dst, const int32_t *__restrict__ p_elabel,
                                           const int32_t *__restrict__ active, const int32_t *__restrict__ g_node_ptr,
                                           const int32_t *__restrict__ out_ptr, const int32_t *__restrict__ out_items,
                                           const int32_t *__restrict__ g_dst, const int32_t *__restrict__ g_elabel,
                                           int32_t *__restrict__ w, int64_t total) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = segment_of(work_ptr, B, t);
    const int pe0 = p_edge_ptr[b], m = p_edge_ptr[b + 1] - pe0;
    const int local = static_cast<int>(t - work_ptr[b]);
    const int s = local / m, e = pe0 + local % m;
    if (!active[e]) return;
    const int pn0 = p_node_ptr[b], np_ = p_node_ptr[b + 1] - pn0;
    const int32_t *row = values + val_ptr[b] + static_cast<int64_t>(s) * np_;
    const int gn0 = g_node_ptr[b];
    const int u = gn0 + row[p_src[e] - pn0], v = gn0 + row[p_dst[e] - pn0];
    const int lab = p_elabel[e];
    for (int p = out_ptr[u]; p < out_ptr[u + 1]; ++p) {
        const int it = out_items[p];
        const int d = g_dst[it];
        if (d > v) break;
        if (d == v && g_elabel[it] == lab) atomicAdd(w + it, 1);
    }
}

extern "C" int dn4gl_subiso_node_weights(int32_t B, const int32_t *val_ptr, const int32_t *values, int64_t total,
                                         const int32_t *g_node_ptr, int64_t Ng, int32_t *weights, void *stream) {
    DN_ARG(B >= 0 && total >= 0 && Ng >= 0 && (Ng == 0 || weights != nullptr));
    cudaStream_t st = as_stream(stream);
    if (Ng > 0) DN_CUDA(cudaMemsetAsync(weights, 0, static_cast<size_t>(Ng) * sizeof(int32_t), st));
    if (total == 0) return DN4GL_OK;
    DN_ARG(val_ptr && values && g_node_ptr);
    subiso_node_weights_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, st>>>(B, val_ptr, values, g_node_ptr,
                                                                                            weights, total);
    DN_LAUNCHED();
    return DN4GL_OK;
}

extern "C" int dn4gl_subiso_edge_weights(int32_t B, const int32_t *work_ptr, int64_t total_work, const int32_t *val_ptr,
                                         const int32_t *values, const int32_t *p_node_ptr, const int32_t *p_edge_ptr,
                                         const int32_t *p_src, const int32_t *p_dst, const int32_t *p_elabel, int64_t Ep,
                                         int32_t *active_ws, const int32_t *g_node_ptr, const int32_t *g_out_ptr,
                                         const int32_t *g_out_items, const int32_t *g_dst, const int32_t *g_elabel,
                                         int64_t Eg, int32_t *weights, void *stream) {
    DN_ARG(B >= 0 && total_work >= 0 && Ep >= 0 && Eg >= 0 && (Eg == 0 || weights != nullptr));
    cudaStream_t st = as_stream(stream);
    if (Eg > 0) DN_CUDA(cudaMemsetAsync(weights, 0, static_cast<size_t>(Eg) * sizeof(int32_t), st));
    if (total_work == 0 || Ep == 0 || Eg == 0) return DN4GL_OK;
    DN_ARG(work_ptr && val_ptr && values && p_node_ptr && p_edge_ptr && p_src && p_dst && p_elabel && active_ws &&
           g_node_ptr && g_out_ptr && g_out_items && g_dst && g_elabel);
    pattern_active_edges_kernel<<<static_cast<unsigned>(ceil_div64(Ep, 256)), 256, 0, st>>>(B, p_edge_ptr, p_src, p_dst,
                                                                                          active_ws, Ep);
    subiso_edge_weights_kernel<<<static_cast<unsigned>(ceil_div64(total_work, 256)), 256, 0, st>>>(
        B, work_ptr, val_ptr, values, p_node_ptr, p_edge_ptr, p_src, p_dst, p_elabel, active_ws, g_node_ptr, g_out_ptr,
        g_out_items, g_dst, g_elabel, weights, total_work);
    DN_LAUNCHED_N(2);
    return DN4GL_OK;
}
