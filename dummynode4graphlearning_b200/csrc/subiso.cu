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

// ---------------------------------------------------------------------------------------------------------------
// get_conjugate_subisomorphisms (subgraph_isomorphism/utils/graph.py:294-330) followed by the g_eid gather of
// convert_to_conjugate (train.py:546-556 / 577-587): every subisomorphism (a map of pattern NODES to graph nodes) becomes
// a map of pattern EDGE slots to graph EDGE ids -- the vertices of the conjugate graphs.  Slot q stands for the q-th
// distinct (u, v) pair of the pattern in order of first appearance; its label set is the LAST run of consecutive pattern
// edges with that pair (dict semantics, as in the edge weights above); the slot's value is the last edge, in
// (src, dst, id) order, between the mapped endpoints whose label is in the set, else -- and for the slots beyond the
// number of distinct pairs -- position 0 of that order (the reference's zero-initialised matrix), mapped to its edge id.
// One thread per (subisomorphism, slot).  Output ids are graph-LOCAL edge ids, laid out like the work items
// (work_ptr[b] + s * m_b + q).
__global__ void subiso_conjugate_kernel(int B, const int32_t *__restrict__ work_ptr, const int32_t *__restrict__ val_ptr,
                                        const int32_t *__restrict__ values, const int32_t *__restrict__ p_node_ptr,
                                        const int32_t *__restrict__ p_edge_ptr, const int32_t *__restrict__ p_src,
                                        const int32_t *__restrict__ p_dst, const int32_t *__restrict__ p_elabel,
                                        const int32_t *__restrict__ active, const int32_t *__restrict__ g_node_ptr,
                                        const int32_t *__restrict__ g_edge_ptr, const int32_t *__restrict__ out_ptr,
                                        const int32_t *__restrict__ out_items, const int32_t *__restrict__ g_dst,
                                        const int32_t *__restrict__ g_elabel, int32_t *__restrict__ conj, int64_t total) {
    const int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= total) return;
    const int b = segment_of(work_ptr, B, t);
    const int pe0 = p_edge_ptr[b], m = p_edge_ptr[b + 1] - pe0;
    const int local = static_cast<int>(t - work_ptr[b]);
    const int s = local / m, q = local % m;
    const int gn0 = g_node_ptr[b], ge0 = g_edge_ptr[b];
    const int first_pos = out_ptr[gn0];                      // position 0 of the graph's (src, dst, id) order
    int result_pos = first_pos;
    // the q-th distinct pair in order of first appearance
    int f = -1, seen = 0;
    for (int e = pe0; e < pe0 + m && f < 0; ++e) {
        bool is_first = true;
        for (int e2 = pe0; e2 < e; ++e2)
            if (p_src[e2] == p_src[e] && p_dst[e2] == p_dst[e]) { is_first = false; break; }
        if (is_first) {
            if (seen == q) f = e;
            ++seen;
        }
    }
    if (f >= 0) {
        const int pn0 = p_node_ptr[b], np_ = p_node_ptr[b + 1] - pn0;
        const int32_t *row = values + val_ptr[b] + static_cast<int64_t>(s) * np_;
        const int pu = p_src[f], pv = p_dst[f];
        const int u = gn0 + row[pu - pn0], v = gn0 + row[pv - pn0];
        for (int p = out_ptr[u]; p < out_ptr[u + 1]; ++p) {
            const int it = out_items[p];
            const int d = g_dst[it];
            if (d > v) break;
            if (d != v) continue;
            const int lab = g_elabel[it];
            for (int e = f; e < pe0 + m; ++e)               // labels of the pair's last run
                if (active[e] && p_src[e] == pu && p_dst[e] == pv && p_elabel[e] == lab) { result_pos = p; break; }
        }
    }
    conj[t] = out_items[result_pos] - ge0;
}

extern "C" int dn4gl_subiso_conjugate(int32_t B, const int32_t *work_ptr, int64_t total_work, const int32_t *val_ptr,
                                      const int32_t *values, const int32_t *p_node_ptr, const int32_t *p_edge_ptr,
                                      const int32_t *p_src, const int32_t *p_dst, const int32_t *p_elabel, int64_t Ep,
                                      int32_t *active_ws, const int32_t *g_node_ptr, const int32_t *g_edge_ptr,
                                      const int32_t *g_out_ptr, const int32_t *g_out_items, const int32_t *g_dst,
                                      const int32_t *g_elabel, int32_t *conj, void *stream) {
    DN_ARG(B >= 0 && total_work >= 0 && Ep >= 0);
    if (total_work == 0) return DN4GL_OK;
    DN_ARG(work_ptr && val_ptr && values && p_node_ptr && p_edge_ptr && p_src && p_dst && p_elabel && active_ws &&
           g_node_ptr && g_edge_ptr && g_out_ptr && g_out_items && g_dst && g_elabel && conj && Ep > 0);
    cudaStream_t st = as_stream(stream);
    pattern_active_edges_kernel<<<static_cast<unsigned>(ceil_div64(Ep, 256)), 256, 0, st>>>(B, p_edge_ptr, p_src, p_dst,
                                                                                          active_ws, Ep);
    subiso_conjugate_kernel<<<static_cast<unsigned>(ceil_div64(total_work, 256)), 256, 0, st>>>(
        B, work_ptr, val_ptr, values, p_node_ptr, p_edge_ptr, p_src, p_dst, p_elabel, active_ws, g_node_ptr, g_edge_ptr,
        g_out_ptr, g_out_items, g_dst, g_elabel, conj, total_work);
    DN_LAUNCHED_N(2);
    return DN4GL_OK;
}
