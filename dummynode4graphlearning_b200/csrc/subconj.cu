// libdn4gl.so -- a5: edge-to-vertex ("conjugate") transform, subgraph-isomorphism flavour.
//
// Replaces convert_conjugate_graph, DGL branch (subgraph_isomorphism/utils/graph.py:77-175; the igraph branch :177-267
// implements the same rule), called per graph from convert_to_conjugate (train.py:564-593), on a whole block-diagonal
// batch at once.  Reference rule, per graph:
//   1. vertices: edges with EQUAL edata["id"] merge into one conjugate vertex (:86-101); it inherits the attributes of
//      the merged edge with the smallest position (id2vertex[id] = min e); vertices are numbered by ascending id
//      (ids that do not occur are removed afterwards, :167-170, order-preserving);
//   2. edges: for e in edge order and e' ascending over the in-edges of src(e), the candidate (id[e'] -> id[e]) is
//      kept iff the key (id[e'], label[src(e)], id[e]) was not seen before (:116-131, python set); the conjugate edge
//      inherits the attributes of the shared vertex src(e).
// After dummy augmentation (train.py:404-435) the n u->d edges share one id and the n d->u edges another, so the n^2
// (in(d) x out(d)) candidates collapse to a single edge: "first occurrence" is the whole difficulty.
//
// GPU formulation (bit-exact, order included):
//   count : first_e[g, id] = min e (atomicMin), presence flags scanned -> global conjugate vertex number of every
//           (g, id); ev[e] = vertex of edge e; cand_off = scan of indeg(src(e)).
//   mark  : an open-addressing hash table keyed by (ev[e'], label[src e], ev[e]) whose slot VALUE is the candidate
//           itself, packed (e << 32 | e') -- exactly the reference's candidate order -- combined with atomicMin, so each
//           key ends with its first occurrence whatever the thread schedule (deterministic).  Lanes of a warp holding
//           the same key are pre-aggregated with __match_any_sync (the dummy node alone produces n^2 equal keys).
//           A candidate survives iff the table holds itself; survivors are scanned in candidate order.
//   fill  : compacted write of (src, dst, shared vertex) and of the vertex origins.
#include "common.cuh"

constexpr unsigned long long SC_EMPTY = ~0ull;

struct SubConjWs {
    int32_t *first_e;   // [B * id_bound]   smallest edge position carrying (g, id), INT_MAX if absent
    int32_t *rank;      // [B * id_bound + 1] exclusive scan of presence = global conjugate vertex number
    char *scan_ws;
    size_t scan_bytes;
};

static bool carve_subconj(void *ws, size_t ws_bytes, int64_t slots, int64_t scan_n, SubConjWs *c) {
    WsCarver w(ws, ws_bytes);
    c->first_e = w.take<int32_t>(slots + 1);
    c->rank = w.take<int32_t>(slots + 2);
    c->scan_bytes = dn4gl_scan_workspace_bytes(scan_n);
    c->scan_ws = w.take<char>(c->scan_bytes);
    return c->first_e && c->rank && c->scan_ws;
}

static int64_t subconj_scan_n(int32_t B, int64_t E, int32_t id_bound, int64_t ncand) {
    int64_t n = static_cast<int64_t>(B) * id_bound + 1;
    if (E + 1 > n) n = E + 1;
    if (ncand + 1 > n) n = ncand + 1;
    return n;
}

extern "C" size_t dn4gl_sub_conj_workspace_bytes(int32_t B, int64_t E, int32_t id_bound) {
    auto a = [](size_t n) { return align_up(n * sizeof(int32_t), 256); };
    const int64_t slots = static_cast<int64_t>(B) * id_bound;
    return a(slots + 1) + a(slots + 2) + dn4gl_scan_workspace_bytes(subconj_scan_n(B, E, id_bound, 0)) + 256;
}

__global__ void sc_fill(int32_t *p, int64_t n, int32_t v) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void sc_first_edge(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ eid, int64_t E,
                              int id_bound, int32_t *__restrict__ first_e, int32_t *__restrict__ err_flag) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int id = eid[e];
    if (id < 0 || id >= id_bound) {
        if (err_flag) atomicExch(err_flag, DN4GL_ELIMIT);
        return;
    }
    const int g = segment_of(edge_ptr, B, e);
    atomicMin(first_e + static_cast<int64_t>(g) * id_bound + id, static_cast<int32_t>(e));
}

__global__ void sc_present(const int32_t *__restrict__ first_e, int64_t n, int32_t *__restrict__ flag) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = first_e[i] != INT32_MAX;
}

__global__ void sc_edge_vertex(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ src,
                               const int32_t *__restrict__ eid, int64_t E, int id_bound, const int32_t *__restrict__ rank,
                               const int32_t *__restrict__ in_ptr, int32_t *__restrict__ ev, int32_t *__restrict__ cand_cnt) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int g = segment_of(edge_ptr, B, e);
    int id = eid[e];
    id = id < 0 ? 0 : (id >= id_bound ? id_bound - 1 : id);   // out-of-range ids were reported through err_flag
    ev[e] = rank[static_cast<int64_t>(g) * id_bound + id];
    const int u = src[e];
    cand_cnt[e] = in_ptr[u + 1] - in_ptr[u];
}

__global__ void sc_node_ptr(int B, int id_bound, const int32_t *__restrict__ rank, int32_t *__restrict__ o_node_ptr) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= B) o_node_ptr[g] = rank[static_cast<int64_t>(g) * id_bound];
}

extern "C" int dn4gl_sub_conj_count(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *eid,
                                    int64_t N, int64_t E, int32_t id_bound, const int32_t *in_ptr, int32_t *ev,
                                    int32_t *cand_off, int32_t *o_node_ptr, void *ws, size_t ws_bytes,
                                    int32_t *err_flag, void *stream) {
    DN_ARG(B >= 0 && N >= 0 && E >= 0 && id_bound >= 1 && edge_ptr && cand_off && o_node_ptr);
    DN_ARG(E == 0 || (src && eid && in_ptr && ev));
    DN_ARG(static_cast<int64_t>(B) * id_bound < (1ll << 31));
    cudaStream_t st = as_stream(stream);
    const int64_t slots = static_cast<int64_t>(B) * id_bound;
    SubConjWs c;
    if (!carve_subconj(ws, ws_bytes, slots, subconj_scan_n(B, E, id_bound, 0), &c)) {
        dn4gl_set_error("dn4gl_sub_conj_count: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    int launched = 0;
    if (slots > 0) {
        sc_fill<<<static_cast<unsigned>(ceil_div64(slots, 256)), 256, 0, st>>>(c.first_e, slots, INT32_MAX);
        ++launched;
    }
    if (E > 0) {
        sc_first_edge<<<static_cast<unsigned>(ceil_div64(E, 256)), 256, 0, st>>>(B, edge_ptr, eid, E, id_bound, c.first_e,
                                                                                 err_flag);
        ++launched;
    }
    if (slots > 0) {
        sc_present<<<static_cast<unsigned>(ceil_div64(slots, 256)), 256, 0, st>>>(c.first_e, slots, c.rank);
        ++launched;
    }
    if (launched) DN_LAUNCHED_N(launched);
    int rc = dn4gl_exclusive_scan_i32(c.rank, c.rank, slots, c.scan_ws, c.scan_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    sc_node_ptr<<<(B + 1 + 255) / 256, 256, 0, st>>>(B, id_bound, c.rank, o_node_ptr);
    DN_LAUNCHED();
    if (E > 0) {
        sc_edge_vertex<<<static_cast<unsigned>(ceil_div64(E, 256)), 256, 0, st>>>(B, edge_ptr, src, eid, E, id_bound, c.rank,
                                                                                  in_ptr, ev, cand_off);
        DN_LAUNCHED();
    }
    return dn4gl_exclusive_scan_i32(cand_off, cand_off, E, c.scan_ws, c.scan_bytes, stream);
}

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long sc_hash(int a, int b, int c) {
    unsigned long long x = (static_cast<unsigned long long>(static_cast<unsigned>(a)) << 32) | static_cast<unsigned>(b);
    x ^= static_cast<unsigned long long>(static_cast<unsigned>(c)) * 0x9E3779B97F4A7C15ull;
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}

struct ScKeyCtx {
    const int32_t *__restrict__ src;
    const int32_t *__restrict__ vlabel;
    const int32_t *__restrict__ ev;
};
__device__ __forceinline__ bool sc_same_key(const ScKeyCtx &k, unsigned long long val, int ku, int kl, int kv) {
    const int e = static_cast<int>(val >> 32), ep = static_cast<int>(val & 0xffffffffu);
    return k.ev[ep] == ku && k.ev[e] == kv && k.vlabel[k.src[e]] == kl;
}

// one warp per edge e; lanes stride the in-list of src(e).  INSERT = true: atomicMin the candidate into its key's slot;
// INSERT = false: keep[c] = (slot value == this candidate).
template <bool INSERT>
__global__ void __launch_bounds__(256)
sc_candidates(const int32_t *__restrict__ src, const int32_t *__restrict__ vlabel, int64_t E,
              const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid, const int32_t *__restrict__ ev,
              const int32_t *__restrict__ cand_off, unsigned long long *__restrict__ table, unsigned long long mask,
              int32_t *__restrict__ keep) {
    const int lane = threadIdx.x & 31;
    const ScKeyCtx kc{src, vlabel, ev};
    for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; e < E;
         e += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const int u = src[e];
        const int kl = vlabel[u], kv = ev[e];
        const int beg = in_ptr[u], end = in_ptr[u + 1];
        const int c0 = cand_off[e];
        for (int p0 = beg; p0 < end; p0 += 32) {
            const int p = p0 + lane;
            const bool act = p < end;
            int ep = 0, ku = -1;
            if (act) { ep = in_eid[p]; ku = ev[ep]; }
            const unsigned long long val = (static_cast<unsigned long long>(static_cast<unsigned>(e)) << 32) |
                                           static_cast<unsigned>(ep);
            const unsigned active = __ballot_sync(0xffffffffu, act);
            if (!act) continue;
            // lanes with the same key (same ku: kl and kv are warp-uniform); the lowest lane holds the smallest e'
            const unsigned peers = __match_any_sync(active, ku);
            const bool leader = (__ffs(peers) - 1) == lane;
            if (INSERT) {
                if (!leader) continue;
                unsigned long long h = sc_hash(ku, kl, kv) & mask;
                while (true) {
                    unsigned long long old = atomicCAS(table + h, SC_EMPTY, val);
                    if (old == SC_EMPTY) break;
                    if (sc_same_key(kc, old, ku, kl, kv)) { atomicMin(table + h, val); break; }
                    h = (h + 1) & mask;
                }
            } else {
                bool k = false;
                if (leader) {   // only the first lane of a group can be the first occurrence
                    unsigned long long h = sc_hash(ku, kl, kv) & mask;
                    while (true) {
                        const unsigned long long cur = table[h];
                        if (cur == val) { k = true; break; }
                        if (cur == SC_EMPTY || sc_same_key(kc, cur, ku, kl, kv)) break;
                        h = (h + 1) & mask;
                    }
                }
                keep[c0 + (p - beg)] = k ? 1 : 0;
            }
        }
    }
}

__global__ void sc_edge_ptr(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ cand_off,
                            const int32_t *__restrict__ keep_scan, int64_t E, int32_t *__restrict__ o_edge_ptr) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g > B) return;
    const int e = edge_ptr[g];
    o_edge_ptr[g] = keep_scan[cand_off[e]];   // cand_off has E+1 entries, keep_scan ncand+1
}

extern "C" int dn4gl_sub_conj_mark(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *vlabel,
                                   int64_t E, const int32_t *in_ptr, const int32_t *in_eid, const int32_t *ev,
                                   const int32_t *cand_off, int64_t ncand, void *table, int64_t table_slots,
                                   int32_t *keep_scan, int32_t *o_edge_ptr, void *ws, size_t ws_bytes, void *stream) {
    DN_ARG(B >= 0 && E >= 0 && ncand >= 0 && ncand < (1ll << 31) && edge_ptr && cand_off && keep_scan && o_edge_ptr);
    DN_ARG(table_slots >= 2 && (table_slots & (table_slots - 1)) == 0 && table_slots >= 2 * ncand && table);
    DN_ARG(E == 0 || (src && vlabel && in_ptr && in_eid && ev));
    cudaStream_t st = as_stream(stream);
    const size_t scan_bytes = dn4gl_scan_workspace_bytes(ncand + 1);
    if (ws == nullptr || ws_bytes < scan_bytes) {
        dn4gl_set_error("dn4gl_sub_conj_mark: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    DN_CUDA(cudaMemsetAsync(table, 0xff, static_cast<size_t>(table_slots) * sizeof(unsigned long long), st));
    if (E > 0 && ncand > 0) {
        const int64_t warps = E < 148 * 64 ? E : 148 * 64;
        const unsigned grid = static_cast<unsigned>(ceil_div64(warps * 32, 256));
        const unsigned long long mask = static_cast<unsigned long long>(table_slots - 1);
        sc_candidates<true><<<grid, 256, 0, st>>>(src, vlabel, E, in_ptr, in_eid, ev, cand_off,
                                                  static_cast<unsigned long long *>(table), mask, nullptr);
        sc_candidates<false><<<grid, 256, 0, st>>>(src, vlabel, E, in_ptr, in_eid, ev, cand_off,
                                                   static_cast<unsigned long long *>(table), mask, keep_scan);
        DN_LAUNCHED_N(2);
    }
    // the scan reads keep[0 .. ncand) and writes keep_scan[0 .. ncand]
    int rc = dn4gl_exclusive_scan_i32(keep_scan, keep_scan, ncand, ws, ws_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    sc_edge_ptr<<<(B + 1 + 255) / 256, 256, 0, st>>>(B, edge_ptr, cand_off, keep_scan, E, o_edge_ptr);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sc_fill_edges(const int32_t *__restrict__ src, int64_t E, const int32_t *__restrict__ in_ptr,
              const int32_t *__restrict__ in_eid, const int32_t *__restrict__ ev, const int32_t *__restrict__ cand_off,
              const int32_t *__restrict__ keep_scan, int32_t *__restrict__ o_src, int32_t *__restrict__ o_dst,
              int32_t *__restrict__ o_e_shared) {
    const int lane = threadIdx.x & 31;
    for (int64_t e = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; e < E;
         e += (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5) {
        const int u = src[e], kv = ev[e];
        const int beg = in_ptr[u], end = in_ptr[u + 1];
        const int c0 = cand_off[e];
        for (int p = beg + lane; p < end; p += 32) {
            const int c = c0 + (p - beg);
            const int q = keep_scan[c];
            if (keep_scan[c + 1] != q) {
                o_src[q] = ev[in_eid[p]];
                o_dst[q] = kv;
                o_e_shared[q] = u;
            }
        }
    }
}

__global__ void sc_fill_vertices(const int32_t *__restrict__ first_e, const int32_t *__restrict__ rank, int64_t slots,
                                 int32_t *__restrict__ o_v_origin) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= slots) return;
    const int f = first_e[i];
    if (f != INT32_MAX) o_v_origin[rank[i]] = f;
}

extern "C" int dn4gl_sub_conj_fill(int32_t B, const int32_t *src, int64_t E, int32_t id_bound, const int32_t *in_ptr,
                                   const int32_t *in_eid, const int32_t *ev, const int32_t *cand_off,
                                   const int32_t *keep_scan, int32_t *o_src, int32_t *o_dst, int32_t *o_v_origin,
                                   int32_t *o_e_shared, void *ws, size_t ws_bytes, void *stream) {
    DN_ARG(B >= 0 && E >= 0 && id_bound >= 1);
    cudaStream_t st = as_stream(stream);
    const int64_t slots = static_cast<int64_t>(B) * id_bound;
    SubConjWs c;
    if (!carve_subconj(ws, ws_bytes, slots, 1, &c)) {   // same carve as the count phase: first_e / rank are still there
        dn4gl_set_error("dn4gl_sub_conj_fill: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    int launched = 0;
    if (slots > 0 && o_v_origin) {
        sc_fill_vertices<<<static_cast<unsigned>(ceil_div64(slots, 256)), 256, 0, st>>>(c.first_e, c.rank, slots, o_v_origin);
        ++launched;
    }
    if (E > 0 && o_src) {
        DN_ARG(src && in_ptr && in_eid && ev && cand_off && keep_scan && o_dst && o_e_shared);
        const int64_t warps = E < 148 * 64 ? E : 148 * 64;
        sc_fill_edges<<<static_cast<unsigned>(ceil_div64(warps * 32, 256)), 256, 0, st>>>(
            src, E, in_ptr, in_eid, ev, cand_off, keep_scan, o_src, o_dst, o_e_shared);
        ++launched;
    }
    if (launched) DN_LAUNCHED_N(launched);
    return DN4GL_OK;
}
