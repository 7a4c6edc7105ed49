// libdn4gl.so -- integer graph transforms: dummy-node augmentation (both reference flavours),
// edge-to-vertex ("conjugate") transform (both flavours), PyG-style coalesce.
// All outputs are bit-exact with the reference's Python (checked against oracle/ + tests/golden/).
#include "common.cuh"

int dn4gl_sort_rows(const int32_t *row_ptr, int64_t N, int32_t *items, const int32_t *primary, int32_t *worklist,
                    int32_t *work_count, int32_t *err_flag, cudaStream_t st, const int32_t *val, int32_t *col,
                    bool work_count_zeroed);

// ===========================================================================================
// a1: tu_data_processing.py:186-214.  One thread per output element; the graph of an element is
// found by binary search in the (closed-form) output offsets  node: node_ptr[g]+g,
// edge: edge_ptr[g]+2*node_ptr[g].
struct TuDummyArgs {
    int B;
    const int32_t *node_ptr, *edge_ptr, *src, *dst, *vlabel, *elabel;
    int64_t N, E;
    int32_t *o_node_ptr, *o_edge_ptr, *o_src, *o_dst, *o_vlabel, *o_vdummy, *o_elabel, *o_edummy;
};

__device__ __forceinline__ int find_graph_by_out_node(const int32_t *__restrict__ node_ptr, int B, int64_t i) {
    int lo = 0, hi = B;  // node_ptr[lo]+lo <= i < node_ptr[hi]+hi
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(node_ptr[mid]) + mid <= i) lo = mid; else hi = mid;
    }
    return lo;
}
__device__ __forceinline__ int find_graph_by_out_edge(const int32_t *__restrict__ node_ptr,
                                                      const int32_t *__restrict__ edge_ptr, int B, int64_t j) {
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(edge_ptr[mid]) + 2ll * node_ptr[mid] <= j) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void tu_add_dummy_kernel(TuDummyArgs a) {
    const int64_t n_out = a.N + a.B, e_out = a.E + 2 * a.N;
    int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t <= a.B) {
        a.o_node_ptr[t] = a.node_ptr[t] + static_cast<int32_t>(t);
        a.o_edge_ptr[t] = a.edge_ptr[t] + 2 * a.node_ptr[t];
    }
    if (t < n_out) {
        int g = find_graph_by_out_node(a.node_ptr, a.B, t);
        int64_t local = t - (a.node_ptr[g] + g);
        int n = a.node_ptr[g + 1] - a.node_ptr[g];
        bool dummy = (local == n);
        a.o_vlabel[t] = dummy ? 0 : a.vlabel[a.node_ptr[g] + local];
        a.o_vdummy[t] = dummy ? 1 : 0;
    }
    if (t < e_out) {
        int g = find_graph_by_out_edge(a.node_ptr, a.edge_ptr, a.B, t);
        int n0 = a.node_ptr[g], n = a.node_ptr[g + 1] - n0;
        int e0 = a.edge_ptr[g], m = a.edge_ptr[g + 1] - e0;
        int64_t local = t - (static_cast<int64_t>(e0) + 2ll * n0);
        int no = n0 + g;  // output node offset of graph g
        if (local < m) {
            a.o_src[t] = a.src[e0 + local] - n0 + no;
            a.o_dst[t] = a.dst[e0 + local] - n0 + no;
            a.o_elabel[t] = a.elabel[e0 + local];
            a.o_edummy[t] = 0;
        } else {
            int k = static_cast<int>(local - m);
            int v = k >> 1;
            bool to_dummy = (k & 1);  // even: (n, v)   odd: (v, n)      line 193
            a.o_src[t] = to_dummy ? no + v : no + n;
            a.o_dst[t] = to_dummy ? no + n : no + v;
            a.o_elabel[t] = 0;
            a.o_edummy[t] = 1;
        }
    }
}

extern "C" int dn4gl_tu_add_dummy(int32_t B, const int32_t *node_ptr, const int32_t *edge_ptr, const int32_t *src,
                                  const int32_t *dst, const int32_t *vlabel, const int32_t *elabel, int64_t N,
                                  int64_t E, int32_t *o_node_ptr, int32_t *o_edge_ptr, int32_t *o_src,
                                  int32_t *o_dst, int32_t *o_vlabel, int32_t *o_vdummy, int32_t *o_elabel,
                                  int32_t *o_edummy, void *stream) {
    DN_ARG(B >= 0 && N >= 0 && E >= 0 && E + 2 * N < INT32_MAX);
    DN_ARG(node_ptr && edge_ptr && o_node_ptr && o_edge_ptr);
    DN_ARG(N == 0 || (vlabel && o_vlabel && o_vdummy && o_src && o_dst && o_elabel && o_edummy));
    DN_ARG(E == 0 || (src && dst && elabel));
    TuDummyArgs a{B, node_ptr, edge_ptr, src, dst, vlabel, elabel, N, E, o_node_ptr, o_edge_ptr,
                  o_src, o_dst, o_vlabel, o_vdummy, o_elabel, o_edummy};
    int64_t total = E + 2 * N;
    if (N + B > total) total = N + B;
    if (B + 1 > total) total = B + 1;
    tu_add_dummy_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, as_stream(stream)>>>(a);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// ===========================================================================================
// a4: train.py:404-474 (GraphAdj branch).  Blocked dummy edges [u->d]*n then [d->u]*n.
struct SubDummyArgs {
    int B;
    const int32_t *node_ptr, *edge_ptr, *src, *dst, *vid, *vlabel, *eid, *elabel, *e_isrev;
    int64_t N, E;
    int max_nv, max_nvl, max_ne, max_nel;
    int32_t *o_node_ptr, *o_edge_ptr, *o_src, *o_dst, *o_vid, *o_vlabel, *o_vdummy, *o_eid, *o_elabel, *o_edummy,
        *o_erev;
};

__global__ void sub_add_dummy_kernel(SubDummyArgs a) {
    const int64_t n_out = a.N + a.B, e_out = a.E + 2 * a.N;
    int64_t t = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t <= a.B) {
        a.o_node_ptr[t] = a.node_ptr[t] + static_cast<int32_t>(t);
        a.o_edge_ptr[t] = a.edge_ptr[t] + 2 * a.node_ptr[t];
    }
    if (t < n_out) {
        int g = find_graph_by_out_node(a.node_ptr, a.B, t);
        int64_t local = t - (a.node_ptr[g] + g);
        int n = a.node_ptr[g + 1] - a.node_ptr[g];
        bool dummy = (local == n);
        int64_t v = a.node_ptr[g] + local;
        a.o_vid[t] = dummy ? a.max_nv : a.vid[v];           // train.py:419
        a.o_vlabel[t] = dummy ? a.max_nvl : a.vlabel[v];    // :420
        a.o_vdummy[t] = dummy ? 1 : 0;                      // :421 (old rows zero-filled by DGL)
    }
    if (t < e_out) {
        int g = find_graph_by_out_edge(a.node_ptr, a.edge_ptr, a.B, t);
        int n0 = a.node_ptr[g], n = a.node_ptr[g + 1] - n0;
        int e0 = a.edge_ptr[g], m = a.edge_ptr[g + 1] - e0;
        int64_t local = t - (static_cast<int64_t>(e0) + 2ll * n0);
        int no = n0 + g;
        if (local < m) {
            int64_t e = e0 + local;
            a.o_src[t] = a.src[e] - n0 + no;
            a.o_dst[t] = a.dst[e] - n0 + no;
            a.o_eid[t] = a.eid[e];
            a.o_elabel[t] = a.elabel[e];
            a.o_edummy[t] = 0;
            a.o_erev[t] = a.e_isrev ? a.e_isrev[e] : 0;
        } else {
            int k = static_cast<int>(local - m);
            bool from_dummy = (k >= n);  // first n: u -> d, last n: d -> u      :409-426
            int v = from_dummy ? k - n : k;
            a.o_src[t] = from_dummy ? no + n : no + v;
            a.o_dst[t] = from_dummy ? no + v : no + n;
            a.o_eid[t] = a.max_ne + (from_dummy ? 1 : 0);       // :412-413
            a.o_elabel[t] = a.max_nel + (from_dummy ? 1 : 0);   // :414-415
            a.o_edummy[t] = 1;
            a.o_erev[t] = from_dummy ? 1 : 0;                   // :431
        }
    }
}

extern "C" int dn4gl_sub_add_dummy(int32_t B, const int32_t *node_ptr, const int32_t *edge_ptr, const int32_t *src,
                                   const int32_t *dst, const int32_t *vid, const int32_t *vlabel, const int32_t *eid,
                                   const int32_t *elabel, const int32_t *e_isrev, int64_t N, int64_t E,
                                   int32_t max_nv, int32_t max_nvl, int32_t max_ne, int32_t max_nel,
                                   int32_t *o_node_ptr, int32_t *o_edge_ptr, int32_t *o_src, int32_t *o_dst,
                                   int32_t *o_vid, int32_t *o_vlabel, int32_t *o_vdummy, int32_t *o_eid,
                                   int32_t *o_elabel, int32_t *o_edummy, int32_t *o_erev, void *stream) {
    DN_ARG(B >= 0 && N >= 0 && E >= 0 && E + 2 * N < INT32_MAX);
    DN_ARG(node_ptr && edge_ptr && o_node_ptr && o_edge_ptr);
    DN_ARG(N == 0 || (vid && vlabel && o_vid && o_vlabel && o_vdummy && o_src && o_dst && o_eid && o_elabel &&
                      o_edummy && o_erev));
    DN_ARG(E == 0 || (src && dst && eid && elabel));
    SubDummyArgs a{B, node_ptr, edge_ptr, src, dst, vid, vlabel, eid, elabel, e_isrev, N, E,
                   max_nv, max_nvl, max_ne, max_nel, o_node_ptr, o_edge_ptr, o_src, o_dst,
                   o_vid, o_vlabel, o_vdummy, o_eid, o_elabel, o_edummy, o_erev};
    int64_t total = E + 2 * N;
    if (N + B > total) total = N + B;
    if (B + 1 > total) total = B + 1;
    sub_add_dummy_kernel<<<static_cast<unsigned>(ceil_div64(total, 256)), 256, 0, as_stream(stream)>>>(a);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// ===========================================================================================
// a2: tu_data_processing.py:223-338, closed-form survivor rule (DESIGN.md "conjugate transform"):
// candidates are (e' -> e) for e in edge order, e' ascending over the in-edges of s = src(e).
// With unique edge IDs the (uid,label,vid) key set (269-273) never fires; after all IS_DUMMY
// edges collapse onto the graph's first dummy edge D (302-312) a candidate survives iff
//   e', e real                      -> always
//   e' dummy, e real   (D -> e)     -> e' is the FIRST dummy in-edge of s
//   e' real,  e dummy  (e' -> D)    -> e  is the FIRST dummy out-edge of s
//   both dummy         (D -> D)     -> never (306)
// which reproduces "keep the first occurrence" (313-317) without a hash set.
struct ConjWs {
    int32_t *egraph;         // [E] graph of each edge
    int32_t *first_dummy;    // [B] smallest dummy edge id per graph (INT_MAX if none)
    int32_t *fd_in, *fd_out; // [N] first dummy in-/out-edge per node (INT_MAX if none)
    int32_t *real_in;        // [N] number of non-dummy in-edges
    int32_t *kept;           // [E+1] survivor flag of each edge-as-vertex, then its scan
    char *scan_ws;
    size_t scan_bytes;
};

static bool carve_conj_ws(void *ws, size_t ws_bytes, int32_t B, int64_t N, int64_t E, ConjWs *c) {
    WsCarver w(ws, ws_bytes);
    c->egraph = w.take<int32_t>(E + 1);
    c->first_dummy = w.take<int32_t>(B + 1);
    c->fd_in = w.take<int32_t>(N + 1);
    c->fd_out = w.take<int32_t>(N + 1);
    c->real_in = w.take<int32_t>(N + 1);
    c->kept = w.take<int32_t>(E + 2);
    c->scan_bytes = dn4gl_scan_workspace_bytes(E + 1);
    c->scan_ws = w.take<char>(c->scan_bytes);
    return c->egraph && c->first_dummy && c->fd_in && c->fd_out && c->real_in && c->kept && c->scan_ws;
}

extern "C" size_t dn4gl_conj_workspace_bytes(int32_t B, int64_t N, int64_t E) {
    auto a = [](size_t n) { return align_up(n * sizeof(int32_t), 256); };
    return a(E + 1) + a(B + 1) + 3 * a(N + 1) + a(E + 2) + dn4gl_scan_workspace_bytes(E + 1) + 256;
}

__global__ void fill_i32(int32_t *p, int64_t n, int32_t v) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

__global__ void conj_edge_pass1(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ src,
                                const int32_t *__restrict__ dst, const int32_t *__restrict__ isd, int64_t E,
                                ConjWs c) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int g = segment_of(edge_ptr, B, e);
    c.egraph[e] = g;
    bool d = isd && isd[e];
    if (d) {
        atomicMin(c.first_dummy + g, static_cast<int32_t>(e));
        atomicMin(c.fd_in + dst[e], static_cast<int32_t>(e));
        atomicMin(c.fd_out + src[e], static_cast<int32_t>(e));
    } else {
        atomicAdd(c.real_in + dst[e], 1);
    }
}

// cand_cnt[e] (written into cand_off[e]) and kept[e]
__global__ void conj_edge_pass2(const int32_t *__restrict__ src, const int32_t *__restrict__ isd, int64_t E,
                                ConjWs c, int32_t *__restrict__ cand_cnt) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    int s = src[e];
    bool d = isd && isd[e];
    int cnt;
    if (!d) cnt = c.real_in[s] + (c.fd_in[s] != INT32_MAX ? 1 : 0);
    else cnt = (c.fd_out[s] == e) ? c.real_in[s] : 0;
    cand_cnt[e] = cnt;
    c.kept[e] = (!d || c.first_dummy[c.egraph[e]] == e) ? 1 : 0;
}

__global__ void conj_graph_offsets(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ kept_scan,
                                   const int32_t *__restrict__ cand_off, int32_t *__restrict__ o_node_ptr,
                                   int32_t *__restrict__ o_edge_ptr) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g <= B) {
        o_node_ptr[g] = kept_scan[edge_ptr[g]];
        o_edge_ptr[g] = cand_off[edge_ptr[g]];
    }
}

// newid[e] = global conjugate vertex id that edge e maps to
__global__ void conj_newid(const int32_t *__restrict__ isd, int64_t E, ConjWs c, int32_t *__restrict__ newid) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    bool d = isd && isd[e];
    int rep = d ? c.first_dummy[c.egraph[e]] : static_cast<int32_t>(e);
    newid[e] = c.kept[rep];  // kept[] holds the exclusive scan by now
}

extern "C" int dn4gl_tu_conjugate_count(int32_t B, const int32_t *node_ptr, const int32_t *edge_ptr,
                                        const int32_t *src, const int32_t *dst, const int32_t *e_isdummy, int64_t N,
                                        int64_t E, const int32_t *in_ptr, const int32_t *in_eid, int32_t *cand_off,
                                        int32_t *newid, int32_t *o_node_ptr, int32_t *o_edge_ptr, void *ws,
                                        size_t ws_bytes, void *stream) {
    (void)node_ptr; (void)in_ptr; (void)in_eid;
    DN_ARG(B >= 0 && N >= 0 && E >= 0 && edge_ptr && cand_off && o_node_ptr && o_edge_ptr);
    DN_ARG(E == 0 || (src && dst && newid));
    cudaStream_t st = as_stream(stream);
    ConjWs c;
    if (!carve_conj_ws(ws, ws_bytes, B, N, E, &c)) {
        dn4gl_set_error("dn4gl_tu_conjugate_count: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    auto blocks = [](int64_t n) { return static_cast<unsigned>(ceil_div64(n > 0 ? n : 1, 256)); };
    fill_i32<<<blocks(B + 1), 256, 0, st>>>(c.first_dummy, B + 1, INT32_MAX);
    fill_i32<<<blocks(N + 1), 256, 0, st>>>(c.fd_in, N + 1, INT32_MAX);
    fill_i32<<<blocks(N + 1), 256, 0, st>>>(c.fd_out, N + 1, INT32_MAX);
    DN_CUDA(cudaMemsetAsync(c.real_in, 0, static_cast<size_t>(N + 1) * sizeof(int32_t), st));
    DN_LAUNCHED_N(3);
    if (E > 0) {
        conj_edge_pass1<<<blocks(E), 256, 0, st>>>(B, edge_ptr, src, dst, e_isdummy, E, c);
        conj_edge_pass2<<<blocks(E), 256, 0, st>>>(src, e_isdummy, E, c, cand_off);
        DN_LAUNCHED_N(2);
    }
    int rc = dn4gl_exclusive_scan_i32(cand_off, cand_off, E, c.scan_ws, c.scan_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    rc = dn4gl_exclusive_scan_i32(c.kept, c.kept, E, c.scan_ws, c.scan_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    conj_graph_offsets<<<blocks(B + 1), 256, 0, st>>>(B, edge_ptr, c.kept, cand_off, o_node_ptr, o_edge_ptr);
    if (E > 0) conj_newid<<<blocks(E), 256, 0, st>>>(e_isdummy, E, c, newid);
    DN_LAUNCHED_N(E > 0 ? 2 : 1);
    return DN4GL_OK;
}

__global__ void conj_fill_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ isd, int64_t E,
                                 const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid,
                                 const int32_t *__restrict__ cand_off, const int32_t *__restrict__ newid, ConjWs c,
                                 int32_t *__restrict__ o_src, int32_t *__restrict__ o_dst,
                                 int32_t *__restrict__ o_v_origin, int32_t *__restrict__ o_e_shared,
                                 int64_t cap_v, int64_t cap_e, int32_t *__restrict__ err_flag) {
    int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= E) return;
    bool d = isd && isd[e];
    // conjugate vertex <- original edge (survivors only; kept[] is the scan, so compare neighbours)
    if (c.kept[e + 1] != c.kept[e]) {
        if (c.kept[e] < cap_v) o_v_origin[c.kept[e]] = static_cast<int32_t>(e);
        else if (err_flag) atomicExch(err_flag, DN4GL_ECAPACITY);
    }
    int pos = cand_off[e];
    if (cand_off[e + 1] == pos) return;
    if (cand_off[e + 1] > cap_e) {          // the caller's buffers (sized from a host-side hint) are too small: write nothing
        if (err_flag) atomicExch(err_flag, DN4GL_ECAPACITY);
        return;
    }
    int s = src[e];
    int me = newid[e];
    int fd_in = c.fd_in[s];
    for (int p = in_ptr[s]; p < in_ptr[s + 1]; ++p) {
        int ep = in_eid[p];
        bool dp = isd && isd[ep];
        bool keep = d ? !dp : (!dp || ep == fd_in);
        if (keep) {
            o_src[pos] = newid[ep];
            o_dst[pos] = me;
            o_e_shared[pos] = s;
            ++pos;
        }
    }
}

extern "C" int dn4gl_tu_conjugate_fill(int32_t B, const int32_t *node_ptr, const int32_t *edge_ptr,
                                       const int32_t *src, const int32_t *dst, const int32_t *e_isdummy, int64_t N,
                                       int64_t E, const int32_t *in_ptr, const int32_t *in_eid,
                                       const int32_t *cand_off, const int32_t *newid, const int32_t *o_node_ptr,
                                       const int32_t *o_edge_ptr, int32_t *o_src, int32_t *o_dst,
                                       int32_t *o_v_origin, int32_t *o_e_shared, int64_t cap_v, int64_t cap_e,
                                       int32_t *err_flag, void *ws, size_t ws_bytes, void *stream) {
    (void)node_ptr; (void)edge_ptr; (void)dst; (void)o_node_ptr; (void)o_edge_ptr;
    DN_ARG(B >= 0 && N >= 0 && E >= 0 && cap_v >= 0 && cap_e >= 0);
    if (E == 0) return DN4GL_OK;
    DN_ARG(src && in_ptr && in_eid && cand_off && newid && o_v_origin);
    ConjWs c;
    if (!carve_conj_ws(ws, ws_bytes, B, N, E, &c)) {  // same workspace as _count, contents preserved
        dn4gl_set_error("dn4gl_tu_conjugate_fill: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    conj_fill_kernel<<<static_cast<unsigned>(ceil_div64(E, 128)), 128, 0, as_stream(stream)>>>(
        src, e_isdummy, E, in_ptr, in_eid, cand_off, newid, c, o_src, o_dst, o_v_origin, o_e_shared, cap_v, cap_e, err_flag);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// ===========================================================================================
// a2 + a3 in closed form for CONJ_ graphs (dummy + edge-to-vertex + PyG canonicalisation), round 2.
//
// After tu_add_dummy every node t has exactly one dummy out-edge (t -> n) and one dummy in-edge (n -> t), all dummy
// edges merge into ONE conjugate vertex D (tu_data_processing.py:289-318), and PyG's read_tu_data then removes self
// loops and sorts / merges the edge list (graph_neural_networks/dataset.py:151).  For a graph with m real edges the
// canonical CONJ_ structure is therefore known row by row without generating, sorting and compacting candidates:
//   vertices   real edge e  -> conjugate vertex e (its position among the graph's edges), D = vertex m
//   out(e)     = { e2 real : src(e2) = dst(e), e2 != e } ascending, then D        (e2 = e only for a self loop)
//   in(e)      = { e2 real : dst(e2) = src(e), e2 != e } ascending, then D
//   out(D) = in(D) = all real edges 0 .. m-1                                     ((D, D) is dropped, :311-318)
// No pair occurs twice (one dummy edge per node and direction; parallel real edges are distinct vertices), so the rows
// are exactly the rows of the (row, col)-sorted, coalesced edge_index, and the in-rows the rows of its transpose in
// ascending source order -- the same CSR pair dn4gl_tu_conjugate_* + dn4gl_coalesce + two dn4gl_build_csr produce
// (tests/test_zzz_conj_direct_gpu.py compares them bit for bit), with 3 launches instead of ~35 and no sort.
// Needs: every graph has at least one node (else it has no D), the real graph's CSR by source and by destination
// (items in ascending edge id).  Global conjugate vertex id of real edge e of graph g: e + g;  D_g = edge_ptr[g+1] + g.
__global__ void conj_direct_lens_kernel(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ src,
                                        const int32_t *__restrict__ dst, const int32_t *__restrict__ out_ptr,
                                        const int32_t *__restrict__ in_ptr, int64_t E, int32_t *__restrict__ len_out,
                                        int32_t *__restrict__ len_in, int32_t *__restrict__ o_node_ptr) {
    const int64_t u = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // conjugate vertex
    if (u <= B) o_node_ptr[u] = edge_ptr[u] + static_cast<int32_t>(u);
    if (u >= E + B) return;
    // graph of u: largest g with edge_ptr[g] + g <= u
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(edge_ptr[mid]) + mid <= u) lo = mid; else hi = mid;
    }
    const int g = lo;
    const int64_t e = u - g;
    if (e == edge_ptr[g + 1]) {                 // D_g
        const int m = edge_ptr[g + 1] - edge_ptr[g];
        len_out[u] = m;
        len_in[u] = m;
        return;
    }
    const int s = src[e], t = dst[e], self = (s == t) ? 1 : 0;
    len_out[u] = out_ptr[t + 1] - out_ptr[t] + 1 - self;
    len_in[u] = in_ptr[s + 1] - in_ptr[s] + 1 - self;
}

// rows of the real vertices: one THREAD per (vertex, direction) -- rows hold ~5 entries, the loads of a row are independent
// of each other, and 3e5 threads are resident at once, so the kernel takes a few dependent-load latencies in total (the
// first version, a warp per vertex with a binary search over the graph offsets each, took 73 us at C2).  The vertex's graph
// comes from the edge -> graph map (dn4gl_segment_ids_i32 over edge_ptr).
__global__ void __launch_bounds__(256)
conj_direct_rows_kernel(const int32_t *__restrict__ edge2graph, const int32_t *__restrict__ edge_ptr,
                        const int32_t *__restrict__ src, const int32_t *__restrict__ dst, const int32_t *__restrict__ elabel,
                        const int32_t *__restrict__ out_ptr, const int32_t *__restrict__ out_eid,
                        const int32_t *__restrict__ in_ptr, const int32_t *__restrict__ in_eid, int64_t E,
                        const int32_t *__restrict__ rp_out, const int32_t *__restrict__ rp_in,
                        int32_t *__restrict__ col_out, int32_t *__restrict__ col_in, int32_t *__restrict__ o_vlabel) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= 2 * E) return;
    const int64_t e = i >> 1;
    const bool inward = i & 1;
    const int g = edge2graph[e];
    const int64_t u = e + g;
    const int32_t Dg = edge_ptr[g + 1] + g;
    const int node = inward ? src[e] : dst[e];                    // in-row: in-edges of src(e); out-row: out-edges of dst(e)
    const int32_t *ptr = inward ? in_ptr : out_ptr, *items = inward ? in_eid : out_eid;
    int32_t *col = inward ? col_in : col_out;
    const int b0 = ptr[node], n = ptr[node + 1] - b0;
    int w = (inward ? rp_in : rp_out)[u];
    for (int k = 0; k < n; ++k) {
        const int e2 = items[b0 + k];
        if (e2 != e) col[w++] = e2 + g;                           // e itself appears only for a self loop
    }
    col[w] = Dg;
    if (!inward) o_vlabel[u] = elabel ? elabel[e] : 1;
}

// rows of the merged dummy vertices D_g: all real edges of the graph, both directions; one warp per graph
__global__ void __launch_bounds__(256)
conj_direct_dummy_rows_kernel(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ rp_out,
                              const int32_t *__restrict__ rp_in, int32_t *__restrict__ col_out,
                              int32_t *__restrict__ col_in, int32_t *__restrict__ o_vlabel) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= B) return;
    const int m = edge_ptr[g + 1] - edge_ptr[g];
    const int32_t first = edge_ptr[g] + g, u = edge_ptr[g + 1] + g;
    const int wo = rp_out[u], wi = rp_in[u];
    for (int k = lane; k < m; k += 32) {
        col_out[wo + k] = first + k;
        col_in[wi + k] = first + k;
    }
    if (lane == 0) o_vlabel[u] = 0;
}

extern "C" int dn4gl_tu_conj_direct_lens(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *dst,
                                         const int32_t *out_ptr, const int32_t *in_ptr, int64_t E, int32_t *len_out,
                                         int32_t *len_in, int32_t *o_node_ptr, void *stream) {
    DN_ARG(B >= 1 && E >= 0 && edge_ptr && out_ptr && in_ptr && len_out && len_in && o_node_ptr && (E == 0 || (src && dst)));
    const int64_t V = E + B;
    conj_direct_lens_kernel<<<static_cast<unsigned>(ceil_div64(V + 1, 256)), 256, 0, as_stream(stream)>>>(
        B, edge_ptr, src, dst, out_ptr, in_ptr, E, len_out, len_in, o_node_ptr);
    DN_LAUNCHED();
    return DN4GL_OK;
}

extern "C" int dn4gl_tu_conj_direct_fill(int32_t B, const int32_t *edge_ptr, const int32_t *edge2graph, const int32_t *src,
                                         const int32_t *dst, const int32_t *elabel, const int32_t *out_ptr,
                                         const int32_t *out_eid, const int32_t *in_ptr, const int32_t *in_eid, int64_t E,
                                         const int32_t *rp_out, const int32_t *rp_in, int32_t *col_out, int32_t *col_in,
                                         int32_t *o_vlabel, void *stream) {
    DN_ARG(B >= 1 && E >= 0 && edge_ptr && out_ptr && in_ptr && rp_out && rp_in && o_vlabel);
    DN_ARG(E == 0 || (edge2graph && src && dst && out_eid && in_eid && col_out && col_in));
    cudaStream_t st = as_stream(stream);
    if (E > 0)
        conj_direct_rows_kernel<<<static_cast<unsigned>(ceil_div64(2 * E, 256)), 256, 0, st>>>(
            edge2graph, edge_ptr, src, dst, elabel, out_ptr, out_eid, in_ptr, in_eid, E, rp_out, rp_in, col_out, col_in, o_vlabel);
    conj_direct_dummy_rows_kernel<<<static_cast<unsigned>(ceil_div64(static_cast<int64_t>(B) * 32, 256)), 256, 0, st>>>(
        B, edge_ptr, rp_out, rp_in, col_out, col_in, o_vlabel);
    DN_LAUNCHED_N(E > 0 ? 2 : 1);
    return DN4GL_OK;
}

// ===========================================================================================
// a3: PyG read_tu_data [ext, torch-geometric 2.0.2]: remove_self_loops + coalesce.
// Input: the edge list (src, dst in edge-id order) and its by-src CSR (row_ptr, items = edge ids
// from dn4gl_build_csr(key=src)).  Rows are sorted in place by (dst, edge id), self loops and
// repeated (src,dst) pairs are flagged out, survivors are compacted in (src, dst) order.
// one thread per CSR position p (rows are already sorted by (dst, edge id)): keep[p] = not a self loop and not a
// repeat of the previous destination in the same row
// (the row of position p is src[items[p]]: no search over row_ptr)
__global__ void coalesce_flags(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ items,
                               const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t N, int64_t E,
                               int32_t *__restrict__ keep) {
    int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= E) return;
    const int it = items[p];
    const int r = src[it], d = dst[it];
    int prev = -1;
    if (p > 0) {
        const int itp = items[p - 1];
        if (src[itp] == r) prev = dst[itp];
    }
    keep[p] = (d != r && d != prev) ? 1 : 0;
}

__global__ void coalesce_compact(const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ items,
                                 const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t N, int64_t E,
                                 const int32_t *__restrict__ keep_scan, int32_t *__restrict__ o_src,
                                 int32_t *__restrict__ o_dst, int32_t *__restrict__ o_first) {
    int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= E) return;
    int q = keep_scan[p];
    if (keep_scan[p + 1] != q) {
        const int it = items[p];
        o_src[q] = src[it];
        o_dst[q] = dst[it];
        o_first[q] = it;
    }
}

extern "C" size_t dn4gl_coalesce_workspace_bytes(int64_t N, int64_t E) {
    return align_up(static_cast<size_t>(N > 0 ? N : 1) * sizeof(int32_t), 256) + 256 +
           dn4gl_scan_workspace_bytes(E + 1);
}

extern "C" int dn4gl_coalesce(const int32_t *src, const int32_t *dst, int64_t N, int64_t E, const int32_t *row_ptr, int32_t *items,
                              int32_t *keep_scan, int32_t *o_src, int32_t *o_dst, int32_t *o_first, void *ws,
                              size_t ws_bytes, int32_t *err_flag, void *stream) {
    DN_ARG(N >= 0 && E >= 0 && row_ptr && keep_scan);
    cudaStream_t st = as_stream(stream);
    if (E == 0) {
        DN_CUDA(cudaMemsetAsync(keep_scan, 0, sizeof(int32_t), st));
        return DN4GL_OK;
    }
    DN_ARG(src && dst && items && o_src && o_dst && o_first);
    WsCarver w(ws, ws_bytes);
    int32_t *worklist = w.take<int32_t>(N);
    int32_t *work_count = w.take<int32_t>(1);
    size_t scan_bytes = dn4gl_scan_workspace_bytes(E + 1);
    char *scan_ws = w.take<char>(scan_bytes);
    if (!worklist || !work_count || !scan_ws) {
        dn4gl_set_error("dn4gl_coalesce: workspace too small");
        return DN4GL_EWORKSPACE;
    }
    int rc = dn4gl_sort_rows(row_ptr, N, items, dst, worklist, work_count, err_flag, st, nullptr, nullptr, false);
    if (rc != DN4GL_OK) return rc;
    coalesce_flags<<<static_cast<unsigned>(ceil_div64(E, 256)), 256, 0, st>>>(row_ptr, items, src, dst, N, E, keep_scan);
    DN_LAUNCHED();
    rc = dn4gl_exclusive_scan_i32(keep_scan, keep_scan, E, scan_ws, scan_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    coalesce_compact<<<static_cast<unsigned>(ceil_div64(E, 256)), 256, 0, st>>>(row_ptr, items, src, dst, N, E, keep_scan,
                                                                                 o_src, o_dst, o_first);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// Tail padding for a compacted edge list whose length only the device knows: entries [count[0], cap) of a and b are
// set to `fill` (the index of a trash row behind the last real row), so that every consumer can run with the host-known
// capacity `cap` instead of waiting for a device->host read of the count (transforms.pyg_canonicalize, deferred count).
__global__ void pad_tail_kernel(const int32_t *__restrict__ count, int64_t cap, int32_t *__restrict__ a, int32_t *__restrict__ b,
                                int32_t fill) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < cap && i >= count[0]) {
        a[i] = fill;
        if (b) b[i] = fill;
    }
}

extern "C" int dn4gl_pad_tail_i32(const int32_t *count, int64_t cap, int32_t *a, int32_t *b, int32_t fill, void *stream) {
    DN_ARG(cap >= 0 && count != nullptr && (cap == 0 || a != nullptr));
    if (cap == 0) return DN4GL_OK;
    pad_tail_kernel<<<static_cast<unsigned>(ceil_div64(cap, 256)), 256, 0, as_stream(stream)>>>(count, cap, a, b, fill);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// ===========================================================================================
// the remaining augmentation flags of subgraph_isomorphism/train.py on the same flat batch layout.
//
// add_reversed_edges (train.py:291-345, GraphAdj branch): every graph's m edges are followed by their m reversals
// (v, u) with id = max_ne + position-in-graph, label + max_nel and is_reversed = 1; the original rows keep
// is_reversed = 0 (DGL zero-fills the new column).  One thread per OUTPUT edge, closed-form offsets (2 * edge_ptr).
__global__ void sub_add_reversed_kernel(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ src,
                                        const int32_t *__restrict__ dst, const int32_t *__restrict__ eid,
                                        const int32_t *__restrict__ elabel, int64_t E, int max_ne, int max_nel,
                                        int32_t *__restrict__ o_edge_ptr, int32_t *__restrict__ o_src,
                                        int32_t *__restrict__ o_dst, int32_t *__restrict__ o_eid,
                                        int32_t *__restrict__ o_elabel, int32_t *__restrict__ o_erev) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i <= B) o_edge_ptr[i] = 2 * edge_ptr[i];
    if (i >= 2 * E) return;
    // graph g with 2 * edge_ptr[g] <= i < 2 * edge_ptr[g + 1]
    int lo = 0, hi = B;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (2ll * __ldg(edge_ptr + mid) <= i) lo = mid; else hi = mid;
    }
    const int e0 = __ldg(edge_ptr + lo), m = __ldg(edge_ptr + lo + 1) - e0;
    const int j = static_cast<int>(i - 2ll * e0);
    if (j < m) {
        const int e = e0 + j;
        o_src[i] = src[e]; o_dst[i] = dst[e]; o_eid[i] = eid[e]; o_elabel[i] = elabel[e]; o_erev[i] = 0;
    } else {
        const int k = j - m, e = e0 + k;
        o_src[i] = dst[e]; o_dst[i] = src[e]; o_eid[i] = max_ne + k; o_elabel[i] = elabel[e] + max_nel; o_erev[i] = 1;
    }
}

extern "C" int dn4gl_sub_add_reversed(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *dst,
                                      const int32_t *eid, const int32_t *elabel, int64_t E, int32_t max_ne, int32_t max_nel,
                                      int32_t *o_edge_ptr, int32_t *o_src, int32_t *o_dst, int32_t *o_eid,
                                      int32_t *o_elabel, int32_t *o_e_is_reversed, void *stream) {
    DN_ARG(B >= 0 && E >= 0 && 2 * E < INT32_MAX && edge_ptr && o_edge_ptr);
    DN_ARG(E == 0 || (src && dst && eid && elabel && o_src && o_dst && o_eid && o_elabel && o_e_is_reversed));
    const int64_t threads = (2 * E > B + 1) ? 2 * E : B + 1;
    sub_add_reversed_kernel<<<static_cast<unsigned>(ceil_div64(threads, 256)), 256, 0, as_stream(stream)>>>(
        B, edge_ptr, src, dst, eid, elabel, E, max_ne, max_nel, o_edge_ptr, o_src, o_dst, o_eid, o_elabel, o_e_is_reversed);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// remove_loops (train.py:270-288): drops every edge with u == v, order of the survivors preserved (DGL remove_edges).
__global__ void loop_flags_kernel(const int32_t *__restrict__ src, const int32_t *__restrict__ dst, int64_t E,
                                  int32_t *__restrict__ keep) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e < E) keep[e] = (src[e] != dst[e]) ? 1 : 0;
}
__global__ void loop_compact_kernel(const int32_t *__restrict__ keep_scan, int64_t E, int B,
                                    const int32_t *__restrict__ edge_ptr, int32_t *__restrict__ o_edge_ptr,
                                    int32_t *__restrict__ survivors) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e <= B) o_edge_ptr[e] = keep_scan[edge_ptr[e]];
    if (e < E) {
        const int q = keep_scan[e];
        if (keep_scan[e + 1] != q) survivors[q] = static_cast<int32_t>(e);
    }
}

extern "C" int dn4gl_remove_loops_mark(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *dst,
                                       int64_t E, int32_t *keep_scan, int32_t *o_edge_ptr, int32_t *survivors, void *ws,
                                       size_t ws_bytes, void *stream) {
    DN_ARG(B >= 0 && E >= 0 && edge_ptr && keep_scan && o_edge_ptr);
    DN_ARG(E == 0 || (src && dst && survivors));
    cudaStream_t st = as_stream(stream);
    if (E > 0) {
        loop_flags_kernel<<<static_cast<unsigned>(ceil_div64(E, 256)), 256, 0, st>>>(src, dst, E, keep_scan);
        DN_LAUNCHED();
    }
    int rc = dn4gl_exclusive_scan_i32(keep_scan, keep_scan, E, ws, ws_bytes, stream);
    if (rc != DN4GL_OK) return rc;
    const int64_t threads = (E > B + 1) ? E : B + 1;
    loop_compact_kernel<<<static_cast<unsigned>(ceil_div64(threads, 256)), 256, 0, st>>>(keep_scan, E, B, edge_ptr, o_edge_ptr,
                                                                                        survivors);
    DN_LAUNCHED();
    return DN4GL_OK;
}

// compute_largest_eigenvalues (utils/graph.py:41-71): per graph  max_e (out_deg[u] + in_deg[v])  and
// max_e (in_deg[u] + out_deg[v]).  One warp per graph, integer maxima (exact); a graph without edges reports 0.
__global__ void eigen_bounds_kernel(int B, const int32_t *__restrict__ edge_ptr, const int32_t *__restrict__ src,
                                    const int32_t *__restrict__ dst, const int32_t *__restrict__ in_deg,
                                    const int32_t *__restrict__ out_deg, int32_t *__restrict__ node_eig,
                                    int32_t *__restrict__ edge_eig) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= B) return;
    int mn = 0, me = 0;
    for (int e = edge_ptr[g] + lane; e < edge_ptr[g + 1]; e += 32) {
        const int u = src[e], v = dst[e];
        mn = max(mn, out_deg[u] + in_deg[v]);
        me = max(me, in_deg[u] + out_deg[v]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = max(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        me = max(me, __shfl_xor_sync(0xffffffffu, me, o));
    }
    if (lane == 0) { node_eig[g] = mn; edge_eig[g] = me; }
}

extern "C" int dn4gl_sub_eigen_bounds(int32_t B, const int32_t *edge_ptr, const int32_t *src, const int32_t *dst,
                                      const int32_t *in_deg, const int32_t *out_deg, int32_t *node_eig, int32_t *edge_eig,
                                      void *stream) {
    DN_ARG(B >= 0);
    if (B == 0) return DN4GL_OK;
    DN_ARG(edge_ptr && in_deg && out_deg && node_eig && edge_eig);
    eigen_bounds_kernel<<<static_cast<unsigned>(ceil_div64(static_cast<int64_t>(B) * 32, 256)), 256, 0, as_stream(stream)>>>(
        B, edge_ptr, src, dst, in_deg, out_deg, node_eig, edge_eig);
    DN_LAUNCHED();
    return DN4GL_OK;
}
