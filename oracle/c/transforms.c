/*
 * oracle/c/transforms.c -- TEST INFRASTRUCTURE ONLY (CPU checker / CPU baseline).
 *
 * Plain-C restatement of the reference's integer graph transforms on a *batched*
 * (block-diagonal) graph.  Each function follows the reference loop structure and cites
 * the file:line it restates.  Layout convention shared with the CUDA product:
 *   node_ptr[B+1], edge_ptr[B+1]  per-graph offsets into the node / edge arrays
 *   src[E], dst[E]                GLOBAL node ids (local id = global - node_ptr[g])
 * Outputs carry provenance indices so any attribute column can be gathered by the caller:
 *   conj vertex k  <- original edge   v_origin[k]   (global edge index)
 *   conj edge  j   <- shared original vertex e_shared[j] (global node index)
 *
 * Two-phase use: call with the big outputs NULL to obtain the sizes (out_*_ptr), allocate,
 * call again to fill.
 *
 * Parity pinning: checked against the reference's own Python executed under oracle/shims
 * (tests/test_oracle_vs_reference.py, tests/golden/), and SURVEY.md App. B vectors.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* small open-addressing hash set over (a,b,c) int triples; used for the reference's     */
/* ``used_keys`` python sets                                                             */
typedef struct { int64_t *k0; int32_t *k2; uint8_t *used; size_t cap; } tripset;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static int tripset_init(tripset *s, size_t n) {
    size_t cap = 16; while (cap < 2 * n + 2) cap <<= 1;
    s->cap = cap;
    s->k0 = (int64_t *)malloc(cap * sizeof(int64_t));
    s->k2 = (int32_t *)malloc(cap * sizeof(int32_t));
    s->used = (uint8_t *)calloc(cap, 1);
    return (s->k0 && s->k2 && s->used) ? 0 : -1;
}
static void tripset_free(tripset *s) { free(s->k0); free(s->k2); free(s->used); }
/* returns 1 if newly inserted, 0 if already present */
static int tripset_add(tripset *s, int32_t a, int32_t b, int32_t c) {
    int64_t k0 = ((int64_t)a << 32) | (uint32_t)b;
    size_t h = (size_t)(mix64((uint64_t)k0 * 31u + (uint32_t)c)) & (s->cap - 1);
    while (s->used[h]) {
        if (s->k0[h] == k0 && s->k2[h] == c) return 0;
        h = (h + 1) & (s->cap - 1);
    }
    s->used[h] = 1; s->k0[h] = k0; s->k2[h] = c;
    return 1;
}

/* in-edge lists (ascending edge id) for the nodes of one graph: CSR over local ids.      */
/* restates igraph ``sorted(g.incident(v, "in"))`` / DGL ``incidence_matrix("in")[v]``    */
static void build_in_lists(int n, int m, const int *src, const int *dst, int node0,
                           int *in_ptr /* n+1 */, int *in_e /* m */) {
    (void)src;
    memset(in_ptr, 0, (size_t)(n + 1) * sizeof(int));
    for (int e = 0; e < m; ++e) in_ptr[dst[e] - node0 + 1]++;
    for (int v = 0; v < n; ++v) in_ptr[v + 1] += in_ptr[v];
    int *cur = (int *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    memcpy(cur, in_ptr, (size_t)n * sizeof(int));
    for (int e = 0; e < m; ++e) in_e[cur[dst[e] - node0]++] = e; /* ascending e */
    free(cur);
}

/* ------------------------------------------------------------------------------------ */
/* a1: dummy augmentation, classification flavour                                        */
/* graph_classification/data_processing/tu_data_processing.py:186-214                    */
/*   node n = dummy (LABEL 0, IS_DUMMY 1); after the m real edges come 2n dummy edges     */
/*   INTERLEAVED (n,v),(v,n) for v = 0..n-1 (line 193); ID = position (213-214).          */
/* out sizes: N+B nodes, E+2N edges.                                                      */
int orc_tu_add_dummy(int B, const int *node_ptr, const int *edge_ptr,
                     const int *src, const int *dst, const int *vlabel, const int *elabel,
                     int *o_node_ptr, int *o_edge_ptr, int *o_src, int *o_dst,
                     int *o_vlabel, int *o_vdummy, int *o_elabel, int *o_edummy) {
    int no = 0, eo = 0;
    for (int g = 0; g < B; ++g) {
        int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
        int e0 = edge_ptr[g], m = edge_ptr[g + 1] - e0;
        o_node_ptr[g] = no; o_edge_ptr[g] = eo;
        for (int v = 0; v < n; ++v) { o_vlabel[no + v] = vlabel[n0 + v]; o_vdummy[no + v] = 0; }
        o_vlabel[no + n] = 0; o_vdummy[no + n] = 1;
        for (int e = 0; e < m; ++e) {
            o_src[eo + e] = src[e0 + e] - n0 + no;
            o_dst[eo + e] = dst[e0 + e] - n0 + no;
            o_elabel[eo + e] = elabel[e0 + e]; o_edummy[eo + e] = 0;
        }
        for (int v = 0; v < n; ++v) {
            int k = eo + m + 2 * v;
            o_src[k] = no + n; o_dst[k] = no + v;           /* (n, v) */
            o_src[k + 1] = no + v; o_dst[k + 1] = no + n;   /* (v, n) */
            o_elabel[k] = o_elabel[k + 1] = 0; o_edummy[k] = o_edummy[k + 1] = 1;
        }
        no += n + 1; eo += m + 2 * n;
    }
    o_node_ptr[B] = no; o_edge_ptr[B] = eo;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a2: edge-to-vertex ("conjugate") transform, classification flavour                    */
/* tu_data_processing.py:223-338, for graphs produced by load_graph_data_from_TUDatadir   */
/* (es["ID"] = range(E), so step 1's id-merge is the identity, lines 228-242).            */
/*  step 2 (259-274): for e in edge order, for e' in ascending in-edges of src(e):        */
/*           candidate (e' -> e) tagged with the shared vertex src(e); the (uid,label,vid)*/
/*           key set is kept although it cannot fire with unique ids.                     */
/*  step 3 (289-318): all IS_DUMMY edges collapse onto the first one; (D,D) dropped;      */
/*           (uid,vid) duplicates dropped keeping the first occurrence.                   */
/*  step 5 (333-336): merged-away vertices deleted, survivors keep relative order.        */
/* e_isdummy may be NULL (LINE_ graphs: no IS_DUMMY attribute).                           */
int orc_tu_conjugate(int B, const int *node_ptr, const int *edge_ptr,
                     const int *src, const int *dst, const int *vlabel, const int *e_isdummy,
                     int *o_node_ptr, int *o_edge_ptr,
                     int *o_src, int *o_dst, int *o_v_origin, int *o_e_shared) {
    int fill = (o_src != NULL);
    int vo = 0, eo = 0;
    for (int g = 0; g < B; ++g) {
        int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
        int e0 = edge_ptr[g], m = edge_ptr[g + 1] - e0;
        const int *s = src + e0, *d = dst + e0;
        o_node_ptr[g] = vo; o_edge_ptr[g] = eo;
        if (m == 0) continue;                              /* line 228/243: empty conj graph */
        int *in_ptr = (int *)malloc((size_t)(n + 1) * sizeof(int));
        int *in_e = (int *)malloc((size_t)m * sizeof(int));
        build_in_lists(n, m, s, d, n0, in_ptr, in_e);
        /* dummy representative + vertex renumbering after delete_vertices */
        int first_dummy = -1;
        int *newid = (int *)malloc((size_t)m * sizeof(int));
        int kept = 0;
        for (int e = 0; e < m; ++e) {
            int fl = e_isdummy ? e_isdummy[e0 + e] : 0;
            if (fl && first_dummy < 0) first_dummy = e;
            if (!fl || e == first_dummy) {
                newid[e] = kept;
                if (fill) o_v_origin[vo + kept] = e0 + e;
                kept++;
            } else newid[e] = -1;
        }
        size_t ncand = 0;
        for (int e = 0; e < m; ++e) { int u = s[e] - n0; ncand += (size_t)(in_ptr[u + 1] - in_ptr[u]); }
        tripset keys1, keys2;
        tripset_init(&keys1, ncand); tripset_init(&keys2, ncand + 1);
        if (first_dummy >= 0) tripset_add(&keys2, first_dummy, first_dummy, 0);   /* line 306 */
        int cnt = 0;
        for (int e = 0; e < m; ++e) {
            int u = s[e] - n0;
            int elabel = vlabel[n0 + u];
            for (int p = in_ptr[u]; p < in_ptr[u + 1]; ++p) {
                int ep = in_e[p];
                if (!tripset_add(&keys1, ep, elabel, e)) continue;                 /* 269-273 */
                int uid = ep, vid = e;
                if (first_dummy >= 0) {
                    if (e_isdummy[e0 + uid]) uid = first_dummy;                    /* 309-312 */
                    if (e_isdummy[e0 + vid]) vid = first_dummy;
                    if (!tripset_add(&keys2, uid, vid, 0)) continue;               /* 313-317 */
                }
                if (fill) {
                    o_src[eo + cnt] = vo + newid[uid];
                    o_dst[eo + cnt] = vo + newid[vid];
                    o_e_shared[eo + cnt] = n0 + u;
                }
                cnt++;
            }
        }
        tripset_free(&keys1); tripset_free(&keys2);
        free(in_ptr); free(in_e); free(newid);
        vo += kept; eo += cnt;
    }
    o_node_ptr[B] = vo; o_edge_ptr[B] = eo;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a4: dummy augmentation, subgraph-isomorphism flavour                                  */
/* subgraph_isomorphism/train.py:404-474 (GraphAdj branch) + dataset.py:1238-1293         */
/*  +1 node {id:max_nv, label:max_nvl, is_dummy:1}; +2n edges BLOCKED [u->d]*n, [d->u]*n  */
/*  with id = max_ne / max_ne+1, label = max_nel / max_nel+1, is_dummy=1,                 */
/*  is_reversed = 0..0 1..1; pre-existing rows zero-filled for the new columns.           */
/* e_isrev_in may be NULL (no REVFLAG column before augmentation).                        */
int orc_sub_add_dummy(int B, const int *node_ptr, const int *edge_ptr,
                      const int *src, const int *dst,
                      const int *vid, const int *vlabel, const int *eid, const int *elabel,
                      const int *e_isrev_in,
                      int max_nv, int max_nvl, int max_ne, int max_nel,
                      int *o_node_ptr, int *o_edge_ptr, int *o_src, int *o_dst,
                      int *o_vid, int *o_vlabel, int *o_vdummy,
                      int *o_eid, int *o_elabel, int *o_edummy, int *o_erev) {
    int no = 0, eo = 0;
    for (int g = 0; g < B; ++g) {
        int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
        int e0 = edge_ptr[g], m = edge_ptr[g + 1] - e0;
        o_node_ptr[g] = no; o_edge_ptr[g] = eo;
        for (int v = 0; v < n; ++v) {
            o_vid[no + v] = vid[n0 + v]; o_vlabel[no + v] = vlabel[n0 + v]; o_vdummy[no + v] = 0;
        }
        o_vid[no + n] = max_nv; o_vlabel[no + n] = max_nvl; o_vdummy[no + n] = 1;   /* 416-423 */
        for (int e = 0; e < m; ++e) {
            o_src[eo + e] = src[e0 + e] - n0 + no; o_dst[eo + e] = dst[e0 + e] - n0 + no;
            o_eid[eo + e] = eid[e0 + e]; o_elabel[eo + e] = elabel[e0 + e];
            o_edummy[eo + e] = 0; o_erev[eo + e] = e_isrev_in ? e_isrev_in[e0 + e] : 0;
        }
        for (int v = 0; v < n; ++v) {                                             /* 424-433 */
            int a = eo + m + v, b = eo + m + n + v;
            o_src[a] = no + v; o_dst[a] = no + n;
            o_src[b] = no + n; o_dst[b] = no + v;
            o_eid[a] = max_ne; o_eid[b] = max_ne + 1;
            o_elabel[a] = max_nel; o_elabel[b] = max_nel + 1;
            o_edummy[a] = o_edummy[b] = 1;
            o_erev[a] = 0; o_erev[b] = 1;
        }
        no += n + 1; eo += m + 2 * n;
    }
    o_node_ptr[B] = no; o_edge_ptr[B] = eo;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a5: conjugate transform, subgraph-isomorphism flavour                                 */
/* subgraph_isomorphism/utils/graph.py:74-175 (DGL branch) == :177-267 (igraph branch)    */
/*  vertices = distinct EDGEID values (merged by id equality; attribute row taken from    */
/*  the smallest edge index carrying the id, lines 86-101); candidates as in a2 but keyed */
/*  (uid, label(src(e)), vid) on the *ids* and deduplicated keeping the first (116-131);  */
/*  ids that no edge carries are removed with order-preserving compaction (167-170).      */
/* eid values are per-graph ids (>= 0).                                                   */
int orc_sub_conjugate(int B, const int *node_ptr, const int *edge_ptr,
                      const int *src, const int *dst, const int *vlabel, const int *eid,
                      int *o_node_ptr, int *o_edge_ptr,
                      int *o_src, int *o_dst, int *o_v_origin, int *o_e_shared) {
    int fill = (o_src != NULL);
    int vo = 0, eo = 0;
    for (int g = 0; g < B; ++g) {
        int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
        int e0 = edge_ptr[g], m = edge_ptr[g + 1] - e0;
        const int *s = src + e0, *d = dst + e0, *ids = eid + e0;
        o_node_ptr[g] = vo; o_edge_ptr[g] = eo;
        if (m == 0) continue;
        int maxid = 0;
        for (int e = 0; e < m; ++e) if (ids[e] > maxid) maxid = ids[e];
        int nid = maxid + 1;
        int *id2vertex = (int *)malloc((size_t)nid * sizeof(int));
        int *newid = (int *)malloc((size_t)nid * sizeof(int));
        for (int i = 0; i < nid; ++i) id2vertex[i] = -1;
        for (int e = 0; e < m; ++e) if (id2vertex[ids[e]] < 0) id2vertex[ids[e]] = e;  /* min e */
        int kept = 0;
        for (int i = 0; i < nid; ++i) {
            if (id2vertex[i] >= 0) {
                newid[i] = kept;
                if (fill) o_v_origin[vo + kept] = e0 + id2vertex[i];
                kept++;
            } else newid[i] = -1;
        }
        int *in_ptr = (int *)malloc((size_t)(n + 1) * sizeof(int));
        int *in_e = (int *)malloc((size_t)m * sizeof(int));
        build_in_lists(n, m, s, d, n0, in_ptr, in_e);
        size_t ncand = 0;
        for (int e = 0; e < m; ++e) { int u = s[e] - n0; ncand += (size_t)(in_ptr[u + 1] - in_ptr[u]); }
        tripset keys; tripset_init(&keys, ncand);
        int cnt = 0;
        for (int e = 0; e < m; ++e) {
            int u = s[e] - n0;
            int vid = ids[e], elabel = vlabel[n0 + u];
            for (int p = in_ptr[u]; p < in_ptr[u + 1]; ++p) {
                int uid = ids[in_e[p]];
                if (!tripset_add(&keys, uid, elabel, vid)) continue;
                if (fill) {
                    o_src[eo + cnt] = vo + newid[uid];
                    o_dst[eo + cnt] = vo + newid[vid];
                    o_e_shared[eo + cnt] = n0 + u;
                }
                cnt++;
            }
        }
        tripset_free(&keys);
        free(in_ptr); free(in_e); free(id2vertex); free(newid);
        vo += kept; eo += cnt;
    }
    o_node_ptr[B] = vo; o_edge_ptr[B] = eo;
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/* a3: PyG ``read_tu_data`` edge canonicalisation [ext: torch-geometric 2.0.2, restated   */
/* from its published behaviour -- SURVEY.md App. C]: remove_self_loops, then             */
/* coalesce(edge_index, edge_attr) = sort by (row, col) and SUM the attributes of         */
/* duplicate pairs.  Edge attributes here are one-hot labels, so the summed attribute of  */
/* a merged edge is the per-label multiplicity; it is returned as o_first (index of the   */
/* first original edge in eid order) + o_mult[(E_out) x R] counts.                        */
typedef struct { int64_t key; int32_t e; } keyed;
static int keyed_cmp(const void *a, const void *b) {
    const keyed *x = (const keyed *)a, *y = (const keyed *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return (x->e > y->e) - (x->e < y->e);
}
/* returns number of output edges; pass o_src == NULL to count only */
int64_t orc_pyg_coalesce(int64_t E, const int *src, const int *dst, const int *elabel0, int R,
                         int *o_src, int *o_dst, int *o_first, int *o_mult) {
    keyed *k = (keyed *)malloc((size_t)(E > 0 ? E : 1) * sizeof(keyed));
    int64_t m = 0;
    for (int64_t e = 0; e < E; ++e) {
        if (src[e] == dst[e]) continue;
        k[m].key = ((int64_t)src[e] << 32) | (uint32_t)dst[e]; k[m].e = (int32_t)e; m++;
    }
    qsort(k, (size_t)m, sizeof(keyed), keyed_cmp);
    int64_t out = -1;
    for (int64_t i = 0; i < m; ++i) {
        if (i == 0 || k[i].key != k[i - 1].key) {
            out++;
            if (o_src) {
                o_src[out] = src[k[i].e]; o_dst[out] = dst[k[i].e]; o_first[out] = k[i].e;
                if (o_mult) memset(o_mult + out * R, 0, (size_t)R * sizeof(int));
            }
        }
        if (o_src && o_mult && elabel0) o_mult[out * R + elabel0[k[i].e]]++;
    }
    free(k);
    return out + 1;
}

/* ------------------------------------------------------------------------------------ */
/* CSR by destination, stable in edge-id order: what DGL's gspmm / torch-scatter iterate  */
/* over (row v lists in-edges ascending).                                                 */
int orc_csr_by_dst(int64_t N, int64_t E, const int *src, const int *dst,
                   int *row_ptr, int *col, int *eid) {
    memset(row_ptr, 0, (size_t)(N + 1) * sizeof(int));
    for (int64_t e = 0; e < E; ++e) row_ptr[dst[e] + 1]++;
    for (int64_t v = 0; v < N; ++v) row_ptr[v + 1] += row_ptr[v];
    int *cur = (int *)malloc((size_t)(N > 0 ? N : 1) * sizeof(int));
    memcpy(cur, row_ptr, (size_t)N * sizeof(int));
    for (int64_t e = 0; e < E; ++e) { int p = cur[dst[e]]++; col[p] = src[e]; eid[p] = (int)e; }
    free(cur);
    return 0;
}

/* sum aggregation out[v] = self_scale*x[v] + sum_{e: dst(e)=v} x[src(e)] in eid order     */
/* (torch-scatter CPU scatter_add_ order / DGL copy_u-sum); fp32 sequential adds.          */
int orc_spmm_sum_f32(int64_t N, int64_t E, int D, const int *src, const int *dst,
                     const float *x, float self_scale, float *out) {
    for (int64_t i = 0; i < N * D; ++i) out[i] = 0.0f;
    for (int64_t e = 0; e < E; ++e) {
        const float *xs = x + (int64_t)src[e] * D; float *o = out + (int64_t)dst[e] * D;
        for (int j = 0; j < D; ++j) o[j] += xs[j];
    }
    if (self_scale != 0.0f)
        for (int64_t i = 0; i < N * D; ++i) out[i] += self_scale * x[i];
    return 0;
}
