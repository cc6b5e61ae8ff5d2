/* treelab.c — development aid (not product, not oracle): CPU model of the device's BVH pipeline used to compare BINARY-TREE builders
 * by what matters to k_trace: wide nodes visited and triangles tested per incoherent closest-hit ray, after the same 8-wide collapse
 * (open the largest-area child until 8, spare slots split small leaves; bvh_build.cu k_collapse) and an ordered traversal.
 *   gcc -O2 -march=native -o /tmp/treelab tools/treelab/treelab.c -lm && /tmp/treelab [n_tris] [n_rays] [leaf_max] [builders: lbvh ext sah ploc] [soup|terrain]
 * Builders: lbvh (Morton order, split at the highest differing bit = the device's k_hierarchy), sah (binned top-down, the oracle's
 * kind of tree), ploc (radius 8, the device's k_ploc_*), and lbvh+X experiments. */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float lo[3], hi[3]; } box_t;
typedef struct { box_t b; int left, right; int count; int first; } bnode_t; /* left/right: >= 0 internal node, < 0 leaf ~prim_pos; */

static uint64_t rng_s = 0x9E3779B97F4A7C15ull;
static inline uint64_t rnd64(void) { rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17; return rng_s; }
static inline float rndf(void) { return (float)((rnd64() >> 40) * (1.0 / 16777216.0)); }

static int n_tris;
static float (*V)[3][3]; /* triangle vertices */
static box_t *pbox;      /* primitive boxes */

static inline box_t box_empty(void) { box_t b = {{FLT_MAX, FLT_MAX, FLT_MAX}, {-FLT_MAX, -FLT_MAX, -FLT_MAX}}; return b; }
static inline void box_grow(box_t *a, const box_t *b) { for (int k = 0; k < 3; k++) { if (b->lo[k] < a->lo[k]) a->lo[k] = b->lo[k]; if (b->hi[k] > a->hi[k]) a->hi[k] = b->hi[k]; } }
static inline float box_area(const box_t *b) { float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2]; return dx * dy + dy * dz + dz * dx; }

/* ---- binary tree storage: nodes[0..n_nodes), order[] = primitive at sorted position ---- */
static bnode_t *nodes; static int n_nodes; static int *order; static int root;

static uint64_t expand21(uint32_t x) { uint64_t v = x & 0x1fffff; v = (v | v << 32) & 0x1f00000000ffffull; v = (v | v << 16) & 0x1f0000ff0000ffull; v = (v | v << 8) & 0x100f00f00f00f00full; v = (v | v << 4) & 0x10c30c30c30c30c3ull; v = (v | v << 2) & 0x1249249249249249ull; return v; }
typedef struct { uint64_t key; int prim; } kv_t;
static int kv_cmp(const void *a, const void *b) { const kv_t *x = a, *y = b; return x->key < y->key ? -1 : x->key > y->key ? 1 : (x->prim - y->prim); }
static kv_t *kv;

static void morton_sort(int bits_total, int extended) {
    box_t cb = box_empty();
    for (int i = 0; i < n_tris; i++) { float c[3]; for (int k = 0; k < 3; k++) c[k] = 0.5f * pbox[i].lo[k] + 0.5f * pbox[i].hi[k]; box_t p = {{c[0], c[1], c[2]}, {c[0], c[1], c[2]}}; box_grow(&cb, &p); }
    float maxdiag = 0.f;
    if (extended) for (int i = 0; i < n_tris; i++) { float d = 0; for (int k = 0; k < 3; k++) { float e = pbox[i].hi[k] - pbox[i].lo[k]; d += e * e; } d = sqrtf(d); if (d > maxdiag) maxdiag = d; }
    for (int i = 0; i < n_tris; i++) {
        uint32_t q[3];
        for (int k = 0; k < 3; k++) { float c = 0.5f * pbox[i].lo[k] + 0.5f * pbox[i].hi[k]; float t = (c - cb.lo[k]) / (cb.hi[k] - cb.lo[k]); t = t < 0 ? 0 : t > 1 ? 1 : t; uint32_t v = (uint32_t)(t * 2097152.0f); q[k] = v > 2097151u ? 2097151u : v; }
        uint64_t code = expand21(q[0]) | expand21(q[1]) << 1 | expand21(q[2]) << 2;
        if (extended) { /* Vinkler et al. 2017: interleave a size bit after every `extended` spatial triplets */
            float d = 0; for (int k = 0; k < 3; k++) { float e = pbox[i].hi[k] - pbox[i].lo[k]; d += e * e; } d = sqrtf(d) / maxdiag;
            uint32_t sz = (uint32_t)(d * 1023.0f); uint64_t out = 0; int ob = 0, sb = 9;
            for (int b = 62; b >= 0 && ob < 63; b -= 3) { for (int j = 0; j < 3 && ob < 63; j++) { out = out << 1 | (code >> (b - j) & 1); ob++; } if (((62 - b) / 3 + 1) % extended == 0 && sb >= 0 && ob < 63) { out = out << 1 | (sz >> sb & 1); sb--; ob++; } }
            code = out << (63 - ob);
        }
        kv[i].key = code >> (63 - bits_total); kv[i].prim = i;
    }
    qsort(kv, n_tris, sizeof(kv_t), kv_cmp);
    for (int i = 0; i < n_tris; i++) order[i] = kv[i].prim;
}

static inline uint64_t delta(int i) { uint64_t x = kv[i].key ^ kv[i + 1].key; return x ? (x | 1ull << 63) : (uint64_t)(i ^ (i + 1)); }

/* returns child ref: >= 0 internal node id, < 0 leaf ~pos */
static int lbvh_rec(int lo, int hi) {
    if (lo == hi) return ~lo;
    int split = lo; uint64_t best = 0;
    for (int i = lo; i < hi; i++) { uint64_t d = delta(i); if (d > best) { best = d; split = i; } }
    int id = n_nodes++;
    int l = lbvh_rec(lo, split), r = lbvh_rec(split + 1, hi);
    nodes[id].left = l; nodes[id].right = r; nodes[id].count = hi - lo + 1; nodes[id].first = lo;
    return id;
}
static box_t ref_box(int ref) { return ref < 0 ? pbox[order[~ref]] : nodes[ref].b; }
static void fit_boxes(int ref) { if (ref < 0) return; fit_boxes(nodes[ref].left); fit_boxes(nodes[ref].right); box_t b = ref_box(nodes[ref].left), c = ref_box(nodes[ref].right); box_grow(&b, &c); nodes[ref].b = b; }
static int ref_count(int ref) { return ref < 0 ? 1 : nodes[ref].count; }

/* highest-differing-bit split found by binary search would be equivalent; the linear scan keeps the code obvious (n log n total) */
static void build_lbvh(int bits, int extended) { morton_sort(bits, extended); n_nodes = 0; root = lbvh_rec(0, n_tris - 1); fit_boxes(root); }

/* ---- binned SAH over the Morton order's primitives (order[] is permuted in place) ---- */
static int sah_rec(int lo, int hi) {
    if (lo == hi) return ~lo;
    box_t cb = box_empty(), bb = box_empty();
    for (int i = lo; i <= hi; i++) { const box_t *p = &pbox[order[i]]; box_grow(&bb, p); float c[3]; for (int k = 0; k < 3; k++) c[k] = 0.5f * (p->lo[k] + p->hi[k]); box_t q = {{c[0], c[1], c[2]}, {c[0], c[1], c[2]}}; box_grow(&cb, &q); }
    int n = hi - lo + 1, best_axis = -1, best_bin = 0; float best_cost = FLT_MAX;
    enum { NB = 32 };
    for (int ax = 0; ax < 3; ax++) {
        float ext = cb.hi[ax] - cb.lo[ax]; if (!(ext > 0)) continue;
        box_t bins[NB]; int cnt[NB]; for (int b = 0; b < NB; b++) { bins[b] = box_empty(); cnt[b] = 0; }
        for (int i = lo; i <= hi; i++) { const box_t *p = &pbox[order[i]]; float c = 0.5f * (p->lo[ax] + p->hi[ax]); int b = (int)((c - cb.lo[ax]) / ext * NB); if (b >= NB) b = NB - 1; box_grow(&bins[b], p); cnt[b]++; }
        float ra[NB]; box_t acc = box_empty(); int rc[NB], c2 = 0;
        for (int b = NB - 1; b > 0; b--) { box_grow(&acc, &bins[b]); c2 += cnt[b]; ra[b] = c2 ? box_area(&acc) : 0; rc[b] = c2; }
        acc = box_empty(); int c1 = 0;
        for (int b = 0; b < NB - 1; b++) { box_grow(&acc, &bins[b]); c1 += cnt[b]; if (!c1 || !rc[b + 1]) continue; float cost = box_area(&acc) * c1 + ra[b + 1] * rc[b + 1]; if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; } }
    }
    int mid;
    if (best_axis < 0) mid = lo + n / 2 - 1 + (n == 1);
    else {
        float ext = cb.hi[best_axis] - cb.lo[best_axis]; int i = lo, j = hi;
        while (i <= j) { const box_t *p = &pbox[order[i]]; float c = 0.5f * (p->lo[best_axis] + p->hi[best_axis]); int b = (int)((c - cb.lo[best_axis]) / ext * NB); if (b >= NB) b = NB - 1; if (b <= best_bin) i++; else { int t = order[i]; order[i] = order[j]; order[j] = t; j--; } }
        mid = i - 1; if (mid < lo || mid >= hi) mid = lo + n / 2 - 1;
    }
    int id = n_nodes++;
    int l = sah_rec(lo, mid), r = sah_rec(mid + 1, hi);
    nodes[id].left = l; nodes[id].right = r; nodes[id].count = n; nodes[id].first = lo;
    return id;
}
static void build_sah(void) { for (int i = 0; i < n_tris; i++) order[i] = i; n_nodes = 0; root = sah_rec(0, n_tris - 1); fit_boxes(root); }

static int ploc_over(int *ref, int n, int radius);
/* ---- hybrid (HLBVH with a SAH top, Garanzha et al. 2011): LBVH inside Morton clusters of about `target` primitives, binned SAH over the cluster roots ---- */
static int *items; /* cluster root refs, permuted by the top-level SAH */
static int top_sah_rec(int lo, int hi) {
    if (lo == hi) return items[lo];
    box_t cb = box_empty();
    for (int i = lo; i <= hi; i++) { box_t b = ref_box(items[i]); float c[3]; for (int k = 0; k < 3; k++) c[k] = 0.5f * (b.lo[k] + b.hi[k]); box_t q = {{c[0], c[1], c[2]}, {c[0], c[1], c[2]}}; box_grow(&cb, &q); }
    int n = hi - lo + 1, best_axis = -1, best_bin = 0; float best_cost = FLT_MAX; enum { NB = 32 };
    for (int ax = 0; ax < 3; ax++) {
        float ext = cb.hi[ax] - cb.lo[ax]; if (!(ext > 0)) continue;
        box_t bins[NB]; int cnt[NB]; for (int b = 0; b < NB; b++) { bins[b] = box_empty(); cnt[b] = 0; }
        for (int i = lo; i <= hi; i++) { box_t p = ref_box(items[i]); float c = 0.5f * (p.lo[ax] + p.hi[ax]); int b = (int)((c - cb.lo[ax]) / ext * NB); if (b >= NB) b = NB - 1; box_grow(&bins[b], &p); cnt[b] += ref_count(items[i]); }
        float ra[NB]; box_t acc = box_empty(); int rc[NB], c2 = 0;
        for (int b = NB - 1; b > 0; b--) { box_grow(&acc, &bins[b]); c2 += cnt[b]; ra[b] = c2 ? box_area(&acc) : 0; rc[b] = c2; }
        acc = box_empty(); int c1 = 0;
        for (int b = 0; b < NB - 1; b++) { box_grow(&acc, &bins[b]); c1 += cnt[b]; if (!c1 || !rc[b + 1]) continue; float cost = box_area(&acc) * c1 + ra[b + 1] * rc[b + 1]; if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; } }
    }
    int mid;
    if (best_axis < 0) mid = lo + n / 2 - 1;
    else { float ext = cb.hi[best_axis] - cb.lo[best_axis]; int i = lo, j = hi;
        while (i <= j) { box_t p = ref_box(items[i]); float c = 0.5f * (p.lo[best_axis] + p.hi[best_axis]); int b = (int)((c - cb.lo[best_axis]) / ext * NB); if (b >= NB) b = NB - 1; if (b <= best_bin) i++; else { int t = items[i]; items[i] = items[j]; items[j] = t; j--; } }
        mid = i - 1; if (mid < lo || mid >= hi) mid = lo + n / 2 - 1; }
    int l = top_sah_rec(lo, mid), r = top_sah_rec(mid + 1, hi);
    int id = n_nodes++;
    nodes[id].left = l; nodes[id].right = r; nodes[id].count = ref_count(l) + ref_count(r); nodes[id].first = -1;
    box_t b = ref_box(l), c = ref_box(r); box_grow(&b, &c); nodes[id].b = b;
    return id;
}
static int g_cluster_ploc = 0;   /* 1: PLOC instead of the LBVH split rule inside the clusters */
static void build_hybrid(int bits, int target) {
    morton_sort(bits, 0); n_nodes = 0;
    int lg = 0; while ((1 << lg) < n_tris / target) lg++;
    int shift = bits - lg; if (shift < 0) shift = 0;
    items = malloc(sizeof(int) * n_tris); int n_items = 0;
    for (int lo = 0; lo < n_tris;) { int hi = lo; while (hi + 1 < n_tris && (kv[hi + 1].key >> shift) == (kv[lo].key >> shift)) hi++; int r; if (g_cluster_ploc && hi > lo) { int *ref = malloc(sizeof(int) * (hi - lo + 1)); for (int q = lo; q <= hi; q++) ref[q - lo] = ~q; r = ploc_over(ref, hi - lo + 1, 8); } else { r = lbvh_rec(lo, hi); fit_boxes(r); } items[n_items++] = r; lo = hi + 1; }
    root = top_sah_rec(0, n_items - 1);
    printf("    (hybrid: %d clusters of ~%d primitives)\n", n_items, n_tris / n_items);
    free(items);
}

/* ---- PLOC (Meister & Bittner 2018), radius r, over the Morton order ---- */
static int ploc_over(int *ref, int n, int radius) {   /* agglomerates the clusters ref[0..n) (in order) into one tree; returns its root */
    int *ref2 = malloc(sizeof(int) * n), *nn = malloc(sizeof(int) * n); box_t *cb = malloc(sizeof(box_t) * n), *cb2 = malloc(sizeof(box_t) * n);
    for (int i = 0; i < n; i++) cb[i] = ref_box(ref[i]);
    int *cur = ref;
    while (n > 1) {
        for (int i = 0; i < n; i++) { float best = FLT_MAX; int bj = -1; for (int j = i - radius; j <= i + radius; j++) { if (j == i || j < 0 || j >= n) continue; box_t u = cb[i]; box_grow(&u, &cb[j]); float a = box_area(&u); if (a < best) { best = a; bj = j; } } nn[i] = bj; }
        int m = 0;
        for (int i = 0; i < n; i++) {
            int j = nn[i];
            if (j >= 0 && nn[j] == i) { if (i < j) { int id = n_nodes++; nodes[id].left = cur[i]; nodes[id].right = cur[j]; nodes[id].count = ref_count(cur[i]) + ref_count(cur[j]); nodes[id].first = -1; box_t u = cb[i]; box_grow(&u, &cb[j]); nodes[id].b = u; ref2[m] = id; cb2[m] = u; m++; } }
            else { ref2[m] = cur[i]; cb2[m] = cb[i]; m++; }
        }
        int *t = cur; cur = ref2; ref2 = t; box_t *tb = cb; cb = cb2; cb2 = tb; n = m;
    }
    int r = cur[0];
    if (cur == ref) free(ref2); else free(cur == ref2 ? ref2 : cur);   /* one of the two scratch arrays is the caller's */
    free(nn); free(cb); free(cb2);
    return r;
}
static void build_ploc(int bits, int radius) {
    morton_sort(bits, 0); n_nodes = 0;
    int *ref = malloc(sizeof(int) * n_tris);
    for (int i = 0; i < n_tris; i++) ref[i] = ~i;
    root = ploc_over(ref, n_tris, radius);
}
/* LBVH inside Morton clusters, PLOC over the cluster roots: the device could do this with the kernels it has (k_hierarchy + k_ploc_* on n / 100 items) */
static void build_hybrid_ploc(int bits, int target, int radius) {
    morton_sort(bits, 0); n_nodes = 0;
    int lg = 0; while ((1 << lg) < n_tris / target) lg++;
    int shift = bits - lg; if (shift < 0) shift = 0;
    int *ref = malloc(sizeof(int) * n_tris); int n_items = 0;
    for (int lo = 0; lo < n_tris;) { int hi = lo; while (hi + 1 < n_tris && (kv[hi + 1].key >> shift) == (kv[lo].key >> shift)) hi++; int r = lbvh_rec(lo, hi); fit_boxes(r); ref[n_items++] = r; lo = hi + 1; }
    printf("    (hybrid: %d clusters of ~%d primitives, PLOC r%d on top)\n", n_items, n_tris / n_items, radius);
    root = ploc_over(ref, n_items, radius);
}

/* ---- bottom-up treelet-style improvement experiments go here ---- */
static double sah_cost_binary(void) { double c = 0; double ra = box_area(&nodes[root].b); for (int i = 0; i < n_nodes; i++) c += box_area(&nodes[i].b) / ra; return c; }

/* ---- 8-wide collapse (k_collapse's rule, including the greedy octant slot assignment and the 8-bit plane quantisation) ---- */
typedef struct { box_t cb[8]; box_t qb[8]; int child[8]; /* >= 0 wide node, -1 empty, <= -2: leaf, first packed prim = -(v + 2) */ int cnt[8]; } wnode_t;
static wnode_t *wn; static int n_wide; static int *packed; static int n_packed; static int leaf_max = 2;
static int slot_rule = 0;   /* slot assignment cost: 0 centroid offset . sign (the device), 1 entry corner . sign */
static int open_rule = 0;   /* which child the collapse opens next: 0 largest area (the device), 1 area * count, 2 area * log2(count), 3 largest count */
static void emit_leaves(int ref) { if (ref < 0) { packed[n_packed++] = order[~ref]; return; } emit_leaves(nodes[ref].left); emit_leaves(nodes[ref].right); }
static int collapse(int ref) {
    int id = n_wide++; int c[8], nc;
    if (ref < 0) { c[0] = ref; nc = 1; } else { c[0] = nodes[ref].left; c[1] = nodes[ref].right; nc = 2; }
    for (int phase = 0; phase < 2; phase++) { int limit = phase == 0 ? leaf_max : 1;
        while (nc < 8) { int who = -1; float best = -1; for (int i = 0; i < nc; i++) if (ref_count(c[i]) > limit) { box_t b = ref_box(c[i]); float a = box_area(&b); const float cn = (float)ref_count(c[i]); if (open_rule == 1) a *= cn; else if (open_rule == 2) a *= log2f(cn + 1.f); else if (open_rule == 3) a = cn; if (a > best) { best = a; who = i; } } if (who < 0) break; int r = c[who]; c[who] = nodes[r].left; c[nc++] = nodes[r].right; } }
    box_t cbx[8], nb = box_empty(); for (int i = 0; i < nc; i++) { cbx[i] = ref_box(c[i]); box_grow(&nb, &cbx[i]); }
    /* greedy slot assignment: global minimum of dot(child centre - node centre, sign_s) */
    int slot_of[8], used = 0; for (int i = 0; i < nc; i++) slot_of[i] = -1;
    for (int it = 0; it < nc; it++) { float best = FLT_MAX; int bc = -1, bs = -1;
        for (int i = 0; i < nc; i++) { if (slot_of[i] >= 0) continue; float cc[3]; for (int k = 0; k < 3; k++) cc[k] = 0.5f * (cbx[i].lo[k] + cbx[i].hi[k]) - 0.5f * (nb.lo[k] + nb.hi[k]);
            for (int sl = 0; sl < 8; sl++) { if (used >> sl & 1) continue; float cost = ((sl & 1) ? -cc[0] : cc[0]) + ((sl & 2) ? -cc[1] : cc[1]) + ((sl & 4) ? -cc[2] : cc[2]);
                if (slot_rule == 1) { cost = 0; for (int k = 0; k < 3; k++) cost += (sl >> k & 1) ? -(cbx[i].hi[k] - nb.hi[k]) : (cbx[i].lo[k] - nb.lo[k]); }   /* slot sl is first for directions negative along the axes of its set bits: they enter at hi */
                if (cost < best) { best = cost; bc = i; bs = sl; } } }
        slot_of[bc] = bs; used |= 1 << bs; }
    for (int i = 0; i < 8; i++) { wn[id].child[i] = -1; wn[id].cnt[i] = 0; }
    float scale[3]; for (int k = 0; k < 3; k++) { float sc = (nb.hi[k] - nb.lo[k]) / 255.0f; int e; float m = frexpf(sc, &e); scale[k] = sc > 0 ? ldexpf(1.0f, m == 0.5f ? e - 1 : e) : 1e-30f; }
    /* children are collapsed in slot order so that packed leaves and node ids follow the device's layout */
    for (int sl = 0; sl < 8; sl++) { int i = -1; for (int j = 0; j < nc; j++) if (slot_of[j] == sl) i = j; if (i < 0) continue;
        wn[id].cb[sl] = cbx[i];
        for (int k = 0; k < 3; k++) { wn[id].qb[sl].lo[k] = nb.lo[k] + floorf((cbx[i].lo[k] - nb.lo[k]) / scale[k]) * scale[k]; wn[id].qb[sl].hi[k] = nb.lo[k] + ceilf((cbx[i].hi[k] - nb.lo[k]) / scale[k]) * scale[k]; }
        if (ref_count(c[i]) > leaf_max) { int ch = collapse(c[i]); wn[id].child[sl] = ch; } else { wn[id].child[sl] = -(n_packed + 2); wn[id].cnt[sl] = ref_count(c[i]); emit_leaves(c[i]); } }
    return id;
}

/* ---- traversal: closest hit ---- */
static inline int tri_hit(const float o[3], const float d[3], int prim, float *t) {
    const float *a = V[prim][0], *b = V[prim][1], *c = V[prim][2];
    float e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    float p[3] = {d[1] * e2[2] - d[2] * e2[1], d[2] * e2[0] - d[0] * e2[2], d[0] * e2[1] - d[1] * e2[0]};
    float det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2]; if (det == 0) return 0; float inv = 1 / det;
    float s[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]}; float u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv; if (u < 0 || u > 1) return 0;
    float q[3] = {s[1] * e1[2] - s[2] * e1[1], s[2] * e1[0] - s[0] * e1[2], s[0] * e1[1] - s[1] * e1[0]};
    float v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv; if (v < 0 || u + v > 1) return 0;
    float tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv; if (tt > 1e-4f && tt < *t) { *t = tt; return 1; } return 0;
}
static double nodes_visited, tris_tested;
static int opt_quant = 0, opt_postpone = 0, opt_nearest = 0, opt_gtmin = 0, opt_sorted = 0, opt_childcull = 0;
static inline int box_hit(const box_t *b, const float o[3], const float inv[3], float tmax, float *tn) {
    float t0 = 1e-4f, t1 = tmax; for (int k = 0; k < 3; k++) { float a = (b->lo[k] - o[k]) * inv[k], c = (b->hi[k] - o[k]) * inv[k]; if (a > c) { float x = a; a = c; c = x; } if (a > t0) t0 = a; if (c < t1) t1 = c; } *tn = t0; return t0 <= t1;
}
/* ideal: children near to far by entry distance, leaves of a node before its internal children */
static void trace_ideal(const float o[3], const float d[3]) {
    float inv[3] = {1 / d[0], 1 / d[1], 1 / d[2]}; float tbest = 1e30f;
    struct { int node; float t; } stack[256]; int sp = 0; stack[sp].node = 0; stack[sp].t = 0; sp++;
    while (sp) { sp--; if (stack[sp].t > tbest) continue; const wnode_t *w = &wn[stack[sp].node]; nodes_visited++;
        int idx[8]; float te[8]; int nh = 0;
        for (int i = 0; i < 8; i++) { if (w->child[i] == -1) continue; float t0; if (box_hit(opt_quant ? &w->qb[i] : &w->cb[i], o, inv, tbest, &t0)) { idx[nh] = i; te[nh] = t0; nh++; } }
        for (int a = 1; a < nh; a++) { int ii = idx[a]; float tt = te[a]; int b = a - 1; while (b >= 0 && te[b] < tt) { idx[b + 1] = idx[b]; te[b + 1] = te[b]; b--; } idx[b + 1] = ii; te[b + 1] = tt; } /* descending: nearest last */
        for (int a = nh - 1; a >= 0; a--) { int i = idx[a]; if (w->child[i] <= -2) { int first = -(w->child[i] + 2); for (int q = 0; q < w->cnt[i]; q++) { tris_tested++; tri_hit(o, d, packed[first + q], &tbest); } } }
        for (int a = 0; a < nh; a++) { int i = idx[a]; if (w->child[i] >= 0 && te[a] <= tbest) { stack[sp].node = w->child[i]; stack[sp].t = te[a]; sp++; } }
    }
}
/* the device's loop (trace.cu k_trace): node group G = hit internal children of one node in octant priority, primitive group Gt = hit
 * leaf primitives in slot order; one node step when Gt is empty, then ONE triangle; the rest of Gt is postponed (pushed) while G has work */
typedef struct { int node; uint32_t bits; int kind; /* 0 node group: bits over priorities 0..7, 1 prim group: bits over the node's packed prims (<= 24) */ int pbase; int nearest; float tmin; float te[8]; } grp_t;
static void trace_device(const float o[3], const float d[3]) {
    float inv[3] = {1 / d[0], 1 / d[1], 1 / d[2]}; float tbest = 1e30f;
    const int octinv = ((d[0] >= 0) ? 1 : 0) | ((d[1] >= 0) ? 2 : 0) | ((d[2] >= 0) ? 4 : 0);   /* bit set = positive */
    grp_t stack[512]; int sp = 0; grp_t G = {-1, 1u << 7, 0, 0, -1, 0.f, {0}}, Gt = {0, 0, 1, 0, -1, 0.f, {0}};  /* G.node = -1: the root's pseudo parent */
    for (;;) {
        if (Gt.bits == 0 && G.bits != 0) {
            int pr = 31 - __builtin_clz(G.bits);
            if (opt_nearest && G.nearest >= 0) { pr = G.nearest; G.nearest = -1; }
            if (opt_sorted) { float bt = FLT_MAX; for (int q = 0; q < 8; q++) if ((G.bits >> q & 1) && G.te[q] < bt) { bt = G.te[q]; pr = q; } if (bt > tbest) { G.bits = 0; goto after_node; } }
            if (opt_childcull) { while (G.bits && G.te[31 - __builtin_clz(G.bits)] > tbest) G.bits &= ~(1u << (31 - __builtin_clz(G.bits))); if (!G.bits) goto after_node; pr = 31 - __builtin_clz(G.bits); }
            G.bits &= ~(1u << pr);
            int nid = G.node < 0 ? 0 : wn[G.node].child[pr ^ octinv];
            if (G.bits) stack[sp++] = G;
            const wnode_t *w = &wn[nid]; nodes_visited++;
            G.node = nid; G.bits = 0; G.kind = 0; Gt.node = nid; Gt.bits = 0; Gt.kind = 1;
            int poff = 0, pbase = -1; float te[8]; int hit[8];
            for (int i = 0; i < 8; i++) { hit[i] = 0; if (w->child[i] == -1) continue; hit[i] = box_hit(opt_quant ? &w->qb[i] : &w->cb[i], o, inv, tbest, &te[i]); }
            G.nearest = -1; G.tmin = FLT_MAX;
            for (int i = 0; i < 8; i++) if (hit[i] && w->child[i] >= 0) { G.bits |= 1u << (i ^ octinv); G.te[i ^ octinv] = te[i]; if (te[i] < G.tmin) { G.tmin = te[i]; G.nearest = i ^ octinv; } }
            for (int i = 0; i < 8; i++) { if (w->child[i] <= -2) { if (pbase < 0) pbase = -(w->child[i] + 2); if (hit[i]) for (int q = 0; q < w->cnt[i]; q++) Gt.bits |= 1u << (poff + q); poff += w->cnt[i]; } }
            Gt.pbase = pbase;
        }
        after_node:
        if (Gt.bits != 0) {
            int b = __builtin_ctz(Gt.bits); Gt.bits &= Gt.bits - 1; tris_tested++; tri_hit(o, d, packed[Gt.pbase + b], &tbest);
            if (opt_postpone && Gt.bits != 0 && G.bits != 0) { stack[sp++] = Gt; Gt.bits = 0; }
        }
        if (Gt.bits == 0 && G.bits == 0) { if (!sp) break; grp_t e = stack[--sp]; if (e.kind == 0) { if (opt_gtmin && e.tmin > tbest) continue; G = e; } else Gt = e; }
    }
}
static void (*trace)(const float o[3], const float d[3]) = trace_ideal;

static int n_rays; static float (*RO)[3], (*RD)[3];
static void run_rays(const char *tag) {
    nodes_visited = tris_tested = 0;
    for (int i = 0; i < n_rays; i++) trace(RO[i], RD[i]);
    printf("    %-44s nodes/ray %6.2f tris/ray %6.2f | est. cost (250 n + 70 t) %.0f\n", tag, nodes_visited / n_rays, tris_tested / n_rays, (250 * nodes_visited + 70 * tris_tested) / n_rays);
    fflush(stdout);
}
static void evaluate(const char *name) {
    n_wide = 0; n_packed = 0; collapse(root);
    printf("%-28s binary SAH %.1f | wide nodes %d\n", name, sah_cost_binary(), n_wide);
    trace = trace_ideal; opt_quant = 0; run_rays("ideal order, exact boxes");
    opt_quant = 1; run_rays("ideal order, 8-bit planes");
    trace = trace_device; opt_postpone = 0; run_rays("device loop, octant order, no postponing");
    opt_postpone = 1; run_rays("device loop, octant order, postponing");
    opt_nearest = 1; run_rays("  + nearest child first");
    opt_nearest = 0; opt_gtmin = 1; run_rays("  + group culled on pop by its min entry t");
    opt_nearest = 1; run_rays("  + both");
    opt_nearest = 0; opt_gtmin = 0; opt_sorted = 1; run_rays("  per-child entry t kept: sorted + culled");
    opt_sorted = 0; opt_childcull = 1; run_rays("  per-child entry t kept: octant order + culled"); opt_childcull = 0;
}

int main(int argc, char **argv) {
    n_tris = argc > 1 ? atoi(argv[1]) : 1000000; n_rays = argc > 2 ? atoi(argv[2]) : 100000; leaf_max = argc > 3 ? atoi(argv[3]) : 2;
    const char *only = argc > 4 ? argv[4] : "";
    V = malloc(sizeof(*V) * n_tris); pbox = malloc(sizeof(box_t) * n_tris); kv = malloc(sizeof(kv_t) * n_tris); order = malloc(sizeof(int) * n_tris);
    nodes = malloc(sizeof(bnode_t) * n_tris); wn = malloc(sizeof(wnode_t) * n_tris); packed = malloc(sizeof(int) * n_tris);
    const int terrain = argc > 5 && strstr(argv[5], "terrain");
    if (terrain) {
        /* nx x nx vertices over [0,1]^2, h = 0.1 * 5-octave value noise (tests/scenes.py terrain()); n_tris is rounded to 2 (nx-1)^2 */
        int nx = (int)sqrt(n_tris / 2.0) + 1; n_tris = 2 * (nx - 1) * (nx - 1);
        float *h = calloc((size_t)nx * nx, sizeof(float)); float amp = 0.5f; int freq = 4;
        for (int oct = 0; oct < 5; oct++) { float *g = malloc(sizeof(float) * (freq + 2) * (freq + 2)); for (int i = 0; i < (freq + 2) * (freq + 2); i++) g[i] = rndf();
            for (int z = 0; z < nx; z++) for (int x = 0; x < nx; x++) { float fx = (float)x / (nx - 1) * freq, fz = (float)z / (nx - 1) * freq; int ix = (int)fx < freq ? (int)fx : freq, iz = (int)fz < freq ? (int)fz : freq; float tx = fx - ix, tz = fz - iz; tx = tx * tx * (3 - 2 * tx); tz = tz * tz * (3 - 2 * tz);
                float a = g[iz * (freq + 2) + ix] * (1 - tx) + g[iz * (freq + 2) + ix + 1] * tx, b = g[(iz + 1) * (freq + 2) + ix] * (1 - tx) + g[(iz + 1) * (freq + 2) + ix + 1] * tx; h[z * nx + x] += amp * (a * (1 - tz) + b * tz); }
            free(g); amp *= 0.5f; freq *= 2; }
        int t = 0;
        for (int z = 0; z < nx - 1; z++) for (int x = 0; x < nx - 1; x++) for (int half = 0; half < 2; half++, t++) {
            int vx[3], vz[3]; if (!half) { vx[0] = x; vz[0] = z; vx[1] = x + 1; vz[1] = z; vx[2] = x + 1; vz[2] = z + 1; } else { vx[0] = x; vz[0] = z; vx[1] = x + 1; vz[1] = z + 1; vx[2] = x; vz[2] = z + 1; }
            box_t b = box_empty();
            for (int v = 0; v < 3; v++) { V[t][v][0] = (float)vx[v] / (nx - 1); V[t][v][1] = 0.1f * h[vz[v] * nx + vx[v]]; V[t][v][2] = (float)vz[v] / (nx - 1); for (int k = 0; k < 3; k++) { if (V[t][v][k] < b.lo[k]) b.lo[k] = V[t][v][k]; if (V[t][v][k] > b.hi[k]) b.hi[k] = V[t][v][k]; } }
            pbox[t] = b; }
        free(h);
    } else {
    const float extent = 0.01f;
    for (int i = 0; i < n_tris; i++) { float c[3], e1[3], e2[3]; for (int k = 0; k < 3; k++) { c[k] = rndf(); e1[k] = (rndf() * 2 - 1) * extent; e2[k] = (rndf() * 2 - 1) * extent; }
        box_t b = box_empty(); for (int k = 0; k < 3; k++) { V[i][0][k] = c[k] - (e1[k] + e2[k]) / 3; V[i][1][k] = V[i][0][k] + e1[k]; V[i][2][k] = V[i][0][k] + e2[k]; for (int v = 0; v < 3; v++) { if (V[i][v][k] < b.lo[k]) b.lo[k] = V[i][v][k]; if (V[i][v][k] > b.hi[k]) b.hi[k] = V[i][v][k]; } } pbox[i] = b; }
    }
    RO = malloc(sizeof(*RO) * n_rays); RD = malloc(sizeof(*RD) * n_rays);
    for (int i = 0; i < n_rays; i++) { for (int k = 0; k < 3; k++) RO[i][k] = rndf(); if (terrain) RO[i][1] = RO[i][1] * 0.3f + 0.1f; /* tools/trace_bench.py's terrain rays */ float z = rndf() * 2 - 1, phi = rndf() * 6.2831853f, r = sqrtf(fmaxf(0, 1 - z * z)); RD[i][0] = r * cosf(phi); RD[i][1] = r * sinf(phi); RD[i][2] = z; }
    printf("%s %d triangles, %d rays, leaf_max %d\n", terrain ? "terrain" : "soup", n_tris, n_rays, leaf_max);
    if (!*only || strstr(only, "lbvh")) { build_lbvh(32, 0); evaluate("lbvh 32-bit"); }
    if (strstr(only, "slots")) { build_lbvh(32, 0); for (slot_rule = 0; slot_rule < 2; slot_rule++) evaluate(slot_rule ? "slots by entry corner" : "slots by centroid (device)"); slot_rule = 0; }
    if (strstr(only, "rules")) { build_lbvh(32, 0); const char *nm[4] = {"open largest area (device)", "open largest area * count", "open largest area * log2 count", "open largest count"}; for (open_rule = 0; open_rule < 4; open_rule++) evaluate(nm[open_rule]); open_rule = 0; }
    if (!*only || strstr(only, "ext")) { for (int e = 2; e <= 4; e++) { build_lbvh(40, e); char nm[64]; snprintf(nm, 64, "lbvh extended (size bit / %d)", e); evaluate(nm); } }
    if (!*only || strstr(only, "sah")) { build_sah(); evaluate("binned SAH"); }
    if (strstr(only, "hybrid")) { int sizes[3] = {2, 8, 32}; for (int q = 0; q < 3; q++) { build_hybrid(32, sizes[q]); char nm[64]; snprintf(nm, 64, "lbvh clusters ~%d + SAH top", sizes[q]); evaluate(nm); } }
    if (strstr(only, "hps")) { g_cluster_ploc = 1; int sizes[2] = {8, 32}; for (int q = 0; q < 2; q++) { build_hybrid(32, sizes[q]); char nm[64]; snprintf(nm, 64, "ploc clusters ~%d + SAH top", sizes[q]); evaluate(nm); } g_cluster_ploc = 0; }
    if (strstr(only, "hploc")) { int sizes[2] = {2, 8}; for (int q = 0; q < 2; q++) { build_hybrid_ploc(32, sizes[q], 8); char nm[64]; snprintf(nm, 64, "lbvh clusters ~%d + PLOC top", sizes[q]); evaluate(nm); } }
    if (!*only || (strstr(only, "ploc") && !strstr(only, "hploc"))) { build_ploc(32, 8); evaluate("ploc r8"); }
    if (strstr(only, "plocx")) { build_ploc(32, 16); evaluate("ploc r16"); build_ploc(32, 32); evaluate("ploc r32"); }
    return 0;
}
