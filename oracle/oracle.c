/*
 * oracle.c — CPU restatement of the reference ray-tracing hot path.  See oracle.h for the
 * reference file:line map.  TEST INFRASTRUCTURE ONLY (checker + CPU baseline), never the
 * product path.  PARITY UNPINNED per hit (Embree, the reference's arithmetic, is absent); pinned statistically against the reference's
 * own cbox.png (oracle.h).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -shared -fPIC (see oracle/Makefile).  -ffp-contract=off
 * is REQUIRED: the canonical arithmetic below is defined operation by operation.
 */
#define _GNU_SOURCE
#include "oracle.h"
#include <float.h>
#include <immintrin.h>
#include <math.h>
#include <pthread.h>
#include <sys/mman.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define MOD_PRIMITIVE 1u
#define MOD_TRANSFORM 2u
#define MOD_OPAQUE_ON 4u
#define MOD_OPAQUE_OFF 8u
#define MOD_VISIBILITY 16u
#define MOD_USER_ID 32u

/* ------------------------------------------------------------------------------------ */
/* scene model                                                                           */
/* ------------------------------------------------------------------------------------ */

typedef struct bvh_node {
    float lo[3], hi[3];
    uint32_t left;  /* internal: index of left child (right = left+1); leaf: first prim slot */
    uint32_t count; /* 0 = internal */
} bvh_node;

typedef struct __attribute__((aligned(32))) wnode {
    float lo[3][8], hi[3][8];
    uint32_t child[8];  /* internal child: wnode index; leaf child: first packed triangle */
    uint32_t count[8];  /* 0 = internal child, 1..4 = triangles of a leaf child, WIDE_EMPTY = unused slot (its box is inverted) */
} wnode;
#define WIDE_EMPTY 0xffffffffu
typedef struct ptri { float v0[3], v1[3], v2[3]; uint32_t prim; } ptri;

typedef struct mesh {
    const uint8_t *verts;
    size_t vstride, nverts;
    const uint8_t *indices;
    size_t istride, ntris;
    int built;
    bvh_node *nodes;
    uint32_t n_nodes;
    uint32_t *prims;
    /* mode 2 (the CPU baseline's fast path): the binary tree collapsed to 8-wide nodes tested with AVX2, leaves packed in tree order */
    struct wnode *wide; uint32_t n_wide;
    struct ptri *packed;
    /* curves (build_curve, accel.rs:142-203): float4 control points {x, y, z, radius}, u32 first-control-point index per segment */
    int is_curve, basis;
    const uint8_t *cps; size_t cp_stride, cp_count;
    const uint32_t *segs; size_t nsegs;
} mesh;

/* accel.rs:270-296 */
typedef struct instance {
    float affine[12];
    float inv[12];
    uint32_t user_id, visible;
    int opaque, valid;
    uint64_t mesh;
} instance;

struct oracle_scene {
    mesh *meshes;
    size_t n_meshes, cap_meshes;
    instance *insts;
    size_t n_insts, cap_insts;
};

static void die(const char *msg) {
    fprintf(stderr, "[oracle] fatal: %s\n", msg);
    abort(); /* reference aborts on every error: backend_impl/src/lib.rs:101-131 */
}

oracle_scene *oracle_scene_new(void) { return (oracle_scene *)calloc(1, sizeof(oracle_scene)); }

void oracle_scene_free(oracle_scene *s) {
    if (!s) return;
    for (size_t i = 0; i < s->n_meshes; i++) { free(s->meshes[i].nodes); free(s->meshes[i].prims); free(s->meshes[i].wide); free(s->meshes[i].packed); }
    free(s->meshes);
    free(s->insts);
    free(s);
}

uint64_t oracle_mesh_new(oracle_scene *s) {
    if (s->n_meshes == s->cap_meshes) {
        s->cap_meshes = s->cap_meshes ? 2 * s->cap_meshes : 8;
        s->meshes = (mesh *)realloc(s->meshes, s->cap_meshes * sizeof(mesh));
    }
    memset(&s->meshes[s->n_meshes], 0, sizeof(mesh));
    return s->n_meshes++;
}

void oracle_mesh_set(oracle_scene *s, uint64_t id, const void *vertices, size_t vstride, size_t nverts,
                     const void *indices, size_t istride, size_t ntris) {
    if (id >= s->n_meshes) die("bad mesh id");
    if (istride != 12) die("index stride must be 12 (api/runtime.cpp:191)");
    mesh *m = &s->meshes[id];
    m->verts = (const uint8_t *)vertices; m->vstride = vstride; m->nverts = nverts;
    m->indices = (const uint8_t *)indices; m->istride = istride; m->ntris = ntris;
}

void oracle_curve_set(oracle_scene *s, uint64_t id, int basis, const void *cps, size_t cp_stride, size_t cp_count, const uint32_t *segs, size_t nsegs) {
    if (id >= s->n_meshes) die("bad curve id");
    if (cp_stride < 16) die("cp buffer stride must be >= 16 (cpu/accel.rs:159)");
    if (basis < 0 || basis > 3) die("bad curve basis");
    mesh *m = &s->meshes[id];
    m->is_curve = 1; m->basis = basis; m->cps = (const uint8_t *)cps; m->cp_stride = cp_stride; m->cp_count = cp_count; m->segs = segs; m->nsegs = nsegs;
    m->built = 1;
}

static inline void tri_verts(const mesh *m, uint32_t prim, const float **a, const float **b, const float **c) {
    const uint32_t *ix = (const uint32_t *)(m->indices + (size_t)prim * m->istride);
    *a = (const float *)(m->verts + (size_t)ix[0] * m->vstride);
    *b = (const float *)(m->verts + (size_t)ix[1] * m->vstride);
    *c = (const float *)(m->verts + (size_t)ix[2] * m->vstride);
}

/* ------------------------------------------------------------------------------------ */
/* CPU BVH: binned SAH, leaves <= 4 triangles.  Only an accelerator for the oracle.      */
/* ------------------------------------------------------------------------------------ */

#define NBINS 16
typedef struct { float lo[3], hi[3]; } box3;
static inline void box_empty(box3 *b) { for (int k = 0; k < 3; k++) { b->lo[k] = FLT_MAX; b->hi[k] = -FLT_MAX; } }
static inline void box_grow(box3 *b, const box3 *o) { for (int k = 0; k < 3; k++) { if (o->lo[k] < b->lo[k]) b->lo[k] = o->lo[k]; if (o->hi[k] > b->hi[k]) b->hi[k] = o->hi[k]; } }
static inline float box_area(const box3 *b) {
    float dx = b->hi[0] - b->lo[0], dy = b->hi[1] - b->lo[1], dz = b->hi[2] - b->lo[2];
    if (dx < 0) return 0.f;
    return dx * dy + dy * dz + dz * dx;
}

typedef struct { uint32_t node, first, count; } build_task;
int oracle_hw_threads(void);

/* Binned-SAH binary BVH, built by a pool of threads: a task is one node with its primitive range; large tasks put their two halves
 * back on the shared stack, small ones are finished by the thread that holds them.  Splits depend on the range only, so the tree is the
 * same whatever the schedule (node numbering aside; hits never depend on the tree anyway). */
typedef struct build_ctx {
    mesh *m; const box3 *tb; const float *cen;
    uint32_t n_nodes;                 /* atomic */
    build_task *stack; size_t top, cap; long pending; int waiting, threads;
    pthread_mutex_t mu; pthread_cond_t cv;
} build_ctx;

/* processes one node; returns 1 and the two child tasks when it was split */
static int build_node(build_ctx *c, build_task t, build_task out[2]) {
    mesh *m = c->m; const box3 *tb = c->tb; const float *cen = c->cen;
    bvh_node *nd = &m->nodes[t.node];
    box3 nb, cb; box_empty(&nb); box_empty(&cb);
    for (uint32_t i = t.first; i < t.first + t.count; i++) {
        uint32_t p = m->prims[i];
        box_grow(&nb, &tb[p]);
        for (int k = 0; k < 3; k++) { float v = cen[3 * p + k]; if (v < cb.lo[k]) cb.lo[k] = v; if (v > cb.hi[k]) cb.hi[k] = v; }
    }
    memcpy(nd->lo, nb.lo, 12); memcpy(nd->hi, nb.hi, 12);
    if (t.count <= 4) { nd->left = t.first; nd->count = t.count; return 0; }
    /* binned SAH over the 3 axes */
    int best_axis = -1, best_bin = -1; float best_cost = FLT_MAX;
    for (int ax = 0; ax < 3; ax++) {
        float ext = cb.hi[ax] - cb.lo[ax];
        if (!(ext > 0)) continue;
        box3 bb[NBINS]; uint32_t bc[NBINS];
        for (int j = 0; j < NBINS; j++) { box_empty(&bb[j]); bc[j] = 0; }
        float scale = NBINS / ext;
        for (uint32_t i = t.first; i < t.first + t.count; i++) {
            uint32_t p = m->prims[i];
            int j = (int)((cen[3 * p + ax] - cb.lo[ax]) * scale); if (j >= NBINS) j = NBINS - 1; if (j < 0) j = 0;
            bc[j]++; box_grow(&bb[j], &tb[p]);
        }
        float la[NBINS], ra[NBINS]; uint32_t lc[NBINS], rc[NBINS];
        box3 acc; box_empty(&acc); uint32_t cnt = 0;
        for (int j = 0; j < NBINS - 1; j++) { box_grow(&acc, &bb[j]); cnt += bc[j]; la[j] = box_area(&acc); lc[j] = cnt; }
        box_empty(&acc); cnt = 0;
        for (int j = NBINS - 1; j > 0; j--) { box_grow(&acc, &bb[j]); cnt += bc[j]; ra[j - 1] = box_area(&acc); rc[j - 1] = cnt; }
        for (int j = 0; j < NBINS - 1; j++) {
            if (lc[j] == 0 || rc[j] == 0) continue;
            float cost = la[j] * lc[j] + ra[j] * rc[j];
            if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = j; }
        }
    }
    uint32_t mid;
    if (best_axis < 0) {
        mid = t.first + t.count / 2; /* all centroids coincide: split by position in list */
    } else {
        float ext = cb.hi[best_axis] - cb.lo[best_axis], scale = NBINS / ext;
        uint32_t i = t.first, j = t.first + t.count;
        while (i < j) {
            uint32_t p = m->prims[i];
            int b = (int)((cen[3 * p + best_axis] - cb.lo[best_axis]) * scale); if (b >= NBINS) b = NBINS - 1; if (b < 0) b = 0;
            if (b <= best_bin) i++; else { j--; m->prims[i] = m->prims[j]; m->prims[j] = p; }
        }
        mid = i;
        if (mid == t.first || mid == t.first + t.count) mid = t.first + t.count / 2;
    }
    nd->left = __atomic_fetch_add(&c->n_nodes, 2u, __ATOMIC_RELAXED); nd->count = 0;
    out[0] = (build_task){nd->left, t.first, mid - t.first};
    out[1] = (build_task){nd->left + 1, mid, t.first + t.count - mid};
    return 1;
}

static void build_subtree(build_ctx *c, build_task root) {  /* finishes a small task on a local stack */
    build_task local[128]; int top = 0; local[top++] = root;
    while (top) {
        build_task kids[2];
        if (build_node(c, local[--top], kids)) { if (top + 2 > 128) die("oracle build stack overflow"); local[top++] = kids[0]; local[top++] = kids[1]; }
    }
}

static void *build_worker(void *arg) {
    build_ctx *c = (build_ctx *)arg;
    for (;;) {
        pthread_mutex_lock(&c->mu);
        while (c->top == 0 && c->pending > 0) pthread_cond_wait(&c->cv, &c->mu);
        if (c->top == 0) { pthread_mutex_unlock(&c->mu); return NULL; }   /* nothing queued, nothing running */
        build_task t = c->stack[--c->top];
        pthread_mutex_unlock(&c->mu);
        long delta = -1;
        if (t.count <= 8192) build_subtree(c, t);
        else {
            build_task kids[2];
            if (build_node(c, t, kids)) {
                pthread_mutex_lock(&c->mu);
                if (c->top + 2 > c->cap) { c->cap *= 2; c->stack = (build_task *)realloc(c->stack, c->cap * sizeof(build_task)); }
                c->stack[c->top++] = kids[0]; c->stack[c->top++] = kids[1];
                c->pending += 2;
                pthread_mutex_unlock(&c->mu);
                pthread_cond_broadcast(&c->cv);
            }
        }
        pthread_mutex_lock(&c->mu);
        c->pending += delta;
        const int done = c->pending == 0;
        pthread_mutex_unlock(&c->mu);
        if (done) pthread_cond_broadcast(&c->cv);
    }
}

static void build_wide(mesh *m);

void oracle_mesh_commit(oracle_scene *s, uint64_t id) {
    if (id >= s->n_meshes) die("bad mesh id");
    mesh *m = &s->meshes[id];
    free(m->nodes); free(m->prims); free(m->wide); free(m->packed);
    m->nodes = NULL; m->prims = NULL; m->n_nodes = 0; m->wide = NULL; m->packed = NULL; m->n_wide = 0;
    m->built = 1;
    size_t n = m->ntris;
    if (n == 0) return;
    box3 *tb = (box3 *)malloc(n * sizeof(box3));
    float *cen = (float *)malloc(n * 3 * sizeof(float));
    m->prims = (uint32_t *)malloc(n * sizeof(uint32_t));
    m->nodes = (bvh_node *)malloc((2 * n) * sizeof(bvh_node));
    for (size_t i = 0; i < n; i++) {
        const float *a, *b, *c; tri_verts(m, (uint32_t)i, &a, &b, &c);
        for (int k = 0; k < 3; k++) {
            float lo = fminf(a[k], fminf(b[k], c[k])), hi = fmaxf(a[k], fmaxf(b[k], c[k]));
            tb[i].lo[k] = lo; tb[i].hi[k] = hi; cen[3 * i + k] = 0.5f * lo + 0.5f * hi;
        }
        m->prims[i] = (uint32_t)i;
    }
    build_ctx c; memset(&c, 0, sizeof(c));
    c.m = m; c.tb = tb; c.cen = cen; c.n_nodes = 1;
    c.cap = 256; c.stack = (build_task *)malloc(c.cap * sizeof(build_task));
    c.stack[c.top++] = (build_task){0, 0, (uint32_t)n}; c.pending = 1;
    pthread_mutex_init(&c.mu, NULL); pthread_cond_init(&c.cv, NULL);
    int threads = n < 65536 ? 1 : oracle_hw_threads();
    if (threads > 64) threads = 64;
    if (threads == 1) build_worker(&c);
    else {
        pthread_t th[64];
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, build_worker, &c);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&c.mu); pthread_cond_destroy(&c.cv);
    m->n_nodes = c.n_nodes;
    free(c.stack); free(tb); free(cen);
    build_wide(m);
}

/* ---- 8-wide collapse for mode 2 --------------------------------------------------------------------------------------------------- */
static uint32_t collapse_node(mesh *m, uint32_t bnode, uint32_t *n_wide, uint32_t *n_packed) {
    const uint32_t w = (*n_wide)++;
    uint32_t kids[8]; int nk = 0;
    if (m->nodes[bnode].count) kids[nk++] = bnode;   /* a tree that is one leaf */
    else { kids[nk++] = m->nodes[bnode].left; kids[nk++] = m->nodes[bnode].left + 1; }
    while (nk < 8) {   /* open the internal child with the largest surface until the node is full */
        int best = -1; float ba = -1.f;
        for (int i = 0; i < nk; i++) {
            const bvh_node *k = &m->nodes[kids[i]];
            if (k->count) continue;
            box3 b; memcpy(b.lo, k->lo, 12); memcpy(b.hi, k->hi, 12);
            const float a = box_area(&b);
            if (a > ba) { ba = a; best = i; }
        }
        if (best < 0) break;
        const uint32_t l = m->nodes[kids[best]].left;
        kids[best] = l; kids[nk++] = l + 1;
    }
    /* children first (their subtrees number themselves), then fill this node: the array may have moved */
    uint32_t child[8], count[8];
    for (int i = 0; i < nk; i++) {
        const bvh_node *k = &m->nodes[kids[i]];
        if (k->count) {
            child[i] = *n_packed; count[i] = k->count;
            for (uint32_t j = 0; j < k->count; j++) {
                const uint32_t p = m->prims[k->left + j];
                const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
                ptri *t = &m->packed[(*n_packed)++];
                memcpy(t->v0, a, 12); memcpy(t->v1, b, 12); memcpy(t->v2, c, 12); t->prim = p;
            }
        } else { count[i] = 0; child[i] = collapse_node(m, kids[i], n_wide, n_packed); }
    }
    wnode *nd = &m->wide[w];
    for (int i = 0; i < 8; i++) {
        if (i < nk) {
            const bvh_node *k = &m->nodes[kids[i]];
            for (int a = 0; a < 3; a++) { nd->lo[a][i] = k->lo[a]; nd->hi[a][i] = k->hi[a]; }
            nd->child[i] = child[i]; nd->count[i] = count[i];
        } else {
            for (int a = 0; a < 3; a++) { nd->lo[a][i] = INFINITY; nd->hi[a][i] = -INFINITY; }
            nd->child[i] = 0; nd->count[i] = WIDE_EMPTY;
        }
    }
    return w;
}

/* The two arrays a walk touches at random (150 MB of nodes + 40 MB of triangles for the 1 M-triangle soup) ask for huge pages: with
 * 4 KB pages nearly every node visit is also a TLB miss. */
static void *alloc_random_access(size_t bytes) {
    void *p = NULL;
    const size_t huge = (size_t)2 << 20;
    if (bytes >= huge) {
        if (posix_memalign(&p, huge, (bytes + huge - 1) / huge * huge) != 0) die("out of memory");
#ifdef MADV_HUGEPAGE
        madvise(p, (bytes + huge - 1) / huge * huge, MADV_HUGEPAGE);
#endif
    } else if (posix_memalign(&p, 64, bytes ? bytes : 64) != 0) die("out of memory");
    return p;
}

static void build_wide(mesh *m) {
    if (m->ntris == 0 || m->is_curve) return;
    m->wide = (wnode *)alloc_random_access((size_t)m->n_nodes * sizeof(wnode));
    m->packed = (ptri *)alloc_random_access(m->ntris * sizeof(ptri));
    uint32_t nw = 0, np = 0;
    collapse_node(m, 0, &nw, &np);
    m->n_wide = nw;
    wnode *shrunk = (wnode *)alloc_random_access((size_t)nw * sizeof(wnode));
    memcpy(shrunk, m->wide, (size_t)nw * sizeof(wnode)); free(m->wide); m->wide = shrunk;
}

void oracle_invert_affine(const float m[12], float inv[12]) {
    /* double adjugate inverse of the 3x3 part, translation = -inv*t, one rounding to fp32 */
    double a = m[0], b = m[1], c = m[2], d = m[4], e = m[5], f = m[6], g = m[8], h = m[9], i = m[10];
    double tx = m[3], ty = m[7], tz = m[11];
    double c00 = e * i - f * h, c01 = c * h - b * i, c02 = b * f - c * e;
    double c10 = f * g - d * i, c11 = a * i - c * g, c12 = c * d - a * f;
    double c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    double det = a * c00 + b * c10 + c * c20;
    double r = 1.0 / det;
    double n00 = c00 * r, n01 = c01 * r, n02 = c02 * r;
    double n10 = c10 * r, n11 = c11 * r, n12 = c12 * r;
    double n20 = c20 * r, n21 = c21 * r, n22 = c22 * r;
    inv[0] = (float)n00; inv[1] = (float)n01; inv[2] = (float)n02; inv[3] = (float)(-(n00 * tx + n01 * ty + n02 * tz));
    inv[4] = (float)n10; inv[5] = (float)n11; inv[6] = (float)n12; inv[7] = (float)(-(n10 * tx + n11 * ty + n12 * tz));
    inv[8] = (float)n20; inv[9] = (float)n21; inv[10] = (float)n22; inv[11] = (float)(-(n20 * tx + n21 * ty + n22 * tz));
}

static const float IDENTITY[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};

void oracle_accel_update(oracle_scene *s, uint32_t instance_count, const oracle_mod *mods, size_t n_mods) {
    /* grow with default (invalid) instances / shrink by popping: accel.rs:345-353 */
    if (instance_count > s->cap_insts) {
        s->cap_insts = instance_count;
        s->insts = (instance *)realloc(s->insts, s->cap_insts * sizeof(instance));
    }
    for (size_t i = s->n_insts; i < instance_count; i++) {
        instance *in = &s->insts[i];
        memset(in, 0, sizeof(*in));
        memcpy(in->affine, IDENTITY, sizeof(IDENTITY)); memcpy(in->inv, IDENTITY, sizeof(IDENTITY));
        in->visible = 0xff; in->opaque = 1; in->valid = 0;
    }
    s->n_insts = instance_count;
    for (size_t k = 0; k < n_mods; k++) {
        const oracle_mod *m = &mods[k];
        if (m->index >= s->n_insts) die("modification index out of range");
        instance *in = &s->insts[m->index];
        if (m->flags & MOD_PRIMITIVE) { /* accel.rs:355-377: resets mask/opaque, takes affine + user_id as given */
            if (m->mesh >= s->n_meshes || !s->meshes[m->mesh].built) die("Mesh not built");
            memcpy(in->affine, m->affine, sizeof(in->affine));
            in->visible = 0xff; in->user_id = m->user_id; in->opaque = 1; in->valid = 1; in->mesh = m->mesh;
        }
        if (m->flags & MOD_OPAQUE_ON) in->opaque = 1;
        else if (m->flags & MOD_OPAQUE_OFF) in->opaque = 0;
        if (m->flags & MOD_TRANSFORM) { if (!in->valid) die("TRANSFORM on empty instance"); memcpy(in->affine, m->affine, sizeof(in->affine)); }
        if (m->flags & MOD_VISIBILITY) { if (!in->valid) die("VISIBILITY on empty instance"); in->visible = m->visibility; }
        if (m->flags & MOD_USER_ID) { if (!in->valid) die("USER_ID on empty instance"); in->user_id = m->user_id; }
    }
    for (size_t i = 0; i < s->n_insts; i++) oracle_invert_affine(s->insts[i].affine, s->insts[i].inv);
}

void oracle_instance_transform(const oracle_scene *s, uint32_t i, float a[12]) { memcpy(a, s->insts[i].affine, 48); }
uint32_t oracle_instance_user_id(const oracle_scene *s, uint32_t i) { return s->insts[i].user_id; }
uint32_t oracle_instance_visibility(const oracle_scene *s, uint32_t i) { return s->insts[i].visible; }
uint32_t oracle_instance_count(const oracle_scene *s) { return (uint32_t)s->n_insts; }

/* ------------------------------------------------------------------------------------ */
/* The canonical fp32 arithmetic                                                         */
/* ------------------------------------------------------------------------------------ */

typedef struct ray_frame { /* per (ray, instance) */
    float o[3], d[3];
    int kx, ky, kz;
    float sx, sy, sz;
} ray_frame;

/* world -> object: o' = Minv*o + t (nested fmaf, innermost = z term + translation), d' = Minv*d */
static inline void xform_ray(const float inv[12], const float o[3], const float d[3], ray_frame *f) {
    for (int r = 0; r < 3; r++) {
        const float *m = inv + 4 * r;
        f->o[r] = fmaf(m[0], o[0], fmaf(m[1], o[1], fmaf(m[2], o[2], m[3])));
        f->d[r] = fmaf(m[0], d[0], fmaf(m[1], d[1], m[2] * d[2]));
    }
    int kz = 0;
    if (fabsf(f->d[1]) > fabsf(f->d[0])) kz = 1;
    if (fabsf(f->d[2]) > fabsf(f->d[kz])) kz = 2;
    int kx = (kz + 1) % 3, ky = (kx + 1) % 3;
    if (f->d[kz] < 0.0f) { int t = kx; kx = ky; ky = t; }
    f->kx = kx; f->ky = ky; f->kz = kz;
    f->sx = f->d[kx] / f->d[kz];
    f->sy = f->d[ky] / f->d[kz];
    f->sz = 1.0f / f->d[kz];
}

/* Watertight edge-function test (after Woop, Benthin, Wald 2013) with this fixed op order. */
static inline int canon_tri(const ray_frame *f, float tmin, float tmax, const float *v0, const float *v1, const float *v2,
                            float *t_out, float *u_out, float *v_out) {
    const int kx = f->kx, ky = f->ky, kz = f->kz;
    const float a_x = v0[kx] - f->o[kx], a_y = v0[ky] - f->o[ky], a_z = v0[kz] - f->o[kz];
    const float b_x = v1[kx] - f->o[kx], b_y = v1[ky] - f->o[ky], b_z = v1[kz] - f->o[kz];
    const float c_x = v2[kx] - f->o[kx], c_y = v2[ky] - f->o[ky], c_z = v2[kz] - f->o[kz];
    const float ax = fmaf(-f->sx, a_z, a_x), ay = fmaf(-f->sy, a_z, a_y);
    const float bx = fmaf(-f->sx, b_z, b_x), by = fmaf(-f->sy, b_z, b_y);
    const float cx = fmaf(-f->sx, c_z, c_x), cy = fmaf(-f->sy, c_z, c_y);
    float U = cx * by - cy * bx;
    float V = ax * cy - ay * cx;
    float W = bx * ay - by * ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        U = (float)((double)cx * (double)by - (double)cy * (double)bx);
        V = (float)((double)ax * (double)cy - (double)ay * (double)cx);
        W = (float)((double)bx * (double)ay - (double)by * (double)ax);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return 0;
    const float det = (U + V) + W;
    if (det == 0.0f) return 0;
    const float az = f->sz * a_z, bz = f->sz * b_z, cz = f->sz * c_z;
    const float T = fmaf(U, az, fmaf(V, bz, W * cz));
    const float t = T / det; /* IEEE division: t is correctly rounded from (T, det) */
    if (!(t > tmin && t <= tmax)) return 0;
    const float rdet = 1.0f / det;
    *t_out = t; *u_out = V * rdet; *v_out = W * rdet;
    return 1;
}

int oracle_canonical_triangle(const float o[3], const float d[3], float tmin, float tmax,
                              const float v0[3], const float v1[3], const float v2[3], float *t, float *u, float *v) {
    ray_frame f; xform_ray(IDENTITY, o, d, &f);
    return canon_tri(&f, tmin, tmax, v0, v1, v2, t, u, v);
}

typedef struct best_hit { float t, u, v; uint32_t inst, prim; int found; int curve; } best_hit;

static inline void consider(best_hit *b, float t, float u, float v, uint32_t inst, uint32_t prim) {
    /* closest; ties -> lowest (inst, prim) */
    if (!b->found || t < b->t || (t == b->t && (inst < b->inst || (inst == b->inst && prim < b->prim)))) {
        b->found = 1; b->t = t; b->u = u; b->v = v; b->inst = inst; b->prim = prim; b->curve = 0;
    }
}

/* ---- curves ------------------------------------------------------------------------------------
 * The surface is the sweep of a sphere of radius r(u) along c(u).  A segment is put into the power basis with the frontend's own
 * matrices (lc/src/rtx/curve.rs:88-139).  Linear segments ARE rounded cones (exact).  Cubic segments are cut at u = k/8 into 8 pieces
 * whose rounded cones (radii inflated by the chord-sag bound) only LOCATE a hit; refine_curve_hit() then solves for the point of the
 * true sweep (Newton, double): t agrees with a dense float64 sweep to 1e-5 (tests/test_curves.py).  Embree's round curves (the
 * reference) are intersected iteratively to their own tolerance: PARITY UNPINNED bit for bit, within tolerance by construction.
 * A hit reports prim = segment, bary = (u, -1) (accel.rs:491-494), entry t. */
#define CURVE_PIECES 8
static inline int filter_accept(const oracle_filter *flt, uint32_t inst, uint32_t prim, float u, float v);
static inline void lin4(const float m[4], float div, const float *q0, const float *q1, const float *q2, const float *q3, float out[4]) {
    for (int c = 0; c < 4; c++) out[c] = (((m[0] * q0[c] + m[1] * q1[c]) + m[2] * q2[c]) + m[3] * q3[c]) / div;
}
static void curve_basis(int basis, const float *q0, const float *q1, const float *q2, const float *q3, float a[4][4]) {
    static const float BS[4][4] = {{-1, 3, -3, 1}, {3, -6, 3, 0}, {-3, 0, 3, 0}, {1, 4, 1, 0}};
    static const float CR[4][4] = {{-1, 3, -3, 1}, {2, -5, 4, -1}, {-1, 0, 1, 0}, {0, 2, 0, 0}};
    static const float BZ[4][4] = {{-1, 3, -3, 1}, {3, -6, 3, 0}, {-3, 3, 0, 0}, {1, 0, 0, 0}};
    const float (*M)[4] = basis == 1 ? BS : (basis == 2 ? CR : BZ);
    const float div = basis == 1 ? 6.0f : (basis == 2 ? 2.0f : 1.0f);
    for (int r = 0; r < 4; r++) lin4(M[r], div, q0, q1, q2, q3, a[r]);
}
static inline void curve_point(float a[4][4], float u, float out[4]) {
    for (int c = 0; c < 4; c++) out[c] = ((a[0][c] * u + a[1][c]) * u + a[2][c]) * u + a[3][c];
}
static void curve_piece(const mesh *m, uint32_t seg, uint32_t k, float A[4], float B[4]) {
    const uint32_t first = m->segs[seg];
    const float *q0 = (const float *)(m->cps + (size_t)first * m->cp_stride), *q1 = (const float *)(m->cps + (size_t)(first + 1) * m->cp_stride);
    if (m->basis == 0) { memcpy(A, q0, 16); memcpy(B, q1, 16); return; }
    const float *q2 = (const float *)(m->cps + (size_t)(first + 2) * m->cp_stride), *q3 = (const float *)(m->cps + (size_t)(first + 3) * m->cp_stride);
    float a[4][4];
    curve_basis(m->basis, q0, q1, q2, q3, a);
    const float u0 = (float)k * (1.0f / CURVE_PIECES), u1 = (float)(k + 1) * (1.0f / CURVE_PIECES);
    curve_point(a, u0, A);
    curve_point(a, u1, B);
    /* The cone only locates the hit (refine_curve_hit decides): its radii are inflated by a bound on how far the cubic leaves its chord
     * over the piece — du^2 / 8 * max |c''| per component, c'' linear in u — so that every ray that enters the true sweep here also enters
     * the cone and becomes a candidate. */
    float sag = 0.0f;
    for (int c = 0; c < 4; c++) {
        const float s0 = fabsf(6.0f * a[0][c] * u0 + 2.0f * a[1][c]), s1 = fabsf(6.0f * a[0][c] * u1 + 2.0f * a[1][c]);
        sag += fmaxf(s0, s1);
    }
    sag = sag * (1.0f / (8.0f * CURVE_PIECES * CURVE_PIECES)) * 1.0625f;
    A[3] = fabsf(A[3]) + sag; B[3] = fabsf(B[3]) + sag;
}
static inline float dot3f(const float *a, const float *b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }
/* canonical ray / rounded cone (after I. Quilez' intersector; both lateral roots, unnormalised direction, formed about the point
 * of the ray nearest to sphere A).  Same operation order as trace_device.cuh canonical_cone. */
static int canon_cone(const ray_frame *f, float tmin, float tmax, const float A[4], const float B[4], float *t_out, float *s_out) {
    const float ra = A[3], rb = B[3];
    const float *d = f->d;
    const float dd = dot3f(d, d);
    const float ao[3] = {A[0] - f->o[0], A[1] - f->o[1], A[2] - f->o[2]};
    const float t0 = dot3f(ao, d) / dd;
    const float o[3] = {fmaf(t0, d[0], f->o[0]), fmaf(t0, d[1], f->o[1]), fmaf(t0, d[2], f->o[2])};
    const float ba[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    const float oa[3] = {o[0] - A[0], o[1] - A[1], o[2] - A[2]};
    const float ob[3] = {o[0] - B[0], o[1] - B[1], o[2] - B[2]};
    const float rr = ra - rb;
    const float m0 = dot3f(ba, ba), m1 = dot3f(ba, oa), m2 = dot3f(ba, d), m3 = dot3f(d, oa), m5 = dot3f(oa, oa), m6 = dot3f(ob, d), m7 = dot3f(ob, ob);
    const float d2 = fmaf(-rr, rr, m0);
    int found = 0; float tb = 0.f, sb = 0.f;
    if (d2 > 0.0f) {
        const float k2 = fmaf(d2, dd, -(m2 * m2));
        const float k1 = fmaf(d2, m3, fmaf(-m1, m2, (m2 * rr) * ra));
        const float k0 = fmaf(d2, m5, fmaf(-m1, m1, fmaf(m1 * rr, ra * 2.0f, -(m0 * (ra * ra)))));
        const float h = fmaf(k1, k1, -(k0 * k2));
        if (h >= 0.0f && k2 != 0.0f) {
            const float sq = sqrtf(h);
            for (int root = 0; root < 2; root++) {
                const float tl = (root == 0 ? -sq - k1 : sq - k1) / k2;
                const float y = fmaf(tl, m2, fmaf(-ra, rr, m1));
                const float t = tl + t0;
                if (t > tmin && t <= tmax && y > 0.0f && y < d2 && (!found || t < tb)) { found = 1; tb = t; sb = y / d2; }
            }
        }
    }
    const float h1 = fmaf(m3, m3, -(dd * fmaf(-ra, ra, m5)));
    if (h1 >= 0.0f) {
        const float t = (-m3 - sqrtf(h1)) / dd + t0;
        if (t > tmin && t <= tmax && (!found || t < tb)) { found = 1; tb = t; sb = 0.0f; }
    }
    const float h2 = fmaf(m6, m6, -(dd * fmaf(-rb, rb, m7)));
    if (h2 >= 0.0f) {
        const float t = (-m6 - sqrtf(h2)) / dd + t0;
        if (t > tmin && t <= tmax && (!found || t < tb)) { found = 1; tb = t; sb = 1.0f; }
    }
    *t_out = tb; *s_out = sb;
    return found;
}
int oracle_canonical_cone(const float o[3], const float d[3], float tmin, float tmax, const float A[4], const float B[4], float *t, float *s) {
    ray_frame f; memset(&f, 0, sizeof(f));
    memcpy(f.o, o, 12); memcpy(f.d, d, 12);
    return canon_cone(&f, tmin, tmax, A, B, t, s);
}
/* The cone pieces only locate the hit: a cubic segment's surface is the sweep of the sphere (c(u), r(u)), and the hit a piece reports is
 * refined against it — Newton on  F1 = |p - c(u)|^2 - r(u)^2 = 0,  F2 = (p - c(u)).c'(u) + r(u) r'(u) = 0  (the envelope condition), p = o + t d,
 * in double precision with a fixed operation order (trace_device.cuh refine_curve_hit is the same sequence), at most six iterations from
 * the cone hit.  When the iteration does not settle inside the segment, the first / last piece answers with the sphere that closes the
 * segment (exact, in double) if the ray hits it; otherwise — and when the iteration settles on the far side of the sweep — the piece reports
 * no hit (returns 0).  The cone hit itself is never reported for a cubic segment. */
static int curve_end_sphere(const double o[3], const double d[3], float a[4][4], double ue, float *t_io, float *u_io) {
    double c[4];
    for (int k = 0; k < 4; k++) c[k] = (((double)a[0][k] * ue + (double)a[1][k]) * ue + (double)a[2][k]) * ue + (double)a[3][k];
    const double oc[3] = {c[0] - o[0], c[1] - o[1], c[2] - o[2]};
    const double dd = (d[0] * d[0] + d[1] * d[1]) + d[2] * d[2];
    const double b = (oc[0] * d[0] + oc[1] * d[1]) + oc[2] * d[2];
    const double disc = b * b - dd * (((oc[0] * oc[0] + oc[1] * oc[1]) + oc[2] * oc[2]) - c[3] * c[3]);
    if (!(disc >= 0.0)) return 0;
    *t_io = (float)((b - sqrt(disc)) / dd); *u_io = (float)ue;
    return 1;
}
/* end: -1 the piece starts the segment, +1 it ends it, 0 neither (a one-piece segment cannot occur: cubic segments have 8 pieces) */
static int refine_curve_hit(const ray_frame *f, float a[4][4], int end, float *t_io, float *u_io) {
    double t = (double)*t_io, u = (double)*u_io;
    const double o[3] = {f->o[0], f->o[1], f->o[2]}, d[3] = {f->d[0], f->d[1], f->d[2]};
    double step_t = INFINITY, step_u = INFINITY, qd = 0.0;
    int settled = 1;
    for (int it = 0; it < 6; it++) {
        double c[4], c1[4], c2[4];
        for (int k = 0; k < 4; k++) {
            const double a0 = a[0][k], a1 = a[1][k], a2 = a[2][k], a3 = a[3][k];
            c[k] = ((a0 * u + a1) * u + a2) * u + a3;
            c1[k] = ((3.0 * a0) * u + (2.0 * a1)) * u + a2;
            c2[k] = (6.0 * a0) * u + (2.0 * a1);
        }
        const double q[3] = {(o[0] + t * d[0]) - c[0], (o[1] + t * d[1]) - c[1], (o[2] + t * d[2]) - c[2]};
        const double qq = (q[0] * q[0] + q[1] * q[1]) + q[2] * q[2];
        qd = (q[0] * d[0] + q[1] * d[1]) + q[2] * d[2];
        const double qc1 = (q[0] * c1[0] + q[1] * c1[1]) + q[2] * c1[2];
        const double qc2 = (q[0] * c2[0] + q[1] * c2[1]) + q[2] * c2[2];
        const double dc1 = (d[0] * c1[0] + d[1] * c1[1]) + d[2] * c1[2];
        const double c1c1 = (c1[0] * c1[0] + c1[1] * c1[1]) + c1[2] * c1[2];
        const double F1 = qq - c[3] * c[3];
        const double F2 = qc1 + c[3] * c1[3];
        const double J11 = 2.0 * qd, J12 = -2.0 * F2, J21 = dc1;
        const double J22 = ((qc2 - c1c1) + c1[3] * c1[3]) + c[3] * c2[3];
        const double det = J11 * J22 - J12 * J21;
        if (det == 0.0) { settled = 0; break; }
        step_t = (F1 * J22 - J12 * F2) / det;
        step_u = (J11 * F2 - J21 * F1) / det;
        t = t - step_t;
        u = u - step_u;
        if (!(fabs(u) <= 4.0)) { settled = 0; break; }   /* diverging (also catches NaN) */
        if (fabs(step_t) <= 1e-13 * (fabs(t) + 1.0) && fabs(step_u) <= 1e-13) break;
    }
    if (settled && !(fabs(step_t) <= 1e-7 * (fabs(t) + 1.0) && fabs(step_u) <= 1e-7)) settled = 0;
    /* candidates: the envelope point the iteration settled on (inside the segment, entry side) and — on the segment's first / last piece —
     * the sphere that closes the segment (exact, in double), which the ray may enter before it reaches the envelope; the earlier one is
     * the hit.  Neither: the cone's surface only bulged out of the sweep, or the ray grazes. */
    int found = 0;
    float tb = 0.f, ub = 0.f;
    if (settled && u >= 0.0 && u <= 1.0 && qd < 0.0) { found = 1; tb = (float)t; ub = (float)u; }
    if (end != 0) {
        float te, ue;
        if (curve_end_sphere(o, d, a, end < 0 ? 0.0 : 1.0, &te, &ue) && (!found || te < tb)) { found = 1; tb = te; ub = ue; }
    }
    if (!found) return 0;
    *t_io = tb; *u_io = ub;
    return 1;
}

/* every piece of every segment, brute force; ties: lowest (inst, prim), then lowest u */
static void curve_closest(const mesh *m, const ray_frame *f, float tmin, float tmax, uint32_t inst, best_hit *best, const oracle_filter *flt, int first) {
    const uint32_t pieces = m->basis == 0 ? 1 : CURVE_PIECES;
    const float du = 1.0f / (float)pieces;
    for (uint32_t seg = 0; seg < m->nsegs; seg++)
        for (uint32_t k = 0; k < pieces; k++) {
            float A[4], B[4], t, sl;
            curve_piece(m, seg, k, A, B);
            if (!canon_cone(f, tmin, tmax, A, B, &t, &sl)) continue;
            float u = fmaf(sl, du, (float)k * du);
            if (m->basis != 0) {
                const uint32_t first = m->segs[seg];
                float a[4][4];
                curve_basis(m->basis, (const float *)(m->cps + (size_t)first * m->cp_stride), (const float *)(m->cps + (size_t)(first + 1) * m->cp_stride),
                            (const float *)(m->cps + (size_t)(first + 2) * m->cp_stride), (const float *)(m->cps + (size_t)(first + 3) * m->cp_stride), a);
                if (!refine_curve_hit(f, a, k == 0 ? -1 : (k + 1 == pieces ? 1 : 0), &t, &u)) continue;
                if (!(t > tmin && t <= tmax)) continue;
            }
            if (flt && !filter_accept(flt, inst, seg, u, -1.0f)) continue;
            best_hit *b = best;
            if (!b->found || t < b->t || (t == b->t && (inst < b->inst || (inst == b->inst && (seg < b->prim || (seg == b->prim && b->curve && u < b->u)))))) {
                b->found = 1; b->t = t; b->u = u; b->v = -1.0f; b->inst = inst; b->prim = seg; b->curve = 1;
            }
            if (first) return;
        }
}


/* Reported barycentrics: the winning triangle is re-evaluated once in double (Moeller-Trumbore on the
 * canonical object-space ray, fixed operation order, no contraction) and rounded to fp32.  The fp32 edge
 * functions decide hit/miss and ordering; this only tightens the (u, v) that is handed to shading. */
static inline void refine_bary(const ray_frame *f, const float *a, const float *b, const float *c, float *u_io, float *v_io) {
    const double e1x = (double)b[0] - (double)a[0], e1y = (double)b[1] - (double)a[1], e1z = (double)b[2] - (double)a[2];
    const double e2x = (double)c[0] - (double)a[0], e2y = (double)c[1] - (double)a[1], e2z = (double)c[2] - (double)a[2];
    const double sx = (double)f->o[0] - (double)a[0], sy = (double)f->o[1] - (double)a[1], sz = (double)f->o[2] - (double)a[2];
    const double dx = f->d[0], dy = f->d[1], dz = f->d[2];
    const double px = dy * e2z - dz * e2y, py = dz * e2x - dx * e2z, pz = dx * e2y - dy * e2x;
    const double det = (e1x * px + e1y * py) + e1z * pz;
    if (det == 0.0) return;
    const double inv = 1.0 / det;
    const double u = ((sx * px + sy * py) + sz * pz) * inv;
    const double qx = sy * e1z - sz * e1y, qy = sz * e1x - sx * e1z, qz = sx * e1y - sy * e1x;
    const double v = ((dx * qx + dy * qy) + dz * qz) * inv;
    *u_io = (float)u; *v_io = (float)v;
}

/* conservative slab test in double against the box padded by 2^-18 * Linf(ray origin, box) */
static inline void ray_inverse(const float d[3], double inv[3]) { for (int k = 0; k < 3; k++) inv[k] = d[k] == 0.0f ? 0.0 : 1.0 / (double)d[k]; }
/* inv = ray_inverse(d), hoisted out of the node loop (the same double values as dividing per node) */
static inline int box_cull(const float lo[3], const float hi[3], const float o[3], const float d[3], const double rinv[3], double tmin, double tbest,
                           double *tnear_out) {
    double R = 0.0;
    for (int k = 0; k < 3; k++) { double a = fabs((double)lo[k] - o[k]), b = fabs((double)hi[k] - o[k]); if (a > R) R = a; if (b > R) R = b; }
    const double pad = R * (1.0 / 262144.0) + 1e-30;
    double tn = -INFINITY, tf = INFINITY;
    for (int k = 0; k < 3; k++) {
        double l = (double)lo[k] - pad - o[k], h = (double)hi[k] + pad - o[k];
        if (d[k] == 0.0f) { if (l > 0.0 || h < 0.0) return 1; continue; }
        double inv = rinv[k];
        double t0 = l * inv, t1 = h * inv;
        if (t0 > t1) { double s = t0; t0 = t1; t1 = s; }
        if (t0 > tn) tn = t0;
        if (t1 < tf) tf = t1;
    }
    if (tn > tf) return 1;
    if (tf < tmin) return 1;
    if (tn > tbest) return 1;
    *tnear_out = tn;
    return 0;
}

/* the candidate hook of the batch RayQuery (mirrors trace.cu candidate_commits operation by operation) */
static inline int filter_accept(const oracle_filter *flt, uint32_t inst, uint32_t prim, float u, float v) {
    switch (flt->kind) {
        case 0: return 1;
        case 1: {
            const float w = (1.0f - u) - v;
            const float r2 = flt->radius * flt->radius;
            const float xy = fmaf(w, w, u * u), yz = fmaf(u, u, v * v), xz = fmaf(w, w, v * v);
            return xy < r2 && yz < r2 && xz < r2;
        }
        case 2: { uint32_t b = flt->first_bit[inst] + prim; return (flt->bits[b >> 5] >> (b & 31u)) & 1u; }
        case 4: {
            const float fr = flt->stripe_freq[inst];
            if (fr == 0.0f) return 1;
            const float x = v * fr;
            return (x - floorf(x)) < flt->stripe_keep[inst];
        }
        default: return 0;
    }
}

/* flt == NULL: every hit counts (trace_closest); else candidates are filtered (non-opaque instances of a RayQuery).
 * first != 0: return after the first counted hit. */
static void mesh_closest_filtered(const mesh *m, const ray_frame *f, float tmin, float tmax, uint32_t inst, best_hit *best, int mode,
                                  const oracle_filter *flt, int first) {
    if (m->ntris == 0) return;
    if (mode == 0) {
        for (uint32_t p = 0; p < m->ntris; p++) {
            const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
            float t, u, v;
            if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v) && (!flt || filter_accept(flt, inst, p, u, v))) {
                consider(best, t, u, v, inst, p);
                if (first) return;
            }
        }
        return;
    }
    uint32_t stack[128]; int top = 0; stack[top++] = 0;
    double rinv[3]; ray_inverse(f->d, rinv);
    while (top) {
        const bvh_node *nd = &m->nodes[stack[--top]];
        double tn;
        double tb = best->found ? (double)best->t : (double)tmax;
        if (box_cull(nd->lo, nd->hi, f->o, f->d, rinv, tmin, tb, &tn)) continue;
        if (nd->count) {
            for (uint32_t i = 0; i < nd->count; i++) {
                uint32_t p = m->prims[nd->left + i];
                const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
                float t, u, v;
                if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v) && (!flt || filter_accept(flt, inst, p, u, v))) {
                    consider(best, t, u, v, inst, p);
                    if (first) return;
                }
            }
        } else {
            if (top + 2 > 128) die("oracle BVH stack overflow");
            stack[top++] = nd->left; stack[top++] = nd->left + 1;
        }
    }
}

static void mesh_closest_wide(const mesh *m, const ray_frame *f, float tmin, float tmax, uint32_t inst, best_hit *best);
static int mesh_any_wide(const mesh *m, const ray_frame *f, float tmin, float tmax);
static int have_avx2(void);
static void mesh_closest(const mesh *m, const ray_frame *f, float tmin, float tmax, uint32_t inst, best_hit *best, int mode) {
    if (m->ntris == 0) return;
    if (mode == 2 && m->wide && have_avx2()) { mesh_closest_wide(m, f, tmin, tmax, inst, best); return; }
    if (mode == 0) {
        for (uint32_t p = 0; p < m->ntris; p++) {
            const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
            float t, u, v;
            if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v)) consider(best, t, u, v, inst, p);
        }
        return;
    }
    uint32_t stack[128]; int top = 0; stack[top++] = 0;
    double rinv[3]; ray_inverse(f->d, rinv);
    while (top) {
        const bvh_node *nd = &m->nodes[stack[--top]];
        double tn;
        double tb = best->found ? (double)best->t : (double)tmax;
        if (box_cull(nd->lo, nd->hi, f->o, f->d, rinv, tmin, tb, &tn)) continue;
        if (nd->count) {
            for (uint32_t i = 0; i < nd->count; i++) {
                uint32_t p = m->prims[nd->left + i];
                const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
                float t, u, v;
                if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v)) consider(best, t, u, v, inst, p);
            }
        } else {
            if (top + 2 > 128) die("oracle BVH stack overflow");
            /* near child last so it is popped first */
            const bvh_node *l = &m->nodes[nd->left];
            int ax = 0; float e = nd->hi[0] - nd->lo[0];
            if (nd->hi[1] - nd->lo[1] > e) { ax = 1; e = nd->hi[1] - nd->lo[1]; }
            if (nd->hi[2] - nd->lo[2] > e) ax = 2;
            const bvh_node *r = l + 1;
            int left_first = (l->lo[ax] + l->hi[ax] <= r->lo[ax] + r->hi[ax]) == (f->d[ax] >= 0);
            if (left_first) { stack[top++] = nd->left + 1; stack[top++] = nd->left; }
            else { stack[top++] = nd->left; stack[top++] = nd->left + 1; }
        }
    }
}

static int mesh_any(const mesh *m, const ray_frame *f, float tmin, float tmax, int mode) {
    if (m->ntris == 0) return 0;
    if (mode == 2 && m->wide && have_avx2()) return mesh_any_wide(m, f, tmin, tmax);
    float t, u, v;
    if (mode == 0) {
        for (uint32_t p = 0; p < m->ntris; p++) {
            const float *a, *b, *c; tri_verts(m, p, &a, &b, &c);
            if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v)) return 1;
        }
        return 0;
    }
    uint32_t stack[128]; int top = 0; stack[top++] = 0;
    double rinv[3]; ray_inverse(f->d, rinv);
    while (top) {
        const bvh_node *nd = &m->nodes[stack[--top]];
        double tn;
        if (box_cull(nd->lo, nd->hi, f->o, f->d, rinv, tmin, tmax, &tn)) continue;
        if (nd->count) {
            for (uint32_t i = 0; i < nd->count; i++) {
                const float *a, *b, *c; tri_verts(m, m->prims[nd->left + i], &a, &b, &c);
                if (canon_tri(f, tmin, tmax, a, b, c, &t, &u, &v)) return 1;
            }
        } else {
            if (top + 2 > 128) die("oracle BVH stack overflow");
            stack[top++] = nd->left; stack[top++] = nd->left + 1;
        }
    }
    return 0;
}


/* ---- mode 2: the 8-wide tree, box tests eight at a time (AVX2) -----------------------------------------------------------------------
 * The CPU baseline's fast path: same canonical triangle arithmetic (canon_tri), same tie rule (consider), so the hits are the same
 * bits as modes 0 and 1; only the culling differs — single-precision slabs for the eight children of a node at once, each child box
 * padded by 2^-18 of its L-inf distance from the ray origin (32 times the rounding of the fp32 slab arithmetic, so no box that the
 * exact arithmetic would enter is ever skipped), children visited near to far. */
static int g_has_avx2 = -1;
static int have_avx2(void) {
    if (g_has_avx2 < 0) g_has_avx2 = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma") ? 1 : 0;
    return g_has_avx2;
}

typedef struct wide_ray { __m256 o[3], inv[3]; } wide_ray;

__attribute__((target("avx2,fma"))) static inline void wide_ray_setup(const ray_frame *f, wide_ray *w) {
    for (int k = 0; k < 3; k++) {
        w->o[k] = _mm256_set1_ps(f->o[k]);
        /* a zero direction component: a huge finite reciprocal keeps 0 * inv = 0 (no NaN) and sends every other plane to +-1e30 */
        w->inv[k] = _mm256_set1_ps(f->d[k] == 0.0f ? 1e30f : 1.0f / f->d[k]);
    }
}

/* entry distances of the children of `nd` that [tmin, tbest] may hit; returns the lane mask */
__attribute__((target("avx2,fma"))) static inline unsigned wide_node_test(const wnode *nd, const wide_ray *w, float tmin, float tbest, float tn_out[8]) {
    const __m256 sign = _mm256_set1_ps(-0.0f);
    __m256 l[3], h[3], R = _mm256_setzero_ps();
    for (int k = 0; k < 3; k++) {
        l[k] = _mm256_sub_ps(_mm256_load_ps(nd->lo[k]), w->o[k]);
        h[k] = _mm256_sub_ps(_mm256_load_ps(nd->hi[k]), w->o[k]);
        R = _mm256_max_ps(R, _mm256_max_ps(_mm256_andnot_ps(sign, l[k]), _mm256_andnot_ps(sign, h[k])));
    }
    const __m256 pad = _mm256_add_ps(_mm256_mul_ps(R, _mm256_set1_ps(1.0f / 262144.0f)), _mm256_set1_ps(1e-30f));
    __m256 tn = _mm256_set1_ps(tmin), tf = _mm256_set1_ps(tbest);
    for (int k = 0; k < 3; k++) {
        const __m256 t0 = _mm256_mul_ps(_mm256_sub_ps(l[k], pad), w->inv[k]), t1 = _mm256_mul_ps(_mm256_add_ps(h[k], pad), w->inv[k]);
        tn = _mm256_max_ps(tn, _mm256_min_ps(t0, t1));
        tf = _mm256_min_ps(tf, _mm256_max_ps(t0, t1));
    }
    /* widen the comparison by the rounding of the products themselves (a relative 2^-22 of the larger magnitude) */
    const __m256 slack = _mm256_mul_ps(_mm256_max_ps(_mm256_andnot_ps(sign, tn), _mm256_andnot_ps(sign, tf)), _mm256_set1_ps(1.0f / 4194304.0f));
    _mm256_storeu_ps(tn_out, tn);
    return (unsigned)_mm256_movemask_ps(_mm256_cmp_ps(tn, _mm256_add_ps(tf, slack), _CMP_LE_OQ));
}

typedef struct wide_entry { uint32_t node; float tn; } wide_entry;

__attribute__((target("avx2,fma"))) static void mesh_closest_wide(const mesh *m, const ray_frame *f, float tmin, float tmax, uint32_t inst, best_hit *best) {
    wide_ray w; wide_ray_setup(f, &w);
    wide_entry stack[512]; int top = 0;
    stack[top++] = (wide_entry){0, -INFINITY};
    while (top) {
        const wide_entry e = stack[--top];
        const float tb = best->found ? best->t : tmax;
        if (e.tn > tb) continue;
        const wnode *nd = &m->wide[e.node];
        float tn[8];
        unsigned mask = wide_node_test(nd, &w, tmin, tb, tn);
        const int base = top;
        while (mask) {
            const int i = __builtin_ctz(mask); mask &= mask - 1;
            if (nd->count[i] == WIDE_EMPTY) continue;
            if (nd->count[i]) {
                const ptri *t = &m->packed[nd->child[i]];
                for (uint32_t j = 0; j < nd->count[i]; j++) {
                    float tt, u, v;
                    if (canon_tri(f, tmin, tmax, t[j].v0, t[j].v1, t[j].v2, &tt, &u, &v)) consider(best, tt, u, v, inst, t[j].prim);
                }
            } else {
                if (top + 1 > 512) die("oracle wide BVH stack overflow");
                /* insertion keeps the new entries far-to-near, so the nearest child is popped first */
                int k = top++;
                while (k > base && stack[k - 1].tn < tn[i]) { stack[k] = stack[k - 1]; k--; }
                stack[k] = (wide_entry){nd->child[i], tn[i]};
                { const char *pf = (const char *)&m->wide[nd->child[i]]; _mm_prefetch(pf, _MM_HINT_T0); _mm_prefetch(pf + 64, _MM_HINT_T0); _mm_prefetch(pf + 128, _MM_HINT_T0); _mm_prefetch(pf + 192, _MM_HINT_T0); }
            }
        }
    }
}

__attribute__((target("avx2,fma"))) static int mesh_any_wide(const mesh *m, const ray_frame *f, float tmin, float tmax) {
    wide_ray w; wide_ray_setup(f, &w);
    uint32_t stack[512]; int top = 0;
    stack[top++] = 0;
    while (top) {
        const wnode *nd = &m->wide[stack[--top]];
        float tn[8];
        unsigned mask = wide_node_test(nd, &w, tmin, tmax, tn);
        while (mask) {
            const int i = __builtin_ctz(mask); mask &= mask - 1;
            if (nd->count[i] == WIDE_EMPTY) continue;
            if (nd->count[i]) {
                const ptri *t = &m->packed[nd->child[i]];
                for (uint32_t j = 0; j < nd->count[i]; j++) {
                    float tt, u, v;
                    if (canon_tri(f, tmin, tmax, t[j].v0, t[j].v1, t[j].v2, &tt, &u, &v)) return 1;
                }
            } else {
                if (top + 1 > 512) die("oracle wide BVH stack overflow");
                stack[top++] = nd->child[i];
            }
        }
    }
    return 0;
}

/* AccelImpl::trace_closest — accel.rs:449-509 (hit/miss encoding :485-508) */
static void closest_one(const oracle_scene *s, const oracle_ray *r, uint32_t mask, oracle_hit *h, int mode) {
    best_hit best; memset(&best, 0, sizeof(best));
    for (uint32_t i = 0; i < s->n_insts; i++) {
        const instance *in = &s->insts[i];
        if (!in->valid || (in->visible & mask) == 0) continue;
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        if (s->meshes[in->mesh].is_curve) curve_closest(&s->meshes[in->mesh], &f, r->tmin, r->tmax, i, &best, NULL, 0);
        else mesh_closest(&s->meshes[in->mesh], &f, r->tmin, r->tmax, i, &best, mode);
    }
    if (best.found && best.curve) { h->inst = best.inst; h->prim = best.prim; h->u = best.u; h->v = best.v; h->t = best.t; }
    else if (best.found) {
        const instance *in = &s->insts[best.inst];
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        const float *a, *b, *c; tri_verts(&s->meshes[in->mesh], best.prim, &a, &b, &c);
        refine_bary(&f, a, b, c, &best.u, &best.v);
        h->inst = best.inst; h->prim = best.prim; h->u = best.u; h->v = best.v; h->t = best.t;
    }
    else { h->inst = UINT32_MAX; h->prim = UINT32_MAX; h->u = 0.f; h->v = 0.f; h->t = r->tmax; }
    h->pad = 0;
}

/* AccelImpl::trace_any — accel.rs:511-535 */
static uint32_t any_one(const oracle_scene *s, const oracle_ray *r, uint32_t mask, int mode) {
    for (uint32_t i = 0; i < s->n_insts; i++) {
        const instance *in = &s->insts[i];
        if (!in->valid || (in->visible & mask) == 0) continue;
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        if (s->meshes[in->mesh].is_curve) {
            best_hit b; memset(&b, 0, sizeof(b));
            curve_closest(&s->meshes[in->mesh], &f, r->tmin, r->tmax, i, &b, NULL, 1);
            if (b.found) return 1;
            continue;
        }
        if (mesh_any(&s->meshes[in->mesh], &f, r->tmin, r->tmax, mode)) return 1;
    }
    return 0;
}

/* AccelImpl::ray_query — accel.rs:582-800 (triangles) */
static void query_one(const oracle_scene *s, const oracle_ray *r, uint32_t mask, int first, const oracle_filter *flt, oracle_committed_hit *h, int mode) {
    best_hit best; memset(&best, 0, sizeof(best));
    for (uint32_t i = 0; i < s->n_insts && !(first && best.found); i++) {
        const instance *in = &s->insts[i];
        if (!in->valid || (in->visible & mask) == 0) continue;
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        if (s->meshes[in->mesh].is_curve) curve_closest(&s->meshes[in->mesh], &f, r->tmin, r->tmax, i, &best, in->opaque ? NULL : flt, first);
        else mesh_closest_filtered(&s->meshes[in->mesh], &f, r->tmin, r->tmax, i, &best, mode, in->opaque ? NULL : flt, first);
    }
    if (best.found && best.curve) { h->inst = best.inst; h->prim = best.prim; h->u = best.u; h->v = best.v; h->hit_type = 1; h->t = best.t; }
    else if (best.found) {
        const instance *in = &s->insts[best.inst];
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        const float *a, *b, *c; tri_verts(&s->meshes[in->mesh], best.prim, &a, &b, &c);
        refine_bary(&f, a, b, c, &best.u, &best.v);
        h->inst = best.inst; h->prim = best.prim; h->u = best.u; h->v = best.v; h->hit_type = 1; h->t = best.t;
    } else { h->inst = UINT32_MAX; h->prim = UINT32_MAX; h->u = 0.f; h->v = 0.f; h->hit_type = 0; h->t = 0.f; }
}

/* ------------------------------------------------------------------------------------ */
/* double-precision ground truth (Moeller-Trumbore), with ambiguity flags                */
/* ------------------------------------------------------------------------------------ */

typedef struct truth_state {
    double t_best; uint32_t inst, prim; double u, v; int found;
    double t_second;          /* second-best certain hit */
    double t_uncertain;       /* smallest t of a near-edge candidate (hit or near miss) */
    int boundary;             /* some candidate's t within rounding of tmin/tmax */
} truth_state;

static inline void truth_tri(const double o[3], const double d[3], double tmin, double tmax, const float *a, const float *b, const float *c,
                             uint32_t inst, uint32_t prim, truth_state *st) {
    double e1[3], e2[3], p[3], s[3], q[3];
    for (int k = 0; k < 3; k++) { e1[k] = (double)b[k] - a[k]; e2[k] = (double)c[k] - a[k]; s[k] = o[k] - a[k]; }
    p[0] = d[1] * e2[2] - d[2] * e2[1]; p[1] = d[2] * e2[0] - d[0] * e2[2]; p[2] = d[0] * e2[1] - d[1] * e2[0];
    double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
    if (det == 0.0) return;
    double inv = 1.0 / det;
    double u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
    q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
    double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
    double t = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
    double w = 1.0 - u - v;
    double bmin = fmin(u, fmin(v, w));
    const double EB = 1e-4, ET = 1e-5;
    if (bmin < -EB) return;
    double lo = tmin - fabs(tmin) * ET - 1e-30, hi = tmax + fabs(tmax) * ET;
    if (!(t >= lo && t <= hi)) return;
    if (fabs(t - tmin) <= fabs(tmin) * ET + 1e-30 || fabs(t - tmax) <= fabs(tmax) * ET) st->boundary = 1;
    if (bmin < EB) { if (t < st->t_uncertain) st->t_uncertain = t; }
    if (bmin < 0.0 || !(t > tmin && t <= tmax)) return;
    if (!st->found || t < st->t_best || (t == st->t_best && (inst < st->inst || (inst == st->inst && prim < st->prim)))) {
        if (st->found && st->t_best < st->t_second) st->t_second = st->t_best;
        st->found = 1; st->t_best = t; st->inst = inst; st->prim = prim; st->u = u; st->v = v;
    } else if (t < st->t_second) st->t_second = t;
}

static void truth_one(const oracle_scene *s, const oracle_ray *r, uint32_t mask, oracle_hit *h, uint8_t *amb, int mode) {
    truth_state st; memset(&st, 0, sizeof(st)); st.t_second = INFINITY; st.t_uncertain = INFINITY;
    for (uint32_t i = 0; i < s->n_insts; i++) {
        const instance *in = &s->insts[i];
        if (!in->valid || (in->visible & mask) == 0) continue;
        const mesh *m = &s->meshes[in->mesh];
        if (m->ntris == 0) continue;
        double o[3], d[3]; float of[3], df[3];
        for (int k = 0; k < 3; k++) {
            const float *mm = in->inv + 4 * k;
            o[k] = (double)mm[0] * r->o[0] + (double)mm[1] * r->o[1] + (double)mm[2] * r->o[2] + (double)mm[3];
            d[k] = (double)mm[0] * r->d[0] + (double)mm[1] * r->d[1] + (double)mm[2] * r->d[2];
            of[k] = (float)o[k]; df[k] = (float)d[k];
        }
        if (mode == 0) {
            for (uint32_t p = 0; p < m->ntris; p++) { const float *a, *b, *c; tri_verts(m, p, &a, &b, &c); truth_tri(o, d, r->tmin, r->tmax, a, b, c, i, p, &st); }
        } else {
            uint32_t stack[128]; int top = 0; stack[top++] = 0;
            double rinv[3]; ray_inverse(df, rinv);
            while (top) {
                const bvh_node *nd = &m->nodes[stack[--top]];
                double tn, tb = st.found ? st.t_best * (1.0 + 1e-4) + 1e-30 : (double)r->tmax * (1.0 + 1e-4);
                if (box_cull(nd->lo, nd->hi, of, df, rinv, (double)r->tmin * (1.0 - 1e-4) - 1e-30, tb, &tn)) continue;
                if (nd->count) {
                    for (uint32_t j = 0; j < nd->count; j++) { uint32_t p = m->prims[nd->left + j]; const float *a, *b, *c; tri_verts(m, p, &a, &b, &c); truth_tri(o, d, r->tmin, r->tmax, a, b, c, i, p, &st); }
                } else { if (top + 2 > 128) die("oracle BVH stack overflow"); stack[top++] = nd->left; stack[top++] = nd->left + 1; }
            }
        }
    }
    int ambiguous = st.boundary;
    if (st.found) {
        h->inst = st.inst; h->prim = st.prim; h->u = (float)st.u; h->v = (float)st.v; h->t = (float)st.t_best;
        double lim = st.t_best + fabs(st.t_best) * 1e-5 + 1e-30;
        if (st.t_second <= lim) ambiguous = 1;
        if (st.t_uncertain <= lim) ambiguous = 1;
    } else {
        h->inst = UINT32_MAX; h->prim = UINT32_MAX; h->u = 0.f; h->v = 0.f; h->t = r->tmax;
        if (st.t_uncertain < INFINITY) ambiguous = 1;
    }
    h->pad = 0;
    if (amb) *amb = (uint8_t)ambiguous;
}

/* ------------------------------------------------------------------------------------ */
/* mode 2, closest hit, several rays in flight per thread                                 */
/* ------------------------------------------------------------------------------------ */
/* On incoherent rays a wide-BVH walk is a chain of cache misses (the 1 M-triangle soup's tree is 150 MB).  A thread therefore keeps
 * WALKS rays in flight and advances them in turn, one node each: the prefetches a step issues for the children it pushed have the
 * other rays' steps to complete in.  Per ray the sequence of nodes, triangle tests and `consider` calls is exactly the one of
 * closest_one / mesh_closest_wide (same stack discipline), so the hits are the same bits. */
enum { WALKS = 8, WALK_STACK = 256 };
typedef struct walk {
    uint64_t ray; uint32_t inst; int live, in_mesh;
    ray_frame f; wide_ray w; const mesh *m; best_hit best;
    int top; wide_entry stack[WALK_STACK];
    int n_pending; uint64_t pending[8];   /* leaf children hit by the last node test: first packed triangle << 8 | count */
    const ptri *best_tri;                  /* packed copy of the winning triangle (still in cache when the ray is finished) */
} walk;

/* moves the walk to the next instance it has to traverse (curves and scalar-mode meshes are finished on the spot); 0 = ray done */
__attribute__((target("avx2,fma"))) static int walk_next_instance(walk *k, const oracle_scene *s, const oracle_ray *r, uint32_t mask) {
    for (; k->inst < s->n_insts; k->inst++) {
        const instance *in = &s->insts[k->inst];
        if (!in->valid || (in->visible & mask) == 0) continue;
        const mesh *m = &s->meshes[in->mesh];
        xform_ray(in->inv, r->o, r->d, &k->f);
        if (m->is_curve) { curve_closest(m, &k->f, r->tmin, r->tmax, k->inst, &k->best, NULL, 0); continue; }
        if (m->ntris == 0) continue;
        if (!m->wide) { mesh_closest(m, &k->f, r->tmin, r->tmax, k->inst, &k->best, 1); continue; }
        wide_ray_setup(&k->f, &k->w);
        k->m = m; k->top = 0; k->n_pending = 0; k->stack[k->top++] = (wide_entry){0, -INFINITY};
        _mm_prefetch((const char *)&m->wide[0], _MM_HINT_T0);
        k->in_mesh = 1;
        return 1;
    }
    return 0;
}

__attribute__((target("avx2,fma"))) static void walk_finish(walk *k, const oracle_scene *s, const oracle_ray *r, oracle_hit *h) {
    best_hit *best = &k->best;
    if (best->found && best->curve) { h->inst = best->inst; h->prim = best->prim; h->u = best->u; h->v = best->v; h->t = best->t; }
    else if (best->found) {
        const instance *in = &s->insts[best->inst];
        ray_frame f; xform_ray(in->inv, r->o, r->d, &f);
        const float *a, *b, *c;
        /* the winner's vertices: the packed copy the walk tested (the same floats as the mesh's, and still in cache) when it came from
         * a wide tree, else through the index buffer as closest_one does */
        if (k->best_tri && k->best_tri->prim == best->prim && s->meshes[in->mesh].wide) { a = k->best_tri->v0; b = k->best_tri->v1; c = k->best_tri->v2; }
        else tri_verts(&s->meshes[in->mesh], best->prim, &a, &b, &c);
        refine_bary(&f, a, b, c, &best->u, &best->v);
        h->inst = best->inst; h->prim = best->prim; h->u = best->u; h->v = best->v; h->t = best->t;
    }
    else { h->inst = UINT32_MAX; h->prim = UINT32_MAX; h->u = 0.f; h->v = 0.f; h->t = r->tmax; }
    h->pad = 0;
}

/* one node of one ray: the body of mesh_closest_wide's loop, except that the triangles of the leaf children a node test hits are only
 * PREFETCHED here and tested at the ray's next step (they are cache misses too).  The hit is the same: the set of triangles tested can
 * only grow (a later node is culled against a slightly older tbest), and `consider` does not depend on the order of its calls. */
__attribute__((target("avx2,fma"))) static inline void walk_step(walk *k, const oracle_ray *r) {
    const mesh *m = k->m;
    for (int p = 0; p < k->n_pending; p++) {
        const ptri *t = &m->packed[k->pending[p] >> 8];
        const uint32_t cnt = k->pending[p] & 0xffu;
        for (uint32_t j = 0; j < cnt; j++) {
            float tt, u, v;
            if (canon_tri(&k->f, r->tmin, r->tmax, t[j].v0, t[j].v1, t[j].v2, &tt, &u, &v)) {
                const int had = k->best.found; const uint32_t pi = k->best.inst, pp = k->best.prim;
                consider(&k->best, tt, u, v, k->inst, t[j].prim);
                if (!had || k->best.inst != pi || k->best.prim != pp) k->best_tri = &t[j];
            }
        }
    }
    k->n_pending = 0;
    if (k->top == 0) return;
    const wide_entry e = k->stack[--k->top];
    const float tb = k->best.found ? k->best.t : r->tmax;
    if (e.tn > tb) return;
    const wnode *nd = &m->wide[e.node];
    float tn[8];
    unsigned hit = wide_node_test(nd, &k->w, r->tmin, tb, tn);
    const int base = k->top;
    while (hit) {
        const int i = __builtin_ctz(hit); hit &= hit - 1;
        if (nd->count[i] == WIDE_EMPTY) continue;
        if (nd->count[i]) {
            const char *pf = (const char *)&m->packed[nd->child[i]];
            const char *pe = pf + (size_t)nd->count[i] * sizeof(ptri);
            for (; pf < pe; pf += 64) _mm_prefetch(pf, _MM_HINT_T0);
            _mm_prefetch(pe - 1, _MM_HINT_T0);
            k->pending[k->n_pending++] = ((uint64_t)nd->child[i] << 8) | nd->count[i];
        } else {
            if (k->top + 1 > WALK_STACK) die("oracle wide BVH stack overflow");
            int q = k->top++;
            while (q > base && k->stack[q - 1].tn < tn[i]) { k->stack[q] = k->stack[q - 1]; q--; }
            k->stack[q] = (wide_entry){nd->child[i], tn[i]};
            { const char *pf = (const char *)&m->wide[nd->child[i]]; _mm_prefetch(pf, _MM_HINT_T0); _mm_prefetch(pf + 64, _MM_HINT_T0); _mm_prefetch(pf + 128, _MM_HINT_T0); _mm_prefetch(pf + 192, _MM_HINT_T0); }
        }
    }
}

__attribute__((target("avx2,fma"))) static void closest_block_interleaved(const oracle_scene *s, const oracle_ray *rays, uint64_t i0, uint64_t i1, uint32_t mask, oracle_hit *hits) {
    static __thread walk *ws = NULL;
    if (!ws && posix_memalign((void **)&ws, 64, sizeof(walk) * WALKS) != 0) die("out of memory");
    uint64_t next = i0; int live = 0;
    for (int a = 0; a < WALKS; a++) ws[a].live = 0;
    for (;;) {
        for (int a = 0; a < WALKS; a++) {
            walk *k = &ws[a];
            if (!k->live) {
                if (next >= i1) continue;
                k->ray = next++; k->inst = 0; k->live = 1; k->in_mesh = 0; k->best_tri = NULL; memset(&k->best, 0, sizeof(k->best)); live++;
            }
            const oracle_ray *r = &rays[k->ray];
            if (k->in_mesh && k->top == 0 && k->n_pending == 0) { k->in_mesh = 0; k->inst++; }   /* this instance is exhausted */
            if (!k->in_mesh && !walk_next_instance(k, s, r, mask)) { walk_finish(k, s, r, &hits[k->ray]); k->live = 0; live--; continue; }
            walk_step(k, r);
        }
        if (live == 0 && next >= i1) break;
    }
}

/* ------------------------------------------------------------------------------------ */
/* StreamImpl::parallel_for — stream.rs:185-209: N workers, atomic counter, block = 64    */
/* ------------------------------------------------------------------------------------ */

typedef struct job {
    const oracle_scene *s; const oracle_ray *rays; uint64_t n; uint32_t mask; int mode; int kind;
    oracle_hit *hits; uint32_t *occ; uint8_t *amb;
    uint64_t counter;
    const oracle_filter *flt; oracle_committed_hit *committed; int first;
    const struct pt_job *pt;
} job;
static void pt_pixel(const struct pt_job *pt, const oracle_scene *s, uint64_t pixel);

static void *worker(void *arg) {
    job *j = (job *)arg;
    for (;;) {
        uint64_t i0 = __atomic_fetch_add(&j->counter, 64, __ATOMIC_RELAXED);
        if (i0 >= j->n) break;
        uint64_t i1 = i0 + 64 < j->n ? i0 + 64 : j->n;
        if (j->kind == 0 && j->mode == 2 && have_avx2()) { closest_block_interleaved(j->s, j->rays, i0, i1, j->mask, j->hits); continue; }
        for (uint64_t i = i0; i < i1; i++) {
            if (j->kind == 0) closest_one(j->s, &j->rays[i], j->mask, &j->hits[i], j->mode);
            else if (j->kind == 1) j->occ[i] = any_one(j->s, &j->rays[i], j->mask, j->mode);
            else if (j->kind == 2) truth_one(j->s, &j->rays[i], j->mask, &j->hits[i], j->amb ? &j->amb[i] : NULL, j->mode);
            else if (j->kind == 3) query_one(j->s, &j->rays[i], j->mask, j->first, j->flt, &j->committed[i], j->mode);
            else pt_pixel(j->pt, j->s, i);
        }
    }
    return NULL;
}

int oracle_hw_threads(void) { long n = sysconf(_SC_NPROCESSORS_ONLN); return n > 0 ? (int)n : 1; }

static void run(job *j, int threads) {
    if (threads <= 0) threads = oracle_hw_threads();
    if (threads > 256) threads = 256;
    if (threads == 1 || j->n <= 64) { worker(j); return; }
    pthread_t th[256];
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

void oracle_trace_closest(const oracle_scene *s, const oracle_ray *rays, uint64_t n, uint32_t mask, oracle_hit *hits, int mode, int threads) {
    job j = {s, rays, n, mask, mode, 0, hits, NULL, NULL, 0, NULL, NULL, 0, NULL}; run(&j, threads);
}
void oracle_trace_any(const oracle_scene *s, const oracle_ray *rays, uint64_t n, uint32_t mask, uint32_t *occ, int mode, int threads) {
    job j = {s, rays, n, mask, mode, 1, NULL, occ, NULL, 0, NULL, NULL, 0, NULL}; run(&j, threads);
}
void oracle_trace_closest_f64(const oracle_scene *s, const oracle_ray *rays, uint64_t n, uint32_t mask, oracle_hit *hits, uint8_t *amb, int mode, int threads) {
    job j = {s, rays, n, mask, mode, 2, hits, NULL, amb, 0, NULL, NULL, 0, NULL}; run(&j, threads);
}
void oracle_ray_query(const oracle_scene *s, const oracle_ray *rays, uint64_t n, uint32_t mask, int terminate_on_first, const oracle_filter *filter,
                      oracle_committed_hit *out, int mode, int threads) {
    static const oracle_filter commit_all = {0, 0.f, NULL, NULL, NULL, NULL};
    job j = {s, rays, n, mask, mode, 3, NULL, NULL, NULL, 0, filter ? filter : &commit_all, out, terminate_on_first, NULL}; run(&j, threads);
}

/* ------------------------------------------------------------------------------------ */
/* offset_ray_origin — lc/src/rtx.rs:517-535                                             */
/* ------------------------------------------------------------------------------------ */

void oracle_offset_ray_origin(const float p[3], const float n[3], float out[3]) {
    const float origin = 1.0f / 32.0f, float_scale = 1.0f / 65536.0f, int_scale = 256.0f;
    for (int k = 0; k < 3; k++) {
        int32_t of_i = (int32_t)(int_scale * n[k]);
        int32_t bits; memcpy(&bits, &p[k], 4);
        int32_t p_i = bits + (p[k] < 0.0f ? -of_i : of_i);
        float pf; memcpy(&pf, &p_i, 4);
        out[k] = fabsf(p[k]) < origin ? p[k] + float_scale * n[k] : pf;
    }
}

/* ------------------------------------------------------------------------------------ */
/* examples/path_tracer.rs:247-455 — the caller of the path on config C2                 */
/* ------------------------------------------------------------------------------------ */

typedef struct v3 { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 vadd(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 vdiv(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float vlength(v3 a) { return sqrtf(vdot(a, a)); }
/* lc_normalize(v) = v * rsqrt(dot(v, v)) with rsqrt(x) = 1 / sqrt(x): cpu/codegen/device_math.h:3588, cpu_prelude.h:7 */
static inline v3 vnormalize(v3 a) { return vscale(a, 1.0f / sqrtf(vdot(a, a))); }
static inline v3 voffset(v3 p, v3 n) { float pi[3] = {p.x, p.y, p.z}, ni[3] = {n.x, n.y, n.z}, o[3]; oracle_offset_ray_origin(pi, ni, o); return V(o[0], o[1], o[2]); }

static inline void sincos_2pi(float u, float *s, float *c) {
    const float kf = floorf(u * 4.0f + 0.5f);
    const float r = u - kf * 0.25f;
    const float x = r * 6.28318530717958647692f;
    const float x2 = x * x;
    float sp = fmaf(x2, 2.7557319e-6f, -1.9841270e-4f);
    sp = fmaf(sp, x2, 8.3333333e-3f);
    sp = fmaf(sp, x2, -1.6666667e-1f);
    sp = fmaf(sp * x2, x, x);
    float cp = fmaf(x2, 2.4801587e-5f, -1.3888889e-3f);
    cp = fmaf(cp, x2, 4.1666667e-2f);
    cp = fmaf(cp, x2, -0.5f);
    cp = fmaf(cp, x2, 1.0f);
    const int k = (int)kf & 3;
    *s = k == 0 ? sp : k == 1 ? cp : k == 2 ? -sp : -cp;
    *c = k == 0 ? cp : k == 1 ? -sp : k == 2 ? -cp : sp;
}

static inline float lcg(uint32_t *state) {
    *state = 1664525u * *state + 1013904223u;
    return (float)(*state & 0x00ffffffu) * (1.0f / 16777216.0f);
}

struct pt_job {
    const float *const *vertex_heap; const uint32_t *const *index_heap; float *image; uint32_t *seeds;
    uint32_t width, height, spp, max_depth; float tan_half_fov; int mode;
    uint64_t n_closest, n_any;
    const oracle_filter *flt;  /* non-NULL: path_tracer_cutout.rs (ray queries with a candidate filter) */
};

static void pt_pixel(const struct pt_job *ptc, const oracle_scene *s, uint64_t pixel) {
    struct pt_job *pt = (struct pt_job *)ptc;
    static const float mats[8][3] = {{0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.140f, 0.450f, 0.091f},
                                     {0.630f, 0.065f, 0.050f}, {0.725f, 0.710f, 0.680f}, {0.725f, 0.710f, 0.680f}, {0.000f, 0.000f, 0.000f}};
    const float FRAC_1_PI = 0.318309886183790671537767526745028724f, F32_MAX = 3.40282347e+38f;
    const uint32_t cx = (uint32_t)(pixel % pt->width), cy = (uint32_t)(pixel / pt->width);
    const float frame_size = (float)(pt->width < pt->height ? pt->width : pt->height);
    uint32_t state = pt->seeds[pixel];
    const float rx = lcg(&state), ry = lcg(&state);
    const float px = ((float)cx + rx) / frame_size * 2.0f - 1.0f, py = ((float)cy + ry) / frame_size * 2.0f - 1.0f;
    v3 radiance = V(0.f, 0.f, 0.f);
    uint64_t n_closest = 0, n_any = 0;
    const v3 light_position = V(-0.24f, 1.98f, 0.16f);
    const v3 light_u = vsub(V(-0.24f, 1.98f, -0.22f), light_position), light_v = vsub(V(0.23f, 1.98f, 0.16f), light_position);
    const v3 light_emission = V(17.0f, 12.0f, 4.0f);
    const float light_area = vlength(vcross(light_u, light_v));
    const v3 light_normal = vnormalize(vcross(light_u, light_v));
    for (uint32_t sample = 0; sample < pt->spp; sample++) {
        const v3 cam = V(-0.01f, 0.995f, 5.0f);
        const v3 pix = vadd(cam, V(px * 1.0f * pt->tan_half_fov, py * -1.0f * pt->tan_half_fov, -1.0f));
        v3 ray_o = cam, ray_d = vnormalize(vsub(pix, cam));
        float ray_tmin = 0.0f, ray_tmax = F32_MAX;
        v3 beta = V(1.f, 1.f, 1.f);
        float pdf_bsdf = 0.0f;
        uint32_t depth = 0;
        while (depth < pt->max_depth) {
            oracle_ray r = {{ray_o.x, ray_o.y, ray_o.z}, ray_tmin, {ray_d.x, ray_d.y, ray_d.z}, ray_tmax};
            oracle_hit hit;
            if (pt->flt) {
                oracle_committed_hit ch;
                query_one(s, &r, 0xffu, 0, pt->flt, &ch, pt->mode);
                hit.inst = ch.inst; hit.prim = ch.prim; hit.u = ch.u; hit.v = ch.v; hit.t = ch.t;
            } else closest_one(s, &r, 0xffu, &hit, pt->mode);
            n_closest++;
            if (hit.inst == UINT32_MAX) break;
            const float *vb = pt->vertex_heap[hit.inst];
            const uint32_t *tri = pt->index_heap[hit.inst] + 3 * (size_t)hit.prim;
            const uint32_t i0 = tri[0], i1 = tri[1], i2 = tri[2];
            const v3 p0 = V(vb[3 * i0], vb[3 * i0 + 1], vb[3 * i0 + 2]), p1 = V(vb[3 * i1], vb[3 * i1 + 1], vb[3 * i1 + 2]), p2 = V(vb[3 * i2], vb[3 * i2 + 1], vb[3 * i2 + 2]);
            const v3 p = vadd(vadd(vscale(p0, (1.0f - hit.u) - hit.v), vscale(p1, hit.u)), vscale(p2, hit.v));
            const v3 n = vnormalize(vcross(vsub(p1, p0), vsub(p2, p0)));
            const float cos_wi = -vdot(ray_d, n);
            if (cos_wi < 1e-4f) break;
            const v3 pp = voffset(p, n);
            const v3 albedo = V(mats[hit.inst & 7u][0], mats[hit.inst & 7u][1], mats[hit.inst & 7u][2]);
            if (hit.inst == 7u) {
                if (depth == 0u) radiance = vadd(radiance, light_emission);
                else {
                    const v3 d = vsub(p, ray_o);
                    const float pdf_light = vdot(d, d) / (light_area * cos_wi);
                    const float mis_weight = pdf_bsdf / fmaxf(pdf_bsdf + pdf_light, 1e-4f);
                    radiance = vadd(radiance, vmul(V(mis_weight * beta.x, mis_weight * beta.y, mis_weight * beta.z), light_emission));
                }
                break;
            } else {
                const float ux_light = lcg(&state), uy_light = lcg(&state);
                const v3 p_light = vadd(vadd(light_position, V(ux_light * light_u.x, ux_light * light_u.y, ux_light * light_u.z)),
                                        V(uy_light * light_v.x, uy_light * light_v.y, uy_light * light_v.z));
                const v3 pp_light = voffset(p_light, light_normal);
                const float d_light = vlength(vsub(pp, pp_light));
                const v3 wi_light = vnormalize(vsub(pp_light, pp));
                const v3 so = voffset(pp, n);
                oracle_ray sr = {{so.x, so.y, so.z}, 0.0f, {wi_light.x, wi_light.y, wi_light.z}, d_light};
                int occluded;
                if (pt->flt) { oracle_committed_hit ch; query_one(s, &sr, 0xffu, 1, pt->flt, &ch, pt->mode); occluded = ch.hit_type != 0u; }
                else occluded = (int)any_one(s, &sr, 0xffu, pt->mode);
                n_any++;
                const float cos_wi_light = vdot(wi_light, n);
                const float cos_light = -vdot(light_normal, wi_light);
                if (!occluded && cos_wi_light > 1e-4f && cos_light > 1e-4f) {
                    const float pdf_light = (d_light * d_light) / (light_area * cos_light);
                    const float pdf_b = cos_wi_light * FRAC_1_PI;
                    const float mis_weight = pdf_light / fmaxf(pdf_light + pdf_b, 1e-4f);
                    const v3 bsdf = vscale(vscale(albedo, FRAC_1_PI), cos_wi_light);
                    radiance = vadd(radiance, vdiv(vmul(vscale(vmul(beta, bsdf), mis_weight), light_emission), fmaxf(pdf_light, 1e-4f)));
                }
            }
            const v3 binormal = fabsf(n.x) > fabsf(n.z) ? V(-n.y, n.x, 0.0f) : V(0.0f, -n.z, n.y);
            const v3 tangent = vnormalize(vcross(binormal, n));
            const float ux = lcg(&state), uy = lcg(&state);
            const float rr0 = sqrtf(ux);
            float sphi, cphi;
            sincos_2pi(uy, &sphi, &cphi);
            const v3 local = V(rr0 * cphi, rr0 * sphi, sqrtf(1.0f - ux));
            const v3 new_direction = vadd(vadd(vscale(tangent, local.x), vscale(binormal, local.y)), vscale(n, local.z));
            ray_o = pp; ray_d = new_direction; ray_tmin = 0.0f; ray_tmax = F32_MAX;
            beta = vmul(beta, albedo);
            pdf_bsdf = cos_wi * FRAC_1_PI;
            const float l = vdot(V(0.212671f, 0.715160f, 0.072169f), beta);
            if (l == 0.0f) break;
            const float q = fmaxf(l, 0.05f);
            const float rr = lcg(&state);
            if (rr > q) break;
            beta = vdiv(beta, q);
            depth += 1;
        }
    }
    radiance = vdiv(radiance, (float)pt->spp);
    pt->seeds[pixel] = state;
    if (isnan(radiance.x) || isnan(radiance.y) || isnan(radiance.z)) radiance = V(0.f, 0.f, 0.f);
    radiance = V(fminf(fmaxf(radiance.x, 0.0f), 30.0f), fminf(fmaxf(radiance.y, 0.0f), 30.0f), fminf(fmaxf(radiance.z, 0.0f), 30.0f));
    float *px4 = pt->image + 4 * pixel;
    px4[0] = radiance.x + px4[0]; px4[1] = radiance.y + px4[1]; px4[2] = radiance.z + px4[2]; px4[3] = px4[3] + 1.0f;
    __atomic_fetch_add(&pt->n_closest, n_closest, __ATOMIC_RELAXED);
    __atomic_fetch_add(&pt->n_any, n_any, __ATOMIC_RELAXED);
}

void oracle_path_tracer(const oracle_scene *s, const float *const *vertex_heap, const uint32_t *const *index_heap, float *image_rgba, uint32_t *seed_image,
                        uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov, int threads,
                        uint64_t ray_counts_out[2]) {
    oracle_path_tracer_cutout(s, vertex_heap, index_heap, image_rgba, seed_image, width, height, spp_per_dispatch, max_depth, tan_half_fov, NULL, threads, ray_counts_out);
}

void oracle_path_tracer_cutout(const oracle_scene *s, const float *const *vertex_heap, const uint32_t *const *index_heap, float *image_rgba, uint32_t *seed_image,
                               uint32_t width, uint32_t height, uint32_t spp_per_dispatch, uint32_t max_depth, float tan_half_fov, const oracle_filter *filter,
                               int threads, uint64_t ray_counts_out[2]) {
    struct pt_job pt = {vertex_heap, index_heap, image_rgba, seed_image, width, height, spp_per_dispatch, max_depth, tan_half_fov, 1, 0, 0, filter};
    job j = {s, NULL, (uint64_t)width * height, 0xff, 1, 4, NULL, NULL, NULL, 0, NULL, NULL, 0, &pt};
    run(&j, threads);
    if (ray_counts_out) { ray_counts_out[0] = pt.n_closest; ray_counts_out[1] = pt.n_any; }
}
