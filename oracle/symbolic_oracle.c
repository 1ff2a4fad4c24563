/* symbolic_oracle.c -- symbolic analysis for the oracle's CPU Cholesky.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference factorises with scikit-sparse 0.4.12 -> SuiteSparse CHOLMOD (poetry.lock:582-583; call sites
 * advection_diffusion2D.py:117,193), which is absent from /root/reference and from this image.  This file restates
 * the *published* symbolic phase of CHOLMOD's supernodal method (Chen, Davis, Hager, Rajamanickam, ACM TOMS 35(3),
 * 2008; Davis, "Direct Methods for Sparse Linear Systems", SIAM 2006, ch. 4) on a GENERAL symmetric pattern:
 *
 *   1. elimination tree (Liu's algorithm with path compression),
 *   2. postorder of the tree,
 *   3. column counts of L by marking the row subtrees (cs_ereach-style traversal, O(nnz(L))),
 *   4. maximal supernodes + CHOLMOD's relaxed amalgamation with its documented defaults
 *      (cholmod_common: nrelax = {4, 16, 48}, zrelax = {0.8, 0.1, 0.05}),
 *   5. row structure of every supernode by a bottom-up merge over the supernodal tree.
 *
 * It is independent of the product's own analysis (spdepy_b200/csrc/symbolic.cpp works on the mesh geometry, this
 * works on the CSC pattern of the permuted matrix) and is used by bench.py's `--impl reference` / cpu_baseline legs
 * and by tests/ only.  Nothing under spdepy_b200/ links or loads it.
 *
 * Input: the FULL symmetric pattern (both triangles, diagonal optional) of the already permuted matrix in CSC form.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n, nsuper;
    int *post;      /* post[k] = column of the input matrix that becomes column k */
    int *first;     /* nsuper+1 */
    int64_t *rowptr; /* nsuper+1 */
    int *rows;      /* rows below the pivot block of every supernode, ascending, in the FINAL numbering */
    int *sparent;   /* nsuper */
    int64_t nrows;
    double nnzL, flops;
} OSym;

static int cmp_int(const void *a, const void *b)
{
    const int x = *(const int *)a, y = *(const int *)b;
    return (x > y) - (x < y);
}

void osym_free(OSym *s)
{
    if (!s) return;
    free(s->post); free(s->first); free(s->rowptr); free(s->rows); free(s->sparent);
    free(s);
}

OSym *osym_analyse(int n, const int64_t *Ap, const int32_t *Ai, int relax)
{
    OSym *S = (OSym *)calloc(1, sizeof(OSym));
    S->n = n;
    int *parent = (int *)malloc(sizeof(int) * n), *anc = (int *)malloc(sizeof(int) * n);
    /* 1. elimination tree */
    for (int j = 0; j < n; j++) {
        parent[j] = -1; anc[j] = -1;
        for (int64_t p = Ap[j]; p < Ap[j + 1]; p++) {
            int r = Ai[p];
            if (r >= j) continue;
            while (anc[r] != -1 && anc[r] != j) { const int nx = anc[r]; anc[r] = j; r = nx; }
            if (anc[r] == -1) { anc[r] = j; parent[r] = j; }
        }
    }
    /* 2. postorder (children in increasing order) */
    int *head = (int *)malloc(sizeof(int) * n), *next = (int *)malloc(sizeof(int) * n);
    int *post = (int *)malloc(sizeof(int) * n), *stack = (int *)malloc(sizeof(int) * n);
    for (int j = 0; j < n; j++) head[j] = -1;
    for (int j = n - 1; j >= 0; j--)
        if (parent[j] >= 0) { next[j] = head[parent[j]]; head[parent[j]] = j; }
    int np = 0;
    for (int root = 0; root < n; root++) {
        if (parent[root] != -1) continue;
        int top = 0;
        stack[top++] = root;
        while (top) {
            const int p = stack[top - 1], c = head[p];
            if (c == -1) { post[np++] = p; top--; }
            else { head[p] = next[c]; stack[top++] = c; }
        }
    }
    int *ipost = anc;   /* reuse */
    for (int k = 0; k < n; k++) ipost[post[k]] = k;
    int *par2 = head;   /* reuse: parent in the postordered numbering */
    for (int k = 0; k < n; k++) par2[k] = parent[post[k]] < 0 ? -1 : ipost[parent[post[k]]];
    S->post = post;
    /* 3. column counts: row subtrees of the postordered matrix.  Row i of L has an entry in every column on the
     *    paths from the k with A(i,k) != 0, k < i, up the tree until a column already marked for row i. */
    int *cc = next;     /* reuse */
    int *mark = stack;  /* reuse */
    for (int j = 0; j < n; j++) { cc[j] = 1; mark[j] = -1; }
    for (int i = 0; i < n; i++) {
        mark[i] = i;
        const int oi = post[i];
        for (int64_t p = Ap[oi]; p < Ap[oi + 1]; p++) {
            int k = ipost[Ai[p]];
            if (k >= i) continue;
            while (mark[k] != i) { cc[k]++; mark[k] = i; k = par2[k]; }
        }
    }
    double nnzL = 0, flops = 0;
    for (int j = 0; j < n; j++) { nnzL += cc[j]; flops += (double)cc[j] * cc[j]; }
    S->nnzL = nnzL; S->flops = flops;
    /* 4. maximal supernodes, then relaxed amalgamation of a supernode with the child that ends right before it */
    int *sfirst = (int *)malloc(sizeof(int) * (n + 1));
    int ns = 0;
    sfirst[ns++] = 0;
    for (int j = 0; j + 1 < n; j++)
        if (!(par2[j] == j + 1 && cc[j + 1] == cc[j] - 1)) sfirst[ns++] = j + 1;
    sfirst[ns] = n;
    int *slast = (int *)malloc(sizeof(int) * ns), *sf = (int *)malloc(sizeof(int) * ns), *live = (int *)malloc(sizeof(int) * ns);
    double *zeros = (double *)calloc(ns, sizeof(double));
    char *dead = (char *)calloc(ns, 1);
    int *lead = (int *)malloc(sizeof(int) * ns);    /* column count of the leading column (after merging) */
    for (int s = 0; s < ns; s++) { sf[s] = sfirst[s]; slast[s] = sfirst[s + 1] - 1; lead[s] = cc[sfirst[s]]; }
    int nl = 0;
    for (int s = 0; s < ns; s++) {
        while (relax && nl) {
            const int c = live[nl - 1];
            if (slast[c] != sf[s] - 1 || par2[slast[c]] != sf[s]) break;
            const double ncc = slast[c] - sf[c] + 1, ncs = slast[s] - sf[s] + 1;
            const double mc = lead[c], ms = lead[s];
            const double nc = ncc + ncs, m = ncc + ms;
            const double total = nc * m - nc * (nc - 1) / 2;
            const double have = (ncc * mc - ncc * (ncc - 1) / 2 - zeros[c]) + (ncs * ms - ncs * (ncs - 1) / 2 - zeros[s]);
            const double z = total - have, frac = z / total;
            int merge;
            if (nc <= 4) merge = 1;
            else if (nc <= 16) merge = frac < 0.8;
            else if (nc <= 48) merge = frac < 0.1;
            else merge = frac < 0.05;
            if (z <= 0) merge = 1;
            if (!merge) break;
            sf[s] = sf[c];
            zeros[s] = z;
            lead[s] = (int)m;
            dead[c] = 1;
            nl--;
        }
        live[nl++] = s;
    }
    int nsuper = 0;
    for (int s = 0; s < ns; s++) if (!dead[s]) nsuper++;
    S->nsuper = nsuper;
    S->first = (int *)malloc(sizeof(int) * (nsuper + 1));
    { int q = 0; for (int s = 0; s < ns; s++) if (!dead[s]) S->first[q++] = sf[s]; S->first[q] = n; }
    free(sfirst); free(slast); free(sf); free(live); free(zeros); free(dead); free(lead);
    int *snode_of = (int *)malloc(sizeof(int) * n);
    for (int s = 0; s < nsuper; s++) for (int j = S->first[s]; j < S->first[s + 1]; j++) snode_of[j] = s;
    /* 5. row structures, children merged into parents (supernodes are in postorder) */
    S->rowptr = (int64_t *)calloc(nsuper + 1, sizeof(int64_t));
    S->sparent = (int *)malloc(sizeof(int) * nsuper);
    int *khead = (int *)malloc(sizeof(int) * nsuper), *knext = (int *)malloc(sizeof(int) * nsuper);
    for (int s = 0; s < nsuper; s++) { khead[s] = -1; knext[s] = -1; S->sparent[s] = -1; }
    int64_t cap = (int64_t)n * 4 + 1024, used = 0;
    int *rows = (int *)malloc(sizeof(int) * cap);
    int *tmp = (int *)malloc(sizeof(int) * n);
    for (int j = 0; j < n; j++) mark[j] = -1;
    for (int s = 0; s < nsuper; s++) {
        const int lo = S->first[s], hi = S->first[s + 1];
        int cnt = 0;
        for (int j = lo; j < hi; j++) {
            const int oj = post[j];
            for (int64_t p = Ap[oj]; p < Ap[oj + 1]; p++) {
                const int i = ipost[Ai[p]];
                if (i >= hi && mark[i] != s) { mark[i] = s; tmp[cnt++] = i; }
            }
        }
        for (int c = khead[s]; c != -1; c = knext[c])
            for (int64_t e = S->rowptr[c]; e < S->rowptr[c + 1]; e++) {
                const int i = rows[e];
                if (i >= hi && mark[i] != s) { mark[i] = s; tmp[cnt++] = i; }
            }
        qsort(tmp, cnt, sizeof(int), cmp_int);
        if (used + cnt > cap) { while (used + cnt > cap) cap *= 2; rows = (int *)realloc(rows, sizeof(int) * cap); }
        memcpy(rows + used, tmp, sizeof(int) * cnt);
        S->rowptr[s] = used;
        used += cnt;
        S->rowptr[s + 1] = used;
        if (cnt) {
            const int p = snode_of[tmp[0]];
            S->sparent[s] = p;
            knext[s] = khead[p];
            khead[p] = s;
        }
    }
    S->rows = rows;
    S->nrows = used;
    free(tmp); free(khead); free(knext); free(snode_of);
    free(parent); free(anc); free(head); free(next); free(stack);
    return S;
}

int osym_nsuper(const OSym *s) { return s->nsuper; }
int64_t osym_nrows(const OSym *s) { return s->nrows; }
double osym_nnzL(const OSym *s) { return s->nnzL; }
double osym_flops(const OSym *s) { return s->flops; }
void osym_export(const OSym *s, int32_t *post, int32_t *first, int64_t *rowptr, int32_t *rows, int32_t *sparent)
{
    memcpy(post, s->post, sizeof(int) * s->n);
    memcpy(first, s->first, sizeof(int) * (s->nsuper + 1));
    memcpy(rowptr, s->rowptr, sizeof(int64_t) * (s->nsuper + 1));
    memcpy(rows, s->rows, sizeof(int) * s->nrows);
    memcpy(sparent, s->sparent, sizeof(int) * s->nsuper);
}

/* Extend-add of one child update matrix (lower triangle of the n x n column-major `src`, leading dimension lds) into the
 * parent's frontal matrix held as three column-major blocks: F11 (nc x nc), F21 (nr x nc), F22 (nr x nr).  rel[i] is the
 * position of the child's row i in the parent's front (ascending).  The dense kernel of the multifrontal method's
 * assembly step (Duff & Reid 1983); NumPy's fancy indexing needs three passes over temporaries for the same thing.
 * Single-threaded on purpose: the BLAS pool's workers spin between calls and would fight an OpenMP team for the cores. */
void osym_extend_add(const double *src, int64_t lds, int n, const int32_t *rel, int nc,
                     double *F11, double *F21, int64_t nr, double *F22)
{
    for (int j = 0; j < n; j++) {
        const int rj = rel[j];
        const double *s = src + (int64_t)j * lds;
        if (rj < nc) {
            double *d11 = F11 + (int64_t)rj * nc, *d21 = F21 + (int64_t)rj * nr;
            for (int i = j; i < n; i++) {
                const int ri = rel[i];
                if (ri < nc) d11[ri] += s[i]; else d21[ri - nc] += s[i];
            }
        } else {
            double *d22 = F22 + (int64_t)(rj - nc) * nr;
            for (int i = j; i < n; i++) d22[rel[i] - nc] += s[i];
        }
    }
}
