// symbolic.cpp -- nested dissection, elimination tree, column counts, supernodes and their row
// structures for the 25-point (spatial) / 43-point (space-time) mesh pattern.  Host C++ only;
// runs once per mesh and never touches the GPU.
//
// Algorithms: geometric nested dissection on the box (separator thickness 2 in x,y and 1 in t,
// forced by the 5x5 / 3x3 coupling); Liu's elimination tree with path
// compression; Gilbert-Ng-Peyton column counts (skeleton / least-common-ancestor form); maximal
// supernodes with CHOLMOD-style relaxed amalgamation of a supernode with its last child;
// row structures by a bottom-up merge over the supernodal tree.
#include "symbolic.h"

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <numeric>

namespace spde {

int Symbolic::slot_nbr(int node, int slot) const { return geo.slot_nbr(node, slot); }

namespace {
struct Box { int x0, x1, y0, y1, t0, t1; };
}

void Symbolic::nested_dissection(int leaf)
{
    const int M = geo.M, N = geo.N, T = geo.T, Ns = M * N;
    perm.clear();
    perm.reserve(n);
    auto emit = [&](const Box &b) {
        for (int t = b.t0; t < b.t1; t++)
            for (int y = b.y0; y < b.y1; y++)
                for (int x = b.x0; x < b.x1; x++) perm.push_back(t * Ns + y * M + x);
    };
    std::function<void(const Box &)> nd = [&](const Box &b) {
        const int64_t lx = b.x1 - b.x0, ly = b.y1 - b.y0, lt = b.t1 - b.t0;
        const int64_t vol = lx * ly * lt;
        if (vol <= 0) return;
        int best = -1;
        int64_t bestcost = 0, bestlen = 0;
        if (vol > leaf) {
            const int64_t len[3] = {lx, ly, lt};
            const int thick[3] = {2, 2, 1};
            for (int d = 0; d < 3; d++) {
                if (len[d] < thick[d] + 2) continue;
                const int64_t cost = thick[d] * (vol / len[d]);
                if (best < 0 || cost < bestcost || (cost == bestcost && len[d] > bestlen)) {
                    best = d; bestcost = cost; bestlen = len[d];
                }
            }
        }
        if (best < 0) { emit(b); return; }
        Box l = b, r = b, s = b;
        if (best == 0) { int mid = b.x0 + (int)(lx - 2) / 2; l.x1 = mid; s.x0 = mid; s.x1 = mid + 2; r.x0 = mid + 2; }
        else if (best == 1) { int mid = b.y0 + (int)(ly - 2) / 2; l.y1 = mid; s.y0 = mid; s.y1 = mid + 2; r.y0 = mid + 2; }
        else { int mid = b.t0 + (int)(lt - 1) / 2; l.t1 = mid; s.t0 = mid; s.t1 = mid + 1; r.t0 = mid + 1; }
        nd(l);
        nd(r);
        emit(s);
    };
    if (geo.bc == 2) {
        // periodic in x and y: the wrap couples the two ends, so the strips x<2 and y<2 go last
        nd(Box{2, M, 2, N, 0, T});
        emit(Box{0, 2, 2, N, 0, T});
        emit(Box{0, M, 0, 2, 0, T});
    } else {
        nd(Box{0, M, 0, N, 0, T});
    }
    iperm.assign(n, -1);
    for (int k = 0; k < n; k++) iperm[perm[k]] = k;
}

void Symbolic::etree_postorder()
{
    // Liu's algorithm on the permuted pattern, neighbours enumerated from the mesh geometry
    std::vector<int> par(n, -1), anc(n, -1);
    for (int j = 0; j < n; j++) {
        const int node = perm[j];
        for (int s = 0; s < nslots; s++) {
            const int c = slot_nbr(node, s);
            if (c < 0) continue;
            int r = iperm[c];
            if (r >= j) continue;
            while (anc[r] != -1 && anc[r] != j) { const int nx = anc[r]; anc[r] = j; r = nx; }
            if (anc[r] == -1) { anc[r] = j; par[r] = j; }
        }
    }
    // postorder (children visited in increasing index, so the separator structure is kept)
    std::vector<int> head(n, -1), next(n, -1), post;
    post.reserve(n);
    for (int j = n - 1; j >= 0; j--)
        if (par[j] >= 0) { next[j] = head[par[j]]; head[par[j]] = j; }
    std::vector<int> stack;
    for (int root = 0; root < n; root++) {
        if (par[root] != -1) continue;
        stack.push_back(root);
        while (!stack.empty()) {
            const int p = stack.back();
            const int c = head[p];
            if (c == -1) { post.push_back(p); stack.pop_back(); }
            else { head[p] = next[c]; stack.push_back(c); }
        }
    }
    std::vector<int> ipost(n);
    for (int k = 0; k < n; k++) ipost[post[k]] = k;
    std::vector<int> nperm(n);
    parent.assign(n, -1);
    for (int k = 0; k < n; k++) {
        nperm[k] = perm[post[k]];
        parent[k] = par[post[k]] < 0 ? -1 : ipost[par[post[k]]];
    }
    perm.swap(nperm);
    for (int k = 0; k < n; k++) iperm[perm[k]] = k;
}

void Symbolic::column_counts()
{
    // Gilbert-Ng-Peyton; the ordering is already a postorder, so post[k] = k.
    std::vector<int> delta(n), anc(n), maxfirst(n, -1), prevleaf(n, -1), firstd(n, -1);
    for (int k = 0; k < n; k++) {
        int j = k;
        delta[j] = (firstd[j] == -1) ? 1 : 0;
        for (; j != -1 && firstd[j] == -1; j = parent[j]) firstd[j] = k;
    }
    std::iota(anc.begin(), anc.end(), 0);
    for (int j = 0; j < n; j++) {
        if (parent[j] != -1) delta[parent[j]]--;
        const int node = perm[j];
        for (int s = 0; s < nslots; s++) {
            const int c = slot_nbr(node, s);
            if (c < 0) continue;
            const int i = iperm[c];
            if (i <= j || firstd[j] <= maxfirst[i]) continue;
            maxfirst[i] = firstd[j];
            const int jprev = prevleaf[i];
            prevleaf[i] = j;
            if (jprev == -1) { delta[j]++; continue; }
            int q = jprev;
            while (q != anc[q]) q = anc[q];
            for (int sx = jprev; sx != q;) { const int sp = anc[sx]; anc[sx] = q; sx = sp; }
            delta[j]++;
            delta[q]--;
        }
        if (parent[j] != -1) anc[j] = parent[j];
    }
    colcount = delta;
    for (int j = 0; j < n; j++)
        if (parent[j] != -1) colcount[parent[j]] += colcount[j];
    flops = 0;
    nnzL = 0;
    for (int j = 0; j < n; j++) { flops += (double)colcount[j] * colcount[j]; nnzL += colcount[j]; }
}

void Symbolic::supernodes()
{
    // maximal supernodes: j+1 joins j when parent[j]==j+1 and |L(:,j+1)| == |L(:,j)|-1
    std::vector<int> start;
    start.push_back(0);
    for (int j = 0; j + 1 < n; j++)
        if (!(parent[j] == j + 1 && colcount[j + 1] == colcount[j] - 1)) start.push_back(j + 1);
    int ns = (int)start.size();
    start.push_back(n);
    // relaxed amalgamation with the last child (the child whose columns end right before ours)
    const char *env = getenv("SPDE_RELAX");
    const double zscale = env ? atof(env) : 1.0;
    std::vector<int> sfirst(start.begin(), start.begin() + ns), slast(ns);
    for (int s = 0; s < ns; s++) slast[s] = start[s + 1] - 1;
    std::vector<double> zeros(ns, 0.0);
    std::vector<char> dead(ns, 0);
    // supernode index of each column (before merging)
    std::vector<int> sof(n);
    for (int s = 0; s < ns; s++) for (int j = sfirst[s]; j <= slast[s]; j++) sof[j] = s;
    std::vector<int> live;   // stack of live supernodes seen so far
    for (int s = 0; s < ns; s++) {
      while (!live.empty()) {
        // candidate child: the live supernode that ends at sfirst[s]-1
        const int c = live.back();
        if (slast[c] != sfirst[s] - 1) break;
        if (parent[slast[c]] != sfirst[s]) break;
        const double ncc = slast[c] - sfirst[c] + 1, ncs = slast[s] - sfirst[s] + 1;
        const double mc = colcount[sfirst[c]], ms = colcount[sfirst[s]];
        // entries of the merged trapezoid vs what the two hold now
        const double nc = ncc + ncs, m = ncc + ms;
        const double total = nc * m - nc * (nc - 1) / 2;
        const double have = (ncc * mc - ncc * (ncc - 1) / 2 - zeros[c]) + (ncs * ms - ncs * (ncs - 1) / 2 - zeros[s]);
        const double z = total - have;
        const double frac = z / total;
        bool merge;
        if (nc <= 8) merge = true;
        else if (nc <= 32) merge = frac < 0.5 * zscale;
        else if (nc <= 96) merge = frac < 0.15 * zscale;
        else merge = frac < 0.03 * zscale;
        if (z <= 0) merge = true;
        if (!merge) break;
        // merged supernode keeps index s
        sfirst[s] = sfirst[c];
        zeros[s] = z;
        colcount[sfirst[s]] = (int)m;   // leading column count of the merged (relaxed) supernode
        dead[c] = 1;
        live.pop_back();
      }
      live.push_back(s);
    }
    first.clear();
    for (int s = 0; s < ns; s++) if (!dead[s]) first.push_back(sfirst[s]);
    nsuper = (int)first.size();
    first.push_back(n);
    snode_of.assign(n, 0);
    for (int s = 0; s < nsuper; s++) for (int j = first[s]; j < first[s + 1]; j++) snode_of[j] = s;
}

void Symbolic::structures()
{
    sparent.assign(nsuper, -1);
    rowptr.assign(nsuper + 1, 0);
    rows.clear();
    std::vector<std::vector<int>> kids(nsuper);
    std::vector<int> mark(n, -1);
    std::vector<int> tmp;
    // supernodes are in postorder: children precede parents
    std::vector<int64_t> rp(nsuper + 1, 0);
    std::vector<std::vector<int>> st(nsuper);   // freed as soon as the parent has consumed them
    for (int s = 0; s < nsuper; s++) {
        const int lo = first[s], hi = first[s + 1];   // columns [lo,hi)
        tmp.clear();
        for (int j = lo; j < hi; j++) {
            const int node = perm[j];
            for (int q = 0; q < nslots; q++) {
                const int c = slot_nbr(node, q);
                if (c < 0) continue;
                const int i = iperm[c];
                if (i >= hi && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
            }
        }
        for (int c : kids[s]) {
            for (int i : st[c])
                if (i >= hi && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
            std::vector<int>().swap(st[c]);
        }
        std::sort(tmp.begin(), tmp.end());
        st[s] = tmp;
        rp[s + 1] = rp[s] + (int64_t)tmp.size();
        rows.insert(rows.end(), tmp.begin(), tmp.end());
        if (!tmp.empty()) {
            sparent[s] = snode_of[tmp[0]];
            kids[sparent[s]].push_back(s);
        }
    }
    rowptr = rp;
    // depth and relative indices
    depth.assign(nsuper, 0);
    maxdepth = 0;
    for (int s = nsuper - 1; s >= 0; s--) {
        depth[s] = sparent[s] < 0 ? 0 : depth[sparent[s]] + 1;
        maxdepth = std::max(maxdepth, depth[s]);
    }
    relidx.assign(rows.size(), -1);
    for (int s = 0; s < nsuper; s++) {
        const int p = sparent[s];
        if (p < 0) continue;
        const int pf = first[p], pl = first[p + 1];
        const int ncp = pl - pf;
        const int *prow = rows.data() + rowptr[p];
        const int64_t pnr = rowptr[p + 1] - rowptr[p];
        int64_t cursor = 0;
        for (int64_t e = rowptr[s]; e < rowptr[s + 1]; e++) {
            const int i = rows[e];
            if (i < pl) { relidx[e] = i - pf; continue; }
            while (cursor < pnr && prow[cursor] < i) cursor++;
            relidx[e] = ncp + (int)cursor;   // by construction prow[cursor] == i
        }
    }
}

void Symbolic::analyse(const Geo &g, int leaf)
{
    geo = g;
    n = g.M * g.N * g.T;
    nslots = g.nslots();
    if (leaf <= 0) {
        const char *env = getenv("SPDE_ND_LEAF");
        leaf = env ? atoi(env) : (g.T == 1 ? 32 : 64);
    }
    nested_dissection(leaf);
    etree_postorder();
    column_counts();
    supernodes();
    structures();
}

}  // namespace spde
