// ooc.cu -- streamed evaluation (see ooc.h): segmentation of the supernodal tree, static memory plan,
// per-segment schedules, executor and C ABI.
#include <algorithm>
#include <cstring>
#include <functional>
#include <numeric>

#include "ooc.h"
#include "plan_steps.h"

namespace spde {

namespace {

constexpr int64_t ALIGN = 32;    // pool regions start on 256-byte boundaries
inline int64_t R(int64_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

__global__ void k_scatter_seg(const double *__restrict__ Q, const ScatEntry *__restrict__ ent, long long cnt, int n,
                              int diag_slot, const double *__restrict__ obs_cnt, double tau, double *__restrict__ pool)
{
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < cnt; e += (long long)gridDim.x * blockDim.x) {
        const ScatEntry s = ent[e];
        double v = Q[s.src];
        if (obs_cnt) {
            const long long slot = s.src / n;
            if (slot == diag_slot) v += obs_cnt[s.src - slot * n] * tau;
        }
        pool[s.dst] = v;
    }
}

size_t program_bytes(const Program &P)
{
    auto a = [](size_t b) { return (b + 255) / 256 * 256; };
    return a(P.gemm.size() * sizeof(GemmTask)) + a(P.tiles.size() * sizeof(TileRef)) + a(P.potrf.size() * sizeof(PotrfTask)) +
           a(P.ext.size() * sizeof(ExtTask)) + a(P.gather.size() * sizeof(GatherTask)) + a(P.wtw.size() * sizeof(WtwTask));
}

struct Stage {
    char *base;
    size_t cap, used = 0;
    cudaStream_t st;
    template <class T>
    int put(const T *h, size_t count, T **d)
    {
        *d = nullptr;
        if (!count) return SPDE_OK;
        const size_t bytes = count * sizeof(T);
        used = (used + 255) / 256 * 256;
        if (used + bytes > cap) { set_error("streamed evaluation: staging buffer too small"); return SPDE_ERR_ARG; }
        SPDE_CUDA_CHECK(cudaMemcpyAsync(base + used, h, bytes, cudaMemcpyHostToDevice, st));
        *d = reinterpret_cast<T *>(base + used);
        used += bytes;
        return SPDE_OK;
    }
    int program(Program &P)
    {
        int rc;
        if ((rc = put(P.gemm.data(), P.gemm.size(), &P.d_gemm))) return rc;
        if ((rc = put(P.tiles.data(), P.tiles.size(), &P.d_tiles))) return rc;
        if ((rc = put(P.potrf.data(), P.potrf.size(), &P.d_potrf))) return rc;
        if ((rc = put(P.ext.data(), P.ext.size(), &P.d_ext))) return rc;
        if ((rc = put(P.gather.data(), P.gather.size(), &P.d_gather))) return rc;
        if ((rc = put(P.wtw.data(), P.wtw.size(), &P.d_wtw))) return rc;
        P.uploaded = true;
        return SPDE_OK;
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
void Ooc::segment(int64_t top_bytes)
{
    const Plan &p = *plan;
    const Symbolic &S = p.sym;
    const int ns = S.nsuper;
    std::vector<int64_t> sub(ns);
    std::vector<int> cnt(ns, 1);
    for (int s = 0; s < ns; s++) sub[s] = (int64_t)p.sn[s].ld * p.sn[s].nc * 8;
    for (int s = 0; s < ns; s++)
        if (S.sparent[s] >= 0) { sub[S.sparent[s]] += sub[s]; cnt[S.sparent[s]] += cnt[s]; }
    seg_of.assign(ns, -1);
    segs.clear();
    for (int s = 0; s < ns; s++) {
        const bool top = sub[s] > top_bytes;
        const int par = S.sparent[s];
        const bool broot = !top && (par < 0 || sub[par] > top_bytes);
        if (!top && !broot) continue;
        OocSeg g;
        g.root = s;
        g.top = top;
        if (top) g.nodes = {s};
        else
            for (int v = s - cnt[s] + 1; v <= s; v++) g.nodes.push_back(v);
        for (int v : g.nodes) seg_of[v] = (int)segs.size();
        segs.push_back(std::move(g));
    }
    for (size_t i = 0; i < segs.size(); i++) {
        OocSeg &g = segs[i];
        const int par = S.sparent[g.root];
        g.parent_seg = par < 0 ? -1 : seg_of[par];
        g.dmin = S.depth[g.root];
        g.dmax = g.dmin;
        for (int v : g.nodes) g.dmax = std::max(g.dmax, S.depth[v]);
        g.by_depth.assign(g.dmax - g.dmin + 1, {});
        for (int v : g.nodes) g.by_depth[S.depth[v] - g.dmin].push_back(v);
        g.col0 = S.first[g.nodes.front()];
        g.col1 = S.first[g.root + 1];
        for (int j = g.col0; j < g.col1; j++) g.flops += (double)S.colcount[j] * S.colcount[j];
    }
    for (size_t i = 0; i < segs.size(); i++)
        if (segs[i].parent_seg >= 0) segs[segs[i].parent_seg].kids.push_back((int)i);
    // local layout of every segment (offsets relative to the segment's regions, made absolute by plan_memory)
    osn = p.sn;
    // two-level Takahashi recursion: the outer-block inverses of a front live behind its Yt scratch in the Y region
    const char *envt = getenv("SPDE_SELINV_OUTER");
    const bool sel_outer = envt ? atoi(envt) != 0 : true;
    yoff_of.assign(osn.size(), 0);
    for (OocSeg &g : segs) {
        const int nd = g.dmax - g.dmin + 1;
        std::vector<int64_t> upd_used(nd, 0), front_used(nd, 0), y_used(nd, 0);
        for (int s : g.nodes) {
            SNode &x = osn[s];
            const int d = x.depth - g.dmin;
            x.panel = g.l_size;
            g.l_size += (int64_t)x.ld * x.nc;
            g.l_size += g.l_size & 1;
            x.dinv = g.dinv_size;
            g.dinv_size += (int64_t)x.nblk * NB * NB;
            x.upd = upd_used[d];
            upd_used[d] += (int64_t)x.ldu * x.nr;
            x.front = front_used[d];
            front_used[d] += (int64_t)x.ld * x.ld;
            yoff_of[s] = y_used[d];
            if (sel_outer && x.nblk > 1) {
                x.winv = y_used[d] + ybuf_need(x);
                y_used[d] += ybuf_need(x) + winv_size(x);
                y_used[d] += y_used[d] & 1;
            } else {
                x.winv = -1;
                y_used[d] += (int64_t)x.ld * NB;
            }
        }
        for (int d = 0; d < nd; d++) {
            const int par = (d + g.dmin) & 1;
            g.arena[par] = std::max(g.arena[par], upd_used[d]);
            g.zarena[par] = std::max(g.zarena[par], front_used[d]);
            g.ybuf = std::max(g.ybuf, y_used[d]);
        }
        g.u_size = (int64_t)p.sn[g.root].ldu * p.sn[g.root].nr;
    }
}

// ---------------------------------------------------------------------------------------------
void Ooc::plan_memory(bool backward)
{
    const int m = (int)segs.size();
    auto fws = [&](const OocSeg &g) { return R(g.l_size) + R(g.dinv_size) + R(g.arena[0]) + R(g.arena[1]); };
    auto bws = [&](const OocSeg &g) {
        if (g.top) return R(g.l_size) + R(g.dinv_size) + R(g.zarena[0]) + R(g.zarena[1]) + R(g.ybuf);
        return fws(g) + R(g.zarena[0]) + R(g.zarena[1]) + R(g.ybuf);
    };
    // Liu's child order: segments are numbered in postorder, so children precede parents
    std::vector<int64_t> P(m, 0);
    for (int i = 0; i < m; i++) {
        OocSeg &g = segs[i];
        std::stable_sort(g.kids.begin(), g.kids.end(), [&](int a, int b) { return P[a] - R(segs[a].u_size) > P[b] - R(segs[b].u_size); });
        int64_t below = 0, pk = 0;
        for (int c : g.kids) { pk = std::max(pk, below + P[c]); below += R(segs[c].u_size); }
        pk = std::max(pk, below + fws(g));
        pk = std::max(pk, R(g.u_size) + fws(g));
        P[i] = pk;
    }
    order.clear();
    std::function<void(int)> dfs = [&](int i) {
        for (int c : segs[i].kids) dfs(c);
        order.push_back(i);
    };
    for (int i = 0; i < m; i++)
        if (segs[i].parent_seg < 0) dfs(i);
    // forward simulation
    int64_t top = 0;
    peak_fwd = 0;
    for (int i : order) {
        OocSeg &g = segs[i];
        peak_fwd = std::max(peak_fwd, top + fws(g));
        for (int c : g.kids) top -= R(segs[c].u_size);
        g.stack_U = top;
        top += R(g.u_size);
        peak_fwd = std::max(peak_fwd, top + fws(g));
    }
    // backward simulation (exact reverse order)
    top = 0;
    peak_bwd = 0;
    for (int q = (int)order.size() - 1; q >= 0; q--) {
        OocSeg &g = segs[order[q]];
        peak_bwd = std::max(peak_bwd, top + bws(g));
        if (g.parent_seg >= 0) top -= R(g.u_size);
        for (int c : g.kids) { segs[c].stack_Z = top; top += R(segs[c].u_size); }
        peak_bwd = std::max(peak_bwd, top + bws(g));
    }
    pool_size = R(std::max(peak_fwd, backward ? peak_bwd : (int64_t)0));
    host_size = 0;
    recompute_flops = 0;
    for (int i = 0; i < m; i++) {
        OocSeg &g = segs[i];
        g.off_L = pool_size - R(g.l_size);
        g.off_dinv = g.off_L - R(g.dinv_size);
        g.off_ar[0] = g.off_dinv - R(g.arena[0]);
        g.off_ar[1] = g.off_ar[0] - R(g.arena[1]);
        const int64_t zb = g.top ? g.off_dinv : g.off_ar[1];      // a top segment needs no arenas in the backward pass
        g.off_z[0] = zb - R(g.zarena[0]);
        g.off_z[1] = g.off_z[0] - R(g.zarena[1]);
        g.off_y = g.off_z[1] - R(g.ybuf);
        g.keep = backward && !order.empty() && i == order.back();
        g.host_off = -1;
        if (backward && g.top && !g.keep) { g.host_off = host_size; host_size += pool_size - g.off_dinv; }
        if (backward && !g.top && !g.keep) recompute_flops += g.flops;
    }
}

// ---------------------------------------------------------------------------------------------
void Ooc::build_tables()
{
    const Plan &p = *plan;
    const Symbolic &S = p.sym;
    const int n = S.n;
    // absolute node offsets
    for (OocSeg &g : segs)
        for (int s : g.nodes) {
            SNode &x = osn[s];
            x.panel += g.off_L;
            x.dinv += g.off_dinv;
            x.upd += g.off_ar[x.depth & 1];
            x.front += g.off_z[x.depth & 1];
            yoff_of[s] += g.off_y;
            if (x.winv >= 0) x.winv += g.off_y;
        }
    diagpos.resize(n);
    for (int j = 0; j < n; j++) {
        const int s = S.snode_of[j];
        diagpos[j] = osn[s].panel + (int64_t)(j - osn[s].first) * (osn[s].ld + 1);
    }
    // scatter entries grouped by segment (counting sort on the in-core scatter map)
    const size_t ncand = p.cand_slots.size();
    const int m = (int)segs.size();
    std::vector<int64_t> cnt(m + 1, 0);
    auto owner = [&](long long d) {       // supernode whose panel holds in-core L offset d
        int lo = 0, hi = S.nsuper - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (p.sn[mid].panel <= d) lo = mid; else hi = mid - 1;
        }
        return lo;
    };
    std::vector<int> own(p.qdest.size(), -1);
    for (size_t e = 0; e < p.qdest.size(); e++) {
        if (p.qdest[e] < 0) continue;
        own[e] = owner(p.qdest[e]);
        cnt[seg_of[own[e]] + 1]++;
    }
    for (int i = 0; i < m; i++) { cnt[i + 1] += cnt[i]; segs[i].scat0 = cnt[i]; segs[i].scat1 = cnt[i + 1]; }
    scat.resize(cnt[m]);
    {
        std::vector<int64_t> cur(cnt.begin(), cnt.end() - 1);
        for (size_t ci = 0; ci < ncand; ci++)
            for (int r = 0; r < n; r++) {
                const size_t e = ci * n + r;
                if (own[e] < 0) continue;
                const int s = own[e];
                ScatEntry x;
                x.src = (long long)p.cand_slots[ci] * n + r;
                x.dst = osn[s].panel + (p.qdest[e] - p.sn[s].panel);
                scat[cur[seg_of[s]]++] = x;
            }
    }
    // extraction entries grouped by (segment, depth)
    zptr.assign(m, {});
    std::vector<int64_t> base(m + 1, 0);
    for (int i = 0; i < m; i++) {
        zptr[i].assign(segs[i].dmax - segs[i].dmin + 2, 0);
    }
    for (const ZEntry &z : p.zentries) zptr[seg_of[z.sn]][S.depth[z.sn] - segs[seg_of[z.sn]].dmin + 1]++;
    int64_t run = 0;
    for (int i = 0; i < m; i++) {
        segs[i].zent0 = run;
        int64_t acc = 0;
        for (size_t d = 0; d + 1 < zptr[i].size(); d++) { const int64_t c = zptr[i][d + 1]; zptr[i][d] = acc; acc += c; }
        zptr[i].back() = acc;
        run += acc;
        segs[i].zent1 = run;
    }
    zent.resize(run);
    {
        std::vector<std::vector<int64_t>> cur(m);
        for (int i = 0; i < m; i++) cur[i].assign(zptr[i].begin(), zptr[i].end() - 1);
        for (const ZEntry &z0 : p.zentries) {
            const int i = seg_of[z0.sn];
            ZEntry z = z0;
            z.src = osn[z0.sn].front + (z0.src - p.sn[z0.sn].front);
            zent[segs[i].zent0 + cur[i][S.depth[z0.sn] - segs[i].dmin]++] = z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
void Ooc::build_programs()
{
    const Plan &p = *plan;
    const Symbolic &S = p.sym;
    std::vector<std::vector<int>> kids(S.nsuper);
    for (int s = 0; s < S.nsuper; s++)
        if (S.sparent[s] >= 0) kids[S.sparent[s]].push_back(s);
    const int OUTER = env_int("SPDE_FACTOR_OUTER", spde::OUTER, 1);
    const int splitk_min = env_int("SPDE_SPLITK_MIN", 2048, 8);
    const char *envk = getenv("SPDE_SELINV_KCHUNK");
    const int kchunk = envk ? atoi(envk) : 1024;
    const int kchunk2 = env_int("SPDE_SELINV_KCHUNK2", 0, 0);
    int max_nr = 2;
    for (const SNode &x : osn) max_nr = std::max(max_nr, x.nr);
    ident_base = 2 * (int64_t)S.rows.size();
    const bool overlap_on = env_int("SPDE_OOC_OVERLAP", 1, 0) != 0;
    for (size_t gi = 0; gi < segs.size(); gi++) {
        OocSeg &g = segs[gi];
        g.overlap = overlap_on && g.top && g.host_off >= 0;
        g.chunk_off.clear();
        g.chunk_len.clear();
        // ---- factorisation: levels bottom-up
        {
            Program &P = g.factor;
            for (int d = g.dmax; d >= g.dmin; d--) {
                const std::vector<int> &lev = g.by_depth[d - g.dmin];
                const int sp_u = SP_AR0 + (d & 1), sp_child = SP_AR0 + ((d + 1) & 1);
                int64_t used = 0;
                for (int s : lev) used = std::max(used, osn[s].upd + (int64_t)osn[s].ldu * osn[s].nr);
                zero_launch(P, sp_u, g.off_ar[d & 1], used);
                size_t maxk = 0;
                for (int s : lev) maxk = std::max(maxk, kids[s].size());
                for (size_t r = 0; r < maxk; r++) {
                    Launch L;
                    memset(&L, 0, sizeof L);
                    L.kind = LK_EXTADD;
                    L.task0 = (int64_t)P.ext.size();
                    L.tile0 = (int64_t)P.tiles.size();
                    for (int s : lev) {
                        if (kids[s].size() <= r) continue;
                        const int c = kids[s][r];
                        const SNode &cx = osn[c];
                        // a child of another segment left its update matrix on the stack
                        const long long src = seg_of[c] == (int)gi ? cx.upd : segs[seg_of[c]].stack_U;
                        push_ext_task(P, L, src, cx.ldu, cx.nr, cx.rows, p.rel_base, osn[s], sp_child, sp_u);
                    }
                    L.ntasks = (int)(P.ext.size() - L.task0);
                    L.ntiles = (int)(P.tiles.size() - L.tile0);
                    if (L.ntiles) P.launches.push_back(L);
                }
                LevelBuilder B(P);
                for (int s : lev) {
                    std::vector<Step> q;
                    if (g.overlap) {
                        // a spilled top segment (one front): every outer block of columns goes to the host as soon
                        // as it holds its final values, on the copy stream, under the right-looking update that follows
                        const SNode &x = osn[s];
                        factor_node_steps(B, x, sp_u, OUTER, q, [&](std::vector<Step> &qq, int P0, int P1) {
                            const int c0 = P0 * NB, c1 = std::min(P1 * NB, x.nc);
                            const int64_t off = x.panel + (int64_t)c0 * x.ld, len = (int64_t)(c1 - c0) * x.ld;
                            g.chunk_off.push_back(off);
                            g.chunk_len.push_back(len);
                            qq.push_back(copy_step(0, off, len, g.host_off + (off - g.off_dinv)));
                        });
                        q.push_back(copy_step(0, g.off_dinv, g.dinv_size, g.host_off));
                    } else {
                        factor_node_steps(B, osn[s], sp_u, OUTER, q);
                    }
                    B.seq.push_back(std::move(q));
                }
                B.flush();
            }
        }
        // ---- Takahashi: levels top-down
        {
            Program &P = g.selinv;
            for (int d = g.dmin; d <= g.dmax; d++) {
                const std::vector<int> &lev = g.by_depth[d - g.dmin];
                const int sp_z = SP_Z0 + (d & 1), sp_par = SP_Z0 + ((d + 1) & 1);
                int64_t used = 0;
                for (int s : lev) used = std::max(used, osn[s].front + (int64_t)osn[s].ld * osn[s].ld);
                zero_launch(P, sp_z, g.off_z[d & 1], used);
                {
                    Launch L;
                    memset(&L, 0, sizeof L);
                    L.kind = LK_GATHER;
                    L.task0 = (int64_t)P.gather.size();
                    L.tile0 = (int64_t)P.tiles.size();
                    for (int s : lev) {
                        const SNode &x = osn[s];
                        if (x.parent < 0 || x.nr == 0) continue;
                        if (s == g.root) {
                            // compact Z_RR left on the stack by the parent's segment (identity relative indices)
                            push_gather_task(P, L, x.front, x.ld, x.ncp, x.nr, g.stack_Z, x.ldu, 0, 0, ident_base, sp_par, sp_z);
                        } else {
                            const SNode &q = osn[x.parent];
                            push_gather_task(P, L, x.front, x.ld, x.ncp, x.nr, q.front, q.ld, q.nc, q.ncp,
                                             p.rel_base + x.rows, sp_par, sp_z);
                        }
                    }
                    L.ntasks = (int)(P.gather.size() - L.task0);
                    L.ntiles = (int)(P.tiles.size() - L.tile0);
                    if (L.ntiles) P.launches.push_back(L);
                }
                // seeds of the diagonal blocks: W^T W of the single-block fronts and the outer-block inverses Wf / Wf^T Wf of
                // the others are hoisted in front of the level's recursion -- except for a front whose panel is still
                // arriving from the host, which builds them outer block by outer block behind the wait records
                {
                    int64_t yend = g.off_y;
                    for (int s : lev) yend = std::max(yend, yoff_of[s] + (osn[s].winv >= 0 ? ybuf_need(osn[s]) + winv_size(osn[s]) : (int64_t)osn[s].ld * NB));
                    bool any_multi = false;
                    for (int s : lev) any_multi |= osn[s].winv >= 0;
                    if (any_multi) zero_launch(P, SP_Y, g.off_y, yend);      // (the inverses are lower triangular: upper blocks stay zero)
                }
                if (!g.overlap) {
                    std::vector<const SNode *> single, multi;
                    std::vector<int64_t> ymulti;
                    for (int s : lev) {
                        if (osn[s].winv >= 0) { multi.push_back(&osn[s]); ymulti.push_back(yoff_of[s]); }
                        else single.push_back(&osn[s]);
                    }
                    wtw_level_launch(P, single, sp_z);
                    if (!multi.empty()) winv_level_launches(P, multi, ymulti, sp_z);
                }
                LevelBuilder B(P);
                for (int s : lev) {
                    std::vector<Step> q;
                    const SNode &x = osn[s];
                    const int nblk = x.nblk;
                    if (g.overlap && x.winv >= 0) {
                        // outer blocks are visited last to first; the panel comes back in slices of OUTER block columns, last
                        // slice first on one stream: the slice holding the first column of the outer block is the last one it needs
                        selinv_node_steps_outer(B, x, sp_z, yoff_of[s], kchunk2, q, [&](std::vector<Step> &qq, int k) {
                            qq.push_back(copy_step(1, (k * SEL_OUTER) / OUTER, 0, 0));
                        }, true);
                    } else if (g.overlap) {
                        // block columns are visited last to first: wait for slice c just before its last block column
                        selinv_node_steps(B, x, sp_z, yoff_of[s], splitk_min, kchunk, q, [&](std::vector<Step> &qq, int pb) {
                            const int c = pb / OUTER;
                            if (pb == std::min((c + 1) * OUTER, nblk) - 1) qq.push_back(copy_step(1, c, 0, 0));
                        });
                    } else if (x.winv >= 0) {
                        selinv_node_steps_outer(B, x, sp_z, yoff_of[s], kchunk2, q);
                    } else {
                        selinv_node_steps(B, x, sp_z, yoff_of[s], splitk_min, kchunk, q, true);
                    }
                    B.seq.push_back(std::move(q));
                }
                B.flush();
                const int64_t z0 = zptr[gi][d - g.dmin], z1 = zptr[gi][d - g.dmin + 1];
                if (z1 > z0) {
                    Launch L;
                    memset(&L, 0, sizeof L);
                    L.kind = LK_EXTRACT;
                    L.variant = sp_z;
                    L.a0 = z0;
                    L.a1 = z1;
                    P.launches.push_back(L);
                }
            }
            if (!g.kids.empty()) {
                // push the compact Z_RR of every child segment onto the stack
                const SNode &x = osn[g.root];
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_GATHER;
                L.task0 = (int64_t)P.gather.size();
                L.tile0 = (int64_t)P.tiles.size();
                for (int c : g.kids) {
                    const SNode &cx = osn[segs[c].root];
                    push_gather_task(P, L, segs[c].stack_Z, cx.ldu, 0, cx.nr, x.front, x.ld, x.nc, x.ncp,
                                     p.rel_base + cx.rows, SP_Z0, SP_Z0);
                }
                L.ntasks = (int)(P.gather.size() - L.task0);
                L.ntiles = (int)(P.tiles.size() - L.tile0);
                if (L.ntiles) P.launches.push_back(L);
            }
        }
    }
    built = true;
}

Program &Ooc::solve_program(OocSeg &g, int k, int dir)
{
    std::map<int, Program> &mp = dir ? g.bsolve : g.fsolve;
    auto it = mp.find(k);
    if (it != mp.end()) return it->second;
    Program &P = mp[k];
    const int kp = up2(k);
    const bool blocked = k > 4;
    const int OUTER = env_int("SPDE_SOLVE_OUTER", spde::OUTER, 1);
    if (dir == 0) {
        for (int d = g.dmax; d >= g.dmin; d--) {
            LevelBuilder B(P);
            for (int s : g.by_depth[d - g.dmin]) {
                std::vector<Step> q;
                fsolve_node_steps(B, osn[s], k, kp, blocked, OUTER, q);
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    } else {
        for (int d = g.dmin; d <= g.dmax; d++) {
            LevelBuilder B(P);
            for (int s : g.by_depth[d - g.dmin]) {
                std::vector<Step> q;
                bsolve_node_steps(B, osn[s], k, kp, blocked, OUTER, q);
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    }
    return P;
}

}  // namespace spde

using namespace spde;

// ---------------------------------------------------------------------------------------------
// C ABI

extern "C" int spde_ooc_create(spde_plan *pp, int64_t top_bytes, int want_backward, int build, spde_ooc **out)
{
    if (!pp || !out || top_bytes < 0) { set_error("spde_ooc_create: bad arguments"); return SPDE_ERR_ARG; }
    Ooc *o = new Ooc();
    o->plan = reinterpret_cast<Plan *>(pp);
    o->segment(top_bytes);
    o->plan_memory(want_backward != 0);
    if (build) {
        o->build_tables();
        o->build_programs();
    }
    *out = reinterpret_cast<spde_ooc *>(o);
    return SPDE_OK;
}

extern "C" void spde_ooc_destroy(spde_ooc *oo)
{
    if (!oo) return;
    Ooc *o = reinterpret_cast<Ooc *>(oo);
    cudaFree(o->d_pool); cudaFree(o->d_Xp); cudaFree(o->d_ld); cudaFree(o->d_red); cudaFree(o->d_stage);
    cudaFree(o->d_idx); cudaFree(o->d_perm); cudaFree(o->d_status); cudaFree(o->d_diag);
    if (o->h_pool) cudaFreeHost(o->h_pool);
    if (o->copy_stream) { cudaStreamDestroy(o->copy_stream); cudaEventDestroy(o->copy_fork); cudaEventDestroy(o->copy_done); }
    for (cudaEvent_t e : o->fetch_ev) cudaEventDestroy(e);
    delete o;      // the programs only borrowed slices of the staging buffer
}

/* info ids: 0 segments, 1 top segments, 2 pool bytes, 3 forward peak bytes, 4 backward peak bytes, 5 pinned host
 * bytes, 6 scatter entries, 7 largest segment working set (forward) bytes, 8 factor launches over all segments,
 * 9 selected-inverse launches over all segments */
extern "C" int64_t spde_ooc_info(const spde_ooc *oo, int what)
{
    const Ooc *o = reinterpret_cast<const Ooc *>(oo);
    switch (what) {
    case 0: return (int64_t)o->segs.size();
    case 1: { int64_t c = 0; for (auto &g : o->segs) c += g.top; return c; }
    case 2: return o->pool_size * 8;
    case 3: return o->peak_fwd * 8;
    case 4: return o->peak_bwd * 8;
    case 5: return o->host_size * 8;
    case 6: return (int64_t)o->scat.size();
    case 7: { int64_t w = 0; for (auto &g : o->segs) w = std::max(w, g.l_size + g.dinv_size + g.arena[0] + g.arena[1]); return w * 8; }
    case 8: { int64_t c = 0; for (auto &g : o->segs) c += (int64_t)g.factor.launches.size(); return c; }
    case 9: { int64_t c = 0; for (auto &g : o->segs) c += (int64_t)g.selinv.launches.size(); return c; }
    }
    return -1;
}
extern "C" double spde_ooc_info_d(const spde_ooc *oo, int what)
{
    const Ooc *o = reinterpret_cast<const Ooc *>(oo);
    if (what == 0) return o->recompute_flops;
    if (what == 1) return o->last_ms[0];
    if (what == 2) return o->last_ms[1];
    if (what >= 16 && what - 16 < (int)o->last_ld.size()) return o->last_ld[what - 16];   // log-determinant share of a segment
    return 0.0;
}

/* Host export for the tests and the NumPy interpreter (oracle/plan_emulator.py).
 * seg >= 0: prog 0 factor, 1 forward solve (k), 2 back solve (k), 3 selected inverse; what as spde_plan_export (0..6);
 * seg = -1: what 0 segment table (16 int64 per segment), 1 processing order (int32), 2 scatter entries, 3 diagonal
 * positions, 4 index array (rows | relative indices | identity), 5 extraction entries. */
extern "C" int spde_ooc_export(spde_ooc *oo, int seg, int prog, int k, int what, void *h_out, int64_t *count, int *elem_size)
{
    Ooc &o = *reinterpret_cast<Ooc *>(oo);
    const void *src = nullptr;
    int64_t cnt = 0;
    int es = 0;
    static std::vector<int64_t> tab;
    static std::vector<int> idx;
#define EXP(vec) { src = (vec).data(); cnt = (int64_t)(vec).size(); es = (int)sizeof((vec)[0]); }
    if (seg >= 0) {
        if (seg >= (int)o.segs.size() || !o.built) { set_error("spde_ooc_export: bad segment"); return SPDE_ERR_ARG; }
        OocSeg &g = o.segs[seg];
        Program *P = prog == 0 ? &g.factor : prog == 1 ? &o.solve_program(g, k, 0) : prog == 2 ? &o.solve_program(g, k, 1) : &g.selinv;
        switch (what) {
        case 0: EXP(P->launches) break;
        case 1: EXP(P->gemm) break;
        case 2: EXP(P->tiles) break;
        case 3: EXP(P->potrf) break;
        case 4: EXP(P->ext) break;
        case 5: EXP(P->gather) break;
        case 6: EXP(P->wtw) break;
        case 7:     // slices of an overlapped top segment: (pool offset, doubles, host offset), then [dinv] as the last row
            tab.clear();
            if (g.overlap) {
                for (size_t c = 0; c < g.chunk_off.size(); c++) {
                    tab.push_back(g.chunk_off[c]); tab.push_back(g.chunk_len[c]); tab.push_back(g.host_off + (g.chunk_off[c] - g.off_dinv));
                }
                tab.push_back(g.off_dinv); tab.push_back(g.dinv_size); tab.push_back(g.host_off);
            }
            EXP(tab) break;
        default: set_error("spde_ooc_export: bad what"); return SPDE_ERR_ARG;
        }
    } else {
        switch (what) {
        case 0:
            tab.clear();
            for (auto &g : o.segs) {
                const SNode &x = o.osn[g.root];
                const int64_t row[16] = {g.top, g.keep, g.root, g.parent_seg, g.off_L, g.l_size, g.off_dinv, g.dinv_size,
                                         x.upd, g.u_size, g.stack_U, g.stack_Z, g.scat0, g.scat1, g.col0, g.col1};
                tab.insert(tab.end(), row, row + 16);
            }
            EXP(tab) break;
        case 1: EXP(o.order) break;
        case 2: EXP(o.scat) break;
        case 3: EXP(o.diagpos) break;
        case 4: {
            const Symbolic &S = o.plan->sym;
            idx = S.rows;
            idx.insert(idx.end(), S.relidx.begin(), S.relidx.end());
            int mx = 2;
            for (const SNode &x : o.osn) mx = std::max(mx, x.nr);
            for (int i = 0; i < mx; i++) idx.push_back(i);
            EXP(idx) break;
        }
        case 5: EXP(o.zent) break;
        case 6:
            tab.clear();
            for (auto &g : o.segs) tab.push_back(g.zent0);
            EXP(tab) break;
        default: set_error("spde_ooc_export: bad what"); return SPDE_ERR_ARG;
        }
    }
#undef EXP
    if (count) *count = cnt;
    if (elem_size) *elem_size = es;
    if (h_out && cnt) memcpy(h_out, src, (size_t)cnt * es);
    return SPDE_OK;
}

namespace {

int ensure_ooc_device(Ooc &o, int k, bool backward, size_t stage_need)
{
    Plan &p = *o.plan;
    init_exec_env(p);
    const Symbolic &S = p.sym;
    if (!o.d_pool) {
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_pool, (size_t)o.pool_size * sizeof(double)));
        std::vector<int> idx(S.rows);
        idx.insert(idx.end(), S.relidx.begin(), S.relidx.end());
        int mx = 2;
        for (const SNode &x : o.osn) mx = std::max(mx, x.nr);
        for (int i = 0; i < mx; i++) idx.push_back(i);
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_idx, idx.size() * sizeof(int)));
        SPDE_CUDA_CHECK(cudaMemcpy(o.d_idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_perm, (size_t)S.n * sizeof(int)));
        SPDE_CUDA_CHECK(cudaMemcpy(o.d_perm, S.perm.data(), (size_t)S.n * sizeof(int), cudaMemcpyHostToDevice));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_diag, (size_t)S.n * sizeof(long long)));
        SPDE_CUDA_CHECK(cudaMemcpy(o.d_diag, o.diagpos.data(), (size_t)S.n * sizeof(long long), cudaMemcpyHostToDevice));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_status, sizeof(int)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_ld, std::max<size_t>(o.segs.size(), 1) * sizeof(double)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_red, 1024 * sizeof(double)));
    }
    if (!o.copy_stream) {
        SPDE_CUDA_CHECK(cudaStreamCreateWithFlags(&o.copy_stream, cudaStreamNonBlocking));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&o.copy_fork, cudaEventDisableTiming));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&o.copy_done, cudaEventDisableTiming));
    }
    size_t nev = 1;
    for (const OocSeg &g : o.segs) nev = std::max(nev, g.chunk_off.size());
    while (o.fetch_ev.size() < nev) {
        cudaEvent_t e;
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        o.fetch_ev.push_back(e);
    }
    if (backward && o.host_size > 0 && !o.h_pool) {
        cudaError_t e = cudaHostAlloc((void **)&o.h_pool, (size_t)o.host_size * sizeof(double), cudaHostAllocDefault);
        if (e != cudaSuccess) {
            o.h_pool = nullptr;
            set_error(std::string("streamed evaluation: pinned host pool of ") + std::to_string(o.host_size * 8) + " bytes: " + cudaGetErrorString(e));
            return SPDE_ERR_OOM;
        }
    }
    if (stage_need > (size_t)o.stage_bytes) {
        SPDE_CUDA_CHECK(cudaDeviceSynchronize());
        cudaFree(o.d_stage);
        o.d_stage = nullptr;
        o.stage_bytes = 0;
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_stage, stage_need));
        o.stage_bytes = (int64_t)stage_need;
    }
    const int kp = k + (k & 1);
    const int64_t need = std::max<int64_t>((int64_t)S.n * kp, 2);
    if (o.xp_cap < need) {
        cudaFree(o.d_Xp);
        o.d_Xp = nullptr;
        o.xp_cap = 0;
        SPDE_CUDA_CHECK(cudaMalloc((void **)&o.d_Xp, need * sizeof(double)));
        o.xp_cap = need;
    }
    return SPDE_OK;
}

}  // namespace

/* One streamed pass over the whole tree.
 *   d_Q, d_cnt, tau : as spde_factorize (Q in slot layout, optional diagonal update tau * cnt)
 *   d_X, k, mode    : k right-hand sides (n x k row-major, caller's ordering), solved in place; mode bits as
 *                     spde_solve (1 forward, 2 backward, 4 permute on input, 8 permute on output); k = 0: none
 *   d_Zq            : selected inverse on the pattern of Q (slot layout) or NULL
 *   h_logdet        : log det of the factorised matrix
 * A backward pass runs when mode has bit 2 or d_Zq is given; it needs a plan created with want_backward. */
extern "C" int spde_ooc_run(spde_ooc *oo, const double *d_Q, const double *d_cnt, double tau, double *d_X, int k, int mode,
                            double *d_Zq, double *h_logdet, void *stream)
{
    Ooc &o = *reinterpret_cast<Ooc *>(oo);
    Plan &p = *o.plan;
    if (!o.built) { set_error("spde_ooc_run: plan was created without schedules"); return SPDE_ERR_ARG; }
    if (k < 0 || (k > 0 && !d_X)) { set_error("spde_ooc_run: bad right-hand sides"); return SPDE_ERR_ARG; }
    const bool fsolve = k > 0 && (mode & 1), bsolve = k > 0 && (mode & 2);
    const bool backward = bsolve || d_Zq;
    if (backward && o.peak_bwd > o.pool_size) { set_error("spde_ooc_run: plan was created without a backward pass"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const Symbolic &S = p.sym;
    const int n = S.n, kp = k + (k & 1);
    // staging need: the largest set of tables a segment uses in one pass
    size_t need = 0;
    for (OocSeg &g : o.segs) {
        size_t f = program_bytes(g.factor) + (size_t)(g.scat1 - g.scat0) * sizeof(ScatEntry) + 1024;
        if (fsolve) f += program_bytes(o.solve_program(g, k, 0)) + 256;
        size_t b = 0;
        if (backward) {
            b = program_bytes(g.selinv) * (d_Zq ? 1 : 0) + (size_t)(g.zent1 - g.zent0) * sizeof(ZEntry) * (d_Zq ? 1 : 0) + 1024;
            if (!g.top) b += program_bytes(g.factor) + (size_t)(g.scat1 - g.scat0) * sizeof(ScatEntry);
            if (bsolve) b += program_bytes(o.solve_program(g, k, 1)) + 256;
        }
        need = std::max(need, std::max(f, b));
    }
    int rc = ensure_ooc_device(o, k, backward, need);
    if (rc) return rc;
    ExecCtx ctx;
    for (int i = 0; i < 8; i++) ctx.sp.base[i] = o.d_pool;
    ctx.sp.base[SP_X] = o.d_Xp;
    ctx.sp.idx = o.d_idx;
    ctx.L = o.d_pool; ctx.dinv = o.d_pool; ctx.status = o.d_status; ctx.Zq = d_Zq; ctx.which = 0; ctx.lanes = false;
    // LK_COPY records of the overlapped top segments: no host pool = no backward pass = nothing is parked
    ctx.h_pool = backward ? o.h_pool : nullptr;
    ctx.copy_stream = o.copy_stream; ctx.copy_fork = o.copy_fork; ctx.fetch_ev = &o.fetch_ev;
    const cudaStream_t cs = p.prof_on ? st : o.copy_stream;     // profiled evaluations keep everything on one stream
    // pass-boundary events, destroyed on every exit path (the error returns below included)
    struct PassEvents {
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        ~PassEvents() { for (auto &x : e) if (x) cudaEventDestroy(x); }
    } pev;
    cudaEvent_t *ev = pev.e;
    for (int i = 0; i < 3; i++) SPDE_CUDA_CHECK(cudaEventCreate(&ev[i]));
    SPDE_CUDA_CHECK(cudaMemsetAsync(o.d_status, 0, sizeof(int), st));
    if (k > 0) { rc = launch_perm_in(d_X, o.d_perm, n, k, kp, (mode >> 2) & 1, o.d_Xp, st); if (rc) return rc; }
    SPDE_CUDA_CHECK(cudaEventRecord(ev[0], st));
    const int diag_slot = S.nslots / 2;
    auto scatter_and_factor = [&](OocSeg &g, Stage &sg) -> int {
        ScatEntry *d_sc = nullptr;
        int r = sg.put(o.scat.data() + g.scat0, (size_t)(g.scat1 - g.scat0), &d_sc);
        if (r) return r;
        if ((r = sg.program(g.factor))) return r;
        SPDE_CUDA_CHECK(cudaMemsetAsync(o.d_pool + g.off_L, 0, (size_t)g.l_size * sizeof(double), st));
        const long long cnt = g.scat1 - g.scat0;
        if (cnt > 0) {
            count_launch();
            k_scatter_seg<<<(int)std::min<long long>((cnt + 255) / 256, 148 * 8), 256, 0, st>>>(d_Q, d_sc, cnt, n, diag_slot, d_cnt, tau, o.d_pool);
            SPDE_LAUNCH_CHECK();
        }
        return issue_program_ex(p, g.factor, ctx, st);
    };
    // ---- forward pass
    for (size_t q = 0; q < o.order.size(); q++) {
        OocSeg &g = o.segs[o.order[q]];
        Stage sg{o.d_stage, (size_t)o.stage_bytes, 0, st};
        if ((rc = scatter_and_factor(g, sg))) return rc;
        if ((rc = launch_logdet(o.d_pool, o.d_diag + g.col0, g.col1 - g.col0, o.d_red, o.d_ld + o.order[q], st))) return rc;
        if (fsolve) {
            Program &F = o.solve_program(g, k, 0);
            if ((rc = sg.program(F))) return rc;
            if ((rc = issue_program_ex(p, F, ctx, st))) return rc;
        }
        if (backward && g.host_off >= 0) {
            if (g.overlap) {
                // the slices left on the copy stream during the factorisation; the next segment reuses this memory
                SPDE_CUDA_CHECK(cudaEventRecord(o.copy_done, o.copy_stream));
                SPDE_CUDA_CHECK(cudaStreamWaitEvent(st, o.copy_done, 0));
            } else {
                SPDE_CUDA_CHECK(cudaMemcpyAsync(o.h_pool + g.host_off, o.d_pool + g.off_dinv, (size_t)(o.pool_size - g.off_dinv) * sizeof(double),
                                                cudaMemcpyDeviceToHost, st));
            }
        }
        if (g.u_size > 0)
            SPDE_CUDA_CHECK(cudaMemcpyAsync(o.d_pool + g.stack_U, o.d_pool + o.osn[g.root].upd, (size_t)g.u_size * sizeof(double),
                                            cudaMemcpyDeviceToDevice, st));
    }
    SPDE_CUDA_CHECK(cudaEventRecord(ev[1], st));
    // ---- backward pass
    if (backward) {
        if (d_Zq) SPDE_CUDA_CHECK(cudaMemsetAsync(d_Zq, 0, (size_t)S.nslots * n * sizeof(double), st));
        for (int q = (int)o.order.size() - 1; q >= 0; q--) {
            OocSeg &g = o.segs[o.order[q]];
            Stage sg{o.d_stage, (size_t)o.stage_bytes, 0, st};
            const bool fetch = !g.keep && g.top && g.overlap;
            // tables of this segment's schedules first: the host-to-device copy engine is a FIFO, a table upload queued
            // behind the slices of the panel would hold the compute stream until the whole panel has arrived
            Program *B = bsolve ? &o.solve_program(g, k, 1) : nullptr;
            ZEntry *d_ze = nullptr;
            if (fetch) {
                if (B && (rc = sg.program(*B))) return rc;
                if (d_Zq) {
                    if ((rc = sg.put(o.zent.data() + g.zent0, (size_t)(g.zent1 - g.zent0), &d_ze))) return rc;
                    if ((rc = sg.program(g.selinv))) return rc;
                }
                // the panel comes back slice by slice, last slice first (the order the Takahashi recursion walks the
                // block columns), behind everything that still uses this memory
                SPDE_CUDA_CHECK(cudaEventRecord(o.copy_fork, st));
                SPDE_CUDA_CHECK(cudaStreamWaitEvent(cs, o.copy_fork, 0));
                SPDE_CUDA_CHECK(cudaMemcpyAsync(o.d_pool + g.off_dinv, o.h_pool + g.host_off, (size_t)g.dinv_size * sizeof(double),
                                                cudaMemcpyHostToDevice, cs));
                for (int c = (int)g.chunk_off.size() - 1; c >= 0; c--) {
                    SPDE_CUDA_CHECK(cudaMemcpyAsync(o.d_pool + g.chunk_off[c], o.h_pool + g.host_off + (g.chunk_off[c] - g.off_dinv),
                                                    (size_t)g.chunk_len[c] * sizeof(double), cudaMemcpyHostToDevice, cs));
                    SPDE_CUDA_CHECK(cudaEventRecord(o.fetch_ev[c], cs));
                }
            } else {
                if (!g.keep) {
                    if (g.top) {
                        SPDE_CUDA_CHECK(cudaMemcpyAsync(o.d_pool + g.off_dinv, o.h_pool + g.host_off, (size_t)(o.pool_size - g.off_dinv) * sizeof(double),
                                                        cudaMemcpyHostToDevice, st));
                    } else if ((rc = scatter_and_factor(g, sg))) return rc;
                }
                if (B) {
                    if ((rc = sg.program(*B))) return rc;
                    if ((rc = issue_program_ex(p, *B, ctx, st))) return rc;
                }
                if (d_Zq) {
                    if ((rc = sg.put(o.zent.data() + g.zent0, (size_t)(g.zent1 - g.zent0), &d_ze))) return rc;
                    if ((rc = sg.program(g.selinv))) return rc;
                }
            }
            if (d_Zq) {
                ExecCtx c2 = ctx;
                c2.zent = d_ze;
                if ((rc = issue_program_ex(p, g.selinv, c2, st))) return rc;
            }
            if (fetch) {
                // the solve and the recursion only read the factor, so their order is free: with the panel still arriving
                // the recursion goes first (it waits slice by slice) and the solve runs once everything is there
                SPDE_CUDA_CHECK(cudaStreamWaitEvent(st, o.fetch_ev[0], 0));
                if (B && (rc = issue_program_ex(p, *B, ctx, st))) return rc;
            }
        }
    }
    SPDE_CUDA_CHECK(cudaEventRecord(ev[2], st));
    if (k > 0) { rc = launch_perm_out(o.d_Xp, o.d_perm, n, k, kp, (mode >> 3) & 1, d_X, st); if (rc) return rc; }
    std::vector<double> ld(o.segs.size(), 0.0);
    int h = 0;
    SPDE_CUDA_CHECK(cudaMemcpyAsync(ld.data(), o.d_ld, ld.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    SPDE_CUDA_CHECK(cudaMemcpyAsync(&h, o.d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ev[0], ev[1]); o.last_ms[0] = ms;
    cudaEventElapsedTime(&ms, ev[1], ev[2]); o.last_ms[1] = ms;
    double sum = 0.0;
    for (int i : o.order) sum += ld[i];       // fixed order
    o.last_ld = ld;
    if (h_logdet) *h_logdet = sum;
    if (h) { set_error("matrix is not positive definite (pivot " + std::to_string(h - 1) + " of the permuted matrix)"); return SPDE_ERR_NOT_SPD; }
    return SPDE_OK;
}
