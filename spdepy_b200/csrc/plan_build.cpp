// plan_build.cpp -- host-side construction of the storage layout and of the kernel task lists.
//
// Schedule: supernodes are grouped by depth in the supernodal tree; a child is exactly one level
// below its parent, so update matrices (factorisation) and inverse fronts (Takahashi) live in two
// ping-pong arenas indexed by depth parity.  Inside a level every supernode walks the same kind of
// step sequence (GEMM / POTRF / ...); the heads of all sequences that share a kind are merged into
// one grouped launch, so the launch count is ~ (levels x steps of the widest front), not ~ nsuper.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "plan.h"

namespace spde {

static inline int up2(int x) { return x + (x & 1); }

void Plan::build_layout()
{
    const Symbolic &S = sym;
    const int ns = S.nsuper;
    sn.resize(ns);
    by_depth.assign(S.maxdepth + 1, {});
    std::vector<int64_t> upd_used(S.maxdepth + 1, 0), front_used(S.maxdepth + 1, 0), y_used(S.maxdepth + 1, 0);
    l_size = 0;
    dinv_size = 0;
    for (int s = 0; s < ns; s++) {
        SNode &x = sn[s];
        x.first = S.first[s];
        x.nc = S.first[s + 1] - S.first[s];
        x.nr = (int)(S.rowptr[s + 1] - S.rowptr[s]);
        x.ncp = up2(x.nc);
        x.ld = x.ncp + up2(x.nr);
        x.nblk = (x.nc + NB - 1) / NB;
        x.depth = S.depth[s];
        x.parent = S.sparent[s];
        x.ldu = up2(x.nr);
        x.panel = l_size;
        l_size += (int64_t)x.ld * x.nc;
        l_size += l_size & 1;
        x.dinv = dinv_size;
        dinv_size += (int64_t)x.nblk * NB * NB;
        x.upd = upd_used[x.depth];
        upd_used[x.depth] += (int64_t)x.ldu * x.nr;
        x.front = front_used[x.depth];
        front_used[x.depth] += (int64_t)x.ld * x.ld;
        y_used[x.depth] += (int64_t)x.ld * NB;
        x.rows = S.rowptr[s];
        by_depth[x.depth].push_back(s);
    }
    arena_size[0] = arena_size[1] = zarena_size[0] = zarena_size[1] = 0;
    ybuf_size = 0;
    for (int d = 0; d <= S.maxdepth; d++) {
        arena_size[d & 1] = std::max(arena_size[d & 1], upd_used[d]);
        zarena_size[d & 1] = std::max(zarena_size[d & 1], front_used[d]);
        ybuf_size = std::max(ybuf_size, y_used[d]);
    }
    // scatter map: lower-triangle entries of the ORIGINAL ordering (what CHOLMOD reads), sent to
    // position (max(pr,pc), min(pr,pc)) of the permuted factor
    const int n = S.n;
    cand_slots.clear();
    for (int q = 0; q < S.nslots; q++) {
        bool lower = (q <= S.nslots / 2);
        if (S.geo.bc == 2 || lower) cand_slots.push_back(q);
    }
    qdest.assign((size_t)cand_slots.size() * n, -1);
    diagpos.assign(n, -1);
    for (int r = 0; r < n; r++) {
        for (size_t ci = 0; ci < cand_slots.size(); ci++) {
            const int c = S.slot_nbr(r, cand_slots[ci]);
            if (c < 0 || c > r) continue;
            const int pr = S.iperm[r], pc = S.iperm[c];
            const int col = std::min(pr, pc), row = std::max(pr, pc);
            const int s = S.snode_of[col];
            const SNode &x = sn[s];
            int lr;
            if (row < x.first + x.nc) lr = row - x.first;
            else {
                const int *b = S.rows.data() + S.rowptr[s];
                const int *e = S.rows.data() + S.rowptr[s + 1];
                lr = x.ncp + (int)(std::lower_bound(b, e, row) - b);
            }
            qdest[ci * n + r] = x.panel + (int64_t)(col - x.first) * x.ld + lr;
        }
    }
    for (int j = 0; j < n; j++) {
        const SNode &x = sn[S.snode_of[j]];
        diagpos[j] = x.panel + (int64_t)(j - x.first) * (x.ld + 1);
    }
    // extraction list of the selected inverse, grouped by depth of the owning supernode
    std::vector<int64_t> cnt(S.maxdepth + 2, 0);
    for (size_t ci = 0; ci < cand_slots.size(); ci++)
        for (int r = 0; r < n; r++) {
            if (qdest[ci * n + r] < 0) continue;
            const int c = S.slot_nbr(r, cand_slots[ci]);
            const int col = std::min(S.iperm[r], S.iperm[c]);
            cnt[S.depth[S.snode_of[col]] + 1]++;
        }
    zdepth_ptr.assign(S.maxdepth + 2, 0);
    for (int d = 0; d <= S.maxdepth; d++) zdepth_ptr[d + 1] = zdepth_ptr[d] + cnt[d + 1];
    zentries.resize(zdepth_ptr[S.maxdepth + 1]);
    std::vector<int64_t> cur(zdepth_ptr.begin(), zdepth_ptr.end() - 1);
    for (size_t ci = 0; ci < cand_slots.size(); ci++)
        for (int r = 0; r < n; r++) {
            const long long dst = qdest[ci * n + r];
            if (dst < 0) continue;
            const int q = cand_slots[ci];
            const int c = S.slot_nbr(r, q);
            const int col = std::min(S.iperm[r], S.iperm[c]);
            const int s = S.snode_of[col];
            ZEntry z;
            z.dst = (long long)q * n + r;
            z.dst2 = (c == r) ? -1 : (long long)(S.nslots - 1 - q) * n + c;
            z.src = sn[s].front + (dst - sn[s].panel);   // fronts share the panel's leading dimension
            z.sn = s;
            z.pad = 0;
            zentries[cur[S.depth[s]]++] = z;
        }
}

// ---------------------------------------------------------------------------------------------
namespace {

constexpr int CFG_BM[3] = {128, 128, 64};
constexpr int CFG_BN[3] = {128, 64, 64};

// tile configuration of a grouped launch: the largest tile that still gives >= 2 CTAs per SM
constexpr int kSMs = 148;
long long count_tiles(const GemmTask &t, int cfg)
{
    const int BM = CFG_BM[cfg], BN = CFG_BN[cfg];
    const int tm = (t.M + BM - 1) / BM, tn = (t.N + BN - 1) / BN;
    if (!(t.flags & GF_LOWER)) return (long long)tm * tn;
    long long c = 0;
    for (int tj = 0; tj < tn; tj++)
        for (int ti = 0; ti < tm; ti++) c += !((ti + 1) * BM - 1 < tj * BN);
    return c;
}

struct Step {
    int kind;            // LaunchKind
    int variant;         // GEMM: cfg*4 + akmaj*2 + bkmaj
    int g0, gn;          // range in the level's GEMM pool
    PotrfTask p;
    WtwTask w;
};

struct LevelBuilder {
    Program &prog;
    std::vector<GemmTask> pool;
    std::vector<std::vector<Step>> seq;   // one sequence per supernode of the level
    explicit LevelBuilder(Program &p) : prog(p) {}

    GemmTask task(int sa, long long a, int lda, int sb, long long b, int ldb, int sc, long long c, int ldc,
                  int M, int N, int K, int flags)
    {
        GemmTask t;
        memset(&t, 0, sizeof t);
        t.a = a; t.b = b; t.c = c; t.c2 = 0;
        t.lda = lda; t.ldb = ldb; t.ldc = ldc;
        t.M = M; t.N = N; t.K = K;
        t.flags = flags | sa | (sb << 3) | (sc << 6);
        return t;
    }
    void add_gemm(std::vector<Step> &s, const GemmTask &t, bool akmaj, bool bkmaj, int cfg = -1)
    {
        if (t.M <= 0 || t.N <= 0 || t.K <= 0) return;
        Step st;
        memset(&st, 0, sizeof st);
        st.kind = LK_GEMM;
        // bits 0-1: operand layouts; bits 2-3: forced tile config + 1 (0 = choose per launch)
        st.variant = ((cfg + 1) << 2) + (akmaj ? 2 : 0) + (bkmaj ? 1 : 0);
        st.g0 = (int)pool.size();
        st.gn = 1;
        pool.push_back(t);
        s.push_back(st);
    }
    // small-M (k <= 4 right-hand sides) matrix-vector step; variant = B layout
    void add_gemv(std::vector<Step> &s, const GemmTask &t, bool bkmaj)
    {
        if (t.M <= 0 || t.N <= 0 || t.K <= 0) return;
        Step st;
        memset(&st, 0, sizeof st);
        st.kind = LK_GEMV;
        st.variant = bkmaj ? 1 : 0;
        st.g0 = (int)pool.size();
        st.gn = 1;
        pool.push_back(t);
        s.push_back(st);
    }
    // dense step of a solve: tensor-core GEMM, or the matrix-vector kernel when there are <= 4 columns
    void add_solve(std::vector<Step> &s, const GemmTask &t, bool bkmaj)
    {
        if (t.M <= 4) add_gemv(s, t, bkmaj);
        else add_gemm(s, t, false, bkmaj);
    }
    void emit_gemv_launch(int bk, const std::vector<const Step *> &steps)
    {
        const int TN = bk ? 64 : 256, KC = bk ? GEMV_KC_K : GEMV_KC_N;
        Launch L;
        memset(&L, 0, sizeof L);
        L.kind = LK_GEMV;
        L.variant = bk;
        L.task0 = (int64_t)prog.gemm.size();
        L.tile0 = (int64_t)prog.tiles.size();
        for (const Step *st : steps)
            for (int g = st->g0; g < st->g0 + st->gn; g++) {
                const GemmTask &t = pool[g];
                const int id = (int)(prog.gemm.size() - L.task0);
                prog.gemm.push_back(t);
                const int tn = (t.N + TN - 1) / TN, tk = (t.K + KC - 1) / KC;
                for (int kc = 0; kc < tk; kc++)
                    for (int tj = 0; tj < tn; tj++) prog.tiles.push_back(TileRef{id, kc, tj, 0});
                prog.flops += 2.0 * t.M * t.N * t.K;
            }
        L.ntasks = (int)(prog.gemm.size() - L.task0);
        L.ntiles = (int)(prog.tiles.size() - L.tile0);
        if (L.ntiles > 0) prog.launches.push_back(L);
    }

    // append `t` to the previous GEMM step (same variant) instead of opening a new step
    void join_gemm(std::vector<Step> &s, const GemmTask &t)
    {
        if (t.M <= 0 || t.N <= 0 || t.K <= 0) return;
        pool.push_back(t);
        s.back().gn++;
    }

    void emit_gemm_launch(int key, const std::vector<const Step *> &steps)
    {
        int cfg = (key >> 2) - 1;
        if (cfg < 0) {
            // Tile shape: measured on B200 (tools/tile_sweep.py, profiles/r1_tile_sweep.txt) the 64x64 tile with
            // four warps (3-4 resident CTAs per SM) matches or beats the larger tiles on every problem shape of the
            // schedules -- square, K = 512 panels and skinny N = 64 -- so it is used for all grouped launches
            // (k-tile 32 with a 2-stage ring: +1-2 % on full launches, +16 % on under-filled skinny ones).
            const char *env = getenv("SPDE_TILE");
            cfg = env ? atoi(env) : 2;
            if (cfg < 0 || cfg > 2) cfg = 2;
        }
        const int variant = cfg * 4 + (key & 3);
        const int BM = CFG_BM[cfg], BN = CFG_BN[cfg];
        Launch L;
        memset(&L, 0, sizeof L);
        L.kind = LK_GEMM;
        L.variant = variant;
        L.task0 = (int64_t)prog.gemm.size();
        L.tile0 = (int64_t)prog.tiles.size();
        // longest tiles first: CTAs are dispatched in tile order, so the long-K tiles of the big fronts start
        // early and the short ones fill the tail (LPT packing of one grouped launch)
        std::vector<int> order;
        for (const Step *st : steps)
            for (int g = st->g0; g < st->g0 + st->gn; g++) order.push_back(g);
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return pool[x].K > pool[y].K; });
        for (int g : order) {
            {
                const GemmTask &t = pool[g];
                const int id = (int)(prog.gemm.size() - L.task0);
                prog.gemm.push_back(t);
                const int tm = (t.M + BM - 1) / BM, tn = (t.N + BN - 1) / BN;
                const bool lower = t.flags & GF_LOWER;
                for (int tj = 0; tj < tn; tj++)
                    for (int ti = 0; ti < tm; ti++) {
                        if (lower && (ti + 1) * BM - 1 < tj * BN) continue;
                        prog.tiles.push_back(TileRef{id, ti, tj, 0});
                    }
                prog.flops += 2.0 * t.M * t.N * t.K * (lower ? 0.5 : 1.0);
            }
        }
        L.ntasks = (int)(prog.gemm.size() - L.task0);
        L.ntiles = (int)(prog.tiles.size() - L.tile0);
        if (L.ntiles > 0) prog.launches.push_back(L);
    }

    // merge the per-supernode sequences into grouped launches
    void flush()
    {
        const size_t m = seq.size();
        std::vector<size_t> head(m, 0);
        size_t remaining = 0;
        for (auto &s : seq) remaining += s.size();
        while (remaining) {
            // pick the (kind, variant) shared by most heads
            std::map<std::pair<int, int>, int> votes;
            for (size_t i = 0; i < m; i++)
                if (head[i] < seq[i].size()) votes[{seq[i][head[i]].kind, (seq[i][head[i]].kind == LK_GEMM || seq[i][head[i]].kind == LK_GEMV) ? seq[i][head[i]].variant : 0}]++;
            std::pair<int, int> best{-1, -1};
            int bv = -1;
            for (auto &kv : votes) if (kv.second > bv) { bv = kv.second; best = kv.first; }
            std::vector<const Step *> chosen;
            for (size_t i = 0; i < m; i++) {
                if (head[i] >= seq[i].size()) continue;
                const Step &st = seq[i][head[i]];
                if (st.kind != best.first) continue;
                if ((st.kind == LK_GEMM || st.kind == LK_GEMV) && st.variant != best.second) continue;
                chosen.push_back(&st);
                head[i]++;
                remaining--;
            }
            if (best.first == LK_GEMM) {
                emit_gemm_launch(best.second, chosen);
            } else if (best.first == LK_GEMV) {
                emit_gemv_launch(best.second, chosen);
            } else if (best.first == LK_POTRF) {
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_POTRF;
                L.task0 = (int64_t)prog.potrf.size();
                for (const Step *st : chosen) prog.potrf.push_back(st->p);
                L.ntasks = (int)chosen.size();
                prog.launches.push_back(L);
            } else if (best.first == LK_WTW) {
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_WTW;
                L.task0 = (int64_t)prog.wtw.size();
                for (const Step *st : chosen) prog.wtw.push_back(st->w);
                L.ntasks = (int)chosen.size();
                prog.launches.push_back(L);
            }
        }
        seq.clear();
        pool.clear();
    }
};

void zero_launch(Program &p, int space, int64_t a0, int64_t a1)
{
    if (a1 <= a0) return;
    Launch L;
    memset(&L, 0, sizeof L);
    L.kind = LK_ZERO;
    L.variant = space;
    L.a0 = a0;
    L.a1 = a1;
    p.launches.push_back(L);
}

}  // namespace

// spaces: 0 L, 1 arena0, 2 arena1, 3 dinv, 4 X, 5 ybuf, 6 zarena0, 7 zarena1
enum { SP_L = 0, SP_AR0 = 1, SP_DINV = 3, SP_X = 4, SP_Y = 5, SP_Z0 = 6 };

void Plan::build_factor_program()
{
    Program &P = factor;
    const Symbolic &S = sym;
    std::vector<std::vector<int>> kids(S.nsuper);
    for (int s = 0; s < S.nsuper; s++)
        if (sn[s].parent >= 0) kids[sn[s].parent].push_back(s);
    const char *envl = getenv("SPDE_LOOKAHEAD");
    // Measured on B200 (tools/lookahead_ab.sh, C3): 792 ms per evaluation with the two-lane schedule vs 783 ms without,
    // with or without stream priorities -- the small panel kernels then queue behind resident bulk-GEMM CTAs (no
    // preemption, ~half a tile time per launch), which costs what the overlap gains.  Off by default.
    const bool lookahead = envl ? atoi(envl) != 0 : false;
    const char *envo = getenv("SPDE_FACTOR_OUTER");     // test hook: small outer blocks exercise the look-ahead on small meshes
    const int OUTER = envo ? std::max(1, atoi(envo)) : spde::OUTER;
    for (int d = S.maxdepth; d >= 0; d--) {
        const std::vector<int> &lev = by_depth[d];
        const int sp_u = SP_AR0 + (d & 1), sp_child = SP_AR0 + ((d + 1) & 1);
        int64_t used = 0;
        for (int s : lev) used = std::max(used, sn[s].upd + (int64_t)sn[s].ldu * sn[s].nr);
        zero_launch(P, sp_u, 0, used);
        // extend-add, one round per child rank (deterministic summation order)
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
                const SNode &c = sn[kids[s][r]];
                const SNode &p = sn[s];
                if (c.nr == 0) continue;
                ExtTask e;
                memset(&e, 0, sizeof e);
                e.src = c.upd; e.lds = c.ldu; e.nr = c.nr;
                e.ppanel = p.panel; e.pld = p.ld; e.pnc = p.nc; e.pncp = p.ncp;
                e.pupd = p.upd; e.pldu = p.ldu;
                e.rel = (int)(rel_base + c.rows);
                e.src_space = sp_child; e.dst_space = sp_u;
                const int id = (int)(P.ext.size() - L.task0);
                P.ext.push_back(e);
                const int nt = (c.nr + 31) / 32;
                for (int tj = 0; tj < nt; tj++)
                    for (int ti = tj; ti < nt; ti++) P.tiles.push_back(TileRef{id, ti, tj, 0});
            }
            L.ntasks = (int)(P.ext.size() - L.task0);
            L.ntiles = (int)(P.tiles.size() - L.tile0);
            if (L.ntiles) P.launches.push_back(L);
        }
        // dense partial Cholesky of every front of the level
        int maxblk = 0;
        for (int s : lev) maxblk = std::max(maxblk, sn[s].nblk);
        if (lookahead && maxblk > OUTER) {
            // Two-lane schedule with look-ahead (levels whose fronts span several outer blocks, i.e. the top of the
            // tree): the panel of outer block O+1 -- a chain of small, latency-bound launches -- runs on the main lane
            // while the far part of the trailing update of block O and the update-matrix contribution of block O run
            // on the bulk lane.  Per outer block O:
            //   main: [wait far(O-2)]  near(O-1): block O-1 -> the columns of block O;   panel(O)
            //   bulk: [after panel(O)]  far(O): block O -> the columns beyond block O+1;  U -= L21[:,O] L21[:,O]^T
            // near(O-1) and far(O-2) write the same columns, hence the wait; everything else is disjoint.
            auto sync = [&](int variant, int ev) {
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_SYNC; L.variant = variant; L.a0 = ev;
                P.launches.push_back(L);
            };
            const int maxO = (maxblk + OUTER - 1) / OUTER;
            for (int O = 0; O < maxO; O++) {
                const int P0 = O * OUTER;
                if (O >= 1) {
                    if (O >= 2) sync(2, (O - 2) & 1);
                    LevelBuilder Bn(P);
                    for (int s : lev) {
                        const SNode &x = sn[s];
                        if (x.nblk <= P0) continue;
                        const int mrows = x.ncp + x.nr;
                        const int cPp = (P0 - OUTER) * NB, cP = P0 * NB, cE = std::min((P0 + OUTER) * NB, x.nc);
                        std::vector<Step> q;
                        Bn.add_gemm(q, Bn.task(SP_L, x.panel + cP + (int64_t)cPp * x.ld, x.ld,
                                               SP_L, x.panel + cP + (int64_t)cPp * x.ld, x.ld,
                                               SP_L, x.panel + cP + (int64_t)cP * x.ld, x.ld,
                                               mrows - cP, cE - cP, cP - cPp, GF_NEG | GF_LOWER), false, false);
                        Bn.seq.push_back(std::move(q));
                    }
                    Bn.flush();
                }
                {
                    LevelBuilder Bp(P);
                    for (int s : lev) {
                        const SNode &x = sn[s];
                        if (x.nblk <= P0) continue;
                        const int mrows = x.ncp + x.nr;
                        const int P1 = std::min(P0 + OUTER, x.nblk), cP = P0 * NB;
                        std::vector<Step> q;
                        for (int p = P0; p < P1; p++) {
                            const int c0 = p * NB, b = std::min(NB, x.nc - c0);
                            if (p > P0)
                                Bp.add_gemm(q, Bp.task(SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                                       SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                                       SP_L, x.panel + c0 + (int64_t)c0 * x.ld, x.ld,
                                                       mrows - c0, b, c0 - cP, GF_NEG), false, false);
                            Step st;
                            memset(&st, 0, sizeof st);
                            st.kind = LK_POTRF;
                            st.p.blk = x.panel + c0 + (int64_t)c0 * x.ld;
                            st.p.dinv = x.dinv + (int64_t)p * NB * NB;
                            st.p.ld = x.ld; st.p.b = b; st.p.col0 = x.first + c0;
                            q.push_back(st);
                            const int r0 = (p == x.nblk - 1) ? x.ncp : c0 + NB;
                            Bp.add_gemm(q, Bp.task(SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                                   SP_DINV, st.p.dinv, NB,
                                                   SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                                   mrows - r0, b, b, GF_BETA0), false, false);
                        }
                        Bp.seq.push_back(std::move(q));
                    }
                    Bp.flush();
                }
                sync(0, 0);
                const size_t from = P.launches.size();
                {
                    LevelBuilder Bf(P);
                    for (int s : lev) {
                        const SNode &x = sn[s];
                        if (x.nblk <= P0) continue;
                        const int mrows = x.ncp + x.nr;
                        const int cP = P0 * NB, cE = std::min((P0 + OUTER) * NB, x.nc);
                        const int c2 = std::min((P0 + 2 * OUTER) * NB, x.nc);      // end of the next outer block
                        std::vector<Step> q;
                        bool opened = false;
                        if (c2 < x.nc) {
                            Bf.add_gemm(q, Bf.task(SP_L, x.panel + c2 + (int64_t)cP * x.ld, x.ld,
                                                   SP_L, x.panel + c2 + (int64_t)cP * x.ld, x.ld,
                                                   SP_L, x.panel + c2 + (int64_t)c2 * x.ld, x.ld,
                                                   mrows - c2, x.nc - c2, cE - cP, GF_NEG | GF_LOWER), false, false);
                            opened = true;
                        }
                        if (x.nr > 0) {
                            GemmTask u = Bf.task(SP_L, x.panel + x.ncp + (int64_t)cP * x.ld, x.ld,
                                                 SP_L, x.panel + x.ncp + (int64_t)cP * x.ld, x.ld,
                                                 sp_u, x.upd, x.ldu, x.nr, x.nr, cE - cP, GF_NEG | GF_LOWER);
                            if (opened) Bf.join_gemm(q, u);
                            else Bf.add_gemm(q, u, false, false);
                        }
                        if (!q.empty()) Bf.seq.push_back(std::move(q));
                    }
                    Bf.flush();
                }
                for (size_t i = from; i < P.launches.size(); i++) P.launches[i].lane = 1;
                sync(1, O & 1);
            }
            sync(2, (maxO - 1) & 1);
            continue;
        }
        LevelBuilder B(P);
        for (int s : lev) {
            const SNode &x = sn[s];
            std::vector<Step> q;
            const int mrows = x.ncp + x.nr;     // panel rows in use (the gap row of an odd nc is zero)
            for (int P0 = 0; P0 < x.nblk; P0 += OUTER) {
                const int P1 = std::min(P0 + OUTER, x.nblk);
                const int cP = P0 * NB;
                for (int p = P0; p < P1; p++) {
                    const int c0 = p * NB, b = std::min(NB, x.nc - c0);
                    if (p > P0) {
                        // left-looking update of block column p from the inner blocks of this outer block
                        const int K = c0 - cP;
                        B.add_gemm(q, B.task(SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                             SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                             SP_L, x.panel + c0 + (int64_t)c0 * x.ld, x.ld,
                                             mrows - c0, b, K, GF_NEG), false, false);
                    }
                    Step st;
                    memset(&st, 0, sizeof st);
                    st.kind = LK_POTRF;
                    st.p.blk = x.panel + c0 + (int64_t)c0 * x.ld;
                    st.p.dinv = x.dinv + (int64_t)p * NB * NB;
                    st.p.ld = x.ld; st.p.b = b; st.p.col0 = x.first + c0;
                    q.push_back(st);
                    // rows below the diagonal block: L = A * W^T, in place
                    const int r0 = (p == x.nblk - 1) ? x.ncp : c0 + NB;
                    B.add_gemm(q, B.task(SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                         SP_DINV, st.p.dinv, NB,
                                         SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                         mrows - r0, b, b, GF_BETA0), false, false);
                }
                if (P1 < x.nblk) {
                    // right-looking update of the panel columns beyond this outer block
                    const int cR = P1 * NB, K = cR - cP;
                    B.add_gemm(q, B.task(SP_L, x.panel + cR + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + cR + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + cR + (int64_t)cR * x.ld, x.ld,
                                         mrows - cR, x.nc - cR, K, GF_NEG | GF_LOWER), false, false);
                }
            }
            if (x.nr > 0) {
                // update matrix U -= L21 L21^T (lower), K = all pivot columns
                B.add_gemm(q, B.task(SP_L, x.panel + x.ncp, x.ld, SP_L, x.panel + x.ncp, x.ld,
                                     sp_u, x.upd, x.ldu, x.nr, x.nr, x.nc, GF_NEG | GF_LOWER), false, false);
            }
            B.seq.push_back(std::move(q));
        }
        B.flush();
    }
}

// direction 0: forward (L y = b), 1: backward (L^T x = y)
Program &Plan::solve_program(int k, int dir)
{
    auto key = std::make_pair(k, dir);
    auto it = solve.find(key);
    if (it != solve.end()) return it->second;
    Program &P = solve[key];
    const Symbolic &S = sym;
    const int kp = up2(k);
    const bool blocked = k > 4;          // tensor-core path (see add_solve)
    const char *envo = getenv("SPDE_SOLVE_OUTER");     // test hook: exercise the outer-block path on small meshes
    const int OUTER = envo ? std::max(1, atoi(envo)) : spde::OUTER;
    if (dir == 0) {
        for (int d = S.maxdepth; d >= 0; d--) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                const SNode &x = sn[s];
                std::vector<Step> q;
                const int64_t xs = (int64_t)x.first * kp;
                // Many right-hand sides (tensor-core path): two-level blocking as in the factorisation -- inside an
                // outer block of OUTER*64 pivot columns the update after each 64-block stays inside the outer block
                // (N <= 448), and one K = 512 update per outer block reaches the remaining pivot columns.  A few
                // right-hand sides (matrix-vector path): one update of all remaining columns per 64-block.
                const int outer = blocked ? OUTER : x.nblk;
                for (int P0 = 0; P0 < x.nblk; P0 += outer) {
                    const int P1 = std::min(P0 + outer, x.nblk);
                    const int cEnd = std::min(P1 * NB, x.nc);
                    for (int p = P0; p < P1; p++) {
                        const int c0 = p * NB, b = std::min(NB, x.nc - c0);
                        // y_p = x_p W_p^T (in place)
                        B.add_solve(q, B.task(SP_X, xs + (int64_t)c0 * kp, kp, SP_DINV, x.dinv + (int64_t)p * NB * NB, NB,
                                             SP_X, xs + (int64_t)c0 * kp, kp, k, b, b, GF_BETA0), false);
                        // remaining pivot columns of this outer block
                        const int rest = cEnd - c0 - b;
                        if (rest > 0)
                            B.add_solve(q, B.task(SP_X, xs + (int64_t)c0 * kp, kp,
                                                 SP_L, x.panel + (c0 + b) + (int64_t)c0 * x.ld, x.ld,
                                                 SP_X, xs + (int64_t)(c0 + b) * kp, kp, k, rest, b, GF_NEG), false);
                    }
                    if (cEnd < x.nc) {
                        const int cP = P0 * NB;
                        B.add_solve(q, B.task(SP_X, xs + (int64_t)cP * kp, kp,
                                             SP_L, x.panel + cEnd + (int64_t)cP * x.ld, x.ld,
                                             SP_X, xs + (int64_t)cEnd * kp, kp, k, x.nc - cEnd, cEnd - cP, GF_NEG), false);
                    }
                }
                if (x.nr > 0) {
                    GemmTask t = B.task(SP_X, xs, kp, SP_L, x.panel + x.ncp, x.ld, SP_X, 0, kp, k, x.nr, x.nc,
                                        GF_NEG | GF_SCATTER_C | GF_ATOMIC);
                    t.cidx = (int)x.rows;
                    B.add_solve(q, t, false);
                }
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    } else {
        for (int d = 0; d <= S.maxdepth; d++) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                const SNode &x = sn[s];
                std::vector<Step> q;
                const int64_t xs = (int64_t)x.first * kp;
                if (x.nr > 0) {
                    // x_s -= X[rows below] * L21   (columns of X gathered through the row list)
                    GemmTask t = B.task(SP_X, 0, kp, SP_L, x.panel + x.ncp, x.ld, SP_X, xs, kp, k, x.nc, x.nr,
                                        GF_NEG | GF_GATHER_A);
                    t.aidx = (int)x.rows;
                    B.add_solve(q, t, true);
                }
                // Many right-hand sides: left-looking only inside an outer block (K <= 448), then one right-looking
                // K = 512 update of ALL earlier pivot columns per outer block -- a long-K product on a 64-column block
                // would have k/64 tiles for the whole machine (measured: 1024 samples on C3 at ~5 TFLOP/s).
                const int outer = blocked ? OUTER : x.nblk;
                const int nouter = (x.nblk + outer - 1) / outer;
                for (int o = nouter - 1; o >= 0; o--) {
                    const int P0 = o * outer, P1 = std::min(P0 + outer, x.nblk);
                    const int cEnd = std::min(P1 * NB, x.nc), cP = P0 * NB;
                    for (int p = P1 - 1; p >= P0; p--) {
                        const int c0 = p * NB, b = std::min(NB, x.nc - c0);
                        const int later = cEnd - c0 - b;
                        if (later > 0)   // x_p -= X[later pivot columns of the outer block] * L[later, p]
                            B.add_solve(q, B.task(SP_X, xs + (int64_t)(c0 + b) * kp, kp,
                                                 SP_L, x.panel + (c0 + b) + (int64_t)c0 * x.ld, x.ld,
                                                 SP_X, xs + (int64_t)c0 * kp, kp, k, b, later, GF_NEG), true);
                        // x_p = y_p W_p (in place)
                        B.add_solve(q, B.task(SP_X, xs + (int64_t)c0 * kp, kp, SP_DINV, x.dinv + (int64_t)p * NB * NB, NB,
                                             SP_X, xs + (int64_t)c0 * kp, kp, k, b, b, GF_BETA0), true);
                    }
                    if (cP > 0)          // X[earlier pivot columns] -= X[outer block] * L[outer block, earlier]
                        B.add_solve(q, B.task(SP_X, xs + (int64_t)cP * kp, kp,
                                             SP_L, x.panel + cP, x.ld,
                                             SP_X, xs, kp, k, cP, cEnd - cP, GF_NEG), true);
                }
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    }
    return P;
}

void Plan::build_selinv_program()
{
    if (selinv_built) return;
    selinv_built = true;
    Program &P = selinv;
    const Symbolic &S = sym;
    const char *env = getenv("SPDE_SPLITK_MIN");          // test hook: exercise the split-K path on small meshes
    const int splitk_min = env ? std::max(8, atoi(env)) : 2048;
    const char *envk = getenv("SPDE_SELINV_KCHUNK");
    const int kchunk = envk ? atoi(envk) : 1024;    // measured on C3 (tools/kchunk_sweep.sh): none 786 ms, 1024 781, 512 792, 256 830
    for (int d = 0; d <= S.maxdepth; d++) {
        const std::vector<int> &lev = by_depth[d];
        const int sp_z = SP_Z0 + (d & 1), sp_par = SP_Z0 + ((d + 1) & 1);
        int64_t used = 0;
        for (int s : lev) used = std::max(used, sn[s].front + (int64_t)sn[s].ld * sn[s].ld);
        zero_launch(P, sp_z, 0, used);
        // Z_II of every front <- parent's front
        {
            Launch L;
            memset(&L, 0, sizeof L);
            L.kind = LK_GATHER;
            L.task0 = (int64_t)P.gather.size();
            L.tile0 = (int64_t)P.tiles.size();
            for (int s : lev) {
                const SNode &x = sn[s];
                if (x.parent < 0 || x.nr == 0) continue;
                const SNode &p = sn[x.parent];
                GatherTask g;
                memset(&g, 0, sizeof g);
                g.dst = x.front; g.ldd = x.ld; g.ncp = x.ncp; g.nr = x.nr;
                g.src = p.front; g.lds = p.ld; g.pnc = p.nc; g.pncp = p.ncp;
                g.rel = (int)(rel_base + x.rows);
                g.src_space = sp_par; g.dst_space = sp_z;
                const int id = (int)(P.gather.size() - L.task0);
                P.gather.push_back(g);
                const int nt = (x.nr + 31) / 32;
                for (int tj = 0; tj < nt; tj++)
                    for (int ti = 0; ti < nt; ti++) P.tiles.push_back(TileRef{id, ti, tj, 0});
            }
            L.ntasks = (int)(P.gather.size() - L.task0);
            L.ntiles = (int)(P.tiles.size() - L.tile0);
            if (L.ntiles) P.launches.push_back(L);
        }
        LevelBuilder B(P);
        int64_t yoff = 0;
        for (int s : lev) {
            const SNode &x = sn[s];
            std::vector<Step> q;
            const int mrows = x.ncp + x.nr;
            const int64_t F = x.front;
            const int64_t Y = yoff;
            yoff += (int64_t)x.ld * NB;
            for (int p = x.nblk - 1; p >= 0; p--) {
                const int c0 = p * NB, b = std::min(NB, x.nc - c0);
                const int r0 = (p == x.nblk - 1) ? x.ncp : c0 + NB;
                const int mb = mrows - r0;
                const int64_t W = x.dinv + (int64_t)p * NB * NB;
                const int ldy = up2(std::max(mb, 2));
                // Z_pp = W^T W  (+ correction below)
                Step st;
                memset(&st, 0, sizeof st);
                st.kind = LK_WTW;
                st.w.w = W; st.w.dst = F + c0 + (int64_t)c0 * x.ld; st.w.ldd = x.ld; st.w.b = b; st.w.space = sp_z;
                q.push_back(st);
                if (mb <= 0) continue;
                // Y = L[below,p] * W
                B.add_gemm(q, B.task(SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld, SP_DINV, W, NB,
                                     SP_Y, Y, ldy, mb, b, b, GF_BETA0), false, true);
                // Z[below,p] = -Z[below,below] * Y   (and its transpose into the row block)
                // Skinny product (N <= 64): for the big fronts near the root there are fewer row tiles than
                // SMs, so K is split into chunks that accumulate atomically into the (still zero) block.
                // The K chunks also bound the duration of one tile: a launch ends with a partially filled wave of
                // CTAs, and with K = mb in the thousands one 64x64 tile runs for hundreds of microseconds
                // (profiles/: mean selinv launch ~0.5 ms), so short chunks keep the tail of every launch short.
                int nchunk = 1;
                if (mb >= splitk_min) {
                    const int rowtiles = (mb + 127) / 128;
                    nchunk = std::max(1, std::min((2 * kSMs + rowtiles - 1) / rowtiles, mb / (splitk_min / 4)));
                }
                if (kchunk > 0 && mb > kchunk + kchunk / 2) nchunk = std::max(nchunk, (mb + kchunk - 1) / kchunk);
                int clen = (mb + nchunk - 1) / nchunk;
                clen += clen & 1;
                for (int k0 = 0, ci = 0; k0 < mb; k0 += clen, ci++) {
                    GemmTask t = B.task(sp_z, F + r0 + (int64_t)(r0 + k0) * x.ld, x.ld, SP_Y, Y + k0, ldy,
                                        sp_z, F + r0 + (int64_t)c0 * x.ld, x.ld, mb, b, std::min(clen, mb - k0),
                                        GF_NEG | GF_UPPER_MIRROR | (nchunk == 1 ? GF_BETA0 : GF_ATOMIC));
                    t.c2 = F + c0 + (int64_t)r0 * x.ld;
                    if (ci == 0) B.add_gemm(q, t, false, true);
                    else B.join_gemm(q, t);
                }
                // Z_pp -= Y^T Z[below,p]   (split over K, accumulated atomically)
                const int chunk = 512;
                bool opened = false;
                for (int k0 = 0; k0 < mb; k0 += chunk) {
                    GemmTask u = B.task(SP_Y, Y + k0, ldy, sp_z, F + (r0 + k0) + (int64_t)c0 * x.ld, x.ld,
                                        sp_z, F + c0 + (int64_t)c0 * x.ld, x.ld, b, b, std::min(chunk, mb - k0),
                                        GF_NEG | GF_ATOMIC);
                    if (!opened) { B.add_gemm(q, u, true, true, 2); opened = true; }
                    else B.join_gemm(q, u);
                }
            }
            B.seq.push_back(std::move(q));
        }
        B.flush();
        // copy Z on the pattern of Q out of the fronts of this level
        if (zdepth_ptr[d + 1] > zdepth_ptr[d]) {
            Launch L;
            memset(&L, 0, sizeof L);
            L.kind = LK_EXTRACT;
            L.variant = sp_z;
            L.a0 = zdepth_ptr[d];
            L.a1 = zdepth_ptr[d + 1];
            P.launches.push_back(L);
        }
    }
}

}  // namespace spde
