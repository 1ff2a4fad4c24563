// plan_build.cpp -- host-side construction of the storage layout and of the kernel task lists.
//
// Schedule: supernodes are grouped by depth in the supernodal tree; a child is exactly one level
// below its parent, so update matrices (factorisation) and inverse fronts (Takahashi) live in two
// ping-pong arenas indexed by depth parity.  Inside a level every supernode walks the same kind of
// step sequence (GEMM / POTRF / ...); the heads of all sequences that share a kind are merged into
// one grouped launch, so the launch count is ~ (levels x steps of the widest front), not ~ nsuper.
#include "plan_steps.h"

namespace spde {

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
        x.rows = S.rowptr[s];
        x.winv = -1;
        by_depth[x.depth].push_back(s);
    }
    // full inverses of the outer diagonal blocks (two-level Takahashi recursion), behind the 64x64 inverses
    const char *envt = getenv("SPDE_SELINV_OUTER");      // 0: the 64-column recursion everywhere (validation)
    const bool sel_outer = envt ? atoi(envt) != 0 : true;
    for (int s = 0; s < ns; s++) {
        SNode &x = sn[s];
        if (sel_outer && x.nblk > 1) {
            x.winv = dinv_size;
            dinv_size += winv_size(x);
            dinv_size += dinv_size & 1;
        }
        y_used[x.depth] += x.winv >= 0 ? ybuf_need(x) : (int64_t)x.ld * NB;
    }
    arena_size[0] = arena_size[1] = zarena_size[0] = zarena_size[1] = 0;
    ybuf_size = 0;
    for (int s = 0; s < ns; s++) ybuf_size += winv_scratch(sn[s]);      // doubling scratch of all fronts at once (factor epilogue)
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
    // Diagonal-first panels with outer-block inverses (factor_node_steps_diag) on the fronts that have a winv store.
    // Measured on B200 (profiles/r2_factor_diag_ab.txt): C3 factorisation 261 ms with vs 253 ms without, C2 11.8 vs 10.2 ms
    // -- every launch of the dependent chain costs ~10 us whatever its height, and the variant adds nine launches per outer
    // block (copy, six doubling products, the strip product, the copy back) to save tile rows, not launches.  Off by default.
    winv_from_factor = !lookahead && !envo && env_int("SPDE_FACTOR_DIAG", 0, 0) != 0;
    diag_min_ld = env_int("SPDE_FACTOR_DIAG_MIN", 0, 0);      // only fronts at least this tall (their chain launches are the long ones)
    // POTRF of a diagonal block beside the left-looking update of the rows below it (side lane), see factor_node_steps.
    // Measured on B200 (profiles/r2_factor_diag_ab.txt): C3 factorisation 252.2 ms with vs 253.1 ms without, C2 10.30 vs 10.10 ms
    // -- the two extra fork / join edges per 64 columns cost what the overlap gains.  Off by default.
    const bool potrf_overlap = !lookahead && env_int("SPDE_POTRF_OVERLAP", 0, 0) != 0;
    // The update-matrix arena of level d-1 is the arena the extend-add of level d has just consumed, so it is zeroed on the
    // side lane under the (compute-bound) factorisation of level d instead of in front of level d-1: LK_SYNC records fork
    // and join the lane; executors without lanes run the list in order, which is just as valid.
    const bool zero_ahead = !lookahead && env_int("SPDE_ZERO_AHEAD", 1, 0) != 0;
    auto sync = [&](int variant, int ev) {
        Launch L;
        memset(&L, 0, sizeof L);
        L.kind = LK_SYNC; L.variant = variant; L.a0 = ev;
        P.launches.push_back(L);
    };
    auto arena_used = [&](int d) {
        int64_t used = 0;
        for (int s : by_depth[d]) used = std::max(used, sn[s].upd + (int64_t)sn[s].ldu * sn[s].nr);
        return used;
    };
    for (int d = S.maxdepth; d >= 0; d--) {
        const std::vector<int> &lev = by_depth[d];
        const int sp_u = SP_AR0 + (d & 1), sp_child = SP_AR0 + ((d + 1) & 1);
        if (!zero_ahead || d == S.maxdepth) zero_launch(P, sp_u, 0, arena_used(d));
        else if (arena_used(d) > 0) sync(2, d & 1);
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
        if (zero_ahead && d > 0 && arena_used(d - 1) > 0) {
            sync(0, 0);
            zero_launch(P, sp_child, 0, arena_used(d - 1));
            P.launches.back().lane = 1;
            sync(1, (d - 1) & 1);
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
        int64_t yoff = 0;
        for (int s : lev) {
            std::vector<Step> q;
            if (diag_front(sn[s])) factor_node_steps_diag(B, sn[s], sp_u, yoff, q);
            else factor_node_steps(B, sn[s], sp_u, OUTER, q, potrf_overlap);
            yoff += sn[s].winv >= 0 ? ybuf_need(sn[s]) : (int64_t)sn[s].ld * NB;
            B.seq.push_back(std::move(q));
        }
        B.flush();
    }
    // Epilogue: the outer-block inverses Wf of every multi-block front that did not build them on the way (seven grouped
    // launches for the whole tree: they depend on nothing but the finished panels).  The Takahashi recursion and the
    // triangular solves work on them.
    solve_outer = env_int("SPDE_SOLVE_OUTER_BLOCKS", 1, 0) != 0 && !getenv("SPDE_SOLVE_OUTER");
    for (int s = 0; s < S.nsuper; s++)
        if (sn[s].nblk > 1 && sn[s].winv < 0) solve_outer = false;      // (SPDE_SELINV_OUTER=0: no outer-block inverses anywhere)
    if (solve_outer) {
        std::vector<const SNode *> nodes;
        std::vector<int64_t> toffs;
        int64_t toff = 0;
        for (int s = 0; s < S.nsuper; s++)
            if (sn[s].winv >= 0 && !diag_front(sn[s])) { nodes.push_back(&sn[s]); toffs.push_back(toff); toff += winv_scratch(sn[s]); }
        if (!nodes.empty()) winv_level_launches(P, nodes, toffs, 0, false, false);
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
    if (solve_outer && dir == 0) {
        zero_launch(P, SP_X2, 0, (int64_t)kp * S.n);
        for (int d = S.maxdepth; d >= 0; d--) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                std::vector<Step> q;
                fsolve_node_steps_outer(B, sn[s], k, kp, q);
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    } else if (solve_outer) {
        for (int d = 0; d <= S.maxdepth; d++) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                std::vector<Step> q;
                bsolve_node_steps_outer(B, sn[s], k, kp, q);
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    } else if (dir == 0) {
        for (int d = S.maxdepth; d >= 0; d--) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                std::vector<Step> q;
                fsolve_node_steps(B, sn[s], k, kp, blocked, OUTER, q);
                B.seq.push_back(std::move(q));
            }
            B.flush();
        }
    } else {
        for (int d = 0; d <= S.maxdepth; d++) {
            LevelBuilder B(P);
            for (int s : by_depth[d]) {
                std::vector<Step> q;
                bsolve_node_steps(B, sn[s], k, kp, blocked, OUTER, q);
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
    const int kchunk2 = env_int("SPDE_SELINV_KCHUNK2", 0, 0);      // K chunk of the 512-column products of the two-level recursion
    // (the inverse fronts of level d+1 go where those of level d-1 were, last read by the gather of level d: zeroed on
    // the side lane under the recursion of level d, as the update-matrix arenas of the factorisation)
    const bool zero_ahead = env_int("SPDE_ZERO_AHEAD", 1, 0) != 0;
    auto sync = [&](int variant, int ev) {
        Launch L;
        memset(&L, 0, sizeof L);
        L.kind = LK_SYNC; L.variant = variant; L.a0 = ev;
        P.launches.push_back(L);
    };
    auto zarena_used = [&](int d) {
        int64_t used = 0;
        for (int s : by_depth[d]) used = std::max(used, sn[s].front + (int64_t)sn[s].ld * sn[s].ld);
        return used;
    };
    for (int d = 0; d <= S.maxdepth; d++) {
        const std::vector<int> &lev = by_depth[d];
        const int sp_z = SP_Z0 + (d & 1), sp_par = SP_Z0 + ((d + 1) & 1);
        if (!zero_ahead || d == 0) zero_launch(P, sp_z, 0, zarena_used(d));
        else sync(2, d & 1);
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
                    for (int ti = tj; ti < nt; ti++) P.tiles.push_back(TileRef{id, ti, tj, 0});      // (the kernel mirrors)
            }
            L.ntasks = (int)(P.gather.size() - L.task0);
            L.ntiles = (int)(P.tiles.size() - L.tile0);
            if (L.ntiles) P.launches.push_back(L);
        }
        if (zero_ahead && d < S.maxdepth) {
            sync(0, 0);
            zero_launch(P, sp_par, 0, zarena_used(d + 1));
            P.launches.back().lane = 1;
            sync(1, (d + 1) & 1);
        }
        // seeds of the diagonal blocks, hoisted: W^T W of the single-block fronts in one launch, the outer-block
        // inverses Wf and Wf^T Wf of the others in seven grouped launches
        std::vector<int64_t> yoffs;
        {
            std::vector<const SNode *> single, multi, multi_f;
            std::vector<int64_t> ymulti, ymulti_f;
            int64_t yoff = 0;
            for (int s : lev) {
                yoffs.push_back(yoff);
                if (sn[s].winv >= 0 && (solve_outer || diag_front(sn[s]))) { multi_f.push_back(&sn[s]); ymulti_f.push_back(yoff); yoff += ybuf_need(sn[s]); }
                else if (sn[s].winv >= 0) { multi.push_back(&sn[s]); ymulti.push_back(yoff); yoff += ybuf_need(sn[s]); }
                else { single.push_back(&sn[s]); yoff += (int64_t)sn[s].ld * NB; }
            }
            wtw_level_launch(P, single, sp_z);
            if (!multi.empty()) winv_level_launches(P, multi, ymulti, sp_z, false);
            if (!multi_f.empty()) winv_level_launches(P, multi_f, ymulti_f, sp_z, true);     // (Wf left behind by the factorisation)
        }
        LevelBuilder B(P);
        for (size_t i = 0; i < lev.size(); i++) {
            const SNode &x = sn[lev[i]];
            std::vector<Step> q;
            if (x.winv >= 0) selinv_node_steps_outer(B, x, sp_z, yoffs[i], kchunk2, q);
            else selinv_node_steps(B, x, sp_z, yoffs[i], splitk_min, kchunk, q, true);
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
