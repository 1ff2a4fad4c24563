// plan_steps.h -- building blocks shared by the schedule builders (plan_build.cpp: whole-tree, in-core
// schedules; ooc.cpp: per-segment schedules of the streamed evaluator): the level builder that merges
// per-supernode step sequences into grouped launches, and the step sequence of one supernode for the
// factorisation, the triangular solves and the Takahashi recursion.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>

#include "plan.h"

namespace spde {

static inline int up2(int x) { return x + (x & 1); }
static inline int env_int(const char *name, int dflt, int lo)
{
    const char *e = getenv(name);
    return e ? std::max(lo, atoi(e)) : dflt;
}

// tile shapes of the grouped GEMM configurations; 3 = the warp-specialised bulk-async kernel (N/N layout only)
static constexpr int CFG_BM[4] = {128, 128, 64, 128};
static constexpr int CFG_BN[4] = {128, 64, 64, 64};
static constexpr int CFG_WS = 3;

// tile configuration of a grouped launch: the largest tile that still gives >= 2 CTAs per SM
static constexpr int kSMs = 148;
static inline long long count_tiles(const GemmTask &t, int cfg)
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
    CopyTask cp;
    Launch raw;          // LK_COPY: the launch record itself
    int lane;            // 0 = main lane, 1 = side lane (GEMM steps); LK_SYNC: event number
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
    // lane fork / join record as a step (see LK_SYNC in plan.h): 0 = side lane waits for the main lane, 1 = record event
    // `ev` on the side lane, 2 = main lane waits for event `ev`
    void add_sync(std::vector<Step> &s, int variant, int ev)
    {
        Step st;
        memset(&st, 0, sizeof st);
        st.kind = LK_SYNC; st.variant = variant; st.lane = ev;
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
            // Operands with the tile dimension contiguous on both sides (the factorisation's updates, the Takahashi
            // product, the forward substitution) go to the warp-specialised bulk-async kernel: 34-35 TFLOP/s where the
            // cp.async ring gives 30-32 (tools/gemm_lab.cu, profiles/r2_gemm_lab.txt).  SPDE_GEMM_WS=0 disables it.
            const char *env = getenv("SPDE_TILE");
            cfg = env ? atoi(env) : 2;
            if (cfg < 0 || cfg > 2) cfg = 2;
            const char *ws = getenv("SPDE_GEMM_WS");
            if ((key & 3) == 0 && !(ws && atoi(ws) == 0)) {
                // ... unless the launch has fewer 128x64 tiles than SMs: such launches are latency-bound, and the 64x64 ring
                // kernel (twice the CTAs, shorter prologue) is 10-15 % quicker there (profiles/r2_ws_sweep.txt)
                long long nt = 0;
                for (const Step *st : steps)
                    for (int g = st->g0; g < st->g0 + st->gn && nt < kSMs; g++) nt += count_tiles(pool[g], CFG_WS);
                if (nt >= kSMs) cfg = CFG_WS;
            }
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
        // Under-filled launches: a 64x64 tile is bound by the FP64 tensor pipe of the ONE SM it runs on (1.1 us per 32-deep
        // k-tile, tools/deep_sweep.py), so a launch with fewer tiles than SMs takes as long as its longest K however few
        // tiles it has.  Accumulating tasks of such a launch are cut along K into chunks that accumulate atomically, until
        // the launch has about one tile per SM.
        std::vector<GemmTask> cut;
        {
            static const int splitk = env_int("SPDE_SMALL_SPLITK", 1, 0);
            long long nt = 0;
            int kmax = 0;
            bool ok = splitk != 0 && cfg == 2;
            for (int g : order) {
                nt += count_tiles(pool[g], cfg);
                kmax = std::max(kmax, pool[g].K);
                if (pool[g].flags & GF_BETA0) ok = false;
            }
            int nsplit = (ok && nt > 0 && nt * 2 <= kSMs) ? (int)std::min<long long>(kSMs / nt, kmax / 64) : 1;
            if (nsplit > 1) {
                const int ak = (key >> 1) & 1, bk = key & 1;
                for (int g : order) {
                    const GemmTask &t = pool[g];
                    int clen = (t.K + nsplit - 1) / nsplit;
                    clen = std::max(64, (clen + 31) / 32 * 32);
                    for (int k0 = 0; k0 < t.K; k0 += clen) {
                        GemmTask c = t;
                        c.K = std::min(clen, t.K - k0);
                        c.flags |= GF_ATOMIC;
                        if (t.flags & GF_GATHER_A) c.aidx += k0;
                        else c.a += ak ? (long long)k0 : (long long)k0 * t.lda;
                        c.b += bk ? (long long)k0 : (long long)k0 * t.ldb;
                        cut.push_back(c);
                    }
                }
                order.clear();
                for (size_t i = 0; i < cut.size(); i++) order.push_back((int)i);
            }
        }
        // Wave balance: a launch of T equal-length tiles on S resident slots takes ceil(T/S) tile times; with T a few
        // times S the last, partly filled wave is a sizeable fraction of the launch.  Cutting every task along K into s
        // chunks (atomic accumulation) makes s T shorter tiles: ceil(s T / S) / s tile times.  Applied when it saves more than
        // the atomics cost (~3 % per extra chunk), to launches whose tasks all accumulate (or write a zeroed block, GF_ZDEST).
        if (cut.empty()) {
            static const int wavebal = env_int("SPDE_WAVE_BALANCE", 1, 0);
            const long long S = (cfg == CFG_WS ? 2 : 4) * (long long)kSMs;
            long long T = 0;
            int kmin = 1 << 30;
            bool ok = wavebal != 0 && (cfg == CFG_WS || cfg == 2);
            for (int g : order) {
                T += count_tiles(pool[g], cfg);
                kmin = std::min(kmin, pool[g].K);
                if ((pool[g].flags & GF_BETA0) && !(pool[g].flags & GF_ZDEST)) ok = false;
            }
            if (ok && T >= S && T < 8 * S) {
                int best = 1;
                double bc = (double)((T + S - 1) / S);
                for (int sp = 2; sp <= 4 && kmin / sp >= 128; sp++) {
                    const double c = (double)((T * sp + S - 1) / S) / sp * (1.0 + 0.03 * (sp - 1));
                    if (c < bc * 0.97) { bc = c; best = sp; }
                }
                if (best > 1) {
                    const int ak = (key >> 1) & 1, bk = key & 1;
                    for (int g : order) {
                        const GemmTask &t = pool[g];
                        int clen = (t.K + best - 1) / best;
                        clen = (clen + 31) / 32 * 32;
                        for (int k0 = 0; k0 < t.K; k0 += clen) {
                            GemmTask c = t;
                            c.K = std::min(clen, t.K - k0);
                            c.flags = (c.flags & ~GF_BETA0) | GF_ATOMIC;
                            if (t.flags & GF_GATHER_A) c.aidx += k0;
                            else c.a += ak ? (long long)k0 : (long long)k0 * t.lda;
                            c.b += bk ? (long long)k0 : (long long)k0 * t.ldb;
                            cut.push_back(c);
                        }
                    }
                    order.clear();
                    for (size_t i = 0; i < cut.size(); i++) order.push_back((int)i);
                }
            }
        }
        const std::vector<GemmTask> &src = cut.empty() ? pool : cut;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return src[x].K > src[y].K; });
        for (int g : order) {
            {
                const GemmTask &t = src[g];
                const int id = (int)(prog.gemm.size() - L.task0);
                prog.gemm.push_back(t);
                const int tm = (t.M + BM - 1) / BM, tn = (t.N + BN - 1) / BN;
                const bool lower = t.flags & GF_LOWER;
                // L2-aware raster: CTAs are dispatched in tile order and the ~300 resident ones walk K in step, so an
                // operand panel is fetched from DRAM once per group of resident tiles that share it.  Column-major order
                // makes ~4 tile columns resident: every A row panel is shared by 4 tiles only.  Supertiles of SJ x SI
                // (16 x 16, or all columns x 256/columns for narrow products) share A panels 16-fold and B panels 16-fold.
                static const int sjmax = env_int("SPDE_SUPERTILE", 16, 1);      // (sweep hook: 8 / 16 / 32 columns of tiles)
                const int SJ = std::min(tn, sjmax), SI = std::max(1, 256 / SJ);
                for (int tj0 = 0; tj0 < tn; tj0 += SJ)
                    for (int ti0 = 0; ti0 < tm; ti0 += SI)
                        for (int tj = tj0; tj < std::min(tj0 + SJ, tn); tj++)
                            for (int ti = ti0; ti < std::min(ti0 + SI, tm); ti++) {
                                if (lower && (ti + 1) * BM - 1 < tj * BN) continue;
                                prog.tiles.push_back(TileRef{id, ti, tj, 0});
                            }
                // lower: the trapezoid on and below the diagonal of an M x N block (M >= N)
                prog.flops += 2.0 * t.K * (lower ? (double)t.M * t.N - 0.5 * t.N * std::min(t.M, t.N) : (double)t.M * t.N);
            }
        }
        // Tail split: with T tiles on S resident slots the last T mod S tiles run as a partly filled wave, a whole tile time
        // long.  For launches of many waves (too many for the global K-split above to pay) only THOSE tiles are cut along K
        // into S / (T mod S) chunks each (one-tile tasks that accumulate atomically), so the last wave is full and short.
        {
            const int tailsplit = env_int("SPDE_TAIL_SPLIT", 1, 0);
            const int slots_env = env_int("SPDE_TAIL_SLOTS", 0, 0);      // test hook: pretend the machine has this many slots
            const long long S = slots_env ? slots_env : (cfg == CFG_WS ? 2 : 4) * (long long)kSMs;
            const long long T = (long long)prog.tiles.size() - L.tile0;
            const long long R = T % S;
            if (tailsplit && (cfg == CFG_WS || cfg == 2) && T >= 2 * S && R > 0 && 2 * R <= S) {
                const int ak = (key >> 1) & 1, bk = key & 1;
                const int sp = (int)std::min<long long>(S / R, 8);
                std::vector<TileRef> tail(prog.tiles.end() - R, prog.tiles.end()), keep, split;
                prog.tiles.resize(prog.tiles.size() - R);
                for (const TileRef &tr : tail) {
                    const GemmTask t = prog.gemm[L.task0 + tr.task];
                    const int i0 = tr.ti * BM, j0 = tr.tj * BN;
                    const bool below = !(t.flags & GF_LOWER) || i0 >= j0 + BN - 1;      // no element of the tile is masked
                    const bool accum = !(t.flags & GF_BETA0) || (t.flags & GF_ZDEST);
                    const int nch = std::min(sp, t.K / 128);
                    if (!below || !accum || nch < 2) { keep.push_back(tr); continue; }
                    int clen = (t.K + nch - 1) / nch;
                    clen = (clen + 31) / 32 * 32;
                    for (int k0 = 0; k0 < t.K; k0 += clen) {
                        GemmTask c = t;
                        c.M = std::min(BM, t.M - i0); c.N = std::min(BN, t.N - j0); c.K = std::min(clen, t.K - k0);
                        if (t.flags & GF_GATHER_A) { c.aidx = t.aidx + k0; c.a = t.a + i0; }
                        else c.a = t.a + (ak ? (long long)i0 * t.lda + k0 : (long long)i0 + (long long)k0 * t.lda);
                        c.b = t.b + (bk ? (long long)j0 * t.ldb + k0 : (long long)j0 + (long long)k0 * t.ldb);
                        if (t.flags & GF_SCATTER_C) { c.cidx = t.cidx + j0; c.c = t.c + i0; }
                        else c.c = t.c + i0 + (long long)j0 * t.ldc;
                        c.c2 = t.c2 + j0 + (long long)i0 * t.ldc;
                        c.flags = (t.flags & ~(GF_BETA0 | GF_LOWER)) | GF_ATOMIC;
                        const int id = (int)(prog.gemm.size() - L.task0);
                        prog.gemm.push_back(c);
                        split.push_back(TileRef{id, 0, 0, 0});
                    }
                }
                prog.tiles.insert(prog.tiles.end(), keep.begin(), keep.end());
                prog.tiles.insert(prog.tiles.end(), split.begin(), split.end());
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
            // (GEMM steps: same variant and same lane; lane fork / join records: same kind of record and same event)
            auto subkey = [](const Step &st) {
                if (st.kind == LK_GEMM || st.kind == LK_GEMV || st.kind == LK_SYNC) return st.variant | (st.lane << 16);
                return 0;
            };
            std::map<std::pair<int, int>, int> votes;
            for (size_t i = 0; i < m; i++)
                if (head[i] < seq[i].size()) votes[{seq[i][head[i]].kind, subkey(seq[i][head[i]])}]++;
            std::pair<int, int> best{-1, -1};
            int bv = -1;
            for (auto &kv : votes) if (kv.second > bv) { bv = kv.second; best = kv.first; }
            std::vector<const Step *> chosen;
            for (size_t i = 0; i < m; i++) {
                if (head[i] >= seq[i].size()) continue;
                const Step &st = seq[i][head[i]];
                if (st.kind != best.first || subkey(st) != best.second) continue;
                chosen.push_back(&st);
                head[i]++;
                remaining--;
            }
            if (best.first == LK_GEMM) {
                const size_t from = prog.launches.size();
                emit_gemm_launch(best.second & 0xffff, chosen);
                for (size_t i = from; i < prog.launches.size(); i++) prog.launches[i].lane = best.second >> 16;
            } else if (best.first == LK_SYNC) {
                // the records are global orderings between the two lanes, so one record serves every front at this step
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_SYNC; L.variant = best.second & 0xffff; L.a0 = best.second >> 16;
                prog.launches.push_back(L);
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
            } else if (best.first == LK_COPY) {
                for (const Step *st : chosen) prog.launches.push_back(st->raw);
            } else if (best.first == LK_BCOPY) {
                Launch L;
                memset(&L, 0, sizeof L);
                L.kind = LK_BCOPY;
                L.task0 = (int64_t)prog.bcopy.size();
                L.tile0 = (int64_t)prog.tiles.size();
                for (const Step *st : chosen) {
                    const int id = (int)(prog.bcopy.size() - L.task0);
                    prog.bcopy.push_back(st->cp);
                    const int tr = (st->cp.rows + 511) / 512, tc = (st->cp.cols + 7) / 8;     // 512 x 8 patches
                    for (int tj = 0; tj < tc; tj++)
                        for (int ti = 0; ti < tr; ti++) prog.tiles.push_back(TileRef{id, ti, tj, 0});
                }
                L.ntasks = (int)(prog.bcopy.size() - L.task0);
                L.ntiles = (int)(prog.tiles.size() - L.tile0);
                if (L.ntiles > 0) prog.launches.push_back(L);
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

static inline void zero_launch(Program &p, int space, int64_t a0, int64_t a1)
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

// spaces: 0 L, 1 arena0, 2 arena1, 3 dinv, 4 X, 5 ybuf, 6 zarena0, 7 zarena1
enum { SP_L = 0, SP_AR0 = 1, SP_DINV = 3, SP_X = 4, SP_Y = 5, SP_Z0 = 6 };

// extend-add task of one child update matrix (compact nr x nr, leading dimension lds, at `src` of space
// `src_space`) into its parent's panel / update matrix; appends the task and its 32x32 lower tiles to launch L
static inline void push_ext_task(Program &P, Launch &L, long long src, int lds, int nr, int64_t rows_off, int64_t rel_base,
                                 const SNode &p, int src_space, int dst_space)
{
    if (nr == 0) return;
    ExtTask e;
    memset(&e, 0, sizeof e);
    e.src = src; e.lds = lds; e.nr = nr;
    e.ppanel = p.panel; e.pld = p.ld; e.pnc = p.nc; e.pncp = p.ncp;
    e.pupd = p.upd; e.pldu = p.ldu;
    e.rel = (int)(rel_base + rows_off);
    e.src_space = src_space; e.dst_space = dst_space;
    const int id = (int)(P.ext.size() - L.task0);
    P.ext.push_back(e);
    const int nt = (nr + 31) / 32;
    for (int tj = 0; tj < nt; tj++)
        for (int ti = tj; ti < nt; ti++) P.tiles.push_back(TileRef{id, ti, tj, 0});
}

// gather task of the selected inverse: the nr x nr block of the parent's front addressed by the relative
// indices -> block (ncp.., ncp..) of the matrix at `dst` (leading dimension ldd)
static inline void push_gather_task(Program &P, Launch &L, long long dst, int ldd, int ncp, int nr, long long src, int lds,
                                    int pnc, int pncp, int64_t rel, int src_space, int dst_space)
{
    if (nr == 0) return;
    GatherTask g;
    memset(&g, 0, sizeof g);
    g.dst = dst; g.ldd = ldd; g.ncp = ncp; g.nr = nr;
    g.src = src; g.lds = lds; g.pnc = pnc; g.pncp = pncp;
    g.rel = (int)rel;
    g.src_space = src_space; g.dst_space = dst_space;
    const int id = (int)(P.gather.size() - L.task0);
    P.gather.push_back(g);
    const int nt = (nr + 31) / 32;
    for (int tj = 0; tj < nt; tj++)
        for (int ti = tj; ti < nt; ti++) P.tiles.push_back(TileRef{id, ti, tj, 0});      // (tiles on and below the diagonal: the kernel mirrors)
}

// Dense partial Cholesky of one front (panel already assembled): left-looking GEMM inside an outer block of
// `outer` 64-column blocks, 64x64 POTRF with explicit inverse, TRSM-as-GEMM with that inverse, right-looking GEMM
// beyond the outer block, one SYRK of the update matrix with K = all pivot columns.
// `after_outer(q, P0, P1)` (optional) is called when the block columns P0..P1-1 hold their final values.
// `potrf_overlap`: the left-looking update of block column p is split into its diagonal tile (main lane, what POTRF
// needs) and the rows below (side lane), so that the 41 us POTRF of the diagonal block runs beside the update of the rows
// below instead of after it; the two meet again in front of the TRSM.  Same arithmetic, same results.
template <class Hook>
static inline void factor_node_steps(LevelBuilder &B, const SNode &x, int sp_u, int outer, std::vector<Step> &q, Hook after_outer,
                                     bool potrf_overlap = false)
{
    const int mrows = x.ncp + x.nr;     // panel rows in use (the gap row of an odd nc is zero)
    for (int P0 = 0; P0 < x.nblk; P0 += outer) {
        const int P1 = std::min(P0 + outer, x.nblk);
        const int cP = P0 * NB;
        for (int p = P0; p < P1; p++) {
            const int c0 = p * NB, b = std::min(NB, x.nc - c0);
            bool forked = false;
            if (p > P0) {
                // left-looking update of block column p from the inner blocks of this outer block
                const int K = c0 - cP;
                const int rb = (p == x.nblk - 1) ? x.ncp : c0 + NB;      // first row below the diagonal block (even: tile alignment)
                const int below = mrows - rb;
                if (potrf_overlap && below >= 4 * NB) {
                    B.add_sync(q, 0, 0);
                    B.add_gemm(q, B.task(SP_L, x.panel + rb + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + rb + (int64_t)c0 * x.ld, x.ld,
                                         below, b, K, GF_NEG), false, false);
                    q.back().lane = 1;
                    B.add_sync(q, 1, p & 1);
                    B.add_gemm(q, B.task(SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + c0 + (int64_t)c0 * x.ld, x.ld,
                                         b, b, K, GF_NEG), false, false);
                    forked = true;
                } else {
                    B.add_gemm(q, B.task(SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + c0 + (int64_t)cP * x.ld, x.ld,
                                         SP_L, x.panel + c0 + (int64_t)c0 * x.ld, x.ld,
                                         mrows - c0, b, K, GF_NEG), false, false);
                }
            }
            Step st;
            memset(&st, 0, sizeof st);
            st.kind = LK_POTRF;
            st.p.blk = x.panel + c0 + (int64_t)c0 * x.ld;
            st.p.dinv = x.dinv + (int64_t)p * NB * NB;
            st.p.ld = x.ld; st.p.b = b; st.p.col0 = x.first + c0;
            q.push_back(st);
            if (forked) B.add_sync(q, 2, p & 1);
            // rows below the diagonal block: L = A * W^T, in place
            const int r0 = (p == x.nblk - 1) ? x.ncp : c0 + NB;
            B.add_gemm(q, B.task(SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                 SP_DINV, st.p.dinv, NB,
                                 SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                 mrows - r0, b, b, GF_BETA0), false, false);
        }
        after_outer(q, P0, P1);
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
}

static inline void factor_node_steps(LevelBuilder &B, const SNode &x, int sp_u, int outer, std::vector<Step> &q, bool potrf_overlap = false)
{
    factor_node_steps(B, x, sp_u, outer, q, [](std::vector<Step> &, int, int) {}, potrf_overlap);
}

// Diagonal-first variant for fronts that keep the outer-block inverses (x.winv >= 0, in-core schedules).  Per outer block
// O of 512 columns: the chain of eight (left-looking update, POTRF, TRSM) triples runs on the w x w DIAGONAL block only --
// launches of a few tiles each instead of launches as tall as the front --, the inverse Wf = L_OO^-1 is built by recursive
// doubling (winv_doubling_steps; the Takahashi recursion and the solves reuse it), and the rows below get ONE product
// L[B,O] = A[B,O] Wf^T with N = 512 (computed by 64-column strips into scratch `Y` of the Y space, since strip j reads
// the strips <= j of A, then copied back).  Same flops as the blocked left-looking scheme; the dependent chain of a
// front loses the tall skinny launches.
static inline void factor_node_steps_diag(LevelBuilder &B, const SNode &x, int sp_u, int64_t Y, std::vector<Step> &q);

static inline Step copy_step(int variant, int64_t a0, int64_t a1, int64_t host_off)
{
    Step st;
    memset(&st, 0, sizeof st);
    st.kind = LK_COPY;
    st.raw.kind = LK_COPY; st.raw.variant = variant; st.raw.a0 = a0; st.raw.a1 = a1; st.raw.task0 = host_off;
    return st;
}

// Forward substitution L y = b on the columns of one supernode (X is kp x n, k-major, permuted order).
static inline void fsolve_node_steps(LevelBuilder &B, const SNode &x, int k, int kp, bool blocked, int OUTER, std::vector<Step> &q)
{
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
}

// Back substitution L^T x = y on the columns of one supernode.
static inline void bsolve_node_steps(LevelBuilder &B, const SNode &x, int k, int kp, bool blocked, int OUTER, std::vector<Step> &q)
{
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
}

// Takahashi recursion on one front whose trailing block Z[below,below] is in place: block columns from last to
// first.  Y = scratch of ld x 64 doubles at offset Y of the Y space.
// `before_block(q, p)` (optional) is called before the first step that reads block column p of the factor.
// One launch with the W^T W seeds of the diagonal blocks Z_pp of ALL block columns of the given fronts: they depend
// on nothing but the inverse diagonal blocks, and Z_pp is first read (by its own correction) at step p, so the whole
// level's seeds are hoisted in front of the recursion instead of one small launch per step.
static inline void wtw_level_launch(Program &P, const std::vector<const SNode *> &nodes, int sp_z)
{
    Launch L;
    memset(&L, 0, sizeof L);
    L.kind = LK_WTW;
    L.task0 = (int64_t)P.wtw.size();
    for (const SNode *x : nodes)
        for (int p = 0; p < x->nblk; p++) {
            const int c0 = p * NB;
            WtwTask w;
            memset(&w, 0, sizeof w);
            w.w = x->dinv + (int64_t)p * NB * NB;
            w.dst = x->front + c0 + (int64_t)c0 * x->ld;
            w.ldd = x->ld; w.b = std::min(NB, x->nc - c0); w.space = sp_z;
            P.wtw.push_back(w);
        }
    L.ntasks = (int)(P.wtw.size() - L.task0);
    if (L.ntasks > 0) P.launches.push_back(L);
}

template <class Hook>
static inline void selinv_node_steps(LevelBuilder &B, const SNode &x, int sp_z, int64_t Y, int splitk_min, int kchunk,
                                     std::vector<Step> &q, Hook before_block, bool wtw_hoisted = false)
{
    const int mrows = x.ncp + x.nr;
    const int64_t F = x.front;
    for (int p = x.nblk - 1; p >= 0; p--) {
        before_block(q, p);
        const int c0 = p * NB, b = std::min(NB, x.nc - c0);
        const int r0 = (p == x.nblk - 1) ? x.ncp : c0 + NB;
        const int mb = mrows - r0;
        const int64_t W = x.dinv + (int64_t)p * NB * NB;
        // Y = L[below,p] W is kept TRANSPOSED (Yt: b x mb, leading dimension 64), which makes it an operand with the tile
        // dimension contiguous in the product below (bulk-async kernel) -- a K-contiguous Y would be staged row by row
        // Z_pp = W^T W  (+ correction below)
        Step st;
        memset(&st, 0, sizeof st);
        st.kind = LK_WTW;
        st.w.w = W; st.w.dst = F + c0 + (int64_t)c0 * x.ld; st.w.ldd = x.ld; st.w.b = b; st.w.space = sp_z;
        if (!wtw_hoisted) q.push_back(st);
        if (mb <= 0) continue;
        // Yt = W^T * L[below,p]^T   (A(i,kk) = W(kk,i): K contiguous;  B(j,kk) = L(r0+j, c0+kk): tile dimension contiguous)
        B.add_gemm(q, B.task(SP_DINV, W, NB, SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                             SP_Y, Y, NB, b, mb, b, GF_BETA0), true, false);
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
            GemmTask t = B.task(sp_z, F + r0 + (int64_t)(r0 + k0) * x.ld, x.ld, SP_Y, Y + (int64_t)k0 * NB, NB,
                                sp_z, F + r0 + (int64_t)c0 * x.ld, x.ld, mb, b, std::min(clen, mb - k0),
                                GF_NEG | GF_UPPER_MIRROR | (nchunk == 1 ? (GF_BETA0 | GF_ZDEST) : GF_ATOMIC));
            t.c2 = F + c0 + (int64_t)r0 * x.ld;
            if (ci == 0) B.add_gemm(q, t, false, false);
            else B.join_gemm(q, t);
        }
        // Z_pp -= Y^T Z[below,p]   (split over K, accumulated atomically)
        const int chunk = 512;
        bool opened = false;
        for (int k0 = 0; k0 < mb; k0 += chunk) {
            GemmTask u = B.task(SP_Y, Y + (int64_t)k0 * NB, NB, sp_z, F + (r0 + k0) + (int64_t)c0 * x.ld, x.ld,
                                sp_z, F + c0 + (int64_t)c0 * x.ld, x.ld, b, b, std::min(chunk, mb - k0),
                                GF_NEG | GF_ATOMIC);
            if (!opened) { B.add_gemm(q, u, false, true, 2); opened = true; }
            else B.join_gemm(q, u);
        }
    }
}

static inline void selinv_node_steps(LevelBuilder &B, const SNode &x, int sp_z, int64_t Y, int splitk_min, int kchunk,
                                     std::vector<Step> &q, bool wtw_hoisted = false)
{
    selinv_node_steps(B, x, sp_z, Y, splitk_min, kchunk, q, [](std::vector<Step> &, int) {}, wtw_hoisted);
}

// ---------------------------------------------------------------------------------------------
// Two-level Takahashi recursion (fronts with more than one 64-column block, in-core schedules).
// The pivot columns are walked in OUTER blocks O of SEL_OUTER*64 = 512 columns, last to first.  With
// Wf = L_OO^-1 (the full inverse of the outer diagonal block, built by recursive doubling from the 64x64
// inverses the factorisation leaves in the inverse-block store) and B = the rows below O:
//     Yt     = Wf^T L[B,O]^T                       (w x mb, one strip of 64 rows per task: Wf is lower triangular)
//     Z[B,O] = -Z[B,B] Yt^T                        (mb x w, K = mb: the one big product, N = 512 instead of 64)
//     Z[O,O] = Wf^T Wf - Yt Z[B,O]                 (w x w)
// Same flops as the 64-column recursion (which spends 2 mb w^2 on the rows of O below each of its blocks), but the
// trailing block Z[B,B] is read once per 512 columns instead of once per 64 (8x less DRAM traffic on the dominant
// operand), the chain has 3 dependent launches per 512 columns instead of 24, and most products need no split-K atomics.
static inline int sel_outer_blocks()      // 64-column blocks per outer block (SPDE_SELINV_OUTER_BLOCKS: test hook for small meshes)
{
    return std::min(64, env_int("SPDE_SELINV_OUTER_BLOCKS", 8, 1));
}
#define SEL_OUTER (spde::sel_outer_blocks())
#define SEL_W (SEL_OUTER * NB)

struct OuterBlk { int c0, w, ldw, nb; int64_t wf; };
static inline int n_outer(const SNode &x) { return (x.nblk + SEL_OUTER - 1) / SEL_OUTER; }
static inline OuterBlk outer_block(const SNode &x, int k)
{
    OuterBlk o;
    o.c0 = k * SEL_W;
    o.w = std::min(SEL_W, x.nc - o.c0);
    o.ldw = up2(o.w);
    o.nb = (o.w + NB - 1) / NB;
    o.wf = x.winv + (int64_t)k * SEL_W * SEL_W;
    return o;
}
static inline int64_t winv_size(const SNode &x)
{
    if (x.nblk <= 1) return 0;
    const int no = n_outer(x);
    const int wl = x.nc - (no - 1) * SEL_W;
    return (int64_t)(no - 1) * SEL_W * SEL_W + (int64_t)up2(wl) * wl;
}
static inline int64_t ybuf_need(const SNode &x) { return (int64_t)x.ld * (x.nblk > 1 ? SEL_W : NB); }

// Copy task of one outer block: the 64x64 inverses W_j of its blocks onto the block diagonal of Wf (k_wtw, mode 1).
static inline WtwTask winv_copy_task(const SNode &x, int k)
{
    const OuterBlk o = outer_block(x, k);
    WtwTask w;
    memset(&w, 0, sizeof w);
    w.w = x.dinv + (int64_t)k * SEL_OUTER * NB * NB;
    w.dst = o.wf;
    w.ldd = o.ldw; w.b = o.w; w.space = SP_DINV;
    w.pad = 1;      // mode: plain copy of every W_j (not W^T W)
    return w;
}

// Doubling rounds inv([[A,0],[B,C]]) = [[A^-1,0],[-C^-1 B A^-1, C^-1]] of two grouped products each for the outer blocks
// k_lo..k_hi-1 of one front (block diagonal already copied), then the seeds Z[O,O] = Wf^T Wf.  `toff` = scratch in the Y
// space (<= 128 doubles per pivot column of the blocks treated).
static inline void winv_doubling_steps(LevelBuilder &B, const SNode &x, int k_lo, int k_hi, int64_t toff0, int sp_z, std::vector<Step> &q,
                                       bool doubling = true, bool seeds = true)
{
    if (doubling)
    for (int s = 1; s < SEL_OUTER; s *= 2) {
        // phase 0: T = L[C,A] Wf[A,A]      phase 1: Wf[C,A] = -Wf[C,C] T
        for (int phase = 0; phase < 2; phase++) {
            bool opened = false;
            int64_t toff = toff0;
            for (int k = k_lo; k < k_hi; k++) {
                const OuterBlk o = outer_block(x, k);
                for (int a0 = 0; a0 + s < o.nb; a0 += 2 * s) {
                    const int ra = a0 * NB, rc = (a0 + s) * NB;
                    const int sa = s * NB, sc = std::min(o.w, (a0 + 2 * s) * NB) - rc;
                    const int ldt = up2(sc);
                    GemmTask t;
                    if (phase == 0)
                        t = B.task(SP_L, x.panel + (o.c0 + rc) + (int64_t)(o.c0 + ra) * x.ld, x.ld,
                                   SP_DINV, o.wf + ra + (int64_t)ra * o.ldw, o.ldw,
                                   SP_Y, toff, ldt, sc, sa, sa, GF_BETA0);
                    else
                        t = B.task(SP_DINV, o.wf + rc + (int64_t)rc * o.ldw, o.ldw,
                                   SP_Y, toff, ldt,
                                   SP_DINV, o.wf + rc + (int64_t)ra * o.ldw, o.ldw, sc, sa, sc, GF_NEG | GF_BETA0);
                    toff += (int64_t)ldt * sa;
                    if (!opened) { B.add_gemm(q, t, false, true); opened = true; }
                    else B.join_gemm(q, t);
                }
            }
        }
    }
    for (int k = k_lo; seeds && k < k_hi; k++) {
        const OuterBlk o = outer_block(x, k);
        GemmTask t = B.task(SP_DINV, o.wf, o.ldw, SP_DINV, o.wf, o.ldw,
                            sp_z, x.front + o.c0 + (int64_t)o.c0 * x.ld, x.ld, o.w, o.w, o.w, GF_BETA0);
        if (k == k_lo) B.add_gemm(q, t, true, true);
        else B.join_gemm(q, t);
    }
}

static inline void factor_node_steps_diag(LevelBuilder &B, const SNode &x, int sp_u, int64_t Y, std::vector<Step> &q)
{
    const int mrows = x.ncp + x.nr;
    const int no = n_outer(x);
    for (int k = 0; k < no; k++) {
        const OuterBlk o = outer_block(x, k);
        const int c0 = o.c0, w = o.w, cE = c0 + w;
        for (int p = 0; p < o.nb; p++) {
            const int c = c0 + p * NB, b = std::min(NB, cE - c);
            if (p > 0)      // left-looking update of block column p, rows of the diagonal block only
                B.add_gemm(q, B.task(SP_L, x.panel + c + (int64_t)c0 * x.ld, x.ld,
                                     SP_L, x.panel + c + (int64_t)c0 * x.ld, x.ld,
                                     SP_L, x.panel + c + (int64_t)c * x.ld, x.ld,
                                     cE - c, b, c - c0, GF_NEG), false, false);
            Step st;
            memset(&st, 0, sizeof st);
            st.kind = LK_POTRF;
            st.p.blk = x.panel + c + (int64_t)c * x.ld;
            st.p.dinv = x.dinv + (int64_t)(k * SEL_OUTER + p) * NB * NB;
            st.p.ld = x.ld; st.p.b = b; st.p.col0 = x.first + c;
            q.push_back(st);
            if (cE - c - b > 0)      // rows of the diagonal block below: L = A W^T, in place
                B.add_gemm(q, B.task(SP_L, x.panel + (c + b) + (int64_t)c * x.ld, x.ld,
                                     SP_DINV, st.p.dinv, NB,
                                     SP_L, x.panel + (c + b) + (int64_t)c * x.ld, x.ld,
                                     cE - c - b, b, b, GF_BETA0), false, false);
        }
        // Wf = L_OO^-1
        {
            Step st;
            memset(&st, 0, sizeof st);
            st.kind = LK_WTW;
            st.w = winv_copy_task(x, k);
            q.push_back(st);
            winv_doubling_steps(B, x, k, k + 1, Y, 0, q, true, false);
        }
        const int r0 = (k == no - 1) ? x.ncp : cE;
        const int mb = mrows - r0;
        if (mb > 0) {
            // S = A[B,O] Wf^T by column strips (strip j reads the columns <= 64 (j+1) of A: Wf is lower triangular), then L[B,O] <- S
            const int ldS = up2(mb);
            for (int j = 0; j < o.nb; j++) {
                const int bj = std::min(NB, w - j * NB);
                GemmTask t = B.task(SP_L, x.panel + r0 + (int64_t)c0 * x.ld, x.ld,
                                    SP_DINV, o.wf + j * NB, o.ldw,
                                    SP_Y, Y + (int64_t)j * NB * ldS, ldS, mb, bj, j * NB + bj, GF_BETA0);
                if (j == 0) B.add_gemm(q, t, false, false);
                else B.join_gemm(q, t);
            }
            Step st;
            memset(&st, 0, sizeof st);
            st.kind = LK_BCOPY;
            st.cp.dst = x.panel + r0 + (int64_t)c0 * x.ld; st.cp.ldd = x.ld; st.cp.dst_space = SP_L;
            st.cp.src = Y; st.cp.lds = ldS; st.cp.src_space = SP_Y;
            st.cp.rows = mb; st.cp.cols = w;
            q.push_back(st);
        }
        if (k < no - 1) {
            // right-looking update of the panel columns beyond this outer block
            const int K = w;
            B.add_gemm(q, B.task(SP_L, x.panel + cE + (int64_t)c0 * x.ld, x.ld,
                                 SP_L, x.panel + cE + (int64_t)c0 * x.ld, x.ld,
                                 SP_L, x.panel + cE + (int64_t)cE * x.ld, x.ld,
                                 mrows - cE, x.nc - cE, K, GF_NEG | GF_LOWER), false, false);
        }
    }
    if (x.nr > 0)
        B.add_gemm(q, B.task(SP_L, x.panel + x.ncp, x.ld, SP_L, x.panel + x.ncp, x.ld,
                             sp_u, x.upd, x.ldu, x.nr, x.nr, x.nc, GF_NEG | GF_LOWER), false, false);
}

static inline int64_t winv_scratch(const SNode &x)      // doubling scratch of one front (doubles), see winv_doubling_steps
{
    return x.winv >= 0 ? (int64_t)n_outer(x) * ((int64_t)SEL_W * SEL_W / 4 + SEL_W) : 0;
}

// ---------------------------------------------------------------------------------------------
// Triangular solves on the outer-block inverses (in-core schedules of a plan whose factorisation leaves Wf behind).
// The right-hand sides ping-pong between two buffers, X (space SP_X) and X2 (space SP_X2, which the executor maps to a
// second k x n buffer for solve programs): a product with a 512 x 512 inverse cannot run in place, and with two buffers
// no step reads what a concurrent tile of the same launch writes.
//   forward  (L y = b):   b in X, y accumulates in X2 (zeroed by the program's first launch).  Per outer block O:
//                         X2_O += X_O Wf^T;  X_later -= X2_O L[later,O]^T;  at the end of the front the rows below
//                         receive X[rows] -= X2_front L[below,front]^T (atomic scatter, as before).
//   backward (L^T x = y): y in X2, x goes to X.  Per front: X2_front -= X[rows] L[below,front] (gathered columns of X);
//                         per outer block, last to first: X2_O -= X_later L[later,O];  X_O = X2_O Wf.
// Two dependent launches per 512 pivot columns instead of two per 64; single-block fronts use their 64 x 64 inverse the
// same way.  Works for both the matrix-vector kernel (k <= 4) and the tensor-core tiles.
enum { SP_X2 = 1 };
static inline void fsolve_node_steps_outer(LevelBuilder &B, const SNode &x, int k, int kp, std::vector<Step> &q)
{
    const int64_t xs = (int64_t)x.first * kp;
    const int no = x.winv >= 0 ? n_outer(x) : 1;
    for (int o = 0; o < no; o++) {
        int c0, w, ldw; int64_t wf;
        if (x.winv >= 0) { const OuterBlk ob = outer_block(x, o); c0 = ob.c0; w = ob.w; ldw = ob.ldw; wf = ob.wf; }
        else { c0 = 0; w = x.nc; ldw = NB; wf = x.dinv; }
        const int cE = c0 + w;
        // X2_O += X_O Wf^T      (B(j,kk) = Wf(j,kk): tile dimension contiguous)
        B.add_solve(q, B.task(SP_X, xs + (int64_t)c0 * kp, kp, SP_DINV, wf, ldw,
                              SP_X2, xs + (int64_t)c0 * kp, kp, k, w, w, 0), false);
        if (cE < x.nc)      // X_later -= X2_O L[later,O]^T
            B.add_solve(q, B.task(SP_X2, xs + (int64_t)c0 * kp, kp, SP_L, x.panel + cE + (int64_t)c0 * x.ld, x.ld,
                                  SP_X, xs + (int64_t)cE * kp, kp, k, x.nc - cE, w, GF_NEG), false);
    }
    if (x.nr > 0) {
        GemmTask t = B.task(SP_X2, xs, kp, SP_L, x.panel + x.ncp, x.ld, SP_X, 0, kp, k, x.nr, x.nc,
                            GF_NEG | GF_SCATTER_C | GF_ATOMIC);
        t.cidx = (int)x.rows;
        B.add_solve(q, t, false);
    }
}

static inline void bsolve_node_steps_outer(LevelBuilder &B, const SNode &x, int k, int kp, std::vector<Step> &q)
{
    const int64_t xs = (int64_t)x.first * kp;
    if (x.nr > 0) {
        // X2_front -= X[rows below] L[below,front]   (columns of X gathered through the row list)
        GemmTask t = B.task(SP_X, 0, kp, SP_L, x.panel + x.ncp, x.ld, SP_X2, xs, kp, k, x.nc, x.nr,
                            GF_NEG | GF_GATHER_A);
        t.aidx = (int)x.rows;
        B.add_solve(q, t, true);
    }
    const int no = x.winv >= 0 ? n_outer(x) : 1;
    for (int o = no - 1; o >= 0; o--) {
        int c0, w, ldw; int64_t wf;
        if (x.winv >= 0) { const OuterBlk ob = outer_block(x, o); c0 = ob.c0; w = ob.w; ldw = ob.ldw; wf = ob.wf; }
        else { c0 = 0; w = x.nc; ldw = NB; wf = x.dinv; }
        const int cE = c0 + w;
        if (cE < x.nc)      // X2_O -= X_later L[later,O]     (B(j,kk) = L(cE+kk, c0+j): K contiguous)
            B.add_solve(q, B.task(SP_X, xs + (int64_t)cE * kp, kp, SP_L, x.panel + cE + (int64_t)c0 * x.ld, x.ld,
                                  SP_X2, xs + (int64_t)c0 * kp, kp, k, w, x.nc - cE, GF_NEG), true);
        // X_O = X2_O Wf       (B(j,kk) = Wf(kk,j): K contiguous)
        B.add_solve(q, B.task(SP_X2, xs + (int64_t)c0 * kp, kp, SP_DINV, wf, ldw,
                              SP_X, xs + (int64_t)c0 * kp, kp, k, w, w, GF_BETA0), true);
    }
}

// Hoisted, once per level: Wf of every outer block of the given fronts and the seeds Z[O,O] = Wf^T Wf, in seven grouped
// launches.  `yoff[i]` = scratch of front i in the Y space.
static inline void winv_level_launches(Program &P, const std::vector<const SNode *> &nodes, const std::vector<int64_t> &yoff, int sp_z,
                                       bool from_factor = false, bool seeds = true)
{
    if (!from_factor) {
        Launch L;
        memset(&L, 0, sizeof L);
        L.kind = LK_WTW;
        L.task0 = (int64_t)P.wtw.size();
        for (const SNode *xp : nodes)
            for (int k = 0; k < n_outer(*xp); k++) P.wtw.push_back(winv_copy_task(*xp, k));
        L.ntasks = (int)(P.wtw.size() - L.task0);
        if (L.ntasks > 0) P.launches.push_back(L);
    }
    LevelBuilder B(P);
    for (size_t fi = 0; fi < nodes.size(); fi++) {
        std::vector<Step> q;
        winv_doubling_steps(B, *nodes[fi], 0, n_outer(*nodes[fi]), yoff[fi], sp_z, q, !from_factor, seeds);
        B.seq.push_back(std::move(q));
    }
    B.flush();
}

// The recursion itself on one front (winv >= 0) whose trailing block Z[below,below] is in place and whose seeds are set.
// `before_outer(q, k)` (optional) is called before the first step that reads outer block k of the factor; with
// `inline_winv` the inverse Wf and the seed of an outer block are built right there, in front of its three products
// (streamed evaluator: the panel arrives from the host outer block by outer block), instead of hoisted per level.
template <class Hook>
static inline void selinv_node_steps_outer(LevelBuilder &B, const SNode &x, int sp_z, int64_t Y, int kchunk, std::vector<Step> &q,
                                           Hook before_outer, bool inline_winv)
{
    const int mrows = x.ncp + x.nr;
    const int64_t F = x.front;
    const int no = n_outer(x);
    for (int k = no - 1; k >= 0; k--) {
        const OuterBlk o = outer_block(x, k);
        const int c0 = o.c0, w = o.w, ldY = o.ldw;
        const int r0 = (k == no - 1) ? x.ncp : c0 + w;
        const int mb = mrows - r0;
        before_outer(q, k);
        if (inline_winv) {
            Step st;
            memset(&st, 0, sizeof st);
            st.kind = LK_WTW;
            st.w = winv_copy_task(x, k);
            q.push_back(st);
            winv_doubling_steps(B, x, k, k + 1, Y, sp_z, q);
        }
        if (mb <= 0) continue;
        // Yt = Wf^T L[B,O]^T, one task per strip of 64 rows (rows of strip i only see the blocks i.. of O)
        for (int i = 0; i < o.nb; i++) {
            const int bi = std::min(NB, w - i * NB);
            GemmTask t = B.task(SP_DINV, o.wf + i * NB + (int64_t)i * NB * o.ldw, o.ldw,
                                SP_L, x.panel + r0 + (int64_t)(c0 + i * NB) * x.ld, x.ld,
                                SP_Y, Y + i * NB, ldY, bi, mb, w - i * NB, GF_BETA0);
            if (i == 0) B.add_gemm(q, t, true, false);
            else B.join_gemm(q, t);
        }
        // Z[B,O] = -Z[B,B] Yt^T (and its transpose into the row block).  K is cut into chunks (atomic accumulation into the
        // zeroed block) when the launch would have too few tiles for the machine, and to bound the duration of one tile:
        // a launch ends with a partially filled wave, and a 128x64 tile with K = mb in the thousands runs for a millisecond
        const int tiles0 = ((mb + 127) / 128) * ((w + NB - 1) / NB);
        int nchunk = 1;
        if (tiles0 < 4 * kSMs) nchunk = std::max(1, std::min((4 * kSMs + tiles0 - 1) / tiles0, mb / 256));
        if (kchunk > 0 && mb > kchunk + kchunk / 2) nchunk = std::max(nchunk, (mb + kchunk - 1) / kchunk);
        int clen = (mb + nchunk - 1) / nchunk;
        clen += clen & 1;
        for (int k0 = 0, ci = 0; k0 < mb; k0 += clen, ci++) {
            GemmTask t = B.task(sp_z, F + r0 + (int64_t)(r0 + k0) * x.ld, x.ld, SP_Y, Y + (int64_t)k0 * ldY, ldY,
                                sp_z, F + r0 + (int64_t)c0 * x.ld, x.ld, mb, w, std::min(clen, mb - k0),
                                GF_NEG | GF_UPPER_MIRROR | (nchunk == 1 ? (GF_BETA0 | GF_ZDEST) : GF_ATOMIC));
            t.c2 = F + c0 + (int64_t)r0 * x.ld;
            if (ci == 0) B.add_gemm(q, t, false, false);
            else B.join_gemm(q, t);
        }
        // Z[O,O] -= Yt Z[B,O]   (split over K, accumulated atomically onto the seed)
        const int chunk = 512;
        bool opened = false;
        for (int k0 = 0; k0 < mb; k0 += chunk) {
            GemmTask u = B.task(SP_Y, Y + (int64_t)k0 * ldY, ldY, sp_z, F + (r0 + k0) + (int64_t)c0 * x.ld, x.ld,
                                sp_z, F + c0 + (int64_t)c0 * x.ld, x.ld, w, w, std::min(chunk, mb - k0),
                                GF_NEG | GF_ATOMIC);
            if (!opened) { B.add_gemm(q, u, false, true, 2); opened = true; }
            else B.join_gemm(q, u);
        }
    }
}

static inline void selinv_node_steps_outer(LevelBuilder &B, const SNode &x, int sp_z, int64_t Y, int kchunk, std::vector<Step> &q)
{
    selinv_node_steps_outer(B, x, sp_z, Y, kchunk, q, [](std::vector<Step> &, int) {}, false);
}

}  // namespace spde
