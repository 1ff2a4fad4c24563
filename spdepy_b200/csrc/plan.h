// plan.h -- numeric plan of the supernodal multifrontal method: storage layout, per-level kernel
// task lists ("programs") for factorisation, triangular solves and the selected inverse.
// Built on the host once per mesh from the Symbolic analysis; executed by plan.cu.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "gemm.cuh"
#include "symbolic.h"

namespace spde {

constexpr int NB = 64;        // diagonal-block size of the dense partial Cholesky
constexpr int OUTER = 8;      // inner blocks per outer (right-looking) block -> 512 columns

struct SNode {
    int first, nc, nr, ncp, ld, nblk, depth, parent, ldu;
    int64_t panel;     // offset of the m x nc panel in the L store (ld = ncp + roundup2(nr))
    int64_t dinv;      // offset of the nblk 64x64 inverse diagonal blocks
    int64_t upd;       // offset of the nr x nr update matrix inside arena[depth & 1]
    int64_t front;     // offset of the ld x ld selected-inverse front inside zarena[depth & 1]
    int64_t rows;      // offset into the device index array of the structure rows
    int64_t winv;      // offset (inverse-block store) of the full inverses of the 512-column outer diagonal blocks
                       // used by the Takahashi recursion of fronts with more than one 64-column block; -1 = none
};

struct PotrfTask { long long blk; long long dinv; int ld, b, col0, pad; };
// extend-add of one child update matrix into its parent's front
struct ExtTask {
    long long src; int lds, nr;            // child update matrix (arena of the child's parity)
    long long ppanel; int pld, pnc, pncp;  // parent panel
    long long pupd; int pldu;              // parent update matrix
    int rel;                               // offset of the child's relative indices
    int src_space, dst_space;
};
struct GatherTask {   // selected inverse: child's trailing block <- parent's front
    long long dst; int ldd, ncp, nr;       // child front (zarena of child's parity), trailing block at (ncp,ncp)
    long long src; int lds, pnc, pncp;     // parent front
    int rel;
    int src_space, dst_space;
};
struct WtwTask { long long w; long long dst; int ldd, b, space, pad; };
// rectangular block copy dst(rows x cols, ldd) <- src(rows x cols, lds): the outer-block TRSM of the factorisation is
// computed out of place (scratch in the Y space) and copied back into the panel
struct CopyTask { long long dst, src; int ldd, lds, rows, cols, dst_space, src_space; };

enum LaunchKind : int { LK_GEMM = 0, LK_POTRF, LK_EXTADD, LK_ZERO, LK_GATHER, LK_WTW, LK_EXTRACT, LK_GEMV, LK_SYNC, LK_COPY, LK_BCOPY };
// LK_SYNC (two-lane schedules): variant 0 = the bulk lane waits for everything issued so far on the main lane,
// 1 = record bulk-lane event a0, 2 = the main lane waits for bulk-lane event a0.  The launch list is always a valid
// serial order, so an executor may ignore the lanes (profiling mode, the NumPy interpreter).
// LK_COPY (streamed evaluator only): variant 0 = park a0..a0+a1 doubles of the pool at host offset task0 (on the copy
// stream, after everything issued so far); variant 1 = wait until chunk a0 of the segment's panel has come back from the host.

struct Launch {
    int kind, variant;
    int64_t task0; int ntasks;
    int64_t tile0; int ntiles;
    int64_t a0, a1;   // LK_ZERO: [a0,a1) doubles of space `variant`; LK_EXTRACT: entry range
    int lane, pad;    // 0 = main lane, 1 = bulk lane (trailing updates that overlap the next panel)
};

struct Program {
    std::vector<Launch> launches;
    std::vector<GemmTask> gemm;
    std::vector<TileRef> tiles;
    std::vector<PotrfTask> potrf;
    std::vector<ExtTask> ext;
    std::vector<GatherTask> gather;
    std::vector<WtwTask> wtw;
    std::vector<CopyTask> bcopy;
    // device copies
    GemmTask *d_gemm = nullptr; TileRef *d_tiles = nullptr; PotrfTask *d_potrf = nullptr;
    ExtTask *d_ext = nullptr; GatherTask *d_gather = nullptr; WtwTask *d_wtw = nullptr; CopyTask *d_bcopy = nullptr;
    bool uploaded = false;
    double flops = 0;
    // CUDA graph of the whole schedule, per factor store; re-captured when a base pointer changes
    cudaGraphExec_t graph[2] = {nullptr, nullptr};
    unsigned long long graph_key[2] = {0, 0};
    int runs[2] = {0, 0};
    std::vector<float> last_ms;   // per-launch device time of the last profiled run
};

struct ZEntry { long long dst, dst2; long long src; int sn, pad; };   // selected-inverse extraction

struct Plan {
    Symbolic sym;
    std::vector<SNode> sn;
    std::vector<std::vector<int>> by_depth;
    int64_t l_size = 0, dinv_size = 0, arena_size[2] = {0, 0}, zarena_size[2] = {0, 0}, ybuf_size = 0;
    int max_rhs = 0;
    // scatter map Q slots -> L store
    std::vector<int> cand_slots;           // slots that can hold a lower-triangle entry
    std::vector<long long> qdest;          // cand x n, -1 = unused
    std::vector<long long> diagpos;        // n, position of L(j,j) in the L store (new order)
    std::vector<ZEntry> zentries;          // grouped by depth
    std::vector<int64_t> zdepth_ptr;
    Program factor;
    std::map<std::pair<int, int>, Program> solve;   // (k, direction) -> program
    Program selinv;
    bool selinv_built = false;
    bool winv_from_factor = false;   // the factorisation leaves the outer-block inverses Wf behind (diagonal-first panels) ...
    int diag_min_ld = 0;             // ... on the fronts with at least this many rows
    bool solve_outer = false;        // the factor schedule ends by building Wf of every multi-block front, and the solve
                                     // schedules work on outer blocks with two right-hand-side buffers (plan_steps.h)
    bool diag_front(const SNode &x) const { return winv_from_factor && x.winv >= 0 && x.ld >= diag_min_ld; }
    // device state
    int device_ready = 0;
    double *d_L[2] = {nullptr, nullptr};
    double *d_dinv[2] = {nullptr, nullptr};
    // per factor store (so the schedules of Q and Q + tau S^T S can run concurrently on two streams)
    double *d_arena[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    double *d_zarena[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    double *d_ybuf[2] = {nullptr, nullptr};
    double *d_X = nullptr, *d_X2 = nullptr, *d_red = nullptr;
    int64_t x_cap = 0;
    int *d_idx = nullptr;          // rows | relidx
    int64_t rel_base = 0;
    long long *d_qdest = nullptr, *d_diagpos = nullptr;
    int *d_cand = nullptr, *d_perm = nullptr, *d_status = nullptr;
    ZEntry *d_zentries = nullptr;
    int status[2] = {0, 0}, bad_col[2] = {-1, -1};
    bool factored[2] = {false, false};
    // graph replay machinery: schedules are captured on a private stream and ordered against the caller's
    // stream with two events (the caller's stream may be the legacy default stream, which cannot capture)
    int use_graphs = 1;
    cudaStream_t cap_stream[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool pending[2] = {false, false};   // a graph of this store is in flight and not yet joined with the caller's stream
    bool pending_readonly[2] = {false, false};   // ... and it only READS the factor (Takahashi): solves may run beside it
    // second lane for the triangular solves, so that the latency-bound k = 1 solve of the conditional mean overlaps
    // the Takahashi pass of the same store
    cudaStream_t bulk_stream[2] = {nullptr, nullptr};        // second lane of the factor schedule (look-ahead)
    cudaEvent_t bulk_fork[2] = {nullptr, nullptr}, bulk_ev[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int use_lanes = 1;
    cudaStream_t solve_stream[2] = {nullptr, nullptr};
    cudaEvent_t sev_in[2] = {nullptr, nullptr}, sev_out[2] = {nullptr, nullptr};
    double *d_zq[2] = {nullptr, nullptr};
    int sel_ready[2] = {0, 0};
    // optional per-launch timing (CUDA events), accumulated per (launch kind, GEMM variant)
    bool prof_on = false;
    double prof_ms[8][16] = {{0}};
    double prof_cnt[8][16] = {{0}};

    void build_layout();
    void build_factor_program();
    Program &solve_program(int k, int dir);
    void build_selinv_program();
};

// explicit execution context of a program: base pointers of the eight operand spaces, the factor / inverse-block
// stores the POTRF and W^T W kernels address directly, the status word and the extraction table
struct ExecCtx {
    GemmSpaces sp;
    double *L = nullptr, *dinv = nullptr;
    int *status = nullptr;
    const ZEntry *zent = nullptr;
    double *Zq = nullptr;
    int which = 0;
    bool lanes = false;
    // streamed evaluator: host pool, copy stream and events of the panel traffic (LK_COPY)
    double *h_pool = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t copy_fork = nullptr;
    const std::vector<cudaEvent_t> *fetch_ev = nullptr;
};
int issue_program_ex(Plan &p, Program &P, const ExecCtx &ctx, cudaStream_t st);
int launch_logdet(const double *d_L, const long long *d_diagpos, int n, double *d_partial, double *d_out, cudaStream_t st);
int launch_perm_in(const double *d_X, const int *d_perm, int n, int k, int kp, int use_perm, double *d_Xp, cudaStream_t st);
int launch_perm_out(const double *d_Xp, const int *d_perm, int n, int k, int kp, int use_perm, double *d_X, cudaStream_t st);
void init_exec_env(Plan &p);

}  // namespace spde
