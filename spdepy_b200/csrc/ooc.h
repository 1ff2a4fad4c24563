// ooc.h -- streamed ("out-of-core") evaluation of the supernodal multifrontal method for meshes whose factor,
// update matrices and inverse fronts do not fit in HBM together (BASELINE configs[3], 256x256x100: 260 GB of L).
//
// The supernodal tree is cut into SEGMENTS: the supernodes whose subtree holds more than a threshold of factor
// bytes are "top" segments of one supernode each, processed front by front; every maximal subtree below them is a
// "bottom" segment, processed with the level-grouped schedules of the in-core path.  Segments run depth-first
// (children before parents, in Liu's order) on ONE statically planned device pool:
//   * bottom of the pool: a stack of compact update matrices (forward pass) / compact inverse blocks Z_RR
//     (backward pass) that cross segment boundaries;
//   * top of the pool: the working set of the segment in flight (its factor panels, inverse diagonal blocks,
//     ping-pong arenas; in the backward pass also its inverse fronts and the Y scratch).
// Every address is known when the plan is built, so a segment's schedule is an ordinary Program with absolute
// pool offsets and all eight operand spaces (except X) alias the pool.
// Forward pass (factorise): scatter Q, factor, log-determinant share, forward substitution; the panels of a top
// segment are then spilled to pinned host memory (only if a backward pass follows; slice by slice on a copy stream
// while the factorisation of the same front goes on, LK_COPY records in its schedule), those of a bottom segment are
// dropped.  Backward pass (back substitution + Takahashi selected inverse), segments in exactly the reverse order:
// a top segment's panels come back from the host, a bottom segment is factorised again (its subtree does not
// depend on anything above it).  The last forward segment (the root) stays on the device between the passes.
#pragma once
#include <map>
#include <vector>

#include "plan.h"

namespace spde {

struct ScatEntry { long long src, dst; };   // Q slot array index (slot * n + node) -> pool offset

struct OocSeg {
    std::vector<int> nodes;                   // supernodes, ascending (a contiguous postorder range)
    int root = -1, parent_seg = -1, dmin = 0, dmax = 0, top = 0;
    std::vector<int> kids;                    // child segments in forward processing order
    std::vector<std::vector<int>> by_depth;   // [d - dmin]
    // sizes in doubles
    int64_t l_size = 0, dinv_size = 0, arena[2] = {0, 0}, zarena[2] = {0, 0}, ybuf = 0, u_size = 0;
    // static memory plan: pool offsets in doubles
    int64_t off_L = 0, off_dinv = 0, off_ar[2] = {0, 0}, off_z[2] = {0, 0}, off_y = 0;
    int64_t stack_U = -1, stack_Z = -1;
    int64_t host_off = -1;                    // [dinv | L] in the pinned host pool, -1: not spilled
    // overlapped panel traffic of a spilled top segment: the panel in slices of one outer block of columns
    // (pool offset, doubles); slice c goes to the host as soon as the factorisation has finished its columns and
    // comes back, last slice first, while the Takahashi recursion already works on the slices behind it
    int overlap = 0;
    std::vector<int64_t> chunk_off, chunk_len;
    int keep = 0;                             // factor stays on the device between the passes
    int64_t scat0 = 0, scat1 = 0, zent0 = 0, zent1 = 0;
    int col0 = 0, col1 = 0;                   // column range (new ordering)
    double flops = 0;                         // sum cc^2 of the segment's columns
    Program factor, selinv;
    std::map<int, Program> fsolve, bsolve;
};

struct Ooc {
    Plan *plan = nullptr;
    std::vector<SNode> osn;                   // node records with absolute pool offsets
    std::vector<int> seg_of;                  // supernode -> segment
    std::vector<int64_t> yoff_of;             // supernode -> its scratch in the Y region (Yt, then the outer-block inverses)
    std::vector<OocSeg> segs;
    std::vector<int> order;                   // forward processing order (backward = reverse)
    int64_t pool_size = 0, peak_fwd = 0, peak_bwd = 0, host_size = 0, stage_bytes = 0;
    double recompute_flops = 0;
    bool built = false;
    std::vector<ScatEntry> scat;              // grouped by segment
    std::vector<long long> diagpos;           // per column (new ordering), pool offsets
    std::vector<ZEntry> zent;                 // grouped by (segment, depth)
    std::vector<std::vector<int64_t>> zptr;   // per segment: [d - dmin] -> range start, last = end
    int64_t ident_base = 0;                   // offset of the identity index run inside the index array
    // device state
    double *d_pool = nullptr, *h_pool = nullptr, *d_Xp = nullptr, *d_ld = nullptr, *d_red = nullptr;
    int64_t xp_cap = 0;
    cudaStream_t copy_stream = nullptr;       // panel traffic of the overlapped top segments
    cudaEvent_t copy_fork = nullptr, copy_done = nullptr;
    std::vector<cudaEvent_t> fetch_ev;        // slice c of the segment in flight has come back from the host
    char *d_stage = nullptr;
    int *d_idx = nullptr, *d_perm = nullptr, *d_status = nullptr;
    long long *d_diag = nullptr;
    double last_ms[2] = {0, 0};               // device time of the forward / backward pass of the last run
    std::vector<double> last_ld;              // log-determinant share of every segment in the last run

    void segment(int64_t top_bytes);
    void plan_memory(bool backward);
    void build_tables();
    void build_programs();
    Program &solve_program(OocSeg &g, int k, int dir);
};

}  // namespace spde
