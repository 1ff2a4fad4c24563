// gemm.cuh -- grouped FP64 tensor-core GEMM: the one dense contraction on the hot path.
//
// Blackwell's tcgen05/UMMA has no FP64 kind (SURVEY.md finding 10); FP64 tensor-core math on
// sm_100a is `mma.sync.aligned.m8n8k4.f64` (SASS DMMA.8x8x4).  Every dense block operation of the
// supernodal factorisation, the triangular solves and the Takahashi recursion is expressed as a
// *task* `C (op)= +-A*B` on column-major operands and executed by this kernel:
//   - one launch covers all tasks of a level (grouped GEMM): grid = total number of C tiles,
//     a host-built tile table maps blockIdx.x -> (task, tile row, tile col);
//   - operand tiles are staged global -> shared with 16-byte cp.async (LDGSTS) in a 3-stage
//     ring; the +4 double row padding makes the DMMA fragment loads bank-conflict free;
//   - each warp owns a WMxWN block of 8x8 accumulator tiles held in registers.
// Variants (template): A/B stored with the tile dimension contiguous ("N") or with K contiguous
// ("T"); runtime flags select sign, overwrite, lower-triangle masking, a column gather on A
// (back substitution reads scattered rows of X) and an atomic column scatter on C (forward
// substitution updates scattered rows of X).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spde {

enum : int {
    GF_LOWER = 1 << 9,      // only elements with (row >= col) of C are written, tiles above skipped
    GF_BETA0 = 1 << 10,     // C = +-A*B instead of C += +-A*B
    GF_NEG = 1 << 11,       // C -= A*B
    GF_ATOMIC = 1 << 12,    // atomicAdd into C (several tasks may hit the same column)
    GF_GATHER_A = 1 << 13,  // column kk of A is column idx[aidx+kk] of the A space
    GF_SCATTER_C = 1 << 14, // column j of C is column idx[cidx+j] of the C space
    GF_UPPER_MIRROR = 1 << 15, // also write the transposed block at c2 (keeps selected-inverse fronts symmetric)
    GF_ZDEST = 1 << 16,     // schedule hint (ignored by the kernels): C holds zeros before this launch, so a BETA0 task may be
                            // cut along K into chunks that accumulate atomically
};

struct GemmTask {
    long long a, b, c;   // element offsets into the operand spaces
    long long c2;        // origin of the mirrored block (GF_UPPER_MIRROR): element (r,c) also goes to c2 + c + r*ldc
    int lda, ldb, ldc;
    int M, N, K;
    int flags;           // bits 0-2 A space, 3-5 B space, 6-8 C space, then GF_*
    int aidx, cidx;      // offsets into the index array
    int pad;
};

struct GemmSpaces {
    double *base[8];
    const int *idx;
};

struct TileRef { int task, ti, tj, pad; };

// Programmatic dependent launch (sm_90+): every kernel of a schedule lets its successor start launching
// at once and then waits for its predecessor to complete and flush.  The wait is unconditional and first,
// so completion order stays transitive along the chain (a grid never completes before its predecessor).
__device__ __forceinline__ void pdl_enter()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

constexpr int GEMM_BK = 16;
constexpr int GEMM_STAGES = 3;

template <int BM, int BN, int BKT = GEMM_BK, int STG = GEMM_STAGES>
constexpr int gemm_smem_bytes()
{
    // worst case of the two layouts per operand
    constexpr int a = (BM + 4) * BKT > BM * (BKT + 4) ? (BM + 4) * BKT : BM * (BKT + 4);
    constexpr int b = (BN + 4) * BKT > BN * (BKT + 4) ? (BN + 4) * BKT : BN * (BKT + 4);
    constexpr int pipe = STG * (a + b) * 8;
    constexpr int epi = (BM + 2) * BN * 8;       // accumulator tile staged for the coalesced epilogue
    return pipe > epi ? pipe : epi;
}

// Load one BT x BK operand tile into shared memory.
//  KMAJ=false: element (i,kk) at g[i + col(kk)*ld]   -> smem[kk][BT+4]
//  KMAJ=true : element (i,kk) at g[kk + i*ld]        -> smem[i][BK+4]
template <int BT, bool KMAJ, int NT, int BKT>
__device__ __forceinline__ void load_tile(double *smem, const double *g, int ld, int rows, int k0, int K,
                                          const int *gather, int tid)
{
    if (!KMAJ) {
        constexpr int CH = BT / 2;               // 16-byte chunks per k-column
        for (int id = tid; id < CH * BKT; id += NT) {
            const int kk = id / CH, ic = (id % CH) * 2;
            const int kg = k0 + kk;
            int bytes = 0;
            const double *src = g;
            if (kg < K && ic < rows) {
                const long long col = gather ? (long long)gather[kg] : (long long)kg;
                src = g + ic + col * ld;
                bytes = (rows - ic >= 2) ? 16 : 8;
            }
            cp_async16(smem + kk * (BT + 4) + ic, src, bytes);
        }
    } else {
        constexpr int CH = BKT / 2;          // 16-byte chunks per row
        for (int id = tid; id < CH * BT; id += NT) {
            const int i = id / CH, kc = (id % CH) * 2;
            const int kg = k0 + kc;
            int bytes = 0;
            const double *src = g;
            if (i < rows && kg < K) {
                src = g + kg + (long long)i * ld;
                bytes = (K - kg >= 2) ? 16 : 8;
            }
            cp_async16(smem + i * (BKT + 4) + kc, src, bytes);
        }
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, bool A_KMAJ, bool B_KMAJ, int BKT = GEMM_BK, int STG = GEMM_STAGES>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32)
k_gemm_grouped(const GemmTask *__restrict__ tasks, const TileRef *__restrict__ tiles, GemmSpaces sp)
{
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N;
    constexpr int MT = WM / 8, NTL = WN / 8;
    constexpr int A_ELEMS = A_KMAJ ? BM * (BKT + 4) : (BM + 4) * BKT;
    constexpr int B_ELEMS = B_KMAJ ? BN * (BKT + 4) : (BN + 4) * BKT;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem;
    double *sB = smem + STG * A_ELEMS;

    pdl_enter();
    const TileRef tr = tiles[blockIdx.x];
    const GemmTask tk = tasks[tr.task];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int i0 = tr.ti * BM, j0 = tr.tj * BN;
    const int rowsA = min(BM, tk.M - i0), rowsB = min(BN, tk.N - j0);

    const double *gA = sp.base[tk.flags & 7] + tk.a + (A_KMAJ ? (long long)i0 * tk.lda : (long long)i0);
    const double *gB = sp.base[(tk.flags >> 3) & 7] + tk.b + (B_KMAJ ? (long long)j0 * tk.ldb : (long long)j0);
    const int *gather = (tk.flags & GF_GATHER_A) ? sp.idx + tk.aidx : nullptr;

    double acc[MT][NTL][2];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }

    const int nk = (tk.K + BKT - 1) / BKT;
#pragma unroll
    for (int s = 0; s < STG - 1; s++) {
        if (s < nk) {
            load_tile<BM, A_KMAJ, NT, BKT>(sA + s * A_ELEMS, gA, tk.lda, rowsA, s * BKT, tk.K, gather, tid);
            load_tile<BN, B_KMAJ, NT, BKT>(sB + s * B_ELEMS, gB, tk.ldb, rowsB, s * BKT, tk.K, nullptr, tid);
        }
        cp_async_commit();
    }
    const int lr = lane >> 2, lc = lane & 3;
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STG - 2>();
        __syncthreads();
        {   // prefetch stage kt + STAGES-1 into the slot freed by iteration kt-1
            const int nx = kt + STG - 1;
            if (nx < nk) {
                const int s = nx % STG;
                load_tile<BM, A_KMAJ, NT, BKT>(sA + s * A_ELEMS, gA, tk.lda, rowsA, nx * BKT, tk.K, gather, tid);
                load_tile<BN, B_KMAJ, NT, BKT>(sB + s * B_ELEMS, gB, tk.ldb, rowsB, nx * BKT, tk.K, nullptr, tid);
            }
            cp_async_commit();
        }
        const double *cA = sA + (kt % STG) * A_ELEMS;
        const double *cB = sB + (kt % STG) * B_ELEMS;
#pragma unroll
        for (int k4 = 0; k4 < BKT; k4 += 4) {
            double fa[MT], fb[NTL];
#pragma unroll
            for (int a = 0; a < MT; a++) {
                const int i = wm * WM + a * 8 + lr;
                fa[a] = A_KMAJ ? cA[i * (BKT + 4) + k4 + lc] : cA[(k4 + lc) * (BM + 4) + i];
            }
#pragma unroll
            for (int b = 0; b < NTL; b++) {
                const int j = wn * WN + b * 8 + lr;
                fb[b] = B_KMAJ ? cB[j * (BKT + 4) + k4 + lc] : cB[(k4 + lc) * (BN + 4) + j];
            }
#pragma unroll
            for (int a = 0; a < MT; a++)
#pragma unroll
                for (int b = 0; b < NTL; b++) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
    }
    cp_async_wait<0>();

    // epilogue: accumulators -> shared memory (column-major tile, LD = BM+2 keeps the stores conflict
    // free) -> coalesced 16-byte read-modify-write of C.  Each thread owns C[row = lane/4][cols 2*(lane%4), +1]
    // of every 8x8 DMMA tile, which would be 8-byte strided accesses if written straight to global memory.
    constexpr int LDS = BM + 2;
    double *sC = smem;
    __syncthreads();      // every warp is done with the operand stages
    const bool neg = tk.flags & GF_NEG, beta0 = tk.flags & GF_BETA0, lower = tk.flags & GF_LOWER;
    const bool atomic = tk.flags & GF_ATOMIC, mirror = tk.flags & GF_UPPER_MIRROR;
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const double v = neg ? -acc[a][b][e] : acc[a][b][e];
                sC[(wn * WN + b * 8 + lc * 2 + e) * LDS + (wm * WM + a * 8 + lr)] = v;
            }
    __syncthreads();
    double *gC = sp.base[(tk.flags >> 6) & 7] + tk.c;
    const int *scat = (tk.flags & GF_SCATTER_C) ? sp.idx + tk.cidx : nullptr;
    const int mrows = min(BM, tk.M - i0), ncols = min(BN, tk.N - j0);
    constexpr int RP = BM / 2;                 // row pairs per column
    for (int id = tid; id < RP * BN; id += NT) {
        const int cl = id / RP, rl = (id % RP) * 2;
        if (cl >= ncols || rl >= mrows) continue;
        const int r = i0 + rl, c = j0 + cl;
        const bool two = rl + 1 < mrows;
        bool w0 = true, w1 = two;
        if (lower) { w0 = r >= c; w1 = two && (r + 1 >= c); }
        if (!w0 && !w1) continue;
        const double v0 = sC[cl * LDS + rl], v1 = sC[cl * LDS + rl + 1];
        const long long col = scat ? (long long)scat[c] : (long long)c;
        double *p = gC + r + col * tk.ldc;
        if (atomic) {
            if (w0) atomicAdd(p, v0);
            if (w1) atomicAdd(p + 1, v1);
        } else if (w0 && w1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            double2 o = beta0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2 *>(p);
            o.x += v0; o.y += v1;
            *reinterpret_cast<double2 *>(p) = o;
        } else {
            if (w0) p[0] = beta0 ? v0 : p[0] + v0;
            if (w1) p[1] = beta0 ? v1 : p[1] + v1;
        }
    }
    if (mirror) {
        // transposed copy: element (r,c) also goes to c2 + c + r*ldc, contiguous in c
        double *gM = sp.base[(tk.flags >> 6) & 7] + tk.c2;
        constexpr int CP = BN / 2;
        for (int id = tid; id < CP * BM; id += NT) {
            const int rl = id / CP, cl = (id % CP) * 2;
            if (rl >= mrows || cl >= ncols) continue;
            const bool two = cl + 1 < ncols;
            const double v0 = sC[cl * LDS + rl], v1 = two ? sC[(cl + 1) * LDS + rl] : 0.0;
            double *q = gM + (j0 + cl) + (long long)(i0 + rl) * tk.ldc;
            if (atomic) {
                atomicAdd(q, v0);
                if (two) atomicAdd(q + 1, v1);
            } else if (two && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
                double2 o = beta0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2 *>(q);
                o.x += v0; o.y += v1;
                *reinterpret_cast<double2 *>(q) = o;
            } else {
                q[0] = beta0 ? v0 : q[0] + v0;
                if (two) q[1] = beta0 ? v1 : q[1] + v1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-specialised variant for operands with the tile dimension contiguous (A(i,kk) at a + i + col(kk)*lda,
// B(j,kk) at b + j + kk*ldb): every k-column of an operand tile is one contiguous run of doubles, so the tile is
// staged by ONE producer warp with 32 1-D bulk-async copies (cp.async.bulk.shared::cluster.global, SASS UBLKCP) that
// complete on an mbarrier per stage, while WS_CONSUMERS warps do nothing but wait on that barrier, load fragments and
// issue DMMA; a stage is handed back through a second mbarrier, so there is no CTA-wide barrier in the main loop.  (A
// variant in which the producer role rotates over the consumer warps was measured 4-15 % slower: the issuing warp then
// waits for the slowest consumer of the slot it refills.)  Measured on B200
// (tools/gemm_lab.cu, profiles/r2_gemm_lab.txt): 35.3 TFLOP/s on 8192^3 (cuBLAS DGEMM: 35.5; the cp.async ring above:
// 32.5), 34.7 on the K = 512 panel shape (30.9), 35.0 on the 10000 x 10000 x 2401 update of the top fronts (31.6).
//   * tile 128 x 64, eight consumer warps of 32 x 32 (4 x 2), k-tile 16, four stages (100 KB: two CTAs per SM, so the
//     epilogue of one overlaps the main loop of the other);
//   * fragments are fetched as double2 from two ADJACENT rows (columns), which feed two different 8x8 MMA tiles --
//     rows 16p + 2*(lane/4) + {0,1} -- half the shared-memory load instructions for the same bytes; the +4 double
//     column padding keeps these 16-byte loads bank-conflict free;
//   * bulk copies move whole 16-byte units and cannot zero-fill: a tile with an odd row count reads one row more (the
//     operand spaces have even leading dimensions, so the row exists; it only feeds rows / columns of C that are never
//     written), and the k-columns beyond K of the last k-tile are masked in registers by the consumers.
constexpr int WS_BM = 128, WS_BN = 64, WS_BK = 16, WS_STAGES = 4, WS_WARPS_M = 4, WS_WARPS_N = 2;
constexpr int WS_CONSUMERS = WS_WARPS_M * WS_WARPS_N;
constexpr int WS_THREADS = (WS_CONSUMERS + 1) * 32;
constexpr int WS_LDA = WS_BM + 4, WS_LDB = WS_BN + 4;
constexpr int WS_STAGE_ELEMS = (WS_LDA + WS_LDB) * WS_BK;
constexpr int gemm_ws_smem_bytes()
{
    constexpr int pipe = WS_STAGES * WS_STAGE_ELEMS * 8 + 2 * WS_STAGES * 8 + 16;
    constexpr int epi = (WS_BM + 2) * WS_BN * 8;
    return pipe > epi ? pipe : epi;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int STG>
__global__ void __launch_bounds__(WS_THREADS)
k_gemm_ws(const GemmTask *__restrict__ tasks, const TileRef *__restrict__ tiles, GemmSpaces sp)
{
    constexpr int BM = WS_BM, BN = WS_BN, BKT = WS_BK;
    static_assert(STG == WS_STAGES, "shared-memory size is computed for WS_STAGES");
    constexpr int WM = BM / WS_WARPS_M, WN = BN / WS_WARPS_N, MT = WM / 8, NTL = WN / 8;
    static_assert(MT % 2 == 0 && NTL % 2 == 0, "paired fragment loads need an even number of 8x8 tiles per warp");
    static_assert(BKT == 16, "the producer maps lanes 0-15 to the k-columns of A and lanes 16-31 to those of B");
    constexpr int A_ELEMS = WS_LDA * BKT, B_ELEMS = WS_LDB * BKT;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem, *sB = smem + STG * A_ELEMS;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + STG * WS_STAGE_ELEMS);
    unsigned long long *empty = full + STG;

    pdl_enter();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < STG; s++) { mbar_init(full + s, 1); mbar_init(empty + s, WS_CONSUMERS); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == WS_CONSUMERS) {
        // ---- producer warp: lanes 0-15 copy k-column `lane` of A, lanes 16-31 k-column `lane - 16` of B, one bulk copy each
        const TileRef tr = tiles[blockIdx.x];
        const GemmTask tk = tasks[tr.task];
        const int i0 = tr.ti * BM, j0 = tr.tj * BN;
        const int rowsA = min(BM, tk.M - i0), rowsB = min(BN, tk.N - j0);
        const unsigned bytesA = (unsigned)((rowsA + (rowsA & 1)) * 8), bytesB = (unsigned)((rowsB + (rowsB & 1)) * 8);
        const bool isB = lane >= 16;
        const int pk = lane & 15;
        const double *gsrc = isB ? sp.base[(tk.flags >> 3) & 7] + tk.b + j0 : sp.base[tk.flags & 7] + tk.a + i0;
        const long long gstride = isB ? tk.ldb : tk.lda;
        const unsigned pbytes = isB ? bytesB : bytesA;
        const int *gather = (!isB && (tk.flags & GF_GATHER_A)) ? sp.idx + tk.aidx : nullptr;
        double *pdst = isB ? sB + pk * WS_LDB : sA + pk * WS_LDA;
        const int pstage = isB ? B_ELEMS : A_ELEMS;
        const int nk = (tk.K + BKT - 1) / BKT;
        for (int kt = 0; kt < nk; kt++) {
            const int s = kt % STG;
            if (kt >= STG) mbar_wait(empty + s, ((kt / STG) - 1) & 1);
            const int k0 = kt * BKT, kv = min(BKT, tk.K - k0);
            if (lane == 0) mbar_expect_tx(full + s, (bytesA + bytesB) * kv);
            __syncwarp();
            if (pk < kv) {
                const long long col = gather ? (long long)gather[k0 + pk] : (long long)(k0 + pk);
                bulk_g2s(pdst + s * pstage, gsrc + col * gstride, pbytes, full + s);
            }
        }
        return;
    }
    // ---- consumer warps.  Only K is read from the task before the main loop: everything the epilogue needs is read
    // afterwards, which keeps the loop (64 accumulator registers + fragments) inside the 96-register budget of two
    // resident CTAs per SM.
    const int wm = warp % WS_WARPS_M, wn = warp / WS_WARPS_M;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[MT][NTL][2];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    {
        const int K = tasks[tiles[blockIdx.x].task].K;
        const int nk = (K + BKT - 1) / BKT;
        const double *pA = sA + wm * WM + 2 * lr + lc * WS_LDA;
        const double *pB = sB + wn * WN + 2 * lr + lc * WS_LDB;
        for (int kt = 0; kt < nk; kt++) {
            const int s = kt % STG;
            mbar_wait(full + s, (kt / STG) & 1);
            const double *cA = pA + s * A_ELEMS;
            const double *cB = pB + s * B_ELEMS;
            const int kv = K - kt * BKT - lc;            // this lane's k index is valid while k4 < kv
#pragma unroll
            for (int k4 = 0; k4 < BKT; k4 += 4) {
                double fa[MT], fb[NTL];
#pragma unroll
                for (int a = 0; a < MT; a += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(cA + k4 * WS_LDA + a * 8);
                    fa[a] = v.x; fa[a + 1] = v.y;
                }
#pragma unroll
                for (int b = 0; b < NTL; b += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(cB + k4 * WS_LDB + b * 8);
                    fb[b] = v.x; fb[b + 1] = v.y;
                }
                if (k4 >= kv) {                          // k-columns beyond K were never copied: mask them
#pragma unroll
                    for (int a = 0; a < MT; a++) fa[a] = 0.0;
#pragma unroll
                    for (int b = 0; b < NTL; b++) fb[b] = 0.0;
                }
#pragma unroll
                for (int a = 0; a < MT; a++)
#pragma unroll
                    for (int b = 0; b < NTL; b++) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + s);
        }
    }
    // ---- epilogue: accumulators -> shared memory -> coalesced 16-byte read-modify-write of C,
    // same flag semantics as k_gemm_grouped
    constexpr int LDS = BM + 2, NCT = WS_CONSUMERS * 32;
    double *sC = smem;
    asm volatile("bar.sync 1, %0;\n" ::"n"(NCT) : "memory");      // every consumer is done with the operand stages
    // (re-read through volatile pointers so that none of this is kept live across the main loop)
    const volatile TileRef *vr = tiles + blockIdx.x;
    const volatile GemmTask *vt = tasks + vr->task;
    const int flags = vt->flags;
    const int ei0 = vr->ti * BM, ej0 = vr->tj * BN;
    const int erows = min(BM, vt->M - ei0), ecols = min(BN, vt->N - ej0);
    const bool neg = flags & GF_NEG, beta0 = flags & GF_BETA0, lower = flags & GF_LOWER;
    const bool atomic = flags & GF_ATOMIC, mirror = flags & GF_UPPER_MIRROR;
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                // paired loads permute rows and columns inside blocks of 16: MMA tile a holds rows 16(a/2) + 2 lr + (a&1)
                const int r = wm * WM + (a / 2) * 16 + 2 * lr + (a & 1);
                const int c = wn * WN + (b / 2) * 16 + 2 * (lc * 2 + e) + (b & 1);
                const double v = neg ? -acc[a][b][e] : acc[a][b][e];
                sC[c * LDS + r] = v;
            }
    asm volatile("bar.sync 1, %0;\n" ::"n"(NCT) : "memory");
    const int ldc = vt->ldc;
    double *gC = sp.base[(flags >> 6) & 7] + vt->c;
    const int *scat = (flags & GF_SCATTER_C) ? sp.idx + vt->cidx : nullptr;
    constexpr int RP = BM / 2;
    for (int id = tid; id < RP * BN; id += NCT) {
        const int cl = id / RP, rl = (id % RP) * 2;
        if (cl >= ecols || rl >= erows) continue;
        const int r = ei0 + rl, c = ej0 + cl;
        const bool two = rl + 1 < erows;
        bool w0 = true, w1 = two;
        if (lower) { w0 = r >= c; w1 = two && (r + 1 >= c); }
        if (!w0 && !w1) continue;
        const double v0 = sC[cl * LDS + rl], v1 = sC[cl * LDS + rl + 1];
        const long long col = scat ? (long long)scat[c] : (long long)c;
        double *p = gC + r + col * ldc;
        if (atomic) {
            if (w0) atomicAdd(p, v0);
            if (w1) atomicAdd(p + 1, v1);
        } else if (w0 && w1 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
            double2 o = beta0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2 *>(p);
            o.x += v0; o.y += v1;
            *reinterpret_cast<double2 *>(p) = o;
        } else {
            if (w0) p[0] = beta0 ? v0 : p[0] + v0;
            if (w1) p[1] = beta0 ? v1 : p[1] + v1;
        }
    }
    if (mirror) {
        double *gM = sp.base[(flags >> 6) & 7] + vt->c2;
        constexpr int CP = BN / 2;
        for (int id = tid; id < CP * BM; id += NCT) {
            const int rl = id / CP, cl = (id % CP) * 2;
            if (rl >= erows || cl >= ecols) continue;
            const bool two = cl + 1 < ecols;
            const double v0 = sC[cl * LDS + rl], v1 = two ? sC[(cl + 1) * LDS + rl] : 0.0;
            double *q = gM + (ej0 + cl) + (long long)(ei0 + rl) * ldc;
            if (atomic) {
                atomicAdd(q, v0);
                if (two) atomicAdd(q + 1, v1);
            } else if (two && ((reinterpret_cast<uintptr_t>(q) & 15) == 0)) {
                double2 o = beta0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2 *>(q);
                o.x += v0; o.y += v1;
                *reinterpret_cast<double2 *>(q) = o;
            } else {
                q[0] = beta0 ? v0 : q[0] + v0;
                if (two) q[1] = beta0 ? v1 : q[1] + v1;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Small-M variant (M <= 4 right-hand sides): the triangular solves for the conditional mean have one
// column per data replicate, where a 64x64 tensor-core tile would waste 63/64 of its rows.  These
// are HBM-bound matrix-vector products: every element of L is read exactly once, coalesced, and K is
// split across CTAs (atomic accumulation) so that a single wide front still fills the machine.
//   B_KMAJ=false: B(j,kk) at b + j + kk*ldb -> one thread per output column j (256 columns per tile)
//   B_KMAJ=true : B(j,kk) at b + kk + j*ldb -> one warp per output column, lanes stride over kk (64 columns per tile)
constexpr int GEMV_KC_N = 256;     // K chunk of the N-contiguous variant (one thread per column)
constexpr int GEMV_KC_K = 1024;    // K chunk of the K-contiguous variant (one warp per 8 columns)

template <bool B_KMAJ>
__global__ void __launch_bounds__(256) k_gemv_grouped(const GemmTask *__restrict__ tasks, const TileRef *__restrict__ tiles,
                                                      GemmSpaces sp)
{
    constexpr int KC = B_KMAJ ? GEMV_KC_K : GEMV_KC_N;
    __shared__ double sA[4][KC];             // the right-hand-side slice of this K chunk (<= 4 columns)
    pdl_enter();
    const TileRef tr = tiles[blockIdx.x];
    const GemmTask tk = tasks[tr.task];
    const int tid = threadIdx.x;
    const int k0 = tr.ti * KC, kn = min(tk.K, k0 + KC) - k0;
    const double *gA = sp.base[tk.flags & 7] + tk.a;
    const double *gB = sp.base[(tk.flags >> 3) & 7] + tk.b;
    double *gC = sp.base[(tk.flags >> 6) & 7] + tk.c;
    const int *gather = (tk.flags & GF_GATHER_A) ? sp.idx + tk.aidx : nullptr;
    const int *scat = (tk.flags & GF_SCATTER_C) ? sp.idx + tk.cidx : nullptr;
    const bool neg = tk.flags & GF_NEG, beta0 = tk.flags & GF_BETA0;
    const int M = tk.M;
    for (int e = tid; e < 4 * KC; e += 256) {
        const int i = e / KC, kk = e - i * KC;
        double v = 0.0;
        if (i < M && kk < kn) {
            const long long col = gather ? (long long)gather[k0 + kk] : (long long)(k0 + kk);
            v = gA[i + col * tk.lda];
        }
        sA[i][kk] = v;
    }
    __syncthreads();      // also orders the reads of A before the in-place writes of a BETA0 task
    if (!B_KMAJ) {
        const int j = tr.tj * 256 + tid;
        if (j >= tk.N) return;
        const double *bp = gB + j + (long long)k0 * tk.ldb;
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int kk = 0;
        for (; kk + 16 <= kn; kk += 16) {
            double bv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) bv[u] = bp[(long long)(kk + u) * tk.ldb];
#pragma unroll
            for (int u = 0; u < 16; u++)
#pragma unroll
                for (int i = 0; i < 4; i++) acc[i] += sA[i][kk + u] * bv[u];
        }
        for (; kk < kn; kk++) {
            const double bv = bp[(long long)kk * tk.ldb];
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] += sA[i][kk] * bv;
        }
        const long long ccol = scat ? (long long)scat[j] : (long long)j;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < M) {
                const double v = neg ? -acc[i] : acc[i];
                double *p = gC + i + ccol * tk.ldc;
                if (beta0) *p = v; else atomicAdd(p, v);
            }
    } else {
        // eight columns per warp, all in flight together: 8 (x2) independent coalesced loads per lane
        const int w = tid >> 5, lane = tid & 31;
        const int jb = tr.tj * 64 + w * 8;
        if (jb >= tk.N) return;
        const int nc = min(8, tk.N - jb);
        const double *bp = gB + k0 + (long long)jb * tk.ldb;
        double acc[8][4];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int i = 0; i < 4; i++) acc[c][i] = 0.0;
        for (int kk = lane; kk < kn; kk += 64) {
            double bv[8], bw[8];
            const bool second = kk + 32 < kn;
#pragma unroll
            for (int c = 0; c < 8; c++) {
                bv[c] = c < nc ? bp[kk + (long long)c * tk.ldb] : 0.0;
                bw[c] = (second && c < nc) ? bp[kk + 32 + (long long)c * tk.ldb] : 0.0;
            }
#pragma unroll
            for (int c = 0; c < 8; c++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    acc[c][i] += sA[i][kk] * bv[c];
                    if (second) acc[c][i] += sA[i][kk + 32] * bw[c];
                }
        }
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                for (int o = 16; o; o >>= 1) acc[c][i] += __shfl_down_sync(0xffffffffu, acc[c][i], o);
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                if (c >= nc) break;
                const int j = jb + c;
                const long long ccol = scat ? (long long)scat[j] : (long long)j;
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (i < M) {
                        const double v = neg ? -acc[c][i] : acc[c][i];
                        double *p = gC + i + ccol * tk.ldc;
                        if (beta0) *p = v; else atomicAdd(p, v);
                    }
            }
        }
    }
}

}  // namespace spde
