// plan.cu -- device side of the supernodal multifrontal method: small kernels (scatter, POTRF,
// extend-add, gathers, reductions), the program executor and the C ABI for plan / factor / solve
// / selected inverse.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "plan.h"

extern "C" long long spde_launch_count(int reset);

namespace spde {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }

// ---------------------------------------------------------------------------------------------
// K4: Q (slot layout, original ordering) -> zeroed L store (permuted supernodal panels);
// optional diagonal update tau * cnt  (Q_c = Q + tau S^T S, advection_diffusion2D.py:192)
__global__ void k_scatter_q(const double *__restrict__ Q, const long long *__restrict__ qdest,
                            const int *__restrict__ cand, int ncand, int n, int diag_slot,
                            const double *__restrict__ cnt, double tau, double *__restrict__ L)
{
    const long long total = (long long)ncand * n;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long d = qdest[e];
        if (d < 0) continue;
        const int ci = (int)(e / n), node = (int)(e % n);
        const int slot = cand[ci];
        double v = Q[(long long)slot * n + node];
        if (cnt && slot == diag_slot) v += cnt[node] * tau;
        L[d] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// extend-add: child update matrix (lower triangle) -> parent panel / parent update matrix
__global__ void k_extend_add(const ExtTask *__restrict__ tasks, const TileRef *__restrict__ tiles, GemmSpaces sp)
{
    pdl_enter();
    const TileRef tr = tiles[blockIdx.x];
    const ExtTask t = tasks[tr.task];
    const int *rel = sp.idx + t.rel;
    const double *src = sp.base[t.src_space] + t.src;
    double *dst = sp.base[t.dst_space];
    double *panel = sp.base[0] + t.ppanel;
    const int i = tr.ti * 32 + threadIdx.x;
    if (i >= t.nr) return;
    const int ri = rel[i];
    const int prow = ri < t.pnc ? ri : t.pncp + (ri - t.pnc);
    // the four columns of this thread are handled together: all index / source / destination loads are issued
    // before the first dependent use, so each thread keeps 12 memory operations in flight (HBM-bound kernel)
    double v[4], old[4];
    double *q[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int j = tr.tj * 32 + threadIdx.y + 8 * u;
        q[u] = nullptr;
        if (j < t.nr && j <= i) {
            const int rj = rel[j];
            v[u] = src[i + (long long)j * t.lds];
            q[u] = rj < t.pnc ? panel + prow + (long long)rj * t.pld
                              : dst + t.pupd + (ri - t.pnc) + (long long)(rj - t.pnc) * t.pldu;
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (q[u]) old[u] = *q[u];
#pragma unroll
    for (int u = 0; u < 4; u++) if (q[u]) *q[u] = old[u] + v[u];
}

// ---------------------------------------------------------------------------------------------
// POTRF of one diagonal block (b <= 64) in shared memory + explicit inverse of the factor.
// One CTA of 512 threads per block.  This kernel sits on the critical path of every front (one launch
// per 64 pivot columns), so it is organised for latency (tools/potrf_probe.py gives the phase clocks):
//   * right-looking column sweep, one barrier per column, 8 threads per row, four independent updates in
//     flight per thread; the loop body is kept short because with many warps the sweep is issue-bound;
//   * W = L^-1 by recursive block inversion, inv([[A,0],[B,C]]) = [[A^-1,0],[-C^-1 B A^-1, C^-1]]: eight
//     8x8 triangular inverses by substitution (one thread per column, column kept in registers), then three
//     doubling levels of two small dense products each -- 7 barriers instead of a 63-step substitution.
constexpr int POTRF_THREADS = 512;
template <bool BENCH, int SWEEP = 2>
__global__ void __launch_bounds__(POTRF_THREADS) k_potrf_t(const PotrfTask *__restrict__ tasks, double *__restrict__ L,
                                                           double *__restrict__ dinv, int *__restrict__ status,
                                                           long long *__restrict__ clk)
{
#define POTRF_MARK(k) do { if (BENCH && threadIdx.x == 0 && blockIdx.x == 0) clk[k] = clock64(); } while (0)
    // lower triangle: the block / its factor; strict upper triangle: W^T (W = L^-1); wd: diag(W)
    __shared__ double a[NB][NB + 1];
    __shared__ double tmp[NB * NB / 4];
    __shared__ double wd[NB];
    __shared__ int bad;
    pdl_enter();
    POTRF_MARK(0);
    const PotrfTask t = tasks[blockIdx.x];
    double *blk = L + t.blk;
    const int b = t.b, tid = threadIdx.x;
    if (tid == 0) bad = -1;
    for (int e = tid; e < NB * NB; e += POTRF_THREADS) {
        const int i = e % NB, j = e / NB;
        a[i][j] = (i < b && j < b && i >= j) ? blk[i + (long long)j * t.ld] : 0.0;
    }
    __syncthreads();
    POTRF_MARK(1);
    if (SWEEP == 2) {
        // Register-resident sweep (round 2).  Thread (row i = tid % 64, group ty = tid / 64) keeps the entries A(i, ty + 8k),
        // k = 0..7, of its row in registers for the whole sweep.  Per column j: the owners of column j publish its current
        // (unscaled) values to a double-buffered shared column, ONE barrier, then every thread reads the pivot and the
        // handful of column entries it needs, forms 1/sqrt(d) itself and updates its registers -- one shared-memory round
        // trip and one barrier on the dependent chain of a column (the shared-memory sweep below has two round trips and a
        // two-pass update loop): ~450 instead of ~840 cycles per column (tools/potrf_probe.py).
        constexpr int G = POTRF_THREADS / NB;       // 8 column groups
        static_assert(G == 8 && NB == 64, "register layout of the sweep");
        __shared__ double colbuf[2][NB];
        const int i = tid & 63, ty = tid >> 6;
        double r[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int c = ty + G * k; r[k] = (c <= i && i < b) ? a[i][c] : 0.0; }
#pragma unroll
        for (int kj = 0; kj < 8; kj++) {
#pragma unroll 1
            for (int own = 0; own < G; own++) {
                const int j = kj * G + own;
                if (j >= b) break;                               // uniform per CTA
                double *cb = colbuf[j & 1];
                if (ty == own && i >= j) cb[i] = r[kj];
                __syncthreads();
                const double d = cb[j];
                const bool ok = d > 0.0;
                const double inv = ok ? rsqrt(d) : nan("");     // 1/l_jj
                const double li = (i > j && i < b) ? cb[i] * inv : 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const int c = ty + G * k;
                    if (c > j && c <= i) r[k] -= li * (cb[c] * inv);
                }
                if (ty == 0 && i > j && i < b) a[i][j] = li;
                if (tid == 0) { a[j][j] = d * inv; wd[j] = inv; if (!ok && bad < 0) bad = j; }
            }
        }
        if (tid >= b && tid < NB) wd[tid] = 0.0;
        __syncthreads();
    } else {
        // rows i = tid % 64; the threads of a row split its columns c = j+1+ty, +G, ... (G column groups)
        constexpr int G = POTRF_THREADS / NB;
        const int i = tid & 63, ty = tid >> 6;
        double *ai = &a[i][0];
        for (int j = 0; j < b; j++) {
            const double d = a[j][j];
            const bool ok = d > 0.0;
            const double inv = ok ? rsqrt(d) : nan("");     // 1/l_jj; l_jj = d * inv
            const bool mine = i > j && i < b;
            double li = 0.0;
            if (mine) {
                li = ai[j] * inv;
                // four independent updates in flight per round (loads first, then the dependent arithmetic)
                for (int c0 = j + 1 + ty; c0 <= i; c0 += 4 * G) {
                    double p[4], q[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = c0 + G * u;
                        p[u] = c <= i ? a[c][j] : 0.0;
                        q[u] = c <= i ? ai[c] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int c = c0 + G * u;
                        if (c <= i) ai[c] = q[u] - li * (p[u] * inv);
                    }
                }
            }
            __syncthreads();                       // column j has been read by everyone
            if (ty == 0 && mine) ai[j] = li;
            if (tid == 0) { a[j][j] = d * inv; wd[j] = inv; if (!ok && bad < 0) bad = j; }
        }
        if (tid >= b && tid < NB) wd[tid] = 0.0;
        __syncthreads();
    }
    POTRF_MARK(2);
    // ---- W = L^-1.  W(i,c) for i > c lives at a[c][i]; diag(W) = 1/l_ii is already in wd.
    if (tid < NB) {
        // level 0: 8x8 diagonal blocks, thread = (block, column), the column stays in registers
        const int o = tid & ~7, c = tid & 7;
        double w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0.0;
        if (o + c < b) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < c || o + i >= b) continue;
                if (i == c) { w[i] = wd[o + i]; continue; }
                double sum = 0.0;
#pragma unroll
                for (int k = 0; k < 8; k++)
                    if (k >= c && k < i) sum += a[o + i][o + k] * w[k];
                w[i] = -sum * wd[o + i];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i > c) a[o + c][o + i] = w[i];
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 8; s < NB; s *= 2) {
        if (s >= b) break;                          // the remaining off-diagonal blocks are empty (uniform per CTA)
        // pairs (A at o, C at o+s, B = L[o+s.., o..]); outputs indexed e = (pair, i, c)
        const int per = s * s, total = (NB / (2 * s)) * per;
        for (int e = tid; e < total; e += POTRF_THREADS) {      // T = B * W_A
            const int pr = e / per, r = e - pr * per, i = r / s, c = r - i * s;
            const int o = pr * 2 * s;
            double s0 = 0.0, s1 = 0.0;
            if (o + s + i < b) {
                const double *bi = &a[o + s + i][o], *wc = &a[o + c][o];
                s0 = bi[c] * wd[o + c];
                int k = c + 1;
                for (; k + 1 < s; k += 2) { s0 += bi[k] * wc[k]; s1 += bi[k + 1] * wc[k + 1]; }
                if (k < s) s0 += bi[k] * wc[k];
            }
            tmp[e] = s0 + s1;
        }
        __syncthreads();
        for (int e = tid; e < total; e += POTRF_THREADS) {      // W21 = -W_C * T
            const int pr = e / per, r = e - pr * per, i = r / s, c = r - i * s;
            const int o = pr * 2 * s, oc = o + s;
            double s0 = 0.0, s1 = 0.0;
            if (oc + i < b) {
                const double *tc = &tmp[pr * per + c];
                s0 = wd[oc + i] * tc[i * s];
                int k = 0;
                for (; k + 1 < i; k += 2) { s0 += a[oc + k][oc + i] * tc[k * s]; s1 += a[oc + k + 1][oc + i] * tc[(k + 1) * s]; }
                if (k < i) s0 += a[oc + k][oc + i] * tc[k * s];
            }
            a[o + c][oc + i] = -(s0 + s1);
        }
        __syncthreads();
    }
    POTRF_MARK(3);
    for (int e = tid; e < NB * NB; e += POTRF_THREADS) {
        const int i = e % NB, j = e / NB;
        if (!BENCH && i < b && j < b) blk[i + (long long)j * t.ld] = (i >= j) ? a[i][j] : 0.0;
        double wv = 0.0;
        if (i < b && j < b) wv = (i > j) ? a[j][i] : (i == j ? wd[i] : 0.0);
        dinv[t.dinv + e] = wv;
    }
    if (tid == 0 && bad >= 0) atomicCAS(status, 0, t.col0 + bad + 1);
    POTRF_MARK(4);
#undef POTRF_MARK
}

// ---------------------------------------------------------------------------------------------
// K6: log-determinant, two deterministic stages (warp shuffles, fixed summation order)
__global__ void k_logdet_partial(const double *__restrict__ L, const long long *__restrict__ diagpos, int n,
                                 double *__restrict__ partial)
{
    double s = 0.0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) s += log(L[diagpos[j]]);
    __shared__ double sh[32];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = s;
    }
}
__global__ void k_sum_final(const double *__restrict__ partial, int m, double scale, double *__restrict__ out)
{
    double s = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) s += partial[j];
    __shared__ double sh[32];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = s * scale;
    }
}

// ---------------------------------------------------------------------------------------------
// permutations between the caller's row-major n x k block (original ordering) and the solver's
// k-major, permuted work array Xp[kp x n]
__global__ void k_perm_in(const double *__restrict__ X, const int *__restrict__ perm, int n, int k, int kp,
                          int use_perm, double *__restrict__ Xp)
{
    const long long total = (long long)n * kp;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(e % kp), j = (int)(e / kp);
        const int src = use_perm ? perm[j] : j;
        Xp[e] = p < k ? X[(long long)src * k + p] : 0.0;
    }
}
__global__ void k_perm_out(const double *__restrict__ Xp, const int *__restrict__ perm, int n, int k, int kp,
                           int use_perm, double *__restrict__ X)
{
    const long long total = (long long)n * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(e % k), j = (int)(e / k);
        const int dst = use_perm ? perm[j] : j;
        X[(long long)dst * k + p] = Xp[(long long)j * kp + p];
    }
}

// ---------------------------------------------------------------------------------------------
// selected inverse helpers
__global__ void k_selinv_gather(const GatherTask *__restrict__ tasks, const TileRef *__restrict__ tiles, GemmSpaces sp)
{
    // The source front is symmetric, so only the tiles on and below the diagonal are read (ti >= tj in the tile table); an
    // off-diagonal tile is written twice, as is and transposed through shared memory -- 12 instead of 16 bytes moved per
    // element of the child's trailing block.
    __shared__ double tile[32][33];
    pdl_enter();
    const TileRef tr = tiles[blockIdx.x];
    const GatherTask t = tasks[tr.task];
    const int *rel = sp.idx + t.rel;
    const double *src = sp.base[t.src_space] + t.src;
    double *dst = sp.base[t.dst_space] + t.dst;
    const int i = tr.ti * 32 + threadIdx.x;
    const bool vi = i < t.nr;
    int ri = vi ? rel[i] : 0;
    ri = ri < t.pnc ? ri : t.pncp + (ri - t.pnc);
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {       // four independent gathers in flight per thread
        const int j = tr.tj * 32 + threadIdx.y + 8 * u;
        v[u] = 0.0;
        if (vi && j < t.nr) {
            int rj = rel[j];
            rj = rj < t.pnc ? rj : t.pncp + (rj - t.pnc);
            v[u] = src[ri + (long long)rj * t.lds];
        }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int jl = threadIdx.y + 8 * u, j = tr.tj * 32 + jl;
        if (vi && j < t.nr) dst[(t.ncp + i) + (long long)(t.ncp + j) * t.ldd] = v[u];
        tile[jl][threadIdx.x] = v[u];
    }
    if (tr.ti == tr.tj) return;         // (uniform per CTA)
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int a = threadIdx.y + 8 * u;
        const int ii = tr.ti * 32 + a, jj = tr.tj * 32 + threadIdx.x;
        if (ii < t.nr && jj < t.nr) dst[(t.ncp + jj) + (long long)(t.ncp + ii) * t.ldd] = tile[threadIdx.x][a];
    }
}
__global__ void __launch_bounds__(256) k_wtw(const WtwTask *__restrict__ tasks, const double *__restrict__ dinv, GemmSpaces sp)
{
    __shared__ double w[NB][NB + 1];
    pdl_enter();
    const WtwTask t = tasks[blockIdx.x];
    double *dst = sp.base[t.space] + t.dst;
    if (t.pad == 1) {      // copy mode: the inverses W_j of the 64-column blocks of an outer block onto the block diagonal of Wf
        const int nbl = (t.b + NB - 1) / NB;
        for (int e = threadIdx.x; e < nbl * NB * NB; e += 256) {
            const int j = e / (NB * NB), r = e % (NB * NB), i = r % NB, c = r / NB;
            const int bj = min(NB, t.b - j * NB);
            if (i < bj && c < bj) dst[(j * NB + i) + (long long)(j * NB + c) * t.ldd] = dinv[t.w + e];
        }
        return;
    }
    for (int e = threadIdx.x; e < NB * NB; e += 256) w[e % NB][e / NB] = dinv[t.w + e];
    __syncthreads();
    // thread = (row i, group of 16 columns): 16 independent accumulators, so the FP64 pipe latency (~40 cycles per
    // dependent FMA) is hidden instead of serialised along each dot product.  W is lower triangular with explicit
    // zeros above the diagonal, so k starts at i.
    const int i = threadIdx.x & 63, jg = (threadIdx.x >> 6) * 16;
    double acc[16];
#pragma unroll
    for (int u = 0; u < 16; u++) acc[u] = 0.0;
    if (i < t.b && jg < t.b) {
        for (int k = i; k < t.b; k++) {
            const double wi = w[k][i];
#pragma unroll
            for (int u = 0; u < 16; u++) acc[u] += wi * w[k][jg + u];
        }
#pragma unroll
        for (int u = 0; u < 16; u++)
            if (jg + u < t.b) dst[i + (long long)(jg + u) * t.ldd] = acc[u];
    }
}
// rectangular block copy (LK_BCOPY): one CTA per 512 x 8 patch, coalesced along the rows
__global__ void __launch_bounds__(256) k_block_copy(const CopyTask *__restrict__ tasks, const TileRef *__restrict__ tiles, GemmSpaces sp)
{
    pdl_enter();
    const TileRef tr = tiles[blockIdx.x];
    const CopyTask t = tasks[tr.task];
    const double *src = sp.base[t.src_space] + t.src;
    double *dst = sp.base[t.dst_space] + t.dst;
    const int r0 = tr.ti * 512, c0 = tr.tj * 8;
    for (int e = threadIdx.x; e < 512 * 8; e += 256) {
        const int r = r0 + (e & 511), c = c0 + (e >> 9);
        if (r < t.rows && c < t.cols) dst[r + (long long)c * t.ldd] = src[r + (long long)c * t.lds];
    }
}
__global__ void k_extract(const ZEntry *__restrict__ ent, long long a0, long long a1, const double *__restrict__ zar,
                          double *__restrict__ Zq)
{
    pdl_enter();
    for (long long e = a0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < a1; e += (long long)gridDim.x * blockDim.x) {
        const ZEntry z = ent[e];
        const double v = zar[z.src];
        Zq[z.dst] = v;
        if (z.dst2 >= 0) Zq[z.dst2] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// executor

// Launch with the programmatic-stream-serialisation attribute: the kernel may begin launching while its
// predecessor in the stream (or captured graph) drains; pdl_enter() inside every schedule kernel restores
// the data dependency.  Measured on B200 (C2: 38.1 ms with vs 35.2 ms without, C3: 860 vs 850 ms) the early
// launch does not pay for these schedules, so plain stream order is the default; SPDE_PDL=1 enables it.
static int g_pdl = 0;
static int g_potrf_sweep = 2;      // SPDE_POTRF_SWEEP=1: the shared-memory column sweep of round 1
template <class... KArgs, class... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int BM, int BN, int WMn, int WNn, bool AK, bool BK_, int BKT = GEMM_BK, int STG = GEMM_STAGES>
static cudaError_t launch_gemm_variant(const Launch &L, const Program &P, const GemmSpaces &sp, cudaStream_t st)
{
    auto kern = k_gemm_grouped<BM, BN, WMn, WNn, AK, BK_, BKT, STG>;
    constexpr int smem = gemm_smem_bytes<BM, BN, BKT, STG>();
    static bool attr_done[64] = {false};      // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_done[dev] = true;
    }
    return launch_pdl(kern, dim3(L.ntiles), dim3(WMn * WNn * 32), smem, st, P.d_gemm + L.task0, P.d_tiles + L.tile0, sp);
}

static cudaError_t launch_gemm(const Launch &L, const Program &P, const GemmSpaces &sp, cudaStream_t st)
{
    const int cfg = L.variant / 4, ak = (L.variant >> 1) & 1, bk = L.variant & 1;
#define V(c, A, B)                                                                                     \
    if (cfg == c && ak == A && bk == B) {                                                              \
        if (c == 0) return launch_gemm_variant<128, 128, 2, 4, A, B>(L, P, sp, st);                     \
        if (c == 1) return launch_gemm_variant<128, 64, 4, 2, A, B>(L, P, sp, st);                      \
        return launch_gemm_variant<64, 64, 2, 2, A, B, 32, 2>(L, P, sp, st);   /* k-tile 32, 2 stages */  \
    }
    V(0, false, false) V(0, false, true) V(0, true, false) V(0, true, true)
    V(1, false, false) V(1, false, true) V(1, true, false) V(1, true, true)
    V(2, false, false) V(2, false, true) V(2, true, false) V(2, true, true)
#undef V
    // tuning-only tile shapes (NT layout), reachable through spde_gemm_single
    if (ak == 0 && bk == 0) {
        if (cfg == 3) {     // warp-specialised bulk-async kernel (k_gemm_ws): the production kernel of the N/N layout
            static bool ws_attr[64] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            dev &= 63;
            if (!ws_attr[dev]) {
                cudaError_t e = cudaFuncSetAttribute(k_gemm_ws<WS_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_ws_smem_bytes());
                if (e != cudaSuccess) return e;
                ws_attr[dev] = true;
            }
            return launch_pdl(k_gemm_ws<WS_STAGES>, dim3(L.ntiles), dim3(WS_THREADS), (size_t)gemm_ws_smem_bytes(), st,
                              P.d_gemm + L.task0, P.d_tiles + L.tile0, sp);
        }
        if (cfg == 4) return launch_gemm_variant<128, 128, 4, 4, false, false>(L, P, sp, st);   // 16 warps, 32x32
        if (cfg == 5) return launch_gemm_variant<64, 128, 2, 4, false, false>(L, P, sp, st);    // 8 warps, 32x32
        if (cfg == 6) return launch_gemm_variant<128, 128, 4, 2, false, false>(L, P, sp, st);   // 8 warps, 32x64
        if (cfg == 7) return launch_gemm_variant<256, 64, 8, 2, false, false>(L, P, sp, st);    // 16 warps, 32x32
        if (cfg == 8) return launch_gemm_variant<128, 64, 2, 4, false, false>(L, P, sp, st);    // 8 warps, 64x16
        if (cfg == 9) return launch_gemm_variant<64, 64, 2, 2, false, false, 32, 2>(L, P, sp, st);    // k-tile 32, 2 stages
        if (cfg == 10) return launch_gemm_variant<64, 64, 2, 2, false, false, 32, 3>(L, P, sp, st);   // k-tile 32, 3 stages
        if (cfg == 11) return launch_gemm_variant<64, 64, 2, 2, false, false, 16, 4>(L, P, sp, st);   // 4 stages
        // deep pipelines for launches with about one CTA per SM (the skinny products of the panel chain): a CTA alone on
        // its SM exposes one memory latency per k-tile unless several k-tiles are in flight
        if (cfg == 12) return launch_gemm_variant<64, 64, 2, 2, false, false, 32, 4>(L, P, sp, st);   // k-tile 32, 4 stages
        if (cfg == 13) return launch_gemm_variant<64, 64, 2, 2, false, false, 64, 2>(L, P, sp, st);   // k-tile 64, 2 stages
        if (cfg == 14) return launch_gemm_variant<64, 64, 2, 2, false, false, 64, 3>(L, P, sp, st);   // k-tile 64, 3 stages
        if (cfg == 15) return launch_gemm_variant<64, 64, 2, 2, false, false, 32, 6>(L, P, sp, st);   // k-tile 32, 6 stages
    }
    return cudaErrorInvalidValue;
}

template <class T>
static cudaError_t upload(const std::vector<T> &h, T **d)
{
    if (h.empty()) { *d = nullptr; return cudaSuccess; }
    cudaError_t e = cudaMalloc((void **)d, h.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

static int upload_program(Program &P)
{
    if (P.uploaded) return SPDE_OK;
    SPDE_CUDA_CHECK(upload(P.gemm, &P.d_gemm));
    SPDE_CUDA_CHECK(upload(P.tiles, &P.d_tiles));
    SPDE_CUDA_CHECK(upload(P.potrf, &P.d_potrf));
    SPDE_CUDA_CHECK(upload(P.ext, &P.d_ext));
    SPDE_CUDA_CHECK(upload(P.gather, &P.d_gather));
    SPDE_CUDA_CHECK(upload(P.wtw, &P.d_wtw));
    SPDE_CUDA_CHECK(upload(P.bcopy, &P.d_bcopy));
    P.uploaded = true;
    return SPDE_OK;
}

static void free_program(Program &P)
{
    for (int w = 0; w < 2; w++) if (P.graph[w]) { cudaGraphExecDestroy(P.graph[w]); P.graph[w] = nullptr; }
    cudaFree(P.d_gemm); cudaFree(P.d_tiles); cudaFree(P.d_potrf); cudaFree(P.d_ext); cudaFree(P.d_gather); cudaFree(P.d_wtw); cudaFree(P.d_bcopy);
    P.uploaded = false;
}

// (solve programs of a plan with outer-block solves: operand space 1 is the second right-hand-side buffer X2 instead of
// the update-matrix arena, which only the factorisation uses)
static GemmSpaces spaces_of(Plan &p, int which, bool solve = false)
{
    GemmSpaces sp;
    sp.base[0] = p.d_L[which];
    sp.base[1] = (solve && p.solve_outer) ? p.d_X2 : p.d_arena[which][0];
    sp.base[2] = p.d_arena[which][1];
    sp.base[3] = p.d_dinv[which];
    sp.base[4] = p.d_X;
    sp.base[5] = p.d_ybuf[which];
    sp.base[6] = p.d_zarena[which][0];
    sp.base[7] = p.d_zarena[which][1];
    sp.idx = p.d_idx;
    return sp;
}

static int issue_program(Plan &p, Program &P, int which, cudaStream_t st, double *d_Zq)
{
    int rc = upload_program(P);
    if (rc) return rc;
    ExecCtx ctx;
    ctx.sp = spaces_of(p, which, &P != &p.factor && &P != &p.selinv);
    ctx.L = p.d_L[which]; ctx.dinv = p.d_dinv[which]; ctx.status = p.d_status + which;
    ctx.zent = p.d_zentries; ctx.Zq = d_Zq; ctx.which = which; ctx.lanes = true;
    return issue_program_ex(p, P, ctx, st);
}

// Issue the launches of an (uploaded) program against explicit base pointers: the in-core stores of a plan
// (issue_program) or the memory pool of the streamed evaluator (ooc.cu).
int issue_program_ex(Plan &p, Program &P, const ExecCtx &ctx, cudaStream_t st)
{
    const GemmSpaces &sp = ctx.sp;
    const int which = ctx.which;
    double *const d_Zq = ctx.Zq;
    std::vector<cudaEvent_t> ev;
    if (p.prof_on) {
        ev.resize(P.launches.size() + 1);
        for (auto &e : ev) cudaEventCreate(&e);
        cudaEventRecord(ev[0], st);
    }
    size_t li = 0;
    const int pdl_saved = g_pdl;
    if (p.prof_on) g_pdl = 0;          // per-launch events need plain stream order
    struct Restore { int v; ~Restore() { g_pdl = v; } } restore{pdl_saved};
    // two-lane schedules: the bulk lane gets its own stream, ordered against the main lane by LK_SYNC records
    const bool lanes = ctx.lanes && p.use_lanes && !p.prof_on;
    const cudaStream_t main_st = st;
    cudaStream_t bulk_st = st;
    if (lanes) {
        if (!p.bulk_stream[which]) {
            // lowest priority: CTAs of the main lane (the latency-bound panel chain) are dispatched first whenever
            // an SM slot frees up, the bulk GEMMs fill what is left
            int least = 0, greatest = 0;
            SPDE_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            SPDE_CUDA_CHECK(cudaStreamCreateWithPriority(&p.bulk_stream[which], cudaStreamNonBlocking, least));
            SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.bulk_fork[which], cudaEventDisableTiming));
            for (int e = 0; e < 2; e++) SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.bulk_ev[which][e], cudaEventDisableTiming));
        }
        bulk_st = p.bulk_stream[which];
    }
    for (const Launch &L : P.launches) {
        if (L.kind == LK_SYNC) {
            if (lanes) {
                if (L.variant == 0) {
                    SPDE_CUDA_CHECK(cudaEventRecord(p.bulk_fork[which], main_st));
                    SPDE_CUDA_CHECK(cudaStreamWaitEvent(bulk_st, p.bulk_fork[which], 0));
                } else if (L.variant == 1) {
                    SPDE_CUDA_CHECK(cudaEventRecord(p.bulk_ev[which][L.a0 & 1], bulk_st));
                } else {
                    SPDE_CUDA_CHECK(cudaStreamWaitEvent(main_st, p.bulk_ev[which][L.a0 & 1], 0));
                }
            }
            if (p.prof_on) cudaEventRecord(ev[++li], st);
            continue;
        }
        if (L.kind == LK_COPY) {
            if (L.variant == 0 && !ctx.h_pool) {
                // forward-only run: nothing is parked
            } else if (L.variant == 0) {
                // park a finished slice of the panel in pinned host memory while the factorisation goes on
                const cudaStream_t cs = p.prof_on ? main_st : ctx.copy_stream;
                if (cs != main_st) {
                    SPDE_CUDA_CHECK(cudaEventRecord(ctx.copy_fork, main_st));
                    SPDE_CUDA_CHECK(cudaStreamWaitEvent(cs, ctx.copy_fork, 0));
                }
                SPDE_CUDA_CHECK(cudaMemcpyAsync(ctx.h_pool + L.task0, sp.base[0] + L.a0, (size_t)L.a1 * sizeof(double), cudaMemcpyDeviceToHost, cs));
            } else {
                SPDE_CUDA_CHECK(cudaStreamWaitEvent(main_st, (*ctx.fetch_ev)[L.a0], 0));
            }
            if (p.prof_on) cudaEventRecord(ev[++li], main_st);
            continue;
        }
        st = (lanes && L.lane == 1) ? bulk_st : main_st;
        switch (L.kind) {
        case LK_GEMM:
            SPDE_CUDA_CHECK(launch_gemm(L, P, sp, st));
            break;
        case LK_GEMV:
            if (L.variant) SPDE_CUDA_CHECK(launch_pdl(k_gemv_grouped<true>, dim3(L.ntiles), dim3(256), 0, st, P.d_gemm + L.task0, P.d_tiles + L.tile0, sp));
            else SPDE_CUDA_CHECK(launch_pdl(k_gemv_grouped<false>, dim3(L.ntiles), dim3(256), 0, st, P.d_gemm + L.task0, P.d_tiles + L.tile0, sp));
            break;
        case LK_POTRF:
            if (g_potrf_sweep == 1)
                SPDE_CUDA_CHECK(launch_pdl(k_potrf_t<false, 1>, dim3(L.ntasks), dim3(POTRF_THREADS), 0, st, P.d_potrf + L.task0, ctx.L, ctx.dinv,
                                           ctx.status, (long long *)nullptr));
            else
                SPDE_CUDA_CHECK(launch_pdl(k_potrf_t<false, 2>, dim3(L.ntasks), dim3(POTRF_THREADS), 0, st, P.d_potrf + L.task0, ctx.L, ctx.dinv,
                                           ctx.status, (long long *)nullptr));
            break;
        case LK_EXTADD:
            SPDE_CUDA_CHECK(launch_pdl(k_extend_add, dim3(L.ntiles), dim3(32, 8), 0, st, P.d_ext + L.task0, P.d_tiles + L.tile0, sp));
            break;
        case LK_ZERO:
            SPDE_CUDA_CHECK(cudaMemsetAsync(sp.base[L.variant] + L.a0, 0, (size_t)(L.a1 - L.a0) * sizeof(double), st));
            break;
        case LK_GATHER:
            SPDE_CUDA_CHECK(launch_pdl(k_selinv_gather, dim3(L.ntiles), dim3(32, 8), 0, st, P.d_gather + L.task0, P.d_tiles + L.tile0, sp));
            break;
        case LK_WTW:
            SPDE_CUDA_CHECK(launch_pdl(k_wtw, dim3(L.ntasks), dim3(256), 0, st, P.d_wtw + L.task0, (const double *)ctx.dinv, sp));
            break;
        case LK_BCOPY:
            SPDE_CUDA_CHECK(launch_pdl(k_block_copy, dim3(L.ntiles), dim3(256), 0, st, P.d_bcopy + L.task0, P.d_tiles + L.tile0, sp));
            break;
        case LK_EXTRACT: {
            const long long cnt = L.a1 - L.a0;
            SPDE_CUDA_CHECK(launch_pdl(k_extract, dim3((int)std::min<long long>((cnt + 255) / 256, 148 * 16)), dim3(256), 0, st,
                                       ctx.zent, (long long)L.a0, (long long)L.a1, (const double *)sp.base[L.variant], d_Zq));
            break;
        }
        }
        SPDE_LAUNCH_CHECK();
        if (L.kind != LK_ZERO) count_launch();
        if (p.prof_on) cudaEventRecord(ev[++li], st);
    }
    st = main_st;
    if (p.prof_on) {
        SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
        P.last_ms.assign(P.launches.size(), 0.f);
        for (size_t i = 0; i < P.launches.size(); i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            P.last_ms[i] = ms;
            const Launch &L = P.launches[i];
            if (L.kind == LK_SYNC || L.kind == LK_COPY) continue;
            const int v = L.kind == LK_GEMM ? L.variant : 0;
            const int kd = L.kind == LK_BCOPY ? (int)LK_GATHER : L.kind;      // block copies are accounted with the gathers
            p.prof_ms[kd][v] += ms;
            p.prof_cnt[kd][v] += 1;
        }
        for (auto &e : ev) cudaEventDestroy(e);
    }
    return SPDE_OK;
}

// host wrappers used by the streamed evaluator (ooc.cu)
int launch_logdet(const double *d_L, const long long *d_diagpos, int n, double *d_partial, double *d_out, cudaStream_t st)
{
    const int nb = 256;
    count_launch(2);
    k_logdet_partial<<<nb, 256, 0, st>>>(d_L, d_diagpos, n, d_partial);
    k_sum_final<<<1, 1024, 0, st>>>(d_partial, nb, 2.0, d_out);
    SPDE_LAUNCH_CHECK();
    return SPDE_OK;
}
int launch_perm_in(const double *d_X, const int *d_perm, int n, int k, int kp, int use_perm, double *d_Xp, cudaStream_t st)
{
    count_launch();
    k_perm_in<<<148 * 8, 256, 0, st>>>(d_X, d_perm, n, k, kp, use_perm, d_Xp);
    SPDE_LAUNCH_CHECK();
    return SPDE_OK;
}
int launch_perm_out(const double *d_Xp, const int *d_perm, int n, int k, int kp, int use_perm, double *d_X, cudaStream_t st)
{
    count_launch();
    k_perm_out<<<148 * 8, 256, 0, st>>>(d_Xp, d_perm, n, k, kp, use_perm, d_X);
    SPDE_LAUNCH_CHECK();
    return SPDE_OK;
}
void init_exec_env(Plan &p);

static void init_gemm_attributes()
{
    static bool done[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (done[dev]) return;
    done[dev] = true;
#define A(BM, BN, WM, WN)                                                                                          \
    cudaFuncSetAttribute(k_gemm_grouped<BM, BN, WM, WN, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>()); \
    cudaFuncSetAttribute(k_gemm_grouped<BM, BN, WM, WN, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>());  \
    cudaFuncSetAttribute(k_gemm_grouped<BM, BN, WM, WN, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>());  \
    cudaFuncSetAttribute(k_gemm_grouped<BM, BN, WM, WN, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BM, BN>());
    A(128, 128, 2, 4) A(128, 64, 4, 2)
#undef A
    // (the 64x64 production tile sets its attribute at first launch, launch_gemm_variant)
}

// Run a schedule: first call plainly, second call captured into a CUDA graph, later calls replayed.  The
// schedules are static per mesh (fixed task lists, fixed workspace pointers), so replay removes the host
// cost of thousands of launches per evaluation.
// make the caller's stream wait for the in-flight graph of a factor store
static int join_store(Plan &p, int which, cudaStream_t st)
{
    if (p.pending[which]) {
        SPDE_CUDA_CHECK(cudaStreamWaitEvent(st, p.ev_out[which], 0));
        p.pending[which] = false;
        p.pending_readonly[which] = false;
    }
    return SPDE_OK;
}

// lane 0: factorisation / Takahashi (private stream cap_stream, join may be deferred); lane 1: triangular solves
// (private stream solve_stream, always joined before returning; skips the join with an in-flight read-only graph)
static int run_program(Plan &p, Program &P, int which, cudaStream_t st, double *d_Zq, bool defer_join = false, int lane = 0)
{
    int rc = SPDE_OK;
    if (lane == 0 || !p.pending_readonly[which]) rc = join_store(p, which, st);
    if (rc) return rc;
    if (p.prof_on || !p.use_graphs) return issue_program(p, P, which, st, d_Zq);
    GemmSpaces sp = spaces_of(p, which, &P != &p.factor && &P != &p.selinv);
    unsigned long long key = 1469598103934665603ull;
    for (int i = 0; i < 8; i++) key = (key ^ (unsigned long long)(uintptr_t)sp.base[i]) * 1099511628211ull;
    key = (key ^ (unsigned long long)(uintptr_t)d_Zq) * 1099511628211ull;
    if (P.graph[which] && P.graph_key[which] != key) {
        cudaGraphExecDestroy(P.graph[which]);
        P.graph[which] = nullptr;
    }
    if (!p.cap_stream[which]) {
        int least = 0, greatest = 0;
        SPDE_CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        SPDE_CUDA_CHECK(cudaStreamCreateWithPriority(&p.cap_stream[which], cudaStreamNonBlocking, greatest));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.ev_in[which], cudaEventDisableTiming));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.ev_out[which], cudaEventDisableTiming));
    }
    if (lane == 1 && !p.solve_stream[which]) {
        SPDE_CUDA_CHECK(cudaStreamCreateWithFlags(&p.solve_stream[which], cudaStreamNonBlocking));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.sev_in[which], cudaEventDisableTiming));
        SPDE_CUDA_CHECK(cudaEventCreateWithFlags(&p.sev_out[which], cudaEventDisableTiming));
    }
    cudaStream_t cs = lane ? p.solve_stream[which] : p.cap_stream[which];
    if (!P.graph[which]) {
        if (P.runs[which]++ == 0) return issue_program(p, P, which, st, d_Zq);   // warm run: uploads, lazy init
        SPDE_CUDA_CHECK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        const long long before = spde_launch_count(0);
        rc = issue_program(p, P, which, cs, d_Zq);
        count_launch((int)(before - spde_launch_count(0)));     // counted again at every replay
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(cs, &g);
        if (rc) { if (g) cudaGraphDestroy(g); return rc; }
        SPDE_CUDA_CHECK(e);
        SPDE_CUDA_CHECK(cudaGraphInstantiate(&P.graph[which], g, 0));
        cudaGraphDestroy(g);
        P.graph_key[which] = key;
    }
    int nk = 0;
    for (const Launch &L : P.launches) nk += (L.kind != LK_ZERO && L.kind != LK_SYNC && L.kind != LK_COPY);
    count_launch(nk);
    if (lane == 1) {
        SPDE_CUDA_CHECK(cudaEventRecord(p.sev_in[which], st));
        SPDE_CUDA_CHECK(cudaStreamWaitEvent(cs, p.sev_in[which], 0));
        SPDE_CUDA_CHECK(cudaGraphLaunch(P.graph[which], cs));
        SPDE_CUDA_CHECK(cudaEventRecord(p.sev_out[which], cs));
        SPDE_CUDA_CHECK(cudaStreamWaitEvent(st, p.sev_out[which], 0));
        return SPDE_OK;
    }
    SPDE_CUDA_CHECK(cudaEventRecord(p.ev_in[which], st));
    SPDE_CUDA_CHECK(cudaStreamWaitEvent(cs, p.ev_in[which], 0));
    SPDE_CUDA_CHECK(cudaGraphLaunch(P.graph[which], cs));
    SPDE_CUDA_CHECK(cudaEventRecord(p.ev_out[which], cs));
    p.pending[which] = true;
    p.pending_readonly[which] = (&P == &p.selinv);
    return defer_join ? SPDE_OK : join_store(p, which, st);
}

void init_exec_env(Plan &p)
{
    init_gemm_attributes();
    const char *env = getenv("SPDE_GRAPHS");
    if (env) p.use_graphs = atoi(env);
    const char *pdl = getenv("SPDE_PDL");
    if (pdl) g_pdl = atoi(pdl);
    const char *ln = getenv("SPDE_LANES");
    if (ln) p.use_lanes = atoi(ln);
    const char *ps = getenv("SPDE_POTRF_SWEEP");
    if (ps) g_potrf_sweep = atoi(ps);
}

static int ensure_device(Plan &p, int which)
{
    init_exec_env(p);
    if (!(p.device_ready & 1)) {
        std::vector<int> idx(p.sym.rows);
        idx.insert(idx.end(), p.sym.relidx.begin(), p.sym.relidx.end());
        SPDE_CUDA_CHECK(upload(idx, &p.d_idx));
        SPDE_CUDA_CHECK(upload(p.qdest, &p.d_qdest));
        SPDE_CUDA_CHECK(upload(p.diagpos, &p.d_diagpos));
        SPDE_CUDA_CHECK(upload(p.cand_slots, &p.d_cand));
        SPDE_CUDA_CHECK(upload(p.sym.perm, &p.d_perm));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_status, 2 * sizeof(int)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_red, 4096 * sizeof(double)));
        p.device_ready |= 1;
    }
    if (!p.d_L[which]) {
        for (int a = 0; a < 2; a++)
            if (p.arena_size[a]) SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_arena[which][a], p.arena_size[a] * sizeof(double)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_L[which], std::max<int64_t>(p.l_size, 2) * sizeof(double)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_dinv[which], std::max<int64_t>(p.dinv_size, 2) * sizeof(double)));
        // (the outer-block inverses behind the 64x64 blocks are lower triangular: their upper blocks are never written)
        SPDE_CUDA_CHECK(cudaMemset(p.d_dinv[which], 0, std::max<int64_t>(p.dinv_size, 2) * sizeof(double)));
        // scratch of the outer-block products (factorisation: doubling rounds and the out-of-place TRSM; Takahashi: Yt)
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_ybuf[which], std::max<int64_t>(p.ybuf_size, 2) * sizeof(double)));
        // (zeroed once: a 16-byte cp.async with src-size 8 on the last row of an odd-height tile names the pad row behind it,
        // which no kernel writes -- it is zero-filled in shared memory, not read, but initcheck would flag the address)
        SPDE_CUDA_CHECK(cudaMemset(p.d_ybuf[which], 0, std::max<int64_t>(p.ybuf_size, 2) * sizeof(double)));
    }
    return SPDE_OK;
}

}  // namespace spde

using namespace spde;

extern "C" int spde_abi_version(void) { return 1; }
extern "C" const char *spde_last_error(void) { return g_err.c_str(); }
extern "C" long long spde_launch_count(int reset)
{
    const long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" int spde_plan_create(int M, int N, int T, int bc, int max_rhs, spde_plan **out)
{
    const Geo geo = geo_from_abi(M, N, T, bc);
    bc = geo.bc;
    if (M < 2 || N < 2 || T < 1 || bc < 1 || bc > 3 || geo.pat > 1 || !out) { set_error("spde_plan_create: bad arguments"); return SPDE_ERR_ARG; }
    if (bc == 2 && (M < 5 || N < 5)) { set_error("periodic meshes need M,N >= 5"); return SPDE_ERR_ARG; }
    Plan *p = new Plan();
    p->max_rhs = max_rhs;
    p->sym.analyse(geo);
    p->rel_base = (int64_t)p->sym.rows.size();
    p->build_layout();
    p->build_factor_program();
    *out = reinterpret_cast<spde_plan *>(p);
    return SPDE_OK;
}

extern "C" void spde_plan_destroy(spde_plan *pp)
{
    if (!pp) return;
    Plan *p = reinterpret_cast<Plan *>(pp);
    for (int a = 0; a < 2; a++) {
        cudaFree(p->d_L[a]); cudaFree(p->d_dinv[a]); cudaFree(p->d_ybuf[a]); cudaFree(p->d_zq[a]);
        for (int b = 0; b < 2; b++) { cudaFree(p->d_arena[a][b]); cudaFree(p->d_zarena[a][b]); }
        if (p->cap_stream[a]) { cudaStreamDestroy(p->cap_stream[a]); cudaEventDestroy(p->ev_in[a]); cudaEventDestroy(p->ev_out[a]); }
        if (p->bulk_stream[a]) {
            cudaStreamDestroy(p->bulk_stream[a]); cudaEventDestroy(p->bulk_fork[a]);
            cudaEventDestroy(p->bulk_ev[a][0]); cudaEventDestroy(p->bulk_ev[a][1]);
        }
        if (p->solve_stream[a]) { cudaStreamDestroy(p->solve_stream[a]); cudaEventDestroy(p->sev_in[a]); cudaEventDestroy(p->sev_out[a]); }
    }
    cudaFree(p->d_X); cudaFree(p->d_X2); cudaFree(p->d_red); cudaFree(p->d_idx); cudaFree(p->d_qdest);
    cudaFree(p->d_diagpos); cudaFree(p->d_cand); cudaFree(p->d_perm); cudaFree(p->d_status); cudaFree(p->d_zentries);
    free_program(p->factor);
    free_program(p->selinv);
    for (auto &kv : p->solve) free_program(kv.second);
    delete p;
}

extern "C" int64_t spde_plan_info(const spde_plan *pp, int what)
{
    const Plan *p = reinterpret_cast<const Plan *>(pp);
    switch (what) {
    case 0: return p->sym.n;
    case 1: return p->sym.nsuper;
    case 2: return p->sym.nnzL;
    case 4: return p->l_size * 8;
    case 5: return (p->arena_size[0] + p->arena_size[1]) * 8;
    case 6: return p->sym.maxdepth + 1;
    case 7: { int m = 0; for (auto &s : p->sn) m = std::max(m, s.nc + s.nr); return m; }
    case 8: return (int64_t)p->factor.launches.size();
    case 9: return (int64_t)p->sym.rows.size();
    case 10: return (p->zarena_size[0] + p->zarena_size[1]) * 8;
    case 11: return (int64_t)p->factor.gemm.size();
    case 12: return (int64_t)p->factor.tiles.size();
    case 13: { int m = 0; for (auto &s : p->sn) m = std::max(m, s.nc); return m; }
    case 15: return p->dinv_size * 8;      // inverse diagonal blocks + outer-block inverses
    case 16: return p->ybuf_size * 8;      // scratch of the outer-block products
    }
    return -1;
}
extern "C" double spde_plan_info_d(const spde_plan *pp, int what)
{
    const Plan *p = reinterpret_cast<const Plan *>(pp);
    if (what == 3) return p->sym.flops;
    if (what == 14) return p->factor.flops;   // flops actually scheduled (relaxed supernodes, full blocks)
    return (double)spde_plan_info(pp, what);
}
extern "C" int spde_plan_perm(const spde_plan *pp, int32_t *h_perm)
{
    const Plan *p = reinterpret_cast<const Plan *>(pp);
    memcpy(h_perm, p->sym.perm.data(), sizeof(int) * p->sym.n);
    return SPDE_OK;
}
extern "C" int spde_plan_supernodes(const spde_plan *pp, int32_t *h_first, int64_t *h_rowptr, int32_t *h_rows, int32_t *h_parent)
{
    const Plan *p = reinterpret_cast<const Plan *>(pp);
    const Symbolic &S = p->sym;
    memcpy(h_first, S.first.data(), sizeof(int) * (S.nsuper + 1));
    memcpy(h_rowptr, S.rowptr.data(), sizeof(int64_t) * (S.nsuper + 1));
    memcpy(h_rows, S.rows.data(), sizeof(int) * S.rows.size());
    memcpy(h_parent, S.sparent.data(), sizeof(int) * S.nsuper);
    return SPDE_OK;
}

// Host-side export of the schedules and the storage layout: used by the tests and by the NumPy
// interpreter in oracle/plan_emulator.py to validate the plan without a GPU.
extern "C" int spde_plan_export(spde_plan *pp, int prog, int k, int what, void *h_out, int64_t *count, int *elem_size)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    Program *P = nullptr;
    if (prog == 0) P = &p.factor;
    else if (prog == 1) P = &p.solve_program(k, 0);
    else if (prog == 2) P = &p.solve_program(k, 1);
    else if (prog == 3) { p.build_selinv_program(); P = &p.selinv; }
    const void *src = nullptr;
    int64_t cnt = 0;
    int es = 0;
#define EXP(vec) { src = (vec).data(); cnt = (int64_t)(vec).size(); es = (int)sizeof((vec)[0]); }
    if (prog >= 0 && prog <= 3) {
        switch (what) {
        case 0: EXP(P->launches) break;
        case 1: EXP(P->gemm) break;
        case 2: EXP(P->tiles) break;
        case 3: EXP(P->potrf) break;
        case 4: EXP(P->ext) break;
        case 5: EXP(P->gather) break;
        case 6: EXP(P->wtw) break;
        case 7: EXP(P->last_ms) break;
        case 8: EXP(P->bcopy) break;
        default: set_error("spde_plan_export: bad what"); return SPDE_ERR_ARG;
        }
    } else if (prog == 4) {   // layout
        static std::vector<int64_t> sizes;
        static std::vector<int> idx;
        switch (what) {
        case 0: sizes = {p.l_size, p.dinv_size, p.arena_size[0], p.arena_size[1], p.zarena_size[0], p.zarena_size[1], p.ybuf_size, p.rel_base,
                         (int64_t)p.solve_outer};
                EXP(sizes) break;
        case 1: EXP(p.qdest) break;
        case 2: EXP(p.cand_slots) break;
        case 3: EXP(p.diagpos) break;
        case 4: idx = p.sym.rows; idx.insert(idx.end(), p.sym.relidx.begin(), p.sym.relidx.end()); EXP(idx) break;
        case 5: EXP(p.zentries) break;
        case 6: EXP(p.zdepth_ptr) break;
        case 7: EXP(p.sym.colcount) break;
        default: set_error("spde_plan_export: bad what"); return SPDE_ERR_ARG;
        }
    } else { set_error("spde_plan_export: bad prog"); return SPDE_ERR_ARG; }
#undef EXP
    if (count) *count = cnt;
    if (elem_size) *elem_size = es;
    if (h_out && cnt) memcpy(h_out, src, (size_t)cnt * es);
    return SPDE_OK;
}

extern "C" int spde_plan_profile(spde_plan *pp, int enable, double *h_out /* 8*16*2 or NULL */, int reset)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (h_out) {
        memcpy(h_out, p.prof_ms, sizeof p.prof_ms);
        memcpy(h_out + 8 * 16, p.prof_cnt, sizeof p.prof_cnt);
    }
    if (reset) { memset(p.prof_ms, 0, sizeof p.prof_ms); memset(p.prof_cnt, 0, sizeof p.prof_cnt); }
    p.prof_on = enable != 0;
    return SPDE_OK;
}

// Asynchronous numeric factorisation: returns as soon as the schedule is enqueued (on the store's private
// stream once its CUDA graph exists), so the factorisations of Q and Q + tau S^T S overlap.  Completion,
// ordering with the caller's stream and the positive-definiteness status come from spde_factor_wait.
extern "C" int spde_factorize_async(spde_plan *pp, int which, const double *d_Q, const double *d_cnt, double tau, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (which < 0 || which > 1) { set_error("spde_factorize: which must be 0 or 1"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_device(p, which);
    if (rc) return rc;
    rc = join_store(p, which, st);
    if (rc) return rc;
    const int n = p.sym.n;
    SPDE_CUDA_CHECK(cudaMemsetAsync(p.d_L[which], 0, (size_t)p.l_size * sizeof(double), st));
    SPDE_CUDA_CHECK(cudaMemsetAsync(p.d_status + which, 0, sizeof(int), st));
    const int ncand = (int)p.cand_slots.size();
    count_launch();
    k_scatter_q<<<148 * 8, 256, 0, st>>>(d_Q, p.d_qdest, p.d_cand, ncand, n, p.sym.nslots / 2, d_cnt, tau, p.d_L[which]);
    SPDE_LAUNCH_CHECK();
    p.factored[which] = true;
    p.status[which] = SPDE_OK;
    p.bad_col[which] = -1;
    return run_program(p, p.factor, which, st, nullptr, true);
}

extern "C" int spde_factor_wait(spde_plan *pp, int which, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = join_store(p, which, st);
    if (rc) return rc;
    int h = 0;
    SPDE_CUDA_CHECK(cudaMemcpyAsync(&h, p.d_status + which, sizeof(int), cudaMemcpyDeviceToHost, st));
    SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
    p.status[which] = h ? SPDE_ERR_NOT_SPD : SPDE_OK;
    p.bad_col[which] = h - 1;
    if (h) { set_error("matrix is not positive definite (pivot " + std::to_string(h - 1) + " of the permuted matrix)"); return SPDE_ERR_NOT_SPD; }
    return SPDE_OK;
}

extern "C" int spde_factorize(spde_plan *pp, int which, const double *d_Q, const double *d_cnt, double tau, void *stream)
{
    int rc = spde_factorize_async(pp, which, d_Q, d_cnt, tau, stream);
    return rc ? rc : spde_factor_wait(pp, which, stream);
}

extern "C" int spde_factor_info(spde_plan *pp, int which, int *h_status, int *h_bad_column)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (h_status) *h_status = p.status[which];
    if (h_bad_column) *h_bad_column = p.bad_col[which];
    return SPDE_OK;
}

static int logdet_impl(spde_plan *pp, int which, double *out, bool device_out, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (!p.factored[which]) { set_error("spde_logdet: not factorised"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    { int rcj = join_store(p, which, st); if (rcj) return rcj; }
    const int nb = std::max(1, std::min(1024, (p.sym.n + 1023) / 1024));
    count_launch(2);
    k_logdet_partial<<<nb, 256, 0, st>>>(p.d_L[which], p.d_diagpos, p.sym.n, p.d_red);
    k_sum_final<<<1, 1024, 0, st>>>(p.d_red, nb, 2.0, device_out ? out : p.d_red + nb);
    SPDE_LAUNCH_CHECK();
    if (device_out) return SPDE_OK;
    SPDE_CUDA_CHECK(cudaMemcpyAsync(out, p.d_red + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
    SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
    return SPDE_OK;
}
extern "C" int spde_logdet(spde_plan *pp, int which, double *h_logdet, void *stream)
{
    return logdet_impl(pp, which, h_logdet, false, stream);
}
// same value written to a device address, without copy or synchronise (a NaN appears there when the factorisation
// broke down; the status itself is still reported by spde_factor_wait / spde_factor_info)
extern "C" int spde_logdet_dev(spde_plan *pp, int which, double *d_logdet, void *stream)
{
    return logdet_impl(pp, which, d_logdet, true, stream);
}

extern "C" int spde_solve(spde_plan *pp, int which, int mode, double *d_X, int k, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (!p.factored[which] || k < 1 || mode < 1 || mode > 15 || !(mode & 3)) { set_error("spde_solve: bad state/arguments"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (!p.pending_readonly[which]) { int rcj = join_store(p, which, st); if (rcj) return rcj; }
    const int n = p.sym.n, kp = k + (k & 1);
    const int64_t need = (int64_t)n * kp;
    if (p.x_cap < need) {
        cudaFree(p.d_X);
        cudaFree(p.d_X2);
        p.d_X = p.d_X2 = nullptr;
        p.x_cap = 0;
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_X, need * sizeof(double)));
        if (p.solve_outer) SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_X2, need * sizeof(double)));
        p.x_cap = need;
    }
    const int grid = 148 * 8;
    count_launch(2);
    // outer-block solves ping-pong between X and X2: the forward pass reads b from X and leaves y in X2, the backward
    // pass reads y from X2 and leaves x in X
    const bool pp2 = p.solve_outer;
    double *in = (pp2 && !(mode & 1)) ? p.d_X2 : p.d_X;           // backward only: the input is y
    double *out = (pp2 && !(mode & 2)) ? p.d_X2 : p.d_X;          // forward only: the output is y
    k_perm_in<<<grid, 256, 0, st>>>(d_X, p.d_perm, n, k, kp, (mode >> 2) & 1, in);
    SPDE_LAUNCH_CHECK();
    int rc;
    if (mode & 1) { rc = run_program(p, p.solve_program(k, 0), which, st, nullptr, false, 1); if (rc) return rc; }
    if (mode & 2) { rc = run_program(p, p.solve_program(k, 1), which, st, nullptr, false, 1); if (rc) return rc; }
    k_perm_out<<<grid, 256, 0, st>>>(out, p.d_perm, n, k, kp, (mode >> 3) & 1, d_X);
    SPDE_LAUNCH_CHECK();
    return SPDE_OK;
}

extern "C" int spde_selinv_start(spde_plan *pp, int which, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (!p.factored[which]) { set_error("spde_selinv: not factorised"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    p.build_selinv_program();
    const size_t zbytes = (size_t)p.sym.nslots * p.sym.n * sizeof(double);
    if (!p.sel_ready[which]) {
        for (int a = 0; a < 2; a++)
            if (p.zarena_size[a]) SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_zarena[which][a], p.zarena_size[a] * sizeof(double)));
        SPDE_CUDA_CHECK(cudaMalloc((void **)&p.d_zq[which], zbytes));
        if (!p.d_zentries) SPDE_CUDA_CHECK(upload(p.zentries, &p.d_zentries));
        p.sel_ready[which] = 1;
    }
    int rc = join_store(p, which, st);
    if (rc) return rc;
    SPDE_CUDA_CHECK(cudaMemsetAsync(p.d_zq[which], 0, zbytes, st));
    return run_program(p, p.selinv, which, st, p.d_zq[which], true);
}

extern "C" int spde_selinv_fetch(spde_plan *pp, int which, double *d_Zq, void *stream)
{
    Plan &p = *reinterpret_cast<Plan *>(pp);
    if (!p.sel_ready[which]) { set_error("spde_selinv_fetch: no selected inverse was started"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    int rc = join_store(p, which, st);
    if (rc) return rc;
    const size_t zbytes = (size_t)p.sym.nslots * p.sym.n * sizeof(double);
    SPDE_CUDA_CHECK(cudaMemcpyAsync(d_Zq, p.d_zq[which], zbytes, cudaMemcpyDeviceToDevice, st));
    return SPDE_OK;
}

extern "C" int spde_selinv(spde_plan *pp, int which, double *d_Zq, void *stream)
{
    int rc = spde_selinv_start(pp, which, stream);
    return rc ? rc : spde_selinv_fetch(pp, which, d_Zq, stream);
}

// ---------------------------------------------------------------------------------------------
// Single dense task through the grouped GEMM kernel: unit test and FP64 roofline probe.
// C (M x N, ldc) (op)= +-A*B with the flag/variant semantics of gemm.cuh; returns the mean device
// time of `reps` launches (CUDA events on `stream`) in *h_ms.
extern "C" int spde_gemm_single(int cfg, int a_kmaj, int b_kmaj, int flags, int M, int N, int K,
                                const double *d_A, int lda, const double *d_B, int ldb, double *d_C, int ldc,
                                int reps, float *h_ms, void *stream)
{
    if (cfg < 0 || cfg > 15 || M < 1 || N < 1 || K < 1) { set_error("spde_gemm_single: bad arguments"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    static const int BMs[16] = {128, 128, 64, 128, 128, 64, 128, 256, 128, 64, 64, 64, 64, 64, 64, 64};
    static const int BNs[16] = {128, 64, 64, 64, 128, 128, 128, 64, 64, 64, 64, 64, 128, 64, 64, 64};
    Program P;
    GemmTask t;
    memset(&t, 0, sizeof t);
    t.lda = lda; t.ldb = ldb; t.ldc = ldc; t.M = M; t.N = N; t.K = K;
    t.flags = (flags & ~511) | 0 | (1 << 3) | (2 << 6);
    P.gemm.push_back(t);
    const int tm = (M + BMs[cfg] - 1) / BMs[cfg], tn = (N + BNs[cfg] - 1) / BNs[cfg];
    for (int tj = 0; tj < tn; tj++)
        for (int ti = 0; ti < tm; ti++) {
            if ((flags & GF_LOWER) && (ti + 1) * BMs[cfg] - 1 < tj * BNs[cfg]) continue;
            P.tiles.push_back(TileRef{0, ti, tj, 0});
        }
    int rc = upload_program(P);
    if (rc) return rc;
    GemmSpaces sp;
    memset(&sp, 0, sizeof sp);
    sp.base[0] = const_cast<double *>(d_A);
    sp.base[1] = const_cast<double *>(d_B);
    sp.base[2] = d_C;
    Launch L;
    memset(&L, 0, sizeof L);
    L.kind = LK_GEMM; L.variant = cfg * 4 + (a_kmaj ? 2 : 0) + (b_kmaj ? 1 : 0);
    L.ntasks = 1; L.ntiles = (int)P.tiles.size();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaError_t err = launch_gemm(L, P, sp, st);   // warm-up / the tested launch when reps == 0
    cudaEventRecord(e0, st);
    for (int i = 0; i < reps && err == cudaSuccess; i++) err = launch_gemm(L, P, sp, st);
    cudaEventRecord(e1, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (h_ms) *h_ms = reps > 0 ? ms / reps : 0.f;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    free_program(P);
    if (err != cudaSuccess || e2 != cudaSuccess) {
        set_error(std::string("spde_gemm_single: ") + cudaGetErrorString(err != cudaSuccess ? err : e2));
        return SPDE_ERR_CUDA;
    }
    return SPDE_OK;
}

// Latency probe of the POTRF kernel: `ntasks` SPD blocks of order b (leading dimension ld), `reps` launches that
// do not write the factor back.  h_us: mean microseconds per launch (CUDA events around the whole train);
// h_clk[0..4]: clock64 of thread 0 of CTA 0 at entry / after load / after sweep / after inverse / at exit.
extern "C" int spde_potrf_bench(int b, int ld, int ntasks, int reps, float *h_us, long long *h_clk, void *stream)
{
    if (b < 1 || b > NB || ld < b || ntasks < 1 || reps < 1) { set_error("spde_potrf_bench: bad arguments"); return SPDE_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t per = (size_t)ld * NB;
    std::vector<double> h(per * ntasks, 0.0);
    for (int t = 0; t < ntasks; t++)
        for (int j = 0; j < b; j++)
            for (int i = j; i < b; i++) h[t * per + i + (size_t)j * ld] = (i == j) ? 4.0 + b : 1.0 / (1.0 + i - j);
    std::vector<PotrfTask> tk(ntasks);
    for (int t = 0; t < ntasks; t++) { tk[t].blk = (long long)(t * per); tk[t].dinv = (long long)t * NB * NB; tk[t].ld = ld; tk[t].b = b; tk[t].col0 = 0; tk[t].pad = 0; }
    double *dL = nullptr, *dW = nullptr; PotrfTask *dT = nullptr; int *dS = nullptr; long long *dC = nullptr;
    SPDE_CUDA_CHECK(cudaMalloc((void **)&dL, h.size() * sizeof(double)));
    SPDE_CUDA_CHECK(cudaMalloc((void **)&dW, (size_t)ntasks * NB * NB * sizeof(double)));
    SPDE_CUDA_CHECK(cudaMalloc((void **)&dT, tk.size() * sizeof(PotrfTask)));
    SPDE_CUDA_CHECK(cudaMalloc((void **)&dS, sizeof(int)));
    SPDE_CUDA_CHECK(cudaMalloc((void **)&dC, 8 * sizeof(long long)));
    SPDE_CUDA_CHECK(cudaMemcpy(dL, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
    SPDE_CUDA_CHECK(cudaMemcpy(dT, tk.data(), tk.size() * sizeof(PotrfTask), cudaMemcpyHostToDevice));
    SPDE_CUDA_CHECK(cudaMemset(dS, 0, sizeof(int)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *psw = getenv("SPDE_POTRF_SWEEP");
    const bool v1 = psw && atoi(psw) == 1;
    if (v1) k_potrf_t<true, 1><<<ntasks, POTRF_THREADS, 0, st>>>(dT, dL, dW, dS, dC);
    else k_potrf_t<true, 2><<<ntasks, POTRF_THREADS, 0, st>>>(dT, dL, dW, dS, dC);
    cudaEventRecord(e0, st);
    for (int r = 0; r < reps; r++) {
        if (v1) k_potrf_t<true, 1><<<ntasks, POTRF_THREADS, 0, st>>>(dT, dL, dW, dS, dC);
        else k_potrf_t<true, 2><<<ntasks, POTRF_THREADS, 0, st>>>(dT, dL, dW, dS, dC);
    }
    cudaEventRecord(e1, st);
    SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (h_us) *h_us = ms * 1e3f / reps;
    if (h_clk) SPDE_CUDA_CHECK(cudaMemcpy(h_clk, dC, 5 * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(dL); cudaFree(dW); cudaFree(dT); cudaFree(dS); cudaFree(dC);
    return SPDE_OK;
}
