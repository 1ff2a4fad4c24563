// gemm_lab.cu -- stand-alone FP64 DMMA GEMM laboratory for sm_100a (tuning tool, not part of the product).
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o gemm_lab tools/gemm_lab.cu
//   ./gemm_lab [M N K]...
//
// C (M x N, column-major) -= A (M x K, column-major) * B^T with B either N x K column-major ("NN": both operands have
// the tile dimension contiguous, the factorisation's shape) or K x N column-major ("NK": B has K contiguous, the
// Takahashi product's shape).  Compares kernel designs on the shapes of the supernodal schedules:
//   * cp.async ring + __syncthreads (the round-1 production kernel of spdepy_b200/csrc/gemm.cuh),
//   * warp-specialised: one producer warp issuing 1-D bulk-async copies (cp.async.bulk -> SASS UBLKCP) that complete on
//     mbarriers, consumer warps doing only fragment loads and DMMA, no CTA-wide barrier in the main loop.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ------------------------------------------------------------------------------------------------------------------
// baseline: the round-1 production design (NN only)
template <int BM, int BN, int WARPS_M, int WARPS_N, int BKT, int STG>
__global__ void __launch_bounds__(WARPS_M *WARPS_N * 32)
k_base(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ C, int ldc, int M, int N, int K)
{
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N, MT = WM / 8, NTL = WN / 8;
    constexpr int A_ELEMS = (BM + 4) * BKT, B_ELEMS = (BN + 4) * BKT;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem, *sB = smem + STG * A_ELEMS;
    const int tm = (M + BM - 1) / BM;
    const int ti = blockIdx.x % tm, tj = blockIdx.x / tm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int i0 = ti * BM, j0 = tj * BN;
    const int rowsA = min(BM, M - i0), rowsB = min(BN, N - j0);
    const double *gA = A + i0, *gB = B + j0;
    double acc[MT][NTL][2];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    auto load = [&](double *s, const double *g, int ld, int rows, int BT, int k0) {
        const int CH = BT / 2;
        for (int id = tid; id < CH * BKT; id += NT) {
            const int kk = id / CH, ic = (id % CH) * 2, kg = k0 + kk;
            int bytes = 0;
            const double *src = g;
            if (kg < K && ic < rows) { src = g + ic + (long long)kg * ld; bytes = (rows - ic >= 2) ? 16 : 8; }
            cp_async16(s + kk * (BT + 4) + ic, src, bytes);
        }
    };
    const int nk = (K + BKT - 1) / BKT;
#pragma unroll
    for (int s = 0; s < STG - 1; s++) {
        if (s < nk) { load(sA + s * A_ELEMS, gA, lda, rowsA, BM, s * BKT); load(sB + s * B_ELEMS, gB, ldb, rowsB, BN, s * BKT); }
        cp_async_commit();
    }
    const int lr = lane >> 2, lc = lane & 3;
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<STG - 2>();
        __syncthreads();
        {
            const int nx = kt + STG - 1;
            if (nx < nk) { const int s = nx % STG; load(sA + s * A_ELEMS, gA, lda, rowsA, BM, nx * BKT); load(sB + s * B_ELEMS, gB, ldb, rowsB, BN, nx * BKT); }
            cp_async_commit();
        }
        const double *cA = sA + (kt % STG) * A_ELEMS, *cB = sB + (kt % STG) * B_ELEMS;
#pragma unroll
        for (int k4 = 0; k4 < BKT; k4 += 4) {
            double fa[MT], fb[NTL];
#pragma unroll
            for (int a = 0; a < MT; a++) fa[a] = cA[(k4 + lc) * (BM + 4) + wm * WM + a * 8 + lr];
#pragma unroll
            for (int b = 0; b < NTL; b++) fb[b] = cB[(k4 + lc) * (BN + 4) + wn * WN + b * 8 + lr];
#pragma unroll
            for (int a = 0; a < MT; a++)
#pragma unroll
                for (int b = 0; b < NTL; b++) dmma884(acc[a][b][0], acc[a][b][1], fa[a], fb[b]);
        }
    }
    cp_async_wait<0>();
    constexpr int LDS = BM + 2;
    double *sC = smem;
    __syncthreads();
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) sC[(wn * WN + b * 8 + lc * 2 + e) * LDS + (wm * WM + a * 8 + lr)] = -acc[a][b][e];
    __syncthreads();
    const int mrows = rowsA, ncols = rowsB;
    constexpr int RP = BM / 2;
    for (int id = tid; id < RP * BN; id += NT) {
        const int cl = id / RP, rl = (id % RP) * 2;
        if (cl >= ncols || rl >= mrows) continue;
        double *p = C + (i0 + rl) + (long long)(j0 + cl) * ldc;
        if (rl + 1 < mrows) {
            double2 o = *reinterpret_cast<const double2 *>(p);
            o.x += sC[cl * LDS + rl]; o.y += sC[cl * LDS + rl + 1];
            *reinterpret_cast<double2 *>(p) = o;
        } else p[0] += sC[cl * LDS + rl];
    }
}

// ------------------------------------------------------------------------------------------------------------------
// warp-specialised bulk-async design
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

template <int BM, int BN, int BKT, int STG, bool B_KMAJ>
constexpr int ws_smem_bytes()
{
    constexpr int a = (BM + 4) * BKT;
    constexpr int b = B_KMAJ ? BN * (BKT + 4) : (BN + 4) * BKT;
    constexpr int pipe = STG * (a + b) * 8 + 2 * STG * 8 + 16;
    constexpr int epi = (BM + 2) * BN * 8;
    return pipe > epi ? pipe : epi;
}

// PAIR: A (and B) fragments are fetched as double2 from two adjacent rows, which feed two different 8x8 MMA tiles
// (rows 16p + 2*lr + {0,1}) -- half the shared-memory load instructions for the same bytes.
template <int BM, int BN, int WARPS_M, int WARPS_N, int BKT, int STG, bool B_KMAJ, bool PAIR>
__global__ void __launch_bounds__((WARPS_M * WARPS_N + 1) * 32)
k_ws(const double *__restrict__ A, int lda, const double *__restrict__ B, int ldb, double *__restrict__ C, int ldc, int M, int N, int K)
{
    constexpr int NC = WARPS_M * WARPS_N;           // consumer warps
    constexpr int WM = BM / WARPS_M, WN = BN / WARPS_N, MT = WM / 8, NTL = WN / 8;
    constexpr int LDA_S = BM + 4;
    constexpr int LDB_S = B_KMAJ ? BKT + 4 : BN + 4;
    constexpr int A_ELEMS = LDA_S * BKT;
    constexpr int B_ELEMS = B_KMAJ ? BN * LDB_S : LDB_S * BKT;
    extern __shared__ __align__(16) double smem[];
    double *sA = smem, *sB = smem + STG * A_ELEMS;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + STG * (A_ELEMS + B_ELEMS));
    unsigned long long *empty = full + STG;
    const int tm = (M + BM - 1) / BM;
    const int ti = blockIdx.x % tm, tj = blockIdx.x / tm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int i0 = ti * BM, j0 = tj * BN;
    const int rowsA = min(BM, M - i0), rowsB = min(BN, N - j0);
    const int nk = (K + BKT - 1) / BKT;
    if (tid == 0) {
        for (int s = 0; s < STG; s++) { mbar_init(full + s, 1); mbar_init(empty + s, NC); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == NC) {
        // ---------------- producer warp: one bulk copy per k-column (N-major operand) or per row (K-major operand)
        const double *gA = A + i0;
        const double *gB = B_KMAJ ? B + (long long)j0 * ldb : B + j0;
        const unsigned bytesA = (unsigned)((rowsA + (rowsA & 1)) * 8);
        const unsigned bytesBn = (unsigned)((rowsB + (rowsB & 1)) * 8);
        for (int kt = 0; kt < nk; kt++) {
            const int s = kt % STG;
            if (kt >= STG) mbar_wait(empty + s, ((kt / STG) - 1) & 1);
            const int k0 = kt * BKT, kv = min(BKT, K - k0);
            const unsigned bytesBk = (unsigned)((kv + (kv & 1)) * 8);
            if (lane == 0) mbar_expect_tx(full + s, bytesA * kv + (B_KMAJ ? bytesBk * rowsB : bytesBn * kv));
            __syncwarp();
            double *dA = sA + s * A_ELEMS, *dB = sB + s * B_ELEMS;
            for (int kk = lane; kk < kv; kk += 32) bulk_g2s(dA + kk * LDA_S, gA + (long long)(k0 + kk) * lda, bytesA, full + s);
            if (B_KMAJ) {
                for (int j = lane; j < rowsB; j += 32) bulk_g2s(dB + j * LDB_S, gB + k0 + (long long)j * ldb, bytesBk, full + s);
            } else {
                for (int kk = lane; kk < kv; kk += 32) bulk_g2s(dB + kk * LDB_S, gB + (long long)(k0 + kk) * ldb, bytesBn, full + s);
            }
        }
        return;
    }
    // ---------------- consumer warps
    const int wm = warp % WARPS_M, wn = warp / WARPS_M;
    const int lr = lane >> 2, lc = lane & 3;
    double acc[MT][NTL][2];
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % STG;
        mbar_wait(full + s, (kt / STG) & 1);
        const double *cA = sA + s * A_ELEMS + wm * WM, *cB = sB + s * B_ELEMS;
        const int kv = min(BKT, K - kt * BKT);
        const bool tail = kv < BKT;
#pragma unroll
        for (int k4 = 0; k4 < BKT; k4 += 4) {
            double fa[MT], fb[NTL];
            if (PAIR) {
#pragma unroll
                for (int a = 0; a < MT; a += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(cA + (k4 + lc) * LDA_S + a * 8 + 2 * lr);
                    fa[a] = v.x; fa[a + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int a = 0; a < MT; a++) fa[a] = cA[(k4 + lc) * LDA_S + a * 8 + lr];
            }
            if (B_KMAJ) {
#pragma unroll
                for (int b = 0; b < NTL; b++) fb[b] = cB[(wn * WN + b * 8 + lr) * LDB_S + k4 + lc];
            } else if (PAIR) {
#pragma unroll
                for (int b = 0; b < NTL; b += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(cB + (k4 + lc) * LDB_S + wn * WN + b * 8 + 2 * lr);
                    fb[b] = v.x; fb[b + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int b = 0; b < NTL; b++) fb[b] = cB[(k4 + lc) * LDB_S + wn * WN + b * 8 + lr];
            }
            if (tail && k4 + lc >= kv) {
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
    // epilogue: accumulators -> shared memory -> coalesced read-modify-write of C (consumer warps only)
    constexpr int LDS = BM + 2;
    double *sC = smem;
    asm volatile("bar.sync 1, %0;\n" ::"n"(NC * 32) : "memory");      // every consumer is done with the operand stages
#pragma unroll
    for (int a = 0; a < MT; a++)
#pragma unroll
        for (int b = 0; b < NTL; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                // row / column of this accumulator inside the CTA tile (PAIR permutes rows and columns inside 16-blocks)
                const int r = PAIR ? (a / 2) * 16 + 2 * lr + (a & 1) : a * 8 + lr;
                const int cq = lc * 2 + e;                       // column inside the 8x8 tile
                const int c = (PAIR && !B_KMAJ) ? (b / 2) * 16 + 2 * cq + (b & 1) : b * 8 + cq;
                sC[(wn * WN + c) * LDS + wm * WM + r] = -acc[a][b][e];
            }
    asm volatile("bar.sync 1, %0;\n" ::"n"(NC * 32) : "memory");
    constexpr int RP = BM / 2;
    for (int id = tid; id < RP * BN; id += NC * 32) {
        const int cl = id / RP, rl = (id % RP) * 2;
        if (cl >= rowsB || rl >= rowsA) continue;
        double *p = C + (i0 + rl) + (long long)(j0 + cl) * ldc;
        if (rl + 1 < rowsA) {
            double2 o = *reinterpret_cast<const double2 *>(p);
            o.x += sC[cl * LDS + rl]; o.y += sC[cl * LDS + rl + 1];
            *reinterpret_cast<double2 *>(p) = o;
        } else p[0] += sC[cl * LDS + rl];
    }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void k_ref(const double *A, int lda, const double *B, int ldb, double *C, int ldc, int M, int N, int K, int bk)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= M || j >= N) return;
    double s = 0.0;
    for (int k = 0; k < K; k++) s += A[i + (long long)k * lda] * (bk ? B[k + (long long)j * ldb] : B[j + (long long)k * ldb]);
    C[i + (long long)j * ldc] -= s;
}

struct Case { const char *name; int bkmaj; void (*launch)(const double *, int, const double *, int, double *, int, int, int, int, cudaStream_t); };

template <int BM, int BN, int WM_, int WN_, int BKT, int STG>
static void launch_base(const double *A, int lda, const double *B, int ldb, double *C, int ldc, int M, int N, int K, cudaStream_t st)
{
    auto kern = k_base<BM, BN, WM_, WN_, BKT, STG>;
    constexpr int a = (BM + 4) * BKT, b = (BN + 4) * BKT;
    constexpr int pipe = STG * (a + b) * 8, epi = (BM + 2) * BN * 8;
    constexpr int smem = pipe > epi ? pipe : epi;
    static bool done = false;
    if (!done) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); done = true; }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    kern<<<tiles, WM_ * WN_ * 32, smem, st>>>(A, lda, B, ldb, C, ldc, M, N, K);
}
template <int BM, int BN, int WM_, int WN_, int BKT, int STG, bool BK_, bool PAIR>
static void launch_ws(const double *A, int lda, const double *B, int ldb, double *C, int ldc, int M, int N, int K, cudaStream_t st)
{
    auto kern = k_ws<BM, BN, WM_, WN_, BKT, STG, BK_, PAIR>;
    constexpr int smem = ws_smem_bytes<BM, BN, BKT, STG, BK_>();
    static bool done = false;
    if (!done) { CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); done = true; }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    kern<<<tiles, (WM_ * WN_ + 1) * 32, smem, st>>>(A, lda, B, ldb, C, ldc, M, N, K);
}

int main(int argc, char **argv)
{
    std::vector<Case> cases = {
        {"base 64x64 2x2w bk32 s2 (r1 production)", 0, launch_base<64, 64, 2, 2, 32, 2>},
        {"base 64x128 2x2w(32x64) bk16 s3 (cutlass shape)", 0, launch_base<64, 128, 2, 2, 16, 3>},
        {"base 128x64 2x2w(64x32) bk16 s3", 0, launch_base<128, 64, 2, 2, 16, 3>},
        {"ws 128x64 2x2c(64x32) bk16 s4", 0, launch_ws<128, 64, 2, 2, 16, 4, false, false>},
        {"ws 128x64 2x2c(64x32) bk16 s4 pair", 0, launch_ws<128, 64, 2, 2, 16, 4, false, true>},
        {"ws 64x128 2x2c(32x64) bk16 s4 pair", 0, launch_ws<64, 128, 2, 2, 16, 4, false, true>},
        {"ws 128x64 2x2c bk32 s3 pair", 0, launch_ws<128, 64, 2, 2, 32, 3, false, true>},
        {"ws 128x128 2x4c(64x32) bk16 s5 pair", 0, launch_ws<128, 128, 2, 4, 16, 5, false, true>},
        {"ws 128x128 4x2c(32x64) bk16 s5 pair", 0, launch_ws<128, 128, 4, 2, 16, 5, false, true>},
        {"ws 64x64 2x2c(32x32) bk32 s3 pair", 0, launch_ws<64, 64, 2, 2, 32, 3, false, true>},
        {"ws 128x64 4x2c(32x32) bk16 s4 pair", 0, launch_ws<128, 64, 4, 2, 16, 4, false, true>},
        {"NK ws 128x64 2x2c bk16 s4 pair", 1, launch_ws<128, 64, 2, 2, 16, 4, true, true>},
        {"NK ws 128x64 2x2c bk32 s3 pair", 1, launch_ws<128, 64, 2, 2, 32, 3, true, true>},
        {"NK ws 64x64 2x2c bk32 s3 pair", 1, launch_ws<64, 64, 2, 2, 32, 3, true, true>},
    };
    std::vector<std::vector<int>> shapes;
    for (int i = 1; i + 2 < argc; i += 3) shapes.push_back({atoi(argv[i]), atoi(argv[i + 1]), atoi(argv[i + 2])});
    if (shapes.empty()) shapes = {{8192, 8192, 8192}, {8192, 8192, 512}, {10000, 10000, 2401}, {12274, 64, 1024}, {4097, 4097, 577}};
    const char *only = getenv("LAB_ONLY");
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    // correctness on a ragged problem first
    {
        const int M = 333, N = 201, K = 150, lda = 334, ldbn = 202, ldbk = 150, ldc = 334;
        std::vector<double> hA((size_t)lda * K), hBn((size_t)ldbn * K), hBk((size_t)ldbk * N), hC((size_t)ldc * N);
        srand(1);
        for (auto &v : hA) v = rand() / (double)RAND_MAX - 0.5;
        for (int k = 0; k < K; k++) for (int j = 0; j < ldbn; j++) hBn[j + (size_t)k * ldbn] = rand() / (double)RAND_MAX - 0.5;
        for (int j = 0; j < N; j++) for (int k = 0; k < K; k++) hBk[k + (size_t)j * ldbk] = hBn[j + (size_t)k * ldbn];
        for (auto &v : hC) v = rand() / (double)RAND_MAX;
        double *dA, *dBn, *dBk, *dC, *dR;
        CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dBn, hBn.size() * 8)); CK(cudaMalloc(&dBk, hBk.size() * 8));
        CK(cudaMalloc(&dC, hC.size() * 8)); CK(cudaMalloc(&dR, hC.size() * 8));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBn, hBn.data(), hBn.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dBk, hBk.data(), hBk.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dR, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice));
        k_ref<<<dim3((M + 127) / 128, N), 128, 0, st>>>(dA, lda, dBn, ldbn, dR, ldc, M, N, K, 0);
        std::vector<double> ref(hC.size()), got(hC.size());
        CK(cudaMemcpyAsync(ref.data(), dR, hC.size() * 8, cudaMemcpyDeviceToHost, st));
        for (auto &c : cases) {
            if (only && !strstr(c.name, only)) continue;
            CK(cudaMemcpyAsync(dC, hC.data(), hC.size() * 8, cudaMemcpyHostToDevice, st));
            c.launch(dA, lda, c.bkmaj ? dBk : dBn, c.bkmaj ? ldbk : ldbn, dC, ldc, M, N, K, st);
            CK(cudaMemcpyAsync(got.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            double err = 0;
            for (int j = 0; j < N; j++) for (int i = 0; i < M; i++) err = fmax(err, fabs(got[i + (size_t)j * ldc] - ref[i + (size_t)j * ldc]));
            // padding rows of C must be untouched
            double pad = 0;
            for (int j = 0; j < N; j++) pad = fmax(pad, fabs(got[M + (size_t)j * ldc] - hC[M + (size_t)j * ldc]));
            printf("check %-50s max err %.3g pad %.3g %s\n", c.name, err, pad, (err < 1e-11 && pad == 0) ? "OK" : "FAIL");
        }
        cudaFree(dA); cudaFree(dBn); cudaFree(dBk); cudaFree(dC); cudaFree(dR);
    }
    for (auto &sh : shapes) {
        const int M = sh[0], N = sh[1], K = sh[2];
        const int lda = M + (M & 1), ldbn = N + (N & 1), ldbk = K + (K & 1), ldc = lda;
        double *dA, *dB, *dC;
        const size_t nb = (size_t)std::max((size_t)ldbn * K, (size_t)ldbk * N);
        CK(cudaMalloc(&dA, (size_t)lda * K * 8)); CK(cudaMalloc(&dB, nb * 8)); CK(cudaMalloc(&dC, (size_t)ldc * N * 8));
        CK(cudaMemset(dA, 0, (size_t)lda * K * 8)); CK(cudaMemset(dB, 0, nb * 8)); CK(cudaMemset(dC, 0, (size_t)ldc * N * 8));
        printf("shape %d x %d x %d\n", M, N, K);
        for (auto &c : cases) {
            if (only && !strstr(c.name, only)) continue;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            c.launch(dA, lda, dB, c.bkmaj ? ldbk : ldbn, dC, ldc, M, N, K, st);
            float best = 1e30f;
            const double fl = 2.0 * M * N * K;
            const int reps = fl > 5e11 ? 2 : 6;
            for (int it = 0; it < 3; it++) {
                cudaEventRecord(e0, st);
                for (int r = 0; r < reps; r++) c.launch(dA, lda, dB, c.bkmaj ? ldbk : ldbn, dC, ldc, M, N, K, st);
                cudaEventRecord(e1, st);
                CK(cudaStreamSynchronize(st));
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                best = fminf(best, ms / reps);
            }
            printf("  %-50s %8.3f ms  %6.2f TFLOP/s\n", c.name, best, fl / best / 1e9);
            cudaEventDestroy(e0); cudaEventDestroy(e1);
        }
        cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
    return 0;
}
