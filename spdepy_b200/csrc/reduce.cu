// reduce.cu -- K8/K9/K11: likelihood reductions and the gradient contraction.
//
//   spde_q_apply           y = Q x on the slot layout (stencil apply; no index arrays)
//   spde_dot               deterministic two-stage sum(x .* y), warp shuffles
//   spde_sddmm             W[slot,node] = alpha * <X[node,:], Y[nbr(node,slot),:]> on the pattern of Q
//                          (turns the Hutchinson traces of advection_diffusion2D.py:204-206 into
//                          weights on the pattern; the same weights come from the Takahashi inverse)
//   spde_assembly_adjoint  d sum(W .* Q) / d A9, d Qs, d Q0 -- the transpose of K3, so that
//                          sum(W .* dQ_i) for *every* parameter i costs one pass instead of one
//                          sparse n x n matrix per parameter (advection_diffusion2D.py:119-182)
//   spde_gemv_t            out = B^T u for the spline-basis chain rule
// All are HBM-bound streaming kernels.
#include "common.cuh"

namespace spde {

__device__ __forceinline__ double warp_sum(double s)
{
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    return s;
}

__device__ __forceinline__ double block_sum(double s, double *sh)
{
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        s = warp_sum(s);
    }
    return s;   // valid in thread 0
}

__global__ void k_q_apply(Geo g, const double *__restrict__ Q, const double *__restrict__ X, int k,
                          double *__restrict__ Y)
{
    const long long n = (long long)g.M * g.N * g.T;
    const long long total = n * k;
    const int ns = g.nslots();
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int node = (int)(e / k), p = (int)(e % k);
        double s = 0.0;
        for (int q = 0; q < ns; q++) {
            const int c = g.slot_nbr(node, q);
            if (c < 0) continue;
            s += Q[(long long)q * n + node] * X[(long long)c * k + p];
        }
        Y[e] = s;
    }
}

__global__ void k_dot_partial(const double *__restrict__ X, const double *__restrict__ Y, long long len,
                              double *__restrict__ partial)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x)
        s += X[e] * Y[e];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// sum_node w[node] * sum_p X[node,p]*Y[node,p]   (tau-gradient of advection_diffusion2D.py:207)
__global__ void k_wdot_partial(const double *__restrict__ X, const double *__restrict__ Y, const double *__restrict__ w,
                               long long n, int k, double *__restrict__ partial)
{
    __shared__ double sh[32];
    const long long len = n * k;
    double s = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x)
        s += w[e / k] * X[e] * Y[e];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// sum (data[i,p] - mu[obs[i],p])^2   (advection_diffusion2D.py:198)
__global__ void k_resid_partial(const double *__restrict__ data, const double *__restrict__ mu, const long long *__restrict__ obs,
                                long long nobs, int r, double *__restrict__ partial)
{
    __shared__ double sh[32];
    const long long len = nobs * r;
    double s = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x) {
        const long long i = e / r;
        const int p = (int)(e % r);
        const double d = data[e] - mu[obs[i] * r + p];
        s += d * d;
    }
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}
// b[obs[i], :] += data[i, :] * tau   (S^T data * tau, advection_diffusion2D.py:194)
__global__ void k_scatter_obs(const double *__restrict__ data, const long long *__restrict__ obs, long long nobs, int r,
                              double tau, double *__restrict__ b)
{
    const long long len = nobs * r;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < len; e += (long long)gridDim.x * blockDim.x)
        atomicAdd(&b[obs[e / r] * r + e % r], data[e] * tau);
}
// diag slot of Q += tau * cnt   (Model.update, model.py:120-124)
__global__ void k_add_diag(double *__restrict__ Qdiag, const double *__restrict__ cnt, double tau, long long n)
{
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        Qdiag[e] += cnt[e] * tau;
}

__global__ void k_final(const double *__restrict__ partial, int m, double *__restrict__ out)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) s += partial[j];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[0] = s;
}

// one warp per node: lanes stride over the k probe columns, one shuffle reduction per slot.  For k <= 128 (the reference's
// nh1 = 100) the node's own row stays in registers and four slots are reduced together, so that four independent
// load -> FMA -> shuffle chains are in flight per warp instead of one (the kernel is latency-bound: every dot product is a
// chain of L2 loads, dependent FP64 adds and five shuffle rounds).
__global__ void k_sddmm(Geo g, const double *__restrict__ X, const double *__restrict__ Y, int k, double alpha,
                        int accumulate, double *__restrict__ W)
{
    const long long n = (long long)g.M * g.N * g.T;
    const int ns = g.nslots();
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    if (k <= 128) {
        for (long long node = warp0; node < n; node += nwarps) {
            double xr[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int p = lane + 32 * u; xr[u] = p < k ? X[node * k + p] : 0.0; }
            for (int q0 = 0; q0 < ns; q0 += 4) {
                int c[4];
                double s[4];
#pragma unroll
                for (int j = 0; j < 4; j++) c[j] = (q0 + j < ns) ? g.slot_nbr((int)node, q0 + j) : -1;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    s[j] = 0.0;
                    if (c[j] >= 0) {
                        const double *y = Y + (long long)c[j] * k;
#pragma unroll
                        for (int u = 0; u < 4; u++) { const int p = lane + 32 * u; if (p < k) s[j] += xr[u] * y[p]; }
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1)
#pragma unroll
                    for (int j = 0; j < 4; j++) s[j] += __shfl_down_sync(0xffffffffu, s[j], o);
                if (lane == 0) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (c[j] >= 0) {
                            const long long o = (long long)(q0 + j) * n + node;
                            W[o] = accumulate ? W[o] + alpha * s[j] : alpha * s[j];
                        }
                }
            }
        }
        return;
    }
    for (long long node = warp0; node < n; node += nwarps) {
        for (int q = 0; q < ns; q++) {
            const int c = g.slot_nbr((int)node, q);
            if (c < 0) continue;
            double s = 0.0;
            for (int p = lane; p < k; p += 32) s += X[node * k + p] * Y[(long long)c * k + p];
            s = warp_sum(s);
            if (lane == 0) {
                const long long o = (long long)q * n + node;
                W[o] = accumulate ? W[o] + alpha * s : alpha * s;
            }
        }
    }
}

// Many probe columns (the Hutchinson estimator, nh1 = 100): patch-tiled version.  One CTA owns an 8 x 8 patch of cells of one
// time slice; the probe rows of the patch (X) and of its halo (Y: radius 2 in the slice, radius 1 or 2 in the slices before
// and after, as the pattern says) are staged in shared memory 32 columns at a time, and every thread accumulates a
// handful of (node, slot) dot products from there.  The warp-per-node kernel above fetches each Y row once per node that
// touches it (43 times, from L2); here it is fetched once per patch.  Same products, summed over the columns in order.
constexpr int SD_PW = 8, SD_PH = 8, SD_KC = 32, SD_THREADS = 256, SD_MAXPAIRS = 19;     // 75 slots x 64 nodes / 256 threads
__global__ void __launch_bounds__(SD_THREADS) k_sddmm_tiled(Geo g, const double *__restrict__ X, const double *__restrict__ Y, int k,
                                                            double alpha, int accumulate, double *__restrict__ W)
{
    extern __shared__ __align__(16) double sd_smem[];
    const int ns = g.nslots(), Ns = g.M * g.N;
    const long long n = (long long)Ns * g.T;
    const int rad_t = (g.T == 1) ? 0 : (g.pat == 1 ? 2 : 1);       // halo radius in the neighbouring slices (0: none)
    const int w0 = SD_PW + 4, h0 = SD_PH + 4, w1 = SD_PW + 2 * rad_t, h1 = SD_PH + 2 * rad_t;
    const int nh0 = w0 * h0, nh1 = rad_t ? w1 * h1 : 0;
    const int nhalo = nh0 + 2 * nh1;                               // [slice t | slice t-1 | slice t+1]
    double *Xs = sd_smem;                                          // [64][KC+1]
    double *Ys = Xs + SD_PW * SD_PH * (SD_KC + 1);                 // [nhalo][KC+1]
    int *hnode = reinterpret_cast<int *>(Ys + (size_t)nhalo * (SD_KC + 1));
    const int px = (g.M + SD_PW - 1) / SD_PW, py = (g.N + SD_PH - 1) / SD_PH;
    const int bid = blockIdx.x;
    const int t = bid / (px * py), pj = (bid / px) % py, pi = bid % px;
    const int i0 = pi * SD_PW, j0 = pj * SD_PH;
    const int tid = threadIdx.x;
    // global node of every halo cell (-1: outside the mesh / the time range)
    for (int h = tid; h < nhalo; h += SD_THREADS) {
        int dt, hh = h, ww, rad;
        if (h < nh0) { dt = 0; ww = w0; rad = 2; }
        else if (h < nh0 + nh1) { dt = -1; hh = h - nh0; ww = w1; rad = rad_t; }
        else { dt = 1; hh = h - nh0 - nh1; ww = w1; rad = rad_t; }
        const int hi = hh % ww, hj = hh / ww;
        const int tt = t + dt;
        int node = -1;
        if (tt >= 0 && tt < g.T) {
            // (cells of the patch row / column beyond the mesh edge have no node; wrapped neighbours under bc = 2)
            int ii = i0 + hi - rad, jj = j0 + hj - rad;
            if (g.bc == 2) {
                ii = ii < 0 ? ii + g.M : (ii >= g.M ? ii - g.M : ii);
                jj = jj < 0 ? jj + g.N : (jj >= g.N ? jj - g.N : jj);
            }
            if (ii >= 0 && ii < g.M && jj >= 0 && jj < g.N) node = tt * Ns + jj * g.M + ii;
        }
        hnode[h] = node;
    }
    // pairs (slot q, patch cell c) of this thread: e = tid + 256 m, c = e % 64 (fastest: coalesced rows of W), q = e / 64
    const int npairs = ns * SD_PW * SD_PH;
    int prow[SD_MAXPAIRS];
    double acc[SD_MAXPAIRS];
#pragma unroll
    for (int m = 0; m < SD_MAXPAIRS; m++) {
        acc[m] = 0.0;
        prow[m] = -1;
        const int e = tid + SD_THREADS * m;
        if (e < npairs) {
            const int c = e % (SD_PW * SD_PH), q = e / (SD_PW * SD_PH);
            const int il = c % SD_PW, jl = c / SD_PW;
            if (i0 + il < g.M && j0 + jl < g.N) {
                int dt, dj, di;
                g.slot_offset(q, dt, dj, di);
                // the neighbour the pattern names must exist (same rule as slot_nbr), and it is a halo cell by construction
                const int node = t * Ns + (j0 + jl) * g.M + (i0 + il);
                if (g.slot_nbr(node, q) >= 0) {
                    if (dt == 0) prow[m] = (jl + dj + 2) * w0 + (il + di + 2);
                    else prow[m] = nh0 + (dt > 0 ? nh1 : 0) + (jl + dj + rad_t) * w1 + (il + di + rad_t);
                }
            }
        }
    }
    for (int k0 = 0; k0 < k; k0 += SD_KC) {
        const int kc = min(SD_KC, k - k0);
        __syncthreads();                      // hnode ready / previous chunk consumed
        for (int e = tid; e < SD_PW * SD_PH * SD_KC; e += SD_THREADS) {
            const int c = e / SD_KC, p = e % SD_KC;
            const int il = c % SD_PW, jl = c / SD_PW;
            double v = 0.0;
            if (p < kc && i0 + il < g.M && j0 + jl < g.N) v = X[((long long)t * Ns + (j0 + jl) * g.M + (i0 + il)) * k + k0 + p];
            Xs[c * (SD_KC + 1) + p] = v;
        }
        for (int e = tid; e < nhalo * SD_KC; e += SD_THREADS) {
            const int h = e / SD_KC, p = e % SD_KC;
            const int node = hnode[h];
            Ys[h * (SD_KC + 1) + p] = (node >= 0 && p < kc) ? Y[(long long)node * k + k0 + p] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int m = 0; m < SD_MAXPAIRS; m++) {
            if (prow[m] < 0) continue;
            const int c = (tid + SD_THREADS * m) % (SD_PW * SD_PH);
            const double *xr = Xs + c * (SD_KC + 1), *yr = Ys + prow[m] * (SD_KC + 1);
            double s = acc[m];
#pragma unroll 8
            for (int p = 0; p < SD_KC; p++) s += xr[p] * yr[p];
            acc[m] = s;
        }
    }
#pragma unroll
    for (int m = 0; m < SD_MAXPAIRS; m++) {
        if (prow[m] < 0) continue;
        const int e = tid + SD_THREADS * m;
        const int c = e % (SD_PW * SD_PH), q = e / (SD_PW * SD_PH);
        const long long node = (long long)t * Ns + (j0 + c / SD_PW) * g.M + (i0 + c % SD_PW);
        const long long o = (long long)q * n + node;
        W[o] = accumulate ? W[o] + alpha * acc[m] : alpha * acc[m];
    }
}

// single-column case (the -1/2 mu mu^T term of every gradient): one thread per node, slot loop unrolled by the
// compiler, every slot one coalesced read-modify-write stream
__global__ void k_sddmm1(Geo g, const double *__restrict__ X, const double *__restrict__ Y, double alpha,
                         int accumulate, double *__restrict__ W)
{
    const long long n = (long long)g.M * g.N * g.T;
    const int ns = g.nslots();
    for (long long node = (long long)blockIdx.x * blockDim.x + threadIdx.x; node < n; node += (long long)gridDim.x * blockDim.x) {
        const double x = alpha * X[node];
        for (int q = 0; q < ns; q++) {
            const int c = g.slot_nbr((int)node, q);
            const long long o = (long long)q * n + node;
            if (c < 0) { if (!accumulate) W[o] = 0.0; continue; }
            const double v = x * Y[c];
            W[o] = accumulate ? W[o] + v : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// adjoint of the assembly.  Phase 1 (space-time only): sums of W over the time blocks.
//   Wd[q][k] = sum_{t>=1} W[9+q][(k,t)]      weights of A^T d A
//   Wu[s][k] = sum_{t<=T-2} W[34+s][(k,t)]   weights of -(Qs iV) A
//   Wl[s][k] = sum_{t>=1} W[s][(k,t)]        weights of -A^T iV Qs  (row k, column k+off(s) at t-1)
//   wq[k]    = sum_{t<=T-2} W[21][(k,t)]     weights of the "+Qs" diagonal
__global__ void k_adj_time_reduce(Geo g, const double *__restrict__ W, double *__restrict__ Wd, double *__restrict__ Wu,
                                  double *__restrict__ Wl, double *__restrict__ wq, double *__restrict__ GQ0, double q0scale)
{
    const int Ns = g.M * g.N;
    const long long n = (long long)Ns * g.T;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;   // 0..42
    if (k >= Ns) return;
    double s = 0.0;
    if (q < 9) {
        for (int t = 1; t < g.T; t++) s += W[(long long)q * n + (long long)t * Ns + k];
        Wl[(long long)q * Ns + k] = s;
    } else if (q < 34) {
        for (int t = 1; t < g.T; t++) s += W[(long long)q * n + (long long)t * Ns + k];
        Wd[(long long)(q - 9) * Ns + k] = s;
        GQ0[(long long)(q - 9) * Ns + k] = q0scale * W[(long long)q * n + k];
        if (q == 21) {
            double u = 0.0;
            for (int t = 0; t < g.T - 1; t++) u += W[(long long)q * n + (long long)t * Ns + k];
            wq[k] = u;
        }
    } else {
        for (int t = 0; t < g.T - 1; t++) s += W[(long long)q * n + (long long)t * Ns + k];
        Wu[(long long)(q - 34) * Ns + k] = s;
    }
}

// Separable model Q = Qt (x) Qs (seperable_spatial_temporal2D.py:82): weights W on the 75-slot pattern -> weights on Qs,
//   Wd[q][k] = sum_t sum_dt Qt[t,t+dt] W[(dt+1)*25+q][(k,t)],   Qt tridiagonal with diagonal (d0, d1, ..., d1, d0), off-diagonal e
__global__ void k_kron_reduce(Geo g, const double *__restrict__ W, double d0, double d1, double e, double *__restrict__ Wd)
{
    const int Ns = g.M * g.N;
    const long long n = (long long)Ns * g.T;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int q = blockIdx.y;   // 0..24
    if (k >= Ns) return;
    double s = 0.0;
    for (int t = 0; t < g.T; t++) {
        const long long node = (long long)t * Ns + k;
        const double dd = (t == 0 || t == g.T - 1) ? d0 : d1;
        double v = dd * W[(long long)(25 + q) * n + node];
        if (t > 0) v += e * W[(long long)q * n + node];
        if (t < g.T - 1) v += e * W[(long long)(50 + q) * n + node];
        s += v;
    }
    Wd[(long long)q * Ns + k] = s;
}

__device__ __forceinline__ double qs_val(double V, double kap) { const double As = V * kap; return (As * (1.0 / V)) * As; }

// Phase 2: per cell c.  With k_s = c + off(s):
//   GA[c,s]  = cs*( d_c * sum_s'' A[c,s''] (Wd[k_s -> k_s''] + Wd[k_s'' -> k_s])
//                   - Qs_c iV (Wu[s][c] + Wl[8-s][k_s]) )
//   Gq[c]    = cs*( iV^2 sum_{s,s''} A[c,s] A[c,s''] Wd[k_s -> k_s''] + wq[c]
//                   - iV sum_s A[c,s] (Wu[s][c] + Wl[8-s][k_s]) )          (d S / d Qs_c)
// Spatial (timed=0): Wd = W, d = iV, cs = 1, no Wu/Wl/wq terms.
__global__ void k_adj_cell(Geo g, int timed, const double *__restrict__ Wd, const double *__restrict__ Wu,
                           const double *__restrict__ Wl, const double *__restrict__ wq, const double *__restrict__ A,
                           const double *__restrict__ kappa, int kvar, double V, double cs,
                           double *__restrict__ GA, double *__restrict__ Gq)
{
    const int Ns = g.M * g.N;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Ns) return;
    const int i = c % g.M, j = c / g.M;
    const double iV = 1.0 / V;
    const double qs = timed ? qs_val(V, kvar ? kappa[c] : kappa[0]) : 0.0;
    const double d = timed ? qs * iV * iV : iV;
    int nb[9];
    double a[9];
#pragma unroll
    for (int s = 0; s < 9; s++) {
        nb[s] = g.nbr(i, j, s % 3 - 1, s / 3 - 1);
        a[s] = A[(long long)s * Ns + c];
    }
    double gq = 0.0;
    for (int s = 0; s < 9; s++) {
        double ga = 0.0;
        if (nb[s] >= 0) {
            const int si = s % 3 - 1, sj = s / 3 - 1;
            double acc = 0.0;
            for (int s2 = 0; s2 < 9; s2++) {
                if (nb[s2] < 0) continue;
                const int di = (s2 % 3 - 1) - si, dj = (s2 / 3 - 1) - sj;   // k_s -> k_s2
                const double w1 = Wd[(long long)((dj + 2) * 5 + (di + 2)) * Ns + nb[s]];
                const double w2 = Wd[(long long)((2 - dj) * 5 + (2 - di)) * Ns + nb[s2]];
                acc += a[s2] * (w1 + w2);
                gq += a[s] * a[s2] * w1;
            }
            ga = d * acc;
            if (timed) {
                const double wul = Wu[(long long)s * Ns + c] + Wl[(long long)(8 - s) * Ns + nb[s]];
                ga -= qs * iV * wul;
                gq -= V * a[s] * wul;   // scaled by iV^2 below
            }
        }
        GA[(long long)s * Ns + c] = cs * ga;
    }
    if (Gq) Gq[c] = timed ? cs * (iV * iV * gq + wq[c]) : 0.0;
}

__global__ void k_gemv_t(const double *__restrict__ B, const double *__restrict__ u, int rows, int cols,
                         double *__restrict__ partial)
{
    // one block column-slice: blockIdx.y = column, blocks over rows; partial[col*gridDim.x + blockIdx.x]
    __shared__ double sh[32];
    const int col = blockIdx.y;
    double s = 0.0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += gridDim.x * blockDim.x)
        s += B[(long long)r * cols + col] * u[r];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) partial[(long long)col * gridDim.x + blockIdx.x] = s;
}
__global__ void k_gemv_final(const double *__restrict__ partial, int nb, double *__restrict__ out)
{
    __shared__ double sh[32];
    double s = 0.0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) s += partial[(long long)blockIdx.x * nb + j];
    s = block_sum(s, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// 64 KiB of device scratch for the reductions, one per device (a process may drive several GPUs: Engine caches one
// engine per device); g_scratch is the scratch of the CURRENT device, selected by ensure_scratch() on every entry
constexpr int kMaxDevices = 64;
static double *g_scratch_of[kMaxDevices] = {nullptr};
static thread_local double *g_scratch = nullptr;
static int ensure_scratch()
{
    int dev = 0;
    SPDE_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { set_error("more than 64 devices"); return SPDE_ERR_ARG; }
    if (!g_scratch_of[dev]) SPDE_CUDA_CHECK(cudaMalloc((void **)&g_scratch_of[dev], 8192 * sizeof(double)));
    g_scratch = g_scratch_of[dev];
    return SPDE_OK;
}

}  // namespace spde

using namespace spde;

extern "C" int spde_q_apply(int M, int N, int T, int bc, const double *d_Q, const double *d_X, int k, double *d_Y, void *stream)
{
    Geo g = geo_from_abi(M, N, T, bc);
    const long long total = (long long)M * N * T * k;
    k_q_apply<<<(int)std::min<long long>((total + 255) / 256, 148 * 32), 256, 0, (cudaStream_t)stream>>>(g, d_Q, d_X, k, d_Y);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

// second stage of a two-stage reduction: the sum of the nb partials goes to host memory (with a stream synchronise) or,
// for the *_dev entry points, straight to a device address -- no copy, no synchronise: the likelihood / gradient
// scalars of one evaluation are collected in one device vector and read back once
static int finish_sum(double *out, bool device_out, cudaStream_t st, int nb)
{
    count_launch(2);
    k_final<<<1, 1024, 0, st>>>(g_scratch, nb, device_out ? out : g_scratch + nb);
    SPDE_LAUNCH_CHECK();
    if (device_out) return SPDE_OK;
    SPDE_CUDA_CHECK(cudaMemcpyAsync(out, g_scratch + nb, sizeof(double), cudaMemcpyDeviceToHost, st));
    SPDE_CUDA_CHECK(cudaStreamSynchronize(st));
    return SPDE_OK;
}

// grid of a streaming reduction: enough CTAs for the machine on long vectors, a handful on short ones (the 30x30 mesh)
static inline int reduce_blocks(int64_t len) { return (int)std::max<int64_t>(1, std::min<int64_t>(1184, (len + 1023) / 1024)); }

static int dot_impl(const double *d_X, const double *d_Y, int64_t len, double *out, bool device_out, void *stream)
{
    int rc = ensure_scratch();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = reduce_blocks(len);
    k_dot_partial<<<nb, 256, 0, st>>>(d_X, d_Y, len, g_scratch);
    return finish_sum(out, device_out, st, nb);
}
extern "C" int spde_dot(const double *d_X, const double *d_Y, int64_t len, double *h_out, void *stream)
{
    return dot_impl(d_X, d_Y, len, h_out, false, stream);
}
extern "C" int spde_dot_dev(const double *d_X, const double *d_Y, int64_t len, double *d_out, void *stream)
{
    return dot_impl(d_X, d_Y, len, d_out, true, stream);
}

extern "C" int spde_sddmm(int M, int N, int T, int bc, const double *d_X, const double *d_Y, int k, double alpha,
                          int accumulate, double *d_W, void *stream)
{
    Geo g = geo_from_abi(M, N, T, bc);
    const long long n = (long long)M * N * T;
    if (k == 1) {
        k_sddmm1<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(g, d_X, d_Y, alpha, accumulate, d_W);
    } else if (k >= 16 && !getenv("SPDE_SDDMM_WARP")) {
        // patch-tiled kernel: shared memory = probe rows of the patch and of its halo, 32 columns at a time
        const int rad_t = (T == 1) ? 0 : (g.pat == 1 ? 2 : 1);
        const int nhalo = (SD_PW + 4) * (SD_PH + 4) + (rad_t ? 2 * (SD_PW + 2 * rad_t) * (SD_PH + 2 * rad_t) : 0);
        const size_t smem = (size_t)(SD_PW * SD_PH + nhalo) * (SD_KC + 1) * sizeof(double) + (size_t)nhalo * sizeof(int);
        static bool attr_done[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        dev &= 63;
        if (!attr_done[dev]) {
            SPDE_CUDA_CHECK(cudaFuncSetAttribute(k_sddmm_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr_done[dev] = true;
        }
        const int px = (M + SD_PW - 1) / SD_PW, py = (N + SD_PH - 1) / SD_PH;
        k_sddmm_tiled<<<px * py * T, SD_THREADS, smem, (cudaStream_t)stream>>>(g, d_X, d_Y, k, alpha, accumulate, d_W);
    } else {
        const long long blocks = std::min<long long>((n * 32 + 255) / 256, 148 * 16);
        k_sddmm<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(g, d_X, d_Y, k, alpha, accumulate, d_W);
    }
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_assembly_adjoint(int M, int N, int T, int bc, const double *d_W, const double *d_A9,
                                     const double *d_kappa, int kvar, double V, double sigma, double dt, int timed,
                                     double *d_work /* 44*Ns doubles */, double *d_GA9, double *d_Gq, double *d_GQ0_25,
                                     void *stream)
{
    Geo g{M, N, T, bc};
    const int Ns = M * N;
    cudaStream_t st = (cudaStream_t)stream;
    if (timed == 1) {
        if (!d_work || !d_GQ0_25 || !d_Gq) { set_error("spde_assembly_adjoint: work buffers required"); return SPDE_ERR_ARG; }
        double *Wd = d_work, *Wu = Wd + (size_t)25 * Ns, *Wl = Wu + (size_t)9 * Ns, *wq = Wl + (size_t)9 * Ns;
        const double cs = 1 / (dt * sigma);
        k_adj_time_reduce<<<dim3(cdiv(Ns, 128), 43), 128, 0, st>>>(g, d_W, Wd, Wu, Wl, wq, d_GQ0_25, cs * (sigma * dt));
        SPDE_LAUNCH_CHECK();
        Geo g2{M, N, 1, bc};
        k_adj_cell<<<cdiv(Ns, 128), 128, 0, st>>>(g2, 1, Wd, Wu, Wl, wq, d_A9, d_kappa, kvar, V, cs, d_GA9, d_Gq);
    } else if (timed == 2) {
        // weights given directly on the pattern of B = A^T (Qs/V^2) A (Q25 layout): the time-collapsed prior,
        // logdet Q = logdet Q0 + (T-1) (Ns log(1/(dt sigma)) + logdet B)
        if (!d_work || !d_Gq) { set_error("spde_assembly_adjoint: work buffers required"); return SPDE_ERR_ARG; }
        SPDE_CUDA_CHECK(cudaMemsetAsync(d_work, 0, sizeof(double) * (size_t)19 * Ns, st));
        double *Wu = d_work, *Wl = Wu + (size_t)9 * Ns, *wq = Wl + (size_t)9 * Ns;
        Geo g2{M, N, 1, bc};
        k_adj_cell<<<cdiv(Ns, 128), 128, 0, st>>>(g2, 1, d_W, Wu, Wl, wq, d_A9, d_kappa, kvar, V, 1.0, d_GA9, d_Gq);
    } else {
        k_adj_cell<<<cdiv(Ns, 128), 128, 0, st>>>(g, 0, d_W, nullptr, nullptr, nullptr, d_A9, d_kappa, kvar, V, 1.0, d_GA9, d_Gq);
    }
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_kron_reduce(int M, int N, int T, int bc, const double *d_W75, double d0, double d1, double e,
                                double *d_Wd25, void *stream)
{
    Geo g = geo_from_abi(M, N, T, bc | (1 << 8));
    if (T < 2) { set_error("spde_kron_reduce: T >= 2 required"); return SPDE_ERR_ARG; }
    k_kron_reduce<<<dim3(cdiv(M * N, 128), 25), 128, 0, (cudaStream_t)stream>>>(g, d_W75, d0, d1, e, d_Wd25);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_gemv_t(const double *d_B, const double *d_u, int rows, int cols, double *d_out, void *stream)
{
    int rc = ensure_scratch();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = 32;
    if ((long long)cols * nb > 8192) { set_error("spde_gemv_t: too many columns"); return SPDE_ERR_ARG; }
    k_gemv_t<<<dim3(nb, cols), 256, 0, st>>>(d_B, d_u, rows, cols, g_scratch);
    k_gemv_final<<<cols, 64, 0, st>>>(g_scratch, nb, d_out);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

static int wdot_impl(const double *d_X, const double *d_Y, const double *d_w, int64_t n, int k, double *out, bool device_out, void *stream)
{
    int rc = ensure_scratch();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = reduce_blocks(n * k);
    k_wdot_partial<<<nb, 256, 0, st>>>(d_X, d_Y, d_w, n, k, g_scratch);
    return finish_sum(out, device_out, st, nb);
}
extern "C" int spde_wdot(const double *d_X, const double *d_Y, const double *d_w, int64_t n, int k, double *h_out, void *stream)
{
    return wdot_impl(d_X, d_Y, d_w, n, k, h_out, false, stream);
}
extern "C" int spde_wdot_dev(const double *d_X, const double *d_Y, const double *d_w, int64_t n, int k, double *d_out, void *stream)
{
    return wdot_impl(d_X, d_Y, d_w, n, k, d_out, true, stream);
}

static int resid_impl(const double *d_data, const double *d_mu, const int64_t *d_obs, int64_t nobs, int r, double *out, bool device_out,
                      void *stream)
{
    int rc = ensure_scratch();
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = reduce_blocks(nobs * r);
    k_resid_partial<<<nb, 256, 0, st>>>(d_data, d_mu, (const long long *)d_obs, nobs, r, g_scratch);
    return finish_sum(out, device_out, st, nb);
}
extern "C" int spde_residual_ss(const double *d_data, const double *d_mu, const int64_t *d_obs, int64_t nobs, int r,
                                double *h_out, void *stream)
{
    return resid_impl(d_data, d_mu, d_obs, nobs, r, h_out, false, stream);
}
extern "C" int spde_residual_ss_dev(const double *d_data, const double *d_mu, const int64_t *d_obs, int64_t nobs, int r,
                                    double *d_out, void *stream)
{
    return resid_impl(d_data, d_mu, d_obs, nobs, r, d_out, true, stream);
}

extern "C" int spde_scatter_obs(const double *d_data, const int64_t *d_obs, int64_t nobs, int r, double tau, double *d_b, void *stream)
{
    const long long len = (long long)nobs * r;
    k_scatter_obs<<<(int)std::min<long long>((len + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(
        d_data, (const long long *)d_obs, nobs, r, tau, d_b);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_add_diag(double *d_Qdiag, const double *d_cnt, double tau, int64_t n, void *stream)
{
    k_add_diag<<<(int)std::min<long long>((n + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream>>>(d_Qdiag, d_cnt, tau, n);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}
