// assemble.cu -- K2/K3: finite-volume stencils and precision assembly in fixed slot layouts.
//
// Compiled with -fmad=false: every product and sum below is a separately rounded IEEE operation
// in the reference's order, so the values match the reference's SciPy/C++ pipeline bit for bit
// (SURVEY.md App. A.4 "Operation order").  These kernels are HBM-bound streaming kernels: one
// thread per cell / node, slot-major arrays so that each slot store is a 256-byte warp
// transaction; grids are sized in whole multiples of the SM count where that matters.
#include "common.cuh"

namespace spde {

// reference slot order C,E,W,N,S,NE,SW,NW,SE (AcH_2D_b1.cpp:121-129) -> offsets
__constant__ int c_rdi[9] = {0, 1, -1, 0, 0, 1, -1, -1, 1};
__constant__ int c_rdj[9] = {0, 0, 0, 1, -1, 1, -1, 1, -1};

__device__ __forceinline__ int gslot(int di, int dj) { return (dj + 1) * 3 + (di + 1); }

// ---------------------------------------------------------------------------------------------
// K2a: diffusion stencil  (AcH_2D_b{1,3}.cpp, AH_2D_b{1,2,3}.cpp)
__global__ void k_ah_stencil(Geo g, double hx, double hy, const double *__restrict__ H, int face,
                             double *__restrict__ out)
{
    const int Ns = g.M * g.N;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const int i = k % g.M, j = k / g.M;
    double w[9];
    double rem = 0.0;
    bool del[9];
#pragma unroll
    for (int s = 0; s < 9; s++) del[s] = false;

    if (!face && g.bc == 3) {
        // AcH_2D_b3.cpp:36-44
        const double H00 = H[0], H01 = H[1], H10 = H[2], H11 = H[3];
        const double hxy = H01 + H10;
        w[0] = -2.0 * hy / hx * H00 - 2.0 * hx / hy * H11 + 0.0;
        w[1] = hy / hx * H00; w[2] = hy / hx * H00;
        w[3] = hx / hy * H11; w[4] = hx / hy * H11;
        w[5] = 1.0 / 4.0 * hxy; w[6] = 1.0 / 4.0 * hxy;
        w[7] = -1.0 / 4.0 * hxy; w[8] = -1.0 / 4.0 * hxy;
    } else {
        double W00, E00, W10, E10, S01, N01, S11, N11;
        if (!face) {
            // AcH_2D_b1.cpp:42-68 (the reference reads H[1][0] for every cross term)
            W00 = H[0]; E00 = H[0]; W10 = H[2]; E10 = H[2];
            S01 = H[2]; N01 = H[2]; S11 = H[3]; N11 = H[3];
            if (i == 0) { W00 = 0.0; W10 = 0.0; } else if (i == g.M - 1) { E00 = 0.0; E10 = 0.0; }
            if (j == 0) { S11 = 0.0; S01 = 0.0; } else if (j == g.N - 1) { N11 = 0.0; N01 = 0.0; }
        } else {
            const double *h = H + (size_t)k * 16;     // [face W,E,S,N][a][b]
            W00 = h[0]; W10 = h[2]; E00 = h[4]; E10 = h[6];
            S01 = h[9]; S11 = h[11]; N01 = h[13]; N11 = h[15];
            if (g.bc == 1 && k == 0) {
                // AH_2D_b1.cpp:34-52: the zeroing uses the previous cell's index, so it only
                // ever lands on the faces of cell 0 before they are read.
                W00 = 0.0; W10 = 0.0;
                S11 = 0.0; S01 = 0.0;
            }
        }
        w[1] = hy / hx * E00 + 1.0 / 4.0 * (N01 - S01);
        w[2] = hy / hx * W00 - 1.0 / 4.0 * (N01 - S01);
        w[3] = hx / hy * N11 + 1.0 / 4.0 * (E10 - W10);
        w[4] = hx / hy * S11 - 1.0 / 4.0 * (E10 - W10);
        w[5] = 1.0 / 4.0 * (N01 + E10);
        w[6] = 1.0 / 4.0 * (S01 + W10);
        w[7] = -1.0 / 4.0 * (N01 + W10);
        w[8] = -1.0 / 4.0 * (S01 + E10);
        if (g.bc == 1) {
            // AcH_2D_b1.cpp:71-118: slots whose clamped target is the cell itself are deleted
            // and their coefficient is moved into `rem` (literal operation order).
            const bool xe = (i == g.M - 1), xw = (i == 0), yn = (j == g.N - 1), ys = (j == 0);
            if (xe) { rem = rem + hy / hx * E00 + 1.0 / 4.0 * (N01 - S01); del[1] = true; }
            if (xw) { rem = rem + hy / hx * W00 - 1.0 / 4.0 * (N01 - S01); del[2] = true; }
            if (yn) { rem = rem + hx / hy * N11 + 1.0 / 4.0 * (E10 - W10); del[3] = true; }
            if (ys) { rem = rem + hx / hy * S11 - 1.0 / 4.0 * (E10 - W10); del[4] = true; }
            if (xe && yn) { rem = rem + 1.0 / 4.0 * (N01 + E10); del[5] = true; }
            if (xw && ys) { rem = rem + 1.0 / 4.0 * (S01 + W10); del[6] = true; }
            if (xw && yn) { rem = rem - 1.0 / 4.0 * (N01 + W10); del[7] = true; }
            if (xe && ys) { rem = rem - 1.0 / 4.0 * (S01 + E10); del[8] = true; }
        }
        w[0] = -hy / hx * (E00 + W00) - hx / hy * (N11 + S11) + rem;
    }

    double o[9];
#pragma unroll
    for (int s = 0; s < 9; s++) o[s] = 0.0;
    o[4] = w[0];
#pragma unroll
    for (int s = 1; s < 9; s++) {
        if (del[s]) continue;
        int di = c_rdi[s], dj = c_rdj[s];
        if (g.bc == 1) {
            // clamp (AcH_2D_b1.cpp:51-68); a clamped corner lands on an edge neighbour and is
            // summed there as a duplicate by the COO->CSC conversion (advection_diffusion2D.py:258)
            int ii = min(max(i + di, 0), g.M - 1), jj = min(max(j + dj, 0), g.N - 1);
            o[gslot(ii - i, jj - j)] += w[s];
        } else if (g.bc == 3) {
            if (g.nbr(i, j, di, dj) >= 0) o[gslot(di, dj)] = w[s];   // AcH_2D_b3.cpp:56-73
        } else {
            o[gslot(di, dj)] = w[s];
        }
    }
#pragma unroll
    for (int s = 0; s < 9; s++) out[(size_t)s * Ns + k] = o[s];
}

// ---------------------------------------------------------------------------------------------
// K2b: upwind advection stencil  (Acw_2D_b{1,2,3}.cpp, Aw_2D_b{1,2,3}.cpp)
__global__ void k_aw_stencil(Geo g, double hx, double hy, const double *__restrict__ G,
                             const double *__restrict__ dG, int face, int diff, int nan_to_zero,
                             double *__restrict__ out)
{
    const int Ns = g.M * g.N;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const int i = k % g.M, j = k / g.M;
    const bool xe = (i == g.M - 1), xw = (i == 0), yn = (j == g.N - 1), ys = (j == 0);
    double v[5];   // C,E,W,N,S
    if (!face) {
        const double G0 = G[0], G1 = G[1];
        if (diff == 1) {
            v[0] = G0 / fabs(G0) * hy;
            v[1] = -(G0 / fabs(G0) - 1.0) * hy / 2;
            v[2] = -(G0 / fabs(G0) + 1.0) * hy / 2;
            v[3] = 0.0; v[4] = 0.0;
        } else if (diff == 2) {
            v[0] = G1 / fabs(G1) * hx;
            v[1] = 0.0; v[2] = 0.0;
            v[3] = -(G1 / fabs(G1) - 1.0) * hx / 2;
            v[4] = -(G1 / fabs(G1) + 1.0) * hx / 2;
        } else {
            v[0] = fabs(G0) * hy + fabs(G1) * hx;
            v[1] = -(fabs(G0) - G0) * hy / 2;
            v[2] = -(fabs(G0) + G0) * hy / 2;
            v[3] = -(fabs(G1) - G1) * hx / 2;
            v[4] = -(fabs(G1) + G1) * hx / 2;
        }
        if (g.bc == 1) {   // Acw_2D_b1.cpp:63-80, applied in every diff mode (App. C-4)
            if (xw) v[0] -= fabs(G0) * hy / 2; else if (xe) v[0] -= fabs(G0) * hy / 2;
            if (ys) v[0] -= fabs(G1) * hx / 2; else if (yn) v[0] -= fabs(G1) * hx / 2;
        }
    } else {
        double g0 = G[(size_t)k * 4 + 0], g1 = G[(size_t)k * 4 + 1], g2 = G[(size_t)k * 4 + 2], g3 = G[(size_t)k * 4 + 3];
        double d0 = 0, d1 = 0, d2 = 0, d3 = 0;
        if (dG) { d0 = dG[(size_t)k * 4 + 0]; d1 = dG[(size_t)k * 4 + 1]; d2 = dG[(size_t)k * 4 + 2]; d3 = dG[(size_t)k * 4 + 3]; }
        if (g.bc == 1) {   // Aw_2D_b1.cpp:47-78: boundary faces zeroed in place (App. C-3)
            if (xe) { g0 = 0.0; d0 = 0.0; }
            if (xw) { g2 = 0.0; d2 = 0.0; }
            if (yn) { g1 = 0.0; d1 = 0.0; }
            if (ys) { g3 = 0.0; d3 = 0.0; }
        }
        if (diff == 1) {
            v[0] = (g0 / fabs(g0) * d0 + d0 + g2 / fabs(g2) * d2 - d2) * hy / 2;
            v[1] = -(g0 / fabs(g0) * d0 - d0) * hy / 2;
            v[2] = -(g2 / fabs(g2) * d2 + d2) * hy / 2;
            v[3] = 0.0; v[4] = 0.0;
        } else if (diff == 2) {
            v[0] = (g1 / fabs(g1) * d1 + d1 + g3 / fabs(g3) * d3 - d3) * hx / 2;
            v[1] = 0.0; v[2] = 0.0;
            v[3] = -(g1 / fabs(g1) * d1 - d1) * hx / 2;
            v[4] = -(g3 / fabs(g3) * d3 + d3) * hx / 2;
        } else {
            v[0] = (fabs(g0) + g0 + fabs(g2) - g2) * hy / 2 + (fabs(g1) + g1 + fabs(g3) - g3) * hx / 2;
            v[1] = -(fabs(g0) - g0) * hy / 2;
            v[2] = -(fabs(g2) + g2) * hy / 2;
            v[3] = -(fabs(g1) - g1) * hx / 2;
            v[4] = -(fabs(g3) + g3) * hx / 2;
        }
    }
    double o[9];
#pragma unroll
    for (int s = 0; s < 9; s++) o[s] = 0.0;
    if (g.bc != 2) {   // neighbours outside the mesh are deleted (bc 1: Acw_2D_b1.cpp:66-79, bc 3: Acw_2D_b3.cpp:62-84)
        if (xe) v[1] = 0.0;
        if (xw) v[2] = 0.0;
        if (yn) v[3] = 0.0;
        if (ys) v[4] = 0.0;
    }
    if (nan_to_zero) {
#pragma unroll
        for (int s = 0; s < 5; s++) if (isnan(v[s])) v[s] = 0.0;
    }
    o[4] = v[0]; o[5] = v[1]; o[3] = v[2]; o[7] = v[3]; o[1] = v[4];
#pragma unroll
    for (int s = 0; s < 9; s++) out[(size_t)s * Ns + k] = o[s];
}

// ---------------------------------------------------------------------------------------------
// K2c: A = f(V, kappa, ah, aw, dt) in the reference's operation order
__global__ void k_combine_A(int Ns, int flavour, double V, double dt, const double *__restrict__ kappa,
                            int kvar, const double *__restrict__ ah, const double *__restrict__ aw,
                            double *__restrict__ A)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const double kap = (flavour <= 2) ? (kvar ? kappa[k] : kappa[0]) : 0.0;
#pragma unroll
    for (int s = 0; s < 9; s++) {
        const double h = ah ? ah[(size_t)s * Ns + k] : 0.0;
        const double w = aw ? aw[(size_t)s * Ns + k] : 0.0;
        double a;
        if (flavour == 0) {
            a = (s == 4) ? V * kap - h : -h;
        } else if (flavour == 1) {
            // Dv + Dv@Dk*dt - Ah*dt + Aw*dt   (advection_diffusion2D.py:104)
            a = (s == 4) ? ((V + (V * kap) * dt) - h * dt) : -(h * dt);
            if (aw) a = a + w * dt;
        } else if (flavour == 2) {
            // Dv + (Dv@Dk - Ah + Aw)*dt        (var_advection_var_diffusion2D.py:103)
            double in = (s == 4) ? (V * kap - h) : -h;
            if (aw) in = in + w;
            a = in * dt;
            if (s == 4) a = V + a;
        } else if (flavour == 3) {
            a = -(h * dt);
        } else if (flavour == 4) {
            a = w * dt;
        } else {
            a = -h;
        }
        A[(size_t)s * Ns + k] = a;
    }
}

// ---------------------------------------------------------------------------------------------
// K3a: out25 = A^T D A, accumulated over the inner (cell) index in ascending order, as SciPy's
// SpGEMM does for `A.T@iDv@Qs@iDv@A` / `A.T@iDv@A`.
__device__ __forceinline__ double qs_of(double V, double iV, double kap) {
    const double As = V * kap;          // Dv@Dk
    return (As * iV) * As;              // As.T@iDv@As   (advection_diffusion2D.py:102-103)
}

__global__ void k_atda(Geo g, const double *__restrict__ A, const double *__restrict__ kappa, int kvar,
                       double V, int mode, double *__restrict__ out)
{
    const int Ns = g.M * g.N;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const int i = k % g.M, j = k / g.M;
    const double iV = 1.0 / V;
    // the (up to) nine cells whose stencil touches k, in ascending cell order
    int cidx[9], cdi[9], cdj[9];
    double left[9];
    int nc = 0;
    for (int dj = -1; dj <= 1; dj++)
        for (int di = -1; di <= 1; di++) {
            int c = g.nbr(i, j, di, dj);
            if (c < 0) continue;
            cidx[nc] = c; cdi[nc] = di; cdj[nc] = dj; nc++;
        }
    if (g.bc == 2) {   // wrapped neighbours are not in ascending order: insertion sort
        for (int a = 1; a < nc; a++) {
            int c = cidx[a], x = cdi[a], y = cdj[a], b = a - 1;
            while (b >= 0 && cidx[b] > c) { cidx[b + 1] = cidx[b]; cdi[b + 1] = cdi[b]; cdj[b + 1] = cdj[b]; b--; }
            cidx[b + 1] = c; cdi[b + 1] = x; cdj[b + 1] = y;
        }
    }
    for (int a = 0; a < nc; a++) {
        const int c = cidx[a];
        const double ack = A[(size_t)gslot(-cdi[a], -cdj[a]) * Ns + c];   // A[c,k]
        double m = ack * iV;                                            // (A^T@iDv)[k,c]
        if (mode == 1) {
            const double q = qs_of(V, iV, kvar ? kappa[c] : kappa[0]);
            m = (m * q) * iV;                                           // (..@Qs)@iDv
        }
        left[a] = m;
    }
    for (int Dj = -2; Dj <= 2; Dj++)
        for (int Di = -2; Di <= 2; Di++) {
            double acc = 0.0;
            for (int a = 0; a < nc; a++) {
                const int ri = Di - cdi[a], rj = Dj - cdj[a];   // offset of the target from cell c
                if (ri < -1 || ri > 1 || rj < -1 || rj > 1) continue;
                acc = acc + left[a] * A[(size_t)gslot(ri, rj) * Ns + cidx[a]];
            }
            out[(size_t)((Dj + 2) * 5 + (Di + 2)) * Ns + k] = acc;
        }
}

// ---------------------------------------------------------------------------------------------
// K3b: block-tridiagonal space-time precision straight into the 43-slot layout
// (advection_diffusion2D.py:112-116).  Pure streaming: 43 coalesced stores per node.
__global__ void k_fill_spacetime(Geo g, const double *__restrict__ AtDA, const double *__restrict__ A,
                                 const double *__restrict__ kappa, int kvar, double V,
                                 const double *__restrict__ Q0, double sigma, double dt, int divide,
                                 double *__restrict__ Q)
{
    const int Ns = g.M * g.N;
    const size_t n = (size_t)Ns * g.T;
    size_t node = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    const int t = (int)(node / Ns), k = (int)(node % Ns);
    const int i = k % g.M, j = k / g.M;
    const double iV = 1.0 / V;
    const double sdt = sigma * dt;
    const double den = dt * sigma;
    const double cmul = 1 / den;
    const double qk = qs_of(V, iV, kvar ? kappa[k] : kappa[0]);
    // lower block: (-A^T@iDv@Qs)[k,k'] = ((-A[k',k])*iV)*Qs[k']
#pragma unroll
    for (int s = 0; s < 9; s++) {
        double v = 0.0;
        if (t > 0) {
            const int di = s % 3 - 1, dj = s / 3 - 1;
            const int kp = g.nbr(i, j, di, dj);
            if (kp >= 0) {
                const double qp = qs_of(V, iV, kvar ? kappa[kp] : kappa[0]);
                v = ((-A[(size_t)(8 - s) * Ns + kp]) * iV) * qp;
                v = divide ? v / den : cmul * v;
            }
        }
        Q[(size_t)s * n + node] = v;
    }
    // diagonal block
#pragma unroll
    for (int q = 0; q < 25; q++) {
        double v;
        if (t == 0) {
            v = sdt * Q0[(size_t)q * Ns + k];                 // sigma*dt*Q0
            if (q == 12) v = v + qk;                            // + Qs
        } else {
            v = AtDA[(size_t)q * Ns + k];
            if (q == 12 && t < g.T - 1) v = v + qk;
        }
        v = divide ? v / den : cmul * v;
        Q[(size_t)(9 + q) * n + node] = v;
    }
    // upper block: (-Qs@iDv@A)[k,k'] = ((-Qs[k])*iV)*A[k,k']
    const double up = (-qk) * iV;
#pragma unroll
    for (int s = 0; s < 9; s++) {
        double v = 0.0;
        if (t < g.T - 1) {
            v = up * A[(size_t)s * Ns + k];
            v = divide ? v / den : cmul * v;
        }
        Q[(size_t)(34 + s) * n + node] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// K3c: separable space-time precision Q = Qt (x) Qs into the 75-slot layout (seperable_spatial_temporal2D.py:82,
// sparse.kron: every entry is the single product Qt[t,t'] * Qs[k,k']).  Qt is the tridiagonal AR(1) precision with
// diagonal (d0, d1, ..., d1, d0) and off-diagonal e (makeQt, :186-211).
__global__ void k_fill_kron(Geo g, const double *__restrict__ Qs, double d0, double d1, double e, double *__restrict__ Q)
{
    const int Ns = g.M * g.N;
    const size_t n = (size_t)Ns * g.T;
    const size_t node = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= n) return;
    const int t = (int)(node / Ns), k = (int)(node % Ns);
    const double dd = (t == 0 || t == g.T - 1) ? d0 : d1;
#pragma unroll 5
    for (int q = 0; q < 25; q++) {
        const double v = Qs[(size_t)q * Ns + k];
        Q[(size_t)q * n + node] = t > 0 ? e * v : 0.0;
        Q[(size_t)(25 + q) * n + node] = dd * v;
        Q[(size_t)(50 + q) * n + node] = t < g.T - 1 ? e * v : 0.0;
    }
}

// ---------------------------------------------------------------------------------------------
// Adjoints of the two face-field stencils (transpose of k_ah_stencil / k_aw_stencil in derivative
// mode).  Given GA9 = d S / d A9 they return d S / d (the eight tensor components the diffusion
// stencil reads) and d S / d (dG of each face), so the gradient with respect to *all* spline
// coefficients is a handful of element-wise products and three small GEMVs with the spline bases
// instead of one stencil evaluation per coefficient.
__global__ void k_ah_adjoint(Geo g, double hx, double hy, const double *__restrict__ GA, double *__restrict__ GH)
{
    const int Ns = g.M * g.N;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const int i = k % g.M, j = k / g.M;
    double ga[9];
#pragma unroll
    for (int s = 0; s < 9; s++) ga[s] = GA[(size_t)s * Ns + k];
    double gw[9];
    gw[0] = ga[4];
#pragma unroll
    for (int s = 1; s < 9; s++) {
        const int di = c_rdi[s], dj = c_rdj[s];
        if (g.bc == 1) {
            const int ii = min(max(i + di, 0), g.M - 1), jj = min(max(j + dj, 0), g.N - 1);
            gw[s] = ga[gslot(ii - i, jj - j)];        // self-target (deleted, folded into rem) reads the centre
        } else if (g.bc == 3) {
            gw[s] = g.nbr(i, j, di, dj) >= 0 ? ga[gslot(di, dj)] : 0.0;
        } else {
            gw[s] = ga[gslot(di, dj)];
        }
    }
    const double rx = hy / hx, ry = hx / hy;
    double W00 = rx * (gw[2] - gw[0]), E00 = rx * (gw[1] - gw[0]);
    double S11 = ry * (gw[4] - gw[0]), N11 = ry * (gw[3] - gw[0]);
    double N01 = 0.25 * (gw[1] - gw[2] + gw[5] - gw[7]);
    double S01 = 0.25 * (-gw[1] + gw[2] + gw[6] - gw[8]);
    double E10 = 0.25 * (gw[3] - gw[4] + gw[5] - gw[8]);
    double W10 = 0.25 * (-gw[3] + gw[4] + gw[6] - gw[7]);
    if (g.bc == 1 && k == 0) { W00 = 0.0; W10 = 0.0; S11 = 0.0; S01 = 0.0; }   // stale-k zeroing (App. C-1)
    GH[(size_t)0 * Ns + k] = W00; GH[(size_t)1 * Ns + k] = E00;
    GH[(size_t)2 * Ns + k] = W10; GH[(size_t)3 * Ns + k] = E10;
    GH[(size_t)4 * Ns + k] = S01; GH[(size_t)5 * Ns + k] = N01;
    GH[(size_t)6 * Ns + k] = S11; GH[(size_t)7 * Ns + k] = N11;
}

// GdG[k][f]: adjoint of the derivative-mode advection stencil (diff=1 for faces E,W; diff=2 for N,S)
__global__ void k_aw_adjoint(Geo g, double hx, double hy, const double *__restrict__ G, const double *__restrict__ GA,
                             double *__restrict__ GdG)
{
    const int Ns = g.M * g.N;
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Ns) return;
    const int i = k % g.M, j = k / g.M;
    const bool xe = (i == g.M - 1), xw = (i == 0), yn = (j == g.N - 1), ys = (j == 0);
    double gf[4];
#pragma unroll
    for (int f = 0; f < 4; f++) gf[f] = G[(size_t)k * 4 + f];
    if (g.bc == 1) { if (xe) gf[0] = 0.0; if (xw) gf[2] = 0.0; if (yn) gf[1] = 0.0; if (ys) gf[3] = 0.0; }
    const double gC = GA[(size_t)4 * Ns + k];
    const double gE = GA[(size_t)5 * Ns + k], gWs = GA[(size_t)3 * Ns + k], gN = GA[(size_t)7 * Ns + k], gS = GA[(size_t)1 * Ns + k];
    const bool cut = g.bc != 2;
    const double s0 = gf[0] / fabs(gf[0]), s2 = gf[2] / fabs(gf[2]), s1 = gf[1] / fabs(gf[1]), s3 = gf[3] / fabs(gf[3]);
    // x faces (diff = 1)
    {
        const bool m0 = !(isnan(s0) || isnan(s2));
        const bool mE = !isnan(s0) && !(cut && xe), mW = !isnan(s2) && !(cut && xw);
        double d0 = 0.0, d2 = 0.0;
        if (m0) { d0 += (s0 + 1.0) * gC; d2 += (s2 - 1.0) * gC; }
        if (mE) d0 -= (s0 - 1.0) * gE;
        if (mW) d2 -= (s2 + 1.0) * gWs;
        GdG[(size_t)k * 4 + 0] = isnan(d0) ? 0.0 : d0 * hy / 2;
        GdG[(size_t)k * 4 + 2] = isnan(d2) ? 0.0 : d2 * hy / 2;
    }
    // y faces (diff = 2)
    {
        const bool m0 = !(isnan(s1) || isnan(s3));
        const bool mN = !isnan(s1) && !(cut && yn), mS = !isnan(s3) && !(cut && ys);
        double d1 = 0.0, d3 = 0.0;
        if (m0) { d1 += (s1 + 1.0) * gC; d3 += (s3 - 1.0) * gC; }
        if (mN) d1 -= (s1 - 1.0) * gN;
        if (mS) d3 -= (s3 + 1.0) * gS;
        GdG[(size_t)k * 4 + 1] = isnan(d1) ? 0.0 : d1 * hx / 2;
        GdG[(size_t)k * 4 + 3] = isnan(d3) ? 0.0 : d3 * hx / 2;
    }
}

}  // namespace spde

using namespace spde;

extern "C" int spde_stencil_adjoint(int M, int N, int bc, double hx, double hy, const double *d_GA9, double *d_GH8,
                                    const double *d_G, double *d_GdG, void *stream)
{
    Geo g{M, N, 1, bc};
    if (d_GH8) {
        k_ah_adjoint<<<cdiv(M * N, 128), 128, 0, (cudaStream_t)stream>>>(g, hx, hy, d_GA9, d_GH8);
        count_launch();
    }
    if (d_G && d_GdG) {
        k_aw_adjoint<<<cdiv(M * N, 128), 128, 0, (cudaStream_t)stream>>>(g, hx, hy, d_G, d_GA9, d_GdG);
        count_launch();
    }
    SPDE_LAUNCH_CHECK();
    return SPDE_OK;
}


extern "C" int spde_ah_stencil(int M, int N, int bc, double hx, double hy, const double *d_H, int face,
                               double *d_ah9, void *stream)
{
    if (M < 2 || N < 2 || bc < 1 || bc > 3) { set_error("spde_ah_stencil: bad mesh/bc"); return SPDE_ERR_ARG; }
    if (bc == 2 && !face) {
        set_error("constant-H periodic stencil is undefined in the reference (AcH_2D_b2.cpp:105)");
        return SPDE_ERR_ARG;
    }
    Geo g{M, N, 1, bc};
    k_ah_stencil<<<cdiv(M * N, 128), 128, 0, (cudaStream_t)stream>>>(g, hx, hy, d_H, face, d_ah9);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_aw_stencil(int M, int N, int bc, double hx, double hy, const double *d_G, const double *d_dG,
                               int face, int diff, int nan_to_zero, double *d_aw9, void *stream)
{
    if (M < 2 || N < 2 || bc < 1 || bc > 3) { set_error("spde_aw_stencil: bad mesh/bc"); return SPDE_ERR_ARG; }
    Geo g{M, N, 1, bc};
    k_aw_stencil<<<cdiv(M * N, 128), 128, 0, (cudaStream_t)stream>>>(g, hx, hy, d_G, d_dG, face, diff, nan_to_zero, d_aw9);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_combine_A(int Ns, int flavour, double V, double dt, const double *d_kappa, int kvar,
                              const double *d_ah9, const double *d_aw9, double *d_A9, void *stream)
{
    if (flavour < 0 || flavour > 5) { set_error("spde_combine_A: bad flavour"); return SPDE_ERR_ARG; }
    k_combine_A<<<cdiv(Ns, 128), 128, 0, (cudaStream_t)stream>>>(Ns, flavour, V, dt, d_kappa, kvar, d_ah9, d_aw9, d_A9);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_atda(int M, int N, int bc, const double *d_A9, const double *d_kappa, int kvar, double V,
                         int mode, double *d_out25, void *stream)
{
    if (bc == 2 && (M < 5 || N < 5)) { set_error("periodic meshes need M,N >= 5"); return SPDE_ERR_ARG; }
    Geo g{M, N, 1, bc};
    k_atda<<<cdiv(M * N, 128), 128, 0, (cudaStream_t)stream>>>(g, d_A9, d_kappa, kvar, V, mode, d_out25);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_fill_kron(int M, int N, int T, int bc, const double *d_Qs25, double d0, double d1, double e,
                              double *d_Q75, void *stream)
{
    if (T < 2) { set_error("spde_fill_kron: T >= 2 required"); return SPDE_ERR_ARG; }
    Geo g = geo_from_abi(M, N, T, bc | (1 << 8));
    const size_t n = (size_t)M * N * T;
    k_fill_kron<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, d_Qs25, d0, d1, e, d_Q75);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}

extern "C" int spde_fill_spacetime(int M, int N, int T, int bc, const double *d_AtDA25, const double *d_A9,
                                   const double *d_kappa, int kvar, double V, const double *d_Q0_25,
                                   double sigma, double dt, int divide, double *d_Q43, void *stream)
{
    if (T < 2) { set_error("spde_fill_spacetime: T >= 2 required"); return SPDE_ERR_ARG; }
    Geo g{M, N, T, bc};
    const int64_t n = (int64_t)M * N * T;
    k_fill_spacetime<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(g, d_AtDA25, d_A9, d_kappa, kvar, V, d_Q0_25,
                                                                    sigma, dt, divide, d_Q43);
    SPDE_LAUNCH_CHECK();
    count_launch();
    return SPDE_OK;
}
