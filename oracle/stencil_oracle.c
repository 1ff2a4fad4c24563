/* oracle/stencil_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by spdepy_b200/).
 *
 * CPU restatement of the reference's twelve finite-volume stencil generators
 * (src/spdepy/spdes/ccode/{AcH,AH,Acw,Aw}_2D_b{1,2,3}.cpp), boundary condition passed as an
 * argument instead of compiled in.  Every generator emits cell-major COO triplets in the
 * reference's fixed slot order and marks a deleted slot with row == nx*ny, exactly as the
 * reference does (the Python caller filters `row != M*N`, advection_diffusion2D.py:236-239).
 * The arithmetic is written in the same operation order as the reference so that the result is
 * bit-identical (checked against oracle/_ref/lib_*.so in tests/test_oracle_stencils.py).
 *
 * Slot order, diffusion (9): C,E,W,N,S,NE,SW,NW,SE        (AcH_2D_b1.cpp:121-129)
 * Slot order, advection (5): C,E,W,N,S                    (Acw_2D_b1.cpp:82-86)
 * Face order of H: 0 W, 1 E, 2 S, 3 N  (AH_2D_b3.cpp:36-44); of G: 0 E, 1 N, 2 W, 3 S
 * (Aw_2D_b3.cpp:50-54).
 *
 * Quirks reproduced on purpose:
 *   C-1  AH_2D_b1.cpp:34-52  the Neumann zeroing of the face tensors uses the *previous* cell's
 *        index, so only cell 0 ever sees zeroed faces.
 *   C-2  AcH_2D_b2.cpp:105   constant H + periodic is an ABI mismatch in the reference (NaN);
 *        bc==2 is rejected here for the constant-H generator (returns 1).
 *   C-3  Aw_2D_b1.cpp:47-78  Neumann boundary faces get G=dG=0 and the derivative modes then
 *        evaluate 0/|0| = NaN; the NaN is kept here (the Python wrapper zeroes it afterwards,
 *        var_advection_var_diffusion2D.py:255).
 *   C-4  Acw_2D_b1.cpp:63-80 the boundary correction of the diagonal is applied in every diff mode.
 */
#include <math.h>
#include <string.h>

static void nbr(int i, int j, int nx, int ny, int bc, int *in, int *ip, int *jn, int *jp)
{
    *in = i - 1; *ip = i + 1; *jn = j - 1; *jp = j + 1;
    if (bc == 1) {               /* clamp: AcH_2D_b1.cpp:51-68 */
        if (i == 0) *in = i; else if (i == nx - 1) *ip = i;
        if (j == 0) *jn = j; else if (j == ny - 1) *jp = j;
    } else if (bc == 2) {        /* wrap: AH_2D_b2.cpp:32-41 */
        if (i == 0) *in = nx - 1; else if (i == nx - 1) *ip = 0;
        if (j == 0) *jn = ny - 1; else if (j == ny - 1) *jp = 0;
    }
}

static void cols9(int *col, int k, int i, int j, int nx, int in, int ip, int jn, int jp)
{
    col[0] = k;
    col[1] = j * nx + ip;  col[2] = j * nx + in;
    col[3] = jp * nx + i;  col[4] = jn * nx + i;
    col[5] = jp * nx + ip; col[6] = jn * nx + in;
    col[7] = jp * nx + in; col[8] = jn * nx + ip;
}

/* Dirichlet deletion pattern shared by the b3 generators (AcH_2D_b3.cpp:56-73). */
static void dirichlet_rows9(int *row, int k, int i, int j, int nx, int ny)
{
    int s, del = nx * ny;
    for (s = 0; s < 9; s++) row[s] = k;
    if (i == 0) { row[2] = del; row[6] = del; row[7] = del; }
    else if (i == nx - 1) { row[1] = del; row[5] = del; row[8] = del; }
    if (j == 0) { row[4] = del; row[6] = del; row[8] = del; }
    else if (j == ny - 1) { row[3] = del; row[5] = del; row[7] = del; }
}

/* Constant 2x2 H (row-major H[4] = H00,H01,H10,H11).  AcH_2D_b1.cpp:19-144, AcH_2D_b3.cpp:19-88. */
int orc_ah_const(int nx, int ny, const double *H, double hx, double hy, int bc,
                 int *row, int *col, double *val)
{
    int i, j, idx = 0, del = nx * ny;
    if (bc == 2) return 1;
    for (j = 0; j < ny; j++) for (i = 0; i < nx; i++, idx += 9) {
        int in, ip, jn, jp, k = j * nx + i;
        nbr(i, j, nx, ny, bc, &in, &ip, &jn, &jp);
        cols9(col + idx, k, i, j, nx, in, ip, jn, jp);
        if (bc == 3) {
            double hxy = H[1] + H[2];
            val[idx] = -2.0 * hy / hx * H[0] - 2.0 * hx / hy * H[3] + 0.0;
            val[idx + 1] = hy / hx * H[0];  val[idx + 2] = hy / hx * H[0];
            val[idx + 3] = hx / hy * H[3];  val[idx + 4] = hx / hy * H[3];
            val[idx + 5] = 1.0 / 4.0 * hxy; val[idx + 6] = 1.0 / 4.0 * hxy;
            val[idx + 7] = -1.0 / 4.0 * hxy; val[idx + 8] = -1.0 / 4.0 * hxy;
            dirichlet_rows9(row + idx, k, i, j, nx, ny);
        } else {
            /* per-face copies; note the reference reads H[1][0] for every cross term */
            double W00 = H[0], E00 = H[0], W10 = H[2], E10 = H[2];
            double S01 = H[2], N01 = H[2], S11 = H[3], N11 = H[3];
            double w[9], rem;
            if (i == 0) { W00 = 0.0; W10 = 0.0; } else if (i == nx - 1) { E00 = 0.0; E10 = 0.0; }
            if (j == 0) { S11 = 0.0; S01 = 0.0; } else if (j == ny - 1) { N11 = 0.0; N01 = 0.0; }
            w[1] = hy / hx * E00 + 1.0 / 4.0 * (N01 - S01);
            w[2] = hy / hx * W00 - 1.0 / 4.0 * (N01 - S01);
            w[3] = hx / hy * N11 + 1.0 / 4.0 * (E10 - W10);
            w[4] = hx / hy * S11 - 1.0 / 4.0 * (E10 - W10);
            w[5] = 1.0 / 4.0 * (N01 + E10);
            w[6] = 1.0 / 4.0 * (S01 + W10);
            w[7] = -1.0 / 4.0 * (N01 + W10);
            w[8] = -1.0 / 4.0 * (S01 + E10);
            /* rem: literal transcription of the sign conventions of AcH_2D_b1.cpp:71-118 */
            rem = 0.0;
            {
                int s;
                const int *c = col + idx;
                int *r = row + idx;
                for (s = 1; s < 9; s++) r[s] = k;
                if (c[1] == k) { rem = rem + hy / hx * E00 + 1.0 / 4.0 * (N01 - S01); r[1] = del; }
                if (c[2] == k) { rem = rem + hy / hx * W00 - 1.0 / 4.0 * (N01 - S01); r[2] = del; }
                if (c[3] == k) { rem = rem + hx / hy * N11 + 1.0 / 4.0 * (E10 - W10); r[3] = del; }
                if (c[4] == k) { rem = rem + hx / hy * S11 - 1.0 / 4.0 * (E10 - W10); r[4] = del; }
                if (c[5] == k) { rem = rem + 1.0 / 4.0 * (N01 + E10); r[5] = del; }
                if (c[6] == k) { rem = rem + 1.0 / 4.0 * (S01 + W10); r[6] = del; }
                if (c[7] == k) { rem = rem - 1.0 / 4.0 * (N01 + W10); r[7] = del; }
                if (c[8] == k) { rem = rem - 1.0 / 4.0 * (S01 + E10); r[8] = del; }
                r[0] = k;
            }
            val[idx] = -hy / hx * (E00 + W00) - hx / hy * (N11 + S11) + rem;
            memcpy(val + idx + 1, w + 1, 8 * sizeof(double));
        }
    }
    return 0;
}

/* Per-cell, per-face H[k][f][a][b] (Ns*16 doubles).  AH_2D_b{1,2,3}.cpp:19-128.
 * The b1 generator mutates its input; a private copy is taken so the caller's array is intact. */
int orc_ah_face(int nx, int ny, const double *Hin, double hx, double hy, int bc,
                int *row, int *col, double *val, double *scratch /* Ns*16 */)
{
    int i, j, idx = 0, del = nx * ny, k = 0;
    double *H = scratch;
    memcpy(H, Hin, (size_t)nx * ny * 16 * sizeof(double));
#define HF(c, f, a, b) H[(size_t)(c) * 16 + (f) * 4 + (a) * 2 + (b)]
    for (j = 0; j < ny; j++) for (i = 0; i < nx; i++, idx += 9) {
        int in, ip, jn, jp, s;
        nbr(i, j, nx, ny, bc, &in, &ip, &jn, &jp);
        if (bc == 1) {           /* stale k: still the previous cell here (quirk C-1) */
            if (i == 0) { HF(k, 0, 0, 0) = 0.0; HF(k, 0, 1, 0) = 0.0; }
            else if (i == nx - 1) { HF(k, 1, 0, 0) = 0.0; HF(k, 1, 1, 0) = 0.0; }
            if (j == 0) { HF(k, 2, 1, 1) = 0.0; HF(k, 2, 0, 1) = 0.0; }
            else if (j == ny - 1) { HF(k, 3, 1, 1) = 0.0; HF(k, 3, 0, 1) = 0.0; }
        }
        k = j * nx + i;
        cols9(col + idx, k, i, j, nx, in, ip, jn, jp);
        {
            double W00 = HF(k, 0, 0, 0), E00 = HF(k, 1, 0, 0), W10 = HF(k, 0, 1, 0), E10 = HF(k, 1, 1, 0);
            double S01 = HF(k, 2, 0, 1), N01 = HF(k, 3, 0, 1), S11 = HF(k, 2, 1, 1), N11 = HF(k, 3, 1, 1);
            double rem = 0.0;
            int *r = row + idx;
            const int *c = col + idx;
            for (s = 0; s < 9; s++) r[s] = k;
            if (bc == 1) {
                if (c[1] == k) { rem = rem + hy / hx * E00 + 1.0 / 4.0 * (N01 - S01); r[1] = del; }
                if (c[2] == k) { rem = rem + hy / hx * W00 - 1.0 / 4.0 * (N01 - S01); r[2] = del; }
                if (c[3] == k) { rem = rem + hx / hy * N11 + 1.0 / 4.0 * (E10 - W10); r[3] = del; }
                if (c[4] == k) { rem = rem + hx / hy * S11 - 1.0 / 4.0 * (E10 - W10); r[4] = del; }
                if (c[5] == k) { rem = rem + 1.0 / 4.0 * (N01 + E10); r[5] = del; }
                if (c[6] == k) { rem = rem + 1.0 / 4.0 * (S01 + W10); r[6] = del; }
                if (c[7] == k) { rem = rem - 1.0 / 4.0 * (N01 + W10); r[7] = del; }
                if (c[8] == k) { rem = rem - 1.0 / 4.0 * (S01 + E10); r[8] = del; }
            } else if (bc == 3) {
                dirichlet_rows9(r, k, i, j, nx, ny);
            }
            val[idx] = -hy / hx * (E00 + W00) - hx / hy * (N11 + S11) + rem;
            val[idx + 1] = hy / hx * E00 + 1.0 / 4.0 * (N01 - S01);
            val[idx + 2] = hy / hx * W00 - 1.0 / 4.0 * (N01 - S01);
            val[idx + 3] = hx / hy * N11 + 1.0 / 4.0 * (E10 - W10);
            val[idx + 4] = hx / hy * S11 - 1.0 / 4.0 * (E10 - W10);
            val[idx + 5] = 1.0 / 4.0 * (N01 + E10);
            val[idx + 6] = 1.0 / 4.0 * (S01 + W10);
            val[idx + 7] = -1.0 / 4.0 * (N01 + W10);
            val[idx + 8] = -1.0 / 4.0 * (S01 + E10);
        }
    }
#undef HF
    return 0;
}

/* Constant velocity G = (wx, wy).  Acw_2D_b{1,2,3}.cpp. diff: 1 = d/dwx, 2 = d/dwy, else value. */
int orc_aw_const(int nx, int ny, const double *G, double hx, double hy, int diff, int bc,
                 int *row, int *col, double *val)
{
    int i, j, idx = 0, del = nx * ny, s;
    for (j = 0; j < ny; j++) for (i = 0; i < nx; i++, idx += 5) {
        int in, ip, jn, jp, k = j * nx + i;
        nbr(i, j, nx, ny, bc, &in, &ip, &jn, &jp);
        if (diff == 1) {
            val[idx] = G[0] / fabs(G[0]) * hy;
            val[idx + 1] = -(G[0] / fabs(G[0]) - 1.0) * hy / 2;
            val[idx + 2] = -(G[0] / fabs(G[0]) + 1.0) * hy / 2;
            val[idx + 3] = 0.0; val[idx + 4] = 0.0;
        } else if (diff == 2) {
            val[idx] = G[1] / fabs(G[1]) * hx;
            val[idx + 1] = 0.0; val[idx + 2] = 0.0;
            val[idx + 3] = -(G[1] / fabs(G[1]) - 1.0) * hx / 2;
            val[idx + 4] = -(G[1] / fabs(G[1]) + 1.0) * hx / 2;
        } else {
            val[idx] = fabs(G[0]) * hy + fabs(G[1]) * hx;
            val[idx + 1] = -(fabs(G[0]) - G[0]) * hy / 2;
            val[idx + 2] = -(fabs(G[0]) + G[0]) * hy / 2;
            val[idx + 3] = -(fabs(G[1]) - G[1]) * hx / 2;
            val[idx + 4] = -(fabs(G[1]) + G[1]) * hx / 2;
        }
        for (s = 0; s < 5; s++) row[idx + s] = k;
        if (bc == 1) {           /* quirk C-4: in every diff mode */
            if (i == 0) { val[idx] -= fabs(G[0]) * hy / 2; row[idx + 2] = del; }
            else if (i == nx - 1) { val[idx] -= fabs(G[0]) * hy / 2; row[idx + 1] = del; }
            if (j == 0) { val[idx] -= fabs(G[1]) * hx / 2; row[idx + 4] = del; }
            else if (j == ny - 1) { val[idx] -= fabs(G[1]) * hx / 2; row[idx + 3] = del; }
        } else if (bc == 3) {
            if (i == nx - 1) row[idx + 1] = del;
            if (i == 0) row[idx + 2] = del;
            if (j == ny - 1) row[idx + 3] = del;
            if (j == 0) row[idx + 4] = del;
        }
        col[idx] = k;
        col[idx + 1] = j * nx + ip; col[idx + 2] = j * nx + in;
        col[idx + 3] = jp * nx + i; col[idx + 4] = jn * nx + i;
    }
    return 0;
}

/* Per-cell face-normal velocities G[k][4], dG[k][4] (faces E,N,W,S).  Aw_2D_b{1,2,3}.cpp.
 * b1 zeroes boundary-face G and dG in place (quirk C-3): done on private copies. */
int orc_aw_face(int nx, int ny, const double *Gin, const double *dGin, double hx, double hy,
                int diff, int bc, int *row, int *col, double *val, double *scratch /* Ns*8 */)
{
    int i, j, idx = 0, del = nx * ny, s;
    size_t ns = (size_t)nx * ny;
    double *G = scratch, *dG = scratch + ns * 4;
    memcpy(G, Gin, ns * 4 * sizeof(double));
    memcpy(dG, dGin, ns * 4 * sizeof(double));
    for (j = 0; j < ny; j++) for (i = 0; i < nx; i++, idx += 5) {
        int in, ip, jn, jp, k = j * nx + i;
        double *g = G + (size_t)k * 4, *d = dG + (size_t)k * 4;
        nbr(i, j, nx, ny, bc, &in, &ip, &jn, &jp);
        col[idx] = k;
        col[idx + 1] = j * nx + ip; col[idx + 2] = j * nx + in;
        col[idx + 3] = jp * nx + i; col[idx + 4] = jn * nx + i;
        for (s = 0; s < 5; s++) row[idx + s] = k;
        if (bc == 1) {
            if (col[idx + 1] == k) { g[0] = 0.0; d[0] = 0.0; row[idx + 1] = del; }
            if (col[idx + 2] == k) { g[2] = 0.0; d[2] = 0.0; row[idx + 2] = del; }
            if (col[idx + 3] == k) { g[1] = 0.0; d[1] = 0.0; row[idx + 3] = del; }
            if (col[idx + 4] == k) { g[3] = 0.0; d[3] = 0.0; row[idx + 4] = del; }
        } else if (bc == 3) {
            if (i == nx - 1) row[idx + 1] = del;
            if (i == 0) row[idx + 2] = del;
            if (j == ny - 1) row[idx + 3] = del;
            if (j == 0) row[idx + 4] = del;
        }
        if (diff == 1) {
            val[idx] = (g[0] / fabs(g[0]) * d[0] + d[0] + g[2] / fabs(g[2]) * d[2] - d[2]) * hy / 2;
            val[idx + 1] = -(g[0] / fabs(g[0]) * d[0] - d[0]) * hy / 2;
            val[idx + 2] = -(g[2] / fabs(g[2]) * d[2] + d[2]) * hy / 2;
            val[idx + 3] = 0.0; val[idx + 4] = 0.0;
        } else if (diff == 2) {
            val[idx] = (g[1] / fabs(g[1]) * d[1] + d[1] + g[3] / fabs(g[3]) * d[3] - d[3]) * hx / 2;
            val[idx + 1] = 0.0; val[idx + 2] = 0.0;
            val[idx + 3] = -(g[1] / fabs(g[1]) * d[1] - d[1]) * hx / 2;
            val[idx + 4] = -(g[3] / fabs(g[3]) * d[3] + d[3]) * hx / 2;
        } else {
            val[idx] = (fabs(g[0]) + g[0] + fabs(g[2]) - g[2]) * hy / 2
                     + (fabs(g[1]) + g[1] + fabs(g[3]) - g[3]) * hx / 2;
            val[idx + 1] = -(fabs(g[0]) - g[0]) * hy / 2;
            val[idx + 2] = -(fabs(g[2]) + g[2]) * hy / 2;
            val[idx + 3] = -(fabs(g[1]) - g[1]) * hx / 2;
            val[idx + 4] = -(fabs(g[3]) + g[3]) * hx / 2;
        }
    }
    return 0;
}
