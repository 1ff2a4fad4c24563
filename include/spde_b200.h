/* spde_b200.h -- C ABI of the B200-native SPDE precision-and-likelihood hot path.
 *
 * One shared library, `spdepy_b200/csrc/libspde_b200.so`, built by nvcc for sm_100a.  Plain C
 * types only: ints, doubles, raw device/host pointers and a `void*` CUDA stream.  Every entry
 * point returns an int status (SPDE_OK or an SPDE_ERR_* code; `spde_last_error()` gives text).
 * All `d_*` pointers are DEVICE pointers owned by the caller; the library owns only the plan and
 * its internal workspaces.  Calls are stream-ordered on the stream passed in (NULL = default).
 *
 * What each group replaces in the reference (berild/spdepy; paths relative to /root/reference):
 *
 *   stencils      the ctypes FFI `AH_new/AH_Row/AH_Col/AH_Val/AH_delete` and `Aw_*`
 *                 (src/spdepy/spdes/ccode/AcH_2D_b1.cpp:170-185, AH_2D_b1.cpp:154-169,
 *                 Acw_2D_b1.cpp:117-132, Aw_2D_b1.cpp:136-151) and the Python wrappers `Ah()/Aw()`
 *                 (src/spdepy/spdes/advection_diffusion2D.py:226-260)
 *   assembly      the scipy.sparse SpGEMM / bmat block stacking of `makeQ`
 *                 (advection_diffusion2D.py:101-116, whittle_matern_anisotropic2D.py:69-73)
 *   plan/factor   `sksparse.cholmod.cholesky(Q)` (advection_diffusion2D.py:117,193; model.py:79,125)
 *   solves        `Factor.solve_A / solve_Lt / apply_Pt / logdet`
 *                 (advection_diffusion2D.py:194-202, model.py:80,126)
 *   selinv        the R/INLA `inla.run -m qinv` subprocess (src/spdepy/rqinv.R:18-70, model.py:89-118)
 *   likelihood    the NumPy reductions of `logLike` (advection_diffusion2D.py:198-208)
 *
 * Slot layouts (all slot-major, i.e. `a[slot*Ns + cell]`, so a warp reads/writes 256 contiguous
 * bytes per slot):
 *   A9   operator stencil, slot = (dj+1)*3 + (di+1), dj,di in {-1,0,1}: value of row `cell` at
 *        column `cell + dj*M + di` (wrapped for bc=2).  Slots whose neighbour is outside the
 *        mesh hold 0; Neumann duplicates (AcH_2D_b1.cpp:71-118) are already folded.
 *   Q25  spatial precision row, slot = (dj+2)*5 + (di+2), dj,di in {-2..2}.
 *   Q43  space-time precision row of node (cell,t): slots 0..8 = 3x3 block to t-1, 9..33 = 5x5
 *        block in t, 34..42 = 3x3 block to t+1; address `q[slot*n + t*Ns + cell]`.
 */
#ifndef SPDE_B200_H
#define SPDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPDE_OK 0
#define SPDE_ERR_NOT_SPD 1   /* a pivot was <= 0 or NaN; see spde_factor_info */
#define SPDE_ERR_OOM 2
#define SPDE_ERR_ARG 3
#define SPDE_ERR_CUDA 4

typedef struct spde_plan spde_plan;

int spde_abi_version(void);
const char *spde_last_error(void);
/* number of kernels this library has launched so far in this process (optionally reset) */
long long spde_launch_count(int reset);

/* ------------------------------------------------------------------ stencils (K2) */

/* 9-point finite-volume stencil of div(H grad).  `face`=0: d_H holds one 2x2 tensor (4 doubles,
 * row-major) [AcH_2D_b{1,3}.cpp]; `face`=1: d_H holds H[cell][4 faces W,E,S,N][2][2]
 * [AH_2D_b{1,2,3}.cpp].  Output d_ah9 in the A9 layout.  bc: 1 Neumann, 2 periodic, 3 Dirichlet.
 * bc=2 with face=0 is rejected: the reference returns NaN there (AcH_2D_b2.cpp:105). */
int spde_ah_stencil(int M, int N, int bc, double hx, double hy, const double *d_H, int face,
                    double *d_ah9, void *stream);

/* 5-point upwind advection stencil.  `face`=0: d_G = (wx, wy) [Acw_2D_b*.cpp]; `face`=1:
 * d_G / d_dG hold face-normal velocities [cell][E,N,W,S] [Aw_2D_b*.cpp] (d_dG may be NULL when
 * diff==3).  diff: 1 = d/dwx, 2 = d/dwy, 3 = value.  `nan_to_zero` applies the filter of
 * var_advection_var_diffusion2D.py:255.  Output d_aw9 in the A9 layout (corner slots 0). */
int spde_aw_stencil(int M, int N, int bc, double hx, double hy, const double *d_G, const double *d_dG,
                    int face, int diff, int nan_to_zero, double *d_aw9, void *stream);

/* A = f(V, kappa, ah, aw, dt) in the reference's operation order.
 * flavour 0: V*kappa - ah                         (whittle_matern2D.py:69)
 * flavour 1: V + (V*kappa)*dt - ah*dt + aw*dt     (advection_diffusion2D.py:104)
 * flavour 2: V + ((V*kappa - ah) + aw)*dt         (var_advection_var_diffusion2D.py:103)
 * flavour 3: -(ah*dt)   flavour 4: aw*dt   flavour 5: -ah   (derivative directions)
 * kvar=0: d_kappa points at one double; kvar=1: one per cell. d_aw9 may be NULL. */
int spde_combine_A(int Ns, int flavour, double V, double dt, const double *d_kappa, int kvar,
                   const double *d_ah9, const double *d_aw9, double *d_A9, void *stream);

/* ------------------------------------------------------------------ assembly (K3) */

/* out25 = A^T D A on the 5x5 pattern, D = 1/V (mode 0, whittle_matern2D.py:70) or
 * D = kappa^2/V evaluated as ((a*iV)*Qs)*iV with Qs = ((V k)*iV)*(V k) (mode 1,
 * advection_diffusion2D.py:103,114).  Inner index accumulated in ascending cell order. */
int spde_atda(int M, int N, int bc, const double *d_A9, const double *d_kappa, int kvar, double V,
              int mode, double *d_out25, void *stream);

/* Block-tridiagonal space-time precision (advection_diffusion2D.py:112-116) into the Q43 layout.
 * d_Q0_25: initial-field precision (Q25 layout); divide=0: (1/(dt*sigma))*x, 1: x/(dt*sigma). */
int spde_fill_spacetime(int M, int N, int T, int bc, const double *d_AtDA25, const double *d_A9,
                        const double *d_kappa, int kvar, double V, const double *d_Q0_25,
                        double sigma, double dt, int divide, double *d_Q43, void *stream);

/* Separable space-time precision Q = Qt (x) Qs (seperable_spatial_temporal2D.py:82) into the Q75 layout: slot =
 * (dt+1)*25 + (dj+2)*5 + (di+2), address q[slot*n + t*Ns + cell].  Qt = tridiag with diagonal (d0, d1, ..., d1, d0) and
 * off-diagonal e (makeQt, :186-211).  Every function that takes `bc` selects this 75-slot pattern when
 * SPDE_PATTERN_KRON is or-ed into it (spde_plan_create, spde_q_apply, spde_sddmm); T >= 2. */
#define SPDE_PATTERN_KRON (1 << 8)
int spde_fill_kron(int M, int N, int T, int bc, const double *d_Qs25, double d0, double d1, double e,
                   double *d_Q75, void *stream);
/* Adjoint of spde_fill_kron with respect to Qs: Wd25[q][cell] = sum_t sum_dt Qt[t,t+dt] * W75[(dt+1)*25+q][(cell,t)]. */
int spde_kron_reduce(int M, int N, int T, int bc, const double *d_W75, double d0, double d1, double e,
                     double *d_Wd25, void *stream);

/* ------------------------------------------------------------------ symbolic plan (host, once per mesh) */

/* Geometric nested dissection + elimination tree + supernodes + level schedule + kernel task
 * lists for the pattern of an M x N (x T) mesh (T=1: Q25 pattern, T>1: Q43 pattern).
 * Replaces CHOLMOD's analyse phase; the permutation is exposed by spde_plan_perm. */
int spde_plan_create(int M, int N, int T, int bc, int max_rhs, spde_plan **out);
void spde_plan_destroy(spde_plan *p);
/* info ids: 0 n, 1 nsuper, 2 nnz(L) stored, 3 flops sum cc^2 (as double bits via spde_plan_info_d),
 * 4 factor bytes, 5 arena bytes, 6 nlevels, 7 max front, 8 n launches (factor) */
int64_t spde_plan_info(const spde_plan *p, int what);
double spde_plan_info_d(const spde_plan *p, int what);
int spde_plan_perm(const spde_plan *p, int32_t *h_perm /* n, new -> old */);
/* host copies of the supernodal structure, for tests / the CPU restatement in oracle/ */
int spde_plan_supernodes(const spde_plan *p, int32_t *h_first /* nsuper+1 */, int64_t *h_rowptr /* nsuper+1 */,
                         int32_t *h_rows /* rowptr[nsuper] */, int32_t *h_parent /* nsuper */);

/* Raw host copy of a schedule (prog 0 factor, 1 forward solve, 2 backward solve, 3 selected inverse;
 * what 0 launches, 1 GEMM tasks, 2 tiles, 3 POTRF, 4 extend-add, 5 gather, 6 W^T W tasks) or of the
 * storage layout (prog 4; what 0 sizes, 1 scatter map, 2 candidate slots, 3 diagonal positions,
 * 4 rows|relative indices, 5 extraction entries, 6 extraction depth pointers, 7 column counts).
 * Pass h_out=NULL to query the count and element size.  Test / inspection interface. */
int spde_plan_export(spde_plan *p, int prog, int k, int what, void *h_out, int64_t *count, int *elem_size);

/* Per-launch device timing of the schedules (CUDA events around every launch; serialises the stream,
 * so never on inside a timed benchmark region).  h_out (may be NULL): 8x16 accumulated milliseconds
 * indexed [launch kind][GEMM variant] followed by 8x16 launch counts.  Kinds: 0 GEMM, 1 POTRF,
 * 2 extend-add, 3 memset, 4 gather, 5 W^T W, 6 extract.  GEMM variant = cfg*4 + a_kmaj*2 + b_kmaj. */
int spde_plan_profile(spde_plan *p, int enable, double *h_out, int reset);

/* ------------------------------------------------------------------ numeric factorisation (K4,K5,K6) */

/* L L^T = P (Q + tau*diag(d_cnt)) P^T.  d_Q in Q25/Q43 layout (only the lower-triangle slots of
 * the original ordering are read, as CHOLMOD does); d_cnt (n doubles, may be NULL) is the diagonal
 * of S^T S (advection_diffusion2D.py:192).  `which` selects one of two factor stores (0: Q, 1: Q_c). */
int spde_factorize(spde_plan *p, int which, const double *d_Q, const double *d_cnt, double tau, void *stream);
/* Asynchronous form: spde_factorize_async enqueues the schedule (on a private stream per store once its CUDA
 * graph exists) and returns; spde_factor_wait orders it before `stream`, synchronises and returns the status.
 * Issuing both stores back to back lets the two factorisations of logLike overlap. */
int spde_factorize_async(spde_plan *p, int which, const double *d_Q, const double *d_cnt, double tau, void *stream);
int spde_factor_wait(spde_plan *p, int which, void *stream);
int spde_factor_info(spde_plan *p, int which, int *h_status, int *h_bad_column);
int spde_logdet(spde_plan *p, int which, double *h_logdet, void *stream);
/* The same value written to a DEVICE address, no copy and no synchronise (the scalars of one logLike evaluation are
 * collected in one device vector and read back once; replaces the per-term NumPy scalars of
 * advection_diffusion2D.py:198 and optim/__init__.py:44). */
int spde_logdet_dev(spde_plan *p, int which, double *d_logdet, void *stream);

/* ------------------------------------------------------------------ solves (K7) */
/* X is n x k, ROW-major on the device (node-major: the k values of a node are contiguous); solved
 * in place.  mode is a bit set: 1 forward substitution with L, 2 back substitution with L^T,
 * 4 apply P to the input (rows taken in the original node ordering), 8 apply P^T to the output.
 * solve_A = 15, solve_Lt = 2, solve_L = 1, apply_Pt(solve_Lt(z)) = 10 (model.py:80). */
int spde_solve(spde_plan *p, int which, int mode, double *d_X, int k, void *stream);

/* ------------------------------------------------------------------ selected inverse (K10) */
/* Takahashi recursion on the supernodal structure; writes Z = (L L^T)^-1 restricted to the
 * pattern of Q into d_Zq (same Q25/Q43 layout, all stored slots filled symmetrically). */
int spde_selinv(spde_plan *p, int which, double *d_Zq, void *stream);
/* split form (start both stores, then fetch both) so the two Takahashi passes overlap */
int spde_selinv_start(spde_plan *p, int which, void *stream);
int spde_selinv_fetch(spde_plan *p, int which, double *d_Zq, void *stream);

/* ------------------------------------------------------------------ streamed evaluation (meshes beyond HBM) */
/* Depth-first, statically planned evaluation for meshes whose factor, update matrices and inverse fronts do not
 * fit in HBM together (256x256x100: 260 GB of L).  Replaces the same reference calls as spde_factorize /
 * spde_logdet / spde_solve / spde_selinv (sksparse.cholmod.cholesky + Factor.logdet / solve_A,
 * advection_diffusion2D.py:193-202) in ONE pass over the supernodal tree: supernodes whose subtree holds more than
 * `top_bytes` of factor are processed front by front and their panels parked in pinned host memory between the
 * forward (factorise, log-determinant, forward substitution) and the backward (back substitution, Takahashi
 * selected inverse) pass; the subtrees below are factorised again in the backward pass instead of being stored.
 * All device memory is one pool whose size is known when the plan is made (spde_ooc_info).
 * The panel of a front-by-front supernode travels in slices of one outer block of columns on a copy stream of the
 * evaluator's own: to the host while the factorisation of the same front goes on, and back, last slice first, while
 * the Takahashi recursion already works on the slices behind it (environment SPDE_OOC_OVERLAP=0 when the schedules
 * are built: whole-panel copies on the caller's stream instead).
 * want_backward = 0: forward pass only (log-determinant, L^-1 P b), smaller pool, no host memory.
 * build = 0: memory plan only (sizes through spde_ooc_info), no schedules -- for choosing top_bytes. */
typedef struct spde_ooc spde_ooc;
int spde_ooc_create(spde_plan *p, int64_t top_bytes, int want_backward, int build, spde_ooc **out);
void spde_ooc_destroy(spde_ooc *o);
/* info ids: 0 segments, 1 top segments, 2 pool bytes, 3 forward peak bytes, 4 backward peak bytes, 5 pinned host
 * bytes, 6 scatter entries, 7 largest forward working set bytes, 8 / 9 launches of all factor / selected-inverse
 * schedules.  info_d ids: 0 flops factorised twice, 1 / 2 device milliseconds of the last forward / backward pass. */
int64_t spde_ooc_info(const spde_ooc *o, int what);
double spde_ooc_info_d(const spde_ooc *o, int what);
/* d_X: k right-hand sides (n x k row-major) solved in place with the mode bits of spde_solve, k = 0: none;
 * d_Zq: selected inverse on the pattern of Q, or NULL; h_logdet: log det (Q + tau diag(cnt)). */
int spde_ooc_run(spde_ooc *o, const double *d_Q, const double *d_cnt, double tau, double *d_X, int k, int mode,
                 double *d_Zq, double *h_logdet, void *stream);
/* host export of the per-segment schedules and tables (tests, oracle/plan_emulator.py); seg >= 0, what = 7: the
 * slices of an overlapped segment as (pool offset, doubles, host offset) triples, inverse diagonal blocks last */
int spde_ooc_export(spde_ooc *o, int seg, int prog, int k, int what, void *h_out, int64_t *count, int *elem_size);

/* ------------------------------------------------------------------ likelihood / gradient reductions (K8,K9,K11) */
/* y = Q x for k right-hand sides (x, y row-major n x k): stencil apply, no index arrays. */
int spde_q_apply(int M, int N, int T, int bc, const double *d_Q, const double *d_X, int k, double *d_Y, void *stream);
/* h_out[0] = sum(X .* Y) over n*k entries (deterministic two-stage reduction). */
int spde_dot(const double *d_X, const double *d_Y, int64_t len, double *h_out, void *stream);
/* h_out[0] = sum_node w[node] * sum_p X[node,p]*Y[node,p] (X, Y row-major n x k). */
int spde_wdot(const double *d_X, const double *d_Y, const double *d_w, int64_t n, int k, double *h_out, void *stream);
/* Device-output forms of the three reductions (d_out[0] on the device, no copy, no synchronise): the likelihood and
 * gradient scalars of one evaluation stay on the device until one read-back at the end (advection_diffusion2D.py:198-223). */
int spde_dot_dev(const double *d_X, const double *d_Y, int64_t len, double *d_out, void *stream);
int spde_wdot_dev(const double *d_X, const double *d_Y, const double *d_w, int64_t n, int k, double *d_out, void *stream);
int spde_residual_ss_dev(const double *d_data, const double *d_mu, const int64_t *d_obs, int64_t nobs, int r,
                         double *d_out, void *stream);
/* h_out[0] = sum_{i,p} (data[i,p] - mu[obs[i],p])^2; data nobs x r, mu n x r, both row-major
 * (advection_diffusion2D.py:198). */
int spde_residual_ss(const double *d_data, const double *d_mu, const int64_t *d_obs, int64_t nobs, int r,
                     double *h_out, void *stream);
/* b[obs[i],:] += tau * data[i,:]   (b = S^T data * tau, advection_diffusion2D.py:194; b zeroed by the caller). */
int spde_scatter_obs(const double *d_data, const int64_t *d_obs, int64_t nobs, int r, double tau, double *d_b, void *stream);
/* d_Qdiag[node] += tau * cnt[node]: Q + tau S^T S on the diagonal slot (model.py:120-124). */
int spde_add_diag(double *d_Qdiag, const double *d_cnt, double tau, int64_t n, void *stream);
/* W[slot,node] (+)= alpha * sum_p X[node,p] * Y[nbr(node,slot),p]  (sampled dense-dense product on the
 * pattern of Q; the Hutchinson weights of advection_diffusion2D.py:204-206). */
int spde_sddmm(int M, int N, int T, int bc, const double *d_X, const double *d_Y, int k, double alpha,
               int accumulate, double *d_W, void *stream);
/* Adjoint of the assembly (transpose of spde_atda / spde_fill_spacetime): given weights W on the
 * pattern of Q, d sum(W .* Q) / d A9 (d_GA9, A9 layout), / d Qs per cell (d_Gq, Ns; Qs = kappa^2 V,
 * advection_diffusion2D.py:103) and / d Q0 (d_GQ0_25).  timed=0: W, Q in the Q25 layout and only
 * d_GA9 is produced.  timed=2: W holds weights on B = A^T (Qs/V^2) A alone (Q25 layout; the diagonal
 * blocks of the space-time prior, whose determinant factorises over time: logdet Q = logdet Q0 +
 * (T-1)(Ns log(1/(dt sigma)) + logdet B)); d_GA9 and d_Gq are produced, d_work: 19*Ns doubles.
 * d_work: 44*Ns doubles of scratch (timed=1).  With these, the trace
 * sum(W .* dQ_i) of every parameter i (advection_diffusion2D.py:119-182, 204-206) is a dot
 * product with that parameter's stencil direction dA_i instead of an n x n sparse matrix. */
int spde_assembly_adjoint(int M, int N, int T, int bc, const double *d_W, const double *d_A9,
                          const double *d_kappa, int kvar, double V, double sigma, double dt, int timed,
                          double *d_work, double *d_GA9, double *d_Gq, double *d_GQ0_25, void *stream);
/* Adjoints of the face-field stencils (transposes of spde_ah_stencil with face=1 and of spde_aw_stencil
 * with face=1 in derivative mode): from d_GA9 = dS/dA9 to d_GH8[8][Ns] = dS/d(W00,E00,W10,E10,S01,N01,S11,N11),
 * the eight tensor components AH_2D_b*.cpp reads, and d_GdG[Ns][4] = dS/d(dG of faces E,N,W,S) for the
 * velocity field d_G (NaN-to-zero filter and boundary rules of Aw_2D_b*.cpp included).  Either output may be NULL. */
int spde_stencil_adjoint(int M, int N, int bc, double hx, double hy, const double *d_GA9, double *d_GH8,
                         const double *d_G, double *d_GdG, void *stream);
/* d_out[c] = sum_r B[r,c] * u[r], B row-major rows x cols (spline-basis chain rule, evalB/evalBH). */
int spde_gemv_t(const double *d_B, const double *d_u, int rows, int cols, double *d_out, void *stream);

/* One dense task through the grouped FP64 tensor-core GEMM (unit test + roofline probe).
 * cfg 0: 128x128 tiles, 1: 128x64, 2: 64x64; a_kmaj/b_kmaj: operand stored with K contiguous;
 * flags: GF_* bits of csrc/gemm.cuh (lower 1<<9, beta0 1<<10, neg 1<<11).  *h_ms = mean device time
 * of `reps` launches after one warm-up launch. */
int spde_gemm_single(int cfg, int a_kmaj, int b_kmaj, int flags, int M, int N, int K,
                     const double *d_A, int lda, const double *d_B, int ldb, double *d_C, int ldc,
                     int reps, float *h_ms, void *stream);

/* Latency probe of the 64x64 POTRF + inverse kernel that sits on the critical path of every front: `ntasks`
 * synthetic SPD blocks of order b, `reps` launches.  *h_us = mean microseconds per launch; h_clk[5] = SM clock
 * of CTA 0 at entry / after the load / after the column sweep / after the block inversion / at exit. */
int spde_potrf_bench(int b, int ld, int ntasks, int reps, float *h_us, long long *h_clk, void *stream);

#ifdef __cplusplus
}
#endif
#endif
