// common.cuh -- shared helpers of the sm_100a kernels (error plumbing, mesh geometry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/spde_b200.h"

namespace spde {

void set_error(const std::string &msg);
void count_launch(int n = 1);   // bookkeeping for spde_launch_count (bench.py's gpu_launches)

#define SPDE_CUDA_CHECK(expr)                                                              \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            spde::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));           \
            return _e == cudaErrorMemoryAllocation ? SPDE_ERR_OOM : SPDE_ERR_CUDA;         \
        }                                                                                  \
    } while (0)

#define SPDE_LAUNCH_CHECK()                                                                \
    do {                                                                                   \
        cudaError_t _e = cudaGetLastError();                                               \
        if (_e != cudaSuccess) {                                                           \
            spde::set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));      \
            return SPDE_ERR_CUDA;                                                          \
        }                                                                                  \
    } while (0)

// Mesh geometry shared by host and device code.  Cell k = j*M + i (x fastest),
// spat2Dtemp_regular_mesh.py:82-92.
struct Geo {
    int M, N, T, bc;
    int pat = 0;     // space-time pattern: 0 = 3x3 | 5x5 | 3x3 (advection-diffusion, 43 slots), 1 = 5x5 | 5x5 | 5x5 (Kronecker
                     // Qt (x) Qs of the separable model, 75 slots, slot = (dt+1)*25 + (dj+2)*5 + (di+2))
    __host__ __device__ int Ns() const { return M * N; }
    // neighbour of cell (i,j) at offset (di,dj); returns -1 when it lies outside the mesh
    // (bc 1 and 3) or the wrapped cell (bc 2, AH_2D_b2.cpp:32-41).
    __host__ __device__ int nbr(int i, int j, int di, int dj) const {
        int ii = i + di, jj = j + dj;
        if (bc == 2) {
            ii = ii < 0 ? ii + M : (ii >= M ? ii - M : ii);
            jj = jj < 0 ? jj + N : (jj >= N ? jj - N : jj);
        } else if (ii < 0 || ii >= M || jj < 0 || jj >= N) {
            return -1;
        }
        return jj * M + ii;
    }
    __host__ __device__ int nslots() const { return T == 1 ? 25 : (pat == 1 ? 75 : 43); }
    // offsets of a precision slot (Q25 / Q43 layout, see spde_b200.h)
    __host__ __device__ void slot_offset(int slot, int &dt, int &dj, int &di) const {
        dt = 0;
        if (T == 1) { dj = slot / 5 - 2; di = slot % 5 - 2; }
        else if (pat == 1) { dt = slot / 25 - 1; const int q = slot % 25; dj = q / 5 - 2; di = q % 5 - 2; }
        else if (slot < 9) { dt = -1; dj = slot / 3 - 1; di = slot % 3 - 1; }
        else if (slot < 34) { const int q = slot - 9; dj = q / 5 - 2; di = q % 5 - 2; }
        else { const int q = slot - 34; dt = 1; dj = q / 3 - 1; di = q % 3 - 1; }
    }
    // node reached from `node` through `slot`, or -1
    __host__ __device__ int slot_nbr(int node, int slot) const {
        const int ns = M * N;
        const int t = node / ns, k = node - t * ns;
        int dt, dj, di;
        slot_offset(slot, dt, dj, di);
        const int tt = t + dt;
        if (tt < 0 || tt >= T) return -1;
        const int c = nbr(k % M, k / M, di, dj);
        return c < 0 ? -1 : tt * ns + c;
    }
};

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// C ABI convention: the boundary-condition argument carries the pattern in bits 8..15 (SPDE_PATTERN_KRON = 1 << 8)
static inline Geo geo_from_abi(int M, int N, int T, int bc)
{
    Geo g{M, N, T, bc & 0xff};
    g.pat = (bc >> 8) & 0xff;
    return g;
}

}  // namespace spde
