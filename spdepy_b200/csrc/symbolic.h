// symbolic.h -- host-side symbolic analysis, done once per mesh (replaces CHOLMOD's analyse phase
// behind sksparse.cholmod.cholesky, advection_diffusion2D.py:117).
#pragma once
#include <stdint.h>
#include <vector>

#include "common.cuh"

namespace spde {

struct Symbolic {
    Geo geo;
    int n = 0;
    int nslots = 0;                   // 25 (spatial) or 43 (space-time)
    std::vector<int> perm, iperm;     // perm[new] = old, iperm[old] = new  (postordered nested dissection)
    std::vector<int> parent;          // elimination tree on the new ordering
    std::vector<int> colcount;        // |L(:,j)| including the diagonal
    // supernodes
    int nsuper = 0;
    std::vector<int> first;           // nsuper+1, column ranges
    std::vector<int> snode_of;        // n
    std::vector<int64_t> rowptr;      // nsuper+1
    std::vector<int> rows;            // structure below each supernode (new indices, ascending)
    std::vector<int> sparent;         // supernodal tree
    std::vector<int> depth;           // depth in the supernodal tree (roots 0)
    std::vector<int> relidx;          // same shape as rows: position of each row in the parent's front
    int maxdepth = 0;
    double flops = 0;                 // sum_j colcount[j]^2
    int64_t nnzL = 0;                 // sum_j colcount[j]

    // neighbour enumeration of the mesh pattern: slot -> neighbour node or -1
    int slot_nbr(int node, int slot) const;
    void analyse(const Geo &g, int leaf = 0);

  private:
    void nested_dissection(int leaf);
    void etree_postorder();
    void column_counts();
    void supernodes();
    void structures();
};

}  // namespace spde
