// GPU link-topology solver: binary reduction/broadcast trees over a weighted link matrix.
//
// Parity (capability, own algorithm): src/kvstore/gpu_topology.h — GetP2PWeight :134-199 (link matrix), KernighanLin :269-400
// (balanced bisection that keeps strongly linked devices together), ComputeTrees :1054-1100 (one tree per root), and the tree form
// consumed by comm_tree.h:51-530.  On an NVSwitch box every pair has the same weight and any balanced tree is optimal; the solver matters
// on PCIe / hybrid-cube-mesh hosts and for unit tests that feed it such matrices.
//
// BuildTree(W, n, root): recursive bisection.  A device set S that contains its sub-root r is split into two halves (sizes differ by at
// most one) by Kernighan-Lin refinement maximising the link weight INSIDE the halves; in the half that does not contain r, the device with
// the heaviest link to r becomes that half's sub-root and the child of r; recurse.  The result is a binomial-like tree of depth
// ceil(log2 n) in which every edge was chosen as the heaviest available cross link, returned as parent[] (parent[root] = -1) plus the
// level (distance in rounds from the leaves) at which each edge fires.
#pragma once
#include <algorithm>
#include <cmath>
#include <numeric>
#include <vector>

namespace gx_rt {

struct TopoTree {
  std::vector<int> parent;       // parent[d], -1 for the root
  std::vector<int> round;        // reduction round in which d sends to its parent (0 = first); root: -1
  int depth = 0;
};

class TopologySolver {
 public:
  TopologySolver(const std::vector<float>& W, int n) : W_(W), n_(n) {}

  TopoTree BuildTree(int root) const {
    TopoTree t;
    t.parent.assign(n_, -1);
    t.round.assign(n_, -1);
    std::vector<int> all(n_);
    std::iota(all.begin(), all.end(), 0);
    int levels = 0;
    for (int m = 1; m < n_; m <<= 1) ++levels;
    t.depth = levels;
    Split(all, root, levels, &t);
    return t;
  }

  // Kernighan-Lin bisection of `set` (|A| = ceil(n/2) contains `pin`): returns the two halves
  void Bisect(const std::vector<int>& set, int pin, std::vector<int>* A, std::vector<int>* B) const {
    const int m = static_cast<int>(set.size());
    const int na = (m + 1) / 2;
    // greedy seed: grow A from `pin` by strongest attachment
    std::vector<char> inA(n_, 0), in(n_, 0);
    for (int d : set) in[d] = 1;
    A->assign(1, pin); inA[pin] = 1;
    while (static_cast<int>(A->size()) < na) {
      int best = -1; float bw = -1.f;
      for (int d : set) {
        if (inA[d]) continue;
        float w = 0.f;
        for (int a : *A) w += Wt(a, d);
        if (w > bw) { bw = w; best = d; }
      }
      A->push_back(best); inA[best] = 1;
    }
    B->clear();
    for (int d : set) if (!inA[d]) B->push_back(d);
    // KL passes: swap the pair with the best gain in cut weight while it is positive (pin stays in A)
    for (int pass = 0; pass < 8; ++pass) {
      float best_gain = 1e-6f; int ia = -1, ib = -1;
      for (size_t i = 0; i < A->size(); ++i) {
        if ((*A)[i] == pin) continue;
        for (size_t j = 0; j < B->size(); ++j) {
          const int a = (*A)[i], b = (*B)[j];
          // D(x) = external - internal;  gain = D(a) + D(b) - 2 w(a,b)
          float Da = 0.f, Db = 0.f;
          for (int x : *B) Da += Wt(a, x);
          for (int x : *A) if (x != a) Da -= Wt(a, x);
          for (int x : *A) Db += Wt(b, x);
          for (int x : *B) if (x != b) Db -= Wt(b, x);
          const float gain = Da + Db - 2.f * Wt(a, b);
          if (gain > best_gain) { best_gain = gain; ia = static_cast<int>(i); ib = static_cast<int>(j); }
        }
      }
      if (ia < 0) break;
      std::swap((*A)[ia], (*B)[ib]);
    }
  }

 private:
  float Wt(int a, int b) const { return W_[static_cast<size_t>(a) * n_ + b]; }

  void Split(const std::vector<int>& set, int r, int level, TopoTree* t) const {
    if (set.size() <= 1) return;
    std::vector<int> A, B;
    Bisect(set, r, &A, &B);
    if (B.empty()) return;
    int sub = B[0]; float bw = -1.f;
    for (int d : B) if (Wt(r, d) > bw) { bw = Wt(r, d); sub = d; }
    t->parent[sub] = r;
    t->round[sub] = level - 1;        // the top-level edge fires last in the reduction
    Split(A, r, level - 1, t);
    Split(B, sub, level - 1, t);
  }

  const std::vector<float>& W_;
  int n_;
};

}  // namespace gx_rt
