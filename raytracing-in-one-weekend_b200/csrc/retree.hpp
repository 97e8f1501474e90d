// retree.hpp — another topology over the reference BVH's LEAVES, for the device walk (host code, no CUDA).
//
// What the reference computes with its tree (FindHitCandidates, SampleBatchJob.cs:403-447): an entity is a candidate iff
// every box on the chain from the root to its leaf passes the slab test (AxisAlignedCuboid.Hit).  BvhNodeData.cs:205-212
// makes an inner node's bounds the exact union (component-wise min / max) of its children's, and the slab test is MONOTONIC
// in the box: with round-to-nearest,
//     mnP <= mnC and mxP >= mxC   =>   fl((mnP - o) * inv) and fl((mxP - o) * inv) bracket C's two products on every axis
// (rounding is monotonic, multiplication by one inv keeps or flips the order of both bounds alike), so P's t_enter <= C's
// and P's t_exit >= C's: a ray that passes a box passes every box that contains it.  The only exception is 0 * inf = NaN
// with a FLAT box (mn == mx == o on an axis whose direction component is 0): min / max drop the NaN and the axis stops
// constraining C while P (mn < o or mx > o there) still sees +-inf.  For boxes with mn < mx on every axis the NaN cases are
// misses of C itself (worked through in DESIGN.md §3.1a).  Hence, for such worlds,
//     the reference's candidate set  ==  the entities of the leaves whose OWN box is hit,
// whatever the inner topology.  The device may therefore walk ANY tree over the same leaves (same leaf boxes, same entity
// ranges) whose inner boxes are unions of what lies below them: every decision the reference makes is reproduced, and
// the number of boxes a ray has to visit becomes a free parameter.  This file builds that tree with the surface-area
// heuristic (full sweep on three axes; binned above kSweepMax leaves), bounded in depth so the walk's stack cannot
// overflow.  The reference's median split (BvhNodeData.cs:166-199) keeps a big entity next to the small ones it overlaps
// until the node is smaller than twice its size; SAH isolates it near the root.
//
// Not applied (returns false, the caller flattens the reference's tree as before) when a leaf box is flat, inverted or not
// finite, when an inner box of the host's tree does not contain its children's boxes (a host may upload anything: then the
// chain matters), or when the tree is malformed (the flattener reports that).
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cstdlib>
#include <limits>
#include <thread>
#include <vector>

#include "rtb.h"

namespace rtb_retree {

struct Box {
  float mn[3], mx[3];
  void grow(const Box& b) {
    for (int k = 0; k < 3; k++) { mn[k] = std::min(mn[k], b.mn[k]); mx[k] = std::max(mx[k], b.mx[k]); }
  }
  double area() const {
    const double dx = (double)mx[0] - mn[0], dy = (double)mx[1] - mn[1], dz = (double)mx[2] - mn[2];
    return dx * dy + dy * dz + dz * dx;
  }
  static Box empty() {
    const float inf = std::numeric_limits<float>::infinity();
    return Box{{inf, inf, inf}, {-inf, -inf, -inf}};
  }
};

struct Leaf {
  Box box;
  float centroid[3];
  int32_t node;        // index in the reference array
  uint32_t weight;     // entities in the leaf
};

constexpr size_t kSweepMax = 8192;   // leaves: above this the split search is binned
constexpr size_t kSweepMaxBigWorld = 256;   // the same for worlds too big for the optimiser (upload time: 1 M triangles 4.1 -> ~1.5 s)
constexpr int kBins = 64;

class Builder {
 public:
  Builder(const rtb_bvh_node* ref, std::vector<Leaf>& leaves, int depth_limit, std::vector<rtb_bvh_node>& out, size_t sweep_max)
      : ref_(ref), leaves_(leaves), depth_limit_(depth_limit), out_(out), sweep_max_(sweep_max) {}

  int32_t build(size_t begin, size_t end, int depth) {
    const int32_t self = (int32_t)out_.size();
    out_.emplace_back();
    const size_t n = end - begin;
    if (n == 1) {
      out_[self] = ref_[leaves_[begin].node];
      out_[self].left = out_[self].right = -1;
      return self;
    }
    size_t mid = split(begin, end, depth);
    int32_t l, r;
    if (n >= kParallelLeaves && threads_ > 1) {
      // big subtrees concurrently, each into its own array, spliced in the serial recursion's order (this node, the left
      // subtree, the right subtree): the same array whatever the thread count.  The two halves of leaves_ are disjoint.
      std::vector<rtb_bvh_node> lout, rout;
      Builder lb(ref_, leaves_, depth_limit_, lout, sweep_max_), rb(ref_, leaves_, depth_limit_, rout, sweep_max_);
      lb.threads_ = threads_ / 2;
      rb.threads_ = threads_ - lb.threads_;
      std::thread left_thread;
      try {
        left_thread = std::thread([&] { lb.build(begin, mid, depth + 1); });
      } catch (...) {                     // no thread to be had: this one does both halves
        lb.threads_ = 1;
        lb.build(begin, mid, depth + 1);
      }
      rb.build(mid, end, depth + 1);
      if (left_thread.joinable()) left_thread.join();
      l = splice(lout);
      r = splice(rout);
    } else {
      l = build(begin, mid, depth + 1);
      r = build(mid, end, depth + 1);
    }
    rtb_bvh_node nd;
    for (int k = 0; k < 3; k++) {
      nd.bounds_min[k] = std::min(out_[l].bounds_min[k], out_[r].bounds_min[k]);
      nd.bounds_max[k] = std::max(out_[l].bounds_max[k], out_[r].bounds_max[k]);
    }
    nd.left = l;
    nd.right = r;
    nd.first_entity = -1;
    nd.entity_count = 0;
    out_[self] = nd;
    return self;
  }

  int threads_ = 1;           // threads this subtree may use

 private:
  static constexpr size_t kParallelLeaves = 32768;
  int32_t splice(const std::vector<rtb_bvh_node>& sub) {
    const int32_t base = (int32_t)out_.size();
    for (rtb_bvh_node nd : sub) {
      if (nd.first_entity < 0) { nd.left += base; nd.right += base; }
      out_.push_back(nd);
    }
    return base;
  }
  static int ceil_log2(size_t n) {
    int b = 0;
    while (((size_t)1 << b) < n) b++;
    return b;
  }
  void sort_axis(size_t begin, size_t end, int axis) {
    std::sort(leaves_.begin() + begin, leaves_.begin() + end, [axis](const Leaf& a, const Leaf& b) {
      if (a.centroid[axis] != b.centroid[axis]) return a.centroid[axis] < b.centroid[axis];
      return a.node < b.node;     // deterministic whatever the sort does with ties
    });
  }
  // the split position in [begin + 1, end - 1]; leaves_[begin, end) is left partitioned accordingly
  size_t split(size_t begin, size_t end, int depth) {
    const size_t n = end - begin;
    Box cb = Box::empty();
    for (size_t i = begin; i < end; i++)
      for (int k = 0; k < 3; k++) { cb.mn[k] = std::min(cb.mn[k], leaves_[i].centroid[k]); cb.mx[k] = std::max(cb.mx[k], leaves_[i].centroid[k]); }
    int widest = 0;
    for (int k = 1; k < 3; k++) if (cb.mx[k] - cb.mn[k] > cb.mx[widest] - cb.mn[widest]) widest = k;
    // out of depth budget (or nothing to tell the leaves apart): halve by count along the widest axis
    if (depth + 1 + ceil_log2(n) >= depth_limit_ || !(cb.mx[widest] > cb.mn[widest])) {
      sort_axis(begin, end, widest);
      return begin + n / 2;
    }
    return n <= sweep_max_ ? sweep_split(begin, end) : binned_split(begin, end, cb, widest);
  }
  size_t sweep_split(size_t begin, size_t end) {
    const size_t n = end - begin;
    double best_cost = std::numeric_limits<double>::infinity();
    int best_axis = 0;
    size_t best_left = n / 2;
    right_area_.resize(n);
    for (int axis = 0; axis < 3; axis++) {
      sort_axis(begin, end, axis);
      Box b = Box::empty();
      for (size_t i = n; i-- > 1;) { b.grow(leaves_[begin + i].box); right_area_[i] = b.area(); }
      b = Box::empty();
      uint64_t wl = 0, wtotal = 0;
      for (size_t i = 0; i < n; i++) wtotal += leaves_[begin + i].weight;
      for (size_t i = 1; i < n; i++) {       // left = [0, i)
        b.grow(leaves_[begin + i - 1].box);
        wl += leaves_[begin + i - 1].weight;
        const double cost = b.area() * (double)wl + right_area_[i] * (double)(wtotal - wl);
        if (cost < best_cost) { best_cost = cost; best_axis = axis; best_left = i; }
      }
    }
    if (best_axis != 2) sort_axis(begin, end, best_axis);
    return begin + best_left;
  }
  size_t binned_split(size_t begin, size_t end, const Box& cb, int widest) {
    const size_t n = end - begin;
    double best_cost = std::numeric_limits<double>::infinity();
    int best_axis = -1, best_bin = 0;
    for (int axis = 0; axis < 3; axis++) {
      const float lo = cb.mn[axis], ext = cb.mx[axis] - cb.mn[axis];
      if (!(ext > 0)) continue;
      Box bins[kBins];
      uint64_t w[kBins] = {};
      for (auto& b : bins) b = Box::empty();
      for (size_t i = begin; i < end; i++) {
        const int k = bin_of(leaves_[i].centroid[axis], lo, ext);
        bins[k].grow(leaves_[i].box);
        w[k] += leaves_[i].weight;
      }
      double ra[kBins];
      uint64_t rw[kBins];
      Box b = Box::empty();
      uint64_t acc = 0;
      for (int k = kBins; k-- > 1;) { b.grow(bins[k]); acc += w[k]; ra[k] = acc ? b.area() : 0.0; rw[k] = acc; }
      b = Box::empty();
      acc = 0;
      for (int k = 1; k < kBins; k++) {      // left = bins [0, k)
        b.grow(bins[k - 1]);
        acc += w[k - 1];
        if (acc == 0 || rw[k] == 0) continue;
        const double cost = b.area() * (double)acc + ra[k] * (double)rw[k];
        if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = k; }
      }
    }
    if (best_axis < 0) {
      sort_axis(begin, end, widest);
      return begin + n / 2;
    }
    const float lo = cb.mn[best_axis], ext = cb.mx[best_axis] - cb.mn[best_axis];
    auto it = std::stable_partition(leaves_.begin() + begin, leaves_.begin() + end,
                                    [&](const Leaf& l) { return bin_of(l.centroid[best_axis], lo, ext) < best_bin; });
    return (size_t)(it - leaves_.begin());
  }
  static int bin_of(float c, float lo, float ext) {
    int k = (int)((c - lo) / ext * (float)kBins);
    return k < 0 ? 0 : k >= kBins ? kBins - 1 : k;
  }

  const rtb_bvh_node* ref_;
  std::vector<Leaf>& leaves_;
  int depth_limit_;
  std::vector<rtb_bvh_node>& out_;
  size_t sweep_max_;
  std::vector<double> right_area_;
};

// Insertion-based optimisation of the built tree (Bittner, Hapala, Havran 2013): a subtree is taken out together with its
// parent and put back where it adds the least surface area to the tree (branch-and-bound search from the root over
// "area its ancestors would grow by + area of the new parent"), largest boxes first, for a few passes.  The sum of the inner
// boxes' areas — the expected number of visits of a ray — only goes down; the leaves are untouched.
class Optimizer {
 public:
  explicit Optimizer(const std::vector<rtb_bvh_node>& in) : n_(in.size()), box_(in.size()), parent_(in.size(), -1), left_(in.size(), -1), right_(in.size(), -1), leaf_(in.size()) {
    for (size_t i = 0; i < n_; i++) {
      for (int k = 0; k < 3; k++) { box_[i].mn[k] = in[i].bounds_min[k]; box_[i].mx[k] = in[i].bounds_max[k]; }
      leaf_[i] = in[i];
      if (in[i].first_entity < 0) {
        left_[i] = in[i].left; right_[i] = in[i].right;
        parent_[in[i].left] = (int32_t)i; parent_[in[i].right] = (int32_t)i;
      }
    }
  }
  double inner_area() const {
    double a = 0;
    for (size_t i = 0; i < n_; i++) if (left_[i] >= 0) a += box_[i].area();
    return a;
  }
  void run(int passes) {
    std::vector<int32_t> order;
    for (int pass = 0; pass < passes; pass++) {
      order.clear();
      for (size_t i = 0; i < n_; i++) if ((int32_t)i != root_ && parent_[i] != root_) order.push_back((int32_t)i);
      std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        const double x = box_[a].area(), y = box_[b].area();
        return x != y ? x > y : a < b;
      });
      const double before = inner_area();
      for (int32_t n : order) {
        if (n == root_ || parent_[n] < 0 || parent_[n] == root_) continue;
        reinsert(n);
      }
      if (!(inner_area() < before * (1.0 - 1e-4))) break;
    }
  }
  int depth() const {
    int deepest = 0;
    std::vector<std::pair<int32_t, int>> st{{root_, 0}};
    while (!st.empty()) {
      auto [i, d] = st.back();
      st.pop_back();
      deepest = std::max(deepest, d);
      if (left_[i] >= 0) { st.push_back({left_[i], d + 1}); st.push_back({right_[i], d + 1}); }
    }
    return deepest;
  }
  // depth-first order, root at 0
  void emit(std::vector<rtb_bvh_node>& out) const {
    out.clear();
    out.reserve(n_);
    emit_node(root_, out);
  }

 private:
  int32_t emit_node(int32_t i, std::vector<rtb_bvh_node>& out) const {
    const int32_t self = (int32_t)out.size();
    out.emplace_back();
    if (left_[i] < 0) {
      out[self] = leaf_[i];
      out[self].left = out[self].right = -1;
      return self;
    }
    const int32_t l = emit_node(left_[i], out);
    const int32_t r = emit_node(right_[i], out);
    rtb_bvh_node nd;
    for (int k = 0; k < 3; k++) { nd.bounds_min[k] = box_[i].mn[k]; nd.bounds_max[k] = box_[i].mx[k]; }
    nd.left = l; nd.right = r; nd.first_entity = -1; nd.entity_count = 0;
    out[self] = nd;
    return self;
  }
  void refit_up(int32_t i) {
    for (; i >= 0; i = parent_[i]) {
      Box b = box_[left_[i]];
      b.grow(box_[right_[i]]);
      box_[i] = b;
    }
  }
  void reinsert(int32_t n) {
    const int32_t p = parent_[n], g = parent_[p];
    const int32_t s = left_[p] == n ? right_[p] : left_[p];
    // take n and its parent out: the sibling moves up
    (left_[g] == p ? left_[g] : right_[g]) = s;
    parent_[s] = g;
    refit_up(g);
    // the position that adds the least area
    const double an = box_[n].area();
    double best = std::numeric_limits<double>::infinity();
    int32_t best_x = s;
    heap_.clear();
    heap_.push_back({0.0, root_});
    while (!heap_.empty()) {
      std::pop_heap(heap_.begin(), heap_.end(), cmp_);
      const Item it = heap_.back();
      heap_.pop_back();
      if (it.induced + an >= best) break;
      Box u = box_[it.node];
      u.grow(box_[n]);
      const double total = it.induced + u.area();
      if (total < best) { best = total; best_x = it.node; }
      if (left_[it.node] >= 0) {
        const double child = total - box_[it.node].area();
        if (child + an < best) {
          heap_.push_back({child, left_[it.node]});
          std::push_heap(heap_.begin(), heap_.end(), cmp_);
          heap_.push_back({child, right_[it.node]});
          std::push_heap(heap_.begin(), heap_.end(), cmp_);
        }
      }
    }
    // p becomes the parent of (best_x, n) where best_x was
    const int32_t x = best_x, gx = parent_[x];
    if (gx < 0) root_ = p; else (left_[gx] == x ? left_[gx] : right_[gx]) = p;
    parent_[p] = gx;
    left_[p] = x; right_[p] = n;
    parent_[x] = p; parent_[n] = p;
    refit_up(p);
  }
  struct Item { double induced; int32_t node; };
  static bool cmp_(const Item& a, const Item& b) { return a.induced != b.induced ? a.induced > b.induced : a.node > b.node; }
  size_t n_;
  std::vector<Box> box_;
  std::vector<int32_t> parent_, left_, right_;
  std::vector<rtb_bvh_node> leaf_;
  std::vector<Item> heap_;
  int32_t root_ = 0;
};

constexpr size_t kOptimizeMaxLeaves = 1u << 16;   // above this the sweep / binned tree is used as built (upload time)
#ifndef RTB_RETREE_PASSES
#define RTB_RETREE_PASSES 8
#endif

// Threads for the re-build of big worlds: RTB_BUILD_THREADS (as the host library's builder), else the hardware's, at most 16.
inline int build_threads() {
  if (const char* e = getenv("RTB_BUILD_THREADS")) return std::max(1, atoi(e));
  const unsigned hw = std::thread::hardware_concurrency();
  return (int)std::min(16u, std::max(1u, hw));
}

// -> true and `out` (root at 0, depth-first order, leaves = the reference's non-empty leaves) when the world qualifies.
// depth_limit: the deepest leaf the device walk's stack allows.
inline bool retree(const rtb_bvh_node* ref, size_t node_count, int depth_limit, std::vector<rtb_bvh_node>& out, int passes = RTB_RETREE_PASSES) {
  out.clear();
  if (!ref || node_count < 3) return false;
  std::vector<uint8_t> seen(node_count, 0);
  std::vector<Leaf> leaves;
  std::vector<int32_t> stack{0};
  while (!stack.empty()) {
    const int32_t n = stack.back();
    stack.pop_back();
    if (n < 0 || (size_t)n >= node_count || seen[n]) return false;     // malformed: the flattener says why
    seen[n] = 1;
    const rtb_bvh_node& nd = ref[n];
    if (nd.first_entity >= 0) {
      if (nd.entity_count <= 0) continue;         // an empty leaf yields no candidate whether its box is hit or not
      Leaf lf;
      for (int k = 0; k < 3; k++) {
        if (!std::isfinite(nd.bounds_min[k]) || !std::isfinite(nd.bounds_max[k]) || !(nd.bounds_min[k] < nd.bounds_max[k])) return false;
        lf.box.mn[k] = nd.bounds_min[k];
        lf.box.mx[k] = nd.bounds_max[k];
        lf.centroid[k] = 0.5f * nd.bounds_min[k] + 0.5f * nd.bounds_max[k];
      }
      lf.node = n;
      lf.weight = (uint32_t)nd.entity_count;
      leaves.push_back(lf);
    } else {
      if (nd.left < 0 || nd.right < 0 || (size_t)nd.left >= node_count || (size_t)nd.right >= node_count) return false;
      for (int32_t c : {nd.left, nd.right}) {
        const rtb_bvh_node& ch = ref[c];
        if (ch.first_entity >= 0 && ch.entity_count <= 0) continue;
        for (int k = 0; k < 3; k++)
          if (!(nd.bounds_min[k] <= ch.bounds_min[k] && nd.bounds_max[k] >= ch.bounds_max[k])) return false;   // the chain matters here
      }
      stack.push_back(nd.right);
      stack.push_back(nd.left);
    }
  }
  if (leaves.size() < 2) return false;
  out.reserve(2 * leaves.size());
  Builder b(ref, leaves, depth_limit, out, leaves.size() <= kOptimizeMaxLeaves ? kSweepMax : kSweepMaxBigWorld);
  b.threads_ = build_threads();
  b.build(0, leaves.size(), 0);
  if (passes > 0 && leaves.size() >= 4 && leaves.size() <= kOptimizeMaxLeaves) {
    Optimizer opt(out);
    opt.run(passes);
    if (opt.depth() <= depth_limit) opt.emit(out);     // a tree the walk's stack cannot hold: keep the one as built
  }
  return true;
}

}  // namespace rtb_retree
