// malio_flatten.hpp — header-only host-side flattener: live ikd-Tree -> malio_map_node snapshot.
//
// Host C++ (the reference's language).  Duck-typed on the reference's node type
//   KD_TREE<PointType>::KD_TREE_NODE            MA_LIO/include/ikd-Tree/ikd_Tree.h:59-82
// so it can be instantiated inside a MA-LIO checkout without this repo depending on PCL.
//
// What it reproduces (all read-only; the tree is not modified):
//  * Push_Down (ikd_Tree.cpp:1371-1466): a pending need_push_down_to_{left,right} on a node rewrites the
//    child's {tree,point}_{downsample_}deleted flags; here the rewritten ("effective") flags are carried
//    down the DFS instead of being stored, which is what KD_TREE::Search would see because it calls
//    Push_Down(root) before it looks at either child (ikd_Tree.cpp:1082-1095).
//  * Search's pruning of deleted subtrees (`root == nullptr || root->tree_deleted`, :1075): they are
//    not exported, so the snapshot is compact.
//  * node_range_* of both children, copied verbatim (calc_box_dist reads them, :1702-1720).
//
// Layout: DFS pre-order, left child of node i at i+1, right child index in link (malio_b200.h).
// Concurrency: take the snapshot while no ikd-Tree rebuild is swapping a subtree, i.e. under the same
// protocol Nearest_Search uses (ikd_Tree.cpp:431-450); see INTEGRATION.md.
#ifndef MALIO_FLATTEN_HPP_
#define MALIO_FLATTEN_HPP_

#include <cstdint>
#include <vector>

#include "malio_b200.h"

namespace malio {

struct FlattenResult {
  uint32_t n_nodes = 0;      // exported (live) nodes
  uint32_t n_points = 0;     // exported nodes whose own point is not deleted
  uint32_t max_depth = 0;    // root has depth 1
  bool overflow = false;     // capacity exceeded (nothing past capacity is written)
};

namespace detail {
struct EffFlags {
  bool tree_deleted, point_deleted, tree_ds_deleted, point_ds_deleted, need_l, need_r;
};

template <class Node>
inline EffFlags stored_flags(const Node* n) {
  return EffFlags{n->tree_deleted, n->point_deleted, n->tree_downsample_deleted,
                  n->point_downsample_deleted, n->need_push_down_to_left, n->need_push_down_to_right};
}

// flags of `child` as Push_Down(parent) would leave them (ikd_Tree.cpp:1383-1395 / 1425-1437)
template <class Node>
inline EffFlags pushed_flags(const EffFlags& parent, bool parent_needs_push, const Node* child) {
  EffFlags c = stored_flags(child);
  if (!parent_needs_push) return c;
  c.tree_ds_deleted = c.tree_ds_deleted | parent.tree_ds_deleted;
  c.point_ds_deleted = c.point_ds_deleted | parent.tree_ds_deleted;
  c.tree_deleted = parent.tree_deleted || c.tree_ds_deleted;
  c.point_deleted = c.tree_deleted || c.point_ds_deleted;
  c.need_l = true;
  c.need_r = true;
  return c;
}

template <class Node>
inline void copy_box(const Node* n, float* box) {
  box[0] = n->node_range_x[0]; box[1] = n->node_range_x[1];
  box[2] = n->node_range_y[0]; box[3] = n->node_range_y[1];
  box[4] = n->node_range_z[0]; box[5] = n->node_range_z[1];
}
}  // namespace detail

// PointFn(const Node*, uint32_t slot): called once per exported node, e.g. to record normal_y / an id.
template <class Node, class PointFn>
FlattenResult flatten_ikdtree(const Node* root, malio_map_node* out, uint32_t capacity, PointFn&& on_node) {
  using detail::EffFlags;
  FlattenResult res;
  if (root == nullptr) return res;
  struct Frame {
    const Node* node;
    EffFlags f;
    uint32_t depth;
    int64_t parent_slot;   // slot whose `link` must receive this node's index (right children only)
  };
  std::vector<Frame> stack;
  stack.reserve(128);
  EffFlags rf = detail::stored_flags(root);
  if (rf.tree_deleted) return res;
  stack.push_back(Frame{root, rf, 1u, -1});
  while (!stack.empty()) {
    Frame fr = stack.back();
    stack.pop_back();
    const uint32_t slot = res.n_nodes;
    if (slot >= capacity || slot > MALIO_LINK_INDEX_MASK) {
      res.overflow = true;
      return res;
    }
    res.n_nodes++;
    if (fr.depth > res.max_depth) res.max_depth = fr.depth;
    if (fr.parent_slot >= 0) out[fr.parent_slot].link |= (slot & MALIO_LINK_INDEX_MASK);

    const Node* n = fr.node;
    malio_map_node& o = out[slot];
    o.x = n->point.x;
    o.y = n->point.y;
    o.z = n->point.z;
    uint32_t link = 0;
    if (fr.f.point_deleted) link |= MALIO_LINK_POINT_DELETED; else res.n_points++;
    for (int k = 0; k < 6; ++k) { o.lbox[k] = 0.0f; o.rbox[k] = 0.0f; }

    const Node* l = n->left_son_ptr;
    const Node* r = n->right_son_ptr;
    EffFlags lf{}, rfl{};
    bool has_l = false, has_r = false;
    if (l != nullptr) {
      lf = detail::pushed_flags(fr.f, fr.f.need_l, l);
      has_l = !lf.tree_deleted;
    }
    if (r != nullptr) {
      rfl = detail::pushed_flags(fr.f, fr.f.need_r, r);
      has_r = !rfl.tree_deleted;
    }
    if (has_l) { link |= MALIO_LINK_HAS_LEFT; detail::copy_box(l, o.lbox); }
    if (has_r) { link |= MALIO_LINK_HAS_RIGHT; detail::copy_box(r, o.rbox); }
    o.link = link;
    on_node(n, slot);
    // pre-order: left must be emitted next, so push right first
    if (has_r) stack.push_back(Frame{r, rfl, fr.depth + 1, (int64_t)slot});
    if (has_l) stack.push_back(Frame{l, lf, fr.depth + 1, -1});
  }
  return res;
}

// Compact form for malio_upload_map_compact(): the same DFS, 16 bytes per node (point + link); the children's boxes
// are not read (the device rebuilds them as the tight boxes of the live points, which is what Update() maintains).
// root_box_out (6 floats, optional) receives the root's node_range_* = the bounding box of the map.
template <class Node, class PointFn>
FlattenResult flatten_ikdtree_compact(const Node* root, malio_map_point* out, uint32_t capacity, PointFn&& on_node,
                                      float* root_box_out = nullptr) {
  using detail::EffFlags;
  FlattenResult res;
  if (root == nullptr) return res;
  struct Frame { const Node* node; EffFlags f; uint32_t depth; int64_t parent_slot; };
  std::vector<Frame> stack;
  stack.reserve(128);
  EffFlags rf = detail::stored_flags(root);
  if (rf.tree_deleted) return res;
  if (root_box_out) detail::copy_box(root, root_box_out);
  stack.push_back(Frame{root, rf, 1u, -1});
  while (!stack.empty()) {
    Frame fr = stack.back();
    stack.pop_back();
    const uint32_t slot = res.n_nodes;
    if (slot >= capacity || slot > MALIO_LINK_INDEX_MASK) { res.overflow = true; return res; }
    res.n_nodes++;
    if (fr.depth > res.max_depth) res.max_depth = fr.depth;
    if (fr.parent_slot >= 0) out[fr.parent_slot].link |= (slot & MALIO_LINK_INDEX_MASK);
    const Node* n = fr.node;
    malio_map_point& o = out[slot];
    o.x = n->point.x; o.y = n->point.y; o.z = n->point.z;
    uint32_t link = 0;
    if (fr.f.point_deleted) link |= MALIO_LINK_POINT_DELETED; else res.n_points++;
    const Node* l = n->left_son_ptr;
    const Node* r = n->right_son_ptr;
    EffFlags lf{}, rfl{};
    bool has_l = false, has_r = false;
    if (l != nullptr) { lf = detail::pushed_flags(fr.f, fr.f.need_l, l); has_l = !lf.tree_deleted; }
    if (r != nullptr) { rfl = detail::pushed_flags(fr.f, fr.f.need_r, r); has_r = !rfl.tree_deleted; }
    if (has_l) link |= MALIO_LINK_HAS_LEFT;
    if (has_r) link |= MALIO_LINK_HAS_RIGHT;
    o.link = link;
    on_node(n, slot);
    if (has_r) stack.push_back(Frame{r, rfl, fr.depth + 1, (int64_t)slot});
    if (has_l) stack.push_back(Frame{l, lf, fr.depth + 1, -1});
  }
  return res;
}

// ---------------------------------------------------------------------------------------------------------------
// Parallel flattening (OpenMP).  The serial DFS above is one dependent pointer chase over ~176-byte nodes: ~50-80 ms for
// a 1M-node tree on one core — far more than everything the GPU does per scan.  DFS pre-order slots are computable
// without walking in order once the exported size of every subtree is known (left child = slot + 1, right child =
// slot + 1 + size(left)), so: (1) walk the top of the tree serially down to subtrees of <= `grain` nodes (KD_TREE_NODE::
// TreeSize bounds them), (2) count the exported nodes of those subtrees in parallel, (3) size and place the top part,
// (4) flatten the subtrees in parallel, each into its own slot range.  The result is identical, byte for byte, to the
// serial functions (tests/test_oracle_cpu.py).  Compile the including translation unit with -fopenmp; without it the
// pragmas are ignored and the work runs serially.
namespace detail {
struct RecFull {
  using type = malio_map_node;
  template <class Node>
  static void boxes(type& o, const Node* l, bool has_l, const Node* r, bool has_r) {
    for (int k = 0; k < 6; ++k) { o.lbox[k] = 0.0f; o.rbox[k] = 0.0f; }
    if (has_l) copy_box(l, o.lbox);
    if (has_r) copy_box(r, o.rbox);
  }
};
struct RecCompact {
  using type = malio_map_point;
  template <class Node>
  static void boxes(type&, const Node*, bool, const Node*, bool) {}
};

template <class Node>
inline void child_flags(const Node* n, const EffFlags& f, const Node*& l, const Node*& r, EffFlags& lf, EffFlags& rf, bool& has_l,
                        bool& has_r) {
  l = n->left_son_ptr; r = n->right_son_ptr;
  has_l = false; has_r = false;
  if (l != nullptr) { lf = pushed_flags(f, f.need_l, l); has_l = !lf.tree_deleted; }
  if (r != nullptr) { rf = pushed_flags(f, f.need_r, r); has_r = !rf.tree_deleted; }
}

// exported nodes of the subtree of n (n itself is exported)
template <class Node>
inline uint32_t count_exported(const Node* root, const EffFlags& rf) {
  struct Fr { const Node* n; EffFlags f; };
  std::vector<Fr> st;
  st.reserve(64);
  st.push_back(Fr{root, rf});
  uint32_t cnt = 0;
  while (!st.empty()) {
    Fr fr = st.back();
    st.pop_back();
    ++cnt;
    const Node *l, *r;
    EffFlags lf{}, rfl{};
    bool hl, hr;
    child_flags(fr.n, fr.f, l, r, lf, rfl, hl, hr);
    if (hr) st.push_back(Fr{r, rfl});
    if (hl) st.push_back(Fr{l, lf});
  }
  return cnt;
}

// serial pre-order flatten of one subtree into out[base ...] (absolute slots), depth of its root = depth0
template <class Rec, class Node, class PointFn>
inline void flatten_subtree(const Node* root, const EffFlags& rf, uint32_t base, uint32_t depth0, typename Rec::type* out,
                            PointFn& on_node, uint32_t& n_points, uint32_t& max_depth) {
  struct Fr { const Node* n; EffFlags f; uint32_t depth; int64_t parent_slot; };
  std::vector<Fr> st;
  st.reserve(128);
  st.push_back(Fr{root, rf, depth0, -1});
  uint32_t slot = base;
  while (!st.empty()) {
    Fr fr = st.back();
    st.pop_back();
    if (fr.depth > max_depth) max_depth = fr.depth;
    if (fr.parent_slot >= 0) out[fr.parent_slot].link |= (slot & MALIO_LINK_INDEX_MASK);
    typename Rec::type& o = out[slot];
    o.x = fr.n->point.x; o.y = fr.n->point.y; o.z = fr.n->point.z;
    uint32_t link = 0;
    if (fr.f.point_deleted) link |= MALIO_LINK_POINT_DELETED; else n_points++;
    const Node *l, *r;
    EffFlags lf{}, rfl{};
    bool hl, hr;
    child_flags(fr.n, fr.f, l, r, lf, rfl, hl, hr);
    if (hl) link |= MALIO_LINK_HAS_LEFT;
    if (hr) link |= MALIO_LINK_HAS_RIGHT;
    Rec::boxes(o, l, hl, r, hr);
    o.link = link;
    on_node(fr.n, slot);
    if (hr) st.push_back(Fr{r, rfl, fr.depth + 1, (int64_t)slot});
    if (hl) st.push_back(Fr{l, lf, fr.depth + 1, -1});
    ++slot;
  }
}

template <class Rec, class Node, class PointFn>
FlattenResult flatten_parallel(const Node* root, typename Rec::type* out, uint32_t capacity, PointFn& on_node, float* root_box_out,
                               uint32_t grain) {
  FlattenResult res;
  if (root == nullptr) return res;
  const EffFlags rf0 = stored_flags(root);
  if (rf0.tree_deleted) return res;
  if (root_box_out) copy_box(root, root_box_out);
  if (grain < 1024) grain = 1024;
  // (1) top part: nodes whose TreeSize exceeds the grain, in pre-order; their small children become work items
  struct Top { const Node* n; EffFlags f; uint32_t depth; int left, right;   // index into tops (>= 0) or -(item + 1), 0 = none
               uint32_t size, slot; };
  struct Item { const Node* n; EffFlags f; uint32_t depth; uint32_t size, slot, n_points, max_depth; };
  std::vector<Top> tops;
  std::vector<Item> items;
  if ((uint32_t)root->TreeSize <= grain) {
    items.push_back(Item{root, rf0, 1u, 0, 0, 0, 0});
  } else {
    struct Fr { const Node* n; EffFlags f; uint32_t depth; int parent; bool is_left; };
    std::vector<Fr> st;
    st.push_back(Fr{root, rf0, 1u, -1, false});
    while (!st.empty()) {
      Fr fr = st.back();
      st.pop_back();
      int ref;
      if ((uint32_t)fr.n->TreeSize <= grain) { items.push_back(Item{fr.n, fr.f, fr.depth, 0, 0, 0, 0}); ref = -(int)items.size(); }
      else {
        tops.push_back(Top{fr.n, fr.f, fr.depth, 0, 0, 0, 0});
        ref = (int)tops.size() - 1;
        const Node *l, *r;
        EffFlags lf{}, rfl{};
        bool hl, hr;
        child_flags(fr.n, fr.f, l, r, lf, rfl, hl, hr);
        if (hr) st.push_back(Fr{r, rfl, fr.depth + 1, ref, false});
        if (hl) st.push_back(Fr{l, lf, fr.depth + 1, ref, true});
      }
      if (fr.parent >= 0) {
        const int enc = ref >= 0 ? ref + 1 : ref;      // tops: index + 1 (> 0), items: -(index + 1) (< 0), 0 = no child
        if (fr.is_left) tops[fr.parent].left = enc; else tops[fr.parent].right = enc;
      }
    }
  }
  // (2) exported size of every work item
  const int64_t n_items = (int64_t)items.size();
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_items; ++i) items[i].size = count_exported(items[i].n, items[i].f);
  // (3) sizes of the top nodes (children come later in `tops`: reverse order is bottom-up), then slots top-down
  auto size_of = [&](int enc) -> uint32_t { return enc == 0 ? 0u : (enc > 0 ? tops[enc - 1].size : items[-enc - 1].size); };
  for (int64_t k = (int64_t)tops.size() - 1; k >= 0; --k) tops[k].size = 1u + size_of(tops[k].left) + size_of(tops[k].right);
  const uint64_t total = tops.empty() ? (uint64_t)items[0].size : (uint64_t)tops[0].size;
  if (total > capacity || total > (uint64_t)MALIO_LINK_INDEX_MASK + 1) { res.overflow = true; return res; }
  auto place = [&](int enc, uint32_t slot) { if (enc > 0) tops[enc - 1].slot = slot; else if (enc < 0) items[-enc - 1].slot = slot; };
  if (tops.empty()) items[0].slot = 0; else tops[0].slot = 0;
  for (size_t k = 0; k < tops.size(); ++k) {   // parents precede their children in `tops`
    place(tops[k].left, tops[k].slot + 1);
    place(tops[k].right, tops[k].slot + 1 + size_of(tops[k].left));
  }
  // top records
  for (size_t k = 0; k < tops.size(); ++k) {
    const Top& t = tops[k];
    typename Rec::type& o = out[t.slot];
    o.x = t.n->point.x; o.y = t.n->point.y; o.z = t.n->point.z;
    uint32_t link = 0;
    if (t.f.point_deleted) link |= MALIO_LINK_POINT_DELETED; else res.n_points++;
    const Node *l, *r;
    EffFlags lf{}, rfl{};
    bool hl, hr;
    child_flags(t.n, t.f, l, r, lf, rfl, hl, hr);
    if (hl) link |= MALIO_LINK_HAS_LEFT;
    if (hr) link |= MALIO_LINK_HAS_RIGHT | ((t.slot + 1 + size_of(t.left)) & MALIO_LINK_INDEX_MASK);
    Rec::boxes(o, l, hl, r, hr);
    o.link = link;
    on_node(t.n, t.slot);
    if (t.depth > res.max_depth) res.max_depth = t.depth;
  }
  // (4) the subtrees, each into its own slot range
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t i = 0; i < n_items; ++i) {
    uint32_t np = 0, md = 0;
    flatten_subtree<Rec>(items[i].n, items[i].f, items[i].slot, items[i].depth, out, on_node, np, md);
    items[i].n_points = np; items[i].max_depth = md;
  }
  for (int64_t i = 0; i < n_items; ++i) { res.n_points += items[i].n_points; if (items[i].max_depth > res.max_depth) res.max_depth = items[i].max_depth; }
  res.n_nodes = (uint32_t)total;
  return res;
}
}  // namespace detail

// Parallel variants: same output as flatten_ikdtree / flatten_ikdtree_compact.  on_node is called concurrently from
// several threads (each slot exactly once).  grain = largest subtree (by KD_TREE_NODE::TreeSize) flattened by one thread.
template <class Node, class PointFn>
FlattenResult flatten_ikdtree_parallel(const Node* root, malio_map_node* out, uint32_t capacity, PointFn&& on_node,
                                       uint32_t grain = 16384) {
  return detail::flatten_parallel<detail::RecFull>(root, out, capacity, on_node, nullptr, grain);
}
template <class Node, class PointFn>
FlattenResult flatten_ikdtree_compact_parallel(const Node* root, malio_map_point* out, uint32_t capacity, PointFn&& on_node,
                                               float* root_box_out = nullptr, uint32_t grain = 16384) {
  return detail::flatten_parallel<detail::RecCompact>(root, out, capacity, on_node, root_box_out, grain);
}

}  // namespace malio
#endif  // MALIO_FLATTEN_HPP_
