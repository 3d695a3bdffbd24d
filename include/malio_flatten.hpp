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

}  // namespace malio
#endif  // MALIO_FLATTEN_HPP_
