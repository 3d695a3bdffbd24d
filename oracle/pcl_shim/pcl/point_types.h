// Minimal stand-in for <pcl/point_types.h>, test infrastructure only.
// Lets the reference's MA_LIO/include/ikd-Tree/ikd_Tree.{h,cpp} compile in place (PCL and Eigen are not
// installed in this image).  The two things ikd_Tree uses: the 48-byte pcl::PointXYZINormal layout
// (ikd_Tree.h:11,22) and Eigen::aligned_allocator (ikd_Tree.h:56).
#pragma once
#include <memory>
namespace pcl {
struct alignas(16) PointXYZINormal {
  float x = 0.f, y = 0.f, z = 0.f, pad0 = 1.f;                       // data[4]
  float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, pad1 = 0.f;  // data_n[4]
  float intensity = 0.f, curvature = 0.f, pad2 = 0.f, pad3 = 0.f;    // data_c[4]
};
static_assert(sizeof(PointXYZINormal) == 48, "PCL PointXYZINormal is 48 bytes");
}  // namespace pcl
namespace Eigen {
template <class T> using aligned_allocator = std::allocator<T>;
}
