// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for <pcl/point_types.h> so that the reference's
// include/ikd-Tree/ikd_Tree.{h,cpp} compile unmodified from /root/reference (see oracle/Makefile).
// ikd_Tree.h:11 includes this header only for the point structs and Eigen::aligned_allocator
// (ikd_Tree.h:62). Layout mirrors pcl::PointXYZINormal (48 bytes: xyz+pad | normal+pad | intensity,curvature,pad).
#pragma once
#include <cstring>
#include <vector>
#include <memory>
namespace Eigen { template <class T> using aligned_allocator = std::allocator<T>; }
namespace pcl {
struct PointXYZ { float x, y, z, pad; };
struct PointXYZI { float x, y, z, pad, intensity, pad2[3]; };
struct PointXYZINormal {
  float x, y, z, pad;
  float normal_x, normal_y, normal_z, pad2;
  float intensity, curvature, pad3[2];
};
}  // namespace pcl
