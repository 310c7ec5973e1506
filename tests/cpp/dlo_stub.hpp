// dlo_stub.hpp — the environment the reference's own OdomNode text needs in order to compile HERE, where ROS, PCL,
// Eigen and Boost do not exist: a class dlo::OdomNode with the members that the extracted method bodies touch
// (declarations after reference include/dlo/odom.h), stand-ins for the handful of library calls they make, and
// plain restatements of the OdomNode helpers that lie outside the registration path (pose propagation, hulls).
// The method BODIES of initializeInputTarget / setInputSources / getNextPose / updateKeyframes / pushSubmapIndices /
// getSubmapKeyframes are NOT in this repository: tests/cpp/Makefile cuts them out of /root/reference/src/dlo/odom.cc
// at build time into tests/cpp/_gen/ (git-ignored) and odom_extract.cpp #includes them verbatim.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <math.h>     // the C++ wrappers put std::abs(float) into the global namespace: OdomNode writes abs(dd) on floats (odom.cc:1145-1153),
#include <stdlib.h>   // which would otherwise bind to int abs(int) here and truncate (any ROS header brings these in for the real build)
#include <limits>
#include <memory>
#include <queue>
#include <thread>
#include <utility>
#include <vector>

#include <nano_gicp/nano_gicp.hpp>

#ifndef NANO_GICP_B200_HAVE_PCL
namespace Eigen {
struct Quaternionf {
  float w_, x_, y_, z_;
  Quaternionf() : w_(1.f), x_(0.f), y_(0.f), z_(0.f) {}
  Quaternionf(float w, float x, float y, float z) : w_(w), x_(x), y_(y), z_(z) {}
  float& w() { return w_; } float& x() { return x_; } float& y() { return y_; } float& z() { return z_; }
  float w() const { return w_; } float x() const { return x_; } float y() const { return y_; } float z() const { return z_; }
  static Quaternionf Identity() { return Quaternionf(); }
  Quaternionf inverse() const { const float n2 = w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; return Quaternionf(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaternionf operator*(const Quaternionf& b) const {
    return Quaternionf(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                       w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
};
}  // namespace Eigen
namespace pcl {
template <class PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  if (&in != &out) out = in;
  for (size_t i = 0; i < out.points.size(); i++) {
    const float x = in.points[i].x, y = in.points[i].y, z = in.points[i].z;
    out.points[i].x = T(0, 0) * x + T(0, 1) * y + T(0, 2) * z + T(0, 3);
    out.points[i].y = T(1, 0) * x + T(1, 1) * y + T(1, 2) * z + T(1, 3);
    out.points[i].z = T(2, 0) * x + T(2, 1) * y + T(2, 2) * z + T(2, 3);
  }
}
}  // namespace pcl
namespace boost { using std::make_shared; }
#endif

typedef pcl::PointXYZI PointType;

namespace dlo {

class OdomNode {
 public:
  // ---- members used by the extracted bodies (types as in reference include/dlo/odom.h:60-150) ----
  double curr_frame_stamp = 0., prev_frame_stamp = 0.;
  pcl::PointCloud<PointType>::Ptr source_cloud, current_scan, current_scan_t, target_cloud;
  pcl::PointCloud<PointType>::Ptr keyframes_cloud, keyframe_cloud, submap_cloud;
  std::vector<std::pair<std::pair<Eigen::Vector3f, Eigen::Quaternionf>, pcl::PointCloud<PointType>::Ptr>> keyframes;
  std::vector<std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>> keyframe_normals;
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> submap_normals;
  std::vector<int> submap_kf_idx_curr, submap_kf_idx_prev, keyframe_convex, keyframe_concave;
  bool submap_hasChanged = true, vf_submap_use_ = true, imu_use_ = false;
  int num_keyframes = 0, submap_knn_ = 10, submap_kcv_ = 10, submap_kcc_ = 10;      // cfg/params.yaml:42-46
  double keyframe_thresh_dist_ = 5.0, keyframe_thresh_rot_ = 45.0;                  // cfg/params.yaml:39-40
  Eigen::Matrix4f T, T_s2s, T_s2s_prev, imu_SE3;
  Eigen::Vector3f pose;
  Eigen::Quaternionf rotq;
  std::thread publish_keyframe_thread;
  nano_gicp::NanoGICP<PointType, PointType> gicp_s2s, gicp;
  nano_gicp::VoxelGrid<PointType> vf_submap;

  // ---- the bodies cut out of the reference at build time ----
  void initializeInputTarget();
  void setInputSources();
  void getNextPose();
  void updateKeyframes();
  void pushSubmapIndices(std::vector<float> dists, int k, std::vector<int> frames);
  void getSubmapKeyframes();

  // ---- outside the registration path: restated / stubbed here ----
  OdomNode() {
    // odom.cc:100-127 with the shipped cfg/params.yaml values
    gicp_s2s.setCorrespondenceRandomness(10); gicp_s2s.setMaxCorrespondenceDistance(1.0);
    gicp_s2s.setMaximumIterations(32); gicp_s2s.setTransformationEpsilon(0.01);
    gicp.setCorrespondenceRandomness(20); gicp.setMaxCorrespondenceDistance(0.5);
    gicp.setMaximumIterations(32); gicp.setTransformationEpsilon(0.01);
    vf_submap.setLeafSize(0.5f, 0.5f, 0.5f);
    T = T_s2s = T_s2s_prev = imu_SE3 = Eigen::Matrix4f::Identity();
    keyframes_cloud.reset(new pcl::PointCloud<PointType>);
    keyframe_cloud.reset(new pcl::PointCloud<PointType>);
  }
  void publishKeyframe() {}
  void integrateIMU() {}
  static Eigen::Quaternionf quat_of(const Eigen::Matrix4f& M) {   // rotation block -> unit quaternion (trace > -1 branch suffices for the tests' yaw-only motion; general form kept)
    const float m00 = M(0, 0), m11 = M(1, 1), m22 = M(2, 2), tr = m00 + m11 + m22;
    Eigen::Quaternionf q;
    if (tr > 0.f) { const float s = std::sqrt(tr + 1.f) * 2.f; q = Eigen::Quaternionf(0.25f * s, (M(2, 1) - M(1, 2)) / s, (M(0, 2) - M(2, 0)) / s, (M(1, 0) - M(0, 1)) / s); }
    else if (m00 > m11 && m00 > m22) { const float s = std::sqrt(1.f + m00 - m11 - m22) * 2.f; q = Eigen::Quaternionf((M(2, 1) - M(1, 2)) / s, 0.25f * s, (M(0, 1) + M(1, 0)) / s, (M(0, 2) + M(2, 0)) / s); }
    else if (m11 > m22) { const float s = std::sqrt(1.f + m11 - m00 - m22) * 2.f; q = Eigen::Quaternionf((M(0, 2) - M(2, 0)) / s, (M(0, 1) + M(1, 0)) / s, 0.25f * s, (M(1, 2) + M(2, 1)) / s); }
    else { const float s = std::sqrt(1.f + m22 - m00 - m11) * 2.f; q = Eigen::Quaternionf((M(1, 0) - M(0, 1)) / s, (M(0, 2) + M(2, 0)) / s, (M(1, 2) + M(2, 1)) / s, 0.25f * s); }
    return q;
  }
  void propagateS2S(Eigen::Matrix4f Tl) {          // odom.cc:925-941: T_s2s = T_s2s_prev * T, then pose / rotation of it
    T_s2s = T_s2s_prev * Tl;
    T_s2s_prev = T_s2s;
  }
  void propagateS2M() {                            // odom.cc:948-968: pose / rotq of the global transform T
    pose[0] = T(0, 3); pose[1] = T(1, 3); pose[2] = T(2, 3);
    rotq = quat_of(T);
  }
  void transformCurrentScan() {                    // odom.cc:973-977
    current_scan_t.reset(new pcl::PointCloud<PointType>);
    pcl::transformPointCloud(*current_scan, *current_scan_t, T);
  }
  // computeConvexHull / computeConcaveHull (odom.cc:1017-1090): the reference asks pcl::ConvexHull / pcl::ConcaveHull
  // (alpha = keyframe_thresh_dist_, odom.cc:95-98) for the hull point indices of the keyframe positions; here the
  // library's own hulls (csrc/submap_select.cpp) answer through the C ABI
  std::vector<float> keyframe_positions() const {
    std::vector<float> xyz;
    for (const auto& k : keyframes) { xyz.push_back(k.first.first[0]); xyz.push_back(k.first.first[1]); xyz.push_back(k.first.first[2]); }
    return xyz;
  }
  void computeConvexHull() {
    if (num_keyframes < 4) return;
    const std::vector<float> xyz = keyframe_positions();
    std::vector<int> out(keyframes.size() + 1);
    const int m = ngicp_submap_convex_hull(xyz.data(), (int)keyframes.size(), out.data(), (int)out.size());
    keyframe_convex.assign(out.begin(), out.begin() + (m > 0 ? m : 0));
  }
  void computeConcaveHull() {
    if (num_keyframes < 5) return;
    const std::vector<float> xyz = keyframe_positions();
    std::vector<int> out(keyframes.size() + 1);
    const int m = ngicp_submap_concave_hull(xyz.data(), (int)keyframes.size(), keyframe_thresh_dist_, out.data(), (int)out.size());
    keyframe_concave.assign(out.begin(), out.begin() + (m > 0 ? m : 0));
  }
};

}  // namespace dlo
