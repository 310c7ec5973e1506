// odom_sequence.cpp — compile-and-run check of the drop-in facade: OdomNode's GICP call sites
// (reference src/dlo/odom.cc:100-120 ctor setup, :472-507 initializeInputTarget, :514-528 setInputSources,
// :792-852 getNextPose) written the way odom.cc writes them, against include/nano_gicp/nano_gicp.hpp.
// Input : a binary file of pre-voxelised scans  [int32 nscans][float T0[16] col-major]{[int32 n][n*8 float]}*
// Output: one JSON line per scan with the S2S / S2M transforms and iteration counts.
#include <cmath>
#include <cstdio>
#include <string>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include <nano_gicp/nano_gicp.hpp>

typedef pcl::PointXYZI PointType;

static void transformPointCloud(const pcl::PointCloud<PointType>& in, pcl::PointCloud<PointType>& out, const Eigen::Matrix4f& T) {
  out = in;
  for (size_t i = 0; i < in.points.size(); i++) {
    const PointType& p = in.points[i];
    out.points[i].x = T(0, 0) * p.x + T(0, 1) * p.y + T(0, 2) * p.z + T(0, 3);
    out.points[i].y = T(1, 0) * p.x + T(1, 1) * p.y + T(1, 2) * p.z + T(1, 3);
    out.points[i].z = T(2, 0) * p.x + T(2, 1) * p.y + T(2, 2) * p.z + T(2, 3);
  }
}

static void print_T(const char* name, const Eigen::Matrix4f& T) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < 16; i++) std::printf("%s%.9g", i ? ", " : "", T.data()[i]);
  std::printf("]");
}

struct OdomNodeLike {
  nano_gicp::NanoGICP<PointType, PointType> gicp_s2s;
  nano_gicp::NanoGICP<PointType, PointType> gicp;
  nano_gicp::VoxelGrid<PointType> vf_submap;
  nano_gicp::KeyframeStore store;     // additive: device-resident keyframes
  bool use_store = false;
  pcl::PointCloud<PointType>::Ptr current_scan, target_cloud, keyframe_cloud, submap_cloud;
  std::vector<std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>>> keyframe_normals;
  std::vector<Eigen::Matrix4d, Eigen::aligned_allocator<Eigen::Matrix4d>> submap_normals;
  Eigen::Matrix4f T, T_s2s, T_s2s_prev;
  bool submap_hasChanged = true;

  OdomNodeLike() {
    // odom.cc:100-120 with the shipped cfg/params.yaml values
    gicp_s2s.setCorrespondenceRandomness(10);
    gicp_s2s.setMaxCorrespondenceDistance(1.0);
    gicp_s2s.setMaximumIterations(32);
    gicp_s2s.setTransformationEpsilon(0.01);
    gicp_s2s.setEuclideanFitnessEpsilon(0.01);
    gicp_s2s.setRANSACIterations(5);
    gicp_s2s.setRANSACOutlierRejectionThreshold(1.0);
    gicp.setCorrespondenceRandomness(20);
    gicp.setMaxCorrespondenceDistance(0.5);
    gicp.setMaximumIterations(32);
    gicp.setTransformationEpsilon(0.01);
    gicp.setEuclideanFitnessEpsilon(0.01);
    gicp.setRANSACIterations(5);
    gicp.setRANSACOutlierRejectionThreshold(1.0);
    std::shared_ptr<int> temp;  // stands in for pcl::Registration<...>::KdTreeReciprocalPtr
    gicp_s2s.setSearchMethodSource(temp, true);
    gicp_s2s.setSearchMethodTarget(temp, true);
    gicp.setSearchMethodSource(temp, true);
    gicp.setSearchMethodTarget(temp, true);
    vf_submap.setLeafSize(0.5f, 0.5f, 0.5f);
    keyframe_cloud.reset(new pcl::PointCloud<PointType>);
  }

  void initializeInputTarget() {
    target_cloud = current_scan;
    gicp_s2s.setInputTarget(target_cloud);
    gicp_s2s.calculateTargetCovariances();
    pcl::PointCloud<PointType>::Ptr first_keyframe(new pcl::PointCloud<PointType>);
    transformPointCloud(*target_cloud, *first_keyframe, T);
    vf_submap.setInputCloud(first_keyframe);
    vf_submap.filter(*first_keyframe);
    *keyframe_cloud = *first_keyframe;
    gicp_s2s.setInputSource(keyframe_cloud);
    gicp_s2s.calculateSourceCovariances();
    keyframe_normals.push_back(gicp_s2s.getSourceCovariances());
    store.push(gicp_s2s);
    submap_cloud = first_keyframe;
    submap_normals = keyframe_normals[0];
  }

  void setInputSources() {
    gicp_s2s.setInputSource(current_scan);
    gicp.registerInputSource(current_scan);
    gicp.source_kdtree_ = gicp_s2s.source_kdtree_;
    gicp.source_covs_.clear();
  }

  void getNextPose(int idx) {
    pcl::PointCloud<PointType>::Ptr aligned(new pcl::PointCloud<PointType>);
    gicp_s2s.align(*aligned);
    Eigen::Matrix4f T_S2S = gicp_s2s.getFinalTransformation();
    T_s2s = T_s2s_prev * T_S2S;   // propagateS2S
    T_s2s_prev = T_s2s;
    gicp.source_covs_ = gicp_s2s.source_covs_;
    gicp_s2s.swapSourceAndTarget();
    if (submap_hasChanged) {
      if (use_store) {
        store.setTarget(gicp, std::vector<int>{0});       // additive path: submap assembled on the device
      } else {
        gicp.setInputTarget(submap_cloud);
        gicp.setTargetCovariances(submap_normals);
      }
      submap_hasChanged = false;
    }
    gicp.align(*aligned, T_s2s);
    T = gicp.getFinalTransformation();
    T_s2s_prev = T;
    std::printf("{\"scan\": %d, \"s2s_iterations\": %d, \"s2s_trials\": %d, \"s2m_iterations\": %d, \"s2m_trials\": %d, \"s2m_converged\": %d, \"aligned_points\": %zu, ",
                idx, gicp_s2s.getLastResult().nr_iterations, gicp_s2s.getLastResult().n_compute_error, gicp.getLastResult().nr_iterations,
                gicp.getLastResult().n_compute_error, (int)gicp.hasConverged(), aligned->points.size());
    print_T("T_S2S", T_S2S);
    std::printf(", ");
    print_T("T", T);
    std::printf("}\n");
  }
};

// a sensor_msgs::PointCloud2 stand-in (ROS is not installed here): Ouster-style 48-byte records, 64 rows
struct FieldLike { std::string name; uint32_t offset; uint8_t datatype; uint32_t count; };
struct Pc2Like {
  uint32_t height = 0, width = 0, point_step = 0, row_step = 0;
  bool is_bigendian = false;
  std::vector<FieldLike> fields;
  std::vector<uint8_t> data;
};
static int check_pointcloud2_path(const pcl::PointCloud<PointType>& scan) {
  Pc2Like msg;
  const size_t n = (scan.points.size() / 64) * 64;
  msg.height = 64; msg.width = (uint32_t)(n / 64); msg.point_step = 48; msg.row_step = msg.width * 48;
  msg.fields = {{"x", 0, 7, 1}, {"y", 4, 7, 1}, {"z", 8, 7, 1}, {"intensity", 16, 7, 1}, {"t", 20, 6, 1}, {"ring", 26, 2, 1}};
  msg.data.assign(n * 48, 0xab);
  pcl::PointCloud<PointType> plain;
  plain.points.assign(scan.points.begin(), scan.points.begin() + (long)n);
  for (size_t i = 0; i < n; i++) {
    std::memcpy(&msg.data[i * 48], &scan.points[i].x, 12);
    std::memcpy(&msg.data[i * 48 + 16], &scan.points[i].intensity, 4);
  }
  nano_gicp::Preprocessor<PointType> pre;
  pre.setCropSize(1.0f); pre.setLeafSize(0.5f);
  pcl::PointCloud<PointType> a, b;
  pre.filter(plain, a);
  pre.filterMsg(msg, b);
  if (a.points.size() != b.points.size() || a.points.empty()) return 1;
  return std::memcmp(a.points.data(), b.points.data(), a.points.size() * sizeof(PointType)) == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scans.bin\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror("open"); return 2; }
  int nscans = 0;
  float T0[16];
  if (std::fread(&nscans, 4, 1, f) != 1 || std::fread(T0, 4, 16, f) != 16) return 2;
  OdomNodeLike node;
  if (!node.gicp.handle() || !node.gicp_s2s.handle()) return 3;   // no GPU: fail loudly
  node.use_store = argc > 2 && std::string(argv[2]) == "store";
  for (int i = 0; i < 16; i++) node.T.data()[i] = T0[i];
  node.T_s2s = node.T;
  node.T_s2s_prev = node.T;
  for (int s = 0; s < nscans; s++) {
    int n = 0;
    if (std::fread(&n, 4, 1, f) != 1) return 2;
    pcl::PointCloud<PointType>::Ptr scan(new pcl::PointCloud<PointType>);
    scan->resize((size_t)n);
    if (std::fread(scan->points.data(), sizeof(PointType), (size_t)n, f) != (size_t)n) return 2;
    node.current_scan = scan;
    if (s == 0) {
      if (check_pointcloud2_path(*scan) != 0) { std::fprintf(stderr, "PointCloud2 path differs from the plain path\n"); return 4; }
      node.initializeInputTarget();
      continue;
    }
    node.setInputSources();
    node.getNextPose(s);
  }
  std::fclose(f);
  return 0;
}
