// odom_extract.cpp — "drops in unchanged", tested on the reference's OWN text: the bodies of
//   dlo::OdomNode::initializeInputTarget / setInputSources   (reference src/dlo/odom.cc:470-528)
//   dlo::OdomNode::getNextPose                               (:790-852)
//   dlo::OdomNode::updateKeyframes                           (:1096-1181)
//   dlo::OdomNode::pushSubmapIndices / getSubmapKeyframes    (:1210-1331, incl. the submap_normals concatenation)
// are cut out of /root/reference at build time (tests/cpp/Makefile -> _gen/*.inc, git-ignored, never committed) and
// compiled verbatim against include/nano_gicp/nano_gicp.hpp; dlo_stub.hpp supplies the class around them.
// Same input / output format as odom_sequence.cpp.
#include "dlo_stub.hpp"
#include <chrono>

#include "_gen/odom_470_528.inc"
#include "_gen/odom_790_852.inc"
#include "_gen/odom_1096_1181.inc"
#include "_gen/odom_1210_1331.inc"

static void print_T(const char* name, const Eigen::Matrix4f& T) {
  std::printf("\"%s\": [", name);
  for (int i = 0; i < 16; i++) std::printf("%s%.9g", i ? ", " : "", T.data()[i]);
  std::printf("]");
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s scans.bin [keyframe_thresh_dist]\n", argv[0]); return 2; }
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror("open"); return 2; }
  int nscans = 0;
  float T0[16];
  if (std::fread(&nscans, 4, 1, f) != 1 || std::fread(T0, 4, 16, f) != 16) return 2;
  dlo::OdomNode node;
  if (!node.gicp.handle() || !node.gicp_s2s.handle()) return 3;   // no GPU: fail loudly
  if (argc > 2) node.keyframe_thresh_dist_ = std::atof(argv[2]);   // cfg/params.yaml:39 (shipped 5 m)
  for (int i = 0; i < 16; i++) node.T.data()[i] = T0[i];
  node.T_s2s = node.T_s2s_prev = node.T;
  node.propagateS2M();
  for (int s = 0; s < nscans; s++) {
    int n = 0;
    if (std::fread(&n, 4, 1, f) != 1) return 2;
    pcl::PointCloud<PointType>::Ptr scan(new pcl::PointCloud<PointType>);
    scan->resize((size_t)n);
    if (std::fread(scan->points.data(), sizeof(PointType), (size_t)n, f) != (size_t)n) return 2;
    // icpCB, odom.cc:629-697: preprocess (done by the caller of this test), then
    node.current_scan = scan;
    node.source_cloud = scan;
    if (s == 0) { node.initializeInputTarget(); node.target_cloud.reset(new pcl::PointCloud<PointType>(*scan)); continue; }
    const auto t_begin = std::chrono::steady_clock::now();
    node.setInputSources();
    node.getNextPose();
    node.updateKeyframes();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    std::printf("{\"scan\": %d, \"ms\": %.4f, \"s2s_iterations\": %d, \"s2m_iterations\": %d, \"keyframes\": %d, \"submap_keyframes\": %zu, \"submap_points\": %zu, \"submap_normals\": %zu, ",
                s, ms, node.gicp_s2s.getLastResult().nr_iterations, node.gicp.getLastResult().nr_iterations, node.num_keyframes,
                node.submap_kf_idx_curr.size(), node.submap_cloud ? node.submap_cloud->points.size() : (size_t)0, node.submap_normals.size());
    print_T("T", node.T);
    std::printf("}\n");
  }
  std::fclose(f);
  return 0;
}
