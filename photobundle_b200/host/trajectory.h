// trajectory.h — same interface and semantics as the reference's src/trajectory.{h,cc}.
#ifndef PBA_HOST_TRAJECTORY_H
#define PBA_HOST_TRAJECTORY_H
#include "compat.h"

class Trajectory {
 public:
  typedef int Id_t;
  Trajectory() {}
  // pose = estimated relative pose from VO; T_w_i = T_w_(i-1) * inv(T_i)   (src/trajectory.cc:7-16)
  void push_back(const Mat44& pose, const Id_t id);
  const Mat44& operator[](size_t i) const { return _data[i].pose; }
  Mat44& operator[](size_t i) { return _data[i].pose; }
  const Mat44& atId(const Id_t id) const;
  Mat44& atId(const Id_t id);
  const Mat44& back() const { return _data.back().pose; }
  EigenAlignedContainer_<Mat44> poses() const;
  EigenAlignedContainer_<Vec3> cameraPositions() const;
  size_t size() const { return _data.size(); }
 private:
  struct PoseWithId { Mat44 pose; Id_t id; };
  std::vector<PoseWithId> _data;
  int find(const Id_t id) const;
};
#endif
