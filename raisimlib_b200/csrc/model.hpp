// Host-side robot model: URDF subset -> constant tables consumed by the CUDA step kernel.
// Replaces the model-building half of raisim::World::addArticulatedSystem(urdf)
// (upstream source is not in the reference snapshot -- closed libraisim.so).
#pragma once
#include <array>
#include <string>
#include <vector>

namespace rsb {

enum JointType { JT_FIXED = 0, JT_REVOLUTE = 1, JT_PRISMATIC = 2, JT_FLOATING = 3 };
enum CollType { CT_SPHERE = 0, CT_BOX = 1, CT_CAPSULE = 2, CT_CYLINDER = 3 };
// contact candidates ("features") of a collision body against the terrain, one contact at most each
enum FeatType { FT_POINT = 0,     // sphere of radius pt_rad at pt_pos (radius 0: a box corner / cylinder rim sample)
                FT_SEGMENT = 1,   // interior of the segment pt_pos .. pt_pos2 swept by radius pt_rad (capsule / cylinder side) against terrain edges
                FT_BOXFACE = 2,   // terrain vertices inside the box of collision body pt_coll (a box resting on a peak)
                FT_RIM = 3 };     // lowest point of the rim circle of a cylinder cap: centre pt_pos, radius pt_rad, axis towards pt_pos2 (side rolling)

struct Frame {
  std::string name;    // link name
  std::string joint;   // name of the joint attaching the link (upstream names frames after joints); empty for the root
  int body;
  std::array<double, 3> pos;
  std::array<double, 9> rot;
};

struct Model {
  int nb = 0, nq = 0, nv = 0, floating = 0, maxdepth = 0;
  std::vector<int> parent, jtype, qidx, vidx, depth, subtree;
  std::vector<double> jpos, jrot, axis, mass, com, inertia, jlimit;   // 3,9,3,1,3,6,2 per body
  std::vector<double> jeffort, jvelocity;                              // <limit effort velocity> per body (1e30 = none); velocity is carried, not enforced
  std::vector<std::string> body_names, joint_names;
  // collision bodies
  std::vector<int> cbody, ctype;
  std::vector<double> csize, cpos, crot;                               // 3,3,9 per collision body
  std::vector<std::string> coll_names;
  // contact candidates: points first (sphere -> 1, capsule -> 2 end spheres, box -> 8 corners, cylinder -> 2 x 4 rim points), then, in
  // collision-body order, one FT_SEGMENT per capsule / cylinder and one FT_BOXFACE per box (height-map terrain only) and two FT_RIM
  // per cylinder (any terrain)
  std::vector<int> pt_body, pt_coll, pt_feat, pt_type;
  std::vector<double> pt_pos, pt_rad, pt_pos2;
  std::vector<Frame> frames;                                            // one per URDF link
  int skipped_collisions = 0;                                           // <mesh> collision bodies whose file could not be read (ignored)
  int mesh_boxes = 0;                                                   // <mesh> collision bodies replaced by their bounding box
  std::string source_dir;                                               // directory of the URDF file (mesh paths resolve against it)
  int ncoll() const { return (int)cbody.size(); }
  int npts() const { return (int)pt_body.size(); }
};

// Throws std::runtime_error with a readable message on malformed input.
Model load_urdf(const std::string& path_or_xml);

// binary cache of the compiled tables (same-architecture; throws std::runtime_error on I/O or format errors)
void save_model(const Model& md, const std::string& path);
Model load_model(const std::string& path);

}  // namespace rsb
