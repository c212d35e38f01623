// URDF subset loader (product, host side).  See model.hpp.
// No XML library exists in the image (SURVEY.md section 0 [PROBE]), hence the small parser below.
#include "model.hpp"

#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>

namespace rsb {
namespace {

// ------------------------------------------------------------------ minimal XML ----------------
struct XmlNode {
  std::string tag;
  std::map<std::string, std::string> attr;
  std::vector<std::unique_ptr<XmlNode>> kids;
  const XmlNode* child(const std::string& t) const {
    for (auto& k : kids) if (k->tag == t) return k.get();
    return nullptr;
  }
  std::vector<const XmlNode*> children(const std::string& t) const {
    std::vector<const XmlNode*> r;
    for (auto& k : kids) if (k->tag == t) r.push_back(k.get());
    return r;
  }
  std::string get(const std::string& k, const std::string& dflt = "") const {
    auto it = attr.find(k);
    return it == attr.end() ? dflt : it->second;
  }
  bool has(const std::string& k) const { return attr.count(k) != 0; }
};

class XmlParser {
 public:
  explicit XmlParser(const std::string& s) : s_(s) {}
  std::unique_ptr<XmlNode> parse() {
    skip_misc();
    auto n = element();
    if (!n) fail("no root element");
    return n;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const std::string& m) const { throw std::runtime_error("XML parse error at byte " + std::to_string(i_) + ": " + m); }
  bool starts(const char* t) const { return s_.compare(i_, std::char_traits<char>::length(t), t) == 0; }
  void skip_ws() { while (i_ < s_.size() && isspace((unsigned char)s_[i_])) i_++; }
  void skip_misc() {   // whitespace, <?...?>, <!--...-->, <!DOCTYPE ...>
    for (;;) {
      skip_ws();
      if (starts("<?")) { size_t e = s_.find("?>", i_); if (e == std::string::npos) fail("unterminated <?"); i_ = e + 2; }
      else if (starts("<!--")) { size_t e = s_.find("-->", i_); if (e == std::string::npos) fail("unterminated comment"); i_ = e + 3; }
      else if (starts("<!")) { size_t e = s_.find('>', i_); if (e == std::string::npos) fail("unterminated <!"); i_ = e + 1; }
      else return;
    }
  }
  std::string name() {
    size_t b = i_;
    while (i_ < s_.size() && (isalnum((unsigned char)s_[i_]) || s_[i_] == '_' || s_[i_] == '-' || s_[i_] == ':' || s_[i_] == '.')) i_++;
    if (b == i_) fail("expected a name");
    return s_.substr(b, i_ - b);
  }
  std::unique_ptr<XmlNode> element() {
    if (i_ >= s_.size() || s_[i_] != '<') return nullptr;
    i_++;
    auto n = std::make_unique<XmlNode>();
    n->tag = name();
    for (;;) {
      skip_ws();
      if (i_ >= s_.size()) fail("unterminated tag <" + n->tag);
      if (s_[i_] == '/') { if (!starts("/>")) fail("bad tag end"); i_ += 2; return n; }
      if (s_[i_] == '>') { i_++; break; }
      std::string k = name();
      skip_ws();
      if (s_[i_] != '=') fail("expected '=' after attribute " + k);
      i_++; skip_ws();
      char q = s_[i_];
      if (q != '"' && q != '\'') fail("expected quoted value for " + k);
      size_t e = s_.find(q, i_ + 1);
      if (e == std::string::npos) fail("unterminated attribute value");
      n->attr[k] = s_.substr(i_ + 1, e - i_ - 1);
      i_ = e + 1;
    }
    for (;;) {   // content
      size_t lt = s_.find('<', i_);
      if (lt == std::string::npos) fail("missing </" + n->tag + ">");
      i_ = lt;
      if (starts("</")) {
        i_ += 2;
        std::string t = name();
        if (t != n->tag) fail("mismatched </" + t + "> for <" + n->tag + ">");
        skip_ws();
        if (s_[i_] != '>') fail("bad closing tag");
        i_++;
        return n;
      }
      if (starts("<!--")) { size_t e = s_.find("-->", i_); if (e == std::string::npos) fail("unterminated comment"); i_ = e + 3; continue; }
      if (starts("<![CDATA[")) { size_t e = s_.find("]]>", i_); if (e == std::string::npos) fail("unterminated CDATA"); i_ = e + 3; continue; }
      if (starts("<?")) { size_t e = s_.find("?>", i_); if (e == std::string::npos) fail("unterminated <?"); i_ = e + 2; continue; }
      n->kids.push_back(element());
    }
  }
};

// ------------------------------------------------------------------ small linear algebra -------
using V3 = std::array<double, 3>;
using M3 = std::array<double, 9>;

M3 mul(const M3& a, const M3& b) {
  M3 r{};
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r[3 * i + j] = a[3 * i] * b[j] + a[3 * i + 1] * b[3 + j] + a[3 * i + 2] * b[6 + j];
  return r;
}
M3 transpose(const M3& a) { return {a[0], a[3], a[6], a[1], a[4], a[7], a[2], a[5], a[8]}; }
V3 mul(const M3& a, const V3& v) { return {a[0] * v[0] + a[1] * v[1] + a[2] * v[2], a[3] * v[0] + a[4] * v[1] + a[5] * v[2], a[6] * v[0] + a[7] * v[1] + a[8] * v[2]}; }
V3 add(const V3& a, const V3& b) { return {a[0] + b[0], a[1] + b[1], a[2] + b[2]}; }
const M3 kEye = {1, 0, 0, 0, 1, 0, 0, 0, 1};

M3 rpy_to_rot(const V3& rpy) {   // URDF fixed-axis rpy: R = Rz(yaw) Ry(pitch) Rx(roll)
  double cr = std::cos(rpy[0]), sr = std::sin(rpy[0]), cp = std::cos(rpy[1]), sp = std::sin(rpy[1]), cy = std::cos(rpy[2]), sy = std::sin(rpy[2]);
  return {cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr,
          sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr,
          -sp, cp * sr, cp * cr};
}

V3 parse_v3(const std::string& s, const V3& dflt) {
  if (s.empty()) return dflt;
  std::istringstream is(s);
  V3 v{};
  if (!(is >> v[0] >> v[1] >> v[2])) throw std::runtime_error("URDF: expected three numbers in \"" + s + "\"");
  return v;
}
double parse_d(const std::string& s, const char* what) {
  char* e = nullptr;
  double v = std::strtod(s.c_str(), &e);
  if (s.empty() || e == s.c_str()) throw std::runtime_error(std::string("URDF: bad number for ") + what);
  return v;
}
void origin_of(const XmlNode* n, V3& pos, M3& rot) {
  pos = {0, 0, 0}; rot = kEye;
  const XmlNode* o = n ? n->child("origin") : nullptr;
  if (!o) return;
  pos = parse_v3(o->get("xyz"), {0, 0, 0});
  rot = rpy_to_rot(parse_v3(o->get("rpy"), {0, 0, 0}));
}

// rigid-body inertia accumulator expressed in the owning body's frame
struct Accum {
  double m = 0;
  V3 c{0, 0, 0};
  M3 I{};   // about c
  static M3 shift(double m, const V3& d) {   // m ((d.d) 1 - d d^T)
    double dd = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    return {m * (dd - d[0] * d[0]), -m * d[0] * d[1], -m * d[0] * d[2],
            -m * d[1] * d[0], m * (dd - d[1] * d[1]), -m * d[1] * d[2],
            -m * d[2] * d[0], -m * d[2] * d[1], m * (dd - d[2] * d[2])};
  }
  void add_body(double m2, const V3& c2, const M3& I2) {
    if (m2 <= 0) return;
    double mt = m + m2;
    V3 cn = {(m * c[0] + m2 * c2[0]) / mt, (m * c[1] + m2 * c2[1]) / mt, (m * c[2] + m2 * c2[2]) / mt};
    M3 s1 = shift(m, {c[0] - cn[0], c[1] - cn[1], c[2] - cn[2]});
    M3 s2 = shift(m2, {c2[0] - cn[0], c2[1] - cn[1], c2[2] - cn[2]});
    for (int k = 0; k < 9; k++) I[k] = I[k] + s1[k] + I2[k] + s2[k];
    m = mt; c = cn;
  }
};

// Axis-aligned bounds of a mesh file (binary or ASCII STL, Wavefront OBJ), scaled.  `filename` may be absolute, file://, relative to
// the URDF's directory, or package://pkg/rest (tried as <dir>/rest, <dir>/../rest and <dir>/../../rest, the usual package layouts).
static bool mesh_bounds(const std::string& dir, std::string filename, const V3& scale, V3& lo, V3& hi) {
  std::vector<std::string> tries;
  if (filename.rfind("file://", 0) == 0) filename = filename.substr(7);
  if (filename.rfind("package://", 0) == 0) {
    std::string rest = filename.substr(10);
    size_t slash = rest.find('/');
    std::string tail = slash == std::string::npos ? rest : rest.substr(slash + 1);
    for (const char* up : {"", "../", "../../"}) { tries.push_back(dir + "/" + up + tail); tries.push_back(dir + "/" + up + rest); }
  } else {
    tries.push_back(filename);
    if (!dir.empty()) tries.push_back(dir + "/" + filename);
  }
  for (const std::string& path : tries) {
    std::ifstream f(path, std::ios::binary);
    if (!f) continue;
    std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    bool any = false;
    auto take = [&](double x, double y, double z) {
      const double v[3] = {x * scale[0], y * scale[1], z * scale[2]};
      for (int k = 0; k < 3; k++) { lo[k] = any ? std::min(lo[k], v[k]) : v[k]; hi[k] = any ? std::max(hi[k], v[k]) : v[k]; }
      any = true;
    };
    const bool obj = path.size() > 4 && (path.substr(path.size() - 4) == ".obj" || path.substr(path.size() - 4) == ".OBJ");
    if (obj) {
      std::istringstream ss(data); std::string line;
      while (std::getline(ss, line)) if (line.size() > 2 && line[0] == 'v' && line[1] == ' ') { double x, y, z; if (std::sscanf(line.c_str() + 2, "%lf %lf %lf", &x, &y, &z) == 3) take(x, y, z); }
    } else if (data.size() >= 84) {
      uint32_t ntri = 0; std::memcpy(&ntri, data.data() + 80, 4);
      if (data.size() == 84 + size_t(ntri) * 50) {           // binary STL: 80-byte header, count, 50 bytes per triangle
        for (uint32_t t = 0; t < ntri; t++) for (int v = 0; v < 3; v++) { float c[3]; std::memcpy(c, data.data() + 84 + size_t(t) * 50 + 12 + 12 * v, 12); take(c[0], c[1], c[2]); }
      } else {                                                 // ASCII STL: "vertex x y z"
        size_t pos = 0;
        while ((pos = data.find("vertex", pos)) != std::string::npos) { double x, y, z; if (std::sscanf(data.c_str() + pos + 6, "%lf %lf %lf", &x, &y, &z) == 3) take(x, y, z); pos += 6; }
      }
    }
    if (any) return true;
  }
  return false;
}

struct Builder {
  Model md;
  std::map<std::string, const XmlNode*> links;
  std::map<std::string, std::vector<const XmlNode*>> kids_of;   // parent link -> joints, file order
  std::vector<Accum> acc;
  std::map<std::string, std::string> joint_of_link;              // child link -> name of the joint that carries it

  int new_body(const std::string& name, int parent, int jt, const V3& jp, const M3& jr, const V3& ax, const std::string& jname, double lo, double hi,
               double effort = 1e30, double velocity = 1e30) {
    int i = md.nb++;
    md.parent.push_back(parent); md.jtype.push_back(jt);
    md.jpos.insert(md.jpos.end(), jp.begin(), jp.end());
    md.jrot.insert(md.jrot.end(), jr.begin(), jr.end());
    md.axis.insert(md.axis.end(), ax.begin(), ax.end());
    md.body_names.push_back(name); md.joint_names.push_back(jname);
    md.jlimit.push_back(lo); md.jlimit.push_back(hi);
    md.jeffort.push_back(effort); md.jvelocity.push_back(velocity);
    acc.emplace_back();
    return i;
  }

  // merge `link` (whose frame sits at (pos, rot) in body coordinates) into body `b`, then descend
  void absorb(int b, const std::string& link_name, const V3& pos, const M3& rot) {
    auto it = links.find(link_name);
    if (it == links.end()) throw std::runtime_error("URDF: joint refers to unknown link '" + link_name + "'");
    const XmlNode* l = it->second;
    md.frames.push_back(Frame{link_name, joint_of_link.count(link_name) ? joint_of_link[link_name] : std::string(), b, pos, rot});
    if (const XmlNode* in = l->child("inertial")) {
      V3 cp; M3 cr;
      origin_of(in, cp, cr);
      const XmlNode* mn = in->child("mass");
      const XmlNode* it2 = in->child("inertia");
      if (!mn || !it2) throw std::runtime_error("URDF: <inertial> of link '" + link_name + "' needs <mass> and <inertia>");
      double m = parse_d(mn->get("value"), "mass");
      double ixx = parse_d(it2->get("ixx"), "ixx"), ixy = parse_d(it2->get("ixy", "0"), "ixy"), ixz = parse_d(it2->get("ixz", "0"), "ixz");
      double iyy = parse_d(it2->get("iyy"), "iyy"), iyz = parse_d(it2->get("iyz", "0"), "iyz"), izz = parse_d(it2->get("izz"), "izz");
      M3 Il = {ixx, ixy, ixz, ixy, iyy, iyz, ixz, iyz, izz};
      M3 Rl = mul(rot, cr);
      acc[b].add_body(m, add(pos, mul(rot, cp)), mul(mul(Rl, Il), transpose(Rl)));
    }
    for (const XmlNode* c : l->children("collision")) {
      V3 cp; M3 cr;
      origin_of(c, cp, cr);
      const XmlNode* g = c->child("geometry");
      if (!g) throw std::runtime_error("URDF: <collision> without <geometry> in link '" + link_name + "'");
      V3 p = add(pos, mul(rot, cp));
      M3 R = mul(rot, cr);
      int type; V3 size{0, 0, 0};
      if (const XmlNode* s = g->child("sphere")) { type = CT_SPHERE; size[0] = parse_d(s->get("radius"), "sphere radius"); }
      else if (const XmlNode* bx = g->child("box")) { type = CT_BOX; V3 sz = parse_v3(bx->get("size"), {0, 0, 0}); size = {0.5 * sz[0], 0.5 * sz[1], 0.5 * sz[2]}; }
      else if (const XmlNode* cp2 = g->child("capsule")) { type = CT_CAPSULE; size[0] = parse_d(cp2->get("radius"), "capsule radius"); size[1] = 0.5 * parse_d(cp2->get("length"), "capsule length"); }
      else if (const XmlNode* cy = g->child("cylinder")) { type = CT_CYLINDER; size[0] = parse_d(cy->get("radius"), "cylinder radius"); size[1] = 0.5 * parse_d(cy->get("length"), "cylinder length"); }
      else if (const XmlNode* me = g->child("mesh")) {
        // mesh collision bodies are not part of this path: a mesh whose file can be read is replaced by its
        // bounding box (centre and half extents in the collision frame, <mesh scale> applied); an unreadable one is skipped and counted
        V3 lo3, hi3;
        if (!mesh_bounds(md.source_dir, me->get("filename"), parse_v3(me->get("scale", "1 1 1"), {1, 1, 1}), lo3, hi3)) { md.skipped_collisions++; continue; }
        type = CT_BOX;
        V3 ctr = {0.5 * (lo3[0] + hi3[0]), 0.5 * (lo3[1] + hi3[1]), 0.5 * (lo3[2] + hi3[2])};
        size = {0.5 * (hi3[0] - lo3[0]), 0.5 * (hi3[1] - lo3[1]), 0.5 * (hi3[2] - lo3[2])};
        p = add(p, mul(R, ctr));
        md.mesh_boxes++;
      }
      else throw std::runtime_error("URDF: unsupported collision geometry in link '" + link_name + "' (sphere, box, capsule, cylinder are supported)");
      int ci = md.ncoll();
      md.cbody.push_back(b); md.ctype.push_back(type);
      md.csize.insert(md.csize.end(), size.begin(), size.end());
      md.cpos.insert(md.cpos.end(), p.begin(), p.end());
      md.crot.insert(md.crot.end(), R.begin(), R.end());
      md.coll_names.push_back(link_name);
      auto add_pt = [&](const V3& local, double rad, int feat) {
        V3 w = add(p, mul(R, local));
        md.pt_body.push_back(b); md.pt_coll.push_back(ci); md.pt_feat.push_back(feat); md.pt_type.push_back(FT_POINT);
        md.pt_pos.insert(md.pt_pos.end(), w.begin(), w.end());
        md.pt_pos2.insert(md.pt_pos2.end(), w.begin(), w.end());
        md.pt_rad.push_back(rad);
      };
      if (type == CT_SPHERE) add_pt({0, 0, 0}, size[0], 0);
      else if (type == CT_CAPSULE) { add_pt({0, 0, -size[1]}, size[0], 0); add_pt({0, 0, size[1]}, size[0], 1); }
      else if (type == CT_CYLINDER) {   // four rim points per end cap (axis = local z)
        for (int k = 0; k < 8; k++) {
          const double cx[4] = {1, 0, -1, 0}, sy[4] = {0, 1, 0, -1};
          add_pt({size[0] * cx[k & 3], size[0] * sy[k & 3], (k & 4) ? size[1] : -size[1]}, 0.0, k);
        }
      }
      else for (int k = 0; k < 8; k++) add_pt({(k & 1) ? size[0] : -size[0], (k & 2) ? size[1] : -size[1], (k & 4) ? size[2] : -size[2]}, 0.0, k);
    }
    auto kit = kids_of.find(link_name);
    if (kit == kids_of.end()) return;
    for (const XmlNode* j : kit->second) {
      V3 op; M3 orot;
      origin_of(j, op, orot);
      std::string type = j->get("type");
      std::string child = j->child("child")->get("link");
      V3 jp = add(pos, mul(rot, op));
      M3 jr = mul(rot, orot);
      if (type == "fixed") { absorb(b, child, jp, jr); continue; }
      int jt;
      if (type == "revolute" || type == "continuous") jt = JT_REVOLUTE;
      else if (type == "prismatic") jt = JT_PRISMATIC;
      else throw std::runtime_error("URDF: unsupported joint type '" + type + "' (joint '" + j->get("name") + "')");
      V3 ax = {1, 0, 0};
      if (const XmlNode* a = j->child("axis")) ax = parse_v3(a->get("xyz"), {1, 0, 0});
      double nrm = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
      if (!(nrm > 0)) throw std::runtime_error("URDF: zero joint axis in joint '" + j->get("name") + "'");
      for (double& v : ax) v /= nrm;
      double lo = -1e30, hi = 1e30;
      const XmlNode* lim = j->child("limit");
      if (lim && type != "continuous") {
        if (lim->has("lower")) lo = parse_d(lim->get("lower"), "limit lower");
        if (lim->has("upper")) hi = parse_d(lim->get("upper"), "limit upper");
      }
      double effort = 1e30, velocity = 1e30;       // <limit effort velocity>: 0 or absent = unlimited (URDFs often carry effort="0" placeholders)
      if (lim) {
        if (lim->has("effort")) { double e = parse_d(lim->get("effort"), "limit effort"); if (e > 0) effort = e; }
        if (lim->has("velocity")) { double v = parse_d(lim->get("velocity"), "limit velocity"); if (v > 0) velocity = v; }
      }
      int nb = new_body(child, b, jt, jp, jr, ax, j->get("name"), lo, hi, effort, velocity);
      absorb(nb, child, {0, 0, 0}, kEye);
    }
  }
};

}  // namespace

// One FT_SEGMENT per capsule / cylinder (its axis, swept by the radius) and one FT_BOXFACE per box, appended after the point
// candidates in collision-body order: the parts of a shape that can touch a height map where no end sphere / corner does
// (a capsule lying across a ridge, a box resting on a peak).  Positions in the body frame.
static void append_shape_features(Model& md) {
  for (int ci = 0; ci < md.ncoll(); ci++) {
    const int type = md.ctype[ci];
    if (type == CT_SPHERE) continue;
    const double* cp = &md.cpos[3 * ci]; const double* R = &md.crot[9 * ci]; const double* sz = &md.csize[3 * ci];
    md.pt_body.push_back(md.cbody[ci]); md.pt_coll.push_back(ci); md.pt_feat.push_back(0);
    if (type == CT_BOX) {
      md.pt_type.push_back(FT_BOXFACE);
      for (int k = 0; k < 3; k++) { md.pt_pos.push_back(cp[k]); md.pt_pos2.push_back(cp[k]); }
      md.pt_rad.push_back(0.0);
    } else {   // capsule, cylinder: axis = local z of the collision frame, half length sz[1]
      md.pt_type.push_back(FT_SEGMENT);
      for (int k = 0; k < 3; k++) { md.pt_pos.push_back(cp[k] - R[3 * k + 2] * sz[1]); md.pt_pos2.push_back(cp[k] + R[3 * k + 2] * sz[1]); }
      md.pt_rad.push_back(sz[0]);
      if (type == CT_CYLINDER)   // the lowest rim point of each cap moves round the rim as the cylinder rolls: no fixed sample can stand in for it
        for (int cap = 0; cap < 2; cap++) {
          const double sg = cap ? 1.0 : -1.0;
          md.pt_body.push_back(md.cbody[ci]); md.pt_coll.push_back(ci); md.pt_feat.push_back(cap); md.pt_type.push_back(FT_RIM);
          for (int k = 0; k < 3; k++) { md.pt_pos.push_back(cp[k] + sg * R[3 * k + 2] * sz[1]); md.pt_pos2.push_back(cp[k] - sg * R[3 * k + 2] * sz[1]); }
          md.pt_rad.push_back(sz[0]);
        }
    }
  }
}

Model load_urdf(const std::string& path_or_xml) {
  std::string text;
  size_t first = path_or_xml.find_first_not_of(" \t\r\n");
  if (first != std::string::npos && path_or_xml[first] == '<') text = path_or_xml;
  else {
    std::ifstream f(path_or_xml);
    if (!f) throw std::runtime_error("cannot open URDF file '" + path_or_xml + "'");
    std::stringstream ss; ss << f.rdbuf();
    text = ss.str();
  }
  XmlParser parser(text);
  std::unique_ptr<XmlNode> root = parser.parse();
  if (root->tag != "robot") throw std::runtime_error("URDF: root element must be <robot>");
  Builder bd;
  if (!(first != std::string::npos && path_or_xml[first] == '<')) {
    size_t sl = path_or_xml.find_last_of('/');
    bd.md.source_dir = sl == std::string::npos ? std::string(".") : path_or_xml.substr(0, sl);
  }
  std::map<std::string, bool> is_child;
  for (const XmlNode* l : root->children("link")) {
    std::string n = l->get("name");
    if (n.empty()) throw std::runtime_error("URDF: <link> without a name");
    if (bd.links.count(n)) throw std::runtime_error("URDF: duplicate link '" + n + "'");
    bd.links[n] = l;
  }
  for (const XmlNode* j : root->children("joint")) {
    const XmlNode *p = j->child("parent"), *c = j->child("child");
    if (!p || !c) throw std::runtime_error("URDF: joint '" + j->get("name") + "' needs <parent> and <child>");
    if (is_child[c->get("link")]) throw std::runtime_error("URDF: link '" + c->get("link") + "' has two parents (kinematic loops are unsupported)");
    is_child[c->get("link")] = true;
    bd.kids_of[p->get("link")].push_back(j);
    if (const XmlNode* ch = j->child("child")) bd.joint_of_link[ch->get("link")] = j->get("name");
  }
  std::string root_link;
  for (const XmlNode* l : root->children("link"))
    if (!is_child[l->get("name")]) {
      if (!root_link.empty()) throw std::runtime_error("URDF: more than one root link ('" + root_link + "', '" + l->get("name") + "')");
      root_link = l->get("name");
    }
  if (root_link.empty()) throw std::runtime_error("URDF: no root link");
  Model& md = bd.md;
  md.floating = (root_link != "world") ? 1 : 0;
  bd.new_body(root_link, -1, md.floating ? JT_FLOATING : JT_FIXED, {0, 0, 0}, kEye, {0, 0, 1}, "root", -1e30, 1e30);
  bd.absorb(0, root_link, {0, 0, 0}, kEye);

  md.qidx.assign(md.nb, 0); md.vidx.assign(md.nb, 0); md.depth.assign(md.nb, 0); md.subtree.assign(md.nb, 1);
  md.nq = md.floating ? 7 : 0; md.nv = md.floating ? 6 : 0;
  for (int i = 1; i < md.nb; i++) {
    md.qidx[i] = md.nq++; md.vidx[i] = md.nv++;
    md.depth[i] = md.depth[md.parent[i]] + 1;
    md.maxdepth = std::max(md.maxdepth, md.depth[i]);
  }
  for (int i = md.nb - 1; i > 0; i--) md.subtree[md.parent[i]] += md.subtree[i];
  md.mass.resize(md.nb); md.com.resize(3 * md.nb); md.inertia.resize(6 * md.nb);
  for (int i = 0; i < md.nb; i++) {
    const Accum& a = bd.acc[i];
    md.mass[i] = a.m;
    for (int k = 0; k < 3; k++) md.com[3 * i + k] = a.c[k];
    md.inertia[6 * i + 0] = a.I[0]; md.inertia[6 * i + 1] = a.I[1]; md.inertia[6 * i + 2] = a.I[2];
    md.inertia[6 * i + 3] = a.I[4]; md.inertia[6 * i + 4] = a.I[5]; md.inertia[6 * i + 5] = a.I[8];
    if (i > 0 && !(a.m > 0)) throw std::runtime_error("URDF: movable body '" + md.body_names[i] + "' has no mass");
  }
  append_shape_features(md);
  return md;
}

// ---- binary model cache: the compiled tables of a description, so that large fleets of workers need not
//      re-parse XML.  Layout: magic "RSBM", version, then every field of Model in declaration order; vectors and strings
//      carry a 64-bit length.  Little-endian, same-architecture cache (not an interchange format).
namespace {
constexpr uint32_t kCacheMagic = 0x4d425352u, kCacheVersion = 4u;
struct Writer {
  std::ofstream f;
  template <class T> void pod(const T& v) { f.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
  template <class T> void vec(const std::vector<T>& v) { pod<uint64_t>(v.size()); if (!v.empty()) f.write(reinterpret_cast<const char*>(v.data()), sizeof(T) * v.size()); }
  void str(const std::string& s) { pod<uint64_t>(s.size()); f.write(s.data(), (std::streamsize)s.size()); }
  void strs(const std::vector<std::string>& v) { pod<uint64_t>(v.size()); for (const auto& s : v) str(s); }
};
struct Reader {
  std::ifstream f;
  template <class T> void pod(T& v) { f.read(reinterpret_cast<char*>(&v), sizeof(T)); if (!f) throw std::runtime_error("model cache: truncated file"); }
  uint64_t len() { uint64_t n; pod(n); if (n > (1u << 24)) throw std::runtime_error("model cache: corrupt length"); return n; }
  template <class T> void vec(std::vector<T>& v) { v.resize(len()); if (!v.empty()) { f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * v.size()); if (!f) throw std::runtime_error("model cache: truncated file"); } }
  void str(std::string& s) { s.resize(len()); if (!s.empty()) { f.read(&s[0], (std::streamsize)s.size()); if (!f) throw std::runtime_error("model cache: truncated file"); } }
  void strs(std::vector<std::string>& v) { v.resize(len()); for (auto& s : v) str(s); }
};
template <class IO, class M> void model_fields(IO& io, M& md) {
  io.pod(md.nb); io.pod(md.nq); io.pod(md.nv); io.pod(md.floating); io.pod(md.maxdepth); io.pod(md.skipped_collisions);
  io.vec(md.parent); io.vec(md.jtype); io.vec(md.qidx); io.vec(md.vidx); io.vec(md.depth); io.vec(md.subtree);
  io.vec(md.jpos); io.vec(md.jrot); io.vec(md.axis); io.vec(md.mass); io.vec(md.com); io.vec(md.inertia); io.vec(md.jlimit);
  io.vec(md.jeffort); io.vec(md.jvelocity); io.pod(md.mesh_boxes);
  io.strs(md.body_names); io.strs(md.joint_names);
  io.vec(md.cbody); io.vec(md.ctype); io.vec(md.csize); io.vec(md.cpos); io.vec(md.crot); io.strs(md.coll_names);
  io.vec(md.pt_body); io.vec(md.pt_coll); io.vec(md.pt_feat); io.vec(md.pt_type); io.vec(md.pt_pos); io.vec(md.pt_pos2); io.vec(md.pt_rad);
}
}  // namespace

void save_model(const Model& md, const std::string& path) {
  Writer w; w.f.open(path, std::ios::binary | std::ios::trunc);
  if (!w.f) throw std::runtime_error("model cache: cannot write '" + path + "'");
  w.pod(kCacheMagic); w.pod(kCacheVersion);
  model_fields(w, const_cast<Model&>(md));
  w.pod<uint64_t>(md.frames.size());
  for (const Frame& fr : md.frames) { w.str(fr.name); w.str(fr.joint); w.pod(fr.body); w.pod(fr.pos); w.pod(fr.rot); }
  w.f.flush();
  if (!w.f) throw std::runtime_error("model cache: write to '" + path + "' failed");
}

Model load_model(const std::string& path) {
  Reader r; r.f.open(path, std::ios::binary);
  if (!r.f) throw std::runtime_error("model cache: cannot open '" + path + "'");
  uint32_t magic = 0, version = 0;
  r.pod(magic); r.pod(version);
  if (magic != kCacheMagic) throw std::runtime_error("model cache: '" + path + "' is not a model cache");
  if (version != kCacheVersion) throw std::runtime_error("model cache: '" + path + "' has version " + std::to_string(version) + ", this library reads " + std::to_string(kCacheVersion));
  Model md;
  model_fields(r, md);
  md.frames.resize(r.len());
  for (Frame& fr : md.frames) { r.str(fr.name); r.str(fr.joint); r.pod(fr.body); r.pod(fr.pos); r.pod(fr.rot); }
  // structural sanity: a cache is trusted input only as far as the sizes agree
  const size_t nb = (size_t)md.nb;
  if (md.nb < 1 || md.parent.size() != nb || md.jtype.size() != nb || md.qidx.size() != nb || md.vidx.size() != nb || md.depth.size() != nb ||
      md.subtree.size() != nb || md.jpos.size() != 3 * nb || md.jrot.size() != 9 * nb || md.axis.size() != 3 * nb || md.mass.size() != nb ||
      md.com.size() != 3 * nb || md.inertia.size() != 6 * nb || md.jlimit.size() != 2 * nb || md.jeffort.size() != nb || md.jvelocity.size() != nb || md.body_names.size() != nb || md.joint_names.size() != nb ||
      md.ctype.size() != md.cbody.size() || md.csize.size() != 3 * md.cbody.size() || md.cpos.size() != 3 * md.cbody.size() || md.crot.size() != 9 * md.cbody.size() ||
      md.pt_coll.size() != md.pt_body.size() || md.pt_feat.size() != md.pt_body.size() || md.pt_pos.size() != 3 * md.pt_body.size() || md.pt_rad.size() != md.pt_body.size() ||
      md.pt_type.size() != md.pt_body.size() || md.pt_pos2.size() != 3 * md.pt_body.size() || md.coll_names.size() != md.cbody.size())
    throw std::runtime_error("model cache: '" + path + "' is inconsistent");
  // derived index tables are recomputed, not trusted (a truncated or edited cache must give a parse error, never an out-of-bounds access)
  {
    int nq = md.floating ? 7 : 0, nv = md.floating ? 6 : 0, maxdepth = 0;
    if (md.floating != 0 && md.floating != 1) throw std::runtime_error("model cache: '" + path + "' is inconsistent (base)");
    std::vector<int> sub(nb, 1);
    for (size_t i = 1; i < nb; i++) {
      if (md.jtype[i] != JT_REVOLUTE && md.jtype[i] != JT_PRISMATIC) throw std::runtime_error("model cache: '" + path + "' is inconsistent (joint types)");
      if (md.qidx[i] != nq++ || md.vidx[i] != nv++ || md.depth[i] != md.depth[md.parent[i]] + 1) throw std::runtime_error("model cache: '" + path + "' is inconsistent (indices)");
      maxdepth = std::max(maxdepth, md.depth[i]);
    }
    for (size_t i = nb - 1; i > 0; i--) sub[md.parent[i]] += sub[i];
    if (md.jtype[0] != (md.floating ? JT_FLOATING : JT_FIXED) || md.depth[0] != 0 || nq != md.nq || nv != md.nv || maxdepth != md.maxdepth || sub != md.subtree)
      throw std::runtime_error("model cache: '" + path + "' is inconsistent (tree)");
  }
  for (size_t k = 0; k < md.pt_body.size(); k++)
    if (md.pt_coll[k] < 0 || md.pt_coll[k] >= md.ncoll() || md.pt_type[k] < FT_POINT || md.pt_type[k] > FT_RIM || md.cbody[md.pt_coll[k]] != md.pt_body[k])
      throw std::runtime_error("model cache: '" + path + "' is inconsistent (candidates)");
  for (int ty : md.ctype) if (ty < CT_SPHERE || ty > CT_CYLINDER) throw std::runtime_error("model cache: '" + path + "' is inconsistent (shapes)");
  for (size_t i = 0; i < nb; i++) if (md.parent[i] >= (int)i || (i > 0 && md.parent[i] < 0)) throw std::runtime_error("model cache: '" + path + "' is inconsistent (parents)");
  for (int b : md.cbody) if (b < 0 || b >= md.nb) throw std::runtime_error("model cache: '" + path + "' is inconsistent (collision bodies)");
  for (int b : md.pt_body) if (b < 0 || b >= md.nb) throw std::runtime_error("model cache: '" + path + "' is inconsistent (points)");
  for (const Frame& fr : md.frames) if (fr.body < 0 || fr.body >= md.nb) throw std::runtime_error("model cache: '" + path + "' is inconsistent (frames)");
  return md;
}

}  // namespace rsb
