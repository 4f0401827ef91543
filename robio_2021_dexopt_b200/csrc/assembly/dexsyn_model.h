// dexsyn_model.h -- synthetic Shadow-hand(+UR-arm) manipulation problems.
//
// The real robot models of the reference are external ROS packages
// (sr_description / ur_description, see dexopt/models/handarm02.xacro:13,37 and
// hand8.xacro:13) and are not available, so every benchmark / parity problem is
// built from a seeded synthetic kinematic tree with the same *shape* as the
// reference's two models:
//   hand8   : 2 wrist joints + fingers 4/4/4/5/5            -> 24 revolute joints
//             (dexopt/models/hand8.srdf:39-71)
//   handarm : 6 serial arm joints + hand8 minus WRJ1        -> 29 revolute joints
//             (dexopt/models/handarm02.srdf:47-85)
//   chain   : serial n-revolute chain (roofline stress config of BASELINE.json)
// plus fixed links, one floating dynamic object and a static floor.
//
// Value distributions follow SURVEY.md section 8(d).  The generator is plain
// integer/floating arithmetic on std::mt19937_64 raw output so the oracle-side
// recorder and the product-side builder see bit-identical models.
#pragma once
#include <cmath>
#include <cstdint>
#include <random>
#include <string>
#include <vector>

namespace dexsyn {

enum JointType { JOINT_FIXED = 0, JOINT_REVOLUTE = 1, JOINT_FLOATING = 2 };

struct Joint {
  int parent = -1;         // index of the parent joint (= parent link); -1: attached to the world
  int type = JOINT_FIXED;
  double origin[7] = {0, 0, 0, 0, 0, 0, 1};  // translation xyz + quaternion xyzw
  double axis[3] = {0, 0, 1};
  double lower = -1.0, upper = 1.0;           // joint limits (revolute)
  double q0 = 0.0;                            // start position (revolute)
  int control = -1;                           // index among the controlled joints, -1 if none
  bool arm = false;                           // "arm_" joints are scaled x3, hand joints x20
  int couple_to = -1;                         // J1 := J2 finger coupling (dexenv_grasp.h:200-210)
  std::string name;
};

struct Plane {
  double n[3];
  double d;
};

struct Shape {
  int kind = 0;  // 0 = convex polyhedron (planes), 1 = z-cylinder
  double center[3] = {0, 0, 0};
  std::vector<Plane> planes;
  std::vector<double> points;  // polyhedron corners xyz (the support mapping of shape.h:214-222)
  double radius = 0, length = 0;
  uint64_t handle = 0;  // in the engine's collision shape table, 0 = not registered (trx_collision_shape_create)
};

// one collision_axes instruction per frame (dexlearn.h:66-131): two links, their shapes, the registered pair
struct CollisionPair {
  int link_a = -1, link_b = -1;
  uint64_t handle = 0;
};

struct Model {
  std::vector<Joint> joints;
  std::vector<int> controlled;  // joint indices in policy order
  int object_joint = -1;        // the floating joint
  int palm_link = -1;
  int floor_link = -1;
  std::vector<int> end_effectors;       // link (= joint) indices; the last one is the floor
  std::vector<Shape> ee_shapes;         // one per end effector
  Shape object_shape;
  // link pairs the allowed-collision matrix does not exempt (dexlearn.h:66-80); here: every finger tip and the
  // floor against the object, and neighbouring finger tips against each other.  Filled by registerShapes().
  std::vector<CollisionPair> collision_pairs;
  bool wants_collision_pairs = false;
  bool floor_pair = true;
  bool friction_cones = false;  // collision_project_2 normals + friction-cone penalties (dexlearn.h:219-233,324-362)
  // dynamic body of the object (physics5.h:78-96)
  double body_center[3] = {0, 0, 0};
  double body_mass = 0.01;
  double body_inverse_inertia[9] = {0};
  double inverse_anchor[7] = {0, 0, 0, 0, 0, 0, 1};
  // policy
  int policy_hidden = 64;  // 0 = single linear layer (turn env), -1 = no policy
  int contact_dims = 9;    // dexlearn.h:17
  bool dropout = false;    // dropout mask supplied as a per-lane parameter
  double dropout_rate = 0.3;
};

// --- deterministic random numbers --------------------------------------------

struct Rng {
  std::mt19937_64 gen;
  explicit Rng(uint64_t seed) : gen(seed) {}
  double uniform() { return (double)(gen() >> 11) * (1.0 / 9007199254740992.0); }
  double uniform(double lo, double hi) { return lo + (hi - lo) * uniform(); }
  double normal() {
    double u1 = uniform(), u2 = uniform();
    if (u1 < 1e-300) u1 = 1e-300;
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
  }
  int integer(int n) { return (int)(gen() % (uint64_t)n); }
};

inline void randomOrigin(Rng &rng, double *o, bool along_chain) {
  o[0] = rng.uniform(-0.05, 0.05) + (along_chain ? 0.03 : 0.0);
  o[1] = rng.uniform(-0.05, 0.05);
  o[2] = rng.uniform(-0.05, 0.05);
  double q[4], n = 0;
  for (int i = 0; i < 4; i++) {
    q[i] = rng.normal();
    n += q[i] * q[i];
  }
  n = std::sqrt(n);
  for (int i = 0; i < 4; i++) o[3 + i] = q[i] / n;
}

inline void randomAxis(Rng &rng, double *a) {
  int k = rng.integer(6);
  a[0] = a[1] = a[2] = 0;
  a[k % 3] = (k < 3) ? 1.0 : -1.0;
}

inline Shape boxShape(double sx, double sy, double sz) {
  Shape s;
  s.kind = 0;
  const double h[3] = {sx * 0.5, sy * 0.5, sz * 0.5};
  for (int a = 0; a < 3; a++)
    for (int sign = -1; sign <= 1; sign += 2) {
      Plane p;
      p.n[0] = p.n[1] = p.n[2] = 0;
      p.n[a] = sign;
      p.d = -h[a];  // dot(n, x) + d <= 0 inside   (dexlearn.h:203-210)
      s.planes.push_back(p);
    }
  for (int i = 0; i < 8; i++) {
    s.points.push_back((i & 1) ? h[0] : -h[0]);
    s.points.push_back((i & 2) ? h[1] : -h[1]);
    s.points.push_back((i & 4) ? h[2] : -h[2]);
  }
  return s;
}

inline Shape cylinderShape(double radius, double length) {
  Shape s;
  s.kind = 1;
  s.radius = radius;
  s.length = length;
  return s;
}

struct ModelOptions {
  std::string kind = "handarm";  // handarm | hand8 | chain
  int chain_joints = 30;         // kind == chain
  int chain_contacts = 16;       // kind == chain: number of end effectors
  int extra_fixed = 12;          // additional fixed links hung into the tree
  int policy_hidden = 64;
  bool dropout = false;
  bool collision = false;  // collision penalties between registered link shapes (dexlearn.h:66-131)
  bool friction = false;   // friction-cone penalties from collision_project_2 normals (dexlearn.h:324-362)
  bool floor_pair = true;  // collision=1: also test the floor against the object.  The object starts resting on the
                           // floor, face on face: the closest points of two parallel faces are not unique, so two
                           // correct implementations may report different witnesses (tests use 0 for exact parity)
  uint64_t seed = 20211227;
};

inline int addJoint(Model &m, Rng &rng, int parent, int type, const std::string &name, bool arm,
                    bool along_chain = true) {
  Joint j;
  j.parent = parent;
  j.type = type;
  j.name = name;
  j.arm = arm;
  randomOrigin(rng, j.origin, along_chain);
  if (type == JOINT_REVOLUTE) {
    randomAxis(rng, j.axis);
    j.lower = -1.0;
    j.upper = 1.0;
    j.q0 = rng.uniform(-0.5, 0.5);
    j.control = (int)m.controlled.size();
    m.controlled.push_back((int)m.joints.size());
  }
  m.joints.push_back(j);
  return (int)m.joints.size() - 1;
}

inline Model makeModel(const ModelOptions &opt) {
  Model m;
  Rng rng(opt.seed);
  m.policy_hidden = opt.policy_hidden;
  m.dropout = opt.dropout;
  m.friction_cones = opt.friction;
  m.wants_collision_pairs = opt.collision;
  m.floor_pair = opt.floor_pair;

  int base = addJoint(m, rng, -1, JOINT_FIXED, "base", false, false);
  std::vector<int> tips;
  int palm = base;

  if (opt.kind == "chain") {
    int prev = base;
    int n = opt.chain_joints;
    for (int i = 0; i < n; i++) {
      prev = addJoint(m, rng, prev, JOINT_REVOLUTE, "chain_" + std::to_string(i), i < 6);
      // contact frames are spread along the chain, nearest the tip first
      if (n - 1 - i < opt.chain_contacts - 1) tips.push_back(prev);
    }
    palm = m.controlled[std::min<size_t>(5, m.controlled.size() - 1)];
  } else {
    int prev = base;
    if (opt.kind == "handarm") {
      for (int i = 0; i < 6; i++) prev = addJoint(m, rng, prev, JOINT_REVOLUTE, "arm_" + std::to_string(i), true);
      prev = addJoint(m, rng, prev, JOINT_REVOLUTE, "WRJ2", false);
    } else {
      prev = addJoint(m, rng, prev, JOINT_REVOLUTE, "WRJ2", false);
      prev = addJoint(m, rng, prev, JOINT_REVOLUTE, "WRJ1", false);
    }
    palm = addJoint(m, rng, prev, JOINT_FIXED, "palm", false);
    const char *fingers[5] = {"FF", "MF", "TH", "RF", "LF"};  // end-effector order of dexenv_grasp.h:45-48
    const int counts[5] = {4, 4, 5, 4, 5};
    for (int f = 0; f < 5; f++) {
      int p = palm;
      std::vector<int> chain;
      for (int k = counts[f]; k >= 1; k--) {
        p = addJoint(m, rng, p, JOINT_REVOLUTE, std::string(fingers[f]) + "J" + std::to_string(k), false);
        chain.push_back(p);
      }
      // J1 := J2 coupling for FF/MF/RF/LF (dexenv_grasp.h:200-210)
      if (f != 2) {
        int j1 = chain[chain.size() - 1], j2 = chain[chain.size() - 2];
        m.joints[j1].couple_to = j2;
      }
      int tip = addJoint(m, rng, p, JOINT_FIXED, std::string(fingers[f]) + "distal", false);
      tips.push_back(tip);
    }
  }
  m.palm_link = palm;

  // fixed links hanging off random existing links (sensor / cover frames of a real URDF)
  for (int i = 0; i < opt.extra_fixed; i++) {
    int parent = rng.integer((int)m.joints.size());
    addJoint(m, rng, parent, JOINT_FIXED, "fixed_" + std::to_string(i), false, false);
  }

  // floor: static box attached to the world
  {
    Joint j;
    j.parent = -1;
    j.type = JOINT_FIXED;
    j.name = "floor";
    j.origin[2] = -0.05;
    m.joints.push_back(j);
    m.floor_link = (int)m.joints.size() - 1;
  }

  // object: floating joint attached to the world (hand8.xacro:24-26, handarm02.xacro:73-75)
  {
    Joint j;
    j.parent = -1;
    j.type = JOINT_FLOATING;
    j.name = "object";
    j.origin[0] = 0.25;
    j.origin[1] = 0.02;
    j.origin[2] = 0.03;
    m.joints.push_back(j);
    m.object_joint = (int)m.joints.size() - 1;
    for (int i = 0; i < 7; i++) m.inverse_anchor[i] = 0;
    m.inverse_anchor[0] = -j.origin[0];
    m.inverse_anchor[1] = -j.origin[1];
    m.inverse_anchor[2] = -j.origin[2];
    m.inverse_anchor[6] = 1;
  }

  bool cyl = (opt.kind == "hand8");
  if (cyl) {
    m.object_shape = cylinderShape(0.04, 0.06);
    m.body_mass = 0.1;
  } else {
    m.object_shape = boxShape(0.06, 0.06, 0.06);
    m.body_mass = 0.01;
  }
  {
    // solid box / cylinder inertia, inverted (diagonal)
    double ixx, iyy, izz;
    if (cyl) {
      double r = 0.04, l = 0.06;
      ixx = iyy = m.body_mass * (3 * r * r + l * l) / 12.0;
      izz = 0.5 * m.body_mass * r * r;
    } else {
      double s = 0.06;
      ixx = iyy = izz = m.body_mass * (s * s + s * s) / 12.0;
    }
    for (int i = 0; i < 9; i++) m.body_inverse_inertia[i] = 0;
    m.body_inverse_inertia[0] = 1.0 / ixx;
    m.body_inverse_inertia[4] = 1.0 / iyy;
    m.body_inverse_inertia[8] = 1.0 / izz;
    m.body_center[0] = 0.001;
    m.body_center[1] = -0.002;
    m.body_center[2] = 0.0015;
  }

  for (int t : tips) {
    m.end_effectors.push_back(t);
    Shape s = boxShape(0.02, 0.015, 0.025);
    s.center[0] = 0.002;
    s.center[1] = 0.0;
    s.center[2] = 0.01;
    m.ee_shapes.push_back(s);
  }
  m.end_effectors.push_back(m.floor_link);
  {
    Shape s = boxShape(1.0, 1.0, 0.1);
    m.ee_shapes.push_back(s);
  }
  return m;
}

// Environment weights (dexopt/src/dexenv_grasp.h:19-48, dexenv.h:57-73)
struct EnvInfo {
  double joint_limit_penalty = 0.5;
  double shape_penalty = 1;
  double contact_point_regularization = 0.1;
  double contact_force_regularization = 0.1;
  double contact_distance_penalty = 5;
  double contact_slip_penalty = 0.1;
  double regularization = 0.01;  // dexenv_grasp.h:133 (weights, biases, activity)
  double command_regularization = 0.01;  // dexenv_grasp.h:195
  double move_goal[3] = {0.0, 0.0, 0.1};
  double move_goal_weight = 2;
  double orientation_goal_weight = 1;
  double time_step = 0.01;   // physics5.h:32
  double gravity[3] = {0, 0, -9.81};
  double decay = 0.01;       // physics5.h:178-186
  // dexenv_grasp.h:25-31,43 (used when the model carries collision pairs / friction cones)
  double collision_penalty = 1;
  double collision_avoidance_distance = 0.01;
  double collision_avoidance_weight = 0.02;
  double slip_avoidance_distance = 0.01;
  double slip_avoidance_weight = 0.1;
  double friction_cone_penalty = 1;
};

}  // namespace dexsyn
