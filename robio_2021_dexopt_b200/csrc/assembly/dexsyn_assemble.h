// dexsyn_assemble.h -- objective assembly of the synthetic dexopt problem.
//
// One backend-generic statement of what the reference records onto its tape in
// DexLearn::_runBatch (dexopt/src/dexlearn.h:393-422) for a policy rollout:
//
//   init FK, simulator.init, env init (random object pose)     dexlearn.h:396-403
//   per frame:
//     simulator.step()                                          physics5.h:572-589
//        body centres, gravity, damping, integrate, floating-joint update,
//        joint velocity controllers, forward kinematics
//     joint-limit penalties                                     dexlearn.h:139-157
//     policy network on the environment's observation           dexenv_grasp.h:50-121, neural.h:141-190
//     control: command regularisation, tanh, couplings, scaling dexenv_grasp.h:186-220
//     per end effector: contact points / force / penalties      dexlearn.h:235-391
//   terminal goals (MoveGoal + RelativeOrientationGoal)         goals.h:71-89,173-191
//
// The backend `B` supplies value handle types and primitive operations that
// map 1:1 onto tractor operators.  Two backends exist:
//   * oracle/ref_build/ref_backend.h   -> the reference's own Var<> recording API
//   * csrc/builder (product)           -> the B200 engine's own tape builder
// so the oracle and the product are fed the same problem, and only the
// execution engines differ.
//
// Collision penalties (collision_axes, dexlearn.h:66-131) and friction-cone penalties
// (collision_project_2 normals, dexlearn.h:219-233,324-362) are recorded when the model
// carries registered shapes (Model::collision_pairs, Model::friction_cones) and the
// backend has the two operators -- the product recorder has, the reference-API recorder
// in this image has not (its operators call Bullet: parity unpinned).
// Random draws (object start pose, dropout masks) enter as per-lane parameters instead
// of in-tape RNG ops (H4).
#pragma once
#include <cmath>
#include <stdexcept>
#include <vector>

#include "dexsyn_model.h"

namespace dexsyn {

template <class B> struct Assembler {
  typedef typename B::S S;
  typedef typename B::W W;
  typedef typename B::V3 V3;
  typedef typename B::Q Q;
  typedef typename B::P P;
  typedef typename B::M3 M3;

  B &b;
  const Model &model;
  EnvInfo env;

  // ---- shared across rollouts (recorded once): policy weights ----
  struct Dense {
    int n_in = 0, n_out = 0;
    std::vector<W> weights;  // [n_in x n_out] row-major, as tensor.h:16-30
    std::vector<W> bias;
  };
  std::vector<Dense> layers;
  bool layers_initialised = false;

  // ---- simulator state of one rollout batch (physics5.h:23-101) ----
  struct Body {
    V3 center, center_global, position, linear_momentum, angular_momentum;
    Q orientation;
    S mass, inverse_mass;
    M3 inverse_inertia;
    P inverse_anchor;
  };
  std::vector<P> link_pose, prev_link_pose, first_link_pose;
  std::vector<S> joint_pos;       // revolute joint positions, by joint index
  std::vector<P> joint_origin;
  std::vector<V3> joint_axis;
  V3 float_position;
  Q float_orientation;
  std::vector<S> command;         // controller commands, by joint index
  bool controllers_active = false;
  Body body;
  S time_step;
  V3 gravity;

  Assembler(B &backend, const Model &m, const EnvInfo &e = EnvInfo()) : b(backend), model(m), env(e) {}

  // ------------------------------------------------------------------ policy

  // neural.h:141-190; weights N(0, stdev^2) supplied by the caller's initial x
  void initLayers(const std::vector<double> &x0) {
    if (layers_initialised) return;
    layers_initialised = true;
    int n_in = (int)model.controlled.size() + 3 + 6 + 3;
    int n_out = (int)model.controlled.size() + (int)model.end_effectors.size() * model.contact_dims;
    std::vector<std::pair<int, int>> shapes;
    if (model.policy_hidden > 0) {
      shapes.push_back({n_in, model.policy_hidden});
      shapes.push_back({model.policy_hidden, n_out});
    } else if (model.policy_hidden == 0) {
      shapes.push_back({n_in, n_out});
    }
    size_t k = 0;
    for (auto &sh : shapes) {
      Dense d;
      d.n_in = sh.first;
      d.n_out = sh.second;
      // DenseLayer::evaluate registers weights then bias as variables (neural.h:156-157)
      for (int i = 0; i < d.n_in * d.n_out; i++) d.weights.push_back(b.weight(x0.at(k++)));
      for (int i = 0; i < d.n_out; i++) d.bias.push_back(b.weight(x0.at(k++)));
      // weight / bias regularisation goals, once (neural.h:159-171)
      for (auto &w : d.weights) b.goalW(b.wmul(w, b.cw(env.regularization)));
      for (auto &w : d.bias) b.goalW(b.wmul(w, b.cw(env.regularization)));
      layers.push_back(d);
    }
  }

  static size_t parameterCount(const Model &model) {
    int n_in = (int)model.controlled.size() + 3 + 6 + 3;
    int n_out = (int)model.controlled.size() + (int)model.end_effectors.size() * model.contact_dims;
    if (model.policy_hidden > 0)
      return (size_t)n_in * model.policy_hidden + model.policy_hidden + (size_t)model.policy_hidden * n_out + n_out;
    if (model.policy_hidden == 0) return (size_t)n_in * n_out + n_out;
    return 0;
  }

  std::vector<S> policy(const std::vector<S> &input, int frame) {
    std::vector<S> x = input;
    for (size_t l = 0; l < layers.size(); l++) {
      auto &d = layers[l];
      x = b.dense(x, d.weights, d.bias, d.n_in, d.n_out);
      bool hidden = (l + 1 < layers.size());
      if (hidden) {
        // ActivityRegularizationLayer, DropoutLayer, tanh (dexenv_grasp.h:137-146)
        for (auto &v : x) b.goal(b.mul(v, b.cs(env.regularization)));
        if (model.dropout)
          for (auto &v : x) v = b.mul(v, b.param());  // mask in {0, 1/(1-rate)} per lane
        for (auto &v : x) v = b.tanh(v);
      }
    }
    return x;
  }

  // ------------------------------------------------------------------ kinematics

  // robot/robotmodel.h:128-148, robot/jointtypes.h:143-153,296-301
  void computeFK() {
    for (size_t j = 0; j < model.joints.size(); j++) {
      auto &jt = model.joints[j];
      P parent = (jt.parent >= 0) ? b.pmul(link_pose[jt.parent], joint_origin[j]) : joint_origin[j];
      switch (jt.type) {
        case JOINT_FIXED:
          link_pose[j] = parent;
          break;
        case JOINT_REVOLUTE:
          link_pose[j] = b.prevolute(parent, joint_pos[j], joint_axis[j]);
          break;
        case JOINT_FLOATING:
          link_pose[j] = b.pmul(b.pmul(parent, b.ptrans(float_position)), b.porient(float_orientation));
          break;
      }
    }
  }

  // GeometryFast::inverse(Pose) (geometry/fast.h:143-147)
  P pinv(const P &pose) {
    Q ri = b.qinv(b.porientation(pose));
    V3 ti = b.qmulv(ri, b.vneg(b.ptranslation(pose)));
    return b.pmul(b.ptrans(ti), b.porient(ri));
  }

  // PhysicsSimulator::_velocity (physics5.h:106-112)
  V3 velocity(int link, const V3 &point) {
    V3 moved = b.pmulv(prev_link_pose[link], b.pmulv(pinv(link_pose[link]), point));
    return b.vmuls(b.vsub(point, moved), b.div(b.cs(1.0), time_step));
  }

  // PhysicsSimulator::applyForce -> _applyImpulse (physics5.h:114-121,672-675)
  void applyForce(const V3 &point, const V3 &force) {
    V3 impulse = b.vmuls(force, time_step);
    body.linear_momentum = b.vadd(body.linear_momentum, impulse);
    body.angular_momentum = b.vadd(body.angular_momentum, b.cross(b.vsub(point, body.center_global), impulse));
  }

  // PhysicsSimulator::step (physics5.h:572-589) with contact detection empty
  // (all-allowed ACM, dexlearn.h:424-438)
  void simStep() {
    // _updateBodyPoses
    body.center_global = b.vadd(body.position, b.qmulv(body.orientation, body.center));
    // _applyGravity
    body.linear_momentum = b.vadd(body.linear_momentum, b.vmuls(gravity, b.mul(body.mass, time_step)));
    // _applyDamping (physics5.h:174-193)
    {
      S lin = b.exp(b.mul(b.log(b.cs(env.decay)), time_step));
      S ang = b.exp(b.mul(b.log(b.cs(env.decay)), time_step));
      body.linear_momentum = b.vmuls(body.linear_momentum, lin);
      body.angular_momentum = b.vmuls(body.angular_momentum, ang);
    }
    prev_link_pose = link_pose;
    // _integrate (physics5.h:334-358)
    {
      V3 local_angular_momentum = b.qmulv(b.qinv(body.orientation), body.angular_momentum);
      V3 linear_velocity = b.vmuls(body.linear_momentum, body.inverse_mass);
      body.position = b.vadd(body.position, b.vmuls(linear_velocity, time_step));
      V3 angular_velocity = b.qmulv(body.orientation, b.mmulv(body.inverse_inertia, local_angular_momentum));
      V3 rotation = b.vmuls(angular_velocity, time_step);
      body.position = b.vsub(body.position, b.cross(rotation, b.qmulv(body.orientation, body.center)));
      body.orientation = b.qaddv(body.orientation, rotation);
      body.angular_momentum = b.qmulv(body.orientation, local_angular_momentum);
    }
    // _updateRobotState (physics5.h:360-371)
    {
      P pose = b.pmul(b.pmul(body.inverse_anchor, b.ptrans(body.position)), b.porient(body.orientation));
      float_position = b.ptranslation(pose);
      float_orientation = b.porientation(pose);
    }
    // _applyControllers, velocity mode (physics5.h:373-391)
    if (controllers_active) {
      for (int j : model.controlled) joint_pos[j] = b.add(joint_pos[j], b.mul(command[j], time_step));
    }
    computeFK();
  }

  // ------------------------------------------------------------------ penalties

  // dexlearn.h:139-157
  void jointLimitPenalties() {
    S w = b.cs(env.joint_limit_penalty);
    for (int j : model.controlled) {
      auto &jt = model.joints[j];
      b.goal(b.mul(b.relu(b.sub(joint_pos[j], b.cs(jt.upper))), w));
      b.goal(b.mul(b.relu(b.sub(b.cs(jt.lower), joint_pos[j])), w));
    }
  }

  // dexlearn.h:159-217
  void shapePenalty(int link, const Shape &shape, const V3 &contact_point) {
    V3 p = b.pmulv(pinv(link_pose[link]), contact_point);
    S f = b.cs(env.shape_penalty);
    if (shape.kind == 1) {
      S px, py, pz;
      b.unpack(p, px, py, pz);
      b.goal(b.mul(f, b.add(b.relu(b.sub(pz, b.cs(shape.length * 0.5))),
                            b.relu(b.sub(b.neg(pz), b.cs(shape.length * 0.5))))));
      const int n = 8;
      for (int i = 0; i < n; i++) {
        double angle = i * M_PI / n;
        S v = b.add(b.mul(px, b.cs(std::sin(angle))), b.mul(py, b.cs(std::cos(angle))));
        b.goal(b.mul(f, b.add(b.relu(b.sub(v, b.cs(shape.radius))),
                              b.relu(b.sub(b.neg(v), b.cs(shape.length))))));
      }
      return;
    }
    for (auto &plane : shape.planes) {
      V3 n = b.pack(b.cs(plane.n[0]), b.cs(plane.n[1]), b.cs(plane.n[2]));
      S dist = b.add(b.dot(n, p), b.cs(plane.d));
      b.goal(b.mul(f, b.relu(dist)));
    }
  }

  // dexlearn.h:66-131: one closest-point query per shape pair; the witnesses are re-attached to their links
  // (local_a, local_b are constants for the derivative, the link poses are not)
  void collisionPenalties() {
    if constexpr (B::kHasCollisionOps) collisionPenaltiesImpl();
    else if (!model.collision_pairs.empty()) throw std::runtime_error("this recorder has no collision operators");
  }
  void collisionPenaltiesImpl() {
    for (auto &cp : model.collision_pairs) {
      const P &pose_a = link_pose[cp.link_a], &pose_b = link_pose[cp.link_b];
      V3 point_a, point_b, axis, local_a, local_b;
      b.collisionAxes(pose_a, pose_b, cp.handle, point_a, point_b, axis, local_a, local_b);
      point_a = b.pmulv(pose_a, local_a);
      point_b = b.pmulv(pose_b, local_b);
      S distance = b.dot(axis, b.vsub(point_a, point_b));
      if (env.collision_penalty) b.goal(b.mul(b.relu(b.neg(distance)), b.cs(env.collision_penalty)));
      if (env.collision_avoidance_weight)
        b.goal(b.mul(b.relu(b.add(b.neg(distance), b.cs(env.collision_avoidance_distance))),
                     b.cs(env.collision_avoidance_weight)));
      if (env.slip_avoidance_weight) {
        V3 p_center = b.vmuls(b.vadd(point_a, point_b), b.cs(0.5));
        V3 v1 = velocity(cp.link_a, p_center);
        V3 v2 = velocity(cp.link_b, p_center);
        b.goalV(b.vmuls(b.vsub(v2, v1), b.mul(b.relu(b.add(b.neg(distance), b.cs(env.slip_avoidance_distance))),
                                              b.cs(env.slip_avoidance_weight))));
      }
    }
  }

  // dexlearn.h:219-233
  void frictionConePenalties(const V3 &contact_force, const V3 &contact_normal, const S &weight) {
    for (int sign : {-1, +1})
      for (int axis = 0; axis < 3; axis++) {
        V3 v = b.pack(b.cs(axis == 0), b.cs(axis == 1), b.cs(axis == 2));
        V3 n = b.vadd(contact_normal, b.cross(contact_normal, b.vmuls(v, b.cs(sign * 0.5))));
        b.goal(b.mul(b.relu(b.neg(b.dot(n, contact_force))), weight));
      }
  }

  // dexlearn.h:324-362
  void frictionCones(int obj_link, int ee_link, const Shape &obj_shape, const Shape &ee_shape, const V3 &c1, const V3 &c2,
                     const V3 &force) {
    {
      V3 normal;
      S distance;
      b.collisionProject(b.pmulv(pinv(link_pose[obj_link]), c2), obj_shape.handle, normal, distance);
      normal = b.qmulv(b.porientation(link_pose[obj_link]), normal);
      if (env.friction_cone_penalty) frictionConePenalties(force, normal, b.cs(env.friction_cone_penalty));
    }
    {
      V3 normal;
      S distance;
      b.collisionProject(b.pmulv(pinv(link_pose[ee_link]), c1), ee_shape.handle, normal, distance);
      normal = b.qmulv(b.porientation(link_pose[ee_link]), normal);
      if (env.friction_cone_penalty) frictionConePenalties(b.vneg(force), normal, b.cs(env.friction_cone_penalty));
    }
  }

  // dexlearn.h:235-391
  void evaluateContact(int ee_index, const std::vector<S> &cp) {
    int ee_link = model.end_effectors[ee_index];
    int obj_link = model.object_joint;
    const Shape &ee_shape = model.ee_shapes[ee_index];
    const Shape &obj_shape = model.object_shape;

    V3 c1 = b.vmuls(b.pack(cp[0], cp[1], cp[2]), b.cs(0.1));
    b.goalV(b.vmuls(c1, b.cs(env.contact_point_regularization)));
    c1 = b.vadd(b.pmulv(link_pose[obj_link],
                        b.pack(b.cs(obj_shape.center[0]), b.cs(obj_shape.center[1]), b.cs(obj_shape.center[2]))),
                c1);

    V3 c2 = b.vmuls(b.pack(cp[3], cp[4], cp[5]), b.cs(0.1));
    b.goalV(b.vmuls(c2, b.cs(env.contact_point_regularization)));
    c2 = b.vadd(b.pmulv(link_pose[ee_link],
                        b.pack(b.cs(ee_shape.center[0]), b.cs(ee_shape.center[1]), b.cs(ee_shape.center[2]))),
                c2);

    S force_scale = b.cs(0.1);
    S fx = b.mul(cp[6], force_scale);
    S fy = b.mul(cp[7], force_scale);
    S fz = b.mul(cp[8], force_scale);
    V3 force = b.pack(fx, fy, fz);
    b.goalV(b.vmuls(force, b.cs(env.contact_force_regularization)));

    shapePenalty(obj_link, obj_shape, c1);
    shapePenalty(ee_link, ee_shape, c2);

    // friction cones: surface normal of each shape at the other side's contact point (dexlearn.h:324-362)
    // friction cones: surface normal of each shape at the other side's contact point (dexlearn.h:324-362)
    if (model.friction_cones) {
      if constexpr (B::kHasCollisionOps) frictionCones(obj_link, ee_link, obj_shape, ee_shape, c1, c2, force);
      else throw std::runtime_error("this recorder has no collision operators");
    }
    {
      V3 d = b.vmuls(b.vsub(c2, c1), b.cs(env.contact_distance_penalty));
      b.goalV(b.vmuls(d, fx));
      b.goalV(b.vmuls(d, fy));
      b.goalV(b.vmuls(d, fz));
    }
    {
      V3 p = b.vmuls(b.vadd(c1, c2), b.cs(0.5));
      V3 v1 = velocity(obj_link, p);
      V3 v2 = velocity(ee_link, p);
      V3 dv = b.vmuls(b.vsub(v2, v1), b.cs(env.contact_slip_penalty));
      b.goalV(b.vmuls(dv, fx));
      b.goalV(b.vmuls(dv, fy));
      b.goalV(b.vmuls(dv, fz));
    }
    applyForce(c1, force);
  }

  // ------------------------------------------------------------------ observation / control

  // dexenv_grasp.h:50-121
  std::vector<S> policyInput() {
    std::vector<S> in;
    for (int j : model.controlled) in.push_back(joint_pos[j]);
    P object_pose = link_pose[model.object_joint];
    Q object_orientation = b.porientation(object_pose);
    V3 object_position = b.ptranslation(object_pose);
    S px, py, pz;
    b.unpack(object_position, px, py, pz);
    S pscale = b.cs(10);
    in.push_back(b.mul(px, pscale));
    in.push_back(b.mul(py, pscale));
    in.push_back(b.mul(pz, pscale));
    S a0, a1, a2;
    b.unpack(b.qmulv(object_orientation, b.pack(b.cs(1), b.cs(0), b.cs(0))), a0, a1, a2);
    in.push_back(a0);
    in.push_back(a1);
    in.push_back(a2);
    b.unpack(b.qmulv(object_orientation, b.pack(b.cs(0), b.cs(1), b.cs(0))), a0, a1, a2);
    in.push_back(a0);
    in.push_back(a1);
    in.push_back(a2);
    b.unpack(b.vsub(b.ptranslation(link_pose[model.palm_link]), object_position), a0, a1, a2);
    in.push_back(a0);
    in.push_back(a1);
    in.push_back(a2);
    return in;
  }

  // dexenv_grasp.h:186-220
  void controlRobot(const std::vector<S> &policy_output) {
    std::vector<S> vel = policy_output;
    for (auto &v : vel) {
      b.goal(b.mul(v, b.cs(env.command_regularization)));
      v = b.tanh(v);
    }
    for (int j : model.controlled) {
      auto &jt = model.joints[j];
      if (jt.couple_to >= 0) vel[jt.control] = vel[model.joints[jt.couple_to].control];
    }
    for (int j : model.controlled) {
      auto &jt = model.joints[j];
      command[j] = b.mul(vel[jt.control], b.cs(jt.arm ? 3.0 : 20.0));
    }
    controllers_active = true;
  }

  // ------------------------------------------------------------------ rollout

  // One rollout batch = dexlearn.h:393-422 followed by the terminal goals of
  // dexenv_grasp.h:167-184.  x0 gives the initial policy parameters.
  void rollout(int frames, const std::vector<double> &x0) {
    size_t nj = model.joints.size();
    link_pose.assign(nj, b.cp(Joint().origin));
    joint_pos.assign(nj, b.cs(0));
    command.assign(nj, b.cs(0));
    joint_origin.clear();
    joint_axis.clear();
    for (auto &jt : model.joints) {
      joint_origin.push_back(b.cp(jt.origin));
      joint_axis.push_back(b.cv(jt.axis[0], jt.axis[1], jt.axis[2]));
    }
    for (int j : model.controlled) joint_pos[j] = b.cs(model.joints[j].q0);
    float_position = b.cv(0, 0, 0);
    float_orientation = b.cq(0, 0, 0, 1);
    time_step = b.cs(env.time_step);
    gravity = b.cv(env.gravity[0], env.gravity[1], env.gravity[2]);
    controllers_active = false;

    // trajectory state(0) FK + simulator.init (physics5.h:552-570)
    computeFK();
    prev_link_pose = link_pose;
    body.center = b.cv(model.body_center[0], model.body_center[1], model.body_center[2]);
    body.mass = b.cs(model.body_mass);
    body.inverse_mass = b.cs(1.0 / model.body_mass);
    body.inverse_inertia = b.cm(model.body_inverse_inertia);
    body.inverse_anchor = b.cp(model.inverse_anchor);
    body.linear_momentum = b.cv(0, 0, 0);
    body.angular_momentum = b.cv(0, 0, 0);
    body.position = b.ptranslation(link_pose[model.object_joint]);
    body.orientation = b.porientation(link_pose[model.object_joint]);
    body.center_global = b.cv(0, 0, 0);

    // env->init: random planar offset and yaw of the object (dexenv_grasp.h:149-165);
    // the three draws are per-lane parameters
    {
      S px = b.param(), py = b.param(), yaw = b.param();
      body.position = b.vadd(body.position, b.pack(px, py, b.cs(0)));
      Q rot = b.qaxis(yaw, b.cv(0, 0, 1));
      body.orientation = b.qmul(rot, body.orientation);
    }

    initLayers(x0);

    for (int frame = 0; frame < frames; frame++) {
      b.frameBegin(frame);
      simStep();
      if (frame == 0) first_link_pose = link_pose;
      jointLimitPenalties();
      if (!layers.empty()) {
        std::vector<S> out = policy(policyInput(), frame);
        controlRobot(out);
        size_t nq = model.controlled.size();
        for (size_t e = 0; e < model.end_effectors.size(); e++) {
          std::vector<S> cp(out.begin() + nq + e * model.contact_dims,
                            out.begin() + nq + (e + 1) * model.contact_dims);
          evaluateContact((int)e, cp);
        }
      }
      collisionPenalties();  // dexlearn.h:419
      b.frameEnd(frame);
    }

    // MoveGoal (goals.h:173-191) and RelativeOrientationGoal (goals.h:71-89)
    {
      int o = model.object_joint;
      V3 moved = b.vsub(b.vsub(b.ptranslation(link_pose[o]), b.ptranslation(first_link_pose[o])),
                        b.cv(env.move_goal[0], env.move_goal[1], env.move_goal[2]));
      b.goalV(b.vmuls(moved, b.cs(env.move_goal_weight)));
      Q qa = b.porientation(first_link_pose[o]);
      Q qb = b.porientation(link_pose[o]);
      V3 res = b.qresidual(b.qmul(b.qinv(qa), qb));
      b.goalV(b.vmuls(b.vsub(b.cv(0, 0, 0), res), b.cs(env.orientation_goal_weight)));
    }
  }
};

}  // namespace dexsyn
