// test_shim.cpp -- the B200 shim (shim/: sortham::Optimizer, CriticManager and the 12 critic plugins over libmppi_b200.so)
// configured from the deployed YAML and driven like SORTHAMController::computeVelocityCommands drives the reference.
//
//   ref: robot_bringup/config/nav2_params.yaml:115,184-293 (controller_frequency, FollowPath block incl. the keys no code
//        reads: vy_min, wz_min, ax_*, ay_*, az_*, PathAngleCritic.mode, TwirlingCritic.twirling_cost_*,
//        CostCritic.trajectory_point_step), :300-312 (local costmap: 3 m x 3 m @ 0.05, robot_radius, inflation layer)
//   ref: nav2_sortham_controller/src/controller.cpp:27-116, optimizer.cpp:35-155, critic_manager.cpp:36-65
//
// Modes:  test_shim describe   configuration only (no device): parameters declared / read, critic table, robot description
//         test_shim cycles     + evalControl cycles against the CPU oracle (needs a GPU): twist and control sequence within
//                              1e-4, dynamic parameter changes, speed limit, handle re-creation, fallback throw
// The oracle is driven through its own C ABI (oracle_*) with a critic table written out by hand below, i.e. NOT produced by
// the shim's describe(): a parameter the shim mis-reads shows up as a mismatch.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "nav2_sortham_controller/optimizer.hpp"

extern "C" {
void oracle_config_default(mppi_config *);
void oracle_critic_default(int32_t, mppi_critic_desc *);
int oracle_create(const mppi_config *, mppi_handle **);
void oracle_destroy(mppi_handle *);
int oracle_reset(mppi_handle *);
int oracle_set_critics(mppi_handle *, const mppi_critic_desc *, int32_t);
int oracle_set_robot(mppi_handle *, const mppi_robot_desc *);
int oracle_set_noise(mppi_handle *, const float *, const float *, const float *);
int oracle_set_speed_limit(mppi_handle *, double, int32_t);
int oracle_get_control_sequence(mppi_handle *, float *, float *, float *);
int oracle_eval_control(mppi_handle *, const mppi_cycle_in *, int32_t, mppi_cycle_out *, float[3]);
}

static int g_failures = 0;
#define EXPECT_TRUE(c) do {if (!(c)) {std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_failures;}} while (0)
#define EXPECT_EQ(a, b) do {const auto _a = (a); const auto _b = (b); if (!(_a == _b)) { \
  std::printf("FAIL %s:%d: %s != %s (%g vs %g)\n", __FILE__, __LINE__, #a, #b, static_cast<double>(_a), static_cast<double>(_b)); ++g_failures;}} while (0)
#define EXPECT_NEAR(a, b, tol) do {const double _a = (a), _b = (b); if (!(std::fabs(_a - _b) <= (tol))) { \
  std::printf("FAIL %s:%d: %s = %.9g vs %s = %.9g (tol %g)\n", __FILE__, __LINE__, #a, _a, #b, _b, static_cast<double>(tol)); ++g_failures;}} while (0)

using rclcpp::ParameterValue;
using Node = rclcpp_lifecycle::LifecycleNode;

// ---- nav2_params.yaml:115 and :184-293, key by key (ParameterValue types as the YAML parser would produce them) -----------
static void load_yaml_overrides(Node & n)
{
  n.set_override("controller_frequency", 30.0);
  const std::string p = "FollowPath.";
  n.set_override(p + "time_steps", 56); n.set_override(p + "model_dt", 0.05); n.set_override(p + "batch_size", 2000);
  n.set_override(p + "vx_std", 0.2); n.set_override(p + "vy_std", 0.2); n.set_override(p + "wz_std", 0.2);
  n.set_override(p + "vx_max", 0.5); n.set_override(p + "vx_min", -0.5); n.set_override(p + "vy_max", 0.5);
  n.set_override(p + "vy_min", -0.5);                                             // dead key
  n.set_override(p + "wz_max", 1.0); n.set_override(p + "wz_min", -1.0);            // wz_min: dead key
  for (const char * k : {"ax_max", "ay_max", "az_max"}) {n.set_override(p + k, 3.0);}     // dead keys
  for (const char * k : {"ax_min", "ay_min", "az_min"}) {n.set_override(p + k, -3.0);}    // dead keys
  n.set_override(p + "iteration_count", 1); n.set_override(p + "prune_distance", 1.7); n.set_override(p + "transform_tolerance", 0.1);
  n.set_override(p + "temperature", 0.3); n.set_override(p + "gamma", 0.015); n.set_override(p + "motion_model", "Omni");
  n.set_override(p + "visualize", false); n.set_override(p + "reset_period", 1.0); n.set_override(p + "regenerate_noises", false);
  n.set_override(p + "TrajectoryVisualizer.trajectory_step", 5); n.set_override(p + "TrajectoryVisualizer.time_step", 3);
  n.set_override(p + "AckermannConstraints.min_turning_r", 0.2);
  n.set_override(p + "critics", std::vector<std::string>{"ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic",
      "PathAlignCritic", "PathFollowCritic", "PathAngleCritic", "PreferForwardCritic", "TwirlingCritic"});
  auto crit = [&](const std::string & c, const std::string & k, ParameterValue v) {n.set_override(p + c + "." + k, std::move(v));};
  crit("ConstraintCritic", "enabled", true); crit("ConstraintCritic", "cost_power", 1); crit("ConstraintCritic", "cost_weight", 4.0);
  crit("GoalCritic", "enabled", true); crit("GoalCritic", "cost_power", 1); crit("GoalCritic", "cost_weight", 5.0);
  crit("GoalCritic", "threshold_to_consider", 1.4);
  crit("GoalAngleCritic", "enabled", true); crit("GoalAngleCritic", "cost_power", 1); crit("GoalAngleCritic", "cost_weight", 3.0);
  crit("GoalAngleCritic", "threshold_to_consider", 0.5);
  crit("PreferForwardCritic", "enabled", true); crit("PreferForwardCritic", "cost_power", 1); crit("PreferForwardCritic", "cost_weight", 5.0);
  crit("PreferForwardCritic", "threshold_to_consider", 0.5);
  // ObstaclesCritic has a block in the YAML but is not in the critics list: never loaded, its keys stay undeclared
  crit("ObstaclesCritic", "enabled", true); crit("ObstaclesCritic", "repulsion_weight", 1.5); crit("ObstaclesCritic", "critical_weight", 20.0);
  crit("CostCritic", "enabled", true); crit("CostCritic", "cost_power", 1); crit("CostCritic", "cost_weight", 3.81);
  crit("CostCritic", "critical_cost", 300.0); crit("CostCritic", "consider_footprint", true); crit("CostCritic", "collision_cost", 1000000.0);
  crit("CostCritic", "near_goal_distance", 1.0); crit("CostCritic", "trajectory_point_step", 2);   // last one: dead key
  crit("PathAlignCritic", "enabled", true); crit("PathAlignCritic", "cost_power", 1); crit("PathAlignCritic", "cost_weight", 14.0);
  crit("PathAlignCritic", "max_path_occupancy_ratio", 0.05); crit("PathAlignCritic", "trajectory_point_step", 4);
  crit("PathAlignCritic", "threshold_to_consider", 0.5); crit("PathAlignCritic", "offset_from_furthest", 20);
  crit("PathAlignCritic", "use_path_orientations", false);
  crit("PathFollowCritic", "enabled", true); crit("PathFollowCritic", "cost_power", 1); crit("PathFollowCritic", "cost_weight", 5.0);
  crit("PathFollowCritic", "offset_from_furthest", 5); crit("PathFollowCritic", "threshold_to_consider", 1.4);
  crit("PathAngleCritic", "enabled", true); crit("PathAngleCritic", "cost_power", 1); crit("PathAngleCritic", "cost_weight", 2.0);
  crit("PathAngleCritic", "offset_from_furthest", 4); crit("PathAngleCritic", "threshold_to_consider", 0.5);
  crit("PathAngleCritic", "max_angle_to_furthest", 1.0); crit("PathAngleCritic", "mode", 0);       // mode: dead key
  crit("TwirlingCritic", "enabled", true); crit("TwirlingCritic", "twirling_cost_power", 5);         // dead keys: Twirling runs
  crit("TwirlingCritic", "twirling_cost_weight", 30.0);                                              // with its defaults 1 / 10.0
}

// ---- what the reference would configure from that YAML, written out by hand (the oracle's input) ----------------------------
static mppi_critic_desc desc(int kind, float weight)
{
  mppi_critic_desc d;
  oracle_critic_default(kind, &d);
  d.kind = kind; d.enabled = 1; d.cost_power = 1; d.cost_weight = weight;
  return d;
}
static std::vector<mppi_critic_desc> expected_critics(float path_align_weight = 14.0f)
{
  std::vector<mppi_critic_desc> v;
  v.push_back(desc(MPPI_CRITIC_CONSTRAINT, 4.0f));
  {auto d = desc(MPPI_CRITIC_COST, 3.81f); d.critical_cost = 300.0f; d.consider_footprint = 1; d.collision_cost = 1000000.0f; d.near_goal_distance = 1.0f; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_GOAL, 5.0f); d.threshold_to_consider = 1.4f; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_GOAL_ANGLE, 3.0f); d.threshold_to_consider = 0.5f; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_PATH_ALIGN, path_align_weight); d.max_path_occupancy_ratio = 0.05f; d.trajectory_point_step = 4; d.threshold_to_consider = 0.5f;
    d.offset_from_furthest = 20; d.use_path_orientations = 0; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_PATH_FOLLOW, 5.0f); d.offset_from_furthest = 5; d.threshold_to_consider = 1.4f; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_PATH_ANGLE, 2.0f); d.offset_from_furthest = 4; d.threshold_to_consider = 0.5f; d.max_angle_to_furthest = 1.0f; d.forward_preference = 1; v.push_back(d);}
  {auto d = desc(MPPI_CRITIC_PREFER_FORWARD, 5.0f); d.threshold_to_consider = 0.5f; v.push_back(d);}
  v.push_back(desc(MPPI_CRITIC_TWIRLING, 10.0f));
  return v;
}
static mppi_config expected_config(int batch = 2000)
{
  mppi_config c;
  oracle_config_default(&c);
  c.batch_size = batch; c.time_steps = 56; c.iteration_count = 1; c.model_dt = 0.05f; c.temperature = 0.3f; c.gamma = 0.015f;
  c.vx_max = 0.5f; c.vx_min = -0.5f; c.vy_max = 0.5f; c.wz_max = 1.0f; c.vx_std = 0.2f; c.vy_std = 0.2f; c.wz_std = 0.2f;
  c.motion_model = MPPI_MODEL_OMNI; c.regenerate_noises = 0; c.ackermann_min_turning_r = 0.2f;
  return c;
}

// ---- the world: nav2_params.yaml:300-312 local costmap (60 x 60 @ 0.05), a wall segment ahead-left of the robot ----------------
constexpr double kRes = 0.05, kRobotRadius = 0.22;   // (the YAML's robot_radius 0.5 fills a 3 m map; a smaller robot keeps free space)
struct World
{
  std::shared_ptr<nav2_costmap_2d::Costmap2D> costmap;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> ros;
  explicit World(bool blocked = false)
  {
    costmap = std::make_shared<nav2_costmap_2d::Costmap2D>(60, 60, kRes, 0.0, 0.0);
    std::vector<geometry_msgs::msg::Point> fp;               // nav2_costmap_2d::makeFootprintFromRadius: 16 points
    for (int i = 0; i < 16; ++i) {
      geometry_msgs::msg::Point pt;
      const double a = i * 2.0 * M_PI / 16;
      pt.x = std::cos(a) * kRobotRadius; pt.y = std::sin(a) * kRobotRadius;
      fp.push_back(pt);
    }
    ros = std::make_shared<nav2_costmap_2d::Costmap2DROS>(costmap, fp);
    auto infl = std::make_shared<nav2_costmap_2d::InflationLayer>(0.55, 3.0, ros->getLayeredCostmap()->getInscribedRadius(), kRes);
    ros->getLayeredCostmap()->getPlugins()->push_back(std::make_shared<nav2_costmap_2d::Layer>());   // e.g. the voxel layer
    ros->getLayeredCostmap()->getPlugins()->push_back(infl);
    std::vector<std::pair<int, int>> lethal;
    if (blocked) {
      for (int y = 0; y < 60; ++y) {for (int x = 0; x < 60; ++x) {lethal.emplace_back(x, y);}}
    } else {
      for (int y = 42; y < 46; ++y) {for (int x = 34; x < 50; ++x) {lethal.emplace_back(x, y);}}
      for (int y = 14; y < 17; ++y) {for (int x = 40; x < 44; ++x) {lethal.emplace_back(x, y);}}
    }
    for (int y = 0; y < 60; ++y) {
      for (int x = 0; x < 60; ++x) {
        double best = 1e9;
        for (auto & l : lethal) {best = std::min(best, std::hypot(double(x - l.first), double(y - l.second)));}
        if (best * kRes <= 0.55) {costmap->setCost(x, y, infl->computeCost(best));}
      }
    }
  }
};

struct FixedGoalChecker : nav2_core::GoalChecker
{
  bool getTolerances(geometry_msgs::msg::Pose & pose_tolerance, geometry_msgs::msg::Twist &) override
  {
    pose_tolerance.position.x = pose_tolerance.position.y = 0.25;
    return true;
  }
};

static geometry_msgs::msg::Quaternion yaw_quat(double yaw)
{
  geometry_msgs::msg::Quaternion q;
  q.z = std::sin(yaw / 2); q.w = std::cos(yaw / 2);
  return q;
}

struct Scene
{
  geometry_msgs::msg::PoseStamped pose;
  geometry_msgs::msg::Twist speed;
  nav_msgs::msg::Path plan;
  geometry_msgs::msg::Pose goal;
  std::vector<float> px, py, pyaw;
  Scene(double x, double y, double yaw, double vx)
  {
    pose.pose.position.x = x; pose.pose.position.y = y; pose.pose.orientation = yaw_quat(yaw);
    speed.linear.x = vx; speed.linear.y = 0.01; speed.angular.z = -0.02;
    for (int i = 0; i < 34; ++i) {                 // prune_distance 1.7 m at 0.05 m spacing, gently curving to the right
      geometry_msgs::msg::PoseStamped p;
      const double s = i * 0.05, h = -0.25 * s;
      p.pose.position.x = x + s * std::cos(h / 2); p.pose.position.y = y + s * std::sin(h / 2);
      p.pose.orientation = yaw_quat(h);
      plan.poses.push_back(p);
      px.push_back(static_cast<float>(p.pose.position.x)); py.push_back(static_cast<float>(p.pose.position.y));
      pyaw.push_back(static_cast<float>(tf2::getYaw(p.pose.orientation)));
    }
    goal = plan.poses.back().pose;
  }
  mppi_cycle_in oracle_in(nav2_costmap_2d::Costmap2D & cm) const
  {
    mppi_cycle_in in{};
    in.pose_x = pose.pose.position.x; in.pose_y = pose.pose.position.y; in.pose_yaw = tf2::getYaw(pose.pose.orientation);
    in.speed_vx = speed.linear.x; in.speed_vy = speed.linear.y; in.speed_wz = speed.angular.z;
    in.goal_x = goal.position.x; in.goal_y = goal.position.y; in.goal_checker_xy_tolerance = 0.25;
    in.path_size = static_cast<int32_t>(px.size()); in.path_x = px.data(); in.path_y = py.data(); in.path_yaw = pyaw.data();
    in.costmap.cells = cm.getCharMap(); in.costmap.size_x = cm.getSizeInCellsX(); in.costmap.size_y = cm.getSizeInCellsY();
    in.costmap.resolution = cm.getResolution(); in.costmap.origin_x = cm.getOriginX(); in.costmap.origin_y = cm.getOriginY();
    return in;
  }
};

static bool same_desc(const mppi_critic_desc & a, const mppi_critic_desc & b) {return std::memcmp(&a, &b, sizeof(a)) == 0;}

// ---- 1. configuration -------------------------------------------------------------------------------------------------------
static void test_describe()
{
  auto node = std::make_shared<Node>("controller_server");
  load_yaml_overrides(*node);
  World world;
  sortham::ParametersHandler handler(node);
  sortham::CriticManager manager;
  manager.on_configure(node, "FollowPath", world.ros, &handler);
  const auto table = manager.describe();
  const auto want = expected_critics();
  EXPECT_EQ(table.size(), want.size());
  for (size_t q = 0; q < std::min(table.size(), want.size()); ++q) {
    if (!same_desc(table[q], want[q])) {std::printf("FAIL critic %zu (kind %d) differs from the hand-written descriptor\n", q, want[q].kind); ++g_failures;}
  }
  // every critic of critics.xml loads by its plugin name and describes itself with the reference's defaults
  pluginlib::ClassLoader<sortham::critics::CriticFunction> loader("nav2_sortham_controller", "sortham::critics::CriticFunction");
  EXPECT_EQ(loader.getDeclaredClasses().size(), size_t(12));
  const char * names[12] = {"ConstraintCritic", "CostCritic", "GoalCritic", "GoalAngleCritic", "ObstaclesCritic", "PathAlignCritic",
    "PathAlignLegacyCritic", "PathAngleCritic", "PathFollowCritic", "PreferForwardCritic", "TwirlingCritic", "VelocityDeadbandCritic"};
  auto bare = std::make_shared<Node>("bare");       // no overrides: code defaults
  sortham::ParametersHandler bare_handler(bare);
  for (int k = 0; k < 12; ++k) {
    std::unique_ptr<sortham::critics::CriticFunction> c(loader.createUnmanagedInstance(std::string("sortham::critics::") + names[k]));
    c->on_configure(bare, "FollowPath", std::string("FollowPath.") + names[k], world.ros, &bare_handler);
    mppi_critic_desc got, def;
    c->describe(got);
    mppi_critic_default(k, &def);
    def.kind = k;
    if (!same_desc(got, def)) {std::printf("FAIL %s: describe() with code defaults differs from mppi_critic_default\n", names[k]); ++g_failures;}
    EXPECT_TRUE(bare->has_parameter(std::string("FollowPath.") + names[k] + ".enabled"));
  }
  // the obstacle critic's own inflation parameters are declared only because an inflation layer exists (obstacles_critic.cpp:78-80)
  EXPECT_TRUE(bare->has_parameter("FollowPath.ObstaclesCritic.cost_scaling_factor"));
  // dead keys are never declared (nobody reads them), live ones are
  for (const char * k : {"FollowPath.vy_min", "FollowPath.wz_min", "FollowPath.ax_max", "FollowPath.az_min", "FollowPath.PathAngleCritic.mode",
      "FollowPath.TwirlingCritic.twirling_cost_weight", "FollowPath.CostCritic.trajectory_point_step", "FollowPath.ObstaclesCritic.enabled"}) {
    EXPECT_TRUE(!node->has_parameter(k));
  }
  EXPECT_TRUE(node->has_parameter("FollowPath.PathAngleCritic.forward_preference"));
  const auto robot = manager.describeRobot();
  EXPECT_EQ(robot.footprint_size, 16);
  EXPECT_NEAR(robot.inscribed_radius, kRobotRadius, 1e-12);
  EXPECT_NEAR(robot.circumscribed_radius, kRobotRadius, 1e-12);
  EXPECT_EQ(robot.inflation_layer_found, 1);
  EXPECT_NEAR(robot.inflation_cost_scaling_factor, 3.0, 0.0);
  EXPECT_EQ(robot.track_unknown, 0);
}

// ---- 2. cycles against the oracle -------------------------------------------------------------------------------------------
struct OracleSide
{
  mppi_handle * h{nullptr};
  std::vector<float> vx, vy, wz;
  OracleSide(const mppi_config & cfg, const std::vector<mppi_critic_desc> & critics, const mppi_robot_desc & robot)
  {
    if (oracle_create(&cfg, &h) != MPPI_OK) {throw std::runtime_error("oracle_create");}
    oracle_set_robot(h, &robot);
    oracle_set_critics(h, critics.data(), static_cast<int32_t>(critics.size()));
    vx.resize(cfg.time_steps); vy = vx; wz = vx;
  }
  ~OracleSide() {if (h) {oracle_destroy(h);}}
  void take_noise_of(sortham::Optimizer & opt)
  {
    const auto & c = opt.settings().base;
    std::vector<float> a(size_t(c.batch_size) * c.time_steps), b(a.size()), d(a.size());
    if (mppi_get_noise(opt.core().handle(), a.data(), b.data(), d.data()) != MPPI_OK) {throw std::runtime_error("mppi_get_noise");}
    oracle_set_noise(h, a.data(), b.data(), d.data());
  }
};

static void compare_cycle(sortham::Optimizer & opt, OracleSide & o, const Scene & sc, World & world, bool shift, const char * label)
{
  FixedGoalChecker checker;
  const auto twist = opt.evalControl(sc.pose, sc.speed, sc.plan, sc.goal, &checker);
  mppi_cycle_in in = sc.oracle_in(*world.costmap);
  mppi_cycle_out out{};
  out.control_vx = o.vx.data(); out.control_vy = o.vy.data(); out.control_wz = o.wz.data();
  float cmd[3];
  if (oracle_eval_control(o.h, &in, shift ? 1 : 0, &out, cmd) != MPPI_OK) {throw std::runtime_error("oracle_eval_control");}
  const double tol = 1e-6;
  auto near = [&](double a, double b, const char * what) {
      if (!(std::fabs(a - b) <= tol + 1e-4 * std::fabs(b))) {std::printf("FAIL %s: %s %.9g vs oracle %.9g\n", label, what, a, b); ++g_failures;}
    };
  near(twist.twist.linear.x, cmd[0], "cmd vx"); near(twist.twist.linear.y, cmd[1], "cmd vy"); near(twist.twist.angular.z, cmd[2], "cmd wz");
  EXPECT_EQ(opt.core().lastCycle().fail_flag, out.fail_flag);
  EXPECT_EQ(opt.core().lastCycle().furthest_reached_path_point, out.furthest_reached_path_point);
  const int T = opt.settings().base.time_steps;
  std::vector<float> gx(T), gy(T), gw(T), ox(T), oy(T), ow(T);
  mppi_get_control_sequence(opt.core().handle(), gx.data(), gy.data(), gw.data());
  oracle_get_control_sequence(o.h, ox.data(), oy.data(), ow.data());
  for (int t = 0; t < T; ++t) {near(gx[t], ox[t], "sequence vx"); near(gy[t], oy[t], "sequence vy"); near(gw[t], ow[t], "sequence wz");}
}

static void test_cycles()
{
  auto node = std::make_shared<Node>("controller_server");
  load_yaml_overrides(*node);
  World world;
  sortham::ParametersHandler handler(node);
  handler.start();                                   // SORTHAMController::activate (controller.cpp:63-68)
  sortham::Optimizer opt;
  opt.initialize(node, "FollowPath", world.ros, &handler);
  // controller_frequency 30 Hz against model_dt 0.05: "period less than model dt" -> no shifting (optimizer.cpp:100-103)
  EXPECT_TRUE(!opt.core().shiftControlSequenceEnabled());
  const mppi_config got = opt.settings().base, want = expected_config();
  EXPECT_EQ(got.batch_size, want.batch_size); EXPECT_EQ(got.time_steps, want.time_steps); EXPECT_EQ(got.motion_model, want.motion_model);
  EXPECT_EQ(got.vx_min, want.vx_min); EXPECT_EQ(got.wz_max, want.wz_max); EXPECT_EQ(got.wz_std, want.wz_std); EXPECT_EQ(got.gamma, want.gamma);
  const auto robot = opt.criticManager().describeRobot();
  {
    OracleSide o(want, expected_critics(), robot);
    o.take_noise_of(opt);
    for (int k = 0; k < 6; ++k) {
      Scene sc(1.0 + 0.02 * k, 1.5 - 0.004 * k, -0.01 * k, 0.05 * k);
      compare_cycle(opt, o, sc, world, false, "yaml config");
    }
    // the visualisation feed (controller.cpp:118-123): only materialised when "visualize" is true (the YAML says false)
    bool threw = false;
    try {opt.getGeneratedTrajectories();} catch (const std::runtime_error &) {threw = true;}
    EXPECT_TRUE(threw);
    EXPECT_EQ(opt.getOptimizedTrajectory().size(), size_t(56) * 3);
    // setSpeedLimit (controller.cpp:130-133 -> optimizer.cpp:428-453): 50 % of the base constraints
    opt.setSpeedLimit(50.0, true);
    oracle_set_speed_limit(o.h, 50.0, 1);
    float lim[4];
    mppi_get_constraints(opt.core().handle(), lim);
    EXPECT_NEAR(lim[0], 0.25, 1e-7); EXPECT_NEAR(lim[1], -0.25, 1e-7); EXPECT_NEAR(lim[3], 0.5, 1e-7);
    compare_cycle(opt, o, Scene(1.12, 1.48, -0.05, 0.3), world, false, "speed limit");
    // ---- dynamic parameter (ros2 param set): lands in the critic's member, the post-callback re-describes and resets
    const size_t creates = opt.reconfigureCount();
    node->set_parameter(rclcpp::Parameter("FollowPath.PathAlignCritic.cost_weight", ParameterValue(9.0)));
    EXPECT_EQ(opt.reconfigureCount(), creates);       // no create-time setting changed: same handle, new critic table, reset
    const auto ec = expected_critics(9.0f);
    oracle_set_critics(o.h, ec.data(), static_cast<int32_t>(ec.size()));
    oracle_reset(o.h);
    o.take_noise_of(opt);                               // reset() redraws the noise (optimizer.cpp:130)
    for (int k = 0; k < 3; ++k) {compare_cycle(opt, o, Scene(1.0 + 0.03 * k, 1.5, 0.0, 0.1 * k), world, false, "after cost_weight change");}
  }
  // ---- a create-time setting changes: the handle is re-created behind the same Optimizer object
  {
    const size_t creates = opt.reconfigureCount();
    node->set_parameter(rclcpp::Parameter("FollowPath.batch_size", ParameterValue(1000)));
    EXPECT_EQ(opt.reconfigureCount(), creates + 1);
    EXPECT_EQ(opt.settings().base.batch_size, 1000);
    OracleSide o(expected_config(1000), expected_critics(9.0f), robot);
    o.take_noise_of(opt);
    for (int k = 0; k < 3; ++k) {compare_cycle(opt, o, Scene(1.0 + 0.03 * k, 1.5, 0.0, 0.1 * k), world, false, "after batch_size change");}
  }
  // ---- every trajectory collides: retry_attempt_limit soft resets, then "Optimizer fail to compute path" (optimizer.cpp:166-183)
  {
    World blocked(true);
    auto node2 = std::make_shared<Node>("controller_server");
    load_yaml_overrides(*node2);
    sortham::ParametersHandler handler2(node2);
    sortham::Optimizer opt2;
    opt2.initialize(node2, "FollowPath", blocked.ros, &handler2);
    FixedGoalChecker checker;
    Scene sc(1.0, 1.5, 0.0, 0.0);
    bool threw = false;
    try {opt2.evalControl(sc.pose, sc.speed, sc.plan, sc.goal, &checker);} catch (const std::runtime_error & e) {
      threw = std::string(e.what()) == "Optimizer fail to compute path";
    }
    EXPECT_TRUE(threw);
  }
  // ---- visualize: true -> the candidate trajectories are materialised and the getters feed TrajectoryVisualizer::add
  {
    auto node4 = std::make_shared<Node>("controller_server");
    load_yaml_overrides(*node4);
    node4->set_override("FollowPath.visualize", true);
    node4->set_override("FollowPath.batch_size", 512);
    sortham::ParametersHandler handler4(node4);
    sortham::Optimizer opt4;
    opt4.initialize(node4, "FollowPath", world.ros, &handler4);
    FixedGoalChecker checker;
    Scene sc(1.0, 1.5, 0.0, 0.1);
    opt4.evalControl(sc.pose, sc.speed, sc.plan, sc.goal, &checker);
    auto & traj = opt4.getGeneratedTrajectories();
    EXPECT_EQ(traj.x.size(), size_t(512) * 56);
    EXPECT_NEAR(traj.x[0], 1.0, 0.05);                     // first pose of the first trajectory sits next to the robot
    const auto best = opt4.getOptimizedTrajectory();
    EXPECT_EQ(best.size(), size_t(56) * 3);
    EXPECT_NEAR(best[0], 1.0, 0.05);
  }
  // ---- an unknown motion model throws at configure time like the reference (optimizer.cpp:421-424)
  {
    auto node3 = std::make_shared<Node>("controller_server");
    load_yaml_overrides(*node3);
    node3->set_override("FollowPath.motion_model", "Hovercraft");
    sortham::ParametersHandler handler3(node3);
    sortham::Optimizer opt3;
    bool threw = false;
    try {opt3.initialize(node3, "FollowPath", world.ros, &handler3);} catch (const std::runtime_error &) {threw = true;}
    EXPECT_TRUE(threw);
  }
}

int main(int argc, char ** argv)
{
  const std::string mode = argc > 1 ? argv[1] : "describe";
  try {
    test_describe();
    if (mode == "cycles") {test_cycles();}
  } catch (const std::exception & e) {
    std::printf("FAIL: uncaught %s\n", e.what());
    ++g_failures;
  }
  std::printf("%s: %d failure(s)\n", mode.c_str(), g_failures);
  return g_failures == 0 ? 0 : 1;
}
