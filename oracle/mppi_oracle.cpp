/*
 * mppi_oracle.cpp -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A dependency-free, single-threaded, strict-fp32 restatement of the optimisation loop of
 * nav2_sortham_controller (soham2560/MPCHoloNavigation), used as the checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * it; nothing under mpcholonavigation_b200/ links, imports or calls it.
 *
 * Parity pinning: every known-answer test the reference holds for this path is ported in
 * tests/test_oracle_golden.py (critics_tests.cpp, optimizer_unit_tests.cpp, motion_model_tests.cpp,
 * utils_test.cpp).  UNPINNED by the reference's own tests (the reference has smoke tests only):
 * updateControlSequence (softmax/gamma), ObstaclesCritic, CostCritic and everything that goes through
 * nav2_costmap_2d (worldToMap, FootprintCollisionChecker, LineIterator, InflationLayer::computeCost).
 * nav2_costmap_2d / angles / tf2 are NOT in /root/reference (apt package ros-humble-navigation2, the
 * fork says 1.1.18); they are restated here from the Nav2 Humble 1.1.x sources' published algorithm.
 * xtensor/xsimd cannot be built here either, so the reference itself cannot be run: "parity unpinned"
 * applies to those rows (DESIGN.md section "Oracle").
 *
 * Citations "ref:" are relative to /root/reference/nav2_sortham_controller/ ; inc/ is
 * include/nav2_sortham_controller/.
 *
 * Canonical arithmetic (what "the reference's result" means where the reference is -ffast-math and
 * therefore not bit-defined): every xtensor expression is evaluated element-wise with C++ usual
 * arithmetic conversions exactly as the expression is typed in the source, reductions and cumsums are
 * sequential in index order, no FMA contraction (build with -ffp-contract=off), sin/cos of the yaw is
 * mppi_det_sincosf (include/mppi_det_math.h).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../include/mppi_b200.h"
#include "../include/mppi_det_math.h"

namespace oracle
{

constexpr unsigned char NO_INFORMATION = 255;              // nav2_costmap_2d/cost_values.hpp
constexpr unsigned char LETHAL_OBSTACLE = 254;
constexpr unsigned char INSCRIBED_INFLATED_OBSTACLE = 253;

// ---------------------------------------------------------------------------------------------
// nav2_costmap_2d restated (Nav2 Humble 1.1.x; not vendored in the reference)
// ---------------------------------------------------------------------------------------------
struct Costmap
{
  std::vector<uint8_t> cells;
  unsigned int size_x{0}, size_y{0};
  double resolution{0.05}, origin_x{0}, origin_y{0};

  // Costmap2D::worldToMap(double, double, unsigned&, unsigned&); call sites ref: inc/tools/utils.hpp:372,
  // src/critics/obstacles_critic.cpp:209, src/critics/cost_critic.cpp:205
  bool worldToMap(double wx, double wy, unsigned int & mx, unsigned int & my) const
  {
    if (wx < origin_x || wy < origin_y) {
      return false;
    }
    const double qx = (wx - origin_x) / resolution;
    const double qy = (wy - origin_y) / resolution;
    // static_cast<unsigned int>(q) is only defined for q < 2^32; a quotient that large is off-map anyway
    if (!(qx < static_cast<double>(size_x)) || !(qy < static_cast<double>(size_y))) {
      return false;
    }
    mx = static_cast<unsigned int>(qx);
    my = static_cast<unsigned int>(qy);
    return mx < size_x && my < size_y;
  }
  unsigned char getCost(unsigned int mx, unsigned int my) const {return cells[my * size_x + mx];}
};

// nav2_util::LineIterator (Bresenham, both end points included)
struct LineIterator
{
  LineIterator(int x0, int y0, int x1, int y1)
  : x_(x0), y_(y0), deltax_(std::abs(x1 - x0)), deltay_(std::abs(y1 - y0)), curpixel_(0)
  {
    if (x1 >= x0) {xinc1_ = 1; xinc2_ = 1;} else {xinc1_ = -1; xinc2_ = -1;}
    if (y1 >= y0) {yinc1_ = 1; yinc2_ = 1;} else {yinc1_ = -1; yinc2_ = -1;}
    if (deltax_ >= deltay_) {
      xinc1_ = 0; yinc2_ = 0;
      den_ = deltax_; num_ = deltax_ / 2; numadd_ = deltay_; numpixels_ = deltax_;
    } else {
      xinc2_ = 0; yinc1_ = 0;
      den_ = deltay_; num_ = deltay_ / 2; numadd_ = deltax_; numpixels_ = deltay_;
    }
  }
  bool isValid() const {return curpixel_ <= numpixels_;}
  void advance()
  {
    num_ += numadd_;
    if (num_ >= den_) {num_ -= den_; x_ += xinc1_; y_ += yinc1_;}
    x_ += xinc2_; y_ += yinc2_;
    curpixel_++;
  }
  int x_, y_, deltax_, deltay_, curpixel_;
  int xinc1_, xinc2_, yinc1_, yinc2_, den_, num_, numadd_, numpixels_;
};

// nav2_costmap_2d::FootprintCollisionChecker<Costmap2D*>
struct FootprintCollisionChecker
{
  const Costmap * costmap{nullptr};
  double pointCost(int x, int y) const {return static_cast<double>(costmap->getCost(x, y));}
  double lineCost(int x0, int x1, int y0, int y1) const
  {
    double line_cost = 0.0;
    double point_cost = -1.0;
    for (LineIterator line(x0, y0, x1, y1); line.isValid(); line.advance()) {
      point_cost = pointCost(line.x_, line.y_);
      if (point_cost == static_cast<double>(LETHAL_OBSTACLE)) {
        return point_cost;
      }
      if (line_cost < point_cost) {
        line_cost = point_cost;
      }
    }
    return line_cost;
  }
  double footprintCost(const std::vector<double> & fx, const std::vector<double> & fy) const
  {
    unsigned int x0, x1, y0, y1;
    double footprint_cost = 0.0;
    if (!costmap->worldToMap(fx[0], fy[0], x0, y0)) {
      return static_cast<double>(LETHAL_OBSTACLE);
    }
    const unsigned int xstart = x0, ystart = y0;
    x1 = x0; y1 = y0;
    for (size_t i = 0; i + 1 < fx.size(); ++i) {
      if (!costmap->worldToMap(fx[i + 1], fy[i + 1], x1, y1)) {
        return static_cast<double>(LETHAL_OBSTACLE);
      }
      footprint_cost = std::max(lineCost(x0, x1, y0, y1), footprint_cost);
      x0 = x1; y0 = y1;
      if (footprint_cost == static_cast<double>(LETHAL_OBSTACLE)) {
        return footprint_cost;
      }
    }
    return std::max(lineCost(xstart, x1, ystart, y1), footprint_cost);
  }
  double footprintCostAtPose(double x, double y, double theta, const mppi_robot_desc & robot) const
  {
    const double cos_th = std::cos(theta);
    const double sin_th = std::sin(theta);
    std::vector<double> fx(robot.footprint_size), fy(robot.footprint_size);
    for (int i = 0; i < robot.footprint_size; ++i) {
      fx[i] = x + (robot.footprint_x[i] * cos_th - robot.footprint_y[i] * sin_th);
      fy[i] = y + (robot.footprint_x[i] * sin_th + robot.footprint_y[i] * cos_th);
    }
    return footprintCost(fx, fy);
  }
};

// nav2_costmap_2d::InflationLayer::computeCost(double distance_in_cells)
static unsigned char inflationComputeCost(
  double distance, double resolution, double inscribed_radius, double cost_scaling_factor)
{
  unsigned char cost = 0;
  if (distance == 0) {
    cost = LETHAL_OBSTACLE;
  } else if (distance * resolution <= inscribed_radius) {
    cost = INSCRIBED_INFLATED_OBSTACLE;
  } else {
    const double factor = std::exp(-1.0 * cost_scaling_factor * (distance * resolution - inscribed_radius));
    cost = static_cast<unsigned char>((INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
  }
  return cost;
}

// ---------------------------------------------------------------------------------------------
// angles / tf2 restated
// ---------------------------------------------------------------------------------------------
static double normalize_angle(double angle)   // ros angles::normalize_angle
{
  const double result = std::fmod(angle + M_PI, 2.0 * M_PI);
  if (result <= 0.0) {return result + M_PI;}
  return result - M_PI;
}
static double shortest_angular_distance(double from, double to) {return normalize_angle(to - from);}

// ref: inc/tools/utils.hpp:258-263 (xt::fmod(angles + M_PI, 2.0 * M_PI), element type float -> double)
static double utils_normalize_angles(float angle)
{
  const double theta = std::fmod(static_cast<double>(angle) + M_PI, 2.0 * M_PI);
  return theta <= 0.0 ? theta + M_PI : theta - M_PI;
}
// ref: inc/tools/utils.hpp:278-284 : normalize_angles(to - from), the subtraction is float - float
static double utils_shortest_angular_distance(float from, float to) {return utils_normalize_angles(to - from);}

// ---------------------------------------------------------------------------------------------
// data model (ref: inc/models/*.hpp) : [B,T] row-major float planes
// ---------------------------------------------------------------------------------------------
struct Plane
{
  std::vector<float> v;
  size_t B{0}, T{0};
  void reset(size_t b, size_t t) {B = b; T = t; v.assign(b * t, 0.0f);}
  float & operator()(size_t b, size_t t) {return v[b * T + t];}
  float operator()(size_t b, size_t t) const {return v[b * T + t];}
};

struct State   // ref: inc/models/state.hpp:30-56
{
  Plane vx, vy, wz, cvx, cvy, cwz;
  double pose_x{0}, pose_y{0}, pose_yaw{0};
  double speed_vx{0}, speed_vy{0}, speed_wz{0};
  void reset(size_t b, size_t t)
  {
    vx.reset(b, t); vy.reset(b, t); wz.reset(b, t); cvx.reset(b, t); cvy.reset(b, t); cwz.reset(b, t);
  }
};
struct Trajectories {Plane x, y, yaws; void reset(size_t b, size_t t) {x.reset(b, t); y.reset(b, t); yaws.reset(b, t);}};
struct ControlSequence {std::vector<float> vx, vy, wz; void reset(size_t t) {vx.assign(t, 0.f); vy.assign(t, 0.f); wz.assign(t, 0.f);}};
struct Path {std::vector<float> x, y, yaws;};
struct Constraints {float vx_max, vx_min, vy, wz;};   // ref: inc/models/constraints.hpp:25-31

struct CriticData   // ref: inc/critic_data.hpp:38-53
{
  const State * state;
  const Trajectories * trajectories;
  const Path * path;
  double goal_x, goal_y;
  std::vector<float> * costs;
  float model_dt;
  bool fail_flag{false};
  double goal_checker_xy_tolerance{-1.0};   // < 0 : goal_checker == nullptr
  int motion_model{MPPI_MODEL_DIFF_DRIVE};
  float ackermann_min_turning_r{0.2f};
  bool path_pts_valid_set{false};
  std::vector<bool> path_pts_valid;
  bool furthest_set{false};
  size_t furthest_reached_path_point{0};
};

// ---------------------------------------------------------------------------------------------
// utils (ref: inc/tools/utils.hpp)
// ---------------------------------------------------------------------------------------------
// ref: utils.hpp:201-224
static bool withinGoalCheckerTolerance(double goal_checker_tol, double rx, double ry, double gx, double gy)
{
  if (goal_checker_tol >= 0.0) {
    const double pose_tolerance_sq = goal_checker_tol * goal_checker_tol;
    const double dx = rx - gx, dy = ry - gy;
    if (dx * dx + dy * dy < pose_tolerance_sq) {return true;}
  }
  return false;
}
// ref: utils.hpp:233-249
static bool withinPositionGoalTolerance(float pose_tolerance, double rx, double ry, double gx, double gy)
{
  const double dist_sq = std::pow(gx - rx, 2) + std::pow(gy - ry, 2);
  const float pose_tolerance_sq = pose_tolerance * pose_tolerance;
  return dist_sq < pose_tolerance_sq;
}

// ref: utils.hpp:292-319
static size_t findPathFurthestReachedPoint(const CriticData & data)
{
  const Trajectories & tr = *data.trajectories;
  const Path & path = *data.path;
  const size_t B = tr.x.B, T = tr.x.T, N = path.x.size();
  size_t max_id_by_trajectories = 0;
  if (T == 0) {return 0;}
  for (size_t i = 0; i < B; i++) {
    size_t min_id_by_path = 0;
    float min_distance_by_path = std::numeric_limits<float>::max();
    const float tx = tr.x(i, T - 1), ty = tr.y(i, T - 1);
    for (size_t j = 0; j < N; j++) {
      const float dx = path.x[j] - tx;
      const float dy = path.y[j] - ty;
      const float cur_dist = dx * dx + dy * dy;
      if (cur_dist < min_distance_by_path) {
        min_distance_by_path = cur_dist;
        min_id_by_path = j;
      }
    }
    max_id_by_trajectories = std::max(max_id_by_trajectories, min_id_by_path);
  }
  return max_id_by_trajectories;
}

// ref: utils.hpp:327-344
static size_t findPathTrajectoryInitialPoint(const CriticData & data)
{
  const Trajectories & tr = *data.trajectories;
  const Path & path = *data.path;
  const float x0 = tr.x(0, 0), y0 = tr.y(0, 0);
  float min_distance_by_path = std::numeric_limits<float>::max();
  size_t min_id = 0;
  for (size_t j = 0; j < path.x.size(); j++) {
    const float dx = path.x[j] - x0;
    const float dy = path.y[j] - y0;
    const float d = dx * dx + dy * dy;
    if (d < min_distance_by_path) {
      min_distance_by_path = d;
      min_id = j;
    }
  }
  return min_id;
}

// ref: utils.hpp:350-355
static void setPathFurthestPointIfNotSet(CriticData & data)
{
  if (!data.furthest_set) {
    data.furthest_reached_path_point = findPathFurthestReachedPoint(data);
    data.furthest_set = true;
  }
}

// ref: utils.hpp:361-394
static void findPathCosts(CriticData & data, const Costmap & costmap, bool is_tracking_unknown)
{
  const Path & path = *data.path;
  const size_t path_segments_count = path.x.size() - 1;
  data.path_pts_valid.assign(path_segments_count, false);
  data.path_pts_valid_set = true;
  unsigned int map_x, map_y;
  for (unsigned int idx = 0; idx < path_segments_count; idx++) {
    if (!costmap.worldToMap(path.x[idx], path.y[idx], map_x, map_y)) {
      data.path_pts_valid[idx] = false;
      continue;
    }
    switch (costmap.getCost(map_x, map_y)) {
      case LETHAL_OBSTACLE: data.path_pts_valid[idx] = false; continue;
      case INSCRIBED_INFLATED_OBSTACLE: data.path_pts_valid[idx] = false; continue;
      case NO_INFORMATION: data.path_pts_valid[idx] = is_tracking_unknown ? true : false; continue;
    }
    data.path_pts_valid[idx] = true;
  }
}
// ref: utils.hpp:400-407
static void setPathCostsIfNotSet(CriticData & data, const Costmap & costmap, bool is_tracking_unknown)
{
  if (!data.path_pts_valid_set) {findPathCosts(data, costmap, is_tracking_unknown);}
}

// ref: utils.hpp:417-434
static float posePointAngle(double pose_xd, double pose_yd, double pose_yawd, double point_x, double point_y,
  bool forward_preference)
{
  const float pose_x = pose_xd;
  const float pose_y = pose_yd;
  const float pose_yaw = pose_yawd;
  const float yaw = atan2f(point_y - pose_y, point_x - pose_x);
  if (!forward_preference) {
    return std::min(
      fabs(shortest_angular_distance(yaw, pose_yaw)),
      fabs(shortest_angular_distance(yaw, normalize_angle(pose_yaw + M_PI))));
  }
  return fabs(shortest_angular_distance(yaw, pose_yaw));
}

// ref: utils.hpp:665-675.  The reference dereferences vec.end() when dist exceeds the last entry (UB);
// the oracle DEFINES that case as "clamp to the last index" (what a large garbage read gives).
static size_t findClosestPathPt(const std::vector<float> & vec, float dist, size_t init = 0)
{
  auto iter = std::lower_bound(vec.begin() + init, vec.end(), dist);
  if (iter == vec.begin() + init) {
    return 0;
  }
  if (iter == vec.end()) {
    return vec.size() - 1;
  }
  if (dist - *(iter - 1) < *iter - dist) {
    return iter - 1 - vec.begin();
  }
  return iter - vec.begin();
}

// ref: utils.hpp:442-605 (9-point quadratic Savitzky-Golay; quirks: index num_sequences-4 is skipped by
// the extra idx++ after the loop, already-filtered neighbours are reused, vy is filtered too)
static void savitskyGolayFilter(ControlSequence & cs, float history[4][3], bool shift_control_sequence)
{
  float filter[9] = {-21.0f, 14.0f, 39.0f, 54.0f, 59.0f, 54.0f, 39.0f, 14.0f, -21.0f};
  for (float & f : filter) {f /= 231.0f;}
  const unsigned int num_sequences = cs.vx.size() - 1;
  if (num_sequences < 20) {return;}
  auto applyFilter = [&](const float (&d)[9]) -> float {
      float s = 0.0f;
      for (int i = 0; i < 9; ++i) {s += d[i] * filter[i];}
      return s;
    };
  auto applyFilterOverAxis = [&](std::vector<float> & s, float h0, float h1, float h2, float h3) {
      unsigned int idx = 0;
      s[idx] = applyFilter({h0, h1, h2, h3, s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 4]});
      idx++;
      s[idx] = applyFilter({h1, h2, h3, s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 4]});
      idx++;
      s[idx] = applyFilter({h2, h3, s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 4]});
      idx++;
      s[idx] = applyFilter({h3, s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 4]});
      for (idx = 4; idx != num_sequences - 4; idx++) {
        s[idx] = applyFilter({s[idx - 4], s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 4]});
      }
      idx++;
      s[idx] = applyFilter({s[idx - 4], s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 3], s[idx + 3]});
      idx++;
      s[idx] = applyFilter({s[idx - 4], s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 2], s[idx + 2], s[idx + 2]});
      idx++;
      s[idx] = applyFilter({s[idx - 4], s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx + 1], s[idx + 1], s[idx + 1], s[idx + 1]});
      idx++;
      s[idx] = applyFilter({s[idx - 4], s[idx - 3], s[idx - 2], s[idx - 1], s[idx], s[idx], s[idx], s[idx], s[idx]});
    };
  applyFilterOverAxis(cs.vx, history[0][0], history[1][0], history[2][0], history[3][0]);
  applyFilterOverAxis(cs.vy, history[0][1], history[1][1], history[2][1], history[3][1]);
  applyFilterOverAxis(cs.wz, history[0][2], history[1][2], history[2][2], history[3][2]);
  const unsigned int offset = shift_control_sequence ? 1 : 0;
  for (int k = 0; k < 3; ++k) {
    history[0][k] = history[1][k]; history[1][k] = history[2][k]; history[2][k] = history[3][k];
  }
  history[3][0] = cs.vx[offset]; history[3][1] = cs.vy[offset]; history[3][2] = cs.wz[offset];
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11; Random123) and
// the noise layout of the K1 kernel: counter = (t/4, plane, global b, stream), key = seed.
// ---------------------------------------------------------------------------------------------
static void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4])
{
  uint32_t c[4] = {ctr_in[0], ctr_in[1], ctr_in[2], ctr_in[3]};
  uint32_t k[2] = {key_in[0], key_in[1]};
  for (int r = 0; r < 10; ++r) {
    if (r > 0) {k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;}
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c[0];
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c[2];
    const uint32_t hi0 = p0 >> 32, lo0 = static_cast<uint32_t>(p0);
    const uint32_t hi1 = p1 >> 32, lo1 = static_cast<uint32_t>(p1);
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

static void boxMuller(uint32_t a, uint32_t b, float & z0, float & z1)
{
  const float u1 = static_cast<float>((a >> 8) + 1u) * 5.9604644775390625e-08f;   // (0,1]
  const float u2 = static_cast<float>(b >> 8) * 5.9604644775390625e-08f;          // [0,1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  mppi_det_sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

static void philoxNoisePlane(
  std::vector<float> & plane, size_t B, size_t T, uint32_t plane_id, float stddev, uint64_t seed,
  uint64_t stream, uint64_t shard_offset)
{
  const uint32_t key[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  for (size_t b = 0; b < B; ++b) {
    for (size_t q = 0; q * 4 < T; ++q) {
      const uint32_t ctr[4] = {static_cast<uint32_t>(q), plane_id, static_cast<uint32_t>(b + shard_offset),
        static_cast<uint32_t>(stream)};
      uint32_t r[4];
      philox4x32_10(ctr, key, r);
      float z[4];
      boxMuller(r[0], r[1], z[0], z[1]);
      boxMuller(r[2], r[3], z[2], z[3]);
      for (size_t j = 0; j < 4 && q * 4 + j < T; ++j) {
        plane[b * T + q * 4 + j] = z[j] * stddev;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// the optimizer
// ---------------------------------------------------------------------------------------------
struct Critic
{
  mppi_critic_desc d;
  // derived at initialize()
  float weight;                    // CostCritic: cost_weight / 254.0f (ref: cost_critic.cpp:34)
  float max_vel, min_vel;          // ConstraintCritic (ref: constraint_critic.cpp:31-38)
  bool reversing_allowed{true};    // PathAngleCritic (ref: path_angle_critic.cpp:26-32; member default true)
  bool forward_preference{true};
  float possibly_inscribed_cost{-1.0f};                      // Cost / Obstacles findCircumscribedCost
  float inflation_scale_factor{0.0f}, inflation_radius{0.0f};  // Obstacles (ref: obstacles_critic.hpp)
};

struct Optimizer
{
  mppi_config cfg;
  Constraints base_constraints, constraints;
  mppi_robot_desc robot;
  std::vector<Critic> critics;
  State state;
  Trajectories traj;
  ControlSequence cs;
  Path path;
  Costmap costmap;
  std::vector<float> costs;
  Plane noises_vx, noises_vy, noises_wz;
  std::vector<std::vector<float>> critic_costs;   // [critic][B] : costs added by each critic in the last iteration
  std::vector<int32_t> cells;
  CriticData data;
  uint64_t noise_stream{0};
  float control_history[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // ref: inc/optimizer.hpp:251
  std::string err;
  // Test switch (oracle_set_wide_reductions): the softmax normaliser and the weighted control sums accumulate in double.
  // The reference types these accumulators as float and leaves their ORDER to xtensor/xsimd under -ffast-math (xt::sum over
  // [B], xt::sum(.., 0) over [B,T]; optimizer.cpp:384-391), so its own bits are not defined.  For the sizes the reference is
  // used at (B ~ 1000) every order agrees to ~1e-6; at B = 262144 a float accumulator in index order drops the softmax tail
  // below half an ulp of the running sum (3e-4 relative, larger than the parity tolerance), so the full-size test compares
  // with this order-free value and reports the index-order float result beside it.
  bool wide_reductions{false};
  // Test switch (oracle_set_iteration_controls): behind the update of iteration `pin_iteration` the control sequence is
  // replaced by the given one (the bits another implementation holds at that point), so that the NEXT iteration's rollout can
  // be compared bit for bit although the two updates differ in the order of their batch reductions.
  int pin_iteration{-1};
  ControlSequence pin_cs;
  // Measurement aid (oracle_get_counters): how many poses the obstacle-type critics visited in the last optimize() and how
  // many of them took the footprint branch (cost_critic.cpp:204-209, obstacles_critic.cpp:214-220) - the "branch fraction"
  // SURVEY 8d asks for beside config 3's numbers.  [0] CostCritic visited, [1] CostCritic footprint, [2] Obstacles visited,
  // [3] Obstacles footprint.
  mutable uint64_t counters[4] = {0, 0, 0, 0};

  bool isHolonomic() const {return cfg.motion_model == MPPI_MODEL_OMNI;}

  // ref: NoiseGenerator::generateNoisedControls noise_generator.cpp:107-122 (xt::random::randn is an
  // unseeded process-global mt19937: there is no reference stream to match, the oracle restates the
  // Philox layout of the CUDA K1 kernel instead)
  void generateNoisedControls()
  {
    const size_t B = cfg.batch_size, T = cfg.time_steps;
    philoxNoisePlane(noises_vx.v, B, T, 0, cfg.vx_std, cfg.seed, noise_stream, cfg.shard_offset);
    philoxNoisePlane(noises_wz.v, B, T, 2, cfg.wz_std, cfg.seed, noise_stream, cfg.shard_offset);
    if (isHolonomic()) {
      philoxNoisePlane(noises_vy.v, B, T, 1, cfg.vy_std, cfg.seed, noise_stream, cfg.shard_offset);
    }
  }

  // ref: Optimizer::reset optimizer.cpp:116-132 + NoiseGenerator::reset noise_generator.cpp:76-95
  void reset()
  {
    const size_t B = cfg.batch_size, T = cfg.time_steps;
    state.reset(B, T);
    cs.reset(T);
    std::memset(control_history, 0, sizeof(control_history));   // optimizer.cpp:120-123
    constraints = base_constraints;
    costs.assign(B, 0.0f);
    traj.reset(B, T);
    noises_vx.reset(B, T); noises_vy.reset(B, T); noises_wz.reset(B, T);
    generateNoisedControls();
    noise_stream++;
  }

  // ref: NoiseGenerator::setNoisedControls noise_generator.cpp:65-74
  void setNoisedControls()
  {
    const size_t B = cfg.batch_size, T = cfg.time_steps;
    for (size_t b = 0; b < B; ++b) {
      for (size_t t = 0; t < T; ++t) {
        state.cvx(b, t) = cs.vx[t] + noises_vx(b, t);
        state.cvy(b, t) = cs.vy[t] + noises_vy(b, t);
        state.cwz(b, t) = cs.wz[t] + noises_wz(b, t);
      }
    }
  }

  // ref: Optimizer::updateStateVelocities optimizer.cpp:251-273, MotionModel::predict motion_models.hpp:53-66
  static void updateStateVelocities(State & s, bool holonomic)
  {
    const size_t B = s.vx.B, T = s.vx.T;
    for (size_t b = 0; b < B; ++b) {
      s.vx(b, 0) = s.speed_vx;
      s.wz(b, 0) = s.speed_wz;
      if (holonomic) {s.vy(b, 0) = s.speed_vy;}
      for (size_t t = 1; t < T; ++t) {
        s.vx(b, t) = s.cvx(b, t - 1);
        s.wz(b, t) = s.cwz(b, t - 1);
        if (holonomic) {s.vy(b, t) = s.cvy(b, t - 1);}
      }
    }
  }

  // ref: Optimizer::integrateStateVelocities(Trajectories&, const State&) optimizer.cpp:313-343
  static void integrateStateVelocities(Trajectories & tr, const State & s, float model_dt, bool holonomic)
  {
    const size_t B = s.vx.B, T = s.vx.T;
    tr.reset(B, T);
    const float initial_yaw = s.pose_yaw;   // const float initial_yaw = tf2::getYaw(...)
    float cos0, sin0;
    mppi_det_sincosf(initial_yaw, &sin0, &cos0);
    for (size_t b = 0; b < B; ++b) {
      // xt::cumsum(state.wz * model_dt, 1) + initial_yaw
      float acc = 0.0f;
      for (size_t t = 0; t < T; ++t) {
        const float term = s.wz(b, t) * model_dt;
        acc = (t == 0) ? term : acc + term;
        tr.yaws(b, t) = acc + initial_yaw;
      }
      float accx = 0.0f, accy = 0.0f;
      for (size_t t = 0; t < T; ++t) {
        float yaw_cos, yaw_sin;
        if (t == 0) {
          yaw_cos = cos0; yaw_sin = sin0;
        } else {
          mppi_det_sincosf(tr.yaws(b, t - 1), &yaw_sin, &yaw_cos);
        }
        float dx = s.vx(b, t) * yaw_cos;
        float dy = s.vx(b, t) * yaw_sin;
        if (holonomic) {
          dx = dx - s.vy(b, t) * yaw_sin;
          dy = dy + s.vy(b, t) * yaw_cos;
        }
        const float tx = dx * model_dt, ty = dy * model_dt;
        accx = (t == 0) ? tx : accx + tx;
        accy = (t == 0) ? ty : accy + ty;
        // state.pose.pose.position.x (double) + cumsum (float) -> double, stored to float
        tr.x(b, t) = static_cast<float>(s.pose_x + static_cast<double>(accx));
        tr.y(b, t) = static_cast<float>(s.pose_y + static_cast<double>(accy));
      }
    }
  }

  // ref: Optimizer::applyControlSequenceConstraints optimizer.cpp:237-249 (+ Ackermann motion_models.hpp:110-117)
  void applyControlSequenceConstraints()
  {
    const size_t T = cs.vx.size();
    for (size_t t = 0; t < T; ++t) {
      if (isHolonomic()) {cs.vy[t] = std::min(std::max(cs.vy[t], -constraints.vy), constraints.vy);}
      cs.vx[t] = std::min(std::max(cs.vx[t], constraints.vx_min), constraints.vx_max);
      cs.wz[t] = std::min(std::max(cs.wz[t], -constraints.wz), constraints.wz);
    }
    if (cfg.motion_model == MPPI_MODEL_ACKERMANN) {
      const float r = cfg.ackermann_min_turning_r;
      for (size_t t = 0; t < T; ++t) {
        if ((std::fabs(cs.vx[t]) / std::fabs(cs.wz[t])) < r) {
          const float sgn = (cs.wz[t] > 0.0f) ? 1.0f : ((cs.wz[t] < 0.0f) ? -1.0f : 0.0f);
          cs.wz[t] = sgn * std::fabs(cs.vx[t]) / r;
        }
      }
    }
  }

  // ref: Optimizer::updateControlSequence optimizer.cpp:362-394
  void updateControlSequence()
  {
    const size_t B = cfg.batch_size, T = cfg.time_steps;
    auto gammaTerm = [&](const Plane & c, const std::vector<float> & seq, float stddev) {
        const float k = cfg.gamma / powf(stddev, 2);
        for (size_t b = 0; b < B; ++b) {
          float sum = 0.0f;
          for (size_t t = 0; t < T; ++t) {
            const float bounded = c(b, t) - seq[t];
            sum += seq[t] * bounded;
          }
          costs[b] += k * sum;
        }
      };
    gammaTerm(state.cvx, cs.vx, cfg.vx_std);
    gammaTerm(state.cwz, cs.wz, cfg.wz_std);
    if (isHolonomic()) {gammaTerm(state.cvy, cs.vy, cfg.vy_std);}

    float cmin = std::numeric_limits<float>::max();
    for (size_t b = 0; b < B; ++b) {cmin = std::min(cmin, costs[b]);}
    std::vector<float> softmaxes(B);
    float sum = 0.0f;
    double wide_sum = 0.0;
    const float neg_inv_temp = -1 / cfg.temperature;
    for (size_t b = 0; b < B; ++b) {
      softmaxes[b] = expf(neg_inv_temp * (costs[b] - cmin));
      sum += softmaxes[b];
      wide_sum += static_cast<double>(softmaxes[b]);
    }
    if (wide_reductions) {sum = static_cast<float>(wide_sum);}
    for (size_t b = 0; b < B; ++b) {softmaxes[b] = softmaxes[b] / sum;}
    auto weighted = [&](const Plane & c, std::vector<float> & seq) {
        for (size_t t = 0; t < T; ++t) {
          if (wide_reductions) {
            double acc = 0.0;
            for (size_t b = 0; b < B; ++b) {acc += static_cast<double>(c(b, t) * softmaxes[b]);}
            seq[t] = static_cast<float>(acc);
          } else {
            float acc = 0.0f;
            for (size_t b = 0; b < B; ++b) {acc += c(b, t) * softmaxes[b];}
            seq[t] = acc;
          }
        }
      };
    weighted(state.cvx, cs.vx);
    weighted(state.cwz, cs.wz);
    if (isHolonomic()) {weighted(state.cvy, cs.vy);}
    applyControlSequenceConstraints();
  }

  // ------------------------------------------------------------------------------------------
  // critics
  // ------------------------------------------------------------------------------------------
  // ref: ObstaclesCritic::findCircumscribedCost obstacles_critic.cpp:53-97, CostCritic cost_critic.cpp:63-106
  float findCircumscribedCost() const
  {
    double result = -1.0;
    if (robot.inflation_layer_found) {
      result = inflationComputeCost(
        robot.circumscribed_radius / costmap.resolution, costmap.resolution, robot.inscribed_radius,
        robot.inflation_cost_scaling_factor);
    }
    return static_cast<float>(result);
  }

  void initializeCritic(Critic & c) const
  {
    const mppi_critic_desc & d = c.d;
    c.weight = d.cost_weight;
    switch (d.kind) {
      case MPPI_CRITIC_CONSTRAINT: {   // ref: constraint_critic.cpp:31-38
          const float vx_max = cfg.vx_max, vy_max = cfg.vy_max, vx_min = cfg.vx_min;
          const float min_sgn = vx_min > 0.0 ? 1.0 : -1.0;
          c.max_vel = sqrtf(vx_max * vx_max + vy_max * vy_max);
          c.min_vel = min_sgn * sqrtf(vx_min * vx_min + vy_max * vy_max);
          break;
        }
      case MPPI_CRITIC_COST:           // ref: cost_critic.cpp:34
        c.weight = d.cost_weight / 254.0f;
        break;
      case MPPI_CRITIC_PATH_ANGLE: {   // ref: path_angle_critic.cpp:23-50
          const float vx_min = cfg.vx_min;
          c.reversing_allowed = true;
          if (fabs(vx_min) < 1e-6) {
            c.reversing_allowed = false;
          } else if (vx_min < 0.0) {
            c.reversing_allowed = true;
          }
          c.forward_preference = d.forward_preference != 0;
          if (!c.reversing_allowed) {c.forward_preference = true;}
          break;
        }
      case MPPI_CRITIC_OBSTACLES:      // ref: obstacles_critic.cpp:70-81 (read only when the layer exists)
        if (robot.inflation_layer_found) {
          c.inflation_scale_factor = d.cost_scaling_factor;
          c.inflation_radius = d.inflation_radius;
        } else {
          c.inflation_scale_factor = 0.0f;
          c.inflation_radius = 0.0f;
        }
        break;
      default: break;
    }
  }

  static void addPow(std::vector<float> & costs, size_t b, double value, unsigned int power)
  {
    // data.costs += xt::pow(expr, power_): std::pow(<float|double>, unsigned) is evaluated in double,
    // the sum float + double is double, the assignment narrows to float
    costs[b] = static_cast<float>(static_cast<double>(costs[b]) + std::pow(value, static_cast<double>(power)));
  }

  // ref: ConstraintCritic::score constraint_critic.cpp:41-75
  void scoreConstraint(const Critic & c, CriticData & data) const
  {
    if (!c.d.enabled) {return;}
    const State & s = *data.state;
    const size_t B = s.vx.B, T = s.vx.T;
    const bool acker = data.motion_model == MPPI_MODEL_ACKERMANN;
    for (size_t b = 0; b < B; ++b) {
      double sum = 0.0;
      for (size_t t = 0; t < T; ++t) {
        const float vx = s.vx(b, t), vy = s.vy(b, t);
        const double sgn = vx > 0.0 ? 1.0 : -1.0;
        const double vel_total = sgn * sqrtf(vx * vx + vy * vy);
        const double out_of_max = std::max(vel_total - c.max_vel, 0.0);
        const double out_of_min = std::max(c.min_vel - vel_total, 0.0);
        double e = out_of_max + out_of_min;
        if (acker) {
          const float wz = s.wz(b, t);
          // xt::maximum(min_r - fabs(vx)/fabs(wz), 0.0); 0/0 is defined here as "no violation" (fmax)
          const float ratio = std::fabs(vx) / std::fabs(wz);
          e = e + std::fmax(static_cast<double>(data.ackermann_min_turning_r - ratio), 0.0);
        }
        sum += e * data.model_dt;
      }
      addPow(*data.costs, b, sum * c.weight, c.d.cost_power);
    }
  }

  // ref: GoalCritic::score goal_critic.cpp:36-55
  void scoreGoal(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    if (!c.d.enabled || !withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return;
    }
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, T = tr.x.T;
    for (size_t b = 0; b < B; ++b) {
      double sum = 0.0;
      for (size_t t = 0; t < T; ++t) {
        const double dx = tr.x(b, t) - data.goal_x;
        const double dy = tr.y(b, t) - data.goal_y;
        sum += std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
      }
      addPow(*data.costs, b, (sum / static_cast<double>(T)) * c.weight, c.d.cost_power);
    }
  }

  // ref: GoalAngleCritic::score goal_angle_critic.cpp:36-50
  void scoreGoalAngle(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    if (!c.d.enabled || !withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return;
    }
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, T = tr.x.T;
    const float goal_yaw = data.path->yaws[data.path->x.size() - 1];
    for (size_t b = 0; b < B; ++b) {
      double sum = 0.0;
      for (size_t t = 0; t < T; ++t) {
        sum += std::fabs(utils_shortest_angular_distance(tr.yaws(b, t), goal_yaw));
      }
      addPow(*data.costs, b, (sum / static_cast<double>(T)) * c.weight, c.d.cost_power);
    }
  }

  // ref: PreferForwardCritic::score prefer_forward_critic.cpp:33-47
  void scorePreferForward(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    if (!c.d.enabled || withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return;
    }
    const size_t B = s.vx.B, T = s.vx.T;
    for (size_t b = 0; b < B; ++b) {
      float sum = 0.0f;
      for (size_t t = 0; t < T; ++t) {
        const float backward_motion = std::max(-s.vx(b, t), 0.0f);
        sum += backward_motion * data.model_dt;
      }
      addPow(*data.costs, b, sum * c.weight, c.d.cost_power);
    }
  }

  // ref: TwirlingCritic::score twirling_critic.cpp:31-42
  void scoreTwirling(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    if (!c.d.enabled || withinGoalCheckerTolerance(data.goal_checker_xy_tolerance, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return;
    }
    const size_t B = s.wz.B, T = s.wz.T;
    for (size_t b = 0; b < B; ++b) {
      float sum = 0.0f;
      for (size_t t = 0; t < T; ++t) {sum += std::fabs(s.wz(b, t));}
      addPow(*data.costs, b, (sum / static_cast<float>(T)) * c.weight, c.d.cost_power);
    }
  }

  // ref: VelocityDeadbandCritic::score velocity_deadband_critic.cpp:41-98
  void scoreVelocityDeadband(const Critic & c, CriticData & data) const
  {
    if (!c.d.enabled) {return;}
    const State & s = *data.state;
    const size_t B = s.vx.B, T = s.vx.T;
    const bool holonomic = data.motion_model == MPPI_MODEL_OMNI;
    for (size_t b = 0; b < B; ++b) {
      float sum = 0.0f;
      for (size_t t = 0; t < T; ++t) {
        float e = std::max(fabsf(c.d.deadband_velocities[0]) - std::fabs(s.vx(b, t)), 0.0f);
        if (holonomic) {e = e + std::max(fabsf(c.d.deadband_velocities[1]) - std::fabs(s.vy(b, t)), 0.0f);}
        e = e + std::max(fabsf(c.d.deadband_velocities[2]) - std::fabs(s.wz(b, t)), 0.0f);
        sum += e * data.model_dt;
      }
      if (c.d.cost_power > 1u) {
        addPow(*data.costs, b, sum * c.weight, c.d.cost_power);
      } else {
        (*data.costs)[b] += sum * c.weight;
      }
    }
  }

  // ref: PathFollowCritic::score path_follow_critic.cpp:35-71
  void scorePathFollow(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    const Path & path = *data.path;
    if (!c.d.enabled || path.x.size() < 2 ||
      withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y))
    {
      return;
    }
    setPathFurthestPointIfNotSet(data);
    setPathCostsIfNotSet(data, costmap, robot.track_unknown != 0);
    const size_t path_size = path.x.size() - 1;
    size_t offseted_idx = std::min(data.furthest_reached_path_point + static_cast<size_t>(c.d.offset_from_furthest), path_size);
    bool valid = false;
    while (!valid && offseted_idx < path_size - 1) {
      valid = data.path_pts_valid[offseted_idx];
      if (!valid) {offseted_idx++;}
    }
    const float path_x = path.x[offseted_idx], path_y = path.y[offseted_idx];
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, T = tr.x.T;
    for (size_t b = 0; b < B; ++b) {
      const float dx = tr.x(b, T - 1) - path_x;
      const float dy = tr.y(b, T - 1) - path_y;
      const double dist = std::sqrt(std::pow(dx, 2) + std::pow(dy, 2));
      addPow(*data.costs, b, c.weight * dist, c.d.cost_power);
    }
  }

  // ref: PathAngleCritic::score path_angle_critic.cpp:58-101
  void scorePathAngle(const Critic & c, CriticData & data) const
  {
    if (!c.d.enabled) {return;}
    const State & s = *data.state;
    if (withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return;
    }
    setPathFurthestPointIfNotSet(data);
    const Path & path = *data.path;
    const size_t offseted_idx = std::min(
      data.furthest_reached_path_point + static_cast<size_t>(c.d.offset_from_furthest), path.x.size() - 1);
    const float goal_x = path.x[offseted_idx];
    const float goal_y = path.y[offseted_idx];
    if (posePointAngle(s.pose_x, s.pose_y, s.pose_yaw, goal_x, goal_y, c.forward_preference) < c.d.max_angle_to_furthest) {
      return;
    }
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, T = tr.x.T;
    for (size_t b = 0; b < B; ++b) {
      double sum = 0.0;
      for (size_t t = 0; t < T; ++t) {
        const float ybp = atan2f(goal_y - tr.y(b, t), goal_x - tr.x(b, t));
        const double yaws = std::fabs(utils_shortest_angular_distance(tr.yaws(b, t), ybp));
        if (c.reversing_allowed && !c.forward_preference) {
          // where(yaws < M_PI_2, ybp, normalize_angles(ybp + M_PI)) is double; to - from is double - float
          const double corrected = yaws < M_PI_2 ? static_cast<double>(ybp) :
            utils_normalize_angles_d(static_cast<double>(ybp) + M_PI);
          const double diff = corrected - static_cast<double>(tr.yaws(b, t));
          sum += std::fabs(utils_normalize_angles_d(diff));
        } else {
          sum += yaws;
        }
      }
      addPow(*data.costs, b, (sum / static_cast<double>(T)) * c.weight, c.d.cost_power);
    }
  }
  static double utils_normalize_angles_d(double angle)
  {
    const double theta = std::fmod(angle + M_PI, 2.0 * M_PI);
    return theta <= 0.0 ? theta + M_PI : theta - M_PI;
  }

  // shared head of both PathAlign critics (ref: path_align_critic.cpp:46-72, path_align_legacy_critic.cpp:46-72)
  bool pathAlignGate(const Critic & c, CriticData & data) const
  {
    const State & s = *data.state;
    if (!c.d.enabled || withinPositionGoalTolerance(c.d.threshold_to_consider, s.pose_x, s.pose_y, data.goal_x, data.goal_y)) {
      return false;
    }
    setPathFurthestPointIfNotSet(data);
    if (data.furthest_reached_path_point < static_cast<size_t>(c.d.offset_from_furthest)) {
      return false;
    }
    setPathCostsIfNotSet(data, costmap, robot.track_unknown != 0);
    const size_t closest_initial_path_point = findPathTrajectoryInitialPoint(data);
    unsigned int invalid_ctr = 0;
    const float range = data.furthest_reached_path_point - closest_initial_path_point;
    for (size_t i = closest_initial_path_point; i < data.furthest_reached_path_point; i++) {
      if (!data.path_pts_valid[i]) {invalid_ctr++;}
      if (static_cast<float>(invalid_ctr) / range > c.d.max_path_occupancy_ratio && invalid_ctr > 2) {
        return false;
      }
    }
    return true;
  }

  // ref: PathAlignCritic::score path_align_critic.cpp:46-136
  void scorePathAlign(const Critic & c, CriticData & data) const
  {
    if (!pathAlignGate(c, data)) {return;}
    const size_t path_segments_count = data.furthest_reached_path_point;
    if (path_segments_count == 0) {return;}   // reference: unsigned wrap-around loop (UB); defined as no-op
    const Path & path = *data.path;
    const Trajectories & tr = *data.trajectories;
    const size_t batch_size = tr.x.B, time_steps = tr.x.T;
    const size_t step = c.d.trajectory_point_step;
    std::vector<float> path_integrated_distances(path_segments_count, 0.0f);
    float dx = 0.0f, dy = 0.0f;
    for (unsigned int i = 1; i != path_segments_count; i++) {
      dx = path.x[i] - path.x[i - 1];
      dy = path.y[i] - path.y[i - 1];
      const float curr_dist = sqrtf(dx * dx + dy * dy);
      path_integrated_distances[i] = path_integrated_distances[i - 1] + curr_dist;
    }
    for (size_t t = 0; t < batch_size; ++t) {
      float traj_integrated_distance = 0.0f;
      float summed_path_dist = 0.0f, dyaw = 0.0f;
      float num_samples = 0.0f;
      size_t path_pt = 0u;
      for (size_t p = step; p < time_steps; p += step) {
        const float Tx = tr.x(t, p), Ty = tr.y(t, p);
        dx = Tx - tr.x(t, p - step);
        dy = Ty - tr.y(t, p - step);
        traj_integrated_distance += sqrtf(dx * dx + dy * dy);
        path_pt = findClosestPathPt(path_integrated_distances, traj_integrated_distance, path_pt);
        if (data.path_pts_valid[path_pt]) {
          dx = path.x[path_pt] - Tx;
          dy = path.y[path_pt] - Ty;
          num_samples += 1.0f;
          if (c.d.use_path_orientations) {
            dyaw = shortest_angular_distance(path.yaws[path_pt], tr.yaws(t, p));
            summed_path_dist += sqrtf(dx * dx + dy * dy + dyaw * dyaw);
          } else {
            summed_path_dist += sqrtf(dx * dx + dy * dy);
          }
        }
      }
      const float cost = num_samples > 0 ? summed_path_dist / num_samples : 0.0f;
      addPow(*data.costs, t, cost * c.weight, c.d.cost_power);
    }
  }

  // ref: PathAlignLegacyCritic::score path_align_legacy_critic.cpp:46-129
  void scorePathAlignLegacy(const Critic & c, CriticData & data) const
  {
    if (!pathAlignGate(c, data)) {return;}
    const Path & path = *data.path;
    const Trajectories & tr = *data.trajectories;
    const size_t batch_size = tr.x.B, time_steps = tr.x.T;
    const size_t step = c.d.trajectory_point_step;
    const size_t traj_pts_eval = floor(time_steps / step);
    const size_t path_segments_count = path.x.size() - 1;
    if (path_segments_count < 1) {return;}
    for (size_t t = 0; t < batch_size; ++t) {
      float summed_dist = 0.0f;
      for (size_t p = step; p < time_steps; p += step) {
        float min_dist_sq = std::numeric_limits<float>::max();
        size_t min_s = 0;
        for (size_t sg = 0; sg < path_segments_count - 1; sg++) {
          const float dx = path.x[sg] - tr.x(t, p);
          const float dy = path.y[sg] - tr.y(t, p);
          float dist_sq;
          if (c.d.use_path_orientations) {
            const float dyaw = shortest_angular_distance(path.yaws[sg], tr.yaws(t, p));
            dist_sq = dx * dx + dy * dy + dyaw * dyaw;
          } else {
            dist_sq = dx * dx + dy * dy;
          }
          if (dist_sq < min_dist_sq) {
            min_dist_sq = dist_sq;
            min_s = sg;
          }
        }
        if (min_s != 0 && data.path_pts_valid[min_s]) {
          summed_dist += sqrtf(min_dist_sq);
        }
      }
      const float cost = summed_dist / traj_pts_eval;
      addPow(*data.costs, t, cost * c.weight, c.d.cost_power);
    }
  }

  // ref: CostCritic::inCollision cost_critic.cpp:175-199 / ObstaclesCritic::inCollision obstacles_critic.cpp:185-201
  bool inCollisionValue(float cost, bool consider_footprint) const
  {
    switch (static_cast<unsigned char>(cost)) {
      case LETHAL_OBSTACLE: return true;
      case INSCRIBED_INFLATED_OBSTACLE: return consider_footprint ? false : true;
      case NO_INFORMATION: return robot.track_unknown ? false : true;
    }
    return false;
  }

  // ref: CostCritic::score cost_critic.cpp:108-168, costAtPose :201-210
  void scoreCost(Critic & c, CriticData & data) const
  {
    if (!c.d.enabled) {return;}
    const bool consider_footprint = c.d.consider_footprint != 0;
    c.possibly_inscribed_cost = findCircumscribedCost();
    const State & s = *data.state;
    const bool near_goal = withinPositionGoalTolerance(c.d.near_goal_distance, s.pose_x, s.pose_y, data.goal_x, data.goal_y);
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, traj_len = tr.x.T;
    FootprintCollisionChecker checker{&costmap};
    std::vector<float> repulsive_cost(B, 0.0f);
    bool all_trajectories_collide = true;
    for (size_t i = 0; i < B; ++i) {
      bool trajectory_collide = false;
      for (size_t j = 0; j < traj_len; j++) {
        const float x = tr.x(i, j), y = tr.y(i, j);
        unsigned int x_i, y_i;
        float pose_cost;
        if (!costmap.worldToMap(x, y, x_i, y_i)) {
          pose_cost = NO_INFORMATION;
        } else {
          pose_cost = checker.pointCost(x_i, y_i);
        }
        ++counters[0];
        if (pose_cost < 1.0f) {continue;}
        float cost = pose_cost;
        if (consider_footprint && (cost >= c.possibly_inscribed_cost || c.possibly_inscribed_cost < 1.0f)) {
          cost = static_cast<float>(checker.footprintCostAtPose(x, y, tr.yaws(i, j), robot));
          ++counters[1];
        }
        if (inCollisionValue(cost, consider_footprint)) {
          trajectory_collide = true;
          break;
        }
        if (pose_cost >= INSCRIBED_INFLATED_OBSTACLE) {
          repulsive_cost[i] += c.d.critical_cost;
        } else if (!near_goal) {
          repulsive_cost[i] += pose_cost;
        }
      }
      if (!trajectory_collide) {
        all_trajectories_collide = false;
      } else {
        repulsive_cost[i] = c.d.collision_cost;
      }
    }
    for (size_t i = 0; i < B; ++i) {
      addPow(*data.costs, i, c.weight * repulsive_cost[i] / traj_len, c.d.cost_power);
    }
    data.fail_flag = all_trajectories_collide;
  }

  // ref: ObstaclesCritic::score obstacles_critic.cpp:114-178, costAtPose :203-224, distanceToObstacle :99-112
  void scoreObstacles(Critic & c, CriticData & data) const
  {
    if (!c.d.enabled) {return;}
    const bool consider_footprint = c.d.consider_footprint != 0;
    c.possibly_inscribed_cost = findCircumscribedCost();
    const State & s = *data.state;
    const bool near_goal = withinPositionGoalTolerance(c.d.near_goal_distance, s.pose_x, s.pose_y, data.goal_x, data.goal_y);
    const Trajectories & tr = *data.trajectories;
    const size_t B = tr.x.B, traj_len = tr.x.T;
    FootprintCollisionChecker checker{&costmap};
    std::vector<float> raw_cost(B, 0.0f), repulsive_cost(B, 0.0f);
    bool all_trajectories_collide = true;
    const float scale_factor = c.inflation_scale_factor;
    const float min_radius = robot.inscribed_radius;
    for (size_t i = 0; i < B; ++i) {
      bool trajectory_collide = false;
      float traj_cost = 0.0f;
      for (size_t j = 0; j < traj_len; j++) {
        const float x = tr.x(i, j), y = tr.y(i, j);
        float cost;
        bool using_footprint = false;
        unsigned int x_i, y_i;
        ++counters[2];
        if (!costmap.worldToMap(x, y, x_i, y_i)) {
          cost = NO_INFORMATION;
        } else {
          cost = checker.pointCost(x_i, y_i);
          if (consider_footprint && (cost >= c.possibly_inscribed_cost || c.possibly_inscribed_cost < 1.0f)) {
            cost = static_cast<float>(checker.footprintCostAtPose(x, y, tr.yaws(i, j), robot));
            using_footprint = true;
            ++counters[3];
          }
        }
        if (cost < 1.0f) {continue;}
        if (inCollisionValue(cost, consider_footprint)) {
          trajectory_collide = true;
          break;
        }
        if (c.inflation_radius == 0.0f || c.inflation_scale_factor == 0.0f) {continue;}
        float dist_to_obj = (scale_factor * min_radius - logf(cost) + logf(253.0f)) / scale_factor;
        if (!using_footprint) {dist_to_obj -= min_radius;}
        if (dist_to_obj < c.d.collision_margin_distance) {
          traj_cost += (c.d.collision_margin_distance - dist_to_obj);
        } else if (!near_goal) {
          repulsive_cost[i] += (c.inflation_radius - dist_to_obj);
        }
      }
      if (!trajectory_collide) {all_trajectories_collide = false;}
      raw_cost[i] = trajectory_collide ? c.d.collision_cost : traj_cost;
    }
    for (size_t i = 0; i < B; ++i) {
      const float v = (c.d.critical_weight * raw_cost[i]) + (c.d.repulsion_weight * repulsive_cost[i] / traj_len);
      addPow(*data.costs, i, v, c.d.cost_power);
    }
    data.fail_flag = all_trajectories_collide;
  }

  // ref: CriticManager::evalTrajectoriesScores critic_manager.cpp:67-76
  void evalTrajectoriesScores(CriticData & data)
  {
    const size_t B = data.costs->size();
    critic_costs.assign(critics.size(), std::vector<float>(B, 0.0f));
    for (size_t q = 0; q < critics.size(); q++) {
      if (data.fail_flag) {break;}
      const std::vector<float> before = *data.costs;
      Critic & c = critics[q];
      switch (c.d.kind) {
        case MPPI_CRITIC_CONSTRAINT: scoreConstraint(c, data); break;
        case MPPI_CRITIC_COST: scoreCost(c, data); break;
        case MPPI_CRITIC_GOAL: scoreGoal(c, data); break;
        case MPPI_CRITIC_GOAL_ANGLE: scoreGoalAngle(c, data); break;
        case MPPI_CRITIC_OBSTACLES: scoreObstacles(c, data); break;
        case MPPI_CRITIC_PATH_ALIGN: scorePathAlign(c, data); break;
        case MPPI_CRITIC_PATH_ALIGN_LEGACY: scorePathAlignLegacy(c, data); break;
        case MPPI_CRITIC_PATH_ANGLE: scorePathAngle(c, data); break;
        case MPPI_CRITIC_PATH_FOLLOW: scorePathFollow(c, data); break;
        case MPPI_CRITIC_PREFER_FORWARD: scorePreferForward(c, data); break;
        case MPPI_CRITIC_TWIRLING: scoreTwirling(c, data); break;
        case MPPI_CRITIC_VELOCITY_DEADBAND: scoreVelocityDeadband(c, data); break;
        default: break;
      }
      // per-critic contribution, recorded in double so that tests can compare the critic's own term
      for (size_t b = 0; b < B; ++b) {
        critic_costs[q][b] = static_cast<float>(static_cast<double>((*data.costs)[b]) - static_cast<double>(before[b]));
      }
    }
  }

  void computeCells()
  {
    const size_t B = traj.x.B, T = traj.x.T;
    cells.assign(B * T, -1);
    for (size_t b = 0; b < B; ++b) {
      for (size_t t = 0; t < T; ++t) {
        unsigned int mx, my;
        if (costmap.worldToMap(traj.x(b, t), traj.y(b, t), mx, my)) {
          cells[b * T + t] = static_cast<int32_t>(my * costmap.size_x + mx);
        }
      }
    }
  }

  void loadCycle(const mppi_cycle_in & in)
  {
    state.pose_x = in.pose_x; state.pose_y = in.pose_y; state.pose_yaw = in.pose_yaw;
    state.speed_vx = in.speed_vx; state.speed_vy = in.speed_vy; state.speed_wz = in.speed_wz;
    path.x.assign(in.path_x, in.path_x + in.path_size);
    path.y.assign(in.path_y, in.path_y + in.path_size);
    path.yaws.assign(in.path_yaw, in.path_yaw + in.path_size);
    costmap.size_x = in.costmap.size_x; costmap.size_y = in.costmap.size_y;
    costmap.resolution = in.costmap.resolution;
    costmap.origin_x = in.costmap.origin_x; costmap.origin_y = in.costmap.origin_y;
    costmap.cells.assign(in.costmap.cells, in.costmap.cells + static_cast<size_t>(in.costmap.size_x) * in.costmap.size_y);
  }

  CriticData makeData(const mppi_cycle_in & in, const State * st, const Trajectories * tr, std::vector<float> * c)
  {
    CriticData d;
    d.state = st; d.trajectories = tr; d.path = &path;
    d.goal_x = in.goal_x; d.goal_y = in.goal_y;
    d.costs = c; d.model_dt = cfg.model_dt;
    d.fail_flag = false;
    d.goal_checker_xy_tolerance = in.goal_checker_xy_tolerance;
    d.motion_model = cfg.motion_model;
    d.ackermann_min_turning_r = cfg.ackermann_min_turning_r;
    return d;
  }

  // ref: Optimizer::prepare optimizer.cpp:185-204 + Optimizer::optimize :157-164
  void optimize(const mppi_cycle_in & in)
  {
    loadCycle(in);
    counters[0] = counters[1] = counters[2] = counters[3] = 0;
    std::fill(costs.begin(), costs.end(), 0.0f);
    data = makeData(in, &state, &traj, &costs);
    for (int i = 0; i < cfg.iteration_count; ++i) {
      // generateNoisedTrajectories optimizer.cpp:227-233
      setNoisedControls();
      updateStateVelocities(state, isHolonomic());
      integrateStateVelocities(traj, state, cfg.model_dt, isHolonomic());
      evalTrajectoriesScores(data);
      updateControlSequence();
      if (i == pin_iteration) {cs = pin_cs;}
      // ref: NoiseGenerator::generateNextNoises noise_generator.cpp:54-63 + noiseThread :97-105: with regenerate_noises the
      // side thread redraws once the current set has been consumed.  The reference races that thread against the next
      // iteration; the deterministic restatement is "every iteration consumes a fresh set" (Philox stream + 1 each time).
      if (cfg.regenerate_noises) {
        generateNoisedControls();
        noise_stream++;
      }
    }
    computeCells();
  }
};

}  // namespace oracle

// =================================================================================================
// C entry points, mirroring include/mppi_b200.h one to one (prefix oracle_)
// =================================================================================================
using oracle::Optimizer;

extern "C" {

void oracle_config_default(mppi_config * c)
{
  std::memset(c, 0, sizeof(*c));
  c->batch_size = 1000; c->time_steps = 56; c->iteration_count = 1;
  c->model_dt = 0.05f; c->temperature = 0.3f; c->gamma = 0.015f;
  c->vx_max = 0.5; c->vx_min = -0.35; c->vy_max = 0.5; c->wz_max = 1.9;
  c->vx_std = 0.2; c->vy_std = 0.2; c->wz_std = 0.4;
  c->motion_model = MPPI_MODEL_DIFF_DRIVE;
  c->ackermann_min_turning_r = 0.2;
  c->regenerate_noises = 0; c->seed = 0; c->device = 0; c->shard_offset = 0; c->shard_total = 0;
}

void oracle_critic_default(int32_t kind, mppi_critic_desc * d)
{
  std::memset(d, 0, sizeof(*d));
  d->kind = kind; d->enabled = 1; d->cost_power = 1;
  d->trajectory_point_step = 4; d->max_path_occupancy_ratio = 0.07; d->use_path_orientations = 0;
  d->max_angle_to_furthest = 1.2; d->forward_preference = 1; d->consider_footprint = 0;
  d->near_goal_distance = 0.5; d->repulsion_weight = 1.5; d->critical_weight = 20.0;
  d->collision_margin_distance = 0.10; d->cost_scaling_factor = 10.0; d->inflation_radius = 0.55;
  d->critical_cost = 300.0;
  switch (kind) {
    case MPPI_CRITIC_CONSTRAINT: d->cost_weight = 4.0; break;
    case MPPI_CRITIC_COST: d->cost_weight = 3.81; d->collision_cost = 1000000.0; break;
    case MPPI_CRITIC_GOAL: d->cost_weight = 5.0; d->threshold_to_consider = 1.4; break;
    case MPPI_CRITIC_GOAL_ANGLE: d->cost_weight = 3.0; d->threshold_to_consider = 0.5; break;
    case MPPI_CRITIC_OBSTACLES: d->collision_cost = 10000.0; break;
    case MPPI_CRITIC_PATH_ALIGN:
    case MPPI_CRITIC_PATH_ALIGN_LEGACY: d->cost_weight = 10.0; d->threshold_to_consider = 0.5; d->offset_from_furthest = 20; break;
    case MPPI_CRITIC_PATH_ANGLE: d->cost_weight = 2.0; d->threshold_to_consider = 0.5; d->offset_from_furthest = 4; break;
    case MPPI_CRITIC_PATH_FOLLOW: d->cost_weight = 5.0; d->threshold_to_consider = 1.4; d->offset_from_furthest = 6; break;
    case MPPI_CRITIC_PREFER_FORWARD: d->cost_weight = 5.0; d->threshold_to_consider = 0.5; break;
    case MPPI_CRITIC_TWIRLING: d->cost_weight = 10.0; break;
    case MPPI_CRITIC_VELOCITY_DEADBAND: d->cost_weight = 35.0; break;
    default: break;
  }
}

int oracle_create(const mppi_config * cfg, Optimizer ** out)
{
  if (!cfg || !out || cfg->batch_size <= 0 || cfg->time_steps <= 0) {return MPPI_E_CONFIG;}
  // ref: Optimizer::setMotionModel optimizer.cpp:412-426 throws for anything but DiffDrive / Omni / Ackermann
  if (cfg->motion_model < MPPI_MODEL_DIFF_DRIVE || cfg->motion_model > MPPI_MODEL_ACKERMANN) {return MPPI_E_CONFIG;}
  Optimizer * o = new Optimizer();
  o->cfg = *cfg;
  o->base_constraints = {cfg->vx_max, cfg->vx_min, cfg->vy_max, cfg->wz_max};
  std::memset(&o->robot, 0, sizeof(o->robot));
  o->reset();
  *out = o;
  return MPPI_OK;
}
void oracle_destroy(Optimizer * o) {delete o;}
int oracle_reset(Optimizer * o) {o->reset(); return MPPI_OK;}

int oracle_set_critics(Optimizer * o, const mppi_critic_desc * critics, int32_t n)
{
  o->critics.clear();
  for (int i = 0; i < n; ++i) {
    oracle::Critic c;
    c.d = critics[i];
    o->initializeCritic(c);
    o->critics.push_back(c);
  }
  return MPPI_OK;
}
int oracle_set_robot(Optimizer * o, const mppi_robot_desc * r)
{
  o->robot = *r;
  for (auto & c : o->critics) {o->initializeCritic(c);}
  return MPPI_OK;
}

// ref: Optimizer::setSpeedLimit optimizer.cpp:428-453
int oracle_set_speed_limit(Optimizer * o, double speed_limit, int32_t percentage)
{
  auto & s = o->constraints;
  const auto & b = o->base_constraints;
  if (speed_limit == 0.0) {   // nav2_costmap_2d::NO_SPEED_LIMIT
    s = b;
  } else {
    const double ratio = percentage ? speed_limit / 100.0 : speed_limit / b.vx_max;
    s.vx_max = b.vx_max * ratio; s.vx_min = b.vx_min * ratio; s.vy = b.vy * ratio; s.wz = b.wz * ratio;
  }
  return MPPI_OK;
}
int oracle_get_constraints(const Optimizer * o, float out4[4])
{
  out4[0] = o->constraints.vx_max; out4[1] = o->constraints.vx_min; out4[2] = o->constraints.vy; out4[3] = o->constraints.wz;
  return MPPI_OK;
}

int oracle_set_noise(Optimizer * o, const float * vx, const float * vy, const float * wz)
{
  const size_t n = static_cast<size_t>(o->cfg.batch_size) * o->cfg.time_steps;
  o->noises_vx.v.assign(vx, vx + n);
  if (vy) {o->noises_vy.v.assign(vy, vy + n);} else {o->noises_vy.v.assign(n, 0.0f);}
  o->noises_wz.v.assign(wz, wz + n);
  return MPPI_OK;
}
int oracle_generate_noise(Optimizer * o, uint64_t stream)
{
  o->noise_stream = stream;
  o->generateNoisedControls();
  o->noise_stream = stream + 1;   // a later reset() draws the next stream
  return MPPI_OK;
}
int oracle_get_noise(Optimizer * o, float * vx, float * vy, float * wz)
{
  const size_t n = o->noises_vx.v.size();
  std::memcpy(vx, o->noises_vx.v.data(), n * 4); std::memcpy(vy, o->noises_vy.v.data(), n * 4);
  std::memcpy(wz, o->noises_wz.v.data(), n * 4);
  return MPPI_OK;
}

int oracle_set_control_sequence(Optimizer * o, const float * vx, const float * vy, const float * wz)
{
  const size_t T = o->cfg.time_steps;
  o->cs.vx.assign(vx, vx + T); o->cs.vy.assign(vy, vy + T); o->cs.wz.assign(wz, wz + T);
  return MPPI_OK;
}
int oracle_get_control_sequence(Optimizer * o, float * vx, float * vy, float * wz)
{
  const size_t T = o->cfg.time_steps;
  std::memcpy(vx, o->cs.vx.data(), T * 4); std::memcpy(vy, o->cs.vy.data(), T * 4); std::memcpy(wz, o->cs.wz.data(), T * 4);
  return MPPI_OK;
}
// ref: Optimizer::shiftControlSequence optimizer.cpp:206-225
int oracle_shift_control_sequence(Optimizer * o)
{
  auto roll = [](std::vector<float> & v) {
      if (v.size() < 2) {return;}
      std::rotate(v.begin(), v.begin() + 1, v.end());
      v[v.size() - 1] = v[v.size() - 2];
    };
  roll(o->cs.vx); roll(o->cs.wz);
  if (o->isHolonomic()) {roll(o->cs.vy);}
  return MPPI_OK;
}

static void oracle_shift_impl(Optimizer * o)
{
  auto roll = [](std::vector<float> & v) {
      if (v.size() < 2) {return;}
      std::rotate(v.begin(), v.begin() + 1, v.end());
      v[v.size() - 1] = v[v.size() - 2];
    };
  roll(o->cs.vx); roll(o->cs.wz);
  if (o->isHolonomic()) {roll(o->cs.vy);}
}

int oracle_optimize(Optimizer * o, const mppi_cycle_in * in, mppi_cycle_out * out);

// ref: Optimizer::evalControl optimizer.cpp:134-155, one attempt (the fallback loop is the caller's): optimize(); if it
// did not fail: savitskyGolayFilter, getControlFromSequenceAsTwist (:396-410), shiftControlSequence (:206-225)
int oracle_eval_control(Optimizer * o, const mppi_cycle_in * in, int32_t shift_control_sequence, mppi_cycle_out * out, float cmd_out[3])
{
  mppi_cycle_out tmp;
  std::memset(&tmp, 0, sizeof(tmp));
  oracle_optimize(o, in, &tmp);
  float cmd[3] = {0.0f, 0.0f, 0.0f};
  if (!tmp.fail_flag) {
    oracle::savitskyGolayFilter(o->cs, o->control_history, shift_control_sequence != 0);
    const unsigned offset = shift_control_sequence ? 1 : 0;
    cmd[0] = o->cs.vx[offset]; cmd[2] = o->cs.wz[offset];
    cmd[1] = o->isHolonomic() ? o->cs.vy[offset] : 0.0f;
    if (shift_control_sequence) {oracle_shift_impl(o);}
  }
  const size_t T = o->cfg.time_steps;
  if (out) {
    if (out->control_vx) {std::memcpy(out->control_vx, o->cs.vx.data(), T * 4);}
    if (out->control_vy) {std::memcpy(out->control_vy, o->cs.vy.data(), T * 4);}
    if (out->control_wz) {std::memcpy(out->control_wz, o->cs.wz.data(), T * 4);}
    out->fail_flag = tmp.fail_flag;
    out->furthest_reached_path_point = tmp.furthest_reached_path_point;
    out->device_ms = 0.0f;
  }
  if (cmd_out) {std::memcpy(cmd_out, cmd, sizeof(cmd));}
  return MPPI_OK;
}
int oracle_set_control_history(Optimizer * o, const float hist12[12]) {std::memcpy(o->control_history, hist12, 48); return MPPI_OK;}
int oracle_get_control_history(Optimizer * o, float hist12[12]) {std::memcpy(hist12, o->control_history, 48); return MPPI_OK;}

int oracle_optimize(Optimizer * o, const mppi_cycle_in * in, mppi_cycle_out * out)
{
  o->optimize(*in);
  const size_t T = o->cfg.time_steps;
  if (out) {
    if (out->control_vx) {std::memcpy(out->control_vx, o->cs.vx.data(), T * 4);}
    if (out->control_vy) {std::memcpy(out->control_vy, o->cs.vy.data(), T * 4);}
    if (out->control_wz) {std::memcpy(out->control_wz, o->cs.wz.data(), T * 4);}
    out->fail_flag = o->data.fail_flag ? 1 : 0;
    out->furthest_reached_path_point = o->data.furthest_set ? static_cast<uint32_t>(o->data.furthest_reached_path_point) : UINT32_MAX;
    out->device_ms = 0.0f;
  }
  return MPPI_OK;
}

int oracle_get_trajectories(Optimizer * o, float * x, float * y, float * yaw)
{
  const size_t n = o->traj.x.v.size();
  std::memcpy(x, o->traj.x.v.data(), n * 4); std::memcpy(y, o->traj.y.v.data(), n * 4); std::memcpy(yaw, o->traj.yaws.v.data(), n * 4);
  return MPPI_OK;
}
int oracle_set_wide_reductions(Optimizer * o, int32_t on)
{
  o->wide_reductions = on != 0;
  return MPPI_OK;
}
int oracle_get_counters(Optimizer * o, uint64_t * out4)
{
  std::memcpy(out4, o->counters, sizeof(o->counters));
  return MPPI_OK;
}
int oracle_set_iteration_controls(Optimizer * o, int32_t iteration, const float * vx, const float * vy, const float * wz)
{
  o->pin_iteration = iteration;
  if (iteration >= 0) {
    const size_t T = o->cfg.time_steps;
    o->pin_cs.vx.assign(vx, vx + T); o->pin_cs.vy.assign(vy, vy + T); o->pin_cs.wz.assign(wz, wz + T);
  }
  return MPPI_OK;
}
int oracle_get_state(Optimizer * o, float * vx, float * vy, float * wz, float * cvx, float * cvy, float * cwz)
{
  const size_t n = o->state.vx.v.size();
  std::memcpy(vx, o->state.vx.v.data(), n * 4); std::memcpy(vy, o->state.vy.v.data(), n * 4); std::memcpy(wz, o->state.wz.v.data(), n * 4);
  std::memcpy(cvx, o->state.cvx.v.data(), n * 4); std::memcpy(cvy, o->state.cvy.v.data(), n * 4); std::memcpy(cwz, o->state.cwz.v.data(), n * 4);
  return MPPI_OK;
}
int oracle_get_cells(Optimizer * o, int32_t * cells)
{
  std::memcpy(cells, o->cells.data(), o->cells.size() * 4);
  return MPPI_OK;
}
int oracle_get_costs(Optimizer * o, float * costs)
{
  std::memcpy(costs, o->costs.data(), o->costs.size() * 4);
  return MPPI_OK;
}
int oracle_get_critic_costs(Optimizer * o, int32_t index, float * costs)
{
  if (index < 0 || static_cast<size_t>(index) >= o->critic_costs.size()) {return MPPI_E_CONFIG;}
  std::memcpy(costs, o->critic_costs[index].data(), o->critic_costs[index].size() * 4);
  return MPPI_OK;
}

// ref: Optimizer::getOptimizedTrajectory optimizer.cpp:345-360 + integrateStateVelocities(xtensor2&, ...) :275-311
int oracle_get_optimized_trajectory(Optimizer * o, double pose_x, double pose_y, double pose_yaw, float * traj_t3)
{
  const size_t T = o->cfg.time_steps;
  const float initial_yaw = pose_yaw;
  const float dt = o->cfg.model_dt;
  std::vector<float> yaws(T);
  float acc = 0.0f;
  for (size_t t = 0; t < T; ++t) {
    const float term = o->cs.wz[t] * dt;
    acc = t == 0 ? term : acc + term;
    yaws[t] = acc + initial_yaw;
  }
  float accx = 0.0f, accy = 0.0f;
  for (size_t t = 0; t < T; ++t) {
    float yc, ys;
    // NOT the one-step lag of the batch overload (:322-329): this overload pairs cos/sin[1:] with yaws[1:]
    // (`yaw_offseted = view(traj_yaws, range(1, _))`, optimizer.cpp:294-299), so step t >= 1 turns by its OWN yaw
    mppi_det_sincosf(t == 0 ? initial_yaw : yaws[t], &ys, &yc);
    float dx = o->cs.vx[t] * yc, dy = o->cs.vx[t] * ys;
    if (o->isHolonomic()) {
      dx = dx - o->cs.vy[t] * ys;
      dy = dy + o->cs.vy[t] * yc;
    }
    const float tx = dx * dt, ty = dy * dt;
    accx = t == 0 ? tx : accx + tx;
    accy = t == 0 ? ty : accy + ty;
    traj_t3[t * 3 + 0] = static_cast<float>(pose_x + static_cast<double>(accx));
    traj_t3[t * 3 + 1] = static_cast<float>(pose_y + static_cast<double>(accy));
    traj_t3[t * 3 + 2] = yaws[t];
  }
  return MPPI_OK;
}

int oracle_integrate_state_velocities(
  Optimizer * o, double pose_x, double pose_y, double pose_yaw, const float * vx, const float * vy,
  const float * wz, float * x, float * y, float * yaw)
{
  const size_t B = o->cfg.batch_size, T = o->cfg.time_steps, n = B * T;
  oracle::State s;
  s.reset(B, T);
  s.pose_x = pose_x; s.pose_y = pose_y; s.pose_yaw = pose_yaw;
  s.vx.v.assign(vx, vx + n); s.vy.v.assign(vy, vy + n); s.wz.v.assign(wz, wz + n);
  oracle::Trajectories tr;
  Optimizer::integrateStateVelocities(tr, s, o->cfg.model_dt, o->isHolonomic());
  std::memcpy(x, tr.x.v.data(), n * 4); std::memcpy(y, tr.y.v.data(), n * 4); std::memcpy(yaw, tr.yaws.v.data(), n * 4);
  return MPPI_OK;
}

int oracle_score_trajectories(
  Optimizer * o, const mppi_cycle_in * in, const float * vx, const float * vy, const float * wz,
  const float * x, const float * y, const float * yaw, float * costs_inout, uint32_t * furthest_inout,
  int32_t * fail_flag_out)
{
  const size_t B = o->cfg.batch_size, T = o->cfg.time_steps, n = B * T;
  o->loadCycle(*in);
  oracle::State s;
  s.reset(B, T);
  s.pose_x = in->pose_x; s.pose_y = in->pose_y; s.pose_yaw = in->pose_yaw;
  s.speed_vx = in->speed_vx; s.speed_vy = in->speed_vy; s.speed_wz = in->speed_wz;
  s.vx.v.assign(vx, vx + n); s.vy.v.assign(vy, vy + n); s.wz.v.assign(wz, wz + n);
  oracle::Trajectories tr;
  tr.reset(B, T);
  tr.x.v.assign(x, x + n); tr.y.v.assign(y, y + n); tr.yaws.v.assign(yaw, yaw + n);
  std::vector<float> costs(costs_inout, costs_inout + B);
  oracle::CriticData d = o->makeData(*in, &s, &tr, &costs);
  if (furthest_inout && *furthest_inout != UINT32_MAX) {
    d.furthest_set = true;
    d.furthest_reached_path_point = *furthest_inout;
  }
  o->evalTrajectoriesScores(d);
  std::memcpy(costs_inout, costs.data(), B * 4);
  if (furthest_inout) {*furthest_inout = d.furthest_set ? static_cast<uint32_t>(d.furthest_reached_path_point) : UINT32_MAX;}
  if (fail_flag_out) {*fail_flag_out = d.fail_flag ? 1 : 0;}
  return MPPI_OK;
}

// ---- helpers exposed for the host-logic and utils tests ----
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {oracle::philox4x32_10(ctr, key, out);}
void oracle_det_sincosf(float x, float * s, float * c) {mppi_det_sincosf(x, s, c);}
double oracle_normalize_angle(double a) {return oracle::normalize_angle(a);}
double oracle_utils_normalize_angles(float a) {return oracle::utils_normalize_angles(a);}
double oracle_utils_shortest_angular_distance(float from, float to) {return oracle::utils_shortest_angular_distance(from, to);}
float oracle_pose_point_angle(double px, double py, double pyaw, double x, double y, int32_t forward_preference)
{
  return oracle::posePointAngle(px, py, pyaw, x, y, forward_preference != 0);
}
int32_t oracle_within_tolerance_checker(double tol, double rx, double ry, double gx, double gy)
{
  return oracle::withinGoalCheckerTolerance(tol, rx, ry, gx, gy) ? 1 : 0;
}
int32_t oracle_within_tolerance(float tol, double rx, double ry, double gx, double gy)
{
  return oracle::withinPositionGoalTolerance(tol, rx, ry, gx, gy) ? 1 : 0;
}
// tf2::getYaw (tf2/impl/utils.h)
double oracle_get_yaw(double qx, double qy, double qz, double qw)
{
  const double sqx = qx * qx, sqy = qy * qy, sqz = qz * qz, sqw = qw * qw;
  const double sarg = -2 * (qx * qz - qw * qy) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999) {return -2 * atan2(qy, qx);}
  if (sarg >= 0.99999) {return 2 * atan2(qy, qx);}
  return atan2(2 * (qx * qy + qw * qz), sqw + sqx - sqy - sqz);
}
size_t oracle_find_closest_path_pt(const float * vec, size_t n, float dist, size_t init)
{
  return oracle::findClosestPathPt(std::vector<float>(vec, vec + n), dist, init);
}
void oracle_savitsky_golay(float * vx, float * vy, float * wz, int32_t T, float history[12], int32_t shift)
{
  oracle::ControlSequence cs;
  cs.vx.assign(vx, vx + T); cs.vy.assign(vy, vy + T); cs.wz.assign(wz, wz + T);
  float h[4][3];
  std::memcpy(h, history, sizeof(h));
  oracle::savitskyGolayFilter(cs, h, shift != 0);
  std::memcpy(history, h, sizeof(h));
  std::memcpy(vx, cs.vx.data(), T * 4); std::memcpy(vy, cs.vy.data(), T * 4); std::memcpy(wz, cs.wz.data(), T * 4);
}
// findPathCosts on a caller-provided map (utils_test.cpp:288-322)
void oracle_find_path_costs(const mppi_cycle_in * in, int32_t track_unknown, uint8_t * valid_out)
{
  Optimizer o;
  oracle_config_default(&o.cfg);
  o.loadCycle(*in);
  oracle::CriticData d;
  d.path = &o.path;
  oracle::findPathCosts(d, o.costmap, track_unknown != 0);
  for (size_t i = 0; i < d.path_pts_valid.size(); ++i) {valid_out[i] = d.path_pts_valid[i] ? 1 : 0;}
}
uint8_t oracle_inflation_compute_cost(double distance_cells, double resolution, double inscribed_radius, double scale)
{
  return oracle::inflationComputeCost(distance_cells, resolution, inscribed_radius, scale);
}
double oracle_footprint_cost_at_pose(const mppi_costmap * cm, const mppi_robot_desc * robot, double x, double y, double theta)
{
  oracle::Costmap c;
  c.size_x = cm->size_x; c.size_y = cm->size_y; c.resolution = cm->resolution; c.origin_x = cm->origin_x; c.origin_y = cm->origin_y;
  c.cells.assign(cm->cells, cm->cells + static_cast<size_t>(cm->size_x) * cm->size_y);
  oracle::FootprintCollisionChecker ch{&c};
  return ch.footprintCostAtPose(x, y, theta, *robot);
}
int32_t oracle_world_to_map(const mppi_costmap * cm, double wx, double wy)
{
  oracle::Costmap c;
  c.size_x = cm->size_x; c.size_y = cm->size_y; c.resolution = cm->resolution; c.origin_x = cm->origin_x; c.origin_y = cm->origin_y;
  unsigned int mx, my;
  if (!c.worldToMap(wx, wy, mx, my)) {return -1;}
  return static_cast<int32_t>(my * cm->size_x + mx);
}

}  // extern "C"
