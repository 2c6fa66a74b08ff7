// mppi_optimizer.hpp -- C++ host mirror of sortham::Optimizer over the C ABI of mppi_b200.h.
//
// Header-only, C++17, no ROS: this is the class a SORTHAMController shim holds instead of the reference's
// sortham::Optimizer (include/nav2_sortham_controller/optimizer.hpp:51-263).  Same member names, same argument meaning,
// same error behaviour (std::runtime_error where the reference throws); ROS message types are replaced by the plain
// structs below and the shim converts (INTEGRATION.md shows the conversion, ~30 lines).
//
//   reference member (optimizer.hpp / optimizer.cpp)          here
//   initialize(parent, name, costmap_ros, params_handler)      initialize(settings, critics, robot)
//   shutdown()                                                 shutdown()
//   evalControl(pose, speed, plan, goal, goal_checker)         evalControl(pose, speed, plan, goal, goal_checker_xy_tolerance, costmap)
//   getGeneratedTrajectories()                                 getGeneratedTrajectories(x, y, yaw)
//   getOptimizedTrajectory()                                   getOptimizedTrajectory(pose) -> [T][3]
//   setSpeedLimit(speed_limit, percentage)                     setSpeedLimit(speed_limit, percentage)
//   reset()                                                    reset()
//   protected: optimize / prepare / fallback / shiftControlSequence / getControlFromSequenceAsTwist / setOffset
//
// The ABI prefix is a macro so that the very same class can be compiled against the CPU oracle in the tests
// (-DMPPI_ABI_PREFIX=oracle_); the product build always binds mppi_.
#ifndef MPPI_OPTIMIZER_HPP_
#define MPPI_OPTIMIZER_HPP_

#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "mppi_b200.h"

#ifndef MPPI_ABI_PREFIX
#define MPPI_ABI_PREFIX mppi_
#endif
#define MPPI_ABI_CAT2(a, b) a##b
#define MPPI_ABI_CAT(a, b) MPPI_ABI_CAT2(a, b)
#define MPPI_ABI(name) MPPI_ABI_CAT(MPPI_ABI_PREFIX, name)

#ifdef MPPI_ABI_DECLARE_PREFIXED
// the oracle exports the same signatures under its own prefix; declare the ones this header calls
extern "C" {
void MPPI_ABI(config_default)(mppi_config *);
int MPPI_ABI(create)(const mppi_config *, mppi_handle **);
void MPPI_ABI(destroy)(mppi_handle *);
int MPPI_ABI(reset)(mppi_handle *);
int MPPI_ABI(set_critics)(mppi_handle *, const mppi_critic_desc *, int32_t);
int MPPI_ABI(set_robot)(mppi_handle *, const mppi_robot_desc *);
int MPPI_ABI(set_speed_limit)(mppi_handle *, double, int32_t);
int MPPI_ABI(get_constraints)(const mppi_handle *, float[4]);
int MPPI_ABI(set_noise)(mppi_handle *, const float *, const float *, const float *);
int MPPI_ABI(set_control_sequence)(mppi_handle *, const float *, const float *, const float *);
int MPPI_ABI(get_control_sequence)(mppi_handle *, float *, float *, float *);
int MPPI_ABI(eval_control)(mppi_handle *, const mppi_cycle_in *, int32_t, mppi_cycle_out *, float[3]);
int MPPI_ABI(get_trajectories)(mppi_handle *, float *, float *, float *);
int MPPI_ABI(get_optimized_trajectory)(mppi_handle *, double, double, double, float *);
}
#endif

namespace mppi_b200
{

struct Pose {double x{0}, y{0}, yaw{0};};          // geometry_msgs::msg::Pose: position + tf2::getYaw(orientation)
struct Twist {double vx{0}, vy{0}, wz{0};};        // geometry_msgs::msg::Twist: linear.x, linear.y, angular.z
struct Path {std::vector<float> x, y, yaw;};       // utils::toTensor(nav_msgs::msg::Path) (utils.hpp:180-192)

// ROS parameters of the optimizer (optimizer.cpp:62-93); names are the parameter names
struct OptimizerSettings
{
  mppi_config base;                 // model_dt, time_steps, batch_size, iteration_count, temperature, gamma, limits, std, model
  unsigned retry_attempt_limit{1};  // "retry_attempt_limit"
  double controller_frequency{20.0};// parent parameter "controller_frequency" (static)
  OptimizerSettings() {MPPI_ABI(config_default)(&base);}
};

class Optimizer
{
public:
  Optimizer() = default;
  Optimizer(const Optimizer &) = delete;
  Optimizer & operator=(const Optimizer &) = delete;
  ~Optimizer() {shutdown();}

  // ref: Optimizer::initialize optimizer.cpp:35-55 (getParams, critic manager, noise generator, reset)
  void initialize(const OptimizerSettings & settings, const std::vector<mppi_critic_desc> & critics, const mppi_robot_desc & robot)
  {
    shutdown();
    settings_ = settings;
    setOffset(settings.controller_frequency);                       // may throw, like the reference
    check(MPPI_ABI(create)(&settings_.base, &handle_), "create");   // throws on an invalid motion model (optimizer.cpp:421-424)
    check(MPPI_ABI(set_robot)(handle_, &robot), "set_robot");
    check(MPPI_ABI(set_critics)(handle_, critics.data(), static_cast<int32_t>(critics.size())), "set_critics");
#ifndef MPPI_ABI_DECLARE_PREFIXED
    // the controller never reads device_ms (the reference has no such output): no event records / read-back per cycle
    check(mppi_set_timing(handle_, 0), "set_timing");
#endif
    vx_.assign(settings_.base.time_steps, 0.0f); vy_ = vx_; wz_ = vx_;
  }

  // ref: Optimizer::shutdown optimizer.cpp:57-60
  void shutdown()
  {
    if (handle_) {MPPI_ABI(destroy)(handle_); handle_ = nullptr;}
  }

  // ref: Optimizer::evalControl optimizer.cpp:134-155.  goal_checker_xy_tolerance < 0 means goal_checker == nullptr.
  Twist evalControl(
    const Pose & robot_pose, const Twist & robot_speed, const Path & plan, const Pose & goal,
    double goal_checker_xy_tolerance, const mppi_costmap & costmap)
  {
    mppi_cycle_in in = prepare(robot_pose, robot_speed, plan, goal, goal_checker_xy_tolerance, costmap);
    float cmd[3];
    optimize(in, cmd);
    // The reference's `do {optimize();} while (fallback(critics_data_.fail_flag));` (optimizer.cpp:143-145) cannot recover:
    // fail_flag is cleared in prepare() only (:198), which the loop does not repeat, so every retry's
    // evalTrajectoriesScores breaks before its first critic (critic_manager.cpp:70-73), the flag stays set, and after
    // retry_attempt_limit soft resets fallback() throws "Optimizer fail to compute path".  The retries' optimize()
    // therefore only computes an update from all-zero costs that the next reset() wipes: the mirror keeps the resets and
    // the throw (same observable behaviour, same state afterwards) and does not launch those cycles.
    const bool fail = last_.fail_flag != 0;
    while (fallback(fail)) {}
    return Twist{cmd[0], cmd[1], cmd[2]};
  }

  // ref: Optimizer::getGeneratedTrajectories optimizer.cpp:455-458 ([B][T] row-major each)
  void getGeneratedTrajectories(std::vector<float> & x, std::vector<float> & y, std::vector<float> & yaw)
  {
    const size_t n = static_cast<size_t>(settings_.base.batch_size) * settings_.base.time_steps;
    x.resize(n); y.resize(n); yaw.resize(n);
    check(MPPI_ABI(get_trajectories)(handle_, x.data(), y.data(), yaw.data()), "get_trajectories");
  }

  // ref: Optimizer::getOptimizedTrajectory optimizer.cpp:345-360 -> [T][3] = (x, y, yaw)
  std::vector<float> getOptimizedTrajectory(const Pose & robot_pose)
  {
    std::vector<float> t3(static_cast<size_t>(settings_.base.time_steps) * 3);
    check(MPPI_ABI(get_optimized_trajectory)(handle_, robot_pose.x, robot_pose.y, robot_pose.yaw, t3.data()), "get_optimized_trajectory");
    return t3;
  }

  // ref: Optimizer::setSpeedLimit optimizer.cpp:428-453 (speed_limit == nav2_costmap_2d::NO_SPEED_LIMIT == 0.0 restores)
  void setSpeedLimit(double speed_limit, bool percentage)
  {
    check(MPPI_ABI(set_speed_limit)(handle_, speed_limit, percentage ? 1 : 0), "set_speed_limit");
  }

  // ref: Optimizer::reset optimizer.cpp:116-132
  void reset() {check(MPPI_ABI(reset)(handle_), "reset"); ++resets_;}
  size_t resetCount() const {return resets_;}

  // ---- access for tests and for the shim -----------------------------------------------------------------------
  mppi_handle * handle() {return handle_;}
  const OptimizerSettings & settings() const {return settings_;}
  bool shiftControlSequenceEnabled() const {return shift_control_sequence_;}
  const mppi_cycle_out & lastCycle() const {return last_;}
  // the control sequence as evalControl left it (control_sequence_, optimizer.hpp:250)
  const std::vector<float> & controlVx() const {return vx_;}
  const std::vector<float> & controlVy() const {return vy_;}
  const std::vector<float> & controlWz() const {return wz_;}
  // parity mode: the reference's injected noise tensors, [B][T] row-major (NoiseGenerator, noise_generator.cpp:65-74)
  void setNoise(const float * vx, const float * vy, const float * wz) {check(MPPI_ABI(set_noise)(handle_, vx, vy, wz), "set_noise");}

protected:
  // ref: Optimizer::prepare optimizer.cpp:185-204 (the device side zeroes costs / flags inside optimize)
  mppi_cycle_in prepare(
    const Pose & robot_pose, const Twist & robot_speed, const Path & plan, const Pose & goal,
    double goal_checker_xy_tolerance, const mppi_costmap & costmap) const
  {
    if (plan.x.size() != plan.y.size() || plan.x.size() != plan.yaw.size()) {throw std::runtime_error("plan arrays differ in length");}
    mppi_cycle_in in{};
    in.pose_x = robot_pose.x; in.pose_y = robot_pose.y; in.pose_yaw = robot_pose.yaw;
    in.speed_vx = robot_speed.vx; in.speed_vy = robot_speed.vy; in.speed_wz = robot_speed.wz;
    in.goal_x = goal.x; in.goal_y = goal.y;
    in.goal_checker_xy_tolerance = goal_checker_xy_tolerance;
    in.path_size = static_cast<int32_t>(plan.x.size());
    in.path_x = plan.x.data(); in.path_y = plan.y.data(); in.path_yaw = plan.yaw.data();
    in.costmap = costmap;
    return in;
  }

  // ref: Optimizer::optimize optimizer.cpp:157-164 + the tail of evalControl (:147-152), one device round trip
  void optimize(const mppi_cycle_in & in, float cmd[3])
  {
    last_ = mppi_cycle_out{};
    last_.control_vx = vx_.data(); last_.control_vy = vy_.data(); last_.control_wz = wz_.data();
    check(MPPI_ABI(eval_control)(handle_, &in, shift_control_sequence_ ? 1 : 0, &last_, cmd), "eval_control");
  }

  // ref: Optimizer::fallback optimizer.cpp:166-183.  The reference's retry counter is a function-local static, i.e.
  // shared by every Optimizer of the process; reproduced as a class-level static.
  bool fallback(bool fail)
  {
    static size_t counter = 0;
    if (!fail) {
      counter = 0;
      return false;
    }
    reset();
    if (++counter > settings_.retry_attempt_limit) {
      counter = 0;
      throw std::runtime_error("Optimizer fail to compute path");
    }
    return true;
  }

  // ref: Optimizer::setOffset optimizer.cpp:95-114
  void setOffset(double controller_frequency)
  {
    const double controller_period = 1.0 / controller_frequency;
    constexpr double eps = 1e-6;
    shift_control_sequence_ = false;
    if ((controller_period + eps) < settings_.base.model_dt) {
      // "Controller period is less then model dt, consider setting it equal" (warning only)
    } else if (std::fabs(controller_period - settings_.base.model_dt) < eps) {
      shift_control_sequence_ = true;   // "Control sequence shifting is ON"
    } else {
      throw std::runtime_error("Controller period more then model dt, set it equal to model dt");
    }
  }

  void check(int status, const char * what) const
  {
    if (status != MPPI_OK) {
      throw std::runtime_error(std::string("mppi ") + what + " failed (status " + std::to_string(status) + ")");
    }
  }

  OptimizerSettings settings_;
  mppi_handle * handle_{nullptr};
  bool shift_control_sequence_{false};
  size_t resets_{0};
  mppi_cycle_out last_{};
  std::vector<float> vx_, vy_, wz_;
};

}  // namespace mppi_b200

#endif  // MPPI_OPTIMIZER_HPP_
