// optimizer.cpp (B200 shim) -- replaces src/optimizer.cpp (and with it noise_generator.cpp, motion_models.hpp and the
// xtensor state).  ref: optimizer.cpp:35-155,345-360,412-458.
#include <cstring>
#include <stdexcept>

#include "nav2_sortham_controller/optimizer.hpp"
#include "tf2/utils.h"

namespace sortham
{

void Optimizer::initialize(
  rclcpp_lifecycle::LifecycleNode::WeakPtr parent, const std::string & name,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler)
{
  parent_ = parent;
  name_ = name;
  costmap_ros_ = costmap_ros;
  costmap_ = costmap_ros_->getCostmap();
  parameters_handler_ = param_handler;
  logger_ = parent_.lock()->get_logger();

  getParams();                                                                          // optimizer.cpp:49
  critic_manager_.on_configure(parent_, name_, costmap_ros_, parameters_handler_);      // :51
  {                                                                                     // :52 NoiseGenerator::initialize
    auto getParam = parameters_handler_->getParamGetter(name_);
    getParam(regenerate_noises_, "regenerate_noises", false);                           // noise_generator.cpp:35
  }
  configureDevice(true);                                                                // create + reset (:54)
  // Costmap2D::getCharMap() stays where it is for the life of the costmap: register it once so that large costmaps are
  // uploaded in place, without a staging copy (mppi_register_costmap_memory); a failure only costs the staging memcpy
  const size_t cells = static_cast<size_t>(costmap_->getSizeInCellsX()) * costmap_->getSizeInCellsY();
  mppi_register_costmap_memory(core_.handle(), costmap_->getCharMap(), cells);
}

void Optimizer::shutdown() {core_.shutdown();}

void Optimizer::getParams()
{
  auto & s = settings_.base;
  auto getParam = parameters_handler_->getParamGetter(name_);
  auto getParentParam = parameters_handler_->getParamGetter("");
  getParam(s.model_dt, "model_dt", 0.05f);
  getParam(s.time_steps, "time_steps", 56);
  getParam(s.batch_size, "batch_size", 1000);
  getParam(s.iteration_count, "iteration_count", 1);
  getParam(s.temperature, "temperature", 0.3f);
  getParam(s.gamma, "gamma", 0.015f);
  getParam(s.vx_max, "vx_max", 0.5);
  getParam(s.vx_min, "vx_min", -0.35);
  getParam(s.vy_max, "vy_max", 0.5);
  getParam(s.wz_max, "wz_max", 1.9);
  getParam(s.vx_std, "vx_std", 0.2);
  getParam(s.vy_std, "vy_std", 0.2);
  getParam(s.wz_std, "wz_std", 0.4);
  getParam(settings_.retry_attempt_limit, "retry_attempt_limit", 1);
  getParam(motion_model_name_, "motion_model", std::string("DiffDrive"));
  setMotionModel(motion_model_name_);
  // The controller's own "visualize" switch (controller.cpp:41): the reference always holds generated_trajectories_ on the
  // host; on the device the [B][T] planes are only written when somebody will read them (getGeneratedTrajectories).
  getParam(visualize_, "visualize", false);
  // any dynamic parameter change ends here (parameters_handler.cpp:66-68 -> optimizer.cpp:88): the reference resets; the
  // shim additionally re-packs what the device holds by value (critic table, create-time settings)
  parameters_handler_->addPostCallback([this]() {setMotionModel(motion_model_name_); configureDevice(false);});
  double controller_frequency;
  getParentParam(controller_frequency, "controller_frequency", 0.0, ParameterType::Static);
  settings_.controller_frequency = controller_frequency;     // setOffset (optimizer.cpp:95-114) runs inside core_.initialize
}

void Optimizer::setMotionModel(const std::string & model)   // optimizer.cpp:412-426
{
  if (model == "DiffDrive") {
    settings_.base.motion_model = MPPI_MODEL_DIFF_DRIVE;
  } else if (model == "Omni") {
    settings_.base.motion_model = MPPI_MODEL_OMNI;
  } else if (model == "Ackermann") {
    settings_.base.motion_model = MPPI_MODEL_ACKERMANN;
    auto getParam = parameters_handler_->getParamGetter(name_ + ".AckermannConstraints");   // motion_models.hpp:93-94
    getParam(ackermann_min_turning_r_, "min_turning_r", 0.2);
  } else {
    throw std::runtime_error(std::string("Model " + model + " is not valid! Valid options are DiffDrive, Omni, or Ackermann"));
  }
}

void Optimizer::configureDevice(bool force_create)
{
  settings_.base.regenerate_noises = regenerate_noises_ ? 1 : 0;
  settings_.base.ackermann_min_turning_r = ackermann_min_turning_r_;
  const bool changed = std::memcmp(&active_, &settings_.base, sizeof(mppi_config)) != 0;
  const auto critics = critic_manager_.describe();
  const auto robot = critic_manager_.describeRobot();
  if (force_create || changed || !core_.handle()) {
    core_.initialize(settings_, critics, robot);   // mppi_create + set_robot + set_critics; throws like the reference
    active_ = settings_.base;
    ++reconfigures_;
    visualize_active_ = false;
  } else {
    if (mppi_set_robot(core_.handle(), &robot) != MPPI_OK ||
      mppi_set_critics(core_.handle(), critics.data(), static_cast<int32_t>(critics.size())) != MPPI_OK)
    {
      throw std::runtime_error(std::string("critic table rejected: ") + mppi_last_error(core_.handle()));
    }
  }
  if (visualize_ != visualize_active_) {
    if (mppi_set_outputs(core_.handle(), visualize_ ? MPPI_WANT_TRAJECTORIES : 0u) != MPPI_OK) {
      throw std::runtime_error(std::string("mppi_set_outputs: ") + mppi_last_error(core_.handle()));
    }
    visualize_active_ = visualize_;
  }
  reset();
}

void Optimizer::reset()
{
  core_.reset();                                    // optimizer.cpp:116-132
  RCLCPP_INFO(logger_, "Optimizer reset");
}

geometry_msgs::msg::TwistStamped Optimizer::evalControl(
  const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
  const nav_msgs::msg::Path & plan, const geometry_msgs::msg::Pose & goal, nav2_core::GoalChecker * goal_checker)
{
  // prepare() (optimizer.cpp:185-204): pose, speed, utils::toTensor(plan) (utils.hpp:180-192), goal
  mppi_b200::Pose pose{robot_pose.pose.position.x, robot_pose.pose.position.y, tf2::getYaw(robot_pose.pose.orientation)};
  mppi_b200::Twist speed{robot_speed.linear.x, robot_speed.linear.y, robot_speed.angular.z};
  mppi_b200::Path path;
  path.x.resize(plan.poses.size()); path.y.resize(plan.poses.size()); path.yaw.resize(plan.poses.size());
  for (size_t i = 0; i < plan.poses.size(); ++i) {
    path.x[i] = static_cast<float>(plan.poses[i].pose.position.x);
    path.y[i] = static_cast<float>(plan.poses[i].pose.position.y);
    path.yaw[i] = static_cast<float>(tf2::getYaw(plan.poses[i].pose.orientation));
  }
  // the goal checker's xy tolerance is all the critics read of it (utils.hpp:201-224, twirling_critic.cpp:33-37)
  double tolerance = -1.0;
  if (goal_checker) {
    geometry_msgs::msg::Pose pose_tolerance;
    geometry_msgs::msg::Twist vel_tolerance;
    if (goal_checker->getTolerances(pose_tolerance, vel_tolerance)) {tolerance = pose_tolerance.position.x;}
    // (utils.hpp:212-215: when getTolerances fails the reference logs and treats the goal as not reached)
  }
  // the caller holds the costmap mutex (controller.cpp:99-100); the library has copied the cells when the call returns
  mppi_costmap cm{};
  cm.cells = costmap_->getCharMap();
  cm.size_x = costmap_->getSizeInCellsX(); cm.size_y = costmap_->getSizeInCellsY();
  cm.resolution = costmap_->getResolution();
  cm.origin_x = costmap_->getOriginX(); cm.origin_y = costmap_->getOriginY();
  last_pose_ = robot_pose.pose;
  const mppi_b200::Twist cmd = core_.evalControl(pose, speed, path, mppi_b200::Pose{goal.position.x, goal.position.y, 0.0}, tolerance, cm);
  geometry_msgs::msg::TwistStamped twist;     // utils::toTwistStamped (optimizer.cpp:396-410)
  twist.header.stamp = plan.header.stamp;
  twist.header.frame_id = "base_link";        // costmap_ros_->getBaseFrameID()
  twist.twist.linear.x = cmd.vx;
  twist.twist.linear.y = isHolonomic() ? cmd.vy : 0.0;
  twist.twist.angular.z = cmd.wz;
  return twist;
}

GeneratedTrajectories & Optimizer::getGeneratedTrajectories()   // optimizer.cpp:455-458
{
  if (!visualize_active_) {throw std::runtime_error("getGeneratedTrajectories: the visualize parameter is false, trajectories are not materialised");}
  generated_.batch_size = settings_.base.batch_size;
  generated_.time_steps = settings_.base.time_steps;
  core_.getGeneratedTrajectories(generated_.x, generated_.y, generated_.yaws);
  return generated_;
}

std::vector<float> Optimizer::getOptimizedTrajectory()
{
  return core_.getOptimizedTrajectory(
    mppi_b200::Pose{last_pose_.position.x, last_pose_.position.y, tf2::getYaw(last_pose_.orientation)});
}

void Optimizer::setSpeedLimit(double speed_limit, bool percentage) {core_.setSpeedLimit(speed_limit, percentage);}

}  // namespace sortham
