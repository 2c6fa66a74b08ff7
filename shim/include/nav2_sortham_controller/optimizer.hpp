// optimizer.hpp (B200 shim) -- sortham::Optimizer with the reference's public interface
// (include/nav2_sortham_controller/optimizer.hpp:51-117), implemented over the C ABI of libmppi_b200.so.
//
// SORTHAMController (src/controller.cpp) holds this class exactly like the reference's: optimizer_.initialize(parent, name,
// costmap_ros, param_handler), .evalControl(pose, speed, plan, goal, goal_checker), .getGeneratedTrajectories(),
// .getOptimizedTrajectory(), .setSpeedLimit(limit, percentage), .reset(), .shutdown() -- controller.cpp compiles unchanged
// except for the two visualisation getters, whose return types are plain row-major arrays instead of xtensor containers
// (TrajectoryVisualizer::add only iterates them, trajectory_visualizer.cpp:57-108).
#ifndef NAV2_SORTHAM_CONTROLLER__OPTIMIZER_HPP_
#define NAV2_SORTHAM_CONTROLLER__OPTIMIZER_HPP_

#include <memory>
#include <string>
#include <vector>

#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include "nav2_costmap_2d/costmap_2d_ros.hpp"
#include "nav2_core/goal_checker.hpp"
#include "geometry_msgs/msg/twist.hpp"
#include "geometry_msgs/msg/pose_stamped.hpp"
#include "geometry_msgs/msg/twist_stamped.hpp"
#include "nav_msgs/msg/path.hpp"

#include "nav2_sortham_controller/critic_manager.hpp"
#include "nav2_sortham_controller/tools/parameters_handler.hpp"
#include "mppi_optimizer.hpp"

namespace sortham
{

// models::Trajectories (models/trajectories.hpp:28-43) as the visualiser consumes it: x, y, yaws, [batch][time] row-major
struct GeneratedTrajectories
{
  std::vector<float> x, y, yaws;
  size_t batch_size{0}, time_steps{0};
};

class Optimizer
{
public:
  Optimizer() = default;
  ~Optimizer() {shutdown();}

  void initialize(
    rclcpp_lifecycle::LifecycleNode::WeakPtr parent, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * dynamic_parameters_handler);
  void shutdown();

  geometry_msgs::msg::TwistStamped evalControl(
    const geometry_msgs::msg::PoseStamped & robot_pose, const geometry_msgs::msg::Twist & robot_speed,
    const nav_msgs::msg::Path & plan, const geometry_msgs::msg::Pose & goal, nav2_core::GoalChecker * goal_checker);

  GeneratedTrajectories & getGeneratedTrajectories();
  std::vector<float> getOptimizedTrajectory();          // [time_steps][3] = x, y, yaw (optimizer.cpp:345-360)

  void setSpeedLimit(double speed_limit, bool percentage);
  void reset();

  // access for tests / the controller's diagnostics
  mppi_b200::Optimizer & core() {return core_;}
  const mppi_b200::OptimizerSettings & settings() const {return settings_;}
  const CriticManager & criticManager() const {return critic_manager_;}
  size_t reconfigureCount() const {return reconfigures_;}

protected:
  void getParams();
  void setMotionModel(const std::string & model);
  /// (re)creates the device handle when a create-time setting changed, re-sends the critic table and the robot description
  void configureDevice(bool force_create);
  bool isHolonomic() const {return settings_.base.motion_model == MPPI_MODEL_OMNI;}

  rclcpp_lifecycle::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2_costmap_2d::Costmap2D * costmap_{nullptr};
  std::string name_;
  ParametersHandler * parameters_handler_{nullptr};
  CriticManager critic_manager_;
  mppi_b200::OptimizerSettings settings_;      // what the parameters say now (dynamic callbacks write here)
  mppi_config active_{};                       // what the device handle was created with
  std::string motion_model_name_;
  bool regenerate_noises_{false};
  bool visualize_{false}, visualize_active_{false};
  float ackermann_min_turning_r_{0.2f};
  mppi_b200::Optimizer core_;
  geometry_msgs::msg::Pose last_pose_;
  GeneratedTrajectories generated_;
  size_t reconfigures_{0};
  rclcpp::Logger logger_{rclcpp::get_logger("SORTHAMController")};
};

}  // namespace sortham

#endif  // NAV2_SORTHAM_CONTROLLER__OPTIMIZER_HPP_
