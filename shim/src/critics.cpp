// critics.cpp (B200 shim) -- initialize() + describe() of the twelve critic plugins; replaces src/critics/*.cpp.
// initialize() reads the same parameters with the same defaults, in the same order, as the reference's initialize() of the
// same class (cited per class); describe() maps the members onto mppi_critic_desc (include/mppi_b200.h).
#include <algorithm>
#include <cmath>

#include "nav2_sortham_controller/critics/critics.hpp"
#include "pluginlib/class_list_macros.hpp"

namespace sortham::critics
{

// ---- ConstraintCritic (constraint_critic.cpp:20-39) ---------------------------------------------------------------------
void ConstraintCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  auto getParentParam = parameters_handler_->getParamGetter(parent_name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 4.0);
  RCLCPP_INFO(logger_, "ConstraintCritic instantiated with %d power and %f weight.", power_, weight_);
  // parent parameters: declared first by the optimizer (vy_max 0.5 there, optimizer.cpp:77), so its value wins over the 0.0
  // given here; the device derives max_vel / min_vel from the same three values in mppi_config
  getParentParam(vx_max_, "vx_max", 0.5);
  getParentParam(vy_max_, "vy_max", 0.0);
  getParentParam(vx_min_, "vx_min", -0.35);
}
float ConstraintCritic::getMaxVelConstraint() const {return sqrtf(vx_max_ * vx_max_ + vy_max_ * vy_max_);}
float ConstraintCritic::getMinVelConstraint() const
{
  const float min_sgn = vx_min_ > 0.0 ? 1.0 : -1.0;
  return min_sgn * sqrtf(vx_min_ * vx_min_ + vy_max_ * vy_max_);
}
void ConstraintCritic::describe(mppi_critic_desc & d) const {describeCommon(d, MPPI_CRITIC_CONSTRAINT, power_, weight_);}

// ---- CostCritic (cost_critic.cpp:22-61) ---------------------------------------------------------------------------------
void CostCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(consider_footprint_, "consider_footprint", false);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 3.81);
  getParam(critical_cost_, "critical_cost", 300.0);
  getParam(collision_cost_, "collision_cost", 1000000.0);
  getParam(near_goal_distance_, "near_goal_distance", 0.5);
  getParam(inflation_layer_name_, "inflation_layer_name", std::string(""));
  // The reference divides the weight by 254 here and again in a dynamic callback (cost_critic.cpp:34-40).  The device
  // does that division itself (mppi_critic_desc::cost_weight is the RAW parameter value), so the member keeps the raw
  // value and the stock dynamic callback registered by getParam is the right one.
  RCLCPP_INFO(
    logger_, "InflationCostCritic instantiated with %d power and %f / %f weights. Critic will collision check based on %s cost.",
    power_, critical_cost_, weight_ / 254.0f, consider_footprint_ ? "footprint" : "circular");
}
void CostCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_COST, power_, weight_);
  d.consider_footprint = consider_footprint_ ? 1 : 0;
  d.critical_cost = critical_cost_;
  d.collision_cost = collision_cost_;
  d.near_goal_distance = near_goal_distance_;
}

// ---- GoalCritic (goal_critic.cpp:23-35) ---------------------------------------------------------------------------------
void GoalCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 5.0);
  getParam(threshold_to_consider_, "threshold_to_consider", 1.4);
  RCLCPP_INFO(logger_, "GoalCritic instantiated with %d power and %f weight.", power_, weight_);
}
void GoalCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_GOAL, power_, weight_);
  d.threshold_to_consider = threshold_to_consider_;
}

// ---- GoalAngleCritic (goal_angle_critic.cpp:20-35) ----------------------------------------------------------------------
void GoalAngleCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 3.0);
  getParam(threshold_to_consider_, "threshold_to_consider", 0.5);
  RCLCPP_INFO(
    logger_, "GoalAngleCritic instantiated with %d power, %f weight, and %f angular threshold.", power_, weight_,
    threshold_to_consider_);
}
void GoalAngleCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_GOAL_ANGLE, power_, weight_);
  d.threshold_to_consider = threshold_to_consider_;
}

// ---- ObstaclesCritic (obstacles_critic.cpp:21-51, 53-97) ----------------------------------------------------------------
void ObstaclesCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(consider_footprint_, "consider_footprint", false);
  getParam(power_, "cost_power", 1);
  getParam(repulsion_weight_, "repulsion_weight", 1.5);
  getParam(critical_weight_, "critical_weight", 20.0);
  getParam(collision_cost_, "collision_cost", 10000.0);
  getParam(collision_margin_distance_, "collision_margin_distance", 0.10);
  getParam(near_goal_distance_, "near_goal_distance", 0.5);
  // findCircumscribedCost (obstacles_critic.cpp:65-80): the critic's OWN copies of the inflation parameters are only
  // declared and read when an InflationLayer is among the costmap plugins; otherwise they stay 0 and repulsion is off
  for (auto & layer : *costmap_ros_->getLayeredCostmap()->getPlugins()) {
    if (!std::dynamic_pointer_cast<nav2_costmap_2d::InflationLayer>(layer)) {continue;}
    getParam(inflation_scale_factor_, "cost_scaling_factor", 10.0);
    getParam(inflation_radius_, "inflation_radius", 0.55);
  }
  RCLCPP_INFO(
    logger_, "ObstaclesCritic instantiated with %d power and %f / %f weights. Critic will collision check based on %s cost.",
    power_, critical_weight_, repulsion_weight_, consider_footprint_ ? "footprint" : "circular");
}
void ObstaclesCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_OBSTACLES, power_, 0.0f);
  d.consider_footprint = consider_footprint_ ? 1 : 0;
  d.repulsion_weight = repulsion_weight_;
  d.critical_weight = critical_weight_;
  d.collision_cost = collision_cost_;
  d.collision_margin_distance = collision_margin_distance_;
  d.near_goal_distance = near_goal_distance_;
  d.cost_scaling_factor = inflation_scale_factor_;
  d.inflation_radius = inflation_radius_;
}

// ---- PathAlignCritic (path_align_critic.cpp:26-44) ----------------------------------------------------------------------
void PathAlignCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 10.0);
  getParam(max_path_occupancy_ratio_, "max_path_occupancy_ratio", 0.07);
  getParam(offset_from_furthest_, "offset_from_furthest", 20);
  getParam(trajectory_point_step_, "trajectory_point_step", 4);
  getParam(threshold_to_consider_, "threshold_to_consider", 0.5);
  getParam(use_path_orientations_, "use_path_orientations", false);
  RCLCPP_INFO(logger_, "ReferenceTrajectoryCritic instantiated with %d power and %f weight", power_, weight_);
}
void PathAlignCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_PATH_ALIGN, power_, weight_);
  d.max_path_occupancy_ratio = max_path_occupancy_ratio_;
  d.offset_from_furthest = offset_from_furthest_;
  d.trajectory_point_step = trajectory_point_step_;
  d.threshold_to_consider = threshold_to_consider_;
  d.use_path_orientations = use_path_orientations_ ? 1 : 0;
}

// ---- PathAlignLegacyCritic (path_align_legacy_critic.cpp:26-44) ---------------------------------------------------------
void PathAlignLegacyCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 10.0);
  getParam(max_path_occupancy_ratio_, "max_path_occupancy_ratio", 0.07);
  getParam(offset_from_furthest_, "offset_from_furthest", 20);
  getParam(trajectory_point_step_, "trajectory_point_step", 4);
  getParam(threshold_to_consider_, "threshold_to_consider", 0.5);
  getParam(use_path_orientations_, "use_path_orientations", false);
  RCLCPP_INFO(logger_, "PathAlignLegacyCritic instantiated with %d power and %f weight", power_, weight_);
}
void PathAlignLegacyCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_PATH_ALIGN_LEGACY, power_, weight_);
  d.max_path_occupancy_ratio = max_path_occupancy_ratio_;
  d.offset_from_furthest = offset_from_furthest_;
  d.trajectory_point_step = trajectory_point_step_;
  d.threshold_to_consider = threshold_to_consider_;
  d.use_path_orientations = use_path_orientations_ ? 1 : 0;
}

// ---- PathAngleCritic (path_angle_critic.cpp:23-56) ----------------------------------------------------------------------
void PathAngleCritic::initialize()
{
  auto getParentParam = parameters_handler_->getParamGetter(parent_name_);
  float vx_min;
  getParentParam(vx_min, "vx_min", -0.35);
  if (fabs(vx_min) < 1e-6) {
    reversing_allowed_ = false;
  } else if (vx_min < 0.0) {
    reversing_allowed_ = true;
  }
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(offset_from_furthest_, "offset_from_furthest", 4);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 2.0);
  getParam(threshold_to_consider_, "threshold_to_consider", 0.5);
  getParam(max_angle_to_furthest_, "max_angle_to_furthest", 1.2);
  getParam(forward_preference_, "forward_preference", true);
  if (!reversing_allowed_) {forward_preference_ = true;}
  RCLCPP_INFO(
    logger_, "PathAngleCritic instantiated with %d power and %f weight. Reversing %s", power_, weight_,
    reversing_allowed_ ? "allowed." : "not allowed.");
}
void PathAngleCritic::describe(mppi_critic_desc & d) const
{
  // the device re-derives "reversing allowed" from mppi_config::vx_min exactly like :27-32 above
  describeCommon(d, MPPI_CRITIC_PATH_ANGLE, power_, weight_);
  d.offset_from_furthest = offset_from_furthest_;
  d.threshold_to_consider = threshold_to_consider_;
  d.max_angle_to_furthest = max_angle_to_furthest_;
  d.forward_preference = forward_preference_ ? 1 : 0;
}

// ---- PathFollowCritic (path_follow_critic.cpp:23-33) --------------------------------------------------------------------
void PathFollowCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(threshold_to_consider_, "threshold_to_consider", 1.4);
  getParam(offset_from_furthest_, "offset_from_furthest", 6);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 5.0);
}
void PathFollowCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_PATH_FOLLOW, power_, weight_);
  d.threshold_to_consider = threshold_to_consider_;
  d.offset_from_furthest = offset_from_furthest_;
}

// ---- PreferForwardCritic (prefer_forward_critic.cpp:20-31) --------------------------------------------------------------
void PreferForwardCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 5.0);
  getParam(threshold_to_consider_, "threshold_to_consider", 0.5);
  RCLCPP_INFO(logger_, "PreferForwardCritic instantiated with %d power and %f weight.", power_, weight_);
}
void PreferForwardCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_PREFER_FORWARD, power_, weight_);
  d.threshold_to_consider = threshold_to_consider_;
}

// ---- TwirlingCritic (twirling_critic.cpp:20-29) -------------------------------------------------------------------------
void TwirlingCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 10.0);
  RCLCPP_INFO(logger_, "TwirlingCritic instantiated with %d power and %f weight.", power_, weight_);
}
void TwirlingCritic::describe(mppi_critic_desc & d) const {describeCommon(d, MPPI_CRITIC_TWIRLING, power_, weight_);}

// ---- VelocityDeadbandCritic (velocity_deadband_critic.cpp:20-40) --------------------------------------------------------
void VelocityDeadbandCritic::initialize()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(power_, "cost_power", 1);
  getParam(weight_, "cost_weight", 35.0);
  std::vector<double> deadband_velocities{0.0, 0.0, 0.0};
  getParam(deadband_velocities, "deadband_velocities", std::vector<double>{0.0, 0.0, 0.0});
  std::transform(
    deadband_velocities.begin(), deadband_velocities.end(), deadband_velocities_.begin(),
    [](double v) {return static_cast<float>(v);});
  RCLCPP_INFO(
    logger_, "VelocityDeadbandCritic instantiated with %u power, %f weight, deadband_velocity [%f,%f,%f]", power_, weight_,
    deadband_velocities_[0], deadband_velocities_[1], deadband_velocities_[2]);
}
void VelocityDeadbandCritic::describe(mppi_critic_desc & d) const
{
  describeCommon(d, MPPI_CRITIC_VELOCITY_DEADBAND, power_, weight_);
  for (int i = 0; i < 3; ++i) {d.deadband_velocities[i] = deadband_velocities_[i];}
}

}  // namespace sortham::critics

PLUGINLIB_EXPORT_CLASS(sortham::critics::ConstraintCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::CostCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::GoalCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::GoalAngleCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::ObstaclesCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::PathAlignCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::PathAlignLegacyCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::PathAngleCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::PathFollowCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::PreferForwardCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::TwirlingCritic, sortham::critics::CriticFunction)
PLUGINLIB_EXPORT_CLASS(sortham::critics::VelocityDeadbandCritic, sortham::critics::CriticFunction)
