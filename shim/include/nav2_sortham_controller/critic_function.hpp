// critic_function.hpp (B200 shim) -- replaces include/nav2_sortham_controller/critic_function.hpp:44-114.
//
// Same base class name, same on_configure() signature and behaviour (reads "<name>.enabled", then calls the pure-virtual
// initialize()), same getName(), still loaded by pluginlib under "sortham::critics::<Name>" (critics.xml).  What changes:
// a built-in critic no longer scores xtensor planes on the host -- score(CriticData &) is gone -- it DESCRIBES itself as a
// POD mppi_critic_desc (include/mppi_b200.h) that the manager packs, in list order, into the device critic table
// (mppi_set_critics).  Dynamic parameter changes land in the members as before (ParametersHandler callbacks); the
// optimizer's post-callback re-describes every critic and resets (optimizer.cpp:88).
#ifndef NAV2_SORTHAM_CONTROLLER__CRITIC_FUNCTION_HPP_
#define NAV2_SORTHAM_CONTROLLER__CRITIC_FUNCTION_HPP_

#include <memory>
#include <string>

#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include "nav2_costmap_2d/costmap_2d_ros.hpp"
#include "nav2_sortham_controller/tools/parameters_handler.hpp"
#include "mppi_b200.h"

namespace sortham::critics
{

class CriticFunction
{
public:
  CriticFunction() = default;
  virtual ~CriticFunction() = default;

  void on_configure(
    rclcpp_lifecycle::LifecycleNode::WeakPtr parent, const std::string & parent_name, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler)
  {
    parent_ = parent;
    logger_ = parent_.lock()->get_logger();
    name_ = name;
    parent_name_ = parent_name;
    costmap_ros_ = costmap_ros;
    costmap_ = costmap_ros_->getCostmap();
    parameters_handler_ = param_handler;
    auto getParam = parameters_handler_->getParamGetter(name_);
    getParam(enabled_, "enabled", true);
    initialize();
  }

  /// reads the critic's parameters into its members (same names and defaults as the reference's initialize())
  virtual void initialize() = 0;

  /// the critic as the device sees it; called at configure time and again after every dynamic parameter change
  virtual void describe(mppi_critic_desc & d) const = 0;

  /// the name filter a critic puts on the inflation layer it looks for ("" = any; only CostCritic has one, cost_critic.cpp:31,79-83)
  virtual std::string inflationLayerName() const {return "";}
  /// does this critic look for an inflation layer at all (Cost / Obstacles: findCircumscribedCost)
  virtual bool usesInflationLayer() const {return false;}

  std::string getName() {return name_;}

protected:
  void describeCommon(mppi_critic_desc & d, int kind, unsigned power, float weight) const
  {
    mppi_critic_default(kind, &d);
    d.kind = kind;
    d.enabled = enabled_ ? 1 : 0;
    d.cost_power = power;
    d.cost_weight = weight;
  }

  bool enabled_{true};
  std::string name_, parent_name_;
  rclcpp_lifecycle::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  nav2_costmap_2d::Costmap2D * costmap_{nullptr};
  ParametersHandler * parameters_handler_{nullptr};
  rclcpp::Logger logger_{rclcpp::get_logger("SORTHAMController")};
};

}  // namespace sortham::critics

#endif  // NAV2_SORTHAM_CONTROLLER__CRITIC_FUNCTION_HPP_
