// critic_manager.hpp (B200 shim) -- replaces include/nav2_sortham_controller/critic_manager.hpp + src/critic_manager.cpp.
// Same class, same on_configure(), same "critics" parameter (static string list), same pluginlib loading of
// "sortham::critics::<Name>" (critic_manager.cpp:36-65).  evalTrajectoriesScores(CriticData &) -- the host loop over
// critics (critic_manager.cpp:67-76) -- becomes describe(): the critic table of the device, in list order; the loop
// itself (including the fail_flag short-circuit) runs inside the kernels.
#ifndef NAV2_SORTHAM_CONTROLLER__CRITIC_MANAGER_HPP_
#define NAV2_SORTHAM_CONTROLLER__CRITIC_MANAGER_HPP_

#include <memory>
#include <string>
#include <vector>

#include "pluginlib/class_loader.hpp"
#include "nav2_sortham_controller/critic_function.hpp"

namespace sortham
{

class CriticManager
{
public:
  CriticManager() = default;
  virtual ~CriticManager() = default;

  void on_configure(
    rclcpp_lifecycle::LifecycleNode::WeakPtr parent, const std::string & name,
    std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler);

  /// the critics list as the device critic table (list order == scoring order)
  std::vector<mppi_critic_desc> describe() const;

  /// what the obstacle-type critics learn from the layered costmap (findCircumscribedCost of Cost / Obstacles critic)
  mppi_robot_desc describeRobot() const;

  const std::vector<std::string> & criticNames() const {return critic_names_;}

protected:
  void getParams();
  virtual void loadCritics();
  std::string getFullName(const std::string & name);

  rclcpp_lifecycle::LifecycleNode::WeakPtr parent_;
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros_;
  std::string name_;
  ParametersHandler * parameters_handler_{nullptr};
  std::vector<std::string> critic_names_;
  std::unique_ptr<pluginlib::ClassLoader<critics::CriticFunction>> loader_;
  std::vector<std::unique_ptr<critics::CriticFunction>> critics_;
  rclcpp::Logger logger_{rclcpp::get_logger("SORTHAMController")};
};

}  // namespace sortham

#endif  // NAV2_SORTHAM_CONTROLLER__CRITIC_MANAGER_HPP_
