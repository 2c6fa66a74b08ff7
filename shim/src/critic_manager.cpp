// critic_manager.cpp (B200 shim) -- see critic_manager.hpp.  ref: src/critic_manager.cpp:20-65.
#include <stdexcept>

#include "nav2_sortham_controller/critic_manager.hpp"
#include "nav2_costmap_2d/inflation_layer.hpp"

namespace sortham
{

void CriticManager::on_configure(
  rclcpp_lifecycle::LifecycleNode::WeakPtr parent, const std::string & name,
  std::shared_ptr<nav2_costmap_2d::Costmap2DROS> costmap_ros, ParametersHandler * param_handler)
{
  parent_ = parent;
  costmap_ros_ = costmap_ros;
  name_ = name;
  logger_ = parent_.lock()->get_logger();
  parameters_handler_ = param_handler;
  getParams();
  loadCritics();
}

void CriticManager::getParams()
{
  auto getParam = parameters_handler_->getParamGetter(name_);
  getParam(critic_names_, "critics", std::vector<std::string>{}, ParameterType::Static);
}

void CriticManager::loadCritics()
{
  if (!loader_) {
    loader_ = std::make_unique<pluginlib::ClassLoader<critics::CriticFunction>>(
      "nav2_sortham_controller", "sortham::critics::CriticFunction");
  }
  critics_.clear();
  for (const auto & name : critic_names_) {
    const std::string fullname = getFullName(name);
    critics_.emplace_back(loader_->createUnmanagedInstance(fullname));
    critics_.back()->on_configure(parent_, name_, name_ + "." + name, costmap_ros_, parameters_handler_);
    RCLCPP_INFO(logger_, "Critic loaded : %s", fullname.c_str());
  }
}

std::string CriticManager::getFullName(const std::string & name) {return "sortham::critics::" + name;}

std::vector<mppi_critic_desc> CriticManager::describe() const
{
  std::vector<mppi_critic_desc> table(critics_.size());
  for (size_t q = 0; q < critics_.size(); ++q) {critics_[q]->describe(table[q]);}
  return table;
}

mppi_robot_desc CriticManager::describeRobot() const
{
  mppi_robot_desc r{};
  const auto footprint = costmap_ros_->getRobotFootprint();          // obstacles_critic.cpp:219, cost_critic.cpp:185
  if (footprint.size() > MPPI_MAX_FOOTPRINT) {throw std::runtime_error("robot footprint has too many vertices for the device table");}
  r.footprint_size = static_cast<int32_t>(footprint.size());
  for (size_t i = 0; i < footprint.size(); ++i) {r.footprint_x[i] = footprint[i].x; r.footprint_y[i] = footprint[i].y;}
  auto * layered = costmap_ros_->getLayeredCostmap();
  r.inscribed_radius = layered->getInscribedRadius();                // obstacles_critic.cpp:102
  r.circumscribed_radius = layered->getCircumscribedRadius();        // obstacles_critic.cpp:59, cost_critic.cpp:68
  r.track_unknown = layered->isTrackingUnknown() ? 1 : 0;            // obstacles_critic.cpp:188, cost_critic.cpp:178
  // the inflation layer findCircumscribedCost looks for: the LAST InflationLayer among the plugins (the loops at
  // obstacles_critic.cpp:64-81 / cost_critic.cpp:74-90 overwrite `result` for every match); CostCritic may filter by name
  std::string want_name;
  bool cost_filters = false, other_user = false;
  for (const auto & c : critics_) {
    if (!c->usesInflationLayer()) {continue;}
    if (!c->inflationLayerName().empty()) {cost_filters = true; want_name = c->inflationLayerName();} else {other_user = true;}
  }
  const nav2_costmap_2d::InflationLayer * any = nullptr, * named = nullptr;
  for (auto & layer : *layered->getPlugins()) {
    auto infl = std::dynamic_pointer_cast<nav2_costmap_2d::InflationLayer>(layer);
    if (!infl) {continue;}
    any = infl.get();
    if (cost_filters && infl->getName() == want_name) {named = infl.get();}
  }
  const nav2_costmap_2d::InflationLayer * pick = cost_filters ? named : any;
  if (cost_filters && other_user && named != any) {
    // one robot description serves both obstacle-type critics; they would see different layers
    throw std::runtime_error("CostCritic.inflation_layer_name selects a different inflation layer than ObstaclesCritic uses: not representable");
  }
  r.inflation_layer_found = pick ? 1 : 0;
  r.inflation_cost_scaling_factor = pick ? pick->getCostScalingFactor() : 0.0;
  return r;
}

}  // namespace sortham
