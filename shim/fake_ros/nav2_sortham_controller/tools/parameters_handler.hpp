// Stand-in for nav2_sortham_controller/tools/parameters_handler.hpp (TEST INFRASTRUCTURE, see ../../fake_ros.hpp).
//
// A real build keeps the reference's own ParametersHandler (include/nav2_sortham_controller/tools/parameters_handler.hpp:40-263,
// src/parameters_handler.cpp) untouched; the shim only uses its public surface:
//   getParamGetter(ns)(setting, name, default[, ParameterType])   declare-if-missing, read, register a dynamic callback
//   addDynamicParamCallback(name, cb) / addPreCallback / addPostCallback / getLock / start / dynamicParamsCallback
// This file provides that surface over the fake node so that the shim can be compiled and driven without ROS.
#pragma once
#include <functional>
#include <mutex>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <vector>

#include "rclcpp/rclcpp.hpp"
#include "rclcpp_lifecycle/lifecycle_node.hpp"
#include "nav2_util/node_utils.hpp"

namespace sortham
{
enum class ParameterType {Dynamic, Static};

class ParametersHandler
{
public:
  ParametersHandler() = default;
  explicit ParametersHandler(const rclcpp_lifecycle::LifecycleNode::WeakPtr & parent) : node_(parent) {}

  // parameters_handler.cpp:36-44: hook the node's on-set-parameters callback
  void start()
  {
    auto node = node_.lock();
    handle_ = node->add_on_set_parameters_callback(
      [this](const std::vector<rclcpp::Parameter> & ps) {return dynamicParamsCallback(ps);});
  }

  // parameters_handler.cpp:46-72: lock, pre-callbacks, per-parameter callbacks, post-callbacks
  rcl_interfaces::msg::SetParametersResult dynamicParamsCallback(std::vector<rclcpp::Parameter> parameters)
  {
    std::lock_guard<std::mutex> guard(mutex_);
    for (auto & cb : pre_) {cb();}
    for (const auto & p : parameters) {
      auto it = callbacks_.find(p.get_name());
      if (it != callbacks_.end()) {it->second(p);}
    }
    for (auto & cb : post_) {cb();}
    return rcl_interfaces::msg::SetParametersResult{};
  }

  auto getParamGetter(const std::string & ns)
  {
    return [this, ns](auto & setting, const std::string & name, auto default_value, ParameterType type = ParameterType::Dynamic) {
             get(setting, ns.empty() ? name : ns + "." + name, std::move(default_value), type);
           };
  }
  template<typename F> void addPostCallback(F && f) {post_.emplace_back(std::forward<F>(f));}
  template<typename F> void addPreCallback(F && f) {pre_.emplace_back(std::forward<F>(f));}
  template<typename F> void addDynamicParamCallback(const std::string & name, F && f) {callbacks_[name] = std::forward<F>(f);}
  std::mutex * getLock() {return &mutex_;}

private:
  template<typename SettingT, typename ParamT>
  void get(SettingT & setting, const std::string & name, ParamT def, ParameterType type)
  {
    auto node = node_.lock();
    nav2_util::declare_parameter_if_not_declared(node, name, rclcpp::ParameterValue(def));
    ParamT in{};
    node->get_parameter(name, in);
    setting = static_cast<SettingT>(in);
    if (type == ParameterType::Dynamic && callbacks_.find(name) == callbacks_.end()) {
      callbacks_[name] = [&setting](const rclcpp::Parameter & p) {
          if constexpr (std::is_same_v<SettingT, bool>) {setting = p.as_bool();}
          else if constexpr (std::is_integral_v<SettingT>) {setting = static_cast<SettingT>(p.as_int());}
          else if constexpr (std::is_floating_point_v<SettingT>) {setting = static_cast<SettingT>(p.as_double());}
          else if constexpr (std::is_same_v<SettingT, std::string>) {setting = p.as_string();}
          else if constexpr (std::is_same_v<SettingT, std::vector<double>>) {setting = p.as_double_array();}
          else if constexpr (std::is_same_v<SettingT, std::vector<std::string>>) {setting = p.as_string_array();}
        };
    }
  }

  std::mutex mutex_;
  rclcpp_lifecycle::LifecycleNode::WeakPtr node_;
  rclcpp::node_interfaces::OnSetParametersCallbackHandle::SharedPtr handle_;
  std::unordered_map<std::string, std::function<void(const rclcpp::Parameter &)>> callbacks_;
  std::vector<std::function<void()>> pre_, post_;
};
}  // namespace sortham
