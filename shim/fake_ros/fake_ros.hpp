// fake_ros.hpp -- header-only stand-ins for the few ROS 2 / Nav2 types the plugin boundary touches.
//
// TEST INFRASTRUCTURE.  This image has no ROS (no rclcpp, nav2_costmap_2d, pluginlib, tf2, geometry_msgs), so the shim under
// shim/ -- the sources a maintainer drops into nav2_sortham_controller in place of optimizer.cpp, critic_manager.cpp and
// src/critics/*.cpp -- is compiled and exercised here against these stand-ins.  Only the members the shim (and the
// reference code around it: controller.cpp, parameters_handler.hpp) actually calls exist.  In a real workspace the include
// paths resolve to the real packages and this directory is not on the include path.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <variant>
#include <vector>

// ------------------------------------------------------------------------------------------------ rclcpp
namespace rclcpp
{
class ParameterValue
{
public:
  using V = std::variant<std::monostate, bool, int64_t, double, std::string, std::vector<std::string>, std::vector<double>,
      std::vector<int64_t>, std::vector<bool>>;
  ParameterValue() = default;
  ParameterValue(bool v) : v_(v) {}
  ParameterValue(int v) : v_(static_cast<int64_t>(v)) {}
  ParameterValue(unsigned v) : v_(static_cast<int64_t>(v)) {}
  ParameterValue(int64_t v) : v_(v) {}
  ParameterValue(float v) : v_(static_cast<double>(v)) {}
  ParameterValue(double v) : v_(v) {}
  ParameterValue(const char * v) : v_(std::string(v)) {}
  ParameterValue(std::string v) : v_(std::move(v)) {}
  ParameterValue(std::vector<std::string> v) : v_(std::move(v)) {}
  ParameterValue(std::vector<double> v) : v_(std::move(v)) {}
  const V & get() const {return v_;}
private:
  V v_;
};

class Parameter
{
public:
  Parameter() = default;
  Parameter(std::string name, ParameterValue value) : name_(std::move(name)), value_(std::move(value)) {}
  const std::string & get_name() const {return name_;}
  bool as_bool() const {return std::get<bool>(value_.get());}
  int64_t as_int() const {return std::get<int64_t>(value_.get());}
  double as_double() const
  {
    if (auto p = std::get_if<int64_t>(&value_.get())) {return static_cast<double>(*p);}   // YAML "5" for a double parameter
    return std::get<double>(value_.get());
  }
  std::string as_string() const {return std::get<std::string>(value_.get());}
  std::vector<std::string> as_string_array() const {return std::get<std::vector<std::string>>(value_.get());}
  std::vector<double> as_double_array() const {return std::get<std::vector<double>>(value_.get());}
  std::vector<int64_t> as_integer_array() const {return std::get<std::vector<int64_t>>(value_.get());}
  std::vector<bool> as_bool_array() const {return std::get<std::vector<bool>>(value_.get());}
  const ParameterValue & get_parameter_value() const {return value_;}
private:
  std::string name_;
  ParameterValue value_;
};

class Logger {public: std::string name;};
inline Logger get_logger(const std::string & n) {return Logger{n};}

class Time
{
public:
  explicit Time(double s = 0.0) : s_(s) {}
  double seconds() const {return s_;}
private:
  double s_;
};
class Duration
{
public:
  explicit Duration(double s = 0.0) : s_(s) {}
  static Duration from_seconds(double s) {return Duration(s);}
  double seconds() const {return s_;}
private:
  double s_;
};
inline Duration operator-(const Time & a, const Time & b) {return Duration(a.seconds() - b.seconds());}
inline bool operator>(const Duration & a, const Duration & b) {return a.seconds() > b.seconds();}
class Clock
{
public:
  using SharedPtr = std::shared_ptr<Clock>;
  Time now() const {return Time(fake_now);}
  double fake_now{0.0};   // the test advances time by hand
};
}  // namespace rclcpp

#define FAKE_ROS_LOG(level, logger, ...) do {if (::fake_ros_verbose()) {std::fprintf(stderr, "[%s] [%s] ", level, (logger).name.c_str()); \
  std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n");}} while (0)
inline bool & fake_ros_verbose() {static bool v = false; return v;}
#define RCLCPP_INFO(logger, ...) FAKE_ROS_LOG("INFO", logger, __VA_ARGS__)
#define RCLCPP_WARN(logger, ...) FAKE_ROS_LOG("WARN", logger, __VA_ARGS__)
#define RCLCPP_ERROR(logger, ...) FAKE_ROS_LOG("ERROR", logger, __VA_ARGS__)
#define RCLCPP_DEBUG(logger, ...) do {} while (0)

namespace rcl_interfaces::msg
{
struct SetParametersResult {bool successful{true}; std::string reason;};
}

namespace rclcpp::node_interfaces
{
struct OnSetParametersCallbackHandle
{
  using SharedPtr = std::shared_ptr<OnSetParametersCallbackHandle>;
  std::function<rcl_interfaces::msg::SetParametersResult(const std::vector<rclcpp::Parameter> &)> callback;
};
}

namespace rclcpp_lifecycle
{
class LifecycleNode : public std::enable_shared_from_this<LifecycleNode>
{
public:
  using SharedPtr = std::shared_ptr<LifecycleNode>;
  using WeakPtr = std::weak_ptr<LifecycleNode>;
  explicit LifecycleNode(std::string name) : name_(std::move(name)), clock_(std::make_shared<rclcpp::Clock>()) {}
  const char * get_name() const {return name_.c_str();}
  rclcpp::Logger get_logger() const {return rclcpp::get_logger(name_);}
  rclcpp::Clock::SharedPtr get_clock() {return clock_;}
  bool has_parameter(const std::string & n) const {return params_.count(n) != 0;}
  // parameter overrides (what the YAML file / launch description provides) win over the declared default
  void set_override(const std::string & n, rclcpp::ParameterValue v) {overrides_[n] = std::move(v);}
  void declare_parameter(const std::string & n, const rclcpp::ParameterValue & def)
  {
    auto o = overrides_.find(n);
    params_[n] = o != overrides_.end() ? o->second : def;
    declared_order_.push_back(n);
  }
  template<typename T>
  bool get_parameter(const std::string & n, T & out) const
  {
    auto it = params_.find(n);
    if (it == params_.end()) {return false;}
    const rclcpp::Parameter p(n, it->second);
    if constexpr (std::is_same_v<T, bool>) {out = p.as_bool();}
    else if constexpr (std::is_integral_v<T>) {out = static_cast<T>(p.as_int());}
    else if constexpr (std::is_floating_point_v<T>) {out = static_cast<T>(p.as_double());}
    else if constexpr (std::is_same_v<T, std::string>) {out = p.as_string();}
    else if constexpr (std::is_same_v<T, std::vector<std::string>>) {out = p.as_string_array();}
    else if constexpr (std::is_same_v<T, std::vector<double>>) {out = p.as_double_array();}
    return true;
  }
  rclcpp::node_interfaces::OnSetParametersCallbackHandle::SharedPtr add_on_set_parameters_callback(
    std::function<rcl_interfaces::msg::SetParametersResult(const std::vector<rclcpp::Parameter> &)> cb)
  {
    auto h = std::make_shared<rclcpp::node_interfaces::OnSetParametersCallbackHandle>();
    h->callback = std::move(cb);
    on_set_.push_back(h);
    return h;
  }
  // `ros2 param set`: stores the value and runs the registered callbacks (what the executor thread would do)
  rcl_interfaces::msg::SetParametersResult set_parameter(const rclcpp::Parameter & p)
  {
    params_[p.get_name()] = p.get_parameter_value();
    rcl_interfaces::msg::SetParametersResult res;
    for (auto & w : on_set_) {if (auto h = w.lock()) {res = h->callback({p});}}
    return res;
  }
  const std::vector<std::string> & declared() const {return declared_order_;}
  const std::map<std::string, rclcpp::ParameterValue> & overrides() const {return overrides_;}
private:
  std::string name_;
  rclcpp::Clock::SharedPtr clock_;
  std::map<std::string, rclcpp::ParameterValue> params_, overrides_;
  std::vector<std::string> declared_order_;
  std::vector<std::weak_ptr<rclcpp::node_interfaces::OnSetParametersCallbackHandle>> on_set_;
};
}  // namespace rclcpp_lifecycle

namespace nav2_util
{
template<typename NodeT>
void declare_parameter_if_not_declared(NodeT node, const std::string & name, const rclcpp::ParameterValue & def)
{
  if (!node->has_parameter(name)) {node->declare_parameter(name, def);}
}
}

// ------------------------------------------------------------------------------------------------ messages, tf2
namespace builtin_interfaces::msg {struct Time {int32_t sec{0}; uint32_t nanosec{0};};}
namespace std_msgs::msg {struct Header {builtin_interfaces::msg::Time stamp; std::string frame_id;};}
namespace geometry_msgs::msg
{
struct Point {double x{0}, y{0}, z{0};};
struct Quaternion {double x{0}, y{0}, z{0}, w{1};};
struct Pose {Point position; Quaternion orientation;};
struct PoseStamped {std_msgs::msg::Header header; Pose pose;};
struct Vector3 {double x{0}, y{0}, z{0};};
struct Twist {Vector3 linear, angular;};
struct TwistStamped {std_msgs::msg::Header header; Twist twist;};
}
namespace nav_msgs::msg {struct Path {std_msgs::msg::Header header; std::vector<geometry_msgs::msg::PoseStamped> poses;};}

namespace tf2
{
// tf2::getYaw(const geometry_msgs::msg::Quaternion &) (tf2/utils.h -> tf2/impl/utils.h getYaw): the same formula with the
// gimbal-lock guards the oracle restates (oracle_get_yaw)
inline double getYaw(const geometry_msgs::msg::Quaternion & q)
{
  const double sqx = q.x * q.x, sqy = q.y * q.y, sqz = q.z * q.z, sqw = q.w * q.w;
  const double sarg = -2.0 * (q.x * q.z - q.w * q.y) / (sqx + sqy + sqz + sqw);
  if (sarg <= -0.99999) {return -2.0 * std::atan2(q.y, q.x);}
  if (sarg >= 0.99999) {return 2.0 * std::atan2(q.y, q.x);}
  return std::atan2(2.0 * (q.x * q.y + q.w * q.z), sqw + sqx - sqy - sqz);
}
}
namespace tf2_ros {class Buffer {};}

// ------------------------------------------------------------------------------------------------ nav2_costmap_2d
namespace nav2_costmap_2d
{
constexpr unsigned char NO_INFORMATION = 255, LETHAL_OBSTACLE = 254, INSCRIBED_INFLATED_OBSTACLE = 253, FREE_SPACE = 0;
constexpr double NO_SPEED_LIMIT = 0.0;

class Costmap2D
{
public:
  using mutex_t = std::recursive_mutex;
  Costmap2D(unsigned sx, unsigned sy, double res, double ox, double oy, unsigned char fill = 0)
  : size_x_(sx), size_y_(sy), res_(res), ox_(ox), oy_(oy), cells_(static_cast<size_t>(sx) * sy, fill) {}
  unsigned char * getCharMap() {return cells_.data();}
  unsigned getSizeInCellsX() const {return size_x_;}
  unsigned getSizeInCellsY() const {return size_y_;}
  double getResolution() const {return res_;}
  double getOriginX() const {return ox_;}
  double getOriginY() const {return oy_;}
  unsigned char getCost(unsigned mx, unsigned my) const {return cells_[static_cast<size_t>(my) * size_x_ + mx];}
  void setCost(unsigned mx, unsigned my, unsigned char c) {cells_[static_cast<size_t>(my) * size_x_ + mx] = c;}
  mutex_t * getMutex() {return &mutex_;}
private:
  unsigned size_x_, size_y_;
  double res_, ox_, oy_;
  std::vector<unsigned char> cells_;
  mutex_t mutex_;
};

class Layer
{
public:
  virtual ~Layer() = default;
  const std::string & getName() const {return name_;}
  std::string name_{"layer"};
};
class InflationLayer : public Layer
{
public:
  InflationLayer(double radius, double scale, double inscribed, double resolution)
  : radius_(radius), scale_(scale), inscribed_(inscribed), resolution_(resolution) {name_ = "inflation_layer";}
  double getCostScalingFactor() const {return scale_;}
  double getInflationRadius() const {return radius_;}
  // InflationLayer::computeCost(double distance_in_cells) (Nav2 Humble inflation_layer.hpp)
  unsigned char computeCost(double distance) const
  {
    unsigned char cost = 0;
    if (distance == 0) {
      cost = LETHAL_OBSTACLE;
    } else if (distance * resolution_ <= inscribed_) {
      cost = INSCRIBED_INFLATED_OBSTACLE;
    } else {
      const double factor = std::exp(-1.0 * scale_ * (distance * resolution_ - inscribed_));
      cost = static_cast<unsigned char>((INSCRIBED_INFLATED_OBSTACLE - 1) * factor);
    }
    return cost;
  }
private:
  double radius_, scale_, inscribed_, resolution_;
};

class LayeredCostmap
{
public:
  std::vector<std::shared_ptr<Layer>> * getPlugins() {return &plugins_;}
  double getInscribedRadius() const {return inscribed_;}
  double getCircumscribedRadius() const {return circumscribed_;}
  bool isTrackingUnknown() const {return track_unknown_;}
  std::vector<std::shared_ptr<Layer>> plugins_;
  double inscribed_{0.0}, circumscribed_{0.0};
  bool track_unknown_{false};
};

class Costmap2DROS
{
public:
  Costmap2DROS(std::shared_ptr<Costmap2D> cm, std::vector<geometry_msgs::msg::Point> footprint)
  : costmap_(std::move(cm)), footprint_(std::move(footprint))
  {
    double mn = 1e300, mx = 0.0;   // nav2_costmap_2d::calculateMinAndMaxDistances, vertices only (enough for the regular polygons used here)
    for (const auto & p : footprint_) {const double d = std::hypot(p.x, p.y); mn = std::min(mn, d); mx = std::max(mx, d);}
    layered_.inscribed_ = footprint_.empty() ? 0.0 : mn;
    layered_.circumscribed_ = mx;
  }
  Costmap2D * getCostmap() {return costmap_.get();}
  LayeredCostmap * getLayeredCostmap() {return &layered_;}
  std::vector<geometry_msgs::msg::Point> getRobotFootprint() const {return footprint_;}
  std::string getGlobalFrameID() const {return "odom";}
private:
  std::shared_ptr<Costmap2D> costmap_;
  std::vector<geometry_msgs::msg::Point> footprint_;
  LayeredCostmap layered_;
};
}  // namespace nav2_costmap_2d

// ------------------------------------------------------------------------------------------------ nav2_core
namespace nav2_core
{
class GoalChecker
{
public:
  virtual ~GoalChecker() = default;
  virtual bool getTolerances(geometry_msgs::msg::Pose & pose_tolerance, geometry_msgs::msg::Twist & vel_tolerance) = 0;
};
}

// ------------------------------------------------------------------------------------------------ pluginlib
namespace pluginlib
{
template<class Base>
struct Registry
{
  static std::map<std::string, std::function<Base *()>> & map() {static std::map<std::string, std::function<Base *()>> m; return m;}
};
class LibraryLoadException : public std::runtime_error {public: using std::runtime_error::runtime_error;};
template<class Base>
class ClassLoader
{
public:
  ClassLoader(std::string package, std::string base) : package_(std::move(package)), base_(std::move(base)) {}
  Base * createUnmanagedInstance(const std::string & lookup_name)
  {
    auto it = Registry<Base>::map().find(lookup_name);
    if (it == Registry<Base>::map().end()) {throw LibraryLoadException("no plugin class " + lookup_name + " (" + base_ + ")");}
    return it->second();
  }
  std::vector<std::string> getDeclaredClasses() const
  {
    std::vector<std::string> out;
    for (auto & kv : Registry<Base>::map()) {out.push_back(kv.first);}
    return out;
  }
private:
  std::string package_, base_;
};
}
#define FAKE_PLUGINLIB_CAT2(a, b) a##b
#define FAKE_PLUGINLIB_CAT(a, b) FAKE_PLUGINLIB_CAT2(a, b)
#define PLUGINLIB_EXPORT_CLASS(Class, Base) \
  namespace {struct FAKE_PLUGINLIB_CAT(FakePluginRegistrar, __LINE__) {FAKE_PLUGINLIB_CAT(FakePluginRegistrar, __LINE__)() { \
    pluginlib::Registry<Base>::map()[#Class] = []() -> Base * {return new Class();};}} FAKE_PLUGINLIB_CAT(g_fake_plugin_registrar_, __LINE__);}
