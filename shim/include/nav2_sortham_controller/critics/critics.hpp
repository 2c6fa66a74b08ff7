// critics.hpp (B200 shim) -- the twelve critic plugin classes of critics.xml:1-53, each reduced to "read my parameters,
// describe myself"; replaces include/nav2_sortham_controller/critics/*.hpp.  Parameter names, types and defaults are the
// reference's (src/critics/*.cpp initialize()); the arithmetic lives in the CUDA kernels (mpcholonavigation_b200/csrc).
#ifndef NAV2_SORTHAM_CONTROLLER__CRITICS__CRITICS_HPP_
#define NAV2_SORTHAM_CONTROLLER__CRITICS__CRITICS_HPP_

#include <array>
#include <string>
#include <vector>

#include "nav2_sortham_controller/critic_function.hpp"

namespace sortham::critics
{

#define SORTHAM_B200_CRITIC_COMMON \
public: \
  void initialize() override; \
  void describe(mppi_critic_desc & d) const override; \
protected: \
  unsigned int power_{0}; \
  float weight_{0};

class ConstraintCritic : public CriticFunction          // constraint_critic.cpp:20-39
{
  SORTHAM_B200_CRITIC_COMMON
  float vx_max_{0}, vy_max_{0}, vx_min_{0};              // read from the PARENT namespace; the device takes them from mppi_config
public:
  float getMaxVelConstraint() const;                     // constraint_critic.hpp:52-53, used by the reference's tests
  float getMinVelConstraint() const;
};

class CostCritic : public CriticFunction                // cost_critic.cpp:22-61
{
  SORTHAM_B200_CRITIC_COMMON
  bool consider_footprint_{false};
  float critical_cost_{0}, collision_cost_{0}, near_goal_distance_{0};
  std::string inflation_layer_name_;
public:
  std::string inflationLayerName() const override {return inflation_layer_name_;}
  bool usesInflationLayer() const override {return true;}
};

class GoalCritic : public CriticFunction                // goal_critic.cpp:23-35
{
  SORTHAM_B200_CRITIC_COMMON
  float threshold_to_consider_{0};
};

class GoalAngleCritic : public CriticFunction           // goal_angle_critic.cpp:20-35
{
  SORTHAM_B200_CRITIC_COMMON
  float threshold_to_consider_{0};
};

class ObstaclesCritic : public CriticFunction           // obstacles_critic.cpp:21-51,78-80
{
  SORTHAM_B200_CRITIC_COMMON
  bool consider_footprint_{false};
  float repulsion_weight_{0}, critical_weight_{0}, collision_cost_{0}, collision_margin_distance_{0}, near_goal_distance_{0};
  float inflation_scale_factor_{0}, inflation_radius_{0};   // only read when an inflation layer exists (obstacles_critic.cpp:78-80)
public:
  bool usesInflationLayer() const override {return true;}
};

class PathAlignCritic : public CriticFunction           // path_align_critic.cpp:26-44
{
  SORTHAM_B200_CRITIC_COMMON
  int offset_from_furthest_{0}, trajectory_point_step_{0};
  float threshold_to_consider_{0}, max_path_occupancy_ratio_{0};
  bool use_path_orientations_{false};
};

class PathAlignLegacyCritic : public CriticFunction     // path_align_legacy_critic.cpp:26-44
{
  SORTHAM_B200_CRITIC_COMMON
  int offset_from_furthest_{0}, trajectory_point_step_{0};
  float threshold_to_consider_{0}, max_path_occupancy_ratio_{0};
  bool use_path_orientations_{false};
};

class PathAngleCritic : public CriticFunction           // path_angle_critic.cpp:23-56
{
  SORTHAM_B200_CRITIC_COMMON
  int offset_from_furthest_{0};
  float threshold_to_consider_{0}, max_angle_to_furthest_{0};
  bool reversing_allowed_{true}, forward_preference_{true};
};

class PathFollowCritic : public CriticFunction          // path_follow_critic.cpp:23-33
{
  SORTHAM_B200_CRITIC_COMMON
  int offset_from_furthest_{0};
  float threshold_to_consider_{0};
};

class PreferForwardCritic : public CriticFunction       // prefer_forward_critic.cpp:20-31
{
  SORTHAM_B200_CRITIC_COMMON
  float threshold_to_consider_{0};
};

class TwirlingCritic : public CriticFunction            // twirling_critic.cpp:20-29
{
  SORTHAM_B200_CRITIC_COMMON
};

class VelocityDeadbandCritic : public CriticFunction    // velocity_deadband_critic.cpp:20-40
{
  SORTHAM_B200_CRITIC_COMMON
  std::array<float, 3> deadband_velocities_{0.0f, 0.0f, 0.0f};
};

#undef SORTHAM_B200_CRITIC_COMMON

}  // namespace sortham::critics

#endif  // NAV2_SORTHAM_CONTROLLER__CRITICS__CRITICS_HPP_
