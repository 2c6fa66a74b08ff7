// test_optimizer.cpp -- the reference's Optimizer tests, restated against the C++ host mirror (include/mppi_optimizer.hpp).
//
//   ref: nav2_sortham_controller/test/optimizer_unit_tests.cpp  FallbackTests :326-348, getControlFromSequenceAsTwistTests
//        :539-575, shiftControlSequenceTests :378-419, setOffset (optimizer.cpp:95-114)
//   ref: nav2_sortham_controller/test/optimizer_smoke_test.cpp :48-116 (400 x 15, three model / critic combinations,
//        consider_footprint = true, bow-tie footprint of test/utils/factory.hpp:116-119, obstacle block of cost 250)
//
// Built twice by tests/test_cpp_host.py: against the CPU oracle (-DMPPI_ABI_PREFIX=oracle_, CPU suite) and against
// libmppi_b200.so (GPU suite).  No gtest in this image: a 20-line EXPECT harness instead.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#include "mppi_optimizer.hpp"

extern "C" void MPPI_ABI(critic_default)(int32_t, mppi_critic_desc *);

static int g_failures = 0;
#define EXPECT_TRUE(c) do {if (!(c)) {std::printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #c); ++g_failures;}} while (0)
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_NEAR(a, b, tol) do {const double _a = (a), _b = (b); if (!(std::fabs(_a - _b) <= (tol))) { \
  std::printf("FAIL %s:%d: %s = %.9g vs %s = %.9g (tol %g)\n", __FILE__, __LINE__, #a, _a, #b, _b, static_cast<double>(tol)); ++g_failures;}} while (0)
#define EXPECT_THROW(stmt) do {bool _t = false; try {stmt;} catch (const std::runtime_error &) {_t = true;} \
  if (!_t) {std::printf("FAIL %s:%d: no throw: %s\n", __FILE__, __LINE__, #stmt); ++g_failures;}} while (0)
#define EXPECT_NO_THROW(stmt) do {try {stmt;} catch (const std::exception & e) { \
  std::printf("FAIL %s:%d: threw %s: %s\n", __FILE__, __LINE__, e.what(), #stmt); ++g_failures;}} while (0)

using mppi_b200::Optimizer;
using mppi_b200::OptimizerSettings;

// exposes the protected members, like the reference's OptimizerTester (optimizer_unit_tests.cpp:37-218)
class OptimizerTester : public Optimizer
{
public:
  bool fallbackWrapper(bool fail) {return fallback(fail);}
  void setOffsetWrapper(double f) {setOffset(f);}
};

static mppi_critic_desc critic(int kind, const std::function<void(mppi_critic_desc &)> & edit = nullptr)
{
  mppi_critic_desc d;
  MPPI_ABI(critic_default)(kind, &d);
  if (edit) {edit(d);}
  return d;
}

// test/utils/factory.hpp:101-131: 40 x 40 cells @ 0.1 m, origin (0, 0), bow-tie "square" footprint of half-size 0.15
struct DummyCostmap
{
  std::vector<uint8_t> cells;
  mppi_costmap view{};
  DummyCostmap()
  : cells(40 * 40, 0)
  {
    view.cells = cells.data(); view.size_x = 40; view.size_y = 40; view.resolution = 0.1; view.origin_x = 0.0; view.origin_y = 0.0;
  }
  void addObstacle(unsigned x, unsigned y, unsigned size, uint8_t cost)   // test/utils/utils.hpp:135-144
  {
    for (unsigned i = x; i < x + size; ++i) {
      for (unsigned j = y; j < y + size; ++j) {cells[j * 40 + i] = cost;}
    }
  }
};

static mppi_robot_desc bowtieRobot()
{
  mppi_robot_desc r;
  std::memset(&r, 0, sizeof(r));
  const double a = 0.15;
  const double xs[4] = {a, -a, a, -a}, ys[4] = {a, -a, -a, a};
  r.footprint_size = 4;
  for (int i = 0; i < 4; ++i) {r.footprint_x[i] = xs[i]; r.footprint_y[i] = ys[i];}
  r.inscribed_radius = a; r.circumscribed_radius = a * std::sqrt(2.0);
  r.inflation_layer_found = 0; r.inflation_cost_scaling_factor = 10.0; r.track_unknown = 0;
  return r;
}

static void testSetOffset()
{
  OptimizerTester t;
  OptimizerSettings s;
  s.base.batch_size = 64; s.base.time_steps = 20;
  s.controller_frequency = 1.0 / s.base.model_dt;          // period == model_dt -> shifting ON
  EXPECT_NO_THROW(t.initialize(s, {}, bowtieRobot()));
  EXPECT_TRUE(t.shiftControlSequenceEnabled());
  s.controller_frequency = 30.0;                           // period 0.033 < model_dt 0.05 -> warning only, shifting off
  EXPECT_NO_THROW(t.initialize(s, {}, bowtieRobot()));
  EXPECT_FALSE(t.shiftControlSequenceEnabled());
  s.controller_frequency = 10.0;                           // period 0.1 > model_dt -> throws (optimizer.cpp:110-112)
  EXPECT_THROW(t.initialize(s, {}, bowtieRobot()));
}

static void testInvalidModelThrows()
{
  OptimizerTester t;
  OptimizerSettings s;
  s.base.batch_size = 64; s.base.time_steps = 20; s.base.motion_model = 7;   // optimizer.cpp:421-424
  EXPECT_THROW(t.initialize(s, {}, bowtieRobot()));
}

static void testFallback()
{
  OptimizerTester t;
  OptimizerSettings s;
  s.base.batch_size = 1000; s.base.time_steps = 50; s.controller_frequency = 30.0; s.retry_attempt_limit = 2;
  t.initialize(s, {}, bowtieRobot());
  // because retry is set to 2, it attempts soft resets 2x before throwing for a hard reset
  EXPECT_FALSE(t.fallbackWrapper(false));
  EXPECT_TRUE(t.fallbackWrapper(true));
  EXPECT_TRUE(t.fallbackWrapper(true));
  EXPECT_THROW(t.fallbackWrapper(true));
}

// evalControl's tail: command index follows the shift flag, vy only for holonomic models, shift semantics
static void testTwistAndShift()
{
  for (int model : {MPPI_MODEL_DIFF_DRIVE, MPPI_MODEL_OMNI}) {
    for (bool shift : {false, true}) {
      OptimizerTester t;
      OptimizerSettings s;
      s.base.batch_size = 32; s.base.time_steps = 10; s.base.motion_model = model;   // T - 1 < 20: the SG filter is a no-op
      s.base.vx_max = 1.0f; s.base.vx_min = -1.0f; s.base.vy_max = 0.6f; s.base.wz_max = 2.0f;
      s.base.temperature = 0.3f;
      s.controller_frequency = shift ? 1.0 / s.base.model_dt : 30.0;
      t.initialize(s, {}, bowtieRobot());   // no critics: every trajectory costs the same, the update is the noise mean
      const size_t n = 32 * 10;
      std::vector<float> zeros(n, 0.0f);
      t.setNoise(zeros.data(), zeros.data(), zeros.data());
      std::vector<float> vx(10), vy(10), wz(10);
      for (int i = 0; i < 10; ++i) {vx[i] = 0.25f + 0.01f * i; vy[i] = 0.5f - 0.01f * i; wz[i] = 0.1f * (i + 1);}
      MPPI_ABI(set_control_sequence)(t.handle(), vx.data(), vy.data(), wz.data());
      DummyCostmap cm;
      mppi_b200::Path plan;
      plan.x = {0.5f, 1.0f}; plan.y = {0.5f, 0.5f}; plan.yaw = {0.0f, 0.0f};
      const auto cmd = t.evalControl({0.5, 0.5, 0.0}, {0, 0, 0}, plan, {1.0, 0.5, 0.0}, -1.0, cm.view);
      const bool hol = model == MPPI_MODEL_OMNI;
      // zero noise + equal weights: the mean sequence is reproduced (wz clipped to +-2, here 0.1..1.0)
      const int o = shift ? 1 : 0;
      EXPECT_NEAR(cmd.vx, vx[o], 1e-6);
      EXPECT_NEAR(cmd.vy, hol ? vy[o] : 0.0, 1e-6);   // "Y should not be populated" for DiffDrive
      EXPECT_NEAR(cmd.wz, wz[o], 1e-6);
      // control_sequence_ after evalControl: rolled by one with the last element repeated when shifting is ON
      for (int i = 0; i < 10; ++i) {
        const int src = shift ? std::min(i + 1, 9) : i;
        EXPECT_NEAR(t.controlVx()[i], vx[src], 1e-6);
        EXPECT_NEAR(t.controlWz()[i], wz[src], 1e-6);
        // vy of a non-holonomic model is neither updated nor shifted
        EXPECT_NEAR(t.controlVy()[i], hol ? vy[src] : vy[i], 1e-6);
      }
    }
  }
}

// optimizer_smoke_test.cpp:48-116
static void testSmoke()
{
  struct Combo {int model; std::vector<int> critics;};
  const std::vector<Combo> combos = {
    {MPPI_MODEL_OMNI, {MPPI_CRITIC_GOAL, MPPI_CRITIC_GOAL_ANGLE, MPPI_CRITIC_OBSTACLES, MPPI_CRITIC_PATH_ALIGN, MPPI_CRITIC_TWIRLING,
        MPPI_CRITIC_PATH_FOLLOW, MPPI_CRITIC_PREFER_FORWARD}},
    {MPPI_MODEL_DIFF_DRIVE, {MPPI_CRITIC_GOAL, MPPI_CRITIC_GOAL_ANGLE, MPPI_CRITIC_COST, MPPI_CRITIC_PATH_ANGLE, MPPI_CRITIC_PATH_FOLLOW,
        MPPI_CRITIC_PREFER_FORWARD}},
    {MPPI_MODEL_ACKERMANN, {MPPI_CRITIC_GOAL, MPPI_CRITIC_GOAL_ANGLE, MPPI_CRITIC_OBSTACLES, MPPI_CRITIC_PATH_ANGLE, MPPI_CRITIC_PATH_FOLLOW,
        MPPI_CRITIC_PREFER_FORWARD}},
  };
  for (const auto & combo : combos) {
    OptimizerTester t;
    OptimizerSettings s;
    s.base.batch_size = 400; s.base.time_steps = 15; s.base.iteration_count = 1; s.base.motion_model = combo.model;
    s.controller_frequency = 30.0;
    std::vector<mppi_critic_desc> critics;
    for (int k : combo.critics) {
      critics.push_back(critic(k, [](mppi_critic_desc & d) {d.consider_footprint = 1;}));
    }
    EXPECT_NO_THROW(t.initialize(s, critics, bowtieRobot()));
    DummyCostmap cm;
    const unsigned offset = 4, obstacle_size = offset * 2;
    cm.addObstacle(20 - offset, 20 - offset, obstacle_size, 250);       // centre cell (20, 20), cost 250
    mppi_b200::Path plan;                                                // getIncrementalDummyPath: 50 points, step 0.1 in x and y
    for (unsigned i = 0; i < 50; ++i) {plan.x.push_back(2.0f + 0.1f * i); plan.y.push_back(2.0f + 0.1f * i); plan.yaw.push_back(0.0f);}
    const mppi_b200::Pose start{2.0, 2.0, 0.0}, goal{plan.x.back(), plan.y.back(), 0.0};
    mppi_b200::Twist cmd;
    EXPECT_NO_THROW(cmd = t.evalControl(start, {0, 0, 0}, plan, goal, -1.0, cm.view));
    EXPECT_TRUE(std::isfinite(cmd.vx) && std::isfinite(cmd.vy) && std::isfinite(cmd.wz));
    EXPECT_TRUE(cmd.vx <= s.base.vx_max + 1e-6 && cmd.vx >= s.base.vx_min - 1e-6 && std::fabs(cmd.wz) <= s.base.wz_max + 1e-6);
    // a second cycle from the warm-started sequence, and the visualisation getters
    EXPECT_NO_THROW(cmd = t.evalControl(start, {cmd.vx, cmd.vy, cmd.wz}, plan, goal, -1.0, cm.view));
    std::vector<float> x, y, yaw;
    EXPECT_NO_THROW(t.getOptimizedTrajectory(start));
    EXPECT_NO_THROW(t.reset());
  }
}

// all trajectories collide -> fallback resets and retries, then "Optimizer fail to compute path" (optimizer.cpp:166-183)
static void testAllCollideThrows()
{
  OptimizerTester t;
  OptimizerSettings s;
  s.base.batch_size = 128; s.base.time_steps = 15; s.base.motion_model = MPPI_MODEL_OMNI; s.retry_attempt_limit = 1;
  s.controller_frequency = 30.0;
  t.initialize(s, {critic(MPPI_CRITIC_OBSTACLES)}, bowtieRobot());
  DummyCostmap cm;
  std::fill(cm.cells.begin(), cm.cells.end(), 254);   // lethal everywhere
  mppi_b200::Path plan;
  plan.x = {2.0f, 2.5f}; plan.y = {2.0f, 2.0f}; plan.yaw = {0.0f, 0.0f};
  EXPECT_THROW(t.evalControl({2.0, 2.0, 0.0}, {0, 0, 0}, plan, {2.5, 2.0, 0.0}, -1.0, cm.view));
  // the reference's retry loop never re-runs prepare(), so fail_flag survives the retries (critic_manager.cpp:70-73):
  // retry_attempt_limit + 1 soft resets, then the throw -- on ANY map, even one that became free meanwhile
  EXPECT_TRUE(t.resetCount() == s.retry_attempt_limit + 1);
  EXPECT_TRUE(t.lastCycle().fail_flag != 0);
  // and the optimizer is usable again afterwards on a free map
  std::fill(cm.cells.begin(), cm.cells.end(), 0);
  EXPECT_NO_THROW(t.evalControl({2.0, 2.0, 0.0}, {0, 0, 0}, plan, {2.5, 2.0, 0.0}, -1.0, cm.view));
}

int main()
{
  testSetOffset();
  testInvalidModelThrows();
  testFallback();
  testTwistAndShift();
  testSmoke();
  testAllCollideThrows();
  if (g_failures) {std::printf("%d failure(s)\n", g_failures); return 1;}
  std::printf("all host-mirror tests passed\n");
  return 0;
}
