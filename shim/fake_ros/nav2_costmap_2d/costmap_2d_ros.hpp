#include "../fake_ros.hpp"
