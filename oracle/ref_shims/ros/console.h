/* rosconsole stand-in: assertions abort, log macros compile to nothing */
#ifndef REF_SHIM_ROS_CONSOLE
#define REF_SHIM_ROS_CONSOLE
#include <cstdio>
#include <cstdlib>
#define ROS_ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::abort(); } } while (0)
#define ROS_ASSERT_MSG(cond, ...) do { if (!(cond)) { std::fprintf(stderr, "ROS_ASSERT_MSG failed: %s (%s:%d): ", #cond, __FILE__, __LINE__); std::fprintf(stderr, __VA_ARGS__); std::abort(); } } while (0)
#define ROS_DEPRECATED __attribute__((deprecated))
#define ROS_DEBUG(...) do {} while (0)
#define ROS_DEBUG_COND(...) do {} while (0)
#define ROS_DEBUG_ONCE(...) do {} while (0)
#define ROS_INFO(...) do {} while (0)
#define ROS_INFO_ONCE(...) do {} while (0)
#define ROS_INFO_COND(...) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_WARN_ONCE(...) do {} while (0)
#define ROS_WARN_COND(...) do {} while (0)
#define ROS_WARN_THROTTLE(...) do {} while (0)
#define ROS_ERROR(...) do {} while (0)
#define ROS_ERROR_ONCE(...) do {} while (0)
#define ROS_ERROR_COND(...) do {} while (0)
#define ROS_ERROR_STREAM(...) do {} while (0)
#define ROS_WARN_STREAM(...) do {} while (0)
#define ROS_INFO_STREAM(...) do {} while (0)
#define ROS_DEBUG_STREAM(...) do {} while (0)
#endif
