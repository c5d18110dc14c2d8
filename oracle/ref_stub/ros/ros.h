// Stand-in for roscpp (absent): FullSystem.h declares node handles / subscribers as members; nothing on the hot path calls them.
#pragma once
#include <string>
namespace ros {
struct Time { double t = 0; double toSec() const { return t; } static Time now() { return Time(); } };
struct Subscriber {}; struct Publisher { template <class M> void publish(const M&) const {} };
struct NodeHandle { NodeHandle() {} NodeHandle(const char*) {} template <class T> bool getParam(const std::string&, T&) const { return false; } template <class M = void, class... A> Subscriber subscribe(const std::string&, int, A...) { return Subscriber(); } template <class M> Publisher advertise(const std::string&, int) { return Publisher(); } };
inline bool ok() { return true; } inline void spinOnce() {} inline void spin() {} inline void init(int&, char**, const std::string&) {}
struct Rate { Rate(double) {} void sleep() {} };
}
#define ROS_INFO(...) do {} while (0)
#define ROS_WARN(...) do {} while (0)
#define ROS_ERROR(...) do {} while (0)
