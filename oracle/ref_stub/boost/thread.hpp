// Stand-in for Boost.Thread (absent): just enough for util/IndexThreadReduce.h to COMPILE.  The oracle/_ref build never runs the
// reference's worker threads (multiThreading=false path: reduce() executes callPerIndex inline).
#pragma once
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
namespace boost {
typedef std::mutex mutex;
template <class M> using unique_lock = std::unique_lock<M>;
typedef std::condition_variable condition_variable;
class thread { public: std::thread t; thread() {} template <class F> explicit thread(F f) : t(f) {} template <class F, class... A> thread(F f, A... a) : t(f, a...) {} void join() { if (t.joinable()) t.join(); } };
using std::function;
// boost::bind: std::bind for callables, plus the one composition the reference uses — bind(&pair::second, _1) < bind(&pair::second, _2) (Reprojector.cpp:129)
template <class F, class... A> auto bind(F f, A... a) -> decltype(std::bind(f, a...)) { return std::bind(f, a...); }
template <class M, class C, int I> struct memb_bind { M C::* p; template <class A, class B> const M& operator()(const A& a, const B& b) const { return pick(a, b, std::integral_constant<int, I>()).*p; }
  template <class A, class B> static const A& pick(const A& a, const B&, std::integral_constant<int, 1>) { return a; } template <class A, class B> static const B& pick(const A&, const B& b, std::integral_constant<int, 2>) { return b; } };
template <class M, class C> memb_bind<M, C, 1> bind(M C::* p, const decltype(std::placeholders::_1)&) { return memb_bind<M, C, 1>{p}; }
template <class M, class C> memb_bind<M, C, 2> bind(M C::* p, const decltype(std::placeholders::_2)&) { return memb_bind<M, C, 2>{p}; }
template <class L, class R> struct less_bind { L l; R r; template <class A, class B> bool operator()(const A& a, const B& b) const { return l(a, b) < r(a, b); } };
template <class M, class C, int I, class M2, class C2, int I2> less_bind<memb_bind<M, C, I>, memb_bind<M2, C2, I2> > operator<(const memb_bind<M, C, I>& l, const memb_bind<M2, C2, I2>& r) { return less_bind<memb_bind<M, C, I>, memb_bind<M2, C2, I2> >{l, r}; }
}
using namespace std::placeholders;
