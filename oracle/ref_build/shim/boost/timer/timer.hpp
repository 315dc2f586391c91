// Minimal stand-in for boost::timer::cpu_timer (oracle build only; used by the
// reference's spy cursor and file utilities, never by arithmetic).
#ifndef ORACLE_SHIM_BOOST_TIMER_HPP_
#define ORACLE_SHIM_BOOST_TIMER_HPP_
#include <stdint.h>
#include <time.h>
namespace boost { namespace timer {
typedef int64_t nanosecond_type;
struct cpu_times { nanosecond_type wall, user, system; };
class cpu_timer {
 public:
  cpu_timer() : stopped_(false) { start(); }
  bool is_stopped() const { return stopped_; }
  void start() { stopped_ = false; acc_ = cpu_times{0, 0, 0}; base_ = now(); }
  void stop() { if (!stopped_) { acc_ = elapsed(); stopped_ = true; } }
  void resume() { if (stopped_) { base_ = now(); stopped_ = false; } }
  cpu_times elapsed() const {
    if (stopped_) return acc_;
    cpu_times n = now();
    return cpu_times{acc_.wall + n.wall - base_.wall, acc_.user + n.user - base_.user,
                     acc_.system};
  }
 private:
  static cpu_times now() {
    timespec w, c;
    clock_gettime(CLOCK_MONOTONIC, &w);
    clock_gettime(CLOCK_PROCESS_CPUTIME_ID, &c);
    return cpu_times{w.tv_sec * 1000000000LL + w.tv_nsec, c.tv_sec * 1000000000LL + c.tv_nsec, 0};
  }
  bool stopped_;
  cpu_times acc_, base_;
};
}}
#endif
