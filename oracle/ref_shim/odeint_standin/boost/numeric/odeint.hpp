// -*- C++ -*-
// STAND-IN for the two Boost.odeint facilities tendon/TendonRobot.cpp uses (test infrastructure).
//
// Boost is not installed in this image.  TendonRobot::tension_shape (TendonRobot.cpp:458-462) needs
//     boost::numeric::odeint::runge_kutta4<State>       and
//     boost::numeric::odeint::integrate_times(stepper, system, x, t_begin, t_end, dt, observer)
// This header restates them from odeint's published algorithm so that the reference's OWN
// tension_shape text compiles (oracle/Makefile, _ref/libtendonrobot_ref.so).  It therefore pins
// everything tension_shape does around the integrator (grid, initial state, observer, result
// assembly, convergence flag) but NOT odeint itself:
//   * integrate_times, stepper overload (odeint/integrate/detail/integrate_times.hpp): observe at
//     every time of the list; between two list times repeat
//         while (less_with_sign(t, t_next, dt))   { h = min_abs(dt, t_next - t); do_step(h); t += h;
//                                                  dt' = max_abs(dt, h) }
//     with less_with_sign(t1, t2, dt > 0) := (t2 - t1) > numeric_limits<T>::epsilon();
//   * runge_kutta4 = explicit_generic_rk<4> with the classical tableau (a21 = a32 = 1/2, a43 = 1,
//     b = 1/6, 1/3, 1/3, 1/6, c = 0, 1/2, 1/2, 1); stages are x + h a_ij k_j, the update is the
//     left-to-right sum x + (h b1) k1 + (h b2) k2 + (h b3) k3 + (h b4) k4 of the generic algebra.
#ifndef IRT_ODEINT_STANDIN
#define IRT_ODEINT_STANDIN

#include <cstddef>
#include <cmath>
#include <limits>

namespace boost { namespace numeric { namespace odeint {

template <class State> class runge_kutta4 {
 public:
  template <class System> void do_step(System system, State &x, double t, double dt) {
    const std::size_t n = x.size();
    k1_.resize(n); k2_.resize(n); k3_.resize(n); k4_.resize(n); xt_.resize(n);
    system(x, k1_, t);
    for (std::size_t i = 0; i < n; i++) xt_[i] = x[i] + (dt * 0.5) * k1_[i];
    system(xt_, k2_, t + dt * 0.5);
    for (std::size_t i = 0; i < n; i++) xt_[i] = x[i] + (dt * 0.5) * k2_[i];
    system(xt_, k3_, t + dt * 0.5);
    for (std::size_t i = 0; i < n; i++) xt_[i] = x[i] + dt * k3_[i];
    system(xt_, k4_, t + dt);
    const double b1 = dt * (1.0 / 6.0), b2 = dt * (1.0 / 3.0), b3 = dt * (1.0 / 3.0), b4 = dt * (1.0 / 6.0);
    for (std::size_t i = 0; i < n; i++)
      x[i] = x[i] + b1 * k1_[i] + b2 * k2_[i] + b3 * k3_[i] + b4 * k4_[i];
  }

 private:
  State k1_, k2_, k3_, k4_, xt_;
};

template <class Stepper, class System, class State, class TimeIterator, class Time, class Observer>
std::size_t integrate_times(Stepper stepper, System system, State &x, TimeIterator t_begin,
                            TimeIterator t_end, Time dt, Observer observer) {
  std::size_t steps = 0;
  Time current_dt = dt;
  if (t_begin == t_end) return 0;
  while (true) {
    Time current_time = *t_begin++;
    observer(x, current_time);
    if (t_begin == t_end) break;
    while ((static_cast<Time>(*t_begin) - current_time) > std::numeric_limits<Time>::epsilon()) {
      const Time rest = static_cast<Time>(*t_begin) - current_time;
      current_dt = (std::abs(dt) < std::abs(rest)) ? dt : rest;              // min_abs
      stepper.do_step(system, x, current_time, current_dt);
      ++steps;
      current_time += current_dt;
      current_dt = (std::abs(dt) > std::abs(current_dt)) ? dt : current_dt;  // max_abs
    }
  }
  return steps;
}

}}}  // namespace boost::numeric::odeint

#endif  // IRT_ODEINT_STANDIN
