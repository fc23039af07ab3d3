// viennacl/tools/timer.hpp -- wall-clock timer (reference: tools/timer.hpp:90-114).
#ifndef VIENNACL_B200_TOOLS_TIMER_HPP
#define VIENNACL_B200_TOOLS_TIMER_HPP
#include <chrono>
namespace viennacl
{
namespace tools
{
  class timer
  {
  public:
    timer() : t0_(std::chrono::steady_clock::now()) {}
    void start() { t0_ = std::chrono::steady_clock::now(); }
    double get() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
  private:
    std::chrono::steady_clock::time_point t0_;
  };
}
}
#endif
