#pragma once
#include <chrono>
namespace boost {
class timer {
public:
    timer() { restart(); }
    void restart() { t0_ = std::chrono::steady_clock::now(); }
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
private:
    std::chrono::steady_clock::time_point t0_;
};
}
