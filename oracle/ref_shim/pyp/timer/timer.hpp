// ORACLE — TEST INFRASTRUCTURE ONLY. MyTimer::Timer scope timers of the reference ("KL", "ORBE", "SMatch": src/Frame.cpp:91,119,135).
// Accumulates wall time per name so that the _ref bench can report the reference's own stage split.
#pragma once
#include <chrono>
#include <map>
#include <mutex>
#include <string>
namespace MyTimer {
struct Registry {
    std::mutex m;
    std::map<std::string, std::pair<double, long>> acc;
    static Registry& get() { static Registry r; return r; }
};
class Timer {
   public:
    explicit Timer(const std::string& name) : name_(name), t0_(std::chrono::steady_clock::now()) {}
    ~Timer() {
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
        Registry& r = Registry::get();
        std::lock_guard<std::mutex> g(r.m);
        auto& e = r.acc[name_];
        e.first += s; e.second += 1;
    }
    void tock() {}
   private:
    std::string name_;
    std::chrono::steady_clock::time_point t0_;
};
}  // namespace MyTimer
