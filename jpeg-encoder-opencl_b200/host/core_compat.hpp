// core_compat.hpp -- the two names of the reference's vendored lib/Core that its driver uses for timing
// (Core::TimeSpan, lib/Core/TimeSpan.hpp:35-79; Core::getCurrentTime, lib/Core/Time.cpp:65-69), so that
// JpegEncoderHost (src/OpenCLProject_JpegEncoder.cpp:28-250) compiles against utils_compat.hpp without the
// Boost-dependent lib/Core tree.  Wall clock in whole microseconds, like the reference's gettimeofday.
#pragma once
#include <chrono>
#include <cstdint>
#include <string>

namespace Core {
class TimeSpan {
    int64_t us_;

public:
    explicit TimeSpan(int64_t us = 0) : us_(us) {}
    int64_t getMicroseconds() const { return us_; }
    double getMilliseconds() const { return us_ / 1e3; }
    double getSeconds() const { return us_ / 1e6; }
    std::string toString(bool appendUnit = true) const {  // seconds with six decimals, as the reference prints them
        char buf[48];
        snprintf(buf, sizeof buf, "%s%lld.%06lld%s", us_ < 0 ? "-" : "", (long long)((us_ < 0 ? -us_ : us_) / 1000000),
                 (long long)((us_ < 0 ? -us_ : us_) % 1000000), appendUnit ? "s" : "");
        return buf;
    }
    TimeSpan operator+(TimeSpan o) const { return TimeSpan(us_ + o.us_); }
    TimeSpan operator-(TimeSpan o) const { return TimeSpan(us_ - o.us_); }
    bool operator<(TimeSpan o) const { return us_ < o.us_; }
    bool operator==(TimeSpan o) const { return us_ == o.us_; }
};
inline TimeSpan getCurrentTime() {
    using namespace std::chrono;
    return TimeSpan(duration_cast<microseconds>(system_clock::now().time_since_epoch()).count());
}
}  // namespace Core
