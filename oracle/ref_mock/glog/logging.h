// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for glog: LOG(x) << ... swallows its operands, CHECK aborts.
#pragma once
#include <cstdlib>
#include <ostream>
namespace mock_glog {
struct NullStream {
    template <class T> NullStream& operator<<(const T&) { return *this; }
    NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}
#define LOG(sev) ::mock_glog::NullStream()
#define VLOG(n) ::mock_glog::NullStream()
#define LOG_IF(sev, c) ::mock_glog::NullStream()
#define CHECK(c) if (!(c)) std::abort(); else ::mock_glog::NullStream()
