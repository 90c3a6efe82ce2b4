// ORACLE — TEST INFRASTRUCTURE ONLY. `pyp` is the reference author's private bundle (fmt / yaml / timer; CMakeLists.txt:13-16),
// absent from /root/reference. fmt::print is only used for diagnostics on the path (src/Object.cpp:32,136, ORBExtractor.cpp:33).
#pragma once
#include <cstdio>
#include <string>
namespace fmt {
template <typename... A> static inline void print(const char* s, A&&...) { std::fputs(s, stderr); }
template <typename... A> static inline void print(const std::string& s, A&&...) { std::fputs(s.c_str(), stderr); }
template <typename... A> static inline std::string format(const char* s, A&&...) { return std::string(s); }
}  // namespace fmt
