// ORACLE — TEST INFRASTRUCTURE ONLY. boost::thread subset used by the reference's headers (include/Object.hpp:4-5,21-23,
// include/MapPoint.hpp:3-9, include/Map.hpp:93), mapped onto the C++17 standard library. Boost is not installed here.
#pragma once
#include <mutex>
#include <shared_mutex>
namespace boost {
using mutex = std::mutex;
using shared_mutex = std::shared_mutex;
template <class M> using shared_lock = std::shared_lock<M>;
template <class M> using unique_lock = std::unique_lock<M>;
}  // namespace boost
