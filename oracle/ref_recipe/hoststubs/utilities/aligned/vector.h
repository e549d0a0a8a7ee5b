// TEST INFRASTRUCTURE ONLY. utilities/aligned/vector.h: std::vector with an aligned allocator in the
// reference; a plain std::vector holds the same values.
#pragma once
#include <vector>
namespace util {
namespace aligned {
template <typename T>
using vector = std::vector<T>;
}  // namespace aligned
}  // namespace util
