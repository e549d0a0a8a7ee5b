// TEST INFRASTRUCTURE ONLY. Stands in for core/geo/box.h:1-28 of the reference: the type alias and
// the two declarations the voxelisation code needs. The real header also pulls scene_data.h
// (assimp-facing containers) and declares ray/box intersection helpers that are not compiled here.
#pragma once
#include <functional>

#include "core/cl/include.h"
#include "glm/glm.hpp"
#include "utilities/aligned/vector.h"
#include "utilities/range.h"

namespace wayverb {
namespace core {
namespace geo {
struct triangle_vec3;
class ray;
using box = util::range<glm::vec3>;
bool overlaps(const box& b, const triangle_vec3& t);
}  // namespace geo
}  // namespace core
}  // namespace wayverb
