// TEST INFRASTRUCTURE ONLY. core/conversions.h (to_vec3 & co.) is included by
// core/spatial_division/range.h; the voxelisation code compiled here uses none of it.
#pragma once
#include "core/cl/include.h"
#include "glm/glm.hpp"
