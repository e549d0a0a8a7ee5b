// TEST INFRASTRUCTURE ONLY. core/src/az_el.cpp includes glm/gtx/transform.hpp and uses none of its
// names (rotations); see ../glm.hpp.
#pragma once
#include "../glm.hpp"
