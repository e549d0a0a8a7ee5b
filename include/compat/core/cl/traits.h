// Overlay for core/cl/traits.h: only the cl_* scalar / vector type names are needed on the
// hot-path boundary.
#pragma once
#include "../../../wayverb_b200/cl_compat.hpp"
