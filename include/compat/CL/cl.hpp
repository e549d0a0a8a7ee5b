// Overlay for CL/cl.hpp, for reference headers that include it directly.
#pragma once
#include "../../wayverb_b200/cl_compat.hpp"
