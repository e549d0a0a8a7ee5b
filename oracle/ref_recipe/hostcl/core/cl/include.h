// TEST INFRASTRUCTURE ONLY. core/cl/include.h for the units that run the reference's HOST templates
// (waveguide::run) on the host: the OpenCL typedefs of ../../../hoststubs/core/cl/include.h plus the
// host-memory stand-in for cl.hpp.
#pragma once
#include "../../../hoststubs/core/cl/include.h"
#include "CL/cl.hpp"
