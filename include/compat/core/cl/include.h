// Overlay for the reference's src/core/include/core/cl/include.h (which includes CL/cl.hpp):
// the few cl:: host names its headers spell out, as handles of libwvb200.so.
#pragma once
#include "../../../wayverb_b200/cl_compat.hpp"
