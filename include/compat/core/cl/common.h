// Overlay for the reference's src/core/include/core/cl/common.h: same names
// (core::compute_context, read_value, write_value, read_from_buffer, items_in_buffer),
// implemented over libwvb200.so. Put `-I include/compat` BEFORE the reference's include
// directories and its unmodified headers resolve "core/cl/common.h" to this file.
#pragma once
#include "../../../wayverb_b200/core.hpp"
