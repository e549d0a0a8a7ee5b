// Overlay for core/exceptions.h:22-30 (value_is_nan / value_is_inf).
#pragma once
#include "../../wayverb_b200/core.hpp"
