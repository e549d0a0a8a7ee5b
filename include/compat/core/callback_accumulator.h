// Overlay for core/callback_accumulator.h.
#pragma once
#include "../../wayverb_b200/core.hpp"
