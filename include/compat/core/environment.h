// Overlay for core/environment.h:6-13.
#pragma once
#include "../../wayverb_b200/core.hpp"
