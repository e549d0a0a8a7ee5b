// Overlay for utilities/aligned/vector.h: util::aligned::vector.
#pragma once
#include "../../../wayverb_b200/core.hpp"
