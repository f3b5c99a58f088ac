// Collection.h -- the variant the reference dispatches on (reference src/Collection.h:28-34),
// without Volumes: brick elements only enter the path through adapters that turn them into
// temporary Points (reference src/Influence.h:810-869) and are out of scope here.
#pragma once
#ifndef OMEGA2D_REFERENCE_HEADERS
#include "Points.h"
#include "Surfaces.h"
#include <variant>

using Collection = std::variant<Points<float>, Surfaces<float>>;
#endif
