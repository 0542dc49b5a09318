// math_constants.h stand-in (see cuda_runtime.h in this directory) -- TEST INFRASTRUCTURE ONLY.
#pragma once
#include <limits>
#define CUDART_INF (std::numeric_limits<double>::infinity())
#define CUDART_NAN (std::numeric_limits<double>::quiet_NaN())
