// Minimal stand-in for <ceres/ceres.h> (see jet.h).
#ifndef REF_SHIM_CERES_CERES
#define REF_SHIM_CERES_CERES
#include "ceres/jet.h"
#endif
