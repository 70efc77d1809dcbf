// See oracle/shim2/ceres/jet.h (the same stand-in Jet serves both shim builds).
#pragma once
#include "../../shim2/ceres/jet.h"
