// Stand-in: the reference names DynamicAutoDiffCostFunction in a using-declaration only.
#pragma once
#include "ceres.h"
