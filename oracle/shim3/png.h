// empty stand-in for <png.h> (JointOptimization.h includes it for CImg's PNG writer, which the stand-in CImg does not have)
#pragma once
