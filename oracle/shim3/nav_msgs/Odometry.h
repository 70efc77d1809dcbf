// empty stand-in: HitLSLAM.cpp includes this header but uses nothing of it
#pragma once
