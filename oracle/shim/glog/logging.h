// Empty stand-in for <glog/logging.h>: kdtree.cpp includes it but uses nothing from it.
#pragma once
