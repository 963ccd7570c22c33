// compatibility include (reference: include/ggnn/base/def.h): DistanceMeasure lives in <ggnn/ggnn.hpp>
#pragma once
#include <ggnn/ggnn.hpp>
