// compatibility include (reference: include/ggnn/base/data.cuh): DataType / DataLocation live in <ggnn/ggnn.hpp>
#pragma once
#include <ggnn/ggnn.hpp>
