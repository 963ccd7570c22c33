// compatibility include (reference: include/ggnn/base/dataset.cuh): GenericDataset / Dataset<T> / Results live in <ggnn/ggnn.hpp>
#pragma once
#include <ggnn/ggnn.hpp>
