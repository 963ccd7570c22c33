// compatibility include (reference: include/ggnn/base/eval.h): Evaluator / Evaluation live in <ggnn/ggnn.hpp>
#pragma once
#include <ggnn/ggnn.hpp>
