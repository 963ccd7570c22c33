// compatibility include: programs written against the reference (`#include <ggnn/base/ggnn.cuh>`) get the
// B200 implementation's host API
#pragma once
#include <ggnn/ggnn.hpp>
