// The per-observation arithmetic of K1 lives in the public header include/rsba_reproj_math.h (one source for
// the device kernels and for the host functor structs of include/rsba_cuda_functors.hpp).
#pragma once
#include "../../include/rsba_reproj_math.h"
#include "common.cuh"
