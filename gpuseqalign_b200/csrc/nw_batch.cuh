// nw_batch.cuh -- many short independent pairs (BASELINE config 3).
#pragma once
#include "nw_common.cuh"
namespace nwb {
}  // namespace nwb
