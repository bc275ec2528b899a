// train.cuh -- training-path state and kernels (Gram accumulation, ridge solve).
#pragma once
#include <cuda_runtime.h>
#include <vector>

namespace sml {

struct TrainState {
    bool active = false;
};

inline void train_release(TrainState &t) { t.active = false; }

}  // namespace sml
