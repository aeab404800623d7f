// convolve.h — K3: partitioned-FFT convolver state (device side in convolve_kernel.cu).
#pragma once
#include <cstddef>
#include <cstdint>

namespace eb {

struct ConvolverState {
    // filled in by convolve_host.cpp
};

} // namespace eb
