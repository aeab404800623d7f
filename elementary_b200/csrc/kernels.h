// kernels.h — host-callable launchers of the device kernels (render_kernel.cu, convolve_kernel.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include "program.h"

namespace eb {

// K1: fused render-sequence kernel, one launch per voice group per block.
int render_niter_for(int tileWidth, int niterOverride);   // elements per lane per sample tile (E = 32*NITER = L*T)
size_t render_smem_bytes(int nSlots, int nOut, int nStateRows, int nParams, int warpsPerCta, int tileWidth, int niterOverride);
cudaError_t launch_render_block(const LaunchParams& P, int warpsPerCta, int niterOverride, cudaStream_t stream);

// K1 for many voice groups of one tile geometry in one launch (descs / tileStart are device pointers).
cudaError_t launch_render_groups(const LaunchParams* descs, const int* tileStart, int nGroups, int totalTiles, int tileWidth,
                                 int maxSlots, int nOut, int maxStateRows, int maxParams, int warpsPerCta, cudaStream_t stream);

// K2: deterministic reduction of per-tile partial mixes into the [nOut][blockSize] mix bus.
cudaError_t launch_mix_reduce(const float* partial, float* out, int nTiles, int nOut, int blockSize, int numSamples, cudaStream_t stream);

} // namespace eb
