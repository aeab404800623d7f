// spec_host.h — per-program specialisation of K1 at run time (EXPERIMENTAL: off unless the option "specialize"
// is set; compiles and links without a GPU, first validated on hardware in round 2).
//
// render_kernel.cu is compiled a second time, by NVRTC, with the render program of one voice group as a compile-time constant
// (EB_SPEC_PROGRAM, see the #ifdef in render_tile): the interpreter's per-op bodies are reused verbatim, dispatch and operand
// decoding fold away.  libnvrtc and libcuda are opened with dlopen, so the library has no load-time dependency on either.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>

#include "program.h"

namespace eb {

struct SpecKernel {
    void* module = nullptr;      // CUmodule
    void* function = nullptr;    // CUfunction of render_block_kernel<NITER, LOGL> specialised for one program
    std::vector<char> cubin;
    ~SpecKernel();
};

// Compile K1 for (tileWidth, niterOverride) against `code` (one single-stage program).  Fills out.cubin; when `load` is set also
// loads the module into the current CUDA context and resolves the kernel.  Returns false and a message in `log` on failure.
bool specialise_compile(const std::vector<uint32_t>& code, int tileWidth, int niterOverride, bool load, SpecKernel& out, std::string& log);

// cuLaunchKernel of a specialised kernel with the same launch geometry the built-in instantiation would get.
cudaError_t specialise_launch(const SpecKernel& k, const LaunchParams& P, int grid, int threads, size_t smem, int perWarpFloats, cudaStream_t stream);

}  // namespace eb
