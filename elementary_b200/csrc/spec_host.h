// spec_host.h — per-program specialisation of K1 at run time (EXPERIMENTAL: off unless the option "specialize"
// is set; compiles and links without a GPU, first validated on hardware in round 2).
//
// render_kernel.cu is compiled a second time, by NVRTC, with the render program of one voice group as a compile-time constant
// (EB_SPEC_PROGRAM, see the #ifdef in render_tile): the interpreter's per-op bodies are reused verbatim, dispatch and operand
// decoding fold away.  libnvrtc and libcuda are opened with dlopen, so the library has no load-time dependency on either.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "program.h"

namespace eb {

struct SpecKernel {
    void* module = nullptr;      // CUmodule
    void* function = nullptr;    // CUfunction of render_block_kernel<NITER, LOGL> specialised for one program
    std::vector<char> cubin;
    std::string loweredName;     // mangled name of the kernel inside the cubin
    ~SpecKernel();
};

// Step 1 — pure compilation (NVRTC only; thread safe, needs no CUDA context, works without a GPU): K1 for (tileWidth,
// niterOverride) against `code` (one single-stage program).  Fills out.cubin / out.loweredName.
bool specialise_compile(const std::vector<uint32_t>& code, int tileWidth, int niterOverride, SpecKernel& out, std::string& log);
// Step 2 — on a thread whose CUDA context is current: load the cubin and resolve the kernel (milliseconds).
bool specialise_load(SpecKernel& k, std::string& log);

// A compilation running on its own thread: the interpreter serves the voice group until the cubin is there, so a live graph edit
// never waits for the compiler.  state: 0 compiling, 1 compiled (cubin ready, not loaded), 2 loaded, -1 failed.
struct SpecJob {
    std::atomic<int> state{0};
    SpecKernel kernel;
    std::string log;
    std::thread worker;
    ~SpecJob() { if (worker.joinable()) worker.join(); }
};
std::shared_ptr<SpecJob> specialise_async(std::vector<uint32_t> code, int tileWidth, int niterOverride);

// cuLaunchKernel of a specialised kernel with the same launch geometry the built-in instantiation would get.
cudaError_t specialise_launch(const SpecKernel& k, const LaunchParams& P, int grid, int threads, size_t smem, int perWarpFloats, cudaStream_t stream);

}  // namespace eb
