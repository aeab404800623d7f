// spec_host.cpp — see spec_host.h.
#include "spec_host.h"

#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>

#include "kernels.h"

namespace eb {

namespace {

struct Api {
    void* hNvrtc = nullptr;
    void* hCuda = nullptr;
    decltype(&nvrtcCreateProgram) createProgram = nullptr;
    decltype(&nvrtcDestroyProgram) destroyProgram = nullptr;
    decltype(&nvrtcAddNameExpression) addNameExpression = nullptr;
    decltype(&nvrtcCompileProgram) compileProgram = nullptr;
    decltype(&nvrtcGetProgramLogSize) getLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) getLog = nullptr;
    decltype(&nvrtcGetCUBINSize) getCubinSize = nullptr;
    decltype(&nvrtcGetCUBIN) getCubin = nullptr;
    decltype(&nvrtcGetLoweredName) getLoweredName = nullptr;
    decltype(&cuModuleLoadData) moduleLoadData = nullptr;
    decltype(&cuModuleUnload) moduleUnload = nullptr;
    decltype(&cuModuleGetFunction) moduleGetFunction = nullptr;
    decltype(&cuFuncSetAttribute) funcSetAttribute = nullptr;
    decltype(&cuLaunchKernel) launchKernel = nullptr;
};

template <typename F>
bool sym(void* h, const char* name, F& f) {
    f = reinterpret_cast<F>(dlsym(h, name));
    return f != nullptr;
}

Api& api() {
    static Api a;
    return a;
}

bool loadNvrtc(std::string& log) {
    Api& a = api();
    if (a.hNvrtc) return true;
    for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
        a.hNvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (a.hNvrtc) break;
    }
    if (!a.hNvrtc) { log = "libnvrtc not found (dlopen)"; return false; }
    const bool ok = sym(a.hNvrtc, "nvrtcCreateProgram", a.createProgram) && sym(a.hNvrtc, "nvrtcDestroyProgram", a.destroyProgram) &&
                    sym(a.hNvrtc, "nvrtcAddNameExpression", a.addNameExpression) && sym(a.hNvrtc, "nvrtcCompileProgram", a.compileProgram) &&
                    sym(a.hNvrtc, "nvrtcGetProgramLogSize", a.getLogSize) && sym(a.hNvrtc, "nvrtcGetProgramLog", a.getLog) &&
                    sym(a.hNvrtc, "nvrtcGetCUBINSize", a.getCubinSize) && sym(a.hNvrtc, "nvrtcGetCUBIN", a.getCubin) &&
                    sym(a.hNvrtc, "nvrtcGetLoweredName", a.getLoweredName);
    if (!ok) { log = "libnvrtc lacks a required entry point"; a.hNvrtc = nullptr; }
    return ok;
}

bool loadDriver(std::string& log) {
    Api& a = api();
    if (a.hCuda) return true;
    a.hCuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!a.hCuda) { log = "libcuda.so.1 not found (no CUDA driver on this machine)"; return false; }
    const bool ok = sym(a.hCuda, "cuModuleLoadData", a.moduleLoadData) && sym(a.hCuda, "cuModuleUnload", a.moduleUnload) &&
                    sym(a.hCuda, "cuModuleGetFunction", a.moduleGetFunction) && sym(a.hCuda, "cuFuncSetAttribute", a.funcSetAttribute) &&
                    sym(a.hCuda, "cuLaunchKernel", a.launchKernel);
    if (!ok) { log = "libcuda lacks a required entry point"; a.hCuda = nullptr; }
    return ok;
}

// the directory this shared library was loaded from: <repo>/elementary_b200 ; the sources travel next to it in csrc/
std::string sourceDir() {
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&sourceDir), &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t slash = p.find_last_of('/');
        return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/csrc";
    }
    return "elementary_b200/csrc";
}

}  // namespace

SpecKernel::~SpecKernel() {
    if (module && api().moduleUnload) api().moduleUnload(static_cast<CUmodule>(module));
}

bool specialise_compile(const std::vector<uint32_t>& code, int tileWidth, int niterOverride, SpecKernel& out, std::string& log) {
    if (!loadNvrtc(log)) return false;
    Api& a = api();
    const std::string dir = sourceDir();
    std::ifstream f(dir + "/render_kernel.cu");
    if (!f) { log = "cannot read " + dir + "/render_kernel.cu"; return false; }
    std::stringstream src;
    src << f.rdbuf();

    std::ostringstream hdr;   // the program as a constant: what EB_SPEC_PROGRAM / EB_SPEC_CODE stand for in render_tile
    hdr << "#pragma once\n#include \"rtc_compat.h\"\nnamespace eb {\n__device__ constexpr uint32_t EB_SPEC_CODE[] = {";
    for (size_t i = 0; i < code.size(); ++i) hdr << (i ? "," : "") << "0x" << std::hex << code[i] << "u";
    hdr << "};\nconstexpr int EB_SPEC_CODE_LEN = " << std::dec << code.size() << ";\n}\n#define EB_SPEC_PROGRAM 1\n";
    const std::string hdrText = hdr.str(), srcText = src.str();
    const char* hdrSrc[] = {hdrText.c_str()};
    const char* hdrNames[] = {"eb_spec_program.h"};

    nvrtcProgram prog = nullptr;
    if (a.createProgram(&prog, srcText.c_str(), "render_kernel.cu", 1, hdrSrc, hdrNames) != NVRTC_SUCCESS) { log = "nvrtcCreateProgram failed"; return false; }
    int logl = 0;
    for (int l = tileWidth; l > 1; l >>= 1) ++logl;
    const int niter = render_niter_for(tileWidth, niterOverride);
    const std::string name = "eb::render_block_kernel<" + std::to_string(niter) + ", " + std::to_string(logl) + ">";
    a.addNameExpression(prog, name.c_str());
    const std::string inc = "--include-path=" + dir;
    const char* opts[] = {"--std=c++20", "--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "-default-device",
                          inc.c_str(), "--pre-include=eb_spec_program.h", "-diag-suppress=186,68,179"};
    const nvrtcResult rc = a.compileProgram(prog, (int) (sizeof(opts) / sizeof(opts[0])), opts);
    size_t n = 0;
    a.getLogSize(prog, &n);
    if (n > 1) { log.resize(n); a.getLog(prog, &log[0]); }
    bool ok = rc == NVRTC_SUCCESS;
    std::string lowered;
    if (ok) {
        size_t sz = 0;
        ok = a.getCubinSize(prog, &sz) == NVRTC_SUCCESS && sz > 0;
        if (ok) { out.cubin.resize(sz); ok = a.getCubin(prog, out.cubin.data()) == NVRTC_SUCCESS; }
        const char* ln = nullptr;
        if (ok && a.getLoweredName(prog, name.c_str(), &ln) == NVRTC_SUCCESS && ln) lowered = ln; else ok = false;
    }
    a.destroyProgram(&prog);
    if (!ok) { if (log.empty()) log = "NVRTC compilation failed"; return false; }
    out.loweredName = lowered;
    return true;
}

bool specialise_load(SpecKernel& k, std::string& log) {
    if (k.function) return true;
    if (k.cubin.empty() || k.loweredName.empty()) { log = "nothing compiled"; return false; }
    if (!loadDriver(log)) return false;
    Api& a = api();
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    if (a.moduleLoadData(&mod, k.cubin.data()) != CUDA_SUCCESS) { log = "cuModuleLoadData failed"; return false; }
    if (a.moduleGetFunction(&fn, mod, k.loweredName.c_str()) != CUDA_SUCCESS) { a.moduleUnload(mod); log = "cuModuleGetFunction failed for " + k.loweredName; return false; }
    k.module = mod;
    k.function = fn;
    return true;
}

std::shared_ptr<SpecJob> specialise_async(std::vector<uint32_t> code, int tileWidth, int niterOverride) {
    auto job = std::make_shared<SpecJob>();
    SpecJob* j = job.get();   // the job outlives its worker: ~SpecJob joins
    job->worker = std::thread([j, code = std::move(code), tileWidth, niterOverride]() {
        const bool ok = specialise_compile(code, tileWidth, niterOverride, j->kernel, j->log);
        j->state.store(ok ? 1 : -1, std::memory_order_release);
    });
    return job;
}

cudaError_t specialise_launch(const SpecKernel& k, const LaunchParams& P, int grid, int threads, size_t smem, int perWarpFloats, cudaStream_t stream) {
    Api& a = api();
    if (!k.function || !a.launchKernel) return cudaErrorInvalidDeviceFunction;
    CUfunction fn = static_cast<CUfunction>(k.function);
    if (a.funcSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int) smem) != CUDA_SUCCESS) return cudaErrorInvalidValue;
    LaunchParams params = P;
    int perWarp = perWarpFloats;
    void* args[] = {&params, &perWarp};
    const CUresult r = a.launchKernel(fn, (unsigned) grid, 1, 1, (unsigned) threads, 1, 1, (unsigned) smem, reinterpret_cast<CUstream>(stream), args, nullptr);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorLaunchFailure;
}

}  // namespace eb
