// kernels.h — host-callable launchers of the device kernels (render_kernel.cu, convolve_kernel.cu).
#pragma once
#include "program.h"
#ifndef __CUDACC_RTC__   // host-callable launchers: not part of a run-time compiled specialisation of K1
#include <cuda_runtime.h>

namespace eb {

// K1: fused render-sequence kernel, one launch per voice group per block.
int render_niter_for(int tileWidth, int niterOverride);   // elements per lane per sample tile (E = 32*NITER = L*T)
size_t render_smem_bytes(int nSlots, int nOut, int nStateRows, int nParams, int warpsPerCta, int tileWidth, int niterOverride);
struct SpecKernel;   // spec_host.h: K1 compiled at run time against one program (nullptr = the built-in interpreter)
cudaError_t launch_render_block(const LaunchParams& P, int warpsPerCta, int niterOverride, cudaStream_t stream, const SpecKernel* spec = nullptr);

// K1 for many voice groups of one tile geometry in one launch (descs / tileStart are device pointers).
cudaError_t launch_render_groups(const LaunchParams* descs, const int* tileStart, int nGroups, int totalTiles, int tileWidth,
                                 int maxSlots, int nOut, int maxStateRows, int maxParams, int warpsPerCta, long long sampleTime, int outOffset, cudaStream_t stream,
                                 int niterOverride = 0);

// K1 for many ONE-VOICE graphs whose programs were cut into pipeline stages (LaunchParams::pipeW): one CTA of `stages` warps per graph.
cudaError_t launch_render_groups_pipe(const LaunchParams* descs, const int* tileStart, int nGroups, int totalTiles, int stages,
                                      int maxSlots, int nOut, int maxStateRows, int maxParams, long long sampleTime, int outOffset, cudaStream_t stream,
                                      int niterOverride = 0);

// K2: deterministic reduction of per-tile partial mixes into the [nOut][blockSize] mix bus.
// scratch: [MIX_REDUCE_MAX_GROUPS][nOut][blockSize] floats, tickets: [nOut * ceil(blockSize/32)] zeroed counters (both may be null:
// single-pass reduction).
constexpr int MIX_REDUCE_MAX_GROUPS = 16;
// Host delivery of the finished mix bus: the kernel that writes the FINAL mix (K2 on one GPU, K4 across GPUs) also stores it into a
// mapped pinned host buffer and, once every CTA has done so (`done` counter), raises `flag = seq` there — Runtime::process polls that
// word instead of paying a D2H copy launch + stream synchronize per block.  out == nullptr: no host delivery.
struct HostDeliver {
    float* out = nullptr;            // device alias of the mapped host buffer [nOut][blockSize]
    uint32_t* flag = nullptr;        // device alias of the host sequence word
    unsigned int* done = nullptr;    // device counter, zero between launches
    uint32_t seq = 0;
};
cudaError_t launch_mix_reduce(const float* partial, float* out, float* scratch, unsigned int* tickets, int nTiles, int nOut, int blockSize,
                              int numSamples, cudaStream_t stream, HostDeliver hd = HostDeliver{});

// K4: the one collective of the path as our own kernel over NVLink/NVSwitch peer memory: K2 leaves this rank's
// partial mix bus in PeerMix::own; world-1 CTAs each push it into one peer's buffer as (sample, epoch) pairs, then every CTA waits for the
// pairs of all sources to arrive in its OWN buffer and sums a share of the samples in rank order — one launch, deterministic, no NCCL
// call on the data path, one NVLink one-way latency between "my partial is ready" and "the peer can use it".  count = 0 makes it a
// cross-GPU stream barrier (flags).
constexpr int MAX_PEERS = 8;
struct PeerMix {
    uint2* slot[MAX_PEERS];       // slot[p]: rank p's exchange buffer [2 parities][MAX_PEERS sources][stride] of (sample bits, epoch) pairs
                                  // (peer-mapped for p != rank): every 8-byte store carries its own arrival flag, so the data needs no
                                  // fence and no separate flag store behind it (the "LL" form NCCL uses for small messages)
    uint32_t* flag[MAX_PEERS];    // flag[p]: rank p's flags [2][MAX_PEERS] — used by the pure barrier (count = 0) only
    float* own;                   // this rank's partial mix bus [stride] (local; K2 leaves it here)
    int rank, world, stride;
};
cudaError_t launch_mix_exchange(const PeerMix& pm, float* mix, int count, uint32_t epoch, int* status, cudaStream_t stream, HostDeliver hd = HostDeliver{});

// A/B builds (-DEB_OPPROF): cycles / dispatches per opcode of the interpreter; zeros in the product library.
cudaError_t debug_opprof_read(unsigned long long* out128, bool reset);

} // namespace eb
#endif   // __CUDACC_RTC__
