// convolve.h — K3: partitioned-FFT convolver for the `convolve` node (device side in convolve_kernel.cu).
//
// Reference: ConvolutionNode (wasm/Convolve.h:23-92) wraps fftconvolver::TwoStageFFTConvolver(head 512, tail 4096)
// (wasm/FFTConvolver/TwoStageFFTConvolver.cpp:74-220, FFTConvolver.cpp:85-204).  Whatever its internal staging, that
// object computes — with zero latency and for any call length — the plain linear convolution of the input stream
// with the impulse response (after dropping trailing taps with |h| < 1e-6, FFTConvolver.cpp:94-98).  Here the same
// convolution is evaluated as ONE uniformly partitioned frequency-domain delay line per channel (partition 512,
// FFT 1024, S = ceil(irLen/512) partitions), the structure of the reference's head stage applied to the whole IR:
// no 8192-point tail bursts, identical work every block, spectra streamed once per block.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace eb {

constexpr int CONV_BLOCK = 512;          // partition / head block size (Convolve.h:49: init(512, 4096, ...))
constexpr int CONV_FFT = 1024;           // segment size 2*B (FFTConvolver.cpp:107)
constexpr int CONV_BINS = 513;           // ComplexSize(1024)
constexpr int CONV_PACKED_BINS = 512;    // stored spectra: bin 0 = (Re X[0], Re X[512]) — both bins are purely real
constexpr int CONV_CH_PER_CTA = 2;       // channels sharing one pass over the IR spectra (measured: profiles/r01_g_*)

struct ConvolverState {
    int partitions = 0;                  // S
    int nv = 0;                          // channels (voices of the group)
    int cur = 0;                         // FDL slot of the block being filled (decrements per block: FFTConvolver.cpp:200)
    int fill = 0;                        // samples already in the current partition's input buffer (:157-162)
    float2* dH = nullptr;                // [S][512]   IR partition spectra (packed)
    float2* dFdl = nullptr;              // [nv][S][512] input spectra ring (frequency-domain delay line)
    float2* dYpre = nullptr;             // [nv][512]  sum over the older partitions, valid while fill > 0 (:168-177)
    float* dOverlap = nullptr;           // [nv][512]
    float* dInBuf = nullptr;             // [nv][512]
    float2* dTw = nullptr;               // [512] exp(-2*pi*i*m/1024)
    bool planOnly = false;
    ~ConvolverState();
    size_t bytesPerChannel() const { return (size_t) partitions * CONV_PACKED_BINS * 8 + CONV_PACKED_BINS * 8 + 2 * CONV_BLOCK * 4; }
};

// Build the device state for `nv` channels sharing one impulse response. Returns false and fills `err` on failure.
bool convolver_init(ConvolverState& st, const float* ir, size_t irLen, int nv, bool planOnly, cudaStream_t stream, std::string& err);

// Copy of channels [c0, c0+n) of `src` (same IR, same position in the partition cycle) — used when a voice group is cut.
bool convolver_clone_range(const ConvolverState& src, int c0, int n, ConvolverState& dst, cudaStream_t stream, std::string& err);

// Root + mix fused behind the inverse transform (graph_host.cpp sets it when the convolver's output feeds nothing but the graph's root,
// the `convolve -> root` shape of a reverb send: BASELINE config 4).  The kernel then applies the root's gain fade exactly like K1's
// OP_ROOT does (Core.h:66-78, GainFade.h:56-72: g = clamp(gain0 + step * sampleIndex) while fading, the target otherwise), writes the
// per-voice output and this channel's row of the per-tile partial mix (one-voice tiles: row = channel) — the K1 launch that would only
// have re-read the block and multiplied it disappears, and the convolver's own output buffer is not written at all.
struct ConvEpilogue {
    bool active = false;
    bool running = true;             // the root's sub-sequence runs this block (else: silence, like the skipped segment in K1)
    float gain0 = 1.0f, step = 0.0f, target = 1.0f;
    int channel = 0;                 // the root's output channel
    int nOut = 1, blockSize = 512;
    float* mixPartial = nullptr;     // [tileBase + channel][nOut][blockSize] or null
    int tileBase = 0;
    float* outVoice = nullptr;       // [voice0 + channel][nOut][outStride] (+ outOffset) or null
    int voice0 = 0, outStride = 512, outOffset = 0;
};

// Convolve `n` more samples (n <= 512 - st.fill) of every channel: in is [channel][inStride] (inStride 0 = one input shared by
// all channels), out is [channel][outStride]; the samples of this call start at `offset`.  Advances st.fill / st.cur.
cudaError_t convolver_process_chunk(ConvolverState& st, const float* in, int inStride, float* out, int outStride, int offset, int n, cudaStream_t stream,
                                    const ConvEpilogue& epi = ConvEpilogue{});

// Algorithmic HBM bytes of one full 512-sample block for one channel.
inline size_t convolver_algorithmic_bytes_per_channel_block(int partitions) {
    return (size_t) 2 * CONV_BLOCK * 4                         // input read + output write
         + (size_t) (partitions - 1) * CONV_PACKED_BINS * 8    // older input spectra read
         + (size_t) CONV_PACKED_BINS * 8                       // newest spectrum written
         + (size_t) 2 * CONV_BLOCK * 4;                        // overlap read + write
}

} // namespace eb
