// convolve_host.cpp — host side of K3: impulse-response partitioning and device state of the convolver.
#include "convolve.h"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>

namespace eb {

namespace {

void freeDev(void* p, bool plan) { if (!p) return; if (plan) std::free(p); else cudaFree(p); }

bool allocDev(void** p, size_t bytes, bool plan, cudaStream_t stream, std::string& err) {
    if (plan) { *p = std::calloc(1, bytes ? bytes : 1); if (!*p) { err = "out of host memory"; return false; } return true; }
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { err = std::string("cudaMalloc convolver: ") + cudaGetErrorString(e); return false; }
    e = cudaMemsetAsync(*p, 0, bytes, stream);
    if (e != cudaSuccess) { err = std::string("cudaMemset convolver: ") + cudaGetErrorString(e); return false; }
    return true;
}

// In-place iterative radix-2 FFT in double (host, init time only): spectra of the IR partitions are computed in
// double like the reference's OouraFFT (AudioFFT.cpp:132-155) and rounded to float once.
void fftDouble(std::vector<std::complex<double>>& a) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double) len;
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                const std::complex<double> w(std::cos(ang * (double) k), std::sin(ang * (double) k));
                const std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

} // namespace

ConvolverState::~ConvolverState() {
    freeDev(dH, planOnly); freeDev(dFdl, planOnly); freeDev(dYpre, planOnly);
    freeDev(dOverlap, planOnly); freeDev(dInBuf, planOnly); freeDev(dTw, planOnly);
}

bool convolver_init(ConvolverState& st, const float* ir, size_t irLen, int nv, bool planOnly, cudaStream_t stream, std::string& err) {
    st.planOnly = planOnly;
    st.nv = nv;
    st.cur = 0;
    st.fill = 0;
    // FFTConvolver.cpp:94-98 / TwoStageFFTConvolver.cpp:93-97: trailing taps below 1e-6 are dropped
    while (irLen > 0 && std::fabs(ir[irLen - 1]) < 0.000001f) --irLen;
    st.partitions = (int) ((irLen + CONV_BLOCK - 1) / CONV_BLOCK);
    if (st.partitions == 0) return true;
    const int S = st.partitions;

    std::vector<float2> H((size_t) S * CONV_PACKED_BINS);
    std::vector<std::complex<double>> buf(CONV_FFT);
    for (int i = 0; i < S; ++i) {
        for (int k = 0; k < CONV_FFT; ++k) {
            const size_t idx = (size_t) i * CONV_BLOCK + k;
            buf[k] = (k < CONV_BLOCK && idx < irLen) ? std::complex<double>((double) ir[idx], 0.0) : std::complex<double>(0.0, 0.0);
        }
        fftDouble(buf);
        for (int b = 1; b < CONV_PACKED_BINS; ++b) H[(size_t) i * CONV_PACKED_BINS + b] = make_float2((float) buf[b].real(), (float) buf[b].imag());
        H[(size_t) i * CONV_PACKED_BINS] = make_float2((float) buf[0].real(), (float) buf[CONV_BLOCK].real());   // packed (H[0], H[512])
    }
    std::vector<float2> tw(512);
    for (int m = 0; m < 512; ++m) {
        const double a = -2.0 * M_PI * (double) m / 1024.0;
        tw[m] = make_float2((float) std::cos(a), (float) std::sin(a));
    }

    if (!allocDev((void**) &st.dH, H.size() * sizeof(float2), planOnly, stream, err)) return false;
    if (!allocDev((void**) &st.dTw, tw.size() * sizeof(float2), planOnly, stream, err)) return false;
    if (!allocDev((void**) &st.dFdl, (size_t) nv * S * CONV_PACKED_BINS * sizeof(float2), planOnly, stream, err)) return false;
    if (!allocDev((void**) &st.dYpre, (size_t) nv * CONV_PACKED_BINS * sizeof(float2), planOnly, stream, err)) return false;
    if (!allocDev((void**) &st.dOverlap, (size_t) nv * CONV_BLOCK * sizeof(float), planOnly, stream, err)) return false;
    if (!allocDev((void**) &st.dInBuf, (size_t) nv * CONV_BLOCK * sizeof(float), planOnly, stream, err)) return false;
    if (planOnly) {
        std::memcpy(st.dH, H.data(), H.size() * sizeof(float2));
        std::memcpy(st.dTw, tw.data(), tw.size() * sizeof(float2));
        return true;
    }
    cudaError_t e = cudaMemcpyAsync(st.dH, H.data(), H.size() * sizeof(float2), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(st.dTw, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);   // H / tw are pageable temporaries
    if (e != cudaSuccess) { err = std::string("convolver upload: ") + cudaGetErrorString(e); return false; }
    return true;
}

bool convolver_clone_range(const ConvolverState& src, int c0, int n, ConvolverState& dst, cudaStream_t stream, std::string& err) {
    dst.planOnly = src.planOnly;
    dst.nv = n; dst.partitions = src.partitions; dst.cur = src.cur; dst.fill = src.fill;
    if (src.partitions == 0) return true;
    const size_t S = (size_t) src.partitions;
    struct Part { void** d; const void* s; size_t bytes; };
    const Part parts[] = {
        {(void**) &dst.dH, src.dH, S * CONV_PACKED_BINS * sizeof(float2)},
        {(void**) &dst.dTw, src.dTw, 512 * sizeof(float2)},
        {(void**) &dst.dFdl, src.dFdl + (size_t) c0 * S * CONV_PACKED_BINS, (size_t) n * S * CONV_PACKED_BINS * sizeof(float2)},
        {(void**) &dst.dYpre, src.dYpre + (size_t) c0 * CONV_PACKED_BINS, (size_t) n * CONV_PACKED_BINS * sizeof(float2)},
        {(void**) &dst.dOverlap, src.dOverlap + (size_t) c0 * CONV_BLOCK, (size_t) n * CONV_BLOCK * sizeof(float)},
        {(void**) &dst.dInBuf, src.dInBuf + (size_t) c0 * CONV_BLOCK, (size_t) n * CONV_BLOCK * sizeof(float)},
    };
    for (const Part& p : parts) {
        if (!allocDev(p.d, p.bytes, src.planOnly, stream, err)) return false;
        if (src.planOnly) std::memcpy(*p.d, p.s, p.bytes);
        else if (cudaMemcpyAsync(*p.d, p.s, p.bytes, cudaMemcpyDeviceToDevice, stream) != cudaSuccess) { err = "convolver clone copy failed"; return false; }
    }
    return true;
}

} // namespace eb
