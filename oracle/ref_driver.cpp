// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" driver around the UNMODIFIED reference engine. It is compiled
// against the reference headers/sources where they lie under /root/reference (see
// oracle/Makefile) into oracle/_ref/libelem_ref.so. Nothing from the reference is
// copied into this repository: this file only *calls* the reference's public API
//   elem::Runtime<float>  (runtime/elem/Runtime.h:40-110)
//   elem::js::parseJSON   (runtime/elem/JSON.h:17)
//   elem::ConvolutionNode (wasm/Convolve.h:23-92, registered as in wasm/Main.cpp:47-49)
// so that the tests can compare the CUDA path and the CPU restatement
// (oracle/elem_oracle.cpp) against the real thing, and so that bench.py can time the
// reference's own CPU path on the GPU box's host cores (the .so travels with gpurun).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include <elem/Runtime.h>
#include <elem/JSON.h>
#include <elem/AudioBufferResource.h>
#include <Convolve.h>
#include <FFT.h>
#include <Metro.h>
#include <SampleTime.h>

namespace {

struct RefRuntime {
    elem::Runtime<float> rt;
    int64_t sampleTime = 0;   // the host-kept clock handed to nodes as userData (wasm/Main.cpp:206-217)
    RefRuntime(double sr, int bs) : rt(sr, bs) {
        // Same registrations the wasm host performs (wasm/Main.cpp:47-61).
        rt.registerNodeType("convolve", [](elem::NodeId const id, double fs, int const bs) {
            return std::make_shared<elem::ConvolutionNode<float>>(id, fs, bs);
        });
        rt.registerNodeType("fft", [](elem::NodeId const id, double fs, int const bs) {
            return std::make_shared<elem::FFTNode<float>>(id, fs, bs);
        });
        rt.registerNodeType("metro", [](elem::NodeId const id, double fs, int const bs) {
            return std::make_shared<elem::MetronomeNode<float>>(id, fs, bs);
        });
        rt.registerNodeType("time", [](elem::NodeId const id, double fs, int const bs) {
            return std::make_shared<elem::SampleTimeNode<float>>(id, fs, bs);
        });
    }
};

int applyJson(RefRuntime* r, const char* json) {
    try {
        auto v = elem::js::parseJSON(std::string(json));
        if (!v.isArray()) return elem::ReturnCode::InvalidInstructionFormat();
        return r->rt.applyInstructions(v.getArray());
    } catch (...) {
        // Mirrors what a C-ABI shim has to do with the reference's exceptions
        // (JSON.h:146-154, Value.h:89-92).
        return elem::ReturnCode::InvalidInstructionFormat();
    }
}

} // namespace

extern "C" {

void* elem_ref_create(double sampleRate, int blockSize) {
    return new RefRuntime(sampleRate, blockSize);
}

void elem_ref_destroy(void* h) { delete static_cast<RefRuntime*>(h); }

int elem_ref_apply_instructions(void* h, const char* json) {
    return applyJson(static_cast<RefRuntime*>(h), json);
}

int elem_ref_add_shared_resource(void* h, const char* name, const float* data, size_t numSamples) {
    auto* r = static_cast<RefRuntime*>(h);
    auto res = std::make_unique<elem::AudioBufferResource>(const_cast<float*>(data), numSamples);
    return r->rt.addSharedResource(std::string(name), std::move(res)) ? 1 : 0;
}

// Planar I/O exactly like Runtime::process (Runtime.h:51-57): in = nIn pointers, out = nOut pointers.
void elem_ref_process(void* h, const float* const* in, size_t nIn, float* const* out, size_t nOut, size_t numSamples) {
    auto* r = static_cast<RefRuntime*>(h);
    r->rt.process(const_cast<const float**>(in), nIn, const_cast<float**>(out), nOut, numSamples, static_cast<void*>(&r->sampleTime));
    r->sampleTime += static_cast<int64_t>(numSamples);   // wasm/Main.cpp:217
}

void elem_ref_set_current_time(void* h, int64_t t) { static_cast<RefRuntime*>(h)->sampleTime = t; }

// Runtime::processQueuedEvents (Runtime.h:64,438-446) relayed the way wasm/Main.cpp:220-231 does: a JSON array of
// {"type","event"} objects, serialised with the reference's own js::serialize.  Returns the bytes needed.
int elem_ref_process_queued_events(void* h, char* buf, size_t cap) {
    auto* r = static_cast<RefRuntime*>(h);
    elem::js::Array batch;
    r->rt.processQueuedEvents([&batch](std::string const& type, elem::js::Value evt) {
        batch.push_back(elem::js::Object({{"type", type}, {"event", evt}}));
    });
    std::string s = elem::js::serialize(elem::js::Value(batch));
    if (buf && cap) {
        const size_t k = s.size() < cap - 1 ? s.size() : cap - 1;
        std::memcpy(buf, s.data(), k);
        buf[k] = 0;
    }
    return (int) s.size();
}

// Convenience for tests: contiguous buffers in[nIn][numSamples], out[nOut][numSamples].
void elem_ref_process_flat(void* h, const float* in, size_t nIn, float* out, size_t nOut, size_t numSamples) {
    std::vector<const float*> ip(nIn);
    std::vector<float*> op(nOut);
    for (size_t i = 0; i < nIn; ++i) ip[i] = in + i * numSamples;
    for (size_t i = 0; i < nOut; ++i) op[i] = out + i * numSamples;
    elem_ref_process(h, ip.data(), nIn, op.data(), nOut, numSamples);
}

int elem_ref_gc(void* h, int32_t* ids, size_t cap) {
    auto* r = static_cast<RefRuntime*>(h);
    auto pruned = r->rt.gc();
    size_t n = 0;
    for (auto id : pruned) { if (n < cap) ids[n] = id; ++n; }
    return (int) n;
}

void elem_ref_reset(void* h) { static_cast<RefRuntime*>(h)->rt.reset(); }

// ---------------------------------------------------------------------------------------
// Multi-instance CPU baseline (SURVEY.md §8d "How the reference CPU path is timed"):
// V independent Runtime<float> instances, each fed `baseJson` and then its own
// `voiceJson[v]` (may be NULL), `threads` host threads each round-robin over its share of
// instances for `blocks` blocks of `numSamples`; wall time (steady_clock) of the steady
// state loop only. Inputs: nIn channels of zeros (or `in` if given, shared by all voices).
// Returns seconds; writes the sum of all output samples of the last block to *checksum so
// the work cannot be elided.
double elem_ref_bench(double sampleRate, int blockSize, int numVoices, int threads,
                      const char* baseJson, const char* const* voiceJson,
                      const char* resName, const float* resData, size_t resLen,
                      const float* in, size_t nIn, size_t nOut, size_t numSamples,
                      int warmupBlocks, int blocks, double* checksum) {
    std::vector<std::unique_ptr<RefRuntime>> rts;
    rts.reserve(numVoices);
    for (int v = 0; v < numVoices; ++v) {
        auto r = std::make_unique<RefRuntime>(sampleRate, blockSize);
        if (resName && resData) {
            r->rt.addSharedResource(std::string(resName),
                std::make_unique<elem::AudioBufferResource>(const_cast<float*>(resData), resLen));
        }
        if (applyJson(r.get(), baseJson) != 0) return -1.0;
        if (voiceJson && voiceJson[v] && applyJson(r.get(), voiceJson[v]) != 0) return -1.0;
        rts.push_back(std::move(r));
    }

    if (threads < 1) threads = 1;
    if (threads > numVoices) threads = numVoices;

    std::vector<float> zeros(nIn * numSamples, 0.0f);
    const float* inBase = in ? in : zeros.data();
    std::vector<double> sums(threads, 0.0);

    auto worker = [&](int t, int nblocks, bool record) {
        std::vector<float> outBuf(nOut * numSamples);
        std::vector<const float*> ip(nIn);
        std::vector<float*> op(nOut);
        for (size_t i = 0; i < nIn; ++i) ip[i] = inBase + i * numSamples;
        for (size_t i = 0; i < nOut; ++i) op[i] = outBuf.data() + i * numSamples;
        double s = 0.0;
        for (int b = 0; b < nblocks; ++b) {
            for (int v = t; v < numVoices; v += threads) {
                rts[v]->rt.process(ip.data(), nIn, op.data(), nOut, numSamples, static_cast<void*>(&rts[v]->sampleTime));
                rts[v]->sampleTime += static_cast<int64_t>(numSamples);
                if (record && b == nblocks - 1)
                    for (float x : outBuf) s += x;
            }
        }
        if (record) sums[t] = s;
    };

    auto runAll = [&](int nblocks, bool record) {
        std::vector<std::thread> th;
        for (int t = 1; t < threads; ++t) th.emplace_back(worker, t, nblocks, record);
        worker(0, nblocks, record);
        for (auto& x : th) x.join();
    };

    if (warmupBlocks > 0) runAll(warmupBlocks, false);
    auto t0 = std::chrono::steady_clock::now();
    runAll(blocks, true);
    auto t1 = std::chrono::steady_clock::now();

    double total = 0.0;
    for (double s : sums) total += s;
    if (checksum) *checksum = total;
    return std::chrono::duration<double>(t1 - t0).count();
}

const char* elem_ref_describe() {
    return "elemaudio/elementary reference engine, Runtime<float> + ConvolutionNode, built from /root/reference in place";
}

} // extern "C"
