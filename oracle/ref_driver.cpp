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
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>

#include <elem/Runtime.h>
#include <elem/JSON.h>
#include <elem/AudioBufferResource.h>
#include <Convolve.h>
#include <FFT.h>
#include <Metro.h>
#include <SampleTime.h>

namespace {

// Two node types registered through the REFERENCE's own plug-in interface (Runtime::registerNodeType, GraphNode.h:20-96) — test
// fixtures for the device-side registration of the CUDA path (elem_b200_register_node_type): the same two types are registered there
// as CUDA text and must render the same samples.
//   "b200test.softclip": y = x / (1 + |x|)                      (stateless, 1 input)
//   "b200test.leaky":    s = x + g * s; y = s   (g = 2nd input)   (one float of state, 2 inputs)
template <typename F>
struct SoftClipNode : public elem::GraphNode<F> {
    using elem::GraphNode<F>::GraphNode;
    void process(elem::BlockContext<F> const& ctx) override {
        if (ctx.numInputChannels < 1) { std::fill_n(ctx.outputData[0], ctx.numSamples, F(0)); return; }
        for (size_t i = 0; i < ctx.numSamples; ++i) { const F x = ctx.inputData[0][i]; ctx.outputData[0][i] = x / (F(1) + std::fabs(x)); }
    }
};
template <typename F>
struct LeakyNode : public elem::GraphNode<F> {
    using elem::GraphNode<F>::GraphNode;
    F s = 0;
    void process(elem::BlockContext<F> const& ctx) override {
        if (ctx.numInputChannels < 2) { std::fill_n(ctx.outputData[0], ctx.numSamples, F(0)); return; }
        for (size_t i = 0; i < ctx.numSamples; ++i) { s = ctx.inputData[0][i] + ctx.inputData[1][i] * s; ctx.outputData[0][i] = s; }
    }
};

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
        rt.registerNodeType("b200test.softclip", [](elem::NodeId const id, double fs, int const bs) { return std::make_shared<SoftClipNode<float>>(id, fs, bs); });
        rt.registerNodeType("b200test.leaky", [](elem::NodeId const id, double fs, int const bs) { return std::make_shared<LeakyNode<float>>(id, fs, bs); });
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
// instances (the reference's own timing loop, cli/Benchmark.cpp:86-111, is one instance on one thread).
//
// Timing discipline (VERDICT r01 weak #4 — the first version created its threads inside the timed region and timed 0.1 s):
//   * the worker threads are created ONCE, pinned to one CPU each (thread t -> the t-th CPU of the process's affinity mask), and
//     each constructs and warms up ITS OWN instances (first-touch: a voice's memory lives on the node of the core that runs it);
//   * every round (warm-up, timed) starts and ends at a barrier; each worker stamps steady_clock right after the start barrier and
//     right after its last block, elapsed = latest end - earliest start, so neither thread creation nor a straggler's start-up
//     is inside — and the slowest thread is;
//   * the timed round renders at least `blocks` blocks and keeps going until `minSeconds` have passed (every thread stops at a
//     block boundary of its own voices).
// Returns the seconds of the timed round; *stepsOut = voice-blocks rendered / numVoices (how many blocks of ALL voices it amounts
// to); *checksum receives the sum of all output samples of the last block so the work cannot be elided.
namespace {
struct Barrier {
    std::mutex m; std::condition_variable cv; int count, waiting = 0; unsigned gen = 0;
    explicit Barrier(int n) : count(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const unsigned g = gen;
        if (++waiting == count) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
}

double elem_ref_bench_stable(double sampleRate, int blockSize, int numVoices, int threads,
                             const char* baseJson, const char* const* voiceJson,
                             const char* resName, const float* resData, size_t resLen,
                             const float* in, size_t nIn, size_t nOut, size_t numSamples,
                             int warmupBlocks, int blocks, double minSeconds, double* stepsOut, double* checksum) {
    if (threads < 1) threads = 1;
    if (threads > numVoices) threads = numVoices;
    std::vector<std::unique_ptr<RefRuntime>> rts(numVoices);
    std::vector<float> zeros(nIn * numSamples, 0.0f);
    const float* inBase = in ? in : zeros.data();
    std::vector<double> sums(threads, 0.0);
    std::atomic<int> failed{0};
    std::atomic<int> roundBlocks{0};
    std::atomic<double> roundSeconds{0.0};
    std::vector<long> voiceBlocks(threads, 0);
    std::atomic<bool> quit{false};
    Barrier bar(threads + 1);
    std::vector<std::chrono::steady_clock::time_point> tStart(threads), tEnd(threads);

    std::vector<int> cpus;
    {
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0)
            for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &set)) cpus.push_back(c);
    }

    auto worker = [&](int t) {
        if (!cpus.empty()) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[t % cpus.size()], &one);
            pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
        }
        for (int v = t; v < numVoices; v += threads) {
            auto r = std::make_unique<RefRuntime>(sampleRate, blockSize);
            if (resName && resData)
                r->rt.addSharedResource(std::string(resName), std::make_unique<elem::AudioBufferResource>(const_cast<float*>(resData), resLen));
            if (applyJson(r.get(), baseJson) != 0) failed = 1;
            if (voiceJson && voiceJson[v] && applyJson(r.get(), voiceJson[v]) != 0) failed = 1;
            rts[v] = std::move(r);
        }
        std::vector<float> outBuf(nOut * numSamples);
        std::vector<const float*> ip(nIn);
        std::vector<float*> op(nOut);
        for (size_t i = 0; i < nIn; ++i) ip[i] = inBase + i * numSamples;
        for (size_t i = 0; i < nOut; ++i) op[i] = outBuf.data() + i * numSamples;
        for (;;) {
            bar.wait();                                   // round start
            if (quit.load()) return;
            tStart[t] = std::chrono::steady_clock::now();
            const int nb = roundBlocks.load();
            const double limit = roundSeconds.load();     // > 0: keep rendering whole blocks of this thread's voices until the time is up
            double s = 0.0;
            long done = 0;
            for (int b = 0; limit > 0.0 || b < nb; ++b) {
                for (int v = t; v < numVoices; v += threads) {
                    rts[v]->rt.process(ip.data(), nIn, op.data(), nOut, numSamples, static_cast<void*>(&rts[v]->sampleTime));
                    rts[v]->sampleTime += static_cast<int64_t>(numSamples);
                    ++done;
                }
                if (limit > 0.0 && b + 1 >= nb && std::chrono::duration<double>(std::chrono::steady_clock::now() - tStart[t]).count() >= limit) break;
            }
            for (float x : outBuf) s += x;
            sums[t] = s;
            voiceBlocks[t] = done;
            tEnd[t] = std::chrono::steady_clock::now();
            bar.wait();                                   // round end
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back(worker, t);

    // elapsed = last worker's end - first worker's start, stamped by the workers themselves: the main thread is not pinned and,
    // with one busy worker per CPU, may not be scheduled again until the round is over
    auto round = [&](int nb) -> double {
        roundBlocks = nb;
        bar.wait();
        bar.wait();
        auto t0 = tStart[0], t1 = tEnd[0];
        for (int t = 1; t < threads; ++t) { t0 = std::min(t0, tStart[t]); t1 = std::max(t1, tEnd[t]); }
        return std::chrono::duration<double>(t1 - t0).count();
    };
    round(std::max(1, warmupBlocks));                     // includes construction: never timed
    // Timed round: at least `blocks` blocks and — when minSeconds > 0 — at least that long.  Time-bounded rather than a block count
    // from a calibration: on a box whose container has a CPU quota the first milliseconds run unthrottled and a calibration
    // underestimates the steady state by the throttling factor (seen: 7x).
    roundSeconds = minSeconds;
    const double secs = round(blocks);
    long totalVoiceBlocks = 0;
    for (long n : voiceBlocks) totalVoiceBlocks += n;
    const double steps = (double) totalVoiceBlocks / (double) numVoices;      // blocks of ALL voices the timed round amounts to
    quit = true;
    bar.wait();
    for (auto& x : th) x.join();

    double total = 0.0;
    for (double s : sums) total += s;
    if (checksum) *checksum = total;
    if (stepsOut) *stepsOut = steps;
    return failed.load() ? -1.0 : secs;
}

double elem_ref_bench(double sampleRate, int blockSize, int numVoices, int threads,
                      const char* baseJson, const char* const* voiceJson,
                      const char* resName, const float* resData, size_t resLen,
                      const float* in, size_t nIn, size_t nOut, size_t numSamples,
                      int warmupBlocks, int blocks, double* checksum) {
    return elem_ref_bench_stable(sampleRate, blockSize, numVoices, threads, baseJson, voiceJson, resName, resData, resLen,
                                 in, nIn, nOut, numSamples, warmupBlocks, blocks, 0.0, nullptr, checksum);
}

// CPU model string of the box (first "model name" of /proc/cpuinfo), for the bench line.
int elem_ref_cpu_model(char* buf, size_t cap) {
    std::string model = "unknown";
    if (FILE* f = std::fopen("/proc/cpuinfo", "r")) {
        char line[512];
        while (std::fgets(line, sizeof(line), f))
            if (!std::strncmp(line, "model name", 10)) {
                const char* c = std::strchr(line, ':');
                if (c) { model = c + 1; while (!model.empty() && (model.back() == '\n' || model.back() == ' ')) model.pop_back(); while (!model.empty() && model.front() == ' ') model.erase(model.begin()); }
                break;
            }
        std::fclose(f);
    }
    if (buf && cap) { std::strncpy(buf, model.c_str(), cap - 1); buf[cap - 1] = 0; }
    return (int) model.size();
}

const char* elem_ref_describe() {
    return "elemaudio/elementary reference engine, Runtime<float> + ConvolutionNode, built from /root/reference in place";
}

} // extern "C"
