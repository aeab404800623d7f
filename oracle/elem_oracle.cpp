// oracle/elem_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain scalar CPU restatement of the reference algorithm for the hot path
//   elem::Runtime<float>::applyInstructions / process  and the builtin node kernels it walks,
// written from the reference's semantics (each function cites the file:line under /root/reference it follows).
// It is deliberately structured like the reference — one block buffer per node, nodes processed one after
// another over the whole block — and shares NO code with the CUDA product (own instruction reader, own graph
// walk, own node kernels), so that a bug in the product cannot hide in a common dependency.
//
// Pinning: this restatement is validated (tests/test_oracle_cpu.py) against
//   (1) the compiled reference itself (oracle/_ref/libelem_ref.so) — bit-exact on every graph of the test-suite
//       (same compiler, flags -O2 -ffp-contract=off and glibc libm), and
//   (2) the reference's own golden vectors (jest snapshots for delay/sdelay/table/taps/maxhold/const math and
//       the known-answer anchors of SURVEY.md Appendix E).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
//
// Build: g++ -std=c++17 -O2 -DNDEBUG -ffp-contract=off (oracle/Makefile) — no -ffast-math, no FMA contraction.

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

constexpr float kEps = FLT_EPSILON;

// ---- helpers/Change.h:12-32 ------------------------------------------------------------------------------
struct Change {
    float lastIn = 0;
    float operator()(float xn) {
        float dt = xn - lastIn;
        lastIn = xn;
        if (dt > 0.0f) return 1.0f;
        if (dt < 0.0f) return -1.0f;
        return 0.0f;
    }
};

// ---- helpers/GainFade.h:10-121 -----------------------------------------------------------------------------
struct Fade {
    float cur = 0, target = 0, step = 0, inStep = 0, outStep = 0;
    static double msToStep(double sr, double ms) { return ms > 1e-6 ? 1.0 / (sr * ms / 1000.0) : 1.0; }
    void update() { step = cur > target ? outStep : inStep; }
    void init(double sr) { cur = 0.0f; target = 1.0f; setIn(sr, 20); setOut(sr, 20); }   // Core.h:80
    void setIn(double sr, double ms) { inStep = (float) msToStep(sr, ms); update(); }
    void setOut(double sr, double ms) { outStep = (float) ((double) -1.0f * msToStep(sr, ms)); update(); }
    void fadeIn() { target = 1.0f; update(); }
    void fadeOut() { target = 0.0f; update(); }
    bool on() const { return target > 0.5f; }
    bool settled() const { return std::abs(target - cur) <= 1e-6f; }
    void process(const float* in, float* out, int n) {   // GainFade.h:56-72
        if (cur == target) {
            for (int i = 0; i < n; ++i) out[i] = in[i] * target;
            return;
        }
        for (int i = 0; i < n; ++i) out[i] = in[i] * std::clamp(cur + step * i, 0.0f, 1.0f);
        cur = std::clamp(cur + step * n, 0.0f, 1.0f);
    }
};


// ---- wasm/FFTConvolver restated ------------------------------------------------------------------------------
// Real FFT with float I/O computed in double (AudioFFT.cpp:132-176 does the same with Ooura's rdft; any exact
// double-precision DFT gives the same values to ~1e-16, i.e. identical floats except on rounding ties).
struct RealFFT {
    size_t n = 0;
    std::vector<double> cosT, sinT;
    void init(size_t size) {
        n = size;
        cosT.resize(n / 2); sinT.resize(n / 2);
        for (size_t k = 0; k < n / 2; ++k) { cosT[k] = std::cos(2.0 * M_PI * (double) k / (double) n); sinT[k] = std::sin(2.0 * M_PI * (double) k / (double) n); }
    }
    void transform(std::vector<double>& re, std::vector<double>& im, bool inverse) const {
        for (size_t i = 1, j = 0; i < n; ++i) {
            size_t bit = n >> 1;
            for (; j & bit; bit >>= 1) j ^= bit;
            j ^= bit;
            if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
        }
        for (size_t len = 2; len <= n; len <<= 1) {
            const size_t step = n / len;
            for (size_t i = 0; i < n; i += len)
                for (size_t k = 0; k < len / 2; ++k) {
                    const double wr = cosT[k * step], wi = inverse ? sinT[k * step] : -sinT[k * step];
                    const double xr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                    const double xi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                    re[i + k + len / 2] = re[i + k] - xr; im[i + k + len / 2] = im[i + k] - xi;
                    re[i + k] += xr; im[i + k] += xi;
                }
        }
    }
    void fft(const float* data, float* ore, float* oim) const {       // AudioFFT::fft: n reals -> n/2+1 bins
        std::vector<double> re(data, data + n), im(n, 0.0);
        transform(re, im, false);
        for (size_t k = 0; k <= n / 2; ++k) { ore[k] = (float) re[k]; oim[k] = (float) im[k]; }
    }
    void ifft(float* data, const float* ire, const float* iim) const { // AudioFFT::ifft: n/2+1 bins -> n reals (normalised)
        std::vector<double> re(n), im(n);
        for (size_t k = 0; k <= n / 2; ++k) { re[k] = ire[k]; im[k] = iim[k]; }
        for (size_t k = n / 2 + 1; k < n; ++k) { re[k] = ire[n - k]; im[k] = -(double) iim[n - k]; }
        transform(re, im, true);
        for (size_t i = 0; i < n; ++i) data[i] = (float) (re[i] / (double) n);
    }
};

struct Spectrum { std::vector<float> re, im; void resize(size_t n) { re.assign(n, 0.0f); im.assign(n, 0.0f); } };

// Utilities.cpp:66-117 (scalar branch)
static void cmac(Spectrum& r, const Spectrum& a, const Spectrum& b) {
    for (size_t i = 0; i < r.re.size(); ++i) {
        r.re[i] += a.re[i] * b.re[i] - a.im[i] * b.im[i];
        r.im[i] += a.re[i] * b.im[i] + a.im[i] * b.re[i];
    }
}

static size_t nextPow2(size_t v) { size_t p = 1; while (p < v) p *= 2; return p; }   // Utilities.h:288-296

// FFTConvolver.cpp:85-204 — uniformly partitioned overlap-add convolution with zero latency
struct FFTConv {
    size_t blockSize = 0, segSize = 0, segCount = 0, bins = 0, current = 0, fill = 0;
    std::vector<Spectrum> segs, segsIR;
    std::vector<float> fftBuffer, overlap, inputBuffer;
    Spectrum pre, conv;
    RealFFT fft;

    void init(size_t bs, const float* ir, size_t irLen) {                          // :85-144
        *this = FFTConv();
        if (bs == 0) return;
        while (irLen > 0 && std::fabs(ir[irLen - 1]) < 0.000001f) --irLen;         // :94-98
        if (irLen == 0) return;
        blockSize = nextPow2(bs);
        segSize = 2 * blockSize;
        segCount = (size_t) std::ceil((float) irLen / (float) blockSize);
        bins = segSize / 2 + 1;
        fft.init(segSize);
        fftBuffer.assign(segSize, 0.0f);
        segs.resize(segCount); segsIR.resize(segCount);
        for (size_t i = 0; i < segCount; ++i) {
            segs[i].resize(bins); segsIR[i].resize(bins);
            const size_t remaining = irLen - i * blockSize;
            const size_t sizeCopy = remaining >= blockSize ? blockSize : remaining;
            std::fill(fftBuffer.begin(), fftBuffer.end(), 0.0f);                   // CopyAndPad
            std::copy_n(ir + i * blockSize, sizeCopy, fftBuffer.begin());
            fft.fft(fftBuffer.data(), segsIR[i].re.data(), segsIR[i].im.data());
        }
        pre.resize(bins); conv.resize(bins);
        overlap.assign(blockSize, 0.0f);
        inputBuffer.assign(blockSize, 0.0f);
    }

    void process(const float* input, float* output, size_t len) {                  // :147-204
        if (segCount == 0) { std::fill_n(output, len, 0.0f); return; }
        size_t processed = 0;
        while (processed < len) {
            const bool wasEmpty = fill == 0;
            const size_t processing = std::min(len - processed, blockSize - fill);
            const size_t pos = fill;
            std::copy_n(input + processed, processing, inputBuffer.begin() + pos);
            std::fill(fftBuffer.begin(), fftBuffer.end(), 0.0f);
            std::copy_n(inputBuffer.begin(), blockSize, fftBuffer.begin());
            fft.fft(fftBuffer.data(), segs[current].re.data(), segs[current].im.data());
            if (wasEmpty) {
                pre.resize(bins);
                for (size_t i = 1; i < segCount; ++i) cmac(pre, segsIR[i], segs[(current + i) % segCount]);
            }
            conv = pre;
            cmac(conv, segs[current], segsIR[0]);
            fft.ifft(fftBuffer.data(), conv.re.data(), conv.im.data());
            for (size_t i = 0; i < processing; ++i) output[processed + i] = fftBuffer[pos + i] + overlap[pos + i];   // Sum
            fill += processing;
            if (fill == blockSize) {
                std::fill(inputBuffer.begin(), inputBuffer.end(), 0.0f);
                fill = 0;
                std::copy_n(fftBuffer.begin() + blockSize, blockSize, overlap.begin());
                current = current > 0 ? current - 1 : segCount - 1;
            }
            processed += processing;
        }
    }
};

// TwoStageFFTConvolver.cpp:74-237 — head (short blocks) + two tail stages whose results are delayed by one / two
// tail blocks; the tail convolution runs inline (doBackgroundProcessing, :234-237)
struct TwoStageConv {
    size_t headBlock = 0, tailBlock = 0, tailInputFill = 0, precalculatedPos = 0;
    FFTConv head, tail0, tail;
    std::vector<float> tailOutput0, tailPre0, tailOutput, tailPre, tailInput, bgInput;

    void init(size_t hb, size_t tb, const float* ir, size_t irLen) {               // :74-135
        *this = TwoStageConv();
        if (hb == 0 || tb == 0) return;
        if (hb > tb) std::swap(hb, tb);
        while (irLen > 0 && std::fabs(ir[irLen - 1]) < 0.000001f) --irLen;
        if (irLen == 0) return;
        headBlock = nextPow2(hb); tailBlock = nextPow2(tb);
        head.init(headBlock, ir, std::min(irLen, tailBlock));
        if (irLen > tailBlock) {
            tail0.init(headBlock, ir + tailBlock, std::min(irLen - tailBlock, tailBlock));
            tailOutput0.assign(tailBlock, 0.0f); tailPre0.assign(tailBlock, 0.0f);
        }
        if (irLen > 2 * tailBlock) {
            tail.init(tailBlock, ir + 2 * tailBlock, irLen - 2 * tailBlock);
            tailOutput.assign(tailBlock, 0.0f); tailPre.assign(tailBlock, 0.0f); bgInput.assign(tailBlock, 0.0f);
        }
        if (!tailPre0.empty() || !tailPre.empty()) tailInput.assign(tailBlock, 0.0f);
    }

    void process(const float* input, float* output, size_t len) {                  // :138-220
        head.process(input, output, len);
        if (tailInput.empty()) return;
        size_t processed = 0;
        while (processed < len) {
            const size_t processing = std::min(len - processed, headBlock - (tailInputFill % headBlock));
            if (!tailPre0.empty()) for (size_t i = 0; i < processing; ++i) output[processed + i] += tailPre0[precalculatedPos + i];
            if (!tailPre.empty()) for (size_t i = 0; i < processing; ++i) output[processed + i] += tailPre[precalculatedPos + i];
            precalculatedPos += processing;
            std::copy_n(input + processed, processing, tailInput.begin() + tailInputFill);
            tailInputFill += processing;
            if (!tailPre0.empty() && tailInputFill % headBlock == 0) {
                const size_t off = tailInputFill - headBlock;
                tail0.process(tailInput.data() + off, tailOutput0.data() + off, headBlock);
                if (tailInputFill == tailBlock) tailPre0.swap(tailOutput0);
            }
            if (!tailPre.empty() && tailInputFill == tailBlock) {
                tailPre.swap(tailOutput);
                bgInput = tailInput;
                tail.process(bgInput.data(), tailOutput.data(), tailBlock);
            }
            if (tailInputFill == tailBlock) { tailInputFill = 0; precalculatedPos = 0; }
            processed += processing;
        }
    }
};

struct PropValue {
    char kind = 'N';   // N number, S string, B bool, A array of numbers, M array of {key: number} objects, U null, J other json
    double num = 0;
    std::string str;
    std::vector<double> arr;
    std::vector<std::map<std::string, double>> objs;
};

struct Resource { std::vector<float> data; };
struct Engine;

struct Node {
    int32_t id = 0;
    std::string type;
    std::vector<std::pair<int32_t, int>> inlets;
    std::vector<float> out;      // one block buffer per node (GraphRenderSequence.h:121-124)
    Engine* eng = nullptr;

    // state / props, by node family
    float phase = 0, acc = 0;                       // phasor, blep
    Change change;
    float value = 1.0f;                             // const (Core.h:166)
    float count = 0, total = 0, z = 0, hold = 0;    // counter, accum, latch
    float mx = 0; uint32_t heldSamples = 0; uint32_t holdTime = 0xFFFFFFFFu;   // maxhold
    uint32_t seed = 0;                              // rand (reference default is std::rand(); tests always set it)
    float z1 = 0, z2 = 0;                           // pole / env (z1), biquad
    double dz = 0, ic1 = 0, ic2 = 0;                // mm1p, svf
    int mode = 0, channel = 0;
    Fade fade;                                      // root
    bool activeProp = false;
    std::vector<float> ring; int writeIndex = 0; int length = 0; bool ringPending = false; std::vector<float> pendingRing;
    std::string tapName;
    std::vector<float> tapPrivate;
    std::shared_ptr<Resource> res, pendingRes;
    std::shared_ptr<TwoStageConv> conv, pendingConv;   // convolve (wasm/Convolve.h)

    // once (Core.h:341-404)
    float armed = 0, gain = 0;
    // seq / seq2 (Core.h:407-573, Seq2.h:35-166)
    std::shared_ptr<std::vector<float>> seqActive, seqPending;
    Change resetChange;
    bool wantsHold = false, wantsLoop = true, firstPulse = false;
    size_t seqOffset = 0, seqIndex = 0, edgeCount2 = 0;
    float holdValue = 0;
    // sparseq (SparSeq.h:17-377)
    std::shared_ptr<std::map<int32_t, float>> spActive;
    struct SpEvent { std::shared_ptr<std::map<int32_t, float>> seq; bool isLoop = false; int32_t ls = -1, le = -1; };
    std::vector<SpEvent> spQueue;
    int32_t loopStart = -1, loopEnd = -1, pendStart = -1, pendEnd = -1; bool hasPending = false;
    bool follow = false; int32_t holdOrder = 0; double tickInterval = 0;
    int32_t edgeCount = -1; size_t samplesSince = 0;
    std::map<int32_t, float>::iterator spHold;
    // sparseq2 (SparSeq2.h:17-141)
    std::shared_ptr<std::map<double, float>> sp2Active, sp2Pending;
    std::map<double, float>::iterator sp2Prev, sp2Next;
    int32_t interpOrder = 0;
    // metro (wasm/Metro.h)
    int64_t intervalSamps = 0;
    float lastOut = 0;
    // analysis nodes (Analyzers.h): latest readout + flag, drained by processEvents
    // SingleWriterSingleReaderQueue(32) of readouts (meter / snapshot): positions + slots, with the reference's
    // full-reads-as-empty arithmetic (SingleWriterSingleReaderQueue.h:31-45,63-83,86-111)
    struct Readout { float a = 0, b = 0; };
    Readout roQueue[32]; size_t roR = 0, roW = 0;
    void roPush(Readout r) { roQueue[roW] = r; roW = (roW + 1) & 31; }       // numFreeSlots() is never 0: the push always lands
    size_t roSize() const { return roW > roR ? roW - roR : ((32 - (roR - roW)) & 31); }
    bool metroFlag = false;
    // scope: MultiChannelRingBuffer(4, 8192) (Analyzers.h:149); capture: (1, bitceil(sr)) + 128-sample scratch (Capture.h:17,96-98)
    std::vector<std::vector<float>> mcRing; size_t mcR = 0, mcW = 0, mcCap = 0;
    void mcInit(size_t ch, size_t cap) { mcRing.assign(ch, std::vector<float>(cap, 0.0f)); mcCap = cap; mcR = mcW = 0; }
    size_t mcFull() const { return mcW > mcR ? mcW - mcR : ((mcCap - (mcR - mcW)) & (mcCap - 1)); }
    size_t mcFree() const { return mcR > mcW ? mcR - mcW : mcCap - (mcW - mcR); }
    void mcWrite(const float* const* data, size_t nch, size_t n) {             // MultiChannelRingBuffer.h:36-62
        const bool move = n >= mcFree();
        for (size_t c = 0; c < std::min(mcRing.size(), nch); ++c) for (size_t i = 0; i < n; ++i) mcRing[c][(mcW + i) & (mcCap - 1)] = data[c][i];
        mcW = (mcW + n) & (mcCap - 1);
        if (move) mcR = (mcW + 1) & (mcCap - 1);
    }
    bool mcRead(float* const* dst, size_t nch, size_t n) {                     // :64-88
        if (mcFull() < n) return false;
        for (size_t c = 0; c < std::min(mcRing.size(), nch); ++c) for (size_t i = 0; i < n; ++i) dst[c][i] = mcRing[c][(mcR + i) & (mcCap - 1)];
        mcR = (mcR + n) & (mcCap - 1);
        return true;
    }
    float scratch[128]; size_t scratchSize = 0; bool relayReady = false; std::vector<float> relayBuffer;
    std::map<std::string, PropValue> props;   // GraphNode::props (GraphNode.h:60-63), what processEvents reads
    std::vector<float> window;                // fft (wasm/FFT.h:49-62)

    int32_t spTickTime(int32_t offset);
    std::map<int32_t, float>::iterator spFind(int32_t tickTime);
};

struct RootSeq { int32_t root; std::vector<int32_t> order; std::vector<int32_t> tapOuts; };
struct RenderSeq { std::vector<RootSeq> subseqs; };

struct Engine {
    double sr;
    int bs;
    std::unordered_map<int32_t, Node> nodes;
    std::set<int32_t> currentRoots;
    std::map<std::string, std::shared_ptr<Resource>> resources;
    std::map<std::string, std::vector<float>> taps;
    std::shared_ptr<RenderSeq> queued, active;
    int64_t sampleTime = 0;   // what the wasm host hands to nodes as userData (wasm/Main.cpp:206-217)

    Engine(double s, int b) : sr(s), bs(b) {}

    static bool known(const std::string& t) {
        static const std::set<std::string> k = {
            "in", "sin", "cos", "tan", "tanh", "asinh", "ln", "log", "log2", "ceil", "floor", "round", "sqrt", "exp", "abs",
            "le", "leq", "ge", "geq", "pow", "eq", "and", "or", "add", "sub", "mul", "div", "mod", "min", "max",
            "root", "const", "phasor", "sphasor", "sr", "counter", "accum", "latch", "maxhold", "rand",
            "delay", "sdelay", "z", "pole", "env", "biquad", "prewarp", "mm1p", "svf", "svfshelf",
            "tapIn", "tapOut", "table", "blepsaw", "blepsquare", "bleptriangle", "meter", "scope", "snapshot", "capture", "fft", "convolve",
            "once", "seq", "seq2", "sparseq", "sparseq2", "time", "metro"};
        return k.count(t) > 0;
    }

    static int bitceil(int n) { if ((n & (n - 1)) == 0) return n; int o = 1; while (o < n) o <<= 1; return o; }   // BitUtils.h:9

    // ---- Runtime.h:294-313 ----
    int createNode(int32_t id, const std::string& type) {
        if (!known(type)) return 1;
        if (nodes.count(id)) return 3;
        Node n;
        n.id = id; n.type = type; n.eng = this; n.out.assign(bs, 0.0f);
        if (type == "root") { n.fade.init(sr); n.channel = -1; }
        if (type == "delay") { n.pendingRing.assign(bs, 0.0f); n.ringPending = true; }                       // Delays.h:56
        if (type == "sdelay") { n.pendingRing.assign(bitceil(bs + bs), 0.0f); n.ringPending = true; n.length = bs; }   // Delays.h:183
        if (type == "tapOut") n.tapPrivate.assign(bs, 0.0f);                                                  // Feedback.h:63-66
        if (type == "metro") n.intervalSamps = (int64_t) std::max(2.0, 1000.0 * 0.001 * sr);                 // Metro.h:14-18,30-33
        if (type == "scope") { n.mcInit(4, 8192); PropValue c1; c1.num = 1; n.props["channels"] = c1; PropValue sz; sz.num = 512; n.props["size"] = sz; }   // Analyzers.h:147-153
        if (type == "capture") n.mcInit(1, (size_t) bitceil((int) (size_t) sr));                              // Capture.h:15-19
        if (type == "fft") { n.mcInit(1, 8192); nodes.emplace(id, std::move(n)); PropValue sz; sz.num = 1024; return setProperty(id, "size", sz); }   // wasm/FFT.h:18-26
        nodes.emplace(id, std::move(n));
        return 0;
    }

    // ---- Runtime.h:336-366 ----
    int appendChild(int32_t p, int32_t c, int ch) {
        if (!nodes.count(p) || !nodes.count(c)) return 2;
        nodes.at(p).inlets.push_back({c, ch});
        return 0;
    }

    // ---- Runtime.h:316-333 and each node's setProperty ----
    int setProperty(int32_t id, const std::string& key, const PropValue& v) {
        auto it = nodes.find(id);
        if (it == nodes.end()) return 2;
        Node& n = it->second;
        const std::string& t = n.type;
        const bool isNum = v.kind == 'N', isStr = v.kind == 'S', isBool = v.kind == 'B';
        if (t == "const" && key == "value") { if (!isNum) return 5; n.value = (float) v.num; }              // Core.h:142-152
        if (t == "root") {                                                                                  // Core.h:33-64
            if (key == "active") { if (!isBool) return 5; if (v.num != 0) n.fade.fadeIn(); else n.fade.fadeOut(); n.activeProp = v.num != 0; }
            if (key == "channel") { if (!isNum) return 8; n.channel = (int) v.num; }
            if (key == "fadeInMs") { if (!isNum) return 5; n.fade.setIn(sr, v.num); }
            if (key == "fadeOutMs") { if (!isNum) return 5; n.fade.setOut(sr, v.num); }
        }
        if (t == "in" && key == "channel") { if (!isNum) return 5; n.channel = (int) v.num; }               // Math.h:95-105
        if (t == "svf" && key == "mode") {                                                                  // SVF.h:30-46
            if (!isStr) return 5;
            if (v.str == "lowpass") n.mode = 0; if (v.str == "bandpass") n.mode = 1; if (v.str == "highpass") n.mode = 2;
            if (v.str == "notch") n.mode = 3; if (v.str == "allpass") n.mode = 4;
        }
        if (t == "svfshelf" && key == "mode") {                                                             // SVFShelf.h:30-44
            if (!isStr) return 5;
            if (v.str == "lowshelf") n.mode = 0; if (v.str == "highshelf") n.mode = 1; if (v.str == "bell" || v.str == "peak") n.mode = 2;
        }
        if (t == "mm1p" && key == "mode") {                                                                 // MultiMode1p.h:50-65
            if (!isStr) return 5;
            if (v.str == "lowpass") n.mode = 0; if (v.str == "highpass") n.mode = 2; if (v.str == "allpass") n.mode = 4;
        }
        if (t == "delay" && key == "size") {                                                                // Delays.h:59-76
            if (!isNum) return 5;
            n.pendingRing.assign(std::max(0, (int) v.num), 0.0f); n.ringPending = true;
        }
        if (t == "sdelay" && key == "size") {                                                               // Delays.h:188-206
            if (!isNum) return 5;
            const int len = std::max(0, (int) v.num);
            n.pendingRing.assign(bitceil(len + bs), 0.0f); n.ringPending = true; n.length = len;
        }
        if (t == "maxhold" && key == "hold") {                                                              // Core.h:292-303
            if (!isNum) return 5;
            n.holdTime = (uint32_t) (sr * 0.001 * v.num);
        }
        if (t == "rand" && key == "seed") { if (!isNum) return 5; n.seed = (uint32_t) v.num; }             // Noise.h:13-23
        if ((t == "tapIn" || t == "tapOut") && key == "name") {                                             // Feedback.h:24-38,71-85
            if (!isStr) return 5;
            n.tapName = v.str;
            if (!taps.count(v.str)) taps[v.str].assign(bs, 0.0f);
        }
        if (t == "scope") {                                                                                 // Analyzers.h:155-179
            if (key == "size") { if (!isNum) return 5; if (v.num < 256 || v.num > 8192) return 6; }
            if (key == "channels") { if (!isNum) return 5; if (v.num < 0 || v.num > 4) return 6; }
            if (key == "name" && !isStr) return 5;
        }
        if (t == "fft") {                                                                                   // wasm/FFT.h:32-71
            if (key == "size") {
                if (!isNum) return 5;
                const int size = (int) v.num;
                if (!(size > 0 && (size & (size - 1)) == 0) || size < 256 || size > 8192) return 6;
                n.window.resize(size);
                for (int i = 0; i < size; ++i) {
                    float const a0 = 0.35875, a1 = 0.48829, a2 = 0.14128, a3 = 0.01168;
                    float const pi = 3.1415926535897932385;
                    float const t1 = a1 * std::cos(2.0 * pi * (i / (double) (size - 1)));
                    float const t2 = a2 * std::cos(4.0 * pi * (i / (double) (size - 1)));
                    float const t3 = a3 * std::cos(6.0 * pi * (i / (double) (size - 1)));
                    n.window[i] = a0 - t1 + t2 - t3;
                }
            }
            if (key == "name" && !isStr) return 5;
        }
        if (t == "once" && key == "arm") {                                                                  // Core.h:345-361
            if (!isBool) return 5;
            if (n.armed == 0.0f) n.armed = (float) (v.num != 0);
        }
        if (t == "seq" || t == "seq2") {                                                                    // Core.h:411-466, Seq2.h:39-84
            if (key == "hold") { if (!isBool) return 5; n.wantsHold = v.num != 0; }
            if (key == "loop") { if (!isBool) return 5; n.wantsLoop = v.num != 0; }
            if (key == "offset") { if (!isNum) return 5; if (v.num < 0.0) return 6; n.seqOffset = (size_t) v.num; }
            if (key == "seq") {
                if (v.kind != 'A') return 5;
                auto d = std::make_shared<std::vector<float>>();
                for (double x : v.arr) d->push_back((float) x);
                n.seqPending = d;
            }
        }
        if (t == "sparseq") {                                                                               // SparSeq.h:40-124
            if (key == "offset") { if (!isNum) return 5; if (v.num < 0.0) return 6; n.seqOffset = (size_t) v.num; }
            if (key == "loop") {
                Node::SpEvent e; e.isLoop = true;
                if (v.kind == 'U' || (isBool && v.num == 0)) { e.ls = -1; e.le = -1; }
                else { if (v.kind != 'A') return 5; e.ls = (int32_t) v.arr.at(0); e.le = (int32_t) v.arr.at(1); }
                n.spQueue.push_back(e);
            }
            if (key == "follow") { if (!isBool) return 5; n.follow = v.num != 0; }
            if (key == "interpolate") { if (!isNum) return 5; n.holdOrder = (int32_t) v.num; }
            if (key == "tickInterval") { if (!isNum) return 5; if (v.num < 0.0) return 6; n.tickInterval = sr * v.num; }
            if (key == "seq") {
                if (v.kind != 'M' && v.kind != 'A') return 5;
                Node::SpEvent e;
                e.seq = std::make_shared<std::map<int32_t, float>>();
                for (auto& o : v.objs) e.seq->insert({(int32_t) o.at("tickTime"), (float) o.at("value")});
                n.spQueue.push_back(e);
            }
        }
        if (t == "sparseq2") {                                                                              // SparSeq2.h:20-56
            if (key == "seq") {
                if (v.kind != 'M' && v.kind != 'A') return 5;
                auto d = std::make_shared<std::map<double, float>>();
                for (auto& o : v.objs) d->insert({o.at("time"), (float) o.at("value")});
                n.sp2Pending = d;
            }
            if (key == "interpolate") { if (!isNum) return 5; n.interpOrder = (int32_t) v.num; }
        }
        if (t == "metro" && key == "interval") {                                                            // Metro.h:20-37
            if (!isNum) return 5;
            if (0 >= v.num) return 6;
            n.intervalSamps = (int64_t) std::max(2.0, v.num * 0.001 * sr);
        }
        if (t == "convolve" && key == "path") {                                                             // wasm/Convolve.h:35-56
            if (!isStr) return 5;
            auto r = resources.find(v.str);
            if (r == resources.end()) return 6;
            auto co = std::make_shared<TwoStageConv>();
            co->init(512, 4096, r->second->data.data(), r->second->data.size());
            n.pendingConv = co;
        }
        if (t == "table" && key == "path") {                                                                // Table.h:20-34
            if (!isStr) return 5;
            auto r = resources.find(v.str);
            if (r == resources.end()) return 6;
            n.pendingRes = r->second;
        }
        n.props[key] = v;   // GraphNode.h:60-63
        return 0;
    }

    // ---- Runtime.h:369-433 ----
    int activateRoots(const std::vector<int32_t>& ids) {
        std::set<int32_t> act;
        for (int32_t id : ids) {
            auto it = nodes.find(id);
            if (it == nodes.end()) return 2;
            if (it->second.type == "root") { it->second.fade.fadeIn(); it->second.activeProp = true; act.insert(id); }
        }
        for (int32_t id : currentRoots) {
            auto it = nodes.find(id);
            if (it == nodes.end() || it->second.type != "root") continue;
            Node& n = it->second;
            if (!act.count(id)) { n.fade.fadeOut(); n.activeProp = false; }
            if (n.fade.on() || !n.fade.settled()) act.insert(id);
        }
        currentRoots.swap(act);
        return 0;
    }

    // ---- Runtime.h:503-518 ----
    void traverse(std::set<int32_t>& visited, std::vector<int32_t>& order, int32_t id) {
        if (visited.count(id)) return;
        for (auto& in : nodes.at(id).inlets) traverse(visited, order, in.first);
        order.push_back(id);
        visited.insert(id);
    }

    // ---- Runtime.h:521-577 ----
    void buildRenderSequence() {
        auto seq = std::make_shared<RenderSeq>();
        std::list<int32_t> sorted;
        for (int32_t id : currentRoots) {
            Node& n = nodes.at(id);
            if (n.type != "root") continue;
            if (n.activeProp) sorted.push_front(id); else sorted.push_back(id);
        }
        std::set<int32_t> visited;
        for (int32_t rid : sorted) {
            RootSeq rs;
            rs.root = rid;
            traverse(visited, rs.order, rid);
            for (int32_t nid : rs.order) if (nodes.at(nid).type == "tapOut") rs.tapOuts.push_back(nid);   // GraphRenderSequence.h:113-115
            seq->subseqs.push_back(std::move(rs));
        }
        queued = seq;
    }

    void processNode(Node& n, const float* const* hostIn, int nHostIn, int ns);
    void process(const float* const* in, int nIn, float* const* out, int nOut, int ns);
};


// SparSeq.h:147-193
int32_t Node::spTickTime(int32_t offset) {
    int32_t tickTime = offset + edgeCount;
    const int32_t ls = loopStart, le = loopEnd;
    if ((ls > -1) && (le > -1) && (tickTime >= le)) {
        const int32_t loopDuration = le - ls;
        if (loopDuration > 0) {
            if (hasPending) {
                loopStart = pendStart; loopEnd = pendEnd; hasPending = false;
                const int32_t nls = loopStart, nle = loopEnd;
                if ((nls == -1) && (nle == -1)) return tickTime;
                if (nle - nls != 0) tickTime = nls + ((tickTime - le) % (nle - nls));
            } else {
                tickTime = ls + ((tickTime - le) % loopDuration);
            }
            edgeCount = tickTime - offset;
        }
    }
    return tickTime;
}

// SparSeq.h:126-145
std::map<int32_t, float>::iterator Node::spFind(int32_t tickTime) {
    if (spActive->empty()) return spActive->end();
    auto it = spActive->upper_bound(tickTime);
    if (it == spActive->begin()) return (it->first == 0) ? it : spActive->end();
    return --it;
}

// ---- one node over one block: the builtins ---------------------------------------------------------------------
void Engine::processNode(Node& n, const float* const* hostIn, int nHostIn, int ns) {
    // Inputs: child buffers, or for a leaf the host input channels (GraphRenderSequence.h:126-135,171-186)
    std::vector<const float*> in;
    if (n.inlets.empty()) { for (int c = 0; c < nHostIn; ++c) in.push_back(hostIn[c]); }
    else for (auto& il : n.inlets) in.push_back(nodes.at(il.first).out.data());
    const int nch = (int) in.size();
    float* out = n.out.data();
    const std::string& t = n.type;
    auto zeros = [&]() { std::fill_n(out, ns, 0.0f); };

    auto unary = [&](float (*f)(float)) { if (nch < 1) return zeros(); for (int i = 0; i < ns; ++i) out[i] = f(in[0][i]); };   // Math.h:9-28
    auto binary = [&](auto f) {                                                                                                 // Math.h:30-57
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) out[i] = f(in[0][i], in[1][i]);
    };
    auto reduce = [&](auto f) {                                                                                                 // Math.h:59-89
        if (nch < 1) return zeros();
        for (int i = 0; i < ns; ++i) out[i] = in[0][i];
        for (int c = 1; c < nch; ++c) for (int i = 0; i < ns; ++i) out[i] = f(out[i], in[c][i]);
    };

    if (t == "sin") return unary([](float x) { return std::sin(x); });
    if (t == "cos") return unary([](float x) { return std::cos(x); });
    if (t == "tan") return unary([](float x) { return std::tan(x); });
    if (t == "tanh") return unary([](float x) { return std::tanh(x); });
    if (t == "asinh") return unary([](float x) { return std::asinh(x); });
    if (t == "ln") return unary([](float x) { return std::log(x); });
    if (t == "log") return unary([](float x) { return std::log10(x); });
    if (t == "log2") return unary([](float x) { return std::log2(x); });
    if (t == "ceil") return unary([](float x) { return std::ceil(x); });
    if (t == "floor") return unary([](float x) { return std::floor(x); });
    if (t == "round") return unary([](float x) { return std::round(x); });
    if (t == "sqrt") return unary([](float x) { return std::sqrt(x); });
    if (t == "exp") return unary([](float x) { return std::exp(x); });
    if (t == "abs") return unary([](float x) { return std::abs(x); });

    if (t == "le") return binary([](float x, float y) { return (float) (x < y); });
    if (t == "leq") return binary([](float x, float y) { return (float) (x <= y); });
    if (t == "ge") return binary([](float x, float y) { return (float) (x > y); });
    if (t == "geq") return binary([](float x, float y) { return (float) (x >= y); });
    if (t == "pow") return binary([](float x, float y) { return (x < 0.0f && y != std::floor(y)) ? 0.0f : std::pow(x, y); });   // Math.h:179-188
    if (t == "eq") return binary([](float x, float y) { return (float) (std::abs(x - y) <= kEps); });                          // Math.h:142-147
    if (t == "and") return binary([](float x, float y) { return (float) (std::abs(1.0f - x) <= kEps && std::abs(1.0f - y) <= kEps); });
    if (t == "or") return binary([](float x, float y) { return (float) (std::abs(1.0f - x) <= kEps || std::abs(1.0f - y) <= kEps); });

    if (t == "add") return reduce([](float x, float y) { return x + y; });
    if (t == "sub") return reduce([](float x, float y) { return x - y; });
    if (t == "mul") return reduce([](float x, float y) { return x * y; });
    if (t == "div") return reduce([](float x, float y) { return y == 0.0f ? 0.0f : x / y; });   // Math.h:135-140
    if (t == "mod") return reduce([](float x, float y) { return std::fmod(x, y); });
    if (t == "min") return reduce([](float x, float y) { return std::min(x, y); });
    if (t == "max") return reduce([](float x, float y) { return std::max(x, y); });

    if (t == "in") {   // Math.h:107-122
        const size_t ch = (size_t) n.channel;
        if (ch >= (size_t) nch) return zeros();
        std::copy_n(in[ch], ns, out);
        return;
    }
    if (t == "meter") {   // Analyzers.h:23-40
        if (nch < 1) return zeros();
        std::copy_n(in[0], ns, out);
        auto mm = std::minmax_element(in[0], in[0] + ns);
        n.roPush({*mm.first, *mm.second});
        return;
    }
    if (t == "snapshot") {   // Analyzers.h:80-106
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) {
            const float l = in[0][i], x = in[1][i];
            if (std::abs(n.z) <= kEps && l > kEps) n.roPush({x, 0.0f});
            n.z = l;
            out[i] = x;
        }
        return;
    }
    if (t == "scope") {   // Analyzers.h:181-201
        if (nch < 1) return zeros();
        std::copy_n(in[0], ns, out);
        n.mcWrite(in.data(), (size_t) nch, (size_t) ns);
        return;
    }
    if (t == "capture") {   // Capture.h:22-58
        if (nch < 2) return zeros();
        std::copy_n(in[1], ns, out);
        for (int i = 0; i < ns; ++i) {
            const bool g = static_cast<bool>(in[0][i]);
            const bool falling = n.change(in[0][i]) < -0.5f;
            if (falling || n.scratchSize >= 128) {
                const float* wd = n.scratch;
                n.mcWrite(&wd, 1, n.scratchSize);
                n.scratchSize = 0;
                if (falling) n.relayReady = true;
            }
            if (g) n.scratch[n.scratchSize++] = in[1][i];
        }
        return;
    }
    if (t == "fft") {   // wasm/FFT.h:73-88
        if (nch < 1) return zeros();
        std::copy_n(in[0], ns, out);
        n.mcWrite(in.data(), 1, (size_t) ns);
        return;
    }
    if (t == "time") {   // wasm/SampleTime.h:16-23
        for (int i = 0; i < ns; ++i) out[i] = (float) static_cast<double>(sampleTime + i);
        return;
    }
    if (t == "metro") {  // wasm/Metro.h:39-55
        const double is = (double) n.intervalSamps;
        for (int i = 0; i < ns; ++i) {
            const double tt = (double) (sampleTime + i) / is;
            const float nextOut = (float) ((tt - std::floor(tt)) < 0.5);
            if (n.lastOut < 0.5f && nextOut >= 0.5f) n.metroFlag = true;
            out[i] = nextOut;
            n.lastOut = nextOut;
        }
        return;
    }
    if (t == "once") {   // Core.h:363-396
        if (nch < 1) return zeros();
        const float isArmed = n.armed;
        for (int i = 0; i < ns; ++i) {
            const float delta = n.change(in[0][i]);
            if (isArmed && delta > 0.5f) { n.gain = 1.0f; n.armed = 0.0f; }
            if (delta < -0.5f) n.gain = 0.0f;
            out[i] = in[0][i] * n.gain;
        }
        return;
    }
    if (t == "seq") {    // Core.h:468-555
        if (n.seqPending) {
            n.seqActive = n.seqPending; n.seqPending.reset();
            n.seqIndex = n.seqIndex % n.seqActive->size();
            if (n.firstPulse) n.holdValue = n.seqActive->at(n.seqIndex);
        }
        if (nch < 1 || !n.seqActive) return zeros();
        const bool hasReset = nch > 1, hold = n.wantsHold, loop = n.wantsLoop;
        for (int i = 0; i < ns; ++i) {
            const float x = in[0][i];
            const float reset = hasReset ? in[1][i] : 0.0f;
            if (n.resetChange(reset) > 0.5f) n.seqIndex = n.seqOffset;
            if (n.change(x) > 0.5f) {
                n.holdValue = n.seqActive->at(std::min(n.seqIndex, n.seqActive->size() - 1));
                n.firstPulse = true;
                if ((++n.seqIndex >= n.seqActive->size()) && loop) n.seqIndex = 0;
            }
            if (n.seqIndex < n.seqActive->size()) out[i] = hold ? n.holdValue : n.holdValue * x;
            else out[i] = hold ? n.holdValue : 0.0f;
        }
        return;
    }
    if (t == "seq2") {   // Seq2.h:87-148
        if (n.seqPending) { n.seqActive = n.seqPending; n.seqPending.reset(); }
        if (nch < 1 || !n.seqActive) return zeros();
        const bool hasReset = nch > 1, hold = n.wantsHold, loop = n.wantsLoop;
        const size_t offset = n.seqOffset;
        for (int i = 0; i < ns; ++i) {
            const float x = in[0][i];
            const float reset = hasReset ? in[1][i] : 0.0f;
            if (n.change(x) > 0.5f) n.edgeCount2++;
            if (n.resetChange(reset) > 0.5f) n.edgeCount2 = 0;
            const size_t size = n.seqActive->size();
            const size_t idx = offset + n.edgeCount2;
            const float nextOut = (idx < size) ? n.seqActive->at(idx)
                                : (loop ? n.seqActive->at(idx % size) : (hold ? n.seqActive->at(size - 1) : 0.0f));
            out[i] = hold ? nextOut : nextOut * x;
        }
        return;
    }
    if (t == "sparseq") {   // SparSeq.h:195-333
        const bool hasReset = nch > 1;
        const int32_t offset = (int32_t) n.seqOffset;
        int32_t tickTime = n.spTickTime(offset);
        if (!n.spQueue.empty()) {
            for (auto& e : n.spQueue) {
                if (e.isLoop) { n.hasPending = true; n.pendStart = e.ls; n.pendEnd = e.le; }
                else n.spActive = e.seq;
            }
            n.spQueue.clear();
            if (n.spActive) n.spHold = n.spFind(tickTime);
        }
        if (n.hasPending) {
            const bool takeImmediately = (n.loopStart == -1 && n.loopEnd == -1) || !n.follow;
            if (takeImmediately) { n.loopStart = n.pendStart; n.loopEnd = n.pendEnd; n.hasPending = false; tickTime = n.spTickTime(offset); }
        }
        if (nch < 1 || !n.spActive) return zeros();
        for (int i = 0; i < ns; ++i) {
            n.samplesSince++;
            const float x = in[0][i];
            const float reset = hasReset ? in[1][i] : 0.0f;
            const bool trig = n.change(x) > 0.5f;
            const bool rst = n.resetChange(reset) > 0.5f;
            if (rst) n.edgeCount = 0;
            if (trig) {
                n.edgeCount = rst ? 0 : n.edgeCount + 1;
                n.samplesSince = 0;
                tickTime = n.spTickTime(offset);
                n.spHold = n.spFind(tickTime);
            }
            if (n.spHold == n.spActive->end()) { out[i] = 0.0f; continue; }
            if (n.holdOrder == 1) {
                auto right = std::next(n.spHold);
                if (right == n.spActive->end()) { out[i] = n.spHold->second; continue; }
                const int32_t tl = n.spHold->first, tr = right->first;
                const float lv = n.spHold->second, rv = right->second;
                double alpha = (double) std::max(0, tickTime - tl) / (double) (tr - tl);
                if (n.tickInterval > 0.0) alpha += (std::min((double) n.samplesSince, n.tickInterval) / n.tickInterval) / (double) (tr - tl);
                out[i] = lv + alpha * (rv - lv);
            } else out[i] = n.spHold->second;
        }
        return;
    }
    if (t == "sparseq2") {  // SparSeq2.h:69-128
        if (n.sp2Pending) {
            n.sp2Active = n.sp2Pending; n.sp2Pending.reset();
            n.sp2Prev = n.sp2Active->end(); n.sp2Next = n.sp2Active->end();
        }
        if (nch < 1 || !n.sp2Active || n.sp2Active->empty()) return zeros();
        const auto end = n.sp2Active->end();
        const bool interp = n.interpOrder == 1;
        for (int i = 0; i < ns; ++i) {
            const double tt = (double) in[0][i];
            const bool update = (n.sp2Prev == end && n.sp2Next == end)
                || (n.sp2Prev != end && tt <= (n.sp2Prev->first + 1e-9))
                || (n.sp2Next != end && tt >= (n.sp2Next->first - 1e-9));
            if (update) {
                n.sp2Next = n.sp2Active->upper_bound(tt);
                n.sp2Prev = (n.sp2Next == n.sp2Active->begin()) ? end : std::prev(n.sp2Next);
            }
            if (n.sp2Prev == end) { out[i] = 0.0f; continue; }
            if (n.sp2Next == end) { out[i] = n.sp2Prev->second; continue; }
            const double alpha = interp ? ((tt - n.sp2Prev->first) / (n.sp2Next->first - n.sp2Prev->first)) : 0.0;
            out[i] = n.sp2Prev->second + (float) alpha * (n.sp2Next->second - n.sp2Prev->second);
        }
        return;
    }
    if (t == "const") { std::fill_n(out, ns, n.value); return; }          // Core.h:154-163
    if (t == "sr") { std::fill_n(out, ns, (float) sr); return; }          // Core.h:173-180

    if (t == "root") {   // Core.h:66-78
        if (nch < 1) return zeros();
        n.fade.process(in[0], out, ns);
        return;
    }
    if (t == "phasor" || t == "sphasor") {   // Core.h:89-131
        const bool withReset = t == "sphasor";
        if (nch < (withReset ? 2 : 1)) return zeros();
        for (int i = 0; i < ns; ++i) {
            if (withReset && n.change(in[1][i]) > 0.5f) n.phase = 0.0f;
            const float step = in[0][i] * (1.0f / (float) sr);
            const float y = n.phase;
            const float next = n.phase + step;
            n.phase = next - std::floor(next);
            out[i] = y;
        }
        return;
    }
    if (t == "counter") {   // Core.h:198-211
        if (nch < 1) return zeros();
        for (int i = 0; i < ns; ++i) {
            if ((1.0f - in[0][i]) <= kEps) { out[i] = n.count; n.count = n.count + 1.0f; continue; }
            n.count = 0.0f; out[i] = 0.0f;
        }
        return;
    }
    if (t == "accum") {   // Core.h:233-243
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) {
            if (n.change(in[1][i]) > 0.5f) n.total = 0.0f;
            n.total += in[0][i];
            out[i] = n.total;
        }
        return;
    }
    if (t == "latch") {   // Core.h:265-281
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) {
            const float l = in[0][i], x = in[1][i];
            if (std::abs(n.z) <= kEps && l > kEps) n.hold = x;
            n.z = l;
            out[i] = n.hold;
        }
        return;
    }
    if (t == "maxhold") {   // Core.h:315-332
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) {
            const float x = in[0][i], reset = in[1][i];
            if (n.change(reset) > 0.5f || ++n.heldSamples >= n.holdTime) { n.mx = x; n.heldSamples = 0; }
            else if (x > n.mx) { n.heldSamples = 0; n.mx = x; }
            out[i] = n.mx;
        }
        return;
    }
    if (t == "rand") {   // Noise.h:25-38
        for (int i = 0; i < ns; ++i) {
            n.seed = 214013u * n.seed + 2531011u;
            out[i] = (int) ((n.seed >> 16) & 0x7FFF) / (float) 0x7FFF;
        }
        return;
    }
    if (t == "pole") {   // Filters.h:27-33
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) { n.z1 = in[1][i] + in[0][i] * n.z1; out[i] = n.z1; }
        return;
    }
    if (t == "env") {   // Filters.h:61-73
        if (nch < 3) return zeros();
        for (int i = 0; i < ns; ++i) {
            const float ap = in[0][i], rp = in[1][i], vn = std::abs(in[2][i]);
            if (std::abs(vn) > n.z1) n.z1 = ap * (n.z1 - vn) + vn; else n.z1 = rp * (n.z1 - vn) + vn;
            out[i] = n.z1;
        }
        return;
    }
    if (t == "biquad") {   // Filters.h:102-114
        if (nch < 6) return zeros();
        for (int i = 0; i < ns; ++i) {
            const float b0 = in[0][i], b1 = in[1][i], b2 = in[2][i], a1 = in[3][i], a2 = in[4][i], x = in[5][i];
            const float y = b0 * x + n.z1;
            n.z1 = b1 * x - a1 * y + n.z2;
            n.z2 = b2 * x - a2 * y;
            out[i] = y;
        }
        return;
    }
    if (t == "prewarp") {   // MultiMode1p.h:23-33
        if (nch < 1) return zeros();
        const double T = 1.0 / sr;
        for (int i = 0; i < ns; ++i) {
            const double twoPi = 2.0 * 3.141592653589793238;
            const double wd = twoPi * (double) in[0][i];
            out[i] = (float) std::tan(wd * T / 2.0);
        }
        return;
    }
    if (t == "mm1p") {   // MultiMode1p.h:78-103
        if (nch < 2) return zeros();
        for (int i = 0; i < ns; ++i) {
            const double g = std::clamp((double) in[0][i], 0.0, 0.9999);
            const float xn = in[1][i];
            const double G = g / (1.0 + g);
            const double v = ((double) xn - n.dz) * G;
            const double lp = v + n.dz;
            n.dz = lp + v;
            if (n.mode == 0) out[i] = (float) lp;
            else if (n.mode == 2) out[i] = xn - (float) lp;
            else out[i] = (float) (lp + lp - xn);
        }
        return;
    }
    if (t == "svf") {   // SVF.h:48-104
        if (nch < 3) return zeros();
        for (int i = 0; i < ns; ++i) {
            const double fc = in[0][i], q = in[1][i];
            const float v0 = in[2][i];
            const double g = std::tan(3.14159265359 * std::clamp(fc, 20.0, sr / 2.0001) / sr);
            const double k = 1.0 / std::clamp(q, 0.25, 20.0);
            const double a1 = 1.0 / (1.0 + g * (g + k)), a2 = g * a1, a3 = g * a2;
            const double v3 = v0 - n.ic2;
            const double v1 = n.ic1 * a1 + v3 * a2;
            const double v2 = n.ic2 + n.ic1 * a2 + v3 * a3;
            n.ic1 = v1 * 2.0 - n.ic1;
            n.ic2 = v2 * 2.0 - n.ic2;
            switch (n.mode) {
                case 0: out[i] = (float) v2; break;
                case 1: out[i] = (float) v1; break;
                case 2: out[i] = (float) (v0 - k * v1 - v2); break;
                case 3: out[i] = (float) (v0 - k * v1); break;
                default: out[i] = (float) (v0 - 2.0 * k * v1); break;
            }
        }
        return;
    }
    if (t == "svfshelf") {   // SVFShelf.h:44-106
        if (nch < 4) return zeros();
        for (int i = 0; i < ns; ++i) {
            const double fc = in[0][i], q = in[1][i], gdb = in[2][i];
            const float v0 = in[3][i];
            const double A = std::pow(10, gdb / 40.0);
            double g = std::tan(3.14159265359 * std::clamp(fc, 20.0, sr / 2.0001) / sr);
            double k = 1.0 / std::clamp(q, 0.25, 20.0);
            if (n.mode == 0) g /= A;
            if (n.mode == 1) g *= A;
            if (n.mode == 2) k /= A;
            const double a1 = 1.0 / (1.0 + g * (g + k)), a2 = g * a1, a3 = g * a2;
            const double v3 = v0 - n.ic2;
            const double v1 = n.ic1 * a1 + v3 * a2;
            const double v2 = n.ic2 + n.ic1 * a2 + v3 * a3;
            n.ic1 = v1 * 2.0 - n.ic1;
            n.ic2 = v2 * 2.0 - n.ic2;
            if (n.mode == 2) out[i] = (float) (v0 + k * (A * A - 1.0) * v1);
            else if (n.mode == 0) out[i] = (float) (v0 + k * (A - 1.0) * v1 + (A * A - 1.0) * v2);
            else out[i] = (float) (A * A * v0 + k * (1.0 - A) * A * v1 + (1.0 - A * A) * v2);
        }
        return;
    }
    if (t == "z") {   // Delays.h:29-34
        if (nch < 1) return zeros();
        for (int i = 0; i < ns; ++i) { const float x = in[0][i]; out[i] = n.z; n.z = x; }
        return;
    }
    if (t == "delay") {   // Delays.h:87-159
        if (n.ringPending) { n.ring.swap(n.pendingRing); n.ringPending = false; n.writeIndex = 0; }
        if (nch < 3) return zeros();
        const int size = (int) n.ring.size();
        float* d = n.ring.data();
        if (size == 0) { std::copy_n(in[0], ns, out); return; }
        for (int i = 0; i < ns; ++i) {
            const float offset = std::clamp(in[0][i], 0.0f, (float) size);
            if (offset <= kEps) {
                const float x = in[2][i];
                d[n.writeIndex] = x; out[i] = x;
                if (++n.writeIndex >= size) n.writeIndex -= size;
                continue;
            }
            const float readFrac = (float) (size + n.writeIndex) - offset;
            const int readLeft = (int) readFrac, readRight = readLeft + 1;
            const float frac = readFrac - std::floor(readFrac);
            const float left = d[readLeft % size], right = d[readRight % size];
            const float o = left + frac * (right - left);
            const float fb = std::clamp(in[1][i], -1.0f, 1.0f);
            d[n.writeIndex] = in[2][i] + fb * o;
            out[i] = o;
            if (++n.writeIndex >= size) n.writeIndex -= size;
        }
        return;
    }
    if (t == "sdelay") {   // Delays.h:221-260
        if (n.ringPending) { n.ring.swap(n.pendingRing); n.ringPending = false; n.writeIndex = 0; }
        const int size = (int) n.ring.size();
        if (nch < 1 || size == 0) return zeros();
        const int mask = size - 1, len = n.length;
        float* d = n.ring.data();
        const int readStart = n.writeIndex - len;
        for (int i = 0; i < ns; ++i) { d[n.writeIndex] = in[0][i]; n.writeIndex = (n.writeIndex + 1) & mask; }
        for (int i = 0; i < ns; ++i) out[i] = d[(size + readStart + i) & mask];
        return;
    }
    if (t == "table") {   // Table.h:39-71
        if (n.pendingRes) { n.res = n.pendingRes; n.pendingRes.reset(); }
        if (nch == 0 || !n.res || n.res->data.empty()) return zeros();
        const int size = (int) n.res->data.size();
        const float* d = n.res->data.data();
        for (int i = 0; i < ns; ++i) {
            const float readPos = std::clamp(in[0][i], 0.0f, 1.0f) * (float) (size - 1);
            const int readLeft = (int) readPos, readRight = readLeft + 1;
            const float frac = readPos - std::floor(readPos);
            const float left = d[readLeft % size], right = d[readRight % size];
            out[i] = left + frac * (right - left);
        }
        return;
    }
    if (t == "blepsaw" || t == "blepsquare" || t == "bleptriangle") {   // Oscillators.h:23-89
        if (nch < 1) return zeros();
        const int mode = t == "blepsaw" ? 0 : (t == "blepsquare" ? 1 : 2);
        const float fsr = (float) sr;
        auto blep = [](float phase, float inc) -> float {
            if (phase < inc) { const float p = phase / inc; return (2.0f - p) * p - 1.0f; }
            if (phase > (1.0f - inc)) { const float p = (phase - 1.0f) / inc; return (p + 2.0f) * p + 1.0f; }
            return 0.0f;
        };
        for (int i = 0; i < ns; ++i) {
            const float inc = in[0][i] / fsr;
            float y;
            if (mode == 0) y = 2.0f * n.phase - 1.0f - blep(n.phase, inc);
            else {
                const float naive = n.phase < 0.5f ? 1.0f : -1.0f;
                const float halfPhase = std::fmod(n.phase + 0.5f, 1.0f);
                const float square = naive + blep(n.phase, inc) - blep(halfPhase, inc);
                if (mode == 1) y = square;
                else { n.acc += 4.0f * inc * square; y = n.acc; }
            }
            n.phase += inc;
            if (n.phase >= 1.0f) n.phase -= 1.0f;
            out[i] = y;
        }
        return;
    }
    if (t == "convolve") {   // wasm/Convolve.h:58-85
        if (n.pendingConv) { n.conv = n.pendingConv; n.pendingConv.reset(); }
        if (nch == 0 || !n.conv) return zeros();
        n.conv->process(in[0], out, (size_t) ns);
        return;
    }
    if (t == "tapIn") {   // Feedback.h:42-52
        auto it = taps.find(n.tapName);
        if (n.tapName.empty() || it == taps.end()) return zeros();
        std::copy_n(it->second.data(), ns, out);
        return;
    }
    if (t == "tapOut") {   // Feedback.h:109-121
        if (nch < 1 || ns > (int) n.tapPrivate.size()) return zeros();
        std::copy_n(in[0], ns, n.tapPrivate.data());
        std::copy_n(in[0], ns, out);
        return;
    }
    zeros();
}

// ---- Runtime.h:275-290 + GraphRenderSequence.h:268-309,212-232 ----
void Engine::process(const float* const* in, int nIn, float* const* out, int nOut, int ns) {
    struct Tick { int64_t& t; int n; ~Tick() { t += n; } } tick{sampleTime, ns};   // wasm/Main.cpp:217
    if (queued) { active = queued; queued.reset(); }
    if (!active) return;
    for (int c = 0; c < nOut; ++c) std::fill_n(out[c], ns, 0.0f);
    for (auto& sq : active->subseqs) {
        Node& root = nodes.at(sq.root);
        const size_t ch = (size_t) root.channel;
        const bool running = root.fade.on() || !root.fade.settled();
        if (!running || ch >= (size_t) nOut) continue;
        for (int32_t nid : sq.order) processNode(nodes.at(nid), in, nIn, ns);
        for (int j = 0; j < ns; ++j) out[ch][j] += root.out[j];
    }
    for (auto& sq : active->subseqs) {   // promoteTapBuffers: GraphRenderSequence.h:200-210
        if (!nodes.at(sq.root).fade.on()) continue;
        for (int32_t nid : sq.tapOuts) {
            Node& n = nodes.at(nid);
            if (n.tapName.empty()) continue;
            std::copy_n(n.tapPrivate.data(), ns, taps[n.tapName].data());
        }
    }
}

// ---- Runtime.h:438-446 -> GraphRenderSequence.h:296-304,189-198: events of active roots, nodes in render order -----
static std::string evNum(float f) { if (!std::isfinite(f)) return "null"; char b[40]; std::snprintf(b, sizeof b, "%.9g", (double) f); return b; }
static std::string evSource(const Node& n) {
    auto it = n.props.find("name");
    if (it == n.props.end() || it->second.kind != 'S') return "null";
    return "\"" + it->second.str + "\"";
}
static std::string evArray(const float* d, size_t n) { std::string o = "["; for (size_t i = 0; i < n; ++i) { if (i) o += ", "; o += evNum(d[i]); } return o + "]"; }

std::string processQueuedEvents(Engine& e) {
    std::string out = "[";
    auto emit = [&](const char* type, const std::string& evt) {
        if (out.size() > 1) out += ", ";
        out += std::string("{\"event\": ") + evt + ", \"type\": \"" + type + "\"}";
    };
    if (!e.active) return "[]";
    for (auto& sq : e.active->subseqs) {
        if (!e.nodes.at(sq.root).activeProp) continue;
        for (int32_t nid : sq.order) {
            Node& n = e.nodes.at(nid);
            if (n.type == "meter" || n.type == "snapshot") {          // Analyzers.h:42-60,108-127
                if (n.roSize() == 0) continue;
                Node::Readout ro;
                while (n.roSize() > 0) { ro = n.roQueue[n.roR]; n.roR = (n.roR + 1) & 31; }
                if (n.type == "meter") emit("meter", "{\"max\": " + evNum(ro.b) + ", \"min\": " + evNum(ro.a) + ", \"source\": " + evSource(n) + "}");
                else emit("snapshot", "{\"data\": " + evNum(ro.a) + ", \"source\": " + evSource(n) + "}");
            } else if (n.type == "scope") {                           // Analyzers.h:203-251
                auto num = [&](const char* k, double d) { auto it = n.props.find(k); return (it != n.props.end() && it->second.kind == 'N') ? it->second.num : d; };
                const size_t size = (size_t) num("size", 512), channels = (size_t) num("channels", 1);
                if (!(n.mcFull() > size)) continue;
                std::vector<std::vector<float>> data(channels, std::vector<float>(size, 0.0f));
                std::vector<float*> ptrs(8, nullptr);
                for (size_t c = 0; c < channels; ++c) ptrs[c] = data[c].data();
                if (!n.mcRead(ptrs.data(), channels, size)) continue;
                std::string arr = "[";
                for (size_t c = 0; c < channels; ++c) { if (c) arr += ", "; arr += evArray(data[c].data(), size); }
                emit("scope", "{\"data\": " + arr + "], \"source\": " + evSource(n) + "}");
            } else if (n.type == "capture") {                         // Capture.h:60-93
                const size_t avail = n.mcFull();
                if (avail > 0) {
                    const size_t cur = n.relayBuffer.size();
                    n.relayBuffer.resize(cur + avail);
                    float* dst = n.relayBuffer.data() + cur;
                    if (!n.mcRead(&dst, 1, avail)) continue;
                }
                if (n.relayReady) {
                    n.relayReady = false;
                    emit("capture", "{\"data\": " + evArray(n.relayBuffer.data(), n.relayBuffer.size()) + ", \"source\": " + evSource(n) + "}");
                    n.relayBuffer.clear();
                }
            } else if (n.type == "fft") {                             // wasm/FFT.h:90-131
                const size_t size = n.window.size();
                if (size == 0 || n.mcFull() < size) continue;
                std::vector<float> x(size), re(size / 2 + 1), im(size / 2 + 1);
                float* dst = x.data();
                n.mcRead(&dst, 1, size);
                for (size_t i = 0; i < size; ++i) x[i] *= n.window[i];
                RealFFT f; f.init(size);
                f.fft(x.data(), re.data(), im.data());
                im[0] = 0.0f; im[size / 2] = 0.0f;
                emit("fft", "{\"data\": {\"imag\": " + evArray(im.data(), im.size()) + ", \"real\": " + evArray(re.data(), re.size()) + "}, \"source\": " + evSource(n) + "}");
            } else if (n.type == "metro") {                           // wasm/Metro.h:58-66
                if (n.metroFlag) { n.metroFlag = false; emit("metro", "{\"source\": " + evSource(n) + "}"); }
            }
        }
    }
    return out + "]";
}

} // namespace

extern "C" {

int elem_oracle_process_queued_events(void* h, char* buf, size_t cap) {
    std::string s = processQueuedEvents(*static_cast<Engine*>(h));
    if (buf && cap) { const size_t k = s.size() < cap - 1 ? s.size() : cap - 1; std::memcpy(buf, s.data(), k); buf[k] = 0; }
    return (int) s.size();
}

void* elem_oracle_create(double sr, int bs) { return new Engine(sr, bs); }
void elem_oracle_destroy(void* h) { delete static_cast<Engine*>(h); }

// Line format produced by oracle/oracle.py:batch_to_text (no JSON parser on purpose).
int elem_oracle_apply_text(void* h, const char* text) {
    auto* e = static_cast<Engine*>(h);
    std::istringstream ss(text);
    std::string line;
    bool rebuild = false;
    while (std::getline(ss, line)) {
        if (line.empty()) continue;
        std::istringstream ls(line);
        int op;
        if (!(ls >> op)) return 8;
        int res = 0;
        if (op == 0) { long long id; std::string type; if (!(ls >> id >> type)) return 8; res = e->createNode((int32_t) id, type); }
        else if (op == 2) { long long p, c; int ch; if (!(ls >> p >> c >> ch)) return 8; res = e->appendChild((int32_t) p, (int32_t) c, ch); }
        else if (op == 3) {
            long long id; std::string key; char kind;
            if (!(ls >> id >> key >> kind)) return 8;
            PropValue v; v.kind = kind;
            std::string rest; std::getline(ls, rest);
            if (!rest.empty() && rest[0] == ' ') rest.erase(0, 1);
            if (kind == 'N' || kind == 'B') v.num = std::strtod(rest.c_str(), nullptr);
            else if (kind == 'A') { std::istringstream as(rest); size_t cnt; as >> cnt; double x; while (v.arr.size() < cnt && (as >> x)) v.arr.push_back(x); }
            else if (kind == 'M') {
                std::istringstream ms(rest); size_t cnt, nk; ms >> cnt >> nk;
                for (size_t i = 0; i < cnt; ++i) { std::map<std::string, double> o; for (size_t k = 0; k < nk; ++k) { std::string kk; double x; ms >> kk >> x; o[kk] = x; } v.objs.push_back(o); }
            }
            else v.str = rest;
            res = e->setProperty((int32_t) id, key, v);
        } else if (op == 4) {
            std::vector<int32_t> ids; long long id;
            while (ls >> id) ids.push_back((int32_t) id);
            res = e->activateRoots(ids);
            rebuild = true;
        } else if (op == 5) { if (rebuild) e->buildRenderSequence(); }
        if (res != 0) return res;
    }
    return 0;
}

int elem_oracle_add_shared_resource(void* h, const char* name, const float* data, size_t n) {
    auto* e = static_cast<Engine*>(h);
    if (e->resources.count(name)) return 0;
    auto r = std::make_shared<Resource>();
    r->data.assign(data, data + n);
    e->resources[name] = r;
    return 1;
}

void elem_oracle_set_current_time(void* h, long long t) { static_cast<Engine*>(h)->sampleTime = t; }

void elem_oracle_process_flat(void* h, const float* in, size_t nIn, float* out, size_t nOut, size_t ns) {
    auto* e = static_cast<Engine*>(h);
    std::vector<const float*> ip(nIn);
    std::vector<float*> op(nOut);
    for (size_t i = 0; i < nIn; ++i) ip[i] = in + i * ns;
    for (size_t i = 0; i < nOut; ++i) op[i] = out + i * ns;
    e->process(ip.data(), (int) nIn, op.data(), (int) nOut, (int) ns);
}

} // extern "C"
