// events_host.cpp — the parts of eb::Engine that talk to the outside between blocks: the cross-GPU exchange plumbing
// (CUDA IPC mapping of the K4 buffers) and Runtime::processQueuedEvents (read-back of the per-voice analysis records
// the kernel leaves in HBM, turned into the reference's event objects).  See graph_host.h.
#include "graph_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace eb {

// ---------------------------------------------------------------------------------------------------------
// K4 plumbing: exchange buffers mapped across the ranks of one box with CUDA IPC (one process per GPU)
int Engine::peerExport(void* handleOut64) {
    Lock lk(mu_);
    if (planOnly_) return fail(rc::CudaError, "plan-only runtime has no device memory to export");
    dsetdev();
    if (!dExchange_) {
        const size_t stride = (size_t) MAX_OUT_CHANNELS * blockSize_;
        exchangeFlagOffset_ = sizeof(uint2) * 2 * MAX_PEERS * stride;                       // (sample, epoch) pairs: [2][MAX_PEERS][stride]
        exchangeOwnOffset_ = exchangeFlagOffset_ + sizeof(uint32_t) * 2 * MAX_PEERS + 64;    // then the flags, then this rank's own partial mix
        exchangeOwnOffset_ = (exchangeOwnOffset_ + 255) & ~(size_t) 255;
        exchangeBytes_ = exchangeOwnOffset_ + sizeof(float) * stride;
        if (!cuda(cudaMalloc(&dExchange_, exchangeBytes_), "cudaMalloc exchange buffer")) return rc::CudaError;
        if (!cuda(cudaMemset(dExchange_, 0, exchangeBytes_), "memset exchange buffer")) return rc::CudaError;
        if (!cuda(cudaMalloc((void**) &dPeerStatus_, sizeof(int)), "cudaMalloc peer status")) return rc::CudaError;
        if (!cuda(cudaMemset(dPeerStatus_, 0, sizeof(int)), "memset peer status")) return rc::CudaError;
    }
    cudaIpcMemHandle_t h;
    if (!cuda(cudaIpcGetMemHandle(&h, dExchange_), "cudaIpcGetMemHandle")) return rc::CudaError;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    std::memcpy(handleOut64, &h, 64);
    return rc::Ok;
}

int Engine::peerAttach(int rank, int world, const void* handles) {
    Lock lk(mu_);
    if (planOnly_) return fail(rc::CudaError, "plan-only runtime cannot attach peers");
    if (world < 1 || world > MAX_PEERS || rank < 0 || rank >= world || !dExchange_) return fail(rc::BadArgument, "peerAttach: bad rank/world, or peerExport was not called");
    dsetdev();
    peer_ = PeerMix{};
    peer_.rank = rank; peer_.world = world; peer_.stride = MAX_OUT_CHANNELS * blockSize_;
    for (int p = 0; p < world; ++p) {
        void* base = dExchange_;
        if (p != rank) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, static_cast<const char*>(handles) + (size_t) p * 64, 64);
            if (!cuda(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle (is peer access available between the GPUs?)")) return rc::CudaError;
            peerMapped_.push_back(base);
        }
        peer_.slot[p] = static_cast<uint2*>(base);
        peer_.flag[p] = reinterpret_cast<uint32_t*>(static_cast<char*>(base) + exchangeFlagOffset_);
        if (p == rank) peer_.own = reinterpret_cast<float*>(static_cast<char*>(base) + exchangeOwnOffset_);
    }
    peerAttached_ = true;
    peerEpoch_ = 0;
    return rc::Ok;
}

int Engine::peerStatus() {
    Lock lk(mu_);
    if (!dPeerStatus_) return 0;
    int st = 0;
    cudaStreamSynchronize(stream_);
    cudaMemcpy(&st, dPeerStatus_, sizeof(int), cudaMemcpyDeviceToHost);
    return st;
}

// ---------------------------------------------------------------------------------------------------------
// Runtime::processQueuedEvents (Runtime.h:438-446) -> GraphRenderSequence::processQueuedEvents (:296-304) ->
// RootRenderSequence::processQueuedEvents (:189-198): for every root sub-sequence whose root has active == true, every
// node's processEvents() in render order.  Here the per-voice records the kernel left in HBM are read back and
// turned into the same event objects, one per voice, as JSON text with the reference's keys plus "voice".
static std::string jsonString(const std::string& s) {
    std::string o = "\"";
    for (unsigned char ch : s) {
        if (ch == '"' || ch == '\\') { o += '\\'; o += (char) ch; }
        else if (ch < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
        else o += (char) ch;
    }
    return o + "\"";
}
static std::string jsonNumber(float f) {
    if (!std::isfinite(f)) return "null";                 // nlohmann::json dumps non-finite numbers as null (JSON.h:175-179)
    char b[40];
    std::snprintf(b, sizeof b, "%.9g", (double) f);
    return b;
}
static std::string sourceOf(const Node& n) {              // getPropertyWithDefault("name", js::Value()) — undefined serialises as null
    auto it = n.props.find("name");
    if (it == n.props.end() || !it->second.isString()) return "null";
    return jsonString(it->second.asString());
}
static void appendFloatArray(std::string& o, const float* d, size_t n) {
    o += '[';
    for (size_t i = 0; i < n; ++i) { if (i) o += ", "; o += jsonNumber(d[i]); }
    o += ']';
}

// AudioFFT::fft semantics (wasm/FFTConvolver/AudioFFT.cpp:132-155): n real float samples -> n/2+1 bins, the transform itself in
// double (the reference runs Ooura's rdft in double), Im of bins 0 and n/2 exactly zero.
static void realFftFloatIO(const std::vector<float>& x, std::vector<float>& re, std::vector<float>& im) {
    const size_t n = x.size();
    std::vector<double> ar(x.begin(), x.end()), ai(n, 0.0);
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(ar[i], ar[j]); std::swap(ai[i], ai[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const double ang = -2.0 * M_PI / (double) len;
        for (size_t k = 0; k < len / 2; ++k) {
            const double wr = std::cos(ang * (double) k), wi = std::sin(ang * (double) k);
            for (size_t i = k; i < n; i += len) {
                const size_t j = i + len / 2;
                const double xr = ar[j] * wr - ai[j] * wi, xi = ar[j] * wi + ai[j] * wr;
                ar[j] = ar[i] - xr; ai[j] = ai[i] - xi;
                ar[i] += xr; ai[i] += xi;
            }
        }
    }
    for (size_t k = 0; k <= n / 2; ++k) { re[k] = (float) ar[k]; im[k] = (float) ai[k]; }
    im[0] = 0.0f; im[n / 2] = 0.0f;
}

int Engine::processQueuedEvents(int vb, int ve, EventFn cb, void* user) {
    Lock lk(mu_);
    if (planOnly_) return rc::Ok;
    dsetdev();
    if (vb < 0) vb = 0;
    if (ve < 0 || ve > numVoices_) ve = numVoices_;
    for (auto& gp : groups_) {
        Group& g = *gp;
        if (!g.active || g.active->evNodes.empty()) continue;      // `if (auto ptr = rtRenderSeq)`: the sequence process() last used
        Program& p = *g.active;
        const int b = std::max(vb, g.v0) - g.v0, e = std::min(ve, g.v0 + g.nv) - g.v0;   // group-relative voice range
        if (b >= e) continue;                                           // the poll does not touch this voice group at all
        const int L = g.tileWidth;
        if (!cuda(cudaStreamSynchronize(stream_), "sync before events")) return rc::CudaError;
        auto readRows = [&](int row, int count, std::vector<uint32_t>& out) -> bool {
            out.resize((size_t) count * g.Vpad);
            return cuda(cudaMemcpy(out.data(), g.dRows + (size_t) row * g.Vpad, sizeof(uint32_t) * out.size(), cudaMemcpyDeviceToHost), "read event rows");
        };
        auto asFloat = [](uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; };
        for (size_t ri = 0; ri < p.rootIds.size(); ++ri) {
            auto rit = g.nodes.find(p.rootIds[ri]);
            if (rit == g.nodes.end()) continue;
            auto ap = rit->second.props.find("active");                                     // GraphRenderSequence.h:192
            if (ap == rit->second.props.end() || !ap->second.isBool() || !ap->second.asBool()) continue;
            for (auto& ev : p.evNodes) {
                if (ev.root != (int) ri) continue;
                auto nit = g.nodes.find(ev.node);
                if (nit == g.nodes.end()) continue;
                Node& n = nit->second;
                const std::string src = sourceOf(n);
                std::vector<uint32_t> rows;
                switch (n.kind) {
                case NodeKind::Meter: {        // Analyzers.h:42-60: the latest readout, if the queue shows any
                    if (!readRows(n.stateRow, 3, rows)) return rc::CudaError;
                    for (int v = 0; v < g.nv; ++v) {
                        // SingleWriterSingleReaderQueue(32).size() == pushes mod 32 (:77-83,:86-98): a queue that was
                        // pushed exactly 32 times since it was drained reads as empty
                        if (rows[(size_t) 2 * g.Vpad + v] % 32u == 0 || v < b || v >= e || !cb) continue;
                        std::string js = "{\"max\": " + jsonNumber(asFloat(rows[(size_t) g.Vpad + v])) + ", \"min\": " + jsonNumber(asFloat(rows[v])) +
                                         ", \"source\": " + src + ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                        cb("meter", js.c_str(), g.v0 + v, user);
                    }
                    if (e > b && !cuda(cudaMemsetAsync(g.dRows + (size_t) (n.stateRow + 2) * g.Vpad + b, 0, sizeof(float) * (e - b), stream_), "reset meter queue")) return rc::CudaError;
                } break;
                case NodeKind::Snapshot: {     // Analyzers.h:108-127
                    if (!readRows(n.stateRow, 3, rows)) return rc::CudaError;
                    for (int v = b; v < e; ++v) {
                        if (rows[(size_t) 2 * g.Vpad + v] % 32u == 0 || !cb) continue;
                        std::string js = "{\"data\": " + jsonNumber(asFloat(rows[(size_t) g.Vpad + v])) + ", \"source\": " + src +
                                         ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                        cb("snapshot", js.c_str(), g.v0 + v, user);
                    }
                    if (e > b && !cuda(cudaMemsetAsync(g.dRows + (size_t) (n.stateRow + 2) * g.Vpad + b, 0, sizeof(float) * (e - b), stream_), "reset snapshot queue")) return rc::CudaError;
                } break;
                case NodeKind::Scope: {        // Analyzers.h:203-251 (FloatType == float branch)
                    auto num = [&](const char* k, double d) { auto it = n.props.find(k); return (it != n.props.end() && it->second.isNumber()) ? it->second.asNumber() : d; };
                    const size_t size = (size_t) num("size", 512), channels = (size_t) num("channels", 1);
                    const uint32_t mask = SCOPE_RING - 1, r = n.scopeR, w = n.scopeW;
                    const size_t full = (w > r) ? (w - r) : (((uint32_t) SCOPE_RING - (r - w)) & mask);
                    if (!(full > size) || !n.ring) break;
                    if (full >= size) {        // ringBuffer.read(...) succeeds
                        // one contiguous [pos][L] slab per (tile, channel), wrapped reads in two pieces; then de-interleave
                        std::vector<float> slab((size_t) size * L);
                        std::vector<std::vector<float>> chans(channels, std::vector<float>(size));
                        const int t0 = b / L, t1 = (e + L - 1) / L;
                        for (int tile = t0; tile < t1 && cb; ++tile) {
                            std::vector<std::vector<float>> tileData(std::min(channels, (size_t) SCOPE_CHANNELS));
                            for (size_t ch = 0; ch < tileData.size(); ++ch) {
                                const float* base = n.ring + ((size_t) tile * SCOPE_CHANNELS + ch) * SCOPE_RING * L;
                                const size_t first = std::min(size, (size_t) SCOPE_RING - r);
                                if (!cuda(cudaMemcpy(slab.data(), base + (size_t) r * L, sizeof(float) * first * L, cudaMemcpyDeviceToHost), "read scope ring")) return rc::CudaError;
                                if (first < size && !cuda(cudaMemcpy(slab.data() + first * L, base, sizeof(float) * (size - first) * L, cudaMemcpyDeviceToHost), "read scope ring (wrap)")) return rc::CudaError;
                                tileData[ch] = slab;
                            }
                            for (int vl = 0; vl < L; ++vl) {
                                const int v = tile * L + vl;
                                if (v < b || v >= e) continue;
                                std::string js = "{\"data\": [";
                                for (size_t ch = 0; ch < channels; ++ch) {
                                    if (ch) js += ", ";
                                    std::vector<float> one(size, 0.0f);   // channels beyond the ring's 4 stay zero like a fresh Float32Array
                                    if (ch < tileData.size()) for (size_t i = 0; i < size; ++i) one[i] = tileData[ch][i * L + vl];
                                    appendFloatArray(js, one.data(), size);
                                }
                                js += "], \"source\": " + src + ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                                cb("scope", js.c_str(), g.v0 + v, user);
                            }
                        }
                        n.scopeR = (r + (uint32_t) size) & mask;
                    }
                } break;
                case NodeKind::Fft: {          // wasm/FFT.h:90-131: one windowed real FFT per poll once `size` samples are waiting
                    const size_t size = n.window.size();
                    const uint32_t mask = SCOPE_RING - 1, r = n.scopeR, w = n.scopeW;
                    const size_t full = (w > r) ? (w - r) : (((uint32_t) SCOPE_RING - (r - w)) & mask);
                    if (size == 0 || full < size || !n.ring) break;
                    std::vector<float> slab(size * (size_t) L), x(size), re(size / 2 + 1), im(size / 2 + 1);
                    const int t0 = b / L, t1 = (e + L - 1) / L;
                    for (int tile = t0; tile < t1 && cb; ++tile) {
                        const float* base = n.ring + (size_t) tile * SCOPE_RING * L;
                        const size_t first = std::min(size, (size_t) SCOPE_RING - r);
                        if (!cuda(cudaMemcpy(slab.data(), base + (size_t) r * L, sizeof(float) * first * L, cudaMemcpyDeviceToHost), "read fft ring")) return rc::CudaError;
                        if (first < size && !cuda(cudaMemcpy(slab.data() + first * L, base, sizeof(float) * (size - first) * L, cudaMemcpyDeviceToHost), "read fft ring (wrap)")) return rc::CudaError;
                        for (int vl = 0; vl < L; ++vl) {
                            const int v = tile * L + vl;
                            if (v < b || v >= e) continue;
                            for (size_t i = 0; i < size; ++i) x[i] = slab[i * L + vl] * n.window[i];
                            realFftFloatIO(x, re, im);
                            std::string js = "{\"data\": {\"imag\": ";
                            appendFloatArray(js, im.data(), im.size());
                            js += ", \"real\": ";
                            appendFloatArray(js, re.data(), re.size());
                            js += "}, \"source\": " + src + ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                            cb("fft", js.c_str(), g.v0 + v, user);
                        }
                    }
                    n.scopeR = (r + (uint32_t) size) & mask;
                } break;
                case NodeKind::Capture: {      // Capture.h:60-93
                    if (!n.ring || !readRows(n.stateRow, 5, rows)) return rc::CudaError;
                    const uint32_t cap = (uint32_t) (n.size - CAPTURE_SCRATCH), mask = cap - 1;
                    if (n.relay.size() < (size_t) g.nv) n.relay.resize((size_t) g.nv);
                    bool touched = false;
                    std::vector<float> col;
                    for (int v = b; v < e; ++v) {
                        const uint32_t w = rows[(size_t) 2 * g.Vpad + v], r = rows[(size_t) 3 * g.Vpad + v];
                        const uint32_t avail = (w > r) ? (w - r) : ((cap - (r - w)) & mask);
                        if (avail > 0) {
                            const float* base = n.ring + (size_t) (v / L) * (size_t) n.size * L + (v % L);
                            col.resize(avail);
                            const uint32_t first = std::min(avail, cap - r);
                            if (!cuda(cudaMemcpy2D(col.data(), sizeof(float), base + (size_t) r * L, sizeof(float) * L, sizeof(float), first, cudaMemcpyDeviceToHost), "read capture ring")) return rc::CudaError;
                            if (first < avail && !cuda(cudaMemcpy2D(col.data() + first, sizeof(float), base, sizeof(float) * L, sizeof(float), avail - first, cudaMemcpyDeviceToHost), "read capture ring (wrap)")) return rc::CudaError;
                            n.relay[v].insert(n.relay[v].end(), col.begin(), col.end());
                            rows[(size_t) 3 * g.Vpad + v] = (r + avail) & mask;
                            touched = true;
                        }
                        if (rows[(size_t) 4 * g.Vpad + v]) {       // relayReady.exchange(false)
                            rows[(size_t) 4 * g.Vpad + v] = 0;
                            touched = true;
                            if (cb) {
                                std::string js = "{\"data\": ";
                                appendFloatArray(js, n.relay[v].data(), n.relay[v].size());
                                js += ", \"source\": " + src + ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                                cb("capture", js.c_str(), g.v0 + v, user);
                            }
                            n.relay[v].clear();
                        }
                    }
                    if (touched && !cuda(cudaMemcpy(g.dRows + (size_t) (n.stateRow + 3) * g.Vpad, rows.data() + (size_t) 3 * g.Vpad, sizeof(uint32_t) * 2 * g.Vpad, cudaMemcpyHostToDevice), "write capture positions")) return rc::CudaError;
                } break;
                case NodeKind::Metro: {        // wasm/Metro.h:58-66
                    if (!n.metroFlag) break;
                    if (b == 0 && e == g.nv) n.metroFlag = false;      // the flag is one per group: cleared when every voice was served
                    for (int v = b; v < e && cb; ++v) {
                        std::string js = "{\"source\": " + src + ", \"voice\": " + std::to_string(g.v0 + v) + "}";
                        cb("metro", js.c_str(), g.v0 + v, user);
                    }
                } break;
                default: break;
                }
            }
        }
    }
    return rc::Ok;
}

} // namespace eb
