// elem/Runtime.h — SOURCE-COMPATIBLE drop-in for the reference's runtime/elem/Runtime.h, backed by libelem_b200.so.
//
// Put this directory BEFORE the reference's `runtime/` on the include path: `#include <elem/Runtime.h>` then finds this file,
// every other `<elem/...>` header (Value.h, JSON.h, Types.h, SharedResource.h, GraphNode.h) still resolves to the reference tree.
// The reference's own callers — cli/Benchmark.cpp, cli/Realtime.cpp, wasm/Main.cpp — compile UNCHANGED against it and run their
// graphs on a B200 (oracle/Makefile builds cli/Benchmark.cpp + cli/BenchmarkMain.cpp this way: oracle/_ref/elembench_b200).
//
// elem::Runtime<FloatType> here has the public surface of Runtime.h:44-110, method for method, over the C ABI of
// include/elem_b200.h.  FloatType = float is the engine's arithmetic (north_star); Runtime<double> converts the audio buffers at
// the boundary (the render itself stays float: the wasm hosts' double runtime is out of scope).
// New knobs that have no counterpart in the reference come from the environment so that unmodified callers can use them:
//   ELEM_B200_VOICES (default 1)   ELEM_B200_DEVICE (default 0)   ELEM_B200_SPECIALIZE (0/1/2, default 1: NVRTC in the background)
#pragma once

#include <cstdlib>
#include <functional>
#include <memory>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

#include <elem/GraphNode.h>
#include <elem/JSON.h>
#include <elem/SharedResource.h>
#include <elem/Types.h>
#include <elem/Value.h>

#include <elem_b200.h>

namespace elem
{

    template <typename FloatType>
    class Runtime
    {
    public:
        // Runtime.h:44,158-166
        Runtime(double sampleRate, int blockSize)
            : blockSize(blockSize)
        {
            numVoices = envInt("ELEM_B200_VOICES", 1);
            h = elem_b200_create(sampleRate, blockSize, numVoices, envInt("ELEM_B200_DEVICE", 0));
            if (h == nullptr)   // no CUDA device: there is no CPU fallback
                throw std::runtime_error(std::string("elem_b200_create: ") + elem_b200_last_error(nullptr));
            elem_b200_set_option(h, "specialize", (double) envInt("ELEM_B200_SPECIALIZE", 1));
        }

        ~Runtime() { elem_b200_destroy(h); }
        Runtime(Runtime const&) = delete;
        Runtime& operator=(Runtime const&) = delete;

        // Runtime.h:48,170-218 — same batch, same return codes (Types.h:51-60); the batch travels as the JSON text the
        // reference's own serializer writes (JSON.h:245-250)
        int applyInstructions(js::Array const& batch)
        {
            auto const text = js::serialize(js::Value(batch));
            return elem_b200_apply_instructions(h, 0, -1, text.data(), text.size());
        }

        // Runtime.h:51-57,275-290 — planar host buffers; with more than one voice the outputs carry the mix bus
        void process(const FloatType** inputChannelData, size_t numInputChannels, FloatType** outputChannelData,
                     size_t numOutputChannels, size_t numSamples, void* userData = nullptr)
        {
            if constexpr (std::is_same<FloatType, float>::value) {
                elem_b200_process(h, inputChannelData, numInputChannels, outputChannelData, numOutputChannels, numSamples, userData);
            } else {
                inF.resize(numInputChannels * numSamples);
                outF.resize(numOutputChannels * numSamples);
                inP.resize(numInputChannels);
                outP.resize(numOutputChannels);
                for (size_t c = 0; c < numInputChannels; ++c) {
                    inP[c] = inF.data() + c * numSamples;
                    for (size_t i = 0; i < numSamples; ++i) inF[c * numSamples + i] = static_cast<float>(inputChannelData[c][i]);
                }
                for (size_t c = 0; c < numOutputChannels; ++c) outP[c] = outF.data() + c * numSamples;
                elem_b200_process(h, inP.data(), numInputChannels, outP.data(), numOutputChannels, numSamples, userData);
                for (size_t c = 0; c < numOutputChannels; ++c)
                    for (size_t i = 0; i < numSamples; ++i) outputChannelData[c][i] = static_cast<FloatType>(outF[c * numSamples + i]);
            }
        }

        // Runtime.h:64,438-446 — the event object is the reference's plus "voice"
        void processQueuedEvents(std::function<void(std::string const&, js::Value)>&& evtCallback)
        {
            elem_b200_process_queued_events(h, [](const char* type, const char* json, void* user) {
                (*static_cast<std::function<void(std::string const&, js::Value)>*>(user))(std::string(type), js::parseJSON(std::string(json)));
            }, &evtCallback);
        }

        void reset() { elem_b200_reset(h); }                                            // Runtime.h:70,449-458

        std::set<NodeId> gc()                                                            // Runtime.h:76,221-272
        {
            std::vector<int32_t> ids(65536);
            int const n = elem_b200_gc(h, 0, ids.data(), ids.size());
            return std::set<NodeId>(ids.begin(), ids.begin() + std::min<size_t>((size_t) n, ids.size()));
        }

        bool addSharedResource(std::string const& name, std::unique_ptr<SharedResource> resource)   // Runtime.h:83,462-465
        {
            std::vector<const float*> ch(resource->numChannels());
            for (size_t i = 0; i < ch.size(); ++i) ch[i] = resource->getChannelData(i).data();
            bool const ok = elem_b200_add_shared_resource(h, name.c_str(), ch.data(), ch.size(), resource->numSamples()) == 1;
            if (ok) sharedResourceMap.add(name, std::move(resource));                   // host mirror: keeps getSharedResourceMapKeys() exact
            return ok;
        }

        void pruneSharedResources()                                                      // Runtime.h:89,467-471
        {
            elem_b200_prune_shared_resources(h);
            std::vector<char> buf(1 << 16);
            elem_b200_list_shared_resources(h, buf.data(), buf.size());
            std::set<std::string> alive;
            for (char* p = buf.data(); *p;) { char* e = p; while (*e && *e != '\n') ++e; alive.insert(std::string(p, e)); p = *e ? e + 1 : e; }
            SharedResourceMap kept;
            for (auto const& k : sharedResourceMap.keys()) if (alive.count(k)) kept.add(k, sharedResourceMap.get(k));
            sharedResourceMap = std::move(kept);
        }

        SharedResourceMap::KeyViewType getSharedResourceMapKeys() { return sharedResourceMap.keys(); }   // Runtime.h:94,473-477

        // Runtime.h:105-106,480-487.  The builtin names are compiled into the fused kernel, so registering one of them again gives
        // NodeTypeAlreadyExists exactly like the reference.  A NEW type cannot be a host GraphNode (its process() would have to run
        // on the CPU inside the GPU's block — there is no CPU fallback); a new type is registered as DEVICE code instead:
        // elem_b200_register_node_type (include/elem_b200.h), also reachable through registerDeviceNodeType below.
        using NodeFactoryFn = std::function<std::shared_ptr<GraphNode<FloatType>>(NodeId const id, double sampleRate, int const blockSize)>;
        int registerNodeType(std::string const& type, NodeFactoryFn&&)
        {
            if (elem_b200_has_node_type(h, type.c_str())) return ReturnCode::NodeTypeAlreadyExists();
            return ReturnCode::InvariantViolation();
        }
        int registerDeviceNodeType(std::string const& type, int numInputs, int numStateFloats, std::string const& cudaBody)
        {
            return elem_b200_register_node_type(h, type.c_str(), numInputs, numStateFloats, cudaBody.c_str());
        }

        js::Object snapshot()                                                            // Runtime.h:110,490-499
        {
            int const n = elem_b200_snapshot(h, 0, nullptr, 0);
            std::string buf((size_t) n + 1, '\0');
            elem_b200_snapshot(h, 0, &buf[0], buf.size());
            buf.resize(std::strlen(buf.c_str()));
            auto v = js::parseJSON(buf);
            return v.isObject() ? v.getObject() : js::Object();
        }

        elem_b200_runtime* handle() { return h; }                                        // the C ABI underneath (voice-axis calls)

    private:
        static int envInt(const char* name, int dflt) { const char* s = std::getenv(name); return s ? std::atoi(s) : dflt; }

        elem_b200_runtime* h = nullptr;
        SharedResourceMap sharedResourceMap;
        int blockSize, numVoices = 1;
        std::vector<float> inF, outF;
        std::vector<const float*> inP;
        std::vector<float*> outP;
    };

} // namespace elem
