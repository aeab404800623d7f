// thread_stress.cpp — the boundary's two-thread contract (include/elem_b200.h "threading"; reference: Runtime.h:133,204,277-285):
// ONE control thread (applyInstructions / setProperty / gc / events / describe) and ONE render thread (process) running
// concurrently on the same runtime.  Built twice by tests/test_threading_cpu.py: with -fsanitize=thread against a TSAN build of the
// host side of the library (plan-only runtime, option plan_dry_run: every host-side step of a block without a GPU), and it is
// also what the GPU stress test drives through the real render path.
//
// usage: thread_stress <device: -1 plan-only | 0..> <seconds> <graphA.json> <graphB.json>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "elem_b200.h"

static std::string slurp(const char* path) {
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s device seconds graphA.json graphB.json\n", argv[0]); return 2; }
    const int device = std::atoi(argv[1]);
    const double seconds = std::atof(argv[2]);
    const std::string ga = slurp(argv[3]), gb = slurp(argv[4]);
    const int voices = 64, bs = 512;
    elem_b200_runtime* rt = elem_b200_create(48000.0, bs, voices, device);
    if (!rt) { std::fprintf(stderr, "create failed: %s\n", elem_b200_last_error(nullptr)); return 3; }
    if (device < 0) elem_b200_set_option(rt, "plan_dry_run", 1.0);
    if (elem_b200_apply_instructions(rt, 0, -1, ga.c_str(), ga.size()) != 0) { std::fprintf(stderr, "apply A failed: %s\n", elem_b200_last_error(rt)); return 4; }

    std::atomic<bool> stop{false};
    std::atomic<long> blocks{0}, edits{0}, failures{0};

    std::thread render([&] {
        std::vector<float> out(bs);
        float* outs[1] = {out.data()};
        while (!stop.load(std::memory_order_relaxed)) {
            int rc;
            if (device < 0) rc = elem_b200_enqueue_block(rt, 0, 1, bs, 4);          // host side of a block only (no GPU here)
            else rc = elem_b200_process(rt, nullptr, 0, outs, 1, bs, nullptr);
            if (rc != 0) { failures++; std::fprintf(stderr, "render rc=%d: %s\n", rc, elem_b200_last_error(rt)); break; }
            blocks++;
        }
    });

    std::thread control([&] {
        const auto t0 = std::chrono::steady_clock::now();
        char buf[1 << 16];
        int32_t ids[4096];
        long n = 0;
        while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
            const std::string& g = (n & 1) ? ga : gb;              // re-rendering alternates two graphs: cross-fades, new nodes, gc
            int rc = elem_b200_apply_instructions(rt, 0, -1, g.c_str(), g.size());
            if (rc != 0 && rc != 3) { failures++; std::fprintf(stderr, "apply rc=%d: %s\n", rc, elem_b200_last_error(rt)); break; }
            // a structural edit of half of the voices: cuts the voice group while it renders (first time), then edits that half
            const char* half = "[[0, 777001, \"const\"], [3, 777001, \"value\", 0.5]]";
            rc = elem_b200_apply_instructions(rt, voices / 2, voices, half, std::char_traits<char>::length(half));
            if (rc != 0 && rc != 3) { failures++; std::fprintf(stderr, "half apply rc=%d\n", rc); break; }
            elem_b200_describe(rt, buf, sizeof(buf));
            elem_b200_gc(rt, 0, ids, 4096);
            elem_b200_gc(rt, voices - 1, ids, 4096);
            elem_b200_process_queued_events(rt, nullptr, nullptr);
            (void) elem_b200_kernel_launches(rt);
            (void) elem_b200_current_time(rt);
            edits++;
            ++n;
            std::this_thread::sleep_for(std::chrono::microseconds(200));
        }
        stop = true;
    });

    control.join();
    render.join();
    std::printf("{\"blocks\": %ld, \"edits\": %ld, \"failures\": %ld}\n", blocks.load(), edits.load(), failures.load());
    elem_b200_destroy(rt);
    return failures.load() == 0 && blocks.load() > 0 && edits.load() > 0 ? 0 : 1;
}
