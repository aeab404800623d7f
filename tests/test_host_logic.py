"""CPU: host-side logic of the product (instruction interpreter, return codes, gc, graph compilation) exercised
through a plan-only runtime (device=-1) and compared, where it applies, with the compiled reference's behaviour."""
import json

import numpy as np
import pytest

from elementary_b200 import Runtime, el, graphs
from oracle import oracle as orc
from cases import CASES

SR, BS = 48000.0, 512


def plan(n_voices=8, **opts):
    return Runtime(SR, BS, n_voices, device=-1, **opts)


BAD_BATCHES = [
    ([[0, 1, "nope"]], 1),
    ([[0, 2, "sin"], [0, 2, "sin"]], 3),
    ([[2, 5, 6, 0]], 2),
    ([[0, 7, "sin"], [2, 7, 99, 0]], 2),
    ([[3, 1234, "value", 1]], 2),
    ([[0, 3, "const"], [3, 3, "value", "x"]], 5),
    ([[0, 4, "table"], [3, 4, "path", "missing"]], 6),
    ([[0, 5, "svf"], [3, 5, "mode", 3]], 5),
    ([[0, 6, "root"], [3, 6, "active", 1]], 5),
    ([[4, [424242]]], 2),
    ([5], 8),
    ([["x", 1, "sin"]], 8),
    ([[0, "id", "sin"]], 8),
    ([[0, 8, "delay"], [3, 8, "size", "big"]], 5),
    # sequencing / control / analysis nodes (Core.h:411-466, Seq2.h:39-84, SparSeq.h:40-124, SparSeq2.h:20-56, Core.h:345-361,
    # wasm/Metro.h:20-37, Analyzers.h:155-179, wasm/FFT.h:32-71): property validation must return the reference's codes
    ([[0, 9, "seq"], [3, 9, "hold", 1]], 5),
    ([[0, 9, "seq"], [3, 9, "loop", "yes"]], 5),
    ([[0, 9, "seq"], [3, 9, "offset", -1]], 6),
    ([[0, 9, "seq"], [3, 9, "offset", "x"]], 5),
    ([[0, 9, "seq"], [3, 9, "seq", 3]], 5),
    ([[0, 9, "seq2"], [3, 9, "seq", "abc"]], 5),
    ([[0, 9, "seq2"], [3, 9, "offset", -0.5]], 6),
    ([[0, 9, "sparseq"], [3, 9, "loop", 4]], 5),
    ([[0, 9, "sparseq"], [3, 9, "follow", 1]], 5),
    ([[0, 9, "sparseq"], [3, 9, "interpolate", True]], 5),
    ([[0, 9, "sparseq"], [3, 9, "tickInterval", -1]], 6),
    ([[0, 9, "sparseq"], [3, 9, "seq", {}]], 5),
    ([[0, 9, "sparseq2"], [3, 9, "seq", 1]], 5),
    ([[0, 9, "sparseq2"], [3, 9, "interpolate", "1"]], 5),
    ([[0, 9, "once"], [3, 9, "arm", 1]], 5),
    ([[0, 9, "metro"], [3, 9, "interval", 0]], 6),
    ([[0, 9, "metro"], [3, 9, "interval", "fast"]], 5),
    ([[0, 9, "scope"], [3, 9, "size", 128]], 6),
    ([[0, 9, "scope"], [3, 9, "channels", 5]], 6),
    ([[0, 9, "scope"], [3, 9, "name", 7]], 5),
    ([[0, 9, "fft"], [3, 9, "size", 300]], 6),
    ([[0, 9, "fft"], [3, 9, "size", 16384]], 6),
    ([[0, 9, "fft"], [3, 9, "size", "1k"]], 5),
    ([[0, 9, "fft"], [3, 9, "name", 1]], 5),
    ([[0, 9, "sampleseq"]], 1),         # sample playback stays out of scope: unknown type here (the reference knows it)
]
OUT_OF_SCOPE_TYPES = {"sampleseq"}


@pytest.mark.parametrize("batch,code", BAD_BATCHES)
def test_error_codes_match_reference(batch, code):
    assert plan().apply_instructions(batch) == code
    if orc.ref_available() and not any(len(i) > 2 and i[2] in OUT_OF_SCOPE_TYPES for i in batch if isinstance(i, list)):
        assert orc.RefRuntime(SR, BS).apply(batch) == code


def test_malformed_json_is_code_8_not_an_exception():
    rt = plan()
    assert rt.apply_instructions("[[0, 1") == 8
    assert rt.apply_instructions("{}") == 8
    assert rt.apply_instructions("") == 8


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_every_case_compiles(case):
    rt = plan(4)
    for k, v in (case["resources"] or {}).items():
        assert rt.add_shared_resource(k, v)
    assert rt.apply_instructions(case["batch"]) == 0, rt.last_error()
    d = rt.describe()
    assert d["groups"][0]["roots"] >= 1 and d["groups"][0]["code_words"] > 6


def test_subsynth32_program_shape():
    rt = plan(4096)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    g = rt.describe()["groups"][0]
    assert g["nodes"] == 32                      # 31 + root
    assert g["state_rows"] == 10                 # 3 phasor + 2 svf doubles (4 rows) + delay index, padded even
    assert g["slots"] <= 6                       # liveness allocation: intermediates, not 32 block buffers
    assert g["ops"] == 11                        # 19 executable nodes fused into 11 ops (chains of element-wise math)
    assert g["params"] == 13                     # the 13 consts of SURVEY.md §8d, staged per block in shared memory
    assert g["tile_width"] == 2                  # 4096 voices -> 2 voices per warp (2048 warps) to fill the machine
    rt = plan(1 << 17)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    assert rt.describe()["groups"][0]["tile_width"] == 32


def test_additive_voice_is_scheduled_into_a_handful_of_slots():
    rt = plan(64)
    assert rt.apply_instructions(graphs.additive64()) == 0
    g = rt.describe()["groups"][0]
    assert g["nodes"] == 387 and g["state_rows"] == 64 and g["params"] == 129
    assert g["slots"] <= 11                      # Sethi-Ullman order + incremental fold (+7 for runs of 8 fused phasors): not 65
    big = plan(1 << 17)
    assert big.apply_instructions(graphs.additive64()) == 0
    assert big.describe()["groups"][0]["slots"] <= 4   # full-width tiles do not group phasors
    rt = plan(64, fuse_chains=0)
    assert rt.apply_instructions(graphs.additive64()) == 0
    assert rt.describe()["groups"][0]["slots"] == 65


def test_shared_resources_are_insert_only_and_prunable():
    rt = plan()
    assert rt.add_shared_resource("one", np.arange(5, dtype=np.float32))
    assert not rt.add_shared_resource("one", np.arange(3, dtype=np.float32))     # SharedResource.h:44-46
    assert rt.add_shared_resource("two", np.arange(5, dtype=np.float32))
    assert sorted(rt.get_shared_resource_map_keys()) == ["one", "two"]
    assert rt.apply_instructions(el.render(el.table({"path": "one"}, el.in_(0)))) == 0
    rt.prune_shared_resources()                                                    # vfs.test.js:57-90
    assert rt.get_shared_resource_map_keys() == ["one"]


def test_gc_matches_reference():
    """gc.test.js semantics: nodes of a replaced graph are collectable only once the new sequence is active."""
    r = el.Renderer()
    a = r.render(el.cycle(220.0))
    b = r.render(el.mul(0.5, el.saw(330.0)))
    rt = plan(2)
    assert rt.apply_instructions(a) == 0 and rt.gc() == []
    assert rt.apply_instructions(b) == 0
    assert rt.gc() == []          # old sequence is still the active one until the next process()
    if orc.ref_available():
        o = orc.RefRuntime(SR, BS)
        assert o.apply(a) == 0 and o.gc() == []


def test_value_only_batches_address_sub_ranges_but_structure_does_not_split_live_groups():
    rt = plan(64)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    assert rt.apply_instructions(graphs.subsynth32_voice_props(5), voices=(5, 6)) == 0
    assert len(rt.describe()["groups"]) == 1
    assert rt.apply_instructions([[0, 99, "sin"]], voices=(5, 6)) == 0          # structural batch: the live group is cut in 3
    d = rt.describe()["groups"]
    assert [(g["v0"], g["nv"]) for g in d] == [(0, 5), (5, 1), (6, 58)] and [g["nodes"] for g in d] == [32, 33, 32]
    ida, _ = graphs.subsynth32_param_ids()
    assert rt.set_property_per_voice(ida, "value", np.linspace(50, 500, 64)) == 0
    assert rt.set_property_per_voice(12345, "value", np.zeros(4)) == 2


def test_fresh_runtime_can_host_different_graphs_per_voice_range():
    rt = plan(10)
    for i in range(5):
        assert rt.apply_instructions(graphs.random_graph(i, 24), voices=(2 * i, 2 * i + 2)) == 0, rt.last_error()
    groups = rt.describe()["groups"]
    assert [g["v0"] for g in groups] == [0, 2, 4, 6, 8] and all(g["nv"] == 2 for g in groups)


def test_el_renderer_is_incremental():
    r = el.Renderer()
    a = r.render(el.cycle(el.const(220.0, key="f")))
    b = r.render(el.cycle(el.const(330.0, key="f")))
    assert sum(1 for i in a if i[0] == 0) == 6           # root, sin, mul, const 2pi, phasor, const f (hashing.test.js.snap)
    assert [i for i in b if i[0] == 0] == []             # nothing new is created ...
    assert [i[2:] for i in b if i[0] == 3] == [["value", 330.0]]   # ... only the keyed const changes
    assert b[-2][0] == 4 and b[-1] == [5]


def test_live_group_split_with_sequencer_and_analysis_nodes_runs_the_migration_code():
    """Plan-only runtimes keep "device" buffers in host memory, so cutting a compiled voice group executes the whole
    migration path (row slicing, ring / scratch cloning of scope, capture and fft, sequence-data sharing, recompilation
    of the moved half) on the CPU; the rendering side of the same scenario is covered on the GPU for the older node kinds."""
    trig = el.train(1500.0)
    seq = el.seq({"seq": [0.2, 0.4, 0.6], "hold": True, "key": "sq"}, trig, 0.0)
    sp = el.sparseq({"seq": [{"value": 1, "tickTime": 0}, {"value": 2, "tickTime": 3}], "loop": [0, 4], "key": "sp"}, trig, 0.0)
    core = el.mul(el.add(seq, sp), el.cycle(el.const(220.0, key="f")))
    g1 = el.fft({"name": "ff"}, el.capture({"name": "cap"}, trig, el.scope({"name": "sc"}, el.meter({"name": "m"}, el.once({"arm": True}, core)))))
    g2 = el.tanh(el.mul(2.0, g1))
    rg = el.Renderer()
    a, b = rg.render(g1), rg.render(g2)
    rt = plan(16, tile_width=4)
    assert rt.apply_instructions(a) == 0, rt.last_error()
    d0 = rt.describe()
    assert len(d0["groups"]) == 1 and d0["groups"][0]["nv"] == 16
    assert rt.apply_instructions(b, voices=(6, 16)) == 7                 # not on a tile boundary (Types.h:58 InvariantViolation)
    assert rt.apply_instructions(b, voices=(8, 16)) == 0, rt.last_error()
    d1 = rt.describe()
    assert [(g["v0"], g["nv"]) for g in d1["groups"]] == [(0, 8), (8, 8)]
    assert d1["groups"][1]["ops"] > d1["groups"][0]["ops"]               # the moved half runs the extended graph
    assert rt.apply_instructions([[3, seq.id(), "seq", [9.0, 8.0]]], voices=(8, 16)) == 0    # and can be re-programmed on its own
    assert rt.gc(0) == [] and isinstance(rt.gc(8), list)


def test_per_program_specialisation_compiles_without_a_gpu():
    """EXPERIMENTAL path (option "specialize", off by default): NVRTC compiles K1 against the render program of a
    voice group as a compile-time constant.  NVRTC is a pure compiler, so the compile step is checked here; launching the result
    is left to the first GPU session that has minutes for it."""
    rt = plan(4096)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    words = rt.program_words(0)
    assert len(words) == rt.describe()["groups"][0]["code_words"] and words[0] & 0xFF == 1      # OP_SEG opens the program
    n, log = rt.specialize_dry_run(0)
    assert n > 50_000, log
    rt2 = plan(2)
    assert rt2.add_shared_resource("ir", np.ones(700, dtype=np.float32))
    assert rt2.apply_instructions(graphs.convolve_channel("ir")) == 0
    n2, log2 = rt2.specialize_dry_run(0)
    assert n2 == -1 and "multi-stage" in log2


@pytest.mark.parametrize("name", ["delay_mod", "table", "taps", "two_roots", "sparseq_interp_loop_offset", "seq_hold",
                                  "scope_passthrough", "capture_passthrough", "bleptriangle", "random_graph_1"])
def test_specialised_kernel_compiles_for_every_kind_of_op(name):
    """The specialised path reuses the interpreter's per-op text (render_ops.inc); this keeps every family of op bodies
    compilable with compile-time operands (all 113 cases x 2 tile widths were swept offline: profiles/r01_r_*)."""
    case = next(c for c in CASES if c["name"] == name)
    rt = plan(64, tile_width=32 if len(name) % 2 else 2)
    for k, v in (case["resources"] or {}).items():
        assert rt.add_shared_resource(k, v)
    assert rt.apply_instructions(case["batch"]) == 0
    n, log = rt.specialize_dry_run(0)
    assert n > 0, log


def test_specialisation_runs_in_the_background_or_on_demand():
    """option "specialize": 2 waits for NVRTC at COMMIT, 1 compiles on a worker thread while the interpreter would keep
    serving; describe() shows the state (0 compiling, 1 cubin ready, 2 loaded on a GPU, -1 failed)."""
    import time
    rt = plan(256, specialize=2)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    g = rt.describe()["groups"][0]
    assert g["spec_state"] == 1 and g["spec_cubin_bytes"] > 50_000
    rt = plan(256, specialize=1)
    assert rt.apply_instructions(graphs.subsynth32()) == 0      # returns at once
    for _ in range(600):
        g = rt.describe()["groups"][0]
        if g["spec_state"] != 0:
            break
        time.sleep(0.05)
    assert g["spec_state"] == 1 and g["spec_cubin_bytes"] > 50_000
    rt = plan(64, specialize=2, specialize_max_words=100)        # too long for the limit: the interpreter stays
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    assert "spec_state" not in rt.describe()["groups"][0]


# ---- Runtime::registerNodeType / snapshot (Runtime.h:105-110) at the boundary -------------------------------------------------
SOFTCLIP = "const float x = in[0]; return x / (1.0f + fabsf(x));"
LEAKY = "s[0] = in[0] + in[1] * s[0]; return s[0];"


def test_register_node_type_return_codes_and_compile():
    from elementary_b200 import el
    rt = plan(8)
    assert rt.has_node_type("svf") and not rt.has_node_type("b200test.softclip")
    assert rt.apply_instructions([[0, 1, "b200test.softclip"]]) == 1            # UnknownNodeType before registration (Runtime.h:304-305)
    assert rt.register_node_type("svf", 1, 0, SOFTCLIP) == 4                     # NodeTypeAlreadyExists for a builtin (Runtime.h:482-483)
    assert rt.register_node_type("b200test.softclip", 1, 0, SOFTCLIP) == 0
    assert rt.register_node_type("b200test.softclip", 1, 0, SOFTCLIP) == 4
    assert rt.register_node_type("b200test.leaky", 2, 1, LEAKY) == 0
    g = el.create_node("b200test.leaky", {}, [el.create_node("b200test.softclip", {}, [el.cycle(220.0)]), el.const(0.5, key="g")])
    assert rt.apply_instructions(el.render(g)) == 0, rt.last_error()             # COMMIT compiles the bodies with NVRTC (no GPU needed)
    d = rt.describe()["groups"][0]
    assert d["spec_state"] == 1 and d["spec_cubin_bytes"] > 10000
    # a body that does not compile fails the COMMIT with the compiler's message
    rt2 = plan(8)
    assert rt2.register_node_type("broken", 1, 0, "return in[0] +;") == 0
    assert rt2.apply_instructions(el.render(el.create_node("broken", {}, [el.cycle(1.0)]))) == 7
    assert "error" in rt2.last_error()


def test_snapshot_lists_nodes_and_props_like_the_reference():
    from elementary_b200 import el
    rt = plan(4)
    g = el.svf({"mode": "highpass"}, el.const(800.0, key="fc"), 1.0, el.cycle(220.0))
    assert rt.apply_instructions(el.render(g)) == 0
    snap = rt.snapshot()
    fc = el.const(0, key="fc").id()
    key = "0x%08x" % (fc & 0xFFFFFFFF)
    assert key in snap and snap[key]["value"] == 800.0
    assert any(v.get("mode") == "highpass" for v in snap.values())
    assert all(k.startswith("0x") and len(k) == 10 for k in snap)


# ---- ingestion at scale: binary batch + property table ---------------------------------------------------
def test_binary_batch_is_the_same_instruction_stream():
    from elementary_b200 import el, graphs
    import numpy as np
    for batch in (graphs.subsynth32(), graphs.random_graph(7, 48), el.render(el.seq({"seq": [1, 2, 3.5], "hold": True, "loop": False}, el.train(5.0))),
                  el.render(el.sparseq({"seq": [{"value": 1, "tickTime": 0}, {"value": 2, "tickTime": 4}], "loop": [0, 8]}, el.train(9.0)))):
        a, b = plan(8), plan(8)
        assert a.apply_instructions(batch) == 0
        blob = el.encode_binary(batch)
        assert b.apply_binary(blob) == 0, b.last_error()
        def masked(w):                       # the device-pointer words of the op headers differ between two runtimes
            w = w.copy(); pc = 0
            while pc + 8 <= len(w):
                n = 8 + ((int(w[pc]) >> 8) & 0xFF); w[pc + 4:pc + 6] = 0
                if (int(w[pc]) & 0xFF) == 0: break
                pc += n
            return w[:pc + 8]
        assert np.array_equal(masked(a.program_words(0)), masked(b.program_words(0)))          # same render program, word for word
        assert a.snapshot() == b.snapshot()
    rt = plan(4)
    assert rt.apply_binary(b"nope") == 8                                       # InvalidInstructionFormat
    assert rt.apply_binary(el.encode_binary([[0, 1, "no_such_type"]])) == 1     # same return codes as the text form
    assert rt.apply_binary(el.encode_binary(graphs.subsynth32())[:-3]) == 8    # truncated


def test_binary_batch_is_smaller_and_faster_to_ingest_than_json():
    import json, time
    from elementary_b200 import el, graphs
    batch = graphs.additive64(110.0, 64)
    text, blob = json.dumps(batch).encode(), el.encode_binary(batch)
    assert len(blob) < len(text)
    rt = plan(1)
    t0 = time.perf_counter(); assert rt.apply_binary(blob) == 0; t_bin = time.perf_counter() - t0
    rt = plan(1)
    t0 = time.perf_counter(); assert rt.apply_instructions(text) == 0; t_txt = time.perf_counter() - t0
    print(f"additive64: json {len(text)} B {t_txt * 1e3:.2f} ms, binary {len(blob)} B {t_bin * 1e3:.2f} ms")


def test_const_table_return_codes():
    import numpy as np
    from elementary_b200 import el, graphs
    rt = plan(64)
    assert rt.apply_instructions(graphs.subsynth32()) == 0
    ida, idb = graphs.subsynth32_param_ids()
    tab = np.stack([55.0 * np.arange(1, 65), 55.0 * 1.007 * np.arange(1, 65)]).astype(np.float32)
    assert rt.set_const_table([ida, idb], tab) == 0
    assert rt.set_const_table([ida, 123456], tab) == 2                         # NodeNotFound
    assert rt.snapshot()["0x%08x" % (ida & 0xFFFFFFFF)]["value"] == float(tab[0, -1])


def Runtime_with_graphs(n):
    from elementary_b200 import graphs
    rt = Runtime(SR, BS, n, device=-1)
    for i in range(n):
        assert rt.apply_instructions(graphs.random_graph(300 + i, 64), voices=(i, i + 1)) == 0, rt.last_error()
    return rt


def test_pipeline_cut_of_one_voice_groups_plan_only():
    """The host side of the warp pipeline (render_groups_pipe_kernel): one-voice groups in a many-groups engine get
    their program cut into pipeline stages; single-group engines, wide tiles and pipeline_stages = 0 do not; the program words of a cut
    program are one [SEG ... END] section per stage whose ring slots are only ever written in one stage and read in later ones."""
    from elementary_b200 import graphs
    n = 6
    rt = Runtime(SR, BS, n, device=-1, pipeline_stages=3)
    for i in range(n):
        assert rt.apply_instructions(graphs.random_graph(300 + i, 64), voices=(i, i + 1)) == 0, rt.last_error()
    groups = rt.describe()["groups"]
    assert len(groups) == n and all(g["pipeline_stages"] == 3 for g in groups), groups
    assert all(g["pipeline_stages"] == 4 for g in Runtime_with_graphs(n).describe()["groups"])       # the default (option pipeline_stages)
    off = Runtime(SR, BS, n, device=-1, pipeline_stages=0)
    for i in range(n):
        assert off.apply_instructions(graphs.random_graph(300 + i, 64), voices=(i, i + 1)) == 0
    assert all(g["pipeline_stages"] == 1 for g in off.describe()["groups"])
    assert all(a["ops"] == b["ops"] and a["state_rows"] == b["state_rows"] for a, b in zip(groups, off.describe()["groups"]))
    single = Runtime(SR, BS, 4, device=-1, pipeline_stages=3)
    assert single.apply_instructions(graphs.random_graph(300, 64)) == 0
    assert single.describe()["groups"][0]["pipeline_stages"] == 1
    # structure of the cut program: sections [SEG][ops][END], outputs of a section never collide with another section's private slots
    words = rt.program_words(0)
    HDR = 8
    sections, pc, cur = [], 0, None
    while pc < len(words):
        w0 = words[pc]
        opcode, nwords, out_slot = w0 & 0xFF, (w0 >> 8) & 0xFF, (w0 >> 16) & 0xFF
        if opcode == 1:                       # OP_SEG opens a section
            cur = {"outs": [], "skip": words[pc + 3], "start": pc}
        elif opcode == 0:                     # OP_END closes it
            if cur is None:
                break
            assert pc - cur["start"] - HDR == cur["skip"], "the SEG header skips exactly its own section"
            sections.append(cur)
            cur = None
        else:
            cur["outs"].append(out_slot)
        pc += HDR + nwords
    assert len(sections) == 3 and all(s["outs"] for s in sections)
    for i in range(3):
        for j in range(i + 1, 3):
            assert not (set(sections[i]["outs"]) & set(sections[j]["outs"])), "stages never share an output slot"


def test_tile_width_choice_dense_mid_range_with_specialised_kernels():
    """Engine::chooseTileWidth: ~2048 warps by default; from 8192 voices up, WITH specialised kernels (64-register narrow geometries),
    ~4096 warps (profiles/r02_aa_midrange_voices_ab.txt); an explicit target_tiles or tile_width always wins."""
    from elementary_b200 import graphs

    def tw(nv, **opts):
        rt = Runtime(SR, BS, nv, device=-1, **opts)
        assert rt.apply_instructions(graphs.plumbing() if hasattr(graphs, "plumbing") else graphs.subsynth32()) == 0, rt.last_error()
        return rt.describe()["groups"][0]["tile_width"]

    assert [tw(v) for v in (45, 4096, 8192, 65536, 131072)] == [1, 2, 4, 32, 32]
    assert [tw(v, specialize=1) for v in (4096, 8192, 16384, 32768, 65536, 131072)] == [2, 2, 4, 8, 16, 32]
    assert tw(65536, specialize=1, target_tiles=2048) == 32 and tw(65536, specialize=1, tile_width=8) == 8


def test_bench_arms_describe_the_same_workload():
    """bench.py: `config` is one dict for both arms (--impl b200 and --impl reference) — the driver compares them."""
    import importlib, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    bench = importlib.import_module("bench")
    a, b = bench.bench_config(4096, 1), bench.bench_config(4096, 1)
    assert a == b and a["voices_total"] == 4096 and "SUBSYNTH32" in a["workload"] and "l2" in a
    assert bench.bench_config(4096, 8)["voices_total"] == 32768
    assert "Msamples/s" in bench.METRIC


def test_process_exit_during_a_background_compile_is_clean():
    """A process that ends while the compile-queue thread is inside NVRTC must exit with code 0 (spec_host.cpp waitForCompilerAtExit:
    the handler registered behind NVRTC's statics waits for the compile in flight).  It used to die at exit, intermittently, with a
    segmentation fault or "realloc(): invalid pointer"."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from elementary_b200 import Runtime, graphs\n"
            "rt = Runtime(48000.0, 512, 64, device=-1, specialize=1)\n"
            "assert rt.apply_instructions(graphs.subsynth32()) == 0\n"
            "print('queued', flush=True)\n" % root)
    for _ in range(3):
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
        assert p.returncode == 0 and "queued" in p.stdout, (p.returncode, p.stderr[-500:])
