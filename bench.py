#!/usr/bin/env python
"""bench.py — headline measurement of the hot path (BASELINE.json configs[1]).

Workload: 4096 SUBSYNTH32 voices (32-node subtractive-synth graph, per-voice f0 = 55*(1 + v mod 40) Hz) per GPU,
512-sample blocks @ 48 kHz.  One "step" = one block of all voices through the fused render kernel (+ the
mix-bus reduction, + the cross-GPU all-reduce of the [1][512] mix bus when N > 1).  Weak scaling: every rank
owns its own 4096 voices; the only data-path collective is the mix-bus reduce.

  value  : Msamples/s, device-resident (state, delay rings and parameters in HBM; no host I/O in the step),
           CUDA events per step, L2 flushed (256 MiB memset) between steps outside the timed events.
  e2e    : same metric through the public API call a user makes (elem_b200_process: host output buffers,
           D2H of the mix bus and the synchronisation inside the timed region; the graph has no audio inputs,
           so h2d_bytes_per_step is 0).
  roofline: K1 render kernel only — algorithmic bytes per launch / mean launch duration measured
           with CUDA events around every K1 launch on its own stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the UNMODIFIED reference engine (oracle/_ref/libelem_ref.so, built from
           /root/reference in place) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, BS = 48000.0, 512
VOICES_PER_GPU = 4096
ALGO_BYTES_PER_VOICE_BLOCK_MIX_ONLY = 4236      # SURVEY.md §8d / DESIGN.md §4: 88 state + 52 params + 4096 delay ring
METRIC = "Msamples/s (voices x 512-sample blocks / s), 4096-voice SUBSYNTH32 per GPU @ 48 kHz"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); mx.append(float(c[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def build_runtime(voices, device, rank, stream_handle=None, **opts):
    import numpy as np
    from elementary_b200 import Runtime, graphs
    rt = Runtime(SR, BS, voices, device=device, **opts)
    if stream_handle is not None:
        rt.set_stream(stream_handle)
    assert rt.apply_instructions(graphs.subsynth32()) == 0, rt.last_error()
    ida, idb = graphs.subsynth32_param_ids()
    f0 = np.array([graphs.subsynth32_f0(rank * voices + v) for v in range(voices)], dtype=np.float64)
    assert rt.set_property_per_voice(ida, "value", f0) == 0
    assert rt.set_property_per_voice(idb, "value", f0 * 1.007) == 0
    return rt


def cpu_reference_run(voices, threads, warmup_blocks, blocks, rank0_voice_offset=0):
    """Time the unmodified reference (or, if oracle/_ref did not travel, nothing) on `threads` host threads."""
    from elementary_b200 import graphs
    from oracle import oracle as orc
    if not orc.ref_available():
        return None
    vb = [graphs.subsynth32_voice_props(rank0_voice_offset + v) for v in range(voices)]
    secs, chk = orc.ref_bench(SR, BS, graphs.subsynth32(), vb, voices, threads, 0, 1, warmup_blocks, blocks)
    if secs <= 0:
        return None
    return {"seconds": secs, "msamples_per_s": voices * BS * blocks / secs / 1e6,
            "voice_blocks_per_s": voices * blocks / secs, "checksum": chk}


def run_reference(args, rank, world, emit=print):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    voices = VOICES_PER_GPU * world
    r = cpu_reference_run(voices, cores, args.warmup, args.steps)
    if r is None:
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libelem_ref.so missing (built only where /root/reference exists)"}))
        return
    ms = r["seconds"] / args.steps * 1e3
    sample = f"all {voices} voices x {args.steps} blocks of 512 on {cores} host threads (whole workload, not a subset)"
    line = {
        "impl": "reference", "metric": METRIC, "value": r["msamples_per_s"], "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{VOICES_PER_GPU} SUBSYNTH32 voices per GPU x {world} (BASELINE.json configs[1]), 512-sample blocks, 48 kHz",
                   "engine": "elem::Runtime<float> (unmodified reference, oracle/_ref), one instance per voice, round-robin per thread"},
        "voice_blocks_per_s": r["voice_blocks_per_s"],
        "cpu_baseline": {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": r["msamples_per_s"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(json.dumps(line))


def run_b200(args, rank, local_rank, world, emit=print):
    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    # NCCL prints its version banner to stdout when NCCL_DEBUG=VERSION (it honours NCCL_DEBUG_FILE only above that level)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    voices = args.voices
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        extra = {"tile_width": args.tile_width} if args.tile_width else {}
        for kv in args.opt:
            k, v = kv.split("=")
            extra[k] = float(v)
        rt = build_runtime(voices, local_rank, rank, stream.cuda_stream, time_kernels=1, **extra)
        mix = torch.as_tensor(rt.mix_device(1), device=f"cuda:{local_rank}")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
        host_mix = torch.empty((1, BS), dtype=torch.float32).pin_memory()
        from elementary_b200.runtime import FLAG_MIX, FLAG_ALLREDUCE
        from elementary_b200.distributed import attach_peer_mix

        # The single collective of the path: the sum of the [1][512] mix bus over the ranks.  Default: the engine's own
        # kernel over NVLink/NVSwitch peer memory (K4, launched by enqueue_block in the same stream); --collective nccl
        # (or a box without peer access) uses torch.distributed.all_reduce instead.
        fused, fused_note = False, None
        if world > 1 and args.collective == "fused":
            try:
                attach_peer_mix(rt)
                fused = True
            except Exception as e:      # reported in the JSON line, never silent
                fused_note = f"peer attach failed ({e}); NCCL all_reduce used"
            ok = torch.tensor([1 if fused else 0], device=f"cuda:{local_rank}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            fused = bool(int(ok[0]))
        step_flags = FLAG_MIX | (FLAG_ALLREDUCE if fused else 0)

        def step_device():
            rt.enqueue_block(0, 1, BS, step_flags)
            if world > 1 and not fused:
                dist.all_reduce(mix)

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- warm-up (also carries every voice past the 20 ms root fade) ----
        for _ in range(max(3, args.warmup)):
            step_device()
        barrier()
        rt.take_kernel_time_ms()

        # ---- value: device-resident, per-step CUDA events, L2 flushed between steps ----
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        launches0 = rt.kernel_launches
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t_wall0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()
            a.record(stream)
            step_device()
            b.record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        dev_ms = sum(a.elapsed_time(b) for a, b in evs)
        k1_ms, k1_n = rt.take_kernel_time_ms()
        launches = rt.kernel_launches - launches0

        # ---- e2e: the public call with host buffers (D2H + sync inside), no flush needed for honesty: flushed too ----
        barrier()
        e2e_s = 0.0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if world == 1:
                out = rt.process(None, 1, BS)
            else:
                step_device()
                if rank == 0:
                    host_mix.copy_(mix[:1], non_blocking=True)
                stream.synchronize()
            e2e_s += time.perf_counter() - t0
        barrier()
        clocks = sampler.stop() if rank == 0 else None      # nvidia-smi samples cover both timed regions (value and e2e)

        peer_fail = rt.peer_status() if fused else 0
        t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device=f"cuda:{local_rank}")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        return
    total_samples = world * voices * BS * args.steps
    value = total_samples / (dev_ms * 1e-3) / 1e6
    e2e_value = total_samples / (e2e_ms * 1e-3) / 1e6
    peak, peak_kind = measured_hbm_peak()
    k1_avg_ms = k1_ms / max(1, k1_n)
    algo_bytes = ALGO_BYTES_PER_VOICE_BLOCK_MIX_ONLY * voices
    achieved = algo_bytes / (k1_avg_ms * 1e-3) / 1e9 if k1_avg_ms > 0 else 0.0
    desc = rt.describe()["groups"][0]
    traffic, ncu = None, None   # from the committed `ncu --set full` capture of this same configuration (profiles/)
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["render_block_kernel"].get(str(voices))
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            ncu = {k: tj[k] for k in ("warp_instructions", "issue_active_pct", "registers_per_thread", "fp64_pipe_pct",
                                      "active_threads_per_warp_inst", "report") if k in tj}
            if "warp_instructions" in tj:
                ncu["warp_instructions_per_voice_sample"] = tj["warp_instructions"] / (voices * BS)
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cv = 16 * cores
        blocks = 1500
        r = cpu_reference_run(cv, cores, 20, blocks)
        if r is not None:
            cpu = {"value": r["msamples_per_s"], "unit": "Msamples/s", "cores": cores, "kind": "reference",
                   "sample": f"{cv} voices (16 per thread) x {blocks} blocks of 512, {r['seconds']:.1f} s on {cores} threads, unmodified reference via oracle/_ref"}

    line = {
        "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{voices} SUBSYNTH32 voices per GPU (BASELINE.json configs[1]: 32-node subtractive synth), 512-sample blocks, 48 kHz",
                   "voices_total": world * voices, "block": BS, "sample_rate": SR,
                   "l2": "flushed between steps (256 MiB memset outside the timed events)",
                   "tile_width": desc["tile_width"], "slots": desc["slots"], "state_rows": desc["state_rows"],
                   "spec": {k: desc[k] for k in ("spec_state", "spec_regs", "spec_local_bytes", "spec_cubin_bytes", "spec_log") if k in desc},
                   "collective": "none (1 GPU)" if world == 1 else
                                 ("K4 mix_exchange_kernel: all-reduce(sum,f32) of the [1][512] mix bus per block over NVLink peer memory, one launch in the render stream"
                                  + ("" if not peer_fail else " — PEER TIMEOUT REPORTED") if fused else
                                  "NCCL all_reduce(sum,f32) of the [1][512] mix bus per block" + (f" ({fused_note})" if fused_note else ""))},
        "voice_blocks_per_s": world * voices * args.steps / (dev_ms * 1e-3),
        "realtime_factor": value * 1e6 / (world * voices * SR),
        "wall_ms_per_step_incl_flush": t_wall / args.steps * 1e3,
        "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 4 * BS,
                "ms_per_step": e2e_ms / args.steps, "api": "elem_b200_process (host out buffers)" if world == 1 else
                       ("elem_b200_enqueue_block (render + K4 cross-GPU mix) + D2H of the mix bus" if fused else "elem_b200_enqueue_block + NCCL all_reduce + D2H of the mix bus")},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_kind": f"of {peak_kind}", "kernel": "render_block_kernel<NITER,LOGL> (K1)",
                     "kernel_ms": k1_avg_ms, "kernel_launches_timed": k1_n, "ncu": ncu,
                     "algorithmic_bytes_per_launch": algo_bytes,
                     "note": "K1 is instruction-issue bound, not HBM bound (intermediates never leave the SM); see DESIGN.md section 4"},
        "clocks": clocks,
        "cpu_baseline": cpu,
    }
    emit(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--voices", type=int, default=VOICES_PER_GPU, help="voices per GPU (default: the BASELINE config)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--collective", default="fused", choices=["fused", "nccl"], help="N > 1: K4 peer-memory kernel (default) or NCCL all_reduce")
    ap.add_argument("--tile-width", type=int, default=0, help="override the voices-per-warp heuristic (exploration only)")
    ap.add_argument("--opt", action="append", default=[], help="extra runtime option key=value (exploration only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: anything a native library writes to fd 1 meanwhile (NCCL's banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    emit = lambda text: lines.append(text)
    try:
        if args.impl == "reference":
            run_reference(args, rank, world, emit)
        else:
            run_b200(args, rank, local_rank, world, emit)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    for text in lines:
        print(text, flush=True)


if __name__ == "__main__":
    main()
