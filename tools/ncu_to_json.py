#!/usr/bin/env python
"""Extract the per-launch numbers bench.py quotes (DRAM traffic, instruction count, issue utilisation) from .ncu-rep files.
Usage: python tools/ncu_to_json.py profiles/ncu_traffic.json  key=rep[:kernel-substring] ...   (key e.g. render_block_kernel/4096)"""
import csv, io, json, subprocess, sys

M = {"dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum", "duration_ns": "gpu__time_duration.sum",
     "warp_instructions": "smsp__inst_executed.sum", "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "registers_per_thread": "launch__registers_per_thread", "fp64_pipe_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
     "active_threads_per_warp_inst": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "local_load_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local_store_sectors": "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum"}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e6, "us": 1e3, "ns": 1.0, "s": 1e9}


def extract(rep, sub):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    for r in rows[2:]:
        if sub and sub not in r[ki]:
            continue
        d = {"kernel": r[ki], "report": rep}
        for k, m in M.items():
            if m in hdr:
                i = hdr.index(m)
                v = float(r[i].replace(",", ""))
                d[k] = v * UNIT.get(units[i], 1.0) if k.startswith("dram") or k == "duration_ns" else v
        return d
    return None


def main():
    dst = sys.argv[1]
    import hashlib, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha256()
    for f in ("render_kernel.cu", "render_ops.inc", "program.h"):      # bench.py quotes these numbers only for a build of the same K1 sources
        h.update(open(os.path.join(root, "elementary_b200", "csrc", f), "rb").read())
    res = {"k1_source_sha16": h.hexdigest()[:16], "note": "per-launch values of ONE launch from `ncu --set full --clock-control none` (cold-cache, serialised); see tools/ncu_to_json.py"}
    for a in sys.argv[2:]:
        key, rest = a.split("=", 1)
        rep, _, sub = rest.partition(":")
        top, _, leaf = key.partition("/")
        res.setdefault(top, {})[leaf or "default"] = extract(rep, sub)
    json.dump(res, open(dst, "w"), indent=1)
    print(json.dumps(res, indent=1)[:1500])


if __name__ == "__main__":
    main()
