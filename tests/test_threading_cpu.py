"""The boundary's threading contract (include/elem_b200.h; reference Runtime.h:133,204,277-285): one control thread and one render
thread work on the same runtime concurrently.  tests/native/thread_stress.cpp does exactly that — graph re-renders with cross-fades,
a live voice-group cut, gc, events and describe on one thread against a block loop on the other — here under ThreadSanitizer against
a TSAN build of the library's host side (plan-only runtime with option plan_dry_run: every host-side step of a block, no GPU).  Any
data race makes TSAN exit with code 66."""
import json
import os
import shutil
import subprocess

import pytest

from elementary_b200 import el, graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "elementary_b200", "csrc")


def write_graphs(tmp_path):
    a, b = tmp_path / "a.json", tmp_path / "b.json"
    a.write_text(json.dumps(graphs.subsynth32()))
    g2 = el.tanh(el.add(graphs.subsynth32_graph(220.0), el.mul(0.2, el.cycle(330.0))))
    b.write_text(json.dumps(el.render(g2)))
    return str(a), str(b)


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("g++") is None, reason="needs the compilers")
def test_control_and_render_threads_are_race_free_under_tsan(tmp_path):
    p = subprocess.run(["make", "-j", "8", "tsan"], cwd=CSRC, capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    exe = os.path.join(ROOT, "build", "tsan", "thread_stress_tsan")
    syms = subprocess.run(["nm", "-D", os.path.join(ROOT, "build", "tsan", "libelem_b200_tsan.so")], capture_output=True, text=True).stdout
    assert "__tsan_" in syms, "the library under test is not TSAN-instrumented"
    a, b = write_graphs(tmp_path)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66 report_signal_unsafe=0")
    r = subprocess.run([exe, "-1", "2.0", a, b], capture_output=True, text=True, env=env, timeout=300)
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-6000:]
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-3000:])
    stats = json.loads(r.stdout.strip().splitlines()[-1])
    assert stats["blocks"] > 100 and stats["edits"] > 10 and stats["failures"] == 0
