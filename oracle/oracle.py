"""ctypes bindings for the two CHECKERS.  TEST INFRASTRUCTURE — never imported by the product package.

* :class:`RefRuntime`  — the unmodified reference engine (``oracle/_ref/libelem_ref.so``, built by
  ``oracle/Makefile`` from /root/reference in place; prebuilt file travels to the GPU box).
* :class:`PortRuntime` — our CPU restatement of the reference algorithm (``oracle/libelem_oracle.so``,
  source ``oracle/elem_oracle.cpp``), validated against RefRuntime and the reference's golden vectors.

Both expose the same small surface: ``apply(batch)``, ``add_shared_resource(name, data)``,
``process(inputs, n_out, n)`` -> ``np.ndarray [n_out, n]`` — one instance == one voice.
Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl reference`` legs may use it.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(_HERE, "_ref", "libelem_ref.so")
PORT_LIB = os.path.join(_HERE, "libelem_oracle.so")

_f32p = C.POINTER(C.c_float)


def ref_available() -> bool:
    return os.path.exists(REF_LIB)


def port_available() -> bool:
    return os.path.exists(PORT_LIB)


_ref_lib = None
_port_lib = None


def _load_ref():
    global _ref_lib
    if _ref_lib is None:
        lib = C.CDLL(REF_LIB)
        lib.elem_ref_create.restype = C.c_void_p
        lib.elem_ref_create.argtypes = [C.c_double, C.c_int]
        lib.elem_ref_destroy.argtypes = [C.c_void_p]
        lib.elem_ref_apply_instructions.restype = C.c_int
        lib.elem_ref_apply_instructions.argtypes = [C.c_void_p, C.c_char_p]
        lib.elem_ref_add_shared_resource.restype = C.c_int
        lib.elem_ref_add_shared_resource.argtypes = [C.c_void_p, C.c_char_p, _f32p, C.c_size_t]
        lib.elem_ref_process_flat.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t]
        lib.elem_ref_set_current_time.argtypes = [C.c_void_p, C.c_int64]
        lib.elem_ref_process_queued_events.restype = C.c_int
        lib.elem_ref_process_queued_events.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        lib.elem_ref_gc.restype = C.c_int
        lib.elem_ref_gc.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.c_size_t]
        lib.elem_ref_bench.restype = C.c_double
        lib.elem_ref_bench.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p),
                                       C.c_char_p, _f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, C.c_size_t,
                                       C.c_int, C.c_int, C.POINTER(C.c_double)]
        lib.elem_ref_bench_stable.restype = C.c_double
        lib.elem_ref_bench_stable.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_char_p, C.POINTER(C.c_char_p),
                                              C.c_char_p, _f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, C.c_size_t,
                                              C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        lib.elem_ref_cpu_model.restype = C.c_int
        lib.elem_ref_cpu_model.argtypes = [C.c_char_p, C.c_size_t]
        _ref_lib = lib
    return _ref_lib


def _load_port():
    global _port_lib
    if _port_lib is None:
        lib = C.CDLL(PORT_LIB)
        lib.elem_oracle_create.restype = C.c_void_p
        lib.elem_oracle_create.argtypes = [C.c_double, C.c_int]
        lib.elem_oracle_destroy.argtypes = [C.c_void_p]
        lib.elem_oracle_apply_text.restype = C.c_int
        lib.elem_oracle_apply_text.argtypes = [C.c_void_p, C.c_char_p]
        lib.elem_oracle_add_shared_resource.restype = C.c_int
        lib.elem_oracle_add_shared_resource.argtypes = [C.c_void_p, C.c_char_p, _f32p, C.c_size_t]
        lib.elem_oracle_process_flat.argtypes = [C.c_void_p, _f32p, C.c_size_t, _f32p, C.c_size_t, C.c_size_t]
        lib.elem_oracle_set_current_time.argtypes = [C.c_void_p, C.c_longlong]
        lib.elem_oracle_process_queued_events.restype = C.c_int
        lib.elem_oracle_process_queued_events.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        _port_lib = lib
    return _port_lib


def _as_f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def batch_to_text(batch: Sequence[list]) -> str:
    """Line format consumed by the restatement (it deliberately has no JSON parser of its own):
    ``0 id type`` / ``2 parent child chan`` / ``3 id key N <number>`` / ``3 id key S <string>`` /
    ``3 id key B <0|1>`` / ``4 id id ...`` / ``5``."""
    lines: List[str] = []
    for ins in batch:
        op = int(ins[0])
        if op == 0:
            lines.append(f"0 {int(ins[1])} {ins[2]}")
        elif op == 2:
            lines.append(f"2 {int(ins[1])} {int(ins[2])} {int(ins[3])}")
        elif op == 3:
            v = ins[3]
            if isinstance(v, bool):
                lines.append(f"3 {int(ins[1])} {ins[2]} B {int(v)}")
            elif isinstance(v, (int, float)):
                lines.append(f"3 {int(ins[1])} {ins[2]} N {float(v)!r}")
            elif isinstance(v, str):
                lines.append(f"3 {int(ins[1])} {ins[2]} S {v}")
            elif v is None:
                lines.append(f"3 {int(ins[1])} {ins[2]} U")
            elif isinstance(v, (list, tuple)) and all(isinstance(x, (int, float)) and not isinstance(x, bool) for x in v):
                lines.append(f"3 {int(ins[1])} {ins[2]} A {len(v)} " + " ".join(repr(float(x)) for x in v))
            elif isinstance(v, (list, tuple)) and v and all(isinstance(x, dict) for x in v):
                keys = sorted(v[0].keys())
                lines.append(f"3 {int(ins[1])} {ins[2]} M {len(v)} {len(keys)} " +
                             " ".join(f"{k} {float(x[k])!r}" for x in v for k in keys))
            else:
                lines.append(f"3 {int(ins[1])} {ins[2]} J {json.dumps(v)}")
        elif op == 4:
            lines.append("4 " + " ".join(str(int(x)) for x in ins[1]))
        elif op == 5:
            lines.append("5")
        else:
            lines.append(f"{op}")
    return "\n".join(lines) + "\n"


class _Base:
    def process(self, inputs: Optional[np.ndarray], n_out: int, n: int) -> np.ndarray:
        if inputs is None or len(inputs) == 0:
            n_in = 0
            inp = np.zeros(1, dtype=np.float32)
        else:
            inp = _as_f32(inputs)
            assert inp.ndim == 2 and inp.shape[1] == n
            n_in = inp.shape[0]
        out = np.zeros((n_out, n), dtype=np.float32)
        self._process(inp.ctypes.data_as(_f32p), n_in, out.ctypes.data_as(_f32p), n_out, n)
        return out

    def render(self, n_blocks: int, n_out: int = 1, n: Optional[int] = None, inputs: Optional[np.ndarray] = None) -> np.ndarray:
        """Run ``n_blocks`` blocks; ``inputs`` is [n_in, n_blocks*n] or None. Returns [n_out, n_blocks*n]."""
        n = n or self.block_size
        outs = []
        for b in range(n_blocks):
            inp = None if inputs is None else np.asarray(inputs)[:, b * n:(b + 1) * n]
            outs.append(self.process(inp, n_out, n))
        return np.concatenate(outs, axis=1)


class RefRuntime(_Base):
    def __init__(self, sample_rate: float = 48000.0, block_size: int = 512):
        self.lib = _load_ref()
        self.sample_rate, self.block_size = sample_rate, block_size
        self.h = C.c_void_p(self.lib.elem_ref_create(sample_rate, block_size))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.elem_ref_destroy(self.h)
            self.h = None

    def apply(self, batch) -> int:
        s = batch if isinstance(batch, str) else json.dumps(batch)
        return self.lib.elem_ref_apply_instructions(self.h, s.encode())

    def add_shared_resource(self, name: str, data) -> bool:
        d = _as_f32(data)
        return bool(self.lib.elem_ref_add_shared_resource(self.h, name.encode(), d.ctypes.data_as(_f32p), d.size))

    def gc(self) -> List[int]:
        buf = (C.c_int32 * 4096)()
        n = self.lib.elem_ref_gc(self.h, buf, 4096)
        return sorted(buf[i] for i in range(min(n, 4096)))

    def _process(self, inp, n_in, out, n_out, n):
        self.lib.elem_ref_process_flat(self.h, inp, n_in, out, n_out, n)

    def set_current_time(self, t: int) -> None:
        self.lib.elem_ref_set_current_time(self.h, int(t))

    def process_queued_events(self) -> list:
        """Runtime::processQueuedEvents relayed like wasm/Main.cpp:220-231: [{"type": ..., "event": {...}}, ...]."""
        buf = C.create_string_buffer(1 << 22)
        self.lib.elem_ref_process_queued_events(self.h, buf, len(buf))
        return json.loads(buf.value.decode() or "[]")


class PortRuntime(_Base):
    def __init__(self, sample_rate: float = 48000.0, block_size: int = 512):
        self.lib = _load_port()
        self.sample_rate, self.block_size = sample_rate, block_size
        self.h = C.c_void_p(self.lib.elem_oracle_create(sample_rate, block_size))

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.elem_oracle_destroy(self.h)
            self.h = None

    def apply(self, batch) -> int:
        if isinstance(batch, str):
            batch = json.loads(batch)
        return self.lib.elem_oracle_apply_text(self.h, batch_to_text(batch).encode())

    def add_shared_resource(self, name: str, data) -> bool:
        d = _as_f32(data)
        return bool(self.lib.elem_oracle_add_shared_resource(self.h, name.encode(), d.ctypes.data_as(_f32p), d.size))

    def _process(self, inp, n_in, out, n_out, n):
        self.lib.elem_oracle_process_flat(self.h, inp, n_in, out, n_out, n)

    def set_current_time(self, t: int) -> None:
        self.lib.elem_oracle_set_current_time(self.h, int(t))

    def process_queued_events(self) -> list:
        buf = C.create_string_buffer(1 << 22)
        self.lib.elem_oracle_process_queued_events(self.h, buf, len(buf))
        return json.loads(buf.value.decode() or "[]")


def ref_bench(sample_rate: float, block_size: int, base_batch, voice_batches: Optional[Sequence], n_voices: int,
              threads: int, n_in: int, n_out: int, warmup_blocks: int, blocks: int,
              resource: Optional[tuple] = None, inputs: Optional[np.ndarray] = None, min_seconds: float = 0.0):
    """Time the reference's own CPU path: returns (seconds, checksum[, steps = blocks of all voices rendered, when min_seconds > 0]).  Threads are pinned, created
    and warmed up outside the timed region, rounds are barrier-bracketed: oracle/ref_driver.cpp:elem_ref_bench_stable."""
    lib = _load_ref()
    base = json.dumps(base_batch).encode()
    if voice_batches is not None:
        arr = (C.c_char_p * n_voices)(*[json.dumps(b).encode() for b in voice_batches])
    else:
        arr = None
    if resource is not None:
        rname, rdata = resource[0].encode(), _as_f32(resource[1])
        rptr, rlen = rdata.ctypes.data_as(_f32p), rdata.size
    else:
        rname, rptr, rlen = None, None, 0
    if inputs is not None:
        inp = _as_f32(inputs)
        iptr = inp.ctypes.data_as(_f32p)
    else:
        iptr = None
    chk = C.c_double(0.0)
    rep = C.c_double(0.0)
    secs = lib.elem_ref_bench_stable(sample_rate, block_size, n_voices, threads, base, arr, rname, rptr, rlen,
                                     iptr, n_in, n_out, block_size, warmup_blocks, blocks, float(min_seconds), C.byref(rep), C.byref(chk))
    if min_seconds > 0:
        return secs, chk.value, rep.value      # the timed round amounts to rep.value blocks of all voices
    return secs, chk.value


def ref_cpu_model() -> str:
    lib = _load_ref()
    buf = C.create_string_buffer(256)
    lib.elem_ref_cpu_model(buf, 256)
    return buf.value.decode(errors="replace")
