"""Extract the reference's OWN golden vectors (jest snapshots of js/packages/offline-renderer/__tests__) for the nodes on the
path into tests/golden/jest_snapshots.json.  Run where /root/reference exists:  python tests/golden/make_jest_golden.py

The snapshots were produced by the reference's wasm build (Runtime<double>, float32 I/O).  tests/test_jest_goldens.py replays the
jest tests that made them — same graphs, same inputs, same process() call pattern — against the oracles (CPU) and the CUDA path (GPU).
"""
import json
import os
import re
import sys

REF = "/root/reference/js/packages/offline-renderer/__tests__/__snapshots__"
FILES = ["sparseq.test.js.snap", "sparseq2.test.js.snap", "time.test.js.snap", "events.test.js.snap",
         "delays.test.js.snap", "maxhold.test.js.snap", "tap.test.js.snap", "vfs.test.js.snap"]
HERE = os.path.dirname(os.path.abspath(__file__))


def parse_value(text):
    """Jest pretty-format -> Python: Float32Array [..] / Array [..] / Object {..} / numbers / strings."""
    text = text.strip()
    text = re.sub(r"\b(Float32Array|Array)\s*\[", "[", text)
    text = re.sub(r"\bObject\s*\{", "{", text)
    text = re.sub(r"\bundefined\b", "null", text)
    text = re.sub(r",(\s*[\]}])", r"\1", text)          # trailing commas
    return json.loads(text)


def main():
    out = {"_source": "js/packages/offline-renderer/__tests__/__snapshots__/*.snap of the reference", "snapshots": {}}
    for fn in FILES:
        path = os.path.join(REF, fn)
        if not os.path.exists(path):
            continue
        src = open(path).read()
        for m in re.finditer(r"exports\[`(.+?)`\] = `\n(.*?)\n`;", src, flags=re.S):
            name, body = m.group(1), m.group(2)
            try:
                out["snapshots"][f"{fn}:{name}"] = parse_value(body)
            except Exception as e:   # a snapshot in a shape we do not need
                print("skipped", fn, name, e, file=sys.stderr)
    with open(os.path.join(HERE, "jest_snapshots.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(len(out["snapshots"]), "snapshots")


if __name__ == "__main__":
    main()
