#!/usr/bin/env python3
"""Regenerate tests/golden/scenes/*.json.gz from the reference's shipped scene files.

The five `*.json.gz` scenes at the reference's repo root are INPUT DATA (serde_json dumps of
`Scene`, reference src/main.rs:93-102, 299-313), not source code.  /root/reference does not exist
on the GPU box, so the parity tests and bench.py read these re-encoded copies instead.  The copy
is a semantic re-serialisation (json.load -> json.dumps with ascending numeric keys -> gzip with
mtime 0), i.e. the same `Scene` value in the same wire format, not the same bytes.

    python tools/import_scenes.py [/root/reference]
"""
import gzip
import json
import os
import sys

SCENES = ["cornell", "cornell2", "scene", "volume", "cloud"]


def canonical(scene):
    for coll in ("objects", "data"):
        c = scene[coll]["collection"]
        scene[coll]["collection"] = {k: c[k] for k in sorted(c, key=int)}
    return scene


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scenes")
    os.makedirs(out_dir, exist_ok=True)
    for name in SCENES:
        with gzip.open(os.path.join(ref, name + ".json.gz"), "rt") as f:
            scene = canonical(json.load(f))
        text = json.dumps(scene, separators=(",", ":"))
        path = os.path.join(out_dir, name + ".json.gz")
        with open(path, "wb") as raw:
            with gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0, compresslevel=9) as gz:
                gz.write(text.encode())
        print(f"{name}: {len(text)} JSON bytes -> {os.path.getsize(path)} gz bytes")


if __name__ == "__main__":
    main()
