"""Regenerates tests/golden/* from the FLP0 CPU model (oracle/).

These are SELF-goldens of the provisional format: they freeze today's bitstream so that a
later edit to the model (or to the CUDA engine checked against it) cannot drift silently.
They are NOT reference vectors — the reference is behind a licensing gate (LICENSING.md)
and ships no fixtures of its own.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cases  # noqa: E402
import oracle_binding  # noqa: E402

orc = oracle_binding.Oracle(os.path.join(HERE, "..", "..", "oracle", "libflp0_oracle.so"))
gold = {"format": "FLP0 v3 (flat channels; layouts: slots / 0x20 one stream / 0x40 exact), block 128x32, max code length 10", "sha256": {}}
for name, build in cases.SMALL:
    for flags in cases.ALL_FLAGS:
        s = orc.encode(build(), flags)
        gold["sha256"][f"{name}@{flags:#04x}"] = {"bytes": int(s.size), "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
with open(os.path.join(HERE, "golden.json"), "w") as f:
    json.dump(gold, f, indent=1, sort_keys=True)
tiny = cases.gradient(96, 40, 3, 99)
np.save(os.path.join(HERE, "tiny_96x40x3.npy"), tiny)
orc.encode(tiny, 0x01).tofile(os.path.join(HERE, "tiny_96x40x3.flp0"))
print("wrote", len(gold["sha256"]), "hashes + tiny vector")
