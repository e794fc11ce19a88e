"""GPU debug helper: print where GPU streams differ from the CPU model for chosen cases."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import flic_b200 as flic, cases, oracle_binding
orc = oracle_binding.Oracle(os.path.join(ROOT, "oracle", "libflp0_oracle.so"))
codec = flic.Codec(0)

def show(tag, img, flags):
    want = orc.encode(img, flags)
    try:
        got = codec.encode(img, flags)
    except flic.FlicError as e:
        print(tag, img.shape, hex(flags), "ERROR", e)
        return False
    if got.size == want.size and np.array_equal(got, want):
        return True
    gw, ww = got[: got.size // 4 * 4].view(np.uint32), want.view(np.uint32)
    n = min(gw.size, ww.size)
    bad = np.nonzero(gw[:n] != ww[:n])[0]
    print(tag, img.shape, hex(flags), "sizes", got.size, want.size, "differing words", bad.size, "first", bad[:12])
    for i in bad[:12]:
        print("   word", int(i), "gpu %08x" % gw[i], "model %08x" % ww[i])
    return False

show("one_pixel", dict(cases.SMALL)["one_pixel"](), 0x01)
for v in (0, 1, 5, 200):
    show(f"1x1 value {v}", np.full((1, 1, 1), v, np.uint8), 0x01)
for shape in ((1, 1, 3), (1, 2, 1), (2, 1, 1), (1, 4, 1), (1, 5, 1), (3, 3, 1), (1, 1, 4), (1, 1, 2)):
    show("flat-ish", np.full(shape, 9, np.uint8), 0x01)
    show("ramp", (np.arange(np.prod(shape)) * 37 % 256).astype(np.uint8).reshape(shape), 0x01)

rng = np.random.default_rng(2024)
nbad = 0
for trial in range(120):
    c = int(rng.integers(1, 5))
    w = int(rng.choice([1, 7, 60, 127, 128, 129, 200, 255, 256, 260, 384, 500]))
    h = int(rng.choice([1, 5, 31, 32, 33, 64, 70]))
    kind = trial % 4
    if kind == 0: img = cases.gradient(w, h, c, 1000 + trial)
    elif kind == 1: img = cases.noise(w, h, c, 1000 + trial)
    elif kind == 2: img = cases.skewed(w, h, c, 1000 + trial)
    else: img = cases.gradient(w, h, c, 1000 + trial, sigma=float(rng.choice([0.0, 0.7, 12.0])))
    if rng.random() < 0.4:
        ch = int(rng.integers(0, c)); x1 = w if rng.random() < 0.5 else max(1, w // 2)
        img = img.copy(); img[:, :x1, ch] = int(rng.integers(0, 256))
    flags = 0x11 if (c >= 3 and rng.random() < 0.5) else 0x01
    flags |= int(rng.choice([0, 0, 0x20, 0x40]))
    ok = show(f"trial {trial} kind {kind}", img, flags)
    if ok:
        try:
            back = codec.decode(codec.encode(img, flags))
            if not np.array_equal(back, img):
                print("trial", trial, img.shape, hex(flags), "DECODE MISMATCH", int(np.argmax(back.ravel() != img.ravel())))
                ok = False
        except flic.FlicError as e:
            print("trial", trial, img.shape, hex(flags), "DECODE ERROR", e); ok = False
    nbad += not ok
print("bad trials:", nbad)
