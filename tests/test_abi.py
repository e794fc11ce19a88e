"""CPU tests of the C-ABI boundary: the library loads, exports every symbol include/flic_b200.h
declares, and its host-only entry points (peek, splice, size queries, argument checks) behave.
No compute call is made here — those need a GPU (test_gpu_parity.py)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "flic_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(flic_[a-z_0-9]+)\s*\(", text)))


def test_exports_match_header(flic):
    lib = C.CDLL(flic.library_path())
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/flic_b200.h but not exported"
    from importlib import import_module
    codec = import_module("fast-losless-image-compression-format_b200.codec")
    assert sorted(codec.EXPORTED) == names


def test_header_is_plain_c_and_links(flic, tmp_path):
    """The boundary is a C ABI: the header must compile as C99 with no CUDA / C++ / torch types, and a C
    caller must link against the library (link only — running it needs a GPU)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = tmp_path / "caller.c"
    src.write_text(
        '#include "flic_b200.h"\n'
        "int main(void) {\n"
        "    flic_ctx *ctx = 0; flic_image_info info;\n"
        "    unsigned char hdr[FLIC_HEADER_BYTES] = {0};\n"
        "    if (flic_create(0, &ctx) != FLIC_OK) return flic_peek(hdr, sizeof hdr, &info) == FLIC_E_FORMAT ? 0 : 1;\n"
        "    flic_destroy(ctx);\n"
        "    return (int)(flic_max_stream_bytes(128, 32, 4) == 0);\n"
        "}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(flic.library_path())
    exe = tmp_path / "caller"
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe), f"-L{libdir}",
                        "-lflicb200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_constants_agree_with_header(flic):
    text = open(os.path.join(ROOT, "include", "flic_b200.h")).read()
    d = dict(re.findall(r"#define (FLIC_[A-Z_]+) (\(?-?\w+\)?)", text))
    assert int(d["FLIC_BLOCK_W"]) == flic.BLOCK_W and int(d["FLIC_BLOCK_H"]) == flic.BLOCK_H
    assert int(d["FLIC_MAX_CODE_LEN"]) == flic.MAX_CODE_LEN
    oh = open(os.path.join(ROOT, "oracle", "flp0_oracle.h")).read()
    assert f"#define FLP0_MAX_CODE_LEN {flic.MAX_CODE_LEN}" in oh


def test_size_queries(flic, oracle):
    lib = flic.load_library()
    for (w, h, c) in [(1, 1, 1), (128, 32, 4), (129, 33, 3), (3840, 2160, 4), (1920, 1080, 3)]:
        assert lib.flic_blocks_per_image(w, h) == -(-w // 128) * -(-h // 32)
        assert flic.max_stream_bytes(w, h, c) == oracle.lib.flp0_max_stream_bytes(w, h, c, 128, 32)


def test_no_device_fails_loudly(flic):
    """On a box without an sm_100 GPU the engine refuses to start — there is no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(flic.FlicError) as e:
        flic.Codec(0)
    assert e.value.code == -5


def test_peek_and_errors(flic, oracle):
    img = cases.gradient(300, 70, 3, 5)
    s = oracle.encode(img, 0x11)
    info = flic.peek(s)
    assert (info["width"], info["height"], info["channels"], info["flags"]) == (300, 70, 3, 0x11)
    assert info["n_blocks"] == 3 * 3 and info["block_w"] == 128 and info["block_h"] == 32
    assert 32 + 4 * (info["n_blocks"] + 1 + info["payload_words"]) == s.size
    for bad in (s[:16], s[:-4]):
        with pytest.raises(flic.FlicError) as e:
            flic.peek(bad)
        assert e.value.code == -3
    t = s.copy(); t[0] ^= 1
    with pytest.raises(flic.FlicError):
        flic.peek(t)
    t = s.copy(); t[7] = 0x02  # predictor id 2 does not exist
    with pytest.raises(flic.FlicError):
        flic.peek(t)
    assert flic.load_library().flic_strerror(-5).decode().startswith("no sm_100")


@pytest.mark.parametrize("shape,cuts", [((100, 300, 3), [64]), ((200, 130, 4), [32, 96, 160]), ((33, 50, 1), [32]),
                                        ((64, 256, 4), [32])])
def test_splice_block_rows(flic, oracle, shape, cuts):
    """Streams of whole block-row runs splice (host code in the library) into exactly the full-image stream."""
    h, w, c = shape
    img = cases.gradient(w, h, c, 31)
    full = oracle.encode(img)
    edges = [0] + cuts + [h]
    parts = [oracle.encode(img[a:b]) for a, b in zip(edges, edges[1:])]
    assert np.array_equal(flic.splice_block_rows(parts), full)
    assert np.array_equal(oracle.decode(flic.splice_block_rows(parts), img.shape), img)


def test_splice_rejects_misaligned_parts(flic, oracle):
    img = cases.gradient(140, 100, 3, 32)
    with pytest.raises(flic.FlicError) as e:  # first part is not a whole number of block rows
        flic.splice_block_rows([oracle.encode(img[:40]), oracle.encode(img[40:])])
    assert e.value.code == -1
    with pytest.raises(flic.FlicError):       # widths differ
        flic.splice_block_rows([oracle.encode(img[:64]), oracle.encode(img[64:, :100])])


@pytest.mark.parametrize("flags", [0x21, 0x41, 0x31, 0x51])
def test_layout_flags_peek_and_splice(flic, oracle, flags):
    """The two optional layouts (one bit stream per block; exact block sizes) are header flags: peek reports them,
    the combination is refused, and block-row parts splice into the full stream in either layout (host code)."""
    img = cases.gradient(300, 100, 3, 33)
    full = oracle.encode(img, flags)
    assert flic.peek(full)["flags"] == flags
    parts = [oracle.encode(img[:64], flags), oracle.encode(img[64:], flags)]
    assert np.array_equal(flic.splice_block_rows(parts), full)
    bad = full.copy(); bad[7] = 0x61      # ONE_STREAM | EXACT
    with pytest.raises(flic.FlicError) as e:
        flic.peek(bad)
    assert e.value.code == -3
    bad[7] = 0x81                           # reserved bit
    with pytest.raises(flic.FlicError):
        flic.peek(bad)
    with pytest.raises(flic.FlicError):     # parts of different layouts do not splice
        flic.splice_block_rows([oracle.encode(img[:64], flags), oracle.encode(img[64:], 0x01)])


def test_splice_plan(flic, oracle):
    """flic_splice_plan: where every part's directory entries and payload land in the spliced stream (what the ranks
    of the multi-GPU path address their sends with) — checked against an actual host splice."""
    img = cases.gradient(260, 200, 4, 34)
    cuts = [(0, 64), (64, 160), (160, 200)]
    parts = [oracle.encode(img[a:b]) for a, b in cuts]
    infos = [flic.peek(p) for p in parts]
    nbs, pws = [i["n_blocks"] for i in infos], [i["payload_words"] for i in infos]
    doff, poff, total = flic.splice_plan(nbs, pws)
    full = flic.splice_block_rows(parts)
    assert total == full.size
    base = 0
    for p, nb, pw, do, po in zip(parts, nbs, pws, doff, poff):
        d_part = np.frombuffer(p[32: 32 + 4 * nb].tobytes(), np.uint32)
        d_full = np.frombuffer(full[do: do + 4 * nb].tobytes(), np.uint32)
        assert np.array_equal(d_full, d_part + np.uint32(base))                          # rebased directory segment
        assert np.array_equal(full[po: po + 4 * pw], p[32 + 4 * (nb + 1): 32 + 4 * (nb + 1) + 4 * pw])  # payload in place
        base += pw
    lib = flic.load_library()
    import ctypes as C
    assert lib.flic_splice_plan(None, None, 0, None, None, None) == -1


def test_host_only_argument_checks(flic):
    """Entry points that can refuse without a device do so with FLIC_E_ARG, never a crash."""
    lib = flic.load_library()
    assert lib.flic_set_option(None, 1, 0) == -1
    assert lib.flic_wait(None, 0) == -1
    assert lib.flic_host_register(None, 0) == -1
    assert lib.flic_encode_submit(None, None, 0, 0, 0, 0, 0, None, 0, None) == -1
    assert lib.flic_decode_submit(None, None, None, 0, None, 0) == -1
    assert lib.flic_strerror(-8).decode().startswith("an operation submitted")
