"""B200-native block codec engine — Python host layer over the C ABI.

STATUS: the reference (wouter-rombouts/fast-losless-image-compression-format)
is behind a licensing gate (see LICENSING.md at the repo root); its source was
not read.  This package therefore drives the engine on the *provisional* FLP0
bitstream specified in DESIGN.md — it is NOT bit-compatible with the reference.

The directory name carries hyphens (the spec's package name), so import it with
``importlib.import_module("fast-losless-image-compression-format_b200")`` or
through the ``flic_b200`` shim at the repo root.
"""
from .build import build_library, library_path  # noqa: F401
from .codec import (  # noqa: F401
    BLOCK_H,
    BLOCK_W,
    FLAG_EXACT,
    FLAG_ONE_STREAM,
    FLAG_SUBGREEN,
    MAX_CODE_LEN,
    OP_DECODE,
    OP_ENCODE,
    PRED_LEFT,
    Codec,
    FlicError,
    load_library,
    max_stream_bytes,
    peek,
    splice_block_rows,
    splice_plan,
)
from . import sharding, workloads  # noqa: F401
