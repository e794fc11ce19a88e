"""C4 block-row split alone (bench.py's c4_split), both transports: torchrun ... tools/c4_probe.py [height] [steps]
Prints one JSON line per transport on rank 0."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import flic_b200

h = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
D = bench.Dist()
flic_b200.build_library()
codec = flic_b200.Codec(D.local)
for mode in ("nccl", "peer"):
    args = argparse.Namespace(flags=1, c4_height=h, c4_mode=mode)
    res = bench.c4_split(D, codec, args, steps, 3)
    if D.rank == 0:
        print(json.dumps({k: res[k] for k in ("transport", "value_GBps", "encode_GBps", "decode_GBps", "ms_per_step",
                                                "compressed_ratio", "round_trip_verified_on_every_rank", "peer_memory_unavailable",
                                                "gpu_launches_this_rank")}), flush=True)
codec.close()
D.close()
