"""Per-launch DRAM traffic of every kernel in an .ncu-rep (ncu --set full), as the JSON bench.py reads for
`roofline.traffic`.  usage: python tools/ncu_traffic.py gpurun_out/r02_full_c2x64.ncu-rep C2x64 RAW_BYTES RATIO > profiles/r02_traffic_c2x64.json"""
import csv
import io
import json
import subprocess
import sys

SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main(path, workload, raw_bytes, ratio):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")

    def val(r, name):
        i = hdr.index(name)
        return float(r[i].replace(",", "")) * SCALE[units[i]]

    kernels = {}
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").split("<")[0].strip()
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        kernels[name] = {"dram_read_bytes": int(rd), "dram_write_bytes": int(wr), "traffic_bytes": int(rd + wr),
                         "ncu_duration_ms": round(val(r, "gpu__time_duration.sum"), 6)}
    enc = [k for k in kernels if k not in ("k_decode", "k_decode_one")]
    doc = {
        "workload": workload,
        "source": f"ncu --set full --clock-control none, one launch of each kernel of a step ({path}; summary in "
                  f"profiles/{path.split('/')[-1].replace('full', 'ncu_full').replace('.ncu-rep', '.txt')})",
        "raw_bytes": raw_bytes,
        "note": "traffic_bytes = dram__bytes_read.sum + dram__bytes_write.sum per launch; algorithmic bytes: encode path "
                f"and decoder (1+r)*N = {int(raw_bytes * (1 + ratio))}",
        "encode_path_traffic_bytes": sum(kernels[k]["traffic_bytes"] for k in enc),
        "kernels": kernels,
    }
    json.dump(doc, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4]))
