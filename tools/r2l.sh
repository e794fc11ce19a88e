for d in 1 2 3 4 6 8; do echo "depth $d"; FLIC_PIPE_DEPTH=$d python tools/e2e_probe.py 2>&1 | sed -n '4,6p'; done
