timeout 300 python -m pytest tests -m gpu -x -q -k "latency" 2>&1 | tail -3
run() { echo "== $*"; env DNNB_LATENCY_MODE=1 PROBE_PREFIX=1 "$@" timeout 200 python tools/batch1_probe.py tf32 40 2>&1 | grep "blocks 0" | awk '{printf "%s %s | ", $3, $5} END {print ""}'; }
run X=1
run X=2
