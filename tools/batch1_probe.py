"""Runs a few batch-1 forwards (device-resident input) so that ncu can capture one of them:
  ncu --set full --clock-control none -s 30 -c 12 python tools/batch1_probe.py [tf32|bf16] [forwards]
DNNB_LATENCY_MODE is taken from the environment."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from configs_bench import device_forward_times, make_net  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
forwards = int(sys.argv[2]) if len(sys.argv) > 2 else 5
net = make_net(1, prec)
ts = device_forward_times(net, 1, forwards, warmup=3)
print("batch-1", prec, "latency_mode", os.environ.get("DNNB_LATENCY_MODE", "0"), "us:", [round(t * 1e3, 1) for t in ts])
print("launches per forward", net.launches_per_forward() if hasattr(net, "launches_per_forward") else "?")
if os.environ.get("PROBE_PREFIX"):
    # in-situ share of each block: median latency of forwards cut after k blocks, k = 1..9
    import statistics
    prev = 0.0
    for k in range(1, net.num_blocks + 1):
        net.debug_block_limit(k)
        med = statistics.median(device_forward_times(net, 1, forwards, warmup=3)) * 1e3
        print("blocks 0..%d: %.1f us  (+%.1f)  %s" % (k - 1, med, med - prev, net.block_info(k - 1)["kernel"]))
        prev = med
    net.debug_block_limit(0)
net.close()
torch.cuda.synchronize()
