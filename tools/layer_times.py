"""Per-block times (each block alone, L2 flushed) at one batch size: python tools/layer_times.py 512 tf32"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from configs_bench import make_net  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
net = make_net(batch, prec)
print(batch, prec, " ".join("%.4f" % t for t in net.time_blocks(3, 10)), "ms; tiles", [net.block_info(i)["tiles"] for i in range(net.num_blocks)])
net.close()
