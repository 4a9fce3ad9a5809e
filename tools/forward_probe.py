"""Runs a few device-resident forwards at one batch size so that ncu can capture one whole forward:
  ncu --set full --clock-control none -k regex:'conv0_kernel|conv_rowpair|conv_umma|conv_pair|maxpool' \\
      -s L -c L python tools/forward_probe.py 512 tf32        (L = launches per forward, printed below)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

from configs_bench import make_net  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32"
forwards = int(sys.argv[3]) if len(sys.argv) > 3 else 3
net = make_net(batch, prec)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
x = torch.rand((batch, 416, 416, 3), device=dev)
y = torch.empty((batch, 13, 13, 125), device=dev)
torch.cuda.synchronize()
for _ in range(forwards):
    net.forward_device(x.data_ptr(), y.data_ptr(), stream.cuda_stream)
stream.synchronize()
print("batch", batch, prec, "launches per forward", net.launches_per_forward(),
      [net.block_info(i)["kernel"] for i in range(net.num_blocks)])
net.close()
