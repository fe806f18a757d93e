"""Single launch of the tcgen05 engine for ncu (small, fixed workload)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
mode = sys.argv[1] if len(sys.argv) > 1 else "tc_precise"
order = int(sys.argv[2]) if len(sys.argv) > 2 else 1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 32 * 64
torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
x = torch.rand(n, 3, device="cuda") * 2 - 1
for _ in range(3):
    y, J, _ = net.query(x, order=order, mode=mode)
torch.cuda.synchronize()
print("ok", float(y.sum()))
