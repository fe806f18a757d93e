"""Scratch: one 2 M-point value+gradient query on the reverse engine for ncu source-level sampling."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import inf3dp
torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
x = torch.rand(1 << 21, 3, device="cuda") * 2 - 1
for _ in range(3):
    net.query(x, order=1, mode="tc_reverse")
torch.cuda.synchronize()
print("ok")
