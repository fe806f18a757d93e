"""Scratch: second-order jet (1 M points) and the fused collision sweep (65 536 waypoints x 64) for ncu."""
import copy, math, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200"))
import bench, inf3dp
dev = torch.device("cuda:0")
net = inf3dp.SingleBVPNet(type="sine", in_features=3)
net.load_state_dict({f"net.net.{i}.0.{k}": torch.from_numpy(np.asarray(v)) for i, (W, b) in enumerate(bench.golden_layers())
                     for k, v in (("weight", W), ("bias", b))})
net.to(dev)
x = torch.rand(1 << 20, 3, device=dev) * 2 - 1
for _ in range(2):
    y, J, H = net.query(x, order=2)
    inf3dp.principal_curvatures(J[:, 0], H[:, 0])
torch.manual_seed(2048)
nets = [inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=o) for o in (1, 1, 1, 1, 4)]
slice_net, quat_net = nets[1].to(dev), nets[4].to(dev)
level = inf3dp.MLPClassifier(num_classes=8).to(dev)
sdf = copy.deepcopy(net)
with torch.no_grad():
    sdf.net.net[4][0].bias -= 0.06
n = 65536
g = torch.Generator().manual_seed(4)
u, v = torch.rand(n, generator=g) * 2 * math.pi, torch.rand(n, generator=g) * 2 * math.pi
way = torch.stack([(0.6 + 0.3 * torch.cos(v)) * torch.cos(u), (0.6 + 0.3 * torch.cos(v)) * torch.sin(u), 0.3 * torch.sin(v)], 1).to(dev)
levels = ((way[:, 2] + 0.3) / 0.6 * 8).clamp(0, 7).long()
maxsl = (torch.rand((n, 8), generator=g) * 0.1 - 0.05).to(dev)
t = torch.linspace(0, 1, 64)
nozzle = (torch.stack([0.5 * t * torch.cos(t * 40), 0.5 * t * torch.sin(t * 40), 0.2 + 4.0 * t], 1) / 150.0 * 20.0).to(dev)
for _ in range(2):
    inf3dp.quat_collision_sweep(way, levels, maxsl, nozzle, quat_net, sdf, slice_net, level, 8, -0.9)
torch.cuda.synchronize()
print("ok")
