import os, sys
import numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/inf-3dp_b200')
import inf3dp
from oracle import siren_oracle as so
for gain in (1.0, 2.0, 3.0, 8.0):
    torch.manual_seed(5)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with torch.no_grad():
        for i in (1, 2, 3): net.net.net[i][0].weight.mul_(gain)
    net.cuda()
    layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(8192, 3, generator=g) * 2 - 1).cuda()
    yo, Jo, _ = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=1)
    for eng in ("tc_reverse", "tc_precise", "simt"):
        y, J, _ = net.query(x, order=1, mode=eng)
        ang = so.angle_deg(J.cpu().numpy()[:, 0], Jo[:, 0])
        print(f"gain {gain} {eng:10s}: finite={bool(torch.isfinite(J).all())} |df|max={np.abs(y.cpu().numpy()-yo).max():.2e} angle max={ang.max():.4f} med={np.median(ang):.5f}  |grad| med={np.median(np.linalg.norm(Jo[:,0],axis=1)):.2e}")
