import sys; sys.path.insert(0, "inf-3dp_b200"); sys.path.insert(0, ".")
import numpy as np, torch, inf3dp
from oracle import siren_oracle as so
for gain in (1.0, 3.0, 8.0, 16.0):
    torch.manual_seed(5)
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with torch.no_grad():
        for i in (1, 2, 3):
            net.net.net[i][0].weight.mul_(gain)
    net.cuda()
    layers = so.params_from_state_dict({k: v.detach().cpu().numpy() for k, v in net.state_dict().items()})
    g = torch.Generator().manual_seed(1)
    x = (torch.rand(4096, 3, generator=g) * 2 - 1).cuda()
    yo, Jo, Ho = so.siren_jet(layers, x.cpu().numpy().astype(np.float64), order=2)
    out = {}
    for eng in ("simt", "tc_reverse", "tc_precise"):
        y, J, H = net.query(x, order=2, mode=eng)
        fin = bool(torch.isfinite(H).all() and torch.isfinite(J).all() and torch.isfinite(y).all())
        Hn = H.cpu().numpy()[:, 0]
        rel = np.linalg.norm((Hn - Ho[:, 0]).reshape(len(Hn), -1), axis=1) / np.linalg.norm(Ho[:, 0].reshape(len(Hn), -1), axis=1)
        out[eng] = (fin, float(np.nanmax(rel)), float(np.nanmedian(rel)))
    print(gain, out, "|H| max", float(np.abs(Ho).max()))
