import sys; sys.path.insert(0, "inf-3dp_b200")
import torch, inf3dp
torch.manual_seed(2048)
net = inf3dp.SingleBVPNet(type="sine", in_features=3).cuda()
for n in (1, 127, 128, 129, 257, 100003, 1 << 22):
    x = torch.rand(n, 3, device="cuda") * 2 - 1
    ya, _, _ = net.query(x, order=0)
    yt, _, _ = net.query(x, order=0, mode="tc_precise")
    y1, _, _ = net.query(x, order=1)
    print(n, float((ya - yt).abs().max()), float((ya - y1).abs().max()), bool(torch.isfinite(ya).all()))
x = torch.rand(1 << 22, 3, device="cuda") * 2 - 1
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
for m in ("auto", "tc_precise"):
    ms = timeit(lambda: net.query(x, order=0, mode=m))
    print(m, ms, (1 << 22) / ms / 1e3, "Mpts/s")
