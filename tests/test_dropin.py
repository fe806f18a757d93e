"""CPU: host-side mirror of the reference interface (no kernels launched)."""
import hashlib

import numpy as np
import pytest
import torch


def _sha(sd):
    h = hashlib.sha256()
    for k in sorted(sd.keys()):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().numpy()).tobytes())
    return h.hexdigest()


def test_same_seed_gives_the_reference_weights(golden):
    import inf3dp
    torch.manual_seed(2048)  # exp_scripts/train_sdf.py:13
    nets = {}
    for name, outf in [("sdf", 1), ("slice", 1), ("lattice1", 1), ("lattice2", 1), ("quat", 4)]:
        nets[name] = inf3dp.SingleBVPNet(type="sine", in_features=3, out_features=outf, final_layer_factor=1)
    want = dict(zip(golden["sha_names"].tolist(), golden["sha_values"].tolist()))
    for name, net in nets.items():
        assert _sha(net.state_dict()) == want[name], name
    assert list(nets["sdf"].state_dict().keys()) == [f"net.net.{i}.0.{p}" for i in range(5) for p in ("weight", "bias")]
    assert nets["sdf"].state_dict()["net.net.0.0.weight"].shape == (256, 3)
    assert nets["quat"].state_dict()["net.net.4.0.weight"].shape == (4, 256)


def test_reference_checkpoint_keys_load(golden):
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    sd = {k[len("sdf/"):]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("sdf/net.")}
    net.load_state_dict(sd)  # strict
    lvl = inf3dp.MLPClassifier(num_classes=8)
    lsd = {k[len("level/"):]: torch.from_numpy(np.asarray(golden[k])) for k in golden.files if k.startswith("level/network")}
    lvl.load_state_dict(lsd)


def test_no_cpu_fallback():
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    with pytest.raises(RuntimeError, match="CUDA"):
        net({"coords": torch.zeros(4, 3)})
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        inf3dp.extract_isocontours(torch.zeros(3, 3), torch.zeros(1, 3, dtype=torch.int32), torch.zeros(3), torch.zeros(1))
    with pytest.raises(NotImplementedError):
        inf3dp.SingleBVPNet(type="sine", in_features=3, mode="nerf")


def test_qupkg_and_dropin_modules_resolve():
    import importlib
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "inf-3dp_b200", "dropin"))
    try:
        q = importlib.import_module("qupkg.extract_isocontours")
        assert callable(q.extract_isocontours) and callable(q.get_fields_intersection_inter_slice)
        m = importlib.import_module("modules")
        assert m.SingleBVPNet.__module__ == "inf3dp.siren"
        d = importlib.import_module("diff_operators")
        assert callable(d.gradient) and callable(d.hessian3d)
    finally:
        sys.path.pop(0)
        for n in ("modules", "diff_operators"):
            sys.modules.pop(n, None)


def test_dropin_is_loudly_inference_only():
    """a reference training loop on the drop-in must not silently 'train' (weights get no gradients)."""
    import sys
    import types
    import inf3dp
    net = inf3dp.SingleBVPNet(type="sine", in_features=3)
    assert not net.training and not any(p.requires_grad for p in net.parameters())
    net.eval()
    with pytest.raises(RuntimeError, match="inference-only"):
        net.train()
    holder = torch.nn.Sequential(net)
    with pytest.raises(RuntimeError, match="inference-only"):
        holder.train()
    lvl = inf3dp.MLPClassifier(num_classes=3)
    with pytest.raises(RuntimeError, match="inference-only"):
        lvl.train()
    net.net.net[0][0].weight.requires_grad_(True)
    with pytest.raises(RuntimeError, match="requires_grad"):
        net({"coords": torch.zeros(2, 3)})
    fake = types.ModuleType("training")
    fake.train = lambda *a, **k: None
    sys.modules["training"] = fake
    sys.modules.setdefault("modules", types.ModuleType("modules"))
    try:
        with pytest.raises(RuntimeError, match="training script"):
            inf3dp.patch_reference()
        assert inf3dp.patch_reference(force=True) is True
    finally:
        sys.modules.pop("training", None)
        sys.modules.pop("modules", None)
