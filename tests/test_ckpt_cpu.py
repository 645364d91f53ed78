"""CPU: checkpoint formats.  Reading both shipped files is covered in test_host_cpu.py; here the
Trainer-format writer (`cns/utils/trainer.py:138-162`), its reader, the unpickling allowlist and - where the
reference tree is present - the unmodified reference loading what we wrote."""
import io
import os
import pickle
import types

import numpy as np
import pytest
import torch

import cns_oracle
import ref_import
from cns_b200 import GraphGenerator, GraphVS, ckpt, synth
from cns_b200.train import make_optimizers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def _trained_net(seed=1):
    torch.manual_seed(seed)
    net = GraphVS(2, 2, 128)
    opts = make_optimizers(net, lr=3e-4, weight_decay=2e-4)
    for _ in range(2):
        for p in net.parameters():
            p.grad = torch.randn_like(p) * 1e-2
        for o in opts.values():
            o.step()
    return net, opts


def test_trainer_checkpoint_round_trip(tmp_path):
    net, opts = _trained_net()
    paths = ckpt.save_checkpoint(7, net, opts, 0.25, True, 3, str(tmp_path))
    assert [os.path.basename(p) for p in paths] == ["checkpoint_last.pth", "checkpoint_best.pth"]
    assert len(ckpt.save_checkpoint(8, net, opts, 0.3, False, 4, str(tmp_path))) == 1
    # the generic loaders see the weights
    sd = ckpt.read_state_dict(paths[1])
    assert list(sd.keys()) == list(net.state_dict().keys())
    assert all(torch.equal(sd[k], v) for k, v in net.state_dict().items())
    # resume: weights, AdamW moments / step counts / hyper-parameters, bookkeeping
    net2 = GraphVS(2, 2, 128)
    opts2 = make_optimizers(net2)
    info = ckpt.load_checkpoint(paths[1], net2, opts2)
    assert (info["epoch"], info["score"], info["is_best"], info["not_improve_epochs"]) == (7, 0.25, True, 3)
    assert info["optimizer_state"]["0"]["__hyper__"]["lr"] == 3e-4
    assert info["optimizer_state"]["0"]["__hyper__"]["weight_decay"] == 2e-4
    assert info["optimizer_state"]["1"]["__hyper__"]["weight_decay"] == 0
    for key in ("0", "1"):
        a = {n: opts[key].state[p] for n, p in net.named_parameters() if p in opts[key].state}
        b = {n: opts2[key].state[p] for n, p in net2.named_parameters() if p in opts2[key].state}
        assert a.keys() == b.keys() and len(a) > 0
        for n in a:
            assert torch.equal(a[n]["exp_avg"], b[n]["exp_avg"]) and torch.equal(a[n]["exp_avg_sq"], b[n]["exp_avg_sq"])
            assert float(a[n]["step"]) == float(b[n]["step"]) == 2.0
    # same next step from the restored state
    for (_, p), (_, q) in zip(net.named_parameters(), net2.named_parameters()):
        g = torch.randn_like(p)
        p.grad, q.grad = g, g.clone()
    for o in list(opts.values()) + list(opts2.values()):
        o.step()
    assert all(torch.equal(p, q) for p, q in zip(net.parameters(), net2.parameters()))


def test_fused_optimizer_state_is_saved_and_restored(tmp_path):
    """FusedClipAdamW keeps flat m / v / t: they must survive a checkpoint (as ordinary AdamW state)."""
    net, _ = _trained_net(2)
    params = list(net.parameters())
    n = sum(p.numel() for p in params)
    fused = types.SimpleNamespace(params=params, m=torch.randn(n), v=torch.rand(n), t=11, lr=2e-4, wd=1e-4,
                                  betas=(0.9, 0.999), eps=1e-8, flat=torch.zeros(n))
    path = ckpt.save_checkpoint(1, net, {}, 1.0, False, 0, str(tmp_path), fused_opt=fused)[0]
    net2 = GraphVS(2, 2, 128)
    fused2 = types.SimpleNamespace(params=list(net2.parameters()), m=torch.zeros(n), v=torch.zeros(n), t=0,
                                   flat=torch.zeros(n))
    ckpt.load_checkpoint(path, net2, fused_opt=fused2)
    assert fused2.t == 11 and torch.equal(fused2.m, fused.m) and torch.equal(fused2.v, fused.v)
    assert torch.equal(fused2.flat, torch.cat([p.detach().reshape(-1) for p in params]))


def test_unpickler_refuses_globals_off_the_allowlist(tmp_path):
    class Evil(object):
        def __reduce__(self):
            return os.system, ("echo pwned > %s" % (tmp_path / "pwned"),)

    path = str(tmp_path / "evil.pth")
    torch.save({"net": Evil()}, path)
    with pytest.raises(Exception) as err:
        ckpt.read_state_dict(path)
    assert not (tmp_path / "pwned").exists()
    assert "allowlist" in str(err.value) or "Unsupported" in str(err.value) or "weights_only" in str(err.value).lower()
    # getattr is only honoured on module stand-ins
    with pytest.raises(pickle.UnpicklingError):
        ckpt._safe_getattr(os, "system")


@needs_ref
def test_unmodified_reference_loads_our_checkpoint(tmp_path, monkeypatch):
    """train_cns.py:63 `torch.load(args.load)["net"]` and Trainer.load_from_checkpoint (trainer.py:333-348) of the
    unmodified reference (on the PyG shim) on a file we wrote: its own GraphVS class comes back and computes the
    same forward as the oracle with our weights."""
    gv, gg, fn = ref_import.load()
    from torch_geometric.data import Batch as RefBatch
    net, opts = _trained_net(3)
    path = ckpt.save_checkpoint(5, net, opts, 0.5, False, 2, str(tmp_path))[0]
    loaded = torch.load(path, map_location="cpu", weights_only=False)
    ref_net = loaded["net"]
    assert type(ref_net) is gv.GraphVS and isinstance(loaded["optimizers"]["0"], torch.optim.AdamW)
    assert list(ref_net.state_dict().keys()) == list(net.state_dict().keys())
    # the optimizers are bound to the loaded module's own parameters (so the reference can resume training)
    ids = {id(p) for p in ref_net.parameters()}
    assert all(id(p) in ids for o in loaded["optimizers"].values() for g in o.param_groups for p in g["params"])
    decay, no_decay = ref_net.get_parameter_groups()
    assert {id(p) for p in decay} == {id(p) for p in loaded["optimizers"]["0"].param_groups[0]["params"]}
    rng = np.random.default_rng(5)
    frames, mine = [], []
    for n in (20, 90):
        pts = synth.project_norm(synth.sample_target_pose(rng), synth.sample_points(rng, n))
        np.random.seed(n)
        frames.append(gg.GraphGenerator(pts).get_data(pts + 0.01, missing_node_indices=np.array([1, 5])))
        np.random.seed(n)
        mine.append(GraphGenerator(pts).get_data(pts + 0.01, missing_node_indices=np.array([1, 5])))
    ref_net.eval()
    with torch.no_grad():
        rv, rn, rh = ref_net(RefBatch.from_data_list(frames), None)
    from cns_b200 import Batch
    ov, on, oh = cns_oracle.forward(net.state_dict(), Batch.from_data_list(mine), None)
    assert float((rv - ov).abs().max()) < 2e-6 and float((rn - on).abs().max()) < 2e-6 and float((rh - oh).abs().max()) < 2e-6
    try:
        from cns.utils.trainer import Trainer
    except Exception as e:                      # tensorboard / tqdm missing in this image
        pytest.skip("reference Trainer not importable here: %s" % e)
    # the reference calls torch.load(path, map_location) and predates torch 2.6's weights_only=True default
    real_load = torch.load
    monkeypatch.setattr(torch, "load", lambda f, map_location=None, **kw: real_load(f, map_location, weights_only=False))
    tr = Trainer.load_from_checkpoint(path, map_location="cpu")
    assert tr.start_epoch == 5 and tr.not_improve_epochs == 2 and type(tr.net) is gv.GraphVS
