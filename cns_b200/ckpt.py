"""Checkpoint loading for both files the reference ships.

* `checkpoints/cns_state_dict.pth`: plain OrderedDict of 43 fp32 tensors.
* `checkpoints/cns.pth`: `{"net": <pickled reference GraphVS module>}` - the format
  `cns/utils/trainer.py:138-162` writes.  Its pickle names classes under `cns.models.*`
  and `torch_geometric.nn.*`, neither of which exists here, so it is read with a
  remapping unpickler: every such class resolves to an inert stand-in (an `nn.Module`
  subclass where the original was one), the tensors are recovered through the ordinary
  `state_dict()` walk and loaded into `cns_b200.GraphVS`.  No reference code is run.
"""
import pickle
import types
from collections import OrderedDict

import torch
import torch.nn as nn

from .model import GraphVS

_REMAP_PREFIXES = ("cns.", "torch_geometric")
_stub_cache = {}


def _noop(*args, **kwargs):
    return None


def _walk_tensors(mod, prefix, out):
    """state_dict()-equivalent walk that only trusts `_parameters/_buffers/_modules`."""
    d = mod.__dict__
    for store in ("_parameters", "_buffers"):
        for k, v in (d.get(store) or {}).items():
            if isinstance(v, torch.Tensor):
                out[prefix + k] = v
    for k, sub in (d.get("_modules") or {}).items():
        if sub is not None:
            _walk_tensors(sub, prefix + k + ".", out)
    return out


def _stub(module: str, name: str):
    key = (module, name)
    if key not in _stub_cache:
        is_module = not ("inspector" in module or "typing" in module)
        if is_module:
            def _getattr(self, attr):
                # bound methods of the original class are pickled as getattr(obj, name)
                try:
                    return nn.Module.__getattr__(self, attr)
                except AttributeError:
                    if attr.startswith("__"):
                        raise
                    return _noop
            cls = type(name, (nn.Module,), {"__module__": module, "forward": _noop, "__getattr__": _getattr})
        else:
            def _setstate(self, state):
                if isinstance(state, dict):
                    self.__dict__.update(state)
                else:
                    self.__dict__["_state"] = state
            cls = type(name, (object,), {"__module__": module, "__init__": lambda self, *a, **k: None,
                                         "__setstate__": _setstate})
        _stub_cache[key] = cls
    return _stub_cache[key]


class _RemapUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith(_REMAP_PREFIXES):
            return _stub(module, name)
        return super().find_class(module, name)


_remap_pickle = types.ModuleType("cns_b200_remap_pickle")
_remap_pickle.Unpickler = _RemapUnpickler
_remap_pickle.load = lambda f, **kw: _RemapUnpickler(f, **kw).load()
_remap_pickle.__name__ = "pickle"


def read_state_dict(path: str, map_location="cpu") -> "OrderedDict[str, torch.Tensor]":
    """The 43-tensor state dict from either checkpoint format."""
    try:
        obj = torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        obj = torch.load(path, map_location=map_location, weights_only=False, pickle_module=_remap_pickle)
    if isinstance(obj, dict) and "net" in obj and not isinstance(obj["net"], torch.Tensor):
        obj = obj["net"]
    if isinstance(obj, nn.Module):
        obj = _walk_tensors(obj, "", OrderedDict())
    if not isinstance(obj, dict):
        raise ValueError("unrecognised checkpoint layout in %s: %s" % (path, type(obj)))
    return OrderedDict((k, v.detach().float()) for k, v in obj.items() if isinstance(v, torch.Tensor))


def load_model(path: str, device="cuda:0", precision: str = "fp32") -> GraphVS:
    """`GraphVSController.__init__` of the reference (benchmark/controller.py:11-21)."""
    net = GraphVS(2, 2, 128, regress_norm=True, precision=precision)
    net.load_state_dict(read_state_dict(path), strict=True)
    return net.to(device).eval()


def save_checkpoint(path, epoch, net: GraphVS, optimizers=None, score=None, is_best=False,
                    not_improve_epochs=0):
    """State-dict flavour of the reference's checkpoint dict (trainer.py:146-153)."""
    torch.save({
        "epoch": epoch, "net": net.state_dict(),
        "optimizers": {k: o.state_dict() for k, o in (optimizers or {}).items()},
        "score": score, "is_best": is_best, "not_improve_epochs": not_improve_epochs}, path)
