"""Checkpoint reading and writing for the formats the reference uses.

* `checkpoints/cns_state_dict.pth`: plain OrderedDict of 43 fp32 tensors.
* `checkpoints/cns.pth` and every `checkpoint_{last,best}.pth` the reference's Trainer writes
  (`cns/utils/trainer.py:138-162`): `{"epoch", "net": <pickled GraphVS module>, "optimizers": {name: AdamW object},
  "score", "is_best", "not_improve_epochs"}`.  The pickle names classes under `cns.models.*` and
  `torch_geometric.nn.*`, neither of which exists here, so it is READ with a remapping unpickler: every such class
  resolves to an inert stand-in (an `nn.Module` subclass where the original was one), everything else must be on a
  short allowlist (torch tensors / modules / optimizers, a few stdlib containers) - no reference code runs and no
  other global is resolved.  The tensors come out through a `state_dict()`-equivalent walk.
* WRITING the same format (`save_checkpoint`, signature of trainer.py:138): the object graph of the shipped
  `cns.pth` is the template - its parameters are replaced by ours and it is pickled back under the ORIGINAL class
  paths (the inverse of the remapping), with real `torch.optim.AdamW` objects bound to those parameters.  The
  unmodified reference (`train_cns.py:63`, `Trainer.load_from_checkpoint` trainer.py:333-348) loads the result.
"""
import os
import pickle
import types
from collections import OrderedDict

import torch
import torch.nn as nn

from .model import GraphVS

_REMAP_PREFIXES = ("cns.", "torch_geometric")
_TEMPLATE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "checkpoints", "cns.pth")
_stub_cache = {}


class _MethodRef(object):
    """`getattr(obj, name)` of a stand-in object (bound methods of the original class are pickled that way, e.g.
    PyG Linear's `_lazy_load_hook`): inert when called, pickled back as the same getattr."""

    def __init__(self, obj, name):
        self.obj, self.name = obj, name

    def __call__(self, *args, **kwargs):
        return None

    def __reduce__(self):
        return getattr, (self.obj, self.name)


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
                try:
                    return nn.Module.__getattr__(self, attr)
                except AttributeError:
                    if attr.startswith("__"):
                        raise
                    return _MethodRef(self, attr)
            cls = type(name, (nn.Module,), {"__module__": module, "forward": lambda self, *a, **k: None,
                                            "__getattr__": _getattr})
        else:
            def _setstate(self, state):
                if isinstance(state, dict):
                    self.__dict__.update(state)
                else:
                    self.__dict__["_state"] = state
            cls = type(name, (object,), {"__module__": module, "__init__": lambda self, *a, **k: None,
                                         "__setstate__": _setstate})
        cls._cns_pickle_path = key
        _stub_cache[key] = cls
    return _stub_cache[key]


def _safe_getattr(obj, name):
    if isinstance(obj, nn.Module) or type(obj) in _stub_cache.values():
        return getattr(obj, name)
    raise pickle.UnpicklingError("getattr on %s refused while reading a checkpoint" % type(obj).__name__)


def _safe_type(*args):
    if len(args) != 1:
        raise pickle.UnpicklingError("type() with %d arguments refused while reading a checkpoint" % len(args))
    return type(args[0])


# what a Trainer checkpoint legitimately references besides the remapped classes
_ALLOWED_MODULE_PREFIXES = ("torch._utils", "torch.nn.modules.", "torch.optim.", "torch.utils.hooks", "torch.storage",
                            "torch._tensor", "torch.nn.parameter")
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"), ("inspect", "Parameter"),
    ("inspect", "_ParameterKind"), ("inspect", "_empty"), ("typing", "Union"), ("typing", "Optional"),
    ("_operator", "getitem"), ("builtins", "set"), ("builtins", "int"), ("builtins", "dict"), ("builtins", "list"),
    ("builtins", "tuple"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "frozenset"),
    ("builtins", "str"), ("builtins", "complex"), ("builtins", "slice"), ("builtins", "range"),
    ("torch", "Tensor"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
}


class _RemapUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith(_REMAP_PREFIXES):
            return _stub(module, name)
        if module == "__builtin__":          # protocol-2 spelling (python 2 names)
            module, name = "builtins", {"long": "int", "unicode": "str"}.get(name, name)
        if module == "builtins" and name == "getattr":
            return _safe_getattr
        if module == "builtins" and name == "type":
            return _safe_type
        if (module, name) in _ALLOWED_GLOBALS or module.startswith(_ALLOWED_MODULE_PREFIXES) or \
                (module == "torch" and (name.endswith("Storage") or name in ("float32", "float64", "int64"))):
            return super().find_class(module, name)
        raise pickle.UnpicklingError("global %s.%s is not on the checkpoint allowlist" % (module, name))


_remap_pickle = types.ModuleType("cns_b200_remap_pickle")
_remap_pickle.Unpickler = _RemapUnpickler
_remap_pickle.load = lambda f, **kw: _RemapUnpickler(f, **kw).load()
_remap_pickle.__name__ = "pickle"


class _ClassPathPickler(pickle._Pickler):
    """Writes the stand-in classes under the import paths they stand for (pure-python pickler: `save_global` is
    virtual there), everything else as usual."""

    def save_global(self, obj, name=None):
        path = obj.__dict__.get("_cns_pickle_path") if isinstance(obj, type) else None
        if path is not None:
            self.write(pickle.GLOBAL + path[0].encode("utf-8") + b"\n" + path[1].encode("utf-8") + b"\n")
            self.memoize(obj)
            return
        super().save_global(obj, name)

    dispatch = dict(pickle._Pickler.dispatch)
    dispatch[type] = lambda self, obj: _ClassPathPickler._save_type(self, obj)

    def _save_type(self, obj):
        if obj.__dict__.get("_cns_pickle_path") is not None:
            return self.save_global(obj)
        return pickle._Pickler.save_type(self, obj)


_classpath_pickle = types.ModuleType("cns_b200_classpath_pickle")
_classpath_pickle.Pickler = _ClassPathPickler
_classpath_pickle.__name__ = "pickle"


def _load_raw(path, map_location="cpu"):
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception as e:      # only a pickled reference module takes the remapping route
        msg = str(e)
        if "cns." not in msg and "torch_geometric" not in msg and "Unsupported" not in msg and "weights_only" not in msg.lower():
            raise
    return torch.load(path, map_location=map_location, weights_only=False, pickle_module=_remap_pickle)


def read_state_dict(path: str, map_location="cpu") -> "OrderedDict[str, torch.Tensor]":
    """The 43-tensor state dict from either checkpoint format."""
    obj = _load_raw(path, map_location)
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


def _reference_module(net: GraphVS, template: str = None):
    """The reference's pickled `GraphVS` object graph (from the shipped `cns.pth`) carrying `net`'s parameters.
    Returns (module stand-in, {state-dict key: its Parameter})."""
    tpl = torch.load(template or _TEMPLATE, map_location="cpu", weights_only=False, pickle_module=_remap_pickle)
    mod = tpl["net"] if isinstance(tpl, dict) else tpl
    named = _walk_tensors(mod, "", OrderedDict())
    sd = net.state_dict()
    if list(named.keys()) != list(sd.keys()):
        raise ValueError("template checkpoint and model disagree on the parameter names")
    for k, p in named.items():
        if tuple(p.shape) != tuple(sd[k].shape):
            raise ValueError("template checkpoint is for another configuration (%s: %s vs %s)" % (
                k, tuple(p.shape), tuple(sd[k].shape)))
        p.data = sd[k].detach().to("cpu", torch.float32).clone()
        p.grad = None
    mod.__dict__["training"] = False
    return mod, named


def _reference_optimizers(net: GraphVS, named, optimizers=None, fused_opt=None):
    """`{"0": AdamW(decayed weights), "1": AdamW(the rest)}` (train_cns.py:65-69) over the template's Parameters,
    carrying the moments of `optimizers` (torch) or `fused_opt` (FusedClipAdamW: flat m / v / t)."""
    by_id = {id(p): k for k, p in net.named_parameters()}
    decay, no_decay = net.get_parameter_groups()
    groups = {"0": [named[by_id[id(p)]] for p in decay], "1": [named[by_id[id(p)]] for p in no_decay]}
    out = {}
    for key, params in groups.items():
        src = (optimizers or {}).get(key)
        hyper = dict(lr=5e-4, weight_decay=1e-4 if key == "0" else 0.0)
        if src is not None:
            g = src.param_groups[0]
            hyper = dict(lr=g["lr"], weight_decay=g["weight_decay"], betas=tuple(g["betas"]), eps=g["eps"])
        elif fused_opt is not None:
            hyper = dict(lr=fused_opt.lr, weight_decay=fused_opt.wd if key == "0" else 0.0, betas=tuple(fused_opt.betas),
                         eps=fused_opt.eps)
        opt = torch.optim.AdamW(params, **hyper)
        if src is not None:
            mine = [p for g in src.param_groups for p in g["params"]]
            for p in mine:
                st = src.state.get(p)
                if st:
                    opt.state[named[by_id[id(p)]]] = {
                        k: (v.detach().to("cpu").clone() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
        out[key] = opt
    if fused_opt is not None and fused_opt.t > 0:
        off = 0
        for p in fused_opt.params:
            n = p.numel()
            tp = named[by_id[id(p)]]
            for opt in out.values():
                if any(tp is q for q in opt.param_groups[0]["params"]):
                    opt.state[tp] = {"step": torch.tensor(float(fused_opt.t)),
                                     "exp_avg": fused_opt.m[off:off + n].view_as(p).detach().cpu().clone(),
                                     "exp_avg_sq": fused_opt.v[off:off + n].view_as(p).detach().cpu().clone()}
            off += n
    return out


def save_checkpoint(epoch, net: GraphVS, optimizers, score, is_best, not_improve_epoches, save_folder,
                    fused_opt=None, template: str = None):
    """`cns/utils/trainer.py:138-162`: writes `checkpoint_last.pth` (and `checkpoint_best.pth` when `is_best`) holding
    `{"epoch", "net": module, "optimizers": {name: AdamW}, "score", "is_best", "not_improve_epochs"}` - the module and
    the optimizers pickled under the reference's class paths, so `torch.load(path)["net"]` in the unmodified
    reference (train_cns.py:63) returns its own `GraphVS` with these weights."""
    mod, named = _reference_module(net, template)
    checkpoint = {
        "epoch": epoch, "net": mod,
        "optimizers": _reference_optimizers(net, named, optimizers, fused_opt),
        "score": score, "is_best": is_best, "not_improve_epochs": not_improve_epoches}
    os.makedirs(save_folder, exist_ok=True)
    paths = [os.path.join(save_folder, "checkpoint_last.pth")]
    if is_best:
        paths.append(os.path.join(save_folder, "checkpoint_best.pth"))
    for path in paths:
        torch.save(checkpoint, path, pickle_module=_classpath_pickle, pickle_protocol=2)
    return paths


def load_checkpoint(path: str, net: GraphVS = None, optimizers=None, fused_opt=None, map_location="cpu"):
    """`Trainer.load_from_checkpoint` (trainer.py:333-348) for files written by the reference or by
    `save_checkpoint`: returns `{"epoch", "state_dict", "optimizer_state": {name: {param key: state}}, "score",
    "is_best", "not_improve_epochs"}`; when given, `net` receives the weights and `optimizers` / `fused_opt` the
    AdamW moments (matched by parameter name)."""
    obj = _load_raw(path, map_location)
    if not isinstance(obj, dict) or "net" not in obj:
        raise ValueError("%s is not a Trainer checkpoint" % path)
    mod = obj["net"]
    named = _walk_tensors(mod, "", OrderedDict()) if isinstance(mod, nn.Module) else OrderedDict(mod)
    sd = OrderedDict((k, v.detach().float()) for k, v in named.items())
    ids = {id(p): k for k, p in named.items()}
    opt_state = {}
    for key, opt in (obj.get("optimizers") or {}).items():
        if isinstance(opt, torch.optim.Optimizer):
            opt_state[key] = {ids[id(p)]: st for p, st in opt.state.items() if id(p) in ids}
            opt_state[key]["__hyper__"] = {k: v for k, v in opt.param_groups[0].items() if k != "params"}
    if net is not None:
        net.load_state_dict(sd, strict=True)
        mine = dict(net.named_parameters())
        for key, states in opt_state.items():
            if optimizers is not None and key in optimizers:
                for g in optimizers[key].param_groups:
                    g.update({k: v for k, v in states.get("__hyper__", {}).items()
                              if k in ("lr", "betas", "eps", "weight_decay", "amsgrad")})
                for name, st in states.items():
                    if name != "__hyper__":
                        optimizers[key].state[mine[name]] = {
                            k: (v.to(mine[name].device) if isinstance(v, torch.Tensor) and v.dim() > 0 else v)
                            for k, v in st.items()}
        if fused_opt is not None:
            off, names = 0, {id(p): k for k, p in mine.items()}
            with torch.no_grad():
                for p in fused_opt.params:
                    n, name = p.numel(), names[id(p)]
                    fused_opt.flat[off:off + n].copy_(sd[name].reshape(-1))
                    for states in opt_state.values():
                        st = states.get(name)
                        if st:
                            fused_opt.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                            fused_opt.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                            fused_opt.t = max(fused_opt.t, int(float(st["step"])))
                    off += n
    return {"epoch": obj.get("epoch"), "state_dict": sd, "optimizer_state": opt_state, "score": obj.get("score"),
            "is_best": obj.get("is_best"), "not_improve_epochs": obj.get("not_improve_epochs")}
