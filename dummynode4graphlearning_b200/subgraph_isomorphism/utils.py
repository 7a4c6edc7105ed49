"""Host utilities the counting models need: activation registry and weight initialisers.

Same names, argument meaning and numeric behaviour as the reference's
``subgraph_isomorphism/utils/act.py:457-489`` (``map_activation_str_to_layer``; activations are
shared singleton modules; ``leaky_relu`` slope is 1/5.5, act.py:27) and
``utils/init.py:18-158`` (``init_weight`` / ``init_module``: Xavier-uniform / Kaiming-normal
variants with the reference's gain table).  Pure host-side setup code: nothing here runs per step.
"""
import math

import torch as th
import torch.nn as nn

LEAKY_RELU_A = 1 / 5.5


class Identity(nn.Module):
    def forward(self, x):
        return x


class Sparsemax(nn.Module):
    """Euclidean projection of every slice along `dim` onto the probability simplex (Martins & Astudillo 2016;
    utils/act.py:210-331).  Same operation sequence as the reference -- shift by the maximum, descending sort, support
    size k = max{j : 1 + j z_(j) > sum_{i<=j} z_(i)}, threshold tau = (sum of the supported z - 1) / k -- so values and
    autograd gradients agree with it."""

    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        z = x.movedim(self.dim, -1)
        z = z - z.max(dim=-1, keepdim=True)[0]
        zs = th.sort(z, dim=-1, descending=True)[0]
        j = th.arange(1, z.size(-1) + 1, device=z.device, dtype=z.dtype)
        in_support = (1 + j * zs > th.cumsum(zs, dim=-1)).to(z.dtype)
        k = (in_support * j).max(dim=-1, keepdim=True)[0]
        tau = ((in_support * zs).sum(dim=-1, keepdim=True) - 1) / k
        # maximum(0, .) rather than clamp: an entry that lands exactly on the threshold gets half the gradient, as in
        # the reference's th.max(zeros, x - taus)
        return th.maximum(th.zeros_like(z), z - tau).movedim(-1, self.dim)

    def extra_repr(self):
        return "dim={}".format(self.dim)


class GumbelSoftmax(nn.Module):
    """utils/act.py:357-371: torch's gumbel_softmax (random).  The reference passes `dim` in the position of torch's
    deprecated `eps` argument, so it always normalises over the last axis; so does this module."""

    def __init__(self, tau=1.0, hard=False, dim=-1):
        super().__init__()
        self.tau, self.hard, self.dim = tau, hard, dim

    def forward(self, x):
        return th.nn.functional.gumbel_softmax(x, tau=self.tau, hard=self.hard, dim=-1)

    def extra_repr(self):
        return "tau={}, hard={}, dim={}".format(self.tau, self.hard, self.dim)


class _Extreme(nn.Module):
    """keep the extreme entries of every slice along `dim` (all of them on ties), zero the rest; with scale_up the kept
    entries are rescaled so that the slice keeps its sum (0 where that ratio is undefined)   (utils/act.py:374-455)."""

    pick = None

    def __init__(self, dim=-1, scale_up=False, inplace=False):
        super().__init__()
        self.dim, self.scale_up, self.inplace = dim, scale_up, inplace

    def forward(self, x):
        ext = type(self).pick(x, dim=self.dim, keepdim=True)[0]
        kept = x * (x == ext).to(x.dtype)
        if self.scale_up:
            scale = x.sum(dim=self.dim, keepdim=True) / kept.sum(dim=self.dim, keepdim=True)
            kept = kept * scale.masked_fill(scale.isnan(), 0.0)
        if self.inplace:
            return x.copy_(kept)
        return kept

    def extra_repr(self):
        return "dim={}, scale_up={}{}".format(self.dim, self.scale_up, ", inplace=True" if self.inplace else "")


class Maximum(_Extreme):
    pick = staticmethod(th.max)


class Minimum(_Extreme):
    pick = staticmethod(th.min)


supported_act_funcs = {
    "none": Identity(),
    "softmax": nn.Softmax(dim=-1),
    "sparsemax": Sparsemax(dim=-1),
    "gumbel_softmax": GumbelSoftmax(dim=-1),
    "maximum": Maximum(dim=-1),
    "minimum": Minimum(dim=-1),
    "sigmoid": nn.Sigmoid(),
    "tanh": nn.Tanh(),
    "relu": nn.ReLU(),
    "relu6": nn.ReLU6(),
    "leaky_relu": nn.LeakyReLU(negative_slope=LEAKY_RELU_A),
    "prelu": nn.PReLU(init=LEAKY_RELU_A),
    "elu": nn.ELU(),
    "celu": nn.CELU(),
    "selu": nn.SELU(),
    "gelu": nn.GELU(),
}


def map_activation_str_to_layer(act_func, **kw):
    if act_func not in supported_act_funcs:
        raise NotImplementedError(act_func)
    act = supported_act_funcs[act_func]
    for k, v in kw.items():
        if hasattr(act, k):
            try:
                setattr(act, k, v)
            except Exception:
                pass
    return act


def calculate_gain(activation):
    if activation in ("none", "maximum", "minimum"):
        kind = "linear"
    elif activation in ("relu", "relu6", "elu", "selu", "celu", "gelu"):
        kind = "relu"
    elif activation in ("leaky_relu", "prelu"):
        kind = "leaky_relu"
    elif activation in ("softmax", "sparsemax", "gumbel_softmax"):
        kind = "sigmoid"
    elif activation in ("sigmoid", "tanh"):
        kind = activation
    else:
        raise NotImplementedError(activation)
    return nn.init.calculate_gain(kind, LEAKY_RELU_A)


def _fans(x):
    if x.dim() < 2:
        x = x.unsqueeze(-1)
    rf = x[0][0].numel() if x.dim() > 2 else 1
    return x.size(1) * rf, x.size(0) * rf


def _uniform(x, gain):
    fan_in, fan_out = _fans(x)
    a = math.sqrt(3.0) * gain * math.sqrt(2.0 / float(fan_in + fan_out))
    return nn.init.uniform_(x, -a, a)


def _normal(x, gain):
    fan_in, _ = _fans(x)
    return nn.init.normal_(x, 0, gain / math.sqrt(fan_in))


_INITS = {
    "zero": lambda x, gain: nn.init.zeros_(x),
    "uniform": _uniform,
    "normal": _normal,
    "orthogonal": lambda x, gain: nn.init.orthogonal_(x, gain=1.0),
}


def init_weight(x, activation="none", init="uniform"):
    if init not in _INITS:
        raise ValueError("init=%s is not supported now." % init)
    if isinstance(x, th.Tensor):
        _INITS[init](x, calculate_gain(activation))


def init_module(x, activation="none", init="uniform"):
    if init not in _INITS:
        raise ValueError("init=%s is not supported now." % init)
    gain = calculate_gain(activation)
    if isinstance(x, nn.Linear):
        _INITS[init](x.weight, gain)
        if x.bias is not None:
            nn.init.zeros_(x.bias)
    elif isinstance(x, (nn.BatchNorm1d, nn.LayerNorm)):
        nn.init.ones_(x.weight)
        nn.init.zeros_(x.bias)


def expand_dimensions(old_module, new_module, pre_pad=True):
    """grow a trained module into a freshly built larger one (utils/dl.py:157-195): every parameter of `new_module`
    that also exists (by name) in `old_module` is zeroed and receives the old values in its trailing corner (pre_pad,
    the layout of the left-padded encodings) or leading corner; parameters without an old counterpart keep their
    fresh initialisation.  Also accepts two tensors."""
    with th.no_grad():
        if isinstance(old_module, th.Tensor):
            if old_module.dim() < 1 or old_module.dim() > 4:
                raise NotImplementedError
            new_module.zero_()
            corner = tuple(slice(-n, None) if pre_pad else slice(0, n) for n in old_module.shape)
            new_module[corner].copy_(old_module)
            return
        old = dict(old_module.named_parameters())
        for name, param in new_module.named_parameters():
            if name in old:
                expand_dimensions(old[name], param, pre_pad)


class OutputDict(dict):
    """Model output: the reference's key set (models/container.py:14-101, basemodel.py:964-980);
    readable both as ``out["pred_c"]`` and ``out.pred_c``; ``None`` entries are kept so that
    ``output["pred_e"] is not None`` tests in train.py:779-800 keep working."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def to_tuple(self):
        return tuple(v for v in self.values() if v is not None)
