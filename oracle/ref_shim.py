"""TEST INFRASTRUCTURE -- runs the reference's own model SOURCE without jax / numpyro.

jax and numpyro cannot be installed in the build image, so the reference cannot run as published.  What can run is its
source: this module puts small stand-ins for the handful of `jax.numpy` / `numpyro` names the reference uses (see the
list in `_JNP` and `_Numpyro*` below -- 38 names in all) into `sys.modules`, imports `bpl` from /root/reference
unmodified, calls the real `fit()` of a predictor class -- so the reference's own data preparation (`parse_teams`,
weights, covariates) runs too -- and intercepts `MCMC.run(...)`.  The captured `_model` callable and its arguments are
then traced in float64 torch at given unconstrained parameter values, giving the log joint density numpyro's
`potential_energy` would compute (site log-probs + log|det J| of the `biject_to` transforms, `LocScaleReparam`,
`plate`, `handlers.scale`, `factor`) and, by autograd, its gradient.

What this pins: every line of the reference's model code and data preparation (which sites, which formulas, indexing,
weights, clipping, the bounds, the low-score term).  What it does not pin: numpyro's own arithmetic -- the stand-ins below
restate it from its documentation, like oracle/models.py does.  Used only by scripts/make_ref_shim_golden.py (in the
build container, where /root/reference exists); the vectors it writes are committed under tests/golden/.
"""
from __future__ import annotations

import contextlib
import inspect
import math
import sys
import types

import numpy as np
import torch

F64 = torch.float64


# ---------------------------------------------------------------------------------------------------- jax.numpy
def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, torch.Tensor) for e in x):
            return torch.stack([_t(e, F64).reshape(()) if not isinstance(e, torch.Tensor) or e.ndim == 0 else e for e in x])
        x = np.asarray(x)
    if isinstance(x, np.ndarray):
        if x.dtype.kind in "iub":
            return torch.from_numpy(x.astype(np.int64) if x.dtype.kind != "b" else x)
        return torch.from_numpy(x.astype(np.float64))
    if isinstance(x, (bool, int)):
        return torch.tensor(x)
    return torch.tensor(float(x), dtype=F64)


def _array(x, dtype=None):
    out = _t(x)
    if dtype is not None:
        s = str(dtype)
        if "int" in s:
            out = out.to(torch.int64)
        elif "float" in s:
            out = out.to(F64)
    return out


def _clip(x, a_min=None, a_max=None):
    x = _t(x, F64)
    if a_min is not None:
        x = torch.clamp(x, min=float(a_min))
    if a_max is not None:
        x = torch.clamp(x, max=float(a_max))
    return x


class _At:
    """`x.at[idx].set(v)`: jax's functional update (returns a new array)."""

    def __init__(self, x):
        self.x = x

    def __getitem__(self, idx):
        x = self.x

        class _Ref:
            @staticmethod
            def set(v):
                out = x.clone()
                out[idx] = _t(v, x.dtype) if not isinstance(v, torch.Tensor) else v.to(x.dtype)
                return out
        return _Ref


torch.Tensor.at = property(lambda self: _At(self))  # only inside this test helper's process
_torch_std = torch.Tensor.std


def _jax_std(self, *args, axis=None, **kw):
    """`x.std(axis=...)` the way jax / numpy mean it: population standard deviation (torch's default is Bessel-corrected)."""
    if axis is not None and not args and not kw:
        return _torch_std(self, dim=axis, unbiased=False)
    return _torch_std(self, *args, **kw)


torch.Tensor.std = _jax_std


def _reduce(fn):
    def f(x, axis=None):
        x = _t(x, F64)
        return fn(x) if axis is None else fn(x, dim=axis)
    return f


_JNP = dict(
    array=_array, asarray=_array, exp=lambda x: torch.exp(_t(x, F64)), log=lambda x: torch.log(_t(x, F64)),
    sqrt=lambda x: torch.sqrt(_t(x, F64)), clip=_clip,
    min=_reduce(lambda x, dim=None: x.min() if dim is None else x.min(dim=dim).values),
    max=_reduce(lambda x, dim=None: x.max() if dim is None else x.max(dim=dim).values),
    sum=_reduce(lambda x, dim=None: x.sum() if dim is None else x.sum(dim=dim)),
    matmul=lambda a, b: torch.matmul(_t(a, F64), _t(b, F64)), dot=lambda a, b: torch.matmul(_t(a, F64), _t(b, F64)),
    concatenate=lambda xs, axis=0: torch.cat([_t(x) for x in xs], dim=axis),
    tile=lambda x, reps: torch.tile(_t(x), tuple(np.atleast_1d(reps))),
    repeat=lambda x, n, axis=None: torch.repeat_interleave(_t(x).reshape(-1) if axis is None else _t(x), int(n),
                                                           dim=0 if axis is None else axis),
    arange=lambda *a, dtype=None: torch.arange(*a), zeros=lambda shape, dtype=None: torch.zeros(shape, dtype=F64),
    ones=lambda shape, dtype=None: torch.ones(shape, dtype=F64), empty=lambda shape, dtype=None: torch.zeros(shape, dtype=F64),
    zeros_like=lambda x: torch.zeros_like(_t(x, F64)), ones_like=lambda x: torch.ones_like(_t(x, F64)),
    ndarray=torch.Tensor, float32="float32", float64="float64", int32="int32", uint32="uint32", uint16="uint16", uint8="uint8",
)


# ---------------------------------------------------------------------------------------------------- distributions
class _Dist:
    support = "real"
    event = 0

    def to_event(self, n):
        self.event = n
        return self


class Normal(_Dist):
    def __init__(self, loc=0.0, scale=1.0):
        self.loc, self.scale = _t(loc, F64), _t(scale, F64)

    def log_prob(self, x):
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - torch.log(self.scale) - 0.5 * math.log(2.0 * math.pi)


class HalfNormal(_Dist):
    support = "positive"

    def __init__(self, scale=1.0):
        self.scale = _t(scale, F64)

    def log_prob(self, x):
        return Normal(0.0, self.scale).log_prob(x) + math.log(2.0)


class Beta(_Dist):
    support = "unit"

    def __init__(self, concentration1, concentration0):
        self.a, self.b = float(concentration1), float(concentration0)

    def log_prob(self, x):
        lbeta = math.lgamma(self.a) + math.lgamma(self.b) - math.lgamma(self.a + self.b)
        return (self.a - 1.0) * torch.log(x) + (self.b - 1.0) * torch.log1p(-x) - lbeta


class Uniform(_Dist):
    support = "interval"

    def __init__(self, low=0.0, high=1.0):
        self.low, self.high = float(low), float(high)

    def log_prob(self, x):
        return torch.full_like(x, -math.log(self.high - self.low))


class Poisson(_Dist):
    def __init__(self, rate):
        self.rate = _t(rate, F64)

    def log_prob(self, k):
        k = _t(k, F64)
        return k * torch.log(self.rate) - self.rate - torch.lgamma(k + 1.0)


class LocScaleReparam:
    def __init__(self, centered=None):
        if centered != 0:
            raise NotImplementedError("only the fully decentred form is used by the reference")


# ---------------------------------------------------------------------------------------------------- tracer
class _Trace:
    def __init__(self, values):
        self.values = values          # site name -> unconstrained torch tensor
        self.lp = torch.zeros((), dtype=F64)
        self.plates, self.scales, self.reparams = [], [], []
        self.sites, self.deterministic = [], {}

    def _scale(self):
        s = None
        for w in self.scales:
            s = w if s is None else s * w
        return s

    def _shape(self, fn_shape):
        shape = tuple(size for _, size in self.plates)
        return shape if len(shape) >= len(fn_shape) else tuple(fn_shape)

    def add(self, lp_elem):
        s = self._scale()
        self.lp = self.lp + ((lp_elem * s) if s is not None else lp_elem).sum()


_STACK: list = []


def _cur() -> _Trace:
    return _STACK[-1]


def _constrain(fn, u):
    """numpyro `biject_to(fn.support)`: value and log|det J|."""
    if fn.support == "real":
        return u, torch.zeros_like(u)
    if fn.support == "positive":  # ExpTransform
        return torch.exp(u), u
    if fn.support in ("unit", "interval"):  # SigmoidTransform (clipped like numpyro's), then an affine map
        fi = torch.finfo(u.dtype)
        y = torch.clamp(torch.sigmoid(u), min=fi.tiny, max=1.0 - fi.eps)
        lj = -torch.nn.functional.softplus(-u) - torch.nn.functional.softplus(u)
        if fn.support == "interval":
            y = fn.low + (fn.high - fn.low) * y
            lj = lj + math.log(fn.high - fn.low)
        return y, lj
    raise NotImplementedError(fn.support)


def sample(name, fn, obs=None, **_):
    tr = _cur()
    if obs is not None:
        tr.add(fn.log_prob(_t(obs)))
        tr.sites.append((name, "obs"))
        return obs
    cfg = next((c[name] for c in reversed(tr.reparams) if name in c), None)
    if cfg is not None:  # LocScaleReparam(centered=0): z ~ N(0, 1), value = loc + scale * z
        assert isinstance(fn, Normal)
        z = tr.values[name + "_decentered"]
        tr.add(Normal(0.0, 1.0).log_prob(z))
        tr.sites.append((name + "_decentered", tuple(z.shape)))
        value = fn.loc + fn.scale * z
        tr.deterministic[name] = value
        return value
    u = tr.values[name]
    x, lj = _constrain(fn, u)
    tr.add(fn.log_prob(x) + lj)
    tr.sites.append((name, tuple(u.shape)))
    return x


def deterministic(name, value):
    _cur().deterministic[name] = value
    return value


def factor(name, log_factor):
    _cur().add(_t(log_factor, F64))


@contextlib.contextmanager
def plate(name, size, **_):
    _cur().plates.append((name, int(size)))
    try:
        yield
    finally:
        _cur().plates.pop()


@contextlib.contextmanager
def scale(fn=None, scale=None):
    _cur().scales.append(_t(scale, F64))
    try:
        yield
    finally:
        _cur().scales.pop()


@contextlib.contextmanager
def reparam(fn=None, config=None):
    _cur().reparams.append(config)
    try:
        yield
    finally:
        _cur().reparams.pop()


# ---------------------------------------------------------------------------------------------------- MCMC capture
class Captured(Exception):
    def __init__(self, model, args, kwargs):
        super().__init__("captured MCMC.run")
        self.model, self.args, self.kwargs = model, args, kwargs


class NUTS:
    def __init__(self, model=None, **kw):
        self.model = model


class MCMC:
    def __init__(self, kernel, **kw):
        self.kernel = kernel

    def run(self, rng_key, *args, **kwargs):
        raise Captured(self.kernel.model, args, kwargs)


def install(reference_root="/root/reference"):
    """Puts the stand-ins into sys.modules and the reference on sys.path.  Refuses to shadow a real jax / numpyro."""
    for real in ("jax", "numpyro"):
        if real in sys.modules and not getattr(sys.modules[real], "_bplx_shim", False):
            raise RuntimeError(f"a real {real} is imported: use it instead of the stand-in")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m._bplx_shim = True
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    jnp = mod("jax.numpy", **_JNP)
    rnd = mod("jax.random", PRNGKey=lambda s: int(s), split=lambda k, n=2: [k] * n, choice=None)
    mod("jax", numpy=jnp, random=rnd, vmap=None)
    dist = mod("numpyro.distributions", Normal=Normal, HalfNormal=HalfNormal, Beta=Beta, Uniform=Uniform, Poisson=Poisson)
    handlers = mod("numpyro.handlers", reparam=reparam, scale=scale)
    rep = mod("numpyro.infer.reparam", LocScaleReparam=LocScaleReparam)
    infer = mod("numpyro.infer", MCMC=MCMC, NUTS=NUTS, reparam=rep)
    mod("numpyro", sample=sample, deterministic=deterministic, factor=factor, plate=plate, distributions=dist,
        handlers=handlers, infer=infer)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)


def capture_fit(predictor, *fit_args, **fit_kwargs) -> Captured:
    """Calls the reference's own `fit` and returns what it handed to `MCMC.run`."""
    try:
        predictor.fit(*fit_args, **fit_kwargs)
    except Captured as c:
        return c
    raise RuntimeError("fit() returned without calling MCMC.run")


def log_density(cap: Captured, values: dict):
    """Log joint density in unconstrained space (= -potential energy) at `values` (site name -> numpy array) and its
    gradient per site, by tracing the captured reference model."""
    vals = {k: torch.tensor(np.asarray(v, dtype=np.float64), requires_grad=True) for k, v in values.items()}
    tr = _Trace(vals)
    _STACK.append(tr)
    # jax takes numpy arrays wherever it takes its own; torch does not (ndarray * Tensor): hand the model tensors, except
    # the covariate matrix, which the model standardises with numpy's own mean / std (population std, unlike torch's)
    names = list(inspect.signature(cap.model).parameters)
    args = [(_t(a) if isinstance(a, np.ndarray) and names[i] != "team_covariates" else a) for i, a in enumerate(cap.args)]
    kwargs = {k: (_t(a) if isinstance(a, np.ndarray) and k != "team_covariates" else a) for k, a in cap.kwargs.items()}
    try:
        cap.model(*args, **kwargs)
    finally:
        _STACK.pop()
    used = {n for n, s in tr.sites if s != "obs"}
    if used != set(vals):
        raise RuntimeError(f"latent sites differ: model has {sorted(used)}, values have {sorted(vals)}")
    grads = torch.autograd.grad(tr.lp, [vals[k] for k in vals], allow_unused=True)
    det = {k: v.detach().numpy() for k, v in tr.deterministic.items()}
    return float(tr.lp.detach()), {k: (g.numpy() if g is not None else np.zeros_like(values[k])) for k, g in zip(vals, grads)}, det
