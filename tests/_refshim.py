"""Runs the REFERENCE'S OWN SOURCE TEXT without JAX / NetKet / Flax (test infrastructure only).

`install()` puts minimal stand-ins for `jax`, `jax.numpy`, `jax.scipy.special`, `jax.tree_util`, `flax.core` and
the handful of `netket` names the reference imports into `sys.modules`, then `import netket_fidelity` from
`/root/reference` executes the reference's files AS THEY ARE (nothing is copied into this repository; the files
are read at test / fixture-generation time):

    netket_fidelity/operator/singlequbit_gates.py     Rx / Ry / Hadamard, get_conns_and_mels_* (vmapped kernels)
    netket_fidelity/utils/sampling_Ustate.py          _logpsi_U_fun, make_logpsi_U_afun
    netket_fidelity/infidelity/overlap/expect.py      infidelity_sampling_MCState (kernel_fun, log_pdf_joint)
    netket_fidelity/infidelity/overlap_U/expect.py    infidelity_sampling_MCState (two logsumexp(b=mels))
    netket_fidelity/infidelity/logic.py, overlap*/operator.py   factory, operator classes, error rules
    test/_infidelity_exact.py, test/_finite_diff.py   exact infidelity, central differences

Arrays are `Arr`, a thin wrapper of float64 / complex128 torch tensors (CPU) with the jnp surface these files use
(`.at[].set/.get`, `.real`, `.reshape`, broadcasting arithmetic, ...); torch autograd plays the role of JAX's AD.
What is NOT reference text and therefore still a restatement of NetKet semantics ([NK]):
`nkjax.expect` (A.3), `nkjax.vjp(conjugate=True)` (A.4), `mpi_mean_jax` (identity on one rank), the RBM apply
function (A.1), `Stats`, the `MCState` container.  Everything between those calls is the reference's code.
"""
from __future__ import annotations

import functools
import importlib
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("NKFB_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "netket_fidelity"))


# ------------------------------------------------------------------------------------------------ arrays
def _t(x):
    """anything -> torch tensor (float64 / complex128 / bool / int64)"""
    if isinstance(x, Arr):
        return x.t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x))
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=torch.float64)
    if isinstance(x, (complex, np.complexfloating)):
        return torch.tensor(complex(x), dtype=torch.complex128)
    a = np.asarray(x)
    if a.dtype == object:
        return torch.stack([_t(v) for v in x])
    return torch.from_numpy(np.ascontiguousarray(a))


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, torch.dtype):
        return dtype
    if dtype is complex:
        return torch.complex128
    if dtype is float:
        return torch.float64
    if dtype is int:
        return torch.int64
    if dtype is bool:
        return torch.bool
    return {"complex128": torch.complex128, "complex64": torch.complex128, "float64": torch.float64,
            "float32": torch.float64, "int64": torch.int64, "int32": torch.int64, "bool": torch.bool}[np.dtype(dtype).name]


def _promote(a, b):
    a, b = _t(a), _t(b)
    if a.dtype != b.dtype:
        d = torch.promote_types(a.dtype, b.dtype)
        a, b = a.to(d), b.to(d)
    return a, b


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, _unwrap_idx(idx)

    def set(self, v):
        out = self.arr.t.clone()
        v = _t(v)
        if v.dtype != out.dtype:
            if v.is_complex() and not out.is_complex():
                raise TypeError("complex value into a real array")
            v = v.to(out.dtype)
        out[self.idx] = v
        return Arr(out)

    def get(self):
        return Arr(self.arr.t[self.idx])


def _unwrap_idx(idx):
    if isinstance(idx, tuple):
        return tuple(_unwrap_idx(i) for i in idx)
    if isinstance(idx, Arr):
        return idx.t
    return idx


class Arr:
    """jnp-array look-alike over a torch tensor."""
    __array_priority__ = 1000

    def __init__(self, t):
        self.t = _t(t)

    # --- numpy interop
    def __array__(self, dtype=None, copy=None):
        a = self.t.detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    # --- shape / dtype
    @property
    def shape(self):
        return tuple(self.t.shape)

    @property
    def ndim(self):
        return self.t.ndim

    @property
    def size(self):
        return self.t.numel()

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def real(self):
        return Arr(self.t.real if self.t.is_complex() else self.t)

    @property
    def imag(self):
        return Arr(self.t.imag if self.t.is_complex() else torch.zeros_like(self.t))

    @property
    def T(self):
        return Arr(self.t.T if self.t.ndim > 1 else self.t)

    @property
    def at(self):
        return _At(self)

    def __len__(self):
        return self.t.shape[0]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        return Arr(self.t.reshape(*shape))

    def conj(self):
        return Arr(self.t.conj().resolve_conj() if self.t.is_complex() else self.t)

    def sum(self, axis=None):
        return Arr(self.t.sum() if axis is None else self.t.sum(axis))

    def mean(self, axis=None):
        return Arr(self.t.mean() if axis is None else self.t.mean(axis))

    def astype(self, dtype):
        return Arr(self.t.to(_dt(dtype)))

    def item(self):
        return self.t.item()

    def __getitem__(self, idx):
        return Arr(self.t[_unwrap_idx(idx)])

    def __iter__(self):
        for i in range(self.t.shape[0]):
            yield Arr(self.t[i])

    def __float__(self):
        return float(self.t)

    def __complex__(self):
        return complex(self.t)

    def __bool__(self):
        return bool(self.t)

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return f"Arr({self.t!r})"

    # --- arithmetic
    def _bin(self, other, fn, rev=False):
        a, b = _promote(self, other)
        return Arr(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __matmul__(self, o): return self._bin(o, torch.matmul)
    def __rmatmul__(self, o): return self._bin(o, torch.matmul, True)
    def __pow__(self, o): return self._bin(o, torch.pow)
    def __neg__(self): return Arr(-self.t)
    def __eq__(self, o): return self._bin(o, torch.eq)
    def __ne__(self, o): return self._bin(o, torch.ne)
    def __lt__(self, o): return self._bin(o, torch.lt)
    def __gt__(self, o): return self._bin(o, torch.gt)
    def __le__(self, o): return self._bin(o, torch.le)
    def __ge__(self, o): return self._bin(o, torch.ge)


def _wrap1(fn):
    def f(x, *a, **k):
        return Arr(fn(_t(x), *a, **k))
    return f


def _logsumexp(a, axis=None, b=None):
    """jax.scipy.special.logsumexp semantics: max-shift on the real part, (complex) sum, principal log."""
    a = _t(a)
    if axis is None:
        a, axis = a.reshape(-1), 0
        if b is not None:
            b = _t(b).reshape(-1)
    amax = a.real.detach().max(dim=axis, keepdim=True).values if a.is_complex() else a.detach().max(dim=axis, keepdim=True).values
    amax = torch.where(torch.isfinite(amax), amax, torch.zeros_like(amax))
    e = torch.exp(a - amax)
    if b is not None:
        bb, e = _promote(b, e)
        e = bb * e
    s = e.sum(dim=axis)
    return Arr(torch.log(s) + amax.squeeze(axis))


def _make_jnp():
    m = types.ModuleType("jax.numpy")

    def asarray(x, dtype=None):
        t = _t(x)
        if t.dtype in (torch.int64, torch.int32) and dtype is None:
            pass
        d = _dt(dtype)
        return Arr(t.to(d) if d is not None else t)

    m.asarray = m.array = asarray
    m.tile = lambda x, reps: Arr(_t(x).repeat(*reps) if _t(x).ndim == len(reps) else _t(x).unsqueeze(0).repeat(*reps))

    def where(c, a, b):
        a, b = _promote(a, b)
        return Arr(torch.where(_t(c).bool(), a, b))

    m.where = where
    m.zeros = lambda shape, dtype=float: Arr(torch.zeros(shape, dtype=_dt(dtype)))
    m.ones = lambda shape, dtype=float: Arr(torch.ones(shape, dtype=_dt(dtype)))
    m.ones_like = lambda x: Arr(torch.ones_like(_t(x)))
    m.zeros_like = lambda x: Arr(torch.zeros_like(_t(x)))
    def ew(ft, fnp):
        # elementwise maths: NumPy (libm, correctly rounded sqrt) unless autograd needs the torch op -- torch's
        # vectorised CPU sqrt is 1 ulp off (sqrt(2.) = 1.414213562373095), which is not what XLA:CPU computes
        def f(x):
            t = _t(x)
            if t.requires_grad:
                return Arr(ft(t))
            if not (t.is_floating_point() or t.is_complex()):
                t = t.to(torch.float64)
            return Arr(torch.from_numpy(np.asarray(fnp(t.numpy()))))
        return f

    for name, (ft, fnp) in dict(cos=(torch.cos, np.cos), sin=(torch.sin, np.sin), exp=(torch.exp, np.exp),
                                log=(torch.log, np.log), sqrt=(torch.sqrt, np.sqrt), tanh=(torch.tanh, np.tanh),
                                cosh=(torch.cosh, np.cosh), absolute=(torch.abs, np.abs), abs=(torch.abs, np.abs),
                                log1p=(torch.log1p, np.log1p)).items():
        setattr(m, name, ew(ft, fnp))
    m.signbit = lambda x: Arr(torch.signbit(_t(x)))
    m.real = lambda x: Arr(x).real
    m.imag = lambda x: Arr(x).imag
    m.conj = lambda x: Arr(x).conj()
    m.sum = lambda x, axis=None: Arr(x).sum(axis)
    m.mean = lambda x, axis=None: Arr(x).mean(axis)
    m.iscomplexobj = lambda x: bool(np.iscomplexobj(np.asarray(x))) if not isinstance(x, Arr) else x.t.is_complex()
    m.iscomplex = lambda x: bool(np.any(np.iscomplex(np.asarray(x))))
    m.isnan = lambda x: Arr(torch.isnan(_t(x)))
    m.pi = np.pi
    m.ndarray = Arr
    m.complex128, m.float64 = complex, float
    return m


def _vmap(fun=None, in_axes=0, out_axes=0):
    """jax.vmap as an explicit loop over axis 0 of the mapped arguments (exactly what vmap means)."""
    if fun is None:
        return functools.partial(_vmap, in_axes=in_axes, out_axes=out_axes)

    @functools.wraps(fun)
    def mapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(len(Arr(a)) for a, ax in zip(args, axes) if ax == 0)
        outs = [fun(*[Arr(a)[i] if ax == 0 else a for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(Arr(torch.stack([_t(o[k]) for o in outs])) for k in range(len(outs[0])))
        return Arr(torch.stack([_t(o) for o in outs]))

    return mapped


def _jit(fun=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def _tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return {k: _tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (tuple, list)):
        return type(tree)(_tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return f(tree, *rest)


def _tree_leaves(tree):
    if isinstance(tree, dict):
        return [l for k in tree for l in _tree_leaves(tree[k])]
    if isinstance(tree, (tuple, list)):
        return [l for v in tree for l in _tree_leaves(v)]
    return [tree]


# ------------------------------------------------------------------------------------------------ NetKet stand-ins [NK]
class Stats:
    """netket.stats.Stats: the fields the reference touches (mean via .replace, variance, error_of_mean).
    `local_values` (not a NetKet field) keeps the per-sample estimator values for the parity tests."""

    def __init__(self, mean, variance, error_of_mean=float("nan"), local_values=None):
        self.mean, self.variance, self.error_of_mean, self.local_values = mean, variance, error_of_mean, local_values

    def replace(self, **kw):
        d = dict(mean=self.mean, variance=self.variance, error_of_mean=self.error_of_mean,
                 local_values=self.local_values)
        d.update(kw)
        return Stats(**d)


def nk_expect(log_pdf, f, pars, samples, n_chains=None):
    """[NK] nkjax.expect: forward mean(L) and Stats; backward = grad of mean_s[dL_s log_pdf + L_s]
    with dL = L - mean held constant.  Realised as a surrogate whose VALUE is the mean and whose autograd
    gradient is that expression."""
    L = _t(f(pars, samples))
    mean = L.mean()
    dL = (L - mean).detach()
    lp = _t(log_pdf(pars, samples))
    surrogate = (dL * lp + L).mean()
    value = mean.detach() + (surrogate - surrogate.detach())
    Ld = L.detach()
    stats = Stats(mean=Arr(mean.detach()), variance=float(((Ld - Ld.mean()) ** 2).mean()),
                  local_values=Ld.numpy().copy())
    return Arr(value), stats


def nk_vjp(fun, *primals, has_aux=False, conjugate=False):
    """[NK] nkjax.vjp(conjugate=True): for a real output and complex parameters the cotangent pulled
    back is dF/dRe(theta) + i dF/dIm(theta) -- torch.autograd's convention for complex leaves."""
    assert conjugate, "the reference always passes conjugate=True"
    leaves_in = _tree_leaves(primals)
    fresh = [_t(l).detach().clone().requires_grad_(True) for l in leaves_in]
    it = iter(fresh)
    prim = _tree_map(lambda _: Arr(next(it)), primals)
    out = fun(*prim)
    aux = None
    if has_aux:
        out, aux = out
    y = _t(out)

    def vjp_fun(ct):
        g = torch.autograd.grad(y, fresh, grad_outputs=_t(ct).to(y.dtype), allow_unused=True)
        g = [gi if gi is not None else torch.zeros_like(fi) for gi, fi in zip(g, fresh)]
        # real parameters: ordinary gradient (imaginary part dropped by autograd already)
        itg = iter(g)
        return _tree_map(lambda _: Arr(next(itg)), primals)

    return (Arr(y.detach()), vjp_fun, aux) if has_aux else (Arr(y.detach()), vjp_fun)


class _Dispatcher:
    """netket.vqs.expect / expect_and_grad: `.dispatch` records every overload; call picks by argument classes."""

    def __init__(self, name):
        self.name, self.overloads = name, []

    def dispatch(self, fn):
        self.overloads.append(fn)
        return fn

    def __call__(self, vstate, op):
        for fn in reversed(self.overloads):
            ann = list(fn.__annotations__.values())
            if isinstance(vstate, ann[0]) and isinstance(op, ann[1]):
                if self.name == "expect_and_grad":
                    return fn(vstate, op, None, mutable=False)
                return fn(vstate, op, None)
        raise TypeError(f"no {self.name} overload for ({type(vstate).__name__}, {type(op).__name__})")


class Hilbert:
    """netket.hilbert.Spin(1/2) / Qubit: local states and the enumeration order NetKet uses (all_states)."""

    def __init__(self, N, local_states=(-1.0, 1.0)):
        self.size, self.local_states = N, np.asarray(local_states, dtype=float)

    @property
    def n_states(self):
        return 2 ** self.size

    def all_states(self):
        idx = np.arange(self.n_states)
        bits = (idx[:, None] >> np.arange(self.size - 1, -1, -1)[None, :]) & 1
        return self.local_states[bits]

    def __eq__(self, other):
        return isinstance(other, Hilbert) and self.size == other.size and np.array_equal(self.local_states, other.local_states)

    def __hash__(self):
        return hash((self.size, tuple(self.local_states)))


class AbstractOperator:
    def __init__(self, hilbert):
        self._hilbert = hilbert

    @property
    def hilbert(self):
        return self._hilbert


class DiscreteJaxOperator(AbstractOperator):
    def to_dense(self):
        """dense matrix from get_conn_padded over all basis states: <x|O|x'> = mel(x, x')."""
        hi = self.hilbert
        X = hi.all_states()
        xp, mels = self.get_conn_padded(Arr(X))
        xp, mels = np.asarray(xp), np.asarray(mels)
        bits = (xp == hi.local_states[1]).astype(np.int64)
        idx = (bits << np.arange(hi.size - 1, -1, -1)).sum(-1)
        D = np.zeros((hi.n_states, hi.n_states), dtype=complex)
        for r in range(hi.n_states):
            for c in range(mels.shape[1]):
                D[r, idx[r, c]] += mels[r, c]
        return D

    def to_sparse(self):
        return self.to_dense()


class Adjoint(AbstractOperator):
    pass


class AbstractObservable:
    def __init__(self, hilbert):
        self._hilbert = hilbert

    @property
    def hilbert(self):
        return self._hilbert


class VariationalState:
    pass


class MCState(VariationalState):
    """Container with the attributes the reference reads (netket.vqs.MCState): samples are GIVEN, not drawn."""

    def __init__(self, sampler=None, model=None, *, apply_fun=None, n_samples=None, variables=None, hilbert=None,
                 samples=None):
        self.sampler, self.n_samples = sampler, n_samples
        self._apply_fun = apply_fun
        self.variables = variables
        self.hilbert = hilbert if hilbert is not None else getattr(sampler, "hilbert", None)
        self._samples = samples

    @property
    def parameters(self):
        return self.variables["params"]

    @parameters.setter
    def parameters(self, p):
        self.variables = {**self.variables, "params": p}

    @property
    def model_state(self):
        return {k: v for k, v in self.variables.items() if k != "params"}

    @property
    def samples(self):
        return self._samples

    def to_array(self, normalize=False):
        lp = np.asarray(self._apply_fun(self.variables, Arr(self.hilbert.all_states())))
        psi = np.exp(lp)
        return psi / np.linalg.norm(psi) if normalize else psi


class FullSumState(VariationalState):
    pass


def rbm_apply(variables, x):
    """[NK] nk.models.RBM.__call__ + nknn.log_cosh on Arr / torch (complex128)."""
    p = variables["params"]
    x = _t(x)
    W = _t(p["Dense"]["kernel"])
    xx = x.to(W.dtype)
    th = xx @ W
    if "bias" in p["Dense"]:
        th = th + _t(p["Dense"]["bias"])
    re = th.real if th.is_complex() else th
    sgn = torch.where(torch.signbit(re), -1.0, 1.0).to(th.dtype)
    y = th * sgn
    lc = y + torch.log1p(torch.exp(-2.0 * y)) - np.log(2.0)
    out = lc.sum(-1)
    if "visible_bias" in p:
        out = out + xx @ _t(p["visible_bias"])
    return Arr(out)


def get_local_kernel_arguments(vstate, op):
    """[NK] jax-operator branch: (samples, operator); the connected elements are generated inside the jitted
    function (overlap_U/expect.py:91-93)."""
    return vstate.samples, op


# ------------------------------------------------------------------------------------------------ installation
_installed = None


def install():
    """Registers the stand-in modules and imports the reference package from REFERENCE_ROOT.
    Returns a namespace with the reference modules and the shim helpers."""
    global _installed
    if _installed is not None:
        return _installed
    if not reference_available():
        raise RuntimeError(f"{REFERENCE_ROOT} is not present")

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        parent, _, child = name.rpartition(".")
        if parent:
            setattr(sys.modules[parent], child, m)
        return m

    for forbidden in ("jax", "netket", "flax"):
        if forbidden in sys.modules and not getattr(sys.modules[forbidden], "__nkfb_shim__", False):
            raise RuntimeError(f"a real `{forbidden}` is already imported; the shim is for containers without it")

    jnp = _make_jnp()
    jax = mod("jax", jit=_jit, vmap=_vmap, __nkfb_shim__=True)
    sys.modules["jax.numpy"] = jnp
    jax.numpy = jnp
    mod("jax.scipy")
    mod("jax.scipy.special", logsumexp=_logsumexp)
    mod("jax.tree_util", tree_map=_tree_map, register_pytree_node_class=lambda c: c, tree_leaves=_tree_leaves)

    def _pop(d, key):
        d2 = {k: v for k, v in d.items() if k != key}
        return d2, d[key]

    flax = mod("flax", __nkfb_shim__=True)
    mod("flax.core", pop=_pop, copy=lambda d, add: {**d, **add})

    expect, expect_and_grad = _Dispatcher("expect"), _Dispatcher("expect_and_grad")
    nk = mod("netket", __nkfb_shim__=True)
    mod("netket.jax", expect=nk_expect, vjp=nk_vjp, HashablePartial=functools.partial,
        maybe_promote_to_complex=lambda *dts: np.result_type(*[np.dtype(complex) if np.issubdtype(np.dtype(d), np.complexfloating) else np.dtype(d) for d in dts]))
    mod("netket.utils", _hide_submodules=lambda *a, **k: None)
    mod("netket.utils.mpi", mpi_mean_jax=lambda x, **k: (x, None), n_nodes=1, rank=0)
    mod("netket.utils.types", DType=object, PyTree=object)
    mod("netket.utils.numbers", is_scalar=lambda x: np.ndim(np.asarray(x)) == 0)
    mod("netket.stats", Stats=Stats)
    mod("netket.vqs", MCState=MCState, FullSumState=FullSumState, VariationalState=VariationalState, expect=expect,
        expect_and_grad=expect_and_grad, get_local_kernel_arguments=get_local_kernel_arguments)
    mod("netket.operator", DiscreteJaxOperator=DiscreteJaxOperator, AbstractOperator=AbstractOperator, Adjoint=Adjoint,
        IsingJax=type("IsingJax", (DiscreteJaxOperator,), {}), spin=types.SimpleNamespace())
    mod("netket.optimizer", identity_preconditioner=lambda *a, **k: None, PreconditionerT=object)
    mod("netket.experimental")
    mod("netket.experimental.observable", AbstractObservable=AbstractObservable,
        Renyi2EntanglementEntropy=type("Renyi2EntanglementEntropy", (), {}))
    mod("netket.driver")
    mod("netket.driver.abstract_variational_driver",
        AbstractVariationalDriver=type("AbstractVariationalDriver", (), {"__init__": lambda self, *a, **k: None}))
    mod("netket.callbacks", ConvergenceStopping=type("ConvergenceStopping", (), {"__init__": lambda self, *a, **k: None}))
    nk.jax = sys.modules["netket.jax"]

    # the reference's hatch-vcs version file does not exist in a source checkout
    mod_version = types.ModuleType("netket_fidelity._version")
    mod_version.version = "0+reference.source"
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.modules["netket_fidelity._version"] = mod_version
    pkg = importlib.import_module("netket_fidelity")

    def load_test_helper(name):
        path = os.path.join(REFERENCE_ROOT, "test", name + ".py")
        spec = importlib.util.spec_from_file_location("nkf_reference_test" + name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m

    ns = types.SimpleNamespace(
        pkg=pkg,
        gates=importlib.import_module("netket_fidelity.operator.singlequbit_gates"),
        sampling_Ustate=importlib.import_module("netket_fidelity.utils.sampling_Ustate"),
        expect_std=importlib.import_module("netket_fidelity.infidelity.overlap.expect"),
        expect_U=importlib.import_module("netket_fidelity.infidelity.overlap_U.expect"),
        logic=importlib.import_module("netket_fidelity.infidelity.logic"),
        exact_helper=load_test_helper("_infidelity_exact"),
        finite_diff=load_test_helper("_finite_diff"),
        Arr=Arr, Hilbert=Hilbert, MCState=MCState, FullSumState=FullSumState, rbm_apply=rbm_apply,
        expect=expect, expect_and_grad=expect_and_grad, tree_map=_tree_map, DiscreteJaxOperator=DiscreteJaxOperator,
    )
    _installed = ns
    return ns


def params_to_arr(p):
    """oracle.rbm.RBMParams -> NetKet parameter tree of Arr (complex128, or float64 for real parameters)."""
    d = {"Dense": {"kernel": Arr(np.asarray(p.kernel))}}
    if p.bias is not None:
        d["Dense"]["bias"] = Arr(np.asarray(p.bias))
    if p.visible_bias is not None:
        d["visible_bias"] = Arr(np.asarray(p.visible_bias))
    return d
