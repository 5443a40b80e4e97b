"""A NumPy-backed stand-in for the small part of the jax API that the reference's hot-path
modules touch, so that the reference's *unmodified* sources (hmm.py, params.py, transition.py,
size_history.py, model.py, data.py under /root/reference/src/phlash) can be executed in this
container, where jax cannot be installed.  Used ONLY by make_golden.py to generate the committed
golden vectors; never imported by the product or by the tests themselves.

Semantics: everything is eager float64 NumPy.  vmap / scan / associative_scan are plain Python
loops (sequential association order), remat/jit are identities.
"""

from __future__ import annotations

import dataclasses
import importlib.machinery
import sys
import types

import numpy as np
import scipy.special
import scipy.stats


# ---------------------------------------------------------------------------- pytrees
class _Static:
    def __class_getitem__(cls, item):
        return cls


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def _dyn_fields(x):
    return [f.name for f in dataclasses.fields(x) if "Static" not in str(f.type)]


def tree_flatten(tree):
    """-> (leaves, rebuild)"""
    if tree is None:
        return [], lambda leaves: None
    if _is_namedtuple(tree):
        parts = [tree_flatten(c) for c in tree]
        cls = type(tree)
        return _join(parts, lambda kids: cls(*kids))
    if isinstance(tree, (tuple, list)):
        parts = [tree_flatten(c) for c in tree]
        cls = type(tree)
        return _join(parts, lambda kids: cls(kids))
    if isinstance(tree, dict):
        keys = sorted(tree)
        parts = [tree_flatten(tree[k]) for k in keys]
        return _join(parts, lambda kids: dict(zip(keys, kids)))
    if dataclasses.is_dataclass(tree) and not isinstance(tree, type):
        names = _dyn_fields(tree)
        parts = [tree_flatten(getattr(tree, n)) for n in names]
        return _join(parts, lambda kids: dataclasses.replace(tree, **dict(zip(names, kids))))
    return [tree], lambda leaves: leaves[0]


def _join(parts, make):
    sizes = [len(p[0]) for p in parts]
    leaves = [leaf for p in parts for leaf in p[0]]

    def rebuild(new_leaves):
        kids, pos = [], 0
        for (_, rb), n in zip(parts, sizes):
            kids.append(rb(new_leaves[pos : pos + n]))
            pos += n
        return make(kids)

    return leaves, rebuild


def tree_map(f, tree, *rest):
    leaves, rebuild = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return rebuild([f(*args) for args in zip(leaves, *others)])


def _tree_index(tree, i):
    return tree_map(lambda a: np.asarray(a)[i], tree)


def _tree_stack(trees):
    if trees[0] is None:
        return None
    return tree_map(lambda *xs: np.stack([np.asarray(x) for x in xs]), *trees)


# ---------------------------------------------------------------------------- transforms
def vmap(f, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = np.asarray(tree_flatten(a)[0][0]).shape[ax]
                break
        outs = []
        for i in range(n):
            call = []
            for a, ax in zip(args, axes):
                if ax is None:
                    call.append(a)
                else:
                    call.append(tree_map(lambda x: np.take(np.asarray(x), i, axis=ax), a))
            outs.append(f(*call))
        stacked = _tree_stack(outs)
        if out_axes != 0:
            stacked = tree_map(lambda x: np.moveaxis(x, 0, out_axes), stacked)
        return stacked

    return wrapped


def scan(f, init, xs, length=None):
    n = length if xs is None else np.asarray(tree_flatten(xs)[0][0]).shape[0]
    carry, ys = init, []
    for i in range(n):
        carry, y = f(carry, None if xs is None else _tree_index(xs, i))
        ys.append(y)
    return carry, _tree_stack(ys)


def associative_scan(op, elems, reverse=False, axis=0):
    x = np.asarray(elems)
    assert axis == 0
    if reverse:
        x = x[::-1]
    out = [x[0]]
    for k in range(1, len(x)):
        out.append(op(x[k], out[-1]) if reverse else op(out[-1], x[k]))
    out = np.stack(out)
    return out[::-1] if reverse else out


def _identity_decorator(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def ravel_pytree(tree):
    leaves, rebuild = tree_flatten(tree)
    shapes = [np.shape(x) for x in leaves]
    flat = np.concatenate([np.ravel(np.asarray(x, dtype=np.float64)) for x in leaves])

    def unravel(v):
        out, pos = [], 0
        for s in shapes:
            k = int(np.prod(s)) if s else 1
            out.append(np.reshape(v[pos : pos + k], s))
            pos += k
        return rebuild(out)

    return flat, unravel


# ---------------------------------------------------------------------------- module assembly
class _NumpyProxy(types.ModuleType):
    def __getattr__(self, name):
        return getattr(np, name)


class _Anything(types.ModuleType):
    """Stub for an absent third-party reader (pysam, tskit, ...): any attribute is a dummy class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return type(name, (), {})


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__dict__.update(attrs)
    return m


def install(reference_src="/root/reference/src"):
    """Put the shim (and stubs for the absent third-party readers) into sys.modules and make
    ``phlash.<submodule>`` importable without running phlash/__init__.py."""
    import jaxtyping  # noqa: F401  (real; import before the fake jax exists)

    jnp = _NumpyProxy("jax.numpy")
    jnp.__spec__ = importlib.machinery.ModuleSpec("jax.numpy", None)
    lax = _mod("jax.lax", scan=scan, associative_scan=associative_scan)
    tree = _mod("jax.tree", map=tree_map)
    tree_util = _mod(
        "jax.tree_util",
        tree_map=tree_map,
        tree_flatten=lambda t: (tree_flatten(t)[0], tree_flatten(t)[1]),
    )
    jsp_special = _mod(
        "jax.scipy.special", logit=scipy.special.logit, expit=scipy.special.expit, xlogy=scipy.special.xlogy
    )
    jsp_stats = _mod("jax.scipy.stats", norm=scipy.stats.norm)
    jsp = _mod("jax.scipy", special=jsp_special, stats=jsp_stats)
    nn = _mod("jax.nn", softplus=lambda x: np.logaddexp(0.0, x))
    flatten_util = _mod("jax.flatten_util", ravel_pytree=ravel_pytree)
    typing_mod = _mod("jax.typing", ArrayLike=object)
    config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax = _mod(
        "jax",
        numpy=jnp,
        lax=lax,
        tree=tree,
        tree_util=tree_util,
        scipy=jsp,
        nn=nn,
        flatten_util=flatten_util,
        typing=typing_mod,
        config=config,
        vmap=vmap,
        jit=_identity_decorator,
        remat=_identity_decorator,
        Array=np.ndarray,
    )
    jax.__path__ = []
    mods = {
        "jax": jax,
        "jax.numpy": jnp,
        "jax.lax": lax,
        "jax.tree": tree,
        "jax.tree_util": tree_util,
        "jax.scipy": jsp,
        "jax.scipy.special": jsp_special,
        "jax.scipy.stats": jsp_stats,
        "jax.nn": nn,
        "jax.flatten_util": flatten_util,
        "jax.typing": typing_mod,
        "jax_dataclasses": _mod("jax_dataclasses", pytree_dataclass=dataclasses.dataclass, Static=_Static),
    }
    for absent in ("demes", "msprime", "pysam", "tskit", "tszip", "intervaltree"):
        stub = _Anything(absent)
        stub.__spec__ = importlib.machinery.ModuleSpec(absent, None)
        mods[absent] = stub
    sys.modules.update(mods)
    pkg = types.ModuleType("phlash")
    pkg.__path__ = [f"{reference_src}/phlash"]
    pkg.__spec__ = importlib.machinery.ModuleSpec("phlash", None, is_package=True)
    sys.modules["phlash"] = pkg
