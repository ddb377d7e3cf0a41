"""A small record-of-arrays container with the semantics the hot path relies on.

Inside an SRL checkout the trainer receives `base.namedarray.NamedArray` objects (base/namedarray.py:221-538);
this module is the stand-alone mirror used when SRL itself is not importable (tests, bench, the GPU box).
`MultiAgentPPOB200` only uses the duck-typed surface both share:

  * attribute / item access by field name, nested records, `None` leaves;
  * `x[slice]` applies the slice to every leaf (base/namedarray.py:346-362);
  * keys iterate in SORTED order (base/namedarray.py:282) -- this fixes the leaf order of flatten();
  * `flatten` / `from_flattened` with dotted names (base/namedarray.py:663-692);
  * `recursive_aggregate` zero-fills leaves that are None in only some elements (base/namedarray.py:588-633);
  * `recursive_apply` maps a function over the leaves (base/namedarray.py:636-660).
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Iterator, List, Optional, Tuple

import numpy as np

try:  # torch is optional for this module
    import torch
    _TENSOR = (np.ndarray, torch.Tensor)
except ImportError:  # pragma: no cover
    torch = None
    _TENSOR = (np.ndarray,)


class NamedArray:

    def __init__(self, **fields):
        object.__setattr__(self, "_fields", {k: fields[k] for k in sorted(fields)})
        object.__setattr__(self, "_metadata", {})

    # -- metadata (not sliced, not aggregated; base/namedarray.py:262-279) -------------------------
    def register_metadata(self, **kw) -> None:
        for k in self._fields:
            if k in kw:  # base/namedarray.py:292-294
                raise KeyError("Keys of metadata should be different from data fields!")
        self._metadata.update(kw)

    def pop_metadata(self, key):
        return self._metadata.pop(key)

    def clear_metadata(self) -> None:
        self._metadata.clear()

    @property
    def metadata(self) -> Dict[str, Any]:
        return self._metadata

    # -- mapping surface --------------------------------------------------------------------------------
    def keys(self):
        return self._fields.keys()

    def values(self):
        return self._fields.values()

    def items(self):
        return self._fields.items()

    def __iter__(self) -> Iterator[Any]:
        """The VALUES in key order, as base/namedarray.py:307-309 (not the keys: `zip(record, other)` pairs leaves)."""
        return iter(self._fields.values())

    def __contains__(self, k) -> bool:
        return k in self._fields

    def __len__(self) -> int:
        """Number of fields (base/namedarray.py:417-418); `length()` is the size along an array dimension."""
        return len(self._fields)

    def length(self, dim: int = 0) -> int:
        """Size of the first non-None leaf along `dim` (base/namedarray.py:420-430)."""
        for v in self._fields.values():
            if v is None:
                continue
            return v.length(dim) if isinstance(v, NamedArray) else v.shape[dim]
        raise IndexError("No entries in the NamedArray.")

    def __getattr__(self, name):
        fields = object.__getattribute__(self, "_fields")
        if name in fields:
            return fields[name]
        meta = object.__getattribute__(self, "_metadata")
        if name in meta:
            return meta[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if name in self._fields:
            self._fields[name] = value
        elif name in self._metadata:
            self._metadata[name] = value
        else:
            raise AttributeError(f"{type(self).__name__} has no field {name!r} (fields are fixed at construction)")

    def __getitem__(self, key):
        if isinstance(key, str):
            return self._fields[key]
        out = {k: (None if v is None else v[key]) for k, v in self._fields.items()}
        new = object.__new__(type(self))
        object.__setattr__(new, "_fields", out)
        object.__setattr__(new, "_metadata", dict(self._metadata))
        return new

    def __setitem__(self, key, value):
        if isinstance(key, str):
            self.__setattr__(key, value)
            return
        for k, v in self._fields.items():
            if v is None:
                continue
            src = value[k] if isinstance(value, NamedArray) else value
            if src is not None:
                v[key] = src

    @property
    def shape(self):
        return {k: (None if v is None else v.shape) for k, v in self._fields.items()}

    def __repr__(self):
        body = ", ".join(f"{k}={'None' if v is None else (v if isinstance(v, NamedArray) else tuple(v.shape))}"
                         for k, v in self._fields.items())
        return f"{type(self).__name__}({body})"


def is_record(x) -> bool:
    """True for this module's NamedArray and for SRL's own (duck-typed: has keys() and item access)."""
    return isinstance(x, NamedArray) or (hasattr(x, "keys") and hasattr(x, "__getitem__") and
                                         not isinstance(x, _TENSOR + (dict,)))


def record_class(x) -> type:
    """The generic record class to rebuild records of x's kind with: SRL's own `base.namedarray.NamedArray` when x is
    one (its `recursive_apply` / `from_flattened` return that class too, base/namedarray.py:636-692, and SRL policies
    test `isinstance(..., NamedArray)` against it), this module's mirror otherwise."""
    for c in type(x).__mro__:
        if c.__name__ == "NamedArray" and c is not NamedArray and c is not object:
            return c
    return NamedArray


def from_dict(values: Optional[Dict[str, Any]], cls: type = NamedArray):
    if values is None or len(values) == 0:
        return None
    return cls(**{k: (from_dict(v, cls) if isinstance(v, dict) else v) for k, v in values.items()})


def flatten(x) -> List[Tuple[str, Any]]:
    """[(dotted name, leaf)] in sorted-key depth-first order; None leaves are kept."""
    out = []
    for k in sorted(x.keys()):
        v = x[k]
        if v is not None and is_record(v):
            out += [(f"{k}.{kk}", vv) for kk, vv in flatten(v)]
        else:
            out.append((k, v))
    return out


def from_flattened(entries: List[Tuple[str, Any]], cls: type = NamedArray):
    tree: Dict[str, Any] = {}
    for name, v in entries:
        node = tree
        parts = name.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = v
    return from_dict(tree, cls)


def recursive_apply(x, fn: Callable, cls: Optional[type] = None):
    """Maps fn over the leaves.  The result is built from `cls` (default: the generic record class of x's own family,
    record_class(x)) -- like the reference's recursive_apply, subclasses such as SampleBatch come back as plain records."""
    if x is None:
        return None
    if is_record(x):
        cls = cls or record_class(x)
        return cls(**{k: recursive_apply(x[k], fn, cls) for k in x.keys()})
    return fn(x)


def _zeros_like(x):
    if is_record(x):
        return NamedArray(**{k: (None if x[k] is None else _zeros_like(x[k])) for k in x.keys()})
    return np.zeros_like(x) if isinstance(x, np.ndarray) else torch.zeros_like(x)


def recursive_aggregate(xs: List[Any], aggregate_fn: Callable):
    """Aggregate a list of records leaf by leaf.  A leaf that is None in some (not all) elements is replaced
    by zeros shaped like a present one first (base/namedarray.py:588-595); None everywhere stays None."""
    present = [x for x in xs if x is not None]
    if not present:
        return None
    if len(present) != len(xs):
        xs = [_zeros_like(present[0]) if x is None else x for x in xs]
    if is_record(xs[0]):
        return NamedArray(**{k: recursive_aggregate([x[k] for x in xs], aggregate_fn) for k in xs[0].keys()})
    return aggregate_fn(xs)


def size_bytes(x) -> int:
    return int(sum(v.nbytes if isinstance(v, np.ndarray) else v.numel() * v.element_size()
                   for _, v in flatten(x) if v is not None))
