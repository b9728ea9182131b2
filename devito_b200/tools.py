"""The handful of `devito.tools` helpers user code imports."""
from functools import wraps

__all__ = ['as_tuple', 'memoized_meth', 'Pickable', 'filter_ordered', 'flatten']


def as_tuple(item, type=None, length=None):
    if item is None:
        t = ()
    elif isinstance(item, (str, bytes)):
        t = (item,)
    else:
        try:
            t = tuple(item)
        except TypeError:
            t = (item,) * (length or 1)
    if length and len(t) != length:
        raise ValueError(f"Tuple needs to be of length {length}")
    if type and not all(isinstance(i, type) for i in t):
        raise TypeError(f"Items need to be of type {type}")
    return t


def memoized_meth(meth):
    cache_name = f'_memo_{meth.__name__}'

    @wraps(meth)
    def wrapper(self, *args, **kwargs):
        cache = self.__dict__.setdefault(cache_name, {})
        key = (args, tuple(sorted(kwargs.items())))
        try:
            return cache[key]
        except KeyError:
            cache[key] = val = meth(self, *args, **kwargs)
            return val
        except TypeError:
            return meth(self, *args, **kwargs)
    return wrapper


class Pickable:
    """Minimal stand-in: objects rebuild from `__rargs__`/`__rkwargs__`."""
    __rargs__ = ()
    __rkwargs__ = ()


def filter_ordered(elements):
    seen = set()
    out = []
    for e in elements:
        if e not in seen:
            seen.add(e)
            out.append(e)
    return out


def flatten(items):
    out = []
    for i in items:
        if isinstance(i, (list, tuple)):
            out.extend(flatten(i))
        else:
            out.append(i)
    return out
