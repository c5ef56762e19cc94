"""Stub: the four helpers the reference imports (host-side list plumbing)."""
from itertools import chain, islice


def chunked(it, n):
    it = iter(it)
    while True:
        c = list(islice(it, n))
        if not c:
            return
        yield c


def first(it, default=None):
    for x in it:
        return x
    return default


def flatten(it):
    return chain.from_iterable(it)


def split_when(it, pred):
    buf = []
    prev = None
    for i, x in enumerate(it):
        if i and pred(prev, x):
            yield buf
            buf = []
        buf.append(x)
        prev = x
    if buf:
        yield buf
