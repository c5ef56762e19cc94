"""Stub: names only; the storage layer is never executed by the golden generator."""
from abc import ABCMeta


class _Meta(ABCMeta):
    def __new__(mcls, name, bases, ns, **kw):
        return super().__new__(mcls, name, bases, ns)

    def __init__(cls, name, bases, ns, **kw):
        super().__init__(name, bases, ns)


class SQLModel(metaclass=_Meta):
    metadata = None


def Field(*a, default=None, **k):
    return default


class Column:
    def __init__(self, *a, **k):
        pass


class LargeBinary:
    pass


class String:
    pass


class TypeDecorator:
    impl = None
    cache_ok = True

    def __class_getitem__(cls, item):
        return cls


class Session:
    def __init__(self, *a, **k):
        pass


def create_engine(*a, **k):
    return None


def select(*a, **k):
    return None
