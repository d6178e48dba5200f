"""``mx.registry`` — name → class registries with ``register`` / ``alias`` / ``create`` helpers
(parity: python/mxnet/registry.py:30-175; used by optimizers, initializers, metrics and user extensions)."""
from __future__ import annotations

import json
import warnings

_REGISTRY = {}


def get_registry(base_class):
    return dict(_REGISTRY.setdefault(base_class, {}))


def get_register_func(base_class, nickname):
    reg = _REGISTRY.setdefault(base_class, {})

    def register(klass, name=None):
        assert issubclass(klass, base_class), "Can only register subclass of %s" % base_class.__name__
        name = (name or klass.__name__).lower()
        if name in reg and reg[name] is not klass:
            warnings.warn("New %s %s.%s registered with name %s is overriding existing %s %s.%s" % (
                nickname, klass.__module__, klass.__name__, name, nickname, reg[name].__module__, reg[name].__name__), UserWarning, stacklevel=2)
        reg[name] = klass
        return klass
    register.__doc__ = "Register %s to the %s factory" % (nickname, nickname)
    return register


def get_alias_func(base_class, nickname):
    register = get_register_func(base_class, nickname)

    def alias(*aliases):
        def reg(klass):
            for a in aliases:
                register(klass, a)
            return klass
        return reg
    return alias


def get_create_func(base_class, nickname):
    reg = _REGISTRY.setdefault(base_class, {})

    def create(*args, **kwargs):
        """Accepts an instance (returned as is), a name, a ``'["name", {kwargs}]'`` JSON string, or ``(name, **kwargs)``."""
        if args:
            name, args = args[0], args[1:]
        else:
            name = kwargs.pop(nickname)
        if isinstance(name, base_class):
            assert not args and not kwargs, "%s is already an instance. Additional arguments are invalid" % nickname
            return name
        if isinstance(name, dict):
            return create(**name)
        assert isinstance(name, str), "%s must be of string type" % nickname
        if name.startswith("["):
            assert not args and not kwargs
            name, kwargs = json.loads(name)
            return create(name, **kwargs)
        if name.startswith("{"):
            assert not args and not kwargs
            return create(**json.loads(name))
        name = name.lower()
        assert name in reg, "%s is not registered. Please register with %s.register first" % (name, nickname)
        return reg[name](*args, **kwargs)
    create.__doc__ = "Create a %s instance from config." % nickname
    return create
