"""Name -> object registries with the reference's semantics (neosr/utils/registry.py:8-107):
registration by `__name__`, duplicate names are an error, lookup falls back to name + "_neosr"."""
from __future__ import annotations


class Registry:
    def __init__(self, name: str) -> None:
        self._name = name
        self._obj_map: dict = {}

    def _do_register(self, name: str, obj, suffix: str | None = None) -> None:
        if isinstance(suffix, str):
            name = name + "_" + suffix
        if name in self._obj_map:
            raise AssertionError(f"An object named '{name}' was already registered in '{self._name}' registry!")
        self._obj_map[name] = obj

    def register(self, obj=None, suffix: str | None = None):
        if obj is None:
            def deco(fn_or_cls):
                self._do_register(fn_or_cls.__name__, fn_or_cls, suffix)
                return fn_or_cls
            return deco
        self._do_register(obj if isinstance(obj, str) else obj.__name__, obj, suffix)
        return None

    def get(self, name: str, suffix: str = "neosr"):
        ret = self._obj_map.get(name)
        if ret is None:
            ret = self._obj_map.get(name + "_" + suffix)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name: str) -> bool:
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())

    def keys(self):
        return self._obj_map.keys()


ARCH_REGISTRY = Registry("arch")
LOSS_REGISTRY = Registry("loss")
MODEL_REGISTRY = Registry("model")
