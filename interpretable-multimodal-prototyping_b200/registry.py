"""Model registry with the reference's surface (medmm/utils/registry.py:7-69, medmm/modeling/models/build.py:3-11):
``MODEL_REGISTRY.register()`` as a decorator, ``build_model(name, verbose=True, **kwargs)``.

``register_into(reference_registry)`` puts this package's factories under the SAME names into the reference's own
``MODEL_REGISTRY``, so ``build_model(cfg.MODEL.NAME, cfg=cfg, num_classes=..., omic_sizes=...)``
in ``MBTRAIN.build_model`` (medmm/engine/mbtrain.py:69-75) returns the B200 model without touching the trainer."""
from __future__ import annotations

from typing import Callable, Dict, Optional


class Registry:
    def __init__(self, name: str):
        self._name = name
        self._obj_map: Dict[str, Callable] = {}

    def _do_register(self, name: str, obj: Callable, force: bool = False) -> None:
        if name in self._obj_map and not force:
            raise KeyError('An object named "{}" was already registered in "{}" registry'.format(name, self._name))
        self._obj_map[name] = obj

    def register(self, obj: Optional[Callable] = None, force: bool = False):
        if obj is None:                       # used as a decorator
            def wrapper(fn_or_class):
                self._do_register(fn_or_class.__name__, fn_or_class, force=force)
                return fn_or_class
            return wrapper
        self._do_register(obj.__name__, obj, force=force)
        return obj

    def get(self, name: str) -> Callable:
        if name not in self._obj_map:
            raise KeyError('Object name "{}" does not exist in "{}" registry'.format(name, self._name))
        return self._obj_map[name]

    def registered_names(self):
        return list(self._obj_map.keys())


MODEL_REGISTRY = Registry("MODEL")


def build_model(name: str, verbose: bool = True, **kwargs):
    avail = MODEL_REGISTRY.registered_names()
    if name not in avail:
        raise ValueError("Model must be one of {}, but got {}".format(avail, name))
    if verbose:
        print("Model name: {}".format(name))
    return MODEL_REGISTRY.get(name)(**kwargs)


def register_into(reference_registry) -> None:
    """Replace the reference's entries by this package's factories (same names, same kwargs)."""
    from . import umeml_gan as _m  # noqa: F401  (registers on import)
    for name in MODEL_REGISTRY.registered_names():
        fn = MODEL_REGISTRY.get(name)
        try:
            reference_registry.register(fn, force=True)
        except TypeError:                     # registries without a `force` switch: overwrite the map entry
            reference_registry._obj_map[name] = fn
