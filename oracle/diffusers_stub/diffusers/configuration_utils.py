"""ConfigMixin / register_to_config: constructor kwargs are recorded in `self.config`."""
import functools
import inspect


class _Config(dict):
    __getattr__ = dict.__getitem__


class ConfigMixin:
    config_name = "config.json"

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        bound = sig.bind(self, *args, **kwargs)
        bound.apply_defaults()
        cfg = {k: v for k, v in bound.arguments.items() if k not in ("self", "kwargs")}
        object.__setattr__(self, "_internal_dict", _Config(cfg))
        init(self, *args, **kwargs)
    return inner
