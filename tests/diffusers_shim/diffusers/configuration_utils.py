"""ConfigMixin / register_to_config: records the constructor arguments in `self.config` (attribute + item access)."""
import functools
import inspect


class FrozenDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        cfg = dict(getattr(self, "_internal_dict", {}))
        cfg.update(kwargs)
        self._internal_dict = FrozenDict(cfg)

    @property
    def config(self):
        return self._internal_dict


def register_to_config(init):
    @functools.wraps(init)
    def inner(self, *args, **kwargs):
        sig = inspect.signature(init)
        params = [p for n, p in sig.parameters.items() if n != "self"]
        cfg = {p.name: p.default for p in params if p.default is not inspect.Parameter.empty}
        for p, a in zip(params, args):
            cfg[p.name] = a
        cfg.update({k: v for k, v in kwargs.items() if not k.startswith("_")})
        init(self, *args, **kwargs)
        ConfigMixin.register_to_config(self, **cfg)

    return inner
