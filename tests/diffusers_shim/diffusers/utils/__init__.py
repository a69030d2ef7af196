import logging as _pylogging
from collections import OrderedDict
from dataclasses import fields, is_dataclass

import torch
from packaging import version


class BaseOutput(OrderedDict):
    """dataclass-backed output with attribute, key and index access (diffusers.utils.BaseOutput)."""

    def __post_init__(self):
        assert is_dataclass(self)
        for f in fields(self):
            v = getattr(self, f.name)
            if v is not None:
                OrderedDict.__setitem__(self, f.name, v)

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def to_tuple(self):
        return tuple(OrderedDict.__getitem__(self, k) for k in self.keys())


class _Logging:
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


logging = _Logging()


def is_torch_version(op: str, v: str) -> bool:
    cur = version.parse(version.parse(torch.__version__).base_version)
    ref = version.parse(v)
    return {">=": cur >= ref, ">": cur > ref, "<=": cur <= ref, "<": cur < ref, "==": cur == ref}[op]
