"""Registry with the surface the reference uses: `@MODELS.register_module()` and
`MODELS.build(cfg_dict)` where cfg['type'] is the class name and the rest are kwargs
(rein/models/backbones/lora_backbone.py:15, rein/models/heads/VFMHead.py:51). BACKBONES is the same
registry object as MODELS, as in mmseg.models.builder.

When the real mmseg is importable the classes are additionally registered there under
'B200<Name>' so a rein config can switch with one `type=` edit.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, Optional


class Registry:
    def __init__(self, name: str):
        self.name = name
        self._modules: Dict[str, type] = {}

    @property
    def module_dict(self) -> Dict[str, type]:
        return self._modules

    def register_module(self, name: Optional[str] = None, force: bool = False, module: Optional[type] = None) -> Callable:
        def _register(cls):
            names = [name] if isinstance(name, str) else (name or [cls.__name__])
            for n in names:
                if n in self._modules and not force and self._modules[n] is not cls:
                    raise KeyError(f"{n} is already registered in {self.name}")
                self._modules[n] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def get(self, key: str) -> Optional[type]:
        return self._modules.get(key)

    def build(self, cfg: Any, **default_args):
        if not isinstance(cfg, dict):
            return cfg  # already built (mmengine accepts modules too)
        if "type" not in cfg:
            raise KeyError(f"cfg for registry {self.name} needs a 'type' key, got {sorted(cfg)}")
        args = dict(cfg)
        for k, v in default_args.items():
            args.setdefault(k, v)
        typ = args.pop("type")
        cls = self._modules.get(typ) if isinstance(typ, str) else typ
        if cls is None:
            raise KeyError(f"{typ} is not in the {self.name} registry (have: {sorted(self._modules)})")
        return cls(**args)


MODELS = Registry("model")
BACKBONES = MODELS
METRICS = Registry("metric")


def register_into_mmseg() -> bool:
    """Mirror our classes into mmseg's registries as 'B200<Name>' when mmseg is installed."""
    try:
        from mmseg.registry import METRICS as MM_METRICS, MODELS as MM_MODELS  # type: ignore
    except Exception:
        return False
    for n, c in MODELS.module_dict.items():
        if not n.startswith("B200"):
            MM_MODELS.register_module(name="B200" + n, module=c, force=True)
    for n, c in METRICS.module_dict.items():
        if not n.startswith("B200"):
            MM_METRICS.register_module(name="B200" + n, module=c, force=True)
    return True


class ConfigDict(dict):
    """dict with attribute access, like mmengine's ConfigDict (`self.test_cfg.stride`)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v
