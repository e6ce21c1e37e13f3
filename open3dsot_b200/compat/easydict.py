"""Attribute-accessible dict with the subset of `easydict.EasyDict` behaviour the reference relies on
(main.py:49, models/base_model.py:20-21): `cfg.key`, `cfg['key']`, `getattr(cfg, 'key', default)`,
nested dicts converted recursively."""
try:  # prefer the real package when present
    from easydict import EasyDict  # type: ignore # noqa: F401
except Exception:  # pragma: no cover - exercised in this image

    class EasyDict(dict):
        def __init__(self, d=None, **kwargs):
            super().__init__()
            d = dict(d or {})
            d.update(kwargs)
            for k, v in d.items():
                self[k] = v

        @classmethod
        def _wrap(cls, v):
            if isinstance(v, dict) and not isinstance(v, cls):
                return cls(v)
            if isinstance(v, (list, tuple)):
                return type(v)(cls._wrap(x) for x in v)
            return v

        def __setitem__(self, k, v):
            super().__setitem__(k, self._wrap(v))

        def __setattr__(self, k, v):
            self[k] = v

        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __delattr__(self, k):
            del self[k]

        def update(self, *a, **kw):
            for k, v in dict(*a, **kw).items():
                self[k] = v
