"""Minimal omegaconf stand-in (attribute dicts, YAML load, merge, dotlist).  Test infrastructure only."""
import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, _wrap(v))

    def __delattr__(self, k):
        del self[k]

    def copy(self):
        return _wrap(_unwrap(self))


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, DictConfig):
        d = DictConfig()
        for k, x in v.items():
            d[k] = x
        return d
    if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
        return ListConfig(_wrap(x) for x in v)
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


def _merge(a, b):
    out = _wrap(_unwrap(a))
    for k, v in b.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = _wrap(_unwrap(v))
    return out


class OmegaConf:
    @staticmethod
    def create(obj=None):
        return _wrap(obj if obj is not None else {})

    @staticmethod
    def load(path):
        with open(path) as f:
            return _wrap(yaml.safe_load(f) or {})

    @staticmethod
    def merge(*cfgs):
        out = DictConfig()
        for c in cfgs:
            out = _merge(out, c)
        return out

    @staticmethod
    def from_dotlist(items):
        out = DictConfig()
        for it in items:
            key, val = it.split("=", 1)
            val = yaml.safe_load(val)
            cur = out
            parts = key.split(".")
            for p in parts[:-1]:
                if p not in cur:
                    cur[p] = DictConfig()
                cur = cur[p]
            cur[parts[-1]] = val
        return out

    @staticmethod
    def from_cli(args=None):
        import sys
        return OmegaConf.from_dotlist(args if args is not None else sys.argv[1:])

    @staticmethod
    def to_container(cfg, **kw):
        return _unwrap(cfg)

    @staticmethod
    def to_yaml(cfg):
        return yaml.safe_dump(_unwrap(cfg))

    @staticmethod
    def save(cfg, path):
        with open(path, "w") as f:
            f.write(OmegaConf.to_yaml(cfg))

    @staticmethod
    def is_missing(cfg, key):
        return key not in cfg
