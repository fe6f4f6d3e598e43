"""Config helpers.  ``DotDict`` keeps the reference's cfg contract (common/utils/utils.py:257-291):
attribute access on nested dicts; ``merge_cfg`` is the one-level-deep user override every reference
env applies in ``__init__`` (ship_ice_env.py:47-56)."""
from __future__ import annotations

import copy


class DotDict(dict):
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __getattr__(self, attr):
        if attr not in self.keys():
            raise AttributeError(attr)
        return self.get(attr)

    @staticmethod
    def to_dict(d):
        return {k: DotDict.to_dict(v) if type(v) is DotDict else v for k, v in d.items()}

    @staticmethod
    def to_dot_dict(d):
        return DotDict({k: DotDict.to_dot_dict(v) if type(v) is dict else v for k, v in d.items()})

    @staticmethod
    def load_from_file(fp):
        import yaml
        with open(fp, "r") as fd:
            cfg = yaml.load(fd, Loader=yaml.Loader)
        return DotDict.to_dot_dict(cfg)


def merge_cfg(base: dict, cfg) -> DotDict:
    out = DotDict.to_dot_dict(copy.deepcopy(DotDict.to_dict(base) if isinstance(base, DotDict) else base))
    if cfg is not None:
        for cfg_type in cfg:
            if type(cfg[cfg_type]) is DotDict or type(cfg[cfg_type]) is dict:
                if cfg_type not in out:
                    out[cfg_type] = DotDict()
                for param in cfg[cfg_type]:
                    out[cfg_type][param] = cfg[cfg_type][param]
            else:
                out[cfg_type] = cfg[cfg_type]
    return out
