"""`net_opt()` — default upscale factor / training flag for arch constructors.

The reference re-parses the option file at import time in every arch module
(neosr/archs/arch_util.py:12-27).  Here the TOML named by `-opt` on the command line is read
once, lazily; without it the defaults are scale 4 / training, and `set_default_scale()` lets
a host program (bench, tests, the model factory) set it explicitly."""
from __future__ import annotations

import sys
import tomllib
from pathlib import Path

_default = {"scale": None, "training": True}


def set_default_scale(scale: int, training: bool = True) -> None:
    _default["scale"], _default["training"] = int(scale), training


def net_opt() -> tuple[int, bool]:
    if _default["scale"] is None:
        scale, training = 4, True
        if "-opt" in sys.argv:
            try:
                path = Path(sys.argv[sys.argv.index("-opt") + 1])
                with path.open("rb") as f:
                    opt = tomllib.load(f)
                scale = int(opt.get("scale", 4))
                training = "train" in opt.get("datasets", {"train": 1})
            except (OSError, IndexError, ValueError, tomllib.TOMLDecodeError):
                pass
        _default["scale"], _default["training"] = scale, training
    return _default["scale"], _default["training"]
