#!/usr/bin/env python
"""Stage the UNMODIFIED reference (muslll/neosr) under the git-ignored `baseline/_ref/` so that it travels to
the GPU box with the repo snapshot (`/root/reference` exists only in the build container).

    python baseline/install_ref.py            # from /root/reference (or $NEOSR_REFERENCE)

First tries the contract's offline install (`pip install --no-index --no-build-isolation --find-links
/opt/wheelhouse --target baseline/_ref <reference>`).  In this image that fails: the reference's build backend is
`poetry-core` (pyproject.toml:46-48), which is neither installed nor in the wheelhouse.  The package is pure Python
with no build step, so the fallback stages exactly what such an install would put there - the `neosr/` package
tree, byte for byte - plus `train.py`, `test.py` and `options/` so the reference's own entry point can be run.
Also writes `_stubs/{pywt,lmdb}.py`: empty stand-ins for two dependencies this image lacks (they are imported at
module load by neosr/losses/wavelet_guided.py and neosr/data/file_client.py but used only by the wavelet-guided
and LMDB code paths, which nothing here exercises).  Nothing under baseline/_ref is tracked by git; nothing in the
product (`neosr_b200/`) imports it.  Users: `bench.py --impl reference`, tests/test_dropin_*.py.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
DEST = HERE / "_ref"


def install(src: Path | None = None, quiet: bool = False) -> str:
    src = Path(src or os.environ.get("NEOSR_REFERENCE", "/root/reference"))
    if not (src / "neosr" / "models" / "image.py").exists():
        raise FileNotFoundError(f"reference tree not found at {src}")
    if DEST.exists():
        shutil.rmtree(DEST)
    how = "pip"
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                        "--find-links", "/opt/wheelhouse", "--target", str(DEST), str(src)],
                       capture_output=True, text=True)
    if r.returncode != 0 or not (DEST / "neosr").exists():
        how = "staged (pip failed: " + (r.stderr.strip().splitlines() or ["?"])[-1][:120] + ")"
        if DEST.exists():
            shutil.rmtree(DEST)
        DEST.mkdir(parents=True)
        shutil.copytree(src / "neosr", DEST / "neosr", ignore=shutil.ignore_patterns("__pycache__"))
    for f in ("train.py", "test.py"):
        shutil.copy2(src / f, DEST / f)
    shutil.copytree(src / "options", DEST / "options", dirs_exist_ok=True)
    stubs = DEST / "_stubs"
    stubs.mkdir(exist_ok=True)
    (stubs / "pywt.py").write_text('"""stub: PyWavelets is not in this image (wavelet-guided path unused)."""\n')
    (stubs / "lmdb.py").write_text('"""stub: lmdb is not in this image (LMDB backend unused)."""\n')
    (DEST / "INSTALLED_FROM").write_text(f"{src}\n{how}\n")
    if not quiet:
        print(f"baseline/_ref: {how}")
    return how


if __name__ == "__main__":
    install(Path(sys.argv[1]) if len(sys.argv) > 1 else None)
