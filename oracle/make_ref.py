"""Bundle the reference's own Python sources for the hot path into oracle/_ref/reference_py.zip (git-ignored; it travels
to the GPU box like a built .so).

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (robertvoy/ComfyUI-Distributed) is pure Python: there is nothing to
compile, so "building" it for the `--impl reference` arm of bench.py means making its code loadable where /root/reference
does not exist.  `python oracle/make_ref.py` (called by __graft_entry__.build() in the build container) packs the
packages the path imports -- utils/, upscale/, api/, nodes/ (*.py only), UNMODIFIED -- into one archive under
oracle/_ref/, which .gitignore lists: no reference source enters the repository or its history.  At run time
`staged_root()` unpacks the archive into a temporary directory and oracle/ref_loader.py / oracle/ref_static_run.py load
the modules from there under the same ComfyUI stand-ins they use against /root/reference.  Nothing under
comfyui-distributed_b200/ reads it.
"""
from __future__ import annotations

import atexit
import os
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("USDU_REFERENCE_SRC", "/root/reference")
ARCHIVE = os.path.join(HERE, "_ref", "reference_py.zip")
PACKAGES = ("utils", "upscale", "api", "nodes")
_unpacked = None


def staged_root() -> str:
    """A directory the reference can be loaded from: the real tree when present, else the unpacked archive, else ''."""
    global _unpacked
    if os.path.isfile(os.path.join(SRC, "upscale", "modes", "static.py")):
        return SRC
    if _unpacked is None and os.path.isfile(ARCHIVE):
        _unpacked = tempfile.mkdtemp(prefix="usdu_ref_")
        atexit.register(shutil.rmtree, _unpacked, ignore_errors=True)
        with zipfile.ZipFile(ARCHIVE) as z:
            z.extractall(_unpacked)
    return _unpacked or ""


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"make_ref: {SRC} not present (GPU box): using {ARCHIVE} as staged")
        return 0
    os.makedirs(os.path.dirname(ARCHIVE), exist_ok=True)
    n = 0
    with zipfile.ZipFile(ARCHIVE, "w", zipfile.ZIP_DEFLATED) as z:
        for pkg in PACKAGES:
            for base, _, files in os.walk(os.path.join(SRC, pkg)):
                for f in sorted(files):
                    if f.endswith(".py"):
                        full = os.path.join(base, f)
                        z.write(full, os.path.relpath(full, SRC))
                        n += 1
    print(f"make_ref: packed {n} reference files into {ARCHIVE}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
