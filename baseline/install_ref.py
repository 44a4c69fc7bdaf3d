"""Install the UNMODIFIED reference next to the repo so that it travels to the GPU box (git-ignored, NOT gpurun-ignored).

The reference (underdoc-wang/MPGCN) is six flat Python modules with no setup.py / pyproject.toml -- `pip install
--no-index --target baseline/_ref /root/reference` fails with "Directory '/root/reference' is not installable" (recorded in
DESIGN.md) -- so "installing" it is copying its files to `baseline/_ref/` and putting that directory on sys.path, exactly how
its own Main.py finds its siblings.  Nothing under baseline/_ref is tracked or edited; bench.py --impl reference and the
trainer drop-in test import it from there.  Run in the build container (where /root/reference is mounted):

    python baseline/install_ref.py
"""
import os
import shutil
import sys

SRC = os.environ.get("MPGCN_REFERENCE_DIR", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def main() -> int:
    if not os.path.isdir(SRC):
        print(f"reference not found at {SRC}: keeping whatever is in {DST}")
        return 0
    os.makedirs(DST, exist_ok=True)
    n = 0
    for f in sorted(os.listdir(SRC)):
        if f.endswith(".py") or f in ("LICENSE", "README.md"):
            shutil.copy2(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    print(f"copied {n} files from {SRC} to {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
