#!/usr/bin/env python3
"""Put the UNMODIFIED reference where the GPU box can import it: baseline/_ref (git-ignored, travels with gpurun).

    python tools/install_reference.py

First the documented way (pip --no-index --target baseline/_ref /root/reference); the reference's build backend is
hatchling, which is neither installed nor in /opt/wheelhouse, so pip cannot build the wheel here.  The reference is a
pure-Python package (SURVEY headline 1: not one compiled file), so the wheel would contain exactly the `betse/` tree:
the fallback copies that tree verbatim.  Nothing under baseline/_ref is product source or oracle; it is used by
tests/test_gpu_dropin.py (reference loop vs drop-in on the same GPU box) and is never committed."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("BETSE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(os.path.join(SRC, "betse", "science")):
        print("reference sources not found at", SRC)
        return 1
    os.makedirs(DST, exist_ok=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
           "--find-links", "/opt/wheelhouse", "--target", DST, SRC]
    res = subprocess.run(cmd, capture_output=True, text=True)
    how = "pip"
    if res.returncode != 0 or not os.path.isdir(os.path.join(DST, "betse", "science")):
        why = (res.stderr.strip().splitlines() or ["?"])[-1]
        print("pip install failed (%s); copying the pure-Python package tree instead" % why)
        shutil.rmtree(os.path.join(DST, "betse"), ignore_errors=True)
        shutil.copytree(os.path.join(SRC, "betse"), os.path.join(DST, "betse"),
                        ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        how = "copytree (pip: %s)" % why
    with open(os.path.join(DST, "INSTALLED_BY"), "w") as f:
        f.write(how + "\n")
    print("reference available at", DST, "via", how)
    return 0


if __name__ == "__main__":
    sys.exit(main())
