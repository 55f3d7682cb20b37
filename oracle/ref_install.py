"""TEST / BENCH INFRASTRUCTURE -- ships the UNMODIFIED reference to the GPU box.

The reference is pure Python and has no build: ``install()`` copies its module files from the
read-only tree at /root/reference (present in the build container only) into the git-ignored
directory ``oracle/_ref/pyellipticfd/`` -- never into history -- so that the snapshot gpurun sends
to the GPU box contains it.  ``oracle/refshim.py`` imports it from there
when /root/reference does not exist.  It is used as the checker and as the CPU reference arm of
``bench.py --impl reference`` (``cpu_baseline.kind = "reference"``); nothing in the product imports it.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref", "pyellipticfd")
SOURCE = os.environ.get("PYELLIPTICFD_REFERENCE", "/root/reference")
MODULES = ["__init__.py", "_ddutils.py", "convex_envelope.py", "ddf.py", "ddg.py", "ddi.py",
           "pointclasses.py", "pucci.py", "solvers.py", "stencils.py"]


def available():
    return os.path.exists(os.path.join(DEST, "ddg.py"))


def install():
    """Copy the reference modules (byte for byte) when the source tree is here; otherwise keep
    whatever a previous build left.  Returns True if ``oracle/_ref`` holds the reference."""
    if os.path.exists(os.path.join(SOURCE, "ddg.py")):
        os.makedirs(DEST, exist_ok=True)
        for name in MODULES:
            src = os.path.join(SOURCE, name)
            if os.path.exists(src):
                shutil.copyfile(src, os.path.join(DEST, name))
    return available()


if __name__ == "__main__":
    print(install(), DEST)
