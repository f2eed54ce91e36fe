"""Install the UNMODIFIED reference package into the git-ignored ``oracle/_ref/`` -- TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py            # in the build container (needs /root/reference)

The reference is pure Python, so "building" it is ``pip install --no-deps --target oracle/_ref`` from a scratch copy
of /root/reference (the checkout itself is read-only; no source file is copied into the tracked tree).  ``oracle/_ref``
is listed in .gitignore but not in .gpurunignore, so it travels to the GPU box with the snapshot, where

* ``bench.py --impl reference`` times the reference's own ``SuperResolutionNet`` on the host cores
  (``cpu_baseline.kind = "reference"``; without ``oracle/_ref`` it falls back to the oracle port, kind "port"),
* ``bench.py``'s ``gpu_eager`` leg runs the same unmodified module through ATen/cuDNN on the B200 (the same-box
  GPU comparator of SURVEY.md section 8d), and
* ``tests/test_reference_live_gpu.py`` compares the CUDA path with the live reference module on the GPU.

Only ``tests/``, ``__graft_entry__`` and ``bench.py`` may import from ``oracle/_ref``; the product never does.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def ref_path():
    """Path to put on sys.path to import the live reference as ``nerve_cl``, or None if it was not installed."""
    return DST if os.path.isdir(os.path.join(DST, "nerve_cl")) else None


def build(force: bool = False) -> bool:
    if not os.path.isdir(REF_SRC):
        return ref_path() is not None
    if ref_path() and not force:
        return True
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, work)
        shutil.rmtree(DST, ignore_errors=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DST, work]
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            sys.stderr.write(proc.stdout[-2000:] + proc.stderr[-2000:])
            return False
    return ref_path() is not None


if __name__ == "__main__":
    ok = build(force="--force" in sys.argv)
    print("oracle/_ref:", "ready" if ok else "NOT available")
    sys.exit(0 if ok else 1)
