"""Stage a verbatim copy of the UNMODIFIED reference into oracle/_ref (git-ignored, NOT gpurun-ignored).

    python -m oracle.stage_ref            # build container only: reads /root/reference

/root/reference does not exist on the GPU box; oracle/_ref travels there with the snapshot exactly like the built
.so files, so that (a) `bench.py --impl reference` can time the reference's own `testing.EulerHeunSamplerDPS`, and
(b) tests/test_reference_integration.py can hand the reference's own operator / EDM objects to the buddy_b200
samplers, as testing/tester.py:32,143-161 does.  Nothing under oracle/_ref is part of the repository's history and
nothing in buddy_b200/ imports it (oracle/ is test infrastructure).  Only the Python packages and the Hydra configs
are staged (no audio examples)."""
import os
import shutil
import sys

SRC = os.environ.get("BUDDY_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PARTS = ["networks", "diff_params", "testing", "utils", "conf", "datasets", "training", "test.py", "requirements.txt"]


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, "networks")):
        return None
    if os.path.isdir(DST) and not force:
        return DST
    tmp = DST + ".tmp"
    shutil.rmtree(tmp, ignore_errors=True)
    os.makedirs(tmp)
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc")
    for p in PARTS:
        s = os.path.join(SRC, p)
        if os.path.isdir(s):
            shutil.copytree(s, os.path.join(tmp, p), ignore=ignore, symlinks=True, ignore_dangling_symlinks=True)
        elif os.path.isfile(s):
            shutil.copy2(s, os.path.join(tmp, p))
    shutil.rmtree(DST, ignore_errors=True)
    os.rename(tmp, DST)
    return DST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
