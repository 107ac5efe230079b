"""Content hash of the CUDA sources a profile was taken from. profiles/ktrace_*.json carry it as "csrc_sha"; bench.py prints
"stale": true next to every number it forwards from such a file when the sources have changed since (the .git directory does not
travel to the GPU box, so this hashes file contents, not commits)."""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def csrc_sha():
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(ROOT, "lumen_b200", "csrc", "*.cu")) + glob.glob(os.path.join(ROOT, "lumen_b200", "csrc", "*.cuh"))
                   + glob.glob(os.path.join(ROOT, "lumen_b200", "csrc", "*.h")) + glob.glob(os.path.join(ROOT, "include", "*.h")))
    for f in files:
        h.update(os.path.basename(f).encode())
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(csrc_sha())
