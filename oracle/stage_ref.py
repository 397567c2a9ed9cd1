"""Stage the UNMODIFIED reference packages of the hot path into git-ignored oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

    python oracle/stage_ref.py            (run by __graft_entry__.build() whenever /root/reference is present)

The reference is pure Python (no native code, no build system), so "building" it means copying
the .py / .yaml files of the four packages on the hot path -- audio, video, audio_video, audio_cues_video -- as they
lie under /root/reference.  Nothing is edited and nothing lands in git history: oracle/_ref/ is listed in .gitignore
(not in .gpurunignore, so the staged files travel to the GPU box, where /root/reference does not exist).
Consumers: oracle/ref_loader.py -> bench.py's CPU arm (`cpu_baseline.kind = "reference"`) and the tests that pin the
oracle ports to the reference on the GPU box.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
PACKAGES = ("audio", "video", "audio_video", "audio_cues_video")
KEEP = (".py", ".yaml", ".yml")
SKIP_DIRS = ("metrics", "plots", "__pycache__", "checkpoints", "logs")


def stage(src=SRC, dst=DST, verbose=False):
    """Copy the four packages; returns the number of files staged (0 when the reference is not present)."""
    if not os.path.isdir(src):
        return 0
    n = 0
    for pkg in PACKAGES:
        for root, dirs, files in os.walk(os.path.join(src, pkg)):
            dirs[:] = [d for d in dirs if d not in SKIP_DIRS]
            rel = os.path.relpath(root, src)
            for f in files:
                if not f.endswith(KEEP):
                    continue
                out_dir = os.path.join(dst, rel)
                os.makedirs(out_dir, exist_ok=True)
                s, d = os.path.join(root, f), os.path.join(out_dir, f)
                if not os.path.exists(d) or os.path.getmtime(d) < os.path.getmtime(s) or os.path.getsize(d) != os.path.getsize(s):
                    shutil.copy2(s, d)
                n += 1
    with open(os.path.join(dst, "STAGED_FROM"), "w") as f:
        f.write(f"{src}\nunmodified copies of {', '.join(PACKAGES)} (*.py, *.yaml); made by oracle/stage_ref.py\n")
    if verbose:
        print(f"staged {n} reference files into {dst}")
    return n


if __name__ == "__main__":
    sys.exit(0 if stage(verbose=True) else 1)
