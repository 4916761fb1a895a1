"""Stage the UNMODIFIED reference files of the hot path under oracle/_ref/ (git-ignored, NOT
gpurun-ignored) so that the GPU box -- where /root/reference does not exist -- can

  * run `bench.py --impl reference` / `cpu_baseline` through the reference's own files
    (`kind: "reference"`), and
  * run the reference's `base_model.py` trainer unmodified on top of the drop-in (tests).

TEST / BENCH INFRASTRUCTURE ONLY.  Nothing is edited: byte-for-byte copies, mirrored at
oracle/_ref/Static/<setting>/..., of
    Static/{transductive,inductive}/{models,load_data,utils,base_model,train}.py
    Static/transductive/data/family, Static/inductive/data/fb237_v2, fb237_v2_ind
The copies never enter git history (`.gitignore: oracle/_ref/`); `__graft_entry__.build()` calls
`stage()` whenever the reference tree is mounted.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC = os.environ.get("REDGNN_REFERENCE", "/root/reference")

FILES = ["models.py", "load_data.py", "utils.py", "base_model.py", "train.py"]
DATA = {"transductive": ["family"], "inductive": ["fb237_v2", "fb237_v2_ind"]}


def source_available():
    return os.path.isfile(os.path.join(SRC, "Static", "transductive", "models.py"))


def staged():
    return os.path.isfile(os.path.join(DEST, "Static", "transductive", "models.py"))


def _copy(src, dst):
    if os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False):
        return
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    shutil.copyfile(src, dst)
    os.chmod(dst, 0o644)


def stage():
    """Copy the files listed above; returns the number of files staged (0 if no source tree)."""
    if not source_available():
        return 0
    n = 0
    for setting in ("transductive", "inductive"):
        base = os.path.join(SRC, "Static", setting)
        for f in FILES:
            _copy(os.path.join(base, f), os.path.join(DEST, "Static", setting, f))
            n += 1
        for ds in DATA[setting]:
            for f in sorted(os.listdir(os.path.join(base, "data", ds))):
                _copy(os.path.join(base, "data", ds, f), os.path.join(DEST, "Static", setting, "data", ds, f))
                n += 1
    return n


if __name__ == "__main__":
    print("staged %d files under %s" % (stage(), DEST))
