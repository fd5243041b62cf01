"""Recipe for baseline/_ref: the reference's hot-path modules, UNMODIFIED, for the reference arm of bench.py
and the drop-in test on the GPU box.

    python baseline/install_ref.py            # in the build container (needs /root/reference)

The reference is a directory of plain Python modules - no setup.py / pyproject, so `pip install --target
baseline/_ref /root/reference` has nothing to build (recorded in DESIGN.md); the install is a byte-for-byte
copy of the files of SURVEY 8a/8f (flow, tps, augmentation, reader, loader, data, params, __init__) into
baseline/_ref/, which is git-ignored (never part of the repo's history) but not gpurun-ignored, so it
travels to the GPU box with the snapshot.  Nothing under video-matting_b200/ reads it.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ("__init__.py", "flow.py", "tps.py", "augmentation.py", "reader.py", "loader.py", "data.py", "params.py")


def install(src=None, quiet=False):
    src = src or os.environ.get("VM_REFERENCE_DIR", "/root/reference")
    if not os.path.isdir(src):
        return False
    os.makedirs(DST, exist_ok=True)
    lines = []
    for name in FILES:
        s, d = os.path.join(src, name), os.path.join(DST, name)
        if not os.path.exists(s):
            continue
        shutil.copyfile(s, d)
        os.chmod(d, 0o644)
        with open(d, "rb") as f:
            lines.append(f"{hashlib.sha256(f.read()).hexdigest()}  {name}")
    with open(os.path.join(DST, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if not quiet:
        print(f"baseline/_ref: {len(lines)} reference modules copied from {src}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
