"""Vendor the UNMODIFIED reference modules of the hot path into oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/build_ref.py            (also called by __graft_entry__.build() when /root/reference is present)

The reference is pure Python (no build step): its module files are copied byte for byte, where they lie, into
oracle/_ref/ keeping the directory layout, so `sys.path.insert(0, "oracle/_ref")` imports them exactly as
`sys.path.insert(0, "/root/reference")` does here.  Nothing is copied into tracked paths; oracle/_ref/ is in .gitignore.
Used by: tests (the second, GPU fp32 oracle of SURVEY section 8c), bench.py --impl reference / cpu_baseline / library_bar.
The third-party imports the reference makes at module scope and that are absent from this image (torchio, monai) are
stubbed by `import_ref()` -- only HD95 (metric.py:29-32) would need them.
"""
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")

FILES = [
    "models/three_d/unet3d.py", "models/three_d/vnet3d.py", "models/three_d/residual_unet3d.py",
    "models/three_d/densevoxelnet3d.py", "models/three_d/highresnet.py", "models/three_d/csrnet.py",
    "models/three_d/Double_Unet.py", "models/three_d/SE.py", "models/three_d/ER_net.py", "models/three_d/RE_net.py",
    "models/sync_batchnorm/batchnorm.py", "models/sync_batchnorm/comm.py", "models/sync_batchnorm/replicate.py",
    "models/sync_batchnorm/batchnorm_reimpl.py",
    "utils/convolution.py", "utils/residual.py", "utils/dilation.py", "utils/loss_function.py", "utils/metric.py",
    "conf/config.yaml", "conf/unet.yaml",
]


def build(verbose=True):
    """Copy the files (only when the reference tree is present, i.e. in the build container).  Returns True when
    oracle/_ref is populated afterwards."""
    if os.path.isdir(REF_SRC):
        n = 0
        for rel in FILES:
            src, dst = os.path.join(REF_SRC, rel), os.path.join(REF_DST, rel)
            if not os.path.exists(src):
                continue
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not os.path.exists(dst) or open(src, "rb").read() != open(dst, "rb").read():
                shutil.copyfile(src, dst)
                n += 1
        if verbose and n:
            print("[oracle/_ref] vendored %d reference files" % n)
    return available()


def available():
    return os.path.exists(os.path.join(REF_DST, "models", "three_d", "unet3d.py"))


def import_ref():
    """Make `import models.three_d.unet3d`, `import utils.loss_function` ... resolve to the vendored reference.
    Returns False when oracle/_ref is absent."""
    if not available():
        return False
    if REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    sys.dont_write_bytecode = True
    for name in ("torchio", "monai", "monai.metrics"):
        sys.modules.setdefault(name, types.ModuleType(name))
    for name in ("thop", "torchvision"):       # imported, never used, by Double_Unet.py
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
            sys.modules[name].profile = None
    if not hasattr(sys.modules["monai.metrics"], "compute_hausdorff_distance"):
        sys.modules["monai.metrics"].compute_hausdorff_distance = None
    return True


if __name__ == "__main__":
    print("oracle/_ref available:", build())
