"""The reference's own PYTHON files of the render path, staged UNMODIFIED under oracle/_ref/ref_py/ (git-ignored;
travels to the GPU box with the snapshot like the compiled reference kernels).  TEST INFRASTRUCTURE ONLY:
used by the drop-in tests (the reference's gswrapper.py / gaussian_splatting.py running on top of this repo's
`gscuda` module), by bench.py's cpu_baseline leg (rendering_python, the path BASELINE config 1 names) and by
tests/report_train_step_c5.py (the Fea2GS_ROPE_AMP head of BASELINE config 5).  Nothing is copied into the repository's
history; /root/reference is only read where it exists (this container), by stage().
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
STAGE = os.path.join(HERE, "_ref", "ref_py")
FILES = [
    "utils/gaussian_splatting.py",          # front end: activations, mapping, dispatch, rendering_python
    "utils/gs_cuda_dmax/gswrapper.py",      # GSCUDA autograd function over `import gscuda`
    "utils/split_and_joint_image.py",       # tiled inference
    "utils/fea2gsropeamp.py",               # Fea2GS_ROPE_AMP head (config 5)
    "utils/fea2gs.py",                      # Fea2GS head
]


def stage(reference: str = "/root/reference") -> bool:
    """Copy the files where the reference tree exists; returns have()."""
    if not os.path.isdir(os.path.join(reference, "utils")):
        return have()
    for rel in FILES:
        dst = os.path.join(STAGE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(reference, rel), dst)
    for d in ("utils", "utils/gs_cuda_dmax"):
        open(os.path.join(STAGE, d, "__init__.py"), "a").close()
    return have()


def have() -> bool:
    return all(os.path.exists(os.path.join(STAGE, rel)) for rel in FILES)


def load(name: str):
    """Import a staged module, e.g. load("utils.gaussian_splatting").  The staged tree shadows any other top-level
    package called `utils` for the duration of the import; the reference's `import gscuda` resolves to the
    repository's top-level gscuda.py."""
    if not have():
        raise FileNotFoundError(f"{STAGE}: reference python files not staged (run __graft_entry__.build() where "
                                "/root/reference exists)")
    root = os.path.dirname(HERE)
    for p in (root, STAGE):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.")]:
        mod = sys.modules[k]
        if not getattr(mod, "__file__", "") or not str(mod.__file__).startswith(STAGE):
            del sys.modules[k]
    return importlib.import_module(name)
