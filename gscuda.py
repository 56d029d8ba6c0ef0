"""Top-level ``gscuda`` module: what the reference's wrappers import
(utils/gs_cuda_dmax/gswrapper.py:19-20 ``import gscuda``).  With the repository root (or an
installed copy of this file next to ``gsasr_b200``) on sys.path, GSASR's gswrapper.py /
gaussian_splatting.py run unmodified on the B200-native kernels."""
from gsasr_b200.gscuda import gs_render, gs_render_backward, get_ksigma, set_ksigma  # noqa: F401
