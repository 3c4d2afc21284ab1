"""CPU oracle for the SPVD sparse point-voxel hot path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / torch-CPU) of the algorithm the reference
runs on its hot path.  It exists to CHECK the CUDA path and to be
timed as the CPU baseline; it is never part of the product:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
    ``--impl reference`` legs may import it;
  * nothing under ``spvd_b200/`` imports it, and the product path raises when the CUDA
    library is missing instead of falling back to this code.

PARITY STATUS: **parity unpinned for the external ops.**  The arithmetic of the path lives in
third-party packages that are neither vendored in /root/reference nor installed in the build
container (torchsparse 2.1.0, torch-scatter 2.1.2, torch_geometric 2.5.2; pins from
/root/reference/requirements.txt:1-2 and experiments/ConditionalGeneration.ipynb:448), and the
reference ships no golden vectors or tests for them.  The semantics below
restate their published behaviour and are anchored on the reference's own call sites
(models/sparse_utils.py:36-134, models/spvd.py).  What IS pinned: every piece of the path that
is plain Python/torch in the reference (models/spvd.py, models/ddpm_unet_attn.py,
models/modules/*, models/utils.py, models/sparse_utils.py:138-258, utils/schedulers.py) was
executed unmodified in the build container on top of this oracle's ``torchsparse`` stand-in
(oracle/shim.py) and its outputs are committed under tests/golden/ together with the generating
script oracle/make_golden.py.
"""
from . import ops  # noqa: F401
