"""Name-compatible home of ``BinOp`` (the reference defines it in models/cu_net_prev_version.py:17-92 and
cu-net-prev-version-bin.py:24,65 imports it from there).

This is NOT the prev-version model: the hand-written memory-efficient DenseNet bottleneck of that file (removed
torch._C / torch._thnn APIs, :94-157, :520-760) is out of scope -- the same network is
``models/cu_net.py``.  What matters for BASELINE.json configs[3] is which tensors BinOp binarises on that model: the
dense-layer 3x3 convs and all heads but the last (72 + 7 tensors for CU-Net-8), which ``BinOp(net)`` reproduces on the
drop-in module (``targets="prev_version"``, see utils/quantize.py; pinned by tests/golden/binop_targets.pt)."""
from ..utils.quantize import BinOp, prev_version_conv_names, target_names  # noqa: F401
