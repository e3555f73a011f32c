"""Name-compatible home of ``BinOp`` (the reference defines it in models/cu_net_prev_version.py:17-92 and
cu-net-prev-version-bin.py:24,65 imports it from there).  The rest of that file (the hand-written
memory-efficient DenseNet bottleneck on removed torch._C / torch._thnn APIs) is out of scope."""
from ..utils.quantize import BinOp  # noqa: F401
