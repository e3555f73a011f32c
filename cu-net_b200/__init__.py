"""cu-net_b200: B200-native (sm_100a) implementation of the CU-Net training/inference hot path.

Layout
  csrc/        hand-written CUDA kernels + the C ABI (include/cunet_b200.h) -> libcunet_b200.so
  lib.py       ctypes binding of the C ABI (fails loudly when the library is missing)
  plan.py      static op plan of the network (virtual-concat segment tables, buffers, schedules)
  engine.py    executes a plan through the C ABI (forward, backward, fused train step)
  models/      drop-in for the reference's models/cu_net.py  (create_cu_net)
  utils/       drop-ins for utils/quantize.py (QuanOp), BinOp and utils/checkpoint.py (Checkpoint)
  pylib/       drop-ins for pylib/Evaluation (get_preds, accuracy, final_preds), HumanAug (flip helpers), HumanPts (heat maps)
  utils/synthetic.py  seeded synthetic batches for cu-net.py / bench.py
"""
__version__ = "0.1.0"
