"""Drop-in for ``pylib/Evaluation.get_preds`` (pylib/Evaluation.py:6-23): heatmap -> landmark decode on the GPU.

``get_preds(scores[N,C,H,W]) -> float[N,C,2]``: argmax over H*W (first maximum), 1-based (x, y), zero where the
maximum is <= 0.  Runs the fused decode kernel of csrc/loss.cu (the same one the fused training step uses)."""
import ctypes as C

import torch

from .. import lib as L


def get_preds(scores):
    assert scores.dim() == 4, "Score maps should be 4-dim"
    if not scores.is_cuda:
        raise L.CunetError("cunet_b200 get_preds runs on CUDA tensors only (no CPU fallback)")
    n, c, h, w = scores.shape
    rows = scores.permute(0, 2, 3, 1).contiguous().float()            # NHWC, ld = C
    keys = torch.zeros(n * c, dtype=torch.int64, device=scores.device)
    loss = torch.zeros(2, dtype=torch.float64, device=scores.device)
    preds = torch.empty(n, c, 2, device=scores.device)
    mp = L.MseParams()
    mp.heads[0] = rows.data_ptr()
    mp.nheads = 1
    mp.target = scores.contiguous().float().data_ptr()                # unused for the decode (loss is discarded)
    mp.N, mp.C, mp.H, mp.W, mp.ld = n, c, h, w, c
    mp.loss, mp.keys, mp.grad_scale, mp.dtype = loss.data_ptr(), keys.data_ptr(), 1.0, L.F32
    lib = L.load()
    L.check(lib.cunet_mse_decode(C.byref(mp), L.stream_ptr()), "cunet_mse_decode")
    L.check(lib.cunet_decode_finalize(C.c_void_p(keys.data_ptr()), C.c_void_p(preds.data_ptr()), n * c, w,
                                      L.stream_ptr()), "cunet_decode_finalize")
    return preds
