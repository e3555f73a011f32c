"""The cu-net.py entry point (reference flags) trains for a few steps on the GPU; the loss goes down."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _entry():
    spec = importlib.util.spec_from_file_location("cu_net_entry", os.path.join(ROOT, "cu-net.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("extra", [[], ["--quant", "bin"], ["--quant", "quan", "--bits_w", "8"],
                                   ["--no-fused", "--dtype", "fp32"]], ids=["fused", "binop", "wig", "module_api"])
def test_entry_trains(extra):
    m = _entry()
    opt = m.parse(["--layer_num", "2", "--order", "1", "--class_num", "16", "--loss_num", "2", "--bs", "2",
                   "--nEpochs", "3", "--iters_per_epoch", "4", "--lr", "1e-3", "--is_train", "true"] + extra)
    hist = m.run(opt)
    assert len(hist) == 3 and all(h == h for h in hist)
    assert hist[-1] < hist[0]


def test_flag_parsing_fixes_bool_quirk():
    m = _entry()
    assert m.parse(["--is_train", "false"]).is_train is False          # reference: type=bool -> True
    assert m.parse([]).loss_num == 2                                     # default 16 clipped to layer_num
