"""GPU parity of the fused cat->BN->ReLU->conv forward kernel (cunet_conv_fwd) through the C ABI."""
import ctypes as C

import pytest
import torch

from tests import ops_ref

pytestmark = pytest.mark.gpu


def _mk(lib, dtype):
    return torch.bfloat16 if dtype == lib.BF16 else torch.float32


def fill_concat(cc, srcs, stats, counts, ups, gamma, beta, rmean, rvar, train, act_bits=0):
    cc.act_bits = act_bits
    cc.nseg = len(srcs)
    for i, (x, st, cnt, up) in enumerate(zip(srcs, stats, counts, ups)):
        cc.seg[i].ptr = x.data_ptr()
        cc.seg[i].stats = st.data_ptr()
        cc.seg[i].inv_count = 1.0 / cnt
        cc.seg[i].C = x.shape[1]
        cc.seg[i].ld = x.shape[1]
        cc.seg[i].up = int(up)
    cc.gamma, cc.beta, cc.rmean, cc.rvar = gamma.data_ptr(), beta.data_ptr(), rmean.data_ptr(), rvar.data_ptr()
    cc.bn_train, cc.eps = int(train), 1e-5


def run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, pool=False, train=True, out_fp32=False,
                 cout_pad=None, seed=0, act_bits=0, identity=False):
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    td = _mk(lib, dtype)
    cout_pad = cout_pad or cout
    srcs, stats, counts = [], [], []
    for c, up in zip(seg_c, ups):
        hh, ww = (h // 2, w // 2) if up else (h, w)
        x = (torch.randn(n * hh * ww, c, generator=g) * 1.3 + 0.4).to(dev).to(td)
        srcs.append(x)
        stats.append(ops_ref.tensor_stats(x))
        counts.append(n * hh * ww)
    cin = sum(seg_c)
    gamma = torch.rand(cin, generator=g).to(dev)
    beta = (torch.randn(cin, generator=g) * 0.2).to(dev)
    rmean = (torch.randn(cin, generator=g) * 0.3).to(dev)
    rvar = (torch.rand(cin, generator=g) + 0.5).to(dev)
    k = 3 if taps == 9 else 1
    weight = ((torch.rand(cout, cin, k, k, generator=g) * 2 - 1) / (cin * k * k) ** 0.5).to(dev)

    # pack weights through the ABI
    nbytes = lib.pack_fwd_bytes(cin, taps, cout_pad, dtype)
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(weight.data_ptr(), wpack.data_ptr(), None, cout, cin, taps, cout_pad)
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)

    rows_out = n * h * w // (4 if pool else 1)
    out_ld = cout_pad if out_fp32 else cout
    out = torch.full((rows_out, out_ld), float("nan"), device=dev, dtype=torch.float32 if out_fp32 else td)
    out_stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev) if not out_fp32 else None
    pidx = torch.zeros(rows_out, cout, dtype=torch.uint8, device=dev) if pool else None

    p = lib.ConvFwdParams()
    fill_concat(p.inp, srcs, stats, counts, ups, gamma, beta, rmean, rvar, train, act_bits)
    if identity:
        p.inp.bn_train = 2
    p.N, p.H, p.W, p.taps = n, h, w, taps
    p.wpack, p.Cout, p.CoutPad = wpack.data_ptr(), cout, cout_pad
    p.out, p.out_ld, p.out_fp32 = out.data_ptr(), out_ld, int(out_fp32)
    p.out_stats = out_stats.data_ptr() if out_stats is not None else None
    p.pool = int(pool)
    p.pool_idx = pidx.data_ptr() if pidx is not None else None
    p.dtype = dtype
    lib.conv_fwd(p)
    torch.cuda.synchronize()

    # reference (fp32 on the same device, same rounded inputs)
    if train:
        scale, shift, _, _ = ops_ref.bn_coeffs(stats, counts, gamma, beta)
    else:
        istd = 1.0 / torch.sqrt(rvar.double() + 1e-5)
        scale = (gamma.double() * istd).float()
        shift = (beta.double() - rmean.double() * gamma.double() * istd).float()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ref, ridx = ops_ref.conv_fwd_ref([s.float() for s in srcs], ups, n, h, w, scale, shift, weight, pool,
                                     act_bits=act_bits, identity=identity)
    return out, ref, out_stats, pidx, ridx


def _relerr(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-12)).item()


CASES = [
    # (name, n,h,w, seg_c, ups, cout, taps, pool, train, out_fp32, cout_pad)
    ("1x1_128", 2, 16, 16, [128], [0], 128, 1, False, True, False, None),
    ("1x1_160_multi", 2, 16, 16, [128, 32], [0, 0], 128, 1, False, True, False, None),
    ("1x1_192_pool", 2, 16, 16, [128, 32, 32], [0, 0, 0], 128, 1, True, True, False, None),
    ("1x1_320_up", 2, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, False, True, False, None),
    ("3x3", 2, 16, 16, [128], [0], 32, 9, False, True, False, None),
    ("3x3_small", 3, 4, 4, [128], [0], 32, 9, False, True, False, None),
    ("head68", 2, 16, 16, [128], [0], 68, 1, False, True, True, 80),
    ("head16_eval", 1, 8, 8, [128], [0], 16, 1, False, False, True, 16),
    ("1x1_tail", 3, 4, 4, [128, 32], [0, 0], 128, 1, False, True, False, None),
    ("1x1_pool_tail", 3, 4, 4, [128, 32, 32], [0, 0, 0], 128, 1, True, True, False, None),
    ("1x1_big", 4, 64, 64, [128, 32, 32], [0, 0, 0], 128, 1, False, True, False, None),
    # dense-layer 3x3, persistent shifted-descriptor kernel in bf16 (conv_fwd3x3.cu): every resolution of the network
    ("3x3_4x4_b24", 24, 4, 4, [128], [0], 32, 9, False, True, False, None),      # the neck: raster of width 8
    ("3x3_4x4_b5_eval", 5, 4, 4, [128], [0], 32, 9, False, False, False, None),
    ("3x3_2x2", 7, 2, 2, [128], [0], 32, 9, False, True, False, None),
    ("3x3_rect_4w", 3, 8, 4, [128], [0], 32, 9, False, True, False, None),
    ("3x3_8x8", 3, 8, 8, [128], [0], 32, 9, False, True, False, None),
    ("3x3_8x8_b24", 24, 8, 8, [128], [0], 32, 9, False, True, False, None),
    ("3x3_32x32", 5, 32, 32, [128], [0], 32, 9, False, True, False, None),
    ("3x3_64x64", 3, 64, 64, [128], [0], 32, 9, False, True, False, None),
    ("3x3_64x64_b24", 24, 64, 64, [128], [0], 32, 9, False, True, False, None),
    ("3x3_rect_eval", 2, 8, 16, [128], [0], 32, 9, False, False, False, None),
    # 1x1, persistent bulk-landing kernel in bf16 (conv_fwd_v2.cu): multi-tile CTAs, upsampled sources, partial tiles
    ("1x1_320_up_b24", 24, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, False, True, False, None),
    ("1x1_288_up_32", 6, 32, 32, [128, 128, 32], [1, 0, 0], 128, 1, False, True, False, None),
    ("1x1_up_tail", 3, 4, 4, [128, 128, 32], [1, 0, 0], 128, 1, False, True, False, None),
    ("1x1_256_b24", 24, 64, 64, [128, 128], [0, 0], 128, 1, False, True, False, None),
    ("1x1_160_eval", 5, 16, 16, [128, 32], [0, 0], 128, 1, False, False, False, None),
    ("head68_b24", 24, 64, 64, [128], [0], 68, 1, False, True, True, 80),
    # third-generation 1x1 forward (conv_fwd_v3.cu): pooled outputs incl. the split stage geometry of 64-wide images,
    # the other bench shapes, the 3 img/GPU strong-scaling shape, a partial last stage
    ("1x1_192_pool_b24", 24, 64, 64, [128, 32, 32], [0, 0, 0], 128, 1, True, True, False, None),
    ("1x1_192_pool_32_b24", 24, 32, 32, [128, 32, 32], [0, 0, 0], 128, 1, True, True, False, None),
    ("1x1_192_pool_8", 5, 8, 8, [128, 32, 32], [0, 0, 0], 128, 1, True, True, False, None),
    ("1x1_192_b24", 24, 64, 64, [128, 32, 32], [0, 0, 0], 128, 1, False, True, False, None),
    ("1x1_288_up_b24", 24, 64, 64, [128, 128, 32], [1, 0, 0], 128, 1, False, True, False, None),
    ("1x1_320_up_32_b24", 24, 32, 32, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, False, True, False, None),
    ("1x1_320_up_16_b24", 24, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, False, True, False, None),
    ("1x1_320_up_b3", 3, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, False, True, False, None),
    ("head16_b16", 16, 64, 64, [128], [0], 16, 1, False, True, True, 16),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_conv_fwd(case, dtype_name):
    from cunet_b200 import lib
    lib.load()
    dtype = lib.F32 if dtype_name == "f32" else lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, pool, train, out_fp32, cout_pad = case
    out, ref, out_stats, pidx, ridx = run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, pool, train,
                                                    out_fp32, cout_pad)
    got = out.float()[:, :cout]
    assert torch.isfinite(got).all(), "kernel left unwritten / non-finite outputs"
    err = _relerr(got, ref)
    tol = 2e-3 if dtype == lib.F32 else 1.5e-2      # tf32 operands / bf16 operands+storage
    assert err < tol, "%s %s rel err %g" % (name, dtype_name, err)
    if out_fp32 and cout_pad and cout_pad > cout:
        assert (out[:, cout:] == 0).all()
    if out_stats is not None:
        st_ref = ops_ref.tensor_stats(got)
        assert _relerr(out_stats, st_ref) < 1e-4
    if pool:
        # argmax positions must agree wherever the maximum is not a (near) tie
        agree = (pidx == ridx).float().mean().item()
        assert agree > (0.999 if dtype == lib.F32 else 0.97), agree


V2_CASES = [c for c in CASES if c[7] == 1 and not c[8]]     # every 1x1 case without pooling


@pytest.mark.parametrize("case", V2_CASES, ids=[c[0] for c in V2_CASES])
def test_conv_fwd_v2(case):
    """The opt-in persistent 1x1 forward kernel (conv_fwd_v2.cu) against the same references, bf16."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, pool, train, out_fp32, cout_pad = case
    old = lib.debug_fwd_v2_min_tiles(1)
    try:
        out, ref, out_stats, _, _ = run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, pool, train, out_fp32,
                                                 cout_pad)
    finally:
        lib.debug_fwd_v2_min_tiles(old)
    got = out.float()[:, :cout]
    assert torch.isfinite(got).all(), "kernel left unwritten / non-finite outputs"
    err = _relerr(got, ref)
    assert err < 1.5e-2, "%s rel err %g" % (name, err)
    if out_fp32 and cout_pad and cout_pad > cout:
        assert (out[:, cout:] == 0).all()
    if out_stats is not None:
        assert _relerr(out_stats, ops_ref.tensor_stats(got)) < 1e-4


V3_CASES = [c for c in CASES if c[7] == 1]     # every 1x1 case (ineligible ones fall through to the other kernels)


@pytest.mark.parametrize("case", V3_CASES, ids=[c[0] for c in V3_CASES])
def test_conv_fwd_v3(case):
    """The third-generation persistent 1x1 forward kernel (conv_fwd_v3.cu) forced on at every size, bf16: outputs,
    the per-channel statistics of the stored values, pooled outputs and argmax bytes."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, pool, train, out_fp32, cout_pad = case
    old = lib.debug_fwd_v3_min_rows(0)
    try:
        out, ref, out_stats, pidx, ridx = run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, pool, train,
                                                       out_fp32, cout_pad)
    finally:
        lib.debug_fwd_v3_min_rows(old)
    got = out.float()[:, :cout]
    assert torch.isfinite(got).all(), "kernel left unwritten / non-finite outputs"
    err = _relerr(got, ref)
    assert err < 1.5e-2, "%s rel err %g" % (name, err)
    if out_fp32 and cout_pad and cout_pad > cout:
        assert (out[:, cout:] == 0).all()
    if out_stats is not None:
        assert _relerr(out_stats, ops_ref.tensor_stats(got)) < 1e-4
    if pool:
        agree = (pidx == ridx).float().mean().item()
        assert agree > 0.97, agree


QUANT_CASES = [
    # the convs that sit behind a QuanInput2d in the wig model (models/cu_net_prev_version_wig.py:96-98, 277-279)
    ("3x3_q8", 2, 16, 16, [128], [0], 32, 9, 8, False, None),
    ("3x3_q8_b24", 24, 64, 64, [128], [0], 32, 9, 8, False, None),
    ("3x3_q4", 3, 8, 8, [128], [0], 32, 9, 4, False, None),
    ("3x3_q6_4x4", 24, 4, 4, [128], [0], 32, 9, 6, False, None),
    ("head68_q8", 2, 16, 16, [128], [0], 68, 1, 8, True, 80),
    ("head68_q8_b24", 24, 64, 64, [128], [0], 68, 1, 8, True, 80),
    ("head16_q8_b24", 24, 64, 64, [128], [0], 16, 1, 8, True, 16),
]


@pytest.mark.parametrize("case", QUANT_CASES, ids=[c[0] for c in QUANT_CASES])
@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_conv_fwd_quantized_activations(case, dtype_name):
    """cunet_concat.act_bits: QuanInput2d fused into the operand transform (csrc/loaders.cuh::ActQuant) for the convs the
    wig model quantizes (3x3 dense-layer convs, heads), against conv(Q(C(relu(bn(x)))))."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.F32 if dtype_name == "f32" else lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, bits, out_fp32, cout_pad = case
    out, ref, out_stats, _, _ = run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, False, True, out_fp32,
                                             cout_pad, act_bits=bits)
    plain, _, _, _, _ = run_conv_fwd(lib, dtype, n, h, w, seg_c, ups, cout, taps, False, True, out_fp32, cout_pad)
    got = out.float()[:, :cout]
    assert torch.isfinite(got).all()
    err = _relerr(got, ref)
    # a value that sits on a rounding tie of the 2^-(bits-1) grid may land on the other side (bf16 / tf32 operands)
    tol = (4e-3 if dtype == lib.F32 else 2.5e-2) * (2 if bits < 8 else 1)
    assert err < tol, "%s %s rel err %g" % (name, dtype_name, err)
    assert _relerr(plain.float()[:, :cout], ref) > 3 * err, "quantization had no effect?"


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(8, 64, 64), (3, 16, 16), (96, 64, 64)])
def test_conv_fwd_identity_input(n, h, w):
    """conv0 over the im2col blocks (cunet_stem_im2col: two dense column blocks of 128 and 32 channels, bn_train == 2: no
    BatchNorm, no ReLU) through the persistent 1x1 kernel; (96, 64, 64) is the row count of the stem at batch 24."""
    from cunet_b200 import lib
    lib.load()
    out, ref, out_stats, _, _ = run_conv_fwd(lib, lib.BF16, n, h, w, [128, 32], [0, 0], 128, 1, identity=True)
    assert torch.isfinite(out.float()).all()
    assert _relerr(out.float(), ref) < 2e-2
    # statistics of the stored (rounded) outputs, as every conv reports them
    assert _relerr(out_stats, ops_ref.tensor_stats(out)) < 1e-4
