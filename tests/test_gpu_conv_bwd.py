"""GPU parity of the backward kernels (cunet_conv_dgrad, cunet_conv_wgrad) through the C ABI."""
import pytest
import torch

from tests import ops_ref
from tests.test_gpu_conv_fwd import fill_concat, _relerr

pytestmark = pytest.mark.gpu


def make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad=None, seed=0):
    """dy_mode: 'plain' | 'bn' | 'pool' (bn form routed through a 2x2 max-pool argmax)."""
    dev = torch.device("cuda")
    g = torch.Generator(device="cpu").manual_seed(seed)
    td = torch.bfloat16 if dtype == lib.BF16 else torch.float32
    cout_pad = cout_pad or cout
    srcs, stats, counts = [], [], []
    for c, up in zip(seg_c, ups):
        hh, ww = (h // 2, w // 2) if up else (h, w)
        x = (torch.randn(n * hh * ww, c, generator=g) * 1.3 + 0.4).to(dev).to(td)
        srcs.append(x)
        stats.append(ops_ref.tensor_stats(x))
        counts.append(n * hh * ww)
    cin = sum(seg_c)
    gamma = (torch.rand(cin, generator=g) + 0.1).to(dev)
    beta = (torch.randn(cin, generator=g) * 0.2).to(dev)
    k = 3 if taps == 9 else 1
    weight = ((torch.rand(cout, cin, k, k, generator=g) * 2 - 1) / (cin * k * k) ** 0.5).to(dev)
    # gradient source of the conv output
    ho, wo = (h // 2, w // 2) if dy_mode == "pool" else (h, w)
    rows = n * ho * wo
    gten = (torch.randn(rows, cout_pad, generator=g) * 0.1).to(dev).to(td)
    if cout_pad > cout:
        gten[:, cout:] = 0
    tten = (torch.randn(rows, cout_pad, generator=g) * 1.1 + 0.3).to(dev).to(td)
    tstats = ops_ref.tensor_stats(tten)
    gstats = ops_ref.gstats_of(gten, tten, tstats, rows)
    pidx = None
    if dy_mode == "pool":
        pidx = torch.randint(0, 4, (rows, cout_pad), generator=g).to(torch.uint8).to(dev)
    return dict(srcs=srcs, stats=stats, counts=counts, gamma=gamma, beta=beta, weight=weight, g=gten, t=tten,
                tstats=tstats, gstats=gstats, pidx=pidx, rows=rows, cin=cin, cout_pad=cout_pad, td=td, dev=dev)


def fill_grad_src(gs, cs, dy_mode):
    gs.g, gs.t = cs["g"].data_ptr(), cs["t"].data_ptr()
    gs.stats, gs.gstats = cs["tstats"].data_ptr(), cs["gstats"].data_ptr()
    gs.pool_idx = cs["pidx"].data_ptr() if cs["pidx"] is not None else None
    gs.inv_count = 1.0 / cs["rows"]
    gs.C, gs.ld = cs["cout_pad"], cs["cout_pad"]
    gs.mode = 0 if dy_mode == "plain" else 1
    gs.pooled = int(dy_mode == "pool")
    gs.eps = 1e-5


def reference(cs, n, h, w, ups, dy_mode, act_bits=0):
    scale, shift, mean, var = ops_ref.bn_coeffs(cs["stats"], cs["counts"], cs["gamma"], cs["beta"])
    istd = 1.0 / torch.sqrt(var + 1e-5)
    coeffs = None if dy_mode == "plain" else ops_ref.grad_coeffs(cs["tstats"], cs["gstats"], cs["rows"])
    dy = ops_ref.grad_src_eval(cs["g"], cs["t"], coeffs, cs["pidx"], n, h, w)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return ops_ref.conv_bwd_ref([s.float() for s in cs["srcs"]], ups, n, h, w, scale, shift, mean, istd,
                                cs["gamma"], cs["weight"], dy, act_bits=act_bits)


CASES = [
    # name, n,h,w, seg_c, ups, cout, taps, dy_mode, cout_pad
    ("1x1_128_plain", 2, 16, 16, [128], [0], 128, 1, "plain", None),
    ("1x1_192_bn", 2, 16, 16, [128, 32, 32], [0, 0, 0], 128, 1, "bn", None),
    ("1x1_192_pool", 2, 16, 16, [128, 32, 32], [0, 0, 0], 128, 1, "pool", None),
    ("1x1_320_up_bn", 2, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, "bn", None),
    ("3x3_bn", 2, 16, 16, [128], [0], 32, 9, "bn", None),
    ("head68_plain", 2, 16, 16, [128], [0], 68, 1, "plain", 80),
    ("1x1_tail_bn", 3, 4, 4, [128, 32], [0, 0], 128, 1, "bn", None),
    ("1x1_up_tail", 3, 4, 4, [128, 128, 32], [1, 0, 0], 128, 1, "bn", None),
    ("3x3_tail", 3, 4, 4, [128], [0], 32, 9, "bn", None),
    ("1x1_big_bn", 4, 64, 64, [128, 32, 32], [0, 0, 0], 128, 1, "bn", None),
    # ---- the shapes bench.py runs (BASELINE.json configs[2]: batch 24): 768 tiles of 128 pixels on 148 persistent
    # CTAs, i.e. ~5 tiles per CTA -- landing-ring wrap across tiles, in-place G over the landed x, double-buffered
    # TMEM across work items.  Segment i of every case accumulates when i is odd, so both `accumulate` values are
    # exercised at every shape.
    ("1x1_320_up_b24", 24, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, "bn", None),   # up_blocks.0 adapter
    ("1x1_288_up_b24", 24, 64, 64, [128, 128, 32], [1, 0, 0], 128, 1, "bn", None),          # up_blocks.0 conv1
    ("1x1_192_pool_b24", 24, 64, 64, [128, 32, 32], [0, 0, 0], 128, 1, "pool", None),       # down_blocks.0 ahead
    ("1x1_256_b24", 24, 64, 64, [128, 128], [0, 0], 128, 1, "bn", None),                    # intermedia adapter
    ("1x1_160_b24", 24, 64, 64, [128, 32], [0, 0], 128, 1, "bn", None),                     # down_blocks.0 conv1
    ("head68_b24", 24, 64, 64, [128], [0], 68, 1, "plain", 80),                             # heat-map head
    ("1x1_320_up_32_b24", 24, 32, 32, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, "bn", None),  # 192 tiles x 3 chunks
    ("1x1_192_pool_32_b24", 24, 32, 32, [128, 32, 32], [0, 0, 0], 128, 1, "pool", None),
    ("1x1_320_up_16_b24", 24, 16, 16, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, "bn", None),  # 48 x 3: split mode
    ("1x1_192_pool_8_b24", 24, 8, 8, [128, 32, 32], [0, 0, 0], 128, 1, "pool", None),
    ("1x1_320_up_b3", 3, 64, 64, [128, 128, 32, 32], [1, 0, 0, 0], 128, 1, "bn", None),     # 3 img/GPU (strong scaling)
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_conv_wgrad(case, dtype_name):
    from cunet_b200 import lib
    lib.load()
    dtype = lib.F32 if dtype_name == "f32" else lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad = case
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad)
    dw = torch.zeros(cout, cs["cin"], taps, device=cs["dev"])
    p = lib.ConvWgradParams()
    fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(p.dy, cs, dy_mode)
    p.N, p.H, p.W, p.taps, p.Cout = n, h, w, taps, cout
    p.dw, p.nsplit, p.dtype = dw.data_ptr(), 0, dtype
    lib.conv_wgrad(p)
    torch.cuda.synchronize()
    _, _, _, dw_ref = reference(cs, n, h, w, ups, dy_mode)
    err = _relerr(dw.reshape(dw_ref.shape[0], dw_ref.shape[1], -1), dw_ref.reshape(dw_ref.shape[0], dw_ref.shape[1], -1))
    tol = 3e-3 if dtype == lib.F32 else 2e-2
    assert err < tol, "%s %s dW rel err %g" % (name, dtype_name, err)


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_conv_dgrad(case, dtype_name):
    from cunet_b200 import lib
    lib.load()
    dtype = lib.F32 if dtype_name == "f32" else lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad = case
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad)
    dev, td = cs["dev"], cs["td"]
    cin = cs["cin"]
    # pack dgrad image
    nbytes = lib.pack_dgrad_bytes(cin, taps, cs["cout_pad"], dtype)
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, cin, taps, cs["cout_pad"])
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)

    p = lib.ConvDgradParams()
    fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(p.dy, cs, dy_mode)
    gen = torch.Generator(device="cpu").manual_seed(99)
    Gs, G0, gst = [], [], []
    for i, x in enumerate(cs["srcs"]):
        accumulate = i % 2          # alternate write / read-modify-write
        g0 = (torch.randn(x.shape, generator=gen) * 0.05).to(dev).to(td) if accumulate else \
            torch.full(x.shape, float("nan"), device=dev, dtype=td)
        G0.append(g0.clone())
        Gs.append(g0)
        st = torch.zeros(2 * x.shape[1], dtype=torch.float64, device=dev)
        gst.append(st)
        p.gacc[i].G = g0.data_ptr()
        p.gacc[i].gstats = st.data_ptr()
        p.gacc[i].ld = x.shape[1]
        p.gacc[i].accumulate = accumulate
    dgamma = torch.zeros(cin, device=dev)
    dbeta = torch.zeros(cin, device=dev)
    p.N, p.H, p.W, p.taps = n, h, w, taps
    p.wpack_dgrad, p.Cout, p.CoutPad = wpack.data_ptr(), cout, cs["cout_pad"]
    p.dgamma, p.dbeta, p.dtype = dgamma.data_ptr(), dbeta.data_ptr(), dtype
    lib.conv_dgrad(p)
    torch.cuda.synchronize()

    outs, dg_ref, db_ref, _ = reference(cs, n, h, w, ups, dy_mode)
    tol = 3e-3 if dtype == lib.F32 else 2.5e-2
    for i, (G, ref) in enumerate(zip(Gs, outs)):
        exp = ref + (G0[i].float() if i % 2 else 0)
        assert torch.isfinite(G.float()).all(), "segment %d has unwritten rows" % i
        err = _relerr(G.float(), exp)
        assert err < tol, "%s %s G[%d] rel err %g" % (name, dtype_name, i, err)
        # gstats receives THIS consumer's share of (sum G, sum G*xhat), i.e. the sums of its own gamma*dz
        st_ref = ops_ref.gstats_of(ref, cs["srcs"][i], cs["stats"][i], cs["counts"][i])
        c = G.shape[1]
        assert _relerr(gst[i][:c], st_ref[:c]) < tol and _relerr(gst[i][c:], st_ref[c:]) < tol, "gstats %d" % i
    assert _relerr(dbeta, db_ref) < tol, "dbeta %g" % _relerr(dbeta, db_ref)
    assert _relerr(dgamma, dg_ref) < tol, "dgamma %g" % _relerr(dgamma, dg_ref)


FUSED_CASES = [
    # name, n, h, w, accumulate
    ("16x16", 2, 16, 16, 0),
    ("4x4_tail", 3, 4, 4, 0),          # 48 pixels: one partial stage
    ("4x4_b24", 24, 4, 4, 1),          # the network's neck: 6 stages, several images per stage
    ("8x8", 3, 8, 8, 1),
    ("32x32", 5, 32, 32, 0),           # 80 stages
    ("64x64", 4, 64, 64, 1),           # 256 stages, > 148 CTAs' worth: multi-stage CTAs, halo = 65 rows
    ("64x64_b24", 24, 64, 64, 0),      # the bench shape: 1536 stages, 11 per CTA
]


@pytest.mark.parametrize("case", FUSED_CASES, ids=[c[0] for c in FUSED_CASES])
def test_conv_bwd3x3_fused(case):
    """cunet_conv_bwd3x3 (one fused launch) == dgrad + wgrad of the 3x3 dense-layer conv, bf16."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.BF16
    name, n, h, w, accumulate = case
    seg_c, ups, cout, taps, dy_mode = [128], [0], 32, 9, "bn"
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, None, seed=3)
    dev, td = cs["dev"], cs["td"]
    nbytes = lib.pack_dgrad_bytes(128, taps, 32, dtype)
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, 128, taps, 32)
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)

    dp = lib.ConvDgradParams()
    fill_concat(dp.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(dp.dy, cs, dy_mode)
    x = cs["srcs"][0]
    gen = torch.Generator(device="cpu").manual_seed(99)
    g0 = (torch.randn(x.shape, generator=gen) * 0.05).to(dev).to(td) if accumulate else \
        torch.full(x.shape, float("nan"), device=dev, dtype=td)
    G0 = g0.clone()
    gst = torch.zeros(256, dtype=torch.float64, device=dev)
    dp.gacc[0].G, dp.gacc[0].gstats, dp.gacc[0].ld, dp.gacc[0].accumulate = g0.data_ptr(), gst.data_ptr(), 128, accumulate
    dgamma = torch.zeros(128, device=dev)
    dbeta = torch.zeros(128, device=dev)
    dp.N, dp.H, dp.W, dp.taps = n, h, w, taps
    dp.wpack_dgrad, dp.Cout, dp.CoutPad = wpack.data_ptr(), cout, 32
    dp.dgamma, dp.dbeta, dp.dtype = dgamma.data_ptr(), dbeta.data_ptr(), dtype

    dw = torch.zeros(cout, 128, taps, device=dev)
    wp = lib.ConvWgradParams()
    fill_concat(wp.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(wp.dy, cs, dy_mode)
    wp.N, wp.H, wp.W, wp.taps, wp.Cout = n, h, w, taps, cout
    wp.dw, wp.nsplit, wp.dtype = dw.data_ptr(), 0, dtype

    lib.conv_bwd3x3(dp, wp)
    torch.cuda.synchronize()

    outs, dg_ref, db_ref, dw_ref = reference(cs, n, h, w, ups, dy_mode)
    tol = 2.5e-2
    exp = outs[0] + (G0.float() if accumulate else 0)
    assert torch.isfinite(g0.float()).all(), "unwritten rows in G"
    err = _relerr(g0.float(), exp)
    assert err < tol, "%s G rel err %g" % (name, err)
    st_ref = ops_ref.gstats_of(outs[0], x, cs["stats"][0], cs["counts"][0])
    assert _relerr(gst[:128], st_ref[:128]) < tol and _relerr(gst[128:], st_ref[128:]) < tol, "gstats"
    assert _relerr(dbeta, db_ref) < tol, "dbeta %g" % _relerr(dbeta, db_ref)
    assert _relerr(dgamma, dg_ref) < tol, "dgamma %g" % _relerr(dgamma, dg_ref)
    errw = _relerr(dw.reshape(cout, 128, -1), dw_ref.reshape(cout, 128, -1))
    assert errw < 2e-2, "%s dW rel err %g" % (name, errw)


F1_CASES = [c for c in CASES if c[7] == 1]      # every 1x1 case (the fused kernel falls back when not eligible)


@pytest.mark.parametrize("case", F1_CASES, ids=[c[0] for c in F1_CASES])
def test_conv_bwd1x1_fused(case):
    """cunet_conv_bwd1x1 (one fused launch, csrc/conv_bwd1x1.cu) == dgrad + wgrad of the 1x1 fused conv, bf16: source
    gradients (write and accumulate segments), their statistics shares, dgamma / dbeta and dW, at every shape of the
    network incl. the bench shapes (upsampled source at 64x64 = split stage geometry, pooled output, heads)."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.BF16
    name, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad = case
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad, seed=7)
    dev, td, cin = cs["dev"], cs["td"], cs["cin"]
    nbytes = lib.pack_dgrad_bytes(cin, taps, cs["cout_pad"], dtype)
    wpack = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, cin, taps, cs["cout_pad"])
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)

    dp = lib.ConvDgradParams()
    fill_concat(dp.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(dp.dy, cs, dy_mode)
    gen = torch.Generator(device="cpu").manual_seed(99)
    Gs, G0, gst = [], [], []
    for i, x in enumerate(cs["srcs"]):
        accumulate = i % 2
        g0 = (torch.randn(x.shape, generator=gen) * 0.05).to(dev).to(td) if accumulate else \
            torch.full(x.shape, float("nan"), device=dev, dtype=td)
        G0.append(g0.clone())
        Gs.append(g0)
        st = torch.zeros(2 * x.shape[1], dtype=torch.float64, device=dev)
        gst.append(st)
        dp.gacc[i].G, dp.gacc[i].gstats, dp.gacc[i].ld, dp.gacc[i].accumulate = g0.data_ptr(), st.data_ptr(), x.shape[1], accumulate
    dgamma = torch.zeros(cin, device=dev)
    dbeta = torch.zeros(cin, device=dev)
    dp.N, dp.H, dp.W, dp.taps = n, h, w, taps
    dp.wpack_dgrad, dp.Cout, dp.CoutPad = wpack.data_ptr(), cout, cs["cout_pad"]
    dp.dgamma, dp.dbeta, dp.dtype = dgamma.data_ptr(), dbeta.data_ptr(), dtype

    dw = torch.zeros(cout, cin, taps, device=dev)
    wp = lib.ConvWgradParams()
    fill_concat(wp.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    fill_grad_src(wp.dy, cs, dy_mode)
    wp.N, wp.H, wp.W, wp.taps, wp.Cout = n, h, w, taps, cout
    wp.dw, wp.nsplit, wp.dtype = dw.data_ptr(), 0, dtype

    xs_before = [x.clone() for x in cs["srcs"]]
    lib.conv_bwd1x1(dp, wp)
    torch.cuda.synchronize()
    for x, x0 in zip(cs["srcs"], xs_before):
        assert torch.equal(x, x0), "the kernel must not modify its inputs"

    outs, dg_ref, db_ref, dw_ref = reference(cs, n, h, w, ups, dy_mode)
    tol = 2.5e-2
    for i, (G, ref) in enumerate(zip(Gs, outs)):
        exp = ref + (G0[i].float() if i % 2 else 0)
        assert torch.isfinite(G.float()).all(), "segment %d has unwritten rows" % i
        err = _relerr(G.float(), exp)
        assert err < tol, "%s G[%d] rel err %g" % (name, i, err)
        st_ref = ops_ref.gstats_of(ref, cs["srcs"][i], cs["stats"][i], cs["counts"][i])
        c = G.shape[1]
        assert _relerr(gst[i][:c], st_ref[:c]) < tol and _relerr(gst[i][c:], st_ref[c:]) < tol, "gstats %d" % i
    assert _relerr(dbeta, db_ref) < tol, "dbeta %g" % _relerr(dbeta, db_ref)
    assert _relerr(dgamma, dg_ref) < tol, "dgamma %g" % _relerr(dgamma, dg_ref)
    errw = _relerr(dw.reshape(cout, cin, -1), dw_ref.reshape(cout, cin, -1))
    assert errw < 2e-2, "%s dW rel err %g" % (name, errw)


QBWD_CASES = [
    # name, n, h, w, cout, taps, dy_mode, cout_pad, bits
    ("3x3_q8", 2, 16, 16, 32, 9, "bn", None, 8),
    ("3x3_q8_b24", 24, 64, 64, 32, 9, "bn", None, 8),
    ("head68_q8", 2, 16, 16, 68, 1, "plain", 80, 8),
    ("head68_q8_b24", 24, 64, 64, 68, 1, "plain", 80, 8),
    ("head16_q4_b8", 8, 64, 64, 16, 1, "plain", 16, 4),
]


@pytest.mark.parametrize("case", QBWD_CASES, ids=[c[0] for c in QBWD_CASES])
@pytest.mark.parametrize("dtype_name", ["f32", "bf16"])
def test_conv_bwd_quantized_activations(case, dtype_name):
    """Backward of a conv behind a QuanInput2d (cunet_concat.act_bits): the filter gradient contracts the QUANTIZED
    activation, the data gradient is straight-through and zero where the activation is >= 1 (utils/quantize.py:58-63);
    through cunet_conv_bwd3x3 / cunet_conv_bwd1x1 (fused kernels in bf16, the generic pair in fp32)."""
    from cunet_b200 import lib
    lib.load()
    dtype = lib.F32 if dtype_name == "f32" else lib.BF16
    name, n, h, w, cout, taps, dy_mode, cout_pad, bits = case
    seg_c, ups = [128], [0]
    cs = make_case(lib, dtype, n, h, w, seg_c, ups, cout, taps, dy_mode, cout_pad, seed=11)
    dev, td = cs["dev"], cs["td"]
    wpack = torch.empty(lib.pack_dgrad_bytes(128, taps, cs["cout_pad"], dtype), dtype=torch.uint8, device=dev)
    desc = lib.PackDesc(cs["weight"].data_ptr(), None, wpack.data_ptr(), cout, 128, taps, cs["cout_pad"])
    desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.uint8).to(dev)
    lib.pack_weights(desc_dev.data_ptr(), 1, dtype)
    dp, wp = lib.ConvDgradParams(), lib.ConvWgradParams()
    for q in (dp, wp):
        fill_concat(q.inp, cs["srcs"], cs["stats"], cs["counts"], ups, cs["gamma"], cs["beta"], cs["gamma"], cs["gamma"],
                    True, bits)
        fill_grad_src(q.dy, cs, dy_mode)
    x = cs["srcs"][0]
    g0 = torch.full(x.shape, float("nan"), device=dev, dtype=td)
    gst = torch.zeros(256, dtype=torch.float64, device=dev)
    dp.gacc[0].G, dp.gacc[0].gstats, dp.gacc[0].ld, dp.gacc[0].accumulate = g0.data_ptr(), gst.data_ptr(), 128, 0
    dgamma, dbeta = torch.zeros(128, device=dev), torch.zeros(128, device=dev)
    dp.N, dp.H, dp.W, dp.taps = n, h, w, taps
    dp.wpack_dgrad, dp.Cout, dp.CoutPad = wpack.data_ptr(), cout, cs["cout_pad"]
    dp.dgamma, dp.dbeta, dp.dtype = dgamma.data_ptr(), dbeta.data_ptr(), dtype
    dw = torch.zeros(cout, 128, taps, device=dev)
    wp.N, wp.H, wp.W, wp.taps, wp.Cout = n, h, w, taps, cout
    wp.dw, wp.nsplit, wp.dtype = dw.data_ptr(), 0, dtype
    (lib.conv_bwd3x3 if taps == 9 else lib.conv_bwd1x1)(dp, wp)
    torch.cuda.synchronize()
    outs, dg_ref, db_ref, dw_ref = reference(cs, n, h, w, ups, dy_mode, act_bits=bits)
    outs0, _, _, dw0 = reference(cs, n, h, w, ups, dy_mode)
    tol = (4e-3 if dtype == lib.F32 else 2.5e-2) * (2 if bits < 8 else 1)
    assert torch.isfinite(g0.float()).all()
    eg = _relerr(g0.float(), outs[0])
    assert eg < tol, "%s %s G rel err %g" % (name, dtype_name, eg)
    assert _relerr(dbeta, db_ref) < tol and _relerr(dgamma, dg_ref) < tol
    ew = _relerr(dw.reshape(cout, 128, -1), dw_ref.reshape(cout, 128, -1))
    assert ew < tol, "%s %s dW rel err %g" % (name, dtype_name, ew)
    # and the quantizer matters: the unquantized reference is measurably different
    assert _relerr(outs0[0], outs[0]) > 3 * eg and _relerr(dw0.reshape(cout, 128, -1), dw_ref.reshape(cout, 128, -1)) > 3 * ew


@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(3, 16, 16), (8, 64, 64), (96, 64, 64)])
def test_conv_wgrad_identity_input(n, h, w):
    """conv0's weight gradient over the im2col blocks (two dense column blocks of 128 and 32 channels, bn_train == 2: no
    BatchNorm / ReLU in front), plain output gradient, dW rows of 147 columns (3*7*7, dw_cin) -- the stem's backward-
    filter call; (96, 64, 64) is its row count at batch 24."""
    from cunet_b200 import lib
    lib.load()
    cs = make_case(lib, lib.BF16, n, h, w, [128, 32], [0, 0], 128, 1, "plain", None)
    dw = torch.zeros(128, 147, device=cs["dev"])
    p = lib.ConvWgradParams()
    fill_concat(p.inp, cs["srcs"], cs["stats"], cs["counts"], [0, 0], cs["gamma"], cs["beta"], cs["gamma"],
                cs["gamma"], True)
    p.inp.bn_train = 2
    fill_grad_src(p.dy, cs, "plain")
    p.N, p.H, p.W, p.taps, p.Cout = n, h, w, 1, 128
    p.dw, p.nsplit, p.dtype, p.dw_cin = dw.data_ptr(), 0, lib.BF16, 147
    lib.conv_wgrad(p)
    torch.cuda.synchronize()
    x = torch.cat([t.float() for t in cs["srcs"]], 1)            # [rows][160]
    ref = (cs["g"].float().t() @ x)[:, :147]                     # dW[co][k] = sum_px dY[px][co] * x[px][k]
    assert _relerr(dw, ref) < 2e-2
