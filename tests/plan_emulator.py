"""Executes a cunet_b200 Plan with plain PyTorch on the CPU, kernel contract by kernel contract.

Test infrastructure: it performs exactly the sequence of (stem, fused conv fwd, loss, dgrad, wgrad, stem bwd)
calls the engine issues to the CUDA library, but each call is the torch statement of the kernel's contract
(tests/ops_ref.py).  Comparing its outputs / gradients with the oracle's autograd validates the op plan, the
segment tables, the backward schedule (first-writer / last-writer flags) and the batch-norm backward algebra
without a GPU.
"""
import torch
import torch.nn.functional as F

from tests import ops_ref

EPS = 1e-5


def run(plan, state, img, heatmap, dtype=torch.float64):
    n = img.shape[0]
    st = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state.items()}
    T, stats, pidx, rows_of = {}, {}, {}, {}

    def put(t, rows, with_stats=True):
        T[t.name] = rows
        rows_of[t.name] = rows.shape[0]
        if with_stats:
            stats[t.name] = ops_ref.tensor_stats(rows)

    # ---- stem (cunet_stem_im2col + conv_fwd(identity) + cunet_stem_pool_fwd)
    w0 = st["features.conv0.weight"]
    y = F.conv2d(img.to(dtype), w0, stride=2, padding=3)
    put(plan.stem_y, ops_ref.nchw_to_rows(y))
    sc0, sh0, mu0, var0 = ops_ref.bn_coeffs([stats["stem.y"]], [rows_of["stem.y"]], st["features.norm0.weight"],
                                            st["features.norm0.bias"])
    sc0, sh0 = sc0.to(dtype), sh0.to(dtype)
    z0 = y * sc0.view(1, -1, 1, 1) + sh0.view(1, -1, 1, 1)
    x0 = F.max_pool2d(F.relu(z0), 2, 2)
    put(plan.stem_x, ops_ref.nchw_to_rows(x0))

    # ---- forward ops
    coefs = {}
    for op in plan.ops:
        srcs = [T[t.name] for t, _ in op.srcs]
        ups = [u for _, u in op.srcs]
        sstats = [stats[t.name] for t, _ in op.srcs]
        counts = [rows_of[t.name] for t, _ in op.srcs]
        gamma, beta = st[op.norm + ".weight"], st[op.norm + ".bias"]
        scale, shift, mean, var = ops_ref.bn_coeffs(sstats, counts, gamma, beta)
        coefs[op.name] = (scale.to(dtype), shift.to(dtype), mean.to(dtype), (1.0 / torch.sqrt(var + EPS)).to(dtype))
        w = st[op.conv + ".weight"]
        out, idx = ops_ref.conv_fwd_ref([s.to(dtype) for s in srcs], ups, n, op.res, op.res, coefs[op.name][0],
                                        coefs[op.name][1], w, op.pool, dtype=dtype)
        if op.kind == "head":
            pad = torch.zeros(out.shape[0], op.cout_pad - op.cout, dtype=dtype)
            out = torch.cat([out, pad], 1)
        put(op.out, out, with_stats=op.kind != "head")
        if op.pool:
            pidx[op.out.name] = idx

    # ---- loss (cunet_mse_decode)
    R = plan.out_res
    heads = [ops_ref.rows_to_nchw(T[h.name], n, R, R, dtype=dtype)[:, :plan.class_num] for h in plan.heads]
    hm = heatmap.to(dtype)
    loss = sum(((o - hm) ** 2).sum() / hm.numel() for o in heads)
    dheads = {}
    for h, o in zip(plan.heads, heads):
        d = ops_ref.nchw_to_rows(2.0 * (o - hm) / hm.numel())
        dheads[h.name] = torch.cat([d, torch.zeros(d.shape[0], plan.head_pad - plan.class_num, dtype=dtype)], 1)

    # ---- backward ops (cunet_conv_dgrad + cunet_conv_wgrad per op, reverse order)
    G, gstats, grads = {}, {}, {}

    def grad_of(t, full_res_op):
        """dT rows at the resolution of the op that produced t (cunet_grad_src evaluation)."""
        if t.name in dheads:
            return dheads[t.name]
        c = ops_ref.grad_coeffs(stats[t.name], gstats[t.name], rows_of[t.name])
        c = tuple(v.to(dtype) for v in c)
        return ops_ref.grad_src_eval(G[t.name], T[t.name], c, pidx.get(t.name), n, full_res_op.res, full_res_op.res)

    for op, flags in plan.backward_schedule():
        dy = grad_of(op.out, op)
        srcs = [T[t.name].to(dtype) for t, _ in op.srcs]
        ups = [u for _, u in op.srcs]
        scale, shift, mean, istd = coefs[op.name]
        gamma = st[op.norm + ".weight"]
        outs, dgamma, dbeta, dw = ops_ref.conv_bwd_ref(srcs, ups, n, op.res, op.res, scale, shift, mean, istd,
                                                       gamma, st[op.conv + ".weight"], dy, dtype=dtype)
        grads[op.norm + ".weight"] = dgamma
        grads[op.norm + ".bias"] = dbeta
        grads[op.conv + ".weight"] = dw
        for (t, _), g, (accumulate, last) in zip(op.srcs, outs, flags):
            if accumulate:
                G[t.name] = G[t.name] + g
            else:
                assert t.name not in G, "first writer flag wrong for %s" % t.name
                G[t.name] = g
            if last:
                gstats[t.name] = ops_ref.gstats_of(G[t.name], T[t.name], stats[t.name], rows_of[t.name])

    # ---- stem backward (cunet_stem_bwd phases 0/1 + conv_wgrad(identity))
    p, q, r = (v.to(dtype) for v in ops_ref.grad_coeffs(stats["stem.x"], gstats["stem.x"], rows_of["stem.x"]))
    dx = ops_ref.rows_to_nchw(p * G["stem.x"] + q * T["stem.x"] + r, n, R, R, dtype=dtype)
    a0 = F.relu(z0)
    _, flat = F.max_pool2d(a0, 2, 2, return_indices=True)
    da = torch.zeros_like(a0).flatten(2)
    da.scatter_(2, flat.flatten(2), dx.flatten(2))
    dz = da.view_as(a0) * (z0 > 0).to(dtype)
    istd0 = (1.0 / torch.sqrt(var0 + EPS)).to(dtype)
    yhat = (y - mu0.to(dtype).view(1, -1, 1, 1)) * istd0.view(1, -1, 1, 1)
    dbeta0 = dz.sum((0, 2, 3))
    dgamma0 = (dz * yhat).sum((0, 2, 3))
    cnt = float(rows_of["stem.y"])
    g0 = st["features.norm0.weight"]
    dy0 = (g0 * istd0).view(1, -1, 1, 1) * (dz - dbeta0.view(1, -1, 1, 1) / cnt - yhat * dgamma0.view(1, -1, 1, 1) / cnt)
    wt = w0.detach().clone().requires_grad_(True)
    (dw0,) = torch.autograd.grad(F.conv2d(img.to(dtype), wt, stride=2, padding=3), [wt], dy0)
    grads["features.norm0.weight"] = dgamma0
    grads["features.norm0.bias"] = dbeta0
    grads["features.conv0.weight"] = dw0
    return heads, loss, grads
