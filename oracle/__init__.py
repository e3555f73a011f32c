"""CPU oracle for the CU-Net hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker / the timed
CPU baseline.  The product path (``cu-net_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Contents
--------
* ``cunet_oracle``     -- restatement of ``models/cu_net.py`` (functional, torch CPU fp32)
* ``quantize_oracle``  -- restatement of ``utils/quantize.py`` (QuanOp, QuanInput) and of
                          ``models/cu_net_prev_version.py:17-92`` (BinOp)
* ``evaluation_oracle``-- restatement of ``pylib/Evaluation.py`` (get_preds, calc_dists, dist_acc, accuracy,
                          final_preds + its own TransformPts, accuracy_origin_res), of the flip helpers of
                          ``pylib/HumanAug.py:177-210`` and of ``pylib/HumanPts.py:35-76`` (pts2heatmap)
* ``synthetic``        -- the seeded synthetic inputs of SURVEY.md §8(d)
* ``ref_loader``       -- executes the *real* reference sources from /root/reference
                          (build container only; never on the GPU box)
* ``gen_golden``       -- script that produced ``tests/golden/*.pt`` from the real reference

Parity pinning: the reference ships no tests / golden vectors, so the
restatements are pinned against outputs of the reference itself, executed in the build
container through ``ref_loader`` and committed as fixtures under ``tests/golden/``.
"""
