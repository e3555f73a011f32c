"""Execute the real reference sources (build container only).

TEST INFRASTRUCTURE -- see oracle/__init__.py.  /root/reference does not exist on the GPU
box; everything here raises ``ReferenceUnavailable`` there and callers must skip.

Recipe: ``models/cu_net.py`` is python-2 source; two tokens stop it
from running under python 3:
  * ``models/cu_net.py:286``  ``print '...'`` statement  -> function-call form
  * ``models/cu_net.py:94``   ``adapter_out_num / 2``    -> floor division
Both are substituted in memory; nothing is copied into this repository.

``pylib/HumanAug.py``, ``pylib/Evaluation.py`` and ``pylib/HumanPts.py`` (validation path and target heat maps) are
python-2 sources as well: one ``print '...'`` statement each in a ``__main__`` block (HumanAug.py:277,
HumanPts.py:336), tab/space mixed indentation in HumanPts.py:43-44 (python 2 counts a tab as eight columns) and the
implicit relative ``import HumanAug`` of Evaluation.py:4.  ``load_reference_pylib`` executes them with those three
things patched in memory.  None of the executed functions relies on python-2 integer division.
"""
import re
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("CUNET_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "models", "cu_net.py"))


def load_reference_cu_net():
    """Return a module object holding the reference's models/cu_net.py namespace."""
    if not available():
        raise ReferenceUnavailable("reference tree not present at %s" % REF_ROOT)
    path = os.path.join(REF_ROOT, "models", "cu_net.py")
    with open(path, "r") as f:
        src = f.read()
    a = "print 'order is larger than the layer number.'"
    b = "adapter_out_num = adapter_out_num / 2"
    assert a in src and b in src, "reference source changed; update the substitutions"
    src = src.replace(a, "print('order is larger than the layer number.')")
    src = src.replace(b, "adapter_out_num = adapter_out_num // 2")
    mod = types.ModuleType("reference_cu_net")
    mod.__file__ = path
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def create_reference_net(class_num, layer_num, order, loss_num,
                         neck_size=4, growth_rate=32, init_chan_num=128):
    """create_cu_net of the reference (models/cu_net.py:362-368), stdout suppressed."""
    mod = load_reference_cu_net()
    with contextlib.redirect_stdout(io.StringIO()):
        net = mod.create_cu_net(neck_size=neck_size, growth_rate=growth_rate,
                                init_chan_num=init_chan_num, class_num=class_num,
                                layer_num=layer_num, order=order, loss_num=loss_num)
    return net


def load_reference_quantize(bits_w, bits_i=8, bits_g=8, exp_dir="/tmp/cunet_ref_exp"):
    """Import the reference's utils/quantize.py (parses argv at import, quantize.py:8-11)."""
    if not available():
        raise ReferenceUnavailable("reference tree not present at %s" % REF_ROOT)
    os.makedirs(exp_dir, exist_ok=True)
    saved_argv = list(sys.argv)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in list(sys.modules)
                  if k == "utils" or k.startswith("utils.") or k == "options"
                  or k.startswith("options.")}
    for k in saved_mods:
        sys.modules.pop(k, None)
    try:
        sys.argv = ["x", "--exp_id", "q", "--exp_dir", exp_dir, "--bits_w", str(bits_w),
                    "--bits_i", str(bits_i), "--bits_g", str(bits_g)]
        sys.path.insert(0, REF_ROOT)
        import importlib
        with contextlib.redirect_stdout(io.StringIO()):
            q = importlib.import_module("utils.quantize")
        return q
    finally:
        sys.argv = saved_argv
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "utils" or k.startswith("utils.") or k == "options" or k.startswith("options."):
                sys.modules.pop(k, None)
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v


def load_reference_pylib():
    """Namespaces of the reference's pylib/HumanAug.py, pylib/Evaluation.py, pylib/HumanPts.py as module objects."""
    if not available():
        raise ReferenceUnavailable("reference tree not present at %s" % REF_ROOT)
    import warnings
    mods = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for name in ("HumanAug", "Evaluation", "HumanPts"):
            path = os.path.join(REF_ROOT, "pylib", name + ".py")
            with open(path, "r") as f:
                src = f.read().expandtabs(8)
            src = re.sub(r"^(\s*)print\s+'([^']*)'\s*$", r"\1print('\2')", src, flags=re.M)
            mod = types.ModuleType("reference_pylib_" + name)
            mod.__file__ = path
            # stand-ins during the exec only: the implicit relative import of Evaluation.py:4, and matplotlib.path
            # (HumanPts.py:6, used by a polygon helper that is not on the path) which this image does not have
            stubs = {}
            if name == "Evaluation":
                stubs["HumanAug"] = mods["HumanAug"]
            if name == "HumanPts" and "matplotlib" not in sys.modules:
                mpl, mpath = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.path")
                mpath.Path = object
                mpl.path = mpath
                stubs.update({"matplotlib": mpl, "matplotlib.path": mpath})
            saved = {k: sys.modules.get(k) for k in stubs}
            sys.modules.update(stubs)
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    exec(compile(src, path, "exec"), mod.__dict__)
            finally:
                for k, v in saved.items():
                    if v is None:
                        sys.modules.pop(k, None)
                    else:
                        sys.modules[k] = v
            mods[name] = mod
    return mods
