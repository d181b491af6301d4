"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the *reference's own* ``lib/decompose.py`` and ``lib/net.py`` from
``/root/reference`` (read-only) under a handful of module shims, so that the
numpy restatement in ``oracle/cp_oracle.py`` can be pinned against the real
reference code and golden vectors can be generated (``oracle/make_golden.py``).

``/root/reference`` only exists in the build container; on the GPU box this
module raises ``ReferenceUnavailable`` and nothing in tests/bench/smoke depends
on it at run time (they use the committed fixtures under ``tests/golden``).

Shims:
  * ``easydict.EasyDict``            -- attr-dict (lib/cfgs.py:1)
  * ``IPython.embed``                -- no-op (lib/decompose.py:9)
  * ``termcolor.colored``            -- identity (lib/utils.py:4)
  * ``sklearn.linear_model.RandomizedLasso`` -- removed upstream, still
    imported by lib/decompose.py:7
  * ``scipy.linalg.pinv(x, 1e-6)``   -- positional cond no longer accepted
    (lib/decompose.py:152); patched after import
  * ``caffe`` / ``caffe.proto.caffe_pb2`` / ``matplotlib`` -- empty stubs so
    that lib/net.py *imports*; its methods are then driven with a duck-typed
    ``self`` (see make_golden.py), no Caffe object is ever constructed.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CP_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


class _EasyDict(dict):
    """Minimal stand-in for easydict.EasyDict (attribute access on a dict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load_reference():
    """Returns (decompose, net, cfgs) modules of the unmodified reference."""
    if _loaded:
        return _loaded["decompose"], _loaded["net"], _loaded["cfgs"]
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, "lib")):
        raise ReferenceUnavailable(REFERENCE_ROOT + " not present")

    if "easydict" not in sys.modules:
        _stub("easydict", EasyDict=_EasyDict)
    if "IPython" not in sys.modules:
        _stub("IPython", embed=lambda *a, **k: None)
    if "termcolor" not in sys.modules:
        _stub("termcolor", colored=lambda s, *a, **k: s)
    if "matplotlib" not in sys.modules:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if "caffe" not in sys.modules:
        caffe = _stub("caffe", TEST=1, TRAIN=0)
        proto = _stub("caffe.proto")
        pb2 = _stub("caffe.proto.caffe_pb2")

        class _Msg:  # attribute sink for builder.py class bodies
            def __init__(self, *a, **k):
                pass

            def __getattr__(self, k):
                return _Msg()

        pb2.SolverParameter = _Msg
        pb2.NetParameter = _Msg
        pb2.LayerParameter = _Msg
        pb2.__getattr__ = lambda k: _Msg  # type: ignore
        proto.caffe_pb2 = pb2
        caffe.proto = proto
    import sklearn.linear_model as _lm

    if not hasattr(_lm, "RandomizedLasso"):
        _lm.RandomizedLasso = None

    # the reference's package is literally called ``lib``; keep any existing
    # module of that name out of the way while importing.
    saved = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    cwd = os.getcwd()
    try:
        cfgs = importlib.import_module("lib.cfgs")
        decompose = importlib.import_module("lib.decompose")
        try:
            net = importlib.import_module("lib.net")
        except Exception as e:  # pragma: no cover - reported by make_golden
            net = None
            _loaded["net_error"] = repr(e)
    finally:
        os.chdir(cwd)
        sys.path.remove(REFERENCE_ROOT)
    # keep them reachable under a private name, restore whatever was there
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        sys.modules["_cp_reference_" + k] = sys.modules.pop(k)
    sys.modules.update(saved)

    import scipy.linalg

    decompose.pinv = lambda x: scipy.linalg.pinv(x, rtol=1e-6)  # decompose.py:152
    _loaded.update(decompose=decompose, net=net, cfgs=cfgs)
    return decompose, net, cfgs
