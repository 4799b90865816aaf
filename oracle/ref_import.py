"""Import the unmodified reference from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

The reference needs three shims; no reference file is modified:
  1. ``torchsummary`` (imported at models/EELUnet.py:4) and ``matplotlib`` (utils/tools.py:3,8)
     are not installed -> stub modules;
  2. /root/reference is read-only -> no bytecode writes;
  3. ``visualize_feature_maps`` (called 9x per forward, models/EELUnet.py:389-462) writes PNGs ->
     replaced by a no-op.  It does not touch numerics.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("EEL_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "EELUnet.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    """Attribute sink: matplotlib.pyplot.<whatever>(...) -> no-op."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def load():
    """Returns (EELUnet class, edge_BceDiceLoss class, Unet class) of the real reference."""
    if not available():
        raise RuntimeError("reference not present at %s (it only exists in the build container)" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import torch  # noqa: F401  (before the stubs: torch's import walks sys.modules with inspect, which chokes on attribute sinks)
    try:
        import torchsummary  # noqa: F401
    except Exception:
        _stub("torchsummary", summary=lambda *a, **k: None)
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        sink = _Anything()

        def _attr(name):
            if name.startswith("__"):          # inspect.getmodule() probes __file__ etc. on every entry of sys.modules
                raise AttributeError(name)
            return sink

        plt.__getattr__ = _attr  # type: ignore[attr-defined]
        mpl.pyplot = plt
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.EELUnet as ref_model
    import models.Unet as ref_unet
    import utils.Loss as ref_loss

    ref_model.visualize_feature_maps = lambda *a, **k: None
    return ref_model.EELUnet, ref_loss.edge_BceDiceLoss, ref_unet.Unet


def load_module():
    """The reference's models.EELUnet python module (for its building blocks)."""
    load()
    import models.EELUnet as ref_model

    return ref_model


def load_evaluate():
    """The reference's ``evaluate`` python module (evaluate.py: ``seg2bnd`` :25, ``boundary_f1_score`` :43, ``evaluate`` :62).
    Its top-level imports pull in comparison models that need timm / mmcv (not installed, out of scope): those three modules
    are replaced by empty stubs -- evaluate()'s arithmetic does not touch them."""
    load()
    for mod, names in (("models.egeunet", ["EGEUNet"]), ("models.malunet", ["MALUNet"]), ("models.unext", ["UNext", "UNext_S"])):
        if mod not in sys.modules:
            _stub(mod, **{n: type(n, (), {}) for n in names})
    import evaluate as ref_evaluate

    return ref_evaluate
