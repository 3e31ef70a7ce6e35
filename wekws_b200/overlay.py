"""Makes the unmodified reference scripts use this implementation.

``wekws/bin/score.py`` and friends do ``from wekws.model.kws_model import init_model``
(score.py:30).  After ``patch_reference()`` that name resolves to wekws_b200's factory, so
``model = init_model(configs['model'])`` / ``load_checkpoint(model, path)`` /
``model(feats)`` run on the fused kernels without editing the reference tree.
"""
import importlib
import sys
import types


def patch_reference() -> bool:
    """Returns True if a real ``wekws`` package was found and patched, False if a thin stub
    package ``wekws.model.kws_model`` was registered instead (reference not importable)."""
    from . import kws_model as ours
    try:
        ref = importlib.import_module("wekws.model.kws_model")
        ref.KWSModel = ours.KWSModel
        ref.init_model = ours.init_model
        return True
    except Exception:
        pkg = sys.modules.setdefault("wekws", types.ModuleType("wekws"))
        sub = sys.modules.setdefault("wekws.model", types.ModuleType("wekws.model"))
        pkg.model = sub
        sys.modules["wekws.model.kws_model"] = ours
        sub.kws_model = ours
        return False
