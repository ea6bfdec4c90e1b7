"""Import the REAL reference loss modules from /root/reference (present in the build container only).

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py to mint fixtures and by the optional
"reference is present" tests.  Never imported by the product, never needed on the GPU box.
"""
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PGICA_REFERENCE_ROOT", "/root/reference")
_PKG = "preference_guided_image_captioning_alignment"


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", _PKG, "models", "components.py"))


def load_components():
    """pkg/models/components.py loaded by file path (torch only; the package __init__ would pull in peft)."""
    path = os.path.join(REFERENCE_ROOT, "src", _PKG, "models", "components.py")
    spec = importlib.util.spec_from_file_location("_pgica_ref_components", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_model_losses():
    """ContrastiveLoss / PreferenceLoss from pkg/models/model.py; `peft` is not installed, so a stub module
    satisfies the import at model.py:45 (LoRA is not on the loss-head path)."""
    if "peft" not in sys.modules:
        stub = types.ModuleType("peft")
        stub.LoraConfig = type("LoraConfig", (), {"__init__": lambda self, *a, **k: None})
        stub.get_peft_model = lambda model, cfg: model
        stub.TaskType = types.SimpleNamespace(FEATURE_EXTRACTION="FEATURE_EXTRACTION", CAUSAL_LM="CAUSAL_LM")
        sys.modules["peft"] = stub
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    mod = importlib.import_module(_PKG + ".models.model")
    return mod.ContrastiveLoss, mod.PreferenceLoss
