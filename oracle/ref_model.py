"""Build the REAL reference model / trainer offline (random-init weights, no downloads).  TEST INFRASTRUCTURE ONLY.

The reference package is pure Python; it is found either in the build container (`/root/reference/src`) or in the
git-ignored install `baseline/_ref` (`python -m pip install --no-index --no-build-isolation --no-deps --target
baseline/_ref /root/reference`), which travels to the GPU box with the repo snapshot.  Its third-party imports that are
not in this image (`peft`, `accelerate`, `mlflow`) are satisfied by inert stubs; the `from_pretrained` constructors it
calls (pkg/models/model.py:126,311-312,505-506) are rebound to random-init models of the same architecture:

    CLIPModel        -> CLIPModel(CLIPConfig())                                   (defaults = ViT-B/32)
    AutoModel        -> GPT2Model(GPT2Config(n_embd, n_layer, n_head))            (text encoder, DialoGPT-medium shape)
    GPT2LMHeadModel  -> GPT2LMHeadModel(GPT2Config(...))                          (caption decoder, tied lm_head)
    tokenizers       -> a stand-in with len() == 50257 that grows with add_special_tokens (decoder vocab 50260)

Nothing in the product imports this module.  Callers: tests/test_install_real.py, tools/cfg5_step.py and the cfg5
leg of bench.py (which times the reference's own Stage-2 step next to the patched one).
"""
import importlib
import os
import sys
import types

_PKG = "preference_guided_image_captioning_alignment"
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_src():
    """Directory to put on sys.path so that `import preference_guided_image_captioning_alignment` finds the reference."""
    for cand in (os.environ.get("PGICA_REFERENCE_SRC"), os.path.join(_ROOT, "baseline", "_ref"),
                 "/root/reference/src"):
        if cand and os.path.isfile(os.path.join(cand, _PKG, "models", "model.py")):
            return cand
    return None


def available():
    return reference_src() is not None


class _FakeTokenizer:
    """len() == 50257 like GPT-2's; pad/sep unset; add_special_tokens grows the vocabulary (model.py:314-324,509-511)."""

    def __init__(self, base=50257):
        self._n = base
        self.pad_token = self.sep_token = self.bos_token = self.eos_token = None
        self.pad_token_id = self.eos_token_id = None

    def __len__(self):
        return self._n

    def add_special_tokens(self, tokens):
        for key in tokens:
            setattr(self, key, tokens[key])
            setattr(self, key + "_id", self._n)
            self._n += 1
        return len(tokens)

    @classmethod
    def from_pretrained(cls, *a, **k):
        return cls()


def _module(name):
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)  # transformers probes find_spec("peft") / ("accelerate")
    return m


def _install_stubs():
    import transformers.trainer_utils  # noqa: F401  (probes optional packages at import: do it before the stubs exist)
    if "peft" not in sys.modules:
        stub = _module("peft")
        stub.LoraConfig = type("LoraConfig", (), {"__init__": lambda self, *a, **k: None})
        stub.get_peft_model = lambda model, cfg: model
        stub.TaskType = types.SimpleNamespace(FEATURE_EXTRACTION="FEATURE_EXTRACTION", CAUSAL_LM="CAUSAL_LM")
        sys.modules["peft"] = stub
    if "accelerate" not in sys.modules:
        acc = _module("accelerate")

        class Accelerator:  # only what the trainer touches (pkg/training/trainer.py:189-201, 472, 492, 510)
            def __init__(self, *a, device=None, **k):
                import torch
                self.device = device or torch.device("cuda" if torch.cuda.is_available() else "cpu")
                self.is_main_process = True
                self.mixed_precision = k.get("mixed_precision", "no")

            def prepare(self, *objs):
                import torch
                out = [o.to(self.device) if isinstance(o, torch.nn.Module) else o for o in objs]
                return out[0] if len(out) == 1 else tuple(out)

            def accumulate(self, model):
                import contextlib
                return contextlib.nullcontext()

            def backward(self, loss):
                loss.backward()

            def clip_grad_norm_(self, params, max_norm):
                import torch
                return torch.nn.utils.clip_grad_norm_(params, max_norm)

            def unwrap_model(self, m):
                return m

            def wait_for_everyone(self):
                pass

        acc.Accelerator = Accelerator
        sys.modules["accelerate"] = acc
    for name in ("mlflow", "wandb"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                m = _module(name)
                m.set_experiment = m.start_run = m.log_params = m.log_metrics = m.end_run = m.init = m.log = \
                    lambda *a, **k: None
                m.run = None
                sys.modules[name] = m


def load_package():
    """Import the reference package (with stubs).  -> the `models.model` module."""
    src = reference_src()
    if src is None:
        raise RuntimeError("reference package not found (neither baseline/_ref nor /root/reference/src)")
    _install_stubs()
    if src not in sys.path:
        sys.path.insert(0, src)
    return importlib.import_module(_PKG + ".models.model")


def build_model(n_embd=1024, n_layer=24, n_head=16, clip_layers=None, seed=42, **model_kwargs):
    """PreferenceGuidedCaptioningModel with random-init backbones (seed as scripts/train.py:312).  Defaults are the
    reference's shapes (GPT-2 Medium decoder + text encoder, CLIP ViT-B/32): 867 M parameters; tests pass a smaller
    GPT-2 to stay within seconds."""
    import torch
    from transformers import CLIPConfig, CLIPModel, GPT2Config, GPT2LMHeadModel, GPT2Model
    mm = load_package()

    def gpt2_cfg():
        return GPT2Config(n_embd=n_embd, n_layer=n_layer, n_head=n_head)

    class _Clip:
        @staticmethod
        def from_pretrained(*a, **k):
            cfg = CLIPConfig()
            if clip_layers is not None:
                cfg.vision_config.num_hidden_layers = clip_layers
                cfg.text_config.num_hidden_layers = min(clip_layers, cfg.text_config.num_hidden_layers)
            return CLIPModel(cfg)

    class _Auto:
        @staticmethod
        def from_pretrained(*a, **k):
            return GPT2Model(gpt2_cfg())

    class _LM:
        @staticmethod
        def from_pretrained(*a, **k):
            return GPT2LMHeadModel(gpt2_cfg())

    saved = {n: getattr(mm, n) for n in ("CLIPModel", "AutoModel", "GPT2LMHeadModel", "AutoTokenizer", "GPT2Tokenizer")}
    mm.CLIPModel, mm.AutoModel, mm.GPT2LMHeadModel = _Clip, _Auto, _LM
    mm.AutoTokenizer = mm.GPT2Tokenizer = _FakeTokenizer
    try:
        torch.manual_seed(seed)
        model = mm.PreferenceGuidedCaptioningModel(**model_kwargs)
    finally:
        for n, v in saved.items():
            setattr(mm, n, v)
    return model


def stage2_batch(batch, seq_len=128, vocab=50257, seed=0, min_len=10, max_len=20, device="cpu"):
    """Synthetic Stage-2 batch with the keys and dtypes of the reference loader (pkg/data/loader.py:487-497):
    right-padded captions of U[min_len, max_len] real tokens, pad id = `vocab` ([PAD], preprocessing.py:182-197,223-231)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    out = {"image": torch.randn(batch, 3, 224, 224, generator=g)}
    for name in ("preferred", "rejected"):
        ids = torch.randint(0, vocab, (batch, seq_len), generator=g)
        lens = torch.randint(min_len, max_len + 1, (batch,), generator=g)
        mask = (torch.arange(seq_len)[None, :] < lens[:, None]).long()
        ids = torch.where(mask.bool(), ids, torch.full_like(ids, vocab))
        out[name + "_ids"], out[name + "_mask"] = ids, mask
    return {k: v.to(device) for k, v in out.items()}


def stage2_micro_step(model, preference_loss, batch):
    """The forward of one Stage-2 micro-step exactly as the trainer issues it (pkg/training/trainer.py:575-603):
    two generation-mode forwards and the preference loss.  -> (loss, preferred_outputs, rejected_outputs)"""
    pref = model(images=batch["image"], caption_ids=batch["preferred_ids"], caption_mask=batch["preferred_mask"],
                 labels=batch["preferred_ids"], mode="generation")
    rej = model(images=batch["image"], caption_ids=batch["rejected_ids"], caption_mask=batch["rejected_mask"],
                labels=batch["rejected_ids"], mode="generation")
    loss = preference_loss(pref["logits"], rej["logits"], batch["preferred_ids"], batch["rejected_ids"],
                           batch["preferred_mask"], batch["rejected_mask"])
    return loss, pref, rej
