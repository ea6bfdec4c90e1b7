"""Retrieval-style scoring on the similarity kernel.

    compute_similarity   PreferenceGuidedCaptioningModel.compute_similarity, pkg/models/model.py:925-954:
                         (B, B) matrix  image_embeddings @ text_embeddings.T / temperature
    paired_scores        the per-pair score the CLIP-score loop collects one `.item()` at a time
                         (pkg/evaluation/metrics.py:380-439: logits_per_image of ONE image and ONE caption =
                         logit_scale * cos(image, caption)), for a whole batch in one launch
    retrieval_ranks      rank of the matching caption among all captions of the batch (recall@k style evaluation)

All three run on `pgica_similarity` / `pgica_gemm_lse` (tcgen05 GEMM, csrc/gemm_lse.cu).  fp32 embeddings go in as their
two-term bf16 split (depth 3*D), so scores agree with the reference's fp32 matmul to ~1e-6 relative; bf16 embeddings
are used as they are.  Forward-only (evaluation); no CPU path.
"""
import torch

from . import functional as F


def _operands(a, b):
    if a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16:
        return a.contiguous(), b.contiguous()
    al, _ = F.split3(a)
    _, br = F.split3(b)
    return al, br


@torch.no_grad()
def compute_similarity(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor,
                       temperature: float = 0.07) -> torch.Tensor:
    """(B_i, D), (B_t, D) embeddings as given (the model normalises them, model.py:826-829) -> (B_i, B_t) fp32
    similarity / temperature."""
    a, b = _operands(image_embeddings, text_embeddings)
    return F.similarity(a, b, 1.0 / float(temperature))


@torch.no_grad()
def paired_scores(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor, logit_scale: float = 100.0,
                  normalize: bool = True) -> torch.Tensor:
    """score[i] = logit_scale * <img_i, txt_i> (cosine when `normalize`): the CLIP-score of pair i.  (B,) fp32.
    Uses the diagonal gather of the GEMM+LSE kernel, so the (B, B) matrix is never written."""
    if normalize:
        _, _, _, al, _ = F.rownorm_fwd(image_embeddings, 1e-12, split=True)
        _, _, _, _, br = F.rownorm_fwd(text_embeddings, 1e-12, split=True)
    else:
        al, br = _operands(image_embeddings, text_embeddings)
    _, diag = F.gemm_lse(al, br, float(logit_scale), None, 0, want_tgt=True)
    return diag


@torch.no_grad()
def retrieval_ranks(image_embeddings: torch.Tensor, text_embeddings: torch.Tensor) -> torch.Tensor:
    """rank[i] = number of captions that score strictly higher than caption i for image i (0 = retrieved first)."""
    sim = compute_similarity(image_embeddings, text_embeddings, 1.0)
    return (sim > sim.diagonal().unsqueeze(1)).sum(dim=1)


def model_compute_similarity(self, images, captions, caption_mask):
    """Bound over PreferenceGuidedCaptioningModel.compute_similarity by install(): same signature and result
    (pkg/models/model.py:925-954), the final matmul on the similarity kernel."""
    outputs = self(images=images, caption_ids=captions, caption_mask=caption_mask, mode="contrastive")
    img, txt = outputs["image_embeddings"], outputs["text_embeddings"]
    if not img.is_cuda:
        return torch.matmul(img, txt.t()) / self.temperature
    return compute_similarity(img, txt, self.temperature)
