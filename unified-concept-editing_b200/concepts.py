"""Host-side concept handling shared by the erase and debias drivers.

Mirrors what the reference does before its arithmetic starts: which U-Net modules are edited
(trainscripts/uce_sd_erase.py:15-20), how a prompt becomes ONE text-embedding row (:26-42),
how CLI concept strings are parsed, broadcast and expanded (:134-190).
"""
from __future__ import annotations

import torch

ART_TEMPLATES = ("painting by {}", "art by {}", "artwork by {}", "picture by {}", "style of {}")
OBJECT_TEMPLATES = ("image of {}", "photo of {}", "portrait of {}", "picture of {}", "painting of {}")


def select_projections(unet):
    """[(qualified name, module)] of every cross-attention key/value projection, in
    named_modules() order — the reference's rule: 'attn2' in the name and the name ends with
    to_k or to_v (uce_sd_erase.py:17-20)."""
    picked = []
    for name, module in unet.named_modules():
        if "attn2" in name and name.endswith(("to_k", "to_v")):
            picked.append((name, module))
    return picked


def concept_row(pipe, prompt: str, device) -> torch.Tensor:
    """The single embedding row the reference keeps for a prompt: the last real token,
    index ``attention_mask.sum() - 2`` (uce_sd_erase.py:29-42).  Returns [K] fp32."""
    emb = pipe.encode_prompt(prompt=prompt, device=device, num_images_per_prompt=1, do_classifier_free_guidance=False)
    tok = pipe.tokenizer(prompt, padding="max_length", max_length=pipe.tokenizer.model_max_length, truncation=True,
                         return_tensors="pt")
    last = int(tok["attention_mask"].sum()) - 2
    return emb[0][0, last, :].to(torch.float32)


def can_batch_encode(pipe) -> bool:
    """True for single-text-encoder pipelines (SD-1.x / 2.x layout) whose tokenizer and encoder can be called directly."""
    return (getattr(pipe, "text_encoder", None) is not None and getattr(pipe, "text_encoder_2", None) is None
            and callable(getattr(pipe, "tokenizer", None)))


def can_use_text_engine(pipe, device) -> bool:
    """The B200 text encoder (uce_b200.clip_text) takes a plain CLIPTextModel: one text encoder, quick-GELU MLP, no attention mask,
    no projection — the SD-1.x layout — on a CUDA device."""
    te = getattr(pipe, "text_encoder", None)
    cfg = getattr(te, "config", None)
    if not can_batch_encode(pipe) or cfg is None or not str(device).startswith("cuda") or not torch.cuda.is_available():
        return False
    return (getattr(cfg, "hidden_act", None) == "quick_gelu" and not getattr(cfg, "use_attention_mask", False)
            and hasattr(te, "state_dict") and getattr(cfg, "hidden_size", 4096) <= 1024)


def embed_concepts_engine(pipe, prompts, device, batch_size: int = 128) -> dict:
    """``embed_concepts_batched`` with the forward on the B200 text-encoder engine (fp32 CUDA kernels, include/clip_text_b200.h): every
    distinct prompt once, ``batch_size`` prompts per forward, the kept rows selected on the device."""
    from .clip_text import ClipTextEngine
    uniq = list(dict.fromkeys(prompts))
    tk = pipe.tokenizer
    eng = ClipTextEngine(pipe.text_encoder.state_dict(), pipe.text_encoder.config.num_attention_heads, device=device, max_batch=batch_size)
    try:
        tok = tk(uniq, padding="max_length", max_length=tk.model_max_length, truncation=True, return_tensors="pt")
        rows = eng.concept_rows(tok["input_ids"], tok["attention_mask"])
    finally:
        eng.close()
    return {p: rows[j] for j, p in enumerate(uniq)}


def embed_concepts_batched(pipe, prompts, device, batch_size: int = 128) -> dict:
    """Same rows as ``embed_concepts`` from BATCHED text-encoder calls: the reference encodes one prompt per forward
    (uce_sd_erase.py:26-42), which dominates an edit of hundreds of concepts.  What encode_prompt does for a plain
    SD-1.x / 2.x pipeline is restated here — tokenizer with max_length padding and truncation, ``text_encoder(ids)[0]`` (the attention mask
    is passed only if the encoder's config asks for it), no clip_skip / LoRA scale / textual inversion — and the kept row is the same
    ``attention_mask.sum() - 2``.  Rows agree with the one-by-one path to fp32 rounding (tests/test_concepts_batched.py)."""
    uniq = []
    for p in prompts:
        if p not in uniq:
            uniq.append(p)
    rows = {}
    tk = pipe.tokenizer
    use_mask = bool(getattr(getattr(pipe.text_encoder, "config", None), "use_attention_mask", False))
    for i in range(0, len(uniq), batch_size):
        chunk = uniq[i:i + batch_size]
        tok = tk(chunk, padding="max_length", max_length=tk.model_max_length, truncation=True, return_tensors="pt")
        ids, mask = tok["input_ids"], tok["attention_mask"]
        out = pipe.text_encoder(ids.to(device), attention_mask=mask.to(device) if use_mask else None)[0]
        last = mask.sum(dim=1) - 2
        for j, p in enumerate(chunk):
            rows[p] = out[j, int(last[j]), :].to(torch.float32)
    return rows


def embed_concepts(pipe, prompts, device, batched: bool | None = None) -> dict:
    """prompt -> [K] row, each distinct prompt encoded once (uce_sd_erase.py:26-28).  A pipeline with a plain SD-1.x CLIP text encoder on
    a CUDA device goes through the B200 text-encoder engine (one batched fp32 forward; UCE_TEXT_ENGINE=0 disables it); ``batched``
    (default: the UCE_BATCHED_ENCODE=1 environment switch) routes other single-encoder pipelines through batched calls of their own
    torch encoder; everything else (SDXL's two encoders, synthetic test pipes) gets one encode_prompt call per prompt, as the
    reference does."""
    import os
    if os.environ.get("UCE_TEXT_ENGINE", "1") != "0" and can_use_text_engine(pipe, device):
        return embed_concepts_engine(pipe, prompts, device)      # SD-1.x text encoder: one batched forward on the B200 engine
    if batched is None:
        batched = os.environ.get("UCE_BATCHED_ENCODE") == "1"
    if batched and can_batch_encode(pipe):
        return embed_concepts_batched(pipe, prompts, device)
    rows = {}
    for p in prompts:
        if p not in rows:
            rows[p] = concept_row(pipe, p, device)
    return rows


def split_concepts(arg: str | None):
    """';'-separated CLI list -> stripped strings (uce_sd_erase.py:134,141,151); None -> []."""
    if arg is None:
        return []
    return [c.strip() for c in arg.split(";")]


def resolve_guides(edit_concepts, guide_arg: str | None, concept_type: str):
    """Guide list paired with the edit list (uce_sd_erase.py:136-145): default 'art' for art
    concepts else the empty prompt; a single guide is broadcast; lengths must match."""
    if guide_arg is None:
        guide_arg = "art" if concept_type == "art" else ""
    guides = [c.strip() for c in guide_arg.split(";")]
    if len(guides) == 1:
        guides = guides * len(edit_concepts)
    if len(guides) != len(edit_concepts):
        raise Exception("Error! The length of erase concepts and their corresponding guide concepts do not match. "
                        "Please make sure they are seperated by ; and are of equal sizes")
    return guides


def expand_prompts(edit_concepts, guide_concepts, concept_type: str):
    """--expand_prompts true: five templated variants per (concept, guide) pair appended after
    the originals, wording by concept type (uce_sd_erase.py:155-190).  Preserve concepts are
    never expanded."""
    templates = ART_TEMPLATES if concept_type == "art" else OBJECT_TEMPLATES
    edits, guides = list(edit_concepts), list(guide_concepts)
    for concept, guide in zip(list(edit_concepts), list(guide_concepts)):
        edits.extend(t.format(concept) for t in templates)
        guides.extend(t.format(guide) for t in templates)
    return edits, guides
