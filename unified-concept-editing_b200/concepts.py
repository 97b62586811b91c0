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


def embed_concepts(pipe, prompts, device) -> dict:
    """prompt -> [K] row, each distinct prompt encoded once (uce_sd_erase.py:26-28)."""
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
