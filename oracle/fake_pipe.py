"""TEST INFRASTRUCTURE — synthetic stand-in for the diffusers pipeline.

The reference's ``UCE()`` (trainscripts/uce_sd_erase.py:12-91,
trainscripts/uce_sd_debias.py:37-149) touches the pipeline only through
``pipe.unet.named_modules()`` (:17), ``pipe.encode_prompt`` (:29-32),
``pipe.tokenizer`` (:34-39) and, for debias, ``pipe(prompt, ...).images`` /
``pipe.unet.load_state_dict`` / ``pipe.to`` (uce_sd_debias.py:19-26,90).  This
module provides exactly that surface on seeded synthetic tensors so the real
reference code, the oracle port and the CUDA path can all be driven with
identical inputs without diffusers, model weights or a tokenizer vocabulary.
"""
from __future__ import annotations

import zlib

import torch
from torch import nn

# (module path, d_out) for every attn2.to_k/to_v in named_modules() order
# (down_blocks, up_blocks, mid_block — SURVEY.md §8a / Appendix A).


def sd14_attn2_names():
    names = []
    for b, d in ((0, 320), (1, 640), (2, 1280)):
        for a in (0, 1):
            names.append((f"down_blocks.{b}.attentions.{a}.transformer_blocks.0.attn2", d))
    for b, d in ((1, 1280), (2, 640), (3, 320)):
        for a in (0, 1, 2):
            names.append((f"up_blocks.{b}.attentions.{a}.transformer_blocks.0.attn2", d))
    names.append(("mid_block.attentions.0.transformer_blocks.0.attn2", 1280))
    return names


def sdxl_attn2_names():
    names = []
    for b, d, nt in ((1, 640, 2), (2, 1280, 10)):
        for a in (0, 1):
            for t in range(nt):
                names.append((f"down_blocks.{b}.attentions.{a}.transformer_blocks.{t}.attn2", d))
    for b, d, nt in ((0, 1280, 10), (1, 640, 2)):
        for a in (0, 1, 2):
            for t in range(nt):
                names.append((f"up_blocks.{b}.attentions.{a}.transformer_blocks.{t}.attn2", d))
    for t in range(10):
        names.append((f"mid_block.attentions.0.transformer_blocks.{t}.attn2", 1280))
    return names


def layer_table(kind: str, k_dim: int | None = None):
    """[(qualified linear name, d_out, K)] for a model family or a custom spec.

    kind: 'sd14' (32 projections, K=768), 'sdxl' (140, K=2048), or
    'tiny:<d1>,<d2>,...' (one attn2 per entry, K from ``k_dim``).
    """
    if kind == "sd14":
        blocks, K = sd14_attn2_names(), 768
    elif kind == "sdxl":
        blocks, K = sdxl_attn2_names(), 2048
    elif kind.startswith("tiny:"):
        ds = [int(x) for x in kind[5:].split(",") if x]
        blocks = [(f"down_blocks.0.attentions.{i}.transformer_blocks.0.attn2", d) for i, d in enumerate(ds)]
        K = k_dim or 64
    else:
        raise ValueError(kind)
    if k_dim is not None:
        K = k_dim
    out = []
    for path, d in blocks:
        out.append((path + ".to_k", d, K))
        out.append((path + ".to_v", d, K))
    return out


def _set_submodule(root: nn.Module, path: str, mod: nn.Module):
    parts = path.split(".")
    cur = root
    for p in parts[:-1]:
        if not hasattr(cur, p):
            cur.add_module(p, nn.Module())
        cur = getattr(cur, p)
    cur.add_module(parts[-1], mod)


def build_unet(table, seed: int = 0, w_scale: float = 0.03, also_other: bool = True) -> nn.Module:
    """nn.Module tree whose named_modules() contains the attn2.to_k/to_v Linears.

    ``also_other`` adds attn1/to_q/to_out decoys that the selection rule
    (uce_sd_erase.py:18) must skip.
    """
    g = torch.Generator().manual_seed(seed)
    unet = nn.Module()
    seen_blocks = set()
    for name, d, K in table:
        lin = nn.Linear(K, d, bias=False)
        with torch.no_grad():
            lin.weight.copy_(torch.randn(d, K, generator=g) * w_scale)
        _set_submodule(unet, name, lin)
        blk = name.rsplit(".", 2)[0]
        if also_other and blk not in seen_blocks:
            seen_blocks.add(blk)
            _set_submodule(unet, blk + ".attn2.to_q", nn.Linear(8, 8, bias=False))
            _set_submodule(unet, blk + ".attn1.to_k", nn.Linear(8, 8, bias=False))
            _set_submodule(unet, blk + ".attn1.to_v", nn.Linear(8, 8, bias=False))
    for p in unet.parameters():
        p.requires_grad_(False)
    return unet


class FakeTokenizer:
    """Whitespace 'tokenizer': mask = BOS + words + EOS, padded to 77.

    Reproduces what uce_sd_erase.py:34-39 needs: ``attention_mask.sum()-2`` is
    the index of the last real token; the empty prompt gives index 0 (BOS).
    """

    model_max_length = 77

    def __call__(self, text, padding=None, max_length=None, truncation=None, return_tensors=None):
        n = min(len(text.split()), self.model_max_length - 2)
        mask = torch.zeros(1, self.model_max_length, dtype=torch.long)
        mask[0, : n + 2] = 1
        return {"attention_mask": mask, "input_ids": torch.zeros(1, self.model_max_length, dtype=torch.long)}


class _Images:
    def __init__(self, images):
        self.images = images


class FakePipe:
    """Synthetic pipeline: seeded [1,77,K] embedding per prompt + fake tokenizer.

    correlated=True draws every token row as 0.8*u + 0.6*randn (cos≈0.64
    between prompts) to exercise conditioning; rows are scaled
    to norm ≈ ``emb_norm`` (CLIP-like 28).
    """

    def __init__(self, table, seed=0, correlated=False, emb_norm=28.0, w_scale=0.03, dtype=torch.float32):
        self.table = list(table)
        self.K = self.table[0][2]
        self.seed = seed
        self.correlated = correlated
        self.emb_norm = emb_norm
        self.unet = build_unet(self.table, seed=seed, w_scale=w_scale)
        self.tokenizer = FakeTokenizer()
        self._cache = {}
        gu = torch.Generator().manual_seed(seed + 7919)
        self._u = torch.randn(self.K, generator=gu)
        self.calls = []          # debias: record of pipe(...) invocations
        self.loaded = []         # debias: state dicts passed to load_state_dict
        self._orig_load = self.unet.load_state_dict
        self.dtype = dtype

    # --- reference surface -------------------------------------------------
    def encode_prompt(self, prompt, device=None, num_images_per_prompt=1, do_classifier_free_guidance=False):
        if prompt not in self._cache:
            g = torch.Generator().manual_seed((zlib.crc32(prompt.encode()) + 1000003 * self.seed) % (2**31))
            e = torch.randn(1, 77, self.K, generator=g)
            if self.correlated:
                e = 0.8 * self._u + 0.6 * e
            e = e * (self.emb_norm / e.norm(dim=-1, keepdim=True))
            self._cache[prompt] = e.to(torch.float32)
        return (self._cache[prompt], None)

    def to(self, *a, **k):  # uce_sd_debias.py:90 — solver tensors are deep copies, unaffected
        return self

    def set_progress_bar_config(self, **k):
        pass

    def __call__(self, prompt, num_inference_steps=20, num_images_per_prompt=10, guidance_scale=7.5, **kw):
        self.calls.append((prompt, num_inference_steps, num_images_per_prompt, guidance_scale))
        return _Images([(prompt, i) for i in range(num_images_per_prompt)])

    # --- helpers for tests ---------------------------------------------------
    def token_row(self, prompt: str) -> torch.Tensor:
        """The single [K] row the reference keeps for ``prompt`` (uce_sd_erase.py:34-42)."""
        idx = int(self.tokenizer(prompt)["attention_mask"].sum()) - 2
        return self.encode_prompt(prompt)[0][0, idx, :].clone()

    def weights(self):
        """[(name, W[d,K])] selected by the reference's rule (uce_sd_erase.py:17-20)."""
        out = []
        for name, m in self.unet.named_modules():
            if "attn2" in name and (name.endswith("to_v") or name.endswith("to_k")):
                out.append((name, m.weight.detach()))
        return out


class ScriptedClip:
    """Deterministic zero-shot 'classifier' for the debias loop.

    ``script[it][edit_concept]`` is the list of top-1 labels returned for the
    images of that concept at ``get_ratios`` call ``it`` (uce_sd_debias.py:27-28).
    """

    def __init__(self, script, edit_concepts):
        self.script = script
        self.edit_concepts = list(edit_concepts)
        self.n_calls = 0

    def __call__(self, images, candidate_labels):
        it = self.n_calls // len(self.edit_concepts)
        concept = images[0][0]
        self.n_calls += 1
        labels = self.script[min(it, len(self.script) - 1)][concept]
        assert len(labels) == len(images)
        out = []
        for lab in labels:
            rest = [c for c in candidate_labels if c != lab]
            out.append([{"label": lab, "score": 0.9}] + [{"label": r, "score": 0.1 / max(1, len(rest))} for r in rest])
        return out


class _StateDictHolder:
    def __init__(self, state):
        self._state = state

    def state_dict(self):
        return dict(self._state)


class FakeGenPipe:
    """Synthetic generation pipeline for the generate_images() drop-in: a tiny U-Net state dict (diffusers names),
    seeded prompt embeddings, and a deterministic stand-in for VAE decode + image post-processing."""

    def __init__(self, cfg, weights, latent_size=16, seed=0):
        self.cfg = cfg
        self.unet = _StateDictHolder(weights)
        self.latent_size = latent_size
        self.seed = seed

    def _emb(self, text, n):
        g = torch.Generator().manual_seed((zlib.crc32(text.encode()) + 7 * self.seed) % (2 ** 31))
        return torch.randn(1, 77, self.cfg["cross_attention_dim"], generator=g).expand(n, -1, -1).contiguous()

    def encode_prompt(self, prompt, device=None, num_images_per_prompt=1, do_classifier_free_guidance=True, **kw):
        return self._emb(prompt, num_images_per_prompt), self._emb("", num_images_per_prompt)

    @staticmethod
    def latents_to_uint8(latents):
        x = torch.sigmoid(latents.float().cpu()[:, :3])
        x = torch.nn.functional.interpolate(x, scale_factor=4, mode="nearest")
        return (x.permute(0, 2, 3, 1) * 255).round().to(torch.uint8).numpy()

    def decode_latents_to_pil(self, latents):
        from PIL import Image
        return [Image.fromarray(a) for a in self.latents_to_uint8(latents)]
