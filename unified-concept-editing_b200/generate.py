"""Edited-model image generation: drop-in for ``generate_images()`` of the reference's
evalscripts/generate-images-sd.py:10-46 with the denoise loop (U-Net, classifier-free guidance, scheduler) and the VAE
decode on the B200 engines.  Text encoding stays with the caller's pipeline object.

Row sharding: rank r of a torch.distributed job takes CSV rows r::world — rows are independent,
so the steady state has no collective."""
from __future__ import annotations

import os

import torch

from .png import save_png
from .schedulers import make_plan
from .unet import UNetEngine, cfg_step
from .unet_spec import SD14


class Denoiser:
    """The hot loop of StableDiffusionPipeline.__call__ for ``images`` samples of ONE prompt on given initial latents."""

    def __init__(self, engine: UNetEngine, images: int):
        assert engine.batch == 2 * images
        self.eng, self.B = engine, images
        dev = engine.device
        shp = (images, 4, engine.H, engine.W)
        self.x = torch.empty(shp, device=dev)
        self.x2 = torch.empty((2 * images, 4, engine.H, engine.W), device=dev)
        self.eps2 = torch.empty_like(self.x2)
        self.saved = torch.empty(shp, device=dev)
        self.hist = [torch.zeros(shp, device=dev) for _ in range(4)]
        self.scratch = torch.zeros(shp, device=dev)

    def run(self, latents: torch.Tensor, ctx_uncond_text: torch.Tensor, steps=50, guidance_scale=7.5, scheduler="pndm"):
        """latents [B,4,H,W] (any float dtype, device or host), ctx [2B,77,D] ordered [uncond | text] -> final latents fp32."""
        self.x.copy_(latents.to(self.x.device, torch.float32))
        ctx = ctx_uncond_text.to(self.x.device, torch.float32).contiguous()
        if guidance_scale <= 1.0:                   # diffusers switches classifier-free guidance off at <= 1: the text-conditioned eps alone
            guidance_scale = 1.0                    # (eps_u + 1 * (eps_t - eps_u); the unconditional half of the batch is still computed)
        self.eng.set_context(ctx)                   # cross-attention K / V^T of the prompt: once per row, not once per step
        hist = []                                   # most recent first
        free = list(self.hist)
        for plan in make_plan(scheduler, steps):
            self.x2[: self.B].copy_(self.x); self.x2[self.B:].copy_(self.x)
            self.eng.forward(self.x2, float(plan.t), None, out=self.eps2)
            if plan.save_sample:
                self.saved.copy_(self.x)
            x_in = self.saved if plan.use_saved_sample else self.x
            if plan.append_eps:
                buf = free.pop() if free else hist.pop()
                eps_out = buf
            else:
                buf, eps_out = None, self.scratch
            cfg_step(self.eps2, guidance_scale, x_in, self.x, plan.coeffs, plan.cx, plan.ce, hist=hist[:3], eps_out=eps_out)
            if buf is not None:
                hist.insert(0, buf)
                if len(hist) > 3:
                    free.append(hist.pop())
        return self.x


class _EngineWeights:
    """The ``pipe.unet`` face of EngineGenerator: ``load_state_dict(state, strict=False)`` overwrites engine parameters in place."""

    def __init__(self, engine):
        self._eng = engine

    def load_state_dict(self, state, strict=False):
        self._eng.load_state_dict(state, strict=strict)


class _Images:
    """``.images`` as the pipeline returns them; ``.images_u8`` — when the VAE engine decoded — the same pixels as one uint8
    [B, H, W, 3] tensor still on the device, for a classifier that takes them without a host round trip."""

    def __init__(self, images, images_u8=None):
        self.images = images
        self.images_u8 = images_u8


class EngineGenerator:
    """The part of a StableDiffusionPipeline that ``get_ratios`` of the debias edit uses (trainscripts/uce_sd_debias.py:14-26) with the
    denoise loop on the B200 engine:  ``gen.unet.load_state_dict(edited_weights, strict=False)`` then
    ``gen(concept, num_inference_steps=, num_images_per_prompt=, guidance_scale=).images``.
    Text encoding stays with the wrapped ``pipe`` (``encode_prompt``); VAE decode too unless a decoder engine is given.
    ``engine`` / ``denoiser`` can be injected (tests); otherwise a UNetEngine is built from ``pipe.unet.state_dict()``."""

    def __init__(self, pipe, num_images_per_prompt, device="cuda:0", torch_dtype=torch.bfloat16, scheduler="pndm", unet_config=SD14,
                 engine=None, denoiser=None, vae_engine=None):
        self.pipe, self.B, self.device, self.dtype, self.scheduler = pipe, int(num_images_per_prompt), device, torch_dtype, scheduler
        self.latent = getattr(pipe, "latent_size", 64)
        self._own = engine is None
        if engine is None:
            engine = UNetEngine(unet_config, batch=2 * self.B, H=self.latent, W=self.latent, device=device)
            engine.load_state_dict({k: v for k, v in pipe.unet.state_dict().items()}, strict=False)
            engine.finalize()
        self.eng = engine
        self.den = denoiser if denoiser is not None else Denoiser(engine, self.B)
        self.vae_eng = vae_engine
        self.unet = _EngineWeights(engine)
        self.calls = 0

    def to(self, *a, **k):                 # uce_sd_debias.py:90 moves the pipeline to bf16: the engine already computes in bf16
        return self

    def set_progress_bar_config(self, **k):
        pass

    def __call__(self, prompt, num_inference_steps=50, num_images_per_prompt=None, guidance_scale=7.5, generator=None, **kw):
        n = self.B if num_images_per_prompt is None else int(num_images_per_prompt)
        if n != self.B:
            raise ValueError(f"this generator was built for {self.B} images per prompt, got {n}")
        text, uncond = self.pipe.encode_prompt(prompt=str(prompt), device=self.device, num_images_per_prompt=n,
                                               do_classifier_free_guidance=True)[:2]
        ctx = torch.cat([uncond, text]).to(torch.float32)
        lat = torch.randn((n, 4, self.latent, self.latent), generator=generator, dtype=self.dtype)     # CPU RNG, pipeline dtype
        out = self.den.run(lat, ctx, steps=num_inference_steps, guidance_scale=guidance_scale, scheduler=self.scheduler)
        u8 = None
        if self.vae_eng is not None:
            u8 = self.vae_eng.decode(out.contiguous())
            images = list(u8.cpu().numpy())
        elif hasattr(self.pipe, "decode_latents_to_pil"):
            images = self.pipe.decode_latents_to_pil(out)
        else:
            images = _decode(self.pipe, out, self.dtype)
        self.calls += 1
        return _Images(images, u8)

    def close(self):
        if self._own:
            self.eng.close()


def _rank_world():
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_rank(), torch.distributed.get_world_size()
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def rows_for_rank(df, from_case, till_case, rank, world):
    """CSV rows this rank generates: the reference's inclusive case filter (evalscripts/generate-images-sd.py:33), then the
    surviving rows dealt round-robin to the ranks (prompts data-parallel, no collective in steady state: SURVEY.md 8e)."""
    kept = [row for _, row in df.iterrows() if from_case <= row.case_number <= till_case]
    return kept[rank::world]


def unet_config_of(pipe, default=SD14):
    """Engine configuration and latent size read off ``pipe.unet.config`` (a diffusers UNet2DConditionModel config); pipelines without
    one (synthetic test pipes) get ``default``.  Architectures the engine does not implement fail here, with a message, instead
    of with a shape mismatch deep in the weight upload (the reference accepts any --model_id; this engine is the SD-1.x U-Net)."""
    uc = getattr(getattr(pipe, "unet", None), "config", None)
    latent = getattr(pipe, "latent_size", None)
    if uc is None:
        return dict(default), (latent if latent is not None else 64)
    get = (lambda k, d=None: uc.get(k, d)) if isinstance(uc, dict) else (lambda k, d=None: getattr(uc, k, d))
    down, up = tuple(get("down_block_types", ())), tuple(get("up_block_types", ()))
    ok_down = all(t in ("CrossAttnDownBlock2D", "DownBlock2D") for t in down)
    ok_up = all(t in ("CrossAttnUpBlock2D", "UpBlock2D") for t in up)
    tl = get("transformer_layers_per_block", 1)
    if not down or not ok_down or not ok_up or get("use_linear_projection", False) or tl not in (1, (1,) * len(down), [1] * len(down)) \
            or get("addition_embed_type") or get("class_embed_type") or get("in_channels", 4) != 4:
        raise NotImplementedError("the B200 U-Net engine implements the SD-1.x UNet2DConditionModel layout (conv proj_in/out, one transformer "
                                  f"layer per block, no added-condition embeddings); this pipeline's U-Net config is not supported: {down} / {up}")
    ch = tuple(get("block_out_channels"))
    ahd = get("attention_head_dim", 8)                    # SD-1.x: this config field is the NUMBER of heads
    heads = ahd if isinstance(ahd, int) else ahd[0]
    cfg = dict(in_channels=4, out_channels=get("out_channels", 4), block_out_channels=ch, layers_per_block=get("layers_per_block", 2),
               down_has_attn=tuple(t == "CrossAttnDownBlock2D" for t in down), up_has_attn=tuple(t == "CrossAttnUpBlock2D" for t in up),
               cross_attention_dim=get("cross_attention_dim", 768), heads=heads, norm_groups=get("norm_num_groups", 32), temb_dim=4 * ch[0])
    return cfg, (latent if latent is not None else get("sample_size", 64))


def generate_images(model_id, uce_model_path, prompts_path, save_path, exp_name="test", device="cuda:0", torch_dtype=torch.bfloat16,
                    guidance_scale=7.5, num_inference_steps=100, num_images_per_prompt=10, from_case=0, till_case=1000000,
                    pipe=None, scheduler="pndm", unet_config=None, engine=None, denoiser=None):
    """Same signature and outputs as the reference (``{save_path}/{exp_name}/{case_number}_{i}.png``); ``pipe`` lets a
    caller inject an already-loaded (or synthetic) pipeline object, ``engine`` / ``denoiser`` an already-built U-Net engine and its
    denoise loop (tests/test_generate_golden.py drives the row loop on CPU with stand-ins)."""
    import pandas as pd
    if pipe is None:
        try:
            from diffusers import DiffusionPipeline
        except ImportError as exc:
            raise SystemExit(f"diffusers is required to load '{model_id}' (text encoder, VAE, weights): {exc}")
        pipe = DiffusionPipeline.from_pretrained(model_id, torch_dtype=torch_dtype, safety_checker=None).to(device)
    # an engine built by this call gets the pipeline's U-Net weights with the UCE artifact laid over them; an engine handed in already
    # holds its weights and gets the artifact's tensors only (load_state_dict(strict=False) semantics, generate-images-sd.py:17-19)
    state = {k: v for k, v in pipe.unet.state_dict().items()} if engine is None else {}
    if uce_model_path is not None:
        from .artifact import load_artifact
        state.update(load_artifact(uce_model_path))
    cfg_pipe, latent = unet_config_of(pipe)
    if unet_config is None:
        unet_config = cfg_pipe
    own_engine = engine is None
    if own_engine:
        eng = UNetEngine(unet_config, batch=2 * num_images_per_prompt, H=latent, W=latent, device=device)
        eng.load_state_dict(state, strict=False)
        eng.finalize()
    else:
        eng = engine
        if state:
            eng.load_state_dict(state, strict=False)
    den = denoiser if denoiser is not None else Denoiser(eng, num_images_per_prompt)
    # VAE decode on the B200 decoder engine whenever the pipeline carries a VAE (UCE_VAE_ENGINE=0 keeps the pipeline's own decode)
    use_vae_engine = getattr(pipe, "vae", None) is not None and os.environ.get("UCE_VAE_ENGINE", "1") != "0"
    vae_eng = _vae_engine(pipe, num_images_per_prompt, latent, device) if use_vae_engine else None

    df = pd.read_csv(prompts_path)
    folder = f"{save_path}/{exp_name}"
    os.makedirs(folder, exist_ok=True)
    rank, world = _rank_world()
    for row in rows_for_rank(df, from_case, till_case, rank, world):
        case_number = row.case_number
        prompt, seed = str(row.prompt), row.evaluation_seed
        text, uncond = pipe.encode_prompt(prompt=prompt, device=device, num_images_per_prompt=num_images_per_prompt,
                                          do_classifier_free_guidance=True)[:2]
        ctx = torch.cat([uncond, text]).to(torch.float32)
        gen = torch.Generator().manual_seed(int(seed))
        lat = torch.randn((num_images_per_prompt, 4, latent, latent), generator=gen, dtype=torch_dtype)   # CPU generator, pipe dtype
        out = den.run(lat, ctx, steps=num_inference_steps, guidance_scale=guidance_scale, scheduler=scheduler)
        if vae_eng is not None:        # decoder on the B200 kernels (csrc/vae_engine.cu): uint8 [B, H, W, 3] leaves the device once
            images = vae_eng.decode(out.contiguous()).cpu()
        else:
            images = pipe.decode_latents_to_pil(out) if hasattr(pipe, "decode_latents_to_pil") else _decode(pipe, out, torch_dtype)
        for num, im in enumerate(images):
            save_png(f"{folder}/{case_number}_{num}.png", im)          # lossless, parallel deflate (csrc/png.cu); same file name as :46
    if own_engine:
        eng.close()
    if vae_eng is not None:
        vae_eng.close()


def _vae_engine(pipe, images, latent, device):
    """VAE decoder engine loaded from ``pipe.vae`` (opt-in, UCE_VAE_ENGINE=1): SD-1.x decoder layout, read off the parameter shapes."""
    from .vae import VAEDecoderEngine
    from .vae_spec import SD14_VAE
    state = {k: v for k, v in pipe.vae.state_dict().items()}
    cfg = dict(SD14_VAE)
    vcfg = getattr(pipe.vae, "config", None)
    for key in ("block_out_channels", "layers_per_block", "scaling_factor"):
        if vcfg is not None and getattr(vcfg, key, None) is not None:
            cfg[key] = tuple(getattr(vcfg, key)) if key == "block_out_channels" else getattr(vcfg, key)
    if vcfg is not None and getattr(vcfg, "norm_num_groups", None) is not None:
        cfg["norm_groups"] = vcfg.norm_num_groups
    ve = VAEDecoderEngine(cfg, batch=images, h=latent, w=latent, device=device)
    ve.load_state_dict(state, strict=False)
    ve.finalize()
    return ve


def _decode(pipe, latents, dtype):
    """vae.decode(latents / 0.18215) -> (x/2+0.5).clamp(0,1) -> uint8 NHWC -> PIL."""
    from PIL import Image
    sf = getattr(pipe.vae.config, "scaling_factor", 0.18215)
    img = pipe.vae.decode((latents / sf).to(dtype)).sample
    img = (img.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1).cpu().numpy()
    return [Image.fromarray((x * 255).round().astype("uint8")) for x in img]
