"""BASELINE cfg5 end to end: the row loop of evalscripts/generate-images-sd.py (generate_images) over a vangogh_prompts.csv-shaped
prompt file — 50 rows (case_number, prompt, evaluation_seed) — at 50 denoise steps and 8 images per prompt, prompts dealt to the
ranks under torchrun, every stage on the B200 kernels: CLIP text tower (clip_text.cu) -> U-Net denoise loop (unet_engine.cu) ->
VAE decoder (vae_engine.cu) -> PNG files (png.cu).

There is no network for checkpoints: the pipeline object is synthetic — seeded random weights of the SD-1.4 architectures (U-Net 860 M
parameters, VAE decoder 49 M, CLIP ViT-L/14 text tower 123 M) and a whitespace tokenizer — so the images are noise, the work is not.

    python scripts/cfg5_generate.py [--rows 12] [--steps 50] [--images 8] [--scheduler ddim] [--out gpurun_out/cfg5.json]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/cfg5_generate.py ...

Rank 0 prints one JSON line: images/s and image-steps/s of the whole job (max wall time over ranks), per-GPU denoise steps/s."""
import argparse
import json
import os
import sys
import tempfile
import time
import zlib

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _Holder:
    def __init__(self, state, config=None):
        self._state, self.config = state, config

    def state_dict(self):
        return self._state


class SyntheticPipe:
    """The attributes generate_images() uses of a StableDiffusionPipeline."""

    latent_size = 64

    def __init__(self, device):
        import transformers
        from uce_b200.clip_text import ClipTextEngine
        from uce_b200.synthetic import unet_random_state
        from uce_b200.unet_spec import SD14
        from uce_b200.vae_spec import SD14_VAE
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from vae_probe import random_weights
        self.unet = _Holder(unet_random_state(SD14, seed=0))
        self.vae = _Holder(random_weights(SD14_VAE, seed=0))
        cfg = transformers.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                                          max_position_embeddings=77, hidden_act="quick_gelu", bos_token_id=49406, eos_token_id=49407, pad_token_id=49407)
        torch.manual_seed(0)
        self.text = ClipTextEngine(transformers.CLIPTextModel(cfg).state_dict(), 12, device=device, max_batch=2)

    def _ids(self, prompt):
        words = [zlib.crc32(w.encode()) % 49000 for w in prompt.lower().split()][:75]
        return torch.tensor([[49406] + words + [49407] * (76 - len(words))], dtype=torch.int32)

    def encode_prompt(self, prompt, device, num_images_per_prompt, do_classifier_free_guidance=True, **_):
        h = self.text.encode(torch.cat([self._ids(prompt), self._ids("")]))                 # [2, 77, 768]: prompt, empty prompt
        n = num_images_per_prompt
        return h[0:1].expand(n, -1, -1).contiguous(), h[1:2].expand(n, -1, -1).contiguous()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12, help="CSV rows (prompts) in the job; vangogh_prompts.csv has 50")
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--scheduler", default="ddim")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = f"cuda:{local}"
    torch.cuda.set_device(local)
    torch.set_grad_enabled(False)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device(dev))
    import pandas as pd
    from uce_b200.generate import generate_images
    pipe = SyntheticPipe(dev)
    tmp = tempfile.mkdtemp(prefix="cfg5_")
    csv = os.path.join(tmp, "prompts.csv")
    pd.DataFrame({"case_number": range(a.rows), "prompt": [f"painting number {i} of a wheatfield with cypresses by vincent van gogh" for i in range(a.rows)],
                  "evaluation_seed": [1000 + 7 * i for i in range(a.rows)], "artist": ["Vincent van Gogh"] * a.rows}).to_csv(csv)
    from uce_b200.unet import UNetEngine
    from uce_b200.unet_spec import SD14
    t_build = time.perf_counter()
    eng = UNetEngine(SD14, batch=2 * a.images, H=64, W=64, device=dev)      # built once per process, as a serving process would
    eng.load_state_dict(pipe.unet.state_dict(), strict=False)
    eng.finalize()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build
    kw = dict(model_id="synthetic-sd14", uce_model_path=None, prompts_path=csv, save_path=tmp, device=dev, guidance_scale=7.5,
              num_inference_steps=a.steps, num_images_per_prompt=a.images, pipe=pipe, scheduler=a.scheduler, engine=eng)
    generate_images(exp_name="warm", till_case=world - 1, **kw)           # one row per rank: graph capture, first launches
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    generate_images(exp_name="run", **kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    mine = len(range(rank, a.rows, world))
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    files = len([f for f in os.listdir(os.path.join(tmp, "run")) if f.endswith(".png")])
    if world > 1:
        torch.distributed.barrier()
    if rank == 0:
        wall = float(t.item())
        line = {"workload": f"cfg5: {a.rows} prompts x {a.images} images, {a.steps}-step {a.scheduler}, 512x512, synthetic SD-1.4 weights",
                "n_gpus": world, "wall_s": wall, "images": a.rows * a.images, "images_per_s": a.rows * a.images / wall,
                "image_steps_per_s": a.rows * a.images * a.steps / wall, "rows_this_rank": mine,
                "unet_steps_per_s_per_gpu": mine * a.steps / dt, "png_files_rank0_sees": files,
                "unet_engine_build_s": t_build,
                "includes": "VAE engine construction, text tower, denoise loop, VAE decode, PNG encode + write; the U-Net engine itself is "
                            "built once per process (unet_engine_build_s)"}
        print(json.dumps(line), flush=True)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "w") as f:
                f.write(json.dumps(line) + "\n")
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
