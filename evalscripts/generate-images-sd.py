#!/usr/bin/env python
"""Drop-in CLI for the reference's evalscripts/generate-images-sd.py (flags and defaults :52-62) with the denoise
loop on the B200 engine.  Extra optional flag: --scheduler {pndm,ddim} (the reference uses the pipeline default,
PNDM for SD-1.4).  Under torchrun the CSV rows are sharded over the ranks."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FLAGS = [
    ("model_id", dict(type=str, default="CompVis/stable-diffusion-v1-4", help="hf repo id for the model you want to test")),
    ("uce_model_path", dict(type=str, default=None, help="path for uce model")),
    ("prompts_path", dict(type=str, required=True, help="path to csv file with prompts")),
    ("save_path", dict(type=str, default="../uce_results/", help="folder where to save images")),
    ("device", dict(type=str, default="cuda:0", help="cuda device to run on")),
    ("exp_name", dict(type=str, default="test_images", help="foldername to save the results")),
    ("guidance_scale", dict(type=float, default=7.5, help="guidance to run eval")),
    ("till_case", dict(type=int, default=1000000, help="continue generating from case_number")),
    ("from_case", dict(type=int, default=0, help="continue generating from case_number")),
    ("num_images_per_prompt", dict(type=int, default=1, help="number of samples per prompt")),
    ("num_inference_steps", dict(type=int, default=50, help="ddim steps of inference used to train")),
    ("scheduler", dict(type=str, default="pndm", choices=["pndm", "ddim"], help="(extension) scheduler of the denoise loop")),
]


def build_parser():
    p = argparse.ArgumentParser(prog="generateImages", description="Generate Images using Diffusers Code")
    for name, kw in FLAGS:
        p.add_argument("--" + name, **kw)
    return p


def main(argv=None):
    a = build_parser().parse_args(argv)
    import torch
    torch.set_grad_enabled(False)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        a.device = f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}"
    from uce_b200.generate import generate_images
    generate_images(model_id=a.model_id, uce_model_path=a.uce_model_path, prompts_path=a.prompts_path, save_path=a.save_path,
                    exp_name=a.exp_name, device=a.device, torch_dtype=torch.bfloat16, guidance_scale=a.guidance_scale,
                    num_inference_steps=a.num_inference_steps, num_images_per_prompt=a.num_images_per_prompt,
                    from_case=a.from_case, till_case=a.till_case, scheduler=a.scheduler)


if __name__ == "__main__":
    main()
