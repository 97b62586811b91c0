#!/usr/bin/env python
"""Drop-in CLI for the reference's trainscripts/uce_sd_debias.py (flags and defaults :155-195)
with every re-solve of the iterative loop on the B200 kernels.  Under torchrun the generation / classification rounds are shared
between the ranks (debias.get_ratios); rank 0 writes the artifact."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FLAGS = [
    ("edit_concepts", dict(type=str, required=True, help="prompts corresponding to concepts to edit separated by ;")),
    ("debias_concepts", dict(type=str, default=None, help="Concepts to debias the edit concepts towards seperated by ;")),
    ("preserve_concepts", dict(type=str, default=None, help="Concepts to preserve seperated by ;")),
    ("model_id", dict(type=str, default="CompVis/stable-diffusion-v1-4", help="Model to run UCE on")),
    ("device", dict(type=str, default="cuda:0", help="cuda devices to train on")),
    ("edit_scale", dict(type=float, default=1, help="scale to edit concepts")),
    ("preserve_scale", dict(type=float, default=1, help="scale to preserve concepts")),
    ("lamb", dict(type=float, default=0.5, help="lambda regularization term for UCE")),
    ("save_dir", dict(type=str, default="../uce_models", help="where to save your uce model weights")),
    ("exp_name", dict(type=str, default=None, help="Use this to name your saved filename")),
    ("desired_ratios", dict(type=float, nargs="+", default=[0.5, 0.5], help="List of desired ratios for debiasing concepts(default: [0.5, 0.5])")),
    ("max_iterations", dict(type=int, default=30, help="Maximum number of iterations to debias using UCE(default: 30)")),
    ("max_diff", dict(type=float, default=0.05, help="Maximum difference allowed as error from desired ratio (default: 0.05)")),
    ("step_size", dict(type=float, default=0.1, help="Step size for v* updates(default: 0.1)")),
    ("num_images_per_prompt", dict(type=int, default=10, help="Number of images per prompt (default: 10)")),
    ("num_inference_steps", dict(type=int, default=20, help="Number of inference steps (default: 20)")),
    ("guidance_scale", dict(type=float, default=7.5, help="Guidance scale (default: 7.5)")),
]


def build_parser():
    p = argparse.ArgumentParser(prog="TrainUCE", description="UCE for erasing concepts in Stable Diffusion")
    for name, kw in FLAGS:
        p.add_argument("--" + name, **kw)
    # not in the reference: which denoise loop the generation rounds of get_ratios() use
    p.add_argument("--generator", type=str, default="pipe", choices=["pipe", "engine"],
                   help="'pipe': the diffusers pipeline, as the reference does; 'engine': the B200 U-Net engine (default: pipe)")
    # not in the reference: which classifier scores the generated images
    p.add_argument("--classifier", type=str, default="engine", choices=["engine", "pipeline"],
                   help="'engine': the B200 CLIP kernels on the pipeline's weights (fp32); 'pipeline': the transformers pipeline itself, "
                        "as the reference does (default: engine)")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    from uce_b200.concepts import split_concepts
    os.makedirs(args.save_dir, exist_ok=True)
    exp_name = args.exp_name if args.exp_name is not None else "uce_test"
    edit = split_concepts(args.edit_concepts)
    debias = split_concepts(args.debias_concepts)
    if len(debias) != len(args.desired_ratios):
        raise Exception("Error! The length of debias concepts and their corresponding desired ratios concepts do not match.")
    preserve = split_concepts(args.preserve_concepts)
    print(f"\n\nEditing: {edit}\n")
    print(f"Debias Across: {debias}\n")
    print(f"Preserving: {preserve}\n")
    import torch
    torch.set_grad_enabled(False)
    try:
        from diffusers import DiffusionPipeline
        from transformers import pipeline
    except ImportError as exc:
        raise SystemExit(f"diffusers/transformers are required to load '{args.model_id}': {exc}")
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # under torchrun the edit concepts of get_ratios() are dealt to the ranks (one all-reduce of the label counts per iteration);
        # every rank solves the same edit, rank 0 writes the artifact
        local = int(os.environ.get("LOCAL_RANK", "0"))
        args.device = f"cuda:{local}"
        torch.distributed.init_process_group("nccl", device_id=torch.device(args.device))
    pipe = DiffusionPipeline.from_pretrained(args.model_id, torch_dtype=torch.float32, safety_checker=None).to(args.device)
    pipe.set_progress_bar_config(disable=True)
    clip = pipeline(task="zero-shot-image-classification", model="openai/clip-vit-base-patch32", torch_dtype=torch.bfloat16, device=args.device)
    if args.classifier == "engine":
        from uce_b200.clip_zero_shot import ClipZeroShotEngine
        clip = ClipZeroShotEngine.from_pipeline(clip, device=args.device)
    from uce_b200.debias import UCE
    UCE(pipe, clip, edit, debias, preserve, args.edit_scale, args.preserve_scale, args.lamb, args.save_dir, exp_name,
        args.max_diff, args.step_size, args.num_images_per_prompt, args.num_inference_steps, args.guidance_scale,
        max_iterations=args.max_iterations, desired_ratios=args.desired_ratios, device=args.device,
        generator="engine" if args.generator == "engine" else None)


if __name__ == "__main__":
    main()
