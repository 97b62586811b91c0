#!/usr/bin/env python
"""Drop-in CLI for the reference's trainscripts/uce_sd_erase.py (same flags and defaults,
:97-112; same prints, :193-195; same artifact, :85-88) with the edit solved on the B200 kernels.

Multi-GPU (not in the reference): launched under torchrun (`python -m torch.distributed.run --nproc-per-node N ...`) the projections
are sharded over the N ranks and gathered once at the end; there is no flag for it, the torchrun environment selects it.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

FLAGS = [
    # name, type, default, required, choices, help
    ("edit_concepts", str, None, True, None, "prompts corresponding to concepts to erase separated by ;"),
    ("guide_concepts", str, None, False, None, "Concepts to guide the erased concepts towards seperated by ;"),
    ("preserve_concepts", str, None, False, None, "Concepts to preserve seperated by ;"),
    ("concept_type", str, None, True, ["art", "object"], "type of concept being erased"),
    ("model_id", str, "CompVis/stable-diffusion-v1-4", False, None, "Model to run UCE on"),
    ("device", str, "cuda:0", False, None, "cuda devices to train on"),
    ("erase_scale", float, 1, False, None, "scale to erase concepts"),
    ("preserve_scale", float, 1, False, None, "scale to preserve concepts"),
    ("lamb", float, 0.5, False, None, "lambda regularization term for UCE"),
    ("expand_prompts", str, "false", False, ["true", "false"], "do you wish to expand your prompts?"),
    ("save_dir", str, "../uce_models", False, None, "where to save your uce model weights"),
    ("exp_name", str, None, False, None, "Use this to name your saved filename"),
]


def build_parser():
    p = argparse.ArgumentParser(prog="TrainUCE", description="UCE for erasing concepts in Stable Diffusion")
    for name, typ, default, required, choices, text in FLAGS:
        p.add_argument("--" + name, type=typ, default=default, required=required, choices=choices, help=text)
    return p


def resolve(args):
    """CLI strings -> (edit, guide, preserve) lists following uce_sd_erase.py:134-190."""
    from uce_b200.concepts import expand_prompts, resolve_guides, split_concepts
    edit = split_concepts(args.edit_concepts)
    guide = resolve_guides(edit, args.guide_concepts, args.concept_type)
    preserve = split_concepts(args.preserve_concepts)
    if args.expand_prompts == "true":
        edit, guide = expand_prompts(edit, guide, args.concept_type)
    return edit, guide, preserve


def load_pipeline(model_id, device, with_vae=False):
    try:
        import torch
        from diffusers import DiffusionPipeline
    except ImportError as exc:   # the text encoder / U-Net container comes from diffusers, exactly as in the reference (:197-200)
        raise SystemExit(f"diffusers is required to load '{model_id}': {exc}")
    kw = dict(torch_dtype=torch.float32, safety_checker=None)
    if not with_vae:
        kw["vae"] = None
    return DiffusionPipeline.from_pretrained(model_id, **kw).to(device)


def main(argv=None):
    args = build_parser().parse_args(argv)
    os.makedirs(args.save_dir, exist_ok=True)
    exp_name = args.exp_name if args.exp_name is not None else "uce_test"
    edit, guide, preserve = resolve(args)
    print(f"\n\nErasing: {edit}\n")
    print(f"Guiding: {guide}\n")
    print(f"Preserving: {preserve}\n")
    import torch
    torch.set_grad_enabled(False)
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        args.device = f"cuda:{local}"
        torch.distributed.init_process_group("nccl", device_id=torch.device(args.device))
    pipe = load_pipeline(args.model_id, args.device)
    from uce_b200.erase import UCE
    UCE(pipe, edit, guide, preserve, args.erase_scale, args.preserve_scale, args.lamb, args.save_dir, exp_name, device=args.device)


if __name__ == "__main__":
    main()
