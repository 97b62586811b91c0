"""CPU tests of the host logic: CLI surface (flags/defaults of the reference scripts), concept
parsing/expansion rules, projection selection, sharding layout, C-ABI symbols."""
import ctypes
import importlib.util
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_erase_cli_flags_and_defaults():
    m = _load("trainscripts/uce_sd_erase.py", "cli_erase")
    a = m.build_parser().parse_args(["--edit_concepts", "Van Gogh; Picasso", "--concept_type", "art"])
    assert (a.model_id, a.device, a.erase_scale, a.preserve_scale, a.lamb) == ("CompVis/stable-diffusion-v1-4", "cuda:0", 1, 1, 0.5)
    assert (a.expand_prompts, a.save_dir, a.exp_name, a.guide_concepts, a.preserve_concepts) == ("false", "../uce_models", None, None, None)
    edit, guide, pres = m.resolve(a)
    assert edit == ["Van Gogh", "Picasso"] and guide == ["art", "art"] and pres == []
    a = m.build_parser().parse_args(["--edit_concepts", "dog", "--concept_type", "object", "--expand_prompts", "true",
                                     "--preserve_concepts", "cat; bird"])
    edit, guide, pres = m.resolve(a)
    assert edit == ["dog", "image of dog", "photo of dog", "portrait of dog", "picture of dog", "painting of dog"]
    assert guide == ["", "image of ", "photo of ", "portrait of ", "picture of ", "painting of "]
    assert pres == ["cat", "bird"]
    a = m.build_parser().parse_args(["--edit_concepts", "a;b", "--guide_concepts", "x;y;z", "--concept_type", "art"])
    with pytest.raises(Exception):
        m.resolve(a)
    with pytest.raises(SystemExit):
        m.build_parser().parse_args(["--edit_concepts", "a", "--concept_type", "unsafe"])


def test_art_expansion_wording():
    from uce_b200.concepts import expand_prompts
    e, g = expand_prompts(["Monet"], ["art"], "art")
    assert e[1:] == ["painting by Monet", "art by Monet", "artwork by Monet", "picture by Monet", "style of Monet"]
    assert g[1:] == ["painting by art", "art by art", "artwork by art", "picture by art", "style of art"]


def test_debias_cli_flags_and_defaults():
    m = _load("trainscripts/uce_sd_debias.py", "cli_debias")
    a = m.build_parser().parse_args(["--edit_concepts", "doctor", "--debias_concepts", "male; female"])
    assert a.desired_ratios == [0.5, 0.5] and a.max_iterations == 30 and a.max_diff == 0.05 and a.step_size == 0.1
    assert (a.num_images_per_prompt, a.num_inference_steps, a.guidance_scale, a.edit_scale) == (10, 20, 7.5, 1)


def test_projection_selection_and_token_rule():
    from oracle.fake_pipe import FakePipe, layer_table
    from uce_b200.concepts import concept_row, select_projections
    pipe = FakePipe(layer_table("sd14"))
    sel = select_projections(pipe.unet)
    assert len(sel) == 32 and all("attn2" in n and n.endswith(("to_k", "to_v")) for n, _ in sel)
    assert sel[0][0] == "down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k"
    assert sel[-1][0] == "mid_block.attentions.0.transformer_blocks.0.attn2.to_v"
    assert sum(m.weight.shape[0] for _, m in sel) == 24960
    assert len(select_projections(FakePipe(layer_table("sdxl")).unet)) == 140
    emb = pipe.encode_prompt("a b c")[0]
    assert torch.equal(concept_row(pipe, "a b c", "cpu"), emb[0, 3])     # BOS + 3 words: last real token index 3
    assert torch.equal(concept_row(pipe, "", "cpu"), pipe.encode_prompt("")[0][0, 0])   # empty prompt keeps BOS


def test_build_rows_keeps_duplicates():
    from uce_b200.erase import build_rows
    rows = {"a": torch.ones(4), "b": 2 * torch.ones(4), "g": torch.zeros(4)}
    C, G, s, ne = build_rows(rows, ["a", "a"], ["g", "g"], ["b", "b", "b"], 2.0, 3.0)
    assert C.shape == (5, 4) and G.shape == (2, 4) and ne == 2 and s == [2.0, 2.0, 3.0, 3.0, 3.0]


def test_pack_layout_roundtrip():
    from uce_b200.sharding import pack_layout, shard_layers
    dims = [320, 640, 1280, 320, 640]
    per, where = pack_layout(dims, 8, 2)
    assert shard_layers(5, 2, 0) == [0, 2, 4] and shard_layers(5, 2, 1) == [1, 3]
    assert per == (320 + 1280 + 640) * 8
    assert where[2] == (0, 320 * 8) and where[3] == (1, 640 * 8)


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/uce_b200.h declares."""
    from uce_b200 import _native
    hdr = open(os.path.join(ROOT, "include", "uce_b200.h")).read()
    declared = set(re.findall(r"\b(uce_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    assert _native.lib().uce_abi_version() == 1


def test_sd_unet_c_abi_exports_every_declared_symbol():
    from uce_b200 import _native, unet
    hdr = open(os.path.join(ROOT, "include", "sd_unet_b200.h")).read()
    declared = set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(unet.SD_SIGNATURES), declared ^ set(unet.SD_SIGNATURES)


def test_sd_vae_c_abi_exports_every_declared_symbol():
    from uce_b200 import _native, vae
    hdr = open(os.path.join(ROOT, "include", "sd_vae_b200.h")).read()
    declared = set(re.findall(r"\b(sd_vae_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(vae.VAE_SIGNATURES), declared ^ set(vae.VAE_SIGNATURES)
    # the ctypes mirror of sd_vae_config has the header's layout (4 + 4 + 4 + 16 + 4 + 4 + 4 bytes, no padding)
    assert ctypes.sizeof(vae.SDVaeConfig) == 40
    cs = vae._cfg_struct(vae.SD14_VAE)
    assert list(cs.block_out_channels) == [128, 256, 512, 512] and cs.n_levels == 4 and abs(cs.scaling_factor - 0.18215) < 1e-7


def test_sd_unet_engine_inventory_is_the_diffusers_unet_layout():
    """Same for the U-Net engine: the native inventory equals the spec the CPU oracle is built on — for SD-1.4 the 859 520 964
    parameters of the published checkpoint under their diffusers names (of which the 32 edited attn2.to_k/to_v matrices are 19 169 280)."""
    from uce_b200 import unet
    from uce_b200.unet_spec import SD14, param_shapes, tiny_config
    for cfg in (SD14, tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)):
        inv = unet.engine_inventory(cfg)
        want = {k: tuple(v) for k, v in param_shapes(cfg).items()}
        assert inv == want, sorted(set(inv) ^ set(want))[:5]
    inv = unet.engine_inventory(SD14)
    count = lambda shp: int(torch.tensor(shp).prod())
    assert sum(count(s) for s in inv.values()) == 859_520_964
    edited = [k for k in inv if "attn2" in k and k.endswith(("to_k.weight", "to_v.weight"))]
    assert len(edited) == 32 and sum(count(inv[k]) for k in edited) == 19_169_280


def test_sd_vae_engine_inventory_is_the_diffusers_decoder_layout():
    """The parameter names / shapes the native decoder engine expects (host-only query) are exactly the spec the CPU oracle is built
    on — the SD-1.x `pipe.vae` decoder layout with its 49 490 179 + 20 parameters — so a real checkpoint loads by name."""
    from uce_b200 import vae
    from uce_b200.vae_spec import SD14_VAE, SD14_VAE_DECODER_PARAMS, decoder_param_shapes, tiny_vae_config
    for cfg in (SD14_VAE, tiny_vae_config(ch=(64, 128), groups=8), tiny_vae_config(ch=(64, 64, 128), groups=8)):
        inv = vae.engine_inventory(cfg)
        want = {k: tuple(v) for k, v in decoder_param_shapes(cfg).items()}
        assert inv == want, set(inv) ^ set(want)
    n = 0
    for shp in vae.engine_inventory(SD14_VAE).values():
        k = 1
        for d in shp:
            k *= d
        n += k
    assert n == SD14_VAE_DECODER_PARAMS + 20


def test_no_gpu_fails_loudly():
    from uce_b200.solver import EditSolver
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        EditSolver(768, 16, "cuda:0")
    with pytest.raises(RuntimeError):
        EditSolver(768, 16, "cpu")
    from uce_b200.unet import UNetEngine
    with pytest.raises(RuntimeError):
        UNetEngine(device="cuda:0")
    from uce_b200.vae import VAEDecoderEngine
    with pytest.raises(RuntimeError):
        VAEDecoderEngine(device="cuda:0")


def _plan(sm_count, dims):
    import ctypes as C
    from uce_b200 import _native
    n = len(dims)
    d = (C.c_int * n)(*dims)
    rows, first = (C.c_int * n)(), (C.c_int * n)()
    total = _native.lib().uce_plan_row_blocks(sm_count, d, n, rows, first)
    return total, list(rows), list(first)


def test_row_block_plan_covers_every_row_once(monkeypatch):
    """Host-side tile plan of the two-block tcgen05 apply (apply_tc3.cu): every row of every projection belongs to exactly one
    CTA, block heights are multiples of 8 in [8,128], and an SD-1.4 edit (uce_sd_erase.py:15-22: 10x320 + 10x640 + 12x1280
    rows) is ONE balanced wave on a 148-SM B200."""
    from uce_b200.synthetic import SD14_DIMS
    monkeypatch.delenv("UCE_TC3_BLOCK_ROWS", raising=False)
    cases = [(148, list(SD14_DIMS)), (148, [320] * 4), (148, [8]), (148, [1000, 24, 136]), (148, [1280] * 96), (132, list(SD14_DIMS)),
             (148, [300] * 40)]
    for sms, dims in cases:
        total, rows, first = _plan(sms, dims)
        assert total > 0
        cta = 0
        for dl, h, f in zip(dims, rows, first):
            assert 8 <= h <= 128 and h % 8 == 0, (dims, rows)
            assert f == cta
            n_cta = -(-dl // (2 * h))
            assert (n_cta - 1) * 2 * h < dl <= n_cta * 2 * h      # the last CTA is not empty, all rows covered
            cta += n_cta
        assert cta == total
    total, rows, _ = _plan(148, list(SD14_DIMS))
    assert total <= 148 and min(rows) >= 64, (total, rows)            # one wave, no short blocks
    per_cta = [2 * h for h in rows]
    assert max(per_cta) <= 1.25 * (sum(SD14_DIMS) / total), per_cta  # balanced: no CTA far above the mean row count
    monkeypatch.setenv("UCE_TC3_BLOCK_ROWS", "40")
    total, rows, _ = _plan(148, [300] * 40)
    assert rows == [40] * 40 and total == 40 * 4
    import ctypes as C
    from uce_b200 import _native
    assert _native.lib().uce_plan_row_blocks(148, None, 1, None, None) == _native.UCE_E_ARG


def test_generation_rows_are_dealt_round_robin_after_the_case_filter():
    """generate-images-sd.py:29-35: inclusive from_case / till_case filter; our ranks then take every world-th surviving row,
    so the union over ranks is exactly the reference's row list, in order, without overlap."""
    import pandas as pd
    from uce_b200.generate import rows_for_rank
    df = pd.DataFrame({"prompt": [f"p{i}" for i in range(11)], "evaluation_seed": list(range(100, 111)), "case_number": [0, 1, 2, 3, 5, 8, 9, 10, 12, 40, 41]})
    ref = [r.case_number for _, r in df.iterrows() if 1 <= r.case_number <= 40]            # the reference's loop
    assert [r.case_number for r in rows_for_rank(df, 1, 40, 0, 1)] == ref
    for world in (2, 3, 8):
        parts = [[r.case_number for r in rows_for_rank(df, 1, 40, rk, world)] for rk in range(world)]
        assert sorted(sum(parts, [])) == ref and sum(len(p) for p in parts) == len(ref)
        assert all(parts[rk] == ref[rk::world] for rk in range(world))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    assert rows_for_rank(df, 100, 200, 0, 2) == []


def test_clip_text_c_abi_exports_every_declared_symbol():
    from uce_b200 import _native, clip_text
    hdr = open(os.path.join(ROOT, "include", "clip_text_b200.h")).read()
    declared = set(re.findall(r"\b(clipt_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(clip_text.SIGNATURES), declared ^ set(clip_text.SIGNATURES)


def test_host_arena_is_one_allocation_with_back_to_back_views():
    """EditSolver.host_arena: the layout uce_edit_host_f32 recognises (projection l + 1 starts where l ends), so a pipeline group
    travels as one copy per direction."""
    from uce_b200.solver import EditSolver
    dims, K = [320, 640, 8, 1280], 96
    buf, views = EditSolver.host_arena(dims, K, pin=False)
    assert buf.numel() == sum(dims) * K and [tuple(v.shape) for v in views] == [(d, K) for d in dims]
    for a, b, d in zip(views, views[1:], dims):
        assert b.data_ptr() == a.data_ptr() + d * K * 4 and a.is_contiguous()
    views[2].fill_(3.0)
    assert float(buf[(320 + 640) * K]) == 3.0 and float(buf[(320 + 640 + 8) * K - 1]) == 3.0


def test_clip_vision_c_abi_exports_every_declared_symbol():
    from uce_b200 import _native, clip_zero_shot
    hdr = open(os.path.join(ROOT, "include", "clip_vision_b200.h")).read()
    declared = set(re.findall(r"\b(clipv_[a-z0-9_]+)\s*\(", hdr))
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert declared == set(clip_zero_shot.SIGNATURES), declared ^ set(clip_zero_shot.SIGNATURES)


def test_clip_engines_refuse_to_run_without_cuda():
    """No CPU fallback: on a box without a GPU the engines raise instead of computing somewhere else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from uce_b200.clip_text import ClipTextEngine
    from uce_b200.clip_zero_shot import ClipVisionEngine
    with pytest.raises(RuntimeError):
        ClipTextEngine({}, 4)
    with pytest.raises(RuntimeError):
        ClipVisionEngine({}, 4)


def test_public_headers_are_plain_c(tmp_path):
    """The drop-in boundary is a C ABI: both headers must compile as C11 on their own (no torch / CUDA types in the signatures)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    src = tmp_path / "hdr.c"
    src.write_text(f'#include "{ROOT}/include/uce_b200.h"\n#include "{ROOT}/include/sd_unet_b200.h"\n#include "{ROOT}/include/sd_vae_b200.h"\n#include "{ROOT}/include/clip_text_b200.h"\n#include "{ROOT}/include/clip_vision_b200.h"\nint main(void) {{ return 0; }}\n')
    r = subprocess.run([gcc, "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_python_function_seam_has_the_reference_signatures():
    """SURVEY.md 8b 'Python fn' seam: tests/golden/signatures.json holds the parameter names, order and defaults of the reference's
    UCE() / get_ratios() / generate_images() (read with inspect from the real modules by oracle/make_signature_golden.py).  Our mirrors
    take the same parameters in the same positions with the same defaults; anything we add comes after them and is optional."""
    import inspect
    import json
    from uce_b200 import debias, erase, generate
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "signatures.json")))
    ours = {"erase.UCE": erase.UCE, "debias.UCE": debias.UCE, "debias.get_ratios": debias.get_ratios, "generate.generate_images": generate.generate_images}
    # defaults the mirrors widen on purpose: get_ratios gives the reference's required parameters defaults so that the weight tensors can
    # also be passed by the keyword alias uce_weights (a reference-style call always passes them)
    widened = {"debias.get_ratios": {"uce_modules", "edit_concepts", "debias_concepts", "desired_ratios", "max_diff"}}
    for name, ref in gold.items():
        params = list(inspect.signature(ours[name]).parameters.values())
        assert [p.name for p in params[: len(ref)]] == [n for n, _ in ref], name
        for p, (n, d) in zip(params, ref):
            if d is None:
                assert p.default is inspect._empty or n in widened.get(name, ()), (name, n)
            else:
                assert repr(p.default) == d, (name, n, repr(p.default), d)
        for p in params[len(ref):]:
            assert p.default is not inspect._empty, (name, p.name)            # our extras are optional


def _build_c_example(tmp_path):
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    from uce_b200 import _native
    libdir = os.path.dirname(_native.LIB_PATH)
    exe = str(tmp_path / "edit_host")
    r = subprocess.run([gcc, "-std=c11", "-O2", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "edit_host.c"),
                        "-L" + libdir, "-luce_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_plain_c_consumer_links_and_fails_loudly_without_a_gpu(tmp_path):
    """examples/edit_host.c uses the boundary the way a non-Python host would: it must compile as C11 against include/uce_b200.h, link
    against libuce_b200.so alone, and — the library has no CPU path — stop with UCE_E_NO_DEVICE (exit 3) on a box without a B200."""
    import subprocess
    exe = _build_c_example(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the gpu-marked twin runs the edit")
    r = subprocess.run([exe, str(tmp_path / "out.safetensors")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert "no B200" in r.stderr and not os.path.exists(tmp_path / "out.safetensors")


@pytest.mark.gpu
def test_plain_c_consumer_edits_on_the_gpu(tmp_path):
    """The same program on a B200: edit from host buffers, normal-equation residual checked in C, artifact readable by safetensors."""
    import subprocess
    from safetensors.torch import load_file
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe, str(tmp_path / "out.safetensors")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    got = load_file(str(tmp_path / "out.safetensors"))
    assert sorted(v.shape for v in got.values()) == [torch.Size([24, 64]), torch.Size([40, 64])] and all(v.dtype == torch.float32 for v in got.values())


def test_zero_shot_host_helpers():
    """Host-side pieces of the zero-shot classifier binding that need no GPU: where the text tower pools (argmax of the ids for the
    original checkpoints whose config says eos id 2, first end-of-text token otherwise — transformers' CLIPTextTransformer.forward) and
    the image batching (PIL-free inputs: uint8 arrays, one 4-D array, a tensor)."""
    import numpy as np
    import torch
    from uce_b200.clip_zero_shot import HYPOTHESIS_TEMPLATE, _to_u8_batch, eos_index
    ids = torch.tensor([[49406, 320, 1125, 49407, 49407, 49407], [49406, 49407, 0, 0, 0, 0], [49406, 5, 6, 7, 8, 49407]])
    assert eos_index(ids, 2).tolist() == [3, 1, 5] and eos_index(ids, None).tolist() == [3, 1, 5]          # legacy: argmax
    ids2 = torch.tensor([[300, 5, 9, 301, 301], [300, 301, 7, 7, 7]])
    assert eos_index(ids2, 301).tolist() == [3, 1]                                                           # first eos token
    a = np.zeros((4, 4, 3), dtype=np.uint8); b = np.full((4, 4, 3), 7, dtype=np.uint8)
    t = _to_u8_batch([a, b], "cpu")
    assert t.dtype == torch.uint8 and tuple(t.shape) == (2, 4, 4, 3) and int(t[1].max()) == 7
    assert tuple(_to_u8_batch(np.stack([a, b]), "cpu").shape) == (2, 4, 4, 3)
    assert tuple(_to_u8_batch(a, "cpu").shape) == (1, 4, 4, 3)
    assert tuple(_to_u8_batch(torch.zeros(3, 8, 8, 3, dtype=torch.uint8), "cpu").shape) == (3, 8, 8, 3)
    with pytest.raises(ValueError):
        _to_u8_batch([np.zeros((4, 4, 3), dtype=np.float32)], "cpu")
    assert HYPOTHESIS_TEMPLATE.format("male") == "This is a photo of male."
