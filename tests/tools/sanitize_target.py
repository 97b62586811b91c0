"""Small end-to-end workload for compute-sanitizer (scripts/sanitize.sh runs it under memcheck / racecheck / synccheck / initcheck):
every CUDA kernel family of the library once, at sizes that finish in seconds under the tool — the low-latency factor, the general
blocked factor, both tcgen05 apply kernels and the SIMT twin, one tiny U-Net call (pair GEMM, implicit-GEMM conv, fused attention,
norms) and one tiny VAE decode.  Results are still checked against the oracle so a tool-induced timing change that exposes a race
shows up as a wrong answer too."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from oracle import uce_oracle as O
from uce_b200.solver import EditSolver
from uce_b200.synthetic import concept_rows, weights

what = set((sys.argv[1] if len(sys.argv) > 1 else "solver,unet,vae,clip").split(","))
if "solver" in what:
    for n_edit, n_pres, K, dims, impl, fimpl in [(10, 20, 256, [136, 64], 7, 0), (40, 20, 256, [136, 300, 72, 8], 7, 1), (10, 20, 256, [136, 64], 4, 0), (40, 20, 256, [136], 4, 1), (70, 10, 256, [200, 72], 5, 0), (10, 20, 256, [48], 1, 0)]:
        rows = concept_rows(n_edit + n_pres + n_edit, K, seed=n_edit)
        C, G = rows[: n_edit + n_pres], rows[n_edit + n_pres:]
        W = weights(dims, K, seed=4)
        s = EditSolver(K, C.shape[0], "cuda:0")
        s.set_apply_impl(impl); s.set_factor_impl(fimpl)
        out = s.edit(C.cuda(), G.cuda(), [1.0] * (n_edit + n_pres), n_edit, 0.5, [w.cuda() for w in W])
        exact = O.erase_exact_f64(W, C[:n_edit], G, C[n_edit:], 1.0, 1.0, 0.5)
        errs = [O.rel_fro(o.cpu(), e) for o, e in zip(out, exact)]
        print("solver", (n_edit, n_pres, K, dims, impl, fimpl), s.info(), errs, flush=True)
        assert max(errs) < 2e-5
        s.close()
if "solver" in what:
    # the host-buffer call (three streams, pinned arenas, block tables as kernel parameters)
    n_edit, n_pres, K, dims = 10, 20, 256, [136, 64, 200, 72]
    rows = concept_rows(n_edit + n_pres + n_edit, K, seed=2)
    C, G = rows[: n_edit + n_pres].pin_memory(), rows[n_edit + n_pres:].pin_memory()
    W = weights(dims, K, seed=5)
    s = EditSolver(K, C.shape[0], "cuda:0")
    _, a_in = EditSolver.host_arena(dims, K); _, a_out = EditSolver.host_arena(dims, K)
    for v, w in zip(a_in, W): v.copy_(w)
    s.edit_host(C, G, [1.0] * (n_edit + n_pres), n_edit, 0.5, a_in, a_out)
    exact = O.erase_exact_f64(W, C[:n_edit], G, C[n_edit:], 1.0, 1.0, 0.5)
    errs = [O.rel_fro(o, e) for o, e in zip(a_out, exact)]
    print("solver (host path)", errs, flush=True)
    assert max(errs) < 2e-5
    s.close()
if "clip" in what:
    import transformers
    from oracle import clip_zero_shot_oracle as Z
    from uce_b200.clip_zero_shot import ClipZeroShotEngine
    vocab = 300
    tc = dict(vocab_size=vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=20,
              hidden_act="quick_gelu", bos_token_id=vocab - 2, eos_token_id=vocab - 1, pad_token_id=vocab - 1)
    vc = dict(hidden_size=96, intermediate_size=192, num_hidden_layers=2, num_attention_heads=6, image_size=64, patch_size=16, hidden_act="quick_gelu")
    ccfg = transformers.CLIPConfig(text_config=tc, vision_config=vc, projection_dim=48)
    torch.manual_seed(0)
    P = transformers.CLIPModel(ccfg).eval().state_dict()
    eng = ClipZeroShotEngine(P, 6, 4, tokenizer=None, eos_token_id=vocab - 1, max_batch=2)
    g = torch.Generator().manual_seed(1)
    imgs = torch.randint(0, 256, (3, 96, 96, 3), generator=g, dtype=torch.uint8)
    ids = torch.full((2, 20), vocab - 1, dtype=torch.long); ids[:, 0] = vocab - 2; ids[0, 1:4] = torch.tensor([5, 6, 7]); ids[1, 1:3] = torch.tensor([8, 9])
    got = eng.logits_per_image(imgs.cuda(), input_ids=ids).cpu()
    ref = Z.logits_per_image(P, Z.preprocess(imgs, size=64), ids, 6, 4, vocab - 1)
    print("clip logits err", float((got - ref).abs().max()), flush=True)
    assert float((got - ref).abs().max()) < 0.05
if "unet" in what:
    from oracle import unet_oracle as U
    from uce_b200.unet import UNetEngine
    from uce_b200.unet_spec import tiny_config
    cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    P = U.random_weights(cfg, seed=3)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 4, 16, 16, generator=g); ctx = torch.randn(2, 77, cfg["cross_attention_dim"], generator=g)
    ref = U.unet_forward(P, x, 481.0, ctx, cfg)
    eng = UNetEngine(cfg, batch=2, H=16, W=16); eng.load_state_dict(P); eng.finalize()
    out = eng.forward(x.cuda(), 481.0, ctx.cuda()).cpu()
    rel = float((out.double() - ref.double()).norm() / ref.double().norm())
    print("unet rel", rel, flush=True)
    assert rel < 5e-2
    eng.close()
if "vae" in what:
    from oracle import vae_oracle as VO
    from uce_b200.vae import VAEDecoderEngine
    from uce_b200.vae_spec import tiny_vae_config
    cfg = tiny_vae_config(ch=(64, 128), groups=8)
    P = VO.random_weights(cfg, seed=3)
    eng = VAEDecoderEngine(cfg, batch=1, h=8, w=8); eng.load_state_dict(P); eng.finalize()
    lat = torch.randn((1, 4, 8, 8), generator=torch.Generator().manual_seed(4)) * cfg["scaling_factor"] * 3.0
    rgb = eng.decode(lat.cuda()).cpu()
    print("vae", tuple(rgb.shape), int(rgb.float().mean()), flush=True)
    eng.close()
print("sanitize target done", flush=True)
