"""Pins oracle/clip_zero_shot_oracle.py (the debias edit's zero-shot image classifier, trainscripts/uce_sd_debias.py:27,245-250;
groundwork for SURVEY.md 8(f) rank 3) against the transformers CLIPModel the reference's pipeline wraps.  CPU only."""
import pytest
import torch

from oracle import clip_zero_shot_oracle as Z

transformers = pytest.importorskip("transformers")


def _model():
    vocab = 300
    tc = dict(vocab_size=vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=20,
              hidden_act="quick_gelu", bos_token_id=vocab - 2, eos_token_id=vocab - 1, pad_token_id=vocab - 1)
    vc = dict(hidden_size=96, intermediate_size=192, num_hidden_layers=3, num_attention_heads=6, image_size=64, patch_size=16, hidden_act="quick_gelu")
    cfg = transformers.CLIPConfig(text_config=tc, vision_config=vc, projection_dim=48)
    torch.manual_seed(0)
    return transformers.CLIPModel(cfg).eval(), cfg


def _ids(cfg, lengths, T=20):
    g = torch.Generator().manual_seed(3)
    t = cfg.text_config
    ids = torch.full((len(lengths), T), t.eos_token_id, dtype=torch.long)
    for b, n in enumerate(lengths):
        ids[b, 0] = t.bos_token_id
        ids[b, 1:1 + n] = torch.randint(0, t.vocab_size - 2, (n,), generator=g)
    return ids


def test_logits_match_transformers_clip_model():
    model, cfg = _model()
    g = torch.Generator().manual_seed(1)
    pixels = torch.randn(5, 3, 64, 64, generator=g)
    ids = _ids(cfg, [6, 7, 3])
    with torch.no_grad():
        ref = model(input_ids=ids, pixel_values=pixels)
    P = model.state_dict()
    got = Z.logits_per_image(P, pixels, ids, cfg.vision_config.num_attention_heads, cfg.text_config.num_attention_heads, cfg.text_config.eos_token_id)
    assert got.shape == ref.logits_per_image.shape == (5, 3)
    assert float((got - ref.logits_per_image).abs().max()) < 1e-4
    im = Z.image_features(P, pixels, cfg.vision_config.num_attention_heads)
    assert float((im - ref.image_embeds * im.norm(dim=-1, keepdim=True)).abs().max()) < 1e-4      # image_embeds are the normalised ones
    # what get_ratios keeps: the top-1 label per image
    labels = ["male", "female", "other"]
    res = Z.classify(got, labels)
    want = ref.logits_per_image.softmax(dim=-1)
    for b, r in enumerate(res):
        assert r[0]["label"] == labels[int(want[b].argmax())]
        assert abs(sum(x["score"] for x in r) - 1.0) < 1e-5 and [x["score"] for x in r] == sorted((x["score"] for x in r), reverse=True)


def test_preprocess_matches_clip_image_processor():
    proc = transformers.CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    g = torch.Generator().manual_seed(2)
    imgs = torch.randint(0, 256, (2, 128, 128, 3), generator=g, dtype=torch.uint8)
    # smooth the noise so resampling differences stay at the rounding level
    sm = torch.nn.functional.avg_pool2d(imgs.permute(0, 3, 1, 2).float(), 9, stride=1, padding=4).round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    ref = torch.tensor(proc(images=[x.numpy() for x in sm], return_tensors="np")["pixel_values"])
    got = Z.preprocess(sm, size=32)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 3.0 / 255 / 0.26 + 1e-6        # <= 3 uint8 levels after normalisation (PIL vs torch bicubic antialiasing)
    same = Z.preprocess(sm[:, :32, :32].contiguous(), size=32)            # no resize: exact
    ref_same = torch.tensor(proc(images=[x.numpy() for x in sm[:, :32, :32]], return_tensors="np")["pixel_values"])
    assert float((same - ref_same).abs().max()) < 1e-6
