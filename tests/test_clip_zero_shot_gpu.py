"""The B200 zero-shot image classifier (include/clip_vision_b200.h) against the oracle pinned to transformers
(oracle/clip_zero_shot_oracle.py) and against transformers' CLIPModel itself: preprocessing, image features, logits, and the
``clip(images, candidate_labels=...)`` call debias.get_ratios makes (trainscripts/uce_sd_debias.py:27)."""
import pytest
import torch

from oracle import clip_zero_shot_oracle as Z

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def _model(vision=None, text=None, proj=48, seed=0):
    vocab = 300
    tc = dict(vocab_size=vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, max_position_embeddings=20,
              hidden_act="quick_gelu", bos_token_id=vocab - 2, eos_token_id=vocab - 1, pad_token_id=vocab - 1)
    vc = dict(hidden_size=96, intermediate_size=192, num_hidden_layers=3, num_attention_heads=6, image_size=64, patch_size=16, hidden_act="quick_gelu")
    tc.update(text or {}); vc.update(vision or {})
    cfg = transformers.CLIPConfig(text_config=tc, vision_config=vc, projection_dim=proj)
    torch.manual_seed(seed)
    model = transformers.CLIPModel(cfg).eval()
    with torch.no_grad():                                  # default init leaves the class token / positions tiny: make every term count
        model.vision_model.embeddings.class_embedding.normal_(0, 0.5)
        model.vision_model.embeddings.position_embedding.weight.normal_(0, 0.3)
    return model, cfg


def _ids(cfg, lengths, T=20, seed=3):
    g = torch.Generator().manual_seed(seed)
    t = cfg.text_config
    ids = torch.full((len(lengths), T), t.eos_token_id, dtype=torch.long)
    for b, n in enumerate(lengths):
        ids[b, 0] = t.bos_token_id
        ids[b, 1:1 + n] = torch.randint(0, t.vocab_size - 2, (n,), generator=g)
    return ids


def _smooth_images(n, H, seed):
    g = torch.Generator().manual_seed(seed)
    imgs = torch.randint(0, 256, (n, H, H, 3), generator=g, dtype=torch.uint8)
    x = torch.nn.functional.avg_pool2d(imgs.permute(0, 3, 1, 2).float(), 9, stride=1, padding=4)
    x = (x - x.mean()) * 3 + 128                          # stretch the contrast back after the blur
    return x.round().clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("vision", [None, dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=2, num_attention_heads=12, image_size=224, patch_size=32)])
def test_image_features_and_logits_match_oracle_and_transformers(vision):
    """A reduced tower and the ViT-B/32 geometry of openai/clip-vit-base-patch32 (768 wide, 12 heads, 7x7 patches of 32, 2 of its 12 layers)."""
    from uce_b200.clip_zero_shot import ClipZeroShotEngine
    model, cfg = _model(vision=vision)
    S = cfg.vision_config.image_size
    g = torch.Generator().manual_seed(1)
    pixels = torch.randn(5, 3, S, S, generator=g)
    ids = _ids(cfg, [6, 7, 3])
    P = model.state_dict()
    vh, th, eos = cfg.vision_config.num_attention_heads, cfg.text_config.num_attention_heads, cfg.text_config.eos_token_id
    eng = ClipZeroShotEngine(P, vh, th, tokenizer=None, eos_token_id=eos, max_batch=2)        # 5 images through a 2-image engine
    feats = eng.vision.image_features(pixels).cpu()
    ref_feats = Z.image_features(P, pixels, vh)
    scale = max(1.0, float(ref_feats.abs().max()))
    assert float((feats - ref_feats).abs().max()) <= 3e-5 * scale, float((feats - ref_feats).abs().max())
    assert eng.vision.launch_count() == 7 + 7 * cfg.vision_config.num_hidden_layers
    rows = eng.rows_at_eos(ids)
    logits = eng.vision.logits(eng.vision.image_features(pixels), rows).cpu()
    ref = Z.logits_per_image(P, pixels, ids, vh, th, eos)
    with torch.no_grad():
        lib = model(input_ids=ids, pixel_values=pixels).logits_per_image
    assert logits.shape == ref.shape == (5, 3)
    assert float((logits - ref).abs().max()) <= 2e-4 and float((logits - lib).abs().max()) <= 2e-4      # logits are O(100 * cos)
    again = eng.vision.image_features(pixels).cpu()
    assert torch.equal(feats, again)                        # bit-reproducible
    one = torch.cat([eng.vision.image_features(pixels[b:b + 1]) for b in range(5)]).cpu()
    assert torch.equal(feats, one)                          # images do not interact


@pytest.mark.parametrize("H,S", [(128, 32), (512, 224), (96, 64), (40, 64)])
def test_preprocess_matches_oracle(H, S):
    """Antialiased bicubic down-sampling (512 -> 224 is the debias loop's case), a non-integer ratio and an up-sampling case: the same
    uint8 level as torch's own antialiased resize except where a value lands within float rounding of .5 (one level, a handful of pixels)."""
    from uce_b200.clip_zero_shot import ClipVisionEngine
    model, cfg = _model(vision=dict(image_size=S, patch_size=S // 4))
    eng = ClipVisionEngine(model.state_dict(), cfg.vision_config.num_attention_heads)
    imgs = _smooth_images(3, H, seed=H)
    got = eng.preprocess(imgs.cuda()).cpu()
    ref = Z.preprocess(imgs, size=S)
    assert got.shape == ref.shape == (3, 3, S, S)
    level = 1.0 / 255 / 0.26
    d = (got - ref).abs()
    assert float(d.max()) <= level * 1.001 + 1e-6, float(d.max())
    assert float((d > 1e-5).float().mean()) <= 2e-3, float((d > 1e-5).float().mean())
    # the same geometry needs no resampling and is exact
    same = eng.preprocess(imgs[:, :S, :S].contiguous().cuda()).cpu() if H >= S else None
    if same is not None:
        assert float((same - Z.preprocess(imgs[:, :S, :S].contiguous(), size=S)).abs().max()) <= 1e-6


def test_preprocess_matches_clip_image_processor():
    """Against the library's own CLIPImageProcessor (PIL resampling in integer arithmetic): within the 3 uint8 levels the oracle is held to."""
    from uce_b200.clip_zero_shot import ClipVisionEngine
    model, cfg = _model(vision=dict(image_size=32, patch_size=8))
    eng = ClipVisionEngine(model.state_dict(), cfg.vision_config.num_attention_heads)
    proc = transformers.CLIPImageProcessor(size={"shortest_edge": 32}, crop_size={"height": 32, "width": 32})
    sm = _smooth_images(2, 128, seed=2)
    ref = torch.tensor(proc(images=[x.numpy() for x in sm], return_tensors="np")["pixel_values"])
    got = eng.preprocess(sm.cuda()).cpu()
    assert float((got - ref).abs().max()) < 3.0 / 255 / 0.26 + 1e-6


def test_callable_like_the_zero_shot_pipeline():
    """``clip(images, candidate_labels=...)`` as debias.get_ratios calls it: per image the labels sorted by score; top-1 equals the oracle's
    on the same images, scores sum to one; PIL-free inputs (uint8 arrays, a device tensor) and a single image all work."""
    from uce_b200.clip_zero_shot import ClipZeroShotEngine
    model, cfg = _model()
    t = cfg.text_config
    words = {"this": 3, "is": 4, "a": 5, "photo": 6, "of": 7, "male.": 8, "female.": 9, "other.": 10}

    class Tok:
        def __call__(self, texts, padding=None, return_tensors=None, **_):
            rows = [[t.bos_token_id] + [words[w] for w in s.lower().split()] + [t.eos_token_id] for s in texts]
            L = max(len(r) for r in rows)
            return {"input_ids": torch.tensor([r + [t.pad_token_id] * (L - len(r)) for r in rows])}

    P = model.state_dict()
    eng = ClipZeroShotEngine(P, cfg.vision_config.num_attention_heads, t.num_attention_heads, tokenizer=Tok(), eos_token_id=t.eos_token_id)
    imgs = _smooth_images(6, 160, seed=5)
    labels = ["male", "female", "other"]
    res = eng([x.numpy() for x in imgs], candidate_labels=labels)
    ids = Tok()([Z.HYPOTHESIS_TEMPLATE.format(c) for c in labels])["input_ids"]
    ref_logits = Z.logits_per_image(P, Z.preprocess(imgs, size=64), ids, cfg.vision_config.num_attention_heads, t.num_attention_heads, t.eos_token_id)
    ref = Z.classify(ref_logits, labels)
    assert len(res) == 6
    for r, w in zip(res, ref):
        assert [x["label"] for x in r] == [x["label"] for x in w] or abs(w[0]["score"] - w[1]["score"]) < 1e-3
        assert abs(sum(x["score"] for x in r) - 1.0) < 1e-5
        for a, b in zip(r, w):
            assert abs(a["score"] - b["score"]) < 5e-3          # a preprocessing level here and there moves a score in the 4th digit
    dev = eng(imgs.cuda(), candidate_labels=labels)               # device tensor straight from the VAE engine: no host round trip
    assert [[x["label"] for x in r] for r in dev] == [[x["label"] for x in r] for r in res]
    single = eng(imgs[0].numpy(), candidate_labels=labels)
    assert [x["label"] for x in single] == [x["label"] for x in res[0]]
