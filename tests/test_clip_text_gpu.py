"""The B200 CLIP text encoder (include/clip_text_b200.h) against the oracle pinned to transformers
(oracle/clip_text_oracle.py) and against transformers' CLIPTextModel itself: fp32 parity, batching, the last-real-token rows, and
the drop-in route embed_concepts() takes for an SD-1.x style pipeline."""
import pytest
import torch

from oracle import clip_text_oracle as CO

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def _model(layers=3, hidden=64, heads=4, vocab=500, ffn=None):
    cfg = transformers.CLIPTextConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=ffn or 4 * hidden, num_hidden_layers=layers,
                                      num_attention_heads=heads, max_position_embeddings=77, hidden_act="quick_gelu",
                                      bos_token_id=vocab - 2, eos_token_id=vocab - 1, pad_token_id=vocab - 1)
    torch.manual_seed(0)
    return transformers.CLIPTextModel(cfg).eval(), cfg


def _prompts(cfg, lengths, seed=1):
    g = torch.Generator().manual_seed(seed)
    ids = torch.full((len(lengths), 77), cfg.eos_token_id, dtype=torch.long)
    mask = torch.zeros((len(lengths), 77), dtype=torch.long)
    for b, n in enumerate(lengths):                         # [BOS, n words, EOS, padding = EOS]
        ids[b, 0] = cfg.bos_token_id
        ids[b, 1:1 + n] = torch.randint(0, cfg.vocab_size - 2, (n,), generator=g)
        mask[b, : n + 2] = 1
    return ids, mask


@pytest.mark.parametrize("layers,hidden,heads,ffn", [(3, 64, 4, None), (2, 128, 2, 320), (2, 768, 12, 3072)])
def test_engine_matches_oracle_and_transformers(layers, hidden, heads, ffn):
    """Small configurations and the CLIP ViT-L/14 text width SD-1.4 uses (768 wide, 12 heads of 64, quick-GELU MLP of 3072)."""
    from uce_b200.clip_text import ClipTextEngine
    model, cfg = _model(layers, hidden, heads, ffn=ffn)
    ids, mask = _prompts(cfg, [1, 2, 7, 0, 75, 13, 40])
    P = model.state_dict()
    ref = CO.encode(P, ids, heads)
    with torch.no_grad():
        lib = model(input_ids=ids).last_hidden_state
    eng = ClipTextEngine(P, heads, max_batch=4)             # 7 prompts through a 4-prompt engine: two forwards
    got = eng.encode(ids).cpu()
    assert got.shape == ref.shape and eng.launch_count() == 2 + 7 * layers
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, scale), float((got - ref).abs().max())
    assert float((got - lib).abs().max()) <= 2e-5 * max(1.0, scale)
    rows = eng.concept_rows(ids, mask).cpu()
    ref_rows = CO.concept_rows(P, ids, mask, heads)
    assert float((rows - ref_rows).abs().max()) <= 2e-5 * max(1.0, scale)
    again = eng.encode(ids).cpu()
    assert torch.equal(got, again)                           # bit-reproducible
    one = torch.cat([eng.encode(ids[b:b + 1]) for b in range(ids.shape[0])]).cpu()
    assert torch.equal(got, one)                             # rows do not interact: batching changes nothing
    eng.close()


def test_embed_concepts_routes_through_the_engine(monkeypatch):
    """embed_concepts() — the call erase.UCE / debias.UCE make — on a pipeline object with a CLIPTextModel takes the engine on a CUDA
    device and returns the rows the reference's one-prompt-at-a-time loop gets (trainscripts/uce_sd_erase.py:26-42)."""
    import types
    from uce_b200 import concepts
    model, cfg = _model(2, 64, 4)
    words = {"van": 3, "gogh": 4, "art": 5, "picasso": 6, "a": 7, "dog": 8}

    class Tok:
        model_max_length = 77

        def __call__(self, texts, padding=None, max_length=None, truncation=None, return_tensors=None):
            single = isinstance(texts, str)
            texts = [texts] if single else list(texts)
            ids = torch.full((len(texts), 77), cfg.eos_token_id, dtype=torch.long)
            mask = torch.zeros((len(texts), 77), dtype=torch.long)
            for b, t in enumerate(texts):
                w = [words[x] for x in t.lower().split()]
                ids[b, 0] = cfg.bos_token_id
                ids[b, 1:1 + len(w)] = torch.tensor(w, dtype=torch.long)
                mask[b, : len(w) + 2] = 1
            return {"input_ids": ids, "attention_mask": mask}

    pipe = types.SimpleNamespace(text_encoder=model, text_encoder_2=None, tokenizer=Tok())

    def encode_prompt(prompt, device=None, num_images_per_prompt=1, do_classifier_free_guidance=False):
        with torch.no_grad():
            return (model(input_ids=pipe.tokenizer(prompt)["input_ids"]).last_hidden_state,)
    pipe.encode_prompt = encode_prompt
    prompts = ["Van Gogh", "Picasso", "art", "", "a dog", "Van Gogh"]
    assert concepts.can_use_text_engine(pipe, "cuda:0")
    rows = concepts.embed_concepts(pipe, prompts, "cuda:0")
    monkeypatch.setenv("UCE_TEXT_ENGINE", "0")
    ref = concepts.embed_concepts(pipe, prompts, "cpu")              # the reference's loop: one encode_prompt per distinct prompt
    assert set(rows) == set(ref) == set(prompts)
    for p in ref:
        assert float((rows[p].cpu() - ref[p]).abs().max()) <= 2e-5, p
