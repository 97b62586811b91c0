"""Pins oracle/clip_text_oracle.py (batched CLIP text encoder + last-token select, groundwork for SURVEY.md 8(f) rank 2) against the
transformers implementation the reference's pipeline calls (trainscripts/uce_sd_erase.py:29-33).  CPU only."""
import pytest
import torch

from oracle import clip_text_oracle as C

transformers = pytest.importorskip("transformers")


def _model(layers=3, hidden=64, heads=4, vocab=500):
    cfg = transformers.CLIPTextConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=4 * hidden, num_hidden_layers=layers,
                                      num_attention_heads=heads, max_position_embeddings=77, hidden_act="quick_gelu",
                                      bos_token_id=vocab - 2, eos_token_id=vocab - 1, pad_token_id=vocab - 1)
    torch.manual_seed(0)
    return transformers.CLIPTextModel(cfg).eval(), cfg


def _prompts(cfg, lengths):
    g = torch.Generator().manual_seed(1)
    ids = torch.full((len(lengths), 77), cfg.eos_token_id, dtype=torch.long)
    mask = torch.zeros((len(lengths), 77), dtype=torch.long)
    for b, n in enumerate(lengths):                         # [BOS, n words, EOS, padding = EOS]
        ids[b, 0] = cfg.bos_token_id
        ids[b, 1:1 + n] = torch.randint(0, cfg.vocab_size - 2, (n,), generator=g)
        mask[b, : n + 2] = 1
    return ids, mask


def test_batched_encoder_matches_transformers():
    model, cfg = _model()
    ids, mask = _prompts(cfg, [1, 2, 7, 0, 75])
    with torch.no_grad():
        ref = model(input_ids=ids).last_hidden_state        # what encode_prompt runs for SD-1.x: no attention_mask
    got = C.encode(model.state_dict(), ids, cfg.num_attention_heads)
    assert got.shape == ref.shape
    assert float((got - ref).abs().max()) < 2e-5
    one_by_one = torch.cat([model(input_ids=ids[b:b + 1]).last_hidden_state for b in range(ids.shape[0])]).detach()
    assert float((got - one_by_one).abs().max()) < 2e-5      # batching changes nothing: rows do not interact


def test_last_real_token_rows():
    model, cfg = _model(layers=2)
    ids, mask = _prompts(cfg, [2, 0, 5])
    rows = C.concept_rows(model.state_dict(), ids, mask, cfg.num_attention_heads)
    with torch.no_grad():
        h = model(input_ids=ids).last_hidden_state
    for b, n in enumerate([2, 0, 5]):
        assert torch.allclose(rows[b], h[b, n], atol=2e-5)   # mask.sum() - 2 = n: the last word; BOS (index 0) for the empty prompt
