"""Batched concept encoding (uce_b200.concepts.embed_concepts_batched) against the one-prompt-per-forward path the
reference takes (trainscripts/uce_sd_erase.py:26-42), on a real transformers CLIPTextModel (random weights, reduced size) behind a
pipeline-shaped object whose encode_prompt does what diffusers' does for SD-1.x.  CPU only."""
import pytest
import torch

transformers = pytest.importorskip("transformers")


class WordTokenizer:
    """Deterministic stand-in for CLIPTokenizer (its vocabulary is not available offline): BOS, one id per word, EOS, EOS padding."""
    model_max_length = 77

    def __init__(self, vocab):
        self.vocab = vocab

    def __call__(self, text, padding=None, max_length=None, truncation=None, return_tensors=None):
        texts = [text] if isinstance(text, str) else list(text)
        ids = torch.full((len(texts), max_length), self.vocab - 1, dtype=torch.long)
        mask = torch.zeros((len(texts), max_length), dtype=torch.long)
        for b, t in enumerate(texts):
            words = t.split()[: max_length - 2]
            ids[b, 0] = self.vocab - 2
            for i, w in enumerate(words):
                ids[b, 1 + i] = sum(ord(c) for c in w) % (self.vocab - 2)
            mask[b, : len(words) + 2] = 1
        return {"input_ids": ids, "attention_mask": mask}


class Pipe:
    def __init__(self):
        vocab = 400
        cfg = transformers.CLIPTextConfig(vocab_size=vocab, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                                          max_position_embeddings=77, hidden_act="quick_gelu", bos_token_id=vocab - 2, eos_token_id=vocab - 1,
                                          pad_token_id=vocab - 1)
        torch.manual_seed(0)
        self.text_encoder = transformers.CLIPTextModel(cfg).eval()
        self.tokenizer = WordTokenizer(vocab)
        self.encode_calls = 0

    def encode_prompt(self, prompt, device=None, num_images_per_prompt=1, do_classifier_free_guidance=False):
        self.encode_calls += 1
        tok = self.tokenizer(prompt, padding="max_length", max_length=77, truncation=True, return_tensors="pt")
        with torch.no_grad():
            return (self.text_encoder(tok["input_ids"])[0], None)


PROMPTS = ["Van Gogh", "art", "", "a painting of a very long prompt " * 20, "Van Gogh", "Monet", "style of Kelly McKernan"]


def test_batched_rows_equal_one_by_one_rows(monkeypatch):
    from uce_b200.concepts import can_batch_encode, embed_concepts, embed_concepts_batched
    pipe = Pipe()
    assert can_batch_encode(pipe)
    with torch.no_grad():
        one = embed_concepts(pipe, PROMPTS, "cpu", batched=False)
        n_calls = pipe.encode_calls
        bat = embed_concepts_batched(pipe, PROMPTS, "cpu", batch_size=3)          # two full chunks and a remainder
    assert n_calls == len(set(PROMPTS)) and pipe.encode_calls == n_calls           # de-duplicated; the batched path never calls encode_prompt
    assert list(one) == list(bat)                                                  # same prompts, same first-seen order
    for p in one:
        assert one[p].shape == bat[p].shape == (64,) and bat[p].dtype == torch.float32
        assert float((one[p] - bat[p]).abs().max()) < 2e-5, p
    # the empty prompt keeps token 0 (BOS): mask.sum() - 2 = 0 (uce_sd_erase.py:34-42)
    with torch.no_grad():
        h = pipe.text_encoder(pipe.tokenizer("", padding="max_length", max_length=77)["input_ids"])[0]
    assert torch.allclose(bat[""], h[0, 0], atol=2e-5)
    # environment switch, and pipelines it does not apply to
    monkeypatch.setenv("UCE_BATCHED_ENCODE", "1")
    with torch.no_grad():
        env = embed_concepts(pipe, PROMPTS, "cpu")
    assert pipe.encode_calls == n_calls and all(float((env[p] - one[p]).abs().max()) < 2e-5 for p in env)      # went through the batched path
    pipe.text_encoder_2 = object()                                                 # SDXL layout: falls back to encode_prompt
    assert not can_batch_encode(pipe)
    with torch.no_grad():
        embed_concepts(pipe, ["art"], "cpu")
    assert pipe.encode_calls == n_calls + 1
