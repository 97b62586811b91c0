"""The UCE artifact (safetensors file of trainscripts/uce_sd_erase.py:85-88 / evalscripts/generate-images-sd.py:17-19) through the
native writer / reader of libuce_b200 (csrc/artifact.cu): checked against the `safetensors` package the reference uses.
Host-only code: no GPU needed."""
import os
import struct

import pytest
import torch

from uce_b200.artifact import load_artifact, save_artifact
from uce_b200._native import UCEError


def _sd14_like(seed=0, dims=((320, 768), (640, 768), (1280, 768))):
    g = torch.Generator().manual_seed(seed)
    names = ["down_blocks.0.attentions.0.transformer_blocks.0.attn2.to_k", "mid_block.attentions.0.transformer_blocks.0.attn2.to_v",
             "up_blocks.1.attentions.2.transformer_blocks.0.attn2.to_k"]
    return {n + ".weight": torch.randn(*d, generator=g) for n, d in zip(names, dims)}


def test_writer_is_byte_identical_to_safetensors(tmp_path):
    from safetensors.torch import save_file
    state = _sd14_like()
    ours, theirs = str(tmp_path / "ours.safetensors"), str(tmp_path / "theirs.safetensors")
    save_artifact(dict(reversed(list(state.items()))), ours)          # insertion order must not matter
    save_file(state, theirs)                                          # exactly the reference's call (uce_sd_erase.py:88)
    assert open(ours, "rb").read() == open(theirs, "rb").read()


def test_reference_reader_loads_our_file_and_we_load_theirs(tmp_path):
    from safetensors.torch import load_file, save_file
    state = _sd14_like(seed=3, dims=((8, 5), (1, 7), (33, 2)))
    p = str(tmp_path / "a.safetensors")
    save_artifact(state, p)
    got = load_file(p)                                                # generate-images-sd.py:17
    assert set(got) == set(state) and all(torch.equal(got[k], state[k]) for k in state)
    q = str(tmp_path / "b.safetensors")
    save_file(state, q, metadata={"format": "pt", "note": "with \"quotes\" and a {brace}"})      # metadata is skipped, not required
    back = load_artifact(q)
    assert set(back) == set(state) and all(torch.equal(back[k], state[k]) for k in state)
    assert all(torch.equal(v, state[k]) for k, v in load_artifact(p).items())


def test_header_layout(tmp_path):
    """8-byte little-endian length, JSON padded with spaces to a multiple of 8, keys in name order, offsets contiguous."""
    import json
    state = _sd14_like(seed=1, dims=((2, 3), (4, 1), (1, 1)))
    p = str(tmp_path / "h.safetensors")
    save_artifact(state, p)
    raw = open(p, "rb").read()
    n = struct.unpack("<Q", raw[:8])[0]
    assert n % 8 == 0 and len(raw) == 8 + n + 4 * sum(t.numel() for t in state.values())
    hdr = json.loads(raw[8:8 + n])
    assert list(hdr) == sorted(state)
    end = 0
    for k in sorted(state):
        assert hdr[k]["dtype"] == "F32" and hdr[k]["shape"] == list(state[k].shape) and hdr[k]["data_offsets"][0] == end
        end = hdr[k]["data_offsets"][1]
    assert end == len(raw) - 8 - n


def test_empty_and_errors(tmp_path):
    from safetensors.torch import load_file
    p = str(tmp_path / "e.safetensors")
    save_artifact({}, p)
    assert load_file(p) == {} and load_artifact(p) == {}
    with pytest.raises(ValueError):
        save_artifact({"x": torch.zeros(3)}, p)                       # not two-dimensional
    with pytest.raises(ValueError):
        save_artifact({"x": torch.zeros(2, 2, dtype=torch.float16)}, p)
    with pytest.raises(UCEError):
        load_artifact(str(tmp_path / "missing.safetensors"))
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(struct.pack("<Q", 1 << 40) + b"{}")
    with pytest.raises(UCEError):
        load_artifact(bad)
    trunc = str(tmp_path / "trunc.safetensors")
    save_artifact(_sd14_like(dims=((4, 4), (4, 4), (4, 4))), trunc)
    data = open(trunc, "rb").read()
    open(trunc, "wb").write(data[:-8])                                # tensor data cut short
    with pytest.raises(UCEError):
        load_artifact(trunc)
    from safetensors.torch import save_file
    half = str(tmp_path / "half.safetensors")
    save_file({"w.weight": torch.zeros(2, 2, dtype=torch.float16)}, half)
    with pytest.raises(UCEError):
        load_artifact(half)                                           # the artifact is fp32 (uce_sd_erase.py:117)
