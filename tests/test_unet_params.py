"""The engine's parameter table (uce_b200.unet_spec.param_shapes, also the C inventory sd_unet_inventory) against the oracle's own,
independently written walk of the diffusers module tree (oracle/unet_params.py), and both against the published size of the SD-1.x
U-Net.  The U-Net oracle draws its random weights from ITS table, so the two sides of the GPU parity tests no longer share one spec."""
from oracle import unet_params as UP
from uce_b200.unet_spec import SD14, param_shapes, tiny_config


def test_sd14_tables_agree_name_by_name_and_count():
    ours, theirs = param_shapes(SD14), UP.unet_named_parameters()
    assert set(ours) == set(theirs), (sorted(set(ours) ^ set(theirs))[:10])
    for k in theirs:
        assert tuple(ours[k]) == tuple(theirs[k]), (k, ours[k], theirs[k])
    assert UP.count(theirs) == 859_520_964 == UP.count(ours)
    edited = [k for k in theirs if ".attn2.to_k." in k or ".attn2.to_v." in k]
    assert len(edited) == 32 and sum(theirs[k][0] for k in edited) == 24_960          # the 32 projections UCE edits


def test_tiny_tables_agree():
    cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    ours = param_shapes(cfg)
    theirs = UP.unet_named_parameters(block_out_channels=(64, 128), cross_attention_dim=64,
                                      down_block_types=("CrossAttnDownBlock2D", "DownBlock2D"), up_block_types=("UpBlock2D", "CrossAttnUpBlock2D"))
    assert {k: tuple(v) for k, v in ours.items()} == {k: tuple(v) for k, v in theirs.items()}
