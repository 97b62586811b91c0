"""Time the VAE decoder engine (csrc/vae_engine.cu) on the BASELINE cfg5 image size: 64 x 64 latents -> 512 x 512, batch 1 and 8,
random SD-1.4-shaped decoder weights.  CUDA events on the launching stream, 3 warm-up decodes, 10 timed.  Also prints the launch
count and the algorithmic work (2.51 TFLOP per 512 x 512 image counting 2 per multiply-add; SURVEY.md 8(f) rank 1 quotes the
1.2 T multiply-adds) so the number can be read against the tensor roofline."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uce_b200.vae import VAEDecoderEngine
from uce_b200.vae_spec import SD14_VAE, decoder_param_shapes


def conv_flops(cfg, h, w):
    """2 * MACs of every convolution / GEMM of one decode (attention included)."""
    ch = list(cfg["block_out_channels"]); top = ch[-1]
    f = 2 * h * w * (9 * 4 * top)
    res = lambda hw, ci, co: 2 * hw * (9 * ci * co + 9 * co * co + (ci * co if ci != co else 0))
    f += 2 * res(h * w, top, top) + 2 * h * w * 4 * top * top + 4 * (h * w) ** 2 * top
    cur, hh, ww = top, h, w
    for i, co in enumerate(reversed(ch)):
        for _ in range(cfg["layers_per_block"] + 1):
            f += res(hh * ww, cur, co); cur = co
        if i != len(ch) - 1:
            hh, ww = 2 * hh, 2 * ww
            f += 2 * hh * ww * 9 * co * co
    return f + 2 * hh * ww * 9 * cur * 3


def random_weights(cfg, seed=0):
    """Seeded decoder weights of the published layout: 1/sqrt(fan_in) kernels, unit norm scales, small biases."""
    import math
    g = torch.Generator().manual_seed(seed)
    P = {}
    for name, shp in decoder_param_shapes(cfg).items():
        if name.endswith(".weight") and len(shp) == 1:
            P[name] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif name.endswith(".bias"):
            P[name] = 0.02 * torch.randn(shp, generator=g)
        else:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            P[name] = torch.randn(shp, generator=g) / math.sqrt(fan_in)
    return P


def main():
    P = random_weights(SD14_VAE, seed=0)
    for batch in (1, 8):
        eng = VAEDecoderEngine(SD14_VAE, batch=batch, h=64, w=64)
        eng.load_state_dict(P); eng.finalize()
        lat = torch.randn((batch, 4, 64, 64), device="cuda") * 0.5
        for _ in range(3):
            eng.decode(lat)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            eng.decode(lat)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        fl = conv_flops(SD14_VAE, 64, 64) * batch
        print(f"vae decode batch {batch}: {ms:.2f} ms per call, {ms / batch:.2f} ms per image, {eng.launch_count()} launches, "
              f"{fl / 1e12:.2f} TFLOP -> {fl / ms / 1e9:.0f} TFLOP/s")
        eng.close()


if __name__ == "__main__":
    main()
