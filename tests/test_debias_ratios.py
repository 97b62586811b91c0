"""get_ratios() of the debias driver (mirror of trainscripts/uce_sd_debias.py:14-35): label fractions -> direction scales with the
dead-band, single process and sharded over ranks (world_size-2 gloo on CPU: concepts dealt round-robin, ONE all-reduce of the
label counts — SURVEY.md 8e)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

EDIT = ["doctor", "nurse", "teacher", "ceo", "chef"]
DEB = ["male", "female"]
LABELS = {"doctor": ["male"] * 8 + ["female"] * 2, "nurse": ["female"] * 9 + ["male"], "teacher": ["male"] * 5 + ["female"] * 5,
          "ceo": ["male"] * 10, "chef": ["female"] * 6 + ["male"] * 4}


class _Images:
    def __init__(self, images):
        self.images = images


class _Pipe:
    """pipe(...) -> .images, pipe.unet.load_state_dict(...): what get_ratios touches (uce_sd_debias.py:17-26)."""
    def __init__(self):
        self.calls = []
        self.unet = self

    def load_state_dict(self, state, strict=True):
        return None

    def __call__(self, prompt, num_inference_steps=20, num_images_per_prompt=10, guidance_scale=7.5):
        self.calls.append(prompt)
        return _Images([(prompt, i) for i in range(num_images_per_prompt)])


def _clip(images, candidate_labels):
    return [[{"label": lab, "score": 0.9}] for lab in LABELS[images[0][0]]]


def _reference_get_ratios(desired, max_diff):
    """Restatement of uce_sd_debias.py:21-35 on the scripted labels."""
    out = []
    for concept in EDIT:
        top1 = np.array(LABELS[concept])
        ratios = np.array([want - (np.sum(top1 == c) / len(top1)) for c, want in zip(DEB, desired)])
        if max(ratios) < max_diff and abs(min(ratios)) < max_diff:
            ratios = 0 * ratios
        out.append(ratios)
    return np.array(out)


def test_get_ratios_matches_the_reference_function():
    """tests/golden/get_ratios.json: the reference's OWN get_ratios (uce_sd_debias.py:14-35) called with the keyword form of its caller
    (:96-107) on a recording pipeline and scripted labels (oracle/make_ratios_golden.py).  Ours, through the same call: same weights
    loaded into pipe.unet (keys, strict=False), same pipe(...) calls with the same keyword arguments, same float64 direction scales."""
    import json
    from oracle.make_ratios_golden import RecordingPipe, modules, scripted_clip
    from uce_b200.debias import get_ratios
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "get_ratios.json")))
    for g in gold:
        case = g["case"]
        pipe = RecordingPipe()
        names, mods = modules()
        r = get_ratios(pipe=pipe, clip=scripted_clip(case["labels"]), uce_module_names=names, uce_modules=mods, edit_concepts=g["edit"],
                       debias_concepts=case["debias"], desired_ratios=case["desired"], max_diff=case["max_diff"], step_size=0.1,
                       num_images_per_prompt=case["n_img"], num_inference_steps=7, guidance_scale=6.5)
        assert pipe.loaded == g["loaded"]
        assert pipe.calls == g["calls"]
        assert type(r).__name__ == g["result_type"] and str(r.dtype) == g["result_dtype"]
        assert np.array_equal(r, np.array(g["result"]))                      # bit-equal, signed zeros of the dead-band aside
        # the weight tensors themselves are accepted in place of the modules (what debias.UCE passes), positionally as well
        pipe2 = RecordingPipe()
        r2 = get_ratios(pipe2, scripted_clip(case["labels"]), names, [m.weight for m in mods], g["edit"], case["debias"], case["desired"],
                        case["max_diff"], 0.1, case["n_img"], 7, 6.5)
        assert np.array_equal(r2, r) and pipe2.loaded == g["loaded"] and pipe2.calls == g["calls"]


def test_get_ratios_single_process_matches_reference_arithmetic():
    from uce_b200.debias import get_ratios
    for desired, max_diff in [((0.5, 0.5), 0.05), ((0.3, 0.7), 0.15), ((0.5, 0.5), 0.35)]:
        pipe = _Pipe()
        got = get_ratios(pipe, _clip, [], [], EDIT, DEB, desired, max_diff)
        assert np.array_equal(got, _reference_get_ratios(desired, max_diff))
        assert pipe.calls == EDIT


def test_get_ratios_hands_device_images_to_a_classifier_that_takes_them():
    """EngineGenerator + VAE engine return the decoded pixels also as one uint8 tensor on the device (`images_u8`); a classifier that
    declares `accepts_device_images` (ClipZeroShotEngine) gets that tensor instead of the host copies, any other `clip` (the
    transformers pipeline of the reference) gets `.images` as before — same ratios either way."""
    from uce_b200.debias import get_ratios

    class Out:
        def __init__(self, prompt, n):
            self.images = [(prompt, i) for i in range(n)]
            self.images_u8 = ("device", prompt, n)

    class Pipe(_Pipe):
        def __call__(self, prompt, num_inference_steps=20, num_images_per_prompt=10, guidance_scale=7.5):
            self.calls.append(prompt)
            return Out(prompt, num_images_per_prompt)

    class DeviceClip:
        accepts_device_images = True

        def __init__(self):
            self.seen = []

        def __call__(self, images, candidate_labels):
            self.seen.append(images)
            return [[{"label": lab, "score": 0.9}] for lab in LABELS[images[1]]]

    clip = DeviceClip()
    got = get_ratios(Pipe(), clip, [], [], EDIT, DEB, (0.5, 0.5), 0.05)
    assert all(isinstance(x, tuple) and x[0] == "device" for x in clip.seen) and [x[1] for x in clip.seen] == EDIT
    assert np.array_equal(got, _reference_get_ratios((0.5, 0.5), 0.05))
    host = get_ratios(Pipe(), _clip, [], [], EDIT, DEB, (0.5, 0.5), 0.05)          # a plain callable: the host images
    assert np.array_equal(host, got)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from uce_b200.debias import get_ratios
        pipe = _Pipe()
        got = get_ratios(pipe, _clip, [], [], EDIT, DEB, (0.5, 0.5), 0.05)
        q.put((rank, got.tolist(), pipe.calls))
    finally:
        dist.destroy_process_group()


def test_get_ratios_sharded_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    ref = _reference_get_ratios((0.5, 0.5), 0.05)
    for rank, got, calls in res:
        assert np.array_equal(np.array(got), ref)          # every rank ends with the same, complete ratio matrix
        assert calls == EDIT[rank::2]                      # and generated / classified only its own concepts
