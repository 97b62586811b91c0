"""The mbarrier protocols of the apply kernels under random, skewed interleavings (CPU models: scripts/protocol_sim_tc3.py for the
default kernel csrc/apply_tc3.cu as committed, scripts/protocol_sim.py for the opt-in high-rank kernel).  A few schedules per chunk
count here; the scripts run thousands.  Also checks that the models have teeth: a seeded protocol bug is found."""
import importlib.util
import os
import random

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "scripts", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_default_apply_kernel_protocol_is_sound():
    m = _load("protocol_sim_tc3")
    rng = random.Random(11)
    for n_act in (2, 1):
        for n_chunks in (1, 4, 5, 6, 7, 12, 24):                # K = 32 .. 768; around the ring depths 3 / 5 / 6
            m.Sim(n_chunks, n_act, rng).run()


def test_default_apply_kernel_model_detects_a_seeded_bug():
    src = open(os.path.join(ROOT, "scripts", "protocol_sim_tc3.py")).read()
    bad = src.replace("yield from self.wait(self.raw_empty[u], (uses - 1) & 1)", "pass")        # addend prefetched into a raw stage still being read
    assert bad != src
    ns = {}
    exec(compile(bad, "protocol_sim_tc3_mutant", "exec"), ns)
    rng = random.Random(12)
    caught = 0
    for _ in range(6):
        try:
            ns["Sim"](24, 2, rng).run()
        except AssertionError as e:
            caught += 1
            assert "overwrote" in str(e)
    assert caught >= 1


def test_highrank_apply_kernel_protocol_is_sound():
    m = _load("protocol_sim")
    rng = random.Random(13)
    for pipelined in (True, False):
        for n_chunks in (1, 3, 4, 5, 8, 16):
            m.Sim(n_chunks, pipelined, rng).run()
    for n_chunks in (1, 2, 3, 4, 7, 16):
        m.SimSS(n_chunks, rng).run()                              # csrc/apply_gemm3x_ss.cu
