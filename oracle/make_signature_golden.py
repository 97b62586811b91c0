"""TEST INFRASTRUCTURE — the Python-function seam: parameter names, order and defaults of the reference's public functions
(``UCE`` of uce_sd_erase.py:12 and uce_sd_debias.py:37, ``get_ratios`` :14, ``generate_images`` of generate-images-sd.py:10), read with
``inspect`` from the reference modules loaded in the build container -> tests/golden/signatures.json.  tests/test_cli_and_host.py requires
our mirrors to accept the same positional and keyword calls.      python -m oracle.make_signature_golden
"""
from __future__ import annotations

import importlib.util
import inspect
import json
import os
import sys
import types

from oracle.ref_harness import REFERENCE_ROOT, _load, reference_available

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "signatures.json")


def sig(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        d = None if p.default is inspect._empty else repr(p.default)
        out.append([p.name, d])
    return out


def main():
    if not reference_available():
        raise SystemExit("the reference tree is not mounted")
    erase = _load("uce_sd_erase.py", "_ref_sig_erase")
    debias = _load("uce_sd_debias.py", "_ref_sig_debias")
    spec = importlib.util.spec_from_file_location("_ref_sig_gen", os.path.join(REFERENCE_ROOT, "evalscripts", "generate-images-sd.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = {"erase.UCE": sig(erase.UCE), "debias.UCE": sig(debias.UCE), "debias.get_ratios": sig(debias.get_ratios),
           "generate.generate_images": sig(gen.generate_images)}
    json.dump(out, open(OUT, "w"), indent=1)
    for k, v in out.items():
        print(k, [n for n, _ in v])


if __name__ == "__main__":
    main()
