"""TEST INFRASTRUCTURE — golden vectors for the command-line seam: what the REFERENCE scripts make of their arguments.

Runs the reference's own ``__main__`` blocks (trainscripts/uce_sd_erase.py:97-200, trainscripts/uce_sd_debias.py:155-243) in the build
container, unmodified, under ``runpy`` with ``sys.argv`` set per case.  ``diffusers`` is not installed, so a stub module stands in whose
``DiffusionPipeline.from_pretrained`` raises a sentinel: both scripts resolve and PRINT their concept lists before loading the model
(:193-200 / :236-243), so everything up to that call — defaults, ``;`` splitting, guide broadcast, prompt expansion, the length checks —
is the reference's code.  The printed lists (or the exception text) are written to tests/golden/cli_cases.json; tests/test_cli_golden.py
holds our CLIs to them.  Only usable where /root/reference is mounted; the fixture travels, this script documents how it was made.

    python -m oracle.make_cli_golden
"""
from __future__ import annotations

import ast
import contextlib
import io
import json
import os
import runpy
import sys
import types

from oracle.ref_harness import REFERENCE_ROOT, reference_available

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cli_cases.json")


class _Loaded(Exception):
    """Raised by the stub pipeline loader: the script got as far as loading the model."""


ERASE_CASES = [
    ["--edit_concepts", "Van Gogh", "--concept_type", "art"],
    ["--edit_concepts", "dog; cat ;  french horn", "--concept_type", "object"],
    ["--edit_concepts", "Monet; Picasso", "--guide_concepts", "impressionism", "--concept_type", "art"],
    ["--edit_concepts", "nudity; violence", "--guide_concepts", "person; peace", "--concept_type", "unsafe"],
    ["--edit_concepts", "Kelly McKernan", "--concept_type", "art", "--expand_prompts", "true", "--preserve_concepts", "Monet;  Rembrandt ; "],
    ["--edit_concepts", "dog;cat", "--guide_concepts", "animal", "--concept_type", "object", "--expand_prompts", "true"],
    ["--edit_concepts", "a; b; c", "--guide_concepts", "x; y", "--concept_type", "object"],                  # length mismatch
    ["--edit_concepts", "Van Gogh", "--concept_type", "art", "--expand_prompts", "True"],                    # only the string 'true' expands
    ["--edit_concepts", "", "--concept_type", "object"],                                                    # empty prompt
]
DEBIAS_CASES = [
    ["--edit_concepts", "doctor; nurse", "--debias_concepts", "male; female"],
    ["--edit_concepts", "ceo", "--debias_concepts", "white; black; asian", "--desired_ratios", "0.4", "0.3", "0.3", "--preserve_concepts", "dog ; cat"],
    ["--edit_concepts", "ceo", "--debias_concepts", "white; black; asian"],                                  # 3 concepts, 2 default ratios
]


def _run(script, argv):
    stub = types.ModuleType("diffusers")

    class DiffusionPipeline:
        @staticmethod
        def from_pretrained(*a, **k):
            raise _Loaded()

    stub.DiffusionPipeline = DiffusionPipeline
    old_mod, old_argv = sys.modules.get("diffusers"), sys.argv
    sys.modules["diffusers"] = stub
    sys.argv = [script] + list(argv)
    buf = io.StringIO()
    case = {"argv": list(argv)}
    cwd = os.getcwd()
    try:
        import tempfile
        with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(buf), contextlib.redirect_stderr(io.StringIO()):
            os.chdir(td)                      # the scripts create --save_dir (default ../uce_models) relative to the cwd
            os.makedirs(os.path.join(td, "work"))
            os.chdir(os.path.join(td, "work"))
            try:
                runpy.run_path(os.path.join(REFERENCE_ROOT, "trainscripts", script), run_name="__main__")
                case["outcome"] = "returned"
            except _Loaded:
                case["outcome"] = "loads_model"
            except SystemExit as e:
                case["outcome"] = "argparse_exit"; case["code"] = e.code
            except Exception as e:            # the scripts raise a bare Exception on length mismatches
                case["outcome"] = "exception"; case["message"] = str(e)
    finally:
        os.chdir(cwd)
        sys.argv = old_argv
        if old_mod is not None:
            sys.modules["diffusers"] = old_mod
        else:
            del sys.modules["diffusers"]
    lists = {}
    for line in buf.getvalue().splitlines():
        for label in ("Erasing", "Guiding", "Preserving", "Editing", "Debias Across"):
            if line.startswith(label + ": "):
                lists[label] = ast.literal_eval(line[len(label) + 2:])
    case["printed"] = lists
    return case


def main():
    if not reference_available():
        raise SystemExit("the reference tree is not mounted: fixtures can only be regenerated in the build container")
    out = {"erase": [_run("uce_sd_erase.py", a) for a in ERASE_CASES], "debias": [_run("uce_sd_debias.py", a) for a in DEBIAS_CASES]}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)
    for kind, cases in out.items():
        for c in cases:
            print(kind, c["outcome"], c["argv"][:4], {k: len(v) for k, v in c["printed"].items()})


if __name__ == "__main__":
    main()
