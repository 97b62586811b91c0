"""The command-line seam against the REAL reference: tests/golden/cli_cases.json holds what the reference's own
``__main__`` blocks printed / raised for a set of argument vectors (made by oracle/make_cli_golden.py, which executes
trainscripts/uce_sd_erase.py:97-200 and trainscripts/uce_sd_debias.py:155-243 unmodified up to the model load).  Our drop-in CLIs must
resolve the same arguments to the same concept lists, refuse what the reference refuses, and print the same three lines."""
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = json.load(open(os.path.join(ROOT, "tests", "golden", "cli_cases.json")))


def _load(rel, alias):
    spec = importlib.util.spec_from_file_location(alias, os.path.join(ROOT, rel))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("case", CASES["erase"], ids=lambda c: " ".join(c["argv"])[:60])
def test_erase_cli_resolves_arguments_like_the_reference(case):
    m = _load("trainscripts/uce_sd_erase.py", "cli_erase_golden")
    if case["outcome"] == "argparse_exit":
        with pytest.raises(SystemExit) as e:
            m.build_parser().parse_args(case["argv"])
        assert e.value.code == case["code"]
        return
    args = m.build_parser().parse_args(case["argv"])
    if case["outcome"] == "exception":
        with pytest.raises(Exception) as e:
            m.resolve(args)
        assert not isinstance(e.value, SystemExit)
        assert str(e.value) == case["message"]
        return
    assert case["outcome"] == "loads_model"
    edit, guide, preserve = m.resolve(args)
    assert edit == case["printed"]["Erasing"]
    assert guide == case["printed"]["Guiding"]
    assert preserve == case["printed"]["Preserving"]


@pytest.mark.parametrize("case", CASES["debias"], ids=lambda c: " ".join(c["argv"])[:60])
def test_debias_cli_resolves_arguments_like_the_reference(case, capsys, monkeypatch):
    """main() is run up to the model load: without diffusers it stops there with SystemExit, after printing the lists (as the reference
    prints them before DiffusionPipeline.from_pretrained, uce_sd_debias.py:232-238)."""
    import sys
    m = _load("trainscripts/uce_sd_debias.py", "cli_debias_golden")
    monkeypatch.setitem(sys.modules, "diffusers", None)          # make the import fail even where diffusers is installed
    monkeypatch.chdir(os.path.join(ROOT, "tests"))
    tmp_save = os.path.join(os.environ.get("TMPDIR", "/tmp"), "uce_cli_golden_models")
    argv = case["argv"] + ["--save_dir", tmp_save]
    if case["outcome"] == "exception":
        with pytest.raises(Exception) as e:
            m.main(argv)
        assert not isinstance(e.value, SystemExit) and str(e.value) == case["message"]
        return
    with pytest.raises(SystemExit):
        m.main(argv)
    out = capsys.readouterr().out
    for label in ("Editing", "Debias Across", "Preserving"):
        assert f"{label}: {case['printed'][label]}" in out, (label, out)


def test_erase_cli_prints_the_reference_lines(capsys, monkeypatch):
    import sys
    m = _load("trainscripts/uce_sd_erase.py", "cli_erase_golden_main")
    monkeypatch.setitem(sys.modules, "diffusers", None)
    case = CASES["erase"][4]                                     # expanded art prompts + preserve list
    tmp_save = os.path.join(os.environ.get("TMPDIR", "/tmp"), "uce_cli_golden_models")
    with pytest.raises(SystemExit):
        m.main(case["argv"] + ["--save_dir", tmp_save, "--device", "cpu"])
    out = capsys.readouterr().out
    for label in ("Erasing", "Guiding", "Preserving"):
        assert f"{label}: {case['printed'][label]}" in out, (label, out)
