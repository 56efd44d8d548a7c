"""dvren_render (diff-volume-renderer_b200/apps/dvren_render.cpp): the JSON -> PPM tool on the B200 runtime.

CPU tests: argument / configuration errors (exit code 1 and the reference's messages, apps/dvren_render/
main.cpp:314-334) and the JSON reader.  GPU tests: the frames the UNMODIFIED reference tool rendered on its CPU
path (tests/golden/cli, generator tests/golden/make_cli_golden.py) are reproduced byte for byte, and the first of
them is the reference's documented example: rays=16 samples=160 (README.md:94), PPM md5 a89e8bdf...."""
import glob
import hashlib
import json
import os
import subprocess

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(REPO, "diff-volume-renderer_b200", "dvren_render")
GOLDEN = os.path.join(REPO, "tests", "golden", "cli")


def run(*args, cwd=None):
    return subprocess.run([TOOL, *args], capture_output=True, text=True, cwd=cwd, timeout=120)


def test_tool_is_built():
    assert os.access(TOOL, os.X_OK), "run __graft_entry__.build()"


def test_usage_and_missing_config(tmp_path):
    r = run()
    assert r.returncode == 1 and "Usage: dvren_render <config.json> [output.ppm]" in r.stderr
    r = run(str(tmp_path / "nope.json"))
    assert r.returncode == 1 and "Error parsing config" in r.stderr and "config file not found" in r.stderr


@pytest.mark.parametrize("text,needle", [
    ("{", "JSON parse error"),
    ('{"render": {"width": 4}}', "key 'height' not found"),
    ('{"render": {"width": 4, "height": 4, "t_far": 1, "dt": 0.1, "max_steps": 4, "sampling_mode": "wild"}, "volume": {}}',
     "unsupported sampling mode: wild"),
    ('{"render": {"width": 4, "height": 4, "t_far": 1, "dt": 0.1, "max_steps": 4}, "volume": {"size": [2, 2], "density": [1]}}',
     "volume.size must contain 3 integers"),
    ('{"render": {"width": 4, "height": 4, "t_far": 1, "dt": 0.1, "max_steps": 4, "camera": {"K": [1, 2, 3]}}, '
     '"volume": {"size": [1, 1, 1], "density": [1]}}', "array length mismatch"),
    ('{"render": {"width": 4, "height": 4, "t_far": 1, "dt": 0.1, "max_steps": 4}, '
     '"volume": {"size": [1, 1, 1], "density": [1], "oob": "mirror"}}', "unsupported oob policy: mirror"),
    ('{"render": {"width": 4, "height": 4, "t_far": 1, "dt": 0.1, "max_steps": 4}, "volume": {"size": [1, 1, 1], "density": [1]}} x',
     "trailing characters"),
])
def test_configuration_errors(tmp_path, text, needle):
    cfg = tmp_path / "bad.json"
    cfg.write_text(text)
    r = run(str(cfg))
    assert r.returncode == 1
    assert "Error parsing config: invalid_argument" in r.stderr or "Error parsing config" in r.stderr
    assert needle in r.stderr, r.stderr


def test_json_reader_accepts_escapes_exponents_and_nesting(tmp_path):
    """A valid document that exercises the reader (string escapes, exponents, nested unused members) must get past
    parsing: without a GPU the run then stops at the first device call, with a GPU it renders."""
    scene = json.loads(open(os.path.join(GOLDEN, "example_2x2x2.json")).read())
    scene["render"]["dt"] = 1e-1
    text = json.dumps(scene)
    extra = r''', "note \u00e9 \"quoted\"\n\t\/": {"nested": [[1, 2.5e+0, -3E-2], {"k": null, "t": true, "f": false}], "s": "\ud83d\ude00"}}'''
    text = text[:-1] + extra
    json.loads(text)   # still a valid document
    cfg = tmp_path / "ok.json"
    cfg.write_text(text)
    r = run(str(cfg), str(tmp_path / "o.ppm"))
    assert "Error parsing config" not in r.stderr, r.stderr
    if r.returncode != 0:   # CPU-only box: no CPU fallback exists
        assert "failed" in r.stderr and ("unsupported" in r.stderr.lower() or "cuda" in r.stderr.lower()), r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", sorted(glob.glob(os.path.join(GOLDEN, "*.json"))), ids=lambda p: os.path.basename(p)[:-5])
def test_frames_match_reference_tool_bytewise(tmp_path, cfg):
    out = tmp_path / "frame.ppm"
    r = run(cfg, str(out))
    assert r.returncode == 0, r.stderr
    want = open(cfg[:-5] + ".ppm", "rb").read()
    got = out.read_bytes()
    assert got == want, f"{sum(a != b for a, b in zip(got, want))} of {len(want)} bytes differ"
    counts = open(cfg[:-5] + ".txt").read().strip()
    assert f"Forward stats: {counts} total_ms=" in r.stdout, r.stdout
    assert "Workspace bytes total=" in r.stdout and f"Wrote \"{out}\"" in r.stdout


@pytest.mark.gpu
def test_reference_example_md5_and_default_output(tmp_path):
    """No output argument: the frame goes to output.path of the scene, relative to the working directory."""
    r = run(os.path.join(GOLDEN, "example_2x2x2.json"), cwd=tmp_path)
    assert r.returncode == 0, r.stderr
    assert "rays=16 samples=160" in r.stdout
    data = (tmp_path / "simple.ppm").read_bytes()
    assert hashlib.md5(data).hexdigest() == "a89e8bdf4a99f69f03cec8f945ef6edb"
