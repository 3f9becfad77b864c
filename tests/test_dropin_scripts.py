"""CPU (build container only: needs the read-only reference checkout): the UNMODIFIED reference training scripts run
over `dropin/` (this repo's `model` / `helper_functions` modules on sys.path instead of the reference's `models/`),
one training step each, with the kernel layer mocked by the oracle (tests/_dropin_runner.py explains why and how).
Because the mock computes the reference's own arithmetic on OUR parameter objects, the step must reproduce the golden
vectors the real reference produced (tests/golden/loops.json): this pins the whole drop-in surface -- class names,
constructor signature, attributes, DataParallel / optimizer / state_dict behaviour -- against the real scripts."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DM_REFERENCE", "/root/reference")
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "loops.json")))


def close(a, b, rtol=2e-4):
    if isinstance(a, list):
        return all(close(x, y, rtol) for x, y in zip(a, b))
    return abs(a - b) <= rtol * max(abs(a), abs(b), 1e-12)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "experiments")), reason="reference checkout not present")
@pytest.mark.parametrize("script,key", [("new_vae.py", "vae"), ("new_gan.py", "gan"), ("new_betavaegan.py", "betavaegan")])
def test_unmodified_reference_script_runs_over_dropin(script, key):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_dropin_runner.py"), script], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=600)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("DROPIN_RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-3000:]
    got = json.loads(lines[-1][len("DROPIN_RESULT "):])
    want = GOLD[key][0]  # first step of the golden trajectory written by the real reference
    for k, v in want.items():
        if isinstance(v, dict):
            assert got[k] == v, (k, got[k], v)
        else:
            assert close(got[k], v), (k, got[k], v)
    if key == "betavaegan":
        assert got["n_state_keys"] == 69  # VAE state_dict entries
