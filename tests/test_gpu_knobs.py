"""Every environment switch of the library selects another hand-written kernel for the same numbers (the work-queue
form of the QAGS kernel, KV1_NSp / KV1_NSm as two launches instead of the paired one, the register-staged DMMA and the DFMA product, the warp-per-row application, serial
K1).  The switches are read once per process, so each one runs tools/knob_case.py in its own subprocess: kernel
matrices and QAGS paths bit for bit against the oracle, matrices and applications within 1e-9 / 1e-12."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KNOBS = [{}, {"TKB_K2_MODE": "queue"}, {"TKB_K2_PAIR": "0"}, {"TKB_K1_SERIAL": "1"}, {"TKB_GEMM_SYNC": "1"}, {"TKB_GEMM_DFMA": "1"},
         {"TKB_APPLY_SMALL_MAX": "8"}]


@pytest.mark.parametrize("knob", KNOBS, ids=lambda k: ",".join("%s=%s" % kv for kv in k.items()) or "default")
def test_switch_keeps_the_numbers(knob):
    env = dict(os.environ)
    for k in list(env):
        if k.startswith("TKB_"):
            del env[k]
    env.update(knob)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "knob_case.py")], env=env, capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip().splitlines()[-1] == "OK", r.stdout[-2000:]
