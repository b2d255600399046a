"""Drop-in check: the reference's own unit tests for the tabular path, run unmodified from /root/reference against
this package registered as `rlpoker` (tools/run_reference_tests.py).  Skipped where the reference is not mounted
(the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'rlpoker', 'tests')), reason='reference not mounted')
def test_reference_unit_tests_pass_against_this_package():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE='1')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'run_reference_tests.py'), REF], env=env,
                       capture_output=True, text=True, timeout=600)
    tail = p.stdout[-3000:] + p.stderr[-2000:]
    assert p.returncode == 0, tail
    assert ' passed' in p.stdout and 'failed' not in p.stdout, tail
