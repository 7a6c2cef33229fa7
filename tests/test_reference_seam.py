"""CPU: the drop-in through the REFERENCE's own factories (`make_network`, `make_renderer`, `net_utils.load_network`) and
process-global `lib.config.cfg` -- the central claim of INTEGRATION.md, guarded against regressions.  Needs the reference
tree (/root/reference, build container only); skipped where it is absent (the GPU box)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir('/root/reference/lib'), reason='the reference tree is only present in the build container')
def test_drop_in_through_the_reference_factories():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'ref_seam_check.py')], cwd=ROOT, capture_output=True, text=True,
                       timeout=300, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + '\n' + p.stderr[-3000:]
    assert 'REF_SEAM_OK 46 1274652' in p.stdout
