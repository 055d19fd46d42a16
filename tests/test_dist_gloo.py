"""CPU, world_size > 1, gloo: the host-side sharding logic of the sharded paths.

The cases themselves live in tests/gloo_cases.py.  Each one runs in a FRESH interpreter: torch's `mp.spawn` + gloo
followed by NumPy LAPACK calls in the same long-lived pytest process deadlocked the documented
`python -m pytest tests -q -m "not gpu"` run (round-1 VERDICT); a subprocess per case keeps the rendezvous, its
helper threads and the forked Manager out of the main test process."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    "test_shard_range_covers_everything",
    "test_tsqr_two_ranks_matches_single_qr",
    "test_stack_layout",
    "test_tsqr_qr_ranks_match_reference_compact_factor",
    "test_tsqr_qr_short_shard_raises_on_every_rank",
]


@pytest.mark.timeout(600)
@pytest.mark.parametrize("case", CASES)
def test_gloo_case(case):
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider",
                        os.path.join(HERE, "gloo_cases.py") + "::" + case],
                       capture_output=True, text=True, timeout=580, env=env, cwd=os.path.dirname(HERE))
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-2000:])
