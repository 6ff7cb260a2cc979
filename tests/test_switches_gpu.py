"""The development switches that select the previous kernels ("Development switches") are read once per
process, so the default test run never takes those paths.  Each case re-runs a few parity tests of this suite in a child
process with one switch set: the K/V-cache cross-attention for every beam width (MB_CROSS_CACHED=1), the mma.sync greedy
cross-attention (MB_XE_TC=0, beams then on the cache), the separate LayerNorm statistics pass (MB_LNSTAT_FUSE=0) and the
byte-load horizontal pass of K9 (MB_K9_HBYTE=1).  The fallbacks stay correct as long as these pass."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ("MB_CROSS_CACHED", "1", "tests/test_trocr_gpu.py tests/test_configs_gpu.py",
     "decoder_logits_and_search_tiny or paths_agree or recognize_chunks_and_beam5 or beam5_matches_oracle_tiny"),
    ("MB_XE_TC", "0", "tests/test_trocr_gpu.py tests/test_configs_gpu.py",
     "decoder_logits_and_search_tiny or paths_agree or recognize_chunks_and_beam5 or beam5_matches_oracle_tiny"),
    ("MB_LNSTAT_FUSE", "0", "tests/test_trocr_gpu.py", "encoder_tiny or trocr_base_end_to_end"),
    ("MB_K9_HBYTE", "1", "tests/test_imgproc_gpu.py", "crop or k9 or K9 or pack"),
]


@pytest.mark.parametrize("var,value,files,select", CASES, ids=[c[0] for c in CASES])
def test_previous_kernels_behind_their_switches(cuda_ctx, var, value, files, select):
    env = dict(os.environ)
    env[var] = value
    cmd = [sys.executable, "-m", "pytest", *files.split(), "-x", "-q", "-k", select, "-p", "no:cacheprovider"]
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-2000:]
    assert r.returncode == 0, f"{var}={value}: {tail}"
    assert " passed" in r.stdout and "no tests ran" not in r.stdout, tail
