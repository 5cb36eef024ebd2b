"""The code paths that the default configuration does not take - kept for A/B measurements and as fall-backs for
shapes the default kernels do not cover - produce the same bits as the default ones:

* RTM_STEP_LAZY=0      the step kernel that reads all 144 rows of every tile and decodes in the consumer warps;
* RTM_STEP_FUSED=0     the two-launch pipeline (decode_tma_kernel + post_kernel);
* RTM_FUSE_POST=0      the three stand-alone entry points back to back;
* RTM_LETTERBOX_IMPL   narrow (16 pixels per thread, table lookup), pixels (one thread per 8 output pixels), direct.

The switches are read once per process, so every variant runs tests/variant_digest.py in a process of its own."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def digest(mode, dtype, **env):
    e = dict(os.environ)
    e.update(env)
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "variant_digest.py"), mode, dtype], capture_output=True, text=True,
                       env=e, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("digest")][-1]
    return line.split()[-1]


def test_fallback_paths_give_the_default_bits():
    base = digest("async", "bf16")
    assert digest("sync", "bf16") == base                                    # ordinary launches
    assert digest("async", "bf16", RTM_STEP_LAZY="0") == base
    assert digest("sync", "bf16", RTM_STEP_FUSED="0") == base
    assert digest("sync", "bf16", RTM_FUSE_POST="0") == base
    assert digest("async", "bf16", RTM_LETTERBOX_IMPL="narrow") == base
    assert digest("async", "bf16", RTM_LETTERBOX_IMPL="pixels") == base
    assert digest("async", "bf16", RTM_LETTERBOX_IMPL="direct") == base


def test_eager_step_kernel_gives_the_default_bits_on_float_heads():
    assert digest("async", "f32", RTM_STEP_LAZY="0") == digest("async", "f32")
