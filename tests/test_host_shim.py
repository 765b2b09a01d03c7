"""The C++ host shim (mirrors of vector.rs / trie.rs / search.rs) -- runs the compiled
tests/host_shim_test.cpp binary.  CPU mode: trie KATs K1-K8 + randomized cross-check against the
oracle's literal trie; GPU mode adds HnswIndex / VectorIndex / SearchEngine over libtss."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "trie-semantic-search_b200", "build", "host_shim_test")


def _run(mode):
    if not os.path.exists(BIN):
        subprocess.run(["make", "-C", os.path.join(ROOT, "trie-semantic-search_b200", "host")],
                       check=True, env={k: v for k, v in os.environ.items() if k not in ("CXX", "CC")})
    p = subprocess.run([BIN, mode], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "HOST_SHIM_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


def test_host_shim_cpu():
    _run("cpu")


@pytest.mark.gpu
def test_host_shim_gpu():
    _run("gpu")
