"""__graft_entry__: build() works without a GPU; smoke() on the GPU."""
import pytest


def test_build_entry(built):
    import __graft_entry__

    __graft_entry__.build()


@pytest.mark.gpu
def test_smoke_entry(built):
    import subprocess
    import sys

    from conftest import REPO

    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=REPO, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout
