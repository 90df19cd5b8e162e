"""CPU tests of the multi-GPU band host logic: the stripe layout arithmetic exported by the library
(svo_band_layout / svo_band_owner, no device needed), its Python mirror, and the handle exchange of the
one-process-per-GPU mode over torch.distributed (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def B():
    from __graft_entry__ import load_package
    return load_package().bands


@pytest.mark.parametrize("res", [(1920, 1024), (3840, 2160), (320, 192), (200, 120), (64, 48), (100, 30)])
@pytest.mark.parametrize("G", [1, 2, 3, 4, 8])
@pytest.mark.parametrize("sr", [0, 16, 32, 64])
def test_layout_partitions_the_screen(B, res, G, sr):
    rx, ry = res
    SR = B.effective_stripe_rows(G, ry, sr)
    assert SR % 16 == 0 and SR > 0
    seen_rows = np.zeros(ry, dtype=int)
    nb = (rx // 16) * (ry // 16)
    seen_blocks = np.zeros(nb, dtype=int)
    for r in range(G):
        lay = B.layout(r, G, rx, ry, sr)
        rows = B.owned_rows(r, G, ry, sr)
        blocks = B.owned_blocks(r, G, rx, ry, sr)
        assert lay["SR"] == SR and lay["rows"] == len(rows)
        assert lay["brows"] * (rx // 16) == len(blocks)
        assert all(B._band_owner(int(y), G, SR) == r for y in rows[::7])
        # local -> global row mapping used by the kernels (BandMap::global_row)
        lr = np.arange(len(rows))
        k = lr // SR
        assert np.array_equal((k * G + r) * SR + (lr - k * SR), rows)
        # whole block rows are a prefix of the local rows, in the same order (BandMap::global_brow)
        q = SR // 16
        lbr = np.arange(lay["brows"])
        gbr = ((lbr // q) * G + r) * q + lbr % q
        assert np.array_equal(np.unique(blocks // max(1, rx // 16)) if len(blocks) else np.zeros(0, int), gbr)
        seen_rows[rows] += 1
        seen_blocks[blocks] += 1
    assert (seen_rows == 1).all() and (seen_blocks == 1).all()


def test_layout_rejects_bad_arguments(B):
    with pytest.raises(ValueError):
        B.layout(2, 2, 64, 64, 0)
    with pytest.raises(ValueError):
        B.layout(0, 9, 64, 64, 0)


def test_handle_exchange_gloo():
    """world_size 2 on CPU: the ranks exchange (fake) IPC handle blobs and agree on a partition of the screen."""
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(ROOT, "tests", "_gloo_band_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
