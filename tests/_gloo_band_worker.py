"""Worker of tests/test_bands_host.py::test_handle_exchange_gloo (launched by torchrun, world_size 2, gloo, CPU only):
the host-side plumbing of the one-process-per-GPU band mode -- layout agreement across ranks and the all_gather of the
IPC handle blobs -- without touching a device."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    svo = load_package()
    B = svo.bands
    # a recognisable fake handle blob per rank (the real ones come from cudaIpcGetMemHandle on the GPU box)
    mine = np.full(B.HANDLE_BYTES, 17 * (rank + 1), dtype=np.uint8)
    mine[:4] = np.frombuffer(np.int32(rank).tobytes(), dtype=np.uint8)
    allh = B.exchange_handles(mine, dist)
    assert allh.shape == (world, B.HANDLE_BYTES)
    for r in range(world):
        assert int(np.frombuffer(allh[r, :4].tobytes(), dtype=np.int32)[0]) == r
        assert (allh[r, 4:] == 17 * (r + 1)).all()
    # every rank derives the same partition and the union of the ranks' rows is the screen, each row exactly once
    for res_x, res_y, sr in ((1920, 1024, 0), (3840, 2160, 0), (3840, 2160, 16), (200, 120, 32)):
        lay = B.layout(rank, world, res_x, res_y, sr)
        rows = B.owned_rows(rank, world, res_y, sr)
        assert lay["rows"] == len(rows)
        import torch
        cover = torch.zeros(res_y, dtype=torch.int32)
        cover[torch.from_numpy(rows)] = 1
        dist.all_reduce(cover)
        assert bool((cover == 1).all()), "rows are not partitioned exactly once"
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
