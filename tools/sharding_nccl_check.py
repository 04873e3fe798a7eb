"""Developer check under torchrun (NCCL): both sharded operators against the single-GPU result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pyrayhf_b200  # noqa: E402
from pyrayhf_b200 import sharding, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    lat, lon = synth.grid_subset(37, seed=5)
    alt, freq = synth.default_alt(), synth.default_freq()
    den, bmag, bpsi = synth.profiles_at(lat, lon, alt)
    for layout in ("interleaved", "contiguous"):
        got = sharding.vertical_forward_operator_sharded(freq, den, bmag, bpsi, alt, 'X', 500, layout=layout, gather_to=None)
        want = pyrayhf_b200.vertical_forward_operator_batched(freq, den, bmag, bpsi, alt, 'X', 500, errors='nan')
        assert np.array_equal(got, want, equal_nan=True), layout
    f5 = np.arange(0.01, 17.41, 0.01)
    got = sharding.vertical_forward_operator_sharded_by_frequency(f5, den[0], bmag[0], bpsi[0], alt, 'O', 2000, gather_to=None)
    want = pyrayhf_b200.vertical_forward_operator(f5, den[0], bmag[0], bpsi[0], alt, 'O', 2000)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = np.isfinite(want)
    assert np.allclose(got[m], want[m], rtol=5e-10, atol=0)      # different tilings of the same rows
    dist.barrier()
    if rank == 0:
        print("sharding over NCCL ok: world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
