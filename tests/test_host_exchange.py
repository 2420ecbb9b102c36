"""CPU-only checks of the multi-GPU exchange's host side: the region layout the library reports
(csrc/comm.cu exchange_geom) against the Python mirror in parallel.py, for every supported
world size, and that nothing here needs a GPU to be queried."""
import pytest

from easygaussiansplatting_b200 import _lib, build
from easygaussiansplatting_b200.parallel import rows_per_rank


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("N,k3", [(0, 16), (1, 16), (128, 1), (129, 4), (70001, 9), (1_000_000, 16), (5_000_000, 16)])
def test_region_layout_matches_python_mirror(world, N, k3):
    build.build()
    lib = _lib.load()
    rpr = rows_per_rank(N, world)
    assert rpr % 128 == 0 and rpr * world >= N and (rpr - 128) * world < max(N, 1) + 128 * world
    floats = 3 * k3 + 11
    total = lib.gsb_exchange_region_bytes(N, k3, world)
    assert total == 4096 + 2 * world * rpr * floats * 4          # control + staging + result
    rows = rpr * world
    off = 4096 + world * rpr * floats * 4
    for seg, k in enumerate((3 * k3, 4, 3, 3, 1)):
        assert lib.gsb_exchange_result_offset(N, k3, world, seg) == off
        assert off % 16 == 0
        off += rows * k * 4
    assert off == total


def test_unsupported_world_sizes_are_refused():
    build.build()
    lib = _lib.load()
    assert lib.gsb_exchange_region_bytes(1000, 16, 0) == 0
    assert lib.gsb_exchange_region_bytes(1000, 16, 9) == 0
    assert lib.gsb_exchange_result_offset(1000, 16, 2, 5) == 0
