"""lp_render --device-build (BLASes + TLAS built on the GPU, host BVH build skipped) writes
the same image, byte for byte, as the default host-built run: the hits do not depend on the
tree (DESIGN.md sections 3 and 5b).  Named to run after the parity suites."""
import numpy as np
import pytest

from test_gpu_cli import read_ppm, run_cli

pytestmark = pytest.mark.gpu


def test_cli_device_build_writes_the_same_image(device, tmp_path):
    a, b = tmp_path / "host.ppm", tmp_path / "device.ppm"
    ia = run_cli("--size", "128x96", "--spp", 12, "--bounces", 4, "--seed", 2, "--out", a)
    ib = run_cli("--size", "128x96", "--spp", 12, "--bounces", 4, "--seed", 2, "--out", b,
                 "--device-build")
    assert ia["rays"] == ib["rays"] > 0
    assert np.array_equal(read_ppm(a), read_ppm(b))
