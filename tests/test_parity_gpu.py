"""-m gpu: CUDA path vs the compiled reference, bit for bit (colour tolerance stated: 0 LSB asked,
<= 1 LSB allowed by north_star; coverage/depth: zero differing pixels)."""
import numpy as np
import pytest

from parity import compare, render_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
@pytest.mark.parametrize("bilinear", [True, False])
def test_wavy_grid(size, bilinear, ref_gpu, cuda_gpu):
    from rsr_b200.scenes import WavyGridScene
    scene = WavyGridScene(n=40, bilinear=bilinear)
    outs = render_both(scene, size, ref_gpu, cuda_gpu)
    diff, maxerr, _ = compare(outs)
    assert np.unique(outs["ref"][0]).size > 1000
    assert diff == 0 and maxerr == 0, f"{diff} differing pixels, max channel error {maxerr}"


@pytest.mark.parametrize("size", [(640, 360), (1920, 1080)])
def test_cubes(size, ref_gpu, cuda_gpu):
    from rsr_b200.scenes import CubesScene
    scene = CubesScene(instances=300)
    outs = render_both(scene, size, ref_gpu, cuda_gpu)
    diff, maxerr, _ = compare(outs)
    assert diff == 0 and maxerr == 0, f"{diff} differing pixels, max channel error {maxerr}"
