"""the reference's bundled data/scene/*.lua files drive BOTH renderers through the reference's own
front-end -- Lua scene -> JSON (its host.lua, the vendored Lua interpreter) -> node graph (src/viewer/compile.cxx,
src/viewer/node/*) -> rglv::GL -> rglv::GPU.  In oracle/_ref/librsr_ref.so the GPU is the CPU rasteriser, in
oracle/_ref/librsr_dropin.so GPU::RunImpl is the C-ABI binding in front of librsrcu.so (rsr_b200/host/rglv_gpu_cuda.cxx):
same scene file, same nodes, frames compared bit for bit."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZE = (640, 360)   # BASELINE.json configs[0]
# (particles.lua advances its simulation with a wall-clock timer, node/particles.cxx:49: not comparable frame to frame)
SCENES = ("colortest", "tucker-and-dino", "instanced-cubes", "render-to-texture", "sdf-polygonization-1",
          "oldschool", "plusrqdq", "auraforlaura", "writer")


@pytest.fixture(scope="module")
def scene_libs(refgl):
    if not refgl.scenes_available() or not refgl.dropin_available():
        pytest.skip("bundled scenes / drop-in library not built (oracle/build_ref.sh)")
    refgl.init_dropin()
    return refgl


@pytest.mark.parametrize("name", SCENES)
def test_bundled_scene_renders_identically_on_the_gpu(scene_libs, name):
    from oracle import scene_ref
    times = (0.0, 1.5)
    wants = scene_ref.frames(name, SIZE, times)   # (the CPU reference in a process of its own: see oracle/scene_ref.py)
    gpu = scene_libs.RefScene(name, dropin=True)
    try:
        for t, want in zip(times, wants):
            got = gpu.render(SIZE, t)
            assert len(np.unique(want)) > 20, "the reference frame is not blank"
            diff = int(np.count_nonzero(want != got))
            assert diff == 0, f"{name} at t={t}: {diff} of {want.size} pixels differ"
    finally:
        gpu.close()


def test_bundled_scene_at_1080p(scene_libs):
    """BASELINE.json configs[1]: the bundled scenes at 1920x1080"""
    from oracle import scene_ref
    for name in ("tucker-and-dino", "instanced-cubes"):
        want = scene_ref.frames(name, (1920, 1080), (0.75,))[0]
        gpu = scene_libs.RefScene(name, dropin=True)
        try:
            got = gpu.render((1920, 1080), 0.75)
            assert np.array_equal(want, got), f"{name}: {np.count_nonzero(want != got)} pixels differ"
        finally:
            gpu.close()
