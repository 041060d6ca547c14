"""helpers: render one scene through the reference oracle and through the CUDA path, compare"""
import numpy as np


def render_both(scene, size, ref_gpu, cuda_gpu, with_depth=False, with_fp=False, **kw):
    w, h = size
    outs = {}
    for name, gl in (("ref", ref_gpu), ("cuda", cuda_gpu)):
        color = np.zeros((h, w), np.uint32)
        depth = np.zeros((h, w), np.float32) if with_depth else None
        extra = {}
        fp = None
        if with_fp:
            fp = np.zeros((h, w, 4), np.float32)
            extra["fp_out"] = fp
        scene.record(gl, size, color, depth, **extra, **kw)
        gl.Run()
        outs[name] = (color, depth, fp)
    return outs


def compare(outs):
    """returns (differing pixels, max 8-bit channel error, differing depth values)"""
    (rc, rd, rf), (cc, cd, cf) = outs["ref"], outs["cuda"]
    diff = int(np.count_nonzero(rc != cc))
    ch = lambda a, s: ((a >> s) & 0xff).astype(np.int32)
    maxerr = max(int(np.abs(ch(rc, s) - ch(cc, s)).max()) for s in (0, 8, 16))
    ddepth = 0
    if rd is not None:
        ddepth = int(np.count_nonzero(rd.view(np.uint32) != cd.view(np.uint32)))
    if rf is not None:
        ddepth += int(np.count_nonzero(rf.view(np.uint32) != cf.view(np.uint32)))
    return diff, maxerr, ddepth


def assert_identical(outs, what=""):
    diff, maxerr, dd = compare(outs)
    assert diff == 0 and maxerr == 0 and dd == 0, \
        f"{what}: {diff} differing colour pixels (max channel error {maxerr} LSB), {dd} differing depth/float values"
