"""helpers: render one scene through the reference oracle and through the CUDA path, compare"""
import numpy as np


def render_both(scene, size, ref_gpu, cuda_gpu, **kw):
    w, h = size
    outs = {}
    for name, gl in (("ref", ref_gpu), ("cuda", cuda_gpu)):
        color = np.zeros((h, w), np.uint32)
        depth = np.zeros((h, w), np.float32) if kw.get("with_depth", False) else None
        args = {k: v for k, v in kw.items() if k != "with_depth"}
        scene.record(gl, size, color, depth, **args)
        gl.Run()
        outs[name] = (color, depth)
    return outs


def compare(outs):
    """returns (differing pixels, max 8-bit channel error, differing depth values)"""
    (rc, rd), (cc, cd) = outs["ref"], outs["cuda"]
    diff = int(np.count_nonzero(rc != cc))
    ch = lambda a, s: ((a >> s) & 0xff).astype(np.int32)
    maxerr = max(int(np.abs(ch(rc, s) - ch(cc, s)).max()) for s in (0, 8, 16))
    ddepth = 0
    if rd is not None:
        ddepth = int(np.count_nonzero(rd.view(np.uint32) != cd.view(np.uint32)))
    return diff, maxerr, ddepth
