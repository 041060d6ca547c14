"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/rsrcu.h declares; the
packed-stream encoder of the Python mirror produces well-formed records; without a CUDA device the
product fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

import rsr_b200 as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "rsrcu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rsrcu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = R.load_library()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"librsrcu.so does not export {n}"
    assert sorted(R.EXPORTED_SYMBOLS) == names


def test_state_struct_layout_matches_header():
    # 4+1 floats, 19 int32, 48 floats, 1 uint32, 32 floats
    assert ctypes.sizeof(R.RsrState) == (5 + 19 + 48 + 1 + 32) * 4


def test_no_device_means_error_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(R.RsrError) as e:
        R.GPU(0)
    assert e.value.code == 1 and "no CPU fallback" in str(e.value)


class Recorder(R.GPU):
    """the recording half of rsr_b200.GPU without a device context"""
    def __init__(self):
        self.direct = False
        self._keep, self._rec = [], bytearray()
        self._state = R.RsrState()
        self._reset_state()
        self._dirty, self.size, self.h = True, (0, 0), None


def parse(data):
    out, p = [], 0
    while p < len(data):
        op, size = struct.unpack_from("<II", data, p)
        assert size >= 8 and size % 8 == 0 and p + size <= len(data)
        out.append((op, data[p + 8:p + size]))
        p += size
    return out


def test_packed_stream_records():
    from rsr_b200.scenes import WavyGridScene
    g = Recorder()
    sc = WavyGridScene(n=8, tex_dim=16)
    out = np.zeros((90, 160), np.uint32)
    sc.record(g, (160, 90), out)
    rec = g.Finish()
    ops = [op for op, _ in parse(rec.data)]
    assert ops[0] == R.OP_BEGIN_FRAME and ops[-1] == R.OP_END_FRAME
    assert ops.count(R.OP_DRAW_ELEMENTS) == 1 and ops.count(R.OP_CLEAR) == 1 and ops.count(R.OP_STORE_TC) == 1
    # state snapshots only when something changed before a command (GL::MaybeUpdateState)
    assert ops.count(R.OP_STATE) == 3
    for op, payload in parse(rec.data):
        if op == R.OP_STATE:
            assert len(payload) == (ctypes.sizeof(R.RsrState) + 7) // 8 * 8   # padded to 8
        if op == R.OP_DRAW_ELEMENTS:
            count, hint, inst, upload, ptr = struct.unpack("<iiiiQ", payload)
            assert count == len(sc.idx) and ptr == rec.keep[-2].ctypes.data or ptr != 0
        if op == R.OP_STORE_TC:
            gamma, w, h, stride, ptr = struct.unpack("<iiiiQ", payload)
            assert (gamma, w, h, stride, ptr) == (1, 160, 90, 160, out.ctypes.data)


def test_matrix_convention_column_major():
    m = np.arange(16, dtype=np.float32).reshape(4, 4)
    assert list(R.GPU._mat(m)) == list(m.T.reshape(16))
