"""Prints the marching-cubes case table in the packed form rsr_b200/csrc/rsrcu.cu holds (kMcTriHost): one uint64 per
corner-sign case, 16 nibbles = the edge ids of up to five triangles, 0xF = end of list.  The table is Bourke's
polygonise table; it is read here from the compiled reference (oracle/_ref/librsr_ref.so, ref_mc_tables) so that the
packed copy is the reference's table by construction; tests/test_march_gpu.py checks the copy against it again."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def packed_table():
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "librsr_ref.so"))
    flags = np.zeros(256, np.int16)
    tri = np.zeros((256, 16), np.int8)
    conn = np.zeros((12, 2), np.uint8)
    lib.ref_mc_tables(flags.ctypes.data_as(C.c_void_p), tri.ctypes.data_as(C.c_void_p), conn.ctypes.data_as(C.c_void_p))
    out = []
    for row in tri:
        v = 0
        for k, e in enumerate(row):
            v |= (int(e) & 0xF if e >= 0 else 0xF) << (4 * k)
        out.append(v)
    return out, flags, conn


if __name__ == "__main__":
    rows, _, _ = packed_table()
    for i in range(0, 256, 4):
        print("\t" + " ".join(f"0x{v:016x}ull," for v in rows[i:i + 4]))
