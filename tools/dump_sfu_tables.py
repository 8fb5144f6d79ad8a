#!/usr/bin/env python
"""Measure the special-function unit of the GPU on the finite input sets the SFU arithmetic uses
(csrc/fb_math.h) and store the result as the tables the CPU oracle evaluates MUFU.EX2 / MUFU.LG2 with.

    python tools/dump_sfu_tables.py            (needs a GPU; writes tests/golden/sfu_b200_{ex2,lg2,lg2b,rcp}.xz)

Each file holds, lzma-compressed, one little-endian int32 per table entry: the difference between the bit
pattern the hardware returned and the bit pattern of a reference value that any IEEE-754 machine reproduces
exactly (a fixed float64 series, `oracle/sfu_tables.py`).  The hardware is within a few ulp of the reference,
so the differences are tiny integers and the files are small; `oracle/sfu_tables.py` adds them back.
This script is a fixture generator: it runs the raw hardware functions through the C ABI's math probe and
never touches the decoder."""
import lzma
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, ROOT)


def probe(ctx, fn, x):
    from fbgnn import _ffi
    out = np.empty_like(x)
    step = 1 << 22
    for i in range(0, x.size, step):
        dx = ctx.asarray(x[i:i + step])
        dy = ctx.empty(dx.shape, np.float32)
        _ffi.call("fbgnn_math_probe", ctx.handle, fn.encode(), dx.ptr, dy.ptr, dx.size)
        out[i:i + step] = dy.numpy()
    return out


def main():
    import fbgnn as F
    from oracle import sfu_tables as T
    ctx = F.default_context()
    outdir = os.path.join(ROOT, "tests", "golden")
    for name, inputs, ref, fn in (("ex2", T.ex2_inputs(), T.ex2_reference(), "mufu_ex2"),
                                  ("lg2", T.lg2_inputs(), T.lg2_reference(), "mufu_lg2"),
                                  ("lg2b", T.lg2b_inputs(), T.lg2b_reference(), "mufu_lg2"),
                                  ("rcp", T.rcp_inputs(), T.rcp_reference(), "mufu_rcp")):
        hw = probe(ctx, fn, inputs)
        delta = hw.view(np.int32).astype(np.int64) - ref.view(np.int32).astype(np.int64)
        assert np.abs(delta).max() < 2 ** 31
        print(name, "entries", hw.size, "max |delta| (ulp)", int(np.abs(delta).max()),
              "mean |delta|", float(np.abs(delta).mean()), "device", ctx.name)
        blob = lzma.compress(delta.astype("<i4").tobytes(), preset=9 | lzma.PRESET_EXTREME)
        path = os.path.join(outdir, f"sfu_b200_{name}.xz")
        with open(path, "wb") as f:
            f.write(blob)
        print("wrote", path, len(blob), "bytes")
        # the second run must give the same bits (the MUFU is a pure function)
        assert np.array_equal(probe(ctx, fn, inputs).view(np.int32), hw.view(np.int32))


if __name__ == "__main__":
    main()
