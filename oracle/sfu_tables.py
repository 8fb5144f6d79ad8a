"""ORACLE (test infrastructure only): the GPU's special-function unit as tables.

The SFU arithmetic of the decoders (feedback-gnn_b200/csrc/fb_math.h) confines the inputs of MUFU.EX2 and
MUFU.LG2 to two finite sets of float32 values:

    ex2 : w = u - 1.5 for the 2^23 + 8193 consecutive float32 values u from 1 - 2^-12 (bits 0x3F7FF000)
    lg2 : the 13 302 542 consecutive float32 values m from sqrt(1/2) (bits 0x3f3504f3) up to 2.0
    rcp : the 2^23 + 1 consecutive float32 values q from 1.0 up to 2.0

``tests/golden/sfu_b200_{ex2,lg2,rcp}.xz`` hold what a B200 returns on them (written by tools/dump_sfu_tables.py
as int32 differences from the reference values below); ``tables()`` rebuilds the two float32 arrays the C oracle
indexes.  The reference values are fixed float64 series evaluated with IEEE add / multiply / divide only, so
every machine rebuilds the same bits."""
import lzma
import os

import numpy as np

EX2_BASE, EX2_COUNT = 0x3F7FF000, (1 << 23) + 8193
LG2_BASE, LG2_COUNT = 0x3f3504f3, 0x40000000 - 0x3f3504f3 + 1
RCP_BASE, RCP_COUNT = 0x3F800000, (1 << 23) + 1
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_cache = {}


def _from_bits(base, count):
    return (np.arange(count, dtype=np.int64) + base).astype(np.int32).view(np.float32)


def ex2_inputs():
    """w = u - 1.5 (exact in float32) for every table entry."""
    return (_from_bits(EX2_BASE, EX2_COUNT) - np.float32(1.5)).astype(np.float32)


def lg2_inputs():
    return _from_bits(LG2_BASE, LG2_COUNT)


def ex2_reference():
    """float32(2^w): Taylor series of exp(w ln 2) in float64, Horner, 22 terms (|w ln 2| < 0.35)."""
    z = ex2_inputs().astype(np.float64) * 0.6931471805599453
    acc = np.ones_like(z)
    for k in range(22, 0, -1):
        acc = 1.0 + acc * z / k
    return acc.astype(np.float32)


def lg2_reference():
    """float32(log2 m): 2/ln2 * atanh((m-1)/(m+1)) as an odd series in float64, 24 terms (|s| <= 1/3)."""
    m = lg2_inputs().astype(np.float64)
    s = (m - 1.0) / (m + 1.0)
    s2 = s * s
    acc = np.zeros_like(s)
    for k in range(23, -1, -1):
        acc = acc * s2 + 1.0 / (2 * k + 1)
    return (acc * s * (2.0 / 0.6931471805599453)).astype(np.float32)


def rcp_inputs():
    return _from_bits(RCP_BASE, RCP_COUNT)


def rcp_reference():
    """float32(1/q): the float64 quotient, correctly rounded twice (harmless: it is only a reference point)."""
    return (1.0 / rcp_inputs().astype(np.float64)).astype(np.float32)


def available():
    return all(os.path.exists(os.path.join(_GOLDEN, f"sfu_b200_{n}.xz")) for n in ("ex2", "lg2", "rcp"))


def _load(name, ref):
    with open(os.path.join(_GOLDEN, f"sfu_b200_{name}.xz"), "rb") as f:
        delta = np.frombuffer(lzma.decompress(f.read()), dtype="<i4")
    assert delta.size == ref.size, f"{name}: table has {delta.size} entries, expected {ref.size}"
    return np.ascontiguousarray((ref.view(np.int32) + delta).view(np.float32))


def tables():
    """(ex2_table, lg2_table, rcp_table) float32 arrays as measured on the hardware."""
    if "t" not in _cache:
        _cache["t"] = (_load("ex2", ex2_reference()), _load("lg2", lg2_reference()), _load("rcp", rcp_reference()))
    return _cache["t"]
