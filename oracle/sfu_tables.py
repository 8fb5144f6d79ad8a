"""ORACLE (test infrastructure only): the GPU's special-function unit as tables.

The SFU arithmetic of the decoders (feedback-gnn_b200/csrc/fb_math.h) confines the inputs of MUFU.EX2, MUFU.LG2 and
MUFU.RCP to finite sets of float32 values:

    ex2  : the 12 582 913 consecutive float32 values u from 0.5 (bits 0x3F000000) up to 1.5
    lg2  : the 13 302 542 consecutive float32 values m from sqrt(1/2) (bits 0x3f3504f3) up to 2.0
    lg2b : the multiples k 2^-24 below sqrt(1/2), k = 1 .. 11 863 283 (what 1 - t can be for a float32 t)
    rcp  : the 2^23 + 1 consecutive float32 values q from 1.0 up to 2.0

``tests/golden/sfu_b200_{ex2,lg2,lg2b,rcp}.xz`` hold what a B200 returns on them (written by
tools/dump_sfu_tables.py as int32 differences from the reference values below); ``tables()`` rebuilds the float32
arrays the C oracle indexes.  The reference values are fixed float64 series evaluated with IEEE add / multiply /
divide only, so every machine rebuilds the same bits."""
import lzma
import os

import numpy as np

EX2_BASE, EX2_COUNT = 0x3F000000, 0x00C00000 + 1
LG2B_COUNT = 11863284
LG2_BASE, LG2_COUNT = 0x3f3504f3, 0x40000000 - 0x3f3504f3 + 1
RCP_BASE, RCP_COUNT = 0x3F800000, (1 << 23) + 1
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_cache = {}


def _from_bits(base, count):
    return (np.arange(count, dtype=np.int64) + base).astype(np.int32).view(np.float32)


def ex2_inputs():
    return _from_bits(EX2_BASE, EX2_COUNT)


def lg2_inputs():
    return _from_bits(LG2_BASE, LG2_COUNT)


def ex2_reference():
    """float32(2^u) = 2 * 2^(u - 1): Taylor series of exp((u - 1) ln 2) in float64, Horner, 22 terms (|.| < 0.35)."""
    z = (ex2_inputs().astype(np.float64) - 1.0) * 0.6931471805599453
    acc = np.ones_like(z)
    for k in range(22, 0, -1):
        acc = 1.0 + acc * z / k
    return (2.0 * acc).astype(np.float32)


def _log2_series(m):
    """log2 of m in [sqrt(1/2), sqrt(2)] (float64): 2/ln2 * atanh((m-1)/(m+1)) as an odd series, 24 terms."""
    s = (m - 1.0) / (m + 1.0)
    s2 = s * s
    acc = np.zeros_like(s)
    for k in range(23, -1, -1):
        acc = acc * s2 + 1.0 / (2 * k + 1)
    return acc * s * (2.0 / 0.6931471805599453)


def lg2b_inputs():
    """k 2^-24, k = 0 .. LG2B_COUNT - 1 (entry 0 is a placeholder equal to entry 1: lg2(0) is never looked up)."""
    k = np.arange(LG2B_COUNT, dtype=np.float64)
    k[0] = 1.0
    return (k * 2.0 ** -24).astype(np.float32)


def lg2b_reference():
    """float32(log2(k 2^-24)): exponent split off exactly (frexp), the mantissa through the same series as lg2."""
    m, e = np.frexp(lg2b_inputs().astype(np.float64))          # m in [0.5, 1)
    low = m < 0.7071067811865476
    m = np.where(low, m * 2.0, m)
    e = np.where(low, e - 1, e)
    return (_log2_series(m) + e).astype(np.float32)


def lg2_reference():
    """float32(log2 m): 2/ln2 * atanh((m-1)/(m+1)) as an odd series in float64, 24 terms (|s| <= 1/3)."""
    return _log2_series(lg2_inputs().astype(np.float64)).astype(np.float32)


def rcp_inputs():
    return _from_bits(RCP_BASE, RCP_COUNT)


def rcp_reference():
    """float32(1/q): the float64 quotient, correctly rounded twice (harmless: it is only a reference point)."""
    return (1.0 / rcp_inputs().astype(np.float64)).astype(np.float32)


def available():
    return all(os.path.exists(os.path.join(_GOLDEN, f"sfu_b200_{n}.xz")) for n in ("ex2", "lg2", "lg2b", "rcp"))


def _load(name, ref):
    with open(os.path.join(_GOLDEN, f"sfu_b200_{name}.xz"), "rb") as f:
        delta = np.frombuffer(lzma.decompress(f.read()), dtype="<i4")
    assert delta.size == ref.size, f"{name}: table has {delta.size} entries, expected {ref.size}"
    return np.ascontiguousarray((ref.view(np.int32) + delta).view(np.float32))


def tables():
    """(ex2_table, lg2_table, lg2b_table, rcp_table) float32 arrays as measured on the hardware."""
    if "t" not in _cache:
        _cache["t"] = (_load("ex2", ex2_reference()), _load("lg2", lg2_reference()), _load("lg2b", lg2b_reference()),
                       _load("rcp", rcp_reference()))
    return _cache["t"]
