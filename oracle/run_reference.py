"""ORACLE (test infrastructure only): import and run the UNMODIFIED reference sources from /root/reference
without TensorFlow.

The hot-path files of the reference (``sionna/fec/ldpc/{decoding_q,feedback_gnn,gnn,codes_q}.py``,
``sionna/channel/pauli.py``, ``sionna/fec/utils.py``, ``sionna/utils/metrics.py``) are loaded from where they lie,
one by one, into a skeleton ``sionna`` package (the real package ``__init__`` files pull in the whole wireless
simulator, Mitsuba and matplotlib).  ``tensorflow`` is the numpy stand-in of ``oracle/tfshim``.  What executes is
the reference's own code; see tfshim's docstring for what that does and does not pin.

Used only by ``tests/golden/make_reference_golden.py`` (in this container, where /root/reference exists) to
write ``tests/golden/ref_*.npz``; the tests read those files and never need the reference itself."""
import importlib.util
import os
import sys
import types

REF = os.environ.get("FBGNN_REFERENCE", "/root/reference")
_loaded = {}


def available():
    return os.path.isdir(os.path.join(REF, "sionna", "fec", "ldpc"))


def _pkg(name):
    m = sys.modules.get(name)
    if m is None:
        m = types.ModuleType(name)
        m.__path__ = []                       # a package, but with nothing to discover on disk
        sys.modules[name] = m
        if "." in name:
            parent, _, leaf = name.rpartition(".")
            setattr(_pkg(parent), leaf, m)
    return m


def _load(modname, relpath):
    """Execute one reference file as module ``modname``."""
    if modname in _loaded:
        return _loaded[modname]
    path = os.path.join(REF, relpath)
    spec = importlib.util.spec_from_file_location(modname, path)
    mod = importlib.util.module_from_spec(spec)
    parent, _, leaf = modname.rpartition(".")
    mod.__package__ = parent
    sys.modules[modname] = mod
    setattr(_pkg(parent), leaf, mod)
    spec.loader.exec_module(mod)
    _loaded[modname] = mod
    return mod


def load():
    """Returns a namespace with the reference's classes / functions of the hot path."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference sources not found under {REF}")
    if "tensorflow" in sys.modules and not getattr(sys.modules["tensorflow"], "__fbgnn_shim__", False):
        raise RuntimeError("a real TensorFlow is loaded; run the reference with it instead of the shim")
    from . import tfshim
    tf = tfshim.install()
    tf.__fbgnn_shim__ = True

    # third-party modules the reference imports at module level but never uses on this path
    plt = types.ModuleType("matplotlib.pyplot")
    for fn in ("grid", "title", "figure", "plot", "show", "xlabel", "ylabel", "legend", "savefig", "semilogy",
               "subplots", "xticks", "yticks", "ylim", "xlim", "tight_layout", "close"):
        setattr(plt, fn, lambda *a, **k: None)
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = plt
    sys.modules.setdefault("matplotlib", mpl)
    sys.modules.setdefault("matplotlib.pyplot", plt)
    ir = types.ModuleType("importlib_resources")
    ir.files = ir.as_file = lambda *a, **k: None
    sys.modules.setdefault("importlib_resources", ir)

    # skeleton of the sionna package: only what the hot-path files import by name
    sn = _pkg("sionna")
    for p in ("sionna.fec", "sionna.fec.ldpc", "sionna.utils", "sionna.channel", "sionna.nr", "sionna.mapping",
              "sionna.signal"):
        _pkg(p)
    _pkg("sionna.fec.ldpc").codes = types.ModuleType("sionna.fec.ldpc.codes")
    sys.modules["sionna.fec.ldpc.codes"] = _pkg("sionna.fec.ldpc").codes
    nru = types.ModuleType("sionna.nr.utils")
    nru.generate_prng_seq = lambda *a, **k: None
    sys.modules["sionna.nr.utils"] = nru
    _pkg("sionna.nr").utils = nru
    enc = types.ModuleType("sionna.fec.ldpc.encoding")
    enc.LDPC5GEncoder = type("LDPC5GEncoder", (), {})
    sys.modules["sionna.fec.ldpc.encoding"] = enc

    utils = _pkg("sionna.utils")
    metrics = _load("sionna.utils.metrics", "sionna/utils/metrics.py")
    for k in ("compute_ber", "compute_bler", "count_errors", "count_block_errors"):
        setattr(utils, k, getattr(metrics, k))
    tensors = _load("sionna.utils.tensors", "sionna/utils/tensors.py")
    utils.expand_to_rank = tensors.expand_to_rank
    utils.log2 = lambda x: tf.math.log(x) / tf.math.log(2.0)
    # BinarySource is constructed by the models and never called
    utils.BinarySource = type("BinarySource", (tf.keras.layers.Layer,), {"call": lambda self, s: None})

    fec_utils = _load("sionna.fec.utils", "sionna/fec/utils.py")
    pauli = _load("sionna.channel.pauli", "sionna/channel/pauli.py")
    ch = _pkg("sionna.channel")
    ch.Pauli = pauli.Pauli
    ch.BinarySymmetricChannel = type("BinarySymmetricChannel", (tf.keras.layers.Layer,), {})

    codes_q = _load("sionna.fec.ldpc.codes_q", "sionna/fec/ldpc/codes_q.py")
    gnn = _load("sionna.fec.ldpc.gnn", "sionna/fec/ldpc/gnn.py")
    decoding_q = _load("sionna.fec.ldpc.decoding_q", "sionna/fec/ldpc/decoding_q.py")
    feedback_gnn = _load("sionna.fec.ldpc.feedback_gnn", "sionna/fec/ldpc/feedback_gnn.py")

    # The binary decoder assumes scipy.sparse.find returns the entries column-major (true up to scipy 1.10;
    # decoding_q.py carries an argsort "fix for scipy>=1.11", decoding.py does not -- SURVEY.md F8).  It is run here in
    # the scipy environment it was written for: its module-level `sp` sees a find() with the old ordering.
    import numpy as np
    import scipy as real_sp
    decoding = _load("sionna.fec.ldpc.decoding", "sionna/fec/ldpc/decoding.py")

    def find_column_major(a):
        i, j, v = real_sp.sparse.find(a)
        order = np.lexsort((i, j))
        return i[order], j[order], v[order]
    sparse_proxy = types.SimpleNamespace(**{k: getattr(real_sp.sparse, k) for k in ("csr_matrix", "issparse", "csc_matrix")},
                                         find=find_column_major)
    decoding.sp = types.SimpleNamespace(sparse=sparse_proxy)

    ns = types.SimpleNamespace(tf=tf, codes_q=codes_q, decoding=decoding, LDPCBPDecoder=decoding.LDPCBPDecoder, fec_utils=fec_utils, pauli=pauli, gnn=gnn, decoding_q=decoding_q,
                               feedback_gnn=feedback_gnn, metrics=metrics,
                               QLDPCBPDecoder=decoding_q.QLDPCBPDecoder, Feedback_GNN=feedback_gnn.Feedback_GNN,
                               Sandwich_BP_GNN_Evaluation_Model=feedback_gnn.Sandwich_BP_GNN_Evaluation_Model,
                               Pauli=pauli.Pauli, load_weights=gnn.load_weights, css_code=codes_q.css_code)
    _loaded["ns"] = ns
    return ns
