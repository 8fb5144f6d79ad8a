"""Generates tests/golden/ref_exit.npz: the EXIT trajectory (``track_exit=True``: ``ie_v`` / ``ie_c``,
decoding.py:955-1000, fec/utils.py:151-218) of the reference's own LDPCBPDecoder, executed unmodified under
oracle/tfshim like tests/golden/make_reference_golden.py does.

    python tests/golden/make_reference_golden_exit.py     (only where /root/reference exists)
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import run_reference as R                      # noqa: E402


def main():
    ns = R.load()
    tf, Q = ns.tf, ns.codes_q
    code = Q.create_QC_GHP_codes(63, Q.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
    B = 16
    rng = np.random.default_rng(3)
    # all-zero codeword over a BI-AWGN-like channel: logits (log p1/p0 in this decoder's input convention) mostly negative
    llr = (-2.0 + 1.5 * rng.standard_normal((B, code.N))).astype(np.float32)
    out = {"llr": llr}
    for cn_type in ("boxplus-phi", "minsum"):
        dec = ns.LDPCBPDecoder(code.hx, track_exit=True, num_iter=6, normalization_factor=0.9, cn_type=cn_type,
                               hard_out=False)
        x = dec(tf.constant(llr))
        out[f"{cn_type}.soft"] = np.asarray(x)
        out[f"{cn_type}.ie_v"] = np.asarray(dec.ie_v)
        out[f"{cn_type}.ie_c"] = np.asarray(dec.ie_c)
        print(cn_type, "ie_v", np.asarray(dec.ie_v), "ie_c", np.asarray(dec.ie_c))
    np.savez_compressed(os.path.join(HERE, "ref_exit.npz"), **out)


if __name__ == "__main__":
    main()
