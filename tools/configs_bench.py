"""Layer-API throughput of BASELINE.json configs[0] and configs[1] on one B200 (development aid; the parity of these
configurations is tested in tests/test_gpu_fullsize.py, the headline metric is bench.py)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np
import fbgnn as F
from fbgnn import _ffi
ctx = F.default_context()
out = {}

# configs[0]: [[882,24]] quaternary BP, batch 1000, 32 iterations, f = 0.625, depolarising p = 0.05, prior p0 = p
c882 = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27, 54, 0]), [0, 1, 6])
for B in (1000, 100000):
    nx, nz = F.Pauli(seed=0).sample_device(B, c882.N, F.pauli_thresholds(0.05))
    gx, gz = _ffi.Graph(c882.hx), _ffi.Graph(c882.hz)
    sx, sz = ctx.empty((gx.m, B), np.uint8), ctx.empty((gz.m, B), np.uint8)
    _ffi.call("fbgnn_syndrome", gx.handle, B, nz.t2(), sx.t2())
    _ffi.call("fbgnn_syndrome", gz.handle, B, nx.t2(), sz.t2())
    dec = F.QLDPCBPDecoder(c882, num_iter=32, normalization_factor=0.625, cn_type="boxplus-phi", stage_one=True)
    prior = float(np.log(3 * 0.95 / 0.05))
    dec.decode_device(None, sx, sz, prior=prior); ctx.sync()
    reps = 20 if B == 1000 else 3
    ctx.timer_start()
    for _ in range(reps):
        dec.decode_device(None, sx, sz, prior=prior)
    ms = ctx.timer_stop() / reps
    out[f"configs[0] [[882,24]] BP4 32 it, B={B}"] = {"ms_per_call": ms, "frames_per_s": B / ms * 1e3}

# configs[1]: [[1270,28]] binary syndrome BP on hx (Z errors) and hz (X errors), 64 iterations, B = 10^5, p0 = 0.2
c1270 = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                             [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7], name="GHP_n1270_k28")
B = 100000
for side, pcm, logical in (("hx", c1270.hx, c1270.hz_perp), ("hz", c1270.hz, c1270.hx_perp)):
    dec = F.LDPCBPDecoder(pcm, is_syndrome=True, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi")
    model = F.BP_BSC_Model(pcm, dec, logical_pcm=logical, p0=0.2)
    model.run(B, 0.04); ctx.sync()
    ctx.timer_start()
    r = model.run(B, 0.04)
    ms = ctx.timer_stop()
    out[f"configs[1] [[1270,28]] binary BP on {side}, 64 it, B={B}, p_b=0.04"] = {"ms_per_call": ms, "frames_per_s": B / ms * 1e3}
print(json.dumps(out, indent=1))
