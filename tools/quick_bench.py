"""Quick device timing of the fused pipeline (development aid; bench.py is the contract)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np
import fbgnn as F

def make(code_name, nG, skip=False):
    if code_name == "c1270":
        code = F.create_QC_GHP_codes(127, np.array([[0,-1,51,52,-1],[-1,0,-1,111,20],[0,-1,98,-1,122],[0,80,-1,119,-1],[-1,0,5,-1,106]]), [0,1,7], name="GHP_n1270_k28")
        wf = "feedback_GNN_n1270_k28_wt_10_80_iter_64_16_mixed.npy"
    else:
        code = F.create_QC_GHP_codes(63, F.create_cyclic_permuting_matrix(7, [27,54,0]), [0,1,6])
        wf = "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"
    G = F.Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean", activation="tanh", use_bias=True,
                       gemm=os.environ.get("FBGNN_LAB_GNN_GEMM", "fma"))
    F.load_weights(G, os.path.join(F.WEIGHTS_DIR, wf))
    d1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    d2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
    return F.Sandwich_BP_GNN_Evaluation_Model(code, [d1]+[d2]*nG, [G]*nG, num_layers=nG+1, skip_inactive=skip)

if __name__ == "__main__":
    ctx = F.default_context()
    print("device", ctx.name, ctx.num_sms, "SMs")
    print("sfu peak GTE/s", ctx.sfu_peak()/1e9, "fma peak Ginstr/s", ctx.fma_peak()/1e9)
    for code_name, nG, skip, B, p in [("c1270",3,False,8192,0.10),("c1270",1,False,8192,0.10),("c1270",3,True,8192,0.10),("c1270",0,False,8192,0.10),("c882",5,False,8192,0.06)]:
        m = make(code_name, nG, skip)
        m.run(B, p, want_counters=True)
        ctx.timer_start()
        reps = 3
        for _ in range(reps):
            r = m.run(B, p, want_flags=False, want_diff=False)
        ms = ctx.timer_stop()
        r = m.run(B, p, want_counters=True)
        print(json.dumps(dict(code=code_name, nG=nG, skip=skip, B=B, p=p, ms_per_batch=ms/reps, frames_per_s=B*reps/(ms*1e-3), counters=r["counters"].tolist())))
