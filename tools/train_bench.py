"""Time of one training iteration of examples/Feedback_GNN.ipynb cell 2 on one B200: [[1270,28]], batch 100,
BP4(64) first stage, feedback GNN + BP4(16, stage_two), loss_from 8, Adam (development aid)."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "feedback-gnn_b200"))
import numpy as np
import fbgnn as F
code = F.create_QC_GHP_codes(127, np.array([[0, -1, 51, 52, -1], [-1, 0, -1, 111, 20], [0, -1, 98, -1, 122],
                                            [0, 80, -1, 119, -1], [-1, 0, 5, -1, 106]]), [0, 1, 7], name="GHP_n1270_k28")
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 100
dec1 = F.QLDPCBPDecoder(code, num_iter=64, normalization_factor=1.0, cn_type="boxplus-phi", stage_one=True)
dec2 = F.QLDPCBPDecoder(code, num_iter=16, normalization_factor=1.0, cn_type="boxplus-phi", stage_two=True)
G = F.Feedback_GNN(code, 20, 40, 2, "mean", "tanh", True)
m1, m2 = F.First_Stage_BP_Model(code, dec1), F.Second_Stage_GNN_BP_Model(code, G, dec2, num_iter=16)
gen = F.BP4_Error_Model(code, dec1, wt=True, seed=5)
t0 = time.time()
xs, zs = [], []
for wt in range(40, 61, 4):
    x, z = gen(20000, wt)
    xs.append(x); zs.append(z)
x, z = np.vstack(xs), np.vstack(zs)
t_gen = time.time() - t0
opt = F.Adam(learning_rate=2e-4)
rng = np.random.default_rng(0)
perm = rng.permutation(len(x))
losses, t_step = [], []
for it in range(min(60, len(x) // bs)):
    idx = perm[it * bs:(it + 1) * bs]
    t1 = time.time()
    loss, bler, fb = F.train_step(m1, m2, opt, x[idx], z[idx])
    t_step.append(time.time() - t1)
    losses.append(loss)
print(json.dumps({"config": "[[1270,28]] BP64 -> GNN -> BP16 stage_two, batch %d" % bs, "dataset_strings": int(len(x)),
                  "dataset_generation_s": t_gen, "ms_per_training_iteration_median": 1e3 * float(np.median(t_step[3:])),
                  "loss_first5_mean": float(np.mean(losses[:5])), "loss_last5_mean": float(np.mean(losses[-5:]))}))
