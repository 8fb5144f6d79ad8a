"""Evaluate BP -> (feedback GNN -> BP) x 5 on the [[882,24]] GHP code.

Same command line and flow as the reference's ``n882.py`` (five rounds of feedback, ``-p`` physical
error rate, ``-id`` GPU); the only change is the import block: the layers come from
``fbgnn`` (CUDA, sm_100a) instead of ``sionna.fec.ldpc`` / TensorFlow.
"""
import os
import sys
import argparse

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "feedback-gnn_b200"))
import numpy as np

argParser = argparse.ArgumentParser()
argParser.add_argument("-p", "--p", help="Physical error rate p to simulate.")
argParser.add_argument("-id", "--gpu_id", help="GPU id", default="0")
argParser.add_argument("--batch_size", type=int, default=5000)
argParser.add_argument("--max_iter", type=int, default=100000)
args = argParser.parse_args()

nG = 5
p = float(args.p)
gpu_num = int(args.gpu_id)
os.environ["FBGNN_DEVICE"] = str(gpu_num)

from fbgnn import QLDPCBPDecoder, Feedback_GNN, load_weights, WEIGHTS_DIR
from fbgnn import Sandwich_BP_GNN_Evaluation_Model
from fbgnn import *
from fbgnn import PlotBER, device_count

print('Number of GPUs available :', device_count())
print('Only GPU number', gpu_num, 'used.')
print(f"Running for {nG} rounds of GNN feedback at p={p} on GPU {gpu_num}.")

GHP_n882_k24 = create_QC_GHP_codes(63, create_cyclic_permuting_matrix(7, [27,54,0]), [0,1,6]) # 18 <= d <= 24
code = GHP_n882_k24

ber_plot = PlotBER()

bs = args.batch_size
max_iter = args.max_iter

G = Feedback_GNN(code=code,
                 num_msg_dims=20,
                 num_hidden_units=40,
                 num_mlp_layers=2,
                 reduce_op="mean",
                 activation="tanh",
                 use_bias=True)
load_weights(G, os.path.join(WEIGHTS_DIR, "feedback_GNN_n882_k24_wt_4_60_iter_64_16_mixed.npy"))

num_iter1 = 64
num_iter2 = 16
factor1 = 1.0
factor2 = 1.0

decoder1 = QLDPCBPDecoder(code=code, num_iter=num_iter1, normalization_factor=factor1, cn_type="boxplus-phi", trainable=False, stage_one=True)
decoder2 = QLDPCBPDecoder(code=code, num_iter=num_iter2, normalization_factor=factor2, cn_type="boxplus-phi", trainable=False, stage_one=True)

# skip_inactive (extension): frames whose correction already matches the syndrome skip the later rounds.
# The reference masks those rounds' updates (feedback_gnn.py:339-340), so every output is identical.
model_eval = Sandwich_BP_GNN_Evaluation_Model(code, [decoder1]+[decoder2]*nG, [G]*nG, num_layers=(nG+1), skip_inactive=True)
ber_plot.simulate(model_eval,
              ebno_dbs=[p],
              batch_size=bs,
              num_target_block_errors=100,
              legend=f"feedback GNN {factor1:.2f} {nG} rounds",
              soft_estimates=True,
              max_mc_iter=max_iter,
              early_stop=True,
              add_bler=True,
              show_fig=False,
              qldpc=True,
              forward_keyboard_interrupt=False)

print(f"at {ber_plot._snrs[1]}, BLER is {ber_plot._bers[1]}")
