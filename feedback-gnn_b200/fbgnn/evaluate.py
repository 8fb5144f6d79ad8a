"""Body shared by the evaluation scripts ``n1270.py`` / ``n882.py`` at the repository root: BP -> (feedback GNN -> BP) x nG
with the shipped weights, driven through ``PlotBER.simulate`` exactly as the reference's scripts drive it
(``n1270.py:41-82`` of the reference: Feedback_GNN(20, 40, 2, "mean", "tanh", True), a 64-iteration and a 16-iteration
``QLDPCBPDecoder`` with ``stage_one=True``, 100 target block errors, early stop)."""
import os

from .decoding_q import QLDPCBPDecoder
from .feedback_gnn import Feedback_GNN, Sandwich_BP_GNN_Evaluation_Model
from .gnn import WEIGHTS_DIR, load_weights
from .utils import PlotBER
from ._ffi import device_count


def evaluate_feedback_gnn(code, weights_file, nG, p, gpu_num=0, batch_size=5000, max_mc_iter=100000, num_iter1=64,
                          num_iter2=16, factor1=1.0, factor2=1.0, num_target_block_errors=100, gnn_gemm="fma"):
    """Run the Monte-Carlo evaluation and print the reference's progress table; returns the ``PlotBER`` object
    (``ber_plot._snrs[1]`` / ``ber_plot._bers[1]`` hold the point and its block-error rate).  ``gnn_gemm="tf32x3"`` (extension):
    the feedback GNN's dense products on the tensor cores."""
    print('Number of GPUs available :', device_count())
    print('Only GPU number', gpu_num, 'used.')
    print(f"Running for {nG} rounds of GNN feedback at p={p} on GPU {gpu_num}.")
    G = Feedback_GNN(code=code, num_msg_dims=20, num_hidden_units=40, num_mlp_layers=2, reduce_op="mean",
                     activation="tanh", use_bias=True, gemm=gnn_gemm)
    load_weights(G, os.path.join(WEIGHTS_DIR, weights_file))
    decoder1 = QLDPCBPDecoder(code=code, num_iter=num_iter1, normalization_factor=factor1, cn_type="boxplus-phi",
                              trainable=False, stage_one=True)
    decoder2 = QLDPCBPDecoder(code=code, num_iter=num_iter2, normalization_factor=factor2, cn_type="boxplus-phi",
                              trainable=False, stage_one=True)
    # skip_inactive (extension): frames whose correction already matches the syndrome skip the later rounds.  The
    # reference masks those rounds' updates (feedback_gnn.py:339-340), so every output is identical.
    model_eval = Sandwich_BP_GNN_Evaluation_Model(code, [decoder1] + [decoder2] * nG, [G] * nG, num_layers=nG + 1,
                                                  skip_inactive=True)
    ber_plot = PlotBER()
    ber_plot.simulate(model_eval, ebno_dbs=[p], batch_size=batch_size, num_target_block_errors=num_target_block_errors,
                      legend=f"feedback GNN {factor1:.2f} {nG} rounds", soft_estimates=True, max_mc_iter=max_mc_iter,
                      early_stop=True, add_bler=True, show_fig=False, qldpc=True, forward_keyboard_interrupt=False)
    print(f"at {ber_plot._snrs[1]}, BLER is {ber_plot._bers[1]}")
    return ber_plot
