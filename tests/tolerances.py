"""Parity thresholds of the GPU suite, in one place.

north_star states: "rel 1e-3 for bf16 compute versus the fp32 reference, and 1e-5 in an fp32 validation mode".
What each entry claims:

* fp32 validation mode: asserted at the 1e-5 class (a few 1e-6 observed; sums over >1e5 pairs reach 1e-5).  MEETS north_star.
* bf16 mode, scalar losses (InfoNCE / ConR / SupCon / the step loss): asserted <= 1e-3.  MEETS north_star.
* bf16 mode, elementwise tensors (embeddings, pair tensor, gradients): every stored activation is rounded to bf16, whose
  unit round-off is 2^-9 = 1.95e-3, so a max-norm error below 1e-3 is impossible for ANY bf16-storing implementation
  (PyTorch autocast included).  These are asserted in two senses, each at <= 3x the value observed on the B200
  (gpurun_out/test_report.txt of the round-2 runs; the observed values are quoted beside each bound):
    *_norm  ||a - b|| / ||b||       -- 2e-3 .. 7e-3: DOES NOT MEET 1e-3, bounded by bf16 storage (one rounding = 1.1e-3 rms)
    *_max   max|a - b| / max|b|     -- 4e-3 .. 2e-2 after 15 layers
  "bit-exact" quantities (masks, bin indices, -inf pattern, featurisation) are asserted with torch.equal in their tests.
"""

TOL = {
    # tests/test_gpu_encoder_15l.py -- 15 layers, L = 66, production geometry.  Observed (round 2, B200), bf16 mode, all
    # three pair dtypes within 10 % of each other (the pair dtype is NOT what limits the 15-add pair chain):
    #   'init' weights (std 0.02): rep 7.9e-3 max / 4.2e-3 norm, residual stream by layer 1.0e-3 -> 3.4e-3, pair tensor by
    #           layer 5.0e-3 -> 7.2e-3, gradients <= 1.4e-2
    #   'wide' weights (std 0.05): rep 1.2e-2 / 1.1e-2, residual stream 4.9e-3 -> 1.1e-2, pair 5.0e-3 -> 1.3e-2, gradients
    #           <= 4.1e-2 except the cancellation-dominated gbf.* sums (7.6e-2; the fp32 REFERENCE itself is 1.6e-4 away
    #           from its float64 evaluation there, a 1e3 amplification of the unit round-off)
    # fp32 mode: compared against the float64 truth, bounded by the reference's own fp32 error (see the test).
    "enc15.fp32": dict(rep_max=6e-5, rep_norm=1e-5, x_layer_norm=1e-5, pair_layer_norm=1e-5),
    "enc15.bf16.init": dict(rep_max=2e-2, rep_norm=1.2e-2, x_layer_norm=1e-2, pair_layer_norm=2e-2, grad_norm=4e-2, grad_max=4e-2),
    "enc15.bf16.wide": dict(rep_max=3.5e-2, rep_norm=3e-2, x_layer_norm=3e-2, pair_layer_norm=3.5e-2, grad_norm=0.2, grad_max=0.25),
    # tests/test_gpu_pair_bias.py -- K1, fixtures and bench sizes (multi-tile persistent loops).  Observed: fp32 out 2.2e-6,
    # grads 5.7e-6; bf16 out 8.8e-3, grads 8e-3 (one outlier 2.4e-2: gbf.means at L = 258, 'pre' weights)
    "k1.fp32": dict(out=1e-5, grad=3e-5),
    "k1.bf16": dict(out=2.5e-2, grad=6e-2),
    # tests/test_gpu_encoder.py -- reference-made fixtures, 2-3 layers.  Observed: fp32 1.1e-6 (toy dims) / 3.6e-6 (production
    # dims); bf16 1.6e-2 / 2.6e-2 (max-norm over outputs and every stored gradient)
    "enc.fp32": 1e-5, "enc.bf16": 4e-2, "slice.fp32": 2e-5, "slice.bf16": 6e-2,
    # tests/test_gpu_pair_attn.py -- K2 vs the float64 oracle.  Observed: fp32 3e-7; bf16 o 1.04e-2, scores 3.1e-3, grads 4.9e-3
    "k2.fp32": dict(o=2e-6, s=2e-6, g=2e-6), "k2.bf16": dict(o=2.5e-2, s=8e-3, g=1.5e-2),
    # tests/test_gpu_contrastive.py -- losses / gradients (norm sense).  Observed: fp32 4e-7 / 1.7e-6; bf16 loss 7.2e-4
    # (MEETS north_star's 1e-3), gradients 4.4e-3 (bf16 operand rounding 2^-9 on both GEMMs: does not meet 1e-3)
    "sim.fp32": (5e-6, 1e-5), "sim.bf16": (1e-3, 1.2e-2),
    # tests/test_gpu_cross_modal.py -- f2 cross-modal fusion.  attention core vs float64 (max-norm), the reference-made fixture
    # (outputs max-norm, gradients norm sense), one layer with the three dropouts replayed.
    "cross.attn.fp32": 1e-5, "cross.attn.bf16": 3e-2,
    "cross.golden.fp32": dict(out=1e-5, grad=3e-5), "cross.golden.bf16": dict(out=3e-2, grad=4e-2),
    "cross.layer.fp32": 3e-5, "cross.layer.bf16": 4e-2,
    # tests/test_gpu_chemberta.py -- f4, 2-layer 512-d RoBERTa vs Hugging Face's own module (output max-norm, gradients norm sense)
    "chemberta.fp32": dict(out=1e-5, grad=3e-5), "chemberta.bf16": dict(out=3e-2, grad=4e-2),
    # tests/test_gpu_hot_path_step.py -- the whole step.  Observed: fp32 loss 0 / grads 8e-6; bf16 loss 2.2e-4, grads 9.7e-3
    "step.fp32": (1e-5, 3e-5), "step.bf16": (1e-3, 3e-2),
}
