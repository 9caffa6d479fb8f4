"""CPU oracle for the volumetric hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic of the reference's hot path
(QingYunA/General-Medical-Image-Segmentation-CNN-Framework) so the CUDA kernels can be checked against it:

    ops.py      conv / conv-transpose / batch-norm / instance-norm / max-pool / activations
                (the reference delegates these to torch.nn; call sites models/three_d/unet3d.py:73-104)
    unet3d.py   functional 3D U-Net forward driven by a reference-format state_dict (unet3d.py:50-71)
    losses.py   cross_entropy_3D, DiceLoss, DiceLossss, BinaryDiceLoss, BCE (utils/loss_function.py)
    metric.py   precision / recall / jaccard / dice counts (utils/metric.py:20-75)
    syncbn.py   cross-replica batch-norm statistics (models/sync_batchnorm/batchnorm.py:48-125)
    window.py   sliding-window sampler / aggregator (torchio 0.20.3 semantics; predict.py:100-147)

Pinning: the reference ships no tests or golden vectors, so each restatement is pinned against
outputs of the reference's own modules imported from /root/reference in the build container; the generating script
and the vectors are committed under tests/golden/.  window.py restates torchio, which is neither vendored in the
reference nor installed here: that one function family is "parity unpinned" and says so in its header.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product package never does: it fails loudly when the CUDA extension is missing.
"""
