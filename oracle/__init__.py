"""CPU oracle of DigiPathAI's ``getSegmentation`` hot path -- TEST INFRASTRUCTURE, not product code.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker / timed CPU baseline.  Nothing under ``digipathai_b200/`` imports it.

PARITY UNPINNED: the reference ships no tests, golden vectors, fixtures, sample slides or weights
(SURVEY.md F3/F6), and none of its numeric dependencies (TensorFlow 1.x, OpenSlide, scikit-image, pydensecrf)
exist in this image, so it cannot be executed here.  The restatement is pinned instead by (a) executing the
reference's own pure-numpy functions ``apply_tta`` / ``transform_prob`` (DigiPathAI/helpers/utils.py:487-522,
extracted at fixture-generation time by tests/golden/make_golden.py) and committing their outputs as golden
vectors, and (b) hand-computable known-answer cases for the stitch arithmetic.  Everything that lives inside
TensorFlow (conv / BN / pooling / softmax semantics) is restated from Keras' documented behaviour; (c) the
DenseNet-121 encoder of ``densenet_ref`` agrees stage by stage (2e-5) with an independent implementation present in
this image, ``torchvision.models.densenet121``, loaded with the same weights (tests/test_oracle_vs_torchvision.py).

Modules: ``pipeline_ref`` (numpy restatement of get_prediction / dataset / TTA / tissue mask), ``densenet_ref``,
``inception_ref``, ``deeplab_ref`` (fp32 PyTorch-CPU restatements of the three Keras graphs of DigiPathAI/models/),
``crf_ref`` (exact mean-field inference of the DenseCRF model post_process_crf configures; pydensecrf itself -- a
permutohedral-lattice approximation of the same model -- is not available: parity unpinned there too), ``lattice_ref``
(the permutohedral-lattice filter restated from the published algorithm), ``jpeg_ref`` (baseline-JPEG tile encoder of the
result files; pinned: its streams are decoded by libjpeg and compared with libjpeg's own encoder).
"""
