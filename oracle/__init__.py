"""CPU oracle for the SuperPoint/MagicPoint inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline, never as the thing shipped.  The product path
(``feature-point-cnn_b200``) never imports this package and fails loudly when its
CUDA library is missing.

The oracle is a restatement, in plain torch-fp32-on-CPU / numpy / C, of the
reference's algorithm (Kolkir/feature-point-cnn):

* ``model.py``     - ``SuperPoint.forward``      (python/src/superpoint.py:91-115,
                     python/src/resnet_blocks.py:14-40)
* ``postproc.py``  - ``restore_prob_map`` / ``get_points`` / ``get_descriptors``
                     (python/src/netutils.py:56-121)
* ``nms_greedy.c`` - ``corners_nms``             (python/src/nms.py:4-53)
* ``weights.py``   - seeded synthetic checkpoints in the ``save_checkpoint`` format
                     (python/src/saveutils.py:54-63)

Parity pinning: the reference ships no golden vectors or asserting tests
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF,
generated in the build container by ``tests/golden/make_golden.py`` (which imports
``/root/reference/python/src``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every oracle function against them.
"""
