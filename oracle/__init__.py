"""CPU oracle for the VS_Seg hot path (TEST INFRASTRUCTURE ONLY).

Pure-torch fp32 restatement of the reference algorithm: the 2.5D attention
U-Net forward (``unet_oracle``), the hardness/attention Dice loss
(``loss_oracle``), MONAI-0.4.0 ``sliding_window_inference`` (``sw_oracle``) and
the foreground Dice metric.  Every function cites the reference file:line it
follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product code under
``vs_seg_b200/`` and ``params/`` never does.

Parity pinning: the reference ships no tests or golden vectors,
so the network and loss restatements are pinned against the reference's own
modules imported in the build container under ``oracle/monai_shim`` (see
``oracle/make_golden.py`` and ``tests/golden/``).  ``sliding_window_inference``
lives in MONAI 0.4.0, which is absent from /root/reference and not installable
offline: that function is restated from the published MONAI 0.4.0 algorithm and
is "parity unpinned" against MONAI itself.  It is frozen by
``tests/golden/sw_geometry.json`` (``oracle/make_sw_golden.py``) and checked in
``tests/test_oracle_sliding_window.py`` against window lists worked by hand, the closed-form
erf weights and the SURVEY.md §8a S1 probe numbers (window counts, w[0]/w[64], multiplicity
histogram of the 384x384x160 / 128^3 benchmark geometry).
"""
