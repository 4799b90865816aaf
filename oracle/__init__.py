"""CPU oracle for the EEL-Unet hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and only as the checker or the timed CPU baseline --
never as a fallback for a CUDA kernel.

Contents
--------
ref_import.py     shims that import the real reference from /root/reference (build container only)
edge_np.py        numpy restatement of the cv2 integer edge-map pipeline (gray, Canny, Sobel, Laplacian)
edge_c.c          plain-C restatement of the same pipeline (compiled by oracle/Makefile -> oracle/_build/)
eelunet_torch.py  fp32 PyTorch-CPU restatement of EELUnet.forward and edge_BceDiceLoss
synth.py          re-export of eel_unet_b200/synth.py (deterministic synthetic "tooth-like" inputs)
resize_np.py      numpy restatement of Pillow's antialiased BILINEAR resize + ToTensor + Normalize (ToothDataset.py:58-61)
metrics_np.py     numpy restatement of evaluate.py's confusion counts / boundary F1

Parity pinning: the reference ships no tests or golden vectors, so the oracle is
pinned against outputs of the reference itself, run in the build container by
``tests/golden/make_golden.py`` (committed) -> ``tests/golden/*.npz`` (committed), and against
OpenCV 4.13 (third-party, the library the reference calls) for the integer edge maps.
"""
