"""CPU oracle for the DCGAN/CGAN train step -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU in fp32, the algorithm of the reference's hot
path (model/DCGAN.py, model/CGAN.py, train/dcgan_trainer.py:155-189,
train/cgan_trainer.py:173-213).  It is the *checker*:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
    ``--impl reference`` legs may import it;
  * nothing under ``jck_generation_b200/`` imports it, and the product path raises
    when the CUDA library is missing instead of falling back to anything here.

Where the arithmetic lives: the reference holds no arithmetic of its own -- every
operation is a call into PyTorch (third-party, not vendored, *unpinned* by the
reference: it has no requirements file).  The oracle therefore calls the same
PyTorch CPU operators (this image: torch 2.11.0+cu128, oneDNN/MKL CPU kernels)
in the same order, with the literals the reference hard-codes (3 image channels,
100 classes, 200-wide label embedding) lifted to keyword arguments so the
BASELINE.json configs the reference cannot express (1x64x64, 10 classes) have a
checker as well.

Pinning status: the reference ships NO tests, golden vectors or fixtures for this
path ("parity unpinned" by the reference itself).  The oracle is
pinned instead against *outputs of the reference itself run in the build
container*: ``oracle/ref_harness.py`` drives the unmodified
``/root/reference/train/dcgan_trainer.py`` / ``cgan_trainer.py`` step loops (stubbing
only the absent ``torchinfo`` / ``matplotlib`` imports, the CIFAR download and the
Inception checkpoint) with replayed random tensors, and ``oracle/make_golden.py``
freezes digests of its losses, per-layer activations/gradients and post-step
weights into ``tests/golden/*.json``.  ``tests/test_oracle_golden.py`` checks the
oracle against those digests everywhere, and bit-for-bit against the live
reference wherever ``/root/reference`` exists.
"""
