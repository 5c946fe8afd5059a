"""CPU oracle for the promonet hot path.

TEST INFRASTRUCTURE ONLY. Nothing under ``promonet_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
the CPU number printed beside the GPU one.

Every function here is a plain-torch / numpy / C restatement of the reference's
arithmetic, citing the reference ``file:line`` it follows.  Pinning status:

* ``oracle.hifigan`` / ``oracle.features`` / ``oracle.fargan`` / ``oracle.dsp``
  (spectrogram part) are pinned against the reference's own modules, imported
  unmodified from ``/root/reference`` through ``oracle.ref_shim`` in the build
  container; the generated vectors live in ``tests/golden`` together with the
  generating script ``oracle/make_golden.py``.
* ``oracle.train`` (discriminators, losses, one GAN training step with AdamW) is pinned
  against the reference's own modules composed as ``promonet/train/core.py:183-369``
  (``oracle/make_golden.py --train`` / ``--train-flags`` -> ``tests/golden/train*.npz``);
  ``oracle.features.grid_sample`` against ``promonet.edit.grid.sample`` (``grid.npz``).
* ``ppgs.sparsify``, the librosa pieces of ``oracle.dsp`` (A-weighting, dB,
  mel basis), ``oracle.penn`` and ``oracle.viterbi`` restate third-party
  packages that are absent from ``/root/reference`` (unpinned versions in
  ``setup.py:14-36``).  They are **parity unpinned**: restated from the
  published algorithms, cross-checked only where an independent implementation
  exists in this image (torchaudio mel filterbank, scipy hann window).
"""
