"""CPU oracle for the speech-lid front-end hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker or the timed CPU baseline.
The product package (``speech-lid_b200/``) never imports it and has no CPU fallback.

Parity status: the reference (kouyt5/speech-lid) ships no tests, golden vectors or
fixtures for this path.  The restatement in
``oracle/frontend_oracle.py`` is therefore pinned against outputs of the reference
itself, executed in the build container (``lid/audio_processor.py`` imported
unmodified from ``/root/reference`` on top of torchaudio 2.11.0; the reference pins
torchaudio 0.12.1, whose equivalence cannot be verified offline).  Those outputs are
committed as ``tests/golden/*.npz`` together with the generating script
``tests/golden/make_golden.py``.  Rows that do not exist in the reference (MFCC, CMVN)
are pinned against ``torchaudio.compliance.kaldi.mfcc`` (MFCC) or are OUR definition
and say "parity unpinned" (CMVN).
"""
