"""CPU oracle for the ForwardTacotron / FastPitch generate path and DSP.wav_to_mel.

TEST INFRASTRUCTURE ONLY.  Nothing under ``forwardtacotron_b200/`` may import
this package: the product path is the sm_100a extension and fails loudly when
it is missing.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py`` (``cpu_baseline`` leg and ``--impl reference`` arm).

Parity status
-------------
* model path (``model_oracle``): the reference ships no test for generate();
  the restatement is pinned against the reference code itself, imported from
  ``/root/reference`` in the build container by ``oracle/make_golden.py`` and
  frozen as fixtures in ``tests/golden/``.
* DSP path (``dsp_oracle``): librosa==0.7.2 (requirements.txt:2) is not
  installed and the golden's input audio is not on disk, so numeric parity of
  ``wav_to_mel`` is pinned against torchaudio's Slaney mel + the shape / dtype
  / clamp-floor facts of ``tests/resources/test_mel.npy`` -> "parity unpinned"
  for the librosa arithmetic itself.
"""
