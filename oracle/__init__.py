"""CPU oracle for the audio-visual hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  It may be imported only by
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- always as the checker (or as the CPU
arm being timed), never as a fallback for the CUDA path.  The product package
``multimodal_lipread_b200`` raises if its CUDA library is missing.

Contents
--------
logmel.py     float64 numpy restatement of the log-mel frontend
              (reference ``audio/utils/audio_processor.py:9-21,48-52,60-64`` and
              ``audio/data_utils/dataset.py:52``; the arithmetic itself lives in
              torchaudio 2.6.0 (pinned, requirements.txt:89): ``functional.spectrogram``,
              ``functional.melscale_fbanks``).
frontend.py   the reference's own torchaudio call sequence (fp32) -- the "port" that is
              timed as the CPU baseline and used as the fp32 parity target.
av_models.py  torch/torchvision restatement of the AV models on the path
              (reference ``audio_video/models/middle_fusion_fast.py:5-42`` ...).

Parity pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4).  The oracle is pinned against outputs of the reference's own
modules imported in the build container -- ``tests/golden/make_golden.py`` is the
generating script, ``tests/golden/*.npz`` the committed vectors.
"""
