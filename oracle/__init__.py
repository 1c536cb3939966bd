"""CPU oracle for the avex embedding hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithm (earthspecies/avex v1.2.0) for
  waveform -> Kaldi log-mel fbank -> BEATs encoder forward   (kaldi_fbank.py, relpos.py, beats_encoder.py)
  waveform -> STFT mel-spectrogram (AudioProcessor)          (melspec.py)
  mel-spectrogram -> EfficientNet-B0 features                (effnet.py)
Every function cites the reference file:line it follows.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` leg
may import this package, and only as the checker / the reported CPU baseline.  The product package
`avex_b200` never imports it and fails loudly when its CUDA library is missing.

Parity pin: the oracle is checked against golden vectors produced by the *reference itself*
(imported from /root/reference in the build container by tests/golden/make_golden.py; vectors and the
generating script are committed under tests/golden/).  See tests/test_oracle_vs_golden.py.
"""
