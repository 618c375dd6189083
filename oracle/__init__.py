"""CPU oracle for the SED hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (numpy float64 / torch-CPU float32), what the
reference `fumchin/bird-sound-event-detecion` computes on the hot path named in
BASELINE.json.  It is the *checker* for the CUDA product under
`bird-sound-event-detecion_b200/`; it is never the thing shipped or measured.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import anything from here.  The product package must not.

Pinning status (see DESIGN.md "Oracle"):
  * oracle.crnn   -- PINNED: checked against the reference's own `models/CRNN.py`
                     executed in the build container (tests/golden/*.npz, produced by
                     tests/make_golden.py, which imports /root/reference/src).
  * oracle.train  -- PINNED for the model/loss/EMA arithmetic via the same reference
                     modules (the loss assembly of src/main.py:train_mt is restated,
                     the script itself cannot be imported: tensorboardX/librosa absent).
  * oracle.frontend, oracle.postproc -- PARITY UNPINNED by the reference (it ships no
                     tests, and librosa / dcase_util are not installable here).  They are
                     cross-checked against independent implementations present in the
                     container (torch.stft, torchaudio melscale_fbanks,
                     scipy.ndimage.median_filter) in tests/test_oracle_*.py.
"""
