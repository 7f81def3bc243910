"""CPU oracle for the pykaldi2 hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a float64 CPU restatement of the reference's
algorithm (numpy / torch-CPU / plain C).  It is the checker for the CUDA path.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product (``pykaldi2_b200``) never
does, and fails loudly if the CUDA library is missing.

Parity status:
  * fbank / CMN / chunking: pinned against the reference's own Python code
    (imported from /root/reference in the build container; golden vectors in
    tests/golden/ made by oracle/make_golden.py).
  * BLSTM: the reference model is torch.nn.LSTM + nn.Linear, so torch-CPU fp32/fp64
    is the reference itself.
  * lattice MMI / LF-MMI forward-backward: the reference delegates to Kaldi via
    PyKaldi, which is neither vendored nor pinned (docker/Dockerfile:57-64) and
    ships no tests or golden vectors -> "parity unpinned" by the reference.
    These restatements are pinned by brute-force path enumeration, by autograd
    (posteriors == dlogZ/dloglikes) and by two independent formulations agreeing.
"""
