"""CPU oracle for the IIC hot path -- TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU (torch fp32/fp64) restatement of the
reference algorithm (xu-ji/IIC).  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The product (``iic_b200``) never imports,
links or executes anything here and has no CPU fallback.

Parity pinning: the reference ships no tests / golden vectors, so
the oracle is pinned against outputs of the reference's own modules run in the
build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
