"""CPU oracle for the learned-compression forward path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU
baseline.  The product path (``neural_image_compression_b200``) never imports
this package and fails loudly when its CUDA library is missing.
"""
