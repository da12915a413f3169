"""CPU oracle of the score path — TEST INFRASTRUCTURE ONLY.

Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py``.  The product (``matrix-factorization-torch_b200/``) never imports this package.
"""
