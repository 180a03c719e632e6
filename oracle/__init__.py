"""CPU oracle for the HGR-Net scoring head.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``hgrnet_b200``) never imports it and has no CPU fallback.
"""
