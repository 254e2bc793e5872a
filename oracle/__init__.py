"""CPU oracle for the MNASNet training step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or the
reported CPU baseline -- never as a product path.  The product (``mnasnet-pytorch_b200``)
raises if its CUDA library is missing; it never routes through here.
"""
