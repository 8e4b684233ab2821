"""CPU oracle for the sharded-DP Llama training step.  TEST INFRASTRUCTURE ONLY.

Nothing under ``automodel_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / reported CPU baseline - never as the thing shipped.
"""
