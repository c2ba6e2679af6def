"""TEST INFRASTRUCTURE — CPU restatements ("oracle") of the reference's hot-path algorithms.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker / reported baseline.  The product
(``bevgen_b200``) never imports it and has no CPU fallback.
"""
