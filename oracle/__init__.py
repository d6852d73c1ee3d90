"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the volume-rendering hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker or the timed CPU
baseline -- never as the thing shipped.  The product path
(``articulated-object-nerf_b200`` / ``aon_b200``) never imports this package.
"""
