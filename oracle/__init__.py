"""TEST INFRASTRUCTURE — not product code.

CPU restatement of the reference's algorithm for the moondream hot path (``moondream_oracle``)
plus, in the build container only, a shim that runs the unmodified reference from
``/root/reference`` (``reference_shim``) to pin the restatement and to generate the golden
fixtures under ``tests/golden``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``moondream_b200/`` does.
"""
