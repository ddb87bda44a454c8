"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the crowddynamics per-timestep agent update.

Nothing under ``oracle/`` is part of the shipped product path.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import it, and there only as the checker / the CPU arm that is timed beside the GPU path.
"""
