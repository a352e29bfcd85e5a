"""CPU oracle for the residual-TD3 navigation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU arm), never on the CUDA product path.

Every function restates, line by line, what the reference does in
``/root/reference/environment.py`` and ``/root/reference/robot.py`` (cited per
function).  The restatement is pinned against golden vectors produced by running
the unmodified reference in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).

Parity status: PINNED for dynamics/step/reset/init+goal seeding, replay ring and
index sampling, TD3 critic/actor/Polyak steps and the Robot per-step hooks.
UNPINNED for the Perlin map generator (``perlin_noise`` is an unvendored,
unversioned third-party dependency that is absent here); maps are inputs.
"""
