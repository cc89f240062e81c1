"""oracle/ -- CPU restatement of ns_gym's non-stationary env-step path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package, and there only as
the *checker* (or as the timed CPU baseline), never as the thing shipped.  Nothing
under ``ns_gym_b200/`` imports ``oracle``.

Contents
--------
gym_restated.py  Restatement of the gymnasium 1.2.1 base-env arithmetic the reference
                 delegates to (CartPole / Acrobot / MountainCar / Continuous MountainCar /
                 Pendulum / FrozenLake / CliffWalking ``step`` + ``reset``, ``TimeLimit``,
                 ``make`` / ``register``).  gymnasium is a third-party dependency of the
                 reference (``uv.lock:958-959``: gymnasium==1.2.1) that is NOT vendored
                 under /root/reference and NOT installed in this image.
                 **Parity status of this file: UNPINNED** -- no reference test asserts a
                 numeric post-step state of any gymnasium env, and the package cannot be
                 imported here to diff against.  It is restated from the published
                 upstream algorithm and cross-checked against the in-tree legacy copy
                 ``ns_gym/benchmark_algorithms/rats-experiments/code/envs/nscartpole_v0.py:92-100``.
ns_port.py       Restatement of the reference's OWN code on the path (schedulers,
                 update functions, NS wrappers, Bridge env).  **Parity status: PINNED**
                 -- checked in this container against the reference itself
                 (``/root/reference`` imported verbatim on top of ``gym_restated`` used as
                 the ``gymnasium`` shim, see ``ref_loader.py``) and against the committed
                 golden vectors under ``tests/golden/`` generated from that import.
ref_loader.py    Imports the real reference (only where /root/reference exists).
streams.py       Pre-drawn random-stream injection stubs shared by oracle / reference.
vector.py        Sync / multi-process vector-env loop (next-step autoreset) used as the
                 timed CPU baseline.
"""
