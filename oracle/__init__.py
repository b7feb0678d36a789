"""CPU oracle for the Floor-Field-Model hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the shipped package (``ffm_b200/``) may import, call, link or execute
anything under ``oracle/``.  The only legitimate users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py``, and there only as the checker / the reported CPU baseline.

Parity status: PINNED.  The restatements here are checked (tests/test_oracle_vs_golden.py)
against golden vectors produced by running the unmodified reference classes
(``/root/reference/model/ffm_core.py``, ``ffm_unified.py``, ``ffm_trained_core.py``,
``ffm_learning_core.py``) under the injected-draw protocol of ``oracle/inject.py``
(generator: ``oracle/make_golden.py``, fixtures: ``tests/golden/*.npz``) and against
the assets the reference ships (``data/maps/simple_room.npy``, ``data/sff/distance_*.npy``,
digests in ``tests/golden/shipped_assets.json``).
"""
