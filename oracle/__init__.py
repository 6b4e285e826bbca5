"""CPU oracle for the depth hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``video_3d_pipeline`` + ``libv3d.so``) never does.

Three things live here:

* ``sgbm``      -- ctypes wrapper around ``sgbm_oracle.c``, a plain-C restatement
                   of the cv2 arithmetic behind ``depth.py:250-406`` (parity
                   pinned against the live ``cv2`` by tests/test_oracle_vs_cv2.py
                   and by tests/golden/).
* ``cv2_chain`` -- the reference's own call chain (``depth.py:257-266, 274-275,
                   315-325, 337-341, 374, 400-403``) calling the same ``cv2``
                   functions; this IS the reference CPU implementation and is
                   what ``bench.py --impl reference`` times.
* ``guided``    -- float64 numpy definition of the guided 1080p->2160p upscale.
                   The reference has no guided filter (``upscale.py:47-59`` is
                   ffmpeg ``scale``), so this oracle is normative and its parity
                   is UNPINNED by the reference.
"""
