"""CPU oracle for the pre -> post -> track hot path.

TEST INFRASTRUCTURE ONLY.  ``oracle/`` restates the reference's algorithm on the CPU so
that the CUDA path can be checked bit-for-bit.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and there
only as the checker or the reported CPU baseline -- never as the thing shipped.  The product
package (``realtime_video_analytics_32streams_b200``) does not import it and fails loudly
when its CUDA library is missing.

Parity pinning: the reference has no tests and no golden vectors, so the
oracle is pinned against outputs of the reference's own functions, produced by importing
``/root/reference/src`` in the build container (``tests/golden/make_golden.py`` ->
``tests/golden/*.npz``), and against the installed OpenCV 4.13.0 / NumPy 2.3.5 at test time.
"""
