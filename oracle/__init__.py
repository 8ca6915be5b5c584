"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the magpy_rv likelihood hot path.

Nothing in the product package (``magpy_rv_b200``) may import, call or execute anything in this
directory.  The only permitted users are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only as the checker or the
timed CPU baseline, never as the thing shipped.

Contents
--------
``ref_loader``   imports the *unmodified* reference (``/root/reference/src/magpy_rv``) with
                 matplotlib / seaborn stubbed.  Works only in the build container
                 (the GPU box has no ``/root/reference``); used to pin the port and to generate
                 ``tests/golden/*.npz``.
``logl_port``    NumPy/SciPy restatement of the reference's algorithm for the path, function by
                 function with reference file:line citations.  Travels to the GPU box.
``make_golden``  script that generated ``tests/golden/*.npz`` from the real reference.

Parity pinning: the port is checked (tests/test_oracle.py) against (i) the one known answer the
reference itself records (tutorial 3: Initial Log Likelihood -1719.4341758413852,
``source/tutorials/3_51_peg_tutorial.ipynb`` cell 14) and (ii) golden vectors produced by running
the real reference in the build container (``make_golden.py``).
"""
